#!/bin/bash
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -30; echo "pytest exit=$?"
timeout 300 python __graft_entry__.py --smoke; echo "smoke exit=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline; echo "bench exit=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --workload gan --no-cpu-baseline; echo "bench gan exit=$?"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_gan.csv python tools/profile_step.py gan > gpurun_out/prof_gan.log 2>&1; echo ncu=$?
python tools/summarize_launches.py gpurun_out/launches_gan.csv | head -32
} > gpurun_out/ci.log 2>&1
tail -90 gpurun_out/ci.log | cut -c1-600
