#!/bin/bash
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5; echo "pytest exit=$?"
timeout 300 python __graft_entry__.py --smoke; echo "smoke exit=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_gan.json; echo "bench exit=$?"; cat gpurun_out/bench_gan.json | cut -c1-3000
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json; echo "bench ref exit=$?"; cat gpurun_out/bench_ref.json | cut -c1-700
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_gan.csv python tools/profile_step.py gan > gpurun_out/prof_gan.log 2>&1; echo ncu=$?
python tools/summarize_launches.py gpurun_out/launches_gan.csv | head -24
} > gpurun_out/ci.log 2>&1
tail -60 gpurun_out/ci.log
