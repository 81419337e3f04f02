#!/bin/bash
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests -x -q -m gpu -s 2>&1 | tail -40; echo "pytest exit=$?"
timeout 300 python __graft_entry__.py --smoke; echo "smoke exit=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline; echo "bench exit=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --workload gan --no-cpu-baseline; echo "bench gan exit=$?"
} > gpurun_out/ci.log 2>&1
tail -70 gpurun_out/ci.log
