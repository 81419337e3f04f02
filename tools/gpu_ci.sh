#!/bin/bash
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -12; echo "pytest exit=$?"
timeout 300 python tools/perf_conv.py conv; echo "perf exit=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline; echo "bench exit=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --workload gan --no-cpu-baseline; echo "bench gan exit=$?"
} > gpurun_out/ci.log 2>&1
tail -40 gpurun_out/ci.log | cut -c1-420
