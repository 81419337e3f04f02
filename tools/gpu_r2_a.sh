#!/bin/bash
# Round-2 GPU session A: full GPU test-suite with the printed parity numbers, smoke, bench with and without the CUDA graph.
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader
nproc
timeout 3000 python -m pytest tests -q -m gpu -s -p no:cacheprovider --durations=15 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit=$?"
grep -E "passed|failed|error" gpurun_out/r2a_pytest.log | tail -5
timeout 300 python __graft_entry__.py --smoke; echo "smoke exit=$?"
timeout 1200 python bench.py --steps 50 --warmup 3 > gpurun_out/r2a_bench_graph.json 2> gpurun_out/r2a_bench_graph.err; echo "bench graph exit=$?"
cut -c1-1500 gpurun_out/r2a_bench_graph.json; tail -5 gpurun_out/r2a_bench_graph.err
timeout 900 python bench.py --steps 50 --warmup 3 --no-graph --no-extras --no-cpu-baseline > gpurun_out/r2a_bench_eager.json 2> gpurun_out/r2a_bench_eager.err; echo "bench eager exit=$?"
cut -c1-900 gpurun_out/r2a_bench_eager.json; tail -5 gpurun_out/r2a_bench_eager.err
} > gpurun_out/r2a.log 2>&1
tail -40 gpurun_out/r2a.log
