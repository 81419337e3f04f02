#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python tools/debug_mem.py 2>&1 | tail -12
timeout 300 python -m pytest tests/test_boundary_gpu.py tests/test_netops_gpu.py tests/test_gan_gpu.py -q -m gpu -x 2>&1 | tail -3
timeout 900 python bench.py --steps 50 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2b_bench_graph.json 2> gpurun_out/r2b_bench_graph.err; echo "bench graph exit=$?"
cut -c1-1200 gpurun_out/r2b_bench_graph.json; tail -3 gpurun_out/r2b_bench_graph.err
PESR_NO_MERGED_S2_DGRAD=1 timeout 900 python bench.py --steps 50 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2b_bench_graph_nomerge.json 2> gpurun_out/r2b_bench_graph_nomerge.err; echo "bench nomerge exit=$?"
cut -c1-300 gpurun_out/r2b_bench_graph_nomerge.json
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_gan.csv python tools/profile_step.py gan > gpurun_out/r2b_prof.log 2>&1; echo ncu=$?
python tools/summarize_launches.py gpurun_out/r2b_launches_gan.csv > gpurun_out/r2b_launches_gan.txt; head -40 gpurun_out/r2b_launches_gan.txt
} > gpurun_out/r2b.log 2>&1
tail -80 gpurun_out/r2b.log
