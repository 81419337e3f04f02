#!/bin/bash
# Round-2 closing run: full GPU suite (with the printed parity numbers), smoke, the default bench, the CPU reference arm
# (short), the launch list of one GAN step, ncu --set full of the hot kernels, in-stream chains.
mkdir -p gpurun_out
{
timeout 3000 python -m pytest tests -q -m gpu -s -p no:cacheprovider --durations=10 > gpurun_out/r02_gputest.log 2>&1; echo "pytest exit=$?"
grep -E "passed|failed" gpurun_out/r02_gputest.log | tail -2
grep -E "^FAILED|^ERROR" gpurun_out/r02_gputest.log | head
timeout 300 python __graft_entry__.py --smoke; echo "smoke exit=$?"
timeout 1500 python bench.py > gpurun_out/r02_bench_gan_1gpu.json 2> gpurun_out/r02_bench_gan_1gpu.err; echo "bench exit=$?"
cut -c1-700 gpurun_out/r02_bench_gan_1gpu.json; tail -3 gpurun_out/r02_bench_gan_1gpu.err
timeout 900 python bench.py --workload pretrain --no-extras --no-cpu-baseline > gpurun_out/r02_bench_pretrain_1gpu.json 2>/dev/null; echo "bench pretrain exit=$?"
cut -c1-300 gpurun_out/r02_bench_pretrain_1gpu.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_cpu.json 2>/dev/null; echo "bench ref exit=$?"
cut -c1-900 gpurun_out/r02_bench_reference_cpu.json
bash tools/gpu_ncu_list.sh r02 | tail -3
bash tools/ncu_full.sh > gpurun_out/r02_ncu_full.log 2>&1
for f in igemm_light igemm_residual wgrad; do python tools/ncu_summary.py gpurun_out/ncu_$f.ncu-rep > gpurun_out/r02_ncu_$f.txt 2>&1; done
head -14 gpurun_out/r02_ncu_igemm_light.txt
timeout 300 python tools/perf_chain.py > gpurun_out/r02_perf_chain.txt 2>&1; tail -12 gpurun_out/r02_perf_chain.txt
} > gpurun_out/r02_final.log 2>&1
tail -60 gpurun_out/r02_final.log
