#!/bin/bash
# Round-2 closing run (second session): full GPU suite with the printed parity numbers, smoke, the default bench, the
# pretrain bench, the launch list of one GAN step, memcheck over the new kernels.
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -q -m gpu -s -p no:cacheprovider --durations=8 > gpurun_out/r02b_gputest.log 2>&1; echo "pytest exit=$?"
grep -E "passed|failed" gpurun_out/r02b_gputest.log | tail -2
grep -E "^FAILED|^ERROR" gpurun_out/r02b_gputest.log | head
timeout 300 python __graft_entry__.py --smoke; echo "smoke exit=$?"
timeout 1500 python bench.py > gpurun_out/r02b_bench_gan_1gpu.json 2> gpurun_out/r02b_bench_gan_1gpu.err; echo "bench exit=$?"
cut -c1-600 gpurun_out/r02b_bench_gan_1gpu.json; tail -3 gpurun_out/r02b_bench_gan_1gpu.err
timeout 900 python bench.py --workload pretrain --no-extras --no-cpu-baseline > gpurun_out/r02b_bench_pretrain_1gpu.json 2>/dev/null; echo "bench pretrain exit=$?"
cut -c1-300 gpurun_out/r02b_bench_pretrain_1gpu.json
bash tools/gpu_ncu_list.sh r02b | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_netops_gpu.py tests/test_split_precision_gpu.py -q -m gpu -p no:cacheprovider -x -k "tensor_cores or patch16 or vgg" > gpurun_out/r02b_memcheck.log 2>&1; echo "memcheck exit=$?"
tail -4 gpurun_out/r02b_memcheck.log
} > gpurun_out/r02b_final.log 2>&1
tail -50 gpurun_out/r02b_final.log | cut -c1-700
