#!/bin/bash
# Evidence for profiles/: memcheck over the kernel tests, ncu --set full of the hot kernels, in-stream chain timings,
# SM-hog probe (debug build), GPU test log with the printed parity numbers.
mkdir -p gpurun_out
{
echo "== memcheck"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_conv_gpu.py tests/test_netops_gpu.py tests/test_io_gpu.py "tests/test_gan_gpu.py::test_gan_step_matches_pinned_oracle" tests/test_generator_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r02_memcheck.log 2>&1; echo "memcheck exit=$?"
tail -5 gpurun_out/r02_memcheck.log
echo "== ncu full"
bash tools/ncu_full.sh > gpurun_out/r02_ncu_full.log 2>&1
for f in igemm_light igemm_residual wgrad; do python tools/ncu_summary.py gpurun_out/ncu_$f.ncu-rep > gpurun_out/r02_ncu_$f.txt 2>&1; done
head -12 gpurun_out/r02_ncu_igemm_light.txt
echo "== perf_chain"
timeout 300 python tools/perf_chain.py > gpurun_out/r02_perf_chain.txt 2>&1; tail -12 gpurun_out/r02_perf_chain.txt
echo "== sm_hog"
PESR_B200_LIB=pesr_b200/libpesr_b200_debug.so timeout 300 python tools/sm_hog_probe.py > gpurun_out/r02_sm_hog.txt 2>&1; tail -8 gpurun_out/r02_sm_hog.txt
} > gpurun_out/r02_evidence.log 2>&1
tail -60 gpurun_out/r02_evidence.log
