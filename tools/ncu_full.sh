#!/bin/bash
# One `ncu --set full` capture of each hot tensor-core kernel on the G trunk shape (B=16, 48x48, 256->256, fp16).
# Run on the GPU box through gpurun; the .ncu-rep files come back in gpurun_out/ and tools/ncu_summary.py turns
# them into the text summaries under profiles/.
mkdir -p gpurun_out
export PESR_B200_LIB=pesr_b200/libpesr_b200_debug.so   # tools/perf_conv.py binds the bring-up hooks (tools/build_debug.sh)
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:conv_igemm --launch-skip 5 --launch-count 1 -o gpurun_out/ncu_igemm_light python tools/perf_conv.py one > gpurun_out/ncu_one.log 2>&1
$NCU -k regex:conv_igemm --launch-skip 18 --launch-count 1 -o gpurun_out/ncu_igemm_residual python tools/perf_conv.py one >> gpurun_out/ncu_one.log 2>&1
$NCU -k regex:conv_wgrad --launch-skip 5 --launch-count 1 -o gpurun_out/ncu_wgrad python tools/perf_conv.py one >> gpurun_out/ncu_one.log 2>&1
ls -la gpurun_out/*.ncu-rep
