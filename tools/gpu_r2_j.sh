#!/bin/bash
# upsampler probe: in-stream timings + ncu --set full of the 96x96 fprop / dgrad / wgrad launches
mkdir -p gpurun_out
python tools/perf_up.py > gpurun_out/r2j_perf_up.txt 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
# launch order in `one` mode: per shape 2x each of 4 igemm variants, then 2x wgrad; 96x96 igemm launches are #8..15
$NCU -k regex:conv_igemm --launch-skip 9 --launch-count 1 -o gpurun_out/r2j_ncu_up2_fprop python tools/perf_up.py one > gpurun_out/r2j_ncu.log 2>&1
$NCU -k regex:conv_igemm --launch-skip 15 --launch-count 1 -o gpurun_out/r2j_ncu_up2_dgrad python tools/perf_up.py one >> gpurun_out/r2j_ncu.log 2>&1
$NCU -k regex:conv_wgrad --launch-skip 3 --launch-count 1 -o gpurun_out/r2j_ncu_up2_wgrad python tools/perf_up.py one >> gpurun_out/r2j_ncu.log 2>&1
cat gpurun_out/r2j_perf_up.txt
ls -la gpurun_out/r2j*
