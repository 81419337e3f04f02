#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
# launch order in `one` mode: 2 launches per (pair, wres) variant, 4 variants per shape; #1 = 64->64 @192x192x32 default path
$NCU -k regex:conv_igemm --launch-skip 1 --launch-count 1 -o gpurun_out/r2_ncu_narrow_64_64 python tools/perf_narrow.py one > gpurun_out/r2_ncu_narrow.log 2>&1
# #9 = 64->128 @96x96x32 with resident weights (default path)
$NCU -k regex:conv_igemm --launch-skip 9 --launch-count 1 -o gpurun_out/r2_ncu_narrow_64_128 python tools/perf_narrow.py one >> gpurun_out/r2_ncu_narrow.log 2>&1
ls -la gpurun_out/r2_ncu_narrow*.ncu-rep
