#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_conv_gpu.py -q -m gpu -p no:cacheprovider -x -k "resident" 2>&1 | tail -8
timeout 300 python tools/phase_times.py gan 10 detail 2>&1 | grep "gan step\|64->64 taps9\|128->64 taps9\|64->128 taps9\|VGG forward\|VGG backward\|D forward\|D backward"
} > gpurun_out/r2wres.log 2>&1
cat gpurun_out/r2wres.log | cut -c1-220
