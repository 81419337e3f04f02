#!/bin/bash
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests/test_split_precision_gpu.py -q -m gpu -p no:cacheprovider -s -k "discriminator or vgg or gan_step" 2>&1 | grep -v "^$" | grep -v "^  \|^    " | tail -150
} > gpurun_out/r2o.log 2>&1
grep "vs free-running\|passed\|failed\|Error\|assert" gpurun_out/r2o.log | cut -c1-300
