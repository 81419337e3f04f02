#!/bin/bash
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests/test_split_precision_gpu.py -q -m gpu -s -p no:cacheprovider 2>&1 | grep -vE "^\s*$" | tail -40
} > gpurun_out/r2_split.log 2>&1
tail -60 gpurun_out/r2_split.log | cut -c1-400
