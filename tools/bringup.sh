#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python tools/bringup.py gen; echo "exit=$?"
timeout 300 python tools/perf_conv.py conv; echo "exit=$?"
timeout 300 python tools/perf_conv.py gen; echo "exit=$?"
} > gpurun_out/bringup2.log 2>&1
tail -80 gpurun_out/bringup2.log
