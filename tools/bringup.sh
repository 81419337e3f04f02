#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python tools/bringup.py disc; echo "exit=$?"
timeout 300 python tools/bringup.py vgg; echo "exit=$?"
timeout 300 python tools/bringup.py gan; echo "exit=$?"
} > gpurun_out/bringup3.log 2>&1
tail -60 gpurun_out/bringup3.log
