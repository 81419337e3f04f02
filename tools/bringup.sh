#!/bin/bash
# Runs the bring-up diagnostics, each group in its own process (a device trap is sticky) with its own timeout.
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,driver_version --format=csv
timeout 120 python tools/bringup.py layout; echo "exit=$?"
timeout 300 python tools/bringup.py fprop; echo "exit=$?"
timeout 300 python tools/bringup.py wgrad; echo "exit=$?"
timeout 200 python tools/bringup.py wgrad 1024 8192; echo "exit=$?"
} > gpurun_out/bringup.log 2>&1
tail -80 gpurun_out/bringup.log
