#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -30; echo "pytest exit=$?"
timeout 300 python __graft_entry__.py --smoke; echo "smoke exit=$?"
timeout 600 python bench.py --steps 10 --warmup 3; echo "bench exit=$?"
} > gpurun_out/ci.log 2>&1
tail -60 gpurun_out/ci.log
