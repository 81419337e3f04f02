#!/bin/bash
mkdir -p gpurun_out
{
for c in 4 8 32; do
NCCL_MAX_CTAS=$c PROBE_SYMM=$([ $c = 4 ] && echo 1 || echo 0) timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) tools/nccl_bw_probe.py 2>&1 | grep -v "OMP_NUM\|^\*\*\*\|^$"
done
NCCL_MAX_CTAS=4 NCCL_PROTO=Simple timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) tools/nccl_bw_probe.py 2>&1 | grep -v "OMP_NUM\|^\*\*\*\|^$" | sed 's/^/PROTO=Simple /'
NCCL_MAX_CTAS=4 NCCL_NTHREADS=512 NCCL_BUFFSIZE=16777216 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) tools/nccl_bw_probe.py 2>&1 | grep -v "OMP_NUM\|^\*\*\*\|^$" | sed 's/^/BUFFSIZE=16M /'
} > gpurun_out/r2q.log 2>&1
cat gpurun_out/r2q.log | cut -c1-300
