#!/bin/bash
# 8-GPU GAN-step bench under different NCCL channel limits (each NCCL channel is one CTA that competes with the
# one-CTA-per-SM persistent conv kernels for an SM).
mkdir -p gpurun_out
run() {
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 8 --steps 20 --warmup 5 --no-extras 2> gpurun_out/nccl_probe.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],3), round(d['value'],1))"
}
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 8 --steps 5 --warmup 3 --no-extras 2>&1 | grep -i "channels\|nvls\|Using network" | head -8
run NCCL_MAX_NCHANNELS=4
run NCCL_MAX_NCHANNELS=8
run NCCL_MAX_NCHANNELS=16
