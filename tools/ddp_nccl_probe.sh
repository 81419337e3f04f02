#!/bin/bash
# 8-GPU GAN-step bench under different NCCL settings (each NCCL channel is one CTA that takes an SM away from the
# one-CTA-per-SM persistent conv kernels).
mkdir -p gpurun_out
run() {
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 8 --steps 10 --warmup 3 --no-extras 2> gpurun_out/nccl_probe.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],3), round(d['value'],1))"
}
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 8 --steps 3 --warmup 3 --no-extras > gpurun_out/nccl_info.log 2>&1
grep -i "nvls\|channels\|algo" gpurun_out/nccl_info.log | grep "\[0\]" | head -12 | cut -c1-200
tail -1 gpurun_out/nccl_info.log | cut -c1-160
run NCCL_ALGO=NVLS NCCL_MAX_NCHANNELS=8
