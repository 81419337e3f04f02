#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_infer_gpu.py tests/test_gan_gpu.py -x -q -m gpu 2>&1 | tail -15; echo "exit=$?"
timeout 300 python train.py --synthetic --phase pretrain --num_epochs 1 --max_iters 4 --log_every 2 --patch_size 24 --num_blocks 4 --num_channels 64 --check_point /tmp/ck 2>&1 | tail -6; echo "train pretrain exit=$?"
timeout 300 python train.py --synthetic --phase train --num_epochs 1 --max_iters 4 --log_every 2 --patch_size 24 --num_blocks 4 --num_channels 64 --check_point /tmp/ck --pretrained_model /tmp/ck/pretrain/model_1.pt 2>&1 | tail -6; echo "train gan exit=$?"
python - <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
from pesr_b200.model import Generator
from pesr_b200 import infer
G = Generator({'depth': 32, 'num_channels': 256, 'res_scale': 0.1}).cuda().eval()
for (h, w) in ((128, 128), (339, 510)):
    x = torch.rand(1, 3, h, w, device='cuda') * 255
    for _ in range(2):
        infer.super_resolve(G, x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        out32, out8 = infer.super_resolve(G, x)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print(f"infer {h}x{w}: {dt*1e3:.2f} ms  {16*h*w/dt/1e6:.1f} HR Mpix/s  {h*w*100505088/dt/1e12:.0f} TFLOP/s  out8 {tuple(out8.shape)}")
PY
} > gpurun_out/bringup5.log 2>&1
tail -50 gpurun_out/bringup5.log
