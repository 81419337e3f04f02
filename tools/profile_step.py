"""Runs warm-up steps, then ONE training step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from pesr_b200 import steps
from pesr_b200.model import VGG, Discriminator, Generator
from pesr_b200.optim import Adam

workload = sys.argv[1] if len(sys.argv) > 1 else "gan"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
torch.manual_seed(0)
G = Generator(bench.OPT).to(dev)
optG = Adam(G.parameters(), lr=5e-5)
lr = torch.rand(16, 3, 48, 48, device=dev) * 255
hr = torch.rand(16, 3, 192, 192, device=dev) * 255
if workload == "gan":
    D = Discriminator(bench.OPT).to(dev)
    V = VGG(pretrained=False).to(dev)
    optD = Adam(D.parameters(), lr=5e-5)
    cfg = dict(steps.DEFAULT_GAN_CFG)
    cfg['target_real'] = torch.ones(16, 1, device=dev)
    cfg['target_fake'] = torch.zeros(16, 1, device=dev)

    def step():
        return steps.gan_step(G, D, V, optG, optD, lr, hr, cfg)
else:
    def step():
        return steps.pretrain_step(G, optG, lr, hr)
for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(nsteps):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
