"""cProfile of the host side of training steps (where does the CPU time per launch go?).  Run on the GPU box."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from pesr_b200 import steps
from pesr_b200.model import VGG, Discriminator, Generator
from pesr_b200.optim import Adam

workload = sys.argv[1] if len(sys.argv) > 1 else "gan"
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
for _ in range(5):
    step()
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"host enqueue {1e3*(t1-t0):.2f} ms, until GPU done {1e3*(t2-t0):.2f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
    torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(40)
