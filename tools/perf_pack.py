"""Time the Generator's weight-pack launch (forward + backward layouts) on the GPU box."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pesr_b200.model import Generator  # noqa: E402
from tools.perf_conv import timeit  # noqa: E402

G = Generator(bench.OPT).cuda()
eng = G.engine()
eng._ensure_packed(torch.device("cuda", 0))
mp = eng.fwd_multi


def run():
    mp.key = None
    mp.run()


for flush in (False, True):
    ms = timeit(run, iters=20, flush=flush)
    n = sum(p.numel() for p in G.parameters())
    print(f"G pack (fwd + bwd layouts, {n/1e6:.1f} M params) flush={int(flush)}: {ms*1e3:.1f} us ({n*8/ms/1e6:.0f} GB/s)")
