"""Timing of the 3-channel edge kernels (im2col3 / col2im3) at the GAN-step shapes (run on the GPU box)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pesr_b200 import ops  # noqa: E402
from tools.perf_conv import timeit  # noqa: E402

print("library:", os.environ.get("PESR_B200_LIB", "in-tree"))
for nb, h, w in ((16, 192, 192), (16, 48, 48)):
    x = torch.rand(nb, 3, h, w, device="cuda") * 255
    col = torch.empty(nb * h * w, 64, device="cuda", dtype=torch.float16)
    for flush in (False, True):
        ms = timeit(lambda: ops.im2col3(x, col), iters=20, flush=flush)
        print(f"  im2col3 {nb}x{h}x{w} flush={int(flush)}: {ms*1e3:.1f} us ({(col.numel()*2 + x.numel()*4)/ms/1e6:.0f} GB/s)")
    z = torch.randn(nb * h * w, 32, device="cuda")
    out = torch.empty(nb, 3, h, w, device="cuda")
    for flush in (False, True):
        ms = timeit(lambda: ops.col2im3(z, 32, nb, h, w, out), iters=20, flush=flush)
        print(f"  col2im3 {nb}x{h}x{w} flush={int(flush)}: {ms*1e3:.1f} us ({(z.numel()*4 + out.numel()*4)/ms/1e6:.0f} GB/s)")
