"""Timing of the HBM-bound helper kernels at the GAN-step shapes (run on the GPU box).
Each case is timed warm (operands L2-resident, as inside a step) and cold (L2 flushed)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pesr_b200 import ops  # noqa: E402
from tools.perf_conv import timeit  # noqa: E402


def case(name, fn, nbytes):
    warm = timeit(fn, iters=20, flush=False)
    cold = timeit(fn, iters=10, flush=True)
    print(f"  {name:46s} warm {warm*1e3:7.1f} us ({nbytes/warm/1e6:7.0f} GB/s)   cold {cold*1e3:7.1f} us ({nbytes/cold/1e6:7.0f} GB/s)",
          flush=True)


def main():
    dev = "cuda"
    print("library:", os.environ.get("PESR_B200_LIB", "in-tree"))
    # split-K reduce of a trunk conv: 8 splits x 9 taps x 256 x 256 fp32
    for splits, co, ci in ((8, 256, 256), (4, 1024, 256), (8, 64, 64), (4, 512, 512)):
        part = torch.randn(splits * 9 * co * ci, device=dev)
        grad = torch.zeros(co, ci, 3, 3, device=dev)
        case(f"wgrad_reduce {splits}x9x{co}x{ci}", lambda: ops.wgrad_reduce(part, splits, 9, co, ci, 0, co, ci, grad), part.numel() * 4 + grad.numel() * 4)
    for npix, c in ((16 * 48 * 48, 256), (16 * 96 * 96, 1024), (16 * 192 * 192, 256), (16 * 96 * 96, 64)):
        x = torch.randn(npix, c, device=dev).half()
        out = torch.zeros(c, device=dev)
        case(f"colsum16 {npix}x{c}", lambda: ops.colsum16(x, npix, c, c, out), x.numel() * 2)
    for npix, c in ((16 * 192 * 192, 64), (16 * 96 * 96, 64), (16 * 96 * 96, 128), (16 * 48 * 48, 256), (16 * 24 * 24, 512), (16 * 12 * 12, 512)):
        y = torch.randn(npix, c, device=dev).half()
        ws = torch.zeros(2 * c, device=dev, dtype=torch.float64)
        mean, rstd = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
        case(f"bn_stats {npix}x{c}", lambda: ops.bn_stats(y, npix, c, ws, mean, rstd), y.numel() * 2)


if __name__ == "__main__":
    main()
