"""Narrow (N = 64 / 128) 3x3 layers of VGG / the Discriminator: in-stream time per launch under the kernel-selection
options (CTA pairs forced / auto, resident weights on / off)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pesr_b200 import _lib, ops  # noqa: E402


ONE = len(sys.argv) > 1 and sys.argv[1] == "one"      # two launches per variant only (for ncu)


def chain(fn, n=20):
    for _ in range(2 if ONE else 3):
        fn()
    torch.cuda.synchronize()
    if ONE:
        return float("nan")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for nb, h, w, cin, cout in ((32, 192, 192, 64, 64), (32, 96, 96, 64, 128), (32, 96, 96, 128, 64), (32, 96, 96, 128, 128)):
    x = torch.randn(nb, h, w, cin, device="cuda").half()
    y = torch.empty(nb, h, w, cout, device="cuda", dtype=torch.float16)
    wp = (torch.randn(9 * cout, cin, device="cuda") / (3 * cin ** 0.5)).half()
    b = torch.randn(cout, device="cuda")
    fl = 2.0 * nb * h * w * cin * cout * 9
    for pair, wres in ((1, 1), (1, 0), (2, 1), (2, 0)):
        _lib.set_option(_lib.OPT_PAIR_MODE, pair)
        _lib.set_option(_lib.OPT_RESIDENT_WEIGHTS, wres)
        d = ops.make_conv_desc(dtype=0, nb=nb, h=h, w=w, cin=cin, cout=cout, srcs=[ops.nhwc_src(x, nb, h, w, cin)], wpacked=wp,
                               bias=b, act=ops.ACT_RELU, out16=y, ld_out16=cout)
        t = chain(lambda: ops.conv_igemm(d))
        print(f"{nb}x{h}x{w} {cin}->{cout}: pair mode {pair} resident weights {wres}: {t:7.1f} us  {fl / t / 1e6:6.0f} TFLOP/s", flush=True)
    _lib.set_option(_lib.OPT_PAIR_MODE, 1)
    _lib.set_option(_lib.OPT_RESIDENT_WEIGHTS, 1)
