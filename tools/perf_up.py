"""In-stream timing of the Generator's upsampler convolutions (model/basic.py:54-60: 256 -> 1024 conv + PixelShuffle(2) at
48x48 and 96x96, B=16) against the same GEMMs with a plain NHWC store: chains of launches under ONE event pair, PDL on.
`python tools/perf_up.py [one]` -- `one` launches each variant twice only (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pesr_b200 import ops  # noqa: E402

ONE = len(sys.argv) > 1 and sys.argv[1] == "one"
nb, c = 16, 256
mk = ops.make_conv_desc


def chain_time(fn, n):
    for _ in range(2 if ONE else 4):
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


for h in (48, 96):
    w = h
    P = nb * h * w
    x = torch.randn(nb, h, w, c, device="cuda").half()
    wf = (torch.randn(9 * 4 * c, c, device="cuda") / (3 * c ** 0.5)).half()          # [tap][cout=1024][cin=256]
    wd = (torch.randn(9 * c, 4 * c, device="cuda") / (3 * (4 * c) ** 0.5)).half()    # [tap][cout=256][cin=1024]
    bias = torch.randn(4 * c, device="cuda")
    up = torch.empty(nb, 2 * h, 2 * w, c, device="cuda", dtype=torch.float16)        # shuffled output
    flat = torch.empty(nb, h, w, 4 * c, device="cuda", dtype=torch.float16)          # plain output / dgrad input
    dz = torch.randn(nb, h, w, 4 * c, device="cuda").half()
    dx = torch.empty(nb, h, w, c, device="cuda", dtype=torch.float16)
    dzu = torch.empty(nb, h // 2, w // 2, 4 * c, device="cuda", dtype=torch.float16)
    src = [ops.nhwc_src(x, nb, h, w, c)]
    fl = 2.0 * P * c * 4 * c * 9
    variants = [
        ("fprop 256->1024 + PixelShuffle store (kEpi 5)",
         mk(dtype=0, nb=nb, h=h, w=w, cin=c, cout=4 * c, srcs=src, wpacked=wf, bias=bias, out16=up, ld_out16=c,
            out_mode=ops.OUT_SHUFFLE2, ps_c=c)),
        ("fprop 256->1024, plain NHWC store (kEpi 1)",
         mk(dtype=0, nb=nb, h=h, w=w, cin=c, cout=4 * c, srcs=src, wpacked=wf, bias=bias, out16=flat, ld_out16=4 * c)),
        ("dgrad 1024->256, plain store (kEpi 1)",
         mk(dtype=0, nb=nb, h=h, w=w, cin=4 * c, cout=c, srcs=[ops.nhwc_src(dz, nb, h, w, 4 * c)], wpacked=wd, out16=dx,
            ld_out16=c)),
        ("dgrad 1024->256 + un-shuffle store (kEpi 6)",
         mk(dtype=0, nb=nb, h=h, w=w, cin=4 * c, cout=c, srcs=[ops.nhwc_src(dz, nb, h, w, 4 * c)], wpacked=wd, out16=dzu,
            ld_out16=4 * c, out_mode=ops.OUT_UNSHUFFLE2)),
    ]
    print(f"upsampler shape {nb} x {h} x {w}, {fl / 1e9:.0f} GFLOP per launch")
    for name, d in variants:
        t = chain_time(lambda: ops.conv_igemm(d), 10)
        print(f"  {name}: {t:.1f} us ({fl / t / 1e6:.0f} TFLOP/s)")
    part = torch.empty(64 * 9 * 4 * c * c if h == 48 else 48 * 9 * 4 * c * c, device="cuda")
    wg = ops.make_wgrad_desc(dtype=0, nb=nb, h=h, w=w, a=dz, a_c=4 * c, m_total=4 * c, b_srcs=src, n_total=c, partials=part)
    splits = [0]

    def run_wg():
        splits[0] = ops.conv_wgrad(wg)
    t = chain_time(run_wg, 10)
    print(f"  wgrad dW[1024][256][9] ({splits[0]} splits): {t:.1f} us ({fl / t / 1e6:.0f} TFLOP/s)")
    del x, up, flat, dz, dx, dzu, part
