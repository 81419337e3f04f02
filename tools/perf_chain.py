"""In-stream cost of the split-K reduction (+ bias gradient) that follows every wgrad launch: time a chain of
[wgrad -> reduce] pairs with ONE event pair and subtract the wgrad-only chain (partials stay L2-hot, PDL active)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pesr_b200 import ops  # noqa: E402
from pesr_b200._lib import check, lib  # noqa: E402

nb, c, h, w = 16, 256, 48, 48
dy16 = torch.randn(nb, h, w, c, device="cuda").half()
x16 = torch.randn(nb, h, w, c, device="cuda").half()
part = torch.empty(32 * 9 * c * c, device="cuda")
d = ops.make_wgrad_desc(dtype=0, nb=nb, h=h, w=w, a=dy16, a_c=c, m_total=c, b_srcs=[ops.nhwc_src(x16, nb, h, w, c)], n_total=c,
                        partials=part)
grad = torch.empty(c, c, 3, 3, device="cuda")
bias = torch.zeros(c, device="cuda")
scale = torch.ones(1, device="cuda")
st = torch.cuda.current_stream().cuda_stream


def chain(mode, n=40):
    for _ in range(n):
        splits = ops.conv_wgrad(d)
        if mode == 1:
            check(lib.pesr_wgrad_reduce_bias(part.data_ptr(), splits, 9, c, c, 0, c, c, 1.0, scale.data_ptr(), 0, grad.data_ptr(),
                                             dy16.data_ptr(), nb * h * w, c, c, 1.0, 0, bias.data_ptr(), bias.data_ptr(), c, st), "rb")
        elif mode == 2:
            ops.wgrad_reduce(part, splits, 9, c, c, 0, c, c, grad, div_dev=scale)
        elif mode == 3:
            ops.wgrad_reduce(part, splits, 9, c, c, 0, c, c, grad, div_dev=scale)
            ops.colsum16(dy16, nb * h * w, c, c, bias, div_dev=scale)


def timed(mode, n=40):
    chain(mode, 5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    chain(mode, n)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


base = timed(0)
print(f"wgrad only: {base:.1f} us per launch")
for mode, name in ((1, "wgrad + reduce_bias"), (2, "wgrad + reduce"), (3, "wgrad + reduce + colsum")):
    t = timed(mode)
    print(f"{name}: {t:.1f} us per pair -> helper costs {t - base:.1f} us in-stream")


def chain_time(fn, n=40):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


print("BatchNorm kernels, back-to-back on L2-hot data (per call):")
for npix, cc in ((16 * 192 * 192, 64), (16 * 96 * 96, 64), (16 * 96 * 96, 128), (16 * 48 * 48, 256), (16 * 24 * 24, 512), (16 * 12 * 12, 512)):
    y = torch.randn(npix, cc, device="cuda").half()
    dz = torch.randn(npix, cc, device="cuda").half()
    a = torch.empty_like(y)
    ws = torch.zeros(2 * cc, device="cuda", dtype=torch.float64)
    mean, rstd = torch.zeros(cc, device="cuda"), torch.ones(cc, device="cuda")
    gam, bet = torch.ones(cc, device="cuda"), torch.zeros(cc, device="cuda")
    dg, db = torch.zeros(cc, device="cuda"), torch.zeros(cc, device="cuda")
    t1 = chain_time(lambda: ops.bn_reduce(y, npix, cc, ws, zero_first=False))
    t2 = chain_time(lambda: ops.bn_lrelu_fwd(y, npix, cc, mean, rstd, gam, bet, a, sums_ws=ws))
    t3 = chain_time(lambda: ops.bn_lrelu_bwd(dz, y, npix, cc, mean, rstd, gam, ws, a, dg, db, zero_first=False))
    mb = npix * cc * 2 / 1e6
    print(f"  {npix}x{cc} ({mb:.0f} MB): stats {t1:.1f} us, apply {t2:.1f} us, backward {t3:.1f} us")

print("trunk conv variants, back-to-back chains of 40 launches (PDL on), per launch:")
xa = torch.randn(nb, h, w, c, device="cuda").half()
xb = torch.empty_like(xa)
wp = (torch.randn(9 * c, c, device="cuda") / (3 * c ** 0.5)).half()
bias_c = torch.randn(c, device="cuda")
s32 = torch.randn(nb, h, w, c, device="cuda")
src_a, src_b = [ops.nhwc_src(xa, nb, h, w, c)], [ops.nhwc_src(xb, nb, h, w, c)]
mk = ops.make_conv_desc
light = [mk(dtype=0, nb=nb, h=h, w=w, cin=c, cout=c, srcs=src_a, wpacked=wp, bias=bias_c, act=1, out16=xb, ld_out16=c),
         mk(dtype=0, nb=nb, h=h, w=w, cin=c, cout=c, srcs=src_b, wpacked=wp, bias=bias_c, act=1, out16=xa, ld_out16=c)]
resid = [mk(dtype=0, nb=nb, h=h, w=w, cin=c, cout=c, srcs=src_a, wpacked=wp, bias=bias_c, alpha=0.1, res32=s32, ld_res32=c,
            out32=s32, ld_out32=c, out16=xb, ld_out16=c),
         mk(dtype=0, nb=nb, h=h, w=w, cin=c, cout=c, srcs=src_b, wpacked=wp, bias=bias_c, alpha=0.1, res32=s32, ld_res32=c,
            out32=s32, ld_out32=c, out16=xa, ld_out16=c)]
mask = [mk(dtype=0, nb=nb, h=h, w=w, cin=c, cout=c, srcs=src_a, wpacked=wp, mask16=dy16, ld_mask16=c, mask_mode=1, out16=xb,
           ld_out16=c),
        mk(dtype=0, nb=nb, h=h, w=w, cin=c, cout=c, srcs=src_b, wpacked=wp, mask16=dy16, ld_mask16=c, mask_mode=1, out16=xa,
           ld_out16=c)]
fl = 2.0 * nb * h * w * c * c * 9
for name, pair in (("light (bias+relu -> fp16)", light), ("residual (fp32 stream)", resid), ("mask (dgrad * relu')", mask)):
    def two():
        ops.conv_igemm(pair[0])
        ops.conv_igemm(pair[1])
    t = chain_time(two, 20) / 2
    print(f"  {name}: {t:.1f} us ({fl / t / 1e6:.0f} TFLOP/s)")
    xa.normal_()
