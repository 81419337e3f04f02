"""GPU bring-up diagnostics for the tcgen05 convolution kernels (not a test: prints error tables).

usage: python tools/bringup.py fprop|wgrad|layout [options]
Each case runs against torch's fp32 conv on the same 16-bit-rounded operands.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.nn.functional as F

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

from pesr_b200 import ops  # noqa: E402
from pesr_b200._lib import lib  # noqa: E402


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def to_nhwc16(x, dtype):
    nb, c, h, w = x.shape
    out = torch.empty(nb, h, w, c, device="cuda", dtype=dtype)
    ops.nchw32_to_nhwc16(x.contiguous(), out)
    return out


def from_nhwc16(x, nb, c, h, w):
    out = torch.empty(nb, c, h, w, device="cuda", dtype=torch.float32)
    ops.nhwc16_to_nchw32(x, out)
    return out


def fprop_case(nb, cin, cout, h, w, dtype=torch.float16, bias=True, act=0, single_tap=None, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(nb, cin, h, w, device="cuda", generator=g)
    wt = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3.0 * cin ** 0.5)
    if single_tap is not None:
        m = torch.zeros(3, 3, device="cuda")
        m.view(-1)[single_tap] = 1
        wt = wt * m
    b = torch.randn(cout, device="cuda", generator=g) if bias else None
    xr = x.to(dtype).float()
    wr = wt.to(dtype).float()
    ref = F.conv2d(xr, wr, b, padding=1)
    if act == 1:
        ref = ref.relu()
    x16 = to_nhwc16(x, dtype)
    chk = rel(from_nhwc16(x16, nb, cin, h, w), xr)
    wp = torch.empty(ops.packed_shape(cout, cin, 3, 0), device="cuda", dtype=dtype)
    ops.pack_weights(wt, 0, wp)
    wp_ref = wr.permute(2, 3, 0, 1).reshape(9 * cout, cin)
    chk_w = rel(wp.float(), wp_ref)
    out16 = torch.zeros(nb, h, w, cout, device="cuda", dtype=dtype)
    out32 = torch.zeros(nb, h, w, cout, device="cuda", dtype=torch.float32)
    d = ops.make_conv_desc(dtype=ops.dt_code(dtype), nb=nb, h=h, w=w, cin=cin, cout=cout,
                           srcs=[ops.nhwc_src(x16, nb, h, w, cin)], wpacked=wp, bias=b, act=act,
                           out16=out16, ld_out16=cout, out32=out32, ld_out32=cout)
    ops.conv_igemm(d)
    torch.cuda.synchronize()
    got32 = out32.permute(0, 3, 1, 2)
    got16 = from_nhwc16(out16, nb, cout, h, w)
    e32, e16 = rel(got32, ref), rel(got16, ref)
    print(f"fprop nb={nb} {cin}->{cout} {h}x{w} {str(dtype)[6:]} tap={single_tap} act={act}: layout_chk={chk:.1e} "
          f"pack_chk={chk_w:.1e} rel32={e32:.3e} rel16={e16:.3e} |ref|={ref.abs().mean().item():.3f} "
          f"|got|={got32.abs().mean().item():.3f}", flush=True)
    if e32 > 1e-3:
        err = (got32 - ref).abs()
        print("   err by image:", [f"{v:.2e}" for v in err.mean(dim=(1, 2, 3)).tolist()])
        print("   err by row  :", [f"{v:.1e}" for v in err.mean(dim=(0, 1, 3)).tolist()][:32])
        print("   err by col  :", [f"{v:.1e}" for v in err.mean(dim=(0, 1, 2)).tolist()][:32])
        ec = err.mean(dim=(0, 2, 3))
        print("   err by cout (first 16):", [f"{v:.1e}" for v in ec.tolist()[:16]], " worst", int(ec.argmax()))
    return e32


def wgrad_case(nb, cin, cout, h, w, dtype=torch.float16, lbo=0, sbo=0, seed=0, single_tap=None):
    lib.pesr_debug_wgrad_desc(lbo, sbo)
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(nb, cin, h, w, device="cuda", generator=g)
    dy = torch.randn(nb, cout, h, w, device="cuda", generator=g)
    xr, dyr = x.to(dtype).float(), dy.to(dtype).float()
    ref = torch.nn.grad.conv2d_weight(xr, (cout, cin, 3, 3), dyr, padding=1)
    x16, dy16 = to_nhwc16(x, dtype), to_nhwc16(dy, dtype)
    taps = ops.TAPS_3X3 if single_tap is None else [ops.TAPS_3X3[single_tap]]
    part = torch.zeros(64 * len(taps) * cout * cin, device="cuda", dtype=torch.float32)
    d = ops.make_wgrad_desc(dtype=ops.dt_code(dtype), nb=nb, h=h, w=w, a=dy16, a_c=cout, m_total=cout,
                            b_srcs=[ops.nhwc_src(x16, nb, h, w, cin)], n_total=cin, taps=taps, partials=part)
    splits = ops.conv_wgrad(d)
    torch.cuda.synchronize()
    if single_tap is None:
        grad = torch.zeros(cout, cin, 3, 3, device="cuda")
        ops.wgrad_reduce(part, splits, 9, cout, cin, ops.WMAP_OIHW, cout, cin, grad)
        torch.cuda.synchronize()
        e = rel(grad, ref)
    else:
        got = part[: splits * cout * cin].view(splits, cout, cin).sum(0)
        e = rel(got, ref[:, :, single_tap // 3, single_tap % 3])
    print(f"wgrad nb={nb} cin={cin} cout={cout} {h}x{w} {str(dtype)[6:]} lbo={lbo} sbo={sbo} tap={single_tap} "
          f"splits={splits}: rel={e:.3e}", flush=True)
    return e


def gen_case(opt, shape, dtype=torch.float16, seed=0, qmatch=True):
    from oracle import pesr_oracle as O
    from pesr_b200.model import Generator
    sd = O.init_generator(opt, seed)
    G = Generator(opt, dtype=dtype)
    G.load_state_dict(sd)
    G = G.cuda()
    g = torch.Generator().manual_seed(seed + 1)
    lr = torch.rand(*shape, generator=g) * 255
    hr = torch.rand(shape[0], 3, shape[2] * 4, shape[3] * 4, generator=g) * 255
    with torch.no_grad():
        sr_inf = G(lr.cuda())
    sr = G(lr.cuda())
    loss = (sr - hr.cuda()).abs().mean()
    loss.backward()
    torch.cuda.synchronize()
    print(f"gen {opt} {shape} {str(dtype)[6:]}: train/infer sr diff {rel(sr_inf, sr.detach()):.2e}")
    for name, kw in (("fp64", dict(dtype=torch.float64)), ("qmatch", dict(dtype=torch.float64, qdtype=dtype))):
        if name == "qmatch" and not qmatch:
            continue
        ol, osr, og = O.pretrain_step(sd, lr, hr, opt, **kw)
        errs = {k: rel(p.grad.cpu(), og[k]) for k, p in G.named_parameters()}
        vals = sorted(errs.values())
        worst = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
        print(f"  vs {name}: sr rel {rel(sr.detach().cpu(), osr):.3e}  loss rel {abs(loss.item() - ol.item()) / ol.item():.3e}  "
              f"grad rel median {vals[len(vals) // 2]:.3e} max {vals[-1]:.3e}")
        print("   worst:", ", ".join(f"{k}={v:.2e}" for k, v in worst))
        for k in ("sub_mean.weight", "sub_mean.bias", "embed.weight", "embed.bias", "upsample.0.weight",
                  "upsample.0.bias", "upsample.2.weight", "upsample.4.weight", "upsample.4.bias", "add_mean.weight",
                  "add_mean.bias", f"body.{opt['depth']}.weight", "body.0.body.0.weight", "body.0.body.2.weight"):
            print(f"     {k}: {errs[k]:.2e}", end="")
        print(flush=True)


def main():
    what = sys.argv[1]
    if what == "gen":
        gen_case({'depth': 2, 'num_channels': 64, 'res_scale': 0.1}, (2, 3, 10, 12))
        gen_case({'depth': 2, 'num_channels': 64, 'res_scale': 0.1}, (2, 3, 10, 12), dtype=torch.bfloat16)
        gen_case({'depth': 4, 'num_channels': 128, 'res_scale': 0.1}, (1, 3, 24, 24))
        gen_case({'depth': 32, 'num_channels': 256, 'res_scale': 0.1}, (1, 3, 16, 16), qmatch=False)
        return
    if what == "layout":
        x = torch.randn(2, 40, 10, 12, device="cuda")
        y = from_nhwc16(to_nhwc16(x, torch.float16), 2, 40, 10, 12)
        print("layout roundtrip rel", rel(y, x.half().float()))
    elif what == "fprop":
        fprop_case(1, 64, 64, 8, 16, bias=False)
        fprop_case(1, 64, 64, 8, 16, bias=False, single_tap=4)
        fprop_case(1, 64, 64, 8, 16, bias=False, single_tap=0)
        fprop_case(1, 64, 64, 8, 16, bias=False, single_tap=8)
        fprop_case(2, 64, 64, 16, 16)
        fprop_case(2, 128, 256, 16, 32, act=1)
        fprop_case(2, 256, 256, 48, 48, act=1)
        fprop_case(2, 256, 256, 48, 48, dtype=torch.bfloat16)
        fprop_case(1, 64, 128, 24, 24)
        fprop_case(1, 512, 512, 12, 12)
        fprop_case(1, 256, 1024, 20, 20)
        fprop_case(16, 256, 256, 48, 48)
    elif what == "wgrad":
        lbo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
        sbo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
        wgrad_case(1, 64, 128, 8, 16, lbo=lbo, sbo=sbo, single_tap=4)
        wgrad_case(1, 64, 128, 8, 16, lbo=lbo, sbo=sbo, single_tap=0)
        wgrad_case(1, 256, 256, 16, 16, lbo=lbo, sbo=sbo, single_tap=4)
        wgrad_case(2, 256, 256, 48, 48, lbo=lbo, sbo=sbo)
        wgrad_case(2, 64, 64, 24, 24, lbo=lbo, sbo=sbo)
        wgrad_case(1, 128, 256, 20, 20, lbo=lbo, sbo=sbo, dtype=torch.bfloat16)
        wgrad_case(16, 256, 256, 48, 48, lbo=lbo, sbo=sbo)


if __name__ == "__main__":
    main()
