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



# bring-up hooks live in the debug build only: run with PESR_B200_LIB=pesr_b200/libpesr_b200_debug.so (tools/build_debug.sh)
from pesr_b200 import _debug as _dbg, _lib as _L   # noqa: E402
_dbg.bind()


def _set_pair_mode_compat(mode):
    """round-1 encoding of the option hook: 0/1/2 pair mode, 20x sub stages, 30x PDL, 40x staged epilogue, 50x specialised epilogue"""
    if mode >= 500: _L.set_option(_L.OPT_SPECIALISED_EPILOGUE, mode - 500)
    elif mode >= 400: _L.set_option(_L.OPT_STAGED_EPILOGUE, mode - 400)
    elif mode >= 300: _L.set_option(_L.OPT_PDL, mode - 300)
    elif mode >= 200: _L.set_option(_L.OPT_SUB_STAGES, mode - 200)
    elif mode >= 100: pass
    else: _L.set_option(_L.OPT_PAIR_MODE, mode)


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


def disc_case(patch=12, nb=4, seed=0, dtype=torch.float16):
    from oracle import pesr_oracle as O
    from pesr_b200.model import Discriminator
    opt = {'patch_size': patch, 'spectral_norm': False}
    sd = O.init_discriminator(opt, seed)
    D = Discriminator(opt, dtype=dtype)
    D.load_state_dict(sd)
    D = D.cuda().train()
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.rand(nb, 3, patch * 4, patch * 4, generator=g) * 255
    R = torch.randn(nb, 1, generator=g)
    xc = x.cuda().requires_grad_(True)
    y = D(xc)
    (y * R.cuda()).sum().backward()
    torch.cuda.synchronize()
    for name, qd in (("fp64", None), ("qmatch", dtype)):
        leaf = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
                for k, v in sd.items()}
        xo = x.double().clone().requires_grad_(True)
        st = []
        yo = O.discriminator_forward(leaf, xo, qdtype=qd, stats_out=st)
        if qd is None:
            stats = st
        names = [k for k in leaf if torch.is_tensor(leaf[k]) and leaf[k].is_floating_point() and leaf[k].requires_grad]
        og = torch.autograd.grad((yo * R.double()).sum(), [leaf[k] for k in names] + [xo])
        ograds = dict(zip(names, og[:-1]))
        errs = {k: rel(p.grad.cpu(), ograds[k]) for k, p in D.named_parameters()}
        vals = sorted(errs.values())
        print(f"disc patch={patch} nb={nb} vs {name}: logits rel {rel(y.detach().cpu(), yo.detach()):.3e} "
              f"dx rel {rel(xc.grad.cpu(), og[-1]):.3e} grad median {vals[len(vals)//2]:.3e} max {vals[-1]:.3e}")
        print("    " + " ".join(f"{k.replace('features.','f').replace('classifier.','c')}={v:.1e}" for k, v in errs.items()))
    mean0, var0, n0 = stats[0]
    rm = D.state_dict()['features.0.1.running_mean'].cpu()
    rv = D.state_dict()['features.0.1.running_var'].cpu()
    print(f"    running_mean rel {rel(rm, 0.1 * mean0):.2e} running_var rel {rel(rv, 0.9 + 0.1 * var0 * n0 / (n0 - 1)):.2e} "
          f"nbt {int(D.state_dict()['features.0.1.num_batches_tracked'])}", flush=True)


def vgg_case(nb=2, side=64, seed=0, dtype=torch.float16):
    from oracle import pesr_oracle as O
    from pesr_b200.model import VGG
    sd = O.init_vgg(seed)
    V = VGG(pretrained=False, dtype=dtype)
    V.load_state_dict(sd)
    V = V.cuda()
    g = torch.Generator().manual_seed(seed + 1)
    sr = torch.rand(nb, 3, side, side, generator=g) * 255
    hr = torch.rand(nb, 3, side, side, generator=g) * 255
    src = sr.cuda().requires_grad_(True)
    f_sr, f_hr = V(src, hr.cuda())
    loss = ((f_sr - f_hr) ** 2).mean()
    loss.backward()
    torch.cuda.synchronize()
    for name, qd in (("fp64", None), ("qmatch", dtype)):
        so = sr.double().clone().requires_grad_(True)
        of_sr, of_hr = O.vgg_forward({k: v.double() for k, v in sd.items()}, so, hr.double(), qdtype=qd)
        ol = O.mse_loss(of_sr, of_hr)
        og, = torch.autograd.grad(ol, so)
        print(f"vgg nb={nb} side={side} vs {name}: f_sr rel {rel(f_sr.detach().cpu(), of_sr.detach()):.3e} f_hr rel "
              f"{rel(f_hr.cpu(), of_hr):.3e} loss rel {abs(loss.item()-ol.item())/ol.item():.3e} dsr rel "
              f"{rel(src.grad.cpu(), og):.3e}", flush=True)


def vgg_debug(nb=2, side=64, seed=0, dtype=torch.float16):
    """Per-layer comparison of the VGG activations with the quantisation-matched oracle."""
    import torch.nn.functional as F
    from oracle import pesr_oracle as O
    from pesr_b200.model import VGG
    sd = O.init_vgg(seed)
    V = VGG(pretrained=False, dtype=dtype)
    V.load_state_dict(sd)
    V = V.cuda()
    g = torch.Generator().manual_seed(seed + 1)
    sr = torch.rand(nb, 3, side, side, generator=g) * 255
    hr = torch.rand(nb, 3, side, side, generator=g) * 255
    with torch.no_grad():
        V(sr.cuda(), hr.cuda())
    pl = V.engine().plans[(nb, side, side)][-1]
    q = lambda t: t.to(dtype).double()
    x = F.conv2d(sr.double(), sd['sub_mean.weight'].double(), sd['sub_mean.bias'].double())
    idx, k = 0, 0
    for op in pl.ops:
        if op[0] == "conv":
            li, cin, cout, ch, cw, outb = op[1], op[2], op[3], op[4], op[5], op[7]
            x = F.conv2d(q(x), q(sd[f'vgg.{idx}.weight']), sd[f'vgg.{idx}.bias'].double(), padding=1)
            if li < 15:
                x = F.relu(x)
            got = from_nhwc16(outb[: nb * ch * cw], nb, cout, ch, cw).cpu()
            print(f"  conv{li} {cin}->{cout} {ch}x{cw}: rel vs oracle(rounded) {rel(got, q(x)):.3e}  vs unrounded {rel(got, x):.3e} "
                  f"|x| {x.abs().mean():.3f} max {x.abs().max():.1f}")
            x = got.double()   # continue from OUR activations so that each layer is checked in isolation
            idx += 2
        else:
            c, ch, cw, outb = op[1], op[2], op[3], op[5]
            x = F.max_pool2d(x, 2, 2)
            got = from_nhwc16(outb[: nb * (ch // 2) * (cw // 2)], nb, c, ch // 2, cw // 2).cpu()
            print(f"  pool {c} {ch}x{cw}: rel {rel(got, x):.3e}")
            x = got.double()
            idx += 1
    sys.stdout.flush()


def gan_case(seed=0, nb=4, patch=12):
    from oracle import pesr_oracle as O
    from pesr_b200 import steps
    from pesr_b200.model import VGG, Discriminator, Generator
    from pesr_b200.optim import Adam
    opt = {'depth': 2, 'num_channels': 64, 'res_scale': 0.1, 'patch_size': patch, 'spectral_norm': False}
    g_sd, d_sd, v_sd = O.init_generator(opt, seed), O.init_discriminator(opt, seed + 1), O.init_vgg(seed + 2)
    G, D, V = Generator(opt), Discriminator(opt), VGG(pretrained=False)
    G.load_state_dict(g_sd), D.load_state_dict(d_sd), V.load_state_dict(v_sd)
    G, D, V = G.cuda(), D.cuda(), V.cuda()
    gen = torch.Generator().manual_seed(seed + 3)
    lr = torch.rand(nb, 3, patch, patch, generator=gen) * 255
    hr = torch.rand(nb, 3, patch * 4, patch * 4, generator=gen) * 255
    optG, optD = Adam(G.parameters(), lr=5e-5), Adam(D.parameters(), lr=5e-5)
    cfg = dict(steps.DEFAULT_GAN_CFG)
    cfg['target_real'] = torch.ones(nb, 1, device="cuda")
    cfg['target_fake'] = torch.zeros(nb, 1, device="cuda")
    g_before = {k: v.detach().clone() for k, v in G.state_dict().items()}
    losses = steps.gan_step(G, D, V, optG, optD, lr.cuda(), hr.cuda(), cfg)
    torch.cuda.synchronize()
    for name, qd in (("fp64", None), ("qmatch", torch.float16)):
        out = O.gan_step(g_sd, d_sd, v_sd, lr, hr, opt, dtype=torch.float64, qdtype=qd)
        ref = torch.stack([out['l1'], out['vgg'], out['g_loss'], out['tv'], out['d_loss']])
        print(f"gan vs {name}: losses got {[f'{v:.5g}' for v in losses.tolist()]} ref {[f'{v:.5g}' for v in ref.tolist()]}")
        ge = {k: rel(p.grad.cpu(), out['g_grads'][k]) for k, p in G.named_parameters()}
        de = {k: rel(p.grad.cpu(), out['d_grads'][k]) for k, p in D.named_parameters()}
        gv, dv = sorted(ge.values()), sorted(de.values())
        print(f"    G grads rel median {gv[len(gv)//2]:.3e} max {gv[-1]:.3e} | D grads (D phase) median {dv[len(dv)//2]:.3e} "
              f"max {dv[-1]:.3e}")
        dpa = {k: rel(p.detach().cpu() - d_sd[k].double(), out['d_params_after'][k] - d_sd[k].double())
               for k, p in D.named_parameters()}
        print(f"    D param update (Adam step 1) rel median {sorted(dpa.values())[len(dpa)//2]:.3e}", flush=True)


def main():
    what = sys.argv[1]
    if what == "disc":
        disc_case()
        disc_case(patch=8, nb=3, seed=4)
        return
    if what == "vgg":
        vgg_debug()
        vgg_case()
        vgg_case(nb=1, side=48, seed=3)
        return
    if what == "gan":
        gan_case()
        return
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
