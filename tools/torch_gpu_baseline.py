"""Informative comparator (NOT the reference arm): the PESR networks written with stock torch.nn modules, run on the
B200 through torch / cuDNN -- the de-facto Blackwell implementation a user of the reference would get by running it
unchanged.  Two precisions: fp32 parameters with TF32 convolutions (torch's default), and bf16
autocast with channels_last.  Prints one JSON object: GAN step, pretrain step, x4 inference, and a per-shape table of
cuDNN convolution times next to this library's kernels for the trunk / upsampler / Discriminator / VGG shapes.

Self-contained on purpose: it imports neither oracle/ (test infrastructure) nor pesr_b200's engines for the torch arm.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
import torch.nn.functional as F

B, P, C, DEPTH = 16, 48, 256, 32


class ResBlock(nn.Module):
    def __init__(s, c):
        super().__init__()
        s.a, s.b = nn.Conv2d(c, c, 3, padding=1), nn.Conv2d(c, c, 3, padding=1)

    def forward(s, x):
        return x + 0.1 * s.b(F.relu(s.a(x)))


class Gen(nn.Module):
    def __init__(s):
        super().__init__()
        s.sub, s.add = nn.Conv2d(3, 3, 1), nn.Conv2d(3, 3, 1)
        s.embed = nn.Conv2d(3, C, 3, padding=1)
        s.body = nn.Sequential(*[ResBlock(C) for _ in range(DEPTH)], nn.Conv2d(C, C, 3, padding=1))
        s.up = nn.Sequential(nn.Conv2d(C, 4 * C, 3, padding=1), nn.PixelShuffle(2), nn.Conv2d(C, 4 * C, 3, padding=1),
                             nn.PixelShuffle(2), nn.Conv2d(C, 3, 3, padding=1))

    def forward(s, x):
        x = s.embed(s.sub(x))
        return s.add(s.up(s.body(x) + x))


class Disc(nn.Module):
    def __init__(s):
        super().__init__()
        cfg = [(3, 64, 1), (64, 64, 2), (64, 128, 1), (128, 128, 2), (128, 256, 1), (256, 256, 2), (256, 512, 1), (512, 512, 2)]
        s.f = nn.Sequential(*[nn.Sequential(nn.Conv2d(a, b, 3, stride=st, padding=1, bias=False), nn.BatchNorm2d(b),
                                            nn.LeakyReLU(0.2, True)) for a, b, st in cfg])
        s.c = nn.Sequential(nn.Linear(512 * 12 * 12, 1024), nn.LeakyReLU(0.2, True), nn.Linear(1024, 1))

    def forward(s, x):
        return s.c(s.f(x).flatten(1))


def vgg35():
    cfg = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 256, 'M', 512, 512, 512, 512, 'M', 512, 512, 512, 512]
    layers, cin = [], 3
    for v in cfg:
        if v == 'M':
            layers.append(nn.MaxPool2d(2, 2))
        else:
            layers += [nn.Conv2d(cin, v, 3, padding=1), nn.ReLU(True)]
            cin = v
    return nn.Sequential(*layers[:35])


def timed(fn, n, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run_mode(mode, steps):
    dev = torch.device("cuda")
    torch.manual_seed(0)
    G, D, V = Gen().to(dev), Disc().to(dev), vgg35().to(dev)
    for p in V.parameters():
        p.requires_grad_(False)
    cl = mode == "bf16_autocast_channels_last"
    if cl:
        G, D, V = (m.to(memory_format=torch.channels_last) for m in (G, D, V))
    oG = torch.optim.Adam(G.parameters(), lr=5e-5, fused=True)
    oD = torch.optim.Adam(D.parameters(), lr=5e-5, fused=True)
    lr = torch.rand(B, 3, P, P, device=dev) * 255
    hr = torch.rand(B, 3, 4 * P, 4 * P, device=dev) * 255
    if cl:
        lr, hr = lr.contiguous(memory_format=torch.channels_last), hr.contiguous(memory_format=torch.channels_last)
    ctx = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if cl else (lambda: torch.autocast("cuda", enabled=False))
    ones = torch.ones(B, 1, device=dev)

    def focal(x):
        p = torch.sigmoid(x)
        return ((1 - p) * F.softplus(-x)).mean()

    def gan():
        for p in D.parameters():
            p.requires_grad_(True)
        oD.zero_grad(set_to_none=True)
        with ctx():
            sr = G(lr)
            dl = F.binary_cross_entropy_with_logits(D(hr).float() - D(sr.detach()).float(), ones)
        dl.backward()
        oD.step()
        for p in D.parameters():
            p.requires_grad_(False)
        oG.zero_grad(set_to_none=True)
        with ctx():
            pf, pr = D(sr).float(), D(hr).float()
            with torch.no_grad():
                fh = V(hr)
            vl = F.mse_loss(V(sr).float(), fh.float()) * 50
            srf = sr.float()
            tv = ((srf[..., :-1] - srf[..., 1:]).abs().sum() + (srf[..., :-1, :] - srf[..., 1:, :]).abs().sum()) * 1e-6
            tot = vl + focal(pf - pr) + tv
        tot.backward()
        oG.step()

    def pre():
        oG.zero_grad(set_to_none=True)
        with ctx():
            loss = (G(lr).float() - hr).abs().mean()
        loss.backward()
        oG.step()
    out = {}
    ms = timed(gan, steps)
    out["gan_step"] = {"ms_per_step": ms, "samples_per_s": B / ms * 1e3}
    ms = timed(pre, steps)
    out["pretrain_step"] = {"ms_per_step": ms, "samples_per_s": B / ms * 1e3}
    G.eval()
    x = torch.rand(1, 3, 339, 510, device=dev) * 255
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    with torch.no_grad(), ctx():
        ms = timed(lambda: G(x), 5, warm=2)
    out["inference_339x510"] = {"ms_per_image": ms, "hr_mpix_per_s": 16 * 339 * 510 / ms / 1e3}
    return out


def conv_table():
    """cuDNN (fp16 channels_last, the fastest torch path) vs this library, per shape: fprop / dgrad / wgrad."""
    from pesr_b200 import ops
    dev = torch.device("cuda")
    shapes = [("G trunk 256->256 @48x48 x16", 16, 256, 256, 48), ("G upsample.0 256->1024 @48x48 x16", 16, 256, 1024, 48),
              ("G upsample.2 256->1024 @96x96 x16", 16, 256, 1024, 96), ("D 64->128 @96x96 x32", 32, 64, 128, 96),
              ("D 128->256 @48x48 x32", 32, 128, 256, 48), ("D 256->512 @24x24 x32", 32, 256, 512, 24),
              ("VGG 64->64 @192x192 x32", 32, 64, 64, 192), ("VGG 128->128 @96x96 x32", 32, 128, 128, 96),
              ("VGG 256->256 @48x48 x32", 32, 256, 256, 48), ("VGG 512->512 @24x24 x32", 32, 512, 512, 24)]
    rows = []
    for name, nb, ci, co, hw in shapes:
        flop = 2.0 * nb * hw * hw * ci * co * 9
        x = torch.randn(nb, ci, hw, hw, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        w = (torch.randn(co, ci, 3, 3, device=dev, dtype=torch.float16) / (3 * ci ** 0.5)).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        y = F.conv2d(x, w, padding=1)
        gy = torch.randn_like(y)
        t_f = timed(lambda: F.conv2d(x, w, padding=1), 20)
        t_d = timed(lambda: torch.autograd.grad(y, x, gy, retain_graph=True), 20)
        t_w = timed(lambda: torch.autograd.grad(y, w, gy, retain_graph=True), 20)
        # this library: NHWC 16-bit operands, packed weights
        x16 = x.detach().permute(0, 2, 3, 1).contiguous()
        wf = torch.empty(9 * co, ci, device=dev, dtype=torch.float16)
        ops.pack_weights(w.detach().float().contiguous(), 0, wf)
        wd = torch.empty(9 * ci, co, device=dev, dtype=torch.float16)
        ops.pack_weights(w.detach().float().contiguous(), 1, wd)
        y16 = torch.empty(nb, hw, hw, co, device=dev, dtype=torch.float16)
        gx16 = torch.empty(nb, hw, hw, ci, device=dev, dtype=torch.float16)
        gy16 = gy.permute(0, 2, 3, 1).contiguous()
        df = ops.make_conv_desc(dtype=0, nb=nb, h=hw, w=hw, cin=ci, cout=co, srcs=[ops.nhwc_src(x16, nb, hw, hw, ci)], wpacked=wf,
                                out16=y16, ld_out16=co)
        dd = ops.make_conv_desc(dtype=0, nb=nb, h=hw, w=hw, cin=co, cout=ci, srcs=[ops.nhwc_src(gy16, nb, hw, hw, co)], wpacked=wd,
                                out16=gx16, ld_out16=ci)
        part = torch.empty(max(9 * co * ci * 8, 148 * 128 * 64 * 4), device=dev)
        dw = ops.make_wgrad_desc(dtype=0, nb=nb, h=hw, w=hw, a=gy16, a_c=co, m_total=co, b_srcs=[ops.nhwc_src(x16, nb, hw, hw, ci)],
                                 n_total=ci, partials=part)
        gw = torch.empty(co, ci, 3, 3, device=dev)

        def ours_w():
            s = ops.conv_wgrad(dw)
            ops.wgrad_reduce(part, s, 9, co, ci, ops.WMAP_OIHW, co, ci, gw)
        o_f, o_d, o_w = timed(lambda: ops.conv_igemm(df), 20), timed(lambda: ops.conv_igemm(dd), 20), timed(ours_w, 20)
        rows.append({"shape": name, "gflop": flop / 1e9,
                     "cudnn_us": {"fprop": t_f * 1e3, "dgrad": t_d * 1e3, "wgrad": t_w * 1e3},
                     "pesr_b200_us": {"fprop": o_f * 1e3, "dgrad": o_d * 1e3, "wgrad_incl_reduce": o_w * 1e3},
                     "pesr_b200_tflops": {"fprop": flop / o_f / 1e9, "dgrad": flop / o_d / 1e9, "wgrad": flop / o_w / 1e9},
                     "cudnn_tflops": {"fprop": flop / t_f / 1e9, "dgrad": flop / t_d / 1e9, "wgrad": flop / t_w / 1e9}})
    return rows


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    torch.backends.cudnn.benchmark = True
    out = {"what": "stock torch.nn PESR on the B200 through cuDNN (informative comparator, see BASELINE.md section 3)",
           "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(), "gpu": torch.cuda.get_device_name(0)}
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    out["fp32_tf32"] = run_mode("fp32_tf32", steps)
    out["bf16_autocast_channels_last"] = run_mode("bf16_autocast_channels_last", steps)
    out["conv_table_fp16_channels_last"] = conv_table()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
