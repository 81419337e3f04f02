"""Kernel schedule of the VGG19 feature extractor used by the perceptual loss (reference:
model/vgg.py:5-28 = sub_mean + torchvision vgg19.features[:35], conv5_4 before its ReLU).

`vgg(sr, hr)` runs BOTH images as one batch of 2N through the implicit-GEMM kernel (bias + ReLU fused;
there is no BatchNorm in VGG19, so batching the two calls is exact), keeps the sr half's activations
for backward, and back-propagates through the sr half only (the reference evaluates hr under no_grad,
model/vgg.py:25-26).  Backward is dgrad only: no optimiser owns VGG's weights (train.py:123-126), so
the weight gradients the reference computes as a side effect of `self.vgg.requires_grad = False` being
a no-op (model/vgg.py:16) are never consumed and are not computed here.
"""
import torch

from . import ops
from .engine_g import PackedWeight, PlanCache, _Plan, _Release, _run_conv
from .ops import ACT_NONE, ACT_RELU

VGG19_CFG = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 256, 'M', 512, 512, 512, 512, 'M', 512, 512, 512, 512]


class VGGEngine:
    def __init__(self, vgg, dtype=torch.float16):
        self.vgg = vgg
        self.dtype = dtype
        self.dt = ops.dt_code(dtype)
        self.plans = PlanCache(who="pesr_b200.VGG")
        self.trace_hook = None     # callable(plan), called at the end of every forward (parity tests read the saved activations)
        self.packed = None
        self.device = None

    def invalidate_packs(self):
        """See GeneratorEngine.invalidate_packs."""
        if self.packed is not None:
            self.fwd_multi.key = None
            self.bwd_multi.key = None

    def _conv_modules(self):
        return [m for m in self.vgg.vgg if isinstance(m, torch.nn.Conv2d)]

    def _ensure_packed(self, device):
        convs = self._conv_modules()
        sentinel = (convs[0].weight.data_ptr(), convs[-1].bias.data_ptr())
        if self.packed is not None and self.device == device and self.sentinel == sentinel:
            return
        self.sentinel, self.device = sentinel, device
        self.plans.clear()
        dt = self.dtype
        self.pk_f = [PackedWeight(convs[0].weight, 4, dt, pad_to=64)] + [PackedWeight(c.weight, 0, dt) for c in convs[1:]]
        self.pk_d = [PackedWeight(convs[0].weight, 6, dt, pad_to=32)] + [PackedWeight(c.weight, 1, dt) for c in convs[1:]]
        self.packed = True
        self.fwd_multi = ops.MultiPack(self.pk_f, device, dt)
        self.bwd_multi = ops.MultiPack(self.pk_d, device, dt)
        self.scale_ws = torch.zeros(4, device=device, dtype=torch.float32)

    def _plan(self, nb, h, w):
        return self.plans.acquire((nb, h, w), lambda: self._new_plan(nb, h, w))

    def _new_plan(self, nb, h, w):
        dev, tdt, dt = self.device, self.dtype, self.dt
        convs = self._conv_modules()
        nb2 = 2 * nb
        pl = _Plan()
        pl.nb, pl.h, pl.w = nb, h, w
        e16 = lambda *s: torch.empty(*s, device=dev, dtype=tdt)  # noqa: E731
        pl.col0 = e16(nb2 * h * w, 64)
        # layer table: ("conv", idx, cin, cout, h, w, in_buf, out_buf) / ("pool", c, h, w, in_buf, out_buf)
        ops_list, fwd = [], []
        cur, ch, cw, cin, li = pl.col0, h, w, 64, 0
        n_conv = len(convs)
        for v in VGG19_CFG:
            if v == 'M':
                out = e16(nb2 * (ch // 2) * (cw // 2), cin)
                ops_list.append(("pool", cin, ch, cw, cur, out))
                cur, ch, cw = out, ch // 2, cw // 2
                continue
            out = e16(nb2 * ch * cw, v)
            last = li == n_conv - 1
            if li == 0:
                d = ops.make_conv_desc(dtype=dt, nb=nb2, h=ch, w=cw, cin=64, cout=v, taps=[(0, 0)],
                                       srcs=[ops.nhwc_src(cur, nb2, ch, cw, 64)], wpacked=self.pk_f[0].buf,
                                       bias=convs[0].bias, act=ACT_RELU, out16=out, ld_out16=v)
            else:
                d = ops.make_conv_desc(dtype=dt, nb=nb2, h=ch, w=cw, cin=cin, cout=v,
                                       srcs=[ops.nhwc_src(cur, nb2, ch, cw, cin)], wpacked=self.pk_f[li].buf,
                                       bias=convs[li].bias, act=ACT_NONE if last else ACT_RELU, out16=out, ld_out16=v)
            ops_list.append(("conv", li, cin, v, ch, cw, cur, out, d))
            cur, cin, li = out, v, li + 1
        pl.ops = ops_list
        pl.feat16, pl.fh, pl.fw, pl.fc = cur, ch, cw, cin
        # ---- backward (sr half = first nb images of every buffer)
        max_elems = max(nb * o[4] * o[5] * max(o[2], o[3]) for o in ops_list if o[0] == "conv")
        pl.gA, pl.gB = e16(max_elems), e16(max_elems)
        pl.Zd = torch.empty(nb * h * w, 32, device=dev, dtype=torch.float32)
        bwd = []
        g_out, g_in = pl.gA, pl.gB      # g_out holds dL/d(conv output, pre-ReLU mask applied)
        for idx in range(len(ops_list) - 1, -1, -1):
            o = ops_list[idx]
            if o[0] == "pool":
                continue   # handled together with the conv that follows it (see below)
            _, li, cin, cout, ch, cw, inb, outb, _d = o
            if li == 0:
                bwd.append(("conv", ops.make_conv_desc(dtype=dt, nb=nb, h=ch, w=cw, cin=cout, cout=32, taps=[(0, 0)],
                                                       srcs=[ops.nhwc_src(g_out, nb, ch, cw, cout)],
                                                       wpacked=self.pk_d[0].buf, out32=pl.Zd, ld_out32=32)))
                break
            prev = ops_list[idx - 1]
            if prev[0] == "conv":
                # input of this conv is the previous conv's ReLU output: fuse relu' into the dgrad epilogue
                bwd.append(("conv", ops.make_conv_desc(dtype=dt, nb=nb, h=ch, w=cw, cin=cout, cout=cin,
                                                       srcs=[ops.nhwc_src(g_out, nb, ch, cw, cout)],
                                                       wpacked=self.pk_d[li].buf, mask16=prev[7], ld_mask16=cin,
                                                       mask_mode=1, out16=g_in, ld_out16=cin)))
                g_out, g_in = g_in, g_out
            else:
                # input is a pooled tensor: dgrad -> dPool, then route through the max-pool (+ relu' of the conv before it)
                _, pc, ph_, pw_, pin, _pout = prev
                bwd.append(("conv", ops.make_conv_desc(dtype=dt, nb=nb, h=ch, w=cw, cin=cout, cout=cin,
                                                       srcs=[ops.nhwc_src(g_out, nb, ch, cw, cout)],
                                                       wpacked=self.pk_d[li].buf, out16=g_in, ld_out16=cin)))
                bwd.append(("pool", pin, g_in, nb, ph_, pw_, pc, g_out))
                # after the pool backward the gradient lives in g_out again (at the pre-pool resolution)
        pl.bwd = bwd
        return pl

    # ------------------------------------------------------------------ forward
    def forward(self, sr, hr, save):
        if sr.shape != hr.shape or sr.dim() != 4 or sr.shape[1] != 3:
            raise ValueError(f"VGG expects two [N,3,H,W] tensors of one shape, got {tuple(sr.shape)} and {tuple(hr.shape)}")
        sr = sr.contiguous().float()
        hr = hr.detach().contiguous().float()
        nb, _, h, w = sr.shape
        self._ensure_packed(sr.device)
        pl = self._plan(nb, h, w)
        self.fwd_multi.run()
        sm = self.vgg.sub_mean
        sm_w = sm.weight.detach().reshape(3, 3)
        P = nb * h * w
        ops.im2col3(sr, pl.col0[:P], affine_a=sm_w, affine_b=sm.bias.detach())
        ops.im2col3(hr, pl.col0[P:], affine_a=sm_w, affine_b=sm.bias.detach())
        stream = torch.cuda.current_stream().cuda_stream
        for o in pl.ops:
            if o[0] == "conv":
                _run_conv(o[8], stream)
            else:
                _, c, ch, cw, inb, outb = o
                ops.maxpool2_fwd(inb, 2 * nb, ch, cw, c, outb)
        f_sr = torch.empty(nb, pl.fc, pl.fh, pl.fw, device=sr.device, dtype=torch.float32)
        f_hr = torch.empty(nb, pl.fc, pl.fh, pl.fw, device=sr.device, dtype=torch.float32)
        n_half = nb * pl.fh * pl.fw
        ops.nhwc16_to_nchw32(pl.feat16[:n_half], f_sr)
        ops.nhwc16_to_nchw32(pl.feat16[n_half:], f_hr)
        if self.trace_hook is not None:
            self.trace_hook(pl)
        if not save:
            return f_sr, f_hr, None
        pl.busy = True
        return f_sr, f_hr, (pl, _Release(pl))

    # ------------------------------------------------------------------ backward (to sr only)
    def backward(self, state, dfeat):
        pl, _release = state
        self.bwd_multi.run()
        nb, h, w = pl.nb, pl.h, pl.w
        dfeat = dfeat.contiguous().float()
        scale = self.scale_ws[1:2]
        ops.amax_scale(dfeat, self.scale_ws, target=16.0)
        ops.nchw32_to_nhwc16(dfeat, pl.gA, mul_dev=scale)
        stream = torch.cuda.current_stream().cuda_stream
        for op in pl.bwd:
            if op[0] == "conv":
                _run_conv(op[1], stream)
            else:
                _, x16, dy16, n, ph_, pw_, pc, dx16 = op
                ops.maxpool2_bwd(x16, dy16, n, ph_, pw_, pc, dx16, relu_mask=True)
        sm_wt = self.vgg.sub_mean.weight.detach().reshape(3, 3).t().contiguous()
        dsr = torch.empty(nb, 3, h, w, device=dfeat.device, dtype=torch.float32)
        ops.col2im3(pl.Zd, 32, nb, h, w, dsr, affine_a=sm_wt, mul=1.0, div_dev=scale, sgn=-1)
        return dsr
