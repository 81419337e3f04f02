"""Split-precision schedule of the VGG19 feature extractor (reference: model/vgg.py:5-28): three tensor-core passes per
convolution on fp16 hi + lo operands (engine_split.py), fp32 activations between the layers; ReLU and 2x2 max-pool act
on fp32.  Backward is the input gradient of the sr branch only (the weights are frozen, see engine_v.py)."""
import torch

from . import ops
from .engine_split import HL, SplitOps, SplitWeight
from .engine_v import VGG19_CFG
from .ops import ACT_RELU


class SplitVGGEngine(SplitOps):
    def __init__(self, vgg):
        self.vgg = vgg
        self.packed = None
        self.device = None
        self.trace_hook = None

    def invalidate_packs(self):
        if self.packed is not None:
            for sw in self.packed.values():
                sw.key = None

    def _convs(self):
        return [m for m in self.vgg.vgg if isinstance(m, torch.nn.Conv2d)]

    def _ensure_packed(self, device):
        convs = self._convs()
        sentinel = (convs[0].weight.data_ptr(), convs[-1].bias.data_ptr())
        if self.packed is not None and self.device == device and self.sentinel == sentinel:
            return
        self.sentinel, self.device = sentinel, device
        pk = {"c0_f": SplitWeight(convs[0].weight, 4, pad_to=64), "c0_d": SplitWeight(convs[0].weight, 6, pad_to=32)}
        for i, c in enumerate(convs[1:], start=1):
            pk[f"c{i}_f"], pk[f"c{i}_d"] = SplitWeight(c.weight, 0), SplitWeight(c.weight, 1)
        self.packed = pk
        self.scale_ws = torch.zeros(4, device=device, dtype=torch.float32)

    def _features(self, x, keep):
        """x: [n,3,h,w] fp32.  Returns (features fp32 NHWC [n*fh*fw][512], fh, fw, saved layer list)."""
        convs = self._convs()
        n, _, h, w = x.shape
        dev = x.device
        sm = self.vgg.sub_mean
        sm_w = sm.weight.detach().reshape(3, 3)
        e32 = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)  # noqa: E731
        col = HL(n * h * w, 64, dev)
        ops.im2col3(x, col.hi, affine_a=sm_w, affine_b=sm.bias.detach())
        ops.im2col3(x, col.lo, affine_a=sm_w, affine_b=sm.bias.detach(), low_part=True)
        saved = []            # ("conv", li, cin, cout, h, w, input HL, output HL or None) / ("pool", c, h, w, x32)
        cur, ch, cw, cin, li = col, h, w, 64, 0
        n_conv = len(convs)
        acc = None
        for v in VGG19_CFG:
            if v == 'M':
                # the pool input is relu(acc) of the previous conv: pooled in fp32, then split for the next conv
                act32 = e32(n * ch * cw, cin)
                ops.affine_split(acc, n * ch * cw, cin, out32=act32, act=ACT_RELU)
                pooled = e32(n * (ch // 2) * (cw // 2), cin)
                ops.maxpool2_f32_fwd(act32, n, ch, cw, cin, pooled)
                nxt = HL(pooled.shape[0], cin, dev)
                ops.split16(pooled, nxt.hi, nxt.lo)
                saved.append(("pool", cin, ch, cw, act32 if keep else None))
                cur, ch, cw = nxt, ch // 2, cw // 2
                continue
            acc = e32(n * ch * cw, v)
            if li == 0:
                self._conv3(cur, "c0_f", n, ch, cw, 64, v, acc, bias=convs[0].bias.detach(), taps=[(0, 0)])
            else:
                self._conv3(cur, f"c{li}_f", n, ch, cw, cin, v, acc, bias=convs[li].bias.detach())
            last = li == n_conv - 1
            out = None
            nxt_is_pool = False
            if not last:
                idx = [k for k, u in enumerate(VGG19_CFG) if u != 'M'][li]
                nxt_is_pool = VGG19_CFG[idx + 1] == 'M'
                if not nxt_is_pool:
                    out = HL(n * ch * cw, v, dev)
                    ops.split16(acc, out.hi, out.lo, act=ACT_RELU)
            saved.append(("conv", li, cin, v, ch, cw, cur if keep else None, out if keep else None))
            if out is not None:
                cur = out
            cin, li = v, li + 1
        return acc, ch, cw, saved

    def forward(self, sr, hr, save):
        if sr.shape != hr.shape or sr.dim() != 4 or sr.shape[1] != 3:
            raise ValueError(f"VGG expects two [N,3,H,W] tensors of one shape, got {tuple(sr.shape)} and {tuple(hr.shape)}")
        sr = sr.contiguous().float()
        hr = hr.detach().contiguous().float()
        self._ensure_packed(sr.device)
        nb = sr.shape[0]
        f_hr32, fh, fw, _ = self._features(hr, keep=False)
        f_sr32, fh, fw, saved = self._features(sr, keep=save)
        c = f_sr32.shape[1]
        f_sr = f_sr32.view(nb, fh, fw, c).permute(0, 3, 1, 2).contiguous()
        f_hr = f_hr32.view(nb, fh, fw, c).permute(0, 3, 1, 2).contiguous()
        return f_sr, f_hr, ((saved, nb, sr.shape[2], sr.shape[3], fh, fw) if save else None)

    def backward(self, state, dfeat):
        saved, nb, h, w, fh, fw = state
        dev = dfeat.device
        e32 = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)  # noqa: E731
        dfeat = dfeat.contiguous().float()
        ops.amax_scale(dfeat, self.scale_ws, target=16.0)
        scale = self.scale_ws[1:2]
        g32 = (dfeat * scale).permute(0, 2, 3, 1).contiguous().view(nb * fh * fw, -1)     # NHWC, carries the range scale
        g = HL(g32.shape[0], g32.shape[1], dev)
        ops.split16(g32, g.hi, g.lo)                    # gradient w.r.t. the last conv's output (no ReLU after it)
        Zd = None
        for idx in range(len(saved) - 1, -1, -1):
            op = saved[idx]
            if op[0] == "pool":
                continue
            _, li, cin, cout, ch, cw, inp, _out = op
            if li == 0:
                Zd = self._conv3(g, "c0_d", nb, ch, cw, cout, 32, e32(nb * ch * cw, 32), taps=[(0, 0)])
                break
            gin32 = self._conv3(g, f"c{li}_d", nb, ch, cw, cout, cin, e32(nb * ch * cw, cin))     # d/d(input of conv li)
            prev = saved[idx - 1]
            if prev[0] == "conv":
                # the input is relu(previous conv): mask with the saved post-ReLU operand pair
                g = HL(nb * ch * cw, cin, dev)
                ops.split16(gin32, g.hi, g.lo, mask_hi=inp.hi, mask_lo=inp.lo, mask_mode=1)
            else:
                # the input is max-pool(relu(previous conv)): route through the pool (first maximum, relu' fused)
                _, pc, ph_, pw_, act32 = prev
                gpre = e32(nb * ph_ * pw_, pc)
                ops.maxpool2_f32_bwd(act32, gin32, nb, ph_, pw_, pc, gpre, relu_mask=True)
                g = HL(nb * ph_ * pw_, pc, dev)
                ops.split16(gpre, g.hi, g.lo)
        sm_wt = self.vgg.sub_mean.weight.detach().reshape(3, 3).t().contiguous()
        dsr = torch.empty(nb, 3, h, w, device=dev, dtype=torch.float32)
        ops.col2im3(Zd, 32, nb, h, w, dsr, affine_a=sm_wt, mul=1.0, div_dev=scale, sgn=-1)
        return dsr
