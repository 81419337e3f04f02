"""Split-precision schedule of the x4 Generator (reference: model/pesr.py:3-38): fp32-grade results on the 16-bit
tensor cores.

Every convolution operand is kept as a pair of fp16 tensors, hi = round16(v) and lo = round16(v - hi) (22 significant
bits), weights included, and every convolution runs as THREE passes of the same implicit-GEMM kernel that the
16-bit schedule uses,

    acc  = hi(x) * hi(w) [+ bias]        acc += lo(x) * hi(w)        acc += hi(x) * lo(w)

accumulating in fp32 (TMEM accumulators, then `res32` / `out32` of the epilogue; the lo*lo term is below 2^-22).
Activations between layers are fp32 in HBM and are split again by `pesr_split16` (which also applies ReLU, the ReLU'
mask of backward and the gradient range scale).  Weight gradients: the split-K kernel runs three times into three
regions of its partials workspace and the ordinary reduction sums all of them.

This is the mode BASELINE.md section 4 / SURVEY.md section 7 (iv) call for to show that the gradient tolerance of the
north_star is a property of 16-bit storage and not of the kernels: with it the Generator's gradients agree with the
FREE-RUNNING fp64 oracle as well as the reference's own fp32 evaluation does.  It costs ~3.5x the 16-bit schedule and is
not the headline path: `Generator(opt, split_precision=True)`.
"""
import torch

from . import ops
from .engine_g import FlatGrads, _shuffle_perm, _unshuffle_perm
from .engine_split import HL as _HL
from .engine_split import SplitOps
from .engine_split import SplitWeight as _SplitWeight
from .engine_split import _shuffle2, _unshuffle2
from .ops import ACT_RELU


import os

# The incoming gradient is scaled (by a power of two) so that its largest element sits at _GRAD_TARGET.  The hi + lo pair
# carries 22 bits only for elements above ~1e-4 of the fp16 normal range: with the 16-bit schedule's target of 16 a gradient
# image with a wide dynamic range (d(GAN loss)/d(sr) through a Discriminator right after its first Adam step: 7.8e-3
# relative error of the Generator gradients) loses its small elements.  2^12 leaves 16x headroom for growth through the
# Generator's backward (measured growth: < 2x) and 8 more bits below.
_GRAD_TARGET = float(os.environ.get("PESR_SPLIT_GRAD_TARGET", "4096"))


class SplitGeneratorEngine(SplitOps):
    def __init__(self, gen):
        self.gen = gen
        self.dtype = torch.float16
        self.packed = None
        self.device = None
        self.param_list = None
        self.grad_hook = None
        self.grad_hook_finish = None
        self.defer_finish = False
        self.trace_hook = None
        self.last_flat = None

    def invalidate_packs(self):
        if self.packed is not None:
            for sw in self.packed.values():
                sw.key = None

    # ------------------------------------------------------------------ parameters
    def _ensure_packed(self, device):
        g = self.gen
        sentinel = (g.embed.weight.data_ptr(), g.add_mean.bias.data_ptr())
        if self.packed is not None and self.device == device and self.sentinel == sentinel:
            return
        self.sentinel, self.device = sentinel, device
        trunk = [(blk.body[0], blk.body[2]) for blk in list(g.body)[:-1]]
        tail = g.body[-1]
        pk = {"embed_f": _SplitWeight(g.embed.weight, 4, pad_to=64), "embed_d": _SplitWeight(g.embed.weight, 6, pad_to=32)}
        for i, (c1, c2) in enumerate(trunk):
            for nm, cv in (("c1", c1), ("c2", c2)):
                pk[f"b{i}{nm}_f"] = _SplitWeight(cv.weight, 0)
                pk[f"b{i}{nm}_d"] = _SplitWeight(cv.weight, 1)
        pk["tail_f"], pk["tail_d"] = _SplitWeight(tail.weight, 0), _SplitWeight(tail.weight, 1)
        for nm, cv in (("up0", g.upsample[0]), ("up2", g.upsample[2])):
            pk[nm + "_f"], pk[nm + "_d"] = _SplitWeight(cv.weight, 2), _SplitWeight(cv.weight, 3)
        pk["up4_f"] = _SplitWeight(g.upsample[4].weight, 5, pad_to=32)
        pk["up4_d"] = _SplitWeight(g.upsample[4].weight, 7, pad_to=64)
        self.packed, self.trunk, self.tail = pk, trunk, tail
        self.flat_grads = FlatGrads(self.param_list)
        self.offsets, self.flat_numel = self.flat_grads.offsets, self.flat_grads.numel
        self.scale_ws = torch.zeros(4, device=device, dtype=torch.float32)

    # ------------------------------------------------------------------ forward
    def forward(self, lr, train, out_u8=False):
        g = self.gen
        if out_u8 or lr.dtype == torch.uint8:
            raise NotImplementedError("split-precision Generator: fp32 NCHW in / out only")
        if lr.dim() != 4 or lr.shape[1] != 3:
            raise ValueError(f"Generator expects [N,3,H,W], got {tuple(lr.shape)}")
        lr = lr.contiguous().float()
        nb, _, h, w = lr.shape
        dev = lr.device
        self._ensure_packed(dev)
        Cn, depth, rs = g.n_feats, g.n_resblock, float(g.res_scale)
        P = nb * h * w
        e32 = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)  # noqa: E731
        st = type("SplitState", (), {})()
        st.nb, st.h, st.w, st.lr = nb, h, w, lr
        sm_w = g.sub_mean.weight.detach().reshape(3, 3)
        am_w = g.add_mean.weight.detach().reshape(3, 3)
        col = _HL(P, 64, dev)
        ops.im2col3(lr, col.hi, affine_a=sm_w, affine_b=g.sub_mean.bias.detach())
        ops.im2col3(lr, col.lo, affine_a=sm_w, affine_b=g.sub_mean.bias.detach(), low_part=True)
        F0 = self._conv3(col, "embed_f", nb, h, w, 64, Cn, e32(P, Cn), bias=g.embed.bias.detach(), taps=[(0, 0)])
        S = F0.clone()
        X = [_HL(P, Cn, dev)]
        ops.split16(S, X[0].hi, X[0].lo)
        T = []
        A32 = e32(P, Cn)
        for i, (c1, c2) in enumerate(self.trunk):
            self._conv3(X[i], f"b{i}c1_f", nb, h, w, Cn, Cn, A32, bias=c1.bias.detach())
            t = _HL(P, Cn, dev)
            ops.split16(A32, t.hi, t.lo, act=ACT_RELU)
            self._conv3(t, f"b{i}c2_f", nb, h, w, Cn, Cn, S, bias=c2.bias.detach(), alpha=rs, res32=S)   # S += rs*(conv2(t)+b)
            x = _HL(P, Cn, dev)
            ops.split16(S, x.hi, x.lo)
            T.append(t)
            X.append(x)
            if not train and i > 0:
                X[i] = T[i - 1] = None         # inference keeps nothing
        R32 = self._conv3(X[depth], "tail_f", nb, h, w, Cn, Cn, e32(P, Cn), bias=self.tail.bias.detach(), res32=F0)
        U0 = _HL(P, Cn, dev)
        ops.split16(R32, U0.hi, U0.lo)
        up0, up2, up4 = g.upsample[0], g.upsample[2], g.upsample[4]
        Q0 = self._conv3(U0, "up0_f", nb, h, w, Cn, 4 * Cn, e32(P, 4 * Cn), bias=_shuffle_perm(up0.bias.detach(), Cn).contiguous())
        U1 = _HL(4 * P, Cn, dev)
        ops.split16(_shuffle2(Q0, nb, h, w, Cn), U1.hi, U1.lo)
        Q1 = self._conv3(U1, "up2_f", nb, 2 * h, 2 * w, Cn, 4 * Cn, e32(4 * P, 4 * Cn),
                         bias=_shuffle_perm(up2.bias.detach(), Cn).contiguous())
        U2 = _HL(16 * P, Cn, dev)
        ops.split16(_shuffle2(Q1, nb, 2 * h, 2 * w, Cn), U2.hi, U2.lo)
        del Q0, Q1
        Z = self._conv3(U2, "up4_f", nb, 4 * h, 4 * w, Cn, 32, e32(16 * P, 32), taps=[(0, 0)])
        sr = torch.empty(nb, 3, 4 * h, 4 * w, device=dev, dtype=torch.float32)
        ypre = e32(nb, 3, 4 * h, 4 * w) if train else None
        ops.col2im3(Z, 32, nb, 4 * h, 4 * w, sr, bias=up4.bias.detach(), affine_a=am_w, affine_b=g.add_mean.bias.detach(),
                    sgn=1, pre=ypre)
        if not train:
            return sr, None
        st.col, st.X, st.T, st.U0, st.U1, st.U2, st.ypre = col, X, T, U0, U1, U2, ypre
        return sr, st

    # ------------------------------------------------------------------ backward
    def backward(self, st, dsr, need_input_grad=False):
        g = self.gen
        Cn, depth, rs = g.n_feats, g.n_resblock, float(g.res_scale)
        nb, h, w, lr = st.nb, st.h, st.w, st.lr
        P = nb * h * w
        dev = dsr.device
        dsr = dsr.contiguous().float()
        up0, up2, up4 = g.upsample[0], g.upsample[2], g.upsample[4]
        e32 = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)  # noqa: E731
        flat = self.flat_grads.get(dev)
        self.last_flat = flat
        off = self.offsets
        grads = {p: flat[off[p]:off[p] + p.numel()].view(p.shape) for p in self.param_list}
        self.wg = e32(3 * max(9 * 4 * Cn * Cn * 2, 9 * Cn * Cn * 8, 148 * max(Cn, 128) * 64))
        ws = self.scale_ws
        scale = ws[1:2]
        am_wt = g.add_mean.weight.detach().reshape(3, 3).t().contiguous()
        sums = torch.zeros(24, device=dev, dtype=torch.float32)
        ops.moments3(dsr, st.ypre, sums[:12])
        grads[g.add_mean.weight].view(-1).copy_(sums[0:9])
        grads[g.add_mean.bias].copy_(sums[9:12])
        torch.mv(am_wt, sums[9:12], out=grads[up4.bias])
        ops.amax_scale(dsr, ws, target=_GRAD_TARGET)
        dcol = _HL(16 * P, 64, dev)
        ops.im2col3(dsr, dcol.hi, affine_a=am_wt, mul_dev=scale, sgn=-1)
        ops.im2col3(dsr, dcol.lo, affine_a=am_wt, mul_dev=scale, sgn=-1, low_part=True)
        # upsample.4 (Cout = 3): dgrad as a GEMM over the flipped im2col of dy; wgrad in the col2im order
        G2 = self._conv3(dcol, "up4_d", nb, 4 * h, 4 * w, 64, Cn, e32(16 * P, Cn), taps=[(0, 0)])
        self._wgrad3(st.U2, Cn, dcol, 64, nb, 4 * h, 4 * w, grads[up4.weight], ops.WMAP_COL_OUT, 3, Cn, scale, taps=[(0, 0)])
        dZ2 = _HL(4 * P, 4 * Cn, dev)
        ops.split16(_unshuffle2(G2, nb, 2 * h, 2 * w, Cn), dZ2.hi, dZ2.lo)
        del G2
        tmp = e32(4 * Cn)
        # upsample.2
        G1 = self._conv3(dZ2, "up2_d", nb, 2 * h, 2 * w, 4 * Cn, Cn, e32(4 * P, Cn))
        self._bias_grad(dZ2, 4 * P, 4 * Cn, tmp, scale)
        grads[up2.bias].copy_(_unshuffle_perm(tmp, Cn))
        self._wgrad3(dZ2, 4 * Cn, st.U1, Cn, nb, 2 * h, 2 * w, grads[up2.weight], ops.WMAP_OIHW_PS, 4 * Cn, Cn, scale)
        dZ1 = _HL(P, 4 * Cn, dev)
        ops.split16(_unshuffle2(G1, nb, h, w, Cn), dZ1.hi, dZ1.lo)
        del G1, dZ2
        # upsample.0: its dgrad is dR, the gradient of (tail(X_depth) + F0)
        dR32 = self._conv3(dZ1, "up0_d", nb, h, w, 4 * Cn, Cn, e32(P, Cn))
        self._bias_grad(dZ1, P, 4 * Cn, tmp, scale)
        grads[up0.bias].copy_(_unshuffle_perm(tmp, Cn))
        self._wgrad3(dZ1, 4 * Cn, st.U0, Cn, nb, h, w, grads[up0.weight], ops.WMAP_OIHW_PS, 4 * Cn, Cn, scale)
        dR = _HL(P, Cn, dev)
        ops.split16(dR32, dR.hi, dR.lo)
        # tail conv
        self._bias_grad(dR, P, Cn, grads[self.tail.bias], scale)
        self._wgrad3(dR, Cn, st.X[depth], Cn, nb, h, w, grads[self.tail.weight], ops.WMAP_OIHW, Cn, Cn, scale)
        gS32 = self._conv3(dR, "tail_d", nb, h, w, Cn, Cn, e32(P, Cn))
        gS, dT, dT32 = _HL(P, Cn, dev), _HL(P, Cn, dev), e32(P, Cn)
        ops.split16(gS32, gS.hi, gS.lo)
        for i in range(depth - 1, -1, -1):
            c1, c2 = self.trunk[i]
            self._bias_grad(gS, P, Cn, grads[c2.bias], scale, mul=rs)
            self._wgrad3(gS, Cn, st.T[i], Cn, nb, h, w, grads[c2.weight], ops.WMAP_OIHW, Cn, Cn, scale, mul=rs)
            self._conv3(gS, f"b{i}c2_d", nb, h, w, Cn, Cn, dT32, alpha=rs)
            ops.split16(dT32, dT.hi, dT.lo, mask_hi=st.T[i].hi, mask_lo=st.T[i].lo, mask_mode=1)
            self._bias_grad(dT, P, Cn, grads[c1.bias], scale)
            self._wgrad3(dT, Cn, st.X[i], Cn, nb, h, w, grads[c1.weight], ops.WMAP_OIHW, Cn, Cn, scale)
            self._conv3(dT, f"b{i}c1_d", nb, h, w, Cn, Cn, gS32, res32=gS32)           # gS_i = gS_{i+1} + conv1^T(dT)
            if i == 0:
                gS32.add_(dR32)                                                        # the global skip: F0 also feeds the tail sum
            ops.split16(gS32, gS.hi, gS.lo)
        # embed (Cin = 3)
        self._bias_grad(gS, P, Cn, grads[g.embed.bias], scale)
        self._wgrad3(gS, Cn, st.col, 64, nb, h, w, grads[g.embed.weight], ops.WMAP_COL_IN, Cn, 3, scale, taps=[(0, 0)])
        Zd = self._conv3(gS, "embed_d", nb, h, w, Cn, 32, e32(P, 32), taps=[(0, 0)])
        dx_sm = e32(nb, 3, h, w)
        ops.col2im3(Zd, 32, nb, h, w, dx_sm, mul=1.0, div_dev=scale, sgn=-1)
        ops.moments3(dx_sm, lr, sums[12:])
        grads[g.sub_mean.weight].view(-1).copy_(sums[12:21])
        grads[g.sub_mean.bias].copy_(sums[21:24])
        if self.grad_hook is not None:
            self.grad_hook(0, self.flat_numel, flat)
            if self.grad_hook_finish is not None and not self.defer_finish:
                self.grad_hook_finish()
        self.wg = None
        dlr = None
        if need_input_grad:
            sm_w = g.sub_mean.weight.detach().reshape(3, 3)
            dlr = torch.einsum("oi,nohw->nihw", sm_w, dx_sm)
        return grads, dlr
