"""Kernel schedule of the x4 Generator (reference: model/pesr.py:3-38, model/basic.py:33-60).

Forward (all activations NHWC 16-bit, fp32 residual stream kept beside them):

    lr --im2col3(+sub_mean)--> col --embed GEMM--> F0 (fp32) , X[0]
    block i:  T[i] = relu(conv1(X[i]) + b1)                      (bias+ReLU fused)
              S    = S + res_scale*(conv2(T[i]) + b2) ; X[i+1] = cvt(S)   (bias, scale, fp32 skip fused)
    U0 = tail(X[depth]) + b + F0                                 (global skip fused)
    U1 = shuffle(conv(U0)) ; U2 = shuffle(conv(U1))              (PixelShuffle fused into the store)
    Z  = U2 x W4 (1x1 GEMM, 27 columns) --col2im3(+bias, +add_mean)--> sr (NCHW fp32)

Backward mirrors it with the same implicit-GEMM kernel on flipped weights (dgrad) and the split-K
MN-major kernel (wgrad); the fp32 gradients land in the reference's OIHW layouts.
"""
import torch

from . import ops
from .ops import ACT_NONE, ACT_RELU, OUT_SHUFFLE2, OUT_UNSHUFFLE2


class PackedWeight:
    """16-bit GEMM operand of one fp32 parameter, re-packed only when the parameter changed."""

    def __init__(self, param, mode, dtype, pad_to=0):
        co, ci, k = param.shape[0], param.shape[1], param.shape[2]
        self.param, self.mode, self.pad_to = param, mode, pad_to
        self.buf = torch.empty(ops.packed_shape(co, ci, k, mode, pad_to), device=param.device, dtype=dtype)
        self.key = None

    def get(self):
        p = self.param
        key = (p.data_ptr(), p._version)
        if key != self.key:
            ops.pack_weights(p.detach(), self.mode, self.buf, self.pad_to)
            self.key = key
        return self.buf


def _unshuffle_perm(vec, c):
    """packed PixelShuffle order (i, j, c) -> reference order c*4 + i*2 + j."""
    return vec.view(4, c).t().reshape(-1)


def _shuffle_perm(vec, c):
    """reference order c*4 + ij -> packed (ij, c)."""
    return vec.view(c, 4).t().reshape(-1)


class _Plan:
    """Buffers + prebuilt launch descriptors for one (nb, h, w, training) configuration."""
    pass


class GeneratorEngine:
    def __init__(self, gen, dtype=torch.float16):
        self.gen = gen
        self.dtype = dtype
        self.dt = ops.dt_code(dtype)
        self.plans = {}
        self.packed = None
        self.device = None
        self.wg_ws = None

    # ------------------------------------------------------------------ parameters
    def _convs(self):
        g = self.gen
        trunk = []
        for blk in g.body[:-1]:
            trunk.append((blk.body[0], blk.body[2]))
        return trunk, g.body[-1]

    def _ensure_packed(self, device):
        if self.packed is not None and self.device == device:
            return
        g, dt = self.gen, self.dtype
        self.device = device
        trunk, tail = self._convs()
        pk = {}
        pk["embed_f"] = PackedWeight(g.embed.weight, 4, dt, pad_to=64)
        pk["embed_d"] = PackedWeight(g.embed.weight, 6, dt, pad_to=32)
        for i, (c1, c2) in enumerate(trunk):
            pk[f"b{i}c1_f"] = PackedWeight(c1.weight, 0, dt)
            pk[f"b{i}c2_f"] = PackedWeight(c2.weight, 0, dt)
            pk[f"b{i}c1_d"] = PackedWeight(c1.weight, 1, dt)
            pk[f"b{i}c2_d"] = PackedWeight(c2.weight, 1, dt)
        pk["tail_f"] = PackedWeight(tail.weight, 0, dt)
        pk["tail_d"] = PackedWeight(tail.weight, 1, dt)
        pk["up0_f"] = PackedWeight(g.upsample[0].weight, 2, dt)
        pk["up0_d"] = PackedWeight(g.upsample[0].weight, 3, dt)
        pk["up2_f"] = PackedWeight(g.upsample[2].weight, 2, dt)
        pk["up2_d"] = PackedWeight(g.upsample[2].weight, 3, dt)
        pk["up4_f"] = PackedWeight(g.upsample[4].weight, 5, dt, pad_to=32)
        pk["up4_d"] = PackedWeight(g.upsample[4].weight, 7, dt, pad_to=64)
        self.packed = pk
        self.scale_ws = torch.zeros(4, device=device, dtype=torch.float32)

    # ------------------------------------------------------------------ plans
    def _plan(self, nb, h, w, train):
        key = (nb, h, w, train)
        pl = self.plans.get(key)
        if pl is not None:
            return pl
        g = self.gen
        C, depth = g.n_feats, g.n_resblock
        dev, dt = self.device, self.dtype
        P = nb * h * w
        pl = _Plan()
        pl.nb, pl.h, pl.w, pl.P, pl.train = nb, h, w, P, train
        e16 = lambda *s: torch.empty(*s, device=dev, dtype=dt)
        e32 = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        pl.col_in = e16(P, 64)
        pl.F0 = e32(P, C)
        pl.S = e32(P, C)
        nx = depth + 1 if train else 2
        pl.X = [e16(P, C) for _ in range(nx)]
        pl.T = [e16(P, C) for _ in range(depth if train else 1)]
        pl.U0 = e16(P, C)
        pl.U1 = e16(4 * P, C)
        pl.U2 = e16(16 * P, C)
        pl.Z = e32(16 * P, 32)
        pl.ypre = e32(nb, 3, 4 * h, 4 * w) if train else None
        pl.generation = 0
        if train:
            pl.dcol = e16(16 * P, 64)
            pl.dZ2 = e16(4 * P, 4 * C)
            pl.dZ1 = e16(P, 4 * C)
            pl.dR16 = e16(P, C)
            pl.gS32 = e32(P, C)
            pl.gS16 = e16(P, C)
            pl.dT16 = e16(P, C)
            pl.dF0 = e16(P, C)
            pl.Zd = e32(P, 32)
            pl.dx_sm = e32(nb, 3, h, w)
            pl.sums = torch.zeros(24, device=dev, dtype=torch.float32)
            wg_elems = max(9 * 4 * C * C * 2, 9 * C * C * 8, 148 * C * 64)
            pl.wg = e32(wg_elems)
        self.plans[key] = pl
        return pl

    def _x(self, pl, i):
        return pl.X[i] if pl.train else pl.X[i % 2]

    def _t(self, pl, i):
        return pl.T[i] if pl.train else pl.T[0]

    # ------------------------------------------------------------------ forward
    def forward(self, lr, train):
        g = self.gen
        assert lr.dim() == 4 and lr.shape[1] == 3, "Generator expects [N,3,H,W]"
        lr = lr.contiguous().float()
        nb, _, h, w = lr.shape
        self._ensure_packed(lr.device)
        pl = self._plan(nb, h, w, train)
        pl.generation += 1
        pk, dt = self.packed, self.dt
        C, depth, rs = g.n_feats, g.n_resblock, float(g.res_scale)
        trunk, tail = self._convs()
        P = pl.P
        sm_w = g.sub_mean.weight.detach().reshape(3, 3).contiguous()
        sm_b = g.sub_mean.bias.detach()
        am_w = g.add_mean.weight.detach().reshape(3, 3).contiguous()
        am_b = g.add_mean.bias.detach()

        ops.im2col3(lr, pl.col_in, affine_a=sm_w, affine_b=sm_b)
        # embed: 1x1 GEMM over the im2col matrix
        ops.conv_igemm(ops.make_conv_desc(
            dtype=dt, nb=nb, h=h, w=w, cin=64, cout=C, taps=[(0, 0)], srcs=[ops.nhwc_src(pl.col_in, nb, h, w, 64)],
            wpacked=pk["embed_f"].get(), bias=g.embed.bias.detach(), out32=pl.F0, ld_out32=C,
            out16=self._x(pl, 0), ld_out16=C))
        for i, (c1, c2) in enumerate(trunk):
            xin, t, xout = self._x(pl, i), self._t(pl, i), self._x(pl, i + 1)
            ops.conv_igemm(ops.make_conv_desc(
                dtype=dt, nb=nb, h=h, w=w, cin=C, cout=C, srcs=[ops.nhwc_src(xin, nb, h, w, C)],
                wpacked=pk[f"b{i}c1_f"].get(), bias=c1.bias.detach(), act=ACT_RELU, out16=t, ld_out16=C))
            ops.conv_igemm(ops.make_conv_desc(
                dtype=dt, nb=nb, h=h, w=w, cin=C, cout=C, srcs=[ops.nhwc_src(t, nb, h, w, C)],
                wpacked=pk[f"b{i}c2_f"].get(), bias=c2.bias.detach(), alpha=rs,
                res32=pl.F0 if i == 0 else pl.S, ld_res32=C, out32=pl.S, ld_out32=C, out16=xout, ld_out16=C))
        ops.conv_igemm(ops.make_conv_desc(
            dtype=dt, nb=nb, h=h, w=w, cin=C, cout=C, srcs=[ops.nhwc_src(self._x(pl, depth), nb, h, w, C)],
            wpacked=pk["tail_f"].get(), bias=tail.bias.detach(), res32=pl.F0, ld_res32=C, out16=pl.U0, ld_out16=C))
        up0, up2, up4 = g.upsample[0], g.upsample[2], g.upsample[4]
        ops.conv_igemm(ops.make_conv_desc(
            dtype=dt, nb=nb, h=h, w=w, cin=C, cout=4 * C, block_n=min(C, 256),
            srcs=[ops.nhwc_src(pl.U0, nb, h, w, C)], wpacked=pk["up0_f"].get(),
            bias=_shuffle_perm(up0.bias.detach(), C).contiguous(), out16=pl.U1, ld_out16=C, out_mode=OUT_SHUFFLE2,
            ps_c=C))
        ops.conv_igemm(ops.make_conv_desc(
            dtype=dt, nb=nb, h=2 * h, w=2 * w, cin=C, cout=4 * C, block_n=min(C, 256),
            srcs=[ops.nhwc_src(pl.U1, nb, 2 * h, 2 * w, C)], wpacked=pk["up2_f"].get(),
            bias=_shuffle_perm(up2.bias.detach(), C).contiguous(), out16=pl.U2, ld_out16=C, out_mode=OUT_SHUFFLE2,
            ps_c=C))
        ops.conv_igemm(ops.make_conv_desc(
            dtype=dt, nb=nb, h=4 * h, w=4 * w, cin=C, cout=32, taps=[(0, 0)],
            srcs=[ops.nhwc_src(pl.U2, nb, 4 * h, 4 * w, C)], wpacked=pk["up4_f"].get(), out32=pl.Z, ld_out32=32))
        sr = torch.empty(nb, 3, 4 * h, 4 * w, device=lr.device, dtype=torch.float32)
        ops.col2im3(pl.Z, 32, nb, 4 * h, 4 * w, sr, bias=up4.bias.detach(), affine_a=am_w, affine_b=am_b, sgn=1,
                    pre=pl.ypre)
        return sr, (pl, pl.generation, lr)

    # ------------------------------------------------------------------ backward
    def backward(self, state, dsr, need_input_grad=False):
        """Returns (grads: dict param -> fp32 gradient in the reference layout, dlr or None)."""
        pl, generation, lr = state
        if generation != pl.generation:
            raise RuntimeError("pesr_b200.Generator: backward through a forward whose activations were overwritten "
                               "by a later forward of the same shape (keep one live graph per shape)")
        g = self.gen
        pk, dt = self.packed, self.dt
        C, depth, rs = g.n_feats, g.n_resblock, float(g.res_scale)
        nb, h, w, P = pl.nb, pl.h, pl.w, pl.P
        trunk, tail = self._convs()
        up0, up2, up4 = g.upsample[0], g.upsample[2], g.upsample[4]
        dsr = dsr.contiguous().float()
        dev = dsr.device
        grads = {}
        ws = self.scale_ws
        scale, wg = ws[1:2], pl.wg

        def wgrad(a, a_c, m_total, b, b_c, n_total, gh, gw, param, map_mode, co, ci, taps=ops.TAPS_3X3, mul=1.0):
            d = ops.make_wgrad_desc(dtype=dt, nb=nb, h=gh, w=gw, a=a, a_c=a_c, m_total=m_total,
                                    b_srcs=[ops.nhwc_src(b, nb, gh, gw, b_c)], n_total=n_total, taps=taps,
                                    partials=wg)
            splits = ops.conv_wgrad(d)
            gr = torch.empty_like(param)
            ops.wgrad_reduce(wg, splits, len(taps), m_total, n_total, map_mode, co, ci, gr, scale=mul, div_dev=scale)
            grads[param] = gr

        def bgrad(x16, npix, c, ldc, mul=1.0):
            out = torch.empty(c, device=dev, dtype=torch.float32)
            ops.colsum16(x16, npix, c, ldc, out, mul=mul, div_dev=scale)
            return out

        # --- add_mean (1x1, trainable in the reference) and the gradient scale
        am_w = g.add_mean.weight.detach().reshape(3, 3)
        ops.moments3(dsr, pl.ypre, pl.sums[:12])
        ops.amax_scale(dsr, ws, target=16.0)
        am_wt = am_w.t().contiguous()
        ops.im2col3(dsr, pl.dcol, affine_a=am_wt, mul_dev=scale, sgn=-1)
        # --- upsample.4 (Cout = 3): dgrad is a 1x1 GEMM over the flipped im2col of dy, stored un-shuffled
        ops.conv_igemm(ops.make_conv_desc(
            dtype=dt, nb=nb, h=4 * h, w=4 * w, cin=64, cout=C, taps=[(0, 0)],
            srcs=[ops.nhwc_src(pl.dcol, nb, 4 * h, 4 * w, 64)], wpacked=pk["up4_d"].get(),
            out16=pl.dZ2, ld_out16=4 * C, out_mode=OUT_UNSHUFFLE2))
        wgrad(pl.U2, C, C, pl.dcol, 64, 64, 4 * h, 4 * w, up4.weight, ops.WMAP_COL_OUT, 3, C, taps=[(0, 0)])
        # --- upsample.2
        ops.conv_igemm(ops.make_conv_desc(
            dtype=dt, nb=nb, h=2 * h, w=2 * w, cin=4 * C, cout=C, srcs=[ops.nhwc_src(pl.dZ2, nb, 2 * h, 2 * w, 4 * C)],
            wpacked=pk["up2_d"].get(), out16=pl.dZ1, ld_out16=4 * C, out_mode=OUT_UNSHUFFLE2))
        wgrad(pl.dZ2, 4 * C, 4 * C, pl.U1, C, C, 2 * h, 2 * w, up2.weight, ops.WMAP_OIHW_PS, 4 * C, C)
        grads[up2.bias] = _unshuffle_perm(bgrad(pl.dZ2, 4 * P, 4 * C, 4 * C), C).contiguous()
        # --- upsample.0 ; its dgrad is dR, the gradient of (tail(X_depth) + F0)
        ops.conv_igemm(ops.make_conv_desc(
            dtype=dt, nb=nb, h=h, w=w, cin=4 * C, cout=C, srcs=[ops.nhwc_src(pl.dZ1, nb, h, w, 4 * C)],
            wpacked=pk["up0_d"].get(), out16=pl.dR16, ld_out16=C))
        wgrad(pl.dZ1, 4 * C, 4 * C, pl.U0, C, C, h, w, up0.weight, ops.WMAP_OIHW_PS, 4 * C, C)
        grads[up0.bias] = _unshuffle_perm(bgrad(pl.dZ1, P, 4 * C, 4 * C), C).contiguous()
        # --- tail conv
        wgrad(pl.dR16, C, C, pl.X[depth], C, C, h, w, tail.weight, ops.WMAP_OIHW, C, C)
        grads[tail.bias] = bgrad(pl.dR16, P, C, C)
        ops.conv_igemm(ops.make_conv_desc(
            dtype=dt, nb=nb, h=h, w=w, cin=C, cout=C, srcs=[ops.nhwc_src(pl.dR16, nb, h, w, C)],
            wpacked=pk["tail_d"].get(), out32=pl.gS32, ld_out32=C, out16=pl.gS16, ld_out16=C))
        # --- residual blocks, last to first
        for i in range(depth - 1, -1, -1):
            c1, c2 = trunk[i]
            wgrad(pl.gS16, C, C, pl.T[i], C, C, h, w, c2.weight, ops.WMAP_OIHW, C, C, mul=rs)
            grads[c2.bias] = bgrad(pl.gS16, P, C, C, mul=rs)
            ops.conv_igemm(ops.make_conv_desc(
                dtype=dt, nb=nb, h=h, w=w, cin=C, cout=C, srcs=[ops.nhwc_src(pl.gS16, nb, h, w, C)],
                wpacked=pk[f"b{i}c2_d"].get(), alpha=rs, mask16=pl.T[i], ld_mask16=C, mask_mode=1,
                out16=pl.dT16, ld_out16=C))
            wgrad(pl.dT16, C, C, pl.X[i], C, C, h, w, c1.weight, ops.WMAP_OIHW, C, C)
            grads[c1.bias] = bgrad(pl.dT16, P, C, C)
            last = i == 0
            ops.conv_igemm(ops.make_conv_desc(
                dtype=dt, nb=nb, h=h, w=w, cin=C, cout=C, srcs=[ops.nhwc_src(pl.dT16, nb, h, w, C)],
                wpacked=pk[f"b{i}c1_d"].get(), res32=pl.gS32, ld_res32=C,
                res16=pl.dR16 if last else None, ld_res16=C,
                out32=None if last else pl.gS32, ld_out32=C, out16=pl.dF0 if last else pl.gS16, ld_out16=C))
        if depth == 0:
            raise NotImplementedError("Generator with depth 0")
        # --- embed (Cin = 3): wgrad against the saved im2col matrix; dgrad as a col2im GEMM
        wgrad(pl.dF0, C, C, pl.col_in, 64, 64, h, w, g.embed.weight, ops.WMAP_COL_IN, C, 3, taps=[(0, 0)])
        grads[g.embed.bias] = bgrad(pl.dF0, P, C, C)
        ops.conv_igemm(ops.make_conv_desc(
            dtype=dt, nb=nb, h=h, w=w, cin=C, cout=32, taps=[(0, 0)], srcs=[ops.nhwc_src(pl.dF0, nb, h, w, C)],
            wpacked=pk["embed_d"].get(), out32=pl.Zd, ld_out32=32))
        ops.col2im3(pl.Zd, 32, nb, h, w, pl.dx_sm, mul=1.0, div_dev=scale, sgn=-1)
        ops.moments3(pl.dx_sm, lr, pl.sums[12:])
        # --- tiny host-side (torch) assembly of the 1x1 MeanShift gradients
        s = pl.sums
        grads[g.add_mean.weight] = s[0:9].reshape(3, 3, 1, 1).clone()
        grads[g.add_mean.bias] = s[9:12].clone()
        grads[up4.bias] = am_wt @ s[9:12]
        grads[g.sub_mean.weight] = s[12:21].reshape(3, 3, 1, 1).clone()
        grads[g.sub_mean.bias] = s[21:24].clone()
        dlr = None
        if need_input_grad:
            sm_w = g.sub_mean.weight.detach().reshape(3, 3)
            dlr = torch.einsum("oi,nohw->nihw", sm_w, pl.dx_sm)
        return grads, dlr
