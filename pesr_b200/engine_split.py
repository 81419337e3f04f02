"""Shared pieces of the split-precision schedules (engine_g_split.py, engine_d_split.py, engine_v_split.py): fp16 hi + lo
operand pairs and the three-pass convolution / weight-gradient primitives built on the ordinary 16-bit kernels.

    acc  = hi(x) * hi(w) [+ bias]        acc += lo(x) * hi(w)        acc += hi(x) * lo(w)          (fp32 accumulation)

hi = round16(v), lo = round16(v - hi) carry 22 significant bits; the dropped lo*lo term is below 2^-22.
"""
import ctypes as C

import torch

from . import ops
from ._lib import check, lib


class SplitWeight:
    """hi / lo 16-bit GEMM operands of one fp32 parameter in layout `mode`, re-packed when the parameter changes."""

    def __init__(self, param, mode, pad_to=0, scale=1.0):
        # scale (a power of two): the operand pair holds param * scale.  Small weights (|w| ~ 0.03) have lo halves in the
        # fp16 SUBNORMAL range (19 significant bits in all); scaled by 256 both halves are normal numbers (22 bits) and the
        # caller folds 1 / scale into the convolution's alpha.
        self.scale = float(scale)
        co, ci, k = param.shape[0], param.shape[1], param.shape[2]
        shape = ops.packed_shape(co, ci, k, mode, pad_to)
        self.param, self.mode, self.pad_to = param, mode, pad_to
        self.hi = torch.empty(shape, device=param.device, dtype=torch.float16)
        self.lo = torch.empty(shape, device=param.device, dtype=torch.float16)
        self.key = None

    def get(self):
        p = self.param
        key = (p.data_ptr(), p._version)
        if key != self.key:
            w = p.detach()
            if self.scale != 1.0:
                w = w * self.scale
            ops.pack_weights(w, self.mode, self.hi, self.pad_to)
            ops.pack_weights(w - w.half().float(), self.mode, self.lo, self.pad_to)
            self.key = key
        return self.hi, self.lo


class HL:
    """A pair of 16-bit NHWC tensors [P][C] holding the high and low parts of an fp32 tensor."""

    def __init__(self, p, c, device):
        self.hi = torch.empty(p, c, device=device, dtype=torch.float16)
        self.lo = torch.empty(p, c, device=device, dtype=torch.float16)


def _shuffle2(t, nb, h, w, c):
    """[nb*h*w][4c] in packed PixelShuffle order (ij, c) -> [nb*2h*2w][c]   (model/basic.py:57,59)."""
    return t.view(nb, h, w, 2, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(nb * 4 * h * w, c).contiguous()


def _unshuffle2(t, nb, h, w, c):
    """inverse: [nb*2h*2w][c] -> [nb*h*w][4c]."""
    return t.view(nb, h, 2, w, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(nb * h * w, 4 * c).contiguous()




class SplitOps:
    """Mixin: needs self.packed (name -> SplitWeight) and self.wg (fp32 workspace, 3 regions) set by the engine."""

    def _conv3(self, x, wname, nb, h, w, cin, cout, out32, bias=None, alpha=1.0, res32=None, taps=ops.TAPS_3X3,
               srcs_fn=None, **desc_kw):
        """out32 = alpha * (conv(x_hi + x_lo, w_hi + w_lo) + bias) + res32, dropping the lo*lo term (three launches).
        srcs_fn(tensor) -> source views (default: one dense NHWC view of nb x h x w x cin); desc_kw: further
        make_conv_desc arguments (parity-plane taps, strided output grids, classes ...)."""
        wh, wl = self.packed[wname].get()
        stream = torch.cuda.current_stream().cuda_stream
        ld = out32.shape[-1]
        if srcs_fn is None:
            srcs_fn = lambda t: [ops.nhwc_src(t, nb, h, w, cin)]      # noqa: E731
        first = True
        for xs, ws in ((x.hi, wh), (x.lo, wh), (x.hi, wl)):
            d = ops.make_conv_desc(dtype=0, nb=nb, h=h, w=w, cin=cin, cout=cout, taps=taps, srcs=srcs_fn(xs), wpacked=ws,
                                   bias=bias if first else None, alpha=alpha, res32=res32 if first else out32, ld_res32=ld,
                                   out32=out32, ld_out32=ld, **desc_kw)
            check(lib.pesr_conv_igemm(C.byref(d), stream), "pesr_conv_igemm")
            first = False
        return out32

    def _wgrad3(self, a, a_c, b, b_c, nb, h, w, param_grad, map_mode, co, ci, scale, mul=1.0, taps=ops.TAPS_3X3,
                b_srcs_fn=None, tap_src=None):
        """weight gradient of one layer from split operands a (dY, nb x h x w x a_c) and b (X): three split-K launches,
        one reduction.  b_srcs_fn(tensor) -> source views of b (default dense; parity planes for stride-2 layers)."""
        ntaps = len(taps)
        region = None
        splits = 0
        if b_srcs_fn is None:
            b_srcs_fn = lambda t: [ops.nhwc_src(t, nb, h, w, b_c)]      # noqa: E731
        for k, (as_, bs) in enumerate(((a.hi, b.hi), (a.lo, b.hi), (a.hi, b.lo))):
            part = self.wg if region is None else self.wg[k * region:]
            d = ops.make_wgrad_desc(dtype=0, nb=nb, h=h, w=w, a=as_, a_c=a_c, m_total=a_c, b_srcs=b_srcs_fn(bs),
                                    n_total=b_c, taps=taps, tap_src=tap_src, partials=part, splits=splits)
            if region is None:
                d.partials_elems = self.wg.numel() // 3
            s = ops.conv_wgrad(d)
            if region is None:
                splits, region = s, s * ntaps * a_c * b_c
            elif s != splits:
                raise RuntimeError("split wgrad: the three passes chose different split factors")
        ops.wgrad_reduce(self.wg, 3 * splits, ntaps, a_c, b_c, map_mode, co, ci, param_grad, scale=mul, div_dev=scale)

    @staticmethod
    def _bias_grad(a, npix, c, out, scale, mul=1.0):
        out.zero_()
        ops.colsum16(a.hi, npix, c, c, out, mul=mul, div_dev=scale, accumulate=True)
        ops.colsum16(a.lo, npix, c, c, out, mul=mul, div_dev=scale, accumulate=True)

