"""Stand-alone forms of the reference's building blocks on the sm_100a kernels.

In the reference `Conv`, `MeanShift` and `ResBlock` are callable modules (model/basic.py:4-17,33-52), so user code can
write `G.embed(x)`, `vgg.sub_mean(x)` or reuse a `ResBlock`.  Inside Generator / Discriminator / VGG the work is
scheduled by the engines (fused epilogues, saved 16-bit activations); the functions here give the same modules a
working `.forward` outside those schedules: NCHW fp32 in and out, 16-bit tensor-core operands with fp32 accumulation,
autograd through the same dgrad / wgrad kernels.  They favour generality over speed (layout conversions at both
ends, no fusion across calls).
"""
import ctypes as C

import torch

from . import ops
from ._lib import check, lib
from .ops import ACT_NONE

_PACK_CACHE = {}


def _packed(weight, mode, dtype, pad_to=0):
    """16-bit GEMM operand of `weight` in layout `mode` (include/pesr_b200.h pesr_pack_weights), cached per version."""
    key = (weight.data_ptr(), weight._version, mode, dtype, pad_to, tuple(weight.shape))
    buf = _PACK_CACHE.get(key)
    if buf is None:
        if len(_PACK_CACHE) > 256:
            _PACK_CACHE.clear()
        co, ci, k = weight.shape[0], weight.shape[1], weight.shape[2]
        buf = torch.empty(ops.packed_shape(co, ci, k, mode, pad_to), device=weight.device, dtype=dtype)
        ops.pack_weights(weight.detach().contiguous(), mode, buf, pad_to)
        _PACK_CACHE[key] = buf
    return buf


_block_n = ops.default_block_n


def _s2_taps():
    from .engine_d import _s2_taps as f
    return f()


def _check(x, weight, stride):
    if not x.is_cuda:
        raise RuntimeError(f"pesr_b200.functional: input is on {x.device}; the B200 path has no CPU fallback")
    co, ci, kh, kw = weight.shape
    if kh != 3 or kw != 3:
        raise NotImplementedError("pesr_b200.functional.conv2d_same: 3x3 kernels only (model/basic.py uses nothing else "
                                  "besides the 1x1 MeanShift, see mean_shift)")
    if x.dim() != 4 or x.shape[1] != ci:
        raise ValueError(f"conv2d_same: input {tuple(x.shape)} does not match weight {tuple(weight.shape)}")
    if stride not in (1, 2):
        raise NotImplementedError("conv2d_same: stride 1 or 2")
    if ci != 3 and ci % 64 != 0:
        raise NotImplementedError("conv2d_same: in_channels must be 3 or a multiple of 64 (tensor-core K block)")
    if co != 3 and co % 32 != 0:
        raise NotImplementedError("conv2d_same: out_channels must be 3 or a multiple of 32")
    if ci == 3 and (co == 3 or stride != 1):
        raise NotImplementedError("conv2d_same: 3-channel input needs stride 1 and out_channels % 32 == 0")
    if co == 3 and stride != 1:
        raise NotImplementedError("conv2d_same: 3-channel output needs stride 1")


class _Conv2dSame(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride, dtype):
        _check(x, weight, stride)
        x = x.contiguous().float()
        nb, ci, h, w = x.shape
        co = weight.shape[0]
        dt = ops.dt_code(dtype)
        dev = x.device
        ho, wo = (h + stride - 1) // stride, (w + stride - 1) // stride
        stream = torch.cuda.current_stream().cuda_stream
        b32 = bias.detach().float().contiguous() if bias is not None else None
        if ci == 3:
            a16 = torch.empty(nb * h * w, 64, device=dev, dtype=dtype)
            ops.im2col3(x, a16)
            srcs, taps, kw, cin_k = [ops.nhwc_src(a16, nb, h, w, 64)], [(0, 0)], {}, 64
            wp = _packed(weight, 4, dtype, pad_to=64)
        else:
            a16 = torch.empty(nb, h, w, ci, device=dev, dtype=dtype)
            ops.nchw32_to_nhwc16(x, a16)
            cin_k = ci
            if stride == 1:
                srcs, taps, kw = [ops.nhwc_src(a16, nb, h, w, ci)], ops.TAPS_3X3, {}
            else:
                from .engine_d import _parity_planes
                taps, tsrc, widx = _s2_taps()
                srcs, kw = _parity_planes(a16, nb, h, w, ci), dict(tap_src=tsrc, tap_widx=widx)
            wp = _packed(weight, 5 if co == 3 else 0, dtype, pad_to=32 if co == 3 else 0)
        if co == 3:
            z = torch.empty(nb * h * w, 32, device=dev, dtype=torch.float32)
            d = ops.make_conv_desc(dtype=dt, nb=nb, h=h, w=w, cin=cin_k, cout=32, taps=[(0, 0)], srcs=srcs, wpacked=wp,
                                   out32=z, ld_out32=32)
            check(lib.pesr_conv_igemm(C.byref(d), stream), "pesr_conv_igemm")
            out = torch.empty(nb, 3, h, w, device=dev, dtype=torch.float32)
            ops.col2im3(z, 32, nb, h, w, out, bias=b32)
        else:
            y = torch.empty(nb, ho, wo, co, device=dev, dtype=torch.float32)
            d = ops.make_conv_desc(dtype=dt, nb=nb, h=ho, w=wo, cin=cin_k, cout=co, block_n=_block_n(co), taps=taps,
                                   srcs=srcs, wpacked=wp, bias=b32, act=ACT_NONE, out32=y, ld_out32=co, **kw)
            check(lib.pesr_conv_igemm(C.byref(d), stream), "pesr_conv_igemm")
            out = y.permute(0, 3, 1, 2).contiguous()
        ctx.save_for_backward(weight)
        ctx.a16, ctx.geom, ctx.dtype, ctx.has_bias = a16, (nb, ci, co, h, w, ho, wo, stride), dtype, bias is not None
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        (weight,) = ctx.saved_tensors
        nb, ci, co, h, w, ho, wo, stride = ctx.geom
        dtype, a16 = ctx.dtype, ctx.a16
        dt = ops.dt_code(dtype)
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        if co % 64 != 0 or stride != 1:
            raise NotImplementedError("pesr_b200.functional.conv2d_same: backward is implemented for stride-1 convolutions "
                                      "with out_channels % 64 == 0 (use the Generator / Discriminator modules, whose "
                                      "schedules cover every layer of the PESR path)")
        dev = dy.device
        dy = dy.contiguous().float()
        stream = torch.cuda.current_stream().cuda_stream
        ws = torch.zeros(4, device=dev, dtype=torch.float32)
        ops.amax_scale(dy, ws, target=16.0)             # power-of-two range scaling of the 16-bit gradient operand
        scale = ws[1:2]
        dy16 = torch.empty(nb, ho, wo, co, device=dev, dtype=dtype)
        ops.nchw32_to_nhwc16(dy, dy16, mul_dev=scale)
        gx = gw = gb = None
        if need_b:
            gb = torch.zeros(co, device=dev, dtype=torch.float32)
            ops.colsum16(dy16, nb * ho * wo, co, co, gb, div_dev=scale)
        if need_w:
            gw = torch.empty_like(weight, dtype=torch.float32)
            part = torch.empty(max(9 * co * max(ci, 64) * 8, 148 * 128 * 64), device=dev, dtype=torch.float32)
            if ci == 3:
                wd = ops.make_wgrad_desc(dtype=dt, nb=nb, h=h, w=w, a=dy16, a_c=co, m_total=co,
                                         b_srcs=[ops.nhwc_src(a16, nb, h, w, 64)], n_total=64, taps=[(0, 0)], partials=part)
                splits = ops.conv_wgrad(wd)
                ops.wgrad_reduce(part, splits, 1, co, 64, ops.WMAP_COL_IN, co, 3, gw, div_dev=scale)
            else:
                wd = ops.make_wgrad_desc(dtype=dt, nb=nb, h=h, w=w, a=dy16, a_c=co, m_total=co,
                                         b_srcs=[ops.nhwc_src(a16, nb, h, w, ci)], n_total=ci, partials=part)
                splits = ops.conv_wgrad(wd)
                ops.wgrad_reduce(part, splits, 9, co, ci, ops.WMAP_OIHW, co, ci, gw, div_dev=scale)
        if need_x:
            if ci == 3:
                zd = torch.empty(nb * h * w, 32, device=dev, dtype=torch.float32)
                d = ops.make_conv_desc(dtype=dt, nb=nb, h=h, w=w, cin=co, cout=32, taps=[(0, 0)],
                                       srcs=[ops.nhwc_src(dy16, nb, h, w, co)], wpacked=_packed(weight, 6, dtype, pad_to=32),
                                       out32=zd, ld_out32=32)
                check(lib.pesr_conv_igemm(C.byref(d), stream), "pesr_conv_igemm")
                gx = torch.empty(nb, 3, h, w, device=dev, dtype=torch.float32)
                ops.col2im3(zd, 32, nb, h, w, gx, div_dev=scale, sgn=-1)
            else:
                g32 = torch.empty(nb, h, w, ci, device=dev, dtype=torch.float32)
                d = ops.make_conv_desc(dtype=dt, nb=nb, h=h, w=w, cin=co, cout=ci, block_n=_block_n(ci),
                                       srcs=[ops.nhwc_src(dy16, nb, h, w, co)], wpacked=_packed(weight, 1, dtype),
                                       out32=g32, ld_out32=ci)
                check(lib.pesr_conv_igemm(C.byref(d), stream), "pesr_conv_igemm")
                gx = (g32 * ws[2:3]).permute(0, 3, 1, 2).contiguous()
        return gx, gw, gb, None, None


def conv2d_same(x, weight, bias=None, stride=1, dtype=torch.float16):
    """F.conv2d(x, weight, bias, stride, padding=1) for the 3x3 convolutions of model/basic.py:4-7 on the implicit-GEMM
    tensor-core kernel: 16-bit operands (`dtype`), fp32 accumulation and fp32 NCHW output."""
    return _Conv2dSame.apply(x, weight, bias, stride, dtype)


class _MeanShift(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        if not x.is_cuda:
            raise RuntimeError(f"pesr_b200.functional: input is on {x.device}; the B200 path has no CPU fallback")
        if x.dim() != 4 or x.shape[1] != 3 or weight.shape != (3, 3, 1, 1):
            raise ValueError("mean_shift expects [N,3,H,W] and a 3x3x1x1 weight (model/basic.py:9-17)")
        x = x.contiguous().float()
        w9 = weight.detach().reshape(9).float().contiguous()
        out = torch.empty_like(x)
        ops.mean_shift(x, w9, bias.detach().float().contiguous() if bias is not None else None, out)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous().float()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            wt = weight.detach().reshape(3, 3).t().contiguous().reshape(9)
            gx = torch.empty_like(dy)
            ops.mean_shift(dy, wt, None, gx)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            sums = torch.zeros(12, device=dy.device, dtype=torch.float32)
            ops.moments3(dy, x, sums)         # sums[0:9] = sum_p dy[o] x[i], sums[9:12] = sum_p dy[o]
            gw = sums[0:9].view(3, 3, 1, 1).clone()
            gb = sums[9:12].clone() if ctx.has_bias else None
        return gx, gw, gb


def mean_shift(x, weight, bias):
    """The 1x1 colour-space affine of model/basic.py:9-17 (weights stay trainable, as in the reference)."""
    return _MeanShift.apply(x, weight, bias)
