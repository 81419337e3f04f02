"""Python-side wrappers over the C-ABI: one function per entry point, taking torch tensors.

Torch is used here for device memory and streams only; all arithmetic happens inside
``libpesr_b200.so``.  Tensors are passed as raw device pointers plus shapes, on torch's current stream.
"""
import ctypes as C

import os

import torch

from . import _lib
from ._lib import (ACT_LRELU, ACT_NONE, ACT_RELU, DT_BF16, DT_F16, OUT_NORMAL, OUT_SHUFFLE2, OUT_UNSHUFFLE2,
                   WMAP_COL_IN, WMAP_COL_OUT, WMAP_OIHW, WMAP_OIHW_PS, ConvDesc, WgradDesc, check, lib)

_DT = {torch.float16: DT_F16, torch.bfloat16: DT_BF16}


def dt_code(dtype):
    try:
        return _DT[dtype]
    except KeyError:
        raise TypeError(f"pesr_b200: activations must be float16 or bfloat16, got {dtype}")


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    # raw cudaStream_t of torch's current stream on the current device; the C call is ~20x cheaper than
    # torch.cuda.current_stream() (which walks the Python device-index helpers) and this runs once per launch
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.PesrError("pesr_b200 has no CPU path: tensors must live on a CUDA device")


def pick_tile(h, w, nb=1, cin=0):
    """128-pixel tile (tile_n, tile_h, tile_w) for nb images of h x w pixels.

    One image per tile (8x16, or 16x8 for narrow maps) unless that pads more than 20% of the MMA rows and the layer
    is wide enough (cin >= 128) for the multi-chunk pipeline stages that multi-image tiles use instead of halo stages:
    then the tile spans several images, e.g. 2 x 8 x 8 for 24x24 maps and 8 x 4 x 4 for 12x12 maps (no padding)."""
    def padded(tn, th, tw):
        return -(-nb // tn) * tn * -(-h // th) * th * -(-w // tw) * tw
    legacy = (1, 8, 16) if (w % 16 == 0 or w > 24) else (1, 16, 8)
    best = legacy
    if cin >= 128 and nb > 1:
        for tn in (2, 4, 8, 16):
            for tw in (16, 8, 4):
                th = 128 // (tn * tw)
                if th >= 1 and tn * th * tw == 128 and tn <= nb and padded(tn, th, tw) < padded(*best):
                    best = (tn, th, tw)
        if padded(*best) > 0.8 * padded(*legacy):
            best = legacy
    return best


TAPS_3X3 = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)]


def default_block_n(c, choices=(256, 128, 64, 32)):
    """Widest supported GEMM-N tile that divides c (e.g. 64 for 192 or 320 channels)."""
    for bn in choices:
        if c % bn == 0:
            return bn
    raise ValueError(f"pesr_b200: {c} channels is not a multiple of {choices[-1]}")


def make_conv_desc(*, dtype, nb, h, w, cin, cout, block_n=None, taps=TAPS_3X3, tap_src=None, tap_widx=None,
                   srcs, wpacked, bias=None, alpha=1.0, alpha_dev=None, res32=None, ld_res32=0, res16=None,
                   ld_res16=0, act=ACT_NONE, mask16=None, ld_mask16=0, mask_mode=0, out32=None, ld_out32=0,
                   out16=None, ld_out16=0, out_mode=OUT_NORMAL, out_h=0, out_w=0, out_sy=1, out_sx=1, out_oy=0,
                   out_ox=0, out_coff=0, ps_c=0, tile=None, aux_mode=0, ksplit=0, b_mn_major=0, split_stride32=0,
                   bn_sums=None, classes=None):
    """Build a ``pesr_conv_desc``.

    ``classes``: optional list of (ntaps, out_oy, out_ox) sub-problems sharing this launch (``taps`` / ``tap_widx`` list
    their taps class by class), see pesr_conv_desc.ncls.

    ``srcs`` is a list of (tensor_or_ptr, src_h, src_w, stride_n, stride_h, stride_w) NHWC views (element
    strides); the descriptor keeps raw pointers, the caller keeps the tensors alive.
    """
    d = ConvDesc()
    d.dtype = dtype
    d.nb, d.h, d.w, d.cin, d.cout = nb, h, w, cin, cout
    d.block_n = block_n or default_block_n(cout)
    tl = tile or pick_tile(h, w, nb, cin if (ksplit <= 1 and not b_mn_major) else 0)
    tn, th, tw = tl if len(tl) == 3 else (1,) + tuple(tl)
    d.tile_n, d.tile_h, d.tile_w = tn, th, tw
    d.ntaps = len(taps)
    for t, (dh, dw) in enumerate(taps):
        d.tap_dh[t], d.tap_dw[t] = dh, dw
        d.tap_src[t] = tap_src[t] if tap_src is not None else 0
        d.tap_widx[t] = tap_widx[t] if tap_widx is not None else t
    d.nsrc = len(srcs)
    for i, (t, sh_, sw_, sn, sh, sw) in enumerate(srcs):
        d.src[i] = t if isinstance(t, int) else t.data_ptr()
        d.src_h[i], d.src_w[i] = sh_, sw_
        d.src_sn[i], d.src_sh[i], d.src_sw[i] = sn, sh, sw
    d.wpacked = _ptr(wpacked)
    d.w_rows = wpacked.shape[0]
    d.bias = _ptr(bias)
    d.alpha = alpha
    d.alpha_dev = _ptr(alpha_dev)
    d.res32, d.ld_res32 = _ptr(res32), ld_res32
    d.res16, d.ld_res16 = _ptr(res16), ld_res16
    d.act = act
    d.mask16, d.ld_mask16, d.mask_mode = _ptr(mask16), ld_mask16, mask_mode
    d.out32, d.ld_out32 = _ptr(out32), ld_out32
    d.out16, d.ld_out16 = _ptr(out16), ld_out16
    d.out_mode, d.out_h, d.out_w = out_mode, out_h, out_w
    d.out_sy, d.out_sx, d.out_oy, d.out_ox, d.out_coff, d.ps_c = out_sy, out_sx, out_oy, out_ox, out_coff, ps_c
    d.aux_mode = aux_mode
    d.ksplit, d.b_mn_major, d.split_stride32 = ksplit, b_mn_major, split_stride32
    d.bn_sums = _ptr(bn_sums)
    if classes:
        if len(classes) > 4 or sum(c[0] for c in classes) != len(taps):
            raise ValueError("make_conv_desc: at most 4 classes whose tap counts sum to len(taps)")
        d.ncls = len(classes)
        for i, (nt, oy, ox) in enumerate(classes):
            d.cls_ntaps[i], d.cls_oy[i], d.cls_ox[i] = nt, oy, ox
    return d


def nhwc_src(t, nb, h, w, c):
    """(tensor, h, w, sn, sh, sw) for a dense NHWC tensor of c channels per pixel."""
    return (t, h, w, h * w * c, w * c, c)


def conv_igemm(desc):
    check(lib.pesr_conv_igemm(C.byref(desc), _stream()), "pesr_conv_igemm")


def make_wgrad_desc(*, dtype, nb, h, w, a, a_c, m_total, b_srcs, n_total, block_n=None, taps=TAPS_3X3, tap_src=None,
                    partials, splits=0, out_mul=0.0, out_div_dev=None):
    d = WgradDesc()
    d.dtype = dtype
    d.nb, d.h, d.w = nb, h, w
    d.m_total, d.n_total = m_total, n_total
    d.block_m = 128
    d.block_n = block_n or default_block_n(n_total, (256, 128, 64))
    d.ntaps = len(taps)
    for t, (dh, dw) in enumerate(taps):
        d.tap_dh[t], d.tap_dw[t] = dh, dw
        d.tap_src[t] = tap_src[t] if tap_src is not None else 0
    d.a, d.a_c = _ptr(a), a_c
    d.nsrc = len(b_srcs)
    for i, (t, bh, bw, sn, sh, sw) in enumerate(b_srcs):
        d.b[i] = t if isinstance(t, int) else t.data_ptr()
        d.b_h[i], d.b_w[i] = bh, bw
        d.b_sn[i], d.b_sh[i], d.b_sw[i] = sn, sh, sw
    d.splits = splits
    d.partials = _ptr(partials)
    d.partials_elems = partials.numel()
    d.out_mul, d.out_div_dev = out_mul, _ptr(out_div_dev)
    return d


def conv_wgrad(desc):
    s = C.c_int32(0)
    check(lib.pesr_conv_wgrad(C.byref(desc), C.byref(s), _stream()), "pesr_conv_wgrad")
    return s.value


def wgrad_reduce(partials, splits, ntaps, m_total, n_total, map_mode, co, ci, grad, scale=1.0, div_dev=None,
                 accumulate=False):
    check(lib.pesr_wgrad_reduce(_ptr(partials), splits, ntaps, m_total, n_total, map_mode, co, ci, scale,
                                _ptr(div_dev), 1 if accumulate else 0, _ptr(grad), _stream()), "pesr_wgrad_reduce")


def pack_weights(w, mode, out, pad_to=0):
    """fp32 OIHW -> packed 16-bit GEMM operand (see include/pesr_b200.h for the modes)."""
    _need_cuda(w, out)
    co, ci, k = w.shape[0], w.shape[1], w.shape[2]
    check(lib.pesr_pack_weights(_ptr(w), co, ci, k, mode, pad_to, dt_code(out.dtype), _ptr(out), _stream()),
          "pesr_pack_weights")
    return out


def packed_shape(co, ci, ksize, mode, pad_to=0):
    taps = ksize * ksize
    if mode in (0, 2):
        return (taps * co, ci)
    if mode in (1, 3):
        return (taps * ci, co)
    if mode == 4:
        return (co, pad_to)
    if mode == 5:
        return (pad_to, ci)
    if mode == 6:
        return (pad_to, co)
    if mode == 7:
        return (ci, pad_to)
    raise ValueError(mode)


def im2col3_u8(src_u8, col, affine_a=None, affine_b=None):
    """im2col3 straight from a uint8 HWC image batch [nb][h][w][3] (utils.imgs_to_tensors fused, utils.py:20-25)."""
    _need_cuda(src_u8, col)
    nb, h, w, c = src_u8.shape
    assert c == 3 and src_u8.dtype == torch.uint8 and src_u8.is_contiguous()
    check(lib.pesr_im2col3(_ptr(src_u8), nb, h, w, _ptr(affine_a), _ptr(affine_b), 0, 1, 4, dt_code(col.dtype),
                           _ptr(col), _stream()), "pesr_im2col3")
    return col


def im2col3(src, col, affine_a=None, affine_b=None, mul_dev=None, sgn=1, pad_affine=False, upper_zeroed=False, low_part=False):
    """upper_zeroed: columns 32..63 of ``col`` already hold zeros (buffer created with torch.zeros and only ever written
    by this function) - they are not stored again.  Not used by the engines: measured on the GAN step it is 0.4 ms
    SLOWER (18.18 -> 18.57 ms) - the conv that follows then fetches the zero halves from HBM instead of finding the
    freshly written lines in L2."""
    _need_cuda(src, col)
    nb, c, h, w = src.shape
    assert c == 3 and src.dtype == torch.float32 and src.is_contiguous()
    check(lib.pesr_im2col3(_ptr(src), nb, h, w, _ptr(affine_a), _ptr(affine_b), _ptr(mul_dev), sgn,
                           (1 if pad_affine else 0) | (2 if upper_zeroed else 0) | (8 if low_part else 0), dt_code(col.dtype),
                           _ptr(col), _stream()), "pesr_im2col3")
    return col


_COL2IM_TILED = os.environ.get("PESR_COL2IM_TILED", "1") != "0"     # A/B knob: shared-memory col2im (default) vs per-pixel gather


def col2im3(z, ldz, nb, h, w, out, bias=None, affine_a=None, affine_b=None, mul=1.0, div_dev=None, sgn=1, pre=None,
            out_u8=None):
    """out (fp32 NCHW) and / or out_u8 (uint8 HWC [nb][h][w][3], clip + round-half-even fused, utils.py:13-18)."""
    _need_cuda(z, out if out is not None else out_u8)
    if _COL2IM_TILED or out_u8 is not None:
        check(lib.pesr_col2im3_tiled(_ptr(z), ldz, nb, h, w, _ptr(bias), _ptr(affine_a), _ptr(affine_b), mul,
                                     _ptr(div_dev), sgn, _ptr(pre), _ptr(out), _ptr(out_u8), _stream()),
              "pesr_col2im3_tiled")
    else:
        check(lib.pesr_col2im3(_ptr(z), ldz, nb, h, w, _ptr(bias), _ptr(affine_a), _ptr(affine_b), mul, _ptr(div_dev),
                               sgn, _ptr(pre), _ptr(out), _stream()), "pesr_col2im3")
    return out if out is not None else out_u8


def split16(src32, hi, lo, act=ACT_NONE, mask_hi=None, mask_lo=None, mask_mode=0, mul=1.0, mul_dev=None):
    """hi = round16(v), lo = round16(v - hi) with v = act(src32 * mul) * act'(mask) (see pesr_split16)."""
    ref = hi if hi is not None else lo
    check(lib.pesr_split16(_ptr(src32), src32.numel(), act, _ptr(mask_hi), _ptr(mask_lo), mask_mode, mul, _ptr(mul_dev),
                           dt_code(ref.dtype), _ptr(hi), _ptr(lo), _stream()), "pesr_split16")


def colmoments32(a32, npix, c, sums64, b32=None):
    """sums64[0] += column sums of a32 [npix][c]; sums64[1] += column sums of a32 * (b32 or a32)   (fp64)."""
    check(lib.pesr_colmoments32(_ptr(a32), _ptr(b32), npix, c, _ptr(sums64), _stream()), "pesr_colmoments32")


def affine_split(a32, npix, c, out32=None, hi=None, lo=None, ka=None, kb=None, kc=None, b32=None, act=ACT_NONE,
                 mask_hi=None, mask_lo=None, mask32=None, mask_mode=0):
    """v = act(ka*a + kb*b + kc) * act'(mask), per-channel fp32 coefficient vectors; out32 = v, hi / lo = split of v."""
    check(lib.pesr_affine_split(_ptr(a32), _ptr(b32), npix, c, _ptr(ka), _ptr(kb), _ptr(kc), act, _ptr(mask_hi),
                                _ptr(mask_lo), _ptr(mask32), mask_mode, DT_F16, _ptr(out32), _ptr(hi), _ptr(lo), _stream()),
          "pesr_affine_split")


def maxpool2_f32_fwd(x32, nb, h, w, c, y32):
    check(lib.pesr_maxpool2_f32_fwd(_ptr(x32), nb, h, w, c, _ptr(y32), _stream()), "pesr_maxpool2_f32_fwd")


def maxpool2_f32_bwd(x32, dy32, nb, h, w, c, dx32, relu_mask=True):
    check(lib.pesr_maxpool2_f32_bwd(_ptr(x32), _ptr(dy32), nb, h, w, c, 1 if relu_mask else 0, _ptr(dx32), _stream()),
          "pesr_maxpool2_f32_bwd")


def mean_shift(x, w9, b3, out):
    _need_cuda(x, out)
    nb, c, h, w = x.shape
    assert c == 3 and x.dtype == torch.float32 and x.is_contiguous()
    check(lib.pesr_mean_shift(_ptr(x), nb, h * w, _ptr(w9), _ptr(b3), _ptr(out), _stream()), "pesr_mean_shift")
    return out


def psnr_y_sse(a, b, sse):
    """sse[n] += exact integer sum of squared Y-channel differences of images a[n], b[n] (utils.py:27-41)."""
    _need_cuda(a, b, sse)
    nb, c, h, w = a.shape
    assert c == 3 and a.shape == b.shape and a.dtype == b.dtype == torch.float32 and sse.dtype == torch.int64
    assert a.is_contiguous() and b.is_contiguous() and sse.numel() >= nb
    check(lib.pesr_psnr_y_sse(_ptr(a), _ptr(b), nb, h * w, _ptr(sse), _stream()), "pesr_psnr_y_sse")
    return sse


def gather_patches(table_dev, nb, patch, scale, lr, hr):
    check(lib.pesr_gather_patches(_ptr(table_dev), nb, patch, scale, _ptr(lr), _ptr(hr), _stream()),
          "pesr_gather_patches")


def u8hwc_to_f32nchw(src_u8, out):
    """uint8 HWC [nb][h][w][3] -> fp32 NCHW (utils.imgs_to_tensors, utils.py:20-25)."""
    _need_cuda(src_u8, out)
    nb, h, w, c = src_u8.shape
    assert c == 3 and src_u8.dtype == torch.uint8 and src_u8.is_contiguous()
    check(lib.pesr_u8hwc_to_f32nchw_batch(_ptr(src_u8), nb, h, w, _ptr(out), _stream()), "pesr_u8hwc_to_f32nchw_batch")
    return out


def nchw32_to_nhwc16(src, dst, ldc=None, mul_dev=None):
    _need_cuda(src, dst)
    nb, c, h, w = src.shape
    assert src.dtype == torch.float32 and src.is_contiguous()
    check(lib.pesr_nchw32_to_nhwc16(_ptr(src), nb, c, h, w, ldc or c, _ptr(mul_dev), dt_code(dst.dtype), _ptr(dst),
                                    _stream()), "pesr_nchw32_to_nhwc16")
    return dst


def nhwc16_to_nchw32(src, dst, ldc=None, mul=1.0, div_dev=None):
    _need_cuda(src, dst)
    nb, c, h, w = dst.shape
    assert dst.dtype == torch.float32 and dst.is_contiguous()
    check(lib.pesr_nhwc16_to_nchw32(_ptr(src), nb, c, h, w, ldc or c, mul, _ptr(div_dev), dt_code(src.dtype),
                                    _ptr(dst), _stream()), "pesr_nhwc16_to_nchw32")
    return dst


def colsum16(x, npix, c, ldc, out, mul=1.0, div_dev=None, accumulate=False):
    check(lib.pesr_colsum16(_ptr(x), npix, c, ldc, mul, _ptr(div_dev), 1 if accumulate else 0, dt_code(x.dtype),
                            _ptr(out), _stream()), "pesr_colsum16")
    return out


def amax_scale(x, ws3, target=16.0):
    assert x.dtype == torch.float32 and x.is_contiguous()
    check(lib.pesr_amax_scale(_ptr(x), x.numel(), target, _ptr(ws3), _stream()), "pesr_amax_scale")


def moments3(a, b, sums12):
    nb, c, h, w = a.shape
    assert c == 3 and a.is_contiguous() and b.is_contiguous()
    check(lib.pesr_moments3(_ptr(a), _ptr(b), nb, h * w, _ptr(sums12), _stream()), "pesr_moments3")
    return sums12


def loss_l1(a, b, loss, grad=None):
    check(lib.pesr_loss_l1(_ptr(a), _ptr(b), a.numel(), _ptr(loss), _ptr(grad), _stream()), "pesr_loss_l1")


def loss_mse(a, b, loss, grad=None):
    check(lib.pesr_loss_mse(_ptr(a), _ptr(b), a.numel(), _ptr(loss), _ptr(grad), _stream()), "pesr_loss_mse")


def loss_tv(y, loss, grad=None):
    h, w = y.shape[-2], y.shape[-1]
    check(lib.pesr_loss_tv(_ptr(y), y.numel() // (h * w), h, w, _ptr(loss), _ptr(grad), _stream()), "pesr_loss_tv")


def loss_gan(a, b, loss, sign_a=1.0, sign_b=-1.0, target=1.0, mode=0, gamma=1.0, grad_a=None, grad_b=None):
    check(lib.pesr_loss_gan(_ptr(a), _ptr(b), a.numel(), sign_a, sign_b, target, mode, gamma, _ptr(loss),
                            _ptr(grad_a), _ptr(grad_b), _stream()), "pesr_loss_gan")


def adam_multi(table, nchunks, lr, beta1, beta2, eps, step, grad_mul=1.0):
    check(lib.pesr_adam_multi(_ptr(table), nchunks, lr, beta1, beta2, eps, step, grad_mul, _stream()),
          "pesr_adam_multi")


def adam_multi_dev(table, nchunks, lr_dev, beta1, beta2, eps, step_dev, grad_mul=1.0):
    check(lib.pesr_adam_multi_dev(_ptr(table), nchunks, _ptr(lr_dev), beta1, beta2, eps, _ptr(step_dev), grad_mul,
                                  _stream()), "pesr_adam_multi_dev")


def bn_reduce(y16, npix, c, sums_ws, groups=1, zero_first=True):
    """sums_ws[g] += (sum y, sum y^2) per channel of group g (npix pixels per group)."""
    check(lib.pesr_bn_reduce(_ptr(y16), npix, c, groups, _ptr(sums_ws), 1 if zero_first else 0, dt_code(y16.dtype),
                             _stream()), "pesr_bn_reduce")


def bn_lrelu_fwd(y16, npix, c, mean, rstd, gamma, beta, a16, slope=0.2, groups=1, sums_ws=None, eps=1e-5, momentum=0.1,
                 running_mean=None, running_var=None, num_batches=None, running_mean_shift=None):
    """sums_ws given (train mode): mean / rstd are derived from the sums in the kernel, stored, and the running statistics
    updated; sums_ws None (eval mode): mean / rstd are inputs."""
    check(lib.pesr_bn_lrelu_fwd(_ptr(y16), npix, c, groups, _ptr(sums_ws), eps, momentum, _ptr(mean), _ptr(rstd),
                                _ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var), _ptr(num_batches),
                                _ptr(running_mean_shift), slope, dt_code(y16.dtype), _ptr(a16), _stream()),
          "pesr_bn_lrelu_fwd")


def bn_stats(y16, npix, c, sums_ws, mean, rstd, running_mean=None, running_var=None, num_batches=None, eps=1e-5,
             momentum=0.1, running_mean_shift=None):
    """Statistics only (tests / stand-alone use): mean, rstd and the running-statistics update of one train-mode call.
    y16 = None: the sums were accumulated by the producing conv (make_conv_desc(bn_sums=...)).  The engines never call
    this: their apply kernel derives mean / rstd from the sums itself."""
    if y16 is not None:
        bn_reduce(y16, npix, c, sums_ws)
    mean64 = sums_ws[:c] / npix
    var = (sums_ws[c:2 * c] / npix - mean64 * mean64).clamp_min(0)
    mean.copy_(mean64)
    rstd.copy_(torch.rsqrt(var + eps))
    if running_mean is not None:
        shift = running_mean_shift if running_mean_shift is not None else 0
        running_mean.mul_(1 - momentum).add_(momentum * (mean64.float() + shift))
        running_var.mul_(1 - momentum).add_(momentum * (var * (npix / max(npix - 1, 1))).float())
    if num_batches is not None:
        num_batches += 1
    sums_ws[:2 * c].zero_()


def bn_lrelu_bwd(dz16, y16, npix, c, mean, rstd, gamma, sums_ws, dy16, dgamma, dbeta, grad_mul=1.0, grad_div_dev=None,
                 accumulate=False, groups=1, zero_first=True):
    check(lib.pesr_bn_lrelu_bwd(_ptr(dz16), _ptr(y16), npix, c, groups, _ptr(mean), _ptr(rstd), _ptr(gamma),
                                _ptr(sums_ws), 1 if zero_first else 0, grad_mul, _ptr(grad_div_dev), dt_code(y16.dtype),
                                _ptr(dy16), _ptr(dgamma), _ptr(dbeta), 1 if accumulate else 0, _stream()),
          "pesr_bn_lrelu_bwd")


def maxpool2_fwd(x16, nb, h, w, c, y16):
    check(lib.pesr_maxpool2_fwd(_ptr(x16), nb, h, w, c, dt_code(x16.dtype), _ptr(y16), _stream()), "pesr_maxpool2_fwd")


def maxpool2_bwd(x16, dy16, nb, h, w, c, dx16, relu_mask=True):
    check(lib.pesr_maxpool2_bwd(_ptr(x16), _ptr(dy16), nb, h, w, c, 1 if relu_mask else 0, dt_code(x16.dtype),
                                _ptr(dx16), _stream()), "pesr_maxpool2_bwd")


def linear_workspace_floats(nb, k, o):
    return int(lib.pesr_linear_workspace_floats(nb, k, o))


_LINEAR_ROWS = 16     # rows per pass of the skinny Linear kernels (pesr_linear_skinny_*); larger batches are chunked


def linear_fwd(x16, w16, bias, nb, k, o, workspace, out32=None, out16=None, act=ACT_NONE):
    x16, out32, out16 = x16.view(nb, k), (out32.view(nb, o) if out32 is not None else None), \
        (out16.view(nb, o) if out16 is not None else None)
    for r in range(0, nb, _LINEAR_ROWS):
        n = min(_LINEAR_ROWS, nb - r)
        check(lib.pesr_linear_skinny_fwd(_ptr(x16[r:]), _ptr(w16), _ptr(bias), n, k, o, act, dt_code(x16.dtype),
                                         _ptr(workspace), _ptr(out32[r:] if out32 is not None else None),
                                         _ptr(out16[r:] if out16 is not None else None), _stream()),
              "pesr_linear_skinny_fwd")


def linear_finalize(partials, ksplit, nb, o, bias, dtype, out32=None, out16=None, act=ACT_NONE):
    check(lib.pesr_linear_finalize(_ptr(partials), ksplit, nb, o, _ptr(bias), act, dt_code(dtype), _ptr(out32),
                                   _ptr(out16), _stream()), "pesr_linear_finalize")


def linear_dgrad(dy32, w16, nb, k, o, dx32):
    dy32, dx32 = dy32.view(nb, o), dx32.view(nb, k)
    for r in range(0, nb, _LINEAR_ROWS):
        n = min(_LINEAR_ROWS, nb - r)
        check(lib.pesr_linear_skinny_dgrad(_ptr(dy32[r:]), _ptr(w16), n, k, o, dt_code(w16.dtype), _ptr(dx32[r:]),
                                           _stream()), "pesr_linear_skinny_dgrad")


def linear_wgrad(dy32, x16, nb, k, o, dw, mul=1.0, div_dev=None, accumulate=False):
    dy32, x16 = dy32.view(nb, o), x16.view(nb, k)
    for r in range(0, nb, 2 * _LINEAR_ROWS):          # the weight-gradient kernel takes up to 32 rows per pass
        n = min(2 * _LINEAR_ROWS, nb - r)
        check(lib.pesr_linear_skinny_wgrad(_ptr(dy32[r:]), _ptr(x16[r:]), n, k, o, mul, _ptr(div_dev),
                                           1 if (accumulate or r > 0) else 0, dt_code(x16.dtype), _ptr(dw), _stream()),
              "pesr_linear_skinny_wgrad")


def cast16(src32, dst16):
    check(lib.pesr_cast16(_ptr(src32), src32.numel(), dt_code(dst16.dtype), _ptr(dst16), _stream()), "pesr_cast16")


def flatten_nchw16(src_nhwc16, nb, hw, c, dst):
    check(lib.pesr_flatten_nchw16(_ptr(src_nhwc16), nb, hw, c, _ptr(dst), _stream()), "pesr_flatten_nchw16")


def unflatten_nchw16(src32, mask16, nb, hw, c, dst16, mul=1.0, mul_dev=None, slope=0.2):
    check(lib.pesr_unflatten_nchw16(_ptr(src32), _ptr(mask16), nb, hw, c, mul, _ptr(mul_dev), slope,
                                    dt_code(dst16.dtype), _ptr(dst16), _stream()), "pesr_unflatten_nchw16")


class MultiPack:
    """All weight packs of one network as one launch (re-run only when a parameter version changed).

    ``companions``: optional list parallel to ``packed_weights`` holding, per forward pack (mode 0 / 2), the
    backward-data pack (mode 1 / 3) of the same parameter or None: both layouts are then written from one read of
    the fp32 tensor (pesr_pack_weights_multi's dst2)."""

    def __init__(self, packed_weights, device, dtype, companions=None):
        import numpy as np
        self.pws = list(packed_weights)
        self.companions = list(companions) if companions is not None else [None] * len(self.pws)
        rows = []
        for pw, cw in zip(self.pws, self.companions):
            p = pw.param
            if cw is not None:
                ok = (cw.param is p and cw.mode == pw.mode + 1 and pw.mode in (0, 2) and p.shape[2] == 3
                      and p.shape[0] % 64 == 0 and p.shape[1] % 32 == 0)
                if not ok:
                    raise ValueError("MultiPack: companion must be the mode+1 pack of the same 3x3 parameter "
                                     "(out channels % 64 == 0, in channels % 32 == 0)")
            rows.append([p.data_ptr(), pw.buf.data_ptr(), p.shape[0], p.shape[1], p.shape[2], pw.mode, pw.pad_to,
                         cw.buf.data_ptr() if cw is not None else 0])
        self.host = np.ascontiguousarray(np.array(rows, dtype=np.int64))
        self.dev = torch.empty(len(rows) * 64, device=device, dtype=torch.uint8)
        self.dt = dt_code(dtype)
        self.uploaded = False
        self.key = None

    def run(self):
        key = tuple(pw.param._version for pw in self.pws)
        if key == self.key:
            return
        check(lib.pesr_pack_weights_multi(self.host.ctypes.data, len(self.pws), _ptr(self.dev),
                                          0 if self.uploaded else 1, self.dt, _stream()), "pesr_pack_weights_multi")
        self.uploaded = True
        self.key = key
        for pw, cw, v in zip(self.pws, self.companions, key):
            pw.key = (pw.param.data_ptr(), v)
            if cw is not None:
                cw.key = (cw.param.data_ptr(), v)
