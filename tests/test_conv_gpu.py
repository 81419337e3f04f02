"""GPU per-op parity: the tcgen05 implicit-GEMM kernels against torch's fp32 convolution on the same
16-bit-rounded operands (accumulation-order error only, so the bar is 2e-5 relative L2 on fp32 outputs)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _nhwc16(x, dtype):
    from pesr_b200 import ops
    nb, c, h, w = x.shape
    out = torch.empty(nb, h, w, c, device="cuda", dtype=dtype)
    ops.nchw32_to_nhwc16(x.contiguous(), out)
    return out


def _nchw32(x16, nb, c, h, w):
    from pesr_b200 import ops
    out = torch.empty(nb, c, h, w, device="cuda", dtype=torch.float32)
    ops.nhwc16_to_nchw32(x16, out)
    return out


CASES = [  # nb, cin, cout, h, w   (G trunk, upsampler, D/VGG channel counts, ragged and tiny grids)
    (2, 256, 256, 48, 48), (1, 256, 1024, 16, 16), (1, 64, 64, 40, 24), (1, 64, 128, 24, 24),
    (2, 128, 128, 13, 19), (1, 512, 512, 12, 12), (1, 128, 256, 5, 3), (3, 64, 64, 1, 1),
    # multi-image pixel tiles (ops.pick_tile: 2 x 8 x 8 for 24x24 maps, 8 x 4 x 4 for 12x12), even and ragged batches
    (4, 256, 512, 24, 24), (16, 512, 512, 12, 12), (3, 128, 128, 12, 12), (5, 256, 256, 6, 6),
]


@pytest.fixture(params=[1, 2, 0], ids=["pair-auto", "pair-forced", "pair-off"])
def pair_mode(request):
    """Runs a test with the CTA-pair (cta_group::2) kernel chosen automatically, forced wherever legal, and disabled."""
    from pesr_b200 import _lib
    _lib.set_option(_lib.OPT_PAIR_MODE, request.param)
    yield request.param
    _lib.set_option(_lib.OPT_PAIR_MODE, 1)


@pytest.mark.parametrize("nb,cin,cout,h,w", CASES)
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_fprop_bias_relu(nb, cin, cout, h, w, dtype, pair_mode):
    from pesr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(nb * 1000 + cin + h)
    x = torch.randn(nb, cin, h, w, device="cuda", generator=g)
    wt = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3 * cin ** 0.5)
    b = torch.randn(cout, device="cuda", generator=g)
    ref = F.conv2d(x.to(dtype).float(), wt.to(dtype).float(), b, padding=1).relu()
    wp = torch.empty(ops.packed_shape(cout, cin, 3, 0), device="cuda", dtype=dtype)
    ops.pack_weights(wt, 0, wp)
    out32 = torch.full((nb, h, w, cout), float("nan"), device="cuda")
    out16 = torch.empty(nb, h, w, cout, device="cuda", dtype=dtype)
    d = ops.make_conv_desc(dtype=ops.dt_code(dtype), nb=nb, h=h, w=w, cin=cin, cout=cout,
                           srcs=[ops.nhwc_src(_nhwc16(x, dtype), nb, h, w, cin)], wpacked=wp, bias=b,
                           act=ops.ACT_RELU, out16=out16, ld_out16=cout, out32=out32, ld_out32=cout)
    ops.conv_igemm(d)
    assert rel_l2(out32.permute(0, 3, 1, 2), ref) < 2e-5
    assert rel_l2(_nchw32(out16, nb, cout, h, w), ref) < (3e-4 if dtype == torch.float16 else 3e-3)


@pytest.mark.parametrize("nb,cin,cout,h,w", CASES[:6])
def test_dgrad_is_fprop_with_flipped_weights(nb, cin, cout, h, w):
    from pesr_b200 import ops
    dtype = torch.float16
    g = torch.Generator(device="cuda").manual_seed(7)
    dy = torch.randn(nb, cout, h, w, device="cuda", generator=g)
    wt = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3 * cout ** 0.5)
    ref = torch.nn.grad.conv2d_input((nb, cin, h, w), wt.half().float(), dy.half().float(), padding=1)
    wp = torch.empty(ops.packed_shape(cout, cin, 3, 1), device="cuda", dtype=dtype)
    ops.pack_weights(wt, 1, wp)
    out32 = torch.empty(nb, h, w, cin, device="cuda")
    d = ops.make_conv_desc(dtype=0, nb=nb, h=h, w=w, cin=cout, cout=cin,
                           srcs=[ops.nhwc_src(_nhwc16(dy, dtype), nb, h, w, cout)], wpacked=wp, out32=out32, ld_out32=cin)
    ops.conv_igemm(d)
    assert rel_l2(out32.permute(0, 3, 1, 2), ref) < 2e-5


def test_epilogue_residual_scale_mask_and_pixel_shuffle(pair_mode):
    from pesr_b200 import ops
    dtype = torch.float16
    nb, c, h, w = 2, 64, 12, 16
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(nb, c, h, w, device="cuda", generator=g)
    res = torch.randn(nb, h, w, c, device="cuda", generator=g)
    mask = torch.randn(nb, c, h, w, device="cuda", generator=g)
    wt = torch.randn(c, c, 3, 3, device="cuda", generator=g) / 24
    b = torch.randn(c, device="cuda", generator=g)
    wp = torch.empty(9 * c, c, device="cuda", dtype=dtype)
    ops.pack_weights(wt, 0, wp)
    out32 = res.clone()
    d = ops.make_conv_desc(dtype=0, nb=nb, h=h, w=w, cin=c, cout=c, srcs=[ops.nhwc_src(_nhwc16(x, dtype), nb, h, w, c)],
                           wpacked=wp, bias=b, alpha=0.1, res32=out32, ld_res32=c, mask16=_nhwc16(mask, dtype),
                           ld_mask16=c, mask_mode=1, out32=out32, ld_out32=c)
    ops.conv_igemm(d)
    ref = (0.1 * F.conv2d(x.half().float(), wt.half().float(), b, padding=1) + res.permute(0, 3, 1, 2)) \
        * (mask.half().float() > 0)
    assert rel_l2(out32.permute(0, 3, 1, 2), ref) < 2e-5
    # PixelShuffle(2) fused store and its inverse
    wt4 = torch.randn(4 * c, c, 3, 3, device="cuda", generator=g) / 24
    wp4 = torch.empty(9 * 4 * c, c, device="cuda", dtype=dtype)
    ops.pack_weights(wt4, 2, wp4)
    up = torch.empty(nb, 2 * h, 2 * w, c, device="cuda", dtype=dtype)
    d = ops.make_conv_desc(dtype=0, nb=nb, h=h, w=w, cin=c, cout=4 * c, block_n=c,
                           srcs=[ops.nhwc_src(_nhwc16(x, dtype), nb, h, w, c)], wpacked=wp4, out16=up, ld_out16=c,
                           out_mode=ops.OUT_SHUFFLE2, ps_c=c)
    ops.conv_igemm(d)
    ref_up = F.pixel_shuffle(F.conv2d(x.half().float(), wt4.half().float(), None, padding=1), 2)
    assert rel_l2(_nchw32(up, nb, c, 2 * h, 2 * w), ref_up) < 3e-4
    # un-shuffle: identity 1x1 GEMM from the shuffled tensor back to [h][w][4c] packed (ij, c)
    eye = torch.eye(c, device="cuda").view(c, c, 1, 1).contiguous()
    wpe = torch.empty(c, c, device="cuda", dtype=dtype)
    ops.pack_weights(eye, 0, wpe)
    back = torch.empty(nb, h, w, 4 * c, device="cuda", dtype=dtype)
    d = ops.make_conv_desc(dtype=0, nb=nb, h=2 * h, w=2 * w, cin=c, cout=c, taps=[(0, 0)],
                           srcs=[ops.nhwc_src(up, nb, 2 * h, 2 * w, c)], wpacked=wpe, out16=back, ld_out16=4 * c,
                           out_mode=ops.OUT_UNSHUFFLE2)
    ops.conv_igemm(d)
    ref_back = F.pixel_unshuffle(_nchw32(up, nb, c, 2 * h, 2 * w), 2)      # channel c*4 + ij
    got = _nchw32(back, nb, 4 * c, h, w).view(nb, 4, c, h, w).permute(0, 2, 1, 3, 4).reshape(nb, 4 * c, h, w)
    assert torch.equal(got, ref_back)


@pytest.mark.parametrize("nb,cin,cout,h,w", [(2, 256, 256, 48, 48), (1, 64, 64, 24, 24), (1, 128, 256, 20, 20),
                                               (2, 64, 128, 7, 5), (1, 256, 1024, 16, 16)])
def test_wgrad_split_k(nb, cin, cout, h, w):
    from pesr_b200 import ops
    dtype = torch.float16
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(nb, cin, h, w, device="cuda", generator=g)
    dy = torch.randn(nb, cout, h, w, device="cuda", generator=g)
    ref = torch.nn.grad.conv2d_weight(x.half().float(), (cout, cin, 3, 3), dy.half().float(), padding=1)
    part = torch.empty(32 * 9 * cout * cin, device="cuda")
    d = ops.make_wgrad_desc(dtype=0, nb=nb, h=h, w=w, a=_nhwc16(dy, dtype), a_c=cout, m_total=cout,
                            b_srcs=[ops.nhwc_src(_nhwc16(x, dtype), nb, h, w, cin)], n_total=cin, partials=part)
    splits = ops.conv_wgrad(d)
    grad = torch.empty(cout, cin, 3, 3, device="cuda")
    ops.wgrad_reduce(part, splits, 9, cout, cin, ops.WMAP_OIHW, cout, cin, grad, scale=0.5)
    assert rel_l2(grad, 0.5 * ref) < 2e-5


def test_im2col_col2im_roundtrip_and_moments():
    from pesr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.rand(2, 3, 9, 11, device="cuda", generator=g) * 255
    A = torch.randn(3, 3, device="cuda", generator=g)
    bvec = torch.randn(3, device="cuda", generator=g)
    col = torch.empty(2 * 9 * 11, 64, device="cuda", dtype=torch.float16)
    ops.im2col3(x, col, affine_a=A, affine_b=bvec)
    xa = torch.einsum("oi,nihw->nohw", A, x) + bvec.view(1, 3, 1, 1)
    ref = F.unfold(xa, 3, padding=1).view(2, 3, 9, 99).permute(0, 3, 2, 1).reshape(198, 27)   # [p][tap*3+c]
    assert rel_l2(col[:, :27].float(), ref.half().float()) < 1e-6
    assert float(col[:, 27:].abs().max()) == 0
    z = torch.randn(198, 32, device="cuda", generator=g)
    out = torch.empty(2, 3, 9, 11, device="cuda")
    ops.col2im3(z, 32, 2, 9, 11, out, sgn=1)
    zz = z[:, :27].view(2, 99, 9, 3).permute(0, 3, 2, 1).reshape(2, 27, 99)     # [n][c*9+tap][p] for fold
    # out[p][c] = sum_tap z[p + off(tap)][tap*3+c]  == correlation gather; check against explicit loops
    ref2 = torch.zeros(2, 3, 9, 11, device="cuda")
    zimg = z[:, :27].view(2, 9, 11, 9, 3)
    for tap in range(9):
        dy, dx = tap // 3 - 1, tap % 3 - 1
        ys, ye = max(0, -dy), min(9, 9 - dy)
        xs, xe = max(0, -dx), min(11, 11 - dx)
        ref2[:, :, ys:ye, xs:xe] += zimg[:, ys + dy:ye + dy, xs + dx:xe + dx, tap, :].permute(0, 3, 1, 2)
    assert rel_l2(out, ref2) < 1e-6
    del zz
    sums = torch.empty(12, device="cuda")
    a, b2 = torch.randn(2, 3, 9, 11, device="cuda", generator=g), torch.randn(2, 3, 9, 11, device="cuda", generator=g)
    ops.moments3(a, b2, sums)
    assert rel_l2(sums[:9].view(3, 3), torch.einsum("nohw,nihw->oi", a, b2)) < 1e-5
    assert rel_l2(sums[9:], a.sum(dim=(0, 2, 3))) < 1e-5


@pytest.mark.parametrize("nb,cin,cout,h,w", [(2, 64, 64, 24, 20), (16, 128, 256, 48, 48), (3, 64, 128, 7, 5)])
def test_bn_statistics_epilogue(nb, cin, cout, h, w, pair_mode):
    """pesr_conv_desc.bn_sums: per-channel sum / sum of squares of the ROUNDED 16-bit outputs, accumulated by the conv
    epilogue (train-mode BatchNorm statistics, model/basic.py:29), against torch on the tensor the kernel stored."""
    from pesr_b200 import ops
    dtype = torch.float16
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(nb, cin, h, w, device="cuda", generator=g)
    wt = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3 * cin ** 0.5)
    wp = torch.empty(9 * cout, cin, device="cuda", dtype=dtype)
    ops.pack_weights(wt, 0, wp)
    y = torch.empty(nb, h, w, cout, device="cuda", dtype=dtype)
    sums = torch.zeros(2 * cout, device="cuda", dtype=torch.float64)
    d = ops.make_conv_desc(dtype=0, nb=nb, h=h, w=w, cin=cin, cout=cout, srcs=[ops.nhwc_src(_nhwc16(x, dtype), nb, h, w, cin)],
                           wpacked=wp, out16=y, ld_out16=cout, bn_sums=sums)
    ops.conv_igemm(d)
    ops.conv_igemm(d)          # accumulates: two passes = twice the sums
    yd = y.double().reshape(-1, cout)
    ref = torch.cat([yd.sum(0), (yd * yd).sum(0)]) * 2
    assert rel_l2(_nchw32(y, nb, cout, h, w), F.conv2d(x.half().float(), wt.half().float(), None, padding=1)) < 3e-4
    assert float((sums[:cout] - ref[:cout]).abs().max()) < 2e-4 * float(yd.abs().sum(0).max())
    assert rel_l2(sums[cout:], ref[cout:]) < 1e-5
    # finalisation only (y16 = None): mean / rstd from the accumulated sums, sums zeroed again
    mean, rstd = torch.empty(cout, device="cuda"), torch.empty(cout, device="cuda")
    sums.mul_(0.5)
    ops.bn_stats(None, nb * h * w, cout, sums, mean, rstd)
    assert float((mean.double() - yd.mean(0)).abs().max()) < 1e-5 * float(yd.abs().mean()) + 1e-6
    assert rel_l2(rstd.double(), 1.0 / torch.sqrt(yd.var(0, unbiased=False) + 1e-5)) < 1e-5
    assert float(sums.abs().max()) == 0.0


@pytest.mark.parametrize("nb,cin,cout,h,w", [(2, 256, 256, 48, 48), (1, 128, 256, 20, 20), (2, 64, 128, 7, 5)])
def test_wgrad_reduce_fused_with_bias_gradient(nb, cin, cout, h, w):
    """pesr_wgrad_reduce_bias: the split-K reduction of the 3x3 weight gradient and the bias gradient (column sums of
    dY) of the same layer in one launch, against torch's conv2d_weight / a plain sum; the call also clears the buffer
    the next call of a chain accumulates into."""
    from pesr_b200 import ops
    from pesr_b200._lib import check, lib
    dtype = torch.float16
    g = torch.Generator(device="cuda").manual_seed(13)
    x = torch.randn(nb, cin, h, w, device="cuda", generator=g)
    dy = torch.randn(nb, cout, h, w, device="cuda", generator=g)
    dy16 = _nhwc16(dy, dtype)
    part = torch.empty(32 * 9 * cout * cin, device="cuda")
    d = ops.make_wgrad_desc(dtype=0, nb=nb, h=h, w=w, a=dy16, a_c=cout, m_total=cout,
                            b_srcs=[ops.nhwc_src(_nhwc16(x, dtype), nb, h, w, cin)], n_total=cin, partials=part)
    splits = ops.conv_wgrad(d)
    grad = torch.empty(cout, cin, 3, 3, device="cuda")
    bias = torch.zeros(cout, device="cuda")
    nxt = torch.full((cout + 3,), 7.0, device="cuda")
    scale = torch.tensor([4.0], device="cuda")
    check(lib.pesr_wgrad_reduce_bias(part.data_ptr(), splits, 9, cout, cin, ops.WMAP_OIHW, cout, cin, 0.5, scale.data_ptr(), 0,
                                     grad.data_ptr(), dy16.data_ptr(), nb * h * w, cout, cout, 2.0, 0, bias.data_ptr(),
                                     nxt.data_ptr(), cout, torch.cuda.current_stream().cuda_stream), "pesr_wgrad_reduce_bias")
    ref_w = torch.nn.grad.conv2d_weight(x.half().float(), (cout, cin, 3, 3), dy.half().float(), padding=1)
    assert rel_l2(grad, 0.125 * ref_w) < 2e-5
    ref_b = dy.half().double().sum(dim=(0, 2, 3)) * 0.5
    assert float((bias.double() - ref_b).abs().max()) < 1e-5 * float(dy.half().double().abs().sum(dim=(0, 2, 3)).max())
    assert float(nxt[:cout].abs().max()) == 0.0 and float(nxt[cout:].min()) == 7.0


@pytest.mark.parametrize("co,ci,mode", [(256, 256, 0), (1024, 256, 2), (64, 64, 0), (128, 96, 0)])
def test_multi_pack_writes_both_layouts_from_one_read(co, ci, mode):
    """pesr_pack_weights_multi with a companion destination (tiled kernel): forward layout (mode 0 / 2) and
    backward-data layout (mode 1 / 3) bit-identical to the single-job packs."""
    from pesr_b200 import ops
    from pesr_b200.engine_g import PackedWeight
    g = torch.Generator(device="cuda").manual_seed(co + ci + mode)
    w = torch.nn.Parameter(torch.randn(co, ci, 3, 3, device="cuda", generator=g))
    for dtype in (torch.float16, torch.bfloat16):
        f, b = PackedWeight(w, mode, dtype), PackedWeight(w, mode + 1, dtype)
        f.buf.fill_(9.0), b.buf.fill_(9.0)
        ops.MultiPack([f], w.device, dtype, companions=[b]).run()
        ref_f = torch.empty_like(f.buf)
        ref_b = torch.empty_like(b.buf)
        ops.pack_weights(w.detach(), mode, ref_f)
        ops.pack_weights(w.detach(), mode + 1, ref_b)
        assert torch.equal(f.buf, ref_f) and torch.equal(b.buf, ref_b)
        assert b.key == (w.data_ptr(), w._version)


@pytest.mark.parametrize("nb,cin,cout,h,w", [(5, 64, 64, 96, 96), (3, 64, 128, 128, 96), (3, 128, 64, 96, 128),
                                             (7, 64, 64, 100, 72)])
@pytest.mark.parametrize("resident", [True, False], ids=["weights-resident", "weights-per-stage"])
def test_narrow_layers_with_resident_weights(nb, cin, cout, h, w, resident):
    """The N = 64 / 128 layers of VGG and the Discriminator (>= 2 tiles per SM, one column tile) keep their whole packed
    weight tensor in shared memory (ConvK::wres_bytes): the light (bias + ReLU) and the mask (backward-data) epilogues
    against torch, and bit-equal to the per-stage weight path."""
    from pesr_b200 import _lib, ops
    dtype = torch.float16
    g = torch.Generator(device="cuda").manual_seed(nb * 31 + cin + h)
    x = torch.randn(nb, cin, h, w, device="cuda", generator=g)
    wt = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3 * cin ** 0.5)
    b = torch.randn(cout, device="cuda", generator=g)
    m = torch.randn(nb, cout, h, w, device="cuda", generator=g)
    ref = F.conv2d(x.half().float(), wt.half().float(), b, padding=1)
    wp = torch.empty(ops.packed_shape(cout, cin, 3, 0), device="cuda", dtype=dtype)
    ops.pack_weights(wt, 0, wp)
    x16, m16 = _nhwc16(x, dtype), _nhwc16(m, dtype)
    outs = []
    for res in ((True, False) if resident else (False,)):
        _lib.set_option(_lib.OPT_RESIDENT_WEIGHTS, 2 if res else 0)          # 2: wherever legal (the default takes it for N = 128)
        try:
            o_relu = torch.empty(nb, h, w, cout, device="cuda", dtype=dtype)
            o_mask = torch.empty(nb, h, w, cout, device="cuda", dtype=dtype)
            ops.conv_igemm(ops.make_conv_desc(dtype=0, nb=nb, h=h, w=w, cin=cin, cout=cout, srcs=[ops.nhwc_src(x16, nb, h, w, cin)],
                                              wpacked=wp, bias=b, act=ops.ACT_RELU, out16=o_relu, ld_out16=cout))
            ops.conv_igemm(ops.make_conv_desc(dtype=0, nb=nb, h=h, w=w, cin=cin, cout=cout, srcs=[ops.nhwc_src(x16, nb, h, w, cin)],
                                              wpacked=wp, bias=b, mask16=m16, ld_mask16=cout, mask_mode=2, out16=o_mask, ld_out16=cout))
            # a second launch reuses the plan (and follows another conv in the stream: PDL prologue with early weights)
            o_again = torch.empty_like(o_relu)
            ops.conv_igemm(ops.make_conv_desc(dtype=0, nb=nb, h=h, w=w, cin=cin, cout=cout, srcs=[ops.nhwc_src(x16, nb, h, w, cin)],
                                              wpacked=wp, bias=b, act=ops.ACT_RELU, out16=o_again, ld_out16=cout))
        finally:
            _lib.set_option(_lib.OPT_RESIDENT_WEIGHTS, 1)
        assert rel_l2(_nchw32(o_relu, nb, cout, h, w), ref.relu()) < 3e-4
        assert rel_l2(_nchw32(o_mask, nb, cout, h, w), ref * torch.where(m.half().float() > 0, 1.0, 0.2)) < 3e-4
        assert torch.equal(o_again, o_relu)
        outs.append((o_relu, o_mask))
    if len(outs) == 2:
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
