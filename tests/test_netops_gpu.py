"""GPU per-op parity of the Discriminator / VGG layer kernels and the fused losses against plain PyTorch
fp32/fp64 references of the same op on the same (16-bit-rounded) operands."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _nhwc16(x, dtype=torch.float16):
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


@pytest.mark.parametrize("nb,cin,cout,h,w", [(2, 64, 64, 48, 48), (1, 128, 128, 24, 20), (2, 256, 256, 13, 9),
                                               (1, 512, 512, 6, 6)])
def test_stride2_fprop_dgrad_wgrad(nb, cin, cout, h, w):
    """model/pesr.py:63 stride-2 convs: parity-plane fprop, four-class dgrad, parity-plane wgrad."""
    from pesr_b200 import ops
    from pesr_b200.engine_d import _parity_planes, _s2_taps
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(nb, cin, h, w, device="cuda", generator=g)
    wt = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3 * cin ** 0.5)
    xr, wr = x.half().float(), wt.half().float()
    ref = F.conv2d(xr, wr, None, stride=2, padding=1)
    ho, wo = ref.shape[2], ref.shape[3]
    x16 = _nhwc16(x)
    taps, srcs, widx = _s2_taps()
    wp = torch.empty(9 * cout, cin, device="cuda", dtype=torch.float16)
    ops.pack_weights(wt, 0, wp)
    out32 = torch.empty(nb, ho, wo, cout, device="cuda")
    ops.conv_igemm(ops.make_conv_desc(dtype=0, nb=nb, h=ho, w=wo, cin=cin, cout=cout, taps=taps, tap_src=srcs,
                                      tap_widx=widx, srcs=_parity_planes(x16, nb, h, w, cin), wpacked=wp, out32=out32,
                                      ld_out32=cout))
    assert rel_l2(out32.permute(0, 3, 1, 2), ref) < 2e-5
    # dgrad with LeakyReLU' mask fused, scattered to the four parity classes
    dy = torch.randn(nb, cout, ho, wo, device="cuda", generator=g)
    mask = torch.randn(nb, cin, h, w, device="cuda", generator=g)
    dyr = dy.half().float()
    ref_dx = torch.nn.grad.conv2d_input((nb, cin, h, w), wr, dyr, stride=2, padding=1) * \
        torch.where(mask.half().float() > 0, 1.0, 0.2)
    wpd = torch.empty(9 * cin, cout, device="cuda", dtype=torch.float16)
    ops.pack_weights(wt, 1, wpd)
    dy16, mask16 = _nhwc16(dy), _nhwc16(mask)
    dx16 = torch.full((nb, h, w, cin), float("nan"), device="cuda", dtype=torch.float16)
    for ph in range(2):
        for pw in range(2):
            gh, gw = (h - ph + 1) // 2, (w - pw + 1) // 2
            if gh == 0 or gw == 0:
                continue
            ys = [(1, 0)] if ph == 0 else [(0, 1), (2, 0)]
            xs = [(1, 0)] if pw == 0 else [(0, 1), (2, 0)]
            t2 = [(oy, ox) for (dy_, oy) in ys for (dx_, ox) in xs]
            wi = [8 - (dy_ * 3 + dx_) for (dy_, oy) in ys for (dx_, ox) in xs]
            ops.conv_igemm(ops.make_conv_desc(dtype=0, nb=nb, h=gh, w=gw, cin=cout, cout=cin, taps=t2, tap_widx=wi,
                                              srcs=[ops.nhwc_src(dy16, nb, ho, wo, cout)], wpacked=wpd, mask16=mask16,
                                              ld_mask16=cin, mask_mode=2, out16=dx16, ld_out16=cin, out_h=h, out_w=w,
                                              out_sy=2, out_sx=2, out_oy=ph, out_ox=pw, aux_mode=1))
    assert rel_l2(_nchw32(dx16, nb, cin, h, w), ref_dx) < 4e-4
    # wgrad
    ref_dw = torch.nn.grad.conv2d_weight(xr, (cout, cin, 3, 3), dyr, stride=2, padding=1)
    part = torch.empty(32 * 9 * cout * cin, device="cuda")
    d = ops.make_wgrad_desc(dtype=0, nb=nb, h=ho, w=wo, a=dy16, a_c=cout, m_total=cout,
                            b_srcs=_parity_planes(x16, nb, h, w, cin), n_total=cin, taps=taps, tap_src=srcs, partials=part)
    splits = ops.conv_wgrad(d)
    dw = torch.empty(cout, cin, 3, 3, device="cuda")
    ops.wgrad_reduce(part, splits, 9, cout, cin, ops.WMAP_OIHW, cout, cin, dw)
    assert rel_l2(dw, ref_dw) < 2e-5


@pytest.mark.parametrize("groups", [1, 2])
@pytest.mark.parametrize("nb,c,h,w", [(4, 64, 24, 24), (2, 512, 3, 3), (16, 128, 12, 12), (3, 24, 5, 7)])
def test_batchnorm_lrelu_forward_backward(nb, c, h, w, groups):
    """model/basic.py:29-30 in train mode against torch's batch_norm in fp64, as the engines run it: statistics sums,
    then ONE apply kernel that derives mean / rstd from the sums and updates the running statistics; `groups` independent
    batches (= separate module calls, each with its own statistics, running statistics updated call after call) in one
    launch; backward with the parameter gradients summed over the groups."""
    from pesr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    nt = groups * nb
    y = torch.randn(nt, c, h, w, device="cuda", generator=g) * 3 + 5 * torch.randn(1, c, 1, 1, device="cuda", generator=g)
    y[nb:] += 2.0                                   # the groups have different statistics
    gamma = torch.rand(c, device="cuda", generator=g) + 0.5
    beta = torch.randn(c, device="cuda", generator=g)
    shift = torch.randn(c, device="cuda", generator=g)
    y16 = _nhwc16(y)
    yr = _nchw32(y16, nt, c, h, w).double().requires_grad_(True)
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    rm64, rv64 = rm.double(), rv.double()
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    ref = torch.cat([F.leaky_relu(F.batch_norm(yr[k * nb:(k + 1) * nb] + shift.double().view(1, c, 1, 1), rm64, rv64, gd, bd,
                                               training=True, momentum=0.1, eps=1e-5), 0.2) for k in range(groups)])
    npix = nb * h * w
    ws = torch.zeros(groups, 2, c, device="cuda", dtype=torch.float64)
    mean, rstd = torch.empty(groups, c, device="cuda"), torch.empty(groups, c, device="cuda")
    nbt = torch.zeros((), device="cuda", dtype=torch.long)
    ops.bn_reduce(y16, npix, c, ws, groups=groups, zero_first=True)
    a16 = torch.empty_like(y16)
    ops.bn_lrelu_fwd(y16, npix, c, mean, rstd, gamma, beta, a16, groups=groups, sums_ws=ws, running_mean=rm, running_var=rv,
                     num_batches=nbt, running_mean_shift=shift)
    assert rel_l2(_nchw32(a16, nt, c, h, w), ref) < 4e-4     # BatchNorm output is invariant to the constant shift
    assert rel_l2(rm, rm64) < 1e-5 and rel_l2(rv, rv64) < 1e-5 and int(nbt) == groups
    for k in range(groups):
        yk = yr[k * nb:(k + 1) * nb].detach()
        assert rel_l2(mean[k], yk.mean(dim=(0, 2, 3))) < 1e-5
        assert rel_l2(rstd[k], 1.0 / torch.sqrt(yk.var(dim=(0, 2, 3), unbiased=False) + 1e-5)) < 1e-5
    # eval-mode form: mean / rstd are inputs, same result
    a16b = torch.empty_like(y16)
    ops.bn_lrelu_fwd(y16, npix, c, mean, rstd, gamma, beta, a16b, groups=groups)
    assert torch.equal(a16b, a16)
    # backward: dz = dL/d(bn out) with lrelu' applied by the producer
    da = torch.randn(nt, c, h, w, device="cuda", generator=g)
    dz = da * torch.where(_nchw32(a16, nt, c, h, w) > 0, 1.0, 0.2)
    dz16 = _nhwc16(dz)
    dzr = _nchw32(dz16, nt, c, h, w).double()
    bn_out = torch.cat([F.batch_norm(yr[k * nb:(k + 1) * nb], None, None, gd, bd, training=True, eps=1e-5) for k in range(groups)])
    gy, gg, gb = torch.autograd.grad(bn_out, [yr, gd, bd], dzr)
    dy16 = torch.empty_like(y16)
    dgam, dbet = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    ops.bn_lrelu_bwd(dz16, y16, npix, c, mean, rstd, gamma, ws, dy16, dgam, dbet, groups=groups, zero_first=True)
    assert rel_l2(_nchw32(dy16, nt, c, h, w), gy) < 5e-4
    assert rel_l2(dgam, gg) < 1e-5 and rel_l2(dbet, gb) < 1e-5
    ops.bn_lrelu_bwd(dz16, y16, npix, c, mean, rstd, gamma, ws, dy16, dgam, dbet, groups=groups, zero_first=True, accumulate=True)
    assert rel_l2(dgam, 2 * gg) < 1e-5 and rel_l2(dbet, 2 * gb) < 1e-5


@pytest.mark.parametrize("nb,c,h,w", [(2, 64, 16, 24), (1, 128, 7, 9)])
def test_maxpool_forward_backward(nb, c, h, w):
    from pesr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(nb, c, h, w, device="cuda", generator=g).relu()      # many exact ties at 0, like VGG
    x16 = _nhwc16(x)
    xr = _nchw32(x16, nb, c, h, w).requires_grad_(True)
    ref = F.max_pool2d(xr, 2, 2)
    y16 = torch.empty(nb, h // 2, w // 2, c, device="cuda", dtype=torch.float16)
    ops.maxpool2_fwd(x16, nb, h, w, c, y16)
    assert torch.equal(_nchw32(y16, nb, c, h // 2, w // 2), ref.detach())
    dy = torch.randn_like(ref)
    dy16 = _nhwc16(dy)
    gx, = torch.autograd.grad(ref, xr, _nchw32(dy16, nb, c, h // 2, w // 2))
    dx16 = torch.full((nb, h, w, c), float("nan"), device="cuda", dtype=torch.float16)
    ops.maxpool2_bwd(x16, dy16, nb, h, w, c, dx16, relu_mask=True)
    assert torch.equal(_nchw32(dx16, nb, c, h, w), gx * (xr.detach() > 0))


@pytest.mark.parametrize("nb,k,o", [(16, 73728, 1024), (4, 4608, 1024), (16, 1024, 1), (3, 1024, 1)])
def test_skinny_linear(nb, k, o):
    """nn.Linear of model/pesr.py:71,73: forward (+bias, +LeakyReLU), dgrad, wgrad."""
    from pesr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(nb, k, device="cuda", generator=g)
    w = torch.randn(o, k, device="cuda", generator=g) / k ** 0.5
    b = torch.randn(o, device="cuda", generator=g)
    x16, w16 = x.half(), torch.empty(o, k, device="cuda", dtype=torch.float16)
    ops.cast16(w, w16)
    assert torch.equal(w16, w.half())
    xr, wr = x16.double(), w16.double()
    ref = F.leaky_relu(xr @ wr.t() + b.double(), 0.2)
    ws = torch.empty(ops.linear_workspace_floats(nb, k, o), device="cuda")
    out32 = torch.empty(nb, o, device="cuda")
    out16 = torch.empty(nb, o, device="cuda", dtype=torch.float16)
    ops.linear_fwd(x16, w16, b, nb, k, o, ws, out32=out32, out16=out16, act=ops.ACT_LRELU)
    assert rel_l2(out32, ref) < 1e-5 and rel_l2(out16.float(), ref) < 4e-4
    dy = torch.randn(nb, o, device="cuda", generator=g)
    dx = torch.empty(nb, k, device="cuda")
    ops.linear_dgrad(dy, w16, nb, k, o, dx)
    assert rel_l2(dx, dy.double() @ wr) < 1e-5
    dw = torch.empty(o, k, device="cuda")
    ops.linear_wgrad(dy, x16, nb, k, o, dw, mul=0.5)
    assert rel_l2(dw, 0.5 * dy.double().t() @ xr) < 1e-5
    ops.linear_wgrad(dy, x16, nb, k, o, dw, mul=0.5, accumulate=True)
    assert rel_l2(dw, dy.double().t() @ xr) < 1e-5


def test_flatten_roundtrip_matches_nchw_view():
    from pesr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(4)
    a = torch.randn(3, 512, 4, 5, device="cuda", generator=g)
    a16 = _nhwc16(a)
    flat = torch.empty(3, 512 * 20, device="cuda", dtype=torch.float16)
    ops.flatten_nchw16(a16, 3, 20, 512, flat)
    assert torch.equal(flat.float(), a.half().float().view(3, -1))          # model/pesr.py:79 `.view(N, -1)`
    d32 = torch.randn(3, 512 * 20, device="cuda", generator=g)
    back = torch.empty(3, 4, 5, 512, device="cuda", dtype=torch.float16)
    ops.unflatten_nchw16(d32, a16, 3, 20, 512, back, mul=2.0)
    ref = (2.0 * d32.view(3, 512, 4, 5)) * torch.where(a.half().float() > 0, 1.0, 0.2)
    assert rel_l2(_nchw32(back, 3, 512, 4, 5), ref) < 4e-4


def test_fused_losses_value_and_gradient():
    """train.py:131-140,213,251 + model/focal_loss.py against autograd on the oracle's formulas."""
    from oracle import pesr_oracle as O
    from pesr_b200 import losses
    from pesr_b200.model import FocalLoss
    g = torch.Generator(device="cuda").manual_seed(6)
    a = (torch.rand(2, 3, 20, 28, device="cuda", generator=g) * 255).requires_grad_(True)
    b = torch.rand(2, 3, 20, 28, device="cuda", generator=g) * 255
    for mine, ref_fn in ((losses.l1_loss, O.l1_loss), (losses.mse_loss, O.mse_loss)):
        a.grad = None
        (mine(a, b) * 3.0).backward()
        ad = a.detach().double().requires_grad_(True)
        ref = ref_fn(ad, b.double()) * 3.0
        gr, = torch.autograd.grad(ref, ad)
        assert abs(float(mine(a, b)) * 3.0 - float(ref)) < 1e-5 * abs(float(ref))
        assert rel_l2(a.grad, gr) < 1e-6
    a.grad = None
    (losses.tv_loss(a) * 1e-6).backward()
    ad = a.detach().double().requires_grad_(True)
    ref = O.tv_loss(ad) * 1e-6
    gr, = torch.autograd.grad(ref, ad)
    assert abs(float(losses.tv_loss(a)) * 1e-6 - float(ref)) < 1e-5 * float(ref)
    assert rel_l2(a.grad, gr) < 1e-6
    pr = (torch.randn(16, 1, device="cuda", generator=g) * 2).requires_grad_(True)
    pf = (torch.randn(16, 1, device="cuda", generator=g) * 2).requires_grad_(True)
    ones = torch.ones(16, 1, device="cuda")
    for gamma in (0.5, 1.0, 2.0):
        for detach in (False, True):
            pr.grad = pf.grad = None
            l = losses.rsgan_focal(pf, pr, gamma, ones, detach_weight=detach)
            l.backward()
            prd, pfd = pr.detach().double().requires_grad_(True), pf.detach().double().requires_grad_(True)
            ref = O.focal_loss(pfd - prd, ones.double(), gamma, detach_weight=detach)
            g1, g2 = torch.autograd.grad(ref, [pfd, prd])
            assert abs(float(l) - float(ref)) < 1e-5
            assert rel_l2(pf.grad, g1) < 1e-5 and rel_l2(pr.grad, g2) < 1e-5
    pr.grad = pf.grad = None
    l = losses.rsgan_bce(pr, pf, ones)
    l.backward()
    prd, pfd = pr.detach().double().requires_grad_(True), pf.detach().double().requires_grad_(True)
    ref = F.binary_cross_entropy_with_logits(prd - pfd, ones.double())
    g1, g2 = torch.autograd.grad(ref, [prd, pfd])
    assert abs(float(l) - float(ref)) < 1e-5 and rel_l2(pr.grad, g1) < 1e-5 and rel_l2(pf.grad, g2) < 1e-5
    # the module surface of model/focal_loss.py: FocalLoss(gamma)(x, t) on an autograd expression
    pr.grad = pf.grad = None
    FocalLoss(1)(pf - pr, ones).backward()
    prd, pfd = pr.detach().double().requires_grad_(True), pf.detach().double().requires_grad_(True)
    g1, = torch.autograd.grad(O.focal_loss(pfd - prd, ones.double(), 1), pfd)
    assert rel_l2(pf.grad, g1) < 1e-5


@pytest.mark.parametrize("nb,k,o", [(16, 73728, 1024), (4, 4608, 1024), (7, 8192, 1024)])
def test_linear_on_the_implicit_gemm_kernel(nb, k, o):
    """Linear(k -> o) forward as a split-K run of pesr_conv_igemm, and its backward-data with the [o][k] weight
    matrix consumed MN-major in place (model/pesr.py:71)."""
    from pesr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(8)
    x16 = torch.randn(nb, k, device="cuda", generator=g).half()
    w16 = (torch.randn(o, k, device="cuda", generator=g) / k ** 0.5).half()
    b = torch.randn(o, device="cuda", generator=g)
    ks = max(1, min(k // 64 // 8, 36))
    part = torch.empty(ks * nb * o, device="cuda")
    ops.conv_igemm(ops.make_conv_desc(dtype=0, nb=1, h=1, w=nb, cin=k, cout=o, block_n=256, taps=[(0, 0)],
                                      srcs=[ops.nhwc_src(x16, 1, 1, nb, k)], wpacked=w16, out32=part, ld_out32=o,
                                      ksplit=ks, split_stride32=nb * o))
    out32 = torch.empty(nb, o, device="cuda")
    out16 = torch.empty(nb, o, device="cuda", dtype=torch.float16)
    ops.linear_finalize(part, ks, nb, o, b, torch.float16, out32=out32, out16=out16, act=ops.ACT_LRELU)
    ref = F.leaky_relu(x16.double() @ w16.double().t() + b.double(), 0.2)
    assert rel_l2(out32, ref) < 1e-5 and rel_l2(out16.float(), ref) < 4e-4
    dy16 = torch.randn(nb, o, device="cuda", generator=g).half()
    dx = torch.empty(nb, k, device="cuda")
    ops.conv_igemm(ops.make_conv_desc(dtype=0, nb=1, h=1, w=nb, cin=o, cout=k, block_n=256, taps=[(0, 0)],
                                      srcs=[ops.nhwc_src(dy16, 1, 1, nb, o)], wpacked=w16, out32=dx, ld_out32=k,
                                      b_mn_major=1))
    assert rel_l2(dx, dy16.double() @ w16.double()) < 1e-5


@pytest.mark.parametrize("rows,kfc,world", [(32, 512 * 9, 1), (8, 512 * 16, 1), (96, 512 * 9, 3), (32, 73728, 1)])
def test_linear_weight_gradient_on_the_tensor_cores(rows, kfc, world):
    """DiscriminatorEngine._fc1_wgrad: dW = dz1^T x flat7 / world through the split-K weight-gradient kernel (rows as
    pixels, dz1 as a 16-bit hi + lo pair stacked along K, scale folded into the store) against fp64; `world` > 1 is the
    data-parallel form (the gathered factors of all ranks)."""
    from pesr_b200.engine_d import DiscriminatorEngine
    g = torch.Generator(device="cuda").manual_seed(11)
    eng = DiscriminatorEngine(None)
    dz1 = torch.randn(rows, 1024, device="cuda", generator=g) * 3e-6          # gradients far below the fp16 normal range
    flat7 = torch.randn(rows, kfc, device="cuda", generator=g).half()
    grad = torch.full((1024, kfc), float("nan"), device="cuda")
    eng._fc1_wgrad(dz1, flat7, kfc, grad, world=world)
    ref = dz1.double().t() @ flat7.double() / world
    e = rel_l2(grad, ref)
    print(f"fc1 wgrad on tensor cores rows={rows} kfc={kfc}: rel-L2 {e:.2e}")
    assert e < 2e-6
