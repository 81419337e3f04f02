"""GPU tests of the edges around the hot path: uint8 image I/O fused into the Generator's first
and last kernels (utils.py:13-25), batched / tiled inference (test.py:101-112), the device-side Y-channel PSNR
(utils.py:10-11,27-41) and the one-launch training-patch gather (data.py:64-126)."""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _gen(opt, seed):
    from oracle import pesr_oracle as O
    from pesr_b200.model import Generator
    sd = O.init_generator(opt, seed)
    G = Generator(opt)
    G.load_state_dict(sd)
    return G.cuda().eval(), sd


@pytest.mark.parametrize("sgn,with_affine", [(1, True), (-1, False)])
def test_col2im3_tiled_matches_per_pixel_kernel(sgn, with_affine):
    """The shared-memory col2im (io_ops.cu) against the round-1 per-pixel kernel and torch, ragged sizes, fused uint8."""
    from pesr_b200._lib import check, lib
    g = torch.Generator().manual_seed(0)
    nb, h, w = 3, 21, 45
    z = torch.randn(nb * h * w, 32, generator=g).cuda() * 40
    bias = torch.randn(3, generator=g).cuda()
    A = (torch.eye(3) + 0.1 * torch.randn(3, 3, generator=g)).cuda().contiguous() if with_affine else None
    B = (torch.randn(3, generator=g) * 100 + 100).cuda() if with_affine else None
    s = torch.cuda.current_stream().cuda_stream
    ptr = lambda t: 0 if t is None else t.data_ptr()   # noqa: E731
    ref, pre_ref = torch.empty(nb, 3, h, w, device="cuda"), torch.empty(nb, 3, h, w, device="cuda")
    check(lib.pesr_col2im3(z.data_ptr(), 32, nb, h, w, bias.data_ptr(), ptr(A), ptr(B), 0.5, 0, sgn, pre_ref.data_ptr(),
                           ref.data_ptr(), s))
    out, pre = torch.empty_like(ref), torch.empty_like(ref)
    out8 = torch.empty(nb, h, w, 3, device="cuda", dtype=torch.uint8)
    check(lib.pesr_col2im3_tiled(z.data_ptr(), 32, nb, h, w, bias.data_ptr(), ptr(A), ptr(B), 0.5, 0, sgn, pre.data_ptr(),
                                 out.data_ptr(), out8.data_ptr(), s))
    assert torch.allclose(out, ref, rtol=1e-6, atol=1e-4) and torch.allclose(pre, pre_ref, rtol=1e-6, atol=1e-4)
    want8 = out.clamp(0, 255).round().permute(0, 2, 3, 1).to(torch.uint8)      # torch.round is half-to-even like numpy
    assert torch.equal(out8, want8)
    # against torch: z[p + sgn*(ky-1,kx-1)][tap*3+c] summed over taps
    zz = z.view(nb, h, w, 32)[..., :27].reshape(nb, h, w, 9, 3).permute(0, 3, 4, 1, 2).double()   # [n][tap][c][h][w]
    acc = torch.zeros(nb, 3, h, w, dtype=torch.float64, device="cuda")
    padded = torch.nn.functional.pad(zz, (1, 1, 1, 1))
    for tap in range(9):
        dy, dx = sgn * (tap // 3 - 1), sgn * (tap % 3 - 1)
        acc += padded[:, tap, :, 1 + dy:1 + dy + h, 1 + dx:1 + dx + w]
    assert rel_l2(pre, acc * 0.5 + bias.double().view(1, 3, 1, 1)) < 1e-6


def test_uint8_in_uint8_out_inference_equals_the_fp32_path():
    from pesr_b200 import infer
    G, _ = _gen({'depth': 2, 'num_channels': 64, 'res_scale': 0.1}, 1)
    g = torch.Generator().manual_seed(2)
    imgs = (torch.rand(3, 19, 27, 3, generator=g) * 255).to(torch.uint8).cuda()
    out8 = infer.super_resolve_u8(G, imgs)
    x = infer.imgs_to_tensor(imgs)
    assert torch.equal(x, imgs.permute(0, 3, 1, 2).float())
    out32, ref8 = infer.super_resolve(G, x)
    assert out8.shape == (3, 76, 108, 3) and torch.equal(out8, ref8)
    assert torch.equal(ref8, out32.clamp(0, 255).round().permute(0, 2, 3, 1).to(torch.uint8))
    one = infer.super_resolve_u8(G, imgs[1])
    assert torch.equal(one, out8[1])


def test_batched_and_chunked_inference_equal_single_images(monkeypatch):
    from pesr_b200 import engine_g, infer
    G, _ = _gen({'depth': 2, 'num_channels': 64, 'res_scale': 0.1}, 1)
    Gp, _ = _gen({'depth': 2, 'num_channels': 64, 'res_scale': 0.1}, 2)
    g = torch.Generator().manual_seed(3)
    x = (torch.rand(5, 3, 14, 22, generator=g) * 255).cuda()
    singles = [infer.super_resolve(G, x[i:i + 1], alpha=0.5, model_psnr=Gp) for i in range(5)]
    monkeypatch.setattr(engine_g, "INFER_CHUNK_PIXELS", 2 * 14 * 22)        # forces chunks of 2, 2, 1 images
    out32, out8 = infer.super_resolve(G, x, alpha=0.5, model_psnr=Gp)
    assert out32.shape == (5, 3, 56, 88) and out8.shape == (5, 56, 88, 3)
    for i, (s32, s8) in enumerate(singles):
        assert torch.equal(out32[i], s32[0]) and torch.equal(out8[i], s8)


def test_tiled_inference_matches_whole_image():
    """Halo-overlapped crops (pesr_b200.infer, tile=): exact up to the fp32 summation order of the kernel variants that
    different crop sizes select."""
    from pesr_b200 import infer
    G, _ = _gen({'depth': 2, 'num_channels': 64, 'res_scale': 0.1}, 1)
    g = torch.Generator().manual_seed(4)
    x = (torch.rand(1, 3, 37, 50, generator=g) * 255).cuda()
    whole, _ = infer.super_resolve(G, x, return_u8=False)
    tiled, _ = infer.super_resolve(G, x, return_u8=False, tile=(16, 24))
    assert infer.receptive_halo(G) == 8
    assert rel_l2(tiled, whole) < 1e-4
    assert float((tiled - whole).abs().max()) < 0.5


def test_psnr_y_matches_numpy_reference_formula():
    """utils.compute_PSNR (utils.py:27-41) restated in numpy; every step is integer-valued after rounding, so the sums
    must agree exactly except on exact rounding ties of rgb2y (65 of 2^24 colours, see io_ops.cu)."""
    from pesr_b200.utils import compute_PSNR, compute_PSNR_sse

    def rgb2y(rgb):
        return np.dot(rgb[..., :3], [65.738 / 256, 129.057 / 256, 25.064 / 256]) + 16

    def sse_ref(out, lbl):
        res = []
        for o, l in zip(out, lbl):
            o = o.numpy().clip(0, 255).round().transpose(1, 2, 0).astype(np.uint8)
            l = l.numpy().clip(0, 255).round().transpose(1, 2, 0).astype(np.uint8)
            d = rgb2y(o).clip(0, 255).round() - rgb2y(l).clip(0, 255).round()
            res.append(float((d ** 2).sum()))
        return res
    g = torch.Generator().manual_seed(5)
    a = torch.rand(3, 3, 33, 47, generator=g) * 300 - 20          # exercises the clipping
    b = (a + torch.randn(3, 3, 33, 47, generator=g) * 6).contiguous()
    sse = compute_PSNR_sse(a.cuda(), b.cuda()).cpu().tolist()
    ref = sse_ref(a, b)
    for got, want in zip(sse, ref):
        assert abs(got - want) <= 64, (got, want)                 # a tie colour moves one pixel's Y by 1: d^2 changes by 2|d|+1
    psnr = float(compute_PSNR(a[:1].cuda(), b[:1].cuda()))
    rmse = np.sqrt(ref[0] / (33 * 47))
    assert abs(psnr - 20 * np.log10(255 / rmse)) < 1e-3
    # exhaustive over all colours of a coarse lattice + every exact tie: the kernel's integer rgb2y
    r, gg, bb = np.meshgrid(np.arange(0, 256, 5.), np.arange(0, 256, 3.), np.arange(0, 256, 7.), indexing='ij')
    cols = np.stack([r.ravel(), gg.ravel(), bb.ravel()], 0)
    n = cols.shape[1]
    img = torch.from_numpy(cols.reshape(1, 3, 1, n)).float()
    zero = torch.zeros_like(img)
    got = int(compute_PSNR_sse(img.cuda(), zero.cuda())[0])
    y = rgb2y(cols.T).clip(0, 255).round()
    want = float(((y - 16) ** 2).sum())
    ties = int((((65738 * cols[0] + 129057 * cols[1] + 25064 * cols[2]) % 256000) == 128000).sum())
    assert abs(got - want) <= 600 * max(ties, 1) and abs(got - want) / want < 1e-6


def test_patch_gather_matches_data_py_semantics():
    """PatchSource.batch (one gather launch) against data.py:64-126 restated in numpy: crop, then transpose (bit 2),
    vertical flip (bit 1), horizontal flip (bit 0), HWC -> CHW float."""
    from pesr_b200.data import PatchSource
    rng = np.random.RandomState(0)
    scale, p, b = 4, 6, 16
    imgs = []
    for (h, w) in [(9, 14), (20, 6), (11, 11)]:
        lr = rng.randint(0, 256, size=(h, w, 3), dtype=np.uint8)
        hr = rng.randint(0, 256, size=(h * scale, w * scale, 3), dtype=np.uint8)
        imgs.append((lr, hr))
    src = PatchSource([(torch.from_numpy(l).cuda(), torch.from_numpy(h).cuda()) for l, h in imgs], p, scale, b, torch.device("cuda"))
    for rep in range(3):
        choices = src.draw()
        if rep == 0:
            choices = [(i % 3, min(c[1], imgs[i % 3][0].shape[0] - p), min(c[2], imgs[i % 3][0].shape[1] - p), i % 8)
                       for i, c in enumerate(choices)]          # every augmentation index at least once
        lr_t, hr_t = src.batch(choices)
        assert lr_t.shape == (b, 3, p, p) and hr_t.shape == (b, 3, p * scale, p * scale)
        for n, (i, y, x, k) in enumerate(choices):
            inp = imgs[i][0][y:y + p, x:x + p, :]
            lbl = imgs[i][1][y * scale:(y + p) * scale, x * scale:(x + p) * scale, :]
            if (k >> 2) & 1:
                inp, lbl = inp.transpose((1, 0, 2)), lbl.transpose((1, 0, 2))
            if (k >> 1) & 1:
                inp, lbl = inp[::-1, :, :], lbl[::-1, :, :]
            if k & 1:
                inp, lbl = inp[:, ::-1, :], lbl[:, ::-1, :]
            assert np.array_equal(lr_t[n].cpu().numpy(), inp.transpose(2, 0, 1).astype(np.float32)), (rep, n, k)
            assert np.array_equal(hr_t[n].cpu().numpy(), lbl.transpose(2, 0, 1).astype(np.float32)), (rep, n, k)
