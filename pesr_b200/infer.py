"""Inference path of test.py:101-112 on the B200 kernels: Generator forward, optional x8 self-ensemble of a
second (PSNR) generator, alpha blend, clip/round/uint8 -- one fused kernel for everything after the forwards, and for
alpha = 1 the uint8 HWC image goes in and comes out of the Generator's own first and last kernels.

Images are independent (test.py processes them one at a time); batches of equally sized images are accepted and run
through one plan in bounded-memory chunks.  `tile=` bounds the activation memory of very large images by running the
Generator on overlapping crops whose halo covers its receptive field (exact up to fp32 summation order)."""
import torch

from . import ops
from ._lib import check, lib


def imgs_to_tensor(img_u8):
    """utils.imgs_to_tensors (utils.py:20-25): HWC uint8 image (numpy array or tensor, [H,W,3] or a batch [N,H,W,3])
    -> [N,3,H,W] fp32 CUDA tensor."""
    t = torch.as_tensor(img_u8)
    if t.dim() == 3:
        t = t.unsqueeze(0)
    if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[3] != 3:
        raise ValueError("expected HWC uint8 RGB image(s)")
    t = t.cuda().contiguous()
    out = torch.empty(t.shape[0], 3, t.shape[1], t.shape[2], device=t.device, dtype=torch.float32)
    return ops.u8hwc_to_f32nchw(t, out)


def _transform(img, i):
    """Variant i of test.py:45-58: bit0 flips W ('vflip'), bit1 flips H ('hflip'), bit2 transposes; pure data
    movement on the 3-channel LR image (the reference round-trips through numpy on the CPU here)."""
    t = img
    if i & 1:
        t = t.flip(3)
    if i & 2:
        t = t.flip(2)
    if i & 4:
        t = t.transpose(2, 3)
    return t.contiguous()


def receptive_halo(model):
    """LR pixels of context one output pixel depends on in each direction: embed (1) + 2 per ResBlock + tail conv (1) +
    upsample.0 (1) + upsample.2 (1/2) + upsample.4 (1/4), model/pesr.py:28-38."""
    return 2 * model.n_resblock + 4


def _forward(model, img, tile=None):
    """Generator forward of [N,3,h,w], whole-image or on halo-overlapped crops of at most tile = (th, tw) LR pixels."""
    if tile is None:
        return model(img)
    th, tw = tile
    n, _, h, w = img.shape
    halo = receptive_halo(model)
    out = torch.empty(n, 3, 4 * h, 4 * w, device=img.device, dtype=torch.float32)
    for y0 in range(0, h, th):
        for x0 in range(0, w, tw):
            y1, x1 = min(h, y0 + th), min(w, x0 + tw)
            ya, xa, yb, xb = max(0, y0 - halo), max(0, x0 - halo), min(h, y1 + halo), min(w, x1 + halo)
            sr = model(img[:, :, ya:yb, xa:xb].contiguous())
            out[:, :, 4 * y0:4 * y1, 4 * x0:4 * x1] = sr[:, :, 4 * (y0 - ya):4 * (y1 - ya), 4 * (x0 - xa):4 * (x1 - xa)]
    return out


def _ensemble(model, img, tile=None):
    """The 8 generator outputs of test.py:57 for ONE image, back to back (transposed variants stay transposed)."""
    first = _forward(model, _transform(img, 0), tile)
    _, _, H, W = first.shape
    ens = torch.empty(8, 3, H * W, device=img.device, dtype=torch.float32)
    ens[0].copy_(first.reshape(3, H * W))
    for i in range(1, 8):
        t = (tile[1], tile[0]) if (tile is not None and i & 4) else tile
        ens[i].copy_(_forward(model, _transform(img, i), t).reshape(3, H * W))
    return ens, H, W


def _blend(perc, ens, H, W, alpha, out32, out8):
    check(lib.pesr_blend_x8_to_u8(perc.data_ptr(), ops._ptr(ens), H, W, float(alpha), 0 if ens is None else 8,
                                  ops._ptr(out32), ops._ptr(out8), torch.cuda.current_stream().cuda_stream),
          "pesr_blend_x8_to_u8")


@torch.no_grad()
def super_resolve(model, img, alpha=1.0, model_psnr=None, return_u8=True, tile=None):
    """test.py:106-112 for a batch [N,3,h,w] of equally sized images (fp32, 0..255): returns
    (out fp32 [N,3,4h,4w], uint8 HWC images [N,4h,4w,3] -- [4h,4w,3] for N = 1 -- or None)."""
    if img.dim() != 4 or img.shape[1] != 3:
        raise ValueError("super_resolve takes images as [N,3,h,w]")
    perc = _forward(model, img, tile).contiguous().float()
    n, _, H, W = perc.shape
    if alpha != 1 and model_psnr is None:
        raise ValueError("alpha != 1 needs the PSNR model (test.py:90-92)")
    out32 = torch.empty(n, 3, H, W, device=perc.device, dtype=torch.float32)
    out8 = torch.empty(n, H, W, 3, device=perc.device, dtype=torch.uint8) if return_u8 else None
    for k in range(n):
        ens = None
        if alpha != 1:
            ens, He, We = _ensemble(model_psnr, img[k:k + 1], tile)
            if (He, We) != (H, W):
                raise ValueError("perceptual and PSNR models disagree on the output size")
        _blend(perc[k], ens, H, W, alpha, out32[k], out8[k] if out8 is not None else None)
    if out8 is not None and n == 1:
        out8 = out8[0]
    return out32, out8


@torch.no_grad()
def super_resolve_u8(model, img_u8):
    """alpha = 1 inference from uint8 HWC image(s) to uint8 HWC image(s) (utils.imgs_to_tensors -> G ->
    utils.tensors_to_imgs, test.py:103-114) with both conversions fused into the Generator's first and last kernels:
    no fp32 image is ever written.  img_u8: [h,w,3] or [N,h,w,3] CUDA uint8 tensor."""
    single = img_u8.dim() == 3
    x = img_u8.unsqueeze(0) if single else img_u8
    if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[3] != 3 or not x.is_cuda:
        raise ValueError("super_resolve_u8 takes CUDA uint8 HWC image(s)")
    was_training = model.training
    out, _ = model.engine().forward(x, train=False, out_u8=True)
    model.train(was_training)
    return out[0] if single else out


@torch.no_grad()
def x8_forward(img, model):
    """test.py:45-74 as a tensor function (mean of the 8 inverse-transformed outputs)."""
    ens, H, W = _ensemble(model, img)
    zero = torch.zeros(1, 3, H, W, device=img.device, dtype=torch.float32)
    out32 = torch.empty(1, 3, H, W, device=img.device, dtype=torch.float32)
    _blend(zero, ens, H, W, 0.0, out32, None)
    return out32
