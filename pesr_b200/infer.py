"""Inference path of test.py:101-112 on the B200 kernels: Generator forward, optional x8 self-ensemble of a
second (PSNR) generator, alpha blend, clip/round/uint8 -- one fused kernel for everything after the forwards."""
import torch

from . import ops
from ._lib import check, lib


def imgs_to_tensor(img_u8):
    """utils.imgs_to_tensors (utils.py:20-25): HWC uint8 (numpy array or tensor) -> [1,3,H,W] fp32 CUDA tensor."""
    t = torch.as_tensor(img_u8)
    if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
        raise ValueError("expected an HWC uint8 RGB image")
    t = t.cuda().contiguous()
    h, w = t.shape[0], t.shape[1]
    out = torch.empty(1, 3, h, w, device=t.device, dtype=torch.float32)
    check(lib.pesr_u8hwc_to_f32nchw(t.data_ptr(), h, w, out.data_ptr(), torch.cuda.current_stream().cuda_stream),
          "pesr_u8hwc_to_f32nchw")
    return out


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


def _ensemble(model, img):
    """The 8 generator outputs of test.py:57, back to back (transposed variants stay transposed)."""
    first = model(_transform(img, 0))
    _, _, H, W = first.shape
    ens = torch.empty(8, 3, H * W, device=img.device, dtype=torch.float32)
    ens[0].copy_(first.reshape(3, H * W))
    for i in range(1, 8):
        ens[i].copy_(model(_transform(img, i)).reshape(3, H * W))
    return ens, H, W


def _blend(perc, ens, H, W, alpha, want_u8):
    out32 = torch.empty(1, 3, H, W, device=perc.device, dtype=torch.float32)
    out8 = torch.empty(H, W, 3, device=perc.device, dtype=torch.uint8) if want_u8 else None
    check(lib.pesr_blend_x8_to_u8(perc.data_ptr(), ops._ptr(ens), H, W, float(alpha), 0 if ens is None else 8,
                                  out32.data_ptr(), ops._ptr(out8), torch.cuda.current_stream().cuda_stream),
          "pesr_blend_x8_to_u8")
    return out32, out8


@torch.no_grad()
def super_resolve(model, img, alpha=1.0, model_psnr=None, return_u8=True):
    """test.py:106-112 for one [1,3,h,w] image: returns (out fp32 [1,3,4h,4w], uint8 HWC image or None)."""
    if img.dim() != 4 or img.shape[0] != 1:
        raise ValueError("super_resolve takes one image [1,3,h,w] (test.py processes one image at a time)")
    perc = model(img).contiguous().float()
    _, _, H, W = perc.shape
    ens = None
    if alpha != 1:
        if model_psnr is None:
            raise ValueError("alpha != 1 needs the PSNR model (test.py:90-92)")
        ens, He, We = _ensemble(model_psnr, img)
        if (He, We) != (H, W):
            raise ValueError("perceptual and PSNR models disagree on the output size")
    return _blend(perc, ens, H, W, alpha, return_u8)


@torch.no_grad()
def x8_forward(img, model):
    """test.py:45-74 as a tensor function (mean of the 8 inverse-transformed outputs)."""
    ens, H, W = _ensemble(model, img)
    zero = torch.zeros(1, 3, H, W, device=img.device, dtype=torch.float32)
    return _blend(zero, ens, H, W, 0.0, False)[0]
