"""Device-side counterparts of the reference's utils.py helpers used around the hot path (utils.py:10-41)."""
import torch

from . import ops


def compute_PSNR_sse(out, lbl, sse=None):
    """Accumulates the exact integer sum of squared Y-channel differences of utils.compute_PSNR (utils.py:27-41) per
    image into `sse` (int64 [N], created when None): rgb -> clip/round -> rgb2y -> clip/round -> diff^2, on the
    device, no host round trip."""
    if out.shape != lbl.shape or out.dim() != 4 or out.shape[1] != 3:
        raise ValueError("compute_PSNR expects two [N,3,H,W] tensors of one shape")
    out, lbl = out.detach().contiguous().float(), lbl.detach().contiguous().float()
    if sse is None:
        sse = torch.zeros(out.shape[0], device=out.device, dtype=torch.int64)
    return ops.psnr_y_sse(out, lbl, sse)


def psnr_from_sse(sse, pixels):
    """20 log10(255 / rmse) per image (float64 device tensor; inf where the images agree exactly)."""
    mse = sse.double() / float(pixels)
    return 20.0 * torch.log10(255.0 / torch.sqrt(mse))


def compute_PSNR(out, lbl):
    """utils.compute_PSNR (utils.py:27-41) as a DEVICE tensor: mean over the batch of the per-image Y-channel PSNR
    (the reference calls it with one image at a time).  Reading it back is the caller's choice (train.py reads one
    value per epoch)."""
    sse = compute_PSNR_sse(out, lbl)
    return psnr_from_sse(sse, out.shape[2] * out.shape[3]).mean()


class PSNRMeter:
    """Validation accumulator for train.py:281-295: one kernel per validation image, one device->host read per epoch."""

    def __init__(self, device):
        self.total = torch.zeros((), device=device, dtype=torch.float64)
        self.count = 0

    def update(self, out, lbl):
        sse = compute_PSNR_sse(out, lbl)
        self.total += psnr_from_sse(sse, out.shape[2] * out.shape[3]).sum()
        self.count += out.shape[0]

    def value(self):
        return float(self.total.item()) / max(self.count, 1)
