"""Fused loss kernels behind the reference's loss call sites (train.py:131-140, 213, 251;
model/focal_loss.py).  Every loss computes its value and its input gradient in one pass over HBM; the
autograd wrappers below only multiply that saved gradient by the incoming (scalar) grad_output."""
import torch
import torch.nn as nn

from . import ops


def _scalar(dev):
    return torch.empty((), device=dev, dtype=torch.float32)


class _DiffLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, kind):
        a = a.contiguous().float()
        b = b.contiguous().float()
        if a.shape != b.shape:
            raise ValueError(f"loss: shape mismatch {tuple(a.shape)} vs {tuple(b.shape)}")
        loss = _scalar(a.device)
        need_a, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        grad = torch.empty_like(a) if (need_a or need_b) else None
        (ops.loss_l1 if kind == 0 else ops.loss_mse)(a, b, loss, grad)
        ctx.grad = grad
        return loss

    @staticmethod
    def backward(ctx, go):
        g = ctx.grad * go if ctx.grad is not None else None
        return (g if ctx.needs_input_grad[0] else None, (-g) if ctx.needs_input_grad[1] else None, None)


def l1_loss(a, b):
    """nn.L1Loss() (train.py:131): mean |a - b|."""
    return _DiffLoss.apply(a, b, 0)


def mse_loss(a, b):
    """F.mse_loss (train.py:136): mean (a - b)^2."""
    return _DiffLoss.apply(a, b, 1)


class _TVLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y):
        y = y.contiguous().float()
        loss = _scalar(y.device)
        grad = torch.empty_like(y) if ctx.needs_input_grad[0] else None
        ops.loss_tv(y, loss, grad)
        ctx.grad = grad
        return loss

    @staticmethod
    def backward(ctx, go):
        return ctx.grad * go


def tv_loss(y):
    """train.py:137-140: SUM of absolute horizontal and vertical neighbour differences."""
    return _TVLoss.apply(y)


class _GanLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, sign_a, sign_b, target, mode, gamma):
        a32 = a.contiguous().float()
        b32 = b.contiguous().float() if b is not None else None
        loss = _scalar(a.device)
        ga = torch.empty_like(a32) if ctx.needs_input_grad[0] else None
        gb = torch.empty_like(b32) if (b is not None and ctx.needs_input_grad[1]) else None
        ops.loss_gan(a32, b32, loss, sign_a=sign_a, sign_b=sign_b, target=target, mode=mode, gamma=gamma, grad_a=ga,
                     grad_b=gb)
        ctx.ga, ctx.gb = ga, gb
        return loss

    @staticmethod
    def backward(ctx, go):
        return (ctx.ga * go if ctx.ga is not None else None, ctx.gb * go if ctx.gb is not None else None,
                None, None, None, None, None)


_TARGET_CACHE = {}


def _uniform_target(t):
    """The reference's targets are constant tensors (train.py:148-149); the kernel takes the constant.
    The value is read back once per (tensor, version) so the hot loop has no device->host sync."""
    if not torch.is_tensor(t):
        return float(t)
    key = (t.data_ptr(), t._version, t.numel())
    hit = _TARGET_CACHE.get(key)
    if hit is None:
        lo, hi = float(t.min()), float(t.max())
        if lo != hi:
            raise ValueError("pesr_b200 GAN losses take a constant target tensor (all zeros or all ones)")
        if len(_TARGET_CACHE) > 64:
            _TARGET_CACHE.clear()
        # the entry keeps the tensor alive: a freed tensor's address could otherwise come back holding another value
        hit = _TARGET_CACHE[key] = (lo, t)
    return hit[0]


def rsgan_bce(pred_a, pred_b, target=1.0):
    """bce_loss_fn(pred_a - pred_b, target) (train.py:213 / :253) in one launch."""
    return _GanLoss.apply(pred_a, pred_b, 1.0, -1.0, _uniform_target(target), 0, 1.0)


def rsgan_focal(pred_a, pred_b, gamma, target=1.0, detach_weight=False):
    """f_loss_fn(pred_a - pred_b, target) (train.py:251) in one launch."""
    return _GanLoss.apply(pred_a, pred_b, 1.0, -1.0, _uniform_target(target), 2 if detach_weight else 1, float(gamma))


class BCEWithLogitsLoss(nn.Module):
    """nn.BCEWithLogitsLoss() for a constant target (train.py:132)."""

    def forward(self, x, t):
        return _GanLoss.apply(x, None, 1.0, 0.0, _uniform_target(t), 0, 1.0)


class L1Loss(nn.Module):
    def forward(self, a, b):
        return l1_loss(a, b)
