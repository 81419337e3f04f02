"""The reference's step bodies (train.py:164-176 pretrain, train.py:202-266 GAN) over the B200 modules.

These are what train.py's loops call once per batch; bench.py times exactly these functions.
Losses stay on the device: callers read them back when they want to log (train.py reads five scalars
per iteration, train.py:262-266 -- here that is one optional stacked copy)."""
import os

import torch

from . import losses

_D_PAIR = os.environ.get("PESR_NO_D_PAIR") != "1"    # A/B knob (tools/ab_env.sh)
# VGG branch on a second stream: OFF by default.  Measured (tools/gpu_r2_d.sh / gpu_r2_e.sh): 17.10 vs 17.19 ms per GAN
# step; the tensor-core kernels claim all 227 KB of an SM's shared memory, so the Discriminator's memory-bound kernels
# cannot be co-resident with them, and with a 176 KB budget the convolutions lose 0.45 ms while the overlap returns 0.19
# (and the SM clock drops under the 1 kW power cap).  PESR_TWO_STREAMS=1 enables it.
_TWO_STREAMS = os.environ.get("PESR_TWO_STREAMS") == "1"
_SIDE_STREAMS = {}


def _side_stream(device):
    s = _SIDE_STREAMS.get(device)
    if s is None:
        s = _SIDE_STREAMS[device] = torch.cuda.Stream(device=device)
    return s


def pretrain_step(G, optim_G, lr, hr, ddp=None):
    """train.py:168-173: sr = G(lr); zero_grad; L1; backward; Adam step.  Returns the loss tensor."""
    sr = G(lr)
    optim_G.zero_grad(set_to_none=True)
    loss = losses.l1_loss(sr, hr)
    loss.backward()
    if ddp is not None:
        ddp.finish()
    optim_G.step()
    return loss


def gan_step(G, D, vgg, optim_G, optim_D, lr, hr, cfg, ddp_g=None, ddp_d=None):
    """train.py:202-259 with the defaults of train.py:64-76 (RSGAN + focal loss).  ``cfg`` carries
    alpha_l1, alpha_vgg, alpha_gan, alpha_tv, fl_gamma, gan_type, focal_loss, target_real, target_fake.
    Returns the five scalars of train.py:262-266 as device tensors."""
    t_real, t_fake = cfg['target_real'], cfg['target_fake']
    # ---- Discriminator phase (train.py:202-229)
    for p in D.parameters():
        p.requires_grad = True
    optim_D.zero_grad(set_to_none=True)
    sr = G(lr)
    # The perceptual branch (VGG forward of sr / hr and, through autograd, its backward) depends on sr and hr only.  It is
    # issued on a second stream: its tensor-core-bound convolutions then overlap the Discriminator's BatchNorm /
    # element-wise / small-layer kernels, which are HBM- and latency-bound and leave the tensor cores idle (a conv CTA and
    # the blocks of those kernels are co-resident on an SM).  Autograd runs each backward node on the stream of its forward
    # and orders the streams where gradients meet, so the VGG backward overlaps the Discriminator's in the same way.
    two = _TWO_STREAMS and sr.is_cuda
    if two:
        main, side = torch.cuda.current_stream(), _side_stream(sr.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            f_sr, f_hr = vgg(sr, hr)
            vgg_loss = losses.mse_loss(f_sr, f_hr) * cfg['alpha_vgg']
    if _D_PAIR and hasattr(D, "forward_pair"):
        # D(hr) then D(sr.detach()) as in train.py:205-207 (same order of the BatchNorm running-statistics updates),
        # executed as one batch of 2N images with one gradient pass / one all-reduce for the parameters
        pred_real, pred_fake = D.forward_pair(hr, sr.detach())
    else:
        pred_real = D(hr)
        pred_fake = D(sr.detach())
    if cfg['gan_type'] == 'SGAN':
        bce = losses.BCEWithLogitsLoss()
        total_D_loss = bce(pred_real, t_real) + bce(pred_fake, t_fake)
    else:
        total_D_loss = losses.rsgan_bce(pred_real, pred_fake, t_real)
    if cfg.get('GP'):
        # train.py:216-226 (`--GP true`, off by default): the penalty needs a second derivative through D, which the kernel
        # schedule does not implement; pesr_b200/gp.py evaluates it with ATen on the same parameters.  Its parameter
        # gradients join the schedule's before anything is reduced across ranks.
        from .gp import gradient_penalty
        d_mod = getattr(D, "module", D)
        eng = d_mod.engine()
        params = list(d_mod.parameters())
        hook, eng.grad_hook = eng.grad_hook, None
        try:
            gp = gradient_penalty(d_mod, hr, sr, u=cfg.get('gp_u'))
            gp_grads = torch.autograd.grad(gp, params, allow_unused=True)
            total_D_loss.backward()
        finally:
            eng.grad_hook = hook
        with torch.no_grad():
            for p, g in zip(params, gp_grads):
                if g is not None:
                    p.grad.add_(g)
        total_D_loss = total_D_loss.detach() + gp.detach()
        if ddp_d is not None:
            ddp_d.allreduce_grads()
    else:
        total_D_loss.backward()
    # The Generator-phase terms that do not involve D (train.py:240-247) are issued BEFORE waiting for D's gradient
    # all-reduce (DataParallel(defer_finish=True): backward returns with the last buckets still in flight and
    # ddp_d.finish() below is the wait): VGG's forward, L1 and TV depend on sr / hr only, so they keep the SMs busy
    # while the 321 MB of Discriminator gradients are on the wire.  Values are unchanged (nothing here reads D or its
    # optimiser).
    optim_G.zero_grad(set_to_none=True)
    l1_loss = losses.l1_loss(sr, hr) * cfg['alpha_l1']
    if not two:
        f_sr, f_hr = vgg(sr, hr)
        vgg_loss = losses.mse_loss(f_sr, f_hr) * cfg['alpha_vgg']
    tv_scale = cfg['alpha_tv'] * (ddp_g.world_size if ddp_g is not None else 1)  # TV is a batch SUM (train.py:137-140)
    tv_loss = losses.tv_loss(sr) * tv_scale
    if ddp_d is not None:
        ddp_d.finish()
    optim_D.step()
    # ---- Generator phase (train.py:234-259)
    for p in D.parameters():
        p.requires_grad = False
    if _D_PAIR and hasattr(D, "forward_pair"):
        pred_fake, pred_real = D.forward_pair(sr, hr)       # D(sr) then D(hr), train.py:237-238, as one batch of 2N
    else:
        pred_fake = D(sr)
        pred_real = D(hr)
    if cfg['gan_type'] == 'SGAN':
        if cfg['focal_loss']:
            G_loss = losses._GanLoss.apply(pred_fake, None, 1.0, 0.0, losses._uniform_target(t_real), 1,
                                           float(cfg['fl_gamma']))
        else:
            G_loss = losses.BCEWithLogitsLoss()(pred_fake, t_real)
    else:
        if cfg['focal_loss']:
            G_loss = losses.rsgan_focal(pred_fake, pred_real, cfg['fl_gamma'], t_real)
        else:
            G_loss = losses.rsgan_bce(pred_fake, pred_real, t_real)
    G_loss = G_loss * cfg['alpha_gan']
    if two:
        main.wait_stream(side)
    total_G_loss = l1_loss + vgg_loss + G_loss + tv_loss
    total_G_loss.backward()
    if ddp_g is not None:
        ddp_g.finish()
    optim_G.step()
    return torch.stack([l1_loss.detach(), vgg_loss.detach(), G_loss.detach(), tv_loss.detach() / max(1, (ddp_g.world_size if ddp_g is not None else 1)),
                        total_D_loss.detach()])


DEFAULT_GAN_CFG = dict(alpha_l1=0.0, alpha_vgg=50.0, alpha_gan=1.0, alpha_tv=1e-6, fl_gamma=1.0, gan_type='RSGAN',
                       focal_loss=True)
