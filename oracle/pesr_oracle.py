"""CPU oracle for the PESR hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-PyTorch (CPU, fp32 or fp64) functional restatement of what the reference computes on the path
BASELINE.json names.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this file; nothing under pesr_b200/ does.

Pinning status: the reference ships no tests, fixtures or golden vectors, so the
oracle is pinned against the reference ITSELF: tests/golden/make_golden.py imports the unmodified
modules from /root/reference (model/pesr.py, model/basic.py, model/focal_loss.py, model/vgg.py with the
two shims documented there), runs them on seeded inputs and commits the outputs under tests/golden/;
tests/test_oracle.py checks every function below against those vectors, and -- when /root/reference
is present -- against the live reference modules as well.

Every function cites the reference lines it restates.  The code is written functionally over a
state_dict (name -> tensor) instead of nn.Module classes, so that the same function serves fp32, fp64
and quantisation-matched ("16-bit operand") evaluation.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

DIV2K_MEAN = (0.4488, 0.4371, 0.4040)  # model/pesr.py:13
VGG_MEAN = (0.485, 0.456, 0.406)       # model/vgg.py:14
VGG_STD = (0.229, 0.224, 0.225)        # model/vgg.py:15 (times rgb_range)
VGG19_CFG = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 256, 'M', 512, 512, 512, 512, 'M', 512, 512, 512, 512]


# ------------------------------------------------------------------------------------------------
# parameter construction in the reference's order (so a torch seed reproduces the reference's init)
# ------------------------------------------------------------------------------------------------
def _conv_params(cin, cout, k, bias=True):
    m = nn.Conv2d(cin, cout, k, padding=k // 2, bias=bias)  # default init == nn.Conv2d.reset_parameters
    return m.weight.detach().clone(), (m.bias.detach().clone() if bias else None)


def _mean_shift_params(mean, std, sign, rgb_range=255):
    _conv_params(3, 3, 1)  # model/basic.py:11 constructs (and so draws random numbers for) a Conv2d first
    std_t = torch.tensor(std, dtype=torch.float32)
    w = (torch.eye(3) / std_t.view(3, 1)).view(3, 3, 1, 1)           # model/basic.py:13-14
    b = sign * rgb_range * torch.tensor(mean, dtype=torch.float32) / std_t  # model/basic.py:15-16
    return w, b


def init_generator(opt, seed=None):
    """State dict of model/pesr.py:3-26 `Generator(opt)` under torch.manual_seed(seed)."""
    if seed is not None:
        torch.manual_seed(seed)
    depth, c = opt['depth'], opt['num_channels']
    sd = {}
    for i in range(depth):                       # model/pesr.py:18-19, model/basic.py:40-44
        for j in (0, 2):
            w, b = _conv_params(c, c, 3)
            sd[f'body.{i}.body.{j}.weight'], sd[f'body.{i}.body.{j}.bias'] = w, b
    sd[f'body.{depth}.weight'], sd[f'body.{depth}.bias'] = _conv_params(c, c, 3)   # model/pesr.py:20
    sd['sub_mean.weight'], sd['sub_mean.bias'] = _mean_shift_params(DIV2K_MEAN, (1., 1., 1.), -1)  # :22
    sd['embed.weight'], sd['embed.bias'] = _conv_params(3, c, 3)                  # :23
    sd['upsample.0.weight'], sd['upsample.0.bias'] = _conv_params(c, 4 * c, 3)    # model/basic.py:56
    sd['upsample.2.weight'], sd['upsample.2.bias'] = _conv_params(c, 4 * c, 3)    # :58
    sd['upsample.4.weight'], sd['upsample.4.bias'] = _conv_params(c, 3, 3)        # :60
    sd['add_mean.weight'], sd['add_mean.bias'] = _mean_shift_params(DIV2K_MEAN, (1., 1., 1.), 1)   # pesr.py:26
    # key order of the reference's state_dict (registration order)
    order = (['sub_mean.weight', 'sub_mean.bias', 'embed.weight', 'embed.bias'] +
             [f'body.{i}.body.{j}.{p}' for i in range(depth) for j in (0, 2) for p in ('weight', 'bias')] +
             [f'body.{depth}.weight', f'body.{depth}.bias'] +
             [f'upsample.{k}.{p}' for k in (0, 2, 4) for p in ('weight', 'bias')] +
             ['add_mean.weight', 'add_mean.bias'])
    return {k: sd[k] for k in order}


D_CHANNELS = [(3, 64, 1), (64, 64, 2), (64, 128, 1), (128, 128, 2), (128, 256, 1), (256, 256, 2), (256, 512, 1),
              (512, 512, 2)]  # model/pesr.py:53-66


def init_discriminator(opt, seed=None):
    """State dict of model/pesr.py:40-75 `Discriminator(opt)` under torch.manual_seed(seed)."""
    if seed is not None:
        torch.manual_seed(seed)
    sd = {}
    for i, (cin, cout, _s) in enumerate(D_CHANNELS):
        w, _ = _conv_params(cin, cout, 3, bias=False)
        sd[f'features.{i}.0.weight'] = w
        sd[f'features.{i}.1.weight'] = torch.ones(cout)
        sd[f'features.{i}.1.bias'] = torch.zeros(cout)
        sd[f'features.{i}.1.running_mean'] = torch.zeros(cout)
        sd[f'features.{i}.1.running_var'] = torch.ones(cout)
        sd[f'features.{i}.1.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)
    side = opt['patch_size'] * 4 // 16                      # model/pesr.py:50,69
    l1 = nn.Linear(512 * side * side, 1024)                 # :71
    l2 = nn.Linear(1024, 1)                                 # :73
    sd['classifier.0.weight'], sd['classifier.0.bias'] = l1.weight.detach().clone(), l1.bias.detach().clone()
    sd['classifier.2.weight'], sd['classifier.2.bias'] = l2.weight.detach().clone(), l2.bias.detach().clone()
    return sd


def init_vgg(seed=None):
    """Random-init VGG19.features[:35] + MeanShift (model/vgg.py:8-15 with weights=None: no network here).
    torchvision's init: conv kaiming_normal_(fan_out, relu), bias 0.  Keys follow the reference:
    vgg.{idx}.weight / vgg.{idx}.bias, sub_mean.weight / sub_mean.bias."""
    if seed is not None:
        torch.manual_seed(seed)
    sd = {}
    idx, cin = 0, 3
    for v in VGG19_CFG:
        if v == 'M':
            idx += 1
            continue
        w = torch.empty(v, cin, 3, 3)
        nn.init.kaiming_normal_(w, mode='fan_out', nonlinearity='relu')
        sd[f'vgg.{idx}.weight'], sd[f'vgg.{idx}.bias'] = w, torch.zeros(v)
        idx += 2
        cin = v
    std = tuple(s * 255 for s in VGG_STD)
    sd['sub_mean.weight'], sd['sub_mean.bias'] = _mean_shift_params(VGG_MEAN, std, -1)
    return sd


# ------------------------------------------------------------------------------------------------
# forward passes
# ------------------------------------------------------------------------------------------------
def _q(t, qdtype):
    """Round to the kernel's 16-bit operand type and back (identity when qdtype is None)."""
    if qdtype is None:
        return t
    return t.to(qdtype).to(t.dtype)


class _RoundSTE(torch.autograd.Function):
    """Operand rounding with a straight-through gradient that is itself rounded: this is where the
    B200 path stores 16-bit tensors in forward and backward."""

    @staticmethod
    def forward(ctx, t, qdtype):
        ctx.qdtype = qdtype
        return t.to(qdtype).to(t.dtype)

    @staticmethod
    def backward(ctx, g):
        return g, None


def _qa(t, qdtype):
    return t if qdtype is None else _RoundSTE.apply(t, qdtype)


# ------------------------------------------------------------------------------------------------
# forward-pinned evaluation (test infrastructure for the GRADIENT gates)
#
# Networks of ReLU / LeakyReLU / max-pool / BatchNorm layers are only piecewise linear: a forward deviation of eps
# flips a fraction ~eps of the activation masks and perturbs the gradient by ~sqrt(eps).  Two evaluations of THIS
# oracle that round at identical points and differ only in the accumulator width (fp32 vs fp64) already disagree by
# 2e-3 in the Discriminator's logits and 6e-2 in its gradients (tests/test_oracle.py::test_rounding_noise_floor),
# because a 1e-7 difference before a 16-bit rounding flips that rounding with probability ~1e-4 and the flips
# compound layer by layer until both paths are, in effect, independently rounded.  A gradient gate against any
# free-running oracle is therefore conditioning-limited, whatever the kernel quality.
#
# `pin` removes the conditioning from the comparison: a dict of activations saved by the implementation under test
# (NCHW tensors, see the key names in each forward below).  The oracle then evaluates every layer at THOSE values
# (identity gradient through the substitution, activation masks taken from the pinned post-activation tensors), so
# its autograd gradient is the exact derivative of the reference network at the implementation's own forward point.
# ------------------------------------------------------------------------------------------------
class _PinAct(torch.autograd.Function):
    """value := the pinned post-activation tensor; gradient := g * act'(pinned) with act' in {1, slope}."""

    @staticmethod
    def forward(ctx, pre, post, slope):
        ctx.save_for_backward(post > 0)
        ctx.slope = slope
        return post.to(pre.dtype).clone()

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return g * torch.where(mask, 1.0, ctx.slope).to(g.dtype), None, None


def _pin(t, pin, key, trace=None):
    """Substitute the pinned value for t (identity gradient); `trace` (dict) records the value used."""
    if pin is not None and key in pin:
        t = t + (pin[key].to(t.dtype) - t).detach()
    if trace is not None and key is not None:
        trace[key] = t.detach()
    return t


def _act(pre, pin, key, slope, trace=None):
    """relu (slope 0) / leaky_relu(slope), evaluated at the pinned post-activation values when given."""
    if pin is not None and key in pin:
        out = _PinAct.apply(pre, pin[key], slope)
    else:
        out = F.relu(pre) if slope == 0 else F.leaky_relu(pre, slope)
    if trace is not None:
        trace[key] = out.detach()
    return out


def generator_forward(sd, x, depth, res_scale, qdtype=None, pin=None, trace=None):
    """model/pesr.py:28-38.  qdtype (torch.float16 / bfloat16) emulates 16-bit conv operands with the
    fp32 residual stream the B200 schedule keeps (quantisation-matched oracle).
    pin keys (16-bit conv operands saved by the implementation): 'x{i}' = input of block i (i = depth: input of the
    tail conv), 't{i}' = relu(conv1) of block i, 'u0' / 'u1' / 'u2' = inputs of upsample.0 / .2 / .4."""
    def conv(t, name, key=None):
        return F.conv2d(_pin(_qa(t, qdtype), pin, key, trace), _q(sd[name + '.weight'], qdtype), sd[name + '.bias'], padding=1)

    x = F.conv2d(x, sd['sub_mean.weight'], sd['sub_mean.bias'])               # :29
    x = conv(x, 'embed')                                                       # :31
    res = x
    for i in range(depth):                                                     # :32, model/basic.py:48-52
        t = _act(conv(res, f'body.{i}.body.0', f'x{i}'), pin, f't{i}', 0.0, trace)
        res = conv(t, f'body.{i}.body.2') * res_scale + res
    res = conv(res, f'body.{depth}', f'x{depth}') + x                          # :32-33
    u = F.pixel_shuffle(conv(res, 'upsample.0', 'u0'), 2)                      # model/basic.py:56-57
    u = F.pixel_shuffle(conv(u, 'upsample.2', 'u1'), 2)                        # :58-59
    u = conv(u, 'upsample.4', 'u2')                                            # :60
    return F.conv2d(u, sd['add_mean.weight'], sd['add_mean.bias'])             # model/pesr.py:36


def discriminator_forward(sd, x, eps=1e-5, qdtype=None, stats_out=None, pin=None, trace=None):
    """model/pesr.py:77-81 in train mode: BatchNorm uses the batch's own (biased) statistics
    (model/basic.py:29).  stats_out (list) receives per-layer (mean, biased var) for the running-stat update.
    pin keys: 'y{i}' = conv output of block i before BatchNorm (as stored, 16-bit), 'a{i}' = block output after
    LeakyReLU (16-bit), 'h1' = classifier.1 output (fp32).  With qdtype the second Linear also sees 16-bit operands,
    as on the B200 path."""
    for i, (_cin, _cout, stride) in enumerate(D_CHANNELS):
        if qdtype is not None and i == 0:
            # the B200 path feeds conv 0 the image minus 127.5 (padding included): BatchNorm cancels the shift
            xin = F.pad(_qa(x - 127.5, qdtype), (1, 1, 1, 1), value=-127.5)
            y = F.conv2d(xin, _q(sd['features.0.0.weight'], qdtype), None, stride=stride)
        else:
            y = F.conv2d(_qa(x, qdtype), _q(sd[f'features.{i}.0.weight'], qdtype), None, stride=stride, padding=1)
        y = _pin(_qa(y, qdtype), pin, f'y{i}', trace)   # the pre-BN tensor is stored in 16 bits by the B200 path
        mean = y.mean(dim=(0, 2, 3))
        var = y.var(dim=(0, 2, 3), unbiased=False)
        if stats_out is not None:
            stats_out.append((mean.detach(), var.detach(), y.numel() // y.shape[1]))
        y = (y - mean.view(1, -1, 1, 1)) * torch.rsqrt(var.view(1, -1, 1, 1) + eps)
        y = y * sd[f'features.{i}.1.weight'].view(1, -1, 1, 1) + sd[f'features.{i}.1.bias'].view(1, -1, 1, 1)
        x = _act(y, pin, f'a{i}', 0.2, trace)
    f = x.reshape(x.shape[0], -1)                                              # NCHW flatten, model/pesr.py:79
    h = _act(F.linear(_qa(f, qdtype), _q(sd['classifier.0.weight'], qdtype), sd['classifier.0.bias']), pin, 'h1', 0.2, trace)
    return F.linear(_qa(h, qdtype), _q(sd['classifier.2.weight'], qdtype), sd['classifier.2.bias'])


def vgg_features(sd, x, qdtype=None, pin=None, trace=None):
    """model/vgg.py:18-22: sub_mean then vgg19.features[:35] (conv5_4 output, before its ReLU).
    pin keys: 'c{k}' = output of conv k (0-based) after its ReLU (the last one: the pre-ReLU features), 16-bit.
    With qdtype the returned features are rounded to 16 bits as well (the B200 path stores them so)."""
    x = F.conv2d(x, sd['sub_mean.weight'], sd['sub_mean.bias'])
    idx = 0
    n_conv = sum(1 for v in VGG19_CFG if v != 'M')
    k = 0
    for v in VGG19_CFG:
        if v == 'M':
            x = F.max_pool2d(x, 2, 2)
            idx += 1
            continue
        x = F.conv2d(_qa(x, qdtype), _q(sd[f'vgg.{idx}.weight'], qdtype), sd[f'vgg.{idx}.bias'], padding=1)
        k += 1
        if k < n_conv:
            x = _act(x, pin, f'c{k - 1}', 0.0, trace)
        else:
            x = _pin(_qa(x, qdtype), pin, f'c{k - 1}', trace)
        idx += 2
    return x


def vgg_forward(sd, sr, hr, qdtype=None, pin=None, trace=None):
    """model/vgg.py:18-28: (features(sr), features(hr) without grad).  pin applies to the sr branch."""
    f_sr = vgg_features(sd, sr, qdtype, pin, trace)
    with torch.no_grad():
        f_hr = vgg_features(sd, hr.detach(), qdtype)
    return f_sr, f_hr


# ------------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------------
def bce_with_logits_torch04(x, t, weight=None):
    """torch 0.4's Python F.binary_cross_entropy_with_logits (the version README.md:22 pins); unlike
    torch>=1.0 the gradient flows through `weight`."""
    max_val = (-x).clamp(min=0)
    loss = x - x * t + max_val + ((-max_val).exp() + (-x - max_val).exp()).log()
    if weight is not None:
        loss = loss * weight
    return loss.mean()


def focal_loss(x, t, gamma, detach_weight=False):
    """model/focal_loss.py:9-13."""
    p = x.sigmoid()
    pt = p * t + (1 - p) * (1 - t)
    w = (1 - pt).pow(gamma)
    if detach_weight:
        w = w.detach()
    return bce_with_logits_torch04(x, t, w)


def tv_loss(y):
    """train.py:137-140 (a SUM, not a mean)."""
    return (y[:, :, :, :-1] - y[:, :, :, 1:]).abs().sum() + (y[:, :, :-1, :] - y[:, :, 1:, :]).abs().sum()


def l1_loss(a, b):
    return (a - b).abs().mean()          # nn.L1Loss(), train.py:131


def mse_loss(a, b):
    return ((a - b) ** 2).mean()         # F.mse_loss, train.py:136


# ------------------------------------------------------------------------------------------------
# optimiser (torch.optim.Adam semantics, train.py:124-126: betas (0.9, 0.999), eps 1e-8, no weight decay)
# ------------------------------------------------------------------------------------------------
def adam_update(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


# ------------------------------------------------------------------------------------------------
# step bodies
# ------------------------------------------------------------------------------------------------
def _leaf(sd, dtype):
    out = {}
    for k, v in sd.items():
        if v.is_floating_point():
            out[k] = v.detach().to(dtype).clone().requires_grad_(not k.endswith(('running_mean', 'running_var')))
        else:
            out[k] = v
    return out


def pretrain_step(g_sd, lr_img, hr_img, opt, dtype=torch.float32, qdtype=None):
    """train.py:164-176 up to (not including) the optimiser step: returns loss, sr and d(loss)/d(param)."""
    sd = _leaf(g_sd, dtype)
    sr = generator_forward(sd, lr_img.to(dtype), opt['depth'], opt['res_scale'], qdtype)
    loss = l1_loss(sr, hr_img.to(dtype))
    names = [k for k in sd]
    grads = torch.autograd.grad(loss, [sd[k] for k in names])
    return loss.detach(), sr.detach(), dict(zip(names, grads))


def gan_step(g_sd, d_sd, v_sd, lr_img, hr_img, opt, lr_rate=5e-5, alpha_l1=0.0, alpha_vgg=50.0, alpha_gan=1.0,
             alpha_tv=1e-6, gamma=1.0, dtype=torch.float32, qdtype=None, focal_detach=False, gan_type='RSGAN',
             focal=True, pins=None, gp_u=None):
    """train.py:202-259 (defaults of train.py:64-76: RSGAN + focal loss; gan_type='SGAN' and focal=False select the
    other branches of train.py:210-213 / :244-253), with one real Adam step on D between the two phases as in
    train.py:229.  Returns a dict of losses, gradients and `sr`.
    pins (forward-pinned evaluation, see above): {'g': pins of the Generator forward, 'd': [pins of the four
    Discriminator forwards in the order D(hr), D(sr.detach()), D(sr), D(hr)], 'v': pins of the VGG sr branch,
    'f_hr': the target features, 'sr': the Generator output}."""
    g = _leaf(g_sd, dtype)
    d = _leaf(d_sd, dtype)
    v = {k: t.detach().to(dtype) if t.is_floating_point() else t for k, t in v_sd.items()}
    lr_img, hr_img = lr_img.to(dtype), hr_img.to(dtype)
    d_train = [k for k in d if d[k].is_floating_point() and d[k].requires_grad]
    out = {}
    pins = pins or {}
    pd = pins.get('d', [None] * 4)
    # ---- D phase, train.py:202-229
    pred_real = discriminator_forward(d, hr_img, qdtype=qdtype, pin=pd[0])
    sr = generator_forward(g, lr_img, opt['depth'], opt['res_scale'], qdtype, pin=pins.get('g'))
    sr = _pin(sr, pins, 'sr')
    pred_fake = discriminator_forward(d, sr.detach(), qdtype=qdtype, pin=pd[1])
    ones = torch.ones_like(pred_real)
    if gan_type == 'SGAN':                                                     # train.py:210-211
        d_loss = (F.binary_cross_entropy_with_logits(pred_real, ones) +
                  F.binary_cross_entropy_with_logits(pred_fake, torch.zeros_like(pred_fake)))
    else:
        d_loss = F.binary_cross_entropy_with_logits(pred_real - pred_fake, ones)   # train.py:213
    if gp_u is not None:                                                       # gradient penalty, train.py:216-226
        x_both = (hr_img * gp_u.to(dtype) + sr.detach() * (1 - gp_u.to(dtype))).detach().requires_grad_(True)
        pred_both = discriminator_forward(d, x_both, qdtype=None)
        grad = torch.autograd.grad(outputs=pred_both, inputs=x_both, grad_outputs=torch.ones_like(pred_both),
                                   retain_graph=True, create_graph=True, only_inputs=True)[0]
        out['gp'] = 10 * ((grad.norm(2, 1).norm(2, 1).norm(2, 1) - 1) ** 2).mean()
        d_loss = d_loss + out['gp']
        out['gp'] = out['gp'].detach()
    d_grads = torch.autograd.grad(d_loss, [d[k] for k in d_train])
    out['d_loss'] = d_loss.detach()
    out['d_grads'] = dict(zip(d_train, d_grads))
    out['pred_real_d'], out['pred_fake_d'] = pred_real.detach(), pred_fake.detach()
    with torch.no_grad():                                                      # optim_D.step(), first step
        for k, gk in zip(d_train, d_grads):
            adam_update(d[k], gk, torch.zeros_like(gk), torch.zeros_like(gk), 1, lr_rate)
    out['d_params_after'] = {k: d[k].detach().clone() for k in d_train}
    # ---- G phase, train.py:234-259
    pred_fake = discriminator_forward(d, sr, qdtype=qdtype, pin=pd[2])
    pred_real = discriminator_forward(d, hr_img, qdtype=qdtype, pin=pd[3])
    l1 = l1_loss(sr, hr_img) * alpha_l1
    f_sr, f_hr = vgg_forward(v, sr, hr_img, qdtype, pin=pins.get('v'))
    if 'f_hr' in pins:
        f_hr = pins['f_hr'].to(dtype)
    vgg_l = mse_loss(f_sr, f_hr) * alpha_vgg
    tv = tv_loss(sr) * alpha_tv
    g_in = pred_fake if gan_type == 'SGAN' else pred_fake - pred_real          # train.py:244-253
    if focal:
        g_l = focal_loss(g_in, ones, gamma, detach_weight=focal_detach) * alpha_gan
    else:
        g_l = F.binary_cross_entropy_with_logits(g_in, ones) * alpha_gan
    total = l1 + vgg_l + g_l + tv
    g_names = [k for k in g]
    g_grads = torch.autograd.grad(total, [g[k] for k in g_names] + [sr], allow_unused=True)
    out.update(l1=l1.detach(), vgg=vgg_l.detach(), g_loss=g_l.detach(), tv=tv.detach(), total_g=total.detach(),
               sr=sr.detach(), g_grads=dict(zip(g_names, g_grads[:-1])), dsr=g_grads[-1],
               pred_fake_g=pred_fake.detach(), pred_real_g=pred_real.detach())
    return out


# ------------------------------------------------------------------------------------------------
# inference
# ------------------------------------------------------------------------------------------------
def x8_forward(fn, img):
    """test.py:45-74: 8 flip/transpose variants -> model -> inverse transforms -> mean.
    Variant i: bit0 = flip W ('vflip'), bit1 = flip H ('hflip'), bit2 = transpose."""
    outs = []
    for i in range(8):
        t = img
        if i & 1:
            t = t.flip(3)
        if i & 2:
            t = t.flip(2)
        if i & 4:
            t = t.transpose(2, 3)
        o = fn(t.contiguous())
        if i & 4:
            o = o.transpose(2, 3)
        if i & 2:
            o = o.flip(2)
        if i & 1:
            o = o.flip(3)
        outs.append(o)
    total = outs[0]
    for o in outs[1:]:
        total = total + o
    return total / 8


def infer(fn_perc, img, alpha=1.0, fn_psnr=None):
    """test.py:106-109."""
    out = fn_perc(img)
    if alpha != 1:
        out = alpha * out + (1 - alpha) * x8_forward(fn_psnr, img)
    return out


def tensors_to_img_u8(x):
    """utils.py:13-18: clip, round half to even (numpy), HWC uint8."""
    a = x.squeeze(0).detach().cpu().numpy()
    return a.clip(0, 255).round().transpose(1, 2, 0).astype('uint8')
