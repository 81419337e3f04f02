"""Generates tests/golden/*.pt by running the UNMODIFIED reference modules from /root/reference on CPU.

Run once in the build container (the GPU box has no /root/reference):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Shims, both outside the arithmetic being pinned:
  * torchvision.models.vgg19(pretrained=True) needs a download -> patched to weights=None, then the
    seeded state dict from oracle.init_vgg is loaded into the reference VGG module.
  * test.py cannot be imported (argparse + imageio at import time): x8_forward's source text is
    extracted from /root/reference/test.py and exec'd with Tensor.cuda patched to the identity.
Weights are NOT stored: they are re-created from the seed (oracle.init_*), and a checksum is kept, so
the vectors also pin the reference's parameter construction order.
"""
import ast
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

import torchvision.models as tvm  # noqa: E402

_orig_vgg19 = tvm.vgg19
tvm.vgg19 = lambda pretrained=True, **kw: _orig_vgg19(weights=None)

import model as ref_model  # noqa: E402  (reference package)
import utils as ref_utils  # noqa: E402
from oracle import pesr_oracle as O  # noqa: E402


def checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values() if v.is_floating_point()))


def gen_case(name, opt, seed, shape):
    torch.manual_seed(seed)
    G = ref_model.Generator(opt)
    sd = {k: v.clone() for k, v in G.state_dict().items()}
    g = torch.Generator().manual_seed(seed + 1)
    lr = torch.rand(*shape, generator=g) * 255
    hr = torch.rand(shape[0], 3, shape[2] * 4, shape[3] * 4, generator=g) * 255
    sr = G(lr)
    loss = torch.nn.L1Loss()(sr, hr)
    loss.backward()
    grads = {k: p.grad.clone() for k, p in G.named_parameters()}
    keep = [k for k, v in grads.items() if v.numel() <= 40000] if opt['num_channels'] <= 64 else [
        'sub_mean.weight', 'sub_mean.bias', 'embed.weight', 'embed.bias', 'body.0.body.0.bias',
        f"body.{opt['depth']}.bias", 'upsample.0.bias', 'upsample.4.weight', 'upsample.4.bias', 'add_mean.weight',
        'add_mean.bias']
    out = dict(opt=opt, seed=seed, shape=shape, keys=list(sd.keys()), shapes={k: tuple(v.shape) for k, v in sd.items()},
               weight_checksum=checksum(sd), sr=sr.detach(), loss=loss.detach(),
               grads={k: grads[k] for k in keep},
               grad_norms={k: float(v.norm()) for k, v in grads.items()})
    torch.save(out, os.path.join(HERE, name))
    print(name, "sr", tuple(sr.shape), "loss", float(loss), "chk", out['weight_checksum'])


def disc_case(name, patch, seed, nb):
    opt = {'patch_size': patch, 'spectral_norm': False}
    torch.manual_seed(seed)
    D = ref_model.Discriminator(opt)
    D.train()
    sd = {k: v.clone() for k, v in D.state_dict().items()}
    g = torch.Generator().manual_seed(seed + 1)
    x = (torch.rand(nb, 3, patch * 4, patch * 4, generator=g) * 255).requires_grad_(True)
    y = D(x)
    loss = torch.nn.BCEWithLogitsLoss()(y, torch.ones_like(y))
    loss.backward()
    grads = {k: p.grad.clone() for k, p in D.named_parameters()}
    after = D.state_dict()
    small = {k: v for k, v in grads.items() if v.numel() <= 20000}
    out = dict(opt=opt, seed=seed, nb=nb, keys=list(sd.keys()), shapes={k: tuple(v.shape) for k, v in sd.items()},
               weight_checksum=checksum(sd), logits=y.detach(), loss=loss.detach(), dx=x.grad.clone(),
               grads=small, grad_norms={k: float(v.norm()) for k, v in grads.items()},
               running_mean_0=after['features.0.1.running_mean'].clone(),
               running_var_0=after['features.0.1.running_var'].clone(),
               running_mean_7=after['features.7.1.running_mean'].clone(),
               running_var_7=after['features.7.1.running_var'].clone())
    torch.save(out, os.path.join(HERE, name))
    print(name, "logits", y.detach().flatten().tolist(), "chk", out['weight_checksum'])


def vgg_case(name, seed, nb, side):
    V = ref_model.VGG()
    sd = O.init_vgg(seed)
    V.load_state_dict(sd)
    g = torch.Generator().manual_seed(seed + 1)
    sr = (torch.rand(nb, 3, side, side, generator=g) * 255).requires_grad_(True)
    hr = torch.rand(nb, 3, side, side, generator=g) * 255
    f_sr, f_hr = V(sr, hr)
    loss = torch.nn.functional.mse_loss(f_sr, f_hr)
    loss.backward()
    out = dict(seed=seed, nb=nb, side=side, keys=list(V.state_dict().keys()), weight_checksum=checksum(sd),
               f_sr=f_sr.detach(), f_hr=f_hr.detach(), loss=loss.detach(), dsr=sr.grad.clone(),
               hr_requires_grad=f_hr.requires_grad)
    torch.save(out, os.path.join(HERE, name))
    print(name, "feat", tuple(f_sr.shape), "loss", float(loss))


def focal_case(name):
    g = torch.Generator().manual_seed(3)
    out = {}
    for gamma in (0.5, 1, 2):
        x = torch.randn(16, 1, generator=g) * 3
        for tval in (1.0, 0.0):
            t = torch.full_like(x, tval)
            with torch.no_grad():  # model/focal_loss.py:13 raises under autograd on torch >= 1.0
                y = ref_model.FocalLoss(gamma)(x, t)
            out[(gamma, tval)] = dict(x=x, value=y)
    torch.save(out, os.path.join(HERE, name))
    print(name, {k: float(v['value']) for k, v in out.items()})


def x8_case(name):
    src = open(os.path.join(REF, "test.py")).read()
    tree = ast.parse(src)
    fn_src = next(ast.get_source_segment(src, n) for n in tree.body
                  if isinstance(n, ast.FunctionDef) and n.name == "x8_forward")
    from functools import reduce
    from torch.autograd import Variable
    ns = dict(torch=torch, Variable=Variable, reduce=reduce)
    exec(fn_src, ns)
    torch.Tensor.cuda = lambda self, *a, **k: self
    g = torch.Generator().manual_seed(5)
    img = torch.rand(1, 3, 6, 9, generator=g) * 255
    wt = torch.randn(3, 3, 3, 3, generator=g)

    def model(t):  # a deliberately non-equivariant "network"
        return torch.nn.functional.conv2d(torch.nn.functional.interpolate(t, scale_factor=2), wt, padding=1)

    out = ns["x8_forward"](img, model)
    a = torch.tensor([[[[0.5, 1.5, 2.5, -3.0, 254.5, 255.5, 300.0, 127.49999]]]]).expand(1, 3, 1, 8).clone()
    img_u8 = ref_utils.tensors_to_imgs([a.clone()])[0]
    torch.save(dict(img=img, wt=wt, out=out, round_in=a, round_out=torch.from_numpy(img_u8.copy())),
               os.path.join(HERE, name))
    print(name, tuple(out.shape), img_u8.reshape(-1)[:8])


if __name__ == "__main__":
    gen_case("gen_small.pt", {'depth': 2, 'num_channels': 64, 'res_scale': 0.1}, 11, (2, 3, 10, 12))
    gen_case("gen_full.pt", {'depth': 32, 'num_channels': 256, 'res_scale': 0.1}, 0, (1, 3, 12, 16))
    disc_case("disc_p12.pt", 12, 21, 4)
    vgg_case("vgg_64.pt", 31, 2, 64)
    focal_case("focal.pt")
    x8_case("x8_round.pt")
