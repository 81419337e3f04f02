"""Split-precision Discriminator backward vs the fp64 oracle with parts of the weights moved off their initial values (bring-up)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pesr_oracle as O  # noqa: E402
from pesr_b200.model import Discriminator  # noqa: E402


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


opt = {'patch_size': 12, 'spectral_norm': False}
nb = 4
base = O.init_discriminator(opt, 1)
g = torch.Generator().manual_seed(1)
x = torch.rand(nb, 3, 48, 48, generator=g) * 255
R = torch.randn(nb, 1, generator=g)
for tag, sel, amp in (("initial", lambda k: False, 0), ("BN gamma +-5e-5", lambda k: k.endswith("1.weight"), 5e-5),
                      ("BN beta +-5e-5", lambda k: k.endswith("1.bias"), 5e-5), ("BN beta +-0.3", lambda k: k.endswith("1.bias"), 0.3),
                      ("conv weights +-5e-5", lambda k: k.endswith("0.weight") and "features" in k, 5e-5),
                      ("classifier +-5e-5", lambda k: "classifier" in k, 5e-5), ("everything +-5e-5", lambda k: True, 5e-5)):
    sd = {k: v.clone() for k, v in base.items()}
    for k in sd:
        if sd[k].is_floating_point() and "running" not in k and sel(k):
            sd[k] += amp * torch.sign(torch.randn(sd[k].shape, generator=g))
    leaf = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
    xo = x.double().clone().requires_grad_(True)
    yo = O.discriminator_forward(leaf, xo)
    names = [k for k in leaf if leaf[k].is_floating_point() and leaf[k].requires_grad]
    og = torch.autograd.grad((yo * R.double()).sum(), [leaf[k] for k in names] + [xo])
    ograds = dict(zip(names, og[:-1]))
    D = Discriminator(opt, split_precision=True)
    D.load_state_dict(sd)
    D = D.cuda().train()
    xc = x.cuda().requires_grad_(True)
    y = D(xc)
    (y * R.cuda()).sum().backward()
    errs = sorted((rel_l2(p.grad.cpu(), ograds[k]), k) for k, p in D.named_parameters())
    print(f"{tag}: logits {rel_l2(y.detach().cpu(), yo.detach()):.2e}, param grads median {errs[len(errs) // 2][0]:.2e} max {errs[-1][0]:.2e} "
          f"({errs[-1][1]}), d/dx {rel_l2(xc.grad.cpu(), og[-1]):.2e}; worst five: {[(k, '%.1e' % e) for e, k in errs[-5:]]}")
