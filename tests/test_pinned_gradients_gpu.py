"""GPU gradient parity of the three networks against the FORWARD-PINNED oracle (oracle/pesr_oracle.py, "forward-pinned
evaluation"): the oracle evaluates the reference network in fp64 at the activations the B200 path itself stored, so its
autograd gradient is the exact derivative the backward kernels have to reproduce -- independent of the mask-flip
conditioning that limits any comparison between two free-running evaluations of a ReLU / LeakyReLU / max-pool /
BatchNorm network (tests/test_oracle.py::test_rounding_noise_floor measures that floor on the oracle alone).

Forward values are gated separately, against the free-running fp64 oracle (test_gan_gpu.py, test_generator_gpu.py,
test_headline_gpu.py).

Gates: every parameter / input gradient within 2e-3 relative L2 (median over the tensors) and 3e-3 (worst tensor).  What
remains is the 16-bit storage of the gradient tensors between layers (one rounding of 2^-11/sqrt(3) = 2.8e-4 per stored
tensor, accumulating like sqrt(#layers)).  Measured on the B200 (round 2): Discriminator median 6.5e-4 / worst 1.25e-3,
VGG d(loss)/d(sr) 7.5e-4, Generator median 3.8e-4 .. 5.1e-4 / worst 6.1e-4, at the headline sizes as well as the small
ones; the numbers are printed.
"""
import pytest
import torch

import pinning
from conftest import rel_l2

pytestmark = pytest.mark.gpu

MED, WORST = 2e-3, 3e-3


def _report(name, errs, extra=""):
    med, worst = errs[len(errs) // 2], errs[-1]
    print(f"{name}: param-grad rel-L2 median {med[0]:.2e} ({med[1]}), worst {worst[0]:.2e} ({worst[1]}) {extra}")
    return med[0], worst[0]


@pytest.mark.parametrize("patch,nb", [(16, 8), (48, 16)], ids=["patch16-b8", "headline-patch48-b16"])
def test_discriminator_gradients_pinned(patch, nb):
    from oracle import pesr_oracle as O
    from pesr_b200.model import Discriminator
    opt = {'patch_size': patch, 'spectral_norm': False}
    sd = O.init_discriminator(opt, 0)
    D = Discriminator(opt)
    D.load_state_dict(sd)
    D = D.cuda().train()
    g = torch.Generator().manual_seed(1)
    side = 4 * patch
    x = torch.rand(nb, 3, side, side, generator=g) * 255
    R = torch.randn(nb, 1, generator=g)
    xc = x.cuda().requires_grad_(True)
    y = D(xc)
    pins = pinning.discriminator_pins(D, nb, side, side)
    (y * R.cuda()).sum().backward()
    leaf = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
            for k, v in sd.items()}
    xo = x.double().clone().requires_grad_(True)
    yo = O.discriminator_forward(leaf, xo, qdtype=torch.float16, pin=pins)
    names = [k for k in leaf if leaf[k].is_floating_point() and leaf[k].requires_grad]
    og = torch.autograd.grad((yo * R.double()).sum(), [leaf[k] for k in names] + [xo])
    ograds = dict(zip(names, og[:-1]))
    e_logits = rel_l2(y.detach().cpu(), yo.detach())
    e_dx = rel_l2(xc.grad.cpu(), og[-1])
    med, worst = _report(f"D patch {patch} nb {nb}", pinning.grad_errors(D.named_parameters(), ograds, rel_l2),
                         f"| logits (from pinned h1) {e_logits:.2e}, dx {e_dx:.2e}")
    assert e_logits < 1e-3          # last Linear on the pinned h1: only its own 16-bit operand rounding remains
    assert med < MED and worst < WORST and e_dx < WORST


@pytest.mark.parametrize("nb,side", [(2, 64), (4, 192)], ids=["64px-b2", "headline-192px-b4"])
def test_vgg_input_gradient_pinned(nb, side):
    from oracle import pesr_oracle as O
    from pesr_b200 import losses
    from pesr_b200.model import VGG
    sd = O.init_vgg(2)
    V = VGG(pretrained=False)
    V.load_state_dict(sd)
    V = V.cuda()
    g = torch.Generator().manual_seed(3)
    sr = torch.rand(nb, 3, side, side, generator=g) * 255
    hr = torch.rand(nb, 3, side, side, generator=g) * 255
    src = sr.cuda().requires_grad_(True)
    f_sr, f_hr = V(src, hr.cuda())
    pins = pinning.vgg_pins(V, nb, side, side)
    loss = losses.mse_loss(f_sr, f_hr)
    loss.backward()
    so = sr.double().clone().requires_grad_(True)
    of_sr = O.vgg_features({k: v.double() for k, v in sd.items()}, so, qdtype=torch.float16, pin=pins)
    ol = O.mse_loss(of_sr, f_hr.double().cpu())       # same target features: the comparison isolates the sr branch
    og, = torch.autograd.grad(ol, so)
    e_f = rel_l2(f_sr.detach().cpu(), of_sr.detach())
    e_l = abs(float(loss) - float(ol)) / float(ol)
    e_g = rel_l2(src.grad.cpu(), og)
    print(f"VGG {side}px nb {nb}: pinned features {e_f:.2e}, loss {e_l:.2e}, d(loss)/d(sr) rel-L2 {e_g:.2e}")
    assert e_f < 1e-6 and e_l < 1e-5        # pinned: identical by construction
    assert e_g < 2e-3                        # 16 stored 16-bit gradient tensors between the loss and the image


@pytest.mark.parametrize("opt,shape", [({'depth': 2, 'num_channels': 64, 'res_scale': 0.1}, (2, 3, 10, 12)),
                                       ({'depth': 3, 'num_channels': 128, 'res_scale': 0.1}, (1, 3, 24, 24)),
                                       ({'depth': 32, 'num_channels': 256, 'res_scale': 0.1}, (2, 3, 48, 48))],
                         ids=["d2c64", "d3c128", "headline-d32c256-b2"])
def test_generator_gradients_pinned(opt, shape):
    from oracle import pesr_oracle as O
    from pesr_b200.model import Generator
    sd = O.init_generator(opt, 3)
    G = Generator(opt)
    G.load_state_dict(sd)
    G = G.cuda()
    g = torch.Generator().manual_seed(4)
    lr = torch.rand(*shape, generator=g) * 255
    R = torch.randn(shape[0], 3, shape[2] * 4, shape[3] * 4, generator=g)
    lr_c = lr.cuda().requires_grad_(True)
    sr = G(lr_c)
    pins = pinning.generator_pins(G, shape[0], shape[2], shape[3])
    (sr * R.cuda()).sum().backward()
    leaf = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    x = lr.double().clone().requires_grad_(True)
    osr = O.generator_forward(leaf, x, opt['depth'], opt['res_scale'], qdtype=torch.float16, pin=pins)
    names = list(leaf)
    og = torch.autograd.grad((osr * R.double()).sum(), [leaf[k] for k in names] + [x])
    ograds = dict(zip(names, og[:-1]))
    e_sr = rel_l2(sr.detach().cpu(), osr.detach())
    e_dx = rel_l2(lr_c.grad.cpu(), og[-1])
    med, worst = _report(f"G {opt['depth']}x{opt['num_channels']} {shape}", pinning.grad_errors(G.named_parameters(), ograds, rel_l2),
                         f"| sr {e_sr:.2e}, d/d(lr) {e_dx:.2e}")
    assert e_sr < 1e-3
    assert med < MED and worst < WORST and e_dx < WORST
