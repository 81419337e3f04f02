"""GPU parity at the BASELINE.json configurations themselves (not scaled-down stand-ins):

  configs[0]  x4 inference of one 128x128 image            configs[4]  x4 inference of a 339x510 (DIV2K-size) image
  configs[1]  L1 pretrain step, batch 16 of 48x48 patches   configs[2]  GAN fine-tune step, batch 16 at 48x48

against the fp32 CPU oracle (the reference's own precision; fp64 at these sizes would take minutes).  Gates are the
north_star's: 1e-3 relative L2 for forward outputs and losses, >= 50 dB PSNR of the SR image.  Gradients at these sizes
are gated in test_pinned_gradients_gpu.py (forward-pinned oracle).  Each test costs 5-40 s of CPU oracle time.
"""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu

OPT = {'patch_size': 48, 'num_channels': 256, 'depth': 32, 'res_scale': 0.1, 'spectral_norm': False}


def _psnr(a, b):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 10 * torch.log10(torch.tensor(255.0 ** 2 / max(mse, 1e-20)))


@pytest.fixture(scope="module")
def gen():
    from oracle import pesr_oracle as O
    from pesr_b200.model import Generator
    sd = O.init_generator(OPT, 0)
    G = Generator(OPT)
    G.load_state_dict(sd)
    return G.cuda(), sd


@pytest.mark.parametrize("h,w", [(128, 128), (339, 510)], ids=["config1-128x128", "config5-339x510"])
def test_inference_matches_oracle_at_baseline_sizes(gen, h, w):
    from oracle import pesr_oracle as O
    G, sd = gen
    torch.set_num_threads(max(1, torch.get_num_threads()))
    g = torch.Generator().manual_seed(11)
    lr = torch.rand(1, 3, h, w, generator=g) * 255
    G.eval()
    with torch.no_grad():
        sr = G(lr.cuda()).cpu()
        osr = O.generator_forward(sd, lr, OPT['depth'], OPT['res_scale'])
    G.train()
    e, p = rel_l2(sr, osr), float(_psnr(sr, osr))
    print(f"inference {h}x{w}: sr rel-L2 {e:.2e}, PSNR vs oracle {p:.1f} dB")
    assert sr.shape == (1, 3, 4 * h, 4 * w)
    assert e < 1e-3 and p >= 50.0


def test_pretrain_step_at_headline_config(gen):
    """train.py:164-176 at B=16, 48x48 -> 192x192, 256 channels, 32 blocks."""
    from oracle import pesr_oracle as O
    from pesr_b200 import losses
    G, sd = gen
    g = torch.Generator().manual_seed(0)
    lr = torch.rand(16, 3, 48, 48, generator=g) * 255
    hr = torch.rand(16, 3, 192, 192, generator=g) * 255
    for p in G.parameters():
        p.grad = None
    sr = G(lr.cuda())
    loss = losses.l1_loss(sr, hr.cuda())
    loss.backward()
    oloss, osr, ograds = O.pretrain_step(sd, lr, hr, OPT)
    e_sr, e_loss, p = rel_l2(sr.detach().cpu(), osr), abs(float(loss) - float(oloss)) / float(oloss), float(_psnr(sr.detach().cpu(), osr))
    # the L1 loss' sign() makes its parameter gradients ill-conditioned; they are reported here
    # and gated with a smooth loss in test_pinned_gradients_gpu.py / test_generator_gpu.py
    errs = sorted(rel_l2(pp.grad.cpu(), ograds[k]) for k, pp in G.named_parameters())
    print(f"pretrain B=16: sr rel-L2 {e_sr:.2e} ({p:.1f} dB), L1 loss rel {e_loss:.2e}, L1-gradient rel-L2 median "
          f"{errs[len(errs) // 2]:.2e} max {errs[-1]:.2e} (vs fp32 oracle, sign()-conditioned)")
    assert e_sr < 1e-3 and p >= 50.0 and e_loss < 1e-3
    assert all(torch.isfinite(pp.grad).all() for pp in G.parameters())


def test_gan_step_at_headline_config(gen):
    """train.py:202-259 at B=16, patch 48: the five losses and sr against the fp32 oracle."""
    from oracle import pesr_oracle as O
    from pesr_b200 import steps
    from pesr_b200.model import VGG, Discriminator
    from pesr_b200.optim import Adam
    G, g_sd = gen
    G.load_state_dict(g_sd)
    d_sd, v_sd = O.init_discriminator(OPT, 1), O.init_vgg(2)
    D, V = Discriminator(OPT), VGG(pretrained=False)
    D.load_state_dict(d_sd), V.load_state_dict(v_sd)
    D, V = D.cuda(), V.cuda()
    gen_ = torch.Generator().manual_seed(0)
    lr = torch.rand(16, 3, 48, 48, generator=gen_) * 255
    hr = torch.rand(16, 3, 192, 192, generator=gen_) * 255
    optG, optD = Adam(G.parameters(), lr=5e-5), Adam(D.parameters(), lr=5e-5)
    cfg = dict(steps.DEFAULT_GAN_CFG)
    cfg['target_real'] = torch.ones(16, 1, device="cuda")
    cfg['target_fake'] = torch.zeros(16, 1, device="cuda")
    got = steps.gan_step(G, D, V, optG, optD, lr.cuda(), hr.cuda(), cfg).cpu()
    out = O.gan_step(g_sd, d_sd, v_sd, lr, hr, OPT)
    ref = torch.stack([out['l1'], out['vgg'], out['g_loss'], out['tv'], out['d_loss']]).float()
    names = ["l1", "vgg", "g", "tv", "d"]
    rels = [abs(float(got[i]) - float(ref[i])) / max(abs(float(ref[i])), 1e-30) for i in range(5)]
    print("GAN step B=16 losses: " + ", ".join(f"{n} {float(got[i]):.6g} (ref {float(ref[i]):.6g}, rel {rels[i]:.1e})"
                                               for i, n in enumerate(names)))
    assert float(got[0]) == 0.0 and float(ref[0]) == 0.0          # alpha_l1 = 0 (train.py:76)
    for i in (1, 3):
        assert rels[i] < 1e-3, names[i]
    # the adversarial terms go through 8 train-mode BatchNorm layers on 16-bit activations, twice (the G loss after D's
    # Adam step): measured 6.4e-4 (g) and 1.2e-5 (d); the G loss is gated at 2e-3 because two free-running evaluations of a
    # 16-bit BatchNorm stack agree on the logits only to ~2e-3 (test_oracle.py::test_rounding_noise_floor)
    assert rels[4] < 1e-3, names[4]
    assert rels[2] < 2e-3, names[2]
    G.load_state_dict(g_sd)      # leave the module-scoped fixture as it was
