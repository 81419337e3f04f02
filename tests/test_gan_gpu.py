"""GPU module-level parity of the Discriminator, the VGG extractor and the full GAN step (train.py:202-259)
against the CPU oracle.

What is gated and how.
  * Forward values (logits, features, the five losses) against the free-running fp64 oracle.
  * Gradients against the FORWARD-PINNED oracle (oracle/pesr_oracle.py): the reference network differentiated in fp64
    at the activations the B200 path stored.  Gradients of LeakyReLU / ReLU / max-pool / BatchNorm networks are only
    piecewise linear in the activations, so a free-running comparison measures mask flips (sqrt(eps) of the forward
    deviation: 5e-2 for D, 9e-2 for VGG even between two accumulator widths of the oracle itself,
    test_oracle.py::test_rounding_noise_floor), not the backward kernels.  Free-running gradient errors are printed for
    information and not asserted.
  * The whole GAN step (train.py:202-259): every loss, every Generator AND Discriminator parameter gradient and the
    Discriminator's parameters after its Adam step, for all four (gan_type, focal_loss) branches.
"""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def test_discriminator_forward_backward_and_running_stats():
    from oracle import pesr_oracle as O
    from pesr_b200.model import Discriminator
    opt = {'patch_size': 16, 'spectral_norm': False}
    sd = O.init_discriminator(opt, 0)
    D = Discriminator(opt)
    D.load_state_dict(sd)
    D = D.cuda().train()
    g = torch.Generator().manual_seed(1)
    nb = 8
    x = torch.rand(nb, 3, 64, 64, generator=g) * 255
    R = torch.randn(nb, 1, generator=g)
    xc = x.cuda().requires_grad_(True)
    y = D(xc)
    assert y.shape == (nb, 1) and y.dtype == torch.float32
    (y * R.cuda()).sum().backward()
    leaf = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
            for k, v in sd.items()}
    xo = x.double().clone().requires_grad_(True)
    stats = []
    yo = O.discriminator_forward(leaf, xo, stats_out=stats)
    names = [k for k in leaf if leaf[k].is_floating_point() and leaf[k].requires_grad]
    og = torch.autograd.grad((yo * R.double()).sum(), [leaf[k] for k in names] + [xo])
    ograds = dict(zip(names, og[:-1]))
    e_fwd = rel_l2(y.detach().cpu(), yo.detach())
    errs = sorted(rel_l2(p.grad.cpu(), ograds[k]) for k, p in D.named_parameters())
    e_dx = rel_l2(xc.grad.cpu(), og[-1])
    print(f"D: logits rel {e_fwd:.2e}, param-grad rel median {errs[len(errs)//2]:.2e} max {errs[-1]:.2e}, dx rel {e_dx:.2e}")
    # 16-bit pre-BN and activation storage through 8 BatchNorm layers; the oracle's own fp16-rounded evaluation is 3.2e-3
    # away from its exact one on this input.  Free-running gradient errors are mask-flip noise (module docstring): not
    # asserted here, the gradient gate is test_pinned_gradients_gpu.py::test_discriminator_gradients_pinned
    assert e_fwd < 6e-3
    st = D.state_dict()
    for i in (0, 3, 7):
        mean, var, n = stats[i]
        assert rel_l2(st[f'features.{i}.1.running_mean'].cpu(), 0.1 * mean) < 2e-3
        assert rel_l2(st[f'features.{i}.1.running_var'].cpu(), 0.9 + 0.1 * var * n / (n - 1)) < 2e-3
        assert int(st[f'features.{i}.1.num_batches_tracked']) == 1
    # frozen-parameter call (G phase, train.py:234-238): only the input gradient flows
    for p in D.parameters():
        p.requires_grad = False
        p.grad = None
    xc.grad = None
    D(xc).sum().backward()
    assert xc.grad is not None and all(p.grad is None for p in D.parameters())
    with pytest.raises(NameError):
        Discriminator({'patch_size': 16, 'spectral_norm': True})


def test_discriminator_forward_pair_equals_two_calls():
    """D.forward_pair(a, b) is two train-mode calls (separate BatchNorm statistics, running statistics updated twice in
    order) executed as one batch of 2N images: same logits, same summed parameter gradients, same buffers -- up to the
    fp32 partial sums of the statistics kernel, whose blocking differs between the batched and the separate launches
    (1e-8 of a mean; amplified to ~1e-3 of the logits by the 16-bit rounding flips of the following layers, see
    test_oracle.py::test_rounding_noise_floor).  The exact equivalence of the batched execution is shown on the
    statistics themselves (running_mean / running_var after both calls) and on the first block's activations."""
    from oracle import pesr_oracle as O
    from pesr_b200.model import Discriminator
    opt = {'patch_size': 16, 'spectral_norm': False}
    sd = O.init_discriminator(opt, 0)
    g = torch.Generator().manual_seed(2)
    nb = 8
    a = (torch.rand(nb, 3, 64, 64, generator=g) * 255).cuda()
    b = (torch.rand(nb, 3, 64, 64, generator=g) * 255).cuda()
    ra, rb = torch.randn(nb, 1, generator=g).cuda(), torch.randn(nb, 1, generator=g).cuda()
    res = []
    for pair in (False, True):
        D = Discriminator(opt)
        D.load_state_dict(sd)
        D = D.cuda().train()
        if pair:
            ya, yb = D.forward_pair(a, b)
            pl = D.engine().pools[(nb, 64, 64, 2)][0]
            first = pl.A[0].clone()
        else:
            ya = D(a)
            first_a = D.engine().pools[(nb, 64, 64, 1)][0].A[0].clone()
            yb = D(b)
            first = torch.cat([first_a, D.engine().pools[(nb, 64, 64, 1)][1].A[0].clone()])
        ((ya * ra).sum() + (yb * rb).sum()).backward()
        res.append((ya.detach(), yb.detach(), {k: p.grad.clone() for k, p in D.named_parameters()},
                    {k: v.clone() for k, v in D.state_dict().items() if "running" in k or "tracked" in k}, first))
    (ya0, yb0, g0, s0, f0), (ya1, yb1, g1, s1, f1) = res
    assert rel_l2(f1.float(), f0.float()) < 1e-5          # block 0 output of both calls: at most a few 1-ulp flips
    assert rel_l2(ya1, ya0) < 3e-3 and rel_l2(yb1, yb0) < 3e-3
    for k in s0:
        if "tracked" in k:
            assert torch.equal(s0[k], s1[k]), k
        else:
            assert torch.allclose(s0[k], s1[k], rtol=2e-3, atol=1e-6), k    # two momentum updates in the same order
    assert rel_l2(s1['features.0.1.running_mean'], s0['features.0.1.running_mean']) < 1e-6
    assert rel_l2(s1['features.0.1.running_var'], s0['features.0.1.running_var']) < 1e-6
    # gradients of the two executions agree as well as two free-running evaluations can (mask flips); their exactness is
    # gated against the forward-pinned oracle (test_gan_step_matches_pinned_oracle runs the pair path)
    print("pair vs two calls: logits", rel_l2(ya1, ya0), rel_l2(yb1, yb0), "fc2 grad", rel_l2(g1['classifier.2.weight'], g0['classifier.2.weight']))


def test_vgg_features_and_input_gradient():
    from oracle import pesr_oracle as O
    from pesr_b200.model import VGG
    sd = O.init_vgg(2)
    V = VGG(pretrained=False)
    V.load_state_dict(sd)
    V = V.cuda()
    g = torch.Generator().manual_seed(3)
    sr = torch.rand(2, 3, 64, 64, generator=g) * 255
    hr = torch.rand(2, 3, 64, 64, generator=g) * 255
    src = sr.cuda().requires_grad_(True)
    f_sr, f_hr = V(src, hr.cuda())
    assert f_sr.shape == (2, 512, 4, 4) and f_sr.requires_grad and not f_hr.requires_grad
    from pesr_b200 import losses
    loss = losses.mse_loss(f_sr, f_hr)
    loss.backward()
    so = sr.double().clone().requires_grad_(True)
    of_sr, of_hr = O.vgg_forward({k: v.double() for k, v in sd.items()}, so, hr.double())
    ol = O.mse_loss(of_sr, of_hr)
    og, = torch.autograd.grad(ol, so)
    e = (rel_l2(f_sr.detach().cpu(), of_sr.detach()), rel_l2(f_hr.cpu(), of_hr), abs(float(loss) - float(ol)) / float(ol),
         rel_l2(src.grad.cpu(), og))
    print(f"VGG: f_sr rel {e[0]:.2e} f_hr rel {e[1]:.2e} loss rel {e[2]:.2e} dsr rel {e[3]:.2e}")
    # d(loss)/d(sr) free-running is mask-flip noise; gated in test_pinned_gradients_gpu.py::test_vgg_input_gradient_pinned
    assert e[0] < 3e-3 and e[1] < 3e-3 and e[2] < 3e-3
    assert all(not p.requires_grad for p in V.parameters())


class _Spy(torch.nn.Module):
    """Records the Generator's output inside steps.gan_step (which returns the losses only)."""

    def __init__(self, G):
        super().__init__()
        self.G, self.out = G, None

    def forward(self, x):
        self.out = self.G(x)
        return self.out


def _gan_step_against_pinned_oracle(opt, nb, gan_type, focal):
    import pinning
    from oracle import pesr_oracle as O
    from pesr_b200 import steps
    from pesr_b200.model import VGG, Discriminator, Generator
    from pesr_b200.optim import Adam
    patch = opt['patch_size']
    g_sd, d_sd, v_sd = O.init_generator(opt, 0), O.init_discriminator(opt, 1), O.init_vgg(2)
    G, D, V = Generator(opt), Discriminator(opt), VGG(pretrained=False)
    G.load_state_dict(g_sd), D.load_state_dict(d_sd), V.load_state_dict(v_sd)
    G, D, V = G.cuda(), D.cuda(), V.cuda()
    gen = torch.Generator().manual_seed(3)
    lr = torch.rand(nb, 3, patch, patch, generator=gen) * 255
    hr = torch.rand(nb, 3, patch * 4, patch * 4, generator=gen) * 255
    lrate = 5e-5
    optG, optD = Adam(G.parameters(), lr=lrate), Adam(D.parameters(), lr=lrate)
    cfg = dict(steps.DEFAULT_GAN_CFG, gan_type=gan_type, focal_loss=focal)
    cfg['target_real'] = torch.ones(nb, 1, device="cuda")
    cfg['target_fake'] = torch.zeros(nb, 1, device="cuda")
    spy = _Spy(G)
    tracer = pinning.StepTracer(G, D, V)
    got = steps.gan_step(spy, D, V, optG, optD, lr.cuda(), hr.cuda(), cfg).cpu()
    tracer.close()
    pins = tracer.gan_step_pins(spy.out)
    ref = O.gan_step(g_sd, d_sd, v_sd, lr, hr, opt, lr_rate=lrate, dtype=torch.float64, qdtype=torch.float16,
                     gan_type=gan_type, focal=focal, pins=pins)
    free = O.gan_step(g_sd, d_sd, v_sd, lr, hr, opt, lr_rate=lrate, dtype=torch.float64, gan_type=gan_type, focal=focal)
    tag = f"GAN step [{gan_type}, focal={focal}, depth {opt['depth']}, {opt['num_channels']} ch, patch {patch}, nb {nb}]"
    keys = ['l1', 'vgg', 'g_loss', 'tv', 'd_loss']
    rp = [abs(float(got[i]) - float(ref[k])) / max(abs(float(ref[k])), 1e-30) for i, k in enumerate(keys)]
    rf = [abs(float(got[i]) - float(free[k])) / max(abs(float(free[k])), 1e-30) for i, k in enumerate(keys)]
    print(f"{tag}: losses {[round(float(v), 6) for v in got]}; rel vs pinned oracle {['%.1e' % v for v in rp]}, "
          f"vs free-running fp64 oracle {['%.1e' % v for v in rf]}")
    assert float(got[0]) == 0.0                                     # alpha_l1 = 0 (train.py:76)
    for i in (1, 2, 3, 4):
        assert rp[i] < 5e-4, keys[i]      # same forward point: what is left is the loss kernels' own arithmetic
        assert rf[i] < 3e-3, keys[i]      # free-running: forward deviation of the 16-bit networks (logits, features)
    assert rel_l2(spy.out.detach().cpu(), free['sr']) < 1e-3
    ge = pinning.grad_errors(G.named_parameters(), ref['g_grads'], rel_l2)
    de = pinning.grad_errors(D.named_parameters(), ref['d_grads'], rel_l2)
    gf = sorted(rel_l2(p.grad.cpu(), free['g_grads'][k]) for k, p in G.named_parameters())
    print(f"{tag}: G grads rel-L2 median {ge[len(ge) // 2][0]:.2e} worst {ge[-1][0]:.2e} ({ge[-1][1]}); "
          f"D grads median {de[len(de) // 2][0]:.2e} worst {de[-1][0]:.2e} ({de[-1][1]}); "
          f"[free-running G median {gf[len(gf) // 2]:.2e}, information only]")
    assert ge[len(ge) // 2][0] < 2e-3 and ge[-1][0] < 3e-3      # measured 1.0e-3 .. 1.6e-3 / 1.3e-3 .. 1.8e-3
    assert de[len(de) // 2][0] < 2e-3 and de[-1][0] < 3e-3      # measured 6.5e-4 / 1.1e-3
    # optim_D.step() (train.py:229): first Adam step = -lr * g / (|g| + eps), i.e. +-lr per element; compare the UPDATE
    bad, total = 0, 0
    for k, p in D.named_parameters():
        u_got = (p.detach().cpu().double() - d_sd[k].double())
        u_ref = (ref['d_params_after'][k].double() - d_sd[k].double())
        bad += int(((u_got - u_ref).abs() > 0.05 * lrate).sum())
        total += u_got.numel()
        assert float(u_got.abs().max()) <= 1.001 * lrate + 1e-7 * float(d_sd[k].abs().max()), k
    print(f"{tag}: D Adam update differs (by > 5% of lr) on {bad} of {total} elements ({bad / total:.2e})")
    assert bad / total < 1e-3             # sign of near-zero gradient elements (measured 2e-4)
    moved = [float((p.detach().cpu() - g_sd[k]).abs().max()) for k, p in G.named_parameters()]
    assert max(moved) < 6.2e-5 and min(moved) > 0   # optim_G.step(): lr = 5e-5 plus fp32 rounding of |p| ~ 114 (MeanShift bias)
    return G, D, V, optG, optD, cfg, lr, hr


@pytest.mark.parametrize("gan_type,focal", [("RSGAN", True), ("RSGAN", False), ("SGAN", True), ("SGAN", False)])
def test_gan_step_matches_pinned_oracle(gan_type, focal):
    opt = {'depth': 2, 'num_channels': 64, 'res_scale': 0.1, 'patch_size': 12, 'spectral_norm': False}
    from pesr_b200 import steps
    G, D, V, optG, optD, cfg, lr, hr = _gan_step_against_pinned_oracle(opt, 4, gan_type, focal)
    # second step runs on re-packed weights and reuses every plan
    got2 = steps.gan_step(G, D, V, optG, optD, lr.cuda(), hr.cuda(), cfg)
    assert torch.isfinite(got2).all()


def test_gan_step_matches_pinned_oracle_full_width():
    """Same gates on the BASELINE networks (32 blocks, 256 channels, patch 48) at batch 2 (fp64 oracle: ~40 s of CPU)."""
    opt = {'depth': 32, 'num_channels': 256, 'res_scale': 0.1, 'patch_size': 48, 'spectral_norm': False}
    _gan_step_against_pinned_oracle(opt, 2, "RSGAN", True)
