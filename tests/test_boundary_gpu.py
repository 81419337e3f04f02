"""GPU tests of the drop-in boundary beyond the headline path: the building-block modules are
callable as in the reference (model/basic.py:4-17,33-52), the Discriminator accepts any per-call batch (train.py:48
leaves --batch_size free), several autograd graphs of one shape may be alive, plan caches stay bounded over many image
sizes (test.py walks whole datasets), channel counts that are multiples of 64 but not powers of two work, the optimiser
state interchanges with torch.optim.Adam, and a CUDA-graph replay of the step equals the eager step."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _q(t):
    return t.half().double()


def test_building_blocks_are_callable_like_the_reference():
    """Conv / MeanShift / ResBlock .forward (model/basic.py:4-17,48-52) incl. G.embed(x), vgg.sub_mean(x)."""
    from pesr_b200.model import VGG, Conv, Generator, MeanShift, ResBlock
    g = torch.Generator().manual_seed(0)
    torch.manual_seed(0)
    x3 = (torch.rand(2, 3, 20, 13, generator=g) * 255).cuda()
    x64 = torch.randn(2, 64, 20, 13, generator=g).cuda()
    # Conv 64 -> 128, forward and backward
    conv = Conv(64, 128, 3).cuda()
    xin = x64.clone().requires_grad_(True)
    y = conv(xin)
    R = torch.randn(y.shape, generator=g).cuda()
    (y * R).sum().backward()
    xo = _q(x64).requires_grad_(True)
    wo, bo = _q(conv.weight.detach()).requires_grad_(True), conv.bias.detach().double().requires_grad_(True)
    yo = F.conv2d(xo, wo, bo, padding=1)
    (yo * R.double()).sum().backward()
    assert y.shape == yo.shape and rel_l2(y, yo) < 1e-5
    assert rel_l2(xin.grad, xo.grad) < 2e-3 and rel_l2(conv.weight.grad, wo.grad) < 2e-3 and rel_l2(conv.bias.grad, bo.grad) < 2e-3
    # stride 2 and a 3-channel output (forward)
    c2 = Conv(64, 64, 3, stride=2).cuda()
    with torch.no_grad():
        assert rel_l2(c2(x64), F.conv2d(_q(x64), _q(c2.weight), c2.bias.double(), stride=2, padding=1)) < 1e-5
        c3 = Conv(64, 3, 3).cuda()
        assert rel_l2(c3(x64), F.conv2d(_q(x64), _q(c3.weight), c3.bias.double(), padding=1)) < 1e-5
    # MeanShift, forward and backward (weights stay trainable, model/basic.py:17)
    ms = MeanShift(255, (0.4488, 0.4371, 0.4040), (1.0, 1.0, 1.0)).cuda()
    xin3 = x3.clone().requires_grad_(True)
    y = ms(xin3)
    R3 = torch.randn(y.shape, generator=g).cuda()
    (y * R3).sum().backward()
    xo = x3.double().requires_grad_(True)
    wo, bo = ms.weight.detach().double().requires_grad_(True), ms.bias.detach().double().requires_grad_(True)
    yo = F.conv2d(xo, wo, bo)
    (yo * R3.double()).sum().backward()
    assert rel_l2(y, yo) < 1e-6 and rel_l2(xin3.grad, xo.grad) < 1e-6
    assert rel_l2(ms.weight.grad, wo.grad) < 1e-4 and rel_l2(ms.bias.grad, bo.grad) < 1e-4
    # ResBlock and the module-attribute calls the reference allows
    rb = ResBlock(64, 3, res_scale=0.1).cuda()
    with torch.no_grad():
        t = F.relu(F.conv2d(_q(x64), _q(rb.body[0].weight), rb.body[0].bias.double(), padding=1))
        ref = F.conv2d(_q(t.float()), _q(rb.body[2].weight), rb.body[2].bias.double(), padding=1) * 0.1 + x64.double()
        assert rel_l2(rb(x64), ref) < 1e-4
        G = Generator({'depth': 1, 'num_channels': 64, 'res_scale': 0.1}).cuda()
        e = G.embed(G.sub_mean(x3))
        ref = F.conv2d(_q(F.conv2d(x3.double(), G.sub_mean.weight.double(), G.sub_mean.bias.double()).float()),
                       _q(G.embed.weight), G.embed.bias.double(), padding=1)
        assert e.shape == (2, 64, 20, 13) and rel_l2(e, ref) < 1e-4
        V = VGG(pretrained=False).cuda()
        assert rel_l2(V.sub_mean(x3), F.conv2d(x3.double(), V.sub_mean.weight.double(), V.sub_mean.bias.double())) < 1e-6


def test_discriminator_batch_larger_than_16():
    from oracle import pesr_oracle as O
    from pesr_b200.model import Discriminator
    opt = {'patch_size': 8, 'spectral_norm': False}
    sd = O.init_discriminator(opt, 0)
    D = Discriminator(opt)
    D.load_state_dict(sd)
    D = D.cuda().train()
    g = torch.Generator().manual_seed(1)
    nb = 40
    x = torch.rand(nb, 3, 32, 32, generator=g) * 255
    R = torch.randn(nb, 1, generator=g)
    y = D(x.cuda())
    (y * R.cuda()).sum().backward()
    yo = O.discriminator_forward({k: v.double() for k, v in sd.items()}, x.double())
    assert y.shape == (nb, 1) and rel_l2(y.detach().cpu(), yo) < 6e-3
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in D.parameters())
    # the classifier gradients are exact functions of the saved activations: check the chunked Linear kernels directly
    pl = D.engine().pools[(nb, 32, 32, 1)][0]
    h1, flat7 = pl.h1_32.double(), pl.flat7.double()
    dlog = R.cuda().double()
    w2 = D.classifier[2].weight.detach().half().double()
    dz1 = (dlog @ w2) * torch.where(h1 > 0, 1.0, 0.2)
    assert rel_l2(D.classifier[2].weight.grad, dlog.t() @ pl.h1_16.double()) < 1e-5
    assert rel_l2(D.classifier[0].weight.grad, dz1.t() @ flat7) < 1e-5
    assert rel_l2(D.classifier[0].bias.grad, dz1.sum(0)) < 1e-5


def test_two_live_generator_graphs_of_one_shape():
    """A forward of the training shape between G(lr) and .backward() (e.g. a validation batch) must not clobber the
    saved activations (the reference keeps any number of graphs alive)."""
    from oracle import pesr_oracle as O
    from pesr_b200.model import Generator
    opt = {'depth': 2, 'num_channels': 64, 'res_scale': 0.1}
    sd = O.init_generator(opt, 0)
    G = Generator(opt)
    G.load_state_dict(sd)
    G = G.cuda()
    g = torch.Generator().manual_seed(2)
    a = (torch.rand(2, 3, 12, 12, generator=g) * 255).cuda()
    b = (torch.rand(2, 3, 12, 12, generator=g) * 255).cuda()
    R = torch.randn(2, 3, 48, 48, generator=g).cuda()

    def grads_of(x):
        for p in G.parameters():
            p.grad = None
        (G(x) * R).sum().backward()
        return [p.grad.clone() for p in G.parameters()]
    ga_alone = grads_of(a)
    for p in G.parameters():
        p.grad = None
    sa = G(a)                     # graph 1 alive ...
    sb = G(b)                     # ... graph 2 of the same shape
    with torch.no_grad():
        G(b)                      # and an inference forward in between
    (sa * R).sum().backward()
    ga = [p.grad.clone() for p in G.parameters()]
    for u, v in zip(ga, ga_alone):
        assert torch.allclose(u, v, rtol=1e-5, atol=1e-6 * float(v.abs().max()))   # fp32 atomics in the bias-gradient sums
    (sb * R).sum().backward()     # accumulates into .grad
    with pytest.raises(RuntimeError, match="live autograd graphs"):
        keep = [G(a) for _ in range(6)]
        del keep


def test_plan_cache_is_bounded_over_many_image_sizes():
    from pesr_b200.model import Generator
    G = Generator({'depth': 2, 'num_channels': 64, 'res_scale': 0.1}).cuda().eval()
    torch.cuda.synchronize()
    peak = []
    with torch.no_grad():
        for i in range(14):
            h, w = 40 + 3 * i, 64 - 2 * i
            y = G(torch.rand(1, 3, h, w, device="cuda") * 255)
            assert y.shape == (1, 3, 4 * h, 4 * w)
            del y
            torch.cuda.synchronize()
            peak.append(torch.cuda.memory_allocated())
    eng = G.engine()
    assert len(eng.plans) <= 5
    assert max(peak[8:]) <= 1.25 * max(peak[:6])      # no monotone growth once the cache is full


def test_generator_with_192_channels():
    """num_channels a multiple of 64 that is not a power of two (GEMM-N tiles of 64)."""
    from oracle import pesr_oracle as O
    from pesr_b200.model import Generator
    opt = {'depth': 1, 'num_channels': 192, 'res_scale': 0.1}
    sd = O.init_generator(opt, 4)
    G = Generator(opt)
    G.load_state_dict(sd)
    G = G.cuda()
    g = torch.Generator().manual_seed(5)
    lr = torch.rand(1, 3, 9, 14, generator=g) * 255
    R = torch.randn(1, 3, 36, 56, generator=g)
    sr = G(lr.cuda())
    (sr * R.cuda()).sum().backward()
    leaf = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    osr = O.generator_forward(leaf, lr.double(), 1, 0.1, qdtype=torch.float16)
    og = dict(zip(leaf, torch.autograd.grad((osr * R.double()).sum(), list(leaf.values()))))
    assert rel_l2(sr.detach().cpu(), osr.detach()) < 1e-3
    errs = sorted(rel_l2(p.grad.cpu(), og[k]) for k, p in G.named_parameters())
    assert errs[len(errs) // 2] < 2e-3 and errs[-1] < 1e-2


def test_adam_state_dict_interchange_and_capturable_mode():
    from pesr_b200.optim import Adam
    g = torch.Generator().manual_seed(6)
    shapes = [(64, 3, 3, 3), (64,), (130, 7), (5,)]
    init = [torch.randn(*s, generator=g) for s in shapes]
    grads = [[torch.randn(*s, generator=g) for s in shapes] for _ in range(5)]

    def make(cls, **kw):
        ps = [torch.nn.Parameter(t.clone().cuda()) for t in init]
        return ps, cls(ps, lr=1e-3, betas=(0.9, 0.999), **kw)

    def run(ps, opt, steps):
        for gs in steps:
            for p, gr in zip(ps, gs):
                p.grad = gr.clone().cuda()
            opt.step()
    p_ref, o_ref = make(torch.optim.Adam)
    run(p_ref, o_ref, grads)
    # ours for 2 steps -> state into torch.optim.Adam for 3 more; torch for 2 -> state into ours for 3 more
    p_a, o_a = make(Adam)
    run(p_a, o_a, grads[:2])
    p_b, o_b = make(torch.optim.Adam)
    with torch.no_grad():
        for q, p in zip(p_b, p_a):
            q.copy_(p)
    o_b.load_state_dict(o_a.state_dict())
    run(p_b, o_b, grads[2:])
    p_c, o_c = make(torch.optim.Adam)
    run(p_c, o_c, grads[:2])
    p_d, o_d = make(Adam)
    with torch.no_grad():
        for q, p in zip(p_d, p_c):
            q.copy_(p)
    run(p_d, o_d, grads[:1])                 # builds (and caches) a pointer table against moment buffers ...
    with torch.no_grad():
        for q, p in zip(p_d, p_c):
            q.copy_(p)
    o_d.load_state_dict(o_c.state_dict())    # ... that load_state_dict replaces: the cache must not survive
    run(p_d, o_d, grads[2:])
    p_e, o_e = make(Adam, capturable=True)   # device-side step count
    run(p_e, o_e, grads)
    for ref, b, d, e in zip(p_ref, p_b, p_d, p_e):
        assert torch.allclose(b, ref, rtol=0, atol=1e-6)      # a few fp32 ulps of |p| ~ 2 over five steps (operation order)
        assert torch.allclose(d, ref, rtol=0, atol=1e-6)
        assert torch.allclose(e, ref, rtol=0, atol=1e-6)
    assert int(o_e.state[p_e[0]]['step']) == 5


@pytest.mark.parametrize("workload", ["pretrain", "gan"])
def test_cuda_graph_replay_equals_eager_step(workload):
    """GraphedStep (pesr_b200/graph.py): three replayed steps give the same losses and parameters as three eager steps."""
    from oracle import pesr_oracle as O
    from pesr_b200 import steps
    from pesr_b200.graph import GraphedStep
    from pesr_b200.model import VGG, Discriminator, Generator
    from pesr_b200.optim import Adam
    opt = {'depth': 2, 'num_channels': 64, 'res_scale': 0.1, 'patch_size': 12, 'spectral_norm': False}
    g_sd, d_sd, v_sd = O.init_generator(opt, 0), O.init_discriminator(opt, 1), O.init_vgg(2)
    gen = torch.Generator().manual_seed(7)
    nb = 4
    batches = [((torch.rand(nb, 3, 12, 12, generator=gen) * 255).cuda(), (torch.rand(nb, 3, 48, 48, generator=gen) * 255).cuda())
               for _ in range(6)]

    def build():
        G, D, V = Generator(opt), Discriminator(opt), VGG(pretrained=False)
        G.load_state_dict(g_sd), D.load_state_dict(d_sd), V.load_state_dict(v_sd)
        G, D, V = G.cuda(), D.cuda(), V.cuda()
        oG, oD = Adam(G.parameters(), lr=5e-5, capturable=True), Adam(D.parameters(), lr=5e-5, capturable=True)
        cfg = dict(steps.DEFAULT_GAN_CFG, target_real=torch.ones(nb, 1, device="cuda"), target_fake=torch.zeros(nb, 1, device="cuda"))
        if workload == "gan":
            return (G, D, V), (oG, oD), lambda lr, hr: steps.gan_step(G, D, V, oG, oD, lr, hr, cfg)
        return (G,), (oG,), lambda lr, hr: steps.pretrain_step(G, oG, lr, hr).reshape(1)
    # eager: 3 warm-up batches (what GraphedStep runs before capturing) + 3 compared steps
    mods_e, _, fn_e = build()
    for lr, hr in batches[:3]:
        fn_e(batches[0][0], batches[0][1])
    eager = [fn_e(lr, hr).clone() for lr, hr in batches[3:]]
    mods_g, opts_g, fn_g = build()
    step = GraphedStep(fn_g, batches[0], modules=mods_g, optimizers=opts_g, warmup=3)
    graphed = [step(lr, hr).clone() for lr, hr in batches[3:]]
    # the Generator-only step is deterministic up to fp32 atomics in the bias-gradient sums; in the GAN step such a
    # 1e-8 difference of a weight is amplified to ~1e-3 of the logits by the 16-bit rounding flips of eight BatchNorm
    # blocks (test_oracle.py::test_rounding_noise_floor) - that is run-to-run noise of ANY two executions, graph or not
    rtol = 1e-5 if workload == "pretrain" else 5e-3
    for a, b in zip(eager, graphed):
        assert torch.allclose(a, b, rtol=rtol, atol=1e-7), (a, b)
    for me, mg in zip(mods_e, mods_g):
        for (k, pe), (_, pg) in zip(me.state_dict().items(), mg.state_dict().items()):
            if pe.is_floating_point():
                assert torch.allclose(pe, pg, rtol=0, atol=2.1e-4), k      # <= 2 Adam sign flips of lr = 5e-5 over 3 steps
    assert all(o.table_builds <= 1 for o in opts_g)
    # an eager forward after replays sees the replayed parameters (the pack caches were invalidated)
    with torch.no_grad():
        a = mods_g[0](batches[0][0])
        mods_e[0].load_state_dict(mods_g[0].state_dict())
        b = mods_e[0](batches[0][0])
    assert torch.equal(a, b)
