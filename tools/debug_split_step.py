"""Which branch of the split-precision GAN step carries the Generator-gradient error?  (bring-up)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pesr_oracle as O  # noqa: E402
from pesr_b200 import steps  # noqa: E402
from pesr_b200.model import VGG, Discriminator, Generator  # noqa: E402
from pesr_b200.optim import Adam  # noqa: E402


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


opt = {'depth': 2, 'num_channels': 64, 'res_scale': 0.1, 'patch_size': 12, 'spectral_norm': False}
nb, patch, lrate = 4, 12, 5e-5
g_sd, d_sd, v_sd = O.init_generator(opt, 0), O.init_discriminator(opt, 1), O.init_vgg(2)
gen = torch.Generator().manual_seed(3)
lr = torch.rand(nb, 3, patch, patch, generator=gen) * 255
hr = torch.rand(nb, 3, patch * 4, patch * 4, generator=gen) * 255
for tag, kw in (("GAN only", dict(alpha_vgg=0.0, alpha_tv=0.0)),):
    lrd = 0.0 if "lr_D" in tag else lrate
    free = O.gan_step(g_sd, d_sd, v_sd, lr, hr, opt, lr_rate=lrd, dtype=torch.float64, **kw)
    G, D, V = Generator(opt, split_precision=True), Discriminator(opt, split_precision=True), VGG(pretrained=False, split_precision=True)
    G.load_state_dict(g_sd), D.load_state_dict(d_sd), V.load_state_dict(v_sd)
    G, D, V = G.cuda(), D.cuda(), V.cuda()
    optG, optD = Adam(G.parameters(), lr=lrate), Adam(D.parameters(), lr=lrd)
    cfg = dict(steps.DEFAULT_GAN_CFG, **kw)
    cfg['target_real'] = torch.ones(nb, 1, device="cuda")
    cfg['target_fake'] = torch.zeros(nb, 1, device="cuda")
    grabbed = {}

    class Spy(torch.nn.Module):
        def __init__(self, net):
            super().__init__()
            self.net = net

        def forward(self, x):
            out = self.net(x)
            grabbed['sr'] = out
            out.register_hook(lambda gr: grabbed.__setitem__('dsr', gr.detach().clone()))
            return out
    got = steps.gan_step(Spy(G), D, V, optG, optD, lr.cuda(), hr.cuda(), cfg).cpu()
    if 'dsr' in grabbed:
        a_, b_ = grabbed['dsr'].cpu().double().flatten(), free['dsr'].double().flatten()
        print(f"   in-step d(loss)/d(sr): rel-L2 {rel_l2(grabbed['dsr'].cpu(), free['dsr']):.2e}, least-squares factor ours/oracle {float((a_ @ b_) / (b_ @ b_)):.6f}; "
              f"sr rel-L2 {rel_l2(grabbed['sr'].detach().cpu(), free['sr']):.2e}")
        for k, p_ in list(G.named_parameters())[:4]:
            a_, b_ = p_.grad.cpu().double().flatten(), free['g_grads'][k].double().flatten()
            print(f"      {k}: factor {float((a_ @ b_) / (b_ @ b_)):.6f}")
    ge = sorted(rel_l2(p.grad.cpu(), free['g_grads'][k]) for k, p in G.named_parameters())
    de = sorted(rel_l2(p.grad.cpu(), free['d_grads'][k]) for k, p in D.named_parameters())
    bad = sum(int(((p.detach().cpu().double() - free['d_params_after'][k].double()).abs() > 0.05 * lrate).sum()) for k, p in D.named_parameters())
    if tag == "GAN only":
        for k, p in D.named_parameters():
            dlt = (p.detach().cpu().double() - free['d_params_after'][k].double()).abs()
            n = int((dlt > 0.05 * lrate).sum())
            g64 = free['d_grads'][k]
            if n:
                idx = (dlt > 0.05 * lrate).nonzero()[:3]
                print(f"   {k}: {n} of {dlt.numel()} differ, max |diff| {float(dlt.max()):.2e}; oracle |grad| there "
                      f"{[float(g64[tuple(i)].abs()) for i in idx]}, ours {[float(p.grad.cpu()[tuple(i)].abs()) for i in idx]}; typical |grad| {float(g64.abs().median()):.2e}")
    if tag == "GAN only":
        # the oracle's Generator phase evaluated with OUR post-step Discriminator weights
        import torch.nn.functional as F
        d_ours = {k: v.detach().cpu().double() for k, v in D.state_dict().items()}
        g = {k: v.double().clone().requires_grad_(True) for k, v in g_sd.items()}
        sr = O.generator_forward(g, lr.double(), opt['depth'], opt['res_scale'])
        pf = O.discriminator_forward(d_ours, sr)
        pr = O.discriminator_forward(d_ours, hr.double())
        gl = O.focal_loss(pf - pr, torch.ones_like(pf), 1.0)
        gg = torch.autograd.grad(gl, list(g.values()))
        e2 = sorted(rel_l2(p.grad.cpu(), gk) for (k, p), gk in zip(G.named_parameters(), gg))
        from pesr_b200 import losses as Ls
        sr_in = sr.detach().float().cuda().requires_grad_(True)
        for pp in D.parameters():
            pp.requires_grad = False
        a, b = D.forward_pair(sr_in, hr.cuda())
        l2 = Ls.rsgan_focal(a, b, 1.0, cfg['target_real'])
        l2.backward()
        sro = sr.detach().clone().requires_grad_(True)
        pf2 = O.discriminator_forward(d_ours, sro)
        pr2 = O.discriminator_forward(d_ours, hr.double())
        gl2 = O.focal_loss(pf2 - pr2, torch.ones_like(pf2), 1.0)
        dsro, = torch.autograd.grad(gl2, sro)
        def d_alone(tag2):
            srx = sr.detach().float().cuda().requires_grad_(True)
            ax, bx = D.forward_pair(srx, hr.cuda())
            Ls.rsgan_focal(ax, bx, 1.0, 1.0).backward()
            print(f"   stepped D instance, {tag2}: d/d(sr) {rel_l2(srx.grad.cpu(), dsro):.2e}")
        eng = D.engine()
        stale = [n for n, sw in eng.packed.items() if sw.key != (sw.param.data_ptr(), sw.param._version)]
        print("   packs whose key is not the parameter's current (ptr, version):", stale)
        d_alone("as left by the step")
        for n, sw in eng.packed.items():
            if n.endswith("_d"):
                sw.key = None
        d_alone("after invalidating the backward-data packs")
        eng.invalidate_packs()
        d_alone("after invalidating every pack")
        sro5 = grabbed['sr'].detach().cpu().double().clone().requires_grad_(True)
        pf5 = O.discriminator_forward(d_ours, sro5)
        d5, = torch.autograd.grad(O.focal_loss(pf5 - pr2.detach(), torch.ones_like(pf5), 1.0), sro5)
        print(f"   oracle D (our post-step weights) at OUR in-step sr vs our in-step d/d(sr): {rel_l2(grabbed['dsr'].cpu(), d5):.2e}; "
              f"oracle at our sr vs oracle at its own sr: {rel_l2(d5, dsro):.2e}; sr std per image {[round(float(v), 4) for v in grabbed['sr'].detach().flatten(1).std(dim=1).cpu()]}")
        D2 = Discriminator(opt, split_precision=True)
        D2.load_state_dict(D.state_dict())
        D2 = D2.cuda().train()
        for pp in D2.parameters():
            pp.requires_grad = False
        sr2 = sr.detach().float().cuda().requires_grad_(True)
        a2, b2 = D2.forward_pair(sr2, hr.cuda())
        Ls.rsgan_focal(a2, b2, 1.0, cfg['target_real']).backward()
        print(f"   a FRESH split Discriminator with the same weights: d/d(sr) {rel_l2(sr2.grad.cpu(), dsro):.2e}")
        D3 = Discriminator(opt)
        D3.load_state_dict(D.state_dict())
        D3 = D3.cuda().train()
        for pp in D3.parameters():
            pp.requires_grad = False
        sr3 = sr.detach().float().cuda().requires_grad_(True)
        a3, b3 = D3.forward_pair(sr3, hr.cuda())
        Ls.rsgan_focal(a3, b3, 1.0, cfg['target_real']).backward()
        print(f"   a fresh 16-bit Discriminator with the same weights: d/d(sr) {rel_l2(sr3.grad.cpu(), dsro):.2e}")
        Rr = torch.randn(nb, 1, generator=torch.Generator().manual_seed(9))
        for mode in ("single", "pair"):
            for wts in ("random", "focal"):
                srx = sr.detach().float().cuda().requires_grad_(True)
                if mode == "single":
                    ax = D2(srx)
                    bx = D2(hr.cuda()).detach()
                else:
                    ax, bx = D2.forward_pair(srx, hr.cuda())
                srox = sr.detach().clone().requires_grad_(True)
                px = O.discriminator_forward(d_ours, srox)
                if wts == "random":
                    (ax * Rr.cuda()).sum().backward()
                    dref, = torch.autograd.grad((px * Rr.double()).sum(), srox)
                else:
                    Ls.rsgan_focal(ax, bx, 1.0, 1.0).backward()
                    dref, = torch.autograd.grad(O.focal_loss(px - pr2.detach(), torch.ones_like(px), 1.0), srox)
                print(f"   fresh split D [{mode} call, {wts} logit gradient]: d/d(sr) {rel_l2(srx.grad.cpu(), dref):.2e}")
        sr4 = sr.detach().float().cuda().requires_grad_(True)
        a4 = D2(sr4)
        (a4 * torch.randn(nb, 1, generator=torch.Generator().manual_seed(9)).cuda()).sum().backward()
        sro4 = sr.detach().clone().requires_grad_(True)
        p4 = O.discriminator_forward(d_ours, sro4)
        d4, = torch.autograd.grad((p4 * torch.randn(nb, 1, generator=torch.Generator().manual_seed(9)).double()).sum(), sro4)
        print(f"   fresh split D, single call, random logit weights: d/d(sr) {rel_l2(sr4.grad.cpu(), d4):.2e}")
        print(f"   D alone on the post-step weights: logits fake {rel_l2(a.detach().cpu(), pf2.detach()):.2e} real {rel_l2(b.detach().cpu(), pr2.detach()):.2e} "
              f"(values {a.detach().cpu().flatten().tolist()} vs {pf2.detach().flatten().tolist()}), loss {float(l2):.6f} vs {float(gl2):.6f}, d/d(sr) {rel_l2(sr_in.grad.cpu(), dsro):.2e}")
        print(f"   vs the oracle's G phase evaluated with OUR post-step D weights: G grads median {e2[len(e2) // 2]:.2e} max {e2[-1]:.2e}")
    print(f"{tag}: G grads median {ge[len(ge) // 2]:.2e} max {ge[-1]:.2e}; D grads median {de[len(de) // 2]:.2e}; D params after Adam "
          f"differing by > 5 % of lr: {bad}; losses {[round(float(v), 6) for v in got]} vs {[round(float(free[k]), 6) for k in ('l1', 'vgg', 'g_loss', 'tv', 'd_loss')]}")
