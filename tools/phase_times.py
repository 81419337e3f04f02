"""In-stream attribution of the GAN step (train.py:202-259) to its phases: CUDA events around every engine forward /
backward and optimiser step of eager steps (the GPU stays the bottleneck: host enqueue is ~8 ms of a ~17 ms step), summed
per label over N steps.  `python tools/phase_times.py [gan|pretrain] [steps]`"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from pesr_b200 import steps
from pesr_b200.model import VGG, Discriminator, Generator
from pesr_b200.optim import Adam

workload = sys.argv[1] if len(sys.argv) > 1 else "gan"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
DETAIL = len(sys.argv) > 3 and sys.argv[3] == "detail"     # also one event pair per library launch (serialises PDL)
dev = torch.device("cuda", 0)
torch.manual_seed(0)
records = []          # (label, e0, e1)
launches = []         # (phase label, launch label, e0, e1)
phase = ["rest"]


def wrap(obj, name, label):
    fn = getattr(obj, name)

    def timed(*a, **k):
        lab = label(*a, **k) if callable(label) else label
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        prev, phase[0] = phase[0], lab
        out = fn(*a, **k)
        phase[0] = prev
        e1.record()
        records.append((lab, e0, e1))
        return out
    setattr(obj, name, timed)


def wrap_library():
    from pesr_b200 import _lib

    def make(name, fn):
        def timed(*a):
            lab = name[5:]
            if name in ("pesr_conv_igemm", "pesr_conv_wgrad"):
                d = a[0]._obj
                if name == "pesr_conv_igemm":
                    lab += f" {d.nb}x{d.h}x{d.w} {d.cin}->{d.cout} taps{d.ntaps}" + (" res32" if d.res32 else "") + \
                        (" mask" if d.mask16 else "") + (f" cls{d.ncls}" if d.ncls > 1 else "")
                else:
                    lab += f" {d.nb}x{d.h}x{d.w} M{d.m_total} N{d.n_total} taps{d.ntaps}"
            elif name in ("pesr_bn_reduce", "pesr_bn_lrelu_fwd"):
                lab += f" npix{a[1]} c{a[2]} groups{a[3]}"
            elif name == "pesr_bn_lrelu_bwd":
                lab += f" npix{a[2]} c{a[3]} groups{a[4]}"
            elif name in ("pesr_maxpool2_fwd", "pesr_maxpool2_bwd"):
                lab += " " + "x".join(str(v) for v in a if isinstance(v, int) and 0 < v < 100000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a)
            e1.record()
            launches.append((phase[0], lab, e0, e1))
            return r
        return timed
    for name in _lib.SIGNATURES:
        if name.startswith("pesr_") and name not in ("pesr_version", "pesr_last_error", "pesr_launch_count"):
            try:
                fn = getattr(_lib.lib, name)
            except AttributeError:
                continue
            setattr(_lib.lib, name, make(name, fn))


if DETAIL:
    wrap_library()


G = Generator(bench.OPT).to(dev)
optG = Adam(G.parameters(), lr=5e-5)
lr = torch.rand(16, 3, 48, 48, device=dev) * 255
hr = torch.rand(16, 3, 192, 192, device=dev) * 255
wrap(G.engine(), "forward", "G forward")
wrap(G.engine(), "backward", "G backward")
wrap(optG, "step", "Adam G")
if workload == "gan":
    D = Discriminator(bench.OPT).to(dev)
    V = VGG(pretrained=False).to(dev)
    optD = Adam(D.parameters(), lr=5e-5)
    cfg = dict(steps.DEFAULT_GAN_CFG)
    cfg['target_real'] = torch.ones(16, 1, device=dev)
    cfg['target_fake'] = torch.zeros(16, 1, device=dev)
    wrap(D.engine(), "forward", "D forward (pair, 32 images)")
    wrap(D.engine(), "backward", lambda st, dl, need_param_grads, need_input_grad:
         "D backward (params, 32 images)" if need_param_grads else "D backward (d/d sr, 16 images)")
    wrap(V.engine(), "forward", "VGG forward (32 images)")
    wrap(V.engine(), "backward", "VGG backward (16 images)")
    wrap(optD, "step", "Adam D")

    def step():
        return steps.gan_step(G, D, V, optG, optD, lr, hr, cfg)
else:
    def step():
        return steps.pretrain_step(G, optG, lr, hr)
for _ in range(4):
    step()
torch.cuda.synchronize()
records.clear()
launches.clear()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(nsteps):
    step()
t1.record()
torch.cuda.synchronize()
total = t0.elapsed_time(t1) / nsteps
acc = collections.OrderedDict()
for lab, e0, e1 in records:
    acc[lab] = acc.get(lab, 0.0) + e0.elapsed_time(e1) / nsteps
print(f"{workload} step, eager, {nsteps} steps: {total:.3f} ms per step")
s = 0.0
for lab, ms in acc.items():
    print(f"  {lab:38s} {ms:7.3f} ms  {100 * ms / total:5.1f} %")
    s += ms
print(f"  {'losses, autograd glue, rest':38s} {total - s:7.3f} ms  {100 * (total - s) / total:5.1f} %")
if DETAIL:
    tab = collections.OrderedDict()
    for ph, lab, e0, e1 in launches:
        k = (ph, lab)
        c, t = tab.get(k, (0, 0.0))
        tab[k] = (c + 1, t + e0.elapsed_time(e1) * 1e3)
    cur = None
    for (ph, lab), (c, t) in tab.items():
        if ph != cur:
            print(f"--- {ph}")
            cur = ph
        print(f"    {lab:64s} x{c / nsteps:5.1f}  {t / nsteps:8.1f} us per step  ({t / c:6.1f} us each)")
