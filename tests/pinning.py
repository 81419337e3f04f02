"""Test helpers: read the activations a pesr_b200 engine saved during a forward, as NCHW float64 CPU tensors keyed
the way oracle.pesr_oracle's `pin=` argument expects them (see "forward-pinned evaluation" there).

The oracle then computes the exact gradient of the REFERENCE network at the implementation's own forward point,
which is what the backward kernels have to reproduce; the forward values themselves are gated separately against
the free-running fp64 oracle.
"""
import torch


def _nchw(t, nb, h, w, c):
    return t.view(-1, c)[: nb * h * w].view(nb, h, w, c).permute(0, 3, 1, 2).double().cpu()


def _latest(pool):
    """The most recently used plan instance of a PlanCache pool."""
    return max(pool, key=lambda q: q.stamp) if isinstance(pool, (list, tuple)) else pool


def d_plan_pins(pl, group=0):
    """Pins of call `group` of a (possibly batched, see DiscriminatorEngine.forward) Discriminator forward."""
    nb = pl.nb
    pins = {}
    for i, (hh, ww) in enumerate(pl.dims):
        co = pl.Y[i].shape[1]
        lo, hi = group * nb * hh * ww, (group + 1) * nb * hh * ww
        pins[f'y{i}'] = _nchw(pl.Y[i][lo:hi], nb, hh, ww, co)
        pins[f'a{i}'] = _nchw(pl.A[i][lo:hi], nb, hh, ww, co)
    pins['h1'] = pl.h1_32[group * nb:(group + 1) * nb].double().cpu()
    return pins


def v_plan_pins(pl):
    """Post-ReLU conv outputs of the sr half (the first nb images of every buffer); 'f_hr' = the hr half's features."""
    nb = pl.nb
    pins = {}
    for o in pl.ops:
        if o[0] != "conv":
            continue
        _, li, _cin, cout, ch, cw, _inb, outb, _d = o
        pins[f'c{li}'] = _nchw(outb, nb, ch, cw, cout)
    n_half = nb * pl.fh * pl.fw
    pins['f_hr'] = _nchw(pl.feat16[n_half:], nb, pl.fh, pl.fw, pl.fc)
    return pins


def g_plan_pins(pl, c):
    nb, h, w = pl.nb, pl.h, pl.w
    pins = {}
    for i, x in enumerate(pl.X):
        pins[f'x{i}'] = _nchw(x, nb, h, w, c)
    for i, t in enumerate(pl.T):
        pins[f't{i}'] = _nchw(t, nb, h, w, c)
    pins['u0'] = _nchw(pl.U0, nb, h, w, c)
    pins['u1'] = _nchw(pl.U1, nb, 2 * h, 2 * w, c)
    pins['u2'] = _nchw(pl.U2, nb, 4 * h, 4 * w, c)
    return pins


def discriminator_pins(D, nb, h, w, which=0):
    """Saved tensors of plan `which` (order of acquisition) of the (nb, h, w) pool of D's engine."""
    return d_plan_pins(D.engine().pools[(nb, h, w, 1)][which])


def vgg_pins(V, nb, h, w):
    return v_plan_pins(_latest(V.engine().plans[(nb, h, w)]))


def generator_pins(G, nb, h, w):
    """16-bit conv operands of the last training-mode Generator forward of this shape."""
    return g_plan_pins(_latest(G.engine().plans[(nb, h, w, True)]), G.n_feats)


class StepTracer:
    """Records the pins of every Generator / Discriminator / VGG forward of a training step, in call order, through
    the engines' `trace_hook` (the buffers are reused by later forwards, so they are copied out immediately)."""

    def __init__(self, G, D=None, V=None):
        self.g, self.d, self.v = [], [], []
        self.engines = []
        G.engine().trace_hook = lambda pl: self.g.append(g_plan_pins(pl, G.n_feats)) if pl.train else None
        self.engines.append(G.engine())
        if D is not None:
            D.engine().trace_hook = lambda pl: self.d.extend(d_plan_pins(pl, g) for g in range(pl.groups))
            self.engines.append(D.engine())
        if V is not None:
            V.engine().trace_hook = lambda pl: self.v.append(v_plan_pins(pl))
            self.engines.append(V.engine())

    def close(self):
        for e in self.engines:
            e.trace_hook = None

    def gan_step_pins(self, sr):
        """The `pins=` argument of oracle.gan_step for a step that ran G once, D four times and VGG once."""
        assert len(self.g) == 1 and len(self.d) == 4 and len(self.v) == 1, (len(self.g), len(self.d), len(self.v))
        v = dict(self.v[0])
        f_hr = v.pop('f_hr')
        return {'g': self.g[0], 'd': self.d, 'v': v, 'f_hr': f_hr, 'sr': sr.detach().double().cpu()}


def grad_errors(named_params, oracle_grads, rel_l2):
    """Sorted (error, name) of every parameter gradient against the oracle's."""
    return sorted((rel_l2(p.grad.cpu(), oracle_grads[k]), k) for k, p in named_params)
