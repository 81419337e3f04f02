"""Split-precision schedule of the Discriminator (reference: model/pesr.py:40-81, model/basic.py:19-31): every
convolution and the two Linear layers run as three tensor-core passes on fp16 hi + lo operand pairs (engine_split.py),
activations between the layers are fp32, and train-mode BatchNorm / LeakyReLU act on fp32 (statistics in fp64).

Same interface as engine_d.DiscriminatorEngine (forward of one call or of a list of calls executed as one batch with
per-call BatchNorm statistics; backward with parameter and / or input gradients), no plan pooling: buffers are allocated
per call.  This is the precision demonstration asked for by the north_star's 1e-3 gradient tolerance, not the headline
path: `Discriminator(opt, split_precision=True)`.
"""
import torch

from . import ops
from .engine_d import D_LAYERS, _parity_planes, _s2_taps
from .engine_g import FlatGrads
from .engine_split import HL, SplitOps, SplitWeight
from .ops import ACT_LRELU


import os

_WSCALE = float(os.environ.get("PESR_SPLIT_WSCALE", "256"))     # power of two applied to the conv weight operands (engine_split)


class _SplitLinearWeight:
    """hi / lo 16-bit copies of a Linear weight [out][in], refreshed when the parameter changes."""

    def __init__(self, param):
        self.param = param
        self.hi = torch.empty(param.shape, device=param.device, dtype=torch.float16)
        self.lo = torch.empty(param.shape, device=param.device, dtype=torch.float16)
        self.key = None

    def get(self):
        p = self.param
        key = (p.data_ptr(), p._version)
        if key != self.key:
            ops.split16(p.detach().contiguous(), self.hi, self.lo)
            self.key = key
        return self.hi, self.lo


class SplitDiscriminatorEngine(SplitOps):
    def __init__(self, disc):
        self.disc = disc
        self.dtype = torch.float16
        self.packed = None
        self.device = None
        self.param_list = None
        self.grad_hook = None
        self.grad_hook_finish = None
        self.defer_finish = False
        self.last_flat = None
        self.trace_hook = None

    def invalidate_packs(self):
        if self.packed is not None:
            for sw in self.packed.values():
                sw.key = None
            self.fc1.key = self.fc2.key = None

    def _ensure_packed(self, device):
        d = self.disc
        sentinel = (d.features[0][0].weight.data_ptr(), d.classifier[2].bias.data_ptr())
        if self.packed is not None and self.device == device and self.sentinel == sentinel:
            return
        self.sentinel, self.device = sentinel, device
        w0 = d.features[0][0].weight
        ws = _WSCALE
        pk = {"c0_f": SplitWeight(w0, 4, pad_to=64, scale=ws), "c0_d": SplitWeight(w0, 6, pad_to=32, scale=ws)}
        for i in range(1, 8):
            w = d.features[i][0].weight
            pk[f"c{i}_f"], pk[f"c{i}_d"] = SplitWeight(w, 0, scale=ws), SplitWeight(w, 1, scale=ws)
        self.packed = pk
        self.fc1, self.fc2 = _SplitLinearWeight(d.classifier[0].weight), _SplitLinearWeight(d.classifier[2].weight)
        self.wg = torch.empty(3 * max(9 * 512 * 512 * 4, 148 * 128 * 64), device=device, dtype=torch.float32)
        self.scale_ws = torch.zeros(4, device=device, dtype=torch.float32)
        self.flat_grads = FlatGrads(self.param_list)
        self.offsets, self.flat_numel = self.flat_grads.offsets, self.flat_grads.numel

    @staticmethod
    def _geometry(h, w):
        dims = []
        for (_ci, _co, s) in D_LAYERS:
            if s == 2:
                if h % 2 or w % 2:
                    raise ValueError("split-precision Discriminator: feature maps must stay even (patch sizes multiple of 4)")
                h, w = h // 2, w // 2
            dims.append((h, w))
        return dims

    # ------------------------------------------------------------------ forward
    def forward(self, xs, save):
        d = self.disc
        single = torch.is_tensor(xs)
        xs = [xs] if single else list(xs)
        for x in xs:
            if x.dim() != 4 or x.shape[1] != 3:
                raise ValueError(f"Discriminator expects [N,3,H,W], got {tuple(x.shape)}")
            if x.shape != xs[0].shape:
                raise ValueError("Discriminator: the calls of one batched forward must have one shape")
        xs = [x.contiguous().float() for x in xs]
        G = len(xs)
        nb, _, h, w = xs[0].shape
        dev = xs[0].device
        self._ensure_packed(dev)
        nt = G * nb
        dims = self._geometry(h, w)
        training = d.training
        e32 = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)  # noqa: E731
        P = nb * h * w
        col0 = HL(nt * h * w, 64, dev)
        for g, x in enumerate(xs):
            ops.im2col3(x, col0.hi[g * P:(g + 1) * P])
            ops.im2col3(x, col0.lo[g * P:(g + 1) * P], low_part=True)
        taps2, srcs2, widx2 = _s2_taps()
        A, Y, stats = [], [], []
        cur = col0
        for i, (ci, co, s) in enumerate(D_LAYERS):
            hh, ww = dims[i]
            y = e32(nt * hh * ww, co)
            if i == 0:
                self._conv3(cur, "c0_f", nt, hh, ww, 64, co, y, taps=[(0, 0)], alpha=1.0 / _WSCALE)
            elif s == 1:
                self._conv3(cur, f"c{i}_f", nt, hh, ww, ci, co, y, alpha=1.0 / _WSCALE)
            else:
                hi_, wi_ = dims[i - 1]
                self._conv3(cur, f"c{i}_f", nt, hh, ww, ci, co, y, alpha=1.0 / _WSCALE, taps=taps2, tap_src=srcs2, tap_widx=widx2,
                            srcs_fn=lambda t, hi_=hi_, wi_=wi_, ci=ci: _parity_planes(t, nt, hi_, wi_, ci))
            bn = d.features[i][1]
            npix = nb * hh * ww
            a = HL(nt * hh * ww, co, dev)
            gamma, beta = bn.weight.detach().double(), bn.bias.detach().double()
            st = []
            for g in range(G):
                rows = slice(g * npix, (g + 1) * npix)
                if training:
                    sums = torch.zeros(2, co, device=dev, dtype=torch.float64)
                    ops.colmoments32(y[rows], npix, co, sums)
                    mean = sums[0] / npix
                    var = (sums[1] / npix - mean * mean).clamp_min(0)
                    with torch.no_grad():
                        bn.running_mean.mul_(1 - bn.momentum).add_((bn.momentum * mean).float())
                        bn.running_var.mul_(1 - bn.momentum).add_((bn.momentum * var * (npix / max(npix - 1, 1))).float())
                        bn.num_batches_tracked += 1
                else:
                    mean, var = bn.running_mean.double(), bn.running_var.double()
                rstd = torch.rsqrt(var + bn.eps)
                ka = (gamma * rstd).float()
                kc = (beta - mean * gamma * rstd).float()
                ops.affine_split(y[rows], npix, co, ka=ka, kc=kc, act=ACT_LRELU, hi=a.hi[rows], lo=a.lo[rows])
                st.append((mean, rstd))
            A.append(a)
            Y.append(y)
            stats.append(st)
            cur = a
        h7, w7 = dims[7]
        kfc = 512 * h7 * w7
        fc1, fc2 = d.classifier[0], d.classifier[2]
        if kfc != fc1.in_features:
            raise RuntimeError(f"pesr_b200.Discriminator: a {h}x{w} input gives {kfc} features but classifier.0 expects "
                               f"{fc1.in_features} (patch_size {d.patch_size}): size mismatch")
        flat = HL(nt, kfc, dev)
        ops.flatten_nchw16(A[7].hi, nt, h7 * w7, 512, flat.hi)
        ops.flatten_nchw16(A[7].lo, nt, h7 * w7, 512, flat.lo)
        # Linear(kfc -> 1024): three split-K passes of the implicit-GEMM kernel, all partials summed by one finalisation
        w1h, w1l = self.fc1.get()
        ks = max(1, min(kfc // 64 // 8, 36))
        part = e32(3 * ks * nt * 1024)
        for k, (xs_, ws_) in enumerate(((flat.hi, w1h), (flat.lo, w1h), (flat.hi, w1l))):
            dsc = ops.make_conv_desc(dtype=0, nb=1, h=1, w=nt, cin=kfc, cout=1024, block_n=256, taps=[(0, 0)],
                                     srcs=[ops.nhwc_src(xs_, 1, 1, nt, kfc)], wpacked=ws_, out32=part[k * ks * nt * 1024:],
                                     ld_out32=1024, ksplit=ks, split_stride32=nt * 1024)
            ops.conv_igemm(dsc)
        h1_32 = e32(nt, 1024)
        ops.linear_finalize(part, 3 * ks, nt, 1024, fc1.bias.detach(), torch.float16, out32=h1_32, act=ACT_LRELU)
        h1 = HL(nt, 1024, dev)
        ops.split16(h1_32, h1.hi, h1.lo)
        w2h, w2l = self.fc2.get()
        ws = e32(max(ops.linear_workspace_floats(16, 1024, 1), 1))
        parts = [e32(nt, 1) for _ in range(3)]
        ops.linear_fwd(h1.hi, w2h, fc2.bias.detach(), nt, 1024, 1, ws, out32=parts[0])
        ops.linear_fwd(h1.lo, w2h, None, nt, 1024, 1, ws, out32=parts[1])
        ops.linear_fwd(h1.hi, w2l, None, nt, 1024, 1, ws, out32=parts[2])
        logits = parts[0] + parts[1] + parts[2]
        outs = logits if single else list(logits.split(nb))
        if save:
            if not training:
                raise NotImplementedError("pesr_b200.Discriminator: backward in eval() mode is not on the PESR path")
            return outs, dict(G=G, nb=nb, h=h, w=w, dims=dims, col0=col0, A=A, Y=Y, stats=stats, flat=flat, h1_32=h1_32, h1=h1,
                              kfc=kfc)
        return outs, None

    # ------------------------------------------------------------------ backward
    def backward(self, state, dlogits, need_param_grads, need_input_grad):
        st = state
        d = self.disc
        if self.grad_hook is not None:
            raise NotImplementedError("pesr_b200: the split-precision schedules are single-GPU (no bucketed all-reduce hooks)")
        G, nb, h, w, dims, kfc = st['G'], st['nb'], st['h'], st['w'], st['dims'], st['kfc']
        single = torch.is_tensor(dlogits)
        dl = [dlogits] if single else list(dlogits)
        need_in = [need_input_grad] * G if isinstance(need_input_grad, bool) else list(need_input_grad)
        gsel = G if need_param_grads else max((g + 1 for g in range(G) if need_in[g]), default=0)
        if gsel == 0:
            return {}, [None] * G
        nbb = gsel * nb
        dev = st['h1_32'].device
        e32 = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)  # noqa: E731
        parts = [(dl[g].contiguous().float() if dl[g] is not None else torch.zeros(nb, 1, device=dev)) for g in range(gsel)]
        dlogits = parts[0] if gsel == 1 else torch.cat(parts, dim=0)
        fc1, fc2 = d.classifier[0], d.classifier[2]
        grads, flat_buf = {}, None
        if need_param_grads:
            flat_buf = self.flat_grads.get(dev)
            off = self.offsets
            grads = {p: flat_buf[off[p]:off[p] + p.numel()].view(p.shape) for p in self.param_list}
        h1_32, h1, flat = st['h1_32'][:nbb], st['h1'], st['flat']
        w1h, w1l = self.fc1.get()
        w2h, w2l = self.fc2.get()
        # ---- classifier: fp32 gradients against 16-bit hi + lo operands (exact products, fp32 accumulation)
        if need_param_grads:
            ops.linear_wgrad(dlogits, h1.hi[:nbb], nbb, 1024, 1, grads[fc2.weight])
            ops.linear_wgrad(dlogits, h1.lo[:nbb], nbb, 1024, 1, grads[fc2.weight], accumulate=True)
            torch.sum(dlogits, dim=0, out=grads[fc2.bias])
        dh1, dh1b = e32(nbb, 1024), e32(nbb, 1024)
        ops.linear_dgrad(dlogits, w2h, nbb, 1024, 1, dh1)
        ops.linear_dgrad(dlogits, w2l, nbb, 1024, 1, dh1b)
        dz1 = (dh1 + dh1b) * torch.where(h1_32 > 0, 1.0, 0.2)
        if need_param_grads:
            ops.linear_wgrad(dz1, flat.hi[:nbb], nbb, kfc, 1024, grads[fc1.weight])
            ops.linear_wgrad(dz1, flat.lo[:nbb], nbb, kfc, 1024, grads[fc1.weight], accumulate=True)
            torch.sum(dz1, dim=0, out=grads[fc1.bias])
        dflat, dflat_b = e32(nbb, kfc), e32(nbb, kfc)
        ops.linear_dgrad(dz1, w1h, nbb, kfc, 1024, dflat)
        ops.linear_dgrad(dz1, w1l, nbb, kfc, 1024, dflat_b)
        dflat += dflat_b
        # every 16-bit gradient operand below carries `scale`, a power of two chosen from max|gradient| and RENEWED at
        # every layer: the lo halves are 2^-11 of the hi halves, so a tensor whose values sit far below the fp16 normal
        # range would silently lose them (gradients shrink and grow by the BatchNorm factors gamma * rstd layer by layer)
        ops.amax_scale(dflat, self.scale_ws, target=16.0)
        scale = self.scale_ws[1:2].clone()
        h7, w7 = dims[7]
        dA = (dflat * scale).view(nbb, 512, h7 * w7).permute(0, 2, 1).reshape(nbb * h7 * w7, 512).contiguous()
        taps2, srcs2, _ = _s2_taps()
        for i in range(7, -1, -1):
            ci, co, s = D_LAYERS[i]
            hh, ww = dims[i]
            npix = nb * hh * ww
            bn, conv = d.features[i][1], d.features[i][0]
            a, y = st['A'][i], st['Y'][i]
            gamma = bn.weight.detach().double()
            dY = HL(nbb * hh * ww, co, dev)
            dZ = e32(nbb * hh * ww, co)
            dgam = torch.zeros(co, device=dev, dtype=torch.float64)
            dbet = torch.zeros(co, device=dev, dtype=torch.float64)
            for g in range(gsel):
                rows = slice(g * npix, (g + 1) * npix)
                mean, rstd = st['stats'][i][g]
                # dZ = dA * lrelu'(activation); BatchNorm backward (model/basic.py:29) as one fp64 reduction + one affine pass
                ops.affine_split(dA[rows], npix, co, out32=dZ[rows], mask_hi=a.hi[rows], mask_lo=a.lo[rows], mask_mode=2)
                sums = torch.zeros(2, co, device=dev, dtype=torch.float64)
                ops.colmoments32(dZ[rows], npix, co, sums, b32=y[rows])
                s1, s2 = sums[0], sums[1]
                dgamma = rstd * (s2 - mean * s1)
                dgam += dgamma
                dbet += s1
                gr = gamma * rstd
                ka = gr.float()
                kb = (-gr * rstd * dgamma / npix).float()
                kc = (-gr * s1 / npix + gr * rstd * mean * dgamma / npix).float()
                ops.affine_split(dZ[rows], npix, co, b32=y[rows], ka=ka, kb=kb, kc=kc, hi=dY.hi[rows], lo=dY.lo[rows])
            if need_param_grads:
                inv_scale = 1.0 / scale.double()
                grads[bn.weight].copy_(dgam * inv_scale)
                grads[bn.bias].copy_(dbet * inv_scale)
                if i == 0:
                    self._wgrad3(dY, 64, st['col0'], 64, nbb, hh, ww, grads[conv.weight], ops.WMAP_COL_IN, 64, 3, scale,
                                 taps=[(0, 0)])
                elif s == 1:
                    self._wgrad3(dY, co, st['A'][i - 1], ci, nbb, hh, ww, grads[conv.weight], ops.WMAP_OIHW, co, ci, scale)
                else:
                    hi_, wi_ = dims[i - 1]
                    self._wgrad3(dY, co, st['A'][i - 1], ci, nbb, hh, ww, grads[conv.weight], ops.WMAP_OIHW, co, ci, scale,
                                 taps=taps2, tap_src=srcs2,
                                 b_srcs_fn=lambda t, hi_=hi_, wi_=wi_, ci=ci: _parity_planes(t, nbb, hi_, wi_, ci))
            if i == 0:
                break
            hi_, wi_ = dims[i - 1]
            dA = e32(nbb * hi_ * wi_, ci)
            if s == 1:
                self._conv3(dY, f"c{i}_d", nbb, hi_, wi_, co, ci, dA, alpha=1.0 / _WSCALE)
            else:
                # backward-data of a stride-2 conv: four parity classes of the input grid in one launch (engine_d.py)
                all_taps, all_widx, classes = [], [], []
                for ph in range(2):
                    for pw in range(2):
                        ys = [(1, 0)] if ph == 0 else [(0, 1), (2, 0)]
                        xs_ = [(1, 0)] if pw == 0 else [(0, 1), (2, 0)]
                        all_taps += [(oy, ox) for (_dy, oy) in ys for (_dx, ox) in xs_]
                        all_widx += [8 - (dy * 3 + dx) for (dy, _oy) in ys for (dx, _ox) in xs_]
                        classes.append((len(ys) * len(xs_), ph, pw))
                self._conv3(dY, f"c{i}_d", nbb, hi_ // 2, wi_ // 2, co, ci, dA, alpha=1.0 / _WSCALE, taps=all_taps, tap_widx=all_widx,
                            srcs_fn=lambda t, hh=hh, ww=ww, co=co: [ops.nhwc_src(t, nbb, hh, ww, co)],
                            out_h=hi_, out_w=wi_, out_sy=2, out_sx=2, aux_mode=1, classes=classes)
            # renew the range scale for the next layer (a NEW tensor: kernels already enqueued keep reading the old one)
            ops.amax_scale(dA, self.scale_ws, target=16.0)
            f = self.scale_ws[1:2]
            dA.mul_(f)
            scale = scale * f
        dxs = [None] * G
        if any(need_in[:gsel]):
            Zd = self._conv3(dY, "c0_d", nbb, h, w, 64, 32, e32(nbb * h * w, 32), taps=[(0, 0)], alpha=1.0 / _WSCALE)
            dx = torch.empty(nbb, 3, h, w, device=dev, dtype=torch.float32)
            ops.col2im3(Zd, 32, nbb, h, w, dx, mul=1.0, div_dev=scale, sgn=-1)
            for g in range(gsel):
                if need_in[g]:
                    dxs[g] = dx[g * nb:(g + 1) * nb]
        self.last_flat = flat_buf
        return grads, dxs
