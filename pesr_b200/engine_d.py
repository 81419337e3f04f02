"""Kernel schedule of the Discriminator (reference: model/pesr.py:40-81, model/basic.py:19-31).

Eight (conv3x3 no-bias -> train-mode BatchNorm -> LeakyReLU 0.2) blocks, stride 1/2 alternating,
NCHW flatten, Linear(512*s*s -> 1024), LeakyReLU, Linear(1024 -> 1).

  * conv 0 (Cin = 3) is an im2col GEMM; stride-2 convs read the four parity planes of their input
    through four TMA tensor maps (no strided gather); their backward-data is four parity sub-convolutions.
  * BatchNorm uses the batch's own statistics of every call (the reference calls D four times per step on
    different batches, train.py:205,208,237,238) and updates the running statistics each time.
  * backward emits parameter gradients into one flat fp32 buffer, last parameter first (see engine_g).
"""
import ctypes as C

import os

import torch

from . import ops
from ._lib import check, lib
from .engine_g import _WGRAD_STREAM, FlatGrads, PackedWeight, SideLane, _Plan, _Release, _run_conv
from .ops import ACT_LRELU

D_LAYERS = [(3, 64, 1), (64, 64, 2), (64, 128, 1), (128, 128, 2), (128, 256, 1), (256, 256, 2), (256, 512, 1),
            (512, 512, 2)]   # model/pesr.py:53-66


def _s2_taps():
    """Taps of a stride-2, pad-1 3x3 conv expressed on the parity planes of its input."""
    taps, srcs, widx = [], [], []
    for dy in range(3):
        for dx in range(3):
            ph, pw = (dy + 1) % 2, (dx + 1) % 2
            taps.append((-1 if dy == 0 else 0, -1 if dx == 0 else 0))
            srcs.append(ph * 2 + pw)
            widx.append(dy * 3 + dx)
    return taps, srcs, widx


def _parity_planes(t, nb, h, w, c):
    """Four NHWC views (ptr, h2, w2, sn, sh, sw) of t[nb][h][w][c]: plane (ph, pw) holds pixels (2a+ph, 2b+pw)."""
    es = t.element_size()
    out = []
    for ph in range(2):
        for pw in range(2):
            out.append((t.data_ptr() + (ph * w + pw) * c * es, (h - ph + 1) // 2, (w - pw + 1) // 2,
                        h * w * c, 2 * w * c, 2 * c))
    return out


# PESR_FUSED_BN: "0" (default) = separate statistics pass over the conv output, "1" = accumulated by the conv epilogue
# (pesr_conv_desc.bn_sums) for layers 1-7, "2" = for every layer.  Measured on the GAN step (tools/ab_env2.sh): 19.19 /
# 19.14 / 19.38 ms, i.e. no gain - the statistics cost the MMA-bound layers' epilogue what the extra pass cost - and the
# fp32 shared-memory atomics make the statistics (hence the logits, at the 1e-3 level after eight BatchNorm layers)
# depend on the warp scheduling order, so the deterministic separate pass stays the default.
_FUSE_BN_STATS = os.environ.get("PESR_FUSED_BN", "0")
_MERGE_S2_DGRAD = os.environ.get("PESR_NO_MERGED_S2_DGRAD") != "1"    # A/B knob: four parity classes in one launch
_FC1_WGRAD_SKINNY = os.environ.get("PESR_FC1_WGRAD_SKINNY") == "1"     # A/B knob: CUDA-core Linear weight gradient (round 1)


class DiscriminatorEngine:
    def __init__(self, disc, dtype=torch.float16):
        self.disc = disc
        self.dtype = dtype
        self.dt = ops.dt_code(dtype)
        self.pools = {}
        self.packed = None
        self.device = None
        self.grad_hook = None
        self.grad_hook_finish = None
        self.grad_hook_flush = None    # launches the all-reduce of the gradient ranges handed over so far (tail bucket)
        self.flat_alloc = None         # allocator of the flat gradient buffer (FlatGrads.alloc)
        self.param_list = None
        self.defer_finish = False  # True: the data-parallel wrapper waits for the all-reduce itself (DataParallel.finish)
        self.last_flat = None
        self.lane = None           # SideLane for the weight-gradient kernels (see engine_g.SideLane)
        self.trace_hook = None     # callable(plan), called at the end of every forward (parity tests read the saved activations)
        # data-parallel training (parallel.DataParallel): the Linear(73728 -> 1024) weight gradient is dz1^T x flat7, a
        # rank-(rows) product.  Instead of all-reducing its 302 MB the ranks all-gather the two factors (19 MB at 8 GPUs)
        # and every rank forms the averaged gradient itself.  fc1_gather(dz1, flat7) -> (dz1_all, flat7_all, wait(), world);
        # grad_hook_skip(lo, hi) tells the wrapper that this range of the flat buffer is not to be reduced.
        self.fc1_gather = None
        self.grad_hook_skip = None
        self.fc1_scratch = {}

    # ------------------------------------------------------------------ parameters
    def _ensure_packed(self, device):
        d = self.disc
        sentinel = (d.features[0][0].weight.data_ptr(), d.classifier[2].bias.data_ptr())
        if self.packed is not None and self.device == device and self.sentinel == sentinel:
            return
        self.sentinel, self.device, self.pools = sentinel, device, {}
        dt = self.dtype
        pk = {"c0_f": PackedWeight(d.features[0][0].weight, 4, dt, pad_to=64),
              "c0_d": PackedWeight(d.features[0][0].weight, 6, dt, pad_to=32)}
        for i in range(1, 8):
            pk[f"c{i}_f"] = PackedWeight(d.features[i][0].weight, 0, dt)
            pk[f"c{i}_d"] = PackedWeight(d.features[i][0].weight, 1, dt)
        self.packed = pk
        self.fwd_packs = [v for k, v in pk.items() if k.endswith("_f")]
        self.bwd_packs = [v for k, v in pk.items() if k.endswith("_d")]
        self.fwd_multi = ops.MultiPack(self.fwd_packs, device, dt)
        self.bwd_multi = ops.MultiPack(self.bwd_packs, device, dt)
        fc1, fc2 = d.classifier[0], d.classifier[2]
        self.w1_16 = torch.empty(fc1.weight.shape, device=device, dtype=dt)
        self.w2_16 = torch.empty(fc2.weight.shape, device=device, dtype=dt)
        self.fc_key = None
        # conv 0 sees the image minus IMG_SHIFT (applied to the padding too, so conv0' = conv0 - IMG_SHIFT*sum(w)
        # exactly): train-mode BatchNorm is invariant to that per-channel constant, and the 16-bit pre-BN tensor
        # no longer carries a mean several times its standard deviation.
        self.img_shift = torch.full((3,), -127.5, device=device, dtype=torch.float32)
        self.c0_shift = torch.zeros(64, device=device, dtype=torch.float32)
        self.c0_key = None
        self.scale_ws = torch.zeros(4, device=device, dtype=torch.float32)
        self.bn_ws = torch.zeros(2 * 512, device=device, dtype=torch.float64)
        self.flat_grads = FlatGrads(self.param_list, alloc=self.flat_alloc)
        self.offsets, self.flat_numel = self.flat_grads.offsets, self.flat_grads.numel

    def invalidate_packs(self):
        """See GeneratorEngine.invalidate_packs."""
        if self.packed is not None:
            self.fwd_multi.key = None
            self.bwd_multi.key = None
            self.fc_key = None
            self.c0_key = None

    def _pack_fc(self):
        fc1, fc2 = self.disc.classifier[0], self.disc.classifier[2]
        key = (fc1.weight.data_ptr(), fc1.weight._version, fc2.weight._version)
        if key != self.fc_key:
            ops.cast16(fc1.weight.detach(), self.w1_16)
            ops.cast16(fc2.weight.detach(), self.w2_16)
            self.fc_key = key

    # ------------------------------------------------------------------ plans
    def _geometry(self, h, w):
        dims = []
        for (_ci, _co, s) in D_LAYERS:
            if s == 2:
                h, w = (h + 1) // 2, (w + 1) // 2
            dims.append((h, w))
        return dims

    def _new_plan(self, nb, h, w, groups):
        """Buffers and forward launch descriptors for `groups` calls of nb images each, executed as ONE batch of
        groups * nb images (convolutions and the classifier do not mix samples; BatchNorm keeps per-group statistics)."""
        d = self.disc
        dev, tdt, dt, pk = self.device, self.dtype, self.dt, self.packed
        pl = _Plan()
        nt = groups * nb
        pl.nb, pl.nt, pl.groups, pl.h, pl.w, pl.busy = nb, nt, groups, h, w, False
        dims = self._geometry(h, w)
        pl.dims = dims
        e16 = lambda *s: torch.empty(*s, device=dev, dtype=tdt)  # noqa: E731
        e32 = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)  # noqa: E731
        pl.col0 = e16(nt * h * w, 64)
        pl.Y = [e16(nt * hh * ww, co) for (hh, ww), (_ci, co, _s) in zip(dims, D_LAYERS)]
        pl.A = [e16(nt * hh * ww, co) for (hh, ww), (_ci, co, _s) in zip(dims, D_LAYERS)]
        pl.mean = [e32(groups, co) for (_ci, co, _s) in D_LAYERS]
        pl.rstd = [e32(groups, co) for (_ci, co, _s) in D_LAYERS]
        # BatchNorm statistics sums [layer][group][2][512] (fp64), cleared by one fill per forward
        pl.bn_sums = torch.zeros(8, groups, 2, 512, device=dev, dtype=torch.float64)
        h7, w7 = dims[7]
        kfc = 512 * h7 * w7
        fc1 = d.classifier[0]
        if kfc != fc1.in_features:
            raise RuntimeError(f"pesr_b200.Discriminator: a {h}x{w} input gives {kfc} features but classifier.0 expects "
                               f"{fc1.in_features} (patch_size {d.patch_size}): size mismatch")
        pl.kfc = kfc
        pl.flat7 = e16(nt, kfc)
        pl.h1_32, pl.h1_16 = e32(nt, 1024), e16(nt, 1024)
        pl.fc_ws = e32(max(ops.linear_workspace_floats(16, kfc, 1024), ops.linear_workspace_floats(16, 1024, 1), 1))
        # Linear(kfc -> 1024) as a split-K run of the implicit-GEMM kernel: the rows are the "pixels" of a 1 x nt image
        pl.fc1_tc = kfc % 64 == 0
        if pl.fc1_tc:
            pl.fc1_ksplit = max(1, min(kfc // 64 // 8, 36))
            pl.fc1_part = e32(pl.fc1_ksplit * nt * 1024)
            pl.fc1_fwd = ops.make_conv_desc(dtype=dt, nb=1, h=1, w=nt, cin=kfc, cout=1024, block_n=256, taps=[(0, 0)],
                                            srcs=[ops.nhwc_src(pl.flat7, 1, 1, nt, kfc)], wpacked=self.w1_16,
                                            out32=pl.fc1_part, ld_out32=1024, ksplit=pl.fc1_ksplit,
                                            split_stride32=nt * 1024)
        taps2, srcs2, widx2 = _s2_taps()
        f = []
        f.append(ops.make_conv_desc(dtype=dt, nb=nt, h=h, w=w, cin=64, cout=64, taps=[(0, 0)],
                                    srcs=[ops.nhwc_src(pl.col0, nt, h, w, 64)], wpacked=pk["c0_f"].buf,
                                    out16=pl.Y[0], ld_out16=64))
        for i in range(1, 8):
            ci, co, s = D_LAYERS[i]
            hi, wi = dims[i - 1]
            ho, wo = dims[i]
            if s == 1:
                f.append(ops.make_conv_desc(dtype=dt, nb=nt, h=ho, w=wo, cin=ci, cout=co,
                                            srcs=[ops.nhwc_src(pl.A[i - 1], nt, hi, wi, ci)], wpacked=pk[f"c{i}_f"].buf,
                                            out16=pl.Y[i], ld_out16=co))
            else:
                f.append(ops.make_conv_desc(dtype=dt, nb=nt, h=ho, w=wo, cin=ci, cout=co, taps=taps2, tap_src=srcs2,
                                            tap_widx=widx2, srcs=_parity_planes(pl.A[i - 1], nt, hi, wi, ci),
                                            wpacked=pk[f"c{i}_f"].buf, out16=pl.Y[i], ld_out16=co))
        pl.fwd = f
        return pl

    def _bwd_scratch(self, nbb, h, w, gsel):
        """Backward buffers for nbb = gsel * nb images (shared by all plan instances of that size)."""
        key = ("bwd", nbb, h, w, gsel)
        sc = self.pools.get(key)
        if sc is None:
            dev, tdt = self.device, self.dtype
            dims = self._geometry(h, w)
            sc = _Plan()
            sc.dZ = [torch.empty(nbb * hh * ww, co, device=dev, dtype=tdt) for (hh, ww), (_ci, co, _s) in zip(dims, D_LAYERS)]
            sc.dY = [torch.empty(nbb * hh * ww, co, device=dev, dtype=tdt) for (hh, ww), (_ci, co, _s) in zip(dims, D_LAYERS)]
            h7, w7 = dims[7]
            sc.dflat32 = torch.empty(nbb, 512 * h7 * w7, device=dev, dtype=torch.float32)
            sc.dh1 = torch.empty(nbb, 1024, device=dev, dtype=torch.float32)
            sc.dz1_16 = torch.empty(nbb, 1024, device=dev, dtype=tdt)
            sc.scale1 = torch.zeros(4, device=dev, dtype=torch.float32)
            sc.scale_tot = torch.ones(1, device=dev, dtype=torch.float32)
            sc.bn_sums = torch.zeros(8, gsel, 2, 512, device=dev, dtype=torch.float64)
            kfc = 512 * h7 * w7
            sc.fc1_dgrad = None
            if kfc % 256 == 0:
                # dflat = dz1 x W1 with W1 [1024][kfc] read MN-major in place (no transposed copy of 75.5 M weights)
                sc.fc1_dgrad = ops.make_conv_desc(dtype=self.dt, nb=1, h=1, w=nbb, cin=1024, cout=kfc, block_n=256,
                                                  taps=[(0, 0)], srcs=[ops.nhwc_src(sc.dz1_16, 1, 1, nbb, 1024)],
                                                  wpacked=self.w1_16, out32=sc.dflat32, ld_out32=kfc, b_mn_major=1)
            sc.Zd = torch.empty(nbb * h * w, 32, device=dev, dtype=torch.float32)
            sc.wg = torch.empty(max(9 * 512 * 512 * 4, 148 * 128 * 64), device=dev, dtype=torch.float32)
            sc.descs = {}
            self.pools[key] = sc
        return sc

    def _bwd_descs(self, pl, sc, nb):
        """dgrad / wgrad descriptors binding the first nb images of plan instance `pl` to the backward scratch `sc`."""
        got = sc.descs.get(id(pl))
        if got is not None:
            return got
        dt, pk = self.dt, self.packed
        dims = pl.dims
        dg, wg = {}, {}
        for i in range(7, 0, -1):
            ci, co, s = D_LAYERS[i]
            hi, wi = dims[i - 1]
            ho, wo = dims[i]
            if s == 1:
                dg[i] = [ops.make_conv_desc(dtype=dt, nb=nb, h=hi, w=wi, cin=co, cout=ci,
                                            srcs=[ops.nhwc_src(sc.dY[i], nb, ho, wo, co)], wpacked=pk[f"c{i}_d"].buf,
                                            mask16=pl.A[i - 1], ld_mask16=ci, mask_mode=2, out16=sc.dZ[i - 1],
                                            ld_out16=ci)]
                wg[i] = ops.make_wgrad_desc(dtype=dt, nb=nb, h=ho, w=wo, a=sc.dY[i], a_c=co, m_total=co,
                                            b_srcs=[ops.nhwc_src(pl.A[i - 1], nb, hi, wi, ci)], n_total=ci,
                                            partials=sc.wg)
            else:
                lst, all_taps, all_widx, classes = [], [], [], []
                for ph in range(2):
                    for pw in range(2):
                        gh, gw = (hi - ph + 1) // 2, (wi - pw + 1) // 2
                        if gh == 0 or gw == 0:
                            continue
                        ys = [(1, 0)] if ph == 0 else [(0, 1), (2, 0)]     # (dy, offset into dY rows)
                        xs = [(1, 0)] if pw == 0 else [(0, 1), (2, 0)]
                        taps = [(oy, ox) for (dy, oy) in ys for (dx, ox) in xs]
                        widx = [8 - (dy * 3 + dx) for (dy, oy) in ys for (dx, ox) in xs]   # mode-1 rows are tap-flipped
                        all_taps += taps
                        all_widx += widx
                        classes.append((len(taps), ph, pw))
                        lst.append(ops.make_conv_desc(
                            dtype=dt, nb=nb, h=gh, w=gw, cin=co, cout=ci, taps=taps, tap_widx=widx,
                            srcs=[ops.nhwc_src(sc.dY[i], nb, ho, wo, co)], wpacked=pk[f"c{i}_d"].buf,
                            mask16=pl.A[i - 1], ld_mask16=ci, mask_mode=2, out16=sc.dZ[i - 1], ld_out16=ci,
                            out_h=hi, out_w=wi, out_sy=2, out_sx=2, out_oy=ph, out_ox=pw, aux_mode=1))
                if _MERGE_S2_DGRAD and hi % 2 == 0 and wi % 2 == 0 and len(classes) == 4:
                    # even maps: the four parity classes share one (hi/2 x wi/2) pixel grid -> ONE launch of 4 x the tiles
                    # (alone each class fills half the machine or less: 65 launches of ~22 us per GAN step in round 1)
                    lst = [ops.make_conv_desc(
                        dtype=dt, nb=nb, h=hi // 2, w=wi // 2, cin=co, cout=ci, taps=all_taps, tap_widx=all_widx,
                        srcs=[ops.nhwc_src(sc.dY[i], nb, ho, wo, co)], wpacked=pk[f"c{i}_d"].buf,
                        mask16=pl.A[i - 1], ld_mask16=ci, mask_mode=2, out16=sc.dZ[i - 1], ld_out16=ci,
                        out_h=hi, out_w=wi, out_sy=2, out_sx=2, aux_mode=1, classes=classes)]
                dg[i] = lst
                taps2, srcs2, _ = _s2_taps()
                wg[i] = ops.make_wgrad_desc(dtype=dt, nb=nb, h=ho, w=wo, a=sc.dY[i], a_c=co, m_total=co,
                                            b_srcs=_parity_planes(pl.A[i - 1], nb, hi, wi, ci), n_total=ci, taps=taps2,
                                            tap_src=srcs2, partials=sc.wg)
        h, w = pl.h, pl.w
        wg[0] = ops.make_wgrad_desc(dtype=dt, nb=nb, h=h, w=w, a=sc.dY[0], a_c=64, m_total=64,
                                    b_srcs=[ops.nhwc_src(pl.col0, nb, h, w, 64)], n_total=64, taps=[(0, 0)],
                                    partials=sc.wg)
        dg[0] = [ops.make_conv_desc(dtype=dt, nb=nb, h=h, w=w, cin=64, cout=32, taps=[(0, 0)],
                                    srcs=[ops.nhwc_src(sc.dY[0], nb, h, w, 64)], wpacked=pk["c0_d"].buf, out32=sc.Zd,
                                    ld_out32=32)]
        sc.descs[id(pl)] = (dg, wg)
        return dg, wg

    def _acquire(self, nb, h, w, groups):
        pool = self.pools.setdefault((nb, h, w, groups), [])
        for pl in pool:
            if not pl.busy:
                return pl
        if len(pool) >= 8:
            raise RuntimeError("pesr_b200.Discriminator: more than 8 live autograd graphs of one input shape")
        pl = self._new_plan(nb, h, w, groups)
        pool.append(pl)
        return pl

    # ------------------------------------------------------------------ forward
    def forward(self, xs, save):
        """xs: one tensor or a list of G tensors [nb,3,h,w] = G separate calls D(xs[0]), D(xs[1]), ... (each with its own
        train-mode BatchNorm statistics, running statistics updated call after call) executed as one batch.
        Returns ([nb,1] logits per call, state)."""
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
        self._ensure_packed(xs[0].device)
        pl = self._acquire(nb, h, w, G)
        nt = pl.nt
        self.fwd_multi.run()
        self._pack_fc()
        training = d.training
        w0 = d.features[0][0].weight
        if (w0.data_ptr(), w0._version) != self.c0_key:
            torch.mul(w0.detach().sum(dim=(1, 2, 3)), 127.5, out=self.c0_shift)
            self.c0_key = (w0.data_ptr(), w0._version)
        P = nb * h * w
        for g, x in enumerate(xs):
            ops.im2col3(x, pl.col0[g * P:(g + 1) * P], affine_b=self.img_shift, pad_affine=True)
        stream = torch.cuda.current_stream().cuda_stream
        if training:
            pl.bn_sums.zero_()
        for i in range(8):
            # train mode: optionally the conv epilogue accumulates the BatchNorm sums of its own (rounded) output
            fuse_stats = training and G == 1 and _FUSE_BN_STATS != "0" and (i > 0 or _FUSE_BN_STATS == "2")
            pl.fwd[i].bn_sums = pl.bn_sums[i].data_ptr() if fuse_stats else None
            _run_conv(pl.fwd[i], stream)
            bn = d.features[i][1]
            hh, ww = pl.dims[i]
            npix, co = nb * hh * ww, D_LAYERS[i][1]
            if training:
                sums = pl.bn_sums[i]
                if not fuse_stats:
                    ops.bn_reduce(pl.Y[i], npix, co, sums, groups=G, zero_first=False)
                ops.bn_lrelu_fwd(pl.Y[i], npix, co, pl.mean[i], pl.rstd[i], bn.weight.detach(), bn.bias.detach(), pl.A[i],
                                 groups=G, sums_ws=sums, eps=bn.eps, momentum=bn.momentum, running_mean=bn.running_mean,
                                 running_var=bn.running_var, num_batches=bn.num_batches_tracked,
                                 running_mean_shift=self.c0_shift if i == 0 else None)
            else:
                pl.mean[i].copy_((bn.running_mean - self.c0_shift if i == 0 else bn.running_mean).expand(G, co))
                pl.rstd[i].copy_(torch.rsqrt(bn.running_var + bn.eps).expand(G, co))
                ops.bn_lrelu_fwd(pl.Y[i], npix, co, pl.mean[i], pl.rstd[i], bn.weight.detach(), bn.bias.detach(), pl.A[i],
                                 groups=G)
        h7, w7 = pl.dims[7]
        ops.flatten_nchw16(pl.A[7], nt, h7 * w7, 512, pl.flat7)
        fc1, fc2 = d.classifier[0], d.classifier[2]
        if pl.fc1_tc:
            _run_conv(pl.fc1_fwd, stream)
            ops.linear_finalize(pl.fc1_part, pl.fc1_ksplit, nt, 1024, fc1.bias.detach(), self.dtype, out32=pl.h1_32,
                                out16=pl.h1_16, act=ACT_LRELU)
        else:
            ops.linear_fwd(pl.flat7, self.w1_16, fc1.bias.detach(), nt, pl.kfc, 1024, pl.fc_ws, out32=pl.h1_32,
                           out16=pl.h1_16, act=ACT_LRELU)
        logits = torch.empty(nt, 1, device=xs[0].device, dtype=torch.float32)
        ops.linear_fwd(pl.h1_16, self.w2_16, fc2.bias.detach(), nt, 1024, 1, pl.fc_ws, out32=logits)
        outs = logits if single else list(logits.split(nb))
        if self.trace_hook is not None:
            self.trace_hook(pl)
        if save:
            if not training:
                raise NotImplementedError("pesr_b200.Discriminator: backward in eval() mode is not on the PESR path")
            pl.busy = True
            return outs, (pl, _Release(pl), xs)
        return outs, None

    # ------------------------------------------------------------------ Linear(kfc -> 1024) weight gradient
    def _fc1_wgrad(self, dz1, flat7, kfc, grad, world=1, stream=None):
        """grad[1024][kfc] = dz1^T x flat7 / world (dz1 fp32 [rows][1024], flat7 16-bit [rows][kfc]) on the tensor cores:
        the split-K weight-gradient kernel with the rows as its "pixels" (64-row patches), dz1 as a 16-bit hi + lo pair
        stacked along K (22 significant bits; flat7 is already 16-bit), written straight into `grad` with the range scale
        and 1/world folded into the store.  302 MB are written once; nothing is read back."""
        rows = dz1.shape[0]
        if _FC1_WGRAD_SKINNY or kfc % 64 != 0 or grad.data_ptr() % 32 != 0 or self.dtype != torch.float16:
            ops.linear_wgrad(dz1, flat7, rows, kfc, 1024, grad, mul=1.0 / world)
            return
        rp = (rows + 63) // 64 * 64
        key = (rows, kfc)
        sc = self.fc1_scratch.get(key)
        if sc is None:
            dev = dz1.device
            sc = _Plan()
            sc.a16 = torch.zeros(2 * rp, 1024, device=dev, dtype=self.dtype)       # [hi rows; lo rows], padding rows stay 0
            sc.b16 = torch.zeros(2 * rp, kfc, device=dev, dtype=self.dtype)        # [flat7; flat7]
            sc.scale = torch.zeros(4, device=dev, dtype=torch.float32)
            self.fc1_scratch[key] = sc
        ops.amax_scale(dz1, sc.scale, target=64.0)
        ops.split16(dz1, sc.a16[:rows], sc.a16[rp:rp + rows], mul_dev=sc.scale[1:2])
        sc.b16[:rows].copy_(flat7)
        sc.b16[rp:rp + rows].copy_(flat7)
        nimg = 2 * rp // 64
        d = ops.make_wgrad_desc(dtype=self.dt, nb=nimg, h=4, w=16, a=sc.a16, a_c=1024, m_total=1024,
                                b_srcs=[ops.nhwc_src(sc.b16, nimg, 4, 16, kfc)], n_total=kfc, taps=[(0, 0)], partials=grad,
                                splits=1, out_mul=1.0 / world, out_div_dev=sc.scale[1:2])
        splits = C.c_int32(0)
        check(lib.pesr_conv_wgrad(C.byref(d), C.byref(splits), stream if stream is not None else
                                  torch.cuda.current_stream().cuda_stream), "pesr_conv_wgrad")
        if splits.value != 1:
            raise RuntimeError("fc1 weight gradient: the kernel did not take the single-split path")

    # ------------------------------------------------------------------ backward
    def backward(self, state, dlogits, need_param_grads, need_input_grad):
        """dlogits: one tensor or a list with one [nb,1] gradient (or None) per call of the batched forward;
        need_input_grad: bool or list of bool per call.  Parameter gradients are the SUM over the calls (the D phase
        back-propagates D(hr) and D(sr) before one optimiser step, train.py:213-216).  Only the leading calls that need
        anything are back-propagated (the G phase needs d/d(sr) of D(sr) only, train.py:237-258).
        Returns (grads dict, list of input gradients per call or None)."""
        pl, _rel, xs = state
        d = self.disc
        G, nb, h, w = pl.groups, pl.nb, pl.h, pl.w
        single = torch.is_tensor(dlogits)
        dl = [dlogits] if single else list(dlogits)
        need_in = [need_input_grad] * G if isinstance(need_input_grad, bool) else list(need_input_grad)
        if need_param_grads:
            gsel = G
        else:
            gsel = max((g + 1 for g in range(G) if need_in[g]), default=0)
        if gsel == 0:
            return {}, [None] * G
        nbb = gsel * nb
        dev = xs[0].device
        parts = [(dl[g].contiguous().float() if dl[g] is not None else torch.zeros(nb, 1, device=dev)) for g in range(gsel)]
        dlogits = parts[0] if gsel == 1 else torch.cat(parts, dim=0)
        sc = self._bwd_scratch(nbb, h, w, gsel)
        self.bwd_multi.run()
        dg, wg = self._bwd_descs(pl, sc, nbb)
        fc1, fc2 = d.classifier[0], d.classifier[2]
        stream = torch.cuda.current_stream().cuda_stream
        grads, flat, hook = {}, None, None
        off = self.offsets
        mark_hi = [self.flat_numel]
        if need_param_grads:
            flat = self.flat_grads.get(dev)
            grads = {p: flat[off[p]:off[p] + p.numel()].view(p.shape) for p in self.param_list}
            hook = self.grad_hook

        def mark(param):
            lo = off[param]
            if hook is not None and lo < mark_hi[0]:
                hook(lo, mark_hi[0], flat)
            mark_hi[0] = lo

        h1_32, h1_16, flat7 = pl.h1_32[:nbb], pl.h1_16[:nbb], pl.flat7[:nbb]
        # parameter gradients are leaves of the backward-data chain: they run on a second stream (engine_g.SideLane)
        lane = None
        if need_param_grads and _WGRAD_STREAM:
            if self.lane is None:
                self.lane = SideLane(dev)
            lane = self.lane
            lane.begin(torch.cuda.current_stream())

        def on_lane(fn):
            if lane is None:
                fn(stream)
            else:
                lane.enter_lane()
                with torch.cuda.stream(lane.stream):
                    fn(lane.stream.cuda_stream)

        def chained():
            if lane is not None:
                lane.after_chain()

        # ---- classifier (fp32 gradients, 16-bit operands)
        if need_param_grads:
            def fc2_grads(_st):
                ops.linear_wgrad(dlogits, h1_16, nbb, 1024, 1, grads[fc2.weight])
                torch.sum(dlogits, dim=0, out=grads[fc2.bias])
                mark(fc2.weight)
            on_lane(fc2_grads)
        ops.linear_dgrad(dlogits, self.w2_16, nbb, 1024, 1, sc.dh1)
        dz1 = sc.dh1 * torch.where(h1_32 > 0, 1.0, 0.2)      # LeakyReLU'(h1), an nbb x 1024 tensor
        chained()
        fc1_deferred = None
        if need_param_grads:
            if hook is not None and self.fc1_gather is not None and self.grad_hook_skip is not None:
                # data parallel: gather the factors now (side stream), form the averaged gradient at the end of backward
                torch.sum(dz1, dim=0, out=grads[fc1.bias])
                mark(fc1.bias)
                fc1_deferred = self.fc1_gather(dz1, flat7)
                self.grad_hook_skip(off[fc1.weight], mark_hi[0])
                mark_hi[0] = off[fc1.weight]
            else:
                def fc1_grads(st):
                    self._fc1_wgrad(dz1, flat7, pl.kfc, grads[fc1.weight], stream=st)
                    torch.sum(dz1, dim=0, out=grads[fc1.bias])
                    mark(fc1.weight)
                on_lane(fc1_grads)
        if sc.fc1_dgrad is not None:
            ops.amax_scale(dz1, sc.scale1, target=16.0)
            sc.dz1_16.copy_(dz1 * sc.scale1[1:2])
            _run_conv(sc.fc1_dgrad, stream)                    # dflat32 carries scale1
        else:
            ops.linear_dgrad(dz1, self.w1_16, nbb, pl.kfc, 1024, sc.dflat32)
            sc.scale1[1:2].fill_(1.0)
        ops.amax_scale(sc.dflat32, self.scale_ws, target=16.0)
        torch.mul(self.scale_ws[1:2], sc.scale1[1:2], out=sc.scale_tot)
        scale = sc.scale_tot                                   # every 16-bit gradient below carries scale1*scale2
        h7, w7 = pl.dims[7]
        ops.unflatten_nchw16(sc.dflat32, pl.A[7], nbb, h7 * w7, 512, sc.dZ[7], mul_dev=self.scale_ws[1:2])
        sc.bn_sums.zero_()
        splits = C.c_int32(0)
        dummy = None
        for i in range(7, -1, -1):
            ci, co, s = D_LAYERS[i]
            hh, ww = pl.dims[i]
            npix = nb * hh * ww
            bn, conv = d.features[i][1], d.features[i][0]
            if need_param_grads:
                dgam, dbet = grads[bn.weight], grads[bn.bias]
            else:
                if dummy is None:
                    dummy = torch.empty(2 * 512, device=dev, dtype=torch.float32)
                dgam, dbet = dummy[:co], dummy[512:512 + co]
            ops.bn_lrelu_bwd(sc.dZ[i], pl.Y[i], npix, co, pl.mean[i], pl.rstd[i], bn.weight.detach(), sc.bn_sums[i],
                             sc.dY[i], dgam, dbet, grad_div_dev=scale, groups=gsel, zero_first=False)
            chained()
            if need_param_grads:
                def conv_grads(st, i=i, ci=ci, co=co, conv=conv):
                    check(lib.pesr_conv_wgrad(C.byref(wg[i]), C.byref(splits), st), "pesr_conv_wgrad")
                    if i == 0:
                        check(lib.pesr_wgrad_reduce(sc.wg.data_ptr(), splits.value, 1, 64, 64, ops.WMAP_COL_IN, 64, 3, 1.0,
                                                    scale.data_ptr(), 0, grads[conv.weight].data_ptr(), st),
                              "pesr_wgrad_reduce")
                    else:
                        check(lib.pesr_wgrad_reduce(sc.wg.data_ptr(), splits.value, 9, co, ci, ops.WMAP_OIHW, co, ci, 1.0,
                                                    scale.data_ptr(), 0, grads[conv.weight].data_ptr(), st),
                              "pesr_wgrad_reduce")
                    mark(conv.weight)
                on_lane(conv_grads)
            if i > 0 or any(need_in[:gsel]):
                for dsc in dg[i]:
                    _run_conv(dsc, stream)
        if lane is not None:
            lane.join()
        if need_param_grads and hook is not None and self.grad_hook_flush is not None:
            self.grad_hook_flush()          # the tail bucket goes on the wire now; DataParallel.finish() waits for it
        if fc1_deferred is not None:
            dz_all, f_all, wait, world = fc1_deferred
            wait()
            self._fc1_wgrad(dz_all, f_all, pl.kfc, grads[fc1.weight], world=world)
        dxs = [None] * G
        if any(need_in[:gsel]):
            dx = torch.empty(nbb, 3, h, w, device=dev, dtype=torch.float32)
            ops.col2im3(sc.Zd, 32, nbb, h, w, dx, mul=1.0, div_dev=scale, sgn=-1)
            for g in range(gsel):
                if need_in[g]:
                    dxs[g] = dx[g * nb:(g + 1) * nb]
        if need_param_grads:
            if mark_hi[0] != 0:
                raise AssertionError("discriminator backward: gradient ranges did not cover the flat buffer")
            if hook is not None and self.grad_hook_finish is not None and not self.defer_finish:
                self.grad_hook_finish()
        self.last_flat = flat
        return grads, dxs
