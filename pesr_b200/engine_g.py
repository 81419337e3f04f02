"""Kernel schedule of the x4 Generator (reference: model/pesr.py:3-38, model/basic.py:33-60).

Forward (all activations NHWC 16-bit, fp32 residual stream kept beside them):

    lr --im2col3(+sub_mean)--> col --embed GEMM--> F0 (fp32) , X[0]
    block i:  T[i] = relu(conv1(X[i]) + b1)                      (bias+ReLU fused)
              S    = S + res_scale*(conv2(T[i]) + b2) ; X[i+1] = cvt(S)   (bias, scale, fp32 skip fused)
    U0 = tail(X[depth]) + b + F0                                 (global skip fused)
    U1 = shuffle(conv(U0)) ; U2 = shuffle(conv(U1))              (PixelShuffle fused into the store)
    Z  = U2 x W4 (1x1 GEMM, 27 columns) --col2im3(+bias, +add_mean)--> sr (NCHW fp32)

Backward mirrors it with the same implicit-GEMM kernel on flipped weights (dgrad) and the split-K
MN-major kernel (wgrad).  Parameter gradients are fp32, in the reference's layouts, and are views of
ONE flat buffer in parameter order; they complete from the end of that buffer towards its start, and
``grad_hook(lo, hi, flat)`` is called as each range completes so that a data-parallel wrapper can start
the NCCL all-reduce of a bucket while the rest of backward is still running.

Launch descriptors are built once per (batch, height, width) plan and replayed; per step the host only
re-packs changed weights and issues the launches.
"""
import collections
import ctypes as C

import os

import torch

from . import ops
from ._lib import check, lib
from .ops import ACT_RELU, OUT_SHUFFLE2, OUT_UNSHUFFLE2


class PackedWeight:
    """16-bit GEMM operand of one fp32 parameter, re-packed only when the parameter changed."""

    def __init__(self, param, mode, dtype, pad_to=0):
        co, ci, k = param.shape[0], param.shape[1], param.shape[2]
        self.param, self.mode, self.pad_to = param, mode, pad_to
        self.buf = torch.empty(ops.packed_shape(co, ci, k, mode, pad_to), device=param.device, dtype=dtype)
        self.key = None

    def get(self):
        p = self.param
        key = (p.data_ptr(), p._version)
        if key != self.key:
            ops.pack_weights(p.detach(), self.mode, self.buf, self.pad_to)
            self.key = key
        return self.buf


def _unshuffle_perm(vec, c):
    """packed PixelShuffle order (ij, c) -> reference order c*4 + ij."""
    return vec.view(4, c).t().reshape(-1)


def _shuffle_perm(vec, c):
    """reference order c*4 + ij -> packed (ij, c)."""
    return vec.view(c, 4).t().reshape(-1)


class _Plan:
    pass


class _Release:
    """Returns a plan instance to its pool when the autograd node that holds it dies."""

    def __init__(self, plan):
        self.plan = plan

    def __del__(self):
        self.plan.busy = False


# Plan cache bounds.  A plan owns every activation buffer of one (batch, height, width) shape -- 16.5 KB per LR pixel
# for an inference plan at 256 channels, i.e. 2.8 GB for one 510x339 image -- so the cache keeps the most recently used
# shapes only (test.py walks datasets of arbitrary image sizes in constant memory in the reference).
PLAN_CACHE_SHAPES = int(os.environ.get("PESR_PLAN_CACHE", "4"))
MAX_LIVE_GRAPHS = 4           # live autograd graphs of one input shape (each owns a full set of saved activations)
# eval-mode batches are processed in chunks of at most this many LR pixels (one plan of the chunk's shape is reused)
INFER_CHUNK_PIXELS = int(os.environ.get("PESR_INFER_CHUNK_PIXELS", str(1 << 20)))


class PlanCache:
    """LRU over shapes; each shape owns a small pool of plan instances (one per live autograd graph)."""

    def __init__(self, max_shapes=PLAN_CACHE_SHAPES, max_live=MAX_LIVE_GRAPHS, who="pesr_b200"):
        self.map = collections.OrderedDict()
        self.max_shapes, self.max_live, self.who = max_shapes, max_live, who
        self.stamp = 0

    def clear(self):
        self.map.clear()

    def __len__(self):
        return len(self.map)

    def __contains__(self, key):
        return key in self.map

    def __getitem__(self, key):
        return self.map[key]

    def acquire(self, key, factory):
        pool = self.map.get(key)
        if pool is None:
            pool = self.map[key] = []
        self.map.move_to_end(key)
        pl = next((q for q in pool if not q.busy), None)
        if pl is None:
            if len(pool) >= self.max_live:
                raise RuntimeError(f"{self.who}: more than {self.max_live} live autograd graphs of input shape {key}; "
                                   "run backward (or drop the outputs) of earlier forwards first")
            self._evict(keep=key)
            pl = factory()
            pl.busy = False
            pool.append(pl)
        self.stamp += 1
        pl.stamp = self.stamp
        return pl

    def _evict(self, keep):
        # called before a new plan is allocated: drop least-recently-used shapes with no live graph
        while len(self.map) > self.max_shapes:
            victim = next((k for k, pool in self.map.items() if k != keep and not any(q.busy for q in pool)), None)
            if victim is None:
                return
            del self.map[victim]


class FlatGrads:
    """The flat fp32 gradient buffer of one network (parameter order, 16-byte aligned tensors).

    One PERSISTENT buffer is reused from step to step, so gradient addresses are stable: the multi-tensor Adam keeps its
    pointer table and a CUDA graph of the step can be replayed.  Reuse is only safe once the previous hand-out has been
    consumed; the parameters' version counter tells (every optimiser step bumps it).  A second backward before the next
    optimiser step (gradient accumulation, two separate D calls) gets a fresh buffer, as autograd may still hold views
    of the first one in its input buffers."""

    def __init__(self, param_list, alloc=None):
        self.params = param_list
        self.alloc = alloc        # optional allocator of the persistent buffer: alloc(numel, device) (parallel.DataParallel
        #                           places it in symmetric memory for the NVLink peer-memory all-reduce)
        self.offsets, off = {}, 0
        for p in param_list:
            self.offsets[p] = off
            off += (p.numel() + 3) // 4 * 4
        self.numel = off
        self.buf = None
        self.handed_at = None

    def get(self, device):
        version = self.params[0]._version
        if self.buf is None or self.buf.device != device:
            self.buf = self.alloc(self.numel, device) if self.alloc is not None else \
                torch.empty(self.numel, device=device, dtype=torch.float32)
            self.handed_at = None
        if self.handed_at == version:
            return torch.empty(self.numel, device=device, dtype=torch.float32)
        self.handed_at = version
        return self.buf


# Weight gradients on a second stream (SideLane): OFF by default.  Measured on the B200 (tools/gpu_r2_h.sh, CUDA-graph
# replay, 50 steps): GAN step 16.90 ms with the lane vs 16.94 ms without, L1 pretrain step 11.17 vs 11.17 ms -- the SM
# clock under the 1 kW power cap drops by as much as the overlap gains (1657 vs 1732 MHz median): the step is power-
# bound, not tail-bound.  PESR_WGRAD_STREAM=1 enables it.
_WGRAD_STREAM = os.environ.get("PESR_WGRAD_STREAM") == "1"


class SideLane:
    """A second CUDA stream for the weight-gradient kernels of a backward pass.

    Backward-data is a serial chain (each dgrad feeds the next); the weight / bias gradients hang off it as leaves that
    nothing in the chain waits for.  On one stream every one of those ~130 launches sits BETWEEN two links of the chain,
    so every kernel boundary exposes the tail of one persistent kernel and the pipeline fill of the next (about 5 us of
    a 33 us trunk convolution).  With the leaves on their own stream the block scheduler fills the SMs one kernel's
    last wave leaves idle with the other stream's CTAs.  Ordering is by events (in a captured CUDA graph they become
    plain edges): the lane waits for the newest chain event before each leaf group, and the chain waits, before
    overwriting a ping-pong buffer, for the leaf work that read it (issued two chain launches earlier)."""

    def __init__(self, device):
        self.stream = torch.cuda.Stream(device=device)
        self.pool, self.used = [], 0

    def _event(self):
        if self.used == len(self.pool):
            self.pool.append(torch.cuda.Event())
        ev = self.pool[self.used]
        self.used += 1
        return ev

    def begin(self, main):
        self.main = main
        self.used = 0
        self.stream.wait_stream(main)
        self.chain_ev, self.chain_dirty = None, False
        self.lane_ev, self.snaps = None, []

    def before_chain(self, lag=2):
        """Call before a chain (backward-data) launch that may overwrite a buffer the lane read `lag` launches ago."""
        if len(self.snaps) >= lag and self.snaps[-lag] is not None:
            self.main.wait_event(self.snaps[-lag])
        self.snaps.append(self.lane_ev)

    def after_chain(self):
        ev = self._event()
        ev.record(self.main)
        self.chain_ev, self.chain_dirty = ev, True

    def enter_lane(self):
        """Make the lane wait for everything the chain has produced so far (no-op if it already did)."""
        if self.chain_dirty:
            self.stream.wait_event(self.chain_ev)
            self.chain_dirty = False

    def after_lane(self):
        ev = self._event()
        ev.record(self.stream)
        self.lane_ev = ev

    def join(self):
        self.main.wait_stream(self.stream)


def _run_conv(desc, stream):
    check(lib.pesr_conv_igemm(C.byref(desc), stream), "pesr_conv_igemm")


_FUSE_BIAS = os.environ.get("PESR_NO_FUSED_BIAS") != "1"    # A/B knob (tools/ab_env.sh)
_DUAL_PACK = os.environ.get("PESR_NO_DUAL_PACK") != "1"     # A/B knob: forward + backward weight layouts in one pass


class GeneratorEngine:
    def __init__(self, gen, dtype=torch.float16):
        self.gen = gen
        self.dtype = dtype
        self.dt = ops.dt_code(dtype)
        self.plans = PlanCache(who="pesr_b200.Generator")
        self.packed = None
        self.device = None
        self.grad_hook = None      # callable(lo, hi, flat) -> None
        self.grad_hook_finish = None
        self.grad_hook_flush = None    # launches the all-reduce of the gradient ranges handed over so far (tail bucket)
        self.flat_alloc = None         # allocator of the flat gradient buffer (FlatGrads.alloc)
        self.param_list = None
        self.last_flat = None
        self.defer_finish = False  # True: the data-parallel wrapper waits for the all-reduce itself (DataParallel.finish)
        self.lane = None           # SideLane for the weight-gradient kernels (created on first backward)
        self.trace_hook = None     # callable(plan), called at the end of every forward (parity tests read the saved activations)

    # ------------------------------------------------------------------ parameters
    def _convs(self):
        g = self.gen
        trunk = [(blk.body[0], blk.body[2]) for blk in list(g.body)[:-1]]
        return trunk, g.body[-1]

    def _ensure_packed(self, device):
        sentinel = (self.gen.embed.weight.data_ptr(), self.gen.add_mean.bias.data_ptr())
        if self.packed is not None and self.device == device and self.sentinel == sentinel:
            return
        self.sentinel = sentinel
        g, dt = self.gen, self.dtype
        self.device = device
        self.plans.clear()
        trunk, tail = self._convs()
        pk = {}
        pk["embed_f"] = PackedWeight(g.embed.weight, 4, dt, pad_to=64)
        pk["embed_d"] = PackedWeight(g.embed.weight, 6, dt, pad_to=32)
        for i, (c1, c2) in enumerate(trunk):
            pk[f"b{i}c1_f"] = PackedWeight(c1.weight, 0, dt)
            pk[f"b{i}c2_f"] = PackedWeight(c2.weight, 0, dt)
            pk[f"b{i}c1_d"] = PackedWeight(c1.weight, 1, dt)
            pk[f"b{i}c2_d"] = PackedWeight(c2.weight, 1, dt)
        pk["tail_f"] = PackedWeight(tail.weight, 0, dt)
        pk["tail_d"] = PackedWeight(tail.weight, 1, dt)
        pk["up0_f"] = PackedWeight(g.upsample[0].weight, 2, dt)
        pk["up0_d"] = PackedWeight(g.upsample[0].weight, 3, dt)
        pk["up2_f"] = PackedWeight(g.upsample[2].weight, 2, dt)
        pk["up2_d"] = PackedWeight(g.upsample[2].weight, 3, dt)
        pk["up4_f"] = PackedWeight(g.upsample[4].weight, 5, dt, pad_to=32)
        pk["up4_d"] = PackedWeight(g.upsample[4].weight, 7, dt, pad_to=64)
        self.packed = pk
        # one launch packs the forward operands and, from the same read of each fp32 weight, the backward-data
        # operands of the 3x3 layers; the remaining backward packs (embed / last conv, GEMM layouts) run before backward
        self.fwd_packs = [v for k, v in pk.items() if k.endswith("_f")]
        comp = [pk[k[:-2] + "_d"] if (v.mode in (0, 2) and _DUAL_PACK) else None for k, v in pk.items() if k.endswith("_f")]
        paired = {id(c) for c in comp if c is not None}
        self.bwd_packs = [v for k, v in pk.items() if k.endswith("_d") and id(v) not in paired]
        self.fwd_multi = ops.MultiPack(self.fwd_packs, device, dt, companions=comp)
        self.bwd_multi = ops.MultiPack(self.bwd_packs, device, dt)
        self.scale_ws = torch.zeros(4, device=device, dtype=torch.float32)
        C_ = g.n_feats
        self.bias_up0 = torch.empty(4 * C_, device=device, dtype=torch.float32)
        self.bias_up2 = torch.empty(4 * C_, device=device, dtype=torch.float32)
        # flat gradient layout: parameter order, each tensor starting on a 16-byte boundary
        self.flat_grads = FlatGrads(self.param_list, alloc=self.flat_alloc)
        self.offsets, self.flat_numel = self.flat_grads.offsets, self.flat_grads.numel

    # ------------------------------------------------------------------ plans
    def _plan(self, nb, h, w, train):
        return self.plans.acquire((nb, h, w, train), lambda: self._new_plan(nb, h, w, train))

    def invalidate_packs(self):
        """Forget which parameter versions the packed 16-bit operands were made from (a CUDA-graph replay updates
        the parameters without touching their Python-side version counters)."""
        if self.packed is not None:
            self.fwd_multi.key = None
            self.bwd_multi.key = None

    def _new_plan(self, nb, h, w, train):
        g = self.gen
        Cn, depth, rs = g.n_feats, g.n_resblock, float(g.res_scale)
        dev, tdt, dt = self.device, self.dtype, self.dt
        pk = self.packed
        P = nb * h * w
        pl = _Plan()
        pl.nb, pl.h, pl.w, pl.P, pl.train = nb, h, w, P, train
        e16 = lambda *s: torch.empty(*s, device=dev, dtype=tdt)  # noqa: E731
        e32 = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)  # noqa: E731
        pl.col_in = e16(P, 64)
        pl.F0 = e32(P, Cn)
        pl.S = e32(P, Cn)
        pl.X = [e16(P, Cn) for _ in range(depth + 1 if train else 2)]
        pl.T = [e16(P, Cn) for _ in range(depth if train else 1)]
        pl.U0 = e16(P, Cn)
        pl.U1 = e16(4 * P, Cn)
        pl.U2 = e16(16 * P, Cn)
        pl.Z = e32(16 * P, 32)
        pl.ypre = e32(nb, 3, 4 * h, 4 * w) if train else None
        X = (lambda i: pl.X[i]) if train else (lambda i: pl.X[i % 2])
        T = (lambda i: pl.T[i]) if train else (lambda i: pl.T[0])
        trunk, tail = self._convs()
        up0, up2, up4 = g.upsample[0], g.upsample[2], g.upsample[4]
        bn = ops.default_block_n(Cn)

        def src(t, hh, ww, c):
            return [ops.nhwc_src(t, nb, hh, ww, c)]

        # ---------------- forward launch list
        f = []
        f.append(ops.make_conv_desc(dtype=dt, nb=nb, h=h, w=w, cin=64, cout=Cn, taps=[(0, 0)],
                                    srcs=src(pl.col_in, h, w, 64), wpacked=pk["embed_f"].buf,
                                    bias=g.embed.bias, out32=pl.F0, ld_out32=Cn, out16=X(0), ld_out16=Cn))
        for i, (c1, c2) in enumerate(trunk):
            f.append(ops.make_conv_desc(dtype=dt, nb=nb, h=h, w=w, cin=Cn, cout=Cn, srcs=src(X(i), h, w, Cn),
                                        wpacked=pk[f"b{i}c1_f"].buf, bias=c1.bias, act=ACT_RELU, out16=T(i),
                                        ld_out16=Cn))
            f.append(ops.make_conv_desc(dtype=dt, nb=nb, h=h, w=w, cin=Cn, cout=Cn, srcs=src(T(i), h, w, Cn),
                                        wpacked=pk[f"b{i}c2_f"].buf, bias=c2.bias, alpha=rs,
                                        res32=pl.F0 if i == 0 else pl.S, ld_res32=Cn, out32=pl.S, ld_out32=Cn,
                                        out16=X(i + 1), ld_out16=Cn))
        f.append(ops.make_conv_desc(dtype=dt, nb=nb, h=h, w=w, cin=Cn, cout=Cn, srcs=src(X(depth), h, w, Cn),
                                    wpacked=pk["tail_f"].buf, bias=tail.bias, res32=pl.F0, ld_res32=Cn,
                                    out16=pl.U0, ld_out16=Cn))
        f.append(ops.make_conv_desc(dtype=dt, nb=nb, h=h, w=w, cin=Cn, cout=4 * Cn, block_n=bn,
                                    srcs=src(pl.U0, h, w, Cn), wpacked=pk["up0_f"].buf, bias=self.bias_up0,
                                    out16=pl.U1, ld_out16=Cn, out_mode=OUT_SHUFFLE2, ps_c=Cn))
        f.append(ops.make_conv_desc(dtype=dt, nb=nb, h=2 * h, w=2 * w, cin=Cn, cout=4 * Cn, block_n=bn,
                                    srcs=src(pl.U1, 2 * h, 2 * w, Cn), wpacked=pk["up2_f"].buf, bias=self.bias_up2,
                                    out16=pl.U2, ld_out16=Cn, out_mode=OUT_SHUFFLE2, ps_c=Cn))
        f.append(ops.make_conv_desc(dtype=dt, nb=nb, h=4 * h, w=4 * w, cin=Cn, cout=32, taps=[(0, 0)],
                                    srcs=src(pl.U2, 4 * h, 4 * w, Cn), wpacked=pk["up4_f"].buf, out32=pl.Z,
                                    ld_out32=32))
        pl.fwd = f
        if not train:
            return pl

        # ---------------- backward buffers and launch list
        pl.dcol = e16(16 * P, 64)
        pl.dZ2 = e16(4 * P, 4 * Cn)
        pl.dZ1 = e16(P, 4 * Cn)
        pl.dR16 = e16(P, Cn)
        pl.gS32 = e32(P, Cn)
        # ping-pong (block parity): the weight-gradient stream may still read a buffer while backward-data runs ahead
        pl.gS16 = [e16(P, Cn), e16(P, Cn)]
        pl.dT16 = [e16(P, Cn), e16(P, Cn)]
        pl.dF0 = e16(P, Cn)
        pl.Zd = e32(P, 32)
        pl.dx_sm = e32(nb, 3, h, w)
        pl.sums = torch.zeros(24, device=dev, dtype=torch.float32)
        pl.bias_tmp = e32(8 * Cn)
        pl.wg = e32(max(9 * 4 * Cn * Cn * 2, 9 * Cn * Cn * 8, 148 * max(Cn, 128) * 64))
        scale = self.scale_ws[1:2]
        b = []   # entries: ("conv", desc) | ("wgrad", desc, reduce-args) | ("bias", args) | ("mark", param) | ("call", fn)

        def wgrad(a, a_c, m_total, bt, b_c, n_total, gh, gw, param, map_mode, co, ci, taps=ops.TAPS_3X3, mul=1.0):
            d = ops.make_wgrad_desc(dtype=dt, nb=nb, h=gh, w=gw, a=a, a_c=a_c, m_total=m_total,
                                    b_srcs=[ops.nhwc_src(bt, nb, gh, gw, b_c)], n_total=n_total, taps=taps,
                                    partials=pl.wg)
            b.append(("wgrad", d, (len(taps), m_total, n_total, map_mode, co, ci, mul), param))

        def bgrad(x16, npix, c, ldc, param, mul=1.0, perm_c=0):
            b.append(("bias", (x16, npix, c, ldc, mul, perm_c), param))

        def conv(**kw):
            b.append(("conv", ops.make_conv_desc(dtype=dt, nb=nb, **kw)))

        # upsample.4 (Cout = 3): dgrad = 1x1 GEMM over the flipped im2col of dy, stored un-shuffled
        conv(h=4 * h, w=4 * w, cin=64, cout=Cn, taps=[(0, 0)], srcs=src(pl.dcol, 4 * h, 4 * w, 64),
             wpacked=pk["up4_d"].buf, out16=pl.dZ2, ld_out16=4 * Cn, out_mode=OUT_UNSHUFFLE2)
        wgrad(pl.U2, Cn, Cn, pl.dcol, 64, 64, 4 * h, 4 * w, up4.weight, ops.WMAP_COL_OUT, 3, Cn, taps=[(0, 0)])
        b.append(("mark", up4.weight))
        # upsample.2
        conv(h=2 * h, w=2 * w, cin=4 * Cn, cout=Cn, srcs=src(pl.dZ2, 2 * h, 2 * w, 4 * Cn), wpacked=pk["up2_d"].buf,
             out16=pl.dZ1, ld_out16=4 * Cn, out_mode=OUT_UNSHUFFLE2)
        bgrad(pl.dZ2, 4 * P, 4 * Cn, 4 * Cn, up2.bias, perm_c=Cn)
        wgrad(pl.dZ2, 4 * Cn, 4 * Cn, pl.U1, Cn, Cn, 2 * h, 2 * w, up2.weight, ops.WMAP_OIHW_PS, 4 * Cn, Cn)
        b.append(("mark", up2.weight))
        # upsample.0: its dgrad is dR, the gradient of (tail(X_depth) + F0)
        conv(h=h, w=w, cin=4 * Cn, cout=Cn, srcs=src(pl.dZ1, h, w, 4 * Cn), wpacked=pk["up0_d"].buf, out16=pl.dR16,
             ld_out16=Cn)
        bgrad(pl.dZ1, P, 4 * Cn, 4 * Cn, up0.bias, perm_c=Cn)
        wgrad(pl.dZ1, 4 * Cn, 4 * Cn, pl.U0, Cn, Cn, h, w, up0.weight, ops.WMAP_OIHW_PS, 4 * Cn, Cn)
        b.append(("mark", up0.weight))
        # tail conv
        bgrad(pl.dR16, P, Cn, Cn, tail.bias)
        wgrad(pl.dR16, Cn, Cn, pl.X[depth], Cn, Cn, h, w, tail.weight, ops.WMAP_OIHW, Cn, Cn)
        b.append(("mark", tail.weight))
        conv(h=h, w=w, cin=Cn, cout=Cn, srcs=src(pl.dR16, h, w, Cn), wpacked=pk["tail_d"].buf, out32=pl.gS32,
             ld_out32=Cn, out16=pl.gS16[depth % 2], ld_out16=Cn)
        for i in range(depth - 1, -1, -1):
            c1, c2 = trunk[i]
            gs_in, gs_out, dT = pl.gS16[(i + 1) % 2], pl.gS16[i % 2], pl.dT16[i % 2]
            bgrad(gs_in, P, Cn, Cn, c2.bias, mul=rs)
            wgrad(gs_in, Cn, Cn, pl.T[i], Cn, Cn, h, w, c2.weight, ops.WMAP_OIHW, Cn, Cn, mul=rs)
            conv(h=h, w=w, cin=Cn, cout=Cn, srcs=src(gs_in, h, w, Cn), wpacked=pk[f"b{i}c2_d"].buf, alpha=rs,
                 mask16=pl.T[i], ld_mask16=Cn, mask_mode=1, out16=dT, ld_out16=Cn)
            bgrad(dT, P, Cn, Cn, c1.bias)
            wgrad(dT, Cn, Cn, pl.X[i], Cn, Cn, h, w, c1.weight, ops.WMAP_OIHW, Cn, Cn)
            b.append(("mark", c1.weight))
            last = i == 0
            conv(h=h, w=w, cin=Cn, cout=Cn, srcs=src(dT, h, w, Cn), wpacked=pk[f"b{i}c1_d"].buf,
                 res32=pl.gS32, ld_res32=Cn, res16=pl.dR16 if last else None, ld_res16=Cn,
                 out32=None if last else pl.gS32, ld_out32=Cn, out16=pl.dF0 if last else gs_out, ld_out16=Cn)
        # embed (Cin = 3): wgrad against the saved im2col matrix; dgrad as a col2im GEMM (for sub_mean's gradients)
        bgrad(pl.dF0, P, Cn, Cn, g.embed.bias)
        wgrad(pl.dF0, Cn, Cn, pl.col_in, 64, 64, h, w, g.embed.weight, ops.WMAP_COL_IN, Cn, 3, taps=[(0, 0)])
        conv(h=h, w=w, cin=Cn, cout=32, taps=[(0, 0)], srcs=src(pl.dF0, h, w, Cn), wpacked=pk["embed_d"].buf,
             out32=pl.Zd, ld_out32=32)
        # a bias gradient (column sums of dY) followed by the 3x3 weight gradient of the same layer runs as ONE launch
        # (pesr_wgrad_reduce_bias); each fused call also zeroes the bias gradient the next one accumulates into
        fused, i = [], 0
        while i < len(b):
            op = b[i]
            nxt = b[i + 1] if i + 1 < len(b) else None
            if (_FUSE_BIAS and op[0] == "bias" and nxt is not None and nxt[0] == "wgrad" and op[1][5] == 0 and nxt[2][0] == 9
                    and nxt[2][3] in (ops.WMAP_OIHW, ops.WMAP_OIHW_PS) and op[1][2] % 8 == 0 and op[1][3] % 8 == 0
                    and 256 % (op[1][2] // 8) == 0 and nxt[2][5] * 36 <= 40 * 1024):
                fused.append(["wgrad_bias", nxt[1], nxt[2], nxt[3], op[1], op[2], None])
                i += 2
            else:
                fused.append(op)
                i += 1
        chain = [op for op in fused if op[0] == "wgrad_bias"]
        for k in range(len(chain) - 1):
            chain[k][6] = chain[k + 1][5]          # the bias parameter whose gradient this call zeroes for the next
        pl.first_fused_bias = chain[0][5] if chain else None
        pl.bwd = fused
        pl.scale = scale
        return pl

    # ------------------------------------------------------------------ forward
    def forward(self, lr, train, out_u8=False):
        """lr: fp32 NCHW [N,3,H,W] in 0..255, or (inference) a uint8 HWC batch [N,H,W,3] read directly by the first
        kernel (utils.imgs_to_tensors fused).  out_u8 (inference): return the uint8 HWC image batch [N,4H,4W,3] written
        by the last kernel (clip + round-half-even of utils.tensors_to_imgs fused) instead of the fp32 NCHW tensor."""
        g = self.gen
        u8_in = lr.dtype == torch.uint8
        if lr.dim() != 4 or (lr.shape[3] if u8_in else lr.shape[1]) != 3:
            raise ValueError(f"Generator expects [N,3,H,W] (or uint8 [N,H,W,3]), got {tuple(lr.shape)}")
        if (u8_in or out_u8) and train:
            raise ValueError("pesr_b200.Generator: uint8 input / output is an inference path (no autograd)")
        if g.n_resblock < 1:
            raise NotImplementedError("pesr_b200.Generator needs depth >= 1")
        lr = lr.contiguous() if u8_in else lr.contiguous().float()
        nb, h, w = (lr.shape[0], lr.shape[1], lr.shape[2]) if u8_in else (lr.shape[0], lr.shape[2], lr.shape[3])
        self._ensure_packed(lr.device)
        if not train and nb > 1 and nb * h * w > INFER_CHUNK_PIXELS:
            # inference over a large batch: images are independent, so the batch runs through ONE plan of a smaller
            # batch size (activation memory stays bounded; every chunk still fills the machine)
            per = max(1, INFER_CHUNK_PIXELS // (h * w))
            parts = [self.forward(lr[i:i + per], False, out_u8)[0] for i in range(0, nb, per)]
            return torch.cat(parts, dim=0), None
        pl = self._plan(nb, h, w, train)
        Cn = g.n_feats
        self.fwd_multi.run()
        up0, up2, up4 = g.upsample[0], g.upsample[2], g.upsample[4]
        self.bias_up0.copy_(_shuffle_perm(up0.bias.detach(), Cn))
        self.bias_up2.copy_(_shuffle_perm(up2.bias.detach(), Cn))
        sm_w = g.sub_mean.weight.detach().reshape(3, 3)
        am_w = g.add_mean.weight.detach().reshape(3, 3)
        if u8_in:
            ops.im2col3_u8(lr, pl.col_in, affine_a=sm_w, affine_b=g.sub_mean.bias.detach())
        else:
            ops.im2col3(lr, pl.col_in, affine_a=sm_w, affine_b=g.sub_mean.bias.detach())
        stream = torch.cuda.current_stream().cuda_stream
        for d in pl.fwd:
            _run_conv(d, stream)
        if out_u8:
            sr = torch.empty(nb, 4 * h, 4 * w, 3, device=lr.device, dtype=torch.uint8)
            ops.col2im3(pl.Z, 32, nb, 4 * h, 4 * w, None, bias=up4.bias.detach(), affine_a=am_w,
                        affine_b=g.add_mean.bias.detach(), sgn=1, out_u8=sr)
        else:
            sr = torch.empty(nb, 3, 4 * h, 4 * w, device=lr.device, dtype=torch.float32)
            ops.col2im3(pl.Z, 32, nb, 4 * h, 4 * w, sr, bias=up4.bias.detach(), affine_a=am_w,
                        affine_b=g.add_mean.bias.detach(), sgn=1, pre=pl.ypre)
        if self.trace_hook is not None:
            self.trace_hook(pl)
        if not train:
            return sr, None
        pl.busy = True
        return sr, (pl, _Release(pl), lr)

    # ------------------------------------------------------------------ backward
    def backward(self, state, dsr, need_input_grad=False):
        """Returns (grads: dict param -> fp32 gradient view in the reference layout, dlr or None)."""
        pl, _release, lr = state
        g = self.gen
        Cn = g.n_feats
        nb, h, w = pl.nb, pl.h, pl.w
        up4 = g.upsample[4]
        dsr = dsr.contiguous().float()
        dev = dsr.device
        self.bwd_multi.run()
        flat = self.flat_grads.get(dev)
        self.last_flat = flat
        off = self.offsets
        grads = {p: flat[off[p]:off[p] + p.numel()].view(p.shape) for p in self.param_list}
        hook = self.grad_hook
        mark_hi = [self.flat_numel]

        def mark(param):
            lo = off[param]
            if hook is not None and lo < mark_hi[0]:
                hook(lo, mark_hi[0], flat)
            mark_hi[0] = lo

        ws = self.scale_ws
        scale = pl.scale
        stream = torch.cuda.current_stream().cuda_stream
        # add_mean (1x1, trainable in the reference, model/basic.py:17) and the dynamic gradient scale
        am_w = g.add_mean.weight.detach().reshape(3, 3)
        am_wt = am_w.t().contiguous()
        ops.moments3(dsr, pl.ypre, pl.sums[:12])
        s = pl.sums
        grads[g.add_mean.weight].view(-1).copy_(s[0:9])
        grads[g.add_mean.bias].copy_(s[9:12])
        torch.mv(am_wt, s[9:12], out=grads[up4.bias])
        mark(up4.bias)
        ops.amax_scale(dsr, ws, target=16.0)
        ops.im2col3(dsr, pl.dcol, affine_a=am_wt, mul_dev=scale, sgn=-1)
        splits_out = C.c_int32(0)
        if pl.first_fused_bias is not None:
            grads[pl.first_fused_bias].zero_()
        lane = None
        if _WGRAD_STREAM:
            if self.lane is None:
                self.lane = SideLane(dev)
            lane = self.lane
            lane.begin(torch.cuda.current_stream())
            wstream = lane.stream.cuda_stream

        def leaf(op, st):
            kind = op[0]
            if kind == "wgrad":
                _, d, (ntaps, m_total, n_total, map_mode, co, ci, mul), param = op
                check(lib.pesr_conv_wgrad(C.byref(d), C.byref(splits_out), st), "pesr_conv_wgrad")
                check(lib.pesr_wgrad_reduce(pl.wg.data_ptr(), splits_out.value, ntaps, m_total, n_total, map_mode, co,
                                            ci, mul, scale.data_ptr(), 0, grads[param].data_ptr(), st),
                      "pesr_wgrad_reduce")
            elif kind == "wgrad_bias":
                _, d, (ntaps, m_total, n_total, map_mode, co, ci, mul), param, (x16, npix, c, ldc, bmul, _p), bparam, znext = op
                check(lib.pesr_conv_wgrad(C.byref(d), C.byref(splits_out), st), "pesr_conv_wgrad")
                zn = grads[znext] if znext is not None else None
                check(lib.pesr_wgrad_reduce_bias(pl.wg.data_ptr(), splits_out.value, ntaps, m_total, n_total, map_mode, co,
                                                 ci, mul, scale.data_ptr(), 0, grads[param].data_ptr(), x16.data_ptr(),
                                                 npix, c, ldc, bmul, self.dt, grads[bparam].data_ptr(),
                                                 zn.data_ptr() if zn is not None else 0, zn.numel() if zn is not None else 0,
                                                 st), "pesr_wgrad_reduce_bias")
            elif kind == "bias":
                _, (x16, npix, c, ldc, mul, perm_c), param = op
                if perm_c:
                    ops.colsum16(x16, npix, c, ldc, pl.bias_tmp[:c], mul=mul, div_dev=scale)
                    grads[param].copy_(_unshuffle_perm(pl.bias_tmp[:c], perm_c))
                else:
                    ops.colsum16(x16, npix, c, ldc, grads[param], mul=mul, div_dev=scale)
            else:
                mark(op[1])

        if lane is None:
            for op in pl.bwd:
                if op[0] == "conv":
                    _run_conv(op[1], stream)
                else:
                    leaf(op, stream)
        else:
            for op in pl.bwd:
                if op[0] == "conv":                     # the backward-data chain stays on the caller's stream
                    lane.before_chain()
                    _run_conv(op[1], stream)
                    lane.after_chain()
                else:                                   # weight / bias gradients and the data-parallel hooks: side lane
                    lane.enter_lane()
                    with torch.cuda.stream(lane.stream):
                        leaf(op, wstream)
                    if op[0] != "mark":
                        lane.after_lane()
            lane.join()
        ops.col2im3(pl.Zd, 32, nb, h, w, pl.dx_sm, mul=1.0, div_dev=scale, sgn=-1)
        ops.moments3(pl.dx_sm, lr, pl.sums[12:])
        grads[g.sub_mean.weight].view(-1).copy_(s[12:21])
        grads[g.sub_mean.bias].copy_(s[21:24])
        mark(self.param_list[0])
        if mark_hi[0] != 0:
            raise AssertionError("generator backward: gradient ranges did not cover the flat buffer")
        if hook is not None and self.grad_hook_flush is not None:
            self.grad_hook_flush()
        if hook is not None and self.grad_hook_finish is not None and not self.defer_finish:
            # autograd may copy the gradient views when it accumulates them: the reduced values must be in place
            self.grad_hook_finish()
        dlr = None
        if need_input_grad:
            sm_w = g.sub_mean.weight.detach().reshape(3, 3)
            dlr = torch.einsum("oi,nohw->nihw", sm_w, pl.dx_sm)
        return grads, dlr
