"""Data parallelism: one process per GPU, gradients averaged over NVLink / NVSwitch, replacing the
reference's single-process nn.DataParallel (train.py:114-118).

Semantics kept from nn.DataParallel: every replica sees an equal shard of the batch, BatchNorm
statistics are per replica, and the optimiser sees the gradient of the global-batch loss: mean losses
become the average of the per-rank gradients; the TV term, a batch SUM in the reference
(train.py:137-140), is multiplied by world_size before averaging (pesr_b200.steps.gan_step).

Overlap: the Generator / Discriminator schedules emit their parameter gradients into one flat fp32
buffer that completes from its end towards its start; each completed range is handed to
``_on_range``, which launches the all-reduce of every full bucket on a high-priority side stream while the
remaining backward kernels keep the SMs busy; the tail bucket is launched at the end of backward and
``finish()`` (called right before optimizer.step()) is the only wait.

The reduction itself is the hand-written two-shot kernel of csrc/comm_ops.cu over a symmetric-memory copy of
the flat buffer (multimem.ld_reduce / multimem.st through the NVSwitch, or peer loads / stores), run on the SMs
the persistent tensor-core kernels leave free (PESR_OPT_RESERVE_SMS); NCCL carries the parameter broadcast,
the all-gather of the Linear(73728 -> 1024) gradient factors, and is the fallback when symmetric memory is
not available.
"""
import os

import torch
import torch.distributed as dist


_FC1_GATHER = os.environ.get("PESR_NO_FC1_GATHER") != "1"    # A/B knob: all-reduce the Linear weight gradient instead
_TRACE = os.environ.get("PESR_DDP_TRACE") == "1"             # event pairs around every wait of the compute stream
TRACE_EVENTS = []                                            # (label, e0, e1): bench.py reports their sum per step


def _traced_wait(label, events):
    """Make the current stream wait for `events`; with PESR_DDP_TRACE=1 the exposed wait is measured on the device."""
    cur = torch.cuda.current_stream()
    if _TRACE:
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
    for ev in events:
        cur.wait_event(ev)
    if _TRACE:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        TRACE_EVENTS.append((label, e0, e1))


def _reserve_sms_for_nccl():
    """The persistent tensor-core kernels size their grids to the whole machine with a static tile schedule: one SM taken
    by a concurrently running NCCL CTA sends the CTA that wanted it to a second wave (tools/sm_hog_probe.py: +20 % per
    step for ONE pinned SM).  So the kernels leave PESR_RESERVE_SMS SMs (default 4) to NCCL, and NCCL is told to use at
    most that many CTAs (NCCL_MAX_CTAS, read when the communicator is created: set it before init_process_group, see
    `nccl_env_defaults`)."""
    if not torch.cuda.is_available():
        return
    from ._lib import lib
    lib.pesr_set_option(5, int(os.environ.get("PESR_RESERVE_SMS", "4")))


def nccl_env_defaults():
    """Environment defaults for data-parallel training; call before torch.distributed.init_process_group."""
    n = os.environ.get("PESR_RESERVE_SMS", "4")
    os.environ.setdefault("NCCL_MAX_CTAS", n if int(n) > 0 else "32")
    os.environ.setdefault("NCCL_MIN_CTAS", "1")


_P2P = os.environ.get("PESR_DDP_NCCL_ONLY") != "1"       # A/B knob: reduce through NCCL instead of the peer-memory kernel
_SYMM_NCCL = os.environ.get("PESR_DDP_SYMM_NCCL") == "1"  # A/B knob: symmetric buffer, but reduced by NCCL
_P2P_CTAS = int(os.environ.get("PESR_DDP_CTAS", "0"))    # CTAs of the peer-memory all-reduce (0: the reserved SMs)


class _SymmetricFlat:
    """The flat gradient buffer of one network in symmetric memory (torch.distributed._symmetric_memory: one allocation per
    rank, every rank's allocation mapped into every process, NVSwitch multicast address when the fabric has one) and the
    hand-written all-reduce over it (csrc/comm_ops.cu)."""

    def __init__(self, numel, device, group, world, rank):
        import ctypes as C

        import torch.distributed._symmetric_memory as symm_mem
        self.tensor = symm_mem.empty(numel, dtype=torch.float32, device=device)
        name = group.group_name if group is not None else dist.group.WORLD.group_name
        self.hdl = symm_mem.rendezvous(self.tensor, name)
        self.world, self.rank = world, rank
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.ptrs = (C.c_uint64 * world)(*ptrs)
        self.pads = None
        if int(self.hdl.signal_pad_size) >= 1280 and os.environ.get("PESR_DDP_TORCH_BARRIER") != "1":
            self.pads = (C.c_uint64 * world)(*[int(p) for p in self.hdl.signal_pad_ptrs])
        self.epoch = 0
        mc = int(self.hdl.multicast_ptr or 0)         # 0 when the fabric has no multicast (then: peer loads / stores)
        self.multicast = mc if os.environ.get("PESR_DDP_NO_MULTIMEM") != "1" else 0
        self.ctas = _P2P_CTAS or max(1, int(os.environ.get("PESR_RESERVE_SMS", "4")))

    def all_reduce_avg(self, lo, hi, stream):
        """Average [lo, hi) of the buffer over the ranks, in place, on `stream` (the current stream)."""
        from ._lib import check, lib
        self.epoch = (self.epoch + 1) & 0x7FFFFFFF
        if self.pads is None:
            self.hdl.barrier(channel=0)      # every rank's backward has produced the range
        check(lib.pesr_allreduce_p2p(self.ptrs, self.pads, self.world, self.rank, self.multicast, lo, hi - lo,
                                     1.0 / self.world, self.ctas, self.epoch, stream.cuda_stream), "pesr_allreduce_p2p")
        if self.pads is None:
            self.hdl.barrier(channel=0)      # every owner has stored its slice everywhere


def _new_comm_stream():
    """High priority: when an SM frees up, the block scheduler places NCCL's pending CTAs before the next convolution's."""
    if os.environ.get("PESR_DDP_LOW_PRIORITY") == "1":       # A/B knob
        return torch.cuda.Stream()
    return torch.cuda.Stream(priority=-1)


class DataParallel(torch.nn.Module):
    """Wraps a pesr_b200 network; exposes ``.module`` like nn.DataParallel (train.py:303,309)."""

    def __init__(self, module, process_group=None, bucket_mb=None, defer_finish=True):
        super().__init__()
        self.module = module
        self.pg = process_group
        self.world_size = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        if bucket_mb is None:
            bucket_mb = float(os.environ.get("PESR_DDP_BUCKET_MB", "8"))
        self.bucket_elems = int(bucket_mb * (1 << 20) // 4)
        self.label = type(module).__name__
        self._pending = None      # (lo, hi, flat) accumulated, not yet reduced
        self._works = []
        self._comm_stream = None
        self._engine = None
        self._symm = None         # _SymmetricFlat once the engine has asked for its flat buffer
        self._symm_failed = None
        if self.world_size > 1:
            self._broadcast_parameters()
            eng = getattr(module, "engine", None)
            if callable(eng):
                e = module.engine()
                e.grad_hook = self._on_range
                e.grad_hook_finish = self.finish
                e.grad_hook_flush = self._flush     # end of backward: launch the tail bucket now, wait for it in finish()
                # the step bodies (pesr_b200.steps) call finish() right before optimizer.step(): the kernels issued
                # between backward and that call (e.g. the Generator-phase VGG forward while D's 321 MB are on the wire)
                # overlap the all-reduce instead of waiting for it inside backward
                e.defer_finish = defer_finish
                self._engine = e
                if _P2P and torch.cuda.is_available() and hasattr(e, "flat_alloc"):
                    e.flat_alloc = self._alloc_flat
                    if getattr(e, "flat_grads", None) is not None:      # engine already built: re-home its buffer
                        e.flat_grads.alloc, e.flat_grads.buf = self._alloc_flat, None
                if hasattr(e, "fc1_gather") and _FC1_GATHER:
                    e.fc1_gather = self._gather_rows
                    e.grad_hook_skip = self._skip_range
            _reserve_sms_for_nccl()

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def forward_pair(self, a, b):
        return self.module.forward_pair(a, b)

    def _broadcast_parameters(self):
        with torch.no_grad():
            for t in list(self.module.parameters()) + list(self.module.buffers()):
                dist.broadcast(t, src=0, group=self.pg)

    # ---- called by the kernel schedules during backward
    def _on_range(self, lo, hi, flat):
        if self._pending is None:
            self._pending = [lo, hi, flat]
        else:
            self._pending[0] = lo
        if self._pending[1] - self._pending[0] >= self.bucket_elems:
            self._flush()

    def _alloc_flat(self, numel, device):
        """FlatGrads.alloc: the persistent gradient buffer lives in symmetric memory (a collective call: every rank reaches
        its first backward of this network at the same point of the step).  Falls back to an ordinary tensor + NCCL."""
        if self._symm is None and self._symm_failed is None:
            try:
                rank = dist.get_rank(self.pg)
                self._symm = _SymmetricFlat(numel, device, self.pg, self.world_size, rank)
            except Exception as e:        # no symmetric memory on this system: NCCL all-reduce
                self._symm_failed = f"{type(e).__name__}: {str(e)[:200]}"
                import sys
                print(f"[pesr_b200.parallel] symmetric memory unavailable ({self._symm_failed}); using NCCL all-reduce",
                      file=sys.stderr, flush=True)
        if self._symm is not None and self._symm.tensor.numel() == numel:
            return self._symm.tensor
        return torch.empty(numel, device=device, dtype=torch.float32)

    def _skip_range(self, lo, hi):
        """[lo, hi) of the flat buffer is filled by the engine with an already averaged gradient: reduce what is pending
        above it and leave the range alone."""
        self._flush()

    def _gather_rows(self, dz1, flat7):
        """all-gather the two factors of the Linear weight gradient on the communication stream.  Returns
        (dz1 of all ranks [world*rows][1024], flat7 of all ranks, event, world)."""
        if self._comm_stream is None:
            self._comm_stream = _new_comm_stream()
        w = self.world_size
        dz1, flat7 = dz1.contiguous(), flat7.contiguous()
        dz_all = torch.empty(w * dz1.shape[0], dz1.shape[1], device=dz1.device, dtype=dz1.dtype)
        f_all = torch.empty(w * flat7.shape[0], flat7.shape[1], device=flat7.device, dtype=flat7.dtype)
        ev = torch.cuda.Event()
        ev.record()
        self._comm_stream.wait_event(ev)
        with torch.cuda.stream(self._comm_stream):
            dist.all_gather_into_tensor(dz_all, dz1, group=self.pg)
            dist.all_gather_into_tensor(f_all, flat7, group=self.pg)
            done = torch.cuda.Event()
            done.record()
        for t in (dz1, flat7, dz_all, f_all):
            t.record_stream(self._comm_stream)
        return dz_all, f_all, (lambda: _traced_wait(self.label + " factor gather", [done])), w

    def _flush(self):
        if self._pending is None:
            return
        lo, hi, flat = self._pending
        self._pending = None
        if self.world_size == 1 or hi <= lo:
            return
        if flat.is_cuda:
            if self._comm_stream is None:
                self._comm_stream = _new_comm_stream()
            ev = torch.cuda.Event()
            ev.record()
            self._comm_stream.wait_event(ev)
            with torch.cuda.stream(self._comm_stream):
                if self._symm is not None and flat.data_ptr() == self._symm.tensor.data_ptr() and not _SYMM_NCCL:
                    self._symm.all_reduce_avg(lo, hi, self._comm_stream)
                else:
                    dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.AVG, group=self.pg)
                done = torch.cuda.Event()
                done.record()
            self._works.append(done)
        else:  # gloo (CPU tests): no AVG op, no streams
            w = dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
            self._works.append((w, flat, lo, hi))

    def finish(self):
        """Make the current stream wait for every outstanding bucket (call before optimizer.step())."""
        self._flush()
        evs = []
        for w in self._works:
            if isinstance(w, tuple):
                work, flat, lo, hi = w
                work.wait()
                flat[lo:hi].div_(self.world_size)
            else:
                evs.append(w)
        if evs:
            _traced_wait(self.label + " all-reduce", evs)
        had_work = bool(self._works)
        self._works = []
        if had_work:
            self._sync_grad_views()

    def _sync_grad_views(self):
        """The all-reduce ran in place on the engine's flat buffer.  Autograd normally keeps the views it was handed as
        `.grad` (then there is nothing to do); if it copied them instead, bring the reduced values to the copies."""
        e = self._engine
        flat = getattr(e, "last_flat", None) if e is not None else None
        if flat is None:
            return
        off = e.offsets
        plist = [p for p in e.param_list if p.grad is not None]
        if not plist:
            return
        base = flat.data_ptr()
        if all(p.grad.data_ptr() == base + 4 * off[p] for p in (plist[0], plist[-1])):
            return
        with torch.no_grad():
            for p in plist:
                if p.grad.data_ptr() != base + 4 * off[p]:
                    p.grad.copy_(flat[off[p]:off[p] + p.numel()].view_as(p))

    # ---- fallback for modules without a flat-gradient schedule
    def allreduce_grads(self):
        if self.world_size == 1:
            return
        grads = [p.grad for p in self.module.parameters() if p.grad is not None]
        for g in grads:
            if g.is_cuda:
                dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self.pg)
            else:
                dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.pg)
                g.div_(self.world_size)
