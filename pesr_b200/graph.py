"""CUDA-graph replay of a whole training step.

The step bodies of pesr_b200.steps issue ~700 launches (kernels of this library plus a few torch element-wise ops) from
Python; the host needs 8-9 ms to enqueue them, which is the wall once the GPU step approaches that.  Every plan
(buffers, TMA descriptors, launch parameters) is static per input shape, programmatic-dependent-launch edges survive
stream capture, the optimiser keeps its step count on the device (pesr_b200.optim.Adam(capturable=True)) and the
gradient buffers are persistent (engine_g.FlatGrads), so one step can be captured once and replayed with a single
cudaGraphLaunch.

    step = GraphedStep(lambda lr, hr: steps.gan_step(G, D, vgg, optG, optD, lr, hr, cfg), (lr0, hr0), modules=(G, D, vgg),
                       optimizers=(optG, optD))
    losses = step(lr, hr)          # copies lr / hr into the static input buffers, replays, returns the static output

What a replay does NOT do is run Python: parameter `_version` counters do not move, so after every replay the engines
are told to forget which versions their packed 16-bit weights were made from (an eager forward that follows, e.g. the
validation pass, re-packs).  The learning rate lives in device memory; when a scheduler changes it the optimiser
refreshes that copy eagerly on the next call (no re-capture needed).
"""
import torch


class GraphedStep:
    def __init__(self, fn, example_inputs, modules=(), optimizers=(), warmup=3, pool=None):
        if not all(t.is_cuda for t in example_inputs):
            raise RuntimeError("GraphedStep: example inputs must be CUDA tensors")
        for opt in optimizers:
            for g in opt.param_groups:
                if not g.get('capturable', False):
                    raise RuntimeError("GraphedStep: optimizers must be pesr_b200.optim.Adam(..., capturable=True) "
                                       "(the step count must live on the device to be replayed)")
        self.fn = fn
        self.modules, self.optimizers = list(modules), list(optimizers)
        self.static_in = [torch.empty_like(t) for t in example_inputs]
        for s, t in zip(self.static_in, example_inputs):
            s.copy_(t)
        # warm-up on a side stream (torch.cuda.graph's requirement): builds every plan, packs, pointer table and
        # brings the host-side caches (weight versions, plan pools) into their steady-state pattern
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(2, warmup)):
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, pool=pool):
            self.static_out = fn(*self.static_in)
        self._lrs = self._current_lrs()
        self.replays = 0

    def _current_lrs(self):
        return [float(g['lr']) for opt in self.optimizers for g in opt.param_groups]

    def _sync_lrs(self):
        lrs = self._current_lrs()
        if lrs != self._lrs:
            for opt in self.optimizers:
                for gi, g in enumerate(opt.param_groups):
                    dev = opt._dev.get(gi)
                    if dev is not None and dev[2] != float(g['lr']):
                        dev[1].fill_(float(g['lr']))
                        dev[2] = float(g['lr'])
            self._lrs = lrs

    def __call__(self, *inputs):
        if len(inputs) != len(self.static_in):
            raise ValueError(f"GraphedStep: expected {len(self.static_in)} inputs")
        for s, t in zip(self.static_in, inputs):
            if t.shape != s.shape:
                raise ValueError(f"GraphedStep: input shape {tuple(t.shape)} differs from the captured {tuple(s.shape)}")
            if t.data_ptr() != s.data_ptr():
                s.copy_(t, non_blocking=True)
        self._sync_lrs()
        self.graph.replay()
        self.replays += 1
        for m in self.modules:            # see the module docstring
            eng = m.engine() if hasattr(m, "engine") else None
            if eng is not None:
                eng.invalidate_packs()
        return self.static_out
