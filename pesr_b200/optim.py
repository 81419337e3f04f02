"""Multi-tensor Adam on one kernel launch per step (torch.optim.Adam semantics, train.py:124-126)."""
import torch

from . import ops

_CHUNK = 1 << 16


class Adam(torch.optim.Optimizer):
    """Drop-in for ``optim.Adam(params, betas=(0.9, 0.999), lr=...)``: same state names (exp_avg, exp_avg_sq, step) so
    ``state_dict()`` is interchangeable with torch.optim.Adam's; no weight decay / amsgrad.

    capturable=True keeps the step count and the learning rate of every group in device memory (like
    torch.optim.Adam(capturable=True)): the launch carries no host-side scalar that changes from step to step, so a
    CUDA graph of the whole training step (pesr_b200.graph.GraphedStep) replays correctly."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, capturable=False):
        # the keys torch.optim.Adam keeps in its param_groups, at the values this kernel implements, so that
        # state_dict() loads into torch.optim.Adam and back
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False,
                                      foreach=None, capturable=capturable, differentiable=False, fused=None,
                                      decoupled_weight_decay=False))
        self._tables = {}
        self._dev = {}          # group index -> (step int32[1], lr float32[1], last lr uploaded)
        self.table_builds = 0   # how often the pointer table had to be rebuilt (0 or 1 per group in steady state)

    # ------------------------------------------------------------------ state interchange
    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        # the cached pointer tables hold the addresses of the OLD moment buffers
        self._tables.clear()
        self._dev.clear()

    def __setstate__(self, state):
        super().__setstate__(state)
        self._tables = {}
        self._dev = {}

    @staticmethod
    def _step_value(st):
        s = st.get('step', 0)
        return int(s.item()) if torch.is_tensor(s) else int(s)      # torch.optim.Adam stores `step` as a tensor

    def _table(self, gi, plist):
        # every address the kernel dereferences is part of the key
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]['exp_avg'].data_ptr(),
                     self.state[p]['exp_avg_sq'].data_ptr(), p.numel()) for p in plist)
        cached = self._tables.get(gi)
        if cached is not None and cached[0] == key:
            return cached[1], cached[2]
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("pesr_b200.optim.Adam: the parameter / gradient addresses changed inside a CUDA-graph "
                               "capture; run warm-up steps with the same (persistent) gradient buffers first")
        rows = []
        for (pp, gp, mp, vp, n) in key:
            for off in range(0, n, _CHUNK):
                cnt = min(_CHUNK, n - off)
                rows.append((pp + 4 * off, gp + 4 * off, mp + 4 * off, vp + 4 * off, cnt))
        host = torch.tensor(rows, dtype=torch.int64).pin_memory()
        table = host.to(plist[0].device, non_blocking=True)
        self._tables[gi] = (key, table, len(rows), host)      # `host` stays alive until the copy has certainly run
        self.table_builds += 1
        return table, len(rows)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group['params'] if p.grad is not None]
            if not plist:
                continue
            if group.get('weight_decay', 0) or group.get('amsgrad', False) or group.get('maximize', False):
                raise NotImplementedError("pesr_b200.optim.Adam: weight_decay / amsgrad / maximize are not on the PESR "
                                          "path (train.py:124-126)")
            for p in plist:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("pesr_b200.optim.Adam: parameters must be contiguous fp32 CUDA tensors")
                if not p.grad.is_contiguous():
                    p.grad = p.grad.contiguous()
                st = self.state[p]
                if 'exp_avg' not in st:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p)
                    st['exp_avg_sq'] = torch.zeros_like(p)
            b1, b2 = group['betas']
            if group.get('capturable', False):
                dev = self._dev.get(gi)
                if dev is None:
                    step0 = self._step_value(self.state[plist[0]])
                    dev = [torch.full((1,), step0, device=plist[0].device, dtype=torch.int32),
                           torch.full((1,), float(group['lr']), device=plist[0].device, dtype=torch.float32),
                           float(group['lr'])]
                    self._dev[gi] = dev
                    for p in plist:
                        self.state[p]['step'] = dev[0]      # one shared device counter (state_dict shows a tensor)
                if dev[2] != float(group['lr']):            # scheduler changed the rate: refresh the device copy
                    if torch.cuda.is_current_stream_capturing():
                        raise RuntimeError("pesr_b200.optim.Adam: learning rate changed inside a CUDA-graph capture")
                    dev[1].fill_(float(group['lr']))
                    dev[2] = float(group['lr'])
                dev[0].add_(1)
                table, n = self._table(gi, plist)
                ops.adam_multi_dev(table, n, dev[1], b1, b2, group['eps'], dev[0])
            else:
                steps = {self._step_value(self.state[p]) for p in plist}
                if len(steps) != 1:
                    raise RuntimeError("pesr_b200.optim.Adam: parameters of one group must share a step count")
                step = steps.pop() + 1
                for p in plist:
                    self.state[p]['step'] = step
                table, n = self._table(gi, plist)
                ops.adam_multi(table, n, group['lr'], b1, b2, group['eps'], step)
            for p in plist:   # the kernel wrote through raw pointers: tell autograd / the weight-pack caches
                torch.autograd.graph.increment_version(p)
        return loss
