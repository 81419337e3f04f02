"""Multi-tensor Adam on one kernel launch per step (torch.optim.Adam semantics, train.py:124-126)."""
import torch

from . import ops

_CHUNK = 1 << 16


class Adam(torch.optim.Optimizer):
    """Drop-in for ``optim.Adam(params, betas=(0.9, 0.999), lr=...)``: same state names (exp_avg,
    exp_avg_sq, step) so ``state_dict()`` stays interchangeable; no weight decay / amsgrad."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._tables = {}

    def _table(self, gi, plist):
        key = tuple((p.data_ptr(), p.grad.data_ptr(), p.numel()) for p in plist)
        cached = self._tables.get(gi)
        if cached is not None and cached[0] == key:
            return cached[1], cached[2]
        rows = []
        for p in plist:
            st = self.state[p]
            pp, gp, mp, vp = p.data_ptr(), p.grad.data_ptr(), st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr()
            n = p.numel()
            for off in range(0, n, _CHUNK):
                cnt = min(_CHUNK, n - off)
                rows.append((pp + 4 * off, gp + 4 * off, mp + 4 * off, vp + 4 * off, cnt))
        table = torch.tensor(rows, dtype=torch.int64).pin_memory().to(plist[0].device, non_blocking=True)
        self._tables[gi] = (key, table, len(rows))
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
            for p in plist:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("pesr_b200.optim.Adam: parameters must be contiguous fp32 CUDA tensors")
                if not p.grad.is_contiguous():
                    p.grad = p.grad.contiguous()
                st = self.state[p]
                if not st:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p)
                    st['exp_avg_sq'] = torch.zeros_like(p)
                st['step'] += 1
            step = self.state[plist[0]]['step']
            if any(self.state[p]['step'] != step for p in plist):
                raise RuntimeError("pesr_b200.optim.Adam: parameters of one group must share a step count")
            table, n = self._table(gi, plist)
            b1, b2 = group['betas']
            ops.adam_multi(table, n, group['lr'], b1, b2, group['eps'], step)
            for p in plist:   # the kernel wrote through raw pointers: tell autograd / the weight-pack caches
                torch.autograd.graph.increment_version(p)
        return loss
