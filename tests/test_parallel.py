"""CPU (gloo, world_size 2) test of the data-parallel gradient exchange: the bucketing hook fed by the kernel
schedules must leave the AVERAGE of the ranks' flat gradient buffers in place, whatever the range sizes."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeEngine:
    grad_hook = None
    grad_hook_finish = None
    grad_hook_flush = None
    grad_hook_skip = None
    fc1_gather = None          # parallel.DataParallel installs the factor all-gather / skip hooks when this attribute exists


class _FakeNet(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(7))
        self._e = _FakeEngine()

    def engine(self):
        return self._e


def _worker(rank, world, port, n, cuts, bucket_mb, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pesr_b200.parallel import DataParallel
        net = _FakeNet()
        with torch.no_grad():
            net.w.fill_(float(rank + 3))
        ddp = DataParallel(net, bucket_mb=bucket_mb)
        assert ddp.world_size == world and ddp.module is net
        assert float(net.w[0]) == 3.0                      # parameters broadcast from rank 0
        flat = torch.arange(n, dtype=torch.float32) * (rank + 1)
        hook = net.engine().grad_hook
        hi = n
        for lo in cuts:                                    # ranges complete from the end towards the start
            hook(lo, hi, flat)
            hi = lo
        net.engine().grad_hook_finish()
        expect = torch.arange(n, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
        ok = torch.allclose(flat, expect)
        # a range the engine fills with an already averaged gradient (the Discriminator's Linear weight, formed from the
        # all-gathered factors): everything around it is reduced, the range itself is left alone; the tail bucket is
        # launched by the end-of-backward flush and only waited for in finish()
        flat2 = torch.arange(n, dtype=torch.float32) * (rank + 1)
        eng = net.engine()
        eng.grad_hook(800, n, flat2)
        eng.grad_hook_skip(300, 800)
        eng.grad_hook(100, 300, flat2)
        eng.grad_hook(0, 100, flat2)
        eng.grad_hook_flush()
        ok = ok and ddp._pending is None
        eng.grad_hook_finish()
        expect2 = expect.clone()
        expect2[300:800] = torch.arange(300, 800, dtype=torch.float32) * (rank + 1)
        ok = ok and torch.allclose(flat2, expect2)
        # fallback path for modules without a flat-gradient schedule
        net.w.grad = torch.full((7,), float(rank))
        ddp.allreduce_grads()
        ok = ok and torch.allclose(net.w.grad, torch.full((7,), (world - 1) / 2))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bucket_mb,cuts", [(1e-4, [900, 500, 499, 10, 0]), (64, [700, 0]), (1e-3, [0])])
def test_bucketed_allreduce_gloo_world2(bucket_mb, cuts):
    world, n = 2, 1000
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, cuts, bucket_mb, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(out.get(r) for r in range(world)), dict(out)


def test_nccl_env_defaults_and_flat_buffer_allocator(monkeypatch):
    """parallel.nccl_env_defaults: NCCL is confined to the SMs the persistent kernels leave free unless the user said
    otherwise; engine_g.FlatGrads takes its persistent buffer from the allocator DataParallel installs (symmetric memory on
    the GPU) and keeps handing out fresh buffers while the previous hand-out is unconsumed."""
    sys.path.insert(0, ROOT)
    from pesr_b200 import parallel
    from pesr_b200.engine_g import FlatGrads
    for k in ("NCCL_MAX_CTAS", "NCCL_MIN_CTAS", "PESR_RESERVE_SMS"):
        monkeypatch.delenv(k, raising=False)
    parallel.nccl_env_defaults()
    assert os.environ["NCCL_MAX_CTAS"] == "4" and os.environ["NCCL_MIN_CTAS"] == "1"
    monkeypatch.setenv("NCCL_MAX_CTAS", "16")
    parallel.nccl_env_defaults()
    assert os.environ["NCCL_MAX_CTAS"] == "16"            # an explicit setting wins
    params = [torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(3, 3))]
    calls = []

    def alloc(numel, device):
        calls.append(numel)
        return torch.zeros(numel, device=device)
    fg = FlatGrads(params, alloc=alloc)
    assert fg.offsets[params[1]] == 8 and fg.numel == 20      # every tensor starts on a 16-byte boundary
    a = fg.get(torch.device("cpu"))
    assert calls == [20] and a.numel() == 20
    b = fg.get(torch.device("cpu"))                            # no optimiser step in between: a fresh buffer, not `a`
    assert b.data_ptr() != a.data_ptr() and calls == [20]
    with torch.no_grad():
        params[0].add_(1.0)                                     # what an optimiser step does to the version counter
    assert fg.get(torch.device("cpu")).data_ptr() == a.data_ptr()
