"""What do a few unavailable SMs cost the step?  A side stream pins n SMs (pesr_debug_sm_hog) for the whole timed
region - the situation a concurrent NCCL all-reduce creates - while the pretrain step runs on the main stream."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pesr_b200 import steps  # noqa: E402
from pesr_b200._lib import check, lib  # noqa: E402
from pesr_b200.model import Generator  # noqa: E402
from pesr_b200.optim import Adam  # noqa: E402
# bring-up hooks live in the debug build only: run with PESR_B200_LIB=pesr_b200/libpesr_b200_debug.so (tools/build_debug.sh)
from pesr_b200 import _debug as _dbg  # noqa: E402
_dbg.bind()

dev = torch.device("cuda", 0)
torch.manual_seed(0)
G = Generator(bench.OPT).to(dev)
opt = Adam(G.parameters(), lr=5e-5)
lr = torch.rand(16, 3, 48, 48, device=dev) * 255
hr = torch.rand(16, 3, 192, 192, device=dev) * 255
for _ in range(5):
    steps.pretrain_step(G, opt, lr, hr)
torch.cuda.synchronize()
side = torch.cuda.Stream()
# second column: the persistent kernels leave `reserve` SMs free (PESR_OPT_RESERVE_SMS, what parallel.DataParallel sets)
for reserve in (0, 2, 4, 8):
    lib.pesr_set_option(5, reserve)
    for n in (0, 1, 2, 4, 8, 16):
        if n:
            check(lib.pesr_debug_sm_hog(n, 400000, side.cuda_stream), "hog")      # 0.4 s
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            steps.pretrain_step(G, opt, lr, hr)
        e1.record()
        torch.cuda.synchronize()
        print(f"reserve {reserve}: {n:2d} SMs pinned: pretrain step {e0.elapsed_time(e1) / 10:.2f} ms", flush=True)

