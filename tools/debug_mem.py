"""Memory trace of the config-5 batch sweep (bench.py side_measurements)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pesr_b200 import infer
from pesr_b200.model import Generator
dev = torch.device("cuda", 0)
G = Generator(bench.OPT).to(dev).eval()
def mem(tag):
    torch.cuda.synchronize()
    eng = G.engine()
    print(f"{tag}: allocated {torch.cuda.memory_allocated()/2**30:.2f} GiB reserved {torch.cuda.memory_reserved()/2**30:.2f} GiB plans {list(eng.plans.map.keys())}", flush=True)
mem("start")
for b in (1, 2, 4, 8, 16):
    xb = torch.rand(b, 3, 339, 510, device=dev) * 255
    out32, out8 = infer.super_resolve(G, xb)
    del out32, out8
    mem(f"b={b}")
