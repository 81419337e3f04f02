#!/usr/bin/env python
"""Headline benchmark: PESR training step throughput (samples/s) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload pretrain|gan] [--impl reference]

A "step" is one pass of the training hot path over one synthetic batch of 16 patches per GPU
(48x48 LR -> 192x192 HR, the configs of BASELINE.json): `pretrain` = train.py:164-176 (G fwd, L1,
bwd, Adam), `gan` = train.py:202-259 (D phase + G phase with VGG/TV/focal-RSGAN losses, two Adams).
Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the reference
(oracle/, the reference itself is a torch-0.4-era script tree that cannot be installed or shipped) on the
host cores for the same metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

G_FWD_GFLOP = 231.564   # per 48x48 sample, BASELINE.md section 2
D_FWD_GFLOP = 7.073
VGG_FWD_GFLOP = 28.665
NCU_TRUNK_CONV_DRAM_BYTES = 20095232   # ncu --set full, profiles/r01_ncu_conv_igemm.txt
OPT = {'patch_size': 48, 'num_channels': 256, 'depth': 32, 'res_scale': 0.1, 'spectral_norm': False}
BATCH = 16


def step_gflop(workload):
    if workload == "pretrain":
        return BATCH * 3 * G_FWD_GFLOP
    return BATCH * (3 * G_FWD_GFLOP + 9 * D_FWD_GFLOP + 3 * VGG_FWD_GFLOP)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                d = json.load(f)
            return dict(tflops=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), hbm=float(d["hbm_gbs"]),
                        source="MEASURED_PEAKS.json bf16_tflops_sustained (of measured)")
        except Exception:
            pass
    return dict(tflops=1400.0, hbm=6650.0, source="B200_PROFILING.md fallback, sustained (of fallback)")


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port)
# ------------------------------------------------------------------------------------------------
def cpu_step_time(workload, batch, reps=1, warm=0):
    """Seconds per oracle step (reference modules restated in oracle/pesr_oracle.py) at `batch` samples."""
    import torch
    from oracle import pesr_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    g_sd = O.init_generator(OPT, 0)
    gen = torch.Generator().manual_seed(0)
    lr = torch.rand(batch, 3, 48, 48, generator=gen) * 255
    hr = torch.rand(batch, 3, 192, 192, generator=gen) * 255
    if workload == "gan":
        d_sd, v_sd = O.init_discriminator(OPT, 1), O.init_vgg(2)

    def one():
        if workload == "pretrain":
            _, _, grads = O.pretrain_step(g_sd, lr, hr, OPT)
            with torch.no_grad():
                for k, gk in grads.items():     # the Adam update of train.py:173 (first step)
                    O.adam_update(g_sd[k], gk, torch.zeros_like(gk), torch.zeros_like(gk), 1, 5e-5)
        else:
            out = O.gan_step(g_sd, d_sd, v_sd, lr, hr, OPT)
            with torch.no_grad():
                for k, gk in out['g_grads'].items():
                    if gk is not None:
                        O.adam_update(g_sd[k], gk, torch.zeros_like(gk), torch.zeros_like(gk), 1, 5e-5)
    for _ in range(warm):
        one()
    t0 = time.perf_counter()
    for _ in range(reps):
        one()
    return (time.perf_counter() - t0) / reps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    t_probe = cpu_step_time(args.workload, 1)
    budget = 150.0
    total_steps = args.steps + args.warmup
    batch = int(max(1, min(BATCH, budget / max(t_probe, 1e-3) / max(total_steps, 1))))
    for _ in range(args.warmup):
        cpu_step_time(args.workload, batch)
    t = cpu_step_time(args.workload, batch, reps=args.steps)
    value = batch / t
    sample = f"{args.steps} steps of a {batch}-sample batch (of {BATCH}) after {args.warmup} warm-up steps, fp32, torch CPU"
    line = {
        "impl": "reference", "metric": metric_name(args.workload), "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3 * BATCH / batch, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.workload, args.gpus),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample,
                         "threads": torch.get_num_threads()},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def metric_name(workload):
    return "GAN train samples/s" if workload == "gan" else "L1 pretrain samples/s"


def config_dict(workload, n):
    name = ("GAN fine-tune step: Generator + Discriminator + VGG19 perceptual loss + RaGAN focal loss, batch 16 at 48x48 LR"
            if workload == "gan" else
            "L1 pretrain phase: Generator fwd+bwd, batch 16 of 48x48 LR -> 192x192 HR synthetic patches")
    return {"workload": name, "per_gpu_batch": BATCH, "global_batch": BATCH * n, "lr_patch": 48, "hr_patch": 192,
            "num_channels": 256, "num_blocks": 32, "parallelism": f"dp{n}",
            "l2": "step working set (>1.2 GB of activations per step) exceeds the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import pesr_b200._lib as L
    from pesr_b200 import steps
    from pesr_b200.model import Generator
    from pesr_b200.optim import Adam
    from pesr_b200.parallel import DataParallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.manual_seed(0)
    G = Generator(OPT).to(dev)
    ddp_g = ddp_d = None
    Gw = G
    if world > 1:
        Gw = DataParallel(G)
        ddp_g = Gw
    optim_G = Adam([p for p in G.parameters() if p.requires_grad], lr=5e-5, betas=(0.9, 0.999))
    cfg = None
    if args.workload == "gan":
        from pesr_b200.model import VGG, Discriminator
        D = Discriminator(OPT).to(dev)
        vgg = VGG(pretrained=False).to(dev)
        Dw = D
        if world > 1:
            Dw = DataParallel(D)
            ddp_d = Dw
        optim_D = Adam(D.parameters(), lr=5e-5, betas=(0.9, 0.999))
        cfg = dict(steps.DEFAULT_GAN_CFG)
        cfg['target_real'] = torch.ones(BATCH, 1, device=dev)
        cfg['target_fake'] = torch.zeros(BATCH, 1, device=dev)

    gen = torch.Generator().manual_seed(1234 + rank)
    n_host = 4
    host_lr = [(torch.rand(BATCH, 3, 48, 48, generator=gen) * 255).pin_memory() for _ in range(n_host)]
    host_hr = [(torch.rand(BATCH, 3, 192, 192, generator=gen) * 255).pin_memory() for _ in range(n_host)]
    dev_lr = [t.to(dev) for t in host_lr]
    dev_hr = [t.to(dev) for t in host_hr]

    def do_step(lr, hr):
        if args.workload == "gan":
            return steps.gan_step(Gw, Dw, vgg, optim_G, optim_D, lr, hr, cfg, ddp_g=ddp_g, ddp_d=ddp_d)
        return steps.pretrain_step(Gw, optim_G, lr, hr, ddp=ddp_g)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = [0.0]

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- resident-input arm (value)
    for i in range(max(args.warmup, 3)):
        do_step(dev_lr[i % n_host], dev_hr[i % n_host])
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    L.launch_count(reset=True)
    ms_total = timed(lambda i: do_step(dev_lr[i % n_host], dev_hr[i % n_host]), args.steps)
    launches = L.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    # CPU time to ENQUEUE one step into an empty stream (no launch-queue back-pressure): GPU-bound if < ms_per_step
    for i in range(3):
        barrier()
        h0 = time.perf_counter()
        do_step(dev_lr[i % n_host], dev_hr[i % n_host])
        host_ms[0] = (time.perf_counter() - h0) * 1e3 if i == 0 else min(host_ms[0], (time.perf_counter() - h0) * 1e3)
    barrier()
    host_enqueue_ms = host_ms[0]

    # ---- end-to-end arm: pinned host batch -> H2D -> step -> D2H loss, every step
    last = {}

    def e2e_step(i):
        lr = host_lr[i % n_host].to(dev, non_blocking=True)
        hr = host_hr[i % n_host].to(dev, non_blocking=True)
        out = do_step(lr, hr)
        last['loss'] = out.detach().float().cpu()   # device -> host read of the step's result
    for i in range(2):
        e2e_step(i)
    ms_e2e = timed(e2e_step, args.steps) / args.steps
    h2d = host_lr[0].numel() * 4 + host_hr[0].numel() * 4
    d2h = int(last['loss'].numel()) * 4

    # ---- roofline leg: same steps with per-launch CUDA events around the tensor-core kernels
    L.profile_enable(True)
    L.profile_read(0), L.profile_read(1)
    for i in range(args.steps):
        do_step(dev_lr[i % n_host], dev_hr[i % n_host])
    ig_ms, ig_n, ig_fl = L.profile_read(0)
    wg_ms, wg_n, wg_fl = L.profile_read(1)
    L.profile_enable(False)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    value = BATCH * world / (ms_step * 1e-3)
    e2e_value = BATCH * world / (ms_e2e * 1e-3)
    achieved = ig_fl / (ig_ms * 1e-3) / 1e12 if ig_ms > 0 else 0.0
    wg_achieved = wg_fl / (wg_ms * 1e-3) / 1e12 if wg_ms > 0 else 0.0
    step_tflops = step_gflop(args.workload) / ms_step   # GFLOP / ms == TFLOP/s
    roofline = {
        "bound": "tensor", "kernel": "conv_igemm_kernel (fprop + dgrad of every conv)", "achieved": achieved,
        "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
        "traffic": NCU_TRUNK_CONV_DRAM_BYTES,
        "traffic_note": "dram__bytes_read+write of ONE trunk-conv launch (16x48x48, 256->256) from profiles/r01_ncu_conv_igemm.txt; "
                        "algorithmic bytes of that launch: 18.9 MB in + 1.2 MB weights + 18.9 MB out (the output stays in the 126 MB L2)",
        "peak_source": peaks["source"], "launches_per_step": ig_n / args.steps,
        "avg_launch_us": ig_ms * 1e3 / max(ig_n, 1), "share_of_step": ig_ms / args.steps / ms_step,
        "wgrad": {"kernel": "conv_wgrad_kernel", "achieved": wg_achieved, "frac": wg_achieved / peaks["tflops"],
                  "launches_per_step": wg_n / args.steps, "share_of_step": wg_ms / args.steps / ms_step},
        "whole_step": {"achieved": step_tflops, "frac": step_tflops / peaks["tflops"],
                       "algorithmic_gflop_per_step": step_gflop(args.workload)},
        "how": "algorithmic FLOPs (2*M*N*K, dense convention) of every launch / CUDA-event time of that launch, "
               "summed over a repeat of the timed steps with per-launch events enabled",
    }
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cb = 2
        t = cpu_step_time(args.workload, cb)
        cpu = {"value": cb / t, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"1 step of a {cb}-sample batch (of {BATCH}), fp32, torch CPU oracle, {t:.1f} s"}
    extras = {}
    if world == 1 and not args.no_extras:
        extras = side_measurements(args, dev, G, optim_G, dev_lr, dev_hr)
    line = {
        "metric": metric_name(args.workload), "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16", "data": "synthetic", "config": config_dict(args.workload, world),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e},
        "gpu_launches": launches, "host_enqueue_ms_per_step": host_enqueue_ms, "roofline": roofline, "cpu_baseline": cpu,
    }
    line.update(extras)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def side_measurements(args, dev, G, optim_G, dev_lr, dev_hr):
    """The other single-GPU configurations of BASELINE.json, measured the same way (CUDA events, warm-up) and
    reported beside the headline: config 2 (L1 pretrain step) and configs 1 / 5 (x4 inference, alpha = 1)."""
    import torch
    from pesr_b200 import infer, steps
    out = {}

    def timed(fn, k, warm=3):
        for i in range(warm):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k
    if args.workload != "pretrain":
        ms = timed(lambda i: steps.pretrain_step(G, optim_G, dev_lr[i % len(dev_lr)], dev_hr[i % len(dev_hr)]), args.steps)
        out["pretrain_step"] = {"metric": "L1 pretrain samples/s", "value": BATCH / (ms * 1e-3), "ms_per_step": ms,
                                "tflops": step_gflop("pretrain") / ms, "config": "BASELINE.json configs[1]"}
    G.eval()
    inf = {}
    for name, (h, w) in (("128x128", (128, 128)), ("339x510", (339, 510))):
        x = torch.rand(1, 3, h, w, device=dev) * 255
        ms = timed(lambda i: infer.super_resolve(G, x), 5, warm=2)
        inf[name] = {"ms_per_image": ms, "hr_mpix_per_s": 16 * h * w / (ms * 1e-3) / 1e6,
                     "tflops": h * w * 100505088 / (ms * 1e-3) / 1e12}
    G.train()
    out["inference_alpha1"] = {"metric": "x4 SR inference HR Mpix/s (fp32 image in, uint8 image out, batch 1)", **inf,
                               "config": "BASELINE.json configs[0] and configs[4] image sizes"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("PESR_BENCH_WORKLOAD", "gan"), choices=["pretrain", "gan"])
    ap.add_argument("--no-extras", action="store_true", help="skip the pretrain-step and inference side measurements")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
