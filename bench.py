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
    """Roofline denominators: the driver-written MEASURED_PEAKS.json (cuBLAS bf16 8192^3: best-of-10 "burst" and 4 s
    back-to-back "sustained"; torch copy bandwidth), else the fallback of B200_PROFILING.md."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                d = json.load(f)
            return dict(tflops=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), burst=float(d["bf16_tflops"]),
                        hbm=float(d["hbm_gbs"]), source="MEASURED_PEAKS.json bf16_tflops_sustained (of measured)")
        except Exception:
            pass
    return dict(tflops=1400.0, burst=1650.0, hbm=6650.0, source="B200_PROFILING.md fallback, sustained (of fallback)")


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
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                try:
                    pw.append(float(r[2]))
                except Exception:
                    pass
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        pw.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_median": pw[len(pw) // 2] if pw else None,
                "power_w_max": pw[-1] if pw else None}


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
    """CPU arm: the oracle port on every host core, FULL 16-sample batches (the same config as the B200 arm), `steps`
    timed steps after `warmup` untimed ones; no rescaling of the measured time."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    batch = BATCH
    t_probe = cpu_step_time(args.workload, 1)
    est = t_probe * BATCH * 0.8 * (args.steps + args.warmup)
    note = ""
    if est > 1500.0:      # only if someone asks for hundreds of CPU steps: keep the run bounded and say so
        batch = int(max(1, min(BATCH, 1500.0 / (t_probe * 0.8 * (args.steps + args.warmup)))))
        note = f" (batch reduced from {BATCH}: {args.steps + args.warmup} full-batch steps would take ~{est:.0f} s)"
    for _ in range(args.warmup):
        cpu_step_time(args.workload, batch)
    t = cpu_step_time(args.workload, batch, reps=args.steps)
    value = batch / t
    sample = (f"{args.steps} steps of a {batch}-sample batch after {args.warmup} warm-up steps, fp32, torch CPU "
              f"({torch.get_num_threads()} threads){note}")
    line = {
        "impl": "reference", "metric": metric_name(args.workload), "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.workload, args.gpus),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample,
                         "threads": torch.get_num_threads(), "batch": batch},
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
def _event_timed(fn, k, barrier=None):
    """CUDA-event time (ms) of k calls of fn(i), synchronised on both sides."""
    import torch
    (barrier or torch.cuda.synchronize)()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(k):
        fn(i)
    e1.record()
    (barrier or torch.cuda.synchronize)()
    return e0.elapsed_time(e1)


def run_b200(args):
    import torch
    import torch.distributed as dist
    import pesr_b200._lib as L
    from pesr_b200 import steps
    from pesr_b200.graph import GraphedStep
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
        from pesr_b200.parallel import nccl_env_defaults
        nccl_env_defaults()       # NCCL_MAX_CTAS = the SMs the persistent kernels leave free (pesr_b200/parallel.py)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.manual_seed(0)
    G = Generator(OPT).to(dev)
    ddp_g = ddp_d = None
    Gw = G
    if world > 1:
        Gw = DataParallel(G)
        ddp_g = Gw
    optim_G = Adam([p for p in G.parameters() if p.requires_grad], lr=5e-5, betas=(0.9, 0.999), capturable=True)
    cfg, D, vgg, optim_D = None, None, None, None
    if args.workload == "gan":
        from pesr_b200.model import VGG, Discriminator
        D = Discriminator(OPT).to(dev)
        vgg = VGG(pretrained=False).to(dev)
        Dw = D
        if world > 1:
            Dw = DataParallel(D)
            ddp_d = Dw
        optim_D = Adam(D.parameters(), lr=5e-5, betas=(0.9, 0.999), capturable=True)
        cfg = dict(steps.DEFAULT_GAN_CFG)
        cfg['target_real'] = torch.ones(BATCH, 1, device=dev)
        cfg['target_fake'] = torch.zeros(BATCH, 1, device=dev)

    gen = torch.Generator().manual_seed(1234 + rank)
    n_host = 4
    host_lr = [(torch.rand(BATCH, 3, 48, 48, generator=gen) * 255).pin_memory() for _ in range(n_host)]
    host_hr = [(torch.rand(BATCH, 3, 192, 192, generator=gen) * 255).pin_memory() for _ in range(n_host)]
    dev_lr = [t.to(dev) for t in host_lr]
    dev_hr = [t.to(dev) for t in host_hr]

    def eager_step(lr, hr):
        if args.workload == "gan":
            return steps.gan_step(Gw, Dw, vgg, optim_G, optim_D, lr, hr, cfg, ddp_g=ddp_g, ddp_d=ddp_d)
        return steps.pretrain_step(Gw, optim_G, lr, hr, ddp=ddp_g).reshape(1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- the step as a CUDA graph (pesr_b200.graph): one cudaGraphLaunch per step instead of ~700 launches from Python
    warm = max(args.warmup, 3)
    for i in range(warm):
        eager_step(dev_lr[i % n_host], dev_hr[i % n_host])
    barrier()
    launches_per_step, graph_note, do_step = None, None, eager_step
    # N > 1 stays eager: capturing the step with ProcessGroupNCCL's all-reduce inside hung on the 2-GPU box (round 2)
    if not args.no_graph and world == 1:
        try:
            L.launch_count(reset=True)
            mods = [m for m in (G, D, vgg) if m is not None]
            opts = [o for o in (optim_G, optim_D) if o is not None]
            gstep = GraphedStep(eager_step, (dev_lr[0], dev_hr[0]), modules=mods, optimizers=opts, warmup=2)
            launches_per_step = L.launch_count() // 3          # 2 warm-up steps + the captured one went through the launchers
            do_step = gstep
        except Exception as e:       # the eager path stays valid; say what happened
            graph_note = f"CUDA-graph capture failed, eager launches used: {type(e).__name__}: {str(e)[:300]}"
            print("[bench] " + graph_note, file=sys.stderr, flush=True)
            torch.cuda.synchronize()
    graphed = do_step is not eager_step
    if launches_per_step is None:
        L.launch_count(reset=True)
        eager_step(dev_lr[0], dev_hr[0])
        launches_per_step = L.launch_count()

    def timed(fn, k):
        ms = _event_timed(fn, k, barrier)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- resident-input arm (value): EXACTLY args.steps timed steps after the warm-up
    for i in range(warm):
        do_step(dev_lr[i % n_host], dev_hr[i % n_host])
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    from pesr_b200 import parallel as _par
    _par.TRACE_EVENTS.clear()
    ms_total = timed(lambda i: do_step(dev_lr[i % n_host], dev_hr[i % n_host]), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    ddp_exposed = None
    if _par.TRACE_EVENTS:        # PESR_DDP_TRACE=1: time the compute stream spent waiting for NCCL, per step (rank 0)
        ddp_exposed = {}
        for lab, e0, e1 in _par.TRACE_EVENTS:
            ddp_exposed[lab] = ddp_exposed.get(lab, 0.0) + e0.elapsed_time(e1) / args.steps
        _par.TRACE_EVENTS.clear()
    # CPU time to ENQUEUE one step into an empty stream (no launch-queue back-pressure): GPU-bound if < ms_per_step
    host_enqueue_ms = None
    for i in range(3):
        barrier()
        h0 = time.perf_counter()
        do_step(dev_lr[i % n_host], dev_hr[i % n_host])
        dt = (time.perf_counter() - h0) * 1e3
        host_enqueue_ms = dt if host_enqueue_ms is None else min(host_enqueue_ms, dt)
    barrier()

    # ---- sustained arm: the same step for >= sustain_seconds, so that a comparison with the SUSTAINED cuBLAS rate
    # (measured over 4 s at whatever clock the 1 kW cap allows) is like for like
    sustained = None
    if args.sustain_seconds > 0:
        n_sus = max(args.steps, int(args.sustain_seconds * 1e3 / ms_step) + 1)
        s2 = ClockSampler(local)
        if rank == 0:
            s2.start()
        ms_sus = timed(lambda i: do_step(dev_lr[i % n_host], dev_hr[i % n_host]), n_sus)
        c2 = s2.stop() if rank == 0 else None
        sustained = {"steps": n_sus, "seconds": ms_sus * 1e-3, "ms_per_step": ms_sus / n_sus,
                     "value": BATCH * world / (ms_sus / n_sus * 1e-3), "clocks": c2}

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

    # ---- roofline leg: eager steps with per-launch CUDA events around the tensor-core kernels
    L.profile_enable(True)
    L.profile_read(0), L.profile_read(1)
    n_prof = min(args.steps, 10)
    for i in range(n_prof):
        eager_step(dev_lr[i % n_host], dev_hr[i % n_host])
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
    sus_tflops = step_gflop(args.workload) / sustained["ms_per_step"] if sustained else None
    roofline = {
        "bound": "tensor", "kernel": "conv_igemm_kernel (fprop + dgrad of every conv)", "achieved": achieved,
        "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"], "frac_of_burst": achieved / peaks["burst"],
        "peak_burst": peaks["burst"],
        "traffic": NCU_TRUNK_CONV_DRAM_BYTES,
        "traffic_note": "dram__bytes_read+write of ONE trunk-conv launch (16x48x48, 256->256) from profiles/r01_ncu_conv_igemm.txt; "
                        "algorithmic bytes of that launch: 18.9 MB in + 1.2 MB weights + 18.9 MB out (the output stays in the 126 MB L2)",
        "peak_source": peaks["source"], "launches_per_step": ig_n / n_prof,
        "avg_launch_us": ig_ms * 1e3 / max(ig_n, 1), "share_of_step": ig_ms / n_prof / ms_step,
        "wgrad": {"kernel": "conv_wgrad_kernel", "achieved": wg_achieved, "frac": wg_achieved / peaks["tflops"],
                  "frac_of_burst": wg_achieved / peaks["burst"], "launches_per_step": wg_n / n_prof,
                  "share_of_step": wg_ms / n_prof / ms_step},
        "whole_step": {"achieved": step_tflops, "frac": step_tflops / peaks["tflops"], "frac_of_burst": step_tflops / peaks["burst"],
                       "sustained_run_achieved": sus_tflops,
                       "sustained_run_frac": sus_tflops / peaks["tflops"] if sus_tflops else None,
                       "sustained_run_frac_of_burst": sus_tflops / peaks["burst"] if sus_tflops else None,
                       "algorithmic_gflop_per_step": step_gflop(args.workload)},
        "how": "algorithmic FLOPs (2*M*N*K, dense convention) of every launch / CUDA-event time of that launch, summed over "
               f"{n_prof} EAGER steps with per-launch events enabled (the events serialise what programmatic dependent "
               "launch overlaps, so the per-kernel rates are lower bounds); whole_step = algorithmic FLOPs of the step / "
               "ms_per_step of the timed region (and of the >= 3 s sustained run)",
    }
    if world == 1 and not args.no_extras:
        try:
            roofline["hbm"] = hbm_kernels(dev, peaks["hbm"])
        except Exception as e:
            roofline["hbm_error"] = f"{type(e).__name__}: {str(e)[:200]}"
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cb = 4
        t = cpu_step_time(args.workload, cb)
        cpu = {"value": cb / t, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"1 step of a {cb}-sample batch (of {BATCH}), fp32, torch CPU oracle, {t:.1f} s"}
    extras = {}
    if world == 1 and not args.no_extras:
        extras = side_measurements(args, dev, G, optim_G, dev_lr, dev_hr)
    cfg_out = config_dict(args.workload, world)
    cfg_out["launch"] = "cuda_graph" if graphed else "eager"
    line = {
        "metric": metric_name(args.workload), "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16", "data": "synthetic", "config": cfg_out,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e},
        "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
        "host_enqueue_ms_per_step": host_enqueue_ms, "sustained": sustained,
        "roofline": roofline, "cpu_baseline": cpu,
        "adam_table_builds": [o.table_builds for o in (optim_G, optim_D) if o is not None],
    }
    if graph_note:
        line["graph_note"] = graph_note
    if world > 1:
        line["ddp"] = {"nccl_max_ctas": os.environ.get("NCCL_MAX_CTAS"), "reserved_sms": os.environ.get("PESR_RESERVE_SMS", "4"),
                       "bucket_mb": os.environ.get("PESR_DDP_BUCKET_MB", "8"),
                       "fc1_factor_gather": os.environ.get("PESR_NO_FC1_GATHER") != "1",
                       "exposed_wait_ms_per_step": ddp_exposed}
    line.update(extras)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def hbm_kernels(dev, peak_gbs):
    """Achieved HBM GB/s of the memory-bound kernels of the step, each timed alone with CUDA events over rotating
    buffers larger than the 126 MB L2 (Adam, BatchNorm + LeakyReLU, the loss reductions, FC1, the
    x8 blend).  bytes = algorithmic bytes of ONE launch (read + written once)."""
    import torch
    from pesr_b200 import ops
    from pesr_b200._lib import check, lib
    f16, f32 = torch.float16, torch.float32
    out = []

    def run(name, nbytes, fn, rot, iters=12, note=None):
        for i in range(3):
            fn(i % rot)
        ms = _event_timed(lambda i: fn(i % rot), iters) / iters
        e = {"kernel": name, "bytes": int(nbytes), "us": ms * 1e3, "achieved": nbytes / (ms * 1e-3) / 1e9, "unit": "GB/s"}
        e["frac"] = e["achieved"] / peak_gbs
        if note:
            e["note"] = note
        out.append(e)
    # multi-tensor Adam over the Generator's 43.09 M parameters: 28 B per parameter
    n = 43_089_947
    p, g, m, v = (torch.randn(n, device=dev) * 1e-2 for _ in range(4))
    v.abs_()
    rows = [(p.data_ptr() + 4 * o, g.data_ptr() + 4 * o, m.data_ptr() + 4 * o, v.data_ptr() + 4 * o, min(1 << 16, n - o))
            for o in range(0, n, 1 << 16)]
    table = torch.tensor(rows, dtype=torch.int64).to(dev)
    run("adam_multi_kernel (G, 43.09 M params)", 28 * n, lambda i: ops.adam_multi(table, len(rows), 5e-5, 0.9, 0.999, 1e-8, 10), 1)
    del p, g, m, v, table
    # BatchNorm + LeakyReLU of the Discriminator's first block: 16 x 192 x 192 x 64 16-bit
    npix, c = BATCH * 192 * 192, 64
    ys = [torch.randn(npix, c, device=dev, dtype=f16) for _ in range(3)]
    a = [torch.empty(npix, c, device=dev, dtype=f16) for _ in range(3)]
    mean, rstd, gam, bet = (torch.zeros(c, device=dev), torch.ones(c, device=dev), torch.ones(c, device=dev), torch.zeros(c, device=dev))
    ws = torch.zeros(1024, device=dev, dtype=torch.float64)
    run("bn_reduce_vec_kernel (statistics), 37.7 M x fp16", 2 * npix * c, lambda i: ops.bn_reduce(ys[i], npix, c, ws), 3)
    run("bn_lrelu_fwd_kernel (finalise + apply), 37.7 M x fp16", 4 * npix * c,
        lambda i: ops.bn_lrelu_fwd(ys[i], npix, c, mean, rstd, gam, bet, a[i], sums_ws=ws), 3)
    dgam, dbet = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
    run("bn_lrelu_bwd (reduce + apply), 37.7 M x fp16", (2 * 2 + 3) * npix * c,
        lambda i: ops.bn_lrelu_bwd(a[i], ys[i], npix, c, mean, rstd, gam, ws, a[(i + 1) % 3], dgam, dbet), 3,
        note="two passes: statistics read dz and y, apply reads dz and y and writes dy")
    del ys, a
    # loss reductions over sr / hr (16 x 3 x 192 x 192 fp32 = 7.1 MB per tensor: L2-resident, latency-bound)
    e = BATCH * 3 * 192 * 192
    xs = [torch.rand(BATCH, 3, 192, 192, device=dev) * 255 for _ in range(24)]
    gr = torch.empty_like(xs[0])
    loss = torch.empty((), device=dev)
    small = "7.1 MB per tensor: at this size the kernel is launch / latency bound, not bandwidth bound"
    run("diff_loss_kernel L1 (value + gradient)", 12 * e, lambda i: ops.loss_l1(xs[2 * i], xs[2 * i + 1], loss, gr), 12, note=small)
    run("diff_loss_kernel MSE on VGG features 16x512x12x12", 12 * BATCH * 512 * 144,
        lambda i: ops.loss_mse(xs[2 * i].view(-1)[:BATCH * 512 * 144], xs[2 * i + 1].view(-1)[:BATCH * 512 * 144], loss,
                               gr.view(-1)[:BATCH * 512 * 144]), 12, note=small)
    run("tv_loss_kernel (value + gradient)", 8 * e, lambda i: ops.loss_tv(xs[i], loss, gr), 24, note=small)
    del xs, gr
    # FC1 of the Discriminator (75.5 M weights, 16 rows): forward on the split-K igemm, weight gradient on CUDA cores
    k, o = 73728, 1024
    w16 = [torch.randn(o, k, device=dev, dtype=f16) * 1e-2 for _ in range(2)]
    x16 = torch.randn(BATCH, k, device=dev, dtype=f16)
    ks = 36
    part = torch.empty(ks * BATCH * o, device=dev)
    descs = [ops.make_conv_desc(dtype=0, nb=1, h=1, w=BATCH, cin=k, cout=o, block_n=256, taps=[(0, 0)],
                                srcs=[ops.nhwc_src(x16, 1, 1, BATCH, k)], wpacked=w, out32=part, ld_out32=o, ksplit=ks,
                                split_stride32=BATCH * o) for w in w16]
    run("FC1 forward (conv_igemm split-K, 151 MB of fp16 weights)", 2 * k * o, lambda i: ops.conv_igemm(descs[i]), 2)
    dy = torch.randn(BATCH, o, device=dev)
    dws = [torch.empty(o, k, device=dev) for _ in range(2)]
    run("linear_wgrad_kernel FC1, CUDA cores (302 MB fp32 gradient written; round-1 path, PESR_FC1_WGRAD_SKINNY=1)", 4 * k * o,
        lambda i: ops.linear_wgrad(dy, x16, BATCH, k, o, dws[i]), 2)
    from pesr_b200.engine_d import DiscriminatorEngine
    eng_fc = DiscriminatorEngine(None)
    dy32 = torch.randn(2 * BATCH, o, device=dev) * 1e-4
    x32r = torch.randn(2 * BATCH, k, device=dev, dtype=f16)
    run("FC1 weight gradient on the tensor cores (conv_wgrad_kernel, 32 rows as hi + lo pairs; 302 MB written)", 4 * k * o,
        lambda i: eng_fc._fc1_wgrad(dy32, x32r, k, dws[i]), 2, note="includes the amax / split / operand copies (4 small launches)")
    del eng_fc
    del w16, dws, part
    # x8 self-ensemble blend + uint8 store at 1356 x 2040 (E6): 9 fp32 reads + 1 byte written per element
    H, W = 1356, 2040
    perc = torch.rand(3, H, W, device=dev) * 255
    ens = torch.rand(8, 3, H * W, device=dev) * 255
    o8 = torch.empty(H, W, 3, device=dev, dtype=torch.uint8)
    s = torch.cuda.current_stream().cuda_stream
    run("blend_x8_to_u8_kernel 1356x2040 (alpha = 0.5)", (9 * 4 + 1) * 3 * H * W,
        lambda i: check(lib.pesr_blend_x8_to_u8(perc.data_ptr(), ens.data_ptr(), H, W, 0.5, 8, 0, o8.data_ptr(), s)), 1,
        note="299 MB working set > L2")
    # 3-channel edges of the Generator at the training shape
    lr = torch.rand(BATCH, 3, 192, 192, device=dev) * 255
    cols = [torch.empty(BATCH * 192 * 192, 64, device=dev, dtype=f16) for _ in range(3)]
    run("im2col3_kernel 16x3x192x192 -> [P][64] fp16", 12 * BATCH * 192 * 192 + 128 * BATCH * 192 * 192,
        lambda i: ops.im2col3(lr, cols[i]), 3)
    zs = [torch.randn(BATCH * 192 * 192, 32, device=dev) for _ in range(3)]
    sr = torch.empty(BATCH, 3, 192, 192, device=dev)
    run("col2im3_tiled_kernel [P][32] fp32 -> 16x3x192x192", (128 + 12) * BATCH * 192 * 192, lambda i: ops.col2im3(zs[i], 32, BATCH, 192, 192, sr), 3)
    return out


def side_measurements(args, dev, G, optim_G, dev_lr, dev_hr):
    """The other single-GPU configurations of BASELINE.json, measured the same way (CUDA events, warm-up) and
    reported beside the headline: config 2 (L1 pretrain step), config 1 (128x128 inference) and config 5 (339x510
    inference: batch sweep 1-32 at alpha = 1, alpha = 0.5 with the x8 self-ensemble, uint8 host image -> uint8 host image)."""
    import torch
    from pesr_b200 import infer, steps
    from pesr_b200.model import Generator
    out = {}
    k = min(args.steps, 20)

    def timed(fn, n, warm=2):
        for i in range(warm):
            fn(i)
        return _event_timed(fn, n) / n
    if args.workload != "pretrain":
        ms = timed(lambda i: steps.pretrain_step(G, optim_G, dev_lr[i % len(dev_lr)], dev_hr[i % len(dev_hr)]), k, warm=3)
        out["pretrain_step"] = {"metric": "L1 pretrain samples/s", "value": BATCH / (ms * 1e-3), "ms_per_step": ms,
                                "tflops": step_gflop("pretrain") / ms, "launch": "eager", "config": "BASELINE.json configs[1]"}
    # split-precision Generator (fp16 hi+lo operands, three tensor-core passes per conv: fp32-grade gradients), same step
    try:
        from pesr_b200.optim import Adam
        torch.manual_seed(2)
        Gs = Generator(OPT, split_precision=True).to(dev)
        oS = Adam(Gs.parameters(), lr=5e-5)
        ms = timed(lambda i: steps.pretrain_step(Gs, oS, dev_lr[i % len(dev_lr)], dev_hr[i % len(dev_hr)]), 5, warm=2)
        out["split_precision_pretrain_step"] = {"metric": "L1 pretrain samples/s, split-precision Generator", "value": BATCH / (ms * 1e-3),
                                                "ms_per_step": ms, "tflops_algorithmic": step_gflop("pretrain") / ms,
                                                "tensor_passes_per_conv": 3, "launch": "eager",
                                                "note": "pesr_b200/engine_g_split.py; tests/test_split_precision_gpu.py"}
        if args.workload == "gan":
            # the whole GAN step with all three networks on the split-precision schedules (engine_d_split / engine_v_split)
            from pesr_b200.model import VGG, Discriminator
            Ds = Discriminator(OPT, split_precision=True).to(dev)
            Vs = VGG(pretrained=False, split_precision=True).to(dev)
            oD = Adam(Ds.parameters(), lr=5e-5)
            cfg_s = dict(steps.DEFAULT_GAN_CFG)
            cfg_s['target_real'] = torch.ones(BATCH, 1, device=dev)
            cfg_s['target_fake'] = torch.zeros(BATCH, 1, device=dev)
            ms = timed(lambda i: steps.gan_step(Gs, Ds, Vs, oS, oD, dev_lr[i % len(dev_lr)], dev_hr[i % len(dev_hr)], cfg_s), 4, warm=2)
            out["split_precision_gan_step"] = {"metric": "GAN train samples/s, split-precision G + D + VGG", "value": BATCH / (ms * 1e-3),
                                               "ms_per_step": ms, "tflops_algorithmic": step_gflop("gan") / ms,
                                               "tensor_passes_per_conv": 3, "launch": "eager",
                                               "note": "pesr_b200/engine_{g,d,v}_split.py; gradients vs the free-running fp64 oracle: "
                                                       "tests/test_split_precision_gpu.py"}
            del Ds, Vs, oD
        del Gs, oS
        torch.cuda.empty_cache()
    except Exception as e:
        out["split_precision_pretrain_step"] = out.get("split_precision_pretrain_step") or {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        out["split_precision_error"] = f"{type(e).__name__}: {str(e)[:200]}"
    G.eval()
    flop_px = 100505088
    inf = {}
    x = torch.rand(1, 3, 128, 128, device=dev) * 255
    ms = timed(lambda i: infer.super_resolve(G, x), 5)
    inf["128x128"] = {"ms_per_image": ms, "hr_mpix_per_s": 16 * 128 * 128 / (ms * 1e-3) / 1e6, "tflops": 128 * 128 * flop_px / (ms * 1e-3) / 1e12}
    h, w = 339, 510
    sweep = {}
    for b in (1, 2, 4, 8, 16, 32):
        xb = torch.rand(b, 3, h, w, device=dev) * 255
        ms = timed(lambda i: infer.super_resolve(G, xb), 3, warm=1)
        sweep[str(b)] = {"ms_per_batch": ms, "hr_mpix_per_s": b * 16 * h * w / (ms * 1e-3) / 1e6,
                         "tflops": b * h * w * flop_px / (ms * 1e-3) / 1e12}
        del xb
    inf["339x510_batch_sweep_alpha1"] = sweep
    # uint8 HWC host image -> uint8 HWC host image (the whole of test.py:103-114 minus PNG decode / encode)
    host_in = (torch.rand(h, w, 3) * 255).to(torch.uint8).pin_memory()
    host_out = torch.empty(4 * h, 4 * w, 3, dtype=torch.uint8).pin_memory()

    def e2e(i):
        o8 = infer.super_resolve_u8(G, host_in.to(dev, non_blocking=True))
        host_out.copy_(o8, non_blocking=True)
    ms = timed(e2e, 5)
    inf["339x510_u8_host_to_host"] = {"ms_per_image": ms, "hr_mpix_per_s": 16 * h * w / (ms * 1e-3) / 1e6,
                                      "h2d_bytes": host_in.numel(), "d2h_bytes": host_out.numel()}
    # alpha = 0.5: perceptual model + x8 self-ensemble of the PSNR model (9 forwards) + fused blend
    torch.manual_seed(1)
    Gp = Generator(OPT).to(dev).eval()
    x1 = torch.rand(1, 3, h, w, device=dev) * 255
    ms = timed(lambda i: infer.super_resolve(G, x1, alpha=0.5, model_psnr=Gp), 3, warm=1)
    inf["339x510_alpha0.5"] = {"ms_per_image": ms, "hr_mpix_per_s": 16 * h * w / (ms * 1e-3) / 1e6,
                               "tflops": 9 * h * w * flop_px / (ms * 1e-3) / 1e12, "forwards": 9}
    del Gp
    G.train()
    out["inference"] = {"metric": "x4 SR inference HR Mpix/s (fp32 image in, fp32 + uint8 image out unless noted)", **inf,
                        "config": "BASELINE.json configs[0] and configs[4]"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("PESR_BENCH_WORKLOAD", "gan"), choices=["pretrain", "gan"])
    ap.add_argument("--no-extras", action="store_true", help="skip the pretrain-step and inference side measurements")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step kernel by kernel instead of replaying a CUDA graph")
    ap.add_argument("--sustain-seconds", type=float, default=3.0, help="length of the extra sustained run (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
