"""Kernel timing probe (CUDA events, warm-up, L2 flush between iterations). Prints TFLOP/s per kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from pesr_b200 import ops

FLUSH = None



# bring-up hooks live in the debug build only: run with PESR_B200_LIB=pesr_b200/libpesr_b200_debug.so (tools/build_debug.sh)
from pesr_b200 import _debug as _dbg, _lib as _L   # noqa: E402
_dbg.bind()


def _set_pair_mode_compat(mode):
    """round-1 encoding of the option hook: 0/1/2 pair mode, 20x sub stages, 30x PDL, 40x staged epilogue, 50x specialised epilogue"""
    if mode >= 500: _L.set_option(_L.OPT_SPECIALISED_EPILOGUE, mode - 500)
    elif mode >= 400: _L.set_option(_L.OPT_STAGED_EPILOGUE, mode - 400)
    elif mode >= 300: _L.set_option(_L.OPT_PDL, mode - 300)
    elif mode >= 200: _L.set_option(_L.OPT_SUB_STAGES, mode - 200)
    elif mode >= 100: pass
    else: _L.set_option(_L.OPT_PAIR_MODE, mode)


def timeit(fn, iters=10, warm=3, flush=True):
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if flush:
            FLUSH.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters


def conv_case(nb, cin, cout, h, w, dtype=torch.float16, block_n=None, **epi):
    x16 = torch.randn(nb, h, w, cin, device="cuda").to(dtype)
    wp = (torch.randn(9 * cout, cin, device="cuda") / (3 * cin ** 0.5)).to(dtype)
    out16 = torch.empty(nb, h, w, cout, device="cuda", dtype=dtype)
    bias = torch.randn(cout, device="cuda")
    d = ops.make_conv_desc(dtype=ops.dt_code(dtype), nb=nb, h=h, w=w, cin=cin, cout=cout,
                           srcs=[ops.nhwc_src(x16, nb, h, w, cin)], wpacked=wp, bias=bias, act=1, out16=out16,
                           ld_out16=cout, block_n=block_n)
    ms = timeit(lambda: ops.conv_igemm(d))
    fl = 2.0 * nb * h * w * cout * cin * 9
    # cuDNN comparator (channels_last fp16)
    xc = torch.randn(nb, cin, h, w, device="cuda", dtype=dtype).contiguous(memory_format=torch.channels_last)
    wc = torch.randn(cout, cin, 3, 3, device="cuda", dtype=dtype).contiguous(memory_format=torch.channels_last)
    ms_c = timeit(lambda: F.conv2d(xc, wc, None, padding=1))
    print(f"fprop nb={nb} {cin}->{cout} {h}x{w}: {ms*1e3:.1f} us  {fl/ms/1e9:.0f} TFLOP/s | cudnn {ms_c*1e3:.1f} us "
          f"{fl/ms_c/1e9:.0f} TFLOP/s", flush=True)


def wgrad_case(nb, cin, cout, h, w, dtype=torch.float16):
    x16 = torch.randn(nb, h, w, cin, device="cuda").to(dtype)
    dy16 = torch.randn(nb, h, w, cout, device="cuda").to(dtype)
    part = torch.empty(16 * 9 * cout * cin, device="cuda", dtype=torch.float32)
    grad = torch.empty(cout, cin, 3, 3, device="cuda")
    d = ops.make_wgrad_desc(dtype=ops.dt_code(dtype), nb=nb, h=h, w=w, a=dy16, a_c=cout, m_total=cout,
                            b_srcs=[ops.nhwc_src(x16, nb, h, w, cin)], n_total=cin, partials=part)
    sp = [0]

    def run():
        sp[0] = ops.conv_wgrad(d)
    ms = timeit(run)
    ms_r = timeit(lambda: ops.wgrad_reduce(part, sp[0], 9, cout, cin, ops.WMAP_OIHW, cout, cin, grad))
    fl = 2.0 * nb * h * w * cout * cin * 9
    xc = torch.randn(nb, cin, h, w, device="cuda", dtype=dtype).contiguous(memory_format=torch.channels_last)
    dyc = torch.randn(nb, cout, h, w, device="cuda", dtype=dtype).contiguous(memory_format=torch.channels_last)
    ms_c = timeit(lambda: torch.nn.grad.conv2d_weight(xc, (cout, cin, 3, 3), dyc, padding=1))
    print(f"wgrad nb={nb} {cin}->{cout} {h}x{w}: {ms*1e3:.1f} us (splits {sp[0]}) {fl/ms/1e9:.0f} TFLOP/s, reduce "
          f"{ms_r*1e3:.1f} us | cudnn {ms_c*1e3:.1f} us {fl/ms_c/1e9:.0f} TFLOP/s", flush=True)


def gen_step(nb=16, dtype=torch.float16):
    from pesr_b200.model import Generator
    opt = {'depth': 32, 'num_channels': 256, 'res_scale': 0.1}
    torch.manual_seed(0)
    G = Generator(opt, dtype=dtype).cuda()
    lr = torch.rand(nb, 3, 48, 48, device="cuda") * 255
    hr = torch.rand(nb, 3, 192, 192, device="cuda") * 255

    def fwd():
        with torch.no_grad():
            G(lr)

    def step():
        for p in G.parameters():
            p.grad = None
        sr = G(lr)
        loss = (sr - hr).abs().mean()
        loss.backward()
    ms_f = timeit(fwd, iters=5, flush=False)
    ms_s = timeit(step, iters=5, flush=False)
    print(f"G fwd nb={nb}: {ms_f:.2f} ms = {nb*231.564/ms_f:.0f} TFLOP/s ; fwd+bwd: {ms_s:.2f} ms = "
          f"{nb*3*231.564/ms_s:.0f} TFLOP/s ({nb/ms_s*1e3:.0f} samples/s)", flush=True)
    import time
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"   host launch time per step {1e3*(t1-t0)/3:.2f} ms, total wall {1e3*(t2-t0)/3:.2f} ms")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what == "one":       # single trunk-conv shape for `ncu --set full` (tools/ncu_full.sh): light, residual, wgrad
        conv_case(16, 256, 256, 48, 48)
        nb, c, h, w = 16, 256, 48, 48
        x16 = torch.randn(nb, h, w, c, device="cuda").half()
        wp = (torch.randn(9 * c, c, device="cuda") / (3 * c ** 0.5)).half()
        out16 = torch.empty(nb, h, w, c, device="cuda", dtype=torch.float16)
        s32 = torch.randn(nb, h, w, c, device="cuda")
        d = ops.make_conv_desc(dtype=0, nb=nb, h=h, w=w, cin=c, cout=c, srcs=[ops.nhwc_src(x16, nb, h, w, c)], wpacked=wp,
                               bias=torch.randn(c, device="cuda"), alpha=0.1, res32=s32, ld_res32=c, out32=s32, ld_out32=c,
                               out16=out16, ld_out16=c)
        print(f"residual-stream conv: {timeit(lambda: ops.conv_igemm(d))*1e3:.1f} us")
        wgrad_case(16, 256, 256, 48, 48)
    if what == "mma":
        from pesr_b200._lib import lib
        out = torch.zeros(2, dtype=torch.int64, device="cuda")
        for pair in (0, 1, 2, 4, 6, 7):     # bit0 pair, bit1 A MN-major, bit2 B MN-major
            for n in (256, 128):
                for blocks in (148,):
                    for stages in (4,):
                        iters = 400
                        out.zero_()
                        rc = lib.pesr_debug_mma_rate(n, iters, stages, pair, blocks, out.data_ptr(), 0)
                        if rc:
                            print("  rc", rc, lib.pesr_last_error())
                        torch.cuda.synchronize()
                        a, b = out.cpu().tolist()
                        print(f"mma rate pair={pair} N={n} blocks={blocks} stages={stages}: issue {a/(iters*4):.1f} cyc/MMA, "
                              f"complete {b/(iters*4):.1f} cyc/MMA (ideal {128*n/256:.0f})", flush=True)
    if what == "wtimeline":
        from pesr_b200._lib import lib
        buf = torch.zeros(64 + 2 * 160, dtype=torch.int64, device="cuda")
        lib.pesr_debug_wgrad_timeline(buf.data_ptr())
        for pairmode in (0, 1):
            lib.pesr_debug_wgrad_desc(-1, pairmode)
            wgrad_case(16, 256, 256, 48, 48)
            torch.cuda.synchronize()
            b = buf.cpu().tolist()
            mhz = (b[61] - b[0]) / max(b[63] - b[62], 1) * 1e3
            print(f"  wgrad pair={pairmode} block0 (SM clock {mhz:.0f} MHz, total {(b[63]-b[62])/1e3:.1f} us): first stage landed "
                  f"{b[9]-b[0]}, 8 k-blocks later {b[11]-b[0]}, MMAs issued {b[10]-b[0]}, acc complete {b[16]-b[0]}, "
                  f"epilogue done {b[17]-b[0]}, end {b[61]-b[0]}")
            st = [b[64 + 2 * i] for i in range(148) if b[64 + 2 * i]]
            en = [b[65 + 2 * i] for i in range(148) if b[65 + 2 * i]]
            t0 = min(st)
            durs = sorted((e - s_) / 1e3 for s_, e in zip(st, en))
            print(f"  per-CTA: starts spread {(max(st)-t0)/1e3:.1f} us, first end {(min(en)-t0)/1e3:.1f} us, last end {(max(en)-t0)/1e3:.1f} us, "
                  f"CTA duration min/median/max {durs[0]:.1f}/{durs[len(durs)//2]:.1f}/{durs[-1]:.1f} us")
        lib.pesr_debug_wgrad_timeline(0)
        buf2 = torch.zeros(64 + 2 * 160, dtype=torch.int64, device="cuda")
        lib.pesr_debug_timeline(buf2.data_ptr())
        for pm in (0, 2):
            _set_pair_mode_compat(pm)
            conv_case(16, 256, 256, 48, 48)
            torch.cuda.synchronize()
            b = buf2.cpu().tolist()
            st = [b[64 + 2 * i] for i in range(148) if b[64 + 2 * i]]
            en = [b[65 + 2 * i] for i in range(148) if b[65 + 2 * i]]
            t0 = min(st)
            durs = sorted((e - s_) / 1e3 for s_, e in zip(st, en))
            print(f"  fprop pair={pm} per-CTA: starts spread {(max(st)-t0)/1e3:.1f} us, first end {(min(en)-t0)/1e3:.1f}, last end {(max(en)-t0)/1e3:.1f} us, "
                  f"CTA duration min/median/max {durs[0]:.1f}/{durs[len(durs)//2]:.1f}/{durs[-1]:.1f} us")
        lib.pesr_debug_timeline(0)
        _set_pair_mode_compat(1)
    if what == "epi":      # chunk-level timeline of the LAST tile's epilogue of block 0 (trunk conv, light and heavy epilogue)
        from pesr_b200._lib import lib
        buf = torch.zeros(64 + 2 * 160, dtype=torch.int64, device="cuda")
        lib.pesr_debug_timeline(buf.data_ptr())
        nb, c, h, w = 16, 256, 48, 48
        x16 = torch.randn(nb, h, w, c, device="cuda").half()
        wp = (torch.randn(9 * c, c, device="cuda") / (3 * c ** 0.5)).half()
        out16 = torch.empty(nb, h, w, c, device="cuda", dtype=torch.float16)
        res32 = torch.randn(nb, h, w, c, device="cuda")
        out32 = torch.empty(nb, h, w, c, device="cuda")
        bias = torch.randn(c, device="cuda")
        variants = {
            "out16 + bias + relu": dict(bias=bias, act=1, out16=out16, ld_out16=c),
            "res32 -> out32 + out16": dict(bias=bias, alpha=0.1, res32=res32, ld_res32=c, out32=out32, ld_out32=c, out16=out16, ld_out16=c),
            "mask16 -> out16": dict(mask16=x16, ld_mask16=c, mask_mode=1, out16=out16, ld_out16=c),
        }
        for label, light, staged in (("specialised+staged", 501, 401), ("specialised direct", 501, 400), ("generic", 500, 400)):
            _set_pair_mode_compat(light)
            _set_pair_mode_compat(staged)
            for name, epi in variants.items():
                d = ops.make_conv_desc(dtype=0, nb=nb, h=h, w=w, cin=c, cout=c, srcs=[ops.nhwc_src(x16, nb, h, w, c)], wpacked=wp, **epi)
                for flush in (True, False):
                    ms = timeit(lambda: ops.conv_igemm(d), flush=flush)
                    torch.cuda.synchronize()
                    b = buf.cpu().tolist()
                    t0 = b[0]
                    print(f"{label:18s} {name:24s} flush={int(flush)}: {ms*1e3:.1f} us | tile0: acc complete {b[16]-t0}, epi done {b[17]-t0} | "
                          f"last tile: acc complete {b[20]-t0}, chunks done at "
                          + " ".join(f"+{b[25+i]-b[20]}" for i in range(8)) + f" | end {b[61]-t0}", flush=True)
        lib.pesr_debug_timeline(0)
        _set_pair_mode_compat(401)
        _set_pair_mode_compat(501)
    if what == "timeline":
        from pesr_b200._lib import lib
        buf = torch.zeros(64, dtype=torch.int64, device="cuda")
        lib.pesr_debug_timeline(buf.data_ptr())
        names = {1: "first TMA issued", 2: "TMAs tile0 issued", 3: "TMAs last tile issued", 8: "t0 acc free", 9: "t0 first stage landed",
                 10: "t0 MMAs issued", 12: "tL acc free", 13: "tL first stage landed", 14: "tL MMAs issued", 16: "t0 acc complete",
                 17: "t0 epilogue done", 20: "tL acc complete", 21: "tL epilogue done", 61: "end"}
        for pair in (0, 1):
            _set_pair_mode_compat(pair)
            for shape in ((16, 256, 256, 48, 48), (16, 256, 1024, 96, 96), (16, 64, 64, 192, 192), (16, 512, 512, 24, 24)):
                conv_case(*shape)
                torch.cuda.synchronize()
                b = buf.cpu().tolist()
                t0 = b[0]
                mhz = (b[61] - b[0]) / max(b[63] - b[62], 1) * 1e3
                print(f"  pair={pair} block-0 timeline (cycles from start, SM clock {mhz:.0f} MHz, total {(b[63]-b[62])/1e3:.1f} us):")
                print("   " + "; ".join(f"{names[i]}={b[i]-t0}" for i in sorted(names) if b[i]))
        lib.pesr_debug_timeline(0)
    if what == "bn128":
        for bn in (256, 128):
            print("block_n", bn)
            conv_case(16, 512, 512, 24, 24, block_n=bn)
            conv_case(32, 512, 512, 24, 24, block_n=bn)
            conv_case(32, 256, 512, 24, 24, block_n=bn)
            conv_case(32, 512, 512, 12, 12, block_n=bn)
            conv_case(16, 512, 512, 12, 12, block_n=bn)
            conv_case(16, 128, 256, 48, 48, block_n=bn)
            conv_case(16, 256, 256, 24, 24, block_n=bn)
            conv_case(16, 256, 256, 48, 48, block_n=bn)
            conv_case(32, 256, 256, 48, 48, block_n=bn)
    if what == "sub":
        from pesr_b200._lib import lib
        _set_pair_mode_compat(0)
        for mode in (201, 202):
            _set_pair_mode_compat(mode)
            print("sub-block stages", mode - 200, "(pair off)")
            conv_case(16, 512, 512, 24, 24)
            conv_case(32, 512, 512, 24, 24)
            conv_case(32, 256, 512, 24, 24)
            conv_case(32, 512, 512, 12, 12)
            conv_case(16, 128, 256, 48, 48)
            conv_case(16, 256, 256, 48, 48)
        _set_pair_mode_compat(1)
        _set_pair_mode_compat(201)
    if what == "pair":
        from pesr_b200._lib import lib
        for mode in (0, 2):
            _set_pair_mode_compat(mode)
            print("pair mode", mode)
            conv_case(16, 256, 256, 48, 48)
            conv_case(16, 256, 1024, 96, 96)
            conv_case(16, 1024, 256, 96, 96)
            conv_case(16, 64, 64, 192, 192)
            conv_case(16, 128, 128, 96, 96)
            conv_case(16, 512, 512, 24, 24)
            conv_case(32, 128, 128, 96, 96)
            conv_case(32, 256, 256, 48, 48)
            conv_case(32, 512, 512, 12, 12)
    if what in ("all", "conv"):
        conv_case(16, 256, 256, 48, 48)
        conv_case(16, 256, 1024, 48, 48)
        conv_case(16, 256, 1024, 96, 96)
        conv_case(16, 1024, 256, 96, 96)
        conv_case(16, 64, 64, 192, 192)
        conv_case(16, 512, 512, 24, 24)
        wgrad_case(16, 256, 256, 48, 48)
        wgrad_case(16, 256, 1024, 96, 96)
    if what in ("all", "gen"):
        gen_step()
