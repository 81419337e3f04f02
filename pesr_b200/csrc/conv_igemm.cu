// Implicit-GEMM convolution for sm_100a: forward and backward-data of every 3x3 / 1x1 conv on the
// PESR hot path (reference: model/basic.py:4-7 `Conv` -> nn.Conv2d -> cuDNN fprop/dgrad).
//
//   D[pixel, cout] = sum_tap sum_cin  X[pixel (+) tap, cin] * Wp[tap][cout][cin]
//
// One persistent CTA per SM, 6 warps:
//   warp 0      TMA producer: per k-block one 4-D activation box (64 ch x tile_w x tile_h x 1 image; conv
//               padding and ragged edges come from TMA's out-of-bounds zero fill) + one 2-D weight box.
//   warp 1      tcgen05.mma issuer (single thread), fp32 accumulators in TMEM, two accumulator buffers
//               so the epilogue of tile i overlaps the MMAs of tile i+1.
//   warps 2..5  epilogue: tcgen05.ld -> bias / scale / residual / activation / ReLU-mask -> global
//               (NHWC 16-bit, optional fp32 copy, optional PixelShuffle(2) / inverse addressing).
#include <type_traits>

#include "common.cuh"
#include "host_util.cuh"

#include <stdlib.h>

namespace pesr {

static constexpr int kTileM = 128;        // pixels per tile == TMEM lanes
static constexpr int kKBlock = 64;        // K elements per pipeline stage (one 128B swizzle row)
static constexpr int kABytes = kTileM * kKBlock * 2;  // 16 KB
static constexpr int kAccStride = 256;    // TMEM columns reserved per accumulator buffer
static constexpr int kNumThreads = 192;
static constexpr int kMaxStages = 8;
#ifdef PESR_DEBUG_HOOKS
static constexpr bool kDebugHooks = true;    // timeline stamps compiled in (libpesr_b200_debug.so only)
#else
static constexpr bool kDebugHooks = false;
#endif

struct ConvMaps {
  CUtensorMap a[PESR_MAX_SRC];
  CUtensorMap b;
};

// Residual epilogue through shared memory (kEpi 4): per epilogue warp a ring of kEpiDepth 4 KB fp32 tiles (32 pixels x
// 32 channels; the residual of the next chunks lands there by cp.async while the current one is processed) and one
// 2 KB 16-bit transposition tile.
// BatchNorm-statistics epilogue (kEpi 7): per epilogue warp one 4 KB fp32 transposition tile, then 2 x 512 floats of
// per-CTA channel sums / sums of squares.
constexpr int kBnEpiBytes = 4 * 4096 + 2 * 512 * 4;
constexpr int kEpiDepth = 3;
constexpr int kEpiWarpBytes = kEpiDepth * 4096 + 2048;
constexpr int kEpiBytes = 4 * kEpiWarpBytes;

struct ConvK {
  int dtype;
  int nb, h, w, cin, cout, block_n, tile_h, tile_w, ntaps;
  int tile_n;   // images per pixel tile (rows of a tile are ordered image, y, x)
  int tiles_h, tiles_w, n_tiles, num_tiles, kblocks_per_tap;
  int stages, stage_bytes;
  int8_t tap_dh[PESR_MAX_TAPS], tap_dw[PESR_MAX_TAPS], tap_src[PESR_MAX_TAPS], tap_widx[PESR_MAX_TAPS];
  const float* bias;
  float alpha;
  const float* alpha_dev;
  const float* res32; int ld_res32;
  const uint16_t* res16; int ld_res16;
  int act;
  const uint16_t* mask16; int ld_mask16; int mask_mode;
  float* out32; int ld_out32;
  uint16_t* out16; int ld_out16;
  int out_mode, out_h, out_w, out_sy, out_sx, out_oy, out_ox, out_coff, ps_c, aux_mode;
  int ksplit, b_mn_major, mn_tiles, pdl_early_b;
  double* bn_sums;   // kEpi 7: [2][cout] global accumulators
  // several K sub-blocks per pipeline stage, staged by ONE activation box + ONE weight box (TMA op count bounds the
  // small-N layers): sub_mode 1 = nsub consecutive 64-channel chunks of one tap, 2 = the three vertical taps of one
  // kernel column out of a (tile_h + 2)-row halo box
  int sub_mode, nsub, a_sub_off, b_sub_off, a_bytes, steps_per_tile;
  // sub_mode 2 with RESIDENT weights (wres_bytes > 0): the layer's whole packed weight tensor (<= ~150 KB: the N = 64 /
  // 128 layers of VGG and the Discriminator) is loaded into shared memory once per CTA, in the order the halo stages
  // consume it, and every pipeline stage carries the activation box only.  Those layers re-fetched 74 KB of weights for
  // every 128-pixel tile - more L2 -> SM traffic per tile than the activations, at 115 B per MMA cycle - and ran at a
  // third of the tensor rate.
  int wres_bytes;
  unsigned long long* dbg;   // optional timeline of block 0 (bring-up): [64] clock64 stamps + [63] = globaltimer ns
  long long split_stride32;
  // several sub-problems ("classes") of one pixel grid in ONE launch: class c owns taps [cls_tap0[c], cls_tap0[c+1]) and
  // the output-grid offset (cls_oy[c], cls_ox[c]); tile index = class * tiles_per_cls + tile-within-class.  The four
  // parity classes of a stride-2 backward-data run this way (each alone fills half the machine or less).
  int ncls, tiles_per_cls;
  int cls_tap0[5];
  int cls_oy[4], cls_ox[4];
};

// kPair = false: one CTA per tile (cta_group::1).
// kPair = true : a cluster of two CTAs drives one 256-pixel x block_n MMA (cta_group::2).  Each CTA stages its own
//   128 pixels of A and HALF of the weight tile, so shared-memory fill and operand-read traffic per SM drop by a
//   third; the leader CTA's thread issues every MMA, commits are multicast to both CTAs' barriers, every TMA load
//   of the pair signals the leader's full barrier, and both epilogues report to the leader's tmem_empty barrier.
// kEpi fixes the epilogue's feature set at compile time.  The generic chunk loop (kEpi 0: every feature a run-time
// branch) is ~14 KB of code and the kernel then runs instruction-fetch-bound: the same work takes 1000 cycles per
// 32-column chunk in the generic kernel and 570 in a specialised one, and even the MMA / TMA warps speed up.
//   0 generic (pixel-shuffle addressing, strided output grids, split-K partials, 16-bit residual, ...)
//   1 light    : bias + activation -> 16-bit NHWC output
//   2 mask     : (bias) * relu'/lrelu' mask of a saved 16-bit activation -> 16-bit output      (backward-data)
//   3 residual : alpha * (acc + bias) + fp32 residual -> fp32 output + 16-bit copy              (residual stream)
//   4          : 3 with coalesced global accesses, transposed through shared memory (CTA-pair kernel only)
//   5, 6       : 1 with the PixelShuffle(2) / un-shuffle output addressing (upsampler forward / backward-data)
//   7          : 1 plus per-channel sum / sum of squares of the rounded outputs (train-mode BatchNorm statistics)
template <bool kPair, int kEpi>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_igemm_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ ConvK p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr bool kStaged = kEpi == 4;
  constexpr bool kCtRes32 = kEpi == 3 || kEpi == 4, kCtMask = kEpi == 2;
  const bool has_res32 = kEpi ? kCtRes32 : p.res32 != nullptr;
  const bool has_out32 = kEpi ? kCtRes32 : p.out32 != nullptr;
  const bool has_mask = kEpi ? kCtMask : p.mask16 != nullptr;
  const bool has_res16 = kEpi ? false : p.res16 != nullptr;
  const bool has_out16 = kEpi ? true : p.out16 != nullptr;

  // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps the role branches and the
  // single-issuer loops on the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int rank = kPair ? (int)cluster_ctarank() : 0;
  const bool leader = rank == 0;
  const int worker = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // tile-loop index of this CTA / pair
  const int nworkers = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int b_rows = kPair ? p.block_n / 2 : p.block_n;                      // weight rows staged by this CTA

  uint8_t* epi_base = smem + (size_t)p.stages * p.stage_bytes;            // staging tiles of the staged epilogue
  uint8_t* wres = epi_base;                                               // resident weights (never with a staged epilogue)
  uint8_t* tail = epi_base + (kStaged ? kEpiBytes : kEpi == 7 ? kBnEpiBytes : 0) + p.wres_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full = empty_bar + kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* wres_bar = reinterpret_cast<uint64_t*>(tail + 192);
  float* bias_s = reinterpret_cast<float*>(tail + 256);  // [2][256]

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; i++) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(&maps.b);
    for (int s = 0; s < p.stages; s++) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], kPair ? 8 : 128);   // pair: one arrival per epilogue warp of both CTAs
    }
    mbar_init(wres_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (kPair) { tmem_alloc2(tmem_ptr, 512); tmem_relinquish2(); }
    else       { tmem_alloc(tmem_ptr, 512);  tmem_relinquish(); }
  }
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // k-blocks of one tile of class c (ncls == 1: the whole tap list)
  auto class_kb = [&](int cls) { return (p.cls_tap0[cls + 1] - p.cls_tap0[cls]) * p.kblocks_per_tap; };
  unsigned long long* dbg = (kDebugHooks && p.dbg && blockIdx.x == 0) ? p.dbg : nullptr;
  unsigned long long t_begin = 0, g_begin = 0;
  if (dbg && threadIdx.x == 0) {
    t_begin = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_begin));
    dbg[0] = t_begin;
    dbg[62] = g_begin;
  }
  if (kDebugHooks && p.dbg && threadIdx.x == 0) {   // per-CTA start / end wall-clock stamps (all blocks)
    unsigned long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    p.dbg[64 + 2 * blockIdx.x] = g;
  }

  // tile -> (class, split, m_tile of THIS CTA, n_tile)
  auto decode = [&](int tile, int& split, int& m_tile, int& n_tile, int& cls) {
    cls = 0;
    if (p.ncls > 1) { cls = tile / p.tiles_per_cls; tile -= cls * p.tiles_per_cls; }
    split = tile / p.mn_tiles;
    const int mn = tile - split * p.mn_tiles;
    n_tile = mn % p.n_tiles;
    m_tile = kPair ? 2 * (mn / p.n_tiles) + rank : mn / p.n_tiles;
  };

  if (warp == 0) {
    // ===================== TMA producer (whole warp, warp-uniform; one elected lane issues) =====================
    // The loop below is on the critical path of the pipeline fill (one warp, dependent-issue latency), so it is
    // specialised per staging mode and walks (tap, channel block) incrementally instead of dividing per step.
    //   MODE 0: one 64-channel k-block per stage (also the CTA-pair path)     3: same with MN-major weights
    //   MODE 1: nsub consecutive channel blocks of one tap per stage          2: halo stage (3 vertical taps)
    auto produce = [&](auto mode_tag) {
      constexpr int MODE = decltype(mode_tag)::value;
      int stage = 0;
      uint32_t phase = 0;
      const int kpt = p.kblocks_per_tap;
      // position inside the K loop: MODE 0/1/3: (t, cb); MODE 2: (cb, dx) with t := dx
      // t is an index into the tap arrays (absolute: the class's first tap is added)
      auto seek = [&](int s, int cls, int& t, int& cb) {
        if (MODE == 2) { cb = s / 3; t = s - cb * 3; }
        else { const int kb = MODE == 1 ? s * p.nsub : s; t = kb / kpt; cb = kb - t * kpt; t += p.cls_tap0[cls]; }
      };
      auto advance = [&](int& t, int& cb) {
        if (MODE == 2) { if (++t == 3) { t = 0; cb++; } }
        else { cb += MODE == 1 ? p.nsub : 1; if (cb >= kpt) { cb = 0; t++; } }
      };
      auto expect = [&](uint64_t* fb) {
        if (kPair) { if (leader) mbar_expect_tx_w(fb, 2u * (uint32_t)p.stage_bytes); }
        else mbar_expect_tx_w(fb, (uint32_t)p.stage_bytes);
      };
      auto load_b = [&](uint8_t* sa, uint64_t* fb, int s, int t, int cb, int n0) {
        if (MODE == 2) {
          tma_load_3d_w(sa + p.a_bytes, &maps.b, fb, cb * kKBlock, t * p.cout + n0, 0);
        } else if (MODE == 1) {
          tma_load_3d_w(sa + p.a_bytes, &maps.b, fb, 0, p.tap_widx[t] * p.cout + n0, cb);
        } else if (MODE == 3) {
          // weights stored [K][N] (N contiguous): one 64(K) x 64(N) box per 64 output columns
          for (int i = 0; i < p.block_n / 64; i++) tma_load_2d_w(sa + kABytes + i * 8192, &maps.b, fb, n0 + i * 64, s * kKBlock);
        } else if (kPair) {
          tma2_load_2d_w(sa + kABytes, &maps.b, fb, cb * kKBlock, p.tap_widx[t] * p.cout + n0);
        } else {
          tma_load_2d_w(sa + kABytes, &maps.b, fb, cb * kKBlock, p.tap_widx[t] * p.cout + n0);
        }
      };
      auto step_range = [&](int split, int cls, int& s0, int& s1) {
        if (MODE == 2) { s0 = 0; s1 = p.steps_per_tile; return; }
        const int total_kb = class_kb(cls);
        if (MODE == 1) { s0 = 0; s1 = total_kb / p.nsub; return; }
        s0 = (int)(((long long)split * total_kb) / p.ksplit);
        s1 = (int)(((long long)(split + 1) * total_kb) / p.ksplit);
      };
      // PDL prologue: the weights do not depend on the preceding kernel, so the first ring of weight boxes is in
      // flight while that kernel drains
      int npre = 0;
      const bool resident = MODE == 2 && p.wres_bytes > 0;
      auto load_resident = [&]() {     // every (channel block, dx) weight box of the layer, in stage order
        mbar_expect_tx_w(wres_bar, (uint32_t)p.wres_bytes);
        for (int s = 0; s < p.steps_per_tile; s++) {
          const int cb = s / 3, t = s - cb * 3;
          tma_load_3d_w(wres + (size_t)s * 3 * p.b_sub_off, &maps.b, wres_bar, cb * kKBlock, t * p.cout, 0);
        }
      };
      if (resident && p.pdl_early_b) load_resident();
      if (p.pdl_early_b && !resident) {
        int split, m_tile, n_tile, cls, s0, s1, t, cb;
        decode(worker, split, m_tile, n_tile, cls);
        step_range(split, cls, s0, s1);
        seek(s0, cls, t, cb);
        npre = s1 - s0 < p.stages ? s1 - s0 : p.stages;
        const int n0 = n_tile * p.block_n + rank * b_rows;
        for (int i = 0; i < npre; i++) {
          expect(&full_bar[i]);
          load_b(smem + (size_t)i * p.stage_bytes, &full_bar[i], s0 + i, t, cb, n0);
          advance(t, cb);
        }
      }
      griddep_wait();
      griddep_launch();
      if (resident && !p.pdl_early_b) load_resident();
      for (int tile = worker; tile < p.num_tiles; tile += nworkers) {
        int split, m_tile, n_tile, cls, s0, s1, t, cb;
        decode(tile, split, m_tile, n_tile, cls);
        step_range(split, cls, s0, s1);
        seek(s0, cls, t, cb);
        const int tw = m_tile % p.tiles_w;
        const int th = (m_tile / p.tiles_w) % p.tiles_h;
        // first image of the tile; >= nb for the padding tile of an odd pair: all OOB -> zeros
        const int img = (m_tile / (p.tiles_w * p.tiles_h)) * p.tile_n;
        const int h0 = th * p.tile_h, w0 = tw * p.tile_w, n0 = n_tile * p.block_n + rank * b_rows;
        int s = s0;
        while (s < s1) {
          // everything that depends on the tap only is computed once per tap (MODE 2: once per channel block)
          const int inner_step = MODE == 1 ? p.nsub : 1;
          const int inner_end = MODE == 2 ? 3 : kpt;
          int& inner = MODE == 2 ? t : cb;
          const int tap = MODE == 2 ? 0 : t;
          const CUtensorMap* ma = &maps.a[MODE == 2 ? 0 : p.tap_src[tap]];
          const int hh = MODE == 2 ? h0 - 1 : h0 + p.tap_dh[tap];
          const int ww = MODE == 2 ? w0 - 1 : w0 + p.tap_dw[tap];
          const int brow = MODE == 2 ? n0 : p.tap_widx[tap] * p.cout + n0;
          for (; inner < inner_end && s < s1; inner += inner_step, s++) {
            mbar_wait_w(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + (size_t)stage * p.stage_bytes;
            uint64_t* fb = &full_bar[stage];
            const bool with_b = npre == 0 && !resident;     // else: the weights of this step were issued in the prologue
            if (with_b || resident) expect(fb); else npre--;
            if (MODE == 2) {
              tma_load_4d_w(sa, ma, fb, cb * kKBlock, ww + t, hh, img);
              if (with_b) tma_load_3d_w(sa + p.a_bytes, &maps.b, fb, cb * kKBlock, t * p.cout + brow, 0);
            } else if (MODE == 1) {
              tma_load_5d_w(sa, ma, fb, 0, ww, hh, img, cb);
              if (with_b) tma_load_3d_w(sa + p.a_bytes, &maps.b, fb, 0, brow, cb);
            } else if (kPair) {
              tma2_load_4d_w(sa, ma, fb, cb * kKBlock, ww, hh, img);
              if (with_b) tma2_load_2d_w(sa + kABytes, &maps.b, fb, cb * kKBlock, brow);
            } else {
              tma_load_4d_w(sa, ma, fb, cb * kKBlock, ww, hh, img);
              if (with_b) load_b(sa, fb, s, t, cb, n0);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
            if (dbg && lane == 0 && tile == worker && s == s0) dbg[1] = clock64();          // first TMA issued
          }
          if (inner >= inner_end) {
            inner = 0;
            if (MODE == 2) cb++; else t++;
          }
        }
        if (dbg && lane == 0) dbg[2 + (tile == worker ? 0 : 1)] = clock64();    // all TMAs of tile 0 / last tile issued
      }
    };
    if (kPair || (p.sub_mode == 0 && !p.b_mn_major)) produce(std::integral_constant<int, 0>{});
    else if (p.sub_mode == 1) produce(std::integral_constant<int, 1>{});
    else if (p.sub_mode == 2) produce(std::integral_constant<int, 2>{});
    else produce(std::integral_constant<int, 3>{});
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only in pair mode) =====================
    if (leader) {
      const uint32_t idesc = make_idesc(kPair ? 256 : kTileM, p.block_n, p.dtype, 0, p.b_mn_major);
      // the accumulator address through a shuffle: the compiler then knows it is warp-uniform and keeps it in a uniform
      // register instead of converting it (R2UR under the elect predicate) in front of every tcgen05.mma
      const uint32_t tmem_base_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = worker; tile < p.num_tiles; tile += nworkers) {
        int cls = 0, tloc = tile;
        if (p.ncls > 1) { cls = tile / p.tiles_per_cls; tloc = tile - cls * p.tiles_per_cls; }
        const int split = tloc / p.mn_tiles;
        const int total_kb = class_kb(cls);
        const int kb0 = (int)(((long long)split * total_kb) / p.ksplit);
        const int kb1 = (int)(((long long)(split + 1) * total_kb) / p.ksplit);
        const int steps_per_tile = p.sub_mode == 2 ? p.steps_per_tile : total_kb / p.nsub;
        if (kPair) mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1); else mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base_u + (uint32_t)(acc * kAccStride);
        if (dbg && lane == 0) dbg[8 + (tile == worker ? 0 : 4)] = clock64();    // MMA warp: accumulator free
        // UMMA shared-memory descriptors: everything but the 14-bit address field (bits 0..13, in 16-byte units) is the same
        // for every K-major operand tile, and the tiles are 1024-byte aligned, so the descriptor of sub-block j, k-slice k
        // is desc_k | (base >> 4) + j * sub + 2 * k: one add / one or per operand instead of a shift-mask-or chain per
        // instruction.  The narrow layers (N = 64: 32 tensor cycles per instruction) are bound by this loop's issue rate.
        const uint64_t desc_k = make_smem_desc(0, 16, 1024);
        if (!kPair && p.sub_mode) {
          const uint32_t a_sub16 = (uint32_t)p.a_sub_off >> 4, b_sub16 = (uint32_t)p.b_sub_off >> 4;
          const bool resident = p.wres_bytes > 0;
          if (resident && tile == worker) {       // the layer's weights, loaded once per CTA
            mbar_wait_w(wres_bar, 0);
            tc_fence_after();
          }
          const uint32_t wres16 = smem_u32(wres) >> 4;
          uint32_t accum = 0;
          for (int st = 0; st < steps_per_tile; st++) {
            mbar_wait_w(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + (size_t)stage * p.stage_bytes);
            uint32_t a16 = a_addr >> 4;
            uint32_t b16 = resident ? wres16 + (uint32_t)st * 3u * b_sub16 : (a_addr + (uint32_t)p.a_bytes) >> 4;
            for (int j = 0; j < p.nsub; j++) {
#pragma unroll
              for (int k = 0; k < kKBlock / 16; k++) {
                umma_f16_w(d_tmem, desc_k | (uint64_t)(a16 + 2 * k), desc_k | (uint64_t)(b16 + 2 * k), idesc, accum);
                accum = 1;
              }
              a16 += a_sub16;
              b16 += b_sub16;
            }
            umma_commit_w(&empty_bar[stage]);
            if (st == steps_per_tile - 1) umma_commit_w(&tmem_full[acc]);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
          continue;
        }
        for (int kb = kb0; kb < kb1; kb++) {
          mbar_wait_w(&full_bar[stage], phase);
          tc_fence_after();
          if (dbg && lane == 0 && kb == kb0) dbg[9 + (tile == worker ? 0 : 4)] = clock64();  // first stage landed
          {
            const uint32_t a_addr = smem_u32(smem + (size_t)stage * p.stage_bytes);
            const uint32_t b_addr = a_addr + kABytes;
            const uint32_t a16 = a_addr >> 4, b16 = b_addr >> 4;
#pragma unroll
            for (int k = 0; k < kKBlock / 16; k++) {
              const uint64_t da = desc_k | (uint64_t)(a16 + 2 * k);
              const uint64_t db = p.b_mn_major ? make_smem_desc(b_addr + k * 2048, 8192, 1024) : (desc_k | (uint64_t)(b16 + 2 * k));
              if (kPair) umma2_f16_w(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              else       umma_f16_w(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            if (kPair) {
              umma2_commit_both_w(&empty_bar[stage]);
              if (kb == kb1 - 1) umma2_commit_both_w(&tmem_full[acc]);
            } else {
              umma_commit_w(&empty_bar[stage]);
              if (kb == kb1 - 1) umma_commit_w(&tmem_full[acc]);
            }
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (dbg && lane == 0) dbg[10 + (tile == worker ? 0 : 4)] = clock64();   // all MMAs of the tile issued
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    if (kStaged) {
      // ---- residual-stream epilogue through shared memory (kEpi 4).  With one thread per pixel row every global access
      // of a warp touches 32 different lines; the fp32 residual read + fp32 write + 16-bit write then cost 320
      // 32-byte wavefronts per 32-column chunk (measured 2.4-2.9 K cycles per chunk, longer than the tile's MMAs).
      // Here global memory is only touched with full-line 128-bit accesses (8 lanes per fp32 pixel row, 4 per 16-bit
      // row) and the row <-> lane transposition goes through an XOR-swizzled tile private to the warp: 96 wavefronts.
      griddep_wait();
      const int q = warp & 3;                       // TMEM lane quarter == rows [32q, 32q+32) of the tile
      const int et = threadIdx.x - 64;
      const int bf = p.dtype;
      const uint32_t ring32 = smem_u32(epi_base + q * kEpiWarpBytes);       // [kEpiDepth][32 rows][128 B]
      const uint32_t t16 = ring32 + kEpiDepth * 4096;                       // [32 rows][ 64 B]
      const int rows_h = 32 / p.tile_w;             // spatial rows of the tile owned by this warp
      float alpha = p.alpha;
      if (p.alpha_dev) alpha *= __ldg(p.alpha_dev);
      // own row (lane) inside the staging tiles, and the XOR swizzles that make both access patterns conflict-free
      const uint32_t own32 = (uint32_t)lane * 128u, own16 = (uint32_t)lane * 64u;
      const uint32_t x32 = (uint32_t)(lane & 7), x16 = (uint32_t)((lane >> 1) & 3);
      // coalesced assignment: fp32 tile: 8 lanes per row (16 B each), rows i*4 + lane/8; 16-bit tile: 4 lanes per row
      const int c32 = lane & 7, r32 = lane >> 3, c16 = lane & 3, r16 = lane >> 2;
      int acc = 0;
      uint32_t acc_phase = 0;
      uint32_t slot = 0;                            // ring slot of the chunk being processed
      for (int tile = worker; tile < p.num_tiles; tile += nworkers) {
        int split, m_tile, n_tile, cls;
        decode(tile, split, m_tile, n_tile, cls);
        const int tw = m_tile % p.tiles_w;
        const int th = (m_tile / p.tiles_w) % p.tiles_h;
        const int img = m_tile / (p.tiles_w * p.tiles_h);
        const int h0 = th * p.tile_h + q * rows_h, w0 = tw * p.tile_w, n0 = n_tile * p.block_n;
        const int nch = p.block_n / 32;
        // pixel index of the rows this lane moves in the coalesced pattern; -1 = outside the image
        int pix32[8], pix16[4];
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int rr = i * 4 + r32, hh = h0 + rr / p.tile_w, ww = w0 + rr % p.tile_w;
          pix32[i] = (hh < p.h && ww < p.w && img < p.nb) ? (img * p.h + hh) * p.w + ww : -1;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int rr = i * 8 + r16, hh = h0 + rr / p.tile_w, ww = w0 + rr % p.tile_w;
          pix16[i] = (hh < p.h && ww < p.w && img < p.nb) ? (img * p.h + hh) * p.w + ww : -1;
        }
        const float* res_base = p.res32 + n0 + c32 * 4;
        float* o32_base = p.out32 + n0 + c32 * 4;
        uint16_t* o16_base = p.out16 + p.out_coff + n0 + c16 * 8;
        // residual of chunk ch -> ring slot s, asynchronously (cp.async, zero-filled outside the image); ONE group per
        // call, so "all but the newest kEpiDepth - 1 groups complete" always means "the chunk being processed landed"
        auto fetch = [&](int ch, uint32_t s) {
          if (ch < nch) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
              const uint32_t rr = (uint32_t)(i * 4 + r32);
              const uint32_t dst = ring32 + s * 4096u + rr * 128u + (((uint32_t)c32 ^ (rr & 7u)) << 4);
              const bool ok = pix32[i] >= 0;
              cp_async16(dst, ok ? res_base + (long long)pix32[i] * p.ld_res32 + ch * 32 : p.res32, ok ? 16u : 0u);
            }
          }
          cp_async_commit();
        };
        {
          uint32_t s = slot;
          for (int c = 0; c < kEpiDepth; c++) {       // independent of the accumulator: in flight during the wait
            fetch(c, s);
            if (++s == kEpiDepth) s = 0;
          }
        }

        float* bs = bias_s + acc * 256;
        for (int i = et; i < p.block_n; i += 128) bs[i] = p.bias ? alpha * __ldg(p.bias + n0 + i) : 0.f;   // alpha * bias
        asm volatile("bar.sync 1, 128;" ::: "memory");

        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        if (dbg && et == 0) dbg[16 + (tile == worker ? 0 : 4)] = clock64();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kAccStride);
        uint32_t v[32], vn[32];
        tmem_ld32(taddr, v);
        tmem_ld_wait();
        for (int ch = 0; ch < nch; ch++) {
          const bool more = ch + 1 < nch;
          const uint32_t t32 = ring32 + slot * 4096u;
          if (more) tmem_ld32(taddr + (ch + 1) * 32, vn);
          cp_async_wait<kEpiDepth - 1>();
          __syncwarp();
          // own row: alpha * acc + alpha * bias + residual
          float f[32];
#pragma unroll
          for (int c = 0; c < 8; c++) {
            const uint4 bb = lds128(smem_u32(bs) + (uint32_t)(ch * 128 + c * 16));
            const uint4 r = lds128(t32 + own32 + (((uint32_t)c ^ x32) << 4));
            f[4 * c] = fmaf(alpha, __uint_as_float(v[4 * c]), __uint_as_float(bb.x)) + __uint_as_float(r.x);
            f[4 * c + 1] = fmaf(alpha, __uint_as_float(v[4 * c + 1]), __uint_as_float(bb.y)) + __uint_as_float(r.y);
            f[4 * c + 2] = fmaf(alpha, __uint_as_float(v[4 * c + 2]), __uint_as_float(bb.z)) + __uint_as_float(r.z);
            f[4 * c + 3] = fmaf(alpha, __uint_as_float(v[4 * c + 3]), __uint_as_float(bb.w)) + __uint_as_float(r.w);
          }
          if (p.act == PESR_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; j++) f[j] = fmaxf(f[j], 0.f);
          } else if (p.act == PESR_ACT_LRELU) {
#pragma unroll
            for (int j = 0; j < 32; j++) f[j] = f[j] > 0.f ? f[j] : 0.2f * f[j];
          }
          // results back into the tiles (own row: in place over the residual)
#pragma unroll
          for (int c = 0; c < 8; c++)
            sts128(t32 + own32 + (((uint32_t)c ^ x32) << 4), __float_as_uint(f[4 * c]), __float_as_uint(f[4 * c + 1]),
                   __float_as_uint(f[4 * c + 2]), __float_as_uint(f[4 * c + 3]));
#pragma unroll
          for (int c = 0; c < 4; c++)
            sts128(t16 + own16 + (((uint32_t)c ^ x16) << 4), pack2(f[8 * c], f[8 * c + 1], bf),
                   pack2(f[8 * c + 2], f[8 * c + 3], bf), pack2(f[8 * c + 4], f[8 * c + 5], bf),
                   pack2(f[8 * c + 6], f[8 * c + 7], bf));
          __syncwarp();
          // coalesced stores: every instruction writes whole 128-byte (fp32) / 64-byte (16-bit) row segments
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const uint32_t rr = (uint32_t)(i * 4 + r32);
            const uint4 o = lds128(t32 + rr * 128u + (((uint32_t)c32 ^ (rr & 7u)) << 4));
            if (pix32[i] >= 0) stg128(o32_base + (long long)pix32[i] * p.ld_out32 + ch * 32, o);
          }
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const uint32_t rr = (uint32_t)(i * 8 + r16);
            const uint4 o = lds128(t16 + rr * 64u + (((uint32_t)c16 ^ ((rr >> 1) & 3u)) << 4));
            if (pix16[i] >= 0) stg128(o16_base + (long long)pix16[i] * p.ld_out16 + ch * 32, o);
          }
          __syncwarp();     // slot and 16-bit tile are free again
          fetch(ch + kEpiDepth, slot);                // refill the slot with the residual kEpiDepth chunks ahead
          if (++slot == kEpiDepth) slot = 0;
          if (more) {
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] = vn[j];
          } else if (tile + 2 * nworkers < p.num_tiles) {
            // hand the accumulator back - only if this CTA will use it again: the cluster-scope release of the arrive
            // waits for the epilogue's outstanding global stores (~1.5 K cycles on the exposed last tile)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
          }
          if (dbg && et == 0 && tile != worker && ch < 8) dbg[25 + ch] = clock64();
        }
        if (dbg && et == 0) dbg[17 + (tile == worker ? 0 : 4)] = clock64();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      cp_async_wait<0>();
    } else {
    // (two warps per TMEM lane quarter were tried: measured 5-10% SLOWER, the extra warps compete with the issue warps)
    griddep_wait();   // bias / alpha / residual / mask may be written by the preceding kernel (PDL)
    const float* const e_res32 = has_res32 ? p.res32 : nullptr;
    const uint16_t* const e_res16 = has_res16 ? p.res16 : nullptr;
    const uint16_t* const e_mask16 = has_mask ? p.mask16 : nullptr;
    float* const e_out32 = has_out32 ? p.out32 : nullptr;
    uint16_t* const e_out16 = has_out16 ? p.out16 : nullptr;
    const int e_out_mode = kEpi == 5 ? (int)PESR_OUT_SHUFFLE2 : kEpi == 6 ? (int)PESR_OUT_UNSHUFFLE2
                         : kEpi ? (int)PESR_OUT_NORMAL : p.out_mode;
    const int e_aux = p.aux_mode;     // only changes the per-tile pixel index: stays a run-time value
    // kEpi 7: channel statistics of this CTA, accumulated in shared memory over all its tiles
    float* cta_sum = reinterpret_cast<float*>(epi_base + 4 * 4096);
    float* cta_sq = cta_sum + 512;
    if (kEpi == 7) {
      for (int i = threadIdx.x - 64; i < 1024; i += 128) cta_sum[i] = 0.f;    // made visible by the first bar.sync 1
    }
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const int et = threadIdx.x - 64;  // 0..127
    const int ch_lo = 0, ch_hi = p.block_n / 32;
    const int px_img = p.tile_h * p.tile_w;            // pixels of one image inside the tile
    const int tn_l = row / px_img, ty = (row % px_img) / p.tile_w, tx = row % p.tile_w;
    const int bf = p.dtype;
    float alpha = p.alpha;
    if (p.alpha_dev) alpha *= __ldg(p.alpha_dev);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = worker; tile < p.num_tiles; tile += nworkers) {
      int split, m_tile, n_tile, cls;
      decode(tile, split, m_tile, n_tile, cls);
      const int out_oy = p.cls_oy[cls], out_ox = p.cls_ox[cls];      // output-grid offset of this tile's class
      const int tw = m_tile % p.tiles_w;
      const int th = (m_tile / p.tiles_w) % p.tiles_h;
      const int img = (m_tile / (p.tiles_w * p.tiles_h)) * p.tile_n + tn_l;
      const int h = th * p.tile_h + ty, w = tw * p.tile_w + tx, n0 = n_tile * p.block_n;
      const bool valid = (h < p.h) && (w < p.w) && (img < p.nb);
      // pixel index used by res32 / res16 / mask16 / out32: the GEMM grid, or (aux_mode 1) the strided output grid
      const long long pix = e_aux
          ? ((long long)img * p.out_h + (h * p.out_sy + out_oy)) * p.out_w + (w * p.out_sx + out_ox)
          : ((long long)img * p.h + h) * p.w + w;

      float* bs = bias_s + acc * 256;
      for (int i = et; i < p.block_n; i += 128) bs[i] = p.bias ? alpha * __ldg(p.bias + n0 + i) : 0.f;   // alpha * bias
      asm volatile("bar.sync 1, 128;" ::: "memory");

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (dbg && et == 0) dbg[16 + (tile == worker ? 0 : 4)] = clock64();       // accumulator complete (MMAs done)
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kAccStride);

      const bool acc_reused = tile + 2 * nworkers < p.num_tiles;      // see the staged epilogue: skip the release otherwise
      if (ch_lo == ch_hi && acc_reused) {   // nothing to drain (block_n == 32, upper half): still release the accumulator
        tc_fence_before();
        if (kPair) {
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
        } else {
          mbar_arrive(&tmem_empty[acc]);
        }
      }
      // software pipeline over 32-column chunks: the TMEM load and the fp32-residual loads of chunk ch+1 are in flight
      // while chunk ch is processed (one exposed latency per tile instead of one per chunk)
      uint32_t v[32], vn[32];
      f32x8 rn[4];
      const float* res_row = (has_res32 && valid) ? e_res32 + pix * p.ld_res32 + n0 : nullptr;
      u32x8 mn[2];
      const uint16_t* mask_row = (has_mask && valid) ? e_mask16 + pix * p.ld_mask16 + n0 : nullptr;
      tmem_ld32(taddr + ch_lo * 32, v);
      if (res_row) {
#pragma unroll
        for (int j = 0; j < 4; j++) rn[j] = ld256_f32(res_row + ch_lo * 32 + 8 * j);
      }
      if (mask_row) {
#pragma unroll
        for (int j = 0; j < 2; j++) mn[j] = ld256_b32(mask_row + ch_lo * 32 + 16 * j);
      }
      tmem_ld_wait();
      if (dbg && et == 0 && tile != worker) dbg[24] = clock64();                // last tile: first chunk in registers
      for (int ch = ch_lo; ch < ch_hi; ch++) {
        f32x8 rc[4];
#pragma unroll
        for (int j = 0; j < 4; j++) rc[j] = rn[j];
        u32x8 mc[2];
#pragma unroll
        for (int j = 0; j < 2; j++) mc[j] = mn[j];
        const bool more = ch + 1 < ch_hi;
        if (more) {
          tmem_ld32(taddr + (ch + 1) * 32, vn);
          if (res_row) {
#pragma unroll
            for (int j = 0; j < 4; j++) rn[j] = ld256_f32(res_row + (ch + 1) * 32 + 8 * j);
          }
          if (mask_row) {
#pragma unroll
            for (int j = 0; j < 2; j++) mn[j] = ld256_b32(mask_row + (ch + 1) * 32 + 16 * j);
          }
        }
        const bool stamp = dbg && et == 0 && tile != worker && ch == 2;
        if (stamp) dbg[40] = clock64();     // prefetch of chunk 3 issued
        float yq[kEpi == 7 ? 32 : 1];      // kEpi 7: the rounded outputs of this row (zeros outside the image)
        if (kEpi == 7) {
#pragma unroll
          for (int j = 0; j < 32; j++) yq[kEpi == 7 ? j : 0] = 0.f;
        }
        if (valid) {
          const int q0 = n0 + ch * 32;
          float f[32];
#pragma unroll
          for (int c = 0; c < 8; c++) {
            // explicit ld.shared: through the generic pointer the compiler emits 32 scalar generic loads here
            const uint4 bb = lds128(smem_u32(bs) + (uint32_t)(ch * 128 + c * 16));
            f[4 * c] = fmaf(alpha, __uint_as_float(v[4 * c]), __uint_as_float(bb.x));
            f[4 * c + 1] = fmaf(alpha, __uint_as_float(v[4 * c + 1]), __uint_as_float(bb.y));
            f[4 * c + 2] = fmaf(alpha, __uint_as_float(v[4 * c + 2]), __uint_as_float(bb.z));
            f[4 * c + 3] = fmaf(alpha, __uint_as_float(v[4 * c + 3]), __uint_as_float(bb.w));
          }
          if (has_res32) {
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
              for (int e = 0; e < 8; e++) f[8 * j + e] += rc[j].v[e];
          }
          if (has_res16) {
#pragma unroll
            for (int j = 0; j < 2; j++) {
              const u32x8 t = ld256_b32(e_res16 + pix * p.ld_res16 + q0 + 16 * j);
#pragma unroll
              for (int e = 0; e < 8; e++) {
                const float2 a = unpack2(t.v[e], bf);
                f[16 * j + 2 * e] += a.x; f[16 * j + 2 * e + 1] += a.y;
              }
            }
          }
          if (p.act == PESR_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; j++) f[j] = fmaxf(f[j], 0.f);
          } else if (p.act == PESR_ACT_LRELU) {
#pragma unroll
            for (int j = 0; j < 32; j++) f[j] = f[j] > 0.f ? f[j] : 0.2f * f[j];
          }
          if (has_mask) {
            const float neg = p.mask_mode == 2 ? 0.2f : 0.f;
#pragma unroll
            for (int j = 0; j < 2; j++) {
#pragma unroll
              for (int e = 0; e < 8; e++) {
                const float2 a = unpack2(mc[j].v[e], bf);
                f[16 * j + 2 * e] *= a.x > 0.f ? 1.f : neg; f[16 * j + 2 * e + 1] *= a.y > 0.f ? 1.f : neg;
              }
            }
          }
          if (stamp) dbg[41] = clock64();   // values computed
          if (has_out32) {
            float* o = e_out32 + split * p.split_stride32 + pix * p.ld_out32 + q0;
#pragma unroll
            for (int j = 0; j < 4; j++) st256_f32(o + 8 * j, f + 8 * j);
          }
          if (has_out16) {
            long long off;
            if (e_out_mode == PESR_OUT_SHUFFLE2) {
              const int ij = q0 / p.ps_c, c0 = q0 % p.ps_c;
              const long long op = ((long long)img * (2 * p.h) + (2 * h + (ij >> 1))) * (2 * p.w) + (2 * w + (ij & 1));
              off = op * p.ld_out16 + p.out_coff + c0;
            } else if (e_out_mode == PESR_OUT_UNSHUFFLE2) {
              const long long op = ((long long)img * (p.h >> 1) + (h >> 1)) * (p.w >> 1) + (w >> 1);
              off = op * p.ld_out16 + p.out_coff + ((h & 1) * 2 + (w & 1)) * p.cout + q0;
            } else {
              const long long op =
                  ((long long)img * p.out_h + (h * p.out_sy + out_oy)) * p.out_w + (w * p.out_sx + out_ox);
              off = op * p.ld_out16 + p.out_coff + q0;
            }
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; j++) pk[j] = pack2(f[2 * j], f[2 * j + 1], bf);
            st256_b32(e_out16 + off, pk);
            st256_b32(e_out16 + off + 16, pk + 8);
            if (kEpi == 7) {
#pragma unroll
              for (int j = 0; j < 16; j++) {
                const float2 a = unpack2(pk[j], bf);
                yq[kEpi == 7 ? 2 * j : 0] = a.x;
                yq[kEpi == 7 ? 2 * j + 1 : 0] = a.y;
              }
            }
          }
        }
        if (kEpi == 7) {
          // column sums over the warp's 32 rows: rows -> swizzled smem tile, then lane L reads column L of every row
          const uint32_t tb = smem_u32(epi_base) + (uint32_t)quarter * 4096u;
          const uint32_t xr = (uint32_t)(lane & 7);
#pragma unroll
          for (int c = 0; c < 8; c++)
            sts128(tb + (uint32_t)lane * 128u + (((uint32_t)c ^ xr) << 4), __float_as_uint(yq[kEpi == 7 ? 4 * c : 0]),
                   __float_as_uint(yq[kEpi == 7 ? 4 * c + 1 : 0]), __float_as_uint(yq[kEpi == 7 ? 4 * c + 2 : 0]),
                   __float_as_uint(yq[kEpi == 7 ? 4 * c + 3 : 0]));
          __syncwarp();
          float s1 = 0.f, s2 = 0.f;
          const uint32_t cpiece = (uint32_t)(lane >> 2), cword = (uint32_t)(lane & 3) * 4u;
#pragma unroll
          for (int r = 0; r < 32; r++) {
            const float t = __uint_as_float(lds32(tb + (uint32_t)r * 128u + ((cpiece ^ (uint32_t)(r & 7)) << 4) + cword));
            s1 += t;
            s2 = fmaf(t, t, s2);
          }
          __syncwarp();
          atomicAdd(cta_sum + n0 + ch * 32 + lane, s1);
          atomicAdd(cta_sq + n0 + ch * 32 + lane, s2);
        }
        if (stamp) dbg[42] = clock64();     // stores issued
        if (more) {
          tmem_ld_wait();
          if (stamp) dbg[43] = clock64();   // next chunk's TMEM load complete
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = vn[j];
        } else if (acc_reused) {
          // every accumulator column of this row has been read: hand the TMEM buffer back to the MMA warp
          tc_fence_before();
          if (kPair) {
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
          } else {
            mbar_arrive(&tmem_empty[acc]);
          }
        }
        if (dbg && et == 0 && tile != worker && ch < 8) dbg[25 + ch] = clock64();   // last tile: chunk ch done
      }
      if (dbg && et == 0) dbg[17 + (tile == worker ? 0 : 4)] = clock64();       // epilogue of the tile done
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (kEpi == 7) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int i = et; i < p.cout; i += 128) {
        atomicAdd(p.bn_sums + i, (double)cta_sum[i]);
        atomicAdd(p.bn_sums + p.cout + i, (double)cta_sq[i]);
      }
    }
    }
  }

  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();
  if (dbg && threadIdx.x == 0) {
    unsigned long long g_end;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_end));
    dbg[61] = clock64();
    dbg[63] = g_end;
  }
  if (kDebugHooks && p.dbg && threadIdx.x == 0) {
    unsigned long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    p.dbg[65 + 2 * blockIdx.x] = g;
  }
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

static int g_sub_mode_enabled = 1;
static int g_staged_enabled = 1;
static int g_light_enabled = 1;
static int g_wres_enabled = -1;   // resident weights for the narrow 3x3 layers: -1 from the environment (PESR_NO_WRES=1 disables)
static unsigned long long* g_dbg_buf = nullptr;
static int g_pair_mode = -1;  // -1: from the environment (PESR_NO_PAIR=1 disables), 0: never, 1: whenever legal

}  // namespace pesr

using namespace pesr;

// Kernel-selection options (include/pesr_b200.h PESR_OPT_*).
extern "C" int pesr_set_option(int option, int value) {
  switch (option) {
    case PESR_OPT_PAIR_MODE: g_pair_mode = value; return 0;
    case PESR_OPT_SUB_STAGES: g_sub_mode_enabled = value; return 0;
    case PESR_OPT_PDL: set_pdl(value); return 0;
    case PESR_OPT_STAGED_EPILOGUE: g_staged_enabled = value; return 0;
    case PESR_OPT_SPECIALISED_EPILOGUE: g_light_enabled = value; return 0;
    case PESR_OPT_RESERVE_SMS: set_reserved_sms(value); return 0;
    case PESR_OPT_RESIDENT_WEIGHTS: g_wres_enabled = value < 0 ? 0 : value > 2 ? 2 : value; return 0;
  }
  set_error("pesr_set_option: unknown option %d", option);
  return PESR_E_ARG;
}

#ifdef PESR_DEBUG_HOOKS
// Bring-up hook: device buffer of 64 uint64 that block 0 of every pesr_conv_igemm launch fills with a timeline.
extern "C" void pesr_debug_timeline(void* buf) { g_dbg_buf = reinterpret_cast<unsigned long long*>(buf); }
#endif

extern "C" int pesr_conv_igemm(const pesr_conv_desc* d_in, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(d_in != nullptr, "conv_igemm: null descriptor");
  pesr_conv_desc dd = *d_in;
  const pesr_conv_desc* d = &dd;
  // Tile-width heuristic (measured, tools/perf_conv.py bn128): 256-wide layers with too few pixel tiles for the CTA-pair
  // kernel run faster as 128-wide tiles with several K sub-blocks per stage (twice the tiles, a third of the TMA ops).
  if (g_pair_mode < 0) {
    const char* e = getenv("PESR_NO_PAIR");
    g_pair_mode = (e && e[0] == '1') ? 0 : 1;
  }
  if (g_wres_enabled < 0) {
    const char* e = getenv("PESR_NO_WRES");
    g_wres_enabled = (e && e[0] == '1') ? 0 : 1;
  }
  if (dd.block_n == 256 && dd.tile_h > 0 && dd.tile_w > 0 && g_sub_mode_enabled && dd.ksplit <= 1 && !dd.b_mn_major) {
    const int tn0 = dd.tile_n > 0 ? dd.tile_n : 1;
    const int mt = ((dd.nb + tn0 - 1) / tn0) * ((dd.h + dd.tile_h - 1) / dd.tile_h) * ((dd.w + dd.tile_w - 1) / dd.tile_w);
    const bool will_pair = dd.ncls <= 1 && (g_pair_mode == 2 || (g_pair_mode == 1 && mt >= 256));
    if (!will_pair && (dd.ntaps == 9 || (dd.cin / kKBlock) % 2 == 0)) dd.block_n = 128;
  }
  PESR_CHECK_ARG(d->dtype == PESR_DT_F16 || d->dtype == PESR_DT_BF16, "conv_igemm: bad dtype %d", d->dtype);
  PESR_CHECK_ARG(d->nb > 0 && d->h > 0 && d->w > 0, "conv_igemm: empty pixel grid %dx%dx%d", d->nb, d->h, d->w);
  PESR_CHECK_ARG(d->cin > 0 && d->cin % kKBlock == 0, "conv_igemm: cin %d must be a multiple of 64", d->cin);
  PESR_CHECK_ARG(d->block_n == 32 || d->block_n == 64 || d->block_n == 128 || d->block_n == 256,
                 "conv_igemm: block_n %d not in {32,64,128,256}", d->block_n);
  PESR_CHECK_ARG(d->cout > 0 && d->cout % d->block_n == 0, "conv_igemm: cout %d not a multiple of block_n %d",
                 d->cout, d->block_n);
  const int tile_n = d->tile_n > 0 ? d->tile_n : 1;
  PESR_CHECK_ARG(d->tile_h > 0 && d->tile_w > 0 && tile_n * d->tile_h * d->tile_w == kTileM,
                 "conv_igemm: tile %dx%dx%d must cover 128 pixels", tile_n, d->tile_h, d->tile_w);
  PESR_CHECK_ARG(d->tile_w <= 256 && d->tile_h <= 256, "conv_igemm: tile too large for a TMA box");
  PESR_CHECK_ARG(d->ntaps >= 1 && d->ntaps <= PESR_MAX_TAPS, "conv_igemm: ntaps %d", d->ntaps);
  PESR_CHECK_ARG(d->nsrc >= 1 && d->nsrc <= PESR_MAX_SRC, "conv_igemm: nsrc %d", d->nsrc);
  PESR_CHECK_ARG(d->wpacked != nullptr && d->w_rows > 0, "conv_igemm: missing packed weights");
  PESR_CHECK_ARG(d->out16 != nullptr || d->out32 != nullptr, "conv_igemm: no output");
  // the epilogue moves 32-byte vectors: 32 channels of a pixel must start on a 32-byte boundary in every tensor
  PESR_CHECK_ARG((!d->out16 || (((uintptr_t)d->out16 % 32) == 0 && d->ld_out16 % 16 == 0 && d->out_coff % 16 == 0)) &&
                     (!d->out32 || (((uintptr_t)d->out32 % 32) == 0 && d->ld_out32 % 8 == 0 && d->split_stride32 % 8 == 0)) &&
                     (!d->res32 || (((uintptr_t)d->res32 % 32) == 0 && d->ld_res32 % 8 == 0)) &&
                     (!d->res16 || (((uintptr_t)d->res16 % 32) == 0 && d->ld_res16 % 16 == 0)) &&
                     (!d->mask16 || (((uintptr_t)d->mask16 % 32) == 0 && d->ld_mask16 % 16 == 0)),
                 "conv_igemm: epilogue tensors must be 32-byte aligned with channel strides that are multiples of 32 bytes");
  for (int t = 0; t < d->ntaps; t++) {
    PESR_CHECK_ARG(d->tap_src[t] >= 0 && d->tap_src[t] < d->nsrc, "conv_igemm: tap %d reads source %d", t,
                   d->tap_src[t]);
    PESR_CHECK_ARG(d->b_mn_major || (d->tap_widx[t] + 1) * d->cout <= d->w_rows,
                   "conv_igemm: tap %d weight rows out of range", t);
  }
  const int ncls = d->ncls > 1 ? d->ncls : 1;
  if (ncls > 1) {
    int tsum = 0;
    for (int c = 0; c < ncls; c++) tsum += d->cls_ntaps[c];
    PESR_CHECK_ARG(ncls <= 4 && tsum == d->ntaps && d->ksplit <= 1 && !d->b_mn_major && d->out_mode == PESR_OUT_NORMAL &&
                       !d->bn_sums, "conv_igemm: a multi-class launch needs <= 4 classes whose tap counts sum to ntaps, "
                                    "whole-K tiles, K-major weights and NORMAL output addressing");
    for (int c = 0; c < ncls; c++) PESR_CHECK_ARG(d->cls_ntaps[c] >= 1, "conv_igemm: class %d has no taps", c);
  }
  if (d->out_mode == PESR_OUT_SHUFFLE2)
    PESR_CHECK_ARG(d->ps_c > 0 && d->ps_c % 32 == 0 && d->cout == 4 * d->ps_c, "conv_igemm: bad ps_c %d", d->ps_c);
  if (d->out_mode == PESR_OUT_UNSHUFFLE2)
    PESR_CHECK_ARG(d->h % 2 == 0 && d->w % 2 == 0, "conv_igemm: unshuffle needs even h, w");

  // CTA-pair (cta_group::2) path: whole-K tiles of a K-major weight matrix, at least two pixel tiles
  if (g_pair_mode < 0) {
    const char* e = getenv("PESR_NO_PAIR");
    g_pair_mode = (e && e[0] == '1') ? 0 : 1;
  }
  if (g_wres_enabled < 0) {
    const char* e = getenv("PESR_NO_WRES");
    g_wres_enabled = (e && e[0] == '1') ? 0 : 1;
  }
  const int img_groups = (d->nb + tile_n - 1) / tile_n;
  const int m_tiles_host = img_groups * ((d->h + d->tile_h - 1) / d->tile_h) * ((d->w + d->tile_w - 1) / d->tile_w);
  // measured on B200 (tools/perf_conv.py pair): the pair kernel wins 4-6% on the 256-wide, many-tile layers (G trunk and
  // upsampler) and loses 8-16% on the small D / VGG layers, so it is used where it wins (mode 2 forces it for tests).
  const bool pair_legal = d->ksplit <= 1 && !d->b_mn_major && d->block_n >= 64 && m_tiles_host >= 2 && ncls == 1;
  const bool pair = pair_legal && (g_pair_mode == 2 || (g_pair_mode == 1 && d->block_n == 256 && m_tiles_host >= 256));

  // staged epilogue (pair kernel): plain NHWC outputs on the GEMM's own pixel grid
  // epilogue specialisation (kEpi of the kernel template)
  const bool gemm_grid = !d->aux_mode && (d->out_h <= 0 || d->out_h == d->h) && (d->out_w <= 0 || d->out_w == d->w) &&
                         d->out_sy <= 1 && d->out_sx <= 1 && d->out_oy == 0 && d->out_ox == 0;
  const bool plain = d->ksplit <= 1 && !d->res16 && d->out16;
  int epi = 0;
  if (g_light_enabled && plain) {
    const bool light = !d->res32 && !d->mask16 && !d->out32;
    if (d->out_mode == PESR_OUT_NORMAL) {
      if (light) epi = 1;
      else if (d->mask16 && !d->res32 && !d->out32) epi = 2;
      else if (d->res32 && d->out32 && !d->mask16) epi = 3;
    } else if (light && !d->aux_mode) {
      epi = d->out_mode == PESR_OUT_SHUFFLE2 ? 5 : 6;
    }
  }
  // residual epilogue through shared memory (pair kernel, outputs on the GEMM's own pixel grid)
  if (d->bn_sums) {
    PESR_CHECK_ARG(epi == 1 && d->cout <= 512, "conv_igemm: bn_sums needs the light epilogue (16-bit NHWC output only) and "
                                               "cout <= 512");
    epi = 7;
  }
  const bool staged = pair && epi == 3 && gemm_grid && g_staged_enabled && tile_n == 1 && 32 % d->tile_w == 0 && d->ld_out16 % 8 == 0 &&
                      d->ld_out32 % 4 == 0 && d->ld_res32 % 4 == 0 && (long long)d->nb * d->h * d->w < (1ll << 31);
  if (staged) epi = 4;

  // multi-sub-block stages (single-CTA kernel only): see ConvK::sub_mode
  int sub_mode = 0, nsub = 1;
  if (!pair && g_sub_mode_enabled && d->ksplit <= 1 && !d->b_mn_major && d->block_n <= (g_sub_mode_enabled >= 2 ? 256 : 128)) {
    bool std9 = d->ntaps == 9 && d->nsrc == 1;
    for (int t = 0; std9 && t < 9; t++)
      std9 = d->tap_dh[t] == t / 3 - 1 && d->tap_dw[t] == t % 3 - 1 && d->tap_src[t] == 0 && d->tap_widx[t] == t;
    if (std9 && d->w_rows == 9 * d->cout && d->block_n <= 128 && tile_n == 1) { sub_mode = 2; nsub = 3; }
    else if ((d->cin / kKBlock) % 2 == 0) { sub_mode = 1; nsub = 2; }
  }

  ConvMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int s = 0; s < PESR_MAX_SRC; s++) {
    const int ss = s < d->nsrc ? s : 0;  // unused slots alias source 0 so every descriptor is valid
    PESR_CHECK_ARG(d->src[ss] != nullptr, "conv_igemm: source %d is null", ss);
    int r;
    if (sub_mode == 1) {
      uint64_t dims[5] = {64, (uint64_t)d->src_w[ss], (uint64_t)d->src_h[ss], (uint64_t)d->nb, (uint64_t)(d->cin / 64)};
      uint64_t str[4] = {(uint64_t)d->src_sw[ss] * 2, (uint64_t)d->src_sh[ss] * 2, (uint64_t)d->src_sn[ss] * 2, 128};
      uint32_t box[5] = {64, (uint32_t)d->tile_w, (uint32_t)d->tile_h, (uint32_t)tile_n, (uint32_t)nsub};
      r = get_tensor_map(&maps.a[s], d->src[ss], d->dtype, 5, dims, str, box);
    } else {
      uint64_t dims[4] = {(uint64_t)d->cin, (uint64_t)d->src_w[ss], (uint64_t)d->src_h[ss], (uint64_t)d->nb};
      uint64_t str[3] = {(uint64_t)d->src_sw[ss] * 2, (uint64_t)d->src_sh[ss] * 2, (uint64_t)d->src_sn[ss] * 2};
      uint32_t box[4] = {(uint32_t)kKBlock, (uint32_t)d->tile_w, (uint32_t)(sub_mode == 2 ? d->tile_h + 2 : d->tile_h),
                         (uint32_t)tile_n};
      r = get_tensor_map(&maps.a[s], d->src[ss], d->dtype, 4, dims, str, box);
    }
    if (r) return r;
  }
  const int ksplit = d->ksplit > 1 ? d->ksplit : 1;
  if (ksplit > 1)
    PESR_CHECK_ARG(d->out32 && !d->out16 && !d->res32 && !d->res16 && !d->mask16 && !d->bias && d->act == 0 &&
                       ksplit <= d->ntaps * (d->cin / kKBlock),
                   "conv_igemm: split-K writes raw fp32 partials only (ksplit %d)", ksplit);
  if (d->b_mn_major) {
    PESR_CHECK_ARG(d->ntaps == 1 && d->block_n % 64 == 0 && d->w_rows == d->cin,
                   "conv_igemm: MN-major weights need one tap, block_n %% 64 == 0 and w_rows == cin");
    uint64_t dims[2] = {(uint64_t)d->cout, (uint64_t)d->w_rows};
    uint64_t str[1] = {(uint64_t)d->cout * 2};
    uint32_t box[2] = {64, 64};
    int r = get_tensor_map(&maps.b, d->wpacked, d->dtype, 2, dims, str, box);
    if (r) return r;
  } else if (sub_mode == 2) {
    // weights [tap = dy*3+dx][cout][cin] viewed as [dy][dx*cout + o][cin]: one box = the 3 vertical taps of column dx
    uint64_t dims[3] = {(uint64_t)d->cin, (uint64_t)(3 * d->cout), 3};
    uint64_t str[2] = {(uint64_t)d->cin * 2, (uint64_t)3 * d->cout * d->cin * 2};
    uint32_t box[3] = {(uint32_t)kKBlock, (uint32_t)d->block_n, 3};
    int r = get_tensor_map(&maps.b, d->wpacked, d->dtype, 3, dims, str, box);
    if (r) return r;
  } else if (sub_mode == 1) {
    uint64_t dims[3] = {64, (uint64_t)d->w_rows, (uint64_t)(d->cin / 64)};
    uint64_t str[2] = {(uint64_t)d->cin * 2, 128};
    uint32_t box[3] = {64, (uint32_t)d->block_n, (uint32_t)nsub};
    int r = get_tensor_map(&maps.b, d->wpacked, d->dtype, 3, dims, str, box);
    if (r) return r;
  } else {
    uint64_t dims[2] = {(uint64_t)d->cin, (uint64_t)d->w_rows};
    uint64_t str[1] = {(uint64_t)d->cin * 2};
    uint32_t box[2] = {(uint32_t)kKBlock, (uint32_t)(pair ? d->block_n / 2 : d->block_n)};
    int r = get_tensor_map(&maps.b, d->wpacked, d->dtype, 2, dims, str, box);
    if (r) return r;
  }

  ConvK k;
  memset(&k, 0, sizeof(k));
  k.dtype = d->dtype;
  k.nb = d->nb; k.h = d->h; k.w = d->w; k.cin = d->cin; k.cout = d->cout; k.block_n = d->block_n;
  k.tile_h = d->tile_h; k.tile_w = d->tile_w; k.ntaps = d->ntaps;
  k.tile_n = tile_n;
  k.tiles_h = (d->h + d->tile_h - 1) / d->tile_h;
  k.tiles_w = (d->w + d->tile_w - 1) / d->tile_w;
  k.n_tiles = d->cout / d->block_n;
  const int m_tiles = img_groups * k.tiles_h * k.tiles_w;
  k.mn_tiles = (pair ? (m_tiles + 1) / 2 : m_tiles) * k.n_tiles;   // pair mode counts tiles of 2 x 128 pixels
  k.ksplit = ksplit;
  k.num_tiles = k.mn_tiles * ksplit * ncls;
  k.ncls = ncls;
  k.tiles_per_cls = k.mn_tiles * ksplit;
  k.cls_tap0[0] = 0;
  for (int c = 0; c < 4; c++) {
    k.cls_tap0[c + 1] = ncls > 1 ? k.cls_tap0[c] + (c < ncls ? d->cls_ntaps[c] : 0) : d->ntaps;
    k.cls_oy[c] = ncls > 1 ? d->cls_oy[c] : d->out_oy;
    k.cls_ox[c] = ncls > 1 ? d->cls_ox[c] : d->out_ox;
  }
  k.b_mn_major = d->b_mn_major ? 1 : 0;
  k.pdl_early_b = (weights_settled(stream) && pdl_enabled()) ? 1 : 0;
  k.dbg = g_dbg_buf;
  k.split_stride32 = d->split_stride32;
  k.kblocks_per_tap = d->cin / kKBlock;
  k.stage_bytes = kABytes + (pair ? d->block_n / 2 : d->block_n) * kKBlock * 2;   // per CTA
  k.sub_mode = sub_mode;
  k.nsub = nsub;
  if (sub_mode) {
    const int b_tile = d->block_n * kKBlock * 2;
    k.a_bytes = sub_mode == 2 ? (d->tile_h + 2) * d->tile_w * 128 : nsub * kABytes;
    k.a_sub_off = sub_mode == 2 ? d->tile_w * 128 : kABytes;
    k.b_sub_off = b_tile;
    k.stage_bytes = k.a_bytes + nsub * b_tile;
    k.steps_per_tile = sub_mode == 2 ? 3 * (d->cin / kKBlock) : d->ntaps * (d->cin / kKBlock) / nsub;
  }
  k.bn_sums = d->bn_sums;
  const int epi_smem = staged ? kEpiBytes : epi == 7 ? kBnEpiBytes : 0;
  int smem_budget = conv_smem_budget() - 4096 - epi_smem;
  // resident weights (ConvK::wres_bytes): one column tile, the whole packed tensor plus >= 3 activation stages fit, and
  // every CTA has at least two tiles to spread the one-off weight load over.  Measured in a chain (tools/perf_narrow.py):
  // 64 -> 128 @96x96x32 49.5 -> 43.5 us, 128 -> 64 unchanged, 64 -> 64 @192x192x32 118.5 -> 121.8 us (that layer is not
  // feed-bound and only pays the prologue), so the mode is taken for N = 128 tiles (g_wres_enabled == 2: wherever legal).
  k.wres_bytes = 0;
  if (sub_mode == 2 && g_wres_enabled && (d->block_n >= 128 || g_wres_enabled >= 2) && k.n_tiles == 1 && epi_smem == 0) {
    const long long wbytes = 9ll * d->cout * d->cin * 2;
    if (wbytes + 3ll * k.a_bytes <= smem_budget && k.num_tiles >= 2 * num_sms() && k.a_bytes % 1024 == 0) {
      k.wres_bytes = (int)wbytes;
      k.stage_bytes = k.a_bytes;
      smem_budget -= k.wres_bytes;
    }
  }
  k.stages = smem_budget / k.stage_bytes;
  if (k.stages > kMaxStages) k.stages = kMaxStages;
  for (int t = 0; t < PESR_MAX_TAPS; t++) {
    k.tap_dh[t] = d->tap_dh[t]; k.tap_dw[t] = d->tap_dw[t]; k.tap_src[t] = d->tap_src[t]; k.tap_widx[t] = d->tap_widx[t];
  }
  k.bias = d->bias; k.alpha = d->alpha; k.alpha_dev = d->alpha_dev;
  k.res32 = d->res32; k.ld_res32 = d->ld_res32;
  k.res16 = reinterpret_cast<const uint16_t*>(d->res16); k.ld_res16 = d->ld_res16;
  k.act = d->act;
  k.mask16 = reinterpret_cast<const uint16_t*>(d->mask16); k.ld_mask16 = d->ld_mask16; k.mask_mode = d->mask_mode;
  k.out32 = d->out32; k.ld_out32 = d->ld_out32;
  k.out16 = reinterpret_cast<uint16_t*>(d->out16); k.ld_out16 = d->ld_out16;
  k.out_mode = d->out_mode;
  k.out_h = d->out_h > 0 ? d->out_h : d->h;
  k.out_w = d->out_w > 0 ? d->out_w : d->w;
  k.out_sy = d->out_sy > 0 ? d->out_sy : 1;
  k.out_sx = d->out_sx > 0 ? d->out_sx : 1;
  k.out_oy = d->out_oy; k.out_ox = d->out_ox; k.out_coff = d->out_coff; k.ps_c = d->ps_c;
  k.aux_mode = d->aux_mode;

  // >= 120 KB of dynamic smem also guarantees one CTA per SM, so the 512-column TMEM allocation never contends.
  size_t smem = (size_t)k.stages * k.stage_bytes + epi_smem + k.wres_bytes + 1024 /*align*/ + 256 /*barriers*/ +
                2 * 256 * sizeof(float);
  if (smem < 120 * 1024) smem = 120 * 1024;
  // kernel variant: CTA pair x epilogue specialisation (see the template comment)
  typedef void (*KernelFn)(const ConvMaps, const ConvK);
  static const KernelFn kernels[2][8] = {
      {conv_igemm_kernel<false, 0>, conv_igemm_kernel<false, 1>, conv_igemm_kernel<false, 2>, conv_igemm_kernel<false, 3>,
       nullptr, conv_igemm_kernel<false, 5>, conv_igemm_kernel<false, 6>, conv_igemm_kernel<false, 7>},
      {conv_igemm_kernel<true, 0>, conv_igemm_kernel<true, 1>, conv_igemm_kernel<true, 2>, conv_igemm_kernel<true, 3>,
       conv_igemm_kernel<true, 4>, conv_igemm_kernel<true, 5>, conv_igemm_kernel<true, 6>, conv_igemm_kernel<true, 7>}};
  static bool attr_set = false;
  if (!attr_set) {
    for (int a = 0; a < 2; a++)
      for (int b = 0; b < 8; b++) {
        if (!kernels[a][b]) continue;
        cudaError_t e = cudaFuncSetAttribute(kernels[a][b], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) {
          set_error("conv_igemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
          return (int)e;
        }
      }
    attr_set = true;
  }
  const bool prof = profiling_enabled();
  if (prof) profile_begin(0, 2.0 * d->nb * d->h * d->w * (double)d->cout * d->cin * d->ntaps, stream);
  {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.blockDim = dim3(kNumThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (pair) {
      attr[na].id = cudaLaunchAttributeClusterDimension;
      attr[na].val.clusterDim.x = 2;
      attr[na].val.clusterDim.y = 1;
      attr[na].val.clusterDim.z = 1;
      na++;
    }
    if (pdl_enabled()) {
      // the kernel's prologue (barriers, TMEM, first weight boxes) overlaps the tail of the preceding kernel
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      na++;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    cudaError_t e;
    if (pair) {
      const int clusters = k.num_tiles < num_sms() / 2 ? k.num_tiles : num_sms() / 2;
      cfg.gridDim = dim3(2 * clusters);
      e = cudaLaunchKernelEx(&cfg, kernels[1][epi], maps, k);
    } else {
      cfg.gridDim = dim3(k.num_tiles < num_sms() ? k.num_tiles : num_sms());
      e = cudaLaunchKernelEx(&cfg, kernels[0][epi], maps, k);
    }
    if (e != cudaSuccess) {
      set_error("conv_igemm: launch failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
  }
  if (prof) profile_end(0, stream);
  count_launch();
  PESR_CHECK_LAUNCH("conv_igemm");
  return 0;
}
