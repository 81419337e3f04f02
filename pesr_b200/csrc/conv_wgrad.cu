// Backward-filter (wgrad) for sm_100a as a split-K tcgen05 GEMM whose K dimension is the pixel index
// (reference: nn.Conv2d backward-filter via cuDNN, model/basic.py:4-7).
//
//   part[split][tap][m][n] = sum_{p in split}  A[p][m] * B[p (+) tap][n]
//
// A = dy (NHWC, M = output channels), B = x (NHWC, N = input channels) shifted by the tap with TMA
// zero fill.  Both operands are "MN-major" for the tensor core: a TMA box of 64 channels x 64 pixels
// lands in smem as 64 rows (one per pixel = K index) of 128 swizzled bytes (64 channels = MN index),
// which is exactly the canonical SWIZZLE_128B MN-major layout; no transpose copy is ever made.
// The fp32 partial sums of each split are written to a workspace; pesr_wgrad_reduce sums them and
// scatters into the reference's OIHW gradient layout.
#include "common.cuh"
#include "host_util.cuh"

#include <stdlib.h>

namespace pesr {

static constexpr int kWgKBlock = 64;               // pixels per pipeline stage
static constexpr int kWgBoxBytes = kWgKBlock * 128;  // one 64-channel x 64-pixel box = 8 KB
static constexpr int kWgThreads = 192;
static constexpr int kWgMaxStages = 8;
static constexpr int kWgAccStride = 256;
#ifdef PESR_DEBUG_HOOKS
static constexpr bool kWgDebugHooks = true;    // timeline stamps compiled in (libpesr_b200_debug.so only)
#else
static constexpr bool kWgDebugHooks = false;
#endif

struct WgMaps {
  CUtensorMap a;
  CUtensorMap b[PESR_MAX_SRC];
};

struct WgK {
  int dtype;
  int nb, h, w, m_total, n_total, block_n, ntaps;
  int tile_h, tile_w, tiles_h, tiles_w, patches;
  int m_tiles, n_tiles, splits, num_items;
  int stages, stage_bytes, a_boxes, b_boxes;
  int lbo, sbo;  // MN-major descriptor strides (bytes)
  int8_t tap_dh[PESR_MAX_TAPS], tap_dw[PESR_MAX_TAPS], tap_src[PESR_MAX_TAPS];
  float* partials;
  float out_mul;               // stored value = sum * out_mul / (*out_div_dev)
  const float* out_div_dev;
  unsigned long long* dbg;
};

// kPair: a cluster of two CTAs drives one 256 (output channels) x block_n MMA (cta_group::2); CTA r stages its own
// 128 rows of dy and HALF of the x tile's channels, which cuts the L2 -> smem bytes per MMA by a third (the
// single-CTA kernel is bound by that traffic: 48 KB per 512 MMA cycles).  Protocol as in conv_igemm.cu.
template <bool kPair>
__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ WgMaps maps, const __grid_constant__ WgK p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler (see conv_igemm.cu)
  const int lane = threadIdx.x & 31;
  const int rank = kPair ? (int)cluster_ctarank() : 0;
  const bool leader = rank == 0;
  const int worker = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int nworkers = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  uint8_t* epi_tiles = smem + (size_t)p.stages * p.stage_bytes;      // 4 x 2 KB transposition tiles of the epilogue warps
  uint8_t* tail = epi_tiles + 4 * 2048;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + kWgMaxStages;
  uint64_t* tmem_full = empty_bar + kWgMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&maps.a);
    for (int i = 0; i < 4; i++) tma_prefetch_desc(&maps.b[i]);
    for (int s = 0; s < p.stages; s++) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], kPair ? 8 : 128);
    }
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
  // PDL: everything above overlapped the tail of the preceding kernel; both operands depend on it
  griddep_wait();
  griddep_launch();

  unsigned long long* dbg = (kWgDebugHooks && p.dbg && blockIdx.x == 0) ? p.dbg : nullptr;
  if (dbg && threadIdx.x == 0) {
    unsigned long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    dbg[0] = clock64();
    dbg[62] = g;
  }
  if (kWgDebugHooks && p.dbg && threadIdx.x == 0) {   // per-CTA start / end wall-clock stamps (all blocks)
    unsigned long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    p.dbg[64 + 2 * blockIdx.x] = g;
  }
  // item -> (split, tap, m_tile, n_tile); split slowest so that concurrently running CTAs share pixels in L2
  auto decode = [&](int item, int& split, int& tap, int& mt, int& nt) {
    nt = item % p.n_tiles;
    int r = item / p.n_tiles;
    mt = r % p.m_tiles;
    r /= p.m_tiles;
    tap = r % p.ntaps;
    split = r / p.ntaps;
    if (kPair) mt = 2 * mt + rank;
  };
  auto patch_range = [&](int split, int& p0, int& p1) {
    p0 = (int)(((long long)split * p.patches) / p.splits);
    p1 = (int)(((long long)(split + 1) * p.patches) / p.splits);
  };

  if (warp == 0) {
    {   // whole warp, warp-uniform control flow; one elected lane issues (see common.cuh)
      int stage = 0;
      uint32_t phase = 0;
      for (int item = worker; item < p.num_items; item += nworkers) {
        int split, tap, mt, nt, p0, p1;
        decode(item, split, tap, mt, nt);
        patch_range(split, p0, p1);
        const CUtensorMap* mb = &maps.b[p.tap_src[tap]];
        const int dh = p.tap_dh[tap], dw = p.tap_dw[tap];
        // patch -> (img, th, tw), walked incrementally (the producer loop is latency-critical: no per-step divisions)
        int tw = p0 % p.tiles_w;
        int th = (p0 / p.tiles_w) % p.tiles_h;
        int img = p0 / (p.tiles_w * p.tiles_h);
        for (int pt = p0; pt < p1; pt++) {
          const int h0 = th * p.tile_h, w0 = tw * p.tile_w;
          mbar_wait_w(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * p.stage_bytes;
          uint8_t* sb = sa + p.a_boxes * kWgBoxBytes;
          if (kPair) {
            // one 5-D box per operand: dim 4 walks the 64-channel chunks, which land back to back in smem
            if (leader) mbar_expect_tx_w(&full_bar[stage], 2u * (uint32_t)p.stage_bytes);
            tma2_load_5d_w(sa, &maps.a, &full_bar[stage], 0, w0, h0, img, mt * 2);
            tma2_load_5d_w(sb, mb, &full_bar[stage], 0, w0 + dw, h0 + dh, img, nt * (p.block_n / 64) + rank * p.b_boxes);
          } else {
            mbar_expect_tx_w(&full_bar[stage], (uint32_t)p.stage_bytes);
            tma_load_5d_w(sa, &maps.a, &full_bar[stage], 0, w0, h0, img, mt * 2);
            tma_load_5d_w(sb, mb, &full_bar[stage], 0, w0 + dw, h0 + dh, img, nt * (p.block_n / 64));
          }
          if (++tw == p.tiles_w) { tw = 0; if (++th == p.tiles_h) { th = 0; img++; } }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(kPair ? 256 : 128, p.block_n, p.dtype, 1, 1);
    // warp-uniform copies (through a shuffle the compiler knows it): the accumulator address and the constant part of the
    // operand descriptors stay in uniform registers instead of being converted in front of every tcgen05.mma
    const uint32_t tmem_base_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint64_t desc_mn = make_smem_desc(0, p.lbo, p.sbo);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = worker; leader && item < p.num_items; item += nworkers) {
      int split, tap, mt, nt, p0, p1;
      decode(item, split, tap, mt, nt);
      patch_range(split, p0, p1);
      if (kPair) mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1); else mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base_u + (uint32_t)(acc * kWgAccStride);
      for (int pt = p0; pt < p1; pt++) {
        mbar_wait_w(&full_bar[stage], phase);
        tc_fence_after();
        if (dbg && lane == 0 && pt == p0) dbg[9] = clock64();
        if (dbg && lane == 0 && pt == p0 + 8) dbg[11] = clock64();
        {
          const uint32_t a_addr = smem_u32(smem + (size_t)stage * p.stage_bytes);
          const uint32_t b_addr = a_addr + p.a_boxes * kWgBoxBytes;
          const uint32_t a16 = a_addr >> 4, b16 = b_addr >> 4;      // 1024-byte aligned tiles: the address field just adds
#pragma unroll
          for (int k = 0; k < kWgKBlock / 16; k++) {
            const uint64_t da = desc_mn | (uint64_t)(a16 + k * 128);
            const uint64_t db = desc_mn | (uint64_t)(b16 + k * 128);
            if (kPair) umma2_f16_w(d_tmem, da, db, idesc, (pt > p0 || k > 0) ? 1u : 0u);
            else       umma_f16_w(d_tmem, da, db, idesc, (pt > p0 || k > 0) ? 1u : 0u);
          }
          if (kPair) {
            umma2_commit_both_w(&empty_bar[stage]);
            if (pt == p1 - 1) umma2_commit_both_w(&tmem_full[acc]);
          } else {
            umma_commit_w(&empty_bar[stage]);
            if (pt == p1 - 1) umma_commit_w(&tmem_full[acc]);
          }
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (dbg && lane == 0) dbg[10] = clock64();
      if (p1 > p0) {
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int ch_lo = 0, ch_hi = p.block_n / 32;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = worker; item < p.num_items; item += nworkers) {
      int split, tap, mt, nt, p0, p1;
      decode(item, split, tap, mt, nt);
      patch_range(split, p0, p1);
      const int m = mt * 128 + row;
      float* dst = p.partials + (((long long)split * p.ntaps + tap) * p.m_total + m) * p.n_total + nt * p.block_n;
      if (p1 <= p0) {  // empty split: its partial is all zeros
        if (m < p.m_total)
          for (int j = ch_lo * 32; j < ch_hi * 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(0, 0, 0, 0);
        continue;
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (dbg && threadIdx.x == 64) dbg[16] = clock64();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kWgAccStride);
      // One thread per accumulator row would store 32 bytes to each of 32 different lines per instruction (the 128 KB
      // tile then takes ~8.8 K cycles, all of it exposed: a CTA usually has a single work item).  The rows go through
      // an XOR-swizzled 2 KB tile private to the warp instead and leave as 64-byte row segments.  (A 4 KB tile with
      // 128-byte segments made the kernel itself 2% faster and the step slower: with 16 KB more shared memory the
      // blocks of the PDL-launched reduction kernel that follows no longer fit beside the running CTA.)
      const uint32_t tb = smem_u32(epi_tiles) + (uint32_t)quarter * 2048u;      // [32 rows][64 B] per warp
      const uint32_t xr = (uint32_t)((lane >> 1) & 3);
      const int c16 = lane & 3, r16 = lane >> 2;          // store pattern: 4 lanes per row (16 B each), rows i*8 + lane/4
      float* dst_w = p.partials + (((long long)split * p.ntaps + tap) * p.m_total + mt * 128 + quarter * 32) * p.n_total +
                     nt * p.block_n + c16 * 4;
      const int rows_left = p.m_total - (mt * 128 + quarter * 32);      // rows of this warp inside the matrix
      float osc = p.out_mul;
      if (p.out_div_dev) osc /= __ldg(p.out_div_dev);
      const bool scaled = osc != 1.f;
      uint32_t v[32], vn[32];
      tmem_ld32(taddr + ch_lo * 32, v);
      tmem_ld_wait();
      for (int ch = ch_lo; ch < ch_hi; ch++) {
        const bool more = ch + 1 < ch_hi;
        if (more) tmem_ld32(taddr + (ch + 1) * 32, vn);
#pragma unroll
        for (int half = 0; half < 2; half++) {            // 16 columns per pass keep the tile at 2 KB per warp
#pragma unroll
          for (int c = 0; c < 4; c++)
            sts128(tb + (uint32_t)lane * 64u + (((uint32_t)c ^ xr) << 4), v[16 * half + 4 * c], v[16 * half + 4 * c + 1],
                   v[16 * half + 4 * c + 2], v[16 * half + 4 * c + 3]);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const uint32_t rr = (uint32_t)(i * 8 + r16);
            uint4 o = lds128(tb + rr * 64u + (((uint32_t)c16 ^ ((rr >> 1) & 3u)) << 4));
            if (scaled) {
              o.x = __float_as_uint(__uint_as_float(o.x) * osc); o.y = __float_as_uint(__uint_as_float(o.y) * osc);
              o.z = __float_as_uint(__uint_as_float(o.z) * osc); o.w = __float_as_uint(__uint_as_float(o.w) * osc);
            }
            if ((int)rr < rows_left) stg128(dst_w + (long long)rr * p.n_total + ch * 32 + half * 16, o);
          }
          __syncwarp();      // the tile is rewritten by the next pass
        }
        if (more) {
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = vn[j];
        } else if (item + 2 * nworkers < p.num_items) {      // the accumulator is only handed back if it is used again
          tc_fence_before();
          if (kPair) {
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
          } else {
            mbar_arrive(&tmem_empty[acc]);
          }
        }
      }
      if (dbg && threadIdx.x == 64) dbg[17] = clock64();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();
  if (dbg && threadIdx.x == 0) {
    unsigned long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    dbg[61] = clock64();
    dbg[63] = g;
  }
  if (kWgDebugHooks && p.dbg && threadIdx.x == 0) {
    unsigned long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    p.dbg[65 + 2 * blockIdx.x] = g;
  }
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
// split-K reduction + scatter into the fp32 OIHW gradient
// ------------------------------------------------------------------------------------------
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, int splits, int ntaps, int m_total, int n_total,
                                    int map_mode, int co, int ci, float scale, const float* __restrict__ inv_scale_dev,
                                    int accumulate, float* __restrict__ grad) {
  griddep_wait();   // PDL: see launch_pdl
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool col = map_mode == PESR_WMAP_COL_IN || map_mode == PESR_WMAP_COL_OUT;   // one thread per (o, i, tap)
  const long long total = (long long)co * ci * (col ? 9 : 1);
  if (idx >= total) return;
  if (inv_scale_dev) scale /= __ldg(inv_scale_dev);
  const long long tap_stride = (long long)m_total * n_total;
  const long long split_stride = tap_stride * ntaps;
  if (map_mode == PESR_WMAP_OIHW || map_mode == PESR_WMAP_OIHW_PS) {
    // thread <-> (m, n): n fastest => coalesced partial reads
    const int n = (int)(idx % ci), m = (int)(idx / ci);
    int o = m;
    if (map_mode == PESR_WMAP_OIHW_PS) {
      const int c_ps = co / 4;
      o = (m % c_ps) * 4 + (m / c_ps);
    }
    float* g = grad + ((long long)o * ci + n) * ntaps;
    for (int t = 0; t < ntaps; t++) {
      float s = 0.f;
      const float* src = part + t * tap_stride + (long long)m * n_total + n;
#pragma unroll 8
      for (int k = 0; k < splits; k++) s += src[k * split_stride];
      s *= scale;
      g[t] = accumulate ? g[t] + s : s;
    }
  } else if (map_mode == PESR_WMAP_COL_IN) {
    // partial [m = o][n = tap*ci + i], 9 taps folded into n; thread <-> (o, n): coalesced over n, every thread sums
    // its splits (these layers have one tile and up to #SM splits: a thread per (o, i) walking 9 taps serially took 54 us)
    const int nn = (int)(idx % (9 * ci)), o = (int)(idx / (9 * ci));
    const int t = nn / ci, i = nn - t * ci;
    float s = 0.f;
    const float* src = part + (long long)o * n_total + nn;
#pragma unroll 4
    for (int k = 0; k < splits; k++) s += src[k * split_stride];
    s *= scale;
    float* g = grad + ((long long)o * ci + i) * 9 + t;
    *g = accumulate ? *g + s : s;
  } else {
    // PESR_WMAP_COL_OUT: partial [m = i][n = tap*co + o]; thread <-> (i, n)
    const int nn = (int)(idx % (9 * co)), i = (int)(idx / (9 * co));
    const int t = nn / co, o = nn - t * co;
    float s = 0.f;
    const float* src = part + (long long)i * n_total + nn;
#pragma unroll 4
    for (int k = 0; k < splits; k++) s += src[k * split_stride];
    s *= scale;
    float* g = grad + ((long long)o * ci + i) * 9 + t;
    *g = accumulate ? *g + s : s;
  }
}

// OIHW / OIHW_PS maps with ntaps == 9: block = one GEMM row m (an output channel), thread = input channel(s).
// Partial reads are coalesced over n and independent across the 9 taps x splits (high memory-level parallelism);
// the 9*ci results of the row are staged in smem and written as one contiguous, coalesced run.
__global__ void __launch_bounds__(256)
wgrad_reduce_rows_kernel(const float* __restrict__ part, int splits, int m_total, int n_total, int map_mode, int co,
                         int ci, float scale, const float* __restrict__ div_dev, int accumulate,
                         float* __restrict__ grad) {
  griddep_wait();   // PDL: see launch_pdl
  extern __shared__ float row_s[];   // [ci][9]
  const int m = blockIdx.x;
  if (div_dev) scale /= __ldg(div_dev);
  const long long tap_stride = (long long)m_total * n_total;
  const long long split_stride = tap_stride * 9;
  for (int n = threadIdx.x; n < ci; n += blockDim.x) {
    float s[9];
#pragma unroll
    for (int t = 0; t < 9; t++) s[t] = 0.f;
    const float* src = part + (long long)m * n_total + n;
    int k = 0;
    for (; k + 4 <= splits; k += 4) {        // 36 independent loads in flight (see wgrad_reduce_bias_kernel)
      float v[4][9];
#pragma unroll
      for (int u = 0; u < 4; u++)
#pragma unroll
        for (int t = 0; t < 9; t++) v[u][t] = src[(k + u) * split_stride + t * tap_stride];
#pragma unroll
      for (int u = 0; u < 4; u++)
#pragma unroll
        for (int t = 0; t < 9; t++) s[t] += v[u][t];
    }
    for (; k < splits; k++) {
#pragma unroll
      for (int t = 0; t < 9; t++) s[t] += src[k * split_stride + t * tap_stride];
    }
#pragma unroll
    for (int t = 0; t < 9; t++) row_s[n * 9 + t] = s[t] * scale;
  }
  __syncthreads();
  int o = m;
  if (map_mode == PESR_WMAP_OIHW_PS) {
    const int c_ps = co / 4;
    o = (m % c_ps) * 4 + (m / c_ps);
  }
  float* g = grad + (long long)o * ci * 9;
  for (int i = threadIdx.x; i < ci * 9; i += blockDim.x) g[i] = accumulate ? g[i] + row_s[i] : row_s[i];
}

// Split-K reduction of a 3x3 weight gradient AND the bias gradient (column sums of dY) in one launch: both are short,
// latency-bound kernels that follow every wgrad launch, so they run side by side as one grid.  Blocks [0, co) reduce one
// output-channel row each (as wgrad_reduce_rows_kernel); the remaining blocks column-sum dY (as colsum16_vec_kernel)
// into a PRE-ZEROED bias gradient.  Block 0 also zeroes `zero_next`, the bias gradient the NEXT call accumulates into.
__global__ void __launch_bounds__(256)
wgrad_reduce_bias_kernel(const float* __restrict__ part, int splits, int m_total, int n_total, int map_mode, int co, int ci,
                         float scale, const float* __restrict__ div_dev, int accumulate, float* __restrict__ grad,
                         const uint4* __restrict__ x, long long npix, int c, int ldv, float bmul, int bf,
                         float* __restrict__ bias_out, float* __restrict__ zero_next, int zero_n) {
  griddep_wait();   // PDL: see launch_pdl
  extern __shared__ float row_s[];   // [ci][9]
  __shared__ float red[2048];
  float inv = 1.f;
  if (div_dev) inv = 1.f / __ldg(div_dev);
  if ((int)blockIdx.x < co) {
    const int m = blockIdx.x;
    scale *= inv;
    if (m == 0 && zero_next)
      for (int i = threadIdx.x; i < zero_n; i += blockDim.x) zero_next[i] = 0.f;
    const long long tap_stride = (long long)m_total * n_total;
    const long long split_stride = tap_stride * 9;
    for (int n = threadIdx.x; n < ci; n += blockDim.x) {
      float s[9];
#pragma unroll
      for (int t = 0; t < 9; t++) s[t] = 0.f;
      const float* src = part + (long long)m * n_total + n;
      int k = 0;
      for (; k + 4 <= splits; k += 4) {      // 36 independent loads in flight per thread: the kernel is L2-latency-bound
                                             // (10.1 -> 7.6 us in-stream, tools/perf_chain.py; 72 in flight: no further gain)
        float v[4][9];
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
          for (int t = 0; t < 9; t++) v[u][t] = src[(k + u) * split_stride + t * tap_stride];
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
          for (int t = 0; t < 9; t++) s[t] += v[u][t];
      }
      for (; k < splits; k++) {
#pragma unroll
        for (int t = 0; t < 9; t++) s[t] += src[k * split_stride + t * tap_stride];
      }
#pragma unroll
      for (int t = 0; t < 9; t++) row_s[n * 9 + t] = s[t] * scale;
    }
    __syncthreads();
    int o = m;
    if (map_mode == PESR_WMAP_OIHW_PS) {
      const int c_ps = co / 4;
      o = (m % c_ps) * 4 + (m / c_ps);
    }
    float* g = grad + (long long)o * ci * 9;
    for (int i = threadIdx.x; i < ci * 9; i += blockDim.x) g[i] = accumulate ? g[i] + row_s[i] : row_s[i];
    return;
  }
  // ---- bias gradient: column sums of dY over this block's pixel rows
  const int bid = blockIdx.x - co, nblk = gridDim.x - co;
  const int tpr = c >> 3, rpp = 256 / tpr;
  const int v = threadIdx.x % tpr, r = threadIdx.x / tpr;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; j++) acc[j] = 0.f;
  for (long long px = (long long)bid * rpp + r; px < npix; px += (long long)nblk * rpp) {
    const uint4 u = x[px * ldv + v];
    const float2 a = unpack2(u.x, bf), b = unpack2(u.y, bf), cc = unpack2(u.z, bf), d = unpack2(u.w, bf);
    acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
    acc[4] += cc.x; acc[5] += cc.y; acc[6] += d.x; acc[7] += d.y;
  }
#pragma unroll
  for (int j = 0; j < 8; j++) red[r * c + v * 8 + j] = acc[j];
  __syncthreads();
  bmul *= inv;
  for (int ch = threadIdx.x; ch < c; ch += 256) {
    float t = 0.f;
    for (int k = 0; k < rpp; k++) t += red[k * c + ch];
    atomicAdd(bias_out + ch, t * bmul);
  }
}

static int g_dbg_lbo = 0, g_dbg_sbo = 0;
static int g_wg_pair = 1;
static unsigned long long* g_wg_dbg = nullptr;

}  // namespace pesr

using namespace pesr;

// Debug hook used by the bring-up tests only: override the MN-major descriptor strides (0 = default).
#ifdef PESR_DEBUG_HOOKS
extern "C" void pesr_debug_wgrad_timeline(void* buf) { g_wg_dbg = reinterpret_cast<unsigned long long*>(buf); }

extern "C" void pesr_debug_wgrad_desc(int lbo_bytes, int sbo_bytes) {
  if (lbo_bytes == -1) { g_wg_pair = sbo_bytes; return; }
  if (lbo_bytes == -2) return;   // (-2, mode): load-skip experiment of the bring-up phase, removed
  g_dbg_lbo = lbo_bytes;
  g_dbg_sbo = sbo_bytes;
}
#endif

extern "C" int pesr_conv_wgrad(const pesr_wgrad_desc* d, int32_t* splits_out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(d != nullptr, "conv_wgrad: null descriptor");
  PESR_CHECK_ARG(d->dtype == PESR_DT_F16 || d->dtype == PESR_DT_BF16, "conv_wgrad: bad dtype %d", d->dtype);
  PESR_CHECK_ARG(d->nb > 0 && d->h > 0 && d->w > 0, "conv_wgrad: empty pixel grid");
  PESR_CHECK_ARG(d->block_m == 128, "conv_wgrad: block_m must be 128 (got %d)", d->block_m);
  PESR_CHECK_ARG(d->block_n == 64 || d->block_n == 128 || d->block_n == 256, "conv_wgrad: block_n %d", d->block_n);
  PESR_CHECK_ARG(d->m_total > 0 && d->n_total > 0 && d->n_total % d->block_n == 0,
                 "conv_wgrad: n_total %d not a multiple of block_n %d", d->n_total, d->block_n);
  PESR_CHECK_ARG(d->m_total % 64 == 0, "conv_wgrad: m_total %d must be a multiple of 64", d->m_total);
  PESR_CHECK_ARG(d->ntaps >= 1 && d->ntaps <= PESR_MAX_TAPS, "conv_wgrad: ntaps %d", d->ntaps);
  PESR_CHECK_ARG(d->nsrc >= 1 && d->nsrc <= PESR_MAX_SRC, "conv_wgrad: nsrc %d", d->nsrc);
  PESR_CHECK_ARG(d->a != nullptr && d->partials != nullptr, "conv_wgrad: null operand");
  PESR_CHECK_ARG(d->a_c % 8 == 0, "conv_wgrad: a_c %d must be a multiple of 8", d->a_c);
  PESR_CHECK_ARG(((uintptr_t)d->partials % 32) == 0 && d->n_total % 8 == 0, "conv_wgrad: partials must be 32-byte aligned");

  WgK k;
  memset(&k, 0, sizeof(k));
  k.dtype = d->dtype;
  k.nb = d->nb; k.h = d->h; k.w = d->w; k.m_total = d->m_total; k.n_total = d->n_total;
  k.block_n = d->block_n; k.ntaps = d->ntaps;
  // 64-pixel patches: wide images use 4x16, narrow ones 8x8
  if (d->w % 16 == 0 || d->w > 24) { k.tile_h = 4; k.tile_w = 16; } else { k.tile_h = 8; k.tile_w = 8; }
  k.tiles_h = (d->h + k.tile_h - 1) / k.tile_h;
  k.tiles_w = (d->w + k.tile_w - 1) / k.tile_w;
  k.patches = d->nb * k.tiles_h * k.tiles_w;
  // CTA-pair path: 256 output-channel rows per MMA; needs m_total % 256 == 0 and a 128-aligned half of block_n
  static int pair_env = -1;
  if (pair_env < 0) {
    const char* e = getenv("PESR_NO_PAIR");
    pair_env = (e && e[0] == '1') ? 0 : 1;
  }
  const bool pair = pair_env == 1 && g_wg_pair != 0 && d->m_total % 256 == 0 && d->block_n == 256;
  k.m_tiles = pair ? d->m_total / 256 : (d->m_total + 127) / 128;
  k.n_tiles = d->n_total / d->block_n;
  const int base_items = d->ntaps * k.m_tiles * k.n_tiles;
  int splits = d->splits;
  if (splits <= 0) {
    splits = (pair ? num_sms() / 2 : num_sms()) / base_items;
    if (splits < 1) splits = 1;
  }
  if (splits > k.patches) splits = k.patches;
  const long long need = (long long)splits * d->ntaps * d->m_total * d->n_total;
  if (need > d->partials_elems) {
    // shrink to what the workspace can hold
    long long per = (long long)d->ntaps * d->m_total * d->n_total;
    splits = (int)(d->partials_elems / per);
    if (splits < 1) {
      set_error("conv_wgrad: workspace holds %lld floats, one split needs %lld", (long long)d->partials_elems, per);
      return PESR_E_WORKSPACE;
    }
  }
  k.splits = splits;
  k.num_items = base_items * splits;
  k.a_boxes = 2;
  k.b_boxes = pair ? d->block_n / 128 : d->block_n / 64;   // per CTA
  k.stage_bytes = (k.a_boxes + k.b_boxes) * kWgBoxBytes;
  k.stages = (conv_smem_budget() - 4096 - 4 * 2048) / k.stage_bytes;
  if (k.stages > kWgMaxStages) k.stages = kWgMaxStages;
  k.lbo = g_dbg_lbo ? g_dbg_lbo : kWgBoxBytes;
  k.sbo = g_dbg_sbo ? g_dbg_sbo : 1024;
  for (int t = 0; t < PESR_MAX_TAPS; t++) {
    k.tap_dh[t] = d->tap_dh[t]; k.tap_dw[t] = d->tap_dw[t]; k.tap_src[t] = d->tap_src[t];
  }
  for (int t = 0; t < d->ntaps; t++)
    PESR_CHECK_ARG(d->tap_src[t] >= 0 && d->tap_src[t] < d->nsrc, "conv_wgrad: tap %d reads source %d", t, d->tap_src[t]);
  k.partials = d->partials;
  k.out_mul = d->out_mul != 0.f ? d->out_mul : 1.f;
  k.out_div_dev = d->out_div_dev;
  k.dbg = g_wg_dbg;

  WgMaps maps;
  memset(&maps, 0, sizeof(maps));
  {
    // 5-D view [chunk][n][h][w][64]: dim 4 strides over the 64-channel chunks so that ONE TMA op stages all of them
    uint64_t dims[5] = {64, (uint64_t)d->w, (uint64_t)d->h, (uint64_t)d->nb, (uint64_t)(d->m_total / 64)};
    uint64_t str[4] = {(uint64_t)d->a_c * 2, (uint64_t)d->w * d->a_c * 2, (uint64_t)d->h * d->w * d->a_c * 2, 128};
    uint32_t box[5] = {64, (uint32_t)k.tile_w, (uint32_t)k.tile_h, 1, (uint32_t)k.a_boxes};
    int r = get_tensor_map(&maps.a, d->a, d->dtype, 5, dims, str, box);
    if (r) return r;
  }
  for (int s = 0; s < PESR_MAX_SRC; s++) {
    const int ss = s < d->nsrc ? s : 0;
    PESR_CHECK_ARG(d->b[ss] != nullptr, "conv_wgrad: b source %d is null", ss);
    uint64_t dims[5] = {64, (uint64_t)d->b_w[ss], (uint64_t)d->b_h[ss], (uint64_t)d->nb, (uint64_t)(d->n_total / 64)};
    uint64_t str[4] = {(uint64_t)d->b_sw[ss] * 2, (uint64_t)d->b_sh[ss] * 2, (uint64_t)d->b_sn[ss] * 2, 128};
    uint32_t box[5] = {64, (uint32_t)k.tile_w, (uint32_t)k.tile_h, 1, (uint32_t)k.b_boxes};
    int r = get_tensor_map(&maps.b[s], d->b[ss], d->dtype, 5, dims, str, box);
    if (r) return r;
  }

  size_t smem = (size_t)k.stages * k.stage_bytes + 4 * 2048 + 1024 + 256;
  if (smem < 120 * 1024) smem = 120 * 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("conv_wgrad: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  const bool prof = profiling_enabled();
  if (prof) profile_begin(1, 2.0 * d->nb * d->h * d->w * (double)d->m_total * d->n_total * d->ntaps, stream);
  {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.blockDim = dim3(kWgThreads);
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
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      na++;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    cudaError_t e;
    if (pair) {
      const int clusters = k.num_items < num_sms() / 2 ? k.num_items : num_sms() / 2;
      cfg.gridDim = dim3(2 * clusters);
      e = cudaLaunchKernelEx(&cfg, conv_wgrad_kernel<true>, maps, k);
    } else {
      cfg.gridDim = dim3(k.num_items < num_sms() ? k.num_items : num_sms());
      e = cudaLaunchKernelEx(&cfg, conv_wgrad_kernel<false>, maps, k);
    }
    if (e != cudaSuccess) {
      set_error("conv_wgrad: launch failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
  }
  if (prof) profile_end(1, stream);
  count_launch();
  PESR_CHECK_LAUNCH("conv_wgrad");
  if (splits_out) *splits_out = splits;
  return 0;
}

extern "C" int pesr_wgrad_reduce(const float* partials, int32_t splits, int32_t ntaps, int32_t m_total,
                                 int32_t n_total, int32_t map_mode, int32_t co, int32_t ci, float scale,
                                 const float* inv_scale_dev, int32_t accumulate, float* grad_oihw, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(partials && grad_oihw, "wgrad_reduce: null pointer");
  PESR_CHECK_ARG(splits >= 1 && ntaps >= 1, "wgrad_reduce: bad splits/ntaps");
  PESR_CHECK_ARG(map_mode >= 0 && map_mode <= 3, "wgrad_reduce: bad map_mode %d", map_mode);
  if (map_mode == PESR_WMAP_OIHW || map_mode == PESR_WMAP_OIHW_PS)
    PESR_CHECK_ARG(co <= m_total && ci <= n_total, "wgrad_reduce: %dx%d does not fit partial %dx%d", co, ci, m_total,
                   n_total);
  if (map_mode == PESR_WMAP_COL_IN)
    PESR_CHECK_ARG(co <= m_total && 9 * ci <= n_total && ntaps == 1, "wgrad_reduce: COL_IN shape mismatch");
  if (map_mode == PESR_WMAP_COL_OUT)
    PESR_CHECK_ARG(ci <= m_total && 9 * co <= n_total && ntaps == 1, "wgrad_reduce: COL_OUT shape mismatch");
  if ((map_mode == PESR_WMAP_OIHW || map_mode == PESR_WMAP_OIHW_PS) && ntaps == 9 && ci * 9 * sizeof(float) <= 48 * 1024) {
    launch_pdl(wgrad_reduce_rows_kernel, co, 256, ci * 9 * sizeof(float), stream, partials, splits, m_total, n_total, map_mode,
               co, ci, scale, inv_scale_dev, accumulate, grad_oihw);
    count_launch();
    PESR_CHECK_LAUNCH("wgrad_reduce");
    return 0;
  }
  const bool col = map_mode == PESR_WMAP_COL_IN || map_mode == PESR_WMAP_COL_OUT;
  const long long total = (long long)co * ci * (col ? 9 : 1);
  const int threads = 256;
  const int blocks = (int)((total + threads - 1) / threads);
  launch_pdl(wgrad_reduce_kernel, blocks, threads, 0, stream, partials, splits, ntaps, m_total, n_total, map_mode, co, ci,
                                                     scale, inv_scale_dev, accumulate, grad_oihw);
  count_launch();
  PESR_CHECK_LAUNCH("wgrad_reduce");
  return 0;
}

extern "C" int pesr_wgrad_reduce_bias(const float* partials, int32_t splits, int32_t ntaps, int32_t m_total,
                                      int32_t n_total, int32_t map_mode, int32_t co, int32_t ci, float scale,
                                      const float* inv_scale_dev, int32_t accumulate, float* grad_oihw, const void* dy16,
                                      int64_t npix, int32_t c, int32_t ldc, float bias_mul, int32_t dtype,
                                      float* bias_grad, float* zero_next, int32_t zero_n, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(partials && grad_oihw && dy16 && bias_grad, "wgrad_reduce_bias: null pointer");
  PESR_CHECK_ARG(splits >= 1 && ntaps == 9 && (map_mode == PESR_WMAP_OIHW || map_mode == PESR_WMAP_OIHW_PS),
                 "wgrad_reduce_bias: 3x3 OIHW gradients only (ntaps %d, map %d)", ntaps, map_mode);
  PESR_CHECK_ARG(co <= m_total && ci <= n_total && ci * 9 * sizeof(float) <= 40 * 1024,
                 "wgrad_reduce_bias: %dx%d does not fit partial %dx%d", co, ci, m_total, n_total);
  const int tpr = c / 8;
  PESR_CHECK_ARG(npix > 0 && c > 0 && c % 8 == 0 && ldc % 8 == 0 && tpr <= 256 && 256 % tpr == 0 &&
                     ((uintptr_t)dy16 % 16) == 0,
                 "wgrad_reduce_bias: dY must be 16-byte aligned with c in {8,16,...,2048} dividing 2048 (c %d, ldc %d)", c, ldc);
  const int rpp = 256 / tpr;
  long long bx = (npix + (long long)rpp * 16 - 1) / ((long long)rpp * 16);
  if (bx > 148 * 4) bx = 148 * 4;
  if (bx < 1) bx = 1;
  launch_pdl(wgrad_reduce_bias_kernel, co + (unsigned)bx, 256, ci * 9 * sizeof(float), stream, partials, splits, m_total,
             n_total, map_mode, co, ci, scale, inv_scale_dev, accumulate, grad_oihw, reinterpret_cast<const uint4*>(dy16),
             (long long)npix, c, ldc / 8, bias_mul, dtype, bias_grad, zero_next, zero_n);
  count_launch();
  PESR_CHECK_LAUNCH("wgrad_reduce_bias");
  return 0;
}
