// HBM-bound layout / edge kernels: weight packing, im2col / col2im for the 3-channel convolutions
// (with the MeanShift affine of model/basic.py:9-17 fused), NCHW<->NHWC conversion, column sums
// (bias gradients), the dynamic gradient scale and the MeanShift parameter gradients.
#include "common.cuh"
#include "host_util.cuh"

#include <stdlib.h>

#include <vector>

namespace pesr {

static inline int blocks_for(long long n, int threads, int cap = 148 * 16) {
  long long b = (n + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ------------------------------------------------------------------------------------------
// weight packing
// ------------------------------------------------------------------------------------------
__global__ void pack_weights_kernel(const float* __restrict__ w, int co, int ci, int taps, int mode, int pad_to,
                                    int bf, uint16_t* __restrict__ out, long long total, int rows, int kdim) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % kdim);
    const int r = (int)(idx / kdim);
    float v = 0.f;
    const int c_ps = co / 4;
    switch (mode) {
      case 0: {  // out[tap][o][i]
        const int tap = r / co, o = r % co;
        v = w[((long long)o * ci + k) * taps + tap];
      } break;
      case 1: {  // out[tap][i][o] = w[o][i][tapf]
        const int tap = r / ci, i = r % ci;
        v = w[((long long)k * ci + i) * taps + (taps - 1 - tap)];
      } break;
      case 2: {  // out[tap][o'][i], o' = ij*C + c <- o = c*4 + ij
        const int tap = r / co, op = r % co;
        const int o = (op % c_ps) * 4 + (op / c_ps);
        v = w[((long long)o * ci + k) * taps + tap];
      } break;
      case 3: {  // out[tap][i][o'] = w[o][i][tapf]
        const int tap = r / ci, i = r % ci;
        const int o = (k % c_ps) * 4 + (k / c_ps);
        v = w[((long long)o * ci + i) * taps + (taps - 1 - tap)];
      } break;
      case 4: {  // out[o][tap*ci + i], K padded
        if (k < taps * ci) {
          const int tap = k / ci, i = k % ci;
          v = w[((long long)r * ci + i) * taps + tap];
        }
      } break;
      case 5: {  // out[tap*co + o][i], rows padded
        if (r < taps * co) {
          const int tap = r / co, o = r % co;
          v = w[((long long)o * ci + k) * taps + tap];
        }
      } break;
      case 6: {  // out[tap*ci + i][o], rows padded
        if (r < taps * ci) {
          const int tap = r / ci, i = r % ci;
          v = w[((long long)k * ci + i) * taps + tap];
        }
      } break;
      case 7: {  // out[i][tap*co + o], K padded
        if (k < taps * co) {
          const int tap = k / co, o = k % co;
          v = w[((long long)o * ci + r) * taps + tap];
        }
      } break;
    }
    out[idx] = from_f32(v, bf);
  }
}

// Multi-tensor variant: `jobs` rows = {src, dst, co, ci, taps, mode, pad_to, first_block}; blockIdx.x is
// mapped to its job by binary search over first_block; each block packs 4096 output elements.
struct PackJob {
  const float* src;
  uint16_t* dst;
  uint16_t* dst2;     // tiled jobs: backward-data layout (mode + 1) written from the same read, or nullptr
  int co, ci, taps, mode, pad_to, rows, kdim;
  int first_block;
  int tiled;          // 1: 3x3, mode 0 / 2, co % 64 == 0, ci % 32 == 0 -> one block per 64 (out) x 32 (in) x 9 tile
};
__device__ __forceinline__ float pack_fetch(const float* __restrict__ w, int co, int ci, int taps, int mode, int r, int k);
__global__ void __launch_bounds__(256)
pack_weights_multi_kernel(const PackJob* __restrict__ jobs, int njobs, int bf) {
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].first_block <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const PackJob j = jobs[lo];
  if (j.tiled) {
    // OIHW reads with one thread per packed element are strided by 9 floats (forward layout) or by 9*ci floats
    // (backward layout): 1.4 TB/s.  Here a block reads 64 x (32 x 9) contiguous floats, keeps the tile in shared
    // memory as 16-bit values and writes BOTH operand layouts in 64 / 128-byte runs.
    // [tap][out][in] as 16-bit values; row stride 34 and plane stride 2178 halves keep the transposed accesses below
    // off the same bank (a plane stride of 64*34 halves is a multiple of 32 words: 9-way conflicts on the fill)
    constexpr int kRow = 34, kPlane = 64 * 34 + 2;
    __shared__ __align__(16) uint16_t tile[9 * kPlane];
    const int tiles_i = j.ci >> 5;
    const int tb = (int)blockIdx.x - j.first_block;
    const int op0 = (tb / tiles_i) * 64, i0 = (tb % tiles_i) * 32;
    const int c_ps = j.co >> 2;
    // one warp per OIHW row (288 contiguous floats), 8 rows per warp: all 72 loads of a thread are issued before the first
    // shared-memory store (the fill was latency-bound: one 128-byte request in flight per warp, 1.8 TB/s for the launch)
    {
      const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
      float vals[8][9];
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const int op = op0 + wrp + 8 * r;
        const int o = j.mode == 2 ? (op % c_ps) * 4 + op / c_ps : op;     // packed PixelShuffle order -> OIHW row
        const float* row = j.src + ((long long)o * j.ci + i0) * 9;
#pragma unroll
        for (int q = 0; q < 9; q++) vals[r][q] = __ldg(row + lane + 32 * q);
      }
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const int op_l = wrp + 8 * r;
#pragma unroll
        for (int q = 0; q < 9; q++) {
          const int rem = lane + 32 * q;
          const int i_l = rem / 9, t = rem - i_l * 9;
          tile[t * kPlane + op_l * kRow + i_l] = from_f32(vals[r][q], bf);
        }
      }
    }
    __syncthreads();
    uint32_t* dst32 = reinterpret_cast<uint32_t*>(j.dst);
    for (int idx = threadIdx.x; idx < 9 * 64 * 16; idx += 256) {          // forward: [tap][out][in], in fastest (pairs)
      const int ip = idx & 15, op_l = (idx >> 4) & 63, t = idx >> 10;
      const uint32_t v = *reinterpret_cast<const uint32_t*>(&tile[t * kPlane + op_l * kRow + 2 * ip]);
      dst32[(((long long)t * j.co + op0 + op_l) * j.ci + i0) / 2 + ip] = v;
    }
    if (j.dst2) {
      uint32_t* d2 = reinterpret_cast<uint32_t*>(j.dst2);
      for (int idx = threadIdx.x; idx < 9 * 32 * 32; idx += 256) {        // backward: [8 - tap][in][out], out fastest (pairs)
        const int opp = idx & 31, i_l = (idx >> 5) & 31, t = idx >> 10;
        const uint32_t lo = tile[t * kPlane + (2 * opp) * kRow + i_l], hi = tile[t * kPlane + (2 * opp + 1) * kRow + i_l];
        d2[(((long long)(8 - t) * j.ci + i0 + i_l) * j.co + op0) / 2 + opp] = lo | (hi << 16);
      }
    }
    return;
  }
  const long long total = (long long)j.rows * j.kdim;
  const long long base = (long long)((int)blockIdx.x - j.first_block) * 4096;
  for (int t = threadIdx.x; t < 4096; t += blockDim.x) {
    const long long idx = base + t;
    if (idx >= total) break;
    const int k = (int)(idx % j.kdim), r = (int)(idx / j.kdim);
    j.dst[idx] = from_f32(pack_fetch(j.src, j.co, j.ci, j.taps, j.mode, r, k), bf);
  }
}
__device__ __forceinline__ float pack_fetch(const float* __restrict__ w, int co, int ci, int taps, int mode, int r, int k) {
  const int c_ps = co / 4;
  switch (mode) {
    case 0: { const int tap = r / co, o = r % co; return w[((long long)o * ci + k) * taps + tap]; }
    case 1: { const int tap = r / ci, i = r % ci; return w[((long long)k * ci + i) * taps + (taps - 1 - tap)]; }
    case 2: { const int tap = r / co, op = r % co; const int o = (op % c_ps) * 4 + (op / c_ps);
              return w[((long long)o * ci + k) * taps + tap]; }
    case 3: { const int tap = r / ci, i = r % ci; const int o = (k % c_ps) * 4 + (k / c_ps);
              return w[((long long)o * ci + i) * taps + (taps - 1 - tap)]; }
    case 4: { if (k < taps * ci) { const int tap = k / ci, i = k % ci; return w[((long long)r * ci + i) * taps + tap]; } return 0.f; }
    case 5: { if (r < taps * co) { const int tap = r / co, o = r % co; return w[((long long)o * ci + k) * taps + tap]; } return 0.f; }
    case 6: { if (r < taps * ci) { const int tap = r / ci, i = r % ci; return w[((long long)k * ci + i) * taps + tap]; } return 0.f; }
    case 7: { if (k < taps * co) { const int tap = k / co, o = k % co; return w[((long long)o * ci + r) * taps + tap]; } return 0.f; }
  }
  return 0.f;
}

// ------------------------------------------------------------------------------------------
// im2col for 3-channel images: 4 threads per pixel; thread `q` produces columns [8q, 8q+8) (taps
// 8q/3 .. (8q+7)/3, each tap loaded once and pushed through the 3x3 affine) and the zero chunk q+4.
// ------------------------------------------------------------------------------------------
__global__ void im2col3_kernel(const float* __restrict__ src, int nb, int h, int w, const float* __restrict__ aff_a,
                               const float* __restrict__ aff_b, const float* __restrict__ mul_dev, int sgn,
                               int pad_affine, int bf, uint4* __restrict__ col) {
  griddep_wait();   // PDL: see launch_pdl
  const long long total = (long long)nb * h * w * 4;
  float A[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  float B[3] = {0, 0, 0};
  if (aff_a)
    for (int i = 0; i < 9; i++) A[i] = __ldg(aff_a + i);
  if (aff_b)
    for (int i = 0; i < 3; i++) B[i] = __ldg(aff_b + i);
  const float mul = mul_dev ? __ldg(mul_dev) : 1.f;
  const long long plane = (long long)h * w;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(idx & 3);
    const long long p = idx >> 2;
    int x, y, n;
    if (total < (1ll << 31)) {      // 32-bit divisions: the 64-bit ones were a third of this kernel's instructions
      const unsigned pu = (unsigned)p, wu = (unsigned)w, r = pu / wu;
      x = (int)(pu - r * wu);
      n = (int)(r / (unsigned)h);
      y = (int)(r - (unsigned)n * (unsigned)h);
    } else {
      x = (int)(p % w);
      y = (int)((p / w) % h);
      n = (int)(p / plane);
    }
    const float* base = src + (long long)n * 3 * plane;
    const uint8_t* base8 = reinterpret_cast<const uint8_t*>(src) + (long long)n * 3 * plane;   // flag 4: uint8 HWC source
    float vals[8];
#pragma unroll
    for (int j = 0; j < 8; j++) vals[j] = 0.f;
    const int k0 = q * 8;
    const int t_lo = k0 / 3, t_hi = min(8, (k0 + 7) / 3);
    for (int tap = t_lo; tap <= t_hi; tap++) {
      const int yy = y + sgn * (tap / 3 - 1), xx = x + sgn * (tap % 3 - 1);
      float v0 = 0.f, v1 = 0.f, v2 = 0.f;
      if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
        const long long off = (long long)yy * w + xx;
        float s0, s1, s2;
        if (pad_affine & 4) {
          const uint8_t* q8 = base8 + off * 3;
          s0 = (float)q8[0]; s1 = (float)q8[1]; s2 = (float)q8[2];
        } else {
          s0 = __ldg(base + off); s1 = __ldg(base + plane + off); s2 = __ldg(base + 2 * plane + off);
        }
        v0 = (A[0] * s0 + A[1] * s1 + A[2] * s2 + B[0]) * mul;
        v1 = (A[3] * s0 + A[4] * s1 + A[5] * s2 + B[1]) * mul;
        v2 = (A[6] * s0 + A[7] * s1 + A[8] * s2 + B[2]) * mul;
      } else if (pad_affine & 1) {  // the affine of a zero-padded pixel (a constant shift of the conv input)
        v0 = B[0] * mul; v1 = B[1] * mul; v2 = B[2] * mul;
      }
      const int kb = tap * 3 - k0;  // column of channel 0 of this tap, relative to the chunk
#pragma unroll
      for (int j = 0; j < 8; j++) {
        if (j == kb) vals[j] = v0;
        if (j == kb + 1) vals[j] = v1;
        if (j == kb + 2) vals[j] = v2;
      }
    }
    if (q == 3) { vals[3] = vals[4] = vals[5] = vals[6] = vals[7] = 0.f; }  // columns 27..31
    uint4 o;
    o.x = pack2(vals[0], vals[1], bf);
    o.y = pack2(vals[2], vals[3], bf);
    o.z = pack2(vals[4], vals[5], bf);
    o.w = pack2(vals[6], vals[7], bf);
    col[p * 8 + q] = o;
    if (!(pad_affine & 2)) col[p * 8 + 4 + q] = make_uint4(0, 0, 0, 0);   // bit 1: columns 32..63 are already zero
  }
}


// Same result through a shared-memory tile: one thread per PIXEL gathers its 27 values (adjacent threads read adjacent
// pixels of every tap / channel plane: coalesced), writes its 128-byte im2col row into an XOR-swizzled 16 KB tile, and
// the block then streams the tile out as contiguous 16-byte pieces.  (The 4-threads-per-pixel kernel above spends its
// time on per-thread tap selection and address arithmetic: 1.4 TB/s on the 75 MB im2col matrix of a 16 x 192 x 192 batch.)
__global__ void __launch_bounds__(128)
im2col3_tile_kernel(const float* __restrict__ src, int nb, int h, int w, const float* __restrict__ aff_a,
                    const float* __restrict__ aff_b, const float* __restrict__ mul_dev, int sgn, int flags, int bf,
                    uint4* __restrict__ col) {
  griddep_wait();   // PDL: see launch_pdl
  __shared__ uint4 tile[128 * 8];
  float A[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  float B[3] = {0, 0, 0};
  if (aff_a)
    for (int i = 0; i < 9; i++) A[i] = __ldg(aff_a + i);
  if (aff_b)
    for (int i = 0; i < 3; i++) B[i] = __ldg(aff_b + i);
  const float mul = mul_dev ? __ldg(mul_dev) : 1.f;
  const long long plane = (long long)h * w;
  const long long total = (long long)nb * plane;
  const long long p0 = (long long)blockIdx.x * 128;
  const long long p = p0 + threadIdx.x;
  const bool u8 = (flags & 4) != 0, pad_affine = (flags & 1) != 0;
  if (p < total) {
    const long long r = p / w;
    const int x = (int)(p - r * w);
    const int n = (int)(r / h);
    const int y = (int)(r - (long long)n * h);
    const float* base = src + (long long)n * 3 * plane;
    const uint8_t* base8 = reinterpret_cast<const uint8_t*>(src) + (long long)n * 3 * plane;
    float v[32];
#pragma unroll
    for (int j = 27; j < 32; j++) v[j] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; tap++) {
      const int yy = y + sgn * (tap / 3 - 1), xx = x + sgn * (tap % 3 - 1);
      float v0 = 0.f, v1 = 0.f, v2 = 0.f;
      if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
        const long long off = (long long)yy * w + xx;
        float s0, s1, s2;
        if (u8) {
          const uint8_t* q8 = base8 + off * 3;
          s0 = (float)q8[0]; s1 = (float)q8[1]; s2 = (float)q8[2];
        } else {
          s0 = __ldg(base + off); s1 = __ldg(base + plane + off); s2 = __ldg(base + 2 * plane + off);
        }
        v0 = (A[0] * s0 + A[1] * s1 + A[2] * s2 + B[0]) * mul;
        v1 = (A[3] * s0 + A[4] * s1 + A[5] * s2 + B[1]) * mul;
        v2 = (A[6] * s0 + A[7] * s1 + A[8] * s2 + B[2]) * mul;
      } else if (pad_affine) {  // the affine of a zero-padded pixel (a constant shift of the conv input)
        v0 = B[0] * mul; v1 = B[1] * mul; v2 = B[2] * mul;
      }
      v[tap * 3] = v0; v[tap * 3 + 1] = v1; v[tap * 3 + 2] = v2;
    }
    if (flags & 8) {        // split precision: emit the LOW part, v - float(round16(v)), of every column
#pragma unroll
      for (int j = 0; j < 27; j++) v[j] -= to_f32(from_f32(v[j], bf), bf);
    }
    const int row = threadIdx.x, sw = row & 7;
#pragma unroll
    for (int c = 0; c < 4; c++)
      tile[row * 8 + (c ^ sw)] = make_uint4(pack2(v[8 * c], v[8 * c + 1], bf), pack2(v[8 * c + 2], v[8 * c + 3], bf),
                                            pack2(v[8 * c + 4], v[8 * c + 5], bf), pack2(v[8 * c + 6], v[8 * c + 7], bf));
#pragma unroll
    for (int c = 4; c < 8; c++) tile[row * 8 + (c ^ sw)] = make_uint4(0, 0, 0, 0);    // columns 32..63
  }
  __syncthreads();
  const long long rows_here = total - p0 < 128 ? total - p0 : 128;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int piece = i * 128 + threadIdx.x;
    const int row = piece >> 3, c = piece & 7;
    if (row < rows_here) col[p0 * 8 + piece] = tile[row * 8 + (c ^ (row & 7))];
  }
}

// ------------------------------------------------------------------------------------------
// col2im for 3-channel outputs: one thread per pixel, writes 3 NCHW planes
// ------------------------------------------------------------------------------------------
__global__ void col2im3_kernel(const float* __restrict__ z, int ldz, int nb, int h, int w,
                               const float* __restrict__ bias, const float* __restrict__ aff_a,
                               const float* __restrict__ aff_b, float mul, const float* __restrict__ div_dev, int sgn,
                               float* __restrict__ pre, float* __restrict__ out) {
  griddep_wait();   // PDL: see launch_pdl
  const long long total = (long long)nb * h * w;
  const long long plane = (long long)h * w;
  float A[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  float B[3] = {0, 0, 0};
  if (aff_a)
    for (int i = 0; i < 9; i++) A[i] = __ldg(aff_a + i);
  if (aff_b)
    for (int i = 0; i < 3; i++) B[i] = __ldg(aff_b + i);
  if (div_dev) mul /= __ldg(div_dev);
  float b0 = 0, b1 = 0, b2 = 0;
  if (bias) { b0 = __ldg(bias); b1 = __ldg(bias + 1); b2 = __ldg(bias + 2); }
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total;
       p += (long long)gridDim.x * blockDim.x) {
    int x, y, n;
    if (total < (1ll << 31)) {
      const unsigned pu = (unsigned)p, wu = (unsigned)w, r = pu / wu;
      x = (int)(pu - r * wu);
      n = (int)(r / (unsigned)h);
      y = (int)(r - (unsigned)n * (unsigned)h);
    } else {
      x = (int)(p % w);
      y = (int)((p / w) % h);
      n = (int)(p / plane);
    }
    float s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
    for (int tap = 0; tap < 9; tap++) {
      const int yy = y + sgn * (tap / 3 - 1), xx = x + sgn * (tap % 3 - 1);
      if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
        const float* zp = z + (((long long)n * h + yy) * w + xx) * ldz + tap * 3;
        s0 += __ldg(zp); s1 += __ldg(zp + 1); s2 += __ldg(zp + 2);
      }
    }
    s0 = s0 * mul + b0; s1 = s1 * mul + b1; s2 = s2 * mul + b2;
    const long long o = (long long)n * 3 * plane + (long long)y * w + x;
    if (pre) { pre[o] = s0; pre[o + plane] = s1; pre[o + 2 * plane] = s2; }
    out[o] = A[0] * s0 + A[1] * s1 + A[2] * s2 + B[0];
    out[o + plane] = A[3] * s0 + A[4] * s1 + A[5] * s2 + B[1];
    out[o + 2 * plane] = A[6] * s0 + A[7] * s1 + A[8] * s2 + B[2];
  }
}

// ------------------------------------------------------------------------------------------
// NCHW fp32 <-> NHWC 16-bit through a 32x32 smem transpose (coalesced on both sides)
// ------------------------------------------------------------------------------------------
__global__ void nchw32_to_nhwc16_kernel(const float* __restrict__ src, int c, long long hw, int ldc,
                                        const float* __restrict__ mul_dev, int bf, uint16_t* __restrict__ dst) {
  griddep_wait();   // PDL: see launch_pdl
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const float mul = mul_dev ? __ldg(mul_dev) : 1.f;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int cc = c0 + i;
    const long long pp = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (cc < c && pp < hw) ? src[((long long)n * c + cc) * hw + pp] * mul : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long pp = p0 + i;
    const int cc = c0 + threadIdx.x;
    if (pp < hw && cc < ldc) dst[((long long)n * hw + pp) * ldc + cc] = from_f32(cc < c ? tile[threadIdx.x][i] : 0.f, bf);
  }
}

__global__ void nhwc16_to_nchw32_kernel(const uint16_t* __restrict__ src, int c, long long hw, int ldc, float mul,
                                        const float* __restrict__ div_dev, int bf, float* __restrict__ dst) {
  griddep_wait();   // PDL: see launch_pdl
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  if (div_dev) mul /= __ldg(div_dev);
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long pp = p0 + i;
    const int cc = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (pp < hw && cc < c) ? to_f32(src[((long long)n * hw + pp) * ldc + cc], bf) * mul : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int cc = c0 + i;
    const long long pp = p0 + threadIdx.x;
    if (cc < c && pp < hw) dst[((long long)n * c + cc) * hw + pp] = tile[threadIdx.x][i];
  }
}

// ------------------------------------------------------------------------------------------
// column sums of a 16-bit [npix][ldc] matrix (bias gradients)
// block = 256 threads: 32 channel-pair lanes x 8 pixel rows; grid.y tiles the channels by 64
// ------------------------------------------------------------------------------------------
__global__ void colsum16_kernel(const uint16_t* __restrict__ x, long long npix, int c, int ldc, float mul,
                                const float* __restrict__ div_dev, int bf, float* __restrict__ out) {
  griddep_wait();   // PDL: see launch_pdl
  __shared__ float red[8][64];
  const int lane = threadIdx.x & 31;
  const int rowi = threadIdx.x >> 5;
  const int c0 = blockIdx.y * 64 + lane * 2;
  float s0 = 0.f, s1 = 0.f;
  if (c0 < c) {
    for (long long p = (long long)blockIdx.x * 8 + rowi; p < npix; p += (long long)gridDim.x * 8) {
      const uint32_t u = *reinterpret_cast<const uint32_t*>(x + p * ldc + c0);
      const float2 f = unpack2(u, bf);
      s0 += f.x;
      s1 += f.y;
    }
  }
  red[rowi][lane * 2] = s0;
  red[rowi][lane * 2 + 1] = s1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
    for (int i = 0; i < 8; i++) s += red[i][threadIdx.x];
    const int cc = blockIdx.y * 64 + threadIdx.x;
    if (div_dev) mul /= __ldg(div_dev);
    if (cc < c) atomicAdd(out + cc, s * mul);
  }
}

// Vectorised variant: 16-byte loads; a block covers 256/(c/8) pixel rows per pass (c/8 must divide 256).
__global__ void __launch_bounds__(256)
colsum16_vec_kernel(const uint4* __restrict__ x, long long npix, int c, int ldv, float mul,
                    const float* __restrict__ div_dev, int bf, float* __restrict__ out) {
  griddep_wait();   // PDL: see launch_pdl
  __shared__ float red[2048];       // [row-in-pass][channel], rows-per-pass * c == 2048
  const int tpr = c >> 3, rpp = 256 / tpr;
  const int v = threadIdx.x % tpr, r = threadIdx.x / tpr;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; j++) acc[j] = 0.f;
  for (long long p = (long long)blockIdx.x * rpp + r; p < npix; p += (long long)gridDim.x * rpp) {
    const uint4 u = x[p * ldv + v];
    const float2 a = unpack2(u.x, bf), b = unpack2(u.y, bf), cc = unpack2(u.z, bf), d = unpack2(u.w, bf);
    acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
    acc[4] += cc.x; acc[5] += cc.y; acc[6] += d.x; acc[7] += d.y;
  }
#pragma unroll
  for (int j = 0; j < 8; j++) red[r * c + v * 8 + j] = acc[j];
  __syncthreads();
  if (div_dev) mul /= __ldg(div_dev);
  for (int ch = threadIdx.x; ch < c; ch += 256) {
    float t = 0.f;
    for (int k = 0; k < rpp; k++) t += red[k * c + ch];
    atomicAdd(out + ch, t * mul);
  }
}

// ------------------------------------------------------------------------------------------
// dynamic gradient scale
// ------------------------------------------------------------------------------------------
__global__ void amax_kernel(const float* __restrict__ x, long long n, unsigned int* __restrict__ ws) {
  griddep_wait();   // PDL: see launch_pdl
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float sm[32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) atomicMax(ws, __float_as_uint(m));  // non-negative floats order like uints
  }
}
__global__ void amax_finalize_kernel(float* ws, float target) {
  griddep_wait();   // PDL: see launch_pdl
  const float m = __uint_as_float(*reinterpret_cast<unsigned int*>(ws));
  float scale = 1.f;
  if (m > 0.f && isfinite(m)) {
    int e;
    frexpf(m, &e);  // m = f * 2^e, f in [0.5, 1)
    int et;
    frexpf(target, &et);
    scale = ldexpf(1.f, et - e);  // m*scale in [target/2.., target)
  }
  ws[1] = scale;
  ws[2] = 1.f / scale;
  *reinterpret_cast<unsigned int*>(ws) = 0u;  // re-arm for the next use
}

// ------------------------------------------------------------------------------------------
// MeanShift parameter gradients: 9 cross moments + 3 sums
// ------------------------------------------------------------------------------------------
__global__ void moments3_kernel(const float* __restrict__ a, const float* __restrict__ b, int nb, long long hw,
                                float* __restrict__ sums) {
  griddep_wait();   // PDL: see launch_pdl
  float acc[12];
#pragma unroll
  for (int i = 0; i < 12; i++) acc[i] = 0.f;
  const long long total = (long long)nb * hw;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long n = idx / hw, p = idx % hw;
    const long long o = n * 3 * hw + p;
    const float a0 = a[o], a1 = a[o + hw], a2 = a[o + 2 * hw];
    const float b0 = b[o], b1 = b[o + hw], b2 = b[o + 2 * hw];
    acc[0] += a0 * b0; acc[1] += a0 * b1; acc[2] += a0 * b2;
    acc[3] += a1 * b0; acc[4] += a1 * b1; acc[5] += a1 * b2;
    acc[6] += a2 * b0; acc[7] += a2 * b1; acc[8] += a2 * b2;
    acc[9] += a0; acc[10] += a1; acc[11] += a2;
  }
  __shared__ float sm[12][8];
#pragma unroll
  for (int i = 0; i < 12; i++) {
    float v = acc[i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[i][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    float v = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) v += sm[threadIdx.x][i];
    atomicAdd(sums + threadIdx.x, v);
  }
}

// ------------------------------------------------------------------------------------------
// inference edges (test.py:101-112, utils.py:13-25)
// ------------------------------------------------------------------------------------------
// out = alpha*perc + (1-alpha)*mean_i T_i^-1(ens_i), i = 0..7 (bit0 = flip W, bit1 = flip H, bit2 = transpose;
// inverses applied transpose -> flipH -> flipW, test.py:60-70); then clip(0,255).round() (half to even, like
// numpy) -> uint8 HWC.  ens holds the 8 generator outputs back to back, the transposed ones as [3][w][h].
__global__ void blend_x8_to_u8_kernel(const float* __restrict__ perc, const float* __restrict__ ens, int h, int w,
                                      float alpha, int n_ens, float* __restrict__ out32, uint8_t* __restrict__ out8) {
  const long long plane = (long long)h * w;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < plane;
       p += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(p % w), y = (int)(p / w);
#pragma unroll
    for (int c = 0; c < 3; c++) {
      float v = perc[c * plane + p];
      if (n_ens > 0) {
        float s = 0.f;
        for (int i = 0; i < n_ens; i++) {
          const int x1 = (i & 1) ? w - 1 - x : x;
          const int y1 = (i & 2) ? h - 1 - y : y;
          const long long idx = (i & 4) ? (long long)x1 * h + y1 : (long long)y1 * w + x1;
          const float e = ens[((long long)i * 3 + c) * plane + idx];
          s = i == 0 ? e : s + e;
        }
        v = alpha * v + (1.f - alpha) * (s / (float)n_ens);
      }
      if (out32) out32[c * plane + p] = v;
      if (out8) out8[p * 3 + c] = (uint8_t)rintf(fminf(fmaxf(v, 0.f), 255.f));
    }
  }
}

// uint8 HWC image -> fp32 NCHW tensor in 0..255 (utils.py:20-25)
__global__ void u8hwc_to_f32nchw_kernel(const uint8_t* __restrict__ src, int h, int w, float* __restrict__ dst) {
  const long long plane = (long long)h * w;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < plane;
       p += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; c++) dst[c * plane + p] = (float)src[p * 3 + c];
  }
}

}  // namespace pesr

using namespace pesr;

extern "C" int pesr_blend_x8_to_u8(const float* perc, const float* ens, int32_t h, int32_t w, float alpha,
                                   int32_t n_ens, float* out32, uint8_t* out8, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(perc && (out32 || out8) && h > 0 && w > 0, "blend_x8_to_u8: bad arguments");
  PESR_CHECK_ARG(n_ens == 0 || (n_ens == 8 && ens), "blend_x8_to_u8: n_ens must be 0 or 8");
  blend_x8_to_u8_kernel<<<blocks_for((long long)h * w, 256), 256, 0, stream>>>(perc, ens, h, w, alpha, n_ens, out32, out8);
  count_launch();
  PESR_CHECK_LAUNCH("blend_x8_to_u8");
  return 0;
}

extern "C" int pesr_u8hwc_to_f32nchw(const uint8_t* src, int32_t h, int32_t w, float* dst, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(src && dst && h > 0 && w > 0, "u8hwc_to_f32nchw: bad arguments");
  u8hwc_to_f32nchw_kernel<<<blocks_for((long long)h * w, 256), 256, 0, stream>>>(src, h, w, dst);
  count_launch();
  PESR_CHECK_LAUNCH("u8hwc_to_f32nchw");
  return 0;
}

namespace pesr {}
using namespace pesr;

extern "C" int pesr_pack_weights(const float* w, int32_t co, int32_t ci, int32_t ksize, int32_t mode, int32_t pad_to,
                                 int32_t dtype, void* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(w && out, "pack_weights: null pointer");
  PESR_CHECK_ARG(ksize == 3 || (ksize == 1 && (mode == 0 || mode == 1)), "pack_weights: ksize %d mode %d", ksize, mode);
  PESR_CHECK_ARG(mode >= 0 && mode <= 7, "pack_weights: mode %d", mode);
  const int taps = ksize * ksize;
  int rows = 0, kdim = 0;
  switch (mode) {
    case 0: case 2: rows = taps * co; kdim = ci; break;
    case 1: case 3: rows = taps * ci; kdim = co; break;
    case 4: rows = co; kdim = pad_to; PESR_CHECK_ARG(pad_to >= taps * ci, "pack_weights: pad_to %d < %d", pad_to, taps * ci); break;
    case 5: rows = pad_to; kdim = ci; PESR_CHECK_ARG(pad_to >= taps * co, "pack_weights: pad_to %d < %d", pad_to, taps * co); break;
    case 6: rows = pad_to; kdim = co; PESR_CHECK_ARG(pad_to >= taps * ci, "pack_weights: pad_to %d < %d", pad_to, taps * ci); break;
    case 7: rows = ci; kdim = pad_to; PESR_CHECK_ARG(pad_to >= taps * co, "pack_weights: pad_to %d < %d", pad_to, taps * co); break;
  }
  if (mode == 2 || mode == 3) PESR_CHECK_ARG(co % 4 == 0, "pack_weights: pixel-shuffle needs co %% 4 == 0");
  const long long total = (long long)rows * kdim;
  pack_weights_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(w, co, ci, taps, mode, pad_to, dtype,
                                                                  reinterpret_cast<uint16_t*>(out), total, rows, kdim);
  count_launch();
  note_weight_write(stream);
  PESR_CHECK_LAUNCH("pack_weights");
  return 0;
}

// jobs_host: njobs rows of 8 int64 {src, dst, co, ci, ksize, mode, pad_to, dst2}; jobs_dev: device scratch of
// njobs * 64 bytes that the call fills (kept by the caller so that repeated calls can skip the upload).
extern "C" int pesr_pack_weights_multi(const int64_t* jobs_host, int32_t njobs, void* jobs_dev, int32_t upload,
                                       int32_t dtype, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(jobs_host && jobs_dev && njobs > 0 && njobs <= 4096, "pack_weights_multi: bad arguments");
  static_assert(sizeof(PackJob) <= 64, "PackJob must fit the 64-byte device slot");
  std::vector<PackJob> jobs(njobs);
  int nblocks = 0;
  for (int i = 0; i < njobs; i++) {
    const int64_t* r = jobs_host + i * 8;
    PackJob j;
    j.src = reinterpret_cast<const float*>(r[0]);
    j.dst = reinterpret_cast<uint16_t*>(r[1]);
    j.dst2 = reinterpret_cast<uint16_t*>(r[7]);
    j.co = (int)r[2]; j.ci = (int)r[3];
    const int ksize = (int)r[4];
    j.taps = ksize * ksize; j.mode = (int)r[5]; j.pad_to = (int)r[6];
    PESR_CHECK_ARG(j.mode >= 0 && j.mode <= 7 && (ksize == 3 || ksize == 1), "pack_weights_multi: job %d mode/ksize", i);
    switch (j.mode) {
      case 0: case 2: j.rows = j.taps * j.co; j.kdim = j.ci; break;
      case 1: case 3: j.rows = j.taps * j.ci; j.kdim = j.co; break;
      case 4: j.rows = j.co; j.kdim = j.pad_to; break;
      case 5: j.rows = j.pad_to; j.kdim = j.ci; break;
      case 6: j.rows = j.pad_to; j.kdim = j.co; break;
      default: j.rows = j.ci; j.kdim = j.pad_to; break;
    }
    j.tiled = (ksize == 3 && (j.mode == 0 || j.mode == 2) && j.co % 64 == 0 && j.ci % 32 == 0) ? 1 : 0;
    PESR_CHECK_ARG(j.dst2 == nullptr || j.tiled,
                   "pack_weights_multi: job %d: a companion backward layout needs a 3x3 mode-0/2 job with co %% 64 == 0, "
                   "ci %% 32 == 0", i);
    j.first_block = nblocks;
    nblocks += j.tiled ? (j.co / 64) * (j.ci / 32) : (int)(((long long)j.rows * j.kdim + 4095) / 4096);
    jobs[i] = j;
  }
  if (upload) {
    // pageable-memory copy: synchronous with respect to the host buffer, stream-ordered on the device
    std::vector<unsigned char> staged((size_t)njobs * 64, 0);
    for (int i = 0; i < njobs; i++) memcpy(staged.data() + (size_t)i * 64, &jobs[i], sizeof(PackJob));
    cudaError_t e = cudaMemcpyAsync(jobs_dev, staged.data(), staged.size(), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) { set_error("pack_weights_multi: upload failed: %s", cudaGetErrorString(e)); return (int)e; }
  }
  pack_weights_multi_kernel<<<nblocks, 256, 0, stream>>>(reinterpret_cast<const PackJob*>(jobs_dev), njobs, dtype);
  count_launch();
  note_weight_write(stream);
  PESR_CHECK_LAUNCH("pack_weights_multi");
  return 0;
}

extern "C" int pesr_im2col3(const float* src, int32_t nb, int32_t h, int32_t w, const float* aff_a,
                            const float* aff_b, const float* mul_dev, int32_t sgn, int32_t pad_affine, int32_t dtype,
                            void* col, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(src && col && nb > 0 && h > 0 && w > 0, "im2col3: bad arguments");
  PESR_CHECK_ARG(sgn == 1 || sgn == -1, "im2col3: sgn must be +-1");
  static int use_old = -1;        // A/B knob: PESR_IM2COL_OLD=1 selects the 4-threads-per-pixel kernel
  if (use_old < 0) {
    const char* e = getenv("PESR_IM2COL_OLD");
    use_old = (e && e[0] == '1') ? 1 : 0;
  }
  if (use_old || (pad_affine & 2)) {
    const long long total = (long long)nb * h * w * 4;
    launch_pdl(im2col3_kernel, blocks_for(total, 256, 148 * 32), 256, 0, stream, src, nb, h, w, aff_a, aff_b, mul_dev, sgn,
               pad_affine, dtype, reinterpret_cast<uint4*>(col));
  } else {
    const long long blocks = ((long long)nb * h * w + 127) / 128;
    PESR_CHECK_ARG(blocks < (1ll << 31), "im2col3: too many pixels");
    launch_pdl(im2col3_tile_kernel, (unsigned)blocks, 128, 0, stream, src, nb, h, w, aff_a, aff_b, mul_dev, sgn, pad_affine,
               dtype, reinterpret_cast<uint4*>(col));
  }
  count_launch();
  PESR_CHECK_LAUNCH("im2col3");
  return 0;
}

extern "C" int pesr_col2im3(const float* z, int32_t ldz, int32_t nb, int32_t h, int32_t w, const float* bias,
                            const float* aff_a, const float* aff_b, float mul_host, const float* div_dev, int32_t sgn,
                            float* pre, float* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(z && out && nb > 0 && h > 0 && w > 0 && ldz >= 27, "col2im3: bad arguments");
  PESR_CHECK_ARG(sgn == 1 || sgn == -1, "col2im3: sgn must be +-1");
  const long long total = (long long)nb * h * w;
  launch_pdl(col2im3_kernel, blocks_for(total, 256, 148 * 32), 256, 0, stream, z, ldz, nb, h, w, bias, aff_a, aff_b, mul_host,
                                                                      div_dev, sgn, pre, out);
  count_launch();
  PESR_CHECK_LAUNCH("col2im3");
  return 0;
}

extern "C" int pesr_nchw32_to_nhwc16(const float* src, int32_t nb, int32_t c, int32_t h, int32_t w, int32_t ldc,
                                     const float* mul_dev, int32_t dtype, void* dst, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(src && dst && nb > 0 && c > 0 && ldc >= c, "nchw32_to_nhwc16: bad arguments");
  PESR_CHECK_ARG(nb <= 65535, "nchw32_to_nhwc16: nb too large");
  const long long hw = (long long)h * w;
  dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((ldc + 31) / 32), (unsigned)nb);
  launch_pdl(nchw32_to_nhwc16_kernel, grid, dim3(32, 8), 0, stream, src, c, hw, ldc, mul_dev, dtype,
                                                           reinterpret_cast<uint16_t*>(dst));
  count_launch();
  PESR_CHECK_LAUNCH("nchw32_to_nhwc16");
  return 0;
}

extern "C" int pesr_nhwc16_to_nchw32(const void* src, int32_t nb, int32_t c, int32_t h, int32_t w, int32_t ldc,
                                     float mul_host, const float* div_dev, int32_t dtype, float* dst, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(src && dst && nb > 0 && c > 0 && ldc >= c, "nhwc16_to_nchw32: bad arguments");
  PESR_CHECK_ARG(nb <= 65535, "nhwc16_to_nchw32: nb too large");
  const long long hw = (long long)h * w;
  dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((c + 31) / 32), (unsigned)nb);
  launch_pdl(nhwc16_to_nchw32_kernel, grid, dim3(32, 8), 0, stream, reinterpret_cast<const uint16_t*>(src), c, hw, ldc,
                                                           mul_host, div_dev, dtype, dst);
  count_launch();
  PESR_CHECK_LAUNCH("nhwc16_to_nchw32");
  return 0;
}

extern "C" int pesr_colsum16(const void* x, int64_t npix, int32_t c, int32_t ldc, float mul_host,
                             const float* div_dev, int32_t accumulate, int32_t dtype, float* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(x && out && npix > 0 && c > 0 && c % 2 == 0 && ldc % 2 == 0, "colsum16: bad arguments");
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * c, stream);
    if (e != cudaSuccess) { set_error("colsum16: memset failed: %s", cudaGetErrorString(e)); return (int)e; }
  }
  const int tpr = c / 8;
  if (c % 8 == 0 && ldc % 8 == 0 && tpr <= 256 && 256 % tpr == 0 && ((uintptr_t)x % 16) == 0) {
    const int rpp = 256 / tpr;
    // measured (tools/perf_helpers.py): 16 rows per thread beats more, smaller blocks on the short inputs (the kernel
    // is launch/tail-latency-bound there); the 8-blocks-per-SM cap matters on the long ones (85 -> 68 us at 302 MB)
    long long bx = (npix + (long long)rpp * 16 - 1) / ((long long)rpp * 16);
    if (bx > 148 * 8) bx = 148 * 8;
    if (bx < 1) bx = 1;
    launch_pdl(colsum16_vec_kernel, (unsigned)bx, 256, 0, stream, reinterpret_cast<const uint4*>(x), npix, c, ldc / 8, mul_host,
                                                         div_dev, dtype, out);
  } else {
    long long bx = (npix + 8 * 64 - 1) / (8 * 64);
    if (bx > 1024) bx = 1024;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, (unsigned)((c + 63) / 64));
    launch_pdl(colsum16_kernel, grid, 256, 0, stream, reinterpret_cast<const uint16_t*>(x), npix, c, ldc, mul_host, div_dev,
                                             dtype, out);
  }
  count_launch();
  PESR_CHECK_LAUNCH("colsum16");
  return 0;
}

extern "C" int pesr_amax_scale(const float* x, int64_t n, float target, float* ws3, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(x && ws3 && n > 0 && target > 0.f, "amax_scale: bad arguments");
  launch_pdl(amax_kernel, blocks_for(n, 256, 148 * 4), 256, 0, stream, x, n, reinterpret_cast<unsigned int*>(ws3));
  launch_pdl(amax_finalize_kernel, 1, 1, 0, stream, ws3, target);
  count_launch(2);
  PESR_CHECK_LAUNCH("amax_scale");
  return 0;
}

extern "C" int pesr_moments3(const float* a, const float* b, int32_t nb, int64_t hw, float* sums12, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(a && b && sums12 && nb > 0 && hw > 0, "moments3: bad arguments");
  cudaError_t e = cudaMemsetAsync(sums12, 0, sizeof(float) * 12, stream);
  if (e != cudaSuccess) { set_error("moments3: memset failed: %s", cudaGetErrorString(e)); return (int)e; }
  launch_pdl(moments3_kernel, blocks_for((long long)nb * hw, 256, 148 * 4), 256, 0, stream, a, b, nb, hw, sums12);
  count_launch();
  PESR_CHECK_LAUNCH("moments3");
  return 0;
}
