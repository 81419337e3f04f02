// Image I/O edges, validation metric and input pipeline of the PESR path (HBM-bound, one coalesced pass each):
//  * MeanShift as a stand-alone op (model/basic.py:9-17, for callers that use the building block directly);
//  * col2im for the 3-channel output conv through a shared-memory tile, with the clip / round-half-even / uint8 HWC
//    store of utils.tensors_to_imgs (utils.py:13-18) fused for inference;
//  * Y-channel PSNR of utils.compute_PSNR (utils.py:10-11,27-41) as an exact integer sum of squared differences;
//  * random crop + flip / transpose augmentation of data.py:64-126 as one gather launch over a device-resident
//    uint8 image cache.
#include "common.cuh"
#include "host_util.cuh"

namespace pesr {

static inline int io_blocks(long long n, int threads, int cap = 148 * 16) {
  long long b = (n + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ------------------------------------------------------------------------------------------
// out[n][o][p] = sum_i w[o][i] * x[n][i][p] + b[o]      (1x1 conv on 3 channels, NCHW fp32)
// ------------------------------------------------------------------------------------------
__global__ void mean_shift_kernel(const float* __restrict__ x, int nb, long long hw, const float* __restrict__ w9,
                                  const float* __restrict__ b3, float* __restrict__ out) {
  griddep_wait();
  float A[9], B[3] = {0.f, 0.f, 0.f};
  for (int i = 0; i < 9; i++) A[i] = __ldg(w9 + i);
  if (b3)
    for (int i = 0; i < 3; i++) B[i] = __ldg(b3 + i);
  const long long total = (long long)nb * hw;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long n = idx / hw, p = idx - n * hw;
    const long long o = n * 3 * hw + p;
    const float s0 = x[o], s1 = x[o + hw], s2 = x[o + 2 * hw];
    out[o] = A[0] * s0 + A[1] * s1 + A[2] * s2 + B[0];
    out[o + hw] = A[3] * s0 + A[4] * s1 + A[5] * s2 + B[1];
    out[o + 2 * hw] = A[6] * s0 + A[7] * s1 + A[8] * s2 + B[2];
  }
}

// ------------------------------------------------------------------------------------------
// col2im for a 3-channel output through shared memory.  One block = an 8 x 32 pixel tile of one image; the 10 x 34
// halo of z rows (27 useful floats each) is staged once (rows padded to 33 floats: bank = row + column, so the
// compute phase, where the 32 lanes of a warp read the same column of 32 consecutive rows, is conflict-free), instead
// of every pixel fetching its 9 neighbours' rows through L1 (col2im3_kernel: 1.1 TB/s).
//   v[c]   = mul * sum_tap z[p + sgn*(ky-1,kx-1)][tap*3+c] + bias[c]
//   out    = A v + B      (fp32 NCHW, optional)            pre = v (optional)
//   out8   = uint8( rint( clamp(A v + B, 0, 255) ) )  HWC  (optional; rintf = round-half-even, like numpy)
// ------------------------------------------------------------------------------------------
static constexpr int kC2iTH = 8, kC2iTW = 32, kC2iRow = 33;

__global__ void __launch_bounds__(256)
col2im3_tiled_kernel(const float* __restrict__ z, int ldz, int nb, int h, int w, const float* __restrict__ bias,
                     const float* __restrict__ aff_a, const float* __restrict__ aff_b, float mul,
                     const float* __restrict__ div_dev, int sgn, float* __restrict__ pre, float* __restrict__ out,
                     uint8_t* __restrict__ out8) {
  griddep_wait();
  __shared__ float tile[(kC2iTH + 2) * (kC2iTW + 2) * kC2iRow];
  const int n = blockIdx.z;
  const int y0 = blockIdx.y * kC2iTH, x0 = blockIdx.x * kC2iTW;
  constexpr int HW = kC2iTW + 2, HP = (kC2iTH + 2) * HW;
  for (int idx = threadIdx.x; idx < HP * 8; idx += 256) {
    const int pix = idx >> 3, q = idx & 7;
    if (q == 7) continue;                      // floats 28..31 of a row are never read
    const int yy = y0 - 1 + pix / HW, xx = x0 - 1 + pix % HW;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (yy >= 0 && yy < h && xx >= 0 && xx < w)
      v = __ldg(reinterpret_cast<const float4*>(z + (((long long)n * h + yy) * w + xx) * ldz) + q);
    float* d = tile + pix * kC2iRow + q * 4;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  float A[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  float B[3] = {0, 0, 0};
  if (aff_a)
    for (int i = 0; i < 9; i++) A[i] = __ldg(aff_a + i);
  if (aff_b)
    for (int i = 0; i < 3; i++) B[i] = __ldg(aff_b + i);
  if (div_dev) mul /= __ldg(div_dev);
  float b0 = 0, b1 = 0, b2 = 0;
  if (bias) { b0 = __ldg(bias); b1 = __ldg(bias + 1); b2 = __ldg(bias + 2); }
  __syncthreads();
  const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
  const int y = y0 + ty, x = x0 + tx;
  if (y >= h || x >= w) return;
  float s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
  for (int tap = 0; tap < 9; tap++) {
    const int ry = ty + 1 + sgn * (tap / 3 - 1), rx = tx + 1 + sgn * (tap % 3 - 1);
    const float* zp = tile + (ry * HW + rx) * kC2iRow + tap * 3;
    s0 += zp[0]; s1 += zp[1]; s2 += zp[2];
  }
  s0 = s0 * mul + b0; s1 = s1 * mul + b1; s2 = s2 * mul + b2;
  const long long plane = (long long)h * w;
  const long long o = (long long)n * 3 * plane + (long long)y * w + x;
  if (pre) { pre[o] = s0; pre[o + plane] = s1; pre[o + 2 * plane] = s2; }
  const float r0 = A[0] * s0 + A[1] * s1 + A[2] * s2 + B[0];
  const float r1 = A[3] * s0 + A[4] * s1 + A[5] * s2 + B[1];
  const float r2 = A[6] * s0 + A[7] * s1 + A[8] * s2 + B[2];
  if (out) { out[o] = r0; out[o + plane] = r1; out[o + 2 * plane] = r2; }
  if (out8) {
    uint8_t* q = out8 + (((long long)n * h + y) * w + x) * 3;
    q[0] = (uint8_t)rintf(fminf(fmaxf(r0, 0.f), 255.f));
    q[1] = (uint8_t)rintf(fminf(fmaxf(r1, 0.f), 255.f));
    q[2] = (uint8_t)rintf(fminf(fmaxf(r2, 0.f), 255.f));
  }
}

// ------------------------------------------------------------------------------------------
// Y-channel PSNR (utils.py:10-11,27-41).  Per pixel: rgb = rint(clamp(v, 0, 255)) (tensors_to_imgs),
// y = (65.738 r + 129.057 g + 25.064 b) / 256 + 16, y = rint(clamp(y, 0, 255)); the squared difference of two such
// integers is accumulated EXACTLY in a 64-bit integer per image, so the result does not depend on the reduction order.
// y is evaluated in exact integer arithmetic (numerator 65738 r + 129057 g + 25064 b over 256000) with
// round-half-even; numpy evaluates the same expression in float64, which differs only on the 65 (of 2^24) colours
// that are exact rounding ties.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int y_of_rgb(float r, float g, float b) {
  const int ri = (int)rintf(fminf(fmaxf(r, 0.f), 255.f));
  const int gi = (int)rintf(fminf(fmaxf(g, 0.f), 255.f));
  const int bi = (int)rintf(fminf(fmaxf(b, 0.f), 255.f));
  const int num = 65738 * ri + 129057 * gi + 25064 * bi;     // y = num / 256000 + 16, num < 2^26
  int q = num / 256000;
  const int rem = num - q * 256000;
  if (rem > 128000 || (rem == 128000 && (q & 1))) q++;        // round half to even
  q += 16;
  return q > 255 ? 255 : q;
}

__global__ void psnr_y_sse_kernel(const float* __restrict__ a, const float* __restrict__ b, int nb, long long hw,
                                  unsigned long long* __restrict__ sse) {
  griddep_wait();
  const int n = blockIdx.y;
  const float* pa = a + (long long)n * 3 * hw;
  const float* pb = b + (long long)n * 3 * hw;
  unsigned long long acc = 0;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += (long long)gridDim.x * blockDim.x) {
    const int d = y_of_rgb(pa[p], pa[p + hw], pa[p + 2 * hw]) - y_of_rgb(pb[p], pb[p + hw], pb[p + 2 * hw]);
    acc += (unsigned long long)(d * d);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ unsigned long long sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += sm[i];
    atomicAdd(sse + n, t);     // integer addition: exact and order-independent
  }
}

// ------------------------------------------------------------------------------------------
// Training patches (data.py:64-126): sample s reads the LR crop [y, y+p) x [x, x+p) of image s and the aligned HR crop
// (scale * y, scale * x, size scale * p) and applies augmentation k (data.py:86-104: bit 2 transpose, then bit 1
// vertical flip, then bit 0 horizontal flip).  Images are uint8 HWC in device memory; the output is NCHW fp32 (0..255),
// exactly what `_to_tensor` + the DataLoader's collate produce.  table: per sample 8 int64 =
// {lr ptr, hr ptr, lr width, hr width, y, x, k, unused}.
// ------------------------------------------------------------------------------------------
__global__ void gather_patches_kernel(const long long* __restrict__ table, int nb, int p, int scale,
                                      float* __restrict__ lr, float* __restrict__ hr) {
  griddep_wait();
  const int s = blockIdx.y;
  const long long* row = table + (long long)s * 8;
  const uint8_t* lsrc = reinterpret_cast<const uint8_t*>(row[0]);
  const uint8_t* hsrc = reinterpret_cast<const uint8_t*>(row[1]);
  const int lw = (int)row[2], hwid = (int)row[3], y0 = (int)row[4], x0 = (int)row[5], k = (int)row[6];
  const int P = p * scale;
  const int n_lr = p * p, n_hr = P * P;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_lr + n_hr; idx += gridDim.x * blockDim.x) {
    const bool is_hr = idx >= n_lr;
    const int e = is_hr ? idx - n_lr : idx;
    const int side = is_hr ? P : p;
    int i = e / side, j = e - i * side;                  // output pixel
    int a = i, b = j;
    if (k & 1) b = side - 1 - b;                         // undo the horizontal flip
    if (k & 2) a = side - 1 - a;                         // undo the vertical flip
    if (k & 4) { const int t = a; a = b; b = t; }        // undo the transpose
    const uint8_t* src = is_hr ? hsrc + ((long long)(y0 * scale + a) * hwid + (x0 * scale + b)) * 3
                               : lsrc + ((long long)(y0 + a) * lw + (x0 + b)) * 3;
    float* dst = (is_hr ? hr + (long long)s * 3 * n_hr : lr + (long long)s * 3 * n_lr) + e;
    const int plane = is_hr ? n_hr : n_lr;
    dst[0] = (float)src[0];
    dst[plane] = (float)src[1];
    dst[2 * plane] = (float)src[2];
  }
}

// ------------------------------------------------------------------------------------------
// Split-precision operands: v = act(src * mul) [* act'(mask_hi + mask_lo)]; hi = round16(v); lo = round16(v - hi).
// hi + lo carries 22 (fp16) / 16 (bf16) significant bits; three tensor-core passes (hi*hi + lo*hi + hi*lo) then
// reproduce an fp32-grade product (pesr_b200/engine_g_split.py).
// ------------------------------------------------------------------------------------------
__global__ void split16_kernel(const float4* __restrict__ src, long long n4, int act, const uint2* __restrict__ mask_hi,
                               const uint2* __restrict__ mask_lo, int mask_mode, float mul,
                               const float* __restrict__ mul_dev, int bf, uint2* __restrict__ hi, uint2* __restrict__ lo) {
  griddep_wait();
  if (mul_dev) mul *= __ldg(mul_dev);
  const float neg = mask_mode == 2 ? 0.2f : 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 s = src[i];
    float v[4] = {s.x * mul, s.y * mul, s.z * mul, s.w * mul};
    if (act == PESR_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 4; j++) v[j] = fmaxf(v[j], 0.f);
    } else if (act == PESR_ACT_LRELU) {
#pragma unroll
      for (int j = 0; j < 4; j++) v[j] = v[j] > 0.f ? v[j] : 0.2f * v[j];
    }
    if (mask_hi) {
      const uint2 mh = mask_hi[i];
      uint2 ml = make_uint2(0, 0);
      if (mask_lo) ml = mask_lo[i];
      const float2 a0 = unpack2(mh.x, bf), a1 = unpack2(mh.y, bf), b0 = unpack2(ml.x, bf), b1 = unpack2(ml.y, bf);
      const float m[4] = {a0.x + b0.x, a0.y + b0.y, a1.x + b1.x, a1.y + b1.y};
#pragma unroll
      for (int j = 0; j < 4; j++) v[j] *= m[j] > 0.f ? 1.f : neg;
    }
    const uint32_t h0 = pack2(v[0], v[1], bf), h1 = pack2(v[2], v[3], bf);
    if (hi) hi[i] = make_uint2(h0, h1);
    if (lo) {
      const float2 r0 = unpack2(h0, bf), r1 = unpack2(h1, bf);
      lo[i] = make_uint2(pack2(v[0] - r0.x, v[1] - r0.y, bf), pack2(v[2] - r1.x, v[3] - r1.y, bf));
    }
  }
}

// ------------------------------------------------------------------------------------------
// fp32 NHWC helpers of the split-precision schedules (engine_d_split.py / engine_v_split.py): activations between the
// three-pass convolutions are fp32, so BatchNorm / LeakyReLU / ReLU / max-pool act on fp32 and only the conv OPERANDS are
// split into hi + lo.
//   colmoments32:  sums[0][ch] += sum_p a[p][ch],  sums[1][ch] += sum_p a[p][ch] * (b ? b[p][ch] : a[p][ch])   (fp64)
//   affine_split:  v = ka[ch]*a + kb[ch]*b + kc[ch]  (kb / b optional; ka / kc optional = 1 / 0), v = act(v),
//                  v *= act'(mask_hi + mask_lo or mask32), outputs: out32 = v, hi = round16(v), lo = round16(v - hi)
//   maxpool2_f32:  2x2 / 2 max-pool and its backward (first maximum in scan order, optional relu' mask), NHWC fp32
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colmoments32_kernel(const float* __restrict__ a, const float* __restrict__ b, long long npix, int c, double* __restrict__ sums) {
  griddep_wait();
  __shared__ double red[2][8][32];
  const int lane = threadIdx.x & 31, rowi = threadIdx.x >> 5;
  const int ch = blockIdx.y * 32 + lane;
  double s0 = 0, s1 = 0;
  if (ch < c) {
    for (long long p = (long long)blockIdx.x * 8 + rowi; p < npix; p += (long long)gridDim.x * 8) {
      const float x = a[p * c + ch];
      const float y = b ? b[p * c + ch] : x;
      s0 += (double)x;
      s1 += (double)x * (double)y;
    }
  }
  red[0][rowi][lane] = s0;
  red[1][rowi][lane] = s1;
  __syncthreads();
  if (threadIdx.x < 64) {
    const int which = threadIdx.x >> 5, l = threadIdx.x & 31;
    double t = 0;
    for (int i = 0; i < 8; i++) t += red[which][i][l];
    const int cc = blockIdx.y * 32 + l;
    if (cc < c) atomicAdd(sums + (long long)which * c + cc, t);
  }
}

__global__ void affine_split_kernel(const float4* __restrict__ a, const float4* __restrict__ b, long long n4, int c4,
                                    const float4* __restrict__ ka, const float4* __restrict__ kb, const float4* __restrict__ kc,
                                    int act, const uint2* __restrict__ mask_hi, const uint2* __restrict__ mask_lo,
                                    const float4* __restrict__ mask32, int mask_mode, int bf, float4* __restrict__ out32,
                                    uint2* __restrict__ hi, uint2* __restrict__ lo) {
  griddep_wait();
  const float neg = mask_mode == 2 ? 0.2f : 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int cc = (int)(i % c4);
    const float4 s = a[i];
    float v[4] = {s.x, s.y, s.z, s.w};
    if (ka) { const float4 k = __ldg(ka + cc); v[0] *= k.x; v[1] *= k.y; v[2] *= k.z; v[3] *= k.w; }
    if (b && kb) {
      const float4 t = b[i], k = __ldg(kb + cc);
      v[0] += k.x * t.x; v[1] += k.y * t.y; v[2] += k.z * t.z; v[3] += k.w * t.w;
    }
    if (kc) { const float4 k = __ldg(kc + cc); v[0] += k.x; v[1] += k.y; v[2] += k.z; v[3] += k.w; }
    if (act == PESR_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 4; j++) v[j] = fmaxf(v[j], 0.f);
    } else if (act == PESR_ACT_LRELU) {
#pragma unroll
      for (int j = 0; j < 4; j++) v[j] = v[j] > 0.f ? v[j] : 0.2f * v[j];
    }
    if (mask_hi) {
      const uint2 mh = mask_hi[i];
      uint2 ml = make_uint2(0, 0);
      if (mask_lo) ml = mask_lo[i];
      const float2 a0 = unpack2(mh.x, bf), a1 = unpack2(mh.y, bf), b0 = unpack2(ml.x, bf), b1 = unpack2(ml.y, bf);
      const float m[4] = {a0.x + b0.x, a0.y + b0.y, a1.x + b1.x, a1.y + b1.y};
#pragma unroll
      for (int j = 0; j < 4; j++) v[j] *= m[j] > 0.f ? 1.f : neg;
    } else if (mask32) {
      const float4 m = mask32[i];
      v[0] *= m.x > 0.f ? 1.f : neg; v[1] *= m.y > 0.f ? 1.f : neg; v[2] *= m.z > 0.f ? 1.f : neg; v[3] *= m.w > 0.f ? 1.f : neg;
    }
    if (out32) out32[i] = make_float4(v[0], v[1], v[2], v[3]);
    if (hi || lo) {
      const uint32_t h0 = pack2(v[0], v[1], bf), h1 = pack2(v[2], v[3], bf);
      if (hi) hi[i] = make_uint2(h0, h1);
      if (lo) {
        const float2 r0 = unpack2(h0, bf), r1 = unpack2(h1, bf);
        lo[i] = make_uint2(pack2(v[0] - r0.x, v[1] - r0.y, bf), pack2(v[2] - r1.x, v[3] - r1.y, bf));
      }
    }
  }
}

__global__ void maxpool2_f32_fwd_kernel(const float4* __restrict__ x, int nb, int h, int w, int c4, float4* __restrict__ y) {
  griddep_wait();
  const int ho = h >> 1, wo = w >> 1;
  const long long total = (long long)nb * ho * wo * c4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % c4);
    long long r = i / c4;
    const int ox = (int)(r % wo); r /= wo;
    const int oy = (int)(r % ho);
    const int n = (int)(r / ho);
    const long long base = (((long long)n * h + 2 * oy) * w + 2 * ox) * c4 + v;
    const float4 q0 = x[base], q1 = x[base + c4], q2 = x[base + (long long)w * c4], q3 = x[base + (long long)w * c4 + c4];
    y[i] = make_float4(fmaxf(fmaxf(q0.x, q1.x), fmaxf(q2.x, q3.x)), fmaxf(fmaxf(q0.y, q1.y), fmaxf(q2.y, q3.y)),
                       fmaxf(fmaxf(q0.z, q1.z), fmaxf(q2.z, q3.z)), fmaxf(fmaxf(q0.w, q1.w), fmaxf(q2.w, q3.w)));
  }
}

__global__ void maxpool2_f32_bwd_kernel(const float4* __restrict__ x, const float4* __restrict__ dy, int nb, int h, int w, int c4,
                                        int relu_mask, float4* __restrict__ dx) {
  griddep_wait();
  const int ho = h >> 1, wo = w >> 1;
  const long long total = (long long)nb * ho * wo * c4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % c4);
    long long r = i / c4;
    const int ox = (int)(r % wo); r /= wo;
    const int oy = (int)(r % ho);
    const int n = (int)(r / ho);
    const long long base = (((long long)n * h + 2 * oy) * w + 2 * ox) * c4 + v;
    const long long offs[4] = {base, base + c4, base + (long long)w * c4, base + (long long)w * c4 + c4};
    float q[4][4], g[4], o[4][4];
    for (int k = 0; k < 4; k++) { const float4 t = x[offs[k]]; q[k][0] = t.x; q[k][1] = t.y; q[k][2] = t.z; q[k][3] = t.w; }
    { const float4 t = dy[i]; g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w; }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int am = 0;
#pragma unroll
      for (int k = 1; k < 4; k++) if (q[k][j] > q[am][j]) am = k;
#pragma unroll
      for (int k = 0; k < 4; k++) o[k][j] = (k == am && (!relu_mask || q[k][j] > 0.f)) ? g[j] : 0.f;
    }
    for (int k = 0; k < 4; k++) dx[offs[k]] = make_float4(o[k][0], o[k][1], o[k][2], o[k][3]);
  }
  if (((h & 1) || (w & 1)) && blockIdx.x == 0) {       // odd trailing row / column (floor pooling drops them)
    for (long long i = threadIdx.x; i < (long long)nb * h * w * c4; i += blockDim.x) {
      const long long pix = i / c4;
      const int xx = (int)(pix % w), yy = (int)((pix / w) % h);
      if (yy >= 2 * ho || xx >= 2 * wo) dx[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// uint8 HWC [nb][h][w][3] -> fp32 NCHW [nb][3][h][w]
__global__ void u8hwc_to_f32nchw_batch_kernel(const uint8_t* __restrict__ src, int nb, long long hw,
                                              float* __restrict__ dst) {
  const long long total = (long long)nb * hw;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long n = idx / hw, p = idx - n * hw;
    const uint8_t* s = src + idx * 3;
    float* d = dst + n * 3 * hw + p;
    d[0] = (float)s[0]; d[hw] = (float)s[1]; d[2 * hw] = (float)s[2];
  }
}

}  // namespace pesr

using namespace pesr;

extern "C" int pesr_mean_shift(const float* x, int32_t nb, int64_t hw, const float* w9, const float* b3, float* out,
                               void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(x && w9 && out && nb > 0 && hw > 0, "mean_shift: bad arguments");
  launch_pdl(mean_shift_kernel, io_blocks((long long)nb * hw, 256), 256, 0, stream, x, nb, (long long)hw, w9, b3, out);
  count_launch();
  PESR_CHECK_LAUNCH("mean_shift");
  return 0;
}

extern "C" int pesr_col2im3_tiled(const float* z, int32_t ldz, int32_t nb, int32_t h, int32_t w, const float* bias,
                                  const float* aff_a, const float* aff_b, float mul_host, const float* div_dev,
                                  int32_t sgn, float* pre, float* out, uint8_t* out_u8, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(z && (out || out_u8) && nb > 0 && h > 0 && w > 0 && ldz >= 28 && ldz % 4 == 0 &&
                     ((uintptr_t)z % 16) == 0, "col2im3_tiled: bad arguments (z must be 16-byte aligned, ldz %% 4 == 0)");
  PESR_CHECK_ARG(sgn == 1 || sgn == -1, "col2im3_tiled: sgn must be +-1");
  PESR_CHECK_ARG(nb <= 65535 && (h + kC2iTH - 1) / kC2iTH <= 65535, "col2im3_tiled: grid too large");
  dim3 grid((unsigned)((w + kC2iTW - 1) / kC2iTW), (unsigned)((h + kC2iTH - 1) / kC2iTH), (unsigned)nb);
  launch_pdl(col2im3_tiled_kernel, grid, 256, 0, stream, z, ldz, nb, h, w, bias, aff_a, aff_b, mul_host, div_dev, sgn, pre,
             out, out_u8);
  count_launch();
  PESR_CHECK_LAUNCH("col2im3_tiled");
  return 0;
}

extern "C" int pesr_psnr_y_sse(const float* a, const float* b, int32_t nb, int64_t hw, unsigned long long* sse,
                               void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(a && b && sse && nb > 0 && nb <= 65535 && hw > 0, "psnr_y_sse: bad arguments");
  dim3 grid((unsigned)io_blocks(hw, 256, 148 * 4), (unsigned)nb);
  launch_pdl(psnr_y_sse_kernel, grid, 256, 0, stream, a, b, nb, (long long)hw, sse);
  count_launch();
  PESR_CHECK_LAUNCH("psnr_y_sse");
  return 0;
}

extern "C" int pesr_gather_patches(const int64_t* table_dev, int32_t nb, int32_t patch, int32_t scale, float* lr,
                                   float* hr, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(table_dev && lr && hr && nb > 0 && nb <= 65535 && patch > 0 && scale > 0, "gather_patches: bad arguments");
  const long long per = (long long)patch * patch * (1 + (long long)scale * scale);
  dim3 grid((unsigned)io_blocks(per, 256, 64), (unsigned)nb);
  launch_pdl(gather_patches_kernel, grid, 256, 0, stream, reinterpret_cast<const long long*>(table_dev), nb, patch, scale, lr,
             hr);
  count_launch();
  PESR_CHECK_LAUNCH("gather_patches");
  return 0;
}

extern "C" int pesr_split16(const float* src, int64_t n, int32_t act, const void* mask_hi, const void* mask_lo,
                            int32_t mask_mode, float mul, const float* mul_dev, int32_t dtype, void* hi, void* lo,
                            void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(src && (hi || lo) && n > 0 && n % 4 == 0 && ((uintptr_t)src % 16) == 0, "split16: bad arguments (n %% 4, 16-byte alignment)");
  launch_pdl(split16_kernel, io_blocks(n / 4, 256), 256, 0, stream, reinterpret_cast<const float4*>(src), (long long)(n / 4), act,
             reinterpret_cast<const uint2*>(mask_hi), reinterpret_cast<const uint2*>(mask_lo), mask_mode, mul, mul_dev, dtype,
             reinterpret_cast<uint2*>(hi), reinterpret_cast<uint2*>(lo));
  count_launch();
  PESR_CHECK_LAUNCH("split16");
  return 0;
}

extern "C" int pesr_colmoments32(const float* a, const float* b, int64_t npix, int32_t c, double* sums, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(a && sums && npix > 0 && c > 0, "colmoments32: bad arguments");
  long long bx = (npix + 8 * 32 - 1) / (8 * 32);
  if (bx > 148 * 8) bx = 148 * 8;
  dim3 grid((unsigned)bx, (unsigned)((c + 31) / 32));
  launch_pdl(colmoments32_kernel, grid, 256, 0, stream, a, b, (long long)npix, c, sums);
  count_launch();
  PESR_CHECK_LAUNCH("colmoments32");
  return 0;
}

extern "C" int pesr_affine_split(const float* a, const float* b, int64_t npix, int32_t c, const float* ka, const float* kb,
                                 const float* kc, int32_t act, const void* mask_hi, const void* mask_lo, const float* mask32,
                                 int32_t mask_mode, int32_t dtype, float* out32, void* hi, void* lo, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(a && (out32 || hi || lo) && npix > 0 && c > 0 && c % 4 == 0 && ((uintptr_t)a % 16) == 0,
                 "affine_split: bad arguments (c %% 4 == 0, 16-byte alignment)");
  const long long n4 = (long long)npix * (c / 4);
  launch_pdl(affine_split_kernel, io_blocks(n4, 256), 256, 0, stream, reinterpret_cast<const float4*>(a),
             reinterpret_cast<const float4*>(b), n4, c / 4, reinterpret_cast<const float4*>(ka), reinterpret_cast<const float4*>(kb),
             reinterpret_cast<const float4*>(kc), act, reinterpret_cast<const uint2*>(mask_hi),
             reinterpret_cast<const uint2*>(mask_lo), reinterpret_cast<const float4*>(mask32), mask_mode, dtype,
             reinterpret_cast<float4*>(out32), reinterpret_cast<uint2*>(hi), reinterpret_cast<uint2*>(lo));
  count_launch();
  PESR_CHECK_LAUNCH("affine_split");
  return 0;
}

extern "C" int pesr_maxpool2_f32_fwd(const float* x, int32_t nb, int32_t h, int32_t w, int32_t c, float* y, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(x && y && nb > 0 && h >= 2 && w >= 2 && c % 4 == 0, "maxpool2_f32_fwd: bad arguments");
  const long long total = (long long)nb * (h / 2) * (w / 2) * (c / 4);
  launch_pdl(maxpool2_f32_fwd_kernel, io_blocks(total, 256), 256, 0, stream, reinterpret_cast<const float4*>(x), nb, h, w, c / 4,
             reinterpret_cast<float4*>(y));
  count_launch();
  PESR_CHECK_LAUNCH("maxpool2_f32_fwd");
  return 0;
}

extern "C" int pesr_maxpool2_f32_bwd(const float* x, const float* dy, int32_t nb, int32_t h, int32_t w, int32_t c,
                                     int32_t relu_mask, float* dx, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(x && dy && dx && nb > 0 && h >= 2 && w >= 2 && c % 4 == 0, "maxpool2_f32_bwd: bad arguments");
  const long long total = (long long)nb * (h / 2) * (w / 2) * (c / 4);
  launch_pdl(maxpool2_f32_bwd_kernel, io_blocks(total, 256), 256, 0, stream, reinterpret_cast<const float4*>(x),
             reinterpret_cast<const float4*>(dy), nb, h, w, c / 4, relu_mask, reinterpret_cast<float4*>(dx));
  count_launch();
  PESR_CHECK_LAUNCH("maxpool2_f32_bwd");
  return 0;
}

extern "C" int pesr_u8hwc_to_f32nchw_batch(const uint8_t* src, int32_t nb, int32_t h, int32_t w, float* dst,
                                           void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(src && dst && nb > 0 && h > 0 && w > 0, "u8hwc_to_f32nchw_batch: bad arguments");
  const long long hw = (long long)h * w;
  u8hwc_to_f32nchw_batch_kernel<<<io_blocks(nb * hw, 256), 256, 0, stream>>>(src, nb, hw, dst);
  count_launch();
  PESR_CHECK_LAUNCH("u8hwc_to_f32nchw_batch");
  return 0;
}
