// HBM-bound layer kernels of the Discriminator and the VGG extractor:
//  * train-mode BatchNorm2d + LeakyReLU(0.2), forward and backward (model/basic.py:29-30 as used by
//    model/pesr.py:53-66), on NHWC 16-bit activations with fp32/fp64 statistics;
//  * 2x2 max-pool forward/backward (torchvision vgg19.features, model/vgg.py:10);
//  * the skinny (batch <= 16 rows per pass) weight-streaming Linear layers of model/pesr.py:71-73.
#include <stdlib.h>

#include "common.cuh"
#include "host_util.cuh"

namespace pesr {

// ------------------------------------------------------------------------------------------
// BatchNorm statistics: per-channel sum and sum of squares of x[npix][c] (16-bit), accumulated in
// double (the pre-BN conv outputs have |mean| >> std on 0..255 images, so E[x^2]-E[x]^2 needs it).
// GROUPS: the tensor holds `groups` independent batches back to back ([groups][npix][c]), each normalised with its own
// statistics -- the Discriminator's two calls of one training phase (D(hr), D(sr): train.py:205-208, 237-238) run as
// one launch per layer this way.  Group g uses sums + g*2*c and mean / rstd + g*c.
// block = 256 threads = 32 channel-pairs x 8 pixel rows; grid = (pixel blocks, channel tiles of 64, groups).
// mode 0: sums[0..c) += sum x,           sums[c..2c) += sum x^2
// mode 1: sums[0..c) += sum dz,          sums[c..2c) += sum dz * xhat   (xhat = (y - mean) * rstd)
// ------------------------------------------------------------------------------------------
__global__ void bn_reduce_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ y, long long npix, int c,
                                 const float* __restrict__ mean, const float* __restrict__ rstd, int mode, int bf,
                                 double* __restrict__ sums) {
  griddep_wait();   // PDL: see launch_pdl
  __shared__ double red[2][8][64];
  const int grp = blockIdx.z;
  x += (long long)grp * npix * c;
  if (y) y += (long long)grp * npix * c;
  sums += (long long)grp * 2 * c;
  const int lane = threadIdx.x & 31;
  const int rowi = threadIdx.x >> 5;
  const int c0 = blockIdx.y * 64 + lane * 2;
  double a0 = 0, a1 = 0, b0 = 0, b1 = 0;
  if (c0 < c) {
    float m0 = 0, m1 = 0, r0 = 1, r1 = 1;
    if (mode == 1) { m0 = mean[grp * c + c0]; m1 = mean[grp * c + c0 + 1]; r0 = rstd[grp * c + c0]; r1 = rstd[grp * c + c0 + 1]; }
    float fa0 = 0, fa1 = 0, fb0 = 0, fb1 = 0;
    int cnt = 0;
    for (long long p = (long long)blockIdx.x * 8 + rowi; p < npix; p += (long long)gridDim.x * 8) {
      const float2 f = unpack2(*reinterpret_cast<const uint32_t*>(x + p * c + c0), bf);
      if (mode == 0) {
        fa0 += f.x; fa1 += f.y; fb0 += f.x * f.x; fb1 += f.y * f.y;
      } else {
        const float2 g = unpack2(*reinterpret_cast<const uint32_t*>(y + p * c + c0), bf);
        fa0 += f.x; fa1 += f.y;
        fb0 += f.x * (g.x - m0) * r0; fb1 += f.y * (g.y - m1) * r1;
      }
      if (++cnt == 32) {  // flush the fp32 partials into double every 32 elements
        a0 += fa0; a1 += fa1; b0 += fb0; b1 += fb1;
        fa0 = fa1 = fb0 = fb1 = 0.f; cnt = 0;
      }
    }
    a0 += fa0; a1 += fa1; b0 += fb0; b1 += fb1;
  }
  red[0][rowi][lane * 2] = a0; red[0][rowi][lane * 2 + 1] = a1;
  red[1][rowi][lane * 2] = b0; red[1][rowi][lane * 2 + 1] = b1;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, ch = threadIdx.x & 63;
    double s = 0;
    for (int i = 0; i < 8; i++) s += red[which][i][ch];
    const int cc = blockIdx.y * 64 + ch;
    if (cc < c) atomicAdd(sums + which * c + cc, s);
  }
}

// Vectorised variant for c in {64,128,256,512}: 16-byte loads, c/8 threads per pixel row; grid = (pixel blocks, groups).
__global__ void __launch_bounds__(256)
bn_reduce_vec_kernel(const uint4* __restrict__ x, const uint4* __restrict__ y, long long npix, int c,
                     const float* __restrict__ mean, const float* __restrict__ rstd, int mode, int bf,
                     double* __restrict__ sums) {
  griddep_wait();   // PDL: see launch_pdl
  __shared__ double red[2][2048];   // [quantity][row-in-pass * c + channel], rows-per-pass * c == 2048
  const int grp = blockIdx.y;
  const int tpr = c >> 3;           // threads per pixel row
  const int rpp = 256 / tpr;        // pixel rows per pass
  x += (long long)grp * npix * tpr;
  if (y) y += (long long)grp * npix * tpr;
  sums += (long long)grp * 2 * c;
  const int v = threadIdx.x % tpr, r = threadIdx.x / tpr;
  float m[8], rs[8];
#pragma unroll
  for (int j = 0; j < 8; j++) { m[j] = 0.f; rs[j] = 1.f; }
  if (mode == 1) {
#pragma unroll
    for (int j = 0; j < 8; j++) { m[j] = mean[grp * c + v * 8 + j]; rs[j] = rstd[grp * c + v * 8 + j]; }
  }
  double da[8], db[8];
  float fa[8], fb[8];
#pragma unroll
  for (int j = 0; j < 8; j++) { da[j] = db[j] = 0.0; fa[j] = fb[j] = 0.f; }
  int cnt = 0;
  // four pixel rows per iteration: the loads are issued together (memory-level parallelism: with one 16-byte load in
  // flight per thread the 296 blocks keep ~1.2 MB in flight, a fifth of what HBM3e needs)
  auto accum = [&](const uint4& u, const uint4& w) {
    const uint32_t ux[4] = {u.x, u.y, u.z, u.w};
    if (mode == 0) {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float2 f = unpack2(ux[j], bf);
        fa[2 * j] += f.x; fa[2 * j + 1] += f.y;
        fb[2 * j] += f.x * f.x; fb[2 * j + 1] += f.y * f.y;
      }
    } else {
      const uint32_t uy[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float2 f = unpack2(ux[j], bf), g = unpack2(uy[j], bf);
        fa[2 * j] += f.x; fa[2 * j + 1] += f.y;
        fb[2 * j] += f.x * (g.x - m[2 * j]) * rs[2 * j];
        fb[2 * j + 1] += f.y * (g.y - m[2 * j + 1]) * rs[2 * j + 1];
      }
    }
    if (++cnt == 64) {
#pragma unroll
      for (int j = 0; j < 8; j++) { da[j] += fa[j]; db[j] += fb[j]; fa[j] = fb[j] = 0.f; }
      cnt = 0;
    }
  };
  const long long step = (long long)gridDim.x * rpp;
  long long p = (long long)blockIdx.x * rpp + r;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  if (mode == 0) {
    for (; p + 7 * step < npix; p += 8 * step) {      // statistics pass: one operand, eight loads in flight
      uint4 u[8];
#pragma unroll
      for (int q = 0; q < 8; q++) u[q] = x[(p + q * step) * tpr + v];
#pragma unroll
      for (int q = 0; q < 8; q++) accum(u[q], zero4);
    }
  }
  for (; p + 3 * step < npix; p += 4 * step) {
    uint4 u[4], w[4];
#pragma unroll
    for (int q = 0; q < 4; q++) u[q] = x[(p + q * step) * tpr + v];
    if (mode == 1) {
#pragma unroll
      for (int q = 0; q < 4; q++) w[q] = y[(p + q * step) * tpr + v];
    } else {
#pragma unroll
      for (int q = 0; q < 4; q++) w[q] = zero4;
    }
#pragma unroll
    for (int q = 0; q < 4; q++) accum(u[q], w[q]);
  }
  for (; p < npix; p += step) accum(x[p * tpr + v], mode == 1 ? y[p * tpr + v] : zero4);
#pragma unroll
  for (int j = 0; j < 8; j++) {
    red[0][r * c + v * 8 + j] = da[j] + fa[j];
    red[1][r * c + v * 8 + j] = db[j] + fb[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * c; i += 256) {
    const int which = i / c, ch = i % c;
    double t = 0;
    for (int k = 0; k < rpp; k++) t += red[which][k * c + ch];
    atomicAdd(sums + which * c + ch, t);
  }
}

// mean / rstd of one channel from its sums (the arithmetic every consumer of the sums repeats identically)
__device__ __forceinline__ void bn_moments(const double* __restrict__ sums, int c, int i, double n, float eps, float& mean,
                                           float& rstd, double& var_out) {
  const double m = sums[i] / n;
  double var = sums[c + i] / n - m * m;
  if (var < 0) var = 0;
  mean = (float)m;
  rstd = (float)(1.0 / sqrt(var + (double)eps));
  var_out = var;
}

// a = lrelu(gamma * (y - mean) * rstd + beta) = lrelu(y * sc + sh), 8 channels (16 bytes) per thread; grid = (blocks, groups).
// sums != NULL (train mode): every block derives mean / rstd of its group from the statistics sums itself (no separate
// finalisation launch); block 0 of each group stores them for backward, and block (0, 0) updates the running statistics
// with momentum, group after group in call order (unbiased variance; run_shift is added to the batch mean).
// sums == NULL (eval mode): mean / rstd are inputs.
__global__ void bn_lrelu_fwd_kernel(const uint4* __restrict__ y, long long nvec, int c, const double* __restrict__ sums,
                                    double n, float eps, float momentum, float* __restrict__ mean, float* __restrict__ rstd,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    float* __restrict__ run_mean, float* __restrict__ run_var,
                                    long long* __restrict__ num_batches, const float* __restrict__ run_shift, float slope,
                                    int bf, uint4* __restrict__ a) {
  griddep_wait();   // PDL: see launch_pdl
  __shared__ float sc[512], sh[512];
  const int grp = blockIdx.y, groups = gridDim.y;
  y += (long long)grp * nvec;
  a += (long long)grp * nvec;
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    float m, r;
    if (sums) {
      double var;
      bn_moments(sums + (long long)grp * 2 * c, c, i, n, eps, m, r, var);
      if (blockIdx.x == 0) { mean[grp * c + i] = m; rstd[grp * c + i] = r; }
    } else {
      m = mean[grp * c + i];
      r = rstd[grp * c + i];
    }
    const float k = gamma[i] * r;
    sc[i] = k;
    sh[i] = beta[i] - m * k;
  }
  if (sums && run_mean && blockIdx.x == 0 && grp == 0) {
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
      float rm = run_mean[i], rv = run_var[i];
      for (int g = 0; g < groups; g++) {
        float m, r;
        double var;
        bn_moments(sums + (long long)g * 2 * c, c, i, n, eps, m, r, var);
        rm = (1.f - momentum) * rm + momentum * (m + (run_shift ? run_shift[i] : 0.f));
        const double unb = n > 1 ? var * n / (n - 1) : var;
        rv = (1.f - momentum) * rv + momentum * (float)unb;
      }
      run_mean[i] = rm;
      run_var[i] = rv;
    }
    if (threadIdx.x == 0 && num_batches) *num_batches += groups;
  }
  __syncthreads();
  const int cv = c >> 3;
  auto apply = [&](long long i, const uint4& u) {
    const int c0 = (int)(i % cv) * 8;
    const uint32_t in[4] = {u.x, u.y, u.z, u.w};
    uint32_t out[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float2 f = unpack2(in[j], bf);
      const int cc = c0 + 2 * j;
      float v0 = f.x * sc[cc] + sh[cc];
      float v1 = f.y * sc[cc + 1] + sh[cc + 1];
      v0 = v0 > 0.f ? v0 : slope * v0;
      v1 = v1 > 0.f ? v1 : slope * v1;
      out[j] = pack2(v0, v1, bf);
    }
    a[i] = make_uint4(out[0], out[1], out[2], out[3]);
  };
  // four independent 16-byte loads in flight per thread (one per thread kept ~4.8 MB in flight on the whole GPU, short of
  // what HBM3e needs: 0.63 of the copy bandwidth)
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < nvec; i += 4 * stride) {
    uint4 u[4];
#pragma unroll
    for (int q = 0; q < 4; q++) u[q] = y[i + q * stride];
#pragma unroll
    for (int q = 0; q < 4; q++) apply(i + q * stride, u[q]);
  }
  for (; i < nvec; i += stride) apply(i, y[i]);
}

// dy = gamma * rstd * (dz - s1/n - xhat * s2/n) = ka*dz + kb*y + kc per channel and group; block (0, 0) also emits
// dgamma = sum_groups s2 * gmul, dbeta = sum_groups s1 * gmul (the groups are passes of one optimiser step).
__global__ void bn_bwd_apply_kernel(const uint4* __restrict__ dz, const uint4* __restrict__ y, long long nvec, int c,
                                    double n, const float* __restrict__ mean, const float* __restrict__ rstd,
                                    const float* __restrict__ gamma, const double* __restrict__ sums, float gmul,
                                    const float* __restrict__ gdiv_dev, int bf, uint4* __restrict__ dy,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate) {
  griddep_wait();   // PDL: see launch_pdl
  __shared__ float ka[512], kb[512], kc[512];
  const int grp = blockIdx.y, groups = gridDim.y;
  dz += (long long)grp * nvec;
  y += (long long)grp * nvec;
  dy += (long long)grp * nvec;
  const int cv = c >> 3;
  if (blockIdx.x == 0 && grp == 0 && dgamma) {
    float mul = gmul;
    if (gdiv_dev) mul /= __ldg(gdiv_dev);
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
      double s1 = 0, s2 = 0;
      for (int g = 0; g < groups; g++) { s1 += sums[(long long)g * 2 * c + i]; s2 += sums[(long long)g * 2 * c + c + i]; }
      const float db = (float)s1 * mul, dgm = (float)s2 * mul;
      dbeta[i] = accumulate ? dbeta[i] + db : db;
      dgamma[i] = accumulate ? dgamma[i] + dgm : dgm;
    }
  }
  const double* sg = sums + (long long)grp * 2 * c;
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    const double k1 = sg[i] / n, k2 = sg[c + i] / n;
    const double g = gamma[i], r = rstd[grp * c + i], m = mean[grp * c + i];
    ka[i] = (float)(g * r);
    kb[i] = (float)(-g * r * r * k2);
    kc[i] = (float)(g * r * (m * r * k2 - k1));
  }
  __syncthreads();
  auto apply = [&](long long i, const uint4& ug, const uint4& uy) {
    const int c0 = (int)(i % cv) * 8;
    const uint32_t g_in[4] = {ug.x, ug.y, ug.z, ug.w}, y_in[4] = {uy.x, uy.y, uy.z, uy.w};
    uint32_t out[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float2 g = unpack2(g_in[j], bf), yy = unpack2(y_in[j], bf);
      const int cc = c0 + 2 * j;
      out[j] = pack2(ka[cc] * g.x + kb[cc] * yy.x + kc[cc], ka[cc + 1] * g.y + kb[cc + 1] * yy.y + kc[cc + 1], bf);
    }
    dy[i] = make_uint4(out[0], out[1], out[2], out[3]);
  };
  // two vector pairs (four 16-byte loads) in flight per thread, see bn_lrelu_fwd_kernel
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + stride < nvec; i += 2 * stride) {
    const uint4 g0 = dz[i], g1 = dz[i + stride], y0 = y[i], y1 = y[i + stride];
    apply(i, g0, y0);
    apply(i + stride, g1, y1);
  }
  if (i < nvec) apply(i, dz[i], y[i]);
}

// ------------------------------------------------------------------------------------------
// 2x2 / stride-2 max-pool on NHWC 16-bit, 8 channels per thread.
// backward: routes dy to the FIRST maximum in (kh, kw) scan order (PyTorch's tie rule) and multiplies by
// relu'(x) (x > 0), which is the mask of the ReLU that precedes every pool in VGG19.
// ------------------------------------------------------------------------------------------
__global__ void maxpool2_fwd_kernel(const uint4* __restrict__ x, int nb, int h, int w, int c, int bf,
                                    uint4* __restrict__ y) {
  griddep_wait();   // PDL: see launch_pdl
  const int cv = c >> 3, ho = h >> 1, wo = w >> 1;
  const long long total = (long long)nb * ho * wo * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    long long r = i / cv;
    const int ox = (int)(r % wo); r /= wo;
    const int oy = (int)(r % ho);
    const int n = (int)(r / ho);
    const long long base = (((long long)n * h + 2 * oy) * w + 2 * ox) * cv + v;
    const uint4 q[4] = {x[base], x[base + cv], x[base + (long long)w * cv], x[base + (long long)w * cv + cv]};
    uint32_t out[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t e[4] = {(&q[0].x)[j], (&q[1].x)[j], (&q[2].x)[j], (&q[3].x)[j]};
      float2 m = unpack2(e[0], bf);
#pragma unroll
      for (int k = 1; k < 4; k++) {
        const float2 f = unpack2(e[k], bf);
        m.x = fmaxf(m.x, f.x);
        m.y = fmaxf(m.y, f.y);
      }
      out[j] = pack2(m.x, m.y, bf);
    }
    y[i] = make_uint4(out[0], out[1], out[2], out[3]);
  }
}

__global__ void maxpool2_bwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, int nb, int h, int w,
                                    int c, int relu_mask, int bf, uint4* __restrict__ dx) {
  griddep_wait();   // PDL: see launch_pdl
  const int cv = c >> 3, ho = h >> 1, wo = w >> 1;
  const long long total = (long long)nb * ho * wo * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    long long r = i / cv;
    const int ox = (int)(r % wo); r /= wo;
    const int oy = (int)(r % ho);
    const int n = (int)(r / ho);
    const long long base = (((long long)n * h + 2 * oy) * w + 2 * ox) * cv + v;
    const long long offs[4] = {base, base + cv, base + (long long)w * cv, base + (long long)w * cv + cv};
    const uint4 q[4] = {x[offs[0]], x[offs[1]], x[offs[2]], x[offs[3]]};
    const uint4 g = dy[i];
    uint32_t o[4][4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float2 f[4];
#pragma unroll
      for (int k = 0; k < 4; k++) f[k] = unpack2((&q[k].x)[j], bf);
      const float2 gg = unpack2((&g.x)[j], bf);
      int ax = 0, ay = 0;
#pragma unroll
      for (int k = 1; k < 4; k++) {
        if (f[k].x > f[ax].x) ax = k;
        if (f[k].y > f[ay].y) ay = k;
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const float vx = (k == ax && (!relu_mask || f[k].x > 0.f)) ? gg.x : 0.f;
        const float vy = (k == ay && (!relu_mask || f[k].y > 0.f)) ? gg.y : 0.f;
        o[k][j] = pack2(vx, vy, bf);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) dx[offs[k]] = make_uint4(o[k][0], o[k][1], o[k][2], o[k][3]);
  }
  // odd trailing row / column (floor pooling drops them): their gradient is zero
  if (((h & 1) || (w & 1)) && blockIdx.x == 0) {
    for (long long i = threadIdx.x; i < (long long)nb * h * w * cv; i += blockDim.x) {
      const long long pix = i / cv;
      const int xx = (int)(pix % w), yy = (int)((pix / w) % h);
      if (yy >= 2 * ho || xx >= 2 * wo) dx[i] = make_uint4(0, 0, 0, 0);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Skinny Linear layers (batch rows processed 16 at a time; weights streamed once per pass).
// Operands: x16 [nb][k] and w16 [o][k] 16-bit (k % 8 == 0); fp32 accumulation.
// ------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(256)
linear_fwd_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ w, int nb, int k, int o, int ksplit,
                  int bf, float* __restrict__ part) {
  // block: 8 outputs x one k-split; warp: an interleaved slice of that split; lane: 8 consecutive k
  constexpr int OT = 8;
  __shared__ float red[8][NB * OT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o0 = blockIdx.x * OT;
  const int split = blockIdx.y;
  const int kchunk = (k / 8 + ksplit - 1) / ksplit * 8;
  const int k0 = split * kchunk, k1 = min(k, k0 + kchunk);
  float acc[NB][OT];
#pragma unroll
  for (int n = 0; n < NB; n++)
#pragma unroll
    for (int j = 0; j < OT; j++) acc[n][j] = 0.f;
  for (int kk = k0 + (warp * 32 + lane) * 8; kk < k1; kk += 256 * 8) {
    float wv[OT][8];
#pragma unroll
    for (int j = 0; j < OT; j++) {
      if (o0 + j < o) {
        const uint4 u = *reinterpret_cast<const uint4*>(w + (long long)(o0 + j) * k + kk);
        const float2 a = unpack2(u.x, bf), b = unpack2(u.y, bf), c = unpack2(u.z, bf), d = unpack2(u.w, bf);
        wv[j][0] = a.x; wv[j][1] = a.y; wv[j][2] = b.x; wv[j][3] = b.y;
        wv[j][4] = c.x; wv[j][5] = c.y; wv[j][6] = d.x; wv[j][7] = d.y;
      } else {
#pragma unroll
        for (int e = 0; e < 8; e++) wv[j][e] = 0.f;
      }
    }
#pragma unroll
    for (int n = 0; n < NB; n++) {
      if (n < nb) {
        const uint4 u = *reinterpret_cast<const uint4*>(x + (long long)n * k + kk);
        const float2 a = unpack2(u.x, bf), b = unpack2(u.y, bf), c = unpack2(u.z, bf), d = unpack2(u.w, bf);
        const float xv[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
#pragma unroll
        for (int j = 0; j < OT; j++)
#pragma unroll
          for (int e = 0; e < 8; e++) acc[n][j] += xv[e] * wv[j][e];
      }
    }
  }
#pragma unroll
  for (int n = 0; n < NB; n++)
#pragma unroll
    for (int j = 0; j < OT; j++) {
      float v = acc[n][j];
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == 0) red[warp][n * OT + j] = v;
    }
  __syncthreads();
  if (threadIdx.x < NB * OT) {
    float s = 0.f;
    for (int i = 0; i < 8; i++) s += red[i][threadIdx.x];
    const int n = threadIdx.x / OT, j = threadIdx.x % OT;
    if (n < nb && o0 + j < o) part[((long long)split * nb + n) * o + o0 + j] = s;
  }
}

// out[n][o] = act(sum_splits part + bias[o]); optional 16-bit copy for the next layer
__global__ void linear_finalize_kernel(const float* __restrict__ part, int ksplit, int nb, int o,
                                       const float* __restrict__ bias, int act, int bf, float* __restrict__ out32,
                                       uint16_t* __restrict__ out16) {
  griddep_wait();   // PDL: see launch_pdl
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb * o) return;
  float s = bias ? bias[i % o] : 0.f;
  for (int k = 0; k < ksplit; k++) s += part[(long long)k * nb * o + i];
  if (act == PESR_ACT_LRELU) s = s > 0.f ? s : 0.2f * s;
  else if (act == PESR_ACT_RELU) s = fmaxf(s, 0.f);
  if (out32) out32[i] = s;
  if (out16) out16[i] = from_f32(s, bf);
}

// dx[n][k] += sum_{o in split} dy[n][o] * w[o][k]   (fp32 atomics; dx pre-zeroed)
template <int NB>
__global__ void __launch_bounds__(256)
linear_dgrad_kernel(const float* __restrict__ dy, const uint16_t* __restrict__ w, int nb, int k, int o, int osplit,
                    int bf, float* __restrict__ dx) {
  extern __shared__ float dys[];  // [ochunk][NB]
  const int ochunk = (o + osplit - 1) / osplit;
  const int o0 = blockIdx.y * ochunk, o1 = min(o, o0 + ochunk);
  for (int i = threadIdx.x; i < (o1 - o0) * NB; i += blockDim.x) {
    const int oo = i / NB, n = i % NB;
    dys[i] = n < nb ? dy[(long long)n * o + o0 + oo] : 0.f;
  }
  __syncthreads();
  const int kk = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (kk >= k) return;
  float acc[NB][8];
#pragma unroll
  for (int n = 0; n < NB; n++)
#pragma unroll
    for (int e = 0; e < 8; e++) acc[n][e] = 0.f;
  for (int oo = o0; oo < o1; oo++) {
    const uint4 u = *reinterpret_cast<const uint4*>(w + (long long)oo * k + kk);
    const float2 a = unpack2(u.x, bf), b = unpack2(u.y, bf), c = unpack2(u.z, bf), d = unpack2(u.w, bf);
    const float wv[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
    const float* dv = dys + (oo - o0) * NB;
#pragma unroll
    for (int n = 0; n < NB; n++) {
      const float g = dv[n];
#pragma unroll
      for (int e = 0; e < 8; e++) acc[n][e] += g * wv[e];
    }
  }
#pragma unroll
  for (int n = 0; n < NB; n++) {
    if (n < nb) {
#pragma unroll
      for (int e = 0; e < 8; e++) atomicAdd(dx + (long long)n * k + kk + e, acc[n][e]);
    }
  }
}

// dw[o][k] (+)= mul * sum_n dy[n][o] * x[n][k]; thread <-> (8 outputs, 4 consecutive k)
template <int NB>
__global__ void __launch_bounds__(256)
linear_wgrad_kernel(const float* __restrict__ dy, const uint16_t* __restrict__ x, int nb, int k, int o, float mul,
                    const float* __restrict__ div_dev, int accumulate, int bf, float* __restrict__ dw) {
  constexpr int OT = 8;
  __shared__ float dys[OT][NB];
  const int o0 = blockIdx.y * OT;
  if (div_dev) mul /= __ldg(div_dev);
  if (threadIdx.x < OT * NB) {
    const int j = threadIdx.x / NB, n = threadIdx.x % NB;
    dys[j][n] = (n < nb && o0 + j < o) ? dy[(long long)n * o + o0 + j] * mul : 0.f;
  }
  __syncthreads();
  const int kk = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (kk >= k) return;
  float acc[OT][4];
#pragma unroll
  for (int j = 0; j < OT; j++)
#pragma unroll
    for (int e = 0; e < 4; e++) acc[j][e] = 0.f;
#pragma unroll
  for (int n = 0; n < NB; n++) {
    if (n < nb) {
      const uint2 u = *reinterpret_cast<const uint2*>(x + (long long)n * k + kk);
      const float2 a = unpack2(u.x, bf), b = unpack2(u.y, bf);
      const float xv[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
      for (int j = 0; j < OT; j++) {
        const float g = dys[j][n];
#pragma unroll
        for (int e = 0; e < 4; e++) acc[j][e] += g * xv[e];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < OT; j++) {
    if (o0 + j < o) {
      float4* dst = reinterpret_cast<float4*>(dw + (long long)(o0 + j) * k + kk);
      float4 v = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
      if (accumulate) {
        const float4 old = *dst;
        v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
      }
      *dst = v;
    }
  }
}

// fp32 -> 16-bit cast of a flat array (Linear weights keep the reference's [out][in] layout)
__global__ void cast16_kernel(const float* __restrict__ src, long long n, int bf, uint16_t* __restrict__ dst) {
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += (long long)gridDim.x * blockDim.x * 2) {
    if (i + 1 < n) {
      *reinterpret_cast<uint32_t*>(dst + i) = pack2(src[i], src[i + 1], bf);
    } else {
      dst[i] = from_f32(src[i], bf);
    }
  }
}

// NHWC 16-bit [nb][hw][c] <-> NCHW-flattened 16-bit [nb][c*hw] (the `.view(N, -1)` of model/pesr.py:79).
// dir 0: dst[n][c*hw + p] = src[n][p][c].
// dir 1 (backward): dst[n][p][c] = cvt(src32[n][c*hw + p] * mul) * lrelu'(mask[n][p][c])  with src32 fp32.
__global__ void flatten_nchw_kernel(const uint16_t* __restrict__ src, int nb, int hw, int c, uint16_t* __restrict__ dst) {
  griddep_wait();   // PDL: see launch_pdl
  const long long total = (long long)nb * hw * c;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % hw);
    const int cc = (int)((i / hw) % c);
    const int n = (int)(i / ((long long)hw * c));
    dst[i] = src[((long long)n * hw + p) * c + cc];
  }
}
__global__ void unflatten_nchw_kernel(const float* __restrict__ src32, const uint16_t* __restrict__ mask, int nb, int hw,
                                      int c, float mul, const float* __restrict__ mul_dev, float slope, int bf,
                                      uint16_t* __restrict__ dst) {
  griddep_wait();   // PDL: see launch_pdl
  const long long total = (long long)nb * hw * c;
  if (mul_dev) mul *= __ldg(mul_dev);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cc = (int)(i % c);
    const int p = (int)((i / c) % hw);
    const int n = (int)(i / ((long long)hw * c));
    float v = src32[((long long)n * c + cc) * hw + p] * mul;
    if (mask) v *= to_f32(mask[i], bf) > 0.f ? 1.f : slope;
    dst[i] = from_f32(v, bf);
  }
}

}  // namespace pesr

using namespace pesr;

static int nblocks(long long n, int threads, int cap = 148 * 16) {
  long long b = (n + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

static void launch_bn_reduce(const void* x16, const void* y16, long long npix, int c, int groups, const float* mean,
                             const float* rstd, int mode, int dtype, double* sums_ws, cudaStream_t stream) {
  if (c == 64 || c == 128 || c == 256 || c == 512) {
    const int rpp = 256 / (c / 8);
    long long bx = (npix + (long long)rpp * 8 - 1) / ((long long)rpp * 8);
    // two blocks per SM over all groups: measured (tools/perf_helpers.py, PESR_BN_BLOCKS) 2-4 us faster per call than 4 or
    // 8 - every block ends with 2*c double atomics onto the same addresses
    static int cap = -1;
    if (cap < 0) {
      const char* e = getenv("PESR_BN_BLOCKS");
      cap = e ? atoi(e) : 148 * 2;
    }
    const long long cap_g = cap / groups > 0 ? cap / groups : 1;
    if (bx > cap_g) bx = cap_g;
    if (bx < 1) bx = 1;
    launch_pdl(bn_reduce_vec_kernel, dim3((unsigned)bx, (unsigned)groups), 256, 0, stream, reinterpret_cast<const uint4*>(x16),
               reinterpret_cast<const uint4*>(y16), npix, c, mean, rstd, mode, dtype, sums_ws);
  } else {
    long long bx = (npix + 8 * 64 - 1) / (8 * 64);
    if (bx > 1184) bx = 1184;
    dim3 grid((unsigned)bx, (unsigned)((c + 63) / 64), (unsigned)groups);
    launch_pdl(bn_reduce_kernel, grid, 256, 0, stream, reinterpret_cast<const uint16_t*>(x16),
                                              reinterpret_cast<const uint16_t*>(y16), npix, c, mean, rstd, mode, dtype,
                                              sums_ws);
  }
}

static int bn_zero(double* sums_ws, int groups, int c, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(sums_ws, 0, sizeof(double) * 2 * (size_t)c * groups, stream);
  if (e != cudaSuccess) { set_error("bn: memset failed: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

extern "C" int pesr_bn_reduce(const void* y16, int64_t npix, int32_t c, int32_t groups, double* sums_ws, int32_t zero_first,
                              int32_t dtype, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(y16 && sums_ws && npix > 0 && c > 0 && c % 8 == 0 && c <= 512 && groups >= 1 && groups <= 8,
                 "bn_reduce: bad arguments");
  if (zero_first) { int r = bn_zero(sums_ws, groups, c, stream); if (r) return r; }
  launch_bn_reduce(y16, nullptr, npix, c, groups, nullptr, nullptr, 0, dtype, sums_ws, stream);
  count_launch();
  PESR_CHECK_LAUNCH("bn_reduce");
  return 0;
}

extern "C" int pesr_bn_lrelu_fwd(const void* y16, int64_t npix, int32_t c, int32_t groups, const double* sums_ws, float eps,
                                 float momentum, float* mean, float* rstd, const float* gamma, const float* beta,
                                 float* running_mean, float* running_var, int64_t* num_batches_tracked,
                                 const float* running_mean_shift, float slope, int32_t dtype, void* a16, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(y16 && a16 && mean && rstd && gamma && beta && npix > 0 && c % 8 == 0 && c <= 512 && groups >= 1 &&
                     groups <= 8, "bn_lrelu_fwd: bad arguments");
  const long long nvec = npix * (c / 8);
  int bx = nblocks(nvec, 256, 148 * 16 / groups);
  launch_pdl(bn_lrelu_fwd_kernel, dim3((unsigned)bx, (unsigned)groups), 256, 0, stream, reinterpret_cast<const uint4*>(y16), nvec, c,
             sums_ws, (double)npix, eps, momentum, mean, rstd, gamma, beta, running_mean, running_var,
             reinterpret_cast<long long*>(num_batches_tracked), running_mean_shift, slope, dtype,
             reinterpret_cast<uint4*>(a16));
  count_launch();
  PESR_CHECK_LAUNCH("bn_lrelu_fwd");
  return 0;
}

extern "C" int pesr_bn_lrelu_bwd(const void* dz16, const void* y16, int64_t npix, int32_t c, int32_t groups,
                                 const float* mean, const float* rstd, const float* gamma, double* sums_ws,
                                 int32_t zero_first, float grad_mul, const float* grad_div_dev, int32_t dtype, void* dy16,
                                 float* dgamma, float* dbeta, int32_t accumulate, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(dz16 && y16 && dy16 && mean && rstd && gamma && sums_ws && npix > 0 && c % 8 == 0 && c <= 512 &&
                     groups >= 1 && groups <= 8, "bn_lrelu_bwd: bad arguments");
  if (zero_first) { int r = bn_zero(sums_ws, groups, c, stream); if (r) return r; }
  launch_bn_reduce(dz16, y16, npix, c, groups, mean, rstd, 1, dtype, sums_ws, stream);
  const long long nvec = npix * (c / 8);
  int bx = nblocks(nvec, 256, 148 * 16 / groups);
  launch_pdl(bn_bwd_apply_kernel, dim3((unsigned)bx, (unsigned)groups), 256, 0, stream,
      reinterpret_cast<const uint4*>(dz16), reinterpret_cast<const uint4*>(y16), nvec, c, (double)npix, mean, rstd, gamma,
      sums_ws, grad_mul, grad_div_dev, dtype, reinterpret_cast<uint4*>(dy16), dgamma, dbeta, accumulate);
  count_launch(2);
  PESR_CHECK_LAUNCH("bn_lrelu_bwd");
  return 0;
}

extern "C" int pesr_maxpool2_fwd(const void* x16, int32_t nb, int32_t h, int32_t w, int32_t c, int32_t dtype,
                                 void* y16, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(x16 && y16 && nb > 0 && h >= 2 && w >= 2 && c % 8 == 0, "maxpool2_fwd: bad arguments");
  const long long total = (long long)nb * (h / 2) * (w / 2) * (c / 8);
  launch_pdl(maxpool2_fwd_kernel, nblocks(total, 256), 256, 0, stream, reinterpret_cast<const uint4*>(x16), nb, h, w, c, dtype,
                                                              reinterpret_cast<uint4*>(y16));
  count_launch();
  PESR_CHECK_LAUNCH("maxpool2_fwd");
  return 0;
}

extern "C" int pesr_maxpool2_bwd(const void* x16, const void* dy16, int32_t nb, int32_t h, int32_t w, int32_t c,
                                 int32_t relu_mask, int32_t dtype, void* dx16, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(x16 && dy16 && dx16 && nb > 0 && h >= 2 && w >= 2 && c % 8 == 0, "maxpool2_bwd: bad arguments");
  const long long total = (long long)nb * (h / 2) * (w / 2) * (c / 8);
  launch_pdl(maxpool2_bwd_kernel, nblocks(total, 256), 256, 0, stream, reinterpret_cast<const uint4*>(x16),
                                                              reinterpret_cast<const uint4*>(dy16), nb, h, w, c,
                                                              relu_mask, dtype, reinterpret_cast<uint4*>(dx16));
  count_launch();
  PESR_CHECK_LAUNCH("maxpool2_bwd");
  return 0;
}

static int pick_ksplit(int k, int o) {
  const int otiles = (o + 7) / 8;
  int ks = (148 * 4 + otiles - 1) / otiles;
  const int max_ks = k / 2048 > 0 ? k / 2048 : 1;
  if (ks > max_ks) ks = max_ks;
  if (ks < 1) ks = 1;
  if (ks > 64) ks = 64;
  return ks;
}

extern "C" int64_t pesr_linear_workspace_floats(int32_t nb, int32_t k, int32_t o) {
  return (int64_t)pick_ksplit(k, o) * nb * o;
}

extern "C" int pesr_linear_skinny_fwd(const void* x16, const void* w16, const float* bias, int32_t nb, int32_t k,
                                      int32_t o, int32_t act, int32_t dtype, float* workspace, float* out32,
                                      void* out16, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(x16 && w16 && workspace && nb > 0 && k > 0 && o > 0 && k % 8 == 0, "linear_fwd: bad arguments");
  PESR_CHECK_ARG(nb <= 16, "linear_fwd: at most 16 rows per call (got %d)", nb);
  const int ks = pick_ksplit(k, o);
  dim3 grid((unsigned)((o + 7) / 8), (unsigned)ks);
  linear_fwd_kernel<16><<<grid, 256, 0, stream>>>(reinterpret_cast<const uint16_t*>(x16),
                                                  reinterpret_cast<const uint16_t*>(w16), nb, k, o, ks, dtype, workspace);
  launch_pdl(linear_finalize_kernel, (nb * o + 255) / 256, 256, 0, stream, workspace, ks, nb, o, bias, act, dtype, out32,
                                                                  reinterpret_cast<uint16_t*>(out16));
  count_launch(2);
  PESR_CHECK_LAUNCH("linear_fwd");
  return 0;
}

extern "C" int pesr_linear_finalize(const float* partials, int32_t ksplit, int32_t nb, int32_t o, const float* bias,
                                    int32_t act, int32_t dtype, float* out32, void* out16, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(partials && ksplit >= 1 && nb > 0 && o > 0 && (out32 || out16), "linear_finalize: bad arguments");
  launch_pdl(linear_finalize_kernel, (nb * o + 255) / 256, 256, 0, stream, partials, ksplit, nb, o, bias, act, dtype, out32,
                                                                  reinterpret_cast<uint16_t*>(out16));
  count_launch();
  PESR_CHECK_LAUNCH("linear_finalize");
  return 0;
}

extern "C" int pesr_linear_skinny_dgrad(const float* dy, const void* w16, int32_t nb, int32_t k, int32_t o,
                                        int32_t dtype, float* dx32, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(dy && w16 && dx32 && nb > 0 && nb <= 16 && k % 8 == 0 && o > 0, "linear_dgrad: bad arguments");
  cudaError_t e = cudaMemsetAsync(dx32, 0, sizeof(float) * (size_t)nb * k, stream);
  if (e != cudaSuccess) { set_error("linear_dgrad: memset failed: %s", cudaGetErrorString(e)); return (int)e; }
  const int kblocks = (k / 8 + 255) / 256;
  int osplit = (148 * 4 + kblocks - 1) / kblocks;
  if (osplit > o) osplit = o;
  if (osplit > 64) osplit = 64;
  if (osplit < 1) osplit = 1;
  const int ochunk = (o + osplit - 1) / osplit;
  dim3 grid((unsigned)kblocks, (unsigned)osplit);
  linear_dgrad_kernel<16><<<grid, 256, ochunk * 16 * sizeof(float), stream>>>(
      dy, reinterpret_cast<const uint16_t*>(w16), nb, k, o, osplit, dtype, dx32);
  count_launch();
  PESR_CHECK_LAUNCH("linear_dgrad");
  return 0;
}

extern "C" int pesr_linear_skinny_wgrad(const float* dy, const void* x16, int32_t nb, int32_t k, int32_t o, float mul,
                                        const float* div_dev, int32_t accumulate, int32_t dtype, float* dw,
                                        void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(dy && x16 && dw && nb > 0 && nb <= 32 && k % 4 == 0 && o > 0, "linear_wgrad: bad arguments");
  dim3 grid((unsigned)((k / 4 + 255) / 256), (unsigned)((o + 7) / 8));
  // up to 32 rows per pass: the 302 MB gradient of the Discriminator's first Linear is written once for the two
  // batched calls of a phase (a second 16-row pass re-reads and re-writes all of it)
  if (nb <= 16)
    linear_wgrad_kernel<16><<<grid, 256, 0, stream>>>(dy, reinterpret_cast<const uint16_t*>(x16), nb, k, o, mul, div_dev,
                                                      accumulate, dtype, dw);
  else
    linear_wgrad_kernel<32><<<grid, 256, 0, stream>>>(dy, reinterpret_cast<const uint16_t*>(x16), nb, k, o, mul, div_dev,
                                                      accumulate, dtype, dw);
  count_launch();
  PESR_CHECK_LAUNCH("linear_wgrad");
  return 0;
}

extern "C" int pesr_cast16(const float* src, int64_t n, int32_t dtype, void* dst, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(src && dst && n > 0, "cast16: bad arguments");
  cast16_kernel<<<nblocks((n + 1) / 2, 256), 256, 0, stream>>>(src, n, dtype, reinterpret_cast<uint16_t*>(dst));
  count_launch();
  note_weight_write(stream);   // cast16 produces the 16-bit Linear weights the FC igemm reads as its B operand
  PESR_CHECK_LAUNCH("cast16");
  return 0;
}

extern "C" int pesr_flatten_nchw16(const void* src_nhwc16, int32_t nb, int32_t hw, int32_t c, void* dst, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(src_nhwc16 && dst && nb > 0 && hw > 0 && c > 0, "flatten_nchw16: bad arguments");
  launch_pdl(flatten_nchw_kernel, nblocks((long long)nb * hw * c, 256), 256, 0, stream, 
      reinterpret_cast<const uint16_t*>(src_nhwc16), nb, hw, c, reinterpret_cast<uint16_t*>(dst));
  count_launch();
  PESR_CHECK_LAUNCH("flatten_nchw16");
  return 0;
}

extern "C" int pesr_unflatten_nchw16(const float* src32_nchw, const void* mask_nhwc16, int32_t nb, int32_t hw, int32_t c,
                                     float mul, const float* mul_dev, float slope, int32_t dtype, void* dst_nhwc16,
                                     void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(src32_nchw && dst_nhwc16 && nb > 0 && hw > 0 && c > 0, "unflatten_nchw16: bad arguments");
  launch_pdl(unflatten_nchw_kernel, nblocks((long long)nb * hw * c, 256), 256, 0, stream, 
      src32_nchw, reinterpret_cast<const uint16_t*>(mask_nhwc16), nb, hw, c, mul, mul_dev, slope, dtype,
      reinterpret_cast<uint16_t*>(dst_nhwc16));
  count_launch();
  PESR_CHECK_LAUNCH("unflatten_nchw16");
  return 0;
}
