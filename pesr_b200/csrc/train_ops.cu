// HBM-bound training kernels: the loss reductions of train.py:131-140 (L1, MSE, TV) with their
// gradients produced in the same pass, the 16-logit relativistic GAN losses (train.py:213,251 and
// model/focal_loss.py:9-13, torch-0.4 gradient semantics) and a multi-tensor Adam (train.py:124-126).
#include "common.cuh"
#include "host_util.cuh"

namespace pesr {

__device__ __forceinline__ float block_sum(float v, float* sm) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x < 32) {
    r = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  __syncthreads();
  return r;  // valid in thread 0
}

// mode 0: L1 mean  (value = mean|a-b|,   grad = sign(a-b)/n)
// mode 1: MSE mean (value = mean (a-b)^2, grad = 2(a-b)/n)
__global__ void diff_loss_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, int mode,
                                 float* __restrict__ loss, float* __restrict__ grad) {
  griddep_wait();   // PDL: see launch_pdl
  __shared__ float sm[32];
  const float inv_n = 1.f / (float)n;
  float acc = 0.f;
  const long long n4 = n >> 2;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  float4* g4 = reinterpret_cast<float4*>(grad);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 x = a4[i], y = b4[i];
    float d[4] = {x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w};
    float g[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (mode == 0) {
        acc += fabsf(d[j]);
        g[j] = d[j] > 0.f ? inv_n : (d[j] < 0.f ? -inv_n : 0.f);
      } else {
        acc += d[j] * d[j];
        g[j] = 2.f * d[j] * inv_n;
      }
    }
    if (grad) g4[i] = make_float4(g[0], g[1], g[2], g[3]);
  }
  if (blockIdx.x == 0) {
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      const float d = a[i] - b[i];
      if (mode == 0) {
        acc += fabsf(d);
        if (grad) grad[i] = d > 0.f ? inv_n : (d < 0.f ? -inv_n : 0.f);
      } else {
        acc += d * d;
        if (grad) grad[i] = 2.f * d * inv_n;
      }
    }
  }
  const float s = block_sum(acc, sm);
  if (threadIdx.x == 0) atomicAdd(loss, s * inv_n);
}

// TV (train.py:137-140): sum |y[..., w] - y[..., w+1]| + sum |y[..., h, :] - y[..., h+1, :]| over [planes][h][w];
// grad[p] = sgn(y[p]-y[right]) - sgn(y[left]-y[p]) + sgn(y[p]-y[down]) - sgn(y[up]-y[p]).
__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }
__global__ void tv_loss_kernel(const float* __restrict__ y, long long planes, int h, int w, float* __restrict__ loss,
                               float* __restrict__ grad) {
  griddep_wait();   // PDL: see launch_pdl
  __shared__ float sm[32];
  const long long total = planes * h * w;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const int r = (int)((i / w) % h);
    const float c = y[i];
    float g = 0.f;
    if (x + 1 < w) { const float d = c - y[i + 1]; acc += fabsf(d); g += sgn(d); }
    if (x > 0) g -= sgn(y[i - 1] - c);
    if (r + 1 < h) { const float d = c - y[i + w]; acc += fabsf(d); g += sgn(d); }
    if (r > 0) g -= sgn(y[i - w] - c);
    if (grad) grad[i] = g;
  }
  const float s = block_sum(acc, sm);
  if (threadIdx.x == 0) atomicAdd(loss, s);
}

// Relativistic GAN losses on n logits (n = batch size, 16 on the headline config): one block.
//  x = sign_a * a + sign_b * b  (RSGAN: D phase x = real - fake, G phase x = fake - real), target t.
//  mode 0: BCE-with-logits mean (train.py:213).   mode 1: focal (model/focal_loss.py) with the torch-0.4
//  gradient (through the weight); mode 2: focal with the weight detached (torch >= 1.0 semantics).
// Outputs: loss (scalar), dx/da and dx/db gradients (each n floats, may be NULL).
__device__ __forceinline__ float softplus_f(float z) { return z > 0.f ? z + log1pf(expf(-z)) : log1pf(expf(z)); }
__device__ __forceinline__ float sigmoid_f(float z) { return 1.f / (1.f + expf(-z)); }
__global__ void gan_loss_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, float sign_a,
                                float sign_b, float t, int mode, float gamma, float* __restrict__ loss,
                                float* __restrict__ ga, float* __restrict__ gb) {
  __shared__ float sm[32];
  float acc = 0.f;
  const float inv_n = 1.f / (float)n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float x = sign_a * a[i] + (b ? sign_b * b[i] : 0.f);
    // bce(x,t) = softplus(x) - t*x ; d/dx = sigmoid(x) - t
    const float bce = softplus_f(x) - t * x;
    const float dbce = sigmoid_f(x) - t;
    float l, g;
    if (mode == 0) {
      l = bce;
      g = dbce;
    } else {
      const float p = sigmoid_f(x);
      const float pt = p * t + (1.f - p) * (1.f - t);
      const float om = fmaxf(1.f - pt, 0.f);
      const float wgt = powf(om, gamma);
      l = wgt * bce;
      g = wgt * dbce;
      if (mode == 1) {
        // d w/dx = gamma*(1-pt)^(gamma-1) * d(1-pt)/dx ; d pt/dx = (2t-1) p (1-p)
        const float dw = om > 0.f ? gamma * powf(om, gamma - 1.f) * (-(2.f * t - 1.f) * p * (1.f - p)) : 0.f;
        g += dw * bce;
      }
    }
    acc += l;
    if (ga) ga[i] = g * inv_n * sign_a;
    if (gb && b) gb[i] = g * inv_n * sign_b;
  }
  const float s = block_sum(acc, sm);
  if (threadIdx.x == 0) *loss = s * inv_n;
}

// ------------------------------------------------------------------------------------------
// multi-tensor Adam: table rows = {p, g, m, v, n} (device pointers as int64), one row per chunk
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void adam_chunk(const long long* __restrict__ table, int c, float lr, float beta1, float beta2,
                                           float eps, float bc1, float bc2_sqrt, float grad_mul) {
  float* p = reinterpret_cast<float*>(table[c * 5 + 0]);
  const float* g = reinterpret_cast<const float*>(table[c * 5 + 1]);
  float* m = reinterpret_cast<float*>(table[c * 5 + 2]);
  float* v = reinterpret_cast<float*>(table[c * 5 + 3]);
  const long long n = table[c * 5 + 4];
  const float step = lr / bc1;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  long long i0 = 0;
  if (vec) {
    const long long n4 = n >> 2;
    for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 pp = reinterpret_cast<float4*>(p)[i];
      const float4 gg = reinterpret_cast<const float4*>(g)[i];
      float4 mm = reinterpret_cast<float4*>(m)[i];
      float4 vv = reinterpret_cast<float4*>(v)[i];
      float* pa = &pp.x; const float* gaa = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float gj = gaa[j] * grad_mul;
        ma[j] = beta1 * ma[j] + (1.f - beta1) * gj;
        va[j] = beta2 * va[j] + (1.f - beta2) * gj * gj;
        pa[j] -= step * ma[j] / (sqrtf(va[j]) / bc2_sqrt + eps);
      }
      reinterpret_cast<float4*>(p)[i] = pp;
      reinterpret_cast<float4*>(m)[i] = mm;
      reinterpret_cast<float4*>(v)[i] = vv;
    }
    i0 = n4 << 2;
  }
  for (long long i = i0 + threadIdx.x; i < n; i += blockDim.x) {
    const float gj = g[i] * grad_mul;
    const float mj = beta1 * m[i] + (1.f - beta1) * gj;
    const float vj = beta2 * v[i] + (1.f - beta2) * gj * gj;
    m[i] = mj;
    v[i] = vj;
    p[i] -= step * mj / (sqrtf(vj) / bc2_sqrt + eps);
  }
}

__global__ void adam_multi_kernel(const long long* __restrict__ table, int nchunks, float lr, float beta1, float beta2,
                                  float eps, float bc1, float bc2_sqrt, float grad_mul) {
  const int c = blockIdx.x;
  if (c >= nchunks) return;
  adam_chunk(table, c, lr, beta1, beta2, eps, bc1, bc2_sqrt, grad_mul);
}

// Same with the step count and the learning rate read from device memory, so that a captured CUDA graph of the
// training step replays with the right bias corrections (torch.optim.Adam(capturable=True) semantics).
__global__ void adam_multi_dev_kernel(const long long* __restrict__ table, int nchunks, const float* __restrict__ lr_dev,
                                      float beta1, float beta2, float eps, const int* __restrict__ step_dev,
                                      float grad_mul) {
  const int c = blockIdx.x;
  if (c >= nchunks) return;
  const double step = (double)__ldg(step_dev);
  const float bc1 = (float)(1.0 - pow((double)beta1, step));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, step));
  adam_chunk(table, c, __ldg(lr_dev), beta1, beta2, eps, bc1, bc2_sqrt, grad_mul);
}

}  // namespace pesr

using namespace pesr;

static int grid_for(long long n, int threads, int per_thread, int cap) {
  long long b = (n + (long long)threads * per_thread - 1) / ((long long)threads * per_thread);
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

extern "C" int pesr_loss_l1(const float* a, const float* b, int64_t n, float* loss, float* grad, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(a && b && loss && n > 0, "loss_l1: bad arguments");
  PESR_CHECK_ARG(((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0) && (!grad || (uintptr_t)grad % 16 == 0),
                 "loss_l1: pointers must be 16-byte aligned");
  cudaMemsetAsync(loss, 0, sizeof(float), stream);
  launch_pdl(diff_loss_kernel, grid_for(n, 256, 16, 148 * 8), 256, 0, stream, a, b, n, 0, loss, grad);
  count_launch();
  PESR_CHECK_LAUNCH("loss_l1");
  return 0;
}

extern "C" int pesr_loss_mse(const float* a, const float* b, int64_t n, float* loss, float* grad, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(a && b && loss && n > 0, "loss_mse: bad arguments");
  PESR_CHECK_ARG(((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0) && (!grad || (uintptr_t)grad % 16 == 0),
                 "loss_mse: pointers must be 16-byte aligned");
  cudaMemsetAsync(loss, 0, sizeof(float), stream);
  launch_pdl(diff_loss_kernel, grid_for(n, 256, 16, 148 * 8), 256, 0, stream, a, b, n, 1, loss, grad);
  count_launch();
  PESR_CHECK_LAUNCH("loss_mse");
  return 0;
}

extern "C" int pesr_loss_tv(const float* y, int64_t planes, int32_t h, int32_t w, float* loss, float* grad,
                            void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(y && loss && planes > 0 && h > 0 && w > 0, "loss_tv: bad arguments");
  cudaMemsetAsync(loss, 0, sizeof(float), stream);
  launch_pdl(tv_loss_kernel, grid_for(planes * h * w, 256, 8, 148 * 8), 256, 0, stream, y, planes, h, w, loss, grad);
  count_launch();
  PESR_CHECK_LAUNCH("loss_tv");
  return 0;
}

extern "C" int pesr_loss_gan(const float* a, const float* b, int32_t n, float sign_a, float sign_b, float target,
                             int32_t mode, float gamma, float* loss, float* grad_a, float* grad_b, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(a && loss && n > 0, "loss_gan: bad arguments");
  PESR_CHECK_ARG(mode >= 0 && mode <= 2, "loss_gan: mode %d", mode);
  gan_loss_kernel<<<1, 128, 0, stream>>>(a, b, n, sign_a, sign_b, target, mode, gamma, loss, grad_a, grad_b);
  count_launch();
  PESR_CHECK_LAUNCH("loss_gan");
  return 0;
}

extern "C" int pesr_adam_multi(const int64_t* table_dev, int32_t nchunks, float lr, float beta1, float beta2,
                               float eps, int32_t step, float grad_mul, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(table_dev && nchunks > 0 && step >= 1, "adam_multi: bad arguments");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_multi_kernel<<<nchunks, 256, 0, stream>>>(reinterpret_cast<const long long*>(table_dev), nchunks, lr, beta1,
                                                 beta2, eps, (float)bc1, (float)sqrt(bc2), grad_mul);
  count_launch();
  PESR_CHECK_LAUNCH("adam_multi");
  return 0;
}

extern "C" int pesr_adam_multi_dev(const int64_t* table_dev, int32_t nchunks, const float* lr_dev, float beta1,
                                   float beta2, float eps, const int32_t* step_dev, float grad_mul, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(table_dev && nchunks > 0 && lr_dev && step_dev, "adam_multi_dev: bad arguments");
  adam_multi_dev_kernel<<<nchunks, 256, 0, stream>>>(reinterpret_cast<const long long*>(table_dev), nchunks, lr_dev, beta1,
                                                     beta2, eps, step_dev, grad_mul);
  count_launch();
  PESR_CHECK_LAUNCH("adam_multi_dev");
  return 0;
}
