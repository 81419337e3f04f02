// Bring-up microbenchmark: issue-rate of back-to-back tcgen05.mma (M=128, N=n, K=16, SS operands) with no loads and
// no barriers inside the loop.  Tells the hardware floor the implicit-GEMM mainloop can be compared with.
#include "../../pesr_b200/csrc/common.cuh"
#include "../../pesr_b200/csrc/host_util.cuh"
#include "../../include/pesr_b200_debug.h"

namespace pesr {

template <bool kPair>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int n, int iters, int distinct_stages, int mn_major, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stage_bytes = 16384 + (kPair ? n / 2 : n) * 128;
  uint64_t& bar = *reinterpret_cast<uint64_t*>(smem + (size_t)distinct_stages * stage_bytes);
  uint32_t& tmem_ptr = *reinterpret_cast<uint32_t*>(smem + (size_t)distinct_stages * stage_bytes + 8);
  for (int i = threadIdx.x; i < distinct_stages * stage_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  if (warp == 1) {
    if (kPair) { tmem_alloc2(&tmem_ptr, 512); tmem_relinquish2(); } else { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  const bool leader = !kPair || cluster_ctarank() == 0;
  if (warp == 0 && leader) {
    const uint32_t idesc = make_idesc(kPair ? 256 : 128, n, 0, mn_major & 1, (mn_major >> 1) & 1);
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
      const uint32_t a_addr = smem_u32(smem + (size_t)(it % distinct_stages) * stage_bytes);
      const uint32_t b_addr = a_addr + 16384;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint64_t da = (mn_major & 1) ? make_smem_desc(a_addr + k * 2048, 8192, 1024) : make_smem_desc(a_addr + k * 32, 16, 1024);
        const uint64_t db = (mn_major & 2) ? make_smem_desc(b_addr + k * 2048, 8192, 1024) : make_smem_desc(b_addr + k * 32, 16, 1024);
        if (kPair) umma2_f16_w(tmem_base, da, db, idesc, (it | k) != 0); else umma_f16_w(tmem_base, da, db, idesc, (it | k) != 0);
      }
    }
    const unsigned long long t1 = clock64();
    if (kPair) umma2_commit_both_w(&bar); else umma_commit_w(&bar);
    mbar_wait(&bar, 0);
    const unsigned long long t2 = clock64();
    if (blockIdx.x == 0 && lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace pesr

using namespace pesr;

// out[0] = cycles to ISSUE iters*4 MMAs, out[1] = cycles until they all completed (block 0 / cluster 0).
extern "C" int pesr_debug_mma_rate(int32_t n, int32_t iters, int32_t stages, int32_t pair_and_major, int32_t blocks,
                                   unsigned long long* out_dev, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int pair = pair_and_major & 1, mn_major = pair_and_major >> 1;   // bit0 pair, bit1 A MN-major, bit2 B MN-major
  PESR_CHECK_ARG(n >= 32 && n <= 256 && n % 16 == 0 && iters > 0 && stages >= 1 && stages <= 4 && out_dev, "mma_rate: bad arguments");
  const size_t smem = (size_t)stages * (16384 + n * 128) + 1024 + 64;
  cudaGetLastError();
  cudaFuncSetAttribute(mma_rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(mma_rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (pair) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(blocks & ~1);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, mma_rate_kernel<true>, n, iters, stages, mn_major, out_dev);
    if (e != cudaSuccess) { set_error("mma_rate: %s", cudaGetErrorString(e)); return (int)e; }
  } else {
    mma_rate_kernel<false><<<blocks, 128, smem, stream>>>(n, iters, stages, mn_major, out_dev);
  }
  PESR_CHECK_LAUNCH("mma_rate");
  return 0;
}

// ------------------------------------------------------------------------------------------
// SM-hog probe: n CTAs that each pin one SM (200 KB of dynamic shared memory: nothing of ours fits beside them) and
// spin for `usec` microseconds.  Used by tools/sm_hog_probe.py to measure what a few unavailable SMs - e.g. the CTAs of
// a concurrent NCCL all-reduce - cost the persistent one-CTA-per-SM kernels with their static tile assignment.
// ------------------------------------------------------------------------------------------
namespace pesr {
__global__ void sm_hog_kernel(unsigned long long nsec) {
  extern __shared__ uint8_t hog_smem[];
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do {
    __nanosleep(2000);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  } while (t - t0 < nsec);
  if (nsec == 0) hog_smem[0] = 1;
}
}  // namespace pesr

extern "C" int pesr_debug_sm_hog(int32_t n_ctas, int64_t usec, void* stream_) {
  using namespace pesr;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(sm_hog_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  sm_hog_kernel<<<n_ctas, 32, 200 * 1024, stream>>>((unsigned long long)usec * 1000ull);
  PESR_CHECK_LAUNCH("debug_sm_hog");
  return 0;
}
