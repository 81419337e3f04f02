// Gradient all-reduce (average) over NVLink 5 / NVSwitch peer memory for the data-parallel step (replaces the
// gradient reduction nn.DataParallel performs on GPU 0, train.py:114-118).
//
// Every rank keeps its flat fp32 gradient buffer in symmetric memory (the same allocation mapped into every process of
// the node).  For a range [off, off + n) of that buffer the kernel runs the two-shot schedule on FEW CTAs (the persistent
// tensor-core kernels leave exactly that many SMs free, PESR_OPT_RESERVE_SMS, so the reduction neither waits for an SM
// nor takes one from a convolution):
//   rank r owns the r-th slice of the range; it loads that slice from every rank's buffer (peer loads over NVLink, all
//   `world` loads of an element in flight together), sums in a fixed rank order, scales by 1/world and stores the result
//   into every rank's buffer (peer stores).  One owner per element: all ranks end up with bit-identical gradients.
// With an NVSwitch multicast mapping the slice is read with multimem.ld_reduce (the switch adds the `world` copies) and
// written with multimem.st (the switch broadcasts): one load and one store per element instead of `world` of each.
// Cross-rank ordering is part of the kernel: before the reduction every rank signals "my range is produced" into every
// peer's signal pad and waits for all peers' signals; after it "my slice is stored everywhere" the same way (release /
// acquire at system scope, a call counter as the flag value so the pads are never reset).  ONE launch per bucket, as ONE
// thread-block cluster: the hardware places the CTAs of a cluster on neighbouring SMs, so the SMs it occupies come in
// pairs - a stray single CTA on a TPC would cost the CTA-pair convolutions (clusters of 2) a whole pair.
#include "common.cuh"
#include "host_util.cuh"

namespace pesr {

struct PeerPtrs {
  float* p[PESR_MAX_PEERS];
};
struct PadPtrs {
  uint32_t* p[PESR_MAX_PEERS];      // every rank's signal pad (NULL entries: the caller orders the ranks)
};
constexpr int kPadSlotA = 256, kPadSlotB = 256 + PESR_MAX_PEERS;      // uint32 indices inside a pad (torch uses the first few)

// All ranks meet: thread p of the first CTA writes `epoch` into slot [slot + rank] of rank p's pad (release: everything
// this GPU wrote before is visible to whoever acquires it) and waits until rank p has written into ours.
__device__ __forceinline__ void cross_rank_barrier(const PadPtrs& pads, int world, int rank, int slot, uint32_t epoch) {
  const int p = threadIdx.x;
  if (p < world && p != rank) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(pads.p[p] + slot + rank), "r"(epoch) : "memory");
    const uint32_t* mine = pads.p[rank] + slot + p;
    uint32_t v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    } while ((int32_t)(v - epoch) < 0);
  }
}
__device__ __forceinline__ void ranks_meet_before(const PadPtrs& pads, int world, int rank, uint32_t epoch) {
  if (pads.p[0] != nullptr) {
    if (cluster_ctarank() == 0) { cross_rank_barrier(pads, world, rank, kPadSlotA, epoch); __syncthreads(); }
    cluster_sync_all();
  }
}
__device__ __forceinline__ void ranks_meet_after(const PadPtrs& pads, int world, int rank, uint32_t epoch) {
  __threadfence_system();
  if (pads.p[0] != nullptr) {
    cluster_sync_all();
    if (cluster_ctarank() == 0) cross_rank_barrier(pads, world, rank, kPadSlotB, epoch);
  }
}

__device__ __forceinline__ float4 ld_sys_v4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys_v4(float* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 multimem_ld_reduce_v4(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_v4(float* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// slice of this rank in float4 units
__device__ __forceinline__ void owner_slice(long long n4, int world, int rank, long long& s0, long long& s1) {
  const long long per = (n4 + world - 1) / world;
  s0 = (long long)rank * per;
  if (s0 > n4) s0 = n4;
  s1 = s0 + per;
  if (s1 > n4) s1 = n4;
}

// U vectors of 16 bytes per thread and iteration: U * W independent peer loads in flight per thread (NVLink round trips are
// several microseconds; four CTAs need ~0.5 MB in flight to reach a few hundred GB/s)
template <int W, int U>
__global__ void __launch_bounds__(1024)
allreduce_p2p_kernel(PeerPtrs ptrs, PadPtrs pads, uint32_t epoch, int world, int rank, long long off, long long n4, float scale) {
  ranks_meet_before(pads, world, rank, epoch);
  long long s0, s1;
  owner_slice(n4, W ? W : world, rank, s0, s1);
  const int w = W ? W : world;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = s0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < s1; i0 += U * stride) {
    if (W) {
      float4 v[U][W ? W : 1];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const long long i = i0 + u * stride;
        if (i < s1) {
#pragma unroll
          for (int k = 0; k < W; k++) v[u][k] = ld_sys_v4(ptrs.p[k] + off + 4 * i);      // summed in rank order below
        }
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const long long i = i0 + u * stride;
        if (i < s1) {
          float4 a = v[u][0];
#pragma unroll
          for (int k = 1; k < W; k++) { a.x += v[u][k].x; a.y += v[u][k].y; a.z += v[u][k].z; a.w += v[u][k].w; }
          a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
#pragma unroll
          for (int k = 0; k < W; k++)      // staggered: the ranks do not hit one peer together
            st_sys_v4(ptrs.p[(rank + k) % (W ? W : 1)] + off + 4 * i, a);
        }
      }
    } else {
      for (int u = 0; u < U; u++) {
        const long long i = i0 + u * stride;
        if (i >= s1) break;
        const long long e = off + 4 * i;
        float4 a = ld_sys_v4(ptrs.p[0] + e);
        for (int k = 1; k < w; k++) {
          const float4 t = ld_sys_v4(ptrs.p[k] + e);
          a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
        }
        a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
        for (int k = 0; k < w; k++) st_sys_v4(ptrs.p[(rank + k) % w] + e, a);
      }
    }
  }
  ranks_meet_after(pads, world, rank, epoch);
}

__global__ void __launch_bounds__(1024)
allreduce_multimem_kernel(float* mc, PadPtrs pads, uint32_t epoch, int world, int rank, long long off, long long n4, float scale) {
  constexpr int U = 8;                               // independent in-switch reductions in flight per thread
  ranks_meet_before(pads, world, rank, epoch);
  long long s0, s1;
  owner_slice(n4, world, rank, s0, s1);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = s0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < s1; i0 += U * stride) {
    float4 a[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long i = i0 + u * stride;
      if (i < s1) a[u] = multimem_ld_reduce_v4(mc + off + 4 * i);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long i = i0 + u * stride;
      if (i < s1) {
        a[u].x *= scale; a[u].y *= scale; a[u].z *= scale; a[u].w *= scale;
        multimem_st_v4(mc + off + 4 * i, a[u]);
      }
    }
  }
  ranks_meet_after(pads, world, rank, epoch);
}

}  // namespace pesr

using namespace pesr;

template <typename... KArgs, typename... Args>
static cudaError_t launch_cluster(void (*kernel)(KArgs...), int ctas, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(1024);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = ctas;          // the whole grid is one cluster
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

extern "C" int pesr_allreduce_p2p(const uint64_t* peer_ptrs, const uint64_t* pad_ptrs, int32_t world, int32_t rank,
                                  uint64_t multicast_ptr, int64_t offset_elems, int64_t count, float scale, int32_t ctas,
                                  uint32_t epoch, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PESR_CHECK_ARG(world >= 2 && world <= PESR_MAX_PEERS && rank >= 0 && rank < world, "allreduce_p2p: world %d rank %d", world, rank);
  PESR_CHECK_ARG(count > 0 && count % 4 == 0 && offset_elems % 4 == 0, "allreduce_p2p: range must be float4-aligned");
  PESR_CHECK_ARG(multicast_ptr != 0 || peer_ptrs != nullptr, "allreduce_p2p: no peer pointers");
  if (ctas < 1) ctas = 1;
  if (ctas > 8) ctas = 8;                   // portable cluster size
  const long long n4 = count / 4;
  PadPtrs pads;
  memset(&pads, 0, sizeof(pads));
  if (pad_ptrs) {
    for (int i = 0; i < world; i++) {
      PESR_CHECK_ARG(pad_ptrs[i] != 0, "allreduce_p2p: signal pad %d is null", i);
      pads.p[i] = reinterpret_cast<uint32_t*>(pad_ptrs[i]);
    }
  }
  cudaError_t e;
  if (multicast_ptr) {
    e = launch_cluster(allreduce_multimem_kernel, ctas, stream, reinterpret_cast<float*>(multicast_ptr), pads, epoch, (int)world,
                       (int)rank, (long long)offset_elems, n4, scale);
  } else {
    PeerPtrs pp;
    memset(&pp, 0, sizeof(pp));
    for (int i = 0; i < world; i++) {
      PESR_CHECK_ARG(peer_ptrs[i] != 0 && peer_ptrs[i] % 16 == 0, "allreduce_p2p: peer pointer %d is null or unaligned", i);
      pp.p[i] = reinterpret_cast<float*>(peer_ptrs[i]);
    }
    switch (world) {
      case 2: e = launch_cluster(allreduce_p2p_kernel<2, 4>, ctas, stream, pp, pads, epoch, (int)world, (int)rank, (long long)offset_elems, n4, scale); break;
      case 4: e = launch_cluster(allreduce_p2p_kernel<4, 2>, ctas, stream, pp, pads, epoch, (int)world, (int)rank, (long long)offset_elems, n4, scale); break;
      case 8: e = launch_cluster(allreduce_p2p_kernel<8, 1>, ctas, stream, pp, pads, epoch, (int)world, (int)rank, (long long)offset_elems, n4, scale); break;
      default: e = launch_cluster(allreduce_p2p_kernel<0, 1>, ctas, stream, pp, pads, epoch, (int)world, (int)rank, (long long)offset_elems, n4, scale); break;
    }
  }
  if (e != cudaSuccess) {
    set_error("allreduce_p2p: launch failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  count_launch();
  PESR_CHECK_LAUNCH("allreduce_p2p");
  return 0;
}
