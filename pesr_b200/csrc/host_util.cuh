// Host-side helpers shared by the launchers: error reporting, launch counting, TMA tensor-map cache.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/pesr_b200.h"

namespace pesr {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
void note_weight_write(cudaStream_t stream);
bool weights_settled(cudaStream_t stream);
bool pdl_enabled();
void set_pdl(int on);
int num_sms();        // SMs available to the persistent tensor-core kernels (device count - reserved)
int device_sms();     // SMs of the current device
void set_reserved_sms(int n);
int conv_smem_budget();

// Optional per-launch profiling of the tensor-core kernels (bench.py's roofline leg): when enabled, the
// launchers bracket each launch with CUDA events on the launch stream and record its algorithmic FLOPs.
// kind 0 = conv_igemm (fprop/dgrad), 1 = conv_wgrad.
bool profiling_enabled();
void profile_begin(int kind, double flops, cudaStream_t stream);
void profile_end(int kind, cudaStream_t stream);

// Encodes (or fetches from the cache) a tiled, 128B-swizzled tensor map over a 16-bit tensor.
// dims/strides are innermost-first; strides_bytes has rank-1 entries (dims 1..rank-1).
// Returns 0 or a PESR_E_* code.
int get_tensor_map(CUtensorMap* out, const void* base, int dtype, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box);
// Same with the element type (PESR_DT_F16 / PESR_DT_BF16 / 2 = fp32) and the swizzle (0 none, 1 32B, 2 64B, 3 128B)
// spelled out: used for the TMA-store maps of the staged epilogue.
int get_tensor_map_ex(CUtensorMap* out, const void* base, int elem, int swizzle, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box);

// Launch with programmatic stream serialization: the grid may become resident while its predecessor in the stream is
// still draining, which hides the launch latency between the ~600 short dependent kernels of a training step.
// EVERY kernel launched through this helper starts with griddep_wait() (common.cuh) before touching global memory.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define PESR_CHECK_ARG(cond, ...)      \
  do {                                 \
    if (!(cond)) {                     \
      pesr::set_error(__VA_ARGS__);    \
      return PESR_E_ARG;               \
    }                                  \
  } while (0)

#define PESR_CHECK_LAUNCH(name)                                                 \
  do {                                                                          \
    cudaError_t e__ = cudaGetLastError();                                       \
    if (e__ != cudaSuccess) {                                                   \
      pesr::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return (int)e__;                                                          \
    }                                                                           \
  } while (0)

}  // namespace pesr
