/*
 * Bring-up / profiling hooks of pesr_b200.  NOT part of the product ABI: they are compiled only into
 * pesr_b200/libpesr_b200_debug.so (tools/build_debug.sh, -DPESR_DEBUG_HOOKS), which the tools under tools/ load
 * through PESR_B200_LIB.  The product library (include/pesr_b200.h) exports none of them and its kernels carry no
 * timeline code.
 */
#ifndef PESR_B200_DEBUG_H
#define PESR_B200_DEBUG_H

#include "pesr_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Device buffer (64 x uint64 + 2 per CTA) that block 0 of each pesr_conv_igemm launch fills with clock64 stamps. */
void pesr_debug_timeline(void* buf);
void pesr_debug_wgrad_timeline(void* buf);   /* same for pesr_conv_wgrad */
/* Override the MN-major smem descriptor strides of pesr_conv_wgrad (0 = built-in). */
void pesr_debug_wgrad_desc(int lbo_bytes, int sbo_bytes);
/* Microbenchmark: issue / completion cycles of iters*4 back-to-back tcgen05.mma (M=128 or pair 256, N=n). */
int pesr_debug_mma_rate(int32_t n, int32_t iters, int32_t stages, int32_t pair_and_major, int32_t blocks,
                        unsigned long long* out_dev, void* stream);
/* n_ctas CTAs that each pin one SM (200 KB of shared memory) and spin for usec microseconds on `stream`: what do a
 * few unavailable SMs (a concurrent NCCL all-reduce) cost the one-CTA-per-SM kernels? (tools/sm_hog_probe.py) */
int pesr_debug_sm_hog(int32_t n_ctas, int64_t usec, void* stream);

#ifdef __cplusplus
}
#endif
#endif
