/*
 * pesr_b200 C-ABI: the drop-in boundary of the B200-native PESR hot path.
 *
 * The reference (thangvubk/PESR) has no FFI of its own: its hot path is a set of torch.nn modules
 * (model/basic.py, model/pesr.py, model/vgg.py, model/focal_loss.py) driven by the step bodies in
 * train.py:164-176, train.py:202-259 and test.py:101-112.  Every torch/cuDNN/cuBLAS/ATen call those
 * modules make on the GPU is replaced by one of the entry points below.  The Python host side
 * (pesr_b200/) binds them with ctypes; INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - plain C: raw device pointers, sizes, a cudaStream_t (passed as void*) last; no torch types.
 *  - every function returns 0 on success, a positive cudaError_t, or a negative PESR_E_* argument error;
 *    pesr_last_error() returns a human-readable message for the calling thread's last failure.
 *  - launches are asynchronous on the given stream; the library never allocates user-visible memory,
 *    workspaces are caller-provided.
 *  - activations are NHWC, 16-bit (dtype 0 = fp16, 1 = bf16) with fp32 accumulation; parameters and
 *    parameter gradients stay fp32 in the reference's own layouts (OIHW conv weights, [out,in] linear).
 */
#ifndef PESR_B200_H
#define PESR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PESR_E_ARG (-1)      /* bad argument / unsupported shape */
#define PESR_E_DRIVER (-2)   /* cuTensorMapEncodeTiled unavailable or failed */
#define PESR_E_WORKSPACE (-3) /* workspace too small */

#define PESR_DT_F16 0
#define PESR_DT_BF16 1

#define PESR_ACT_NONE 0
#define PESR_ACT_RELU 1   /* nn.ReLU, model/pesr.py:11, torchvision vgg19 */
#define PESR_ACT_LRELU 2  /* nn.LeakyReLU(0.2), model/pesr.py:47 */

#define PESR_OUT_NORMAL 0     /* out pixel (n, h*sy+oy, w*sx+ox), channel coff+q */
#define PESR_OUT_SHUFFLE2 1   /* nn.PixelShuffle(2) fused (model/basic.py:57,59), packed channel order (i,j,c) */
#define PESR_OUT_UNSHUFFLE2 2 /* inverse of PixelShuffle(2): the backward of the above */

#define PESR_MAX_TAPS 9
#define PESR_MAX_SRC 4

const char* pesr_last_error(void);
int pesr_version(void);
/* Number of kernels this library has launched since the last reset (bench.py's gpu_launches). */
long long pesr_launch_count(int reset);
/* Per-launch profiling of the tensor-core kernels: when enabled every pesr_conv_igemm (kind 0) and
 * pesr_conv_wgrad (kind 1) launch is bracketed by CUDA events on its stream.  pesr_profile_read
 * synchronises the device and returns the summed kernel time, launch count and algorithmic FLOPs
 * (dense convention 2*M*N*K, padding taps counted), then clears the records of that kind. */
void pesr_profile_enable(int on);
int pesr_profile_read(int kind, double* total_ms, long long* launches, double* flops);
/* sizeof(pesr_conv_desc) for which == 0, sizeof(pesr_wgrad_desc) for 1 (binding self-check). */
int pesr_sizeof(int which);

/* ------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution, forward and backward-data.  Replaces nn.Conv2d forward/dgrad
 * (model/basic.py:4-7 `Conv`; call sites model/pesr.py:20,23, model/basic.py:41,56-60,
 * model/pesr.py:54,63 and torchvision vgg19.features used by model/vgg.py:8-10).
 *
 * GEMM view: M = nb*h*w output pixels (tiles of tile_h x tile_w = 128 pixels), N = cout (tiles of
 * block_n), K = ntaps * cin.  For tap t the A operand is the activation box whose top-left input
 * pixel is (h0 + tap_dh[t], w0 + tap_dw[t]) of source tensor tap_src[t]; out-of-bounds pixels read
 * as zero (TMA OOB fill == the conv's zero padding).  The B operand is rows
 * [tap_widx[t]*cout + n0, +block_n) of the packed K-major weight matrix [w_rows][cin].
 * Backward-data is the same kernel with flipped/transposed packed weights (pesr_pack_weights).
 * Stride-2 convolutions pass the four parity planes of the input as sources 0..3.
 * The same kernel runs the Discriminator's Linear(73728 -> 1024) as a split-K GEMM (ksplit) and its
 * backward-data with the weight matrix read MN-major in place (b_mn_major).
 *
 * Epilogue (per output element, in this order):
 *   v = alpha * [alpha_dev] * (acc + bias[q]);  v += res32;  v += res16;  v = act(v);
 *   v *= act'(mask16)   (mask_mode 1: relu', 2: lrelu'(0.2));
 *   out32 = v (normal addressing);  out16 = cvt(v) (out_mode addressing).
 * ------------------------------------------------------------------------------------------------ */
typedef struct pesr_conv_desc {
  int32_t dtype;            /* PESR_DT_* for src / weights / out16 / res16 / mask16 */
  int32_t nb, h, w;         /* GEMM-M pixel grid */
  int32_t cin, cout;        /* cin multiple of 64 (per tap), cout multiple of block_n */
  int32_t block_n;          /* 32, 64, 128 or 256 */
  int32_t tile_h, tile_w;   /* tile_n * tile_h * tile_w == 128 (tile_n: see the end of the struct) */
  int32_t ntaps;
  int8_t tap_dh[PESR_MAX_TAPS], tap_dw[PESR_MAX_TAPS], tap_src[PESR_MAX_TAPS], tap_widx[PESR_MAX_TAPS];
  /* sources: NHWC views; strides in elements, channel stride 1 */
  const void* src[PESR_MAX_SRC];
  int32_t src_h[PESR_MAX_SRC], src_w[PESR_MAX_SRC];
  int64_t src_sn[PESR_MAX_SRC], src_sh[PESR_MAX_SRC], src_sw[PESR_MAX_SRC];
  int32_t nsrc;
  const void* wpacked;      /* [w_rows][cin], 16-bit, K-major */
  int32_t w_rows;
  /* epilogue */
  const float* bias;        /* [cout] or NULL */
  float alpha;
  const float* alpha_dev;   /* optional device scalar multiplied into alpha */
  const float* res32; int32_t ld_res32;
  const void* res16;  int32_t ld_res16;
  int32_t act;
  const void* mask16; int32_t ld_mask16; int32_t mask_mode;
  float* out32; int32_t ld_out32;
  void* out16;  int32_t ld_out16;  /* channels per pixel of the out16 tensor */
  int32_t out_mode;
  int32_t out_h, out_w;     /* out16 pixel grid (NORMAL mode; 0 = same as h, w) */
  int32_t out_sy, out_sx, out_oy, out_ox, out_coff;
  int32_t ps_c;             /* SHUFFLE2: channels of the shuffled tensor (cout/4) */
  int32_t aux_mode;         /* 0: res32/res16/mask16/out32 are indexed by the GEMM pixel; 1: by the NORMAL-mode
                               output pixel (n, h*sy+oy, w*sx+ox) of the out_h x out_w grid (stride-2 dgrad) */
  int32_t ksplit;           /* > 1: split K across CTAs; raw fp32 partials go to out32 + split*split_stride32 */
  int32_t b_mn_major;       /* 1: wpacked is [K = cin][N = cout] (N contiguous), one tap: y = x * W without a transpose */
  int64_t split_stride32;   /* elements between the partial outputs of successive K splits */
  int32_t tile_n;           /* images per pixel tile (0 = 1): tile_n * tile_h * tile_w == 128.  Small feature maps (24x24,
                               12x12) tile without padding rows as e.g. 2 x 8 x 8 or 8 x 4 x 4 */
  int32_t reserved0;
  double* bn_sums;          /* optional, light epilogue only (16-bit NHWC output, no residual / mask): 2*cout doubles that
                               receive += sum and += sum of squares of the ROUNDED 16-bit outputs per channel - the
                               train-mode BatchNorm statistics of model/basic.py:29 without a second pass over y
                               (consumed by pesr_bn_lrelu_fwd(sums_ws = bn_sums, ...)) */
  int32_t ncls;             /* > 1: the launch covers ncls sub-problems ("classes") of the same h x w pixel grid and the
                               same weights / sources / epilogue tensors: class c owns the next cls_ntaps[c] entries of the
                               tap arrays and writes to output-grid offset (cls_oy[c], cls_ox[c]) instead of (out_oy, out_ox).
                               The four parity classes of a stride-2 backward-data (1 + 2 + 2 + 4 taps) run as one launch
                               this way.  Needs ksplit <= 1, K-major weights, NORMAL output addressing. */
  int8_t cls_ntaps[4], cls_oy[4], cls_ox[4];
} pesr_conv_desc;

int pesr_conv_igemm(const pesr_conv_desc* d, void* stream);

/* Kernel-selection options of pesr_conv_igemm (process-wide; the defaults are the measured-fastest choices).  The
 * parity tests use them to force every kernel variant onto small shapes.  Bring-up probes (timelines, MMA-rate and
 * SM-occupancy microbenchmarks) are NOT part of this library: see include/pesr_b200_debug.h / tools/build_debug.sh. */
#define PESR_OPT_PAIR_MODE 0        /* 0 never use the CTA-pair (cta_group::2) kernel, 1 where it wins (default), 2 whenever legal */
#define PESR_OPT_SUB_STAGES 1       /* 0 one k-block per pipeline stage, 1 multi-k-block / halo stages (default), 2 also for N = 256 */
#define PESR_OPT_PDL 2              /* programmatic dependent launch: 0 off, 1 on (default; PESR_NO_PDL=1 in the environment disables) */
#define PESR_OPT_STAGED_EPILOGUE 3  /* residual epilogue through shared memory (pair kernel): 0 off, 1 on (default) */
#define PESR_OPT_SPECIALISED_EPILOGUE 4 /* compile-time specialised epilogues: 0 generic kernel only, 1 on (default) */
#define PESR_OPT_RESERVE_SMS 5      /* SMs the persistent tensor-core kernels leave free (their grids are sized to #SM - value;
                                       default 0, PESR_RESERVE_SMS in the environment).  Data-parallel training reserves the SMs
                                       NCCL's all-reduce CTAs occupy: with a full-machine static tile schedule ONE SM taken by a
                                       concurrent kernel sends a CTA to a second wave (profiles/r02_sm_hog_probe.txt). */
#define PESR_OPT_RESIDENT_WEIGHTS 6  /* narrow 3x3 layers keep their whole packed weight tensor in shared memory (pesr_conv_igemm):
                                       0 off, 1 where it wins: 128-wide column tiles (default; PESR_NO_WRES=1 disables), 2 wherever legal */
int pesr_set_option(int option, int value);

/* ------------------------------------------------------------------------------------------------
 * Backward-filter as a split-K GEMM with both operands MN-major (pixels are K).
 *   part[split][tap][m][n] = sum over the split's pixels p of  a[p][m0+m] * b[p (+) tap][n0+n]
 * a = dy (M = cout), b = x (N = cin, shifted by the tap with zero fill).  Replaces cuDNN
 * backward-filter for the call sites listed above.  pesr_wgrad_reduce sums the splits, applies
 * `scale` (and 1/scale_dev) and writes/accumulates the fp32 OIHW gradient.
 * ------------------------------------------------------------------------------------------------ */
typedef struct pesr_wgrad_desc {
  int32_t dtype;
  int32_t nb, h, w;         /* pixel grid of `a` */
  int32_t m_total, n_total; /* channels of a / b used; m_total multiple of 64 (tiles of 128, zero-filled past the end); n_total multiple of block_n */
  int32_t block_m;          /* 128 */
  int32_t block_n;          /* 64, 128 or 256 */
  int32_t ntaps;
  int8_t tap_dh[PESR_MAX_TAPS], tap_dw[PESR_MAX_TAPS], tap_src[PESR_MAX_TAPS];
  const void* a; int32_t a_c;   /* NHWC [nb][h][w][a_c] */
  const void* b[PESR_MAX_SRC];
  int32_t b_h[PESR_MAX_SRC], b_w[PESR_MAX_SRC];
  int64_t b_sn[PESR_MAX_SRC], b_sh[PESR_MAX_SRC], b_sw[PESR_MAX_SRC];
  int32_t nsrc;
  int32_t splits;           /* split-K factor (0 = choose) */
  float* partials;          /* [splits][ntaps][m_total][n_total] fp32 workspace */
  int64_t partials_elems;   /* capacity of `partials` in floats */
  float out_mul;            /* the sums are stored multiplied by out_mul / (*out_div_dev): with splits == 1 `partials` can be */
  const float* out_div_dev; /* the gradient itself (0 / NULL = 1; the Linear(73728 -> 1024) weight gradient is written this way) */
} pesr_wgrad_desc;

/* Returns the number of splits actually used in *splits_out (may be NULL). */
int pesr_conv_wgrad(const pesr_wgrad_desc* d, int32_t* splits_out, void* stream);

#define PESR_WMAP_OIHW 0      /* partial (tap, m=o, n=i)        -> grad[o][i][tap]                         */
#define PESR_WMAP_OIHW_PS 1   /* packed shuffle order m=(i,j,c) -> grad[c*4+i*2+j][n][tap]                */
#define PESR_WMAP_COL_IN 2    /* Cin<=7 im2col conv: (m=o, n=tap*ci_n+ci) -> grad[o][ci][tap]             */
#define PESR_WMAP_COL_OUT 3   /* Cout<=3 col2im conv: (m=i, n=tap*co_n+o) -> grad[o][i][tap]              */
/* grad = scale / (*div_dev if given) * sum_splits(partials), scattered per map_mode. */
int pesr_wgrad_reduce(const float* partials, int32_t splits, int32_t ntaps, int32_t m_total, int32_t n_total,
                      int32_t map_mode, int32_t co, int32_t ci, float scale, const float* div_dev,
                      int32_t accumulate, float* grad_oihw, void* stream);

/* pesr_wgrad_reduce for a 3x3 OIHW / OIHW_PS gradient fused with the bias gradient of the same layer (the column sums
 * of dY, pesr_colsum16) in one launch: autograd's bias term of cuDNN backward-filter (model/basic.py:4-7).
 * bias_grad must be zero on entry and receives bias_mul / *inv_scale_dev * sum_pixels dY; zero_next (may be NULL) is a
 * buffer of zero_n floats that the call clears for the NEXT call of a chain to accumulate into. */
int pesr_wgrad_reduce_bias(const float* partials, int32_t splits, int32_t ntaps, int32_t m_total, int32_t n_total,
                           int32_t map_mode, int32_t co, int32_t ci, float scale, const float* inv_scale_dev,
                           int32_t accumulate, float* grad_oihw, const void* dy16, int64_t npix, int32_t c, int32_t ldc,
                           float bias_mul, int32_t dtype, float* bias_grad, float* zero_next, int32_t zero_n,
                           void* stream);

/* Weight packing: fp32 OIHW parameter -> 16-bit K-major GEMM operand (tap = ky*3+kx; tapf = 8-tap).
 *  mode 0 fprop     : out[tap][o][i]                        (rows = 9*co, K = ci)
 *  mode 1 dgrad     : out[tap][i][o] = w[o][i][tapf]        (rows = 9*ci, K = co)
 *  mode 2 fprop, PixelShuffle(2) output order: row o' = (i*2+j)*C + c  <- o = c*4 + i*2 + j
 *  mode 3 dgrad of mode 2: out[tap][i][o'] = w[o][i][tapf], K index o' permuted as in mode 2
 *  mode 4 im2col fprop for tiny Cin : out[o][tap*ci + i], K zero-padded to pad_to      (rows = co)
 *  mode 5 col2im fprop for tiny Cout: out[tap*co + o][i], rows zero-padded to pad_to   (K = ci)
 *  mode 6 dgrad of mode 4 (a col2im GEMM): out[tap*ci + i][o], rows zero-padded to pad_to (K = co)
 *  mode 7 dgrad of mode 5 (an im2col GEMM): out[i][tap*co + o], K zero-padded to pad_to   (rows = ci)
 * ksize 1 is accepted for modes 0/1 (tap == 0).
 */
int pesr_pack_weights(const float* w_oihw, int32_t co, int32_t ci, int32_t ksize, int32_t mode, int32_t pad_to,
                      int32_t dtype, void* out, void* stream);

/* All packs of one network in ONE launch.  jobs_host: njobs rows of 8 int64 {src fp32 OIHW, dst 16-bit, co, ci,
 * ksize, mode, pad_to, dst2}; jobs_dev: caller-owned device scratch of njobs*64 bytes, (re)filled when upload != 0.
 * dst2 (0 = none): for a 3x3 mode-0 / mode-2 job with co % 64 == 0 and ci % 32 == 0, the backward-data layout of the
 * same weight (mode 1 / mode 3) written from the same read of the fp32 tensor. */
int pesr_pack_weights_multi(const int64_t* jobs_host, int32_t njobs, void* jobs_dev, int32_t upload, int32_t dtype,
                            void* stream);

/* ------------------------------------------------------------------------------------------------
 * Layout / edge kernels.  These are HBM-bound; each is a single coalesced pass.
 * ------------------------------------------------------------------------------------------------ */

/* 3-channel NCHW fp32 image -> im2col matrix [nb*h*w][64] (16-bit), column tap*3+c holds
 * affine(src)[c] at pixel p + sgn*(ky-1, kx-1), zero outside the image; columns 27..63 are zero.
 * affine: v = A(3x3, row-major [out][in]) * src + b, applied to in-bounds pixels only (this is how
 * MeanShift, model/basic.py:9-17, composes with the zero padding of the following conv); A/b NULL =
 * identity.  pad_affine = 1 instead applies the affine to the padded zeros too (out-of-image taps read b): a
 * constant shift of the conv input, used to centre the 0..255 image ahead of the Discriminator's first
 * conv + BatchNorm (which is invariant to it).  pad_affine bit 1 (value 2): columns 32..63 of `col` already hold zeros
 * (a buffer that is zeroed once and reused) and are not written, which halves the bytes stored.
 * pad_affine bit 2 (value 4): src_nchw is really a uint8 HWC image batch [nb][h][w][3] (utils.imgs_to_tensors,
 * utils.py:20-25, fused into the first kernel of the inference path).
 * mul_dev: optional device scalar multiplied into the result.
 * Feeds the Cin=3 convs (model/pesr.py:23,54; vgg19.features[0]) and, with sgn=-1, the backward of
 * the Cout=3 conv (model/basic.py:60). */
int pesr_im2col3(const float* src_nchw, int32_t nb, int32_t h, int32_t w, const float* affine_a,
                 const float* affine_b, const float* mul_dev, int32_t sgn, int32_t pad_affine, int32_t dtype, void* col,
                 void* stream);

/* col2im for a 3-channel output: out[n][c][h][w] = affine( sum_tap z[p + sgn*(ky-1,kx-1)][tap*3+c] * mul + bias[c] ).
 * z is fp32 [nb*h*w][ldz].  pre (optional) receives the value before the affine (needed by the
 * MeanShift weight gradient).  mul = mul_host / (*div_dev if given). */
int pesr_col2im3(const float* z, int32_t ldz, int32_t nb, int32_t h, int32_t w, const float* bias,
                 const float* affine_a, const float* affine_b, float mul_host, const float* div_dev, int32_t sgn,
                 float* pre_nchw, float* out_nchw, void* stream);

/* pesr_col2im3 through a shared-memory tile (8 x 32 pixels + halo per block: every z row is fetched once instead of
 * nine times), with an optional fused uint8 HWC store out_u8[n][h][w][3] = uint8(rint(clamp(out, 0, 255))) -- the
 * clip / round-half-even / transpose of utils.tensors_to_imgs (utils.py:13-18).  out_nchw or out_u8 may be NULL (not both).
 * z must be 16-byte aligned with ldz %% 4 == 0 and ldz >= 28. */
int pesr_col2im3_tiled(const float* z, int32_t ldz, int32_t nb, int32_t h, int32_t w, const float* bias,
                       const float* affine_a, const float* affine_b, float mul_host, const float* div_dev, int32_t sgn,
                       float* pre_nchw, float* out_nchw, uint8_t* out_u8, void* stream);

/* Split-precision operands (pesr_b200/engine_g_split.py): v = act(src * mul * [*mul_dev]) * act'(mask_hi + mask_lo)
 * (mask_mode 1: relu', 2: lrelu'(0.2); masks optional); hi = round16(v), lo = round16(v - hi).  hi + lo carries 22 (fp16)
 * significant bits, and hi*hi + lo*hi + hi*lo over three passes of pesr_conv_igemm (fp32 accumulation into out32 via
 * res32) reproduces an fp32-grade product.  n must be a multiple of 4; either output may be NULL.
 * pesr_im2col3 with pad_affine bit 3 (value 8) emits the low part of the im2col matrix in the same way. */
int pesr_split16(const float* src, int64_t n, int32_t act, const void* mask_hi, const void* mask_lo, int32_t mask_mode,
                 float mul, const float* mul_dev, int32_t dtype, void* hi, void* lo, void* stream);

/* fp32 NHWC helpers of the split-precision Discriminator / VGG schedules (activations between the three-pass
 * convolutions are fp32; BatchNorm, LeakyReLU / ReLU and max-pool act on fp32 and only conv operands are split):
 *  pesr_colmoments32: sums[0][ch] += sum_p a[p][ch]; sums[1][ch] += sum_p a[p][ch] * (b ? b : a)[p][ch]   (fp64, [2][c])
 *  pesr_affine_split: v = ka[ch]*a + kb[ch]*b + kc[ch] (per-channel vectors; ka, kb/b, kc optional), v = act(v),
 *    v *= act'(mask) (mask = 16-bit hi + lo pair or an fp32 tensor; mask_mode 1 relu', 2 lrelu'(0.2)); outputs (each
 *    optional): out32 = v, hi = round16(v), lo = round16(v - hi).  BatchNorm forward / backward are this with the
 *    coefficients derived from the moments.
 *  pesr_maxpool2_f32_fwd / _bwd: as pesr_maxpool2_* on fp32. */
int pesr_colmoments32(const float* a, const float* b, int64_t npix, int32_t c, double* sums, void* stream);
int pesr_affine_split(const float* a, const float* b, int64_t npix, int32_t c, const float* ka, const float* kb,
                      const float* kc, int32_t act, const void* mask_hi, const void* mask_lo, const float* mask32,
                      int32_t mask_mode, int32_t dtype, float* out32, void* hi, void* lo, void* stream);
int pesr_maxpool2_f32_fwd(const float* x, int32_t nb, int32_t h, int32_t w, int32_t c, float* y, void* stream);
int pesr_maxpool2_f32_bwd(const float* x, const float* dy, int32_t nb, int32_t h, int32_t w, int32_t c, int32_t relu_mask,
                          float* dx, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Gradient all-reduce (average) over NVLink / NVSwitch peer memory: the reduction nn.DataParallel performs on GPU 0
 * (train.py:114-118), as a two-shot kernel on `ctas` CTAs over a range of a SYMMETRIC fp32 buffer (the same allocation
 * mapped into every process of the node).  peer_ptrs[world]: this process's addresses of every rank's buffer (host
 * array); multicast_ptr != 0: NVSwitch multicast address of the buffer (multimem.ld_reduce / multimem.st are used and
 * peer_ptrs may be NULL).  Element range [offset_elems, offset_elems + count) (multiples of 4) ends up holding
 * scale * sum over ranks on every rank, bit-identical.  pad_ptrs[world]: every rank's signal pad (>= 1280 zero-initialised
 * bytes of symmetric memory); the kernel itself makes the ranks meet before and after the reduction, with `epoch` (a
 * counter the caller increments per call, identical on all ranks) as the flag value.  pad_ptrs NULL: the caller orders
 * the ranks.  One launch, one thread-block cluster of `ctas` (<= 8) CTAs.
 * ------------------------------------------------------------------------------------------------ */
#define PESR_MAX_PEERS 16
int pesr_allreduce_p2p(const uint64_t* peer_ptrs, const uint64_t* pad_ptrs, int32_t world, int32_t rank,
                       uint64_t multicast_ptr, int64_t offset_elems, int64_t count, float scale, int32_t ctas,
                       uint32_t epoch, void* stream);

/* MeanShift as a stand-alone op (model/basic.py:9-17): out[n][o][p] = sum_i w9[o*3+i] * x[n][i][p] + b3[o] on
 * [nb][3][hw] fp32 tensors (b3 may be NULL). */
int pesr_mean_shift(const float* x, int32_t nb, int64_t hw, const float* w9, const float* b3, float* out, void* stream);

/* NCHW fp32 <-> NHWC 16-bit (generic; used at module boundaries and by the tests). */
int pesr_nchw32_to_nhwc16(const float* src, int32_t nb, int32_t c, int32_t h, int32_t w, int32_t ldc,
                          const float* mul_dev, int32_t dtype, void* dst, void* stream);
int pesr_nhwc16_to_nchw32(const void* src, int32_t nb, int32_t c, int32_t h, int32_t w, int32_t ldc, float mul_host,
                          const float* div_dev, int32_t dtype, float* dst, void* stream);

/* out[c] (+)= mul * sum over pixels of x[p][c]   (bias gradients; x is 16-bit [npix][ldc]) */
int pesr_colsum16(const void* x, int64_t npix, int32_t c, int32_t ldc, float mul_host, const float* div_dev,
                  int32_t accumulate, int32_t dtype, float* out, void* stream);

/* Dynamic power-of-two gradient scale: ws[0] = running max|x| (as uint bits), ws[1] = scale,
 * ws[2] = 1/scale, with scale = 2^k chosen so that max|x|*scale lies in (target/2, target].
 * Two launches (max-reduce, finalize); no host sync. */
int pesr_amax_scale(const float* x, int64_t n, float target, float* ws3, void* stream);

/* sums[0..8] = sum_p a[o][p]*b[i][p] (row-major [o][i]), sums[9..11] = sum_p a[o][p]; a, b are
 * [nb][3][hw] fp32.  The MeanShift weight/bias gradients (model/basic.py:17 leaves them trainable). */
int pesr_moments3(const float* a, const float* b, int32_t nb, int64_t hw, float* sums12, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Inference edges (test.py:101-112, utils.py:13-25).
 * pesr_blend_x8_to_u8: out = alpha*perc + (1-alpha)*mean_{i<8} T_i^-1(ens_i) for one image [3][h][w] (fp32, 0..255);
 *   variant i: bit0 = flip W, bit1 = flip H, bit2 = transpose (test.py:45-74); ens = the 8 generator outputs back to
 *   back, transposed variants stored [3][w][h]; n_ens = 0 skips the ensemble (alpha = 1 path).  out32 (optional)
 *   gets the blended fp32 image, out8 (optional) the HWC uint8 image after clip(0,255) and round-half-even.
 * pesr_u8hwc_to_f32nchw: HWC uint8 image -> [3][h][w] fp32 (utils.imgs_to_tensors).
 * ------------------------------------------------------------------------------------------------ */
int pesr_blend_x8_to_u8(const float* perc, const float* ens, int32_t h, int32_t w, float alpha, int32_t n_ens,
                        float* out32, uint8_t* out8, void* stream);
int pesr_u8hwc_to_f32nchw(const uint8_t* src, int32_t h, int32_t w, float* dst, void* stream);
int pesr_u8hwc_to_f32nchw_batch(const uint8_t* src, int32_t nb, int32_t h, int32_t w, float* dst, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Validation metric and input pipeline (train.py:281-295, utils.py:10-11,27-41, data.py:64-126).
 * pesr_psnr_y_sse: sse[n] += sum over pixels of (Y(a) - Y(b))^2 for [nb][3][hw] fp32 images in 0..255, where
 *   Y(v) = rint(clamp(rgb2y(rint(clamp(v, 0, 255))), 0, 255)) exactly as utils.compute_PSNR builds it
 *   (rgb2y = (65.738 r + 129.057 g + 25.064 b) / 256 + 16, evaluated in exact integer arithmetic); the sum is a 64-bit
 *   integer, so PSNR = 20 log10(255 / sqrt(sse / hw)) is bit-reproducible.  The caller zeroes sse.
 * pesr_gather_patches: one launch builds a training batch from a device-resident uint8 HWC image cache: sample s =
 *   LR crop [y, y+patch) x [x, x+patch) of its image and the aligned HR crop (scale * y, scale * x, scale * patch),
 *   augmented as data.py:86-104 (k bit 2: transpose, then bit 1: vertical flip, then bit 0: horizontal flip), written
 *   as NCHW fp32 lr [nb][3][patch][patch] and hr [nb][3][scale*patch][scale*patch].  table_dev: nb rows of 8 int64
 *   {lr image ptr, hr image ptr, lr width, hr width, y, x, k, 0}.
 * ------------------------------------------------------------------------------------------------ */
int pesr_psnr_y_sse(const float* a, const float* b, int32_t nb, int64_t hw, unsigned long long* sse, void* stream);
int pesr_gather_patches(const int64_t* table_dev, int32_t nb, int32_t patch, int32_t scale, float* lr, float* hr,
                        void* stream);

/* ------------------------------------------------------------------------------------------------
 * Loss reductions with the gradient produced in the same pass (train.py:131-140), all fp32.
 *   *loss receives the scalar (it is zeroed by the call); grad (optional) receives d loss / d a.
 * ------------------------------------------------------------------------------------------------ */
int pesr_loss_l1(const float* a, const float* b, int64_t n, float* loss, float* grad, void* stream);  /* nn.L1Loss, train.py:131 */
int pesr_loss_mse(const float* a, const float* b, int64_t n, float* loss, float* grad, void* stream); /* F.mse_loss, train.py:136 */
/* TV (train.py:137-140): SUM of |dx| + |dy| over y viewed as [planes][h][w]. */
int pesr_loss_tv(const float* y, int64_t planes, int32_t h, int32_t w, float* loss, float* grad, void* stream);
/* GAN losses on n logits: x = sign_a*a + sign_b*b (b may be NULL), target t in {0,1}.
 *   mode 0: BCE-with-logits mean (nn.BCEWithLogitsLoss, train.py:132,211,213)
 *   mode 1: FocalLoss(gamma) (model/focal_loss.py:9-13) with torch-0.4 gradient semantics (through the weight)
 *   mode 2: FocalLoss(gamma) with the weight detached (what torch >= 1.0 would compute)
 * grad_a / grad_b (optional) receive d loss / d a, d loss / d b. */
int pesr_loss_gan(const float* a, const float* b, int32_t n, float sign_a, float sign_b, float target, int32_t mode,
                  float gamma, float* loss, float* grad_a, float* grad_b, void* stream);

/* Multi-tensor Adam (torch.optim.Adam semantics, train.py:124-126; no weight decay, no amsgrad).
 * table_dev: nchunks rows of {p, g, m, v (device addresses), n} as int64; g is multiplied by grad_mul first. */
int pesr_adam_multi(const int64_t* table_dev, int32_t nchunks, float lr, float beta1, float beta2, float eps,
                    int32_t step, float grad_mul, void* stream);
/* Same with the learning rate and the (already incremented) step count read from device memory at run time, so that a
 * captured CUDA graph of the training step replays with the right bias corrections (torch.optim.Adam(capturable=True)). */
int pesr_adam_multi_dev(const int64_t* table_dev, int32_t nchunks, const float* lr_dev, float beta1, float beta2,
                        float eps, const int32_t* step_dev, float grad_mul, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Discriminator / VGG layer kernels (HBM-bound).
 * ------------------------------------------------------------------------------------------------ */
/* Train-mode BatchNorm2d + LeakyReLU (model/basic.py:29-30, model/pesr.py:47) on 16-bit NHWC tensors.
 * GROUPS: a tensor may hold `groups` independent batches back to back ([groups][npix][c]); each group is normalised
 * with its own batch statistics, exactly as `groups` separate module calls would (the Discriminator is called twice
 * per phase, on hr and on sr: train.py:205-208, 237-238 -- both calls run as one launch per layer).  Per-group
 * quantities are laid out [groups][...]: sums_ws [groups][2][c] doubles, mean / rstd [groups][c] floats.
 *
 * pesr_bn_reduce: sums_ws[g][0][ch] += sum y, sums_ws[g][1][ch] += sum y^2 (fp64 accumulation).  zero_first != 0 clears
 *   sums_ws first; otherwise the caller provides zeroed sums (one memset for all layers of a pass), or sums already
 *   accumulated by the producing convolution (pesr_conv_desc.bn_sums, groups == 1).
 * pesr_bn_lrelu_fwd: a = LeakyReLU_slope(gamma * (y - mean_g) * rstd_g + beta).  sums_ws != NULL (train mode): mean_g and
 *   rstd_g = 1/sqrt(biased var + eps) are derived from the sums inside the kernel and stored to mean / rstd for backward;
 *   running_mean / running_var (optional) are updated with `momentum` (unbiased variance) once per group, in group
 *   order, and num_batches_tracked += groups.  running_mean_shift (optional, per channel) is added to the batch mean in
 *   the running_mean update only (y was produced from a constant-shifted input).  sums_ws == NULL (eval mode): mean and
 *   rstd are inputs.
 * pesr_bn_lrelu_bwd: given dz = dL/d(bn output) (LeakyReLU' already applied by the producer):
 *   dy = gamma*rstd_g*(dz - mean(dz) - xhat*mean(dz*xhat)) per group; dgamma = sum_groups sum(dz*xhat), dbeta = sum_groups
 *   sum(dz), multiplied by grad_mul / (*grad_div_dev); accumulate != 0 adds them to dgamma / dbeta (a further backward
 *   pass of the same optimiser step).  sums_ws as in pesr_bn_reduce (zero_first). */
int pesr_bn_reduce(const void* y16, int64_t npix_per_group, int32_t c, int32_t groups, double* sums_ws, int32_t zero_first,
                   int32_t dtype, void* stream);
int pesr_bn_lrelu_fwd(const void* y16, int64_t npix_per_group, int32_t c, int32_t groups, const double* sums_ws, float eps,
                      float momentum, float* mean, float* rstd, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, int64_t* num_batches_tracked,
                      const float* running_mean_shift, float slope, int32_t dtype, void* a16, void* stream);
int pesr_bn_lrelu_bwd(const void* dz16, const void* y16, int64_t npix_per_group, int32_t c, int32_t groups,
                      const float* mean, const float* rstd, const float* gamma, double* sums_ws, int32_t zero_first,
                      float grad_mul, const float* grad_div_dev, int32_t dtype, void* dy16, float* dgamma, float* dbeta,
                      int32_t accumulate, void* stream);
/* 2x2/2 max-pool (vgg19.features) on NHWC 16-bit; backward routes to the first maximum in scan order (PyTorch's
 * tie rule) and, with relu_mask, multiplies by relu'(x). */
int pesr_maxpool2_fwd(const void* x16, int32_t nb, int32_t h, int32_t w, int32_t c, int32_t dtype, void* y16, void* stream);
int pesr_maxpool2_bwd(const void* x16, const void* dy16, int32_t nb, int32_t h, int32_t w, int32_t c, int32_t relu_mask,
                      int32_t dtype, void* dx16, void* stream);
/* Skinny Linear (nn.Linear, model/pesr.py:71,73) for nb <= 16 rows: weights [o][k] 16-bit streamed once per pass.
 * fwd: out = act(x w^T + bias) (split-K partials in `workspace`, pesr_linear_workspace_floats floats);
 * dgrad: dx32[nb][k] = dy w;  wgrad: dw[o][k] (+)= mul/(*div_dev) * dy^T x. */
int64_t pesr_linear_workspace_floats(int32_t nb, int32_t k, int32_t o);
int pesr_linear_skinny_fwd(const void* x16, const void* w16, const float* bias, int32_t nb, int32_t k, int32_t o,
                           int32_t act, int32_t dtype, float* workspace, float* out32, void* out16, void* stream);
/* out = act(sum over ksplit partials [ksplit][nb][o] + bias): closes a split-K Linear run through pesr_conv_igemm. */
int pesr_linear_finalize(const float* partials, int32_t ksplit, int32_t nb, int32_t o, const float* bias, int32_t act,
                         int32_t dtype, float* out32, void* out16, void* stream);
int pesr_linear_skinny_dgrad(const float* dy, const void* w16, int32_t nb, int32_t k, int32_t o, int32_t dtype,
                             float* dx32, void* stream);
int pesr_linear_skinny_wgrad(const float* dy, const void* x16, int32_t nb, int32_t k, int32_t o, float mul,
                             const float* div_dev, int32_t accumulate, int32_t dtype, float* dw, void* stream);
int pesr_cast16(const float* src, int64_t n, int32_t dtype, void* dst, void* stream);
/* features.view(N, -1) of model/pesr.py:79: NHWC 16-bit -> NCHW-flattened 16-bit, and its backward
 * (fp32 NCHW-flat gradient * mul * (*mul_dev) * lrelu'(mask) -> NHWC 16-bit). */
int pesr_flatten_nchw16(const void* src_nhwc16, int32_t nb, int32_t hw, int32_t c, void* dst, void* stream);
int pesr_unflatten_nchw16(const float* src32_nchw, const void* mask_nhwc16, int32_t nb, int32_t hw, int32_t c, float mul,
                          const float* mul_dev, float slope, int32_t dtype, void* dst_nhwc16, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PESR_B200_H */
