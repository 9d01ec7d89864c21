/*
 * anystereo_b200 -- C ABI of the B200-native (sm_100a) Any-Stereo iterative cost-volume hot path.
 *
 * Plain pointers + sizes + a CUDA stream; no torch types, no exceptions across the boundary.
 * Every entry point
 *   - takes DEVICE pointers (unless the parameter is documented "host array"),
 *   - launches asynchronously on `stream` (pass the caller's current stream; 0 = legacy default),
 *   - never allocates, never synchronises, is CUDA-graph capturable and re-entrant
 *     (no global mutable state; the caller selects the device),
 *   - returns AS_OK (0), a negative AS_ERR_* argument error, or a positive cudaError_t
 *     observed by cudaGetLastError() right after the launch.
 *
 * Each function names the reference interface it replaces (paths relative to the reference
 * repository, Zhaohuai-L/Any-Stereo).  The Python layer in any-stereo_b200/ binds these with ctypes and
 * re-exposes the reference's operator names (CorrBlock1D, Combined_Geo_Encoding_Volume,
 * build_gwc_volume, corr_sampler.forward/backward, BasicMultiUpdateBlock).
 */
#ifndef ANYSTEREO_B200_H
#define ANYSTEREO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* as_stream_t; /* cudaStream_t */

#define AS_OK 0
#define AS_ERR_BAD_ARG (-1)      /* null pointer / non-positive size                       */
#define AS_ERR_UNSUPPORTED (-2)  /* shape or option outside what the kernels implement     */
#define AS_ERR_INDEX_RANGE (-3)  /* tensor exceeds 32-bit element indexing                  */
#define AS_ERR_ALIGNMENT (-4)    /* pointer / pitch alignment requirement violated          */
#define AS_ERR_DRIVER (-5)       /* could not resolve a CUDA driver entry point (TMA maps)  */

#define AS_MAX_LEVELS 8

/* dtype tags for the dtype-dispatched entry points (reference: AT_DISPATCH_FLOATING_TYPES_AND_HALF,
 * sampler/sampler_kernel.cu:126,157) */
#define AS_DTYPE_F32 0
#define AS_DTYPE_F16 1
#define AS_DTYPE_F64 2

int as_abi_version(void);
const char* as_error_string(int code);
/* compiled-in SM target, e.g. 100 for sm_100a */
int as_compiled_sm(void);

/* ------------------------------------------------------------------------------------------
 * a6  corr_sampler.forward / corr_sampler.backward
 *     replaces sampler/sampler.cpp:24-45 (pybind boundary) and sampler/sampler_kernel.cu:19-166.
 *
 * volume [B,H,W1,W2] (dtype), coords [B,coords_ch,H,W1] float32 (channel 0 is used; the reference
 * also reads an unused channel 1), out [B,2r+1,H,W1] (dtype).
 *   out[b,k,y,x] = (1-f)*V[b,y,x,t+k] + f*V[b,y,x,t+k+1],  t = floor(x0)-r, f = x0-floor(x0),
 * taps outside [0,W2) contribute 0.  `out` is fully written (no pre-zeroing needed).
 * ------------------------------------------------------------------------------------------ */
int as_sampler_fwd(const void* volume, const float* coords, int coords_ch, void* out,
                   int B, int H, int W1, int W2, int radius, int dtype, as_stream_t stream);
/* volume_grad [B,H,W1,W2] is fully written (zero-fill fused; the reference memsets then scatters). */
int as_sampler_bwd(const float* coords, int coords_ch, const void* corr_grad, void* volume_grad,
                   int B, int H, int W1, int W2, int radius, int dtype, as_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a1 + a2  all-pairs row correlation and its average-pool pyramid
 *     replaces CorrBlock1D.corr / __init__ (models/corePrune_RAFT/geometry.py:7-19,46-56) and
 *     Combined_Geo_Encoding_Volume.corr / __init__ (models/coreContinuous_IGEV/geometry.py:14-29,63-72).
 *
 * f1 [B,D,H,W1], f2 [B,D,H,W2] float32 NCHW.  levels[l] (host array of device pointers) receives
 * level l as [B*H*W1][pitch[l]] float32 with valid width W2>>l; pitches in floats, >= width.
 * Level 0 = sum_d f1*f2 (no scaling); level l+1[j] = (level l[2j] + level l[2j+1]) * 0.5.
 * mode: AS_CORR_FP32_SIMT  exact fp32 FMA accumulation on CUDA cores,
 *       AS_CORR_BF16X3     tcgen05 tensor cores, operands split hi+lo bf16, 3 MMAs, fp32 accumulate
 *                          (fp32-parity mode, ~1e-5 relative to max|corr|),
 *       AS_CORR_BF16       tcgen05, single bf16 MMA (fast mode; ~2e-3, reported separately).
 * The tensor-core modes need `workspace` of as_corr1d_workspace_bytes() bytes (256-B aligned).
 * ------------------------------------------------------------------------------------------ */
#define AS_CORR_FP32_SIMT 0
#define AS_CORR_BF16X3 1
#define AS_CORR_BF16 2
size_t as_corr1d_workspace_bytes(int B, int D, int H, int W1, int W2, int mode);
int as_corr1d_build(const float* f1, const float* f2, int B, int D, int H, int W1, int W2,
                    int num_levels, float* const* levels, const int* pitches, int mode,
                    void* workspace, size_t workspace_bytes, as_stream_t stream);
/* one pooling step on its own: in [rows][pitch_in] width w_in -> out [rows][pitch_out] width w_in/2 */
int as_pool1d_halve(const float* in, float* out, long long rows, int w_in, int pitch_in, int pitch_out,
                    as_stream_t stream);
/* adjoints (training, config 5): g_fine += halve^T(g_coarse); dF1,dF2 from dCorr (fp32 SIMT) */
int as_pool1d_halve_bwd_acc(const float* g_coarse, float* g_fine, long long rows, int w_fine,
                            int pitch_coarse, int pitch_fine, as_stream_t stream);
int as_corr1d_bwd(const float* g_corr, int pitch, const float* f1, const float* f2,
                  float* g_f1, float* g_f2, int B, int D, int H, int W1, int W2, as_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a2 (IGEV)  geometry-encoding-volume pyramid
 *     replaces the permute+reshape copy and avg_pool2d at coreContinuous_IGEV/geometry.py:17-25.
 * geo [B,G,Dg,H,W] float32 -> levels[l] = [B*H*W][Dg>>l][G] float32 (disparity-major, group-minor:
 * the 2r+2 taps x G groups one lookup needs are one contiguous, 32-byte aligned run).
 * ------------------------------------------------------------------------------------------ */
int as_geo_pyramid_build(const float* geo, int B, int G, int Dg, int H, int W, int num_levels,
                         float* const* levels, as_stream_t stream);
/* adjoint: g_levels (same layout) -> g_geo [B,G,Dg,H,W] (pool-bwd + inverse permute, one pass) */
int as_geo_pyramid_bwd(const float* const* g_levels, int B, int G, int Dg, int H, int W,
                       int num_levels, float* g_geo, as_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a3  CorrBlock1D.__call__   (models/corePrune_RAFT/geometry.py:24-43 + utils/utils.py:59-72)
 * levels/pitches/widths: host arrays (num_levels entries) describing the pyramid from
 * as_corr1d_build.  disp [B,1,H,W]; coords [B,H,W,1] or NULL (NULL => coords[b,y,x] = x, which is
 * what both model forwards pass: prune_raft_stereo.py:272).  out [B, L*(2r+1), H, W] float32,
 * channel = level*(2r+1)+tap, sample position (coords - disp)/2^level + tap - r.
 * ------------------------------------------------------------------------------------------ */
int as_corr_lookup_fwd(const float* const* levels, const int* widths, const int* pitches,
                       int num_levels, const float* disp, const float* coords, float* out,
                       int B, int H, int W, int radius, as_stream_t stream);
/* adjoint w.r.t. the pyramid levels: g_levels[l] must be zero-initialised [N][pitch]; rows are
 * owned by one pixel so no atomics (sampler/sampler_kernel.cu:83-103 semantics per level). */
int as_corr_lookup_bwd(float* const* g_levels, const int* widths, const int* pitches,
                       int num_levels, const float* disp, const float* coords, const float* g_out,
                       int B, int H, int W, int radius, as_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a4  Combined_Geo_Encoding_Volume.__call__  (models/coreContinuous_IGEV/geometry.py:34-60)
 * geo_levels from as_geo_pyramid_build ([N][Dg>>l][G]); corr levels as above.
 * out [B, L*(G+1)*(2r+1), H, W]; per level: G*(2r+1) geo channels (g*(2r+1)+k, sampled at
 * disp/2^l + k - r along Dg) then (2r+1) corr channels (sampled at (coords-disp)/2^l + k - r).
 * ------------------------------------------------------------------------------------------ */
int as_geo_lookup_fwd(const float* const* geo_levels, int G, int Dg,
                      const float* const* corr_levels, const int* corr_widths, const int* corr_pitches,
                      int num_levels, const float* disp, const float* coords, float* out,
                      int B, int H, int W, int radius, as_stream_t stream);
int as_geo_lookup_bwd(float* const* g_geo_levels, int G, int Dg,
                      float* const* g_corr_levels, const int* corr_widths, const int* corr_pitches,
                      int num_levels, const float* disp, const float* coords, const float* g_out,
                      int B, int H, int W, int radius, as_stream_t stream);
/* the lookup above fused with its only consumer, BasicMotionEncoder.convc1 (1x1, L*(G+1)*9 -> 64)
 * + ReLU (models/coreContinuous_IGEV/update.py:78,85).  The lookup tensor never reaches HBM: features are produced
 * into the shared-memory operand tiles of a tcgen05 MMA.  G must be 8, radius 4, num_levels 1 or 2.
 * w_hi/w_lo: bf16 [64][192] with channel (level l, group g, tap k; g == 8: correlation taps) at
 * K = l*96 + g*10 + k and zeros elsewhere (as_pack_conv_weight_bf16 of the permuted matrix); bias fp32 [64];
 * out_hi/out_lo: bf16 planes [B*H*W][64] = relu(convc1(lookup)); nsplit 3 (bf16x3) or 1 (w_lo/out_lo NULL). */
int as_geo_lookup_convc1(const float* const* geo_levels, int G, int Dg,
                         const float* const* corr_levels, const int* corr_widths, const int* corr_pitches,
                         int num_levels, const float* disp, const float* coords,
                         const void* w_hi, const void* w_lo, const float* bias, int nsplit,
                         void* out_hi, void* out_lo, int B, int H, int W, int radius, as_stream_t stream);
/* Same operator, second kernel generation for the IGEV shape (num_levels == 2): 16-byte gathers, shuffle interpolation,
 * M = 128 real pixels per MMA.  Same arguments and outputs; the weights are packed in the TAP-MAJOR K order instead:
 * channel (level l, group g, tap k) at K = l*96 + k*8 + g, correlation tap k of level l at K = l*96 + 72 + k, and the
 * convc1 BIAS in column K = 81 (the kernel multiplies it by a constant 1; `bias` itself is not read).
 * Returns AS_ERR_UNSUPPORTED (callers then use as_geo_lookup_convc1) unless num_levels == 2, every level buffer is
 * 16-byte aligned and every correlation pitch is a multiple of 4 floats. */
int as_geo_lookup_convc1_tap(const float* const* geo_levels, int G, int Dg,
                             const float* const* corr_levels, const int* corr_widths, const int* corr_pitches,
                             int num_levels, const float* disp, const float* coords,
                             const void* w_hi, const void* w_lo, const float* bias, int nsplit,
                             void* out_hi, void* out_lo, int B, int H, int W, int radius, as_stream_t stream);
/* The RAFT-family twin: CorrBlock1D.__call__ (corePrune_RAFT/geometry.py:24-43) fused with convc1 (L*9 -> 64) + ReLU.
 * num_levels 2 or 4, radius 4.  w_hi/w_lo: bf16 [64][64] with channel (level l, tap k) at K = l*10 + k, zeros elsewhere. */
int as_corr_lookup_convc1(const float* const* corr_levels, const int* corr_widths, const int* corr_pitches,
                          int num_levels, const float* disp, const float* coords,
                          const void* w_hi, const void* w_lo, const float* bias, int nsplit,
                          void* out_hi, void* out_lo, int B, int H, int W, int radius, as_stream_t stream);
/* index-parity probe: integer base tap (floor(x)-r) and fractional weight per pixel for one level.
 * kind 0 = geo position disp/2^l, kind 1 = corr position (coords-disp)/2^l. */
int as_lookup_taps(const float* disp, const float* coords, int B, int H, int W, int radius,
                   int level, int kind, int32_t* tap0, float* frac, as_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a7  build_gwc_volume  (models/coreContinuous_IGEV/submodule.py:253-271)
 * left,right [B,C,H,W] float32 -> out [B,G,maxdisp,H,W] float32, fully written:
 *   out[b,g,d,y,x] = (G/C) * sum_{c in group g} left[b,c,y,x]*right[b,c,y,x-d]  (x>=d), else 0.
 * ------------------------------------------------------------------------------------------ */
int as_gwc_build_fwd(const float* left, const float* right, float* out,
                     int B, int C, int H, int W, int maxdisp, int G, as_stream_t stream);
int as_gwc_build_bwd(const float* g_out, const float* left, const float* right,
                     float* g_left, float* g_right,
                     int B, int C, int H, int W, int maxdisp, int G, as_stream_t stream);

/* SURVEY 8(f)-4 (RAFT feature encoder, models/corePrune_RAFT/extractor.py:126-201 with norm_fn='instance'; ResidualBlock
 * :9-58): nn.InstanceNorm2d(affine=False, track_running_stats=False) on a channels-last tensor, fused with what follows it:
 *   y = (x - mean_hw) * rsqrt(var_hw + eps)  per image and channel (biased variance);  relu != 0: y = max(y, 0);
 *   resid != NULL: y = max(resid + y, 0)   (the block's final relu(x' + y), extractor.py:58).
 * x, resid, out: float32 [B][HW][C] (pixel-major), 16-byte aligned, C % 4 == 0, C <= 1024; out may alias x or resid.
 * workspace: as_instnorm_workspace_bytes(B, C) bytes of device memory (fp64 sums + fp32 mean / rstd), 16-byte aligned. */
size_t as_instnorm_workspace_bytes(int B, int C);
int as_instnorm_nhwc(const float* x, const float* resid, float* out, void* workspace, size_t workspace_bytes, int B,
                     long long HW, int C, float eps, int relu, as_stream_t stream);

/* build_gwc_volume fused with corr_stem = Conv3d(G,G,3,1,1,bias=False) + BatchNorm3d (eval: per-channel
 * affine) + LeakyReLU, and optionally FeatureAtt's multiply (continuous_IGEVstereo.py:262-264; submodule.py:6-32,328-341):
 *   out[b,co,d,y,x] = att[b,co,y,x] * lrelu(scale[co] * conv3d(gwc)[b,co,d,y,x] + shift[co]),  gwc as in as_gwc_build_fwd.
 * conv_weight [G][G][3][3][3] (nn.Conv3d layout), scale/shift [G], att [B][G][H][W] or NULL (= 1), out [B][G][maxdisp][H][W],
 * all float32.  The correlation volume itself never reaches HBM.  G must be 8 and maxdisp <= 48 (AS_ERR_UNSUPPORTED). */
int as_gwc_corr_stem_fwd(const float* left, const float* right, const float* conv_weight, const float* scale,
                         const float* shift, const float* att, float* out,
                         int B, int C, int H, int W, int maxdisp, int G, float negative_slope, as_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a8-a11  per-iteration update block  (models/(all families)/update.py:16-41,73-136)
 *
 * Activations are pixel-major ("NHWC": [B*H*W][C], C contiguous).  A convolution reads up to
 * AS_MAX_SRC sources concatenated along C (this is how torch.cat of update.py:35-36,39,90 is never
 * materialised) and fuses its consumer into the epilogue.
 * ------------------------------------------------------------------------------------------ */
#define AS_MAX_SRC 4
#define AS_LAYOUT_NHWC 0
#define AS_LAYOUT_NCHW 1

/* epilogues */
#define AS_EPI_BIAS 0       /* y = acc + bias                                                       */
#define AS_EPI_BIAS_RELU 1  /* y = relu(acc + bias)                                                 */
#define AS_EPI_GRU_ZR 2     /* Cout = 2*Hd: z = sigmoid(acc+bias+ctx) -> aux0 ; r likewise, r*h -> out
                               (update.py:37-38 and the r*h of :39)                                 */
#define AS_EPI_GRU_Q 3      /* q = tanh(acc+bias+ctx); out = (1-z)*h + z*q  (update.py:39-40)        */

typedef struct as_conv_src {
  const float* ptr; /* activation                                            */
  int channels;     /* channels this source contributes                      */
  int pitch;        /* NHWC: floats between consecutive pixels (>= channels) ; NCHW: ignored */
  int layout;       /* AS_LAYOUT_*                                           */
} as_conv_src;

typedef struct as_conv_desc {
  int B, H, W;          /* output (= input) spatial size, stride 1, zero padding KH/2, KW/2          */
  int KH, KW;           /* 1x1, 3x3, 7x7                                                           */
  int Cout;
  int num_src;
  as_conv_src src[AS_MAX_SRC];
  const float* weight;  /* packed [KH*KW][Cin_total][Cout] from as_pack_conv_weight                */
  const float* bias;    /* [Cout] or NULL                                                          */
  int epilogue;         /* AS_EPI_*                                                                */
  float* out;           /* NHWC [N][out_pitch] at channel offset out_coff, or NCHW [B,Cout,H,W]    */
  int out_pitch, out_coff, out_layout;
  /* GRU epilogues (all NHWC, pitch = Hd = Cout/2 for ZR, Cout for Q unless noted) */
  const float* ctx;     /* ZR: [N][2*Hd] = (cz | cr) ; Q: [N][Hd] = cq   (context terms, update.py:37-39) */
  int ctx_pitch;
  const float* h;       /* hidden state [N][Hd]                                                    */
  float* z;             /* ZR: written ; Q: read                                                   */
  float* save;          /* optional (training): ZR stores r [N][Hd], Q stores q [N][Hd]; NULL otherwise */
} as_conv_desc;

/* nn.Conv2d weight [Cout][Cin][KH][KW] (update.py:29-31,78-82) -> GEMM operand [KH*KW][Cin][Cout];
 * run once per weight version, outside the iteration loop */
int as_pack_conv_weight(const float* w_oihw, float* w_packed, int Cout, int Cin, int KH, int KW,
                        as_stream_t stream);
/* exact-fp32 CUDA-core implicit-GEMM convolution (parity baseline for the tensor-core path) */
int as_conv2d_fp32(const as_conv_desc* desc, as_stream_t stream);

/* helpers between the GRU scales (update.py:94-102), NHWC float32, C contiguous */
int as_pool2x_nhwc(const float* in, float* out, int B, int H, int W, int C, as_stream_t stream);
int as_interp_bilinear_nhwc(const float* in, float* out, int B, int Hin, int Win, int Hout, int Wout,
                            int C, as_stream_t stream);
/* layout moves at the torch boundary: [B,C,H,W] <-> [B*H*W][pitch] (+ channel offset) */
int as_nchw_to_nhwc(const float* in, float* out, int B, int C, int H, int W, int out_pitch,
                    int out_coff, as_stream_t stream);
int as_nhwc_to_nchw(const float* in, float* out, int B, int C, int H, int W, int in_pitch,
                    int in_coff, as_stream_t stream);
/* as as_nchw_to_nhwc with a per-channel bias added: out[n][coff+c] = in[b,c,y,x] + bias[c]
 * (folds a convolution's bias into the loop-invariant GRU context term) */
int as_nchw_to_nhwc_bias(const float* in, const float* bias, float* out, int B, int C, int H, int W,
                         int out_pitch, int out_coff, as_stream_t stream);
/* y[i] = a[i] + b[i] (disp = disp + delta, continuous_IGEVstereo.py:295) */
int as_add_f32(const float* a, const float* b, float* y, long long n, as_stream_t stream);


/* ------------------------------------------------------------------------------------------
 * a8-a11 on tensor cores (tcgen05 + TMA).  Activations are pixel-major bf16 planes [B*H*W][C]:
 * `hi` = round-to-nearest bf16 of the value, `lo` = bf16 of the remainder (fp32-parity mode, nsplit = 3:
 * hi*hi + hi*lo + lo*hi with fp32 accumulation in TMEM); nsplit = 1 is the single-pass fast mode (hi planes only).
 * Under AS_FMT_F16F8 (IEEE-half hi planes + e5m2 pair planes): nsplit = 2 = hi*hi + one e5m2 pass for both cross terms;
 * nsplit = 4 = hi*hi + the weight-residual cross term hi_a*lo_w only (half of the e5m2 pass).
 * Channels per source must be multiples of 64; Cout a multiple of 32, <= 256.
 * ------------------------------------------------------------------------------------------ */
#define AS_UEPI_RELU_SPLIT 0 /* relu(acc+bias) -> hi/lo planes at channel offset out_coff               */
#define AS_UEPI_MOTION 1     /* as above, last channel <- disp  (cat(out, disp), update.py:92)           */
#define AS_UEPI_GRU_ZR 2     /* z=sigmoid(acc+ctx) -> z fp32; r likewise, r*h -> hi/lo   (update.py:37-39) */
#define AS_UEPI_GRU_Q 3      /* h'=(1-z)h + z tanh(acc+ctx) -> out_f32 and hi/lo        (update.py:39-40) */
#define AS_UEPI_DISPHEAD 4   /* u[n][t] = sum_c w2[t][c]*relu(acc+bias)[c]: DispHead.conv2 folded into
                                conv1's epilogue (update.py:23-24); finish with as_disp_delta           */
#define AS_UEPI_LINEAR_F32 5 /* out_f32[n][c] = acc (+ bias if given): plain linear map, fp32 NHWC out;
                                out_pitch/out_coff != 0 place the Cout columns inside a wider fp32 row      */

typedef struct as_umma_src {
  const void* hi; /* bf16 [B*H*W][channels] */
  const void* lo; /* same, or NULL when nsplit == 1 */
  int channels;
} as_umma_src;

typedef struct as_conv_umma_desc {
  int B, H, W, KH, KW; /* 1x1 or 3x3, stride 1, zero padding KH/2 (done by TMA out-of-bounds fill) */
  int Cout;            /* padded N of the GEMM */
  int num_src;
  as_umma_src src[3];
  const void* w_hi; /* bf16 [Cout][KH*KW*Cin_total], K index = tap*Cin_total + cin (as_pack_conv_weight_bf16) */
  const void* w_lo;
  int nsplit, epilogue;
  const float* bias;   /* [Cout] (RELU_SPLIT, MOTION, DISPHEAD)                                  */
  const float* ctx;    /* GRU: context + bias, fp32 [N][ctx_pitch]                               */
  int ctx_pitch;
  const float* h;      /* GRU: hidden state fp32 [N][Hd]                                         */
  float* z;            /* GRU_ZR: written, GRU_Q: read; fp32 [N][Hd]                             */
  float* out_f32;      /* GRU_Q: new hidden state fp32 [N][Hd]                                   */
  void* out_hi;        /* bf16 planes [N][out_pitch] written at channel offset out_coff          */
  void* out_lo;
  int out_pitch, out_coff, cout_valid;
  const float* disp;   /* MOTION: [N]                                                            */
  const float* w2;     /* DISPHEAD: conv2 weight as [9][256] fp32                                */
  float* u;            /* DISPHEAD: [N][9] fp32                                                  */
} as_conv_umma_desc;

int as_conv2d_umma(const as_conv_umma_desc* desc, as_stream_t stream);
/* 16-bit operand format of the tensor-core kernels and of every "hi/lo plane" / packed-weight producer below.
 * AS_FMT_BF16 (default): bf16; with nsplit 3 (hi+lo) this is the fp32-parity mode.  AS_FMT_F16: IEEE half, meant for
 * nsplit 1 -- the single-MMA fast mode with 11-bit mantissas (the analogue of the reference's autocast mixed
 * precision, continuous_IGEVstereo.py:287).  Process-wide; planes and weights must be (re)produced after a switch. */
#define AS_FMT_BF16 0
#define AS_FMT_F16 1
#define AS_FMT_F16F8 2   /* hi = IEEE half, "lo" plane = e5m2 pairs: 2-pass fp32-parity mode (nsplit 2), see csrc/common.cuh */
int as_set_operand_format(int fmt);
int as_get_operand_format(void);
/* nn.Conv2d weight [Cout][Cin][KH][KW] fp32 -> bf16 hi/lo [n_pad][KH*KW*cin_pad] (zero padded rows/channels) */
int as_pack_conv_weight_bf16(const float* w_oihw, void* w_hi, void* w_lo, int Cout, int Cin, int KH, int KW,
                             int n_pad, int cin_pad, as_stream_t stream);
/* fp32 -> bf16 hi/lo planes (lo may be NULL) */
int as_split_f32(const float* in, void* hi, void* lo, long long n, as_stream_t stream);
int as_nchw_to_nhwc_split(const float* in, void* hi, void* lo, int B, int C, int H, int W, int c_pad,
                          as_stream_t stream);
int as_pool2x_nhwc_split(const float* in, void* hi, void* lo, int B, int H, int W, int C, as_stream_t stream);
int as_interp_bilinear_nhwc_split(const float* in, void* hi, void* lo, int B, int Hin, int Win, int Hout,
                                  int Wout, int C, as_stream_t stream);
/* BasicMotionEncoder.convd1 (7x7, 1 -> 64, update.py:80,87) + relu -> bf16 planes [N][out_pitch] at out_coff */
int as_convd1_split(const float* disp, const float* w /*[64][49]*/, const float* bias, void* hi, void* lo,
                    int B, int H, int W, int out_pitch, int out_coff, as_stream_t stream);
/* the same convolution on the tensor cores: im2col rows (49 taps padded to K = 64) built in shared memory, one
 * tcgen05 MMA group per 128-pixel tile.  w_hi/w_lo = as_pack_conv_weight_bf16(convd1.weight viewed as [64][49][1][1],
 * n_pad 64, cin_pad 64).  nsplit 3 (fp32 parity) or 1. */
int as_convd1_umma(const float* disp, const void* w_hi, const void* w_lo, const float* bias, void* out_hi, void* out_lo,
                   int B, int H, int W, int out_pitch, int out_coff, int nsplit, as_stream_t stream);
/* delta[n] = bias2 + sum_t u[n + shift(t)][t]  (zero outside the image): finishes DispHead.conv2 */
int as_disp_delta(const float* u, const float* bias2, float* delta, int B, int H, int W, as_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a13-vi  adjoints of the update block (training, config 5), exact fp32 on CUDA cores.
 * All tensors pixel-major fp32.  "acc" outputs are accumulated into (+=): zero them first.
 * ------------------------------------------------------------------------------------------ */
/* weights for the data gradient: [Cout][Cin][KH][KW] -> [KH*KW (flipped)][Cout][Cin]; then
 * dX = as_conv2d_fp32(src = dY with Cout channels, weight = this, Cout := Cin, no bias)        */
int as_pack_conv_weight_dgrad(const float* w_oihw, float* w_packed, int Cout, int Cin, int KH, int KW,
                              as_stream_t stream);
/* dW[co][ci][ky][kx] += sum_n dY[n][co] * X[n + shift(ky,kx)][ci] ;  db[co] += sum_n dY[n][co] (db may be NULL).
 * X is the concatenation of the descriptor's sources (desc->src, num_src, B/H/W/KH/KW); other fields ignored. */
int as_conv2d_wgrad_fp32(const as_conv_desc* desc, const float* dy, int dy_pitch, int Cout, float* dw_acc,
                         float* db_acc, as_stream_t stream);
/* dx = dy * (y > 0) over n elements with channel pitches (relu of update.py:85-91,23) */
int as_relu_bwd(const float* dy, int dy_pitch, int dy_coff, const float* y, int y_pitch, int y_coff, float* dx,
                int dx_pitch, int dx_coff, long long N, int C, as_stream_t stream);
/* GRU gate adjoints (update.py:37-40), Hd channels per pixel:
 *  step 1: dq_pre = dh'*z*(1-q^2) ; dzr_pre[:, :Hd] = dh'*(q-h)*z*(1-z) ; dh_acc += dh'*(1-z)
 *  step 2: dzr_pre[:, Hd:] = d(rh)*h*r*(1-r) ; dh_acc += d(rh)*r                                */
int as_gru_bwd_gates1(const float* dhn, const float* z, const float* q, const float* h, float* dq_pre,
                      float* dzr_pre, float* dh_acc, long long N, int Hd, as_stream_t stream);
int as_gru_bwd_gates2(const float* drh, int drh_pitch, const float* h, const float* r, float* dzr_pre,
                      float* dh_acc, long long N, int Hd, as_stream_t stream);
/* Tensor-core training path: the fused epilogues of as_conv2d_fp32 (AS_EPI_*) applied to a raw convolution output
 * raw[n][raw_pitch] = acc + bias, as written by as_conv2d_umma with AS_UEPI_LINEAR_F32 (update.py:33-41,85-91,23).
 * Same argument meaning as as_conv_desc: GRU_ZR (Cout = 2*Hd): z <- sigmoid(raw+ctx)[:Hd], save <- r, out <- r*h;
 * GRU_Q: save <- q = tanh(raw+ctx), out <- (1-z)h + zq; BIAS_RELU: out <- max(raw,0); BIAS: out <- raw.
 * out is pixel-major [n][out_pitch] written at channel offset out_coff; save may be NULL. */
int as_conv_epilogue_fp32(const float* raw, int raw_pitch, long long N, int Cout, int epilogue, const float* ctx,
                          int ctx_pitch, const float* h, float* z, float* save, float* out, int out_pitch,
                          int out_coff, as_stream_t stream);
/* BasicMotionEncoder.convd1 (7x7, 1 -> 64, update.py:80,87) for training: relu(conv + bias) as fp32 pixel-major
 * [N][out_pitch] at out_coff (the arithmetic of as_convd1_split), and its weight gradient
 * dw_acc[co][ky*7+kx] += sum_n dY[n][co] * disp[n + shift(ky,kx)] (exact fp32, CUDA cores). */
int as_convd1_fp32(const float* disp, const float* w /*[64][49]*/, const float* bias, float* out, int B, int H, int W,
                   int out_pitch, int out_coff, as_stream_t stream);
int as_convd1_wgrad_fp32(const float* disp, const float* dy, int dy_pitch, int B, int H, int W, float* dw_acc,
                         as_stream_t stream);
/* db[c] += sum_n dy[n][c] (the bias half of as_conv2d_wgrad_fp32) */
int as_bias_grad_fp32(const float* dy, int dy_pitch, int Cout, long long N, float* db_acc, as_stream_t stream);
/* Weight gradient on the tensor cores (tcgen05 + TMA), K = pixels.  Operands are CHANNEL-major 16-bit hi/lo planes
 * [nshift][C][B][H][Wp] (Wp = W rounded up to 8) written by as_transpose_split from pixel-major fp32; copy j is the
 * image shifted horizontally by j - nshift/2 pixels with zeros shifted in (nshift = KW for the inputs, 1 for dY: a TMA
 * box cannot start at a 2-byte offset of the innermost dimension).  A K-block is 64 consecutive pixels of one image
 * row; the vertical tap shift is a TMA coordinate offset with out-of-bounds zero fill.  1x1 / 3x3; every source but
 * the last must have a multiple of 128 channels.
 * dw_acc[co][ci][ky][kx] += sum_n dY[n][co] * X[n + shift(ky,kx)][ci];  ws: [KH*KW][Cout][Cin] fp32 scratch. */
int as_transpose_split(const float* in, int pitch, int coff, int C, int B, int H, int W, void* hi, void* lo, int Wp,
                       int nshift, as_stream_t stream);
typedef struct as_wgrad_umma_desc {
  int B, H, W, KH, KW, Cout;
  int num_src;
  as_umma_src src[3];  /* transposed planes of the concatenated inputs, KW shifted copies each */
  const void* dy_hi;   /* transposed planes of dY [Cout][B][H][Wp] */
  const void* dy_lo;   /* NULL when nsplit == 1 */
  int Wp, nsplit;
  float* ws;
  float* dw_acc;
} as_wgrad_umma_desc;
int as_conv2d_wgrad_umma(const as_wgrad_umma_desc* desc, as_stream_t stream);
/* dst[n][dcoff + c] += src[n][scoff + c] */
int as_add_slice(const float* src, int spitch, int scoff, float* dst, int dpitch, int dcoff, long long N, int C,
                 as_stream_t stream);
int as_pool2x_nhwc_bwd(const float* dy, float* dx_acc, int B, int H, int W, int C, as_stream_t stream);
int as_interp_bilinear_nhwc_bwd(const float* dy, float* dx_acc, int B, int Hin, int Win, int Hout, int Wout,
                                int C, as_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * SURVEY 8(f)-2  arbitrary-scale (LIIF) disparity upsampler after the loop
 * (models/coreContinuous_IGEV/liif.py:575-678 liif_out_multi_scale_Training with
 *  unfold "with_ISU"/"with_v2ISU", pos_dim 0, no cell decoding / local ensemble / quarter sampling;
 *  submodule.py:357-372 context_upsample_multiscale_train; continuous_IGEVstereo.py:192-237).
 * ------------------------------------------------------------------------------------------ */
/* AffinityFeature.forward (liif.py:434-449), 3x3 window, dilation 1: cosine similarity with the 8 neighbours,
 * clipped at 0.  feat fp32 NCHW [B,C,H,W].  aff: fp32 NCHW [B,8,H,W] or NULL.  hi/lo: bf16 planes [B*H*W][c_pad];
 * the 8 affinities are written at channel c_off (multiple of 8) -- i.e. StructureFeature's cat([x, affinity])
 * when the planes already hold x in channels [0, C) (as_nchw_to_nhwc_split).  hi NULL = fp32 output only. */
int as_isu_affinity(const float* feat, int B, int C, int H, int W, float* aff, void* hi, void* lo, int c_pad,
                    int c_off, as_stream_t stream);

/* Per-query part of the upsampler.  The first Linear of the MLP is applied at the SOURCE resolution beforehand
 * (it commutes with the nearest-neighbour gather): P[i] = W1[:, cols of input i] . cat([feat_i, affinity_i]),
 * fp32 [B,h_i,w_i,128] (as_conv2d_umma 1x1 with AS_UEPI_LINEAR_F32).  Per query q = coords[b][q] = (y, x) in [-1,1]:
 *   z1 = relu(wc[0] + sum_i P[i][nearest pixel] + sum_i (wc[1+2i]*rel_y_i + wc[2+2i]*rel_x_i))   (liif.py:108-137)
 *   logits = W4 relu(W3 relu(W2 z1 + b2) + b3) + b4                                          (MLP 128-64-64-9)
 *   out = sum_k softmax(logits)_k * (disp*disp_scale[b]) at the 3x3 neighbourhood of the nearest low-res pixel.
 * w2 [64][128], w3 [64][64], w4 [16][64] (rows 9..15 zero): bf16 hi/lo from as_pack_conv_weight_bf16; b4 padded to 16.
 * logits [B,9,Q] and/or out [B,Q] (either may be NULL).  nsplit 3 = fp32 parity (split bf16), 1 = bf16. */
typedef struct as_liif_query_desc {
  int n_in;            /* 1..3 feature maps                                          */
  const float* P[3];
  int h[3], w[3];
  const float* coords; /* [B][Q][2]                                                  */
  int B, Q;
  const float* wc;     /* [1 + 2*n_in][128]: b1, then W1's relative-coordinate columns */
  const void *w2_hi, *w2_lo, *w3_hi, *w3_lo, *w4_hi, *w4_lo;
  const float *b2, *b3, *b4;
  int nsplit;
  const float* disp;       /* [B][hd][wd] low-resolution disparity (needed for `out`) */
  const float* disp_scale; /* [B] multiplier (4*scale) or NULL                        */
  int hd, wd;
  float* logits;
  float* out;
} as_liif_query_desc;
int as_liif_query(const as_liif_query_desc* desc, as_stream_t stream);

/* context_upsample_multiscale_train (submodule.py:357-372) with caller-supplied weights up_weights [B,9,Q]. */
int as_context_upsample_multiscale(const float* disp_low, const float* up_weights, const float* hr_coord, float* out,
                                   int B, int h, int w, int Q, as_stream_t stream);

/* Adjoint of the above: g_weights [B,9,Q] = g_out * neighbour, g_disp [B,1,h,w] += g_out * weight (either may be NULL).
 * g_disp is zeroed by the call. */
int as_context_upsample_multiscale_bwd(const float* disp_low, const float* up_weights, const float* hr_coord,
                                       const float* g_out, float* g_disp, float* g_weights,
                                       int B, int h, int w, long long Q, as_stream_t stream);
/* The upsampler's feature query in TRAINING (liif.py:108-137 liif_feat, nearest source pixel of every query) on a
 * pixel-major source [B,h,w,C] (C % 4 == 0): out[b,q,:] = src[b, iy(q), ix(q), :] with the reference's index convention
 * (grid_sample nearest, align_corners=False, after the clamp of liif.py:118); hr_coord [B,Q,2] = (y, x) in [-1,1].
 * The adjoint zeroes g_src and accumulates with vector fp32 reductions (summation order not deterministic, like the
 * reference's own index adjoints). */
int as_nearest_gather_fwd(const float* src, const float* hr_coord, float* out, int B, int h, int w, int C, long long Q,
                          as_stream_t stream);
int as_nearest_gather_bwd(const float* g_out, const float* hr_coord, float* g_src, int B, int h, int w, int C, long long Q,
                          as_stream_t stream);

/* First MLP layer of the upsampler in training with everything after the source-resolution product fused (num_maps = 1..3
 * feature maps, C <= 128 hidden channels, C % 4 == 0):
 *   h1[b,q,:] = relu( sum_m ( P_m[b, iy_m(q), ix_m(q), :] + rel_m(q) . Wr_m ) + b1 )
 * P_m [B,h_m,w_m,C] = first-layer weights applied to [feat | affinity] at source resolution; Wr_m [2][C] = the layer's two
 * relative-coordinate columns (y, x) transposed; rel_m = (coord - centre of the picked pixel) * (h_m, w_m).
 * The adjoint takes h1 (for the ReLU mask) and g_out and fills g_P[m], g_Wr[m] [2][C], g_b1 [C] (all zeroed by the call). */
int as_liif_layer1_fwd(int num_maps, const float* const* P, const float* const* Wr, const int* hs, const int* ws,
                       const float* hr_coord, const float* b1, float* out, int B, int C, long long Q, as_stream_t stream);
int as_liif_layer1_bwd(int num_maps, const float* const* P, const float* const* Wr, const int* hs, const int* ws,
                       const float* hr_coord, const float* h1, const float* g_out, float* const* g_P, float* const* g_Wr,
                       float* g_b1, int B, int C, long long Q, as_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * SURVEY 8(f)-3 (first half)  initial-disparity head before the loop
 * (continuous_IGEVstereo.py:267-268: softmax over D of Conv3d(G->1, 3x3x3, pad 1, no bias)(geo_encoding_volume),
 *  then disparity_regression, submodule.py:321-325), fused: cost and probability volumes stay on chip.
 * geo fp32 [B,G,D,H,W]; weight fp32 [1,G,3,3,3]; disp_out [B,1,H,W]; prob_out [B,D,H,W] or NULL.  D <= 64.
 * ------------------------------------------------------------------------------------------ */
int as_init_disparity(const float* geo, const float* weight, float* disp_out, float* prob_out,
                      int B, int G, int D, int H, int W, as_stream_t stream);
/* disparity_regression (submodule.py:321-325): out[b,0,y,x] = sum_d prob[b,d,y,x] * d */
int as_disparity_regression(const float* prob, float* out, int B, int D, int H, int W, as_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ANYSTEREO_B200_H */
