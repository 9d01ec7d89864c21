// SURVEY 8(f)-3, second half: build_gwc_volume fused with corr_stem's first 3-D convolution
// (models/coreContinuous_IGEV/continuous_IGEVstereo.py:262-264; submodule.py:253-271 build_gwc_volume, :6-32 BasicConv =
// Conv3d(G, G, 3, 1, 1, bias=False) + BatchNorm3d (eval: a per-channel affine) + LeakyReLU(0.01), :328-341 FeatureAtt).
//
//   out[b,co,d,y,x] = att[b,co,y,x] * lrelu( scale[co] * sum_{ci,dz,dy,dx} w[co,ci,dz,dy,dx] * gwc[b,ci,d+dz-1,y+dy-1,x+dx-1] + shift[co] )
//   gwc[b,g,d,y,x]  = mean_{c in group g} f1[b,c,y,x] * f2[b,c,y,x-d]   (0 for x < d and outside the volume)
//
// The group-wise correlation volume (368 MB at config 2) never reaches HBM: a CTA owns a 32-pixel x 4-row patch (all D
// disparities, all 8 output channels); it walks the 6 source rows, computing each GWC row (with its 1-voxel halo in x and
// d) from the feature rows into a 3-row ring in shared memory, and as soon as a row's three neighbours are in the ring
// runs the 216-tap stencil of the row above out of shared memory on the CUDA cores
// (N = 8 output channels is no tensor-core shape: an SS-mode MMA costs 75 cycles whatever N is).  Thread = one pixel x 3
// disparities x 8 output channels: per (ci, dy, dx) it reads 5 volume values and 24 weights (warp-uniform 128-bit
// broadcasts) for 72 FMAs; 16 warps per SM.  fp32 throughout, same accumulation order class as a direct convolution.
#include "common.cuh"

namespace {

constexpr int kG = 8;                      // groups == conv channels in and out
constexpr int kTX = 32;                    // output pixels per CTA row (one warp lane each)
constexpr int kRY = 4;                     // output rows per CTA: 6 source rows feed 4 output rows (rolling 3-row window)
constexpr int kDT = 3;                     // disparities per thread
constexpr int kWarps = 16;                 // kWarps * kDT = 48 disparities per CTA
constexpr int kThreads = 32 * kWarps;
constexpr int kDMax = kWarps * kDT;
constexpr int kXH = kTX + 2;               // tile + halo in x
constexpr int kDH = kDMax + 2;             // + halo in d
constexpr int kRowFloats = kG * kDH * kXH;                      // one GWC row [g][d][x]
constexpr int kTileFloats = 3 * kRowFloats;                     // ring of three rows
constexpr int kWFloats = kG * 27 * kG;                          // [ci][dy][dx][dz][co]

__host__ __device__ constexpr int stem_smem_bytes(int C, int D) {
  return 4 * (kTileFloats + kWFloats + C * kXH + C * (kXH + D));
}

// CPG = channels per group as a compile-time constant (12 / 6: the left features of a column stay in registers), 0 = run time
template <int CPG>
__global__ void __launch_bounds__(kThreads, 1)
gwc_corr_stem_kernel(const float* __restrict__ f1, const float* __restrict__ f2, const float* __restrict__ w,
                     const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ att,
                     float* __restrict__ out, int C, int D, int H, int W, int tiles_x, float slope) {
  extern __shared__ __align__(16) float sm[];
  float* tile = sm;                          // [3][kG][kDH][kXH]: source row yy lives in slot (yy + 3) % 3
  float* ws = tile + kTileFloats;            // [ci][dy][dx][dz][co]
  float* s1 = ws + kWFloats;                 // [C][kXH]        left features of the current source row
  float* s2 = s1 + C * kXH;                  // [C][kXH + D]    right features, x0-1-D .. x0+32
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = blockIdx.x % tiles_x, yb = blockIdx.x / tiles_x, b = blockIdx.y;
  const int x0 = tx * kTX, y0 = yb * kRY;
  const int cpg = CPG ? CPG : C / kG;
  const float inv = 1.0f / (float)cpg;
  const int W2 = kXH + D;

  // conv weights [co][ci][dz][dy][dx] -> [ci][dy][dx][dz][co]
  for (int i = tid; i < kWFloats; i += kThreads) {
    const int co = i % kG, dz = (i / kG) % 3, dx = (i / (kG * 3)) % 3, dy = (i / (kG * 9)) % 3, ci = i / (kG * 27);
    ws[i] = __ldg(w + (((co * kG + ci) * 3 + dz) * 3 + dy) * 3 + dx);
  }
  // the d = -1 and d >= D planes of the ring are the convolution's zero padding: written once
  for (int i = tid; i < kTileFloats; i += kThreads) tile[i] = 0.f;

  const int d0 = warp * kDT;
  for (int yy = y0 - 1; yy <= y0 + kRY && yy <= H; ++yy) {
    // ---- source row yy -> ring slot -------------------------------------------------------------------------------
    float* trow = tile + ((yy + 3) % 3) * kRowFloats;
    __syncthreads();                         // the stencil of the previous row has read this slot; s1 / s2 are free
    if (yy < 0 || yy >= H) {                 // rows outside the image: zero padding
      for (int i = tid; i < kRowFloats; i += kThreads) trow[i] = 0.f;
    } else {
      const float* p1 = f1 + ((long long)b * C * H + yy) * W;
      const float* p2 = f2 + ((long long)b * C * H + yy) * W;
#pragma unroll 8
      for (int i = tid; i < C * kXH; i += kThreads) {
        const int c = i / kXH, xx = x0 - 1 + (i - c * kXH);
        s1[i] = (xx >= 0 && xx < W) ? __ldg(p1 + (long long)c * H * W + xx) : 0.f;
      }
#pragma unroll 8
      for (int i = tid; i < C * W2; i += kThreads) {
        const int c = i / W2, xx = x0 - 1 - D + (i - c * W2);
        s2[i] = (xx >= 0 && xx < W) ? __ldg(p2 + (long long)c * H * W + xx) : 0.f;
      }
      __syncthreads();
      // gwc[g][d][x] (submodule.py:263-266: columns x < d stay zero).  Pass 1: thread = (group, tile column, half of the
      // disparities) keeps the column's left features in registers and walks d; pass 2: the two halo columns, one
      // (group, column, d) value per step, spread over all threads.
      {
        const int dh = tid >> 8, g = (tid >> 5) & 7, xi = 1 + lane;              // 512 threads = 2 x 8 x 32
        const int xx = x0 + lane;
        const float* a = s1 + g * cpg * kXH + xi;
        const float* c2 = s2 + g * cpg * W2 + xi + D;
        float* o = trow + (g * kDH + 1) * kXH + xi;
        float av[CPG ? CPG : 1];
        if (CPG) {
#pragma unroll
          for (int c = 0; c < CPG; ++c) av[c] = a[c * kXH];
        }
        const int dlo = dh * ((D + 1) >> 1), dhi = dh ? D : ((D + 1) >> 1);
        for (int dd = dlo; dd < dhi; ++dd) {
          float acc = 0.f;
          if (xx < W && xx >= dd) {
            if (CPG) {
#pragma unroll
              for (int c = 0; c < CPG; ++c) acc = fmaf(av[c], c2[c * W2 - dd], acc);
            } else {
              for (int c = 0; c < cpg; ++c) acc = fmaf(a[c * kXH], c2[c * W2 - dd], acc);
            }
            acc *= inv;
          }
          o[dd * kXH] = acc;
        }
      }
      for (int job = tid; job < 2 * kG * D; job += kThreads) {
        const int side = job & 1, g = (job >> 1) & 7, dd = job >> 4;
        const int xi = side ? kXH - 1 : 0;
        const int xx = x0 - 1 + xi;
        float acc = 0.f;
        if (xx >= 0 && xx < W && xx >= dd) {
          const float* a = s1 + g * cpg * kXH + xi;
          const float* c2 = s2 + g * cpg * W2 + xi + D - dd;
          for (int c = 0; c < cpg; ++c) acc = fmaf(a[c * kXH], c2[c * W2], acc);
          acc *= inv;
        }
        trow[(g * kDH + 1 + dd) * kXH + xi] = acc;
      }
    }
    // ---- output row y = yy - 1 once its three source rows are in the ring -------------------------------------------
    const int y = yy - 1;
    if (y < y0 || y >= H) continue;
    __syncthreads();
    float acc[kDT][kG];
#pragma unroll
    for (int i = 0; i < kDT; ++i)
#pragma unroll
      for (int co = 0; co < kG; ++co) acc[i][co] = 0.f;
    for (int ci = 0; ci < kG; ++ci) {
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const float* tr = tile + ((y + dy - 1 + 3) % 3) * kRowFloats + (ci * kDH + d0) * kXH + lane;   // d index d0-1 (+1 halo)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          float v[kDT + 2];
#pragma unroll
          for (int i = 0; i < kDT + 2; ++i) v[i] = tr[i * kXH + dx];
          const float4* wp = reinterpret_cast<const float4*>(ws + ((ci * 3 + dy) * 3 + dx) * 3 * kG);
#pragma unroll
          for (int dz = 0; dz < 3; ++dz) {
            const float4 wa = wp[dz * 2], wb = wp[dz * 2 + 1];
#pragma unroll
            for (int i = 0; i < kDT; ++i) {
              const float t = v[i + dz];
              acc[i][0] = fmaf(wa.x, t, acc[i][0]); acc[i][1] = fmaf(wa.y, t, acc[i][1]);
              acc[i][2] = fmaf(wa.z, t, acc[i][2]); acc[i][3] = fmaf(wa.w, t, acc[i][3]);
              acc[i][4] = fmaf(wb.x, t, acc[i][4]); acc[i][5] = fmaf(wb.y, t, acc[i][5]);
              acc[i][6] = fmaf(wb.z, t, acc[i][6]); acc[i][7] = fmaf(wb.w, t, acc[i][7]);
            }
          }
        }
      }
    }
    const int x = x0 + lane;
    if (x < W) {
#pragma unroll
      for (int co = 0; co < kG; ++co) {
        const float sc = __ldg(scale + co), sh = __ldg(shift + co);
        const float a = att ? __ldg(att + (((long long)b * kG + co) * H + y) * W + x) : 1.0f;
#pragma unroll
        for (int i = 0; i < kDT; ++i) {
          const int d = d0 + i;
          if (d < D) {
            float v = fmaf(acc[i][co], sc, sh);
            v = v > 0.f ? v : v * slope;
            as_stg_stream(out + ((((long long)b * kG + co) * D + d) * H + y) * W + x, v * a);
          }
        }
      }
    }
  }
}

}  // namespace

extern "C" int as_gwc_corr_stem_fwd(const float* left, const float* right, const float* conv_weight, const float* scale,
                                    const float* shift, const float* att, float* out, int B, int C, int H, int W, int maxdisp,
                                    int num_groups, float negative_slope, as_stream_t stream) {
  if (!left || !right || !conv_weight || !scale || !shift || !out) return AS_ERR_BAD_ARG;
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || maxdisp <= 0) return AS_ERR_BAD_ARG;
  if (num_groups != kG || C % kG != 0 || maxdisp > kDMax) return AS_ERR_UNSUPPORTED;
  const int smem = stem_smem_bytes(C, maxdisp);
  if (smem > 227 * 1024) return AS_ERR_UNSUPPORTED;
  if (B > 65535 || (long long)as_ceil_div(H, kRY) * as_ceil_div(W, kTX) >= (1LL << 31)) return AS_ERR_INDEX_RANGE;
  const int tiles_x = as_ceil_div(W, kTX);
  dim3 grid((unsigned)(tiles_x * as_ceil_div(H, kRY)), (unsigned)B);
  cudaError_t e;
#define AS_STEM_LAUNCH(CPG)                                                                                               \
  do {                                                                                                                    \
    e = cudaFuncSetAttribute(gwc_corr_stem_kernel<CPG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);               \
    if (e != cudaSuccess) return (int)e;                                                                                  \
    gwc_corr_stem_kernel<CPG><<<grid, kThreads, smem, as_cu(stream)>>>(left, right, conv_weight, scale, shift, att, out, C, \
                                                                       maxdisp, H, W, tiles_x, negative_slope);           \
  } while (0)
  if (C == 96) AS_STEM_LAUNCH(12);
  else if (C == 48) AS_STEM_LAUNCH(6);
  else AS_STEM_LAUNCH(0);
#undef AS_STEM_LAUNCH
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}
