// a2: average-pool pyramids.
//  * as_pool1d_halve: one [1,2]/stride-2 average pooling step on a row-pitched volume
//    (F.avg_pool2d at corePrune_RAFT/geometry.py:18, coreContinuous_IGEV/geometry.py:28);
//    (a+b)*0.5 is bit-identical to the reference's result.
//  * as_geo_pyramid_build: the reference's permute(0,3,4,1,2).reshape copy of the geometry volume
//    (coreContinuous_IGEV/geometry.py:18) fused with every pooling level (:24) in ONE pass:
//    [B,G,D,H,W] is read once with 128-byte coalesced rows, transposed through shared memory and
//    written as [pixel][d][g] (+ pooled levels) in fully contiguous runs.
#include "common.cuh"

namespace {

__global__ void pool1d_halve_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows,
                                    int w_out, int pitch_in, int pitch_out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = rows * w_out;
  if (idx >= total) return;
  const long long row = idx / w_out;
  const int j = (int)(idx - row * w_out);
  const float2 v = *reinterpret_cast<const float2*>(in + row * pitch_in + 2 * j);
  out[row * pitch_out + j] = (v.x + v.y) * 0.5f;
}

__global__ void pool1d_halve_bwd_acc_kernel(const float* __restrict__ gc, float* __restrict__ gf, long long rows,
                                            int w_coarse, int pitch_coarse, int pitch_fine) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = rows * w_coarse;
  if (idx >= total) return;
  const long long row = idx / w_coarse;
  const int j = (int)(idx - row * w_coarse);
  const float g = gc[row * pitch_coarse + j] * 0.5f;
  float* d = gf + row * pitch_fine + 2 * j;
  d[0] += g;
  d[1] += g;
}

constexpr int kTX = 32;          // pixels (along x) per CTA
constexpr int kTStride = kTX + 1;

struct GeoOut {
  float* ptr[AS_MAX_LEVELS];
};
struct GeoIn {
  const float* ptr[AS_MAX_LEVELS];
};

// One CTA = 32 pixels of one image row x one chunk of DC disparities x all G groups.  DC is a multiple of
// 2^(L-1), so every pooled level of the chunk is computed from the chunk alone.  The chunk is read as G*DC
// coalesced 128-byte rows (batched 8 deep per warp for memory-level parallelism), transposed through shared
// memory ([d*G+g][33], conflict-free both ways) and written as one contiguous (DC>>l)*G-float run per pixel and
// level.  ~25 KB of smem per CTA -> 8 CTAs/SM.
// GT / DCT > 0: compile-time group count and chunk size (every index division becomes a shift); 0 = runtime values
template <int GT, int DCT>
__global__ void __launch_bounds__(256) geo_pyramid_kernel(const float* __restrict__ geo, GeoOut outs, int G_, int Dg,
                                                          int H, int W, int L, int DC_, int nchunks) {
  const int G = GT > 0 ? GT : G_;
  const int DC = DCT > 0 ? DCT : DC_;
  extern __shared__ float s[];
  const int E0 = DC * G;
  float* cur = s;                                  // [DC*G][33]
  float* nxt = s + (size_t)E0 * kTStride;          // [(DC/2)*G][33]
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int x0 = blockIdx.x * kTX, y = blockIdx.y;
  const int b = blockIdx.z / nchunks, ch = blockIdx.z - b * nchunks;
  const int d0 = ch * DC;
  const int nx = min(kTX, W - x0);
  const long long HW = (long long)H * W;
  const float* src = geo + (long long)b * G * Dg * HW + (long long)y * W + x0 + lane;

  for (int r0 = warp * 8; r0 < E0; r0 += 64) {     // 8 rows per warp per batch
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = r0 + i;
      const int g = row / DC, dd = row - g * DC;
      v[i] = (row < E0 && lane < nx && d0 + dd < Dg) ? __ldg(src + ((long long)g * Dg + d0 + dd) * HW) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = r0 + i;
      if (row < E0) {
        const int g = row / DC, dd = row - g * DC;
        cur[(dd * G + g) * kTStride + lane] = v[i];
      }
    }
  }
  __syncthreads();

  const long long n0 = ((long long)b * H + y) * W + x0;
  int Dl = Dg, dc = DC, dl0 = d0;                  // level width, chunk extent and chunk origin at level l
  for (int l = 0; l < L; ++l) {
    const int valid = min(dc, Dl - dl0);           // disparities of this chunk that exist at level l
    if (valid > 0) {
      const int El = valid * G;
      float* dst = outs.ptr[l] + (n0 * Dl + dl0) * G;
      const long long pstride = (long long)Dl * G;
      for (int px = warp; px < nx; px += 8)          // one warp per pixel record: 128-byte contiguous stores
        for (int e = lane; e < El; e += 32) dst[px * pstride + e] = cur[e * kTStride + px];
    }
    if (l + 1 < L) {
      const int dn = dc >> 1;
      const int En = dn * G;
      for (int i = tid; i < En * kTX; i += 256) {
        const int e = i / kTX, px = i - e * kTX;
        const int d2 = e / G, g = e - d2 * G;
        nxt[e * kTStride + px] = (cur[((2 * d2) * G + g) * kTStride + px] + cur[((2 * d2 + 1) * G + g) * kTStride + px]) * 0.5f;
      }
      __syncthreads();
      float* t = cur; cur = nxt; nxt = t;
      Dl >>= 1; dc = dn; dl0 >>= 1;
    }
  }
}

// adjoint: g_geo[b,g,d,y,x] = sum over levels of the pooled-gradient chain, inverse permute fused
__global__ void __launch_bounds__(256) geo_pyramid_bwd_kernel(GeoIn gl, float* __restrict__ ggeo, int G, int Dg,
                                                              int H, int W, int L) {
  extern __shared__ float s[];   // [Dg*G][33]
  const int E0 = Dg * G;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int x0 = blockIdx.x * kTX, y = blockIdx.y, b = blockIdx.z;
  const int nx = min(kTX, W - x0);
  const long long HW = (long long)H * W;
  const long long n0 = ((long long)b * H + y) * W + x0;
  // accumulate sum_l 2^-l * g_l[d>>l] (dropped odd tails get nothing)
  for (int i = tid; i < nx * E0; i += 256) {
    const int px = i / E0, e = i - px * E0;
    const int d = e / G, g = e - d * G;
    float acc = 0.f, wgt = 1.f;
    int Dl = Dg, dl = d;
    for (int l = 0; l < L; ++l) {
      if (dl < Dl) acc += wgt * gl.ptr[l][((n0 + px) * Dl + dl) * G + g];
      // next level: this d contributes only if it is inside the pooled (even-length) prefix
      const int Dn = Dl >> 1;
      if ((dl >> 1) >= Dn) break;
      dl >>= 1; Dl = Dn; wgt *= 0.5f;
    }
    s[e * kTStride + px] = acc;
  }
  __syncthreads();
  for (int row = warp; row < E0; row += 8) {
    const int g = row / Dg, d = row - g * Dg;
    if (lane < nx)
      ggeo[(((long long)b * G + g) * Dg + d) * HW + (long long)y * W + x0 + lane] = s[(d * G + g) * kTStride + lane];
  }
}

}  // namespace

extern "C" int as_pool1d_halve(const float* in, float* out, long long rows, int w_in, int pitch_in, int pitch_out,
                               as_stream_t stream) {
  if (!in || !out || rows <= 0 || w_in < 0 || pitch_in < w_in || pitch_out < w_in / 2) return AS_ERR_BAD_ARG;
  if ((pitch_in & 1) || (reinterpret_cast<uintptr_t>(in) & 7)) return AS_ERR_ALIGNMENT;
  const int w_out = w_in / 2;
  if (w_out == 0) return AS_OK;
  const long long total = rows * w_out;
  pool1d_halve_kernel<<<(unsigned)as_ceil_div_ll(total, 256), 256, 0, as_cu(stream)>>>(in, out, rows, w_out, pitch_in,
                                                                                   pitch_out);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_pool1d_halve_bwd_acc(const float* g_coarse, float* g_fine, long long rows, int w_fine,
                                       int pitch_coarse, int pitch_fine, as_stream_t stream) {
  if (!g_coarse || !g_fine || rows <= 0 || w_fine < 0 || pitch_fine < w_fine || pitch_coarse < w_fine / 2)
    return AS_ERR_BAD_ARG;
  const int w_coarse = w_fine / 2;
  if (w_coarse == 0) return AS_OK;
  const long long total = rows * w_coarse;
  pool1d_halve_bwd_acc_kernel<<<(unsigned)as_ceil_div_ll(total, 256), 256, 0, as_cu(stream)>>>(
      g_coarse, g_fine, rows, w_coarse, pitch_coarse, pitch_fine);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_geo_pyramid_build(const float* geo, int B, int G, int Dg, int H, int W, int num_levels,
                                    float* const* levels, as_stream_t stream) {
  if (!geo || !levels || B <= 0 || G <= 0 || Dg <= 0 || H <= 0 || W <= 0) return AS_ERR_BAD_ARG;
  if (num_levels < 1 || num_levels > AS_MAX_LEVELS) return AS_ERR_BAD_ARG;
  if (B > 65535 || H > 65535) return AS_ERR_UNSUPPORTED;
  GeoOut o{};
  for (int l = 0; l < num_levels; ++l) {
    if (!levels[l]) return AS_ERR_BAD_ARG;
    o.ptr[l] = levels[l];
  }
  int DC = 16;
  while (DC < (1 << (num_levels - 1))) DC <<= 1;
  const int nchunks = as_ceil_div(Dg, DC);
  if ((long long)B * nchunks > 65535) return AS_ERR_UNSUPPORTED;
  const size_t smem = sizeof(float) * kTStride * ((size_t)DC * G + (size_t)(DC / 2) * G);
  if (smem > 220 * 1024) return AS_ERR_UNSUPPORTED;
  dim3 grid(as_ceil_div(W, kTX), H, B * nchunks);
  cudaError_t e;
  if (G == 8 && DC == 16) {                         // the IGEV shape: constant-folded index arithmetic
    e = cudaFuncSetAttribute(geo_pyramid_kernel<8, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    geo_pyramid_kernel<8, 16><<<grid, 256, smem, as_cu(stream)>>>(geo, o, G, Dg, H, W, num_levels, DC, nchunks);
  } else {
    e = cudaFuncSetAttribute(geo_pyramid_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    geo_pyramid_kernel<0, 0><<<grid, 256, smem, as_cu(stream)>>>(geo, o, G, Dg, H, W, num_levels, DC, nchunks);
  }
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_geo_pyramid_bwd(const float* const* g_levels, int B, int G, int Dg, int H, int W, int num_levels,
                                  float* g_geo, as_stream_t stream) {
  if (!g_geo || !g_levels || B <= 0 || G <= 0 || Dg <= 0 || H <= 0 || W <= 0) return AS_ERR_BAD_ARG;
  if (num_levels < 1 || num_levels > AS_MAX_LEVELS) return AS_ERR_BAD_ARG;
  if (B > 65535 || H > 65535) return AS_ERR_UNSUPPORTED;
  GeoIn in{};
  for (int l = 0; l < num_levels; ++l) {
    if (!g_levels[l]) return AS_ERR_BAD_ARG;
    in.ptr[l] = g_levels[l];
  }
  const size_t smem = sizeof(float) * kTStride * (size_t)Dg * G;
  if (smem > 220 * 1024) return AS_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(geo_pyramid_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid(as_ceil_div(W, kTX), H, B);
  geo_pyramid_bwd_kernel<<<grid, 256, smem, as_cu(stream)>>>(in, g_geo, G, Dg, H, W, num_levels);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}
