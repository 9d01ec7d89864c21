// a2: average-pool pyramids.
//  * as_pool1d_halve: one [1,2]/stride-2 average pooling step on a row-pitched volume
//    (F.avg_pool2d at corePrune_RAFT/geometry.py:18, coreContinuous_IGEV/geometry.py:28);
//    (a+b)*0.5 is bit-identical to the reference's result.
//  * as_geo_pyramid_build: the reference's permute(0,3,4,1,2).reshape copy of the geometry volume
//    (coreContinuous_IGEV/geometry.py:18) fused with every pooling level (:24) in ONE pass:
//    [B,G,D,H,W] is read once with 128-byte coalesced rows, transposed through shared memory and
//    written as [pixel][d][g] (+ pooled levels) in fully contiguous runs.
#include <cstdlib>
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


// 128-bit variant for the IGEV shape (G = 8 groups, 16-disparity chunks, W % 4 == 0, L <= 2).
// The r1 kernel above moves 4 bytes per lane, runs four passes over shared memory (tile in, level 0 out, pooling, level 1
// out) and executes 5,100 warp instructions per CTA: ncu showed it bound by shared-memory wavefronts (14.6 M) and issue
// slots, DRAM 49 % busy.  Here the tile crosses shared memory ONCE:
//   load : a warp instruction reads 4 rows ((d, g) planes) x 128 B as float4 and stores them as 128-bit rows
//          [e = d*8+g][32 px] whose 16-byte chunks are XOR-swizzled with (e >> 2) & 7;
//   store: thread = (d, half of the groups) x 4 pixels: four 128-bit smem reads give it a 4(g) x 4(px) block, i.e. for each
//          of its 4 pixels 16 contiguous bytes of the [pixel][d][g] record; lane = (d, half) makes every store
//          instruction one contiguous 512-byte run of one pixel record.  The pooled level comes from registers: the
//          partner disparity d+1 sits two lanes up (warp shuffle), (a+b)*0.5 is bit-identical to F.avg_pool2d, and the
//          even-d lanes write the level-1 record (256 contiguous bytes per instruction).
template <int L>
__global__ void __launch_bounds__(256) geo_pyramid_v4_kernel(const float* __restrict__ geo, float* __restrict__ out0,
                                                             float* __restrict__ out1, int Dg, int H, int W, int nchunks) {
  constexpr int G = 8, DC = 16, E0 = DC * G;
  __shared__ __align__(16) float tile[E0 * kTX];   // 16 KB
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int x0 = blockIdx.x * kTX, y = blockIdx.y;
  const int b = blockIdx.z / nchunks, ch = blockIdx.z - b * nchunks;
  const int d0 = ch * DC;
  const int nx = min(kTX, W - x0);                 // multiple of 4
  const long long HW = (long long)H * W;
  {
    const int lr = lane >> 3, c4 = lane & 7;       // row within the 4-row group, 16-byte chunk (4 pixels) of the row
    const float* src = geo + (long long)b * G * Dg * HW + (long long)y * W + x0 + c4 * 4;
    float4 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = i * 8 + warp;                  // (disparity, half of the groups)
      const int dd = c >> 1, g = (c & 1) * 4 + lr;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c4 * 4 < nx && d0 + dd < Dg) v[i] = as_ldg_stream(reinterpret_cast<const float4*>(src + ((long long)g * Dg + d0 + dd) * HW));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = i * 8 + warp;
      const int e = (c >> 1) * G + (c & 1) * 4 + lr;
      *reinterpret_cast<float4*>(tile + e * kTX + ((c4 ^ ((e >> 2) & 7)) << 2)) = v[i];
    }
  }
  __syncthreads();
  // reader: warp w owns the pixel quad w (pixels 4w..4w+3); lane = (d = lane >> 1, half = lane & 1)
  const int pq = warp;
  if (pq * 4 >= nx) return;
  const int dd = lane >> 1, half = lane & 1;
  const int e0 = dd * G + half * 4;                // (e0 >> 2) & 7 == lane & 7: the 8 lanes of a quarter warp hit 8 distinct chunks
  float4 r[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) r[j] = *reinterpret_cast<const float4*>(tile + (e0 + j) * kTX + ((pq ^ (lane & 7)) << 2));
  // r[j] = group (half*4 + j) at pixels 4pq..4pq+3 -> per pixel i the 4 groups are contiguous in the record
  const long long n0 = ((long long)b * H + y) * W + x0 + pq * 4;
  const bool dvalid = d0 + dd < Dg;
  float4 px[4];
  px[0] = make_float4(r[0].x, r[1].x, r[2].x, r[3].x);
  px[1] = make_float4(r[0].y, r[1].y, r[2].y, r[3].y);
  px[2] = make_float4(r[0].z, r[1].z, r[2].z, r[3].z);
  px[3] = make_float4(r[0].w, r[1].w, r[2].w, r[3].w);
  if (dvalid) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      as_stg_stream4(reinterpret_cast<float4*>(out0 + ((n0 + i) * Dg + d0 + dd) * G + half * 4), px[i]);
  }
  if (L > 1) {
    const int D1 = Dg >> 1;
    const int d1 = (d0 + dd) >> 1;
    const bool writer = ((dd & 1) == 0) && d1 < D1;   // odd tail of Dg is dropped (avg_pool2d floor semantics)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 o;
      o.x = (px[i].x + __shfl_down_sync(0xffffffffu, px[i].x, 2)) * 0.5f;
      o.y = (px[i].y + __shfl_down_sync(0xffffffffu, px[i].y, 2)) * 0.5f;
      o.z = (px[i].z + __shfl_down_sync(0xffffffffu, px[i].z, 2)) * 0.5f;
      o.w = (px[i].w + __shfl_down_sync(0xffffffffu, px[i].w, 2)) * 0.5f;
      if (writer) as_stg_stream4(reinterpret_cast<float4*>(out1 + ((n0 + i) * D1 + d1) * G + half * 4), o);
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
  bool aligned = as_aligned16(geo) && (W % 4 == 0);
  for (int l = 0; l < num_levels; ++l) aligned = aligned && as_aligned16(levels[l]);
  static const bool use_v4 = !(getenv("AS_GEOPYR_V4") && getenv("AS_GEOPYR_V4")[0] == '0');      // A/B knob
  if (G == 8 && DC == 16 && aligned && use_v4 && num_levels <= 2) {    // the IGEV shape, 128-bit accesses on both sides
    if (num_levels == 1) geo_pyramid_v4_kernel<1><<<grid, 256, 0, as_cu(stream)>>>(geo, o.ptr[0], nullptr, Dg, H, W, nchunks);
    else geo_pyramid_v4_kernel<2><<<grid, 256, 0, as_cu(stream)>>>(geo, o.ptr[0], o.ptr[1], Dg, H, W, nchunks);
  } else if (G == 8 && DC == 16) {                  // the IGEV shape: constant-folded index arithmetic
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
