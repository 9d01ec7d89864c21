// SURVEY 8(f)-2 in TRAINING (the upsampler runs every iteration there, continuous_IGEVstereo.py:296-301): the two gather
// steps of the arbitrary-scale upsampler and their adjoints as kernels.  In the reference (and in a plain ATen restatement)
// they are advanced-indexing / grid_sample(nearest) gathers whose backward is a sort-based index_put: measured 101 ms of a
// 161 ms forward and most of a 357 ms backward for 16 iterations of 4 pairs at 320x736.
//
//   nearest gather   out[b,q,:] = src[b, iy[b,q], ix[b,q], :]          (liif.py:108-137 liif_feat: the nearest source pixel)
//       src is the FIRST LINEAR LAYER ALREADY APPLIED at source resolution (P = W1_slice . [feat | affinity], 128 channels,
//       pixel-major), which is exact because the layer is linear and the gather picks single pixels
//   adjoint          gsrc[b, iy, ix, :] += gout[b,q,:]                  vector fp32 reductions (red.global.add.v4.f32)
//   context upsample out[b,q] = sum_k pad(disp)[b, iy+ky, ix+kx] * w[b,k,q]   (submodule.py:357-372), and its adjoint
//       (gw[b,k,q] = gout * neighbour, gdisp via scalar reductions)
// The index convention is the reference's: grid_sample(mode='nearest', align_corners=False) after its clamp (liif.py:118).
#include "common.cuh"

namespace {

// index F.grid_sample(mode='nearest', align_corners=False) picks for a normalised coordinate after the reference's clamp
__device__ __forceinline__ int nearest_index(float c, int n) {
  c = fminf(fmaxf(c, -1.0f + 1e-6f), 1.0f - 1e-6f);
  const int i = (int)rintf(((c + 1.0f) * (float)n - 1.0f) * 0.5f);
  return min(max(i, 0), n - 1);
}

// thread = (query, 4 channels)
__global__ void __launch_bounds__(256) nearest_gather_fwd_kernel(const float4* __restrict__ src, const float* __restrict__ coord,
                                                                 float4* __restrict__ out, int h, int w, int C4, long long Q,
                                                                 long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long bq = i / C4;
  const int c4 = (int)(i - bq * C4);
  const long long b = bq / Q;
  const int iy = nearest_index(__ldg(coord + bq * 2), h), ix = nearest_index(__ldg(coord + bq * 2 + 1), w);
  out[i] = __ldg(src + ((b * h + iy) * w + ix) * C4 + c4);
}

__global__ void __launch_bounds__(256) nearest_gather_bwd_kernel(const float4* __restrict__ gout, const float* __restrict__ coord,
                                                                 float* __restrict__ gsrc, int h, int w, int C4, long long Q,
                                                                 long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long bq = i / C4;
  const int c4 = (int)(i - bq * C4);
  const long long b = bq / Q;
  const int iy = nearest_index(__ldg(coord + bq * 2), h), ix = nearest_index(__ldg(coord + bq * 2 + 1), w);
  const float4 g = __ldg(gout + i);
  float* p = gsrc + (((b * h + iy) * w + ix) * C4 + c4) * 4;
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(g.x), "f"(g.y), "f"(g.z), "f"(g.w) : "memory");
}

// thread = query: gw[b,k,q] = gout[b,q] * neighbour_k, gdisp[neighbour_k] += gout[b,q] * w[b,k,q]
__global__ void __launch_bounds__(256) context_upsample_bwd_kernel(const float* __restrict__ disp, const float* __restrict__ wts,
                                                                   const float* __restrict__ coord, const float* __restrict__ gout,
                                                                   float* __restrict__ gdisp, float* __restrict__ gw, int h, int w,
                                                                   long long Q, long long total) {
  const long long bq = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (bq >= total) return;
  const long long b = bq / Q, q = bq - b * Q;
  const int iy = nearest_index(__ldg(coord + bq * 2), h), ix = nearest_index(__ldg(coord + bq * 2 + 1), w);
  const float g = __ldg(gout + bq);
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int y = iy + k / 3 - 1, x = ix + k % 3 - 1;
    const bool in = y >= 0 && y < h && x >= 0 && x < w;            // F.pad(.., (1,1,1,1)) zeros (submodule.py:359)
    const long long o = (b * h + y) * w + x;
    const float d = in ? __ldg(disp + o) : 0.f;
    const long long wi = (b * 9 + k) * Q + q;
    if (gw) gw[wi] = g * d;
    if (gdisp && in) atomicAdd(gdisp + o, g * __ldg(wts + wi));
  }
}

// ---- first MLP layer of the upsampler in training, everything after the source-resolution product fused:
//   h1[b,q,:] = relu( sum_m ( P_m[b, iy_m(q), ix_m(q), :] + rel_m(q) . Wr_m ) + b1 )       (liif.py:108-137, 652-678)
// rel_m = (coord - centre of the picked source pixel) * (h_m, w_m): the relative-coordinate inputs of the reference.
struct L1Maps {
  const float* P[3];       // [B][h][w][C] source-resolution first-layer products
  const float* Wr[3];      // [2][C]: the two weight columns of the relative coordinates, transposed
  float* gP[3];            // adjoint targets (zeroed by the launcher)
  float* gWr[3];           // [2][C] accumulators (zeroed by the launcher)
  int h[3], w[3];
  int M;
};

__device__ __forceinline__ float axis_centre(int i, int n) {      // make_coord (liif.py:32-45): -1 + (2 i + 1) / n
  const float r = 1.0f / (float)n;
  return -1.0f + r + (2.0f * r) * (float)i;
}

// thread = (query, 4 channels)
__global__ void __launch_bounds__(256) liif_layer1_fwd_kernel(const L1Maps mp, const float* __restrict__ coord,
                                                              const float* __restrict__ b1, float4* __restrict__ out, int C4,
                                                              long long Q, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long bq = i / C4;
  const int c4 = (int)(i - bq * C4);
  const long long b = bq / Q;
  const float cy = __ldg(coord + bq * 2), cx = __ldg(coord + bq * 2 + 1);
  float4 acc = __ldg(reinterpret_cast<const float4*>(b1) + c4);
#pragma unroll
  for (int m = 0; m < 3; ++m) {
    if (m < mp.M) {
      const int h = mp.h[m], w = mp.w[m];
      const int iy = nearest_index(cy, h), ix = nearest_index(cx, w);
      const float ry = (cy - axis_centre(iy, h)) * (float)h, rx = (cx - axis_centre(ix, w)) * (float)w;
      const float4 p = __ldg(reinterpret_cast<const float4*>(mp.P[m]) + ((b * h + iy) * w + ix) * C4 + c4);
      const float4 wy = __ldg(reinterpret_cast<const float4*>(mp.Wr[m]) + c4);
      const float4 wx = __ldg(reinterpret_cast<const float4*>(mp.Wr[m]) + C4 + c4);
      acc.x += p.x + ry * wy.x + rx * wx.x;
      acc.y += p.y + ry * wy.y + rx * wx.y;
      acc.z += p.z + ry * wy.z + rx * wx.z;
      acc.w += p.w + ry * wy.w + rx * wx.w;
    }
  }
  out[i] = make_float4(fmaxf(acc.x, 0.f), fmaxf(acc.y, 0.f), fmaxf(acc.z, 0.f), fmaxf(acc.w, 0.f));
}

// block = 8 query lanes x 32 channel quads (C = 128); every thread walks kQPT queries, scatters g = gout * (h1 > 0) into the
// source-resolution adjoints and keeps the bias / relative-coordinate weight sums in registers; one reduction per block.
constexpr int kQPT = 32;
__global__ void __launch_bounds__(256) liif_layer1_bwd_kernel(const L1Maps mp, const float* __restrict__ coord,
                                                              const float4* __restrict__ h1, const float4* __restrict__ gout,
                                                              float* __restrict__ gb1, int C4, long long Q, long long BQ) {
  __shared__ float4 red[8][32];
  const int c4 = threadIdx.x & 31, ql = threadIdx.x >> 5;
  float4 sb = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 sy[3], sx[3];
#pragma unroll
  for (int m = 0; m < 3; ++m) sy[m] = sx[m] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long q0 = (long long)blockIdx.x * (8 * kQPT);
  if (c4 < C4) {
    for (int t = 0; t < kQPT; ++t) {
      const long long bq = q0 + t * 8 + ql;
      if (bq >= BQ) break;
      const long long b = bq / Q;
      const float4 hv = __ldg(h1 + bq * C4 + c4), gv = __ldg(gout + bq * C4 + c4);
      const float4 g = make_float4(hv.x > 0.f ? gv.x : 0.f, hv.y > 0.f ? gv.y : 0.f, hv.z > 0.f ? gv.z : 0.f, hv.w > 0.f ? gv.w : 0.f);
      const float cy = __ldg(coord + bq * 2), cx = __ldg(coord + bq * 2 + 1);
      sb.x += g.x; sb.y += g.y; sb.z += g.z; sb.w += g.w;
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        if (m < mp.M) {
          const int h = mp.h[m], w = mp.w[m];
          const int iy = nearest_index(cy, h), ix = nearest_index(cx, w);
          const float ry = (cy - axis_centre(iy, h)) * (float)h, rx = (cx - axis_centre(ix, w)) * (float)w;
          sy[m].x += g.x * ry; sy[m].y += g.y * ry; sy[m].z += g.z * ry; sy[m].w += g.w * ry;
          sx[m].x += g.x * rx; sx[m].y += g.y * rx; sx[m].z += g.z * rx; sx[m].w += g.w * rx;
          float* p = mp.gP[m] + (((b * h + iy) * w + ix) * C4 + c4) * 4;
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(g.x), "f"(g.y), "f"(g.z), "f"(g.w) : "memory");
        }
      }
    }
  }
  // block reduction over the 8 query lanes, then one vector reduction per channel quad and quantity
  auto block_sum = [&](float4 v, float* dst) {
    red[ql][c4] = v;
    __syncthreads();
    if (ql == 0 && c4 < C4) {
      float4 s = red[0][c4];
#pragma unroll
      for (int j = 1; j < 8; ++j) { const float4 o = red[j][c4]; s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w; }
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c4 * 4), "f"(s.x), "f"(s.y), "f"(s.z), "f"(s.w) : "memory");
    }
    __syncthreads();
  };
  block_sum(sb, gb1);
  for (int m = 0; m < mp.M; ++m) {
    block_sum(sy[m], mp.gWr[m]);
    block_sum(sx[m], mp.gWr[m] + C4 * 4);
  }
}

}  // namespace

static int l1_fill(L1Maps& mp, int M, const float* const* P, const float* const* Wr, const int* hs, const int* ws) {
  if (M < 1 || M > 3 || !P || !Wr || !hs || !ws) return AS_ERR_BAD_ARG;
  mp.M = M;
  for (int m = 0; m < M; ++m) {
    if (!P[m] || !Wr[m] || hs[m] <= 0 || ws[m] <= 0 || !as_aligned16(P[m]) || !as_aligned16(Wr[m])) return AS_ERR_BAD_ARG;
    mp.P[m] = P[m]; mp.Wr[m] = Wr[m]; mp.h[m] = hs[m]; mp.w[m] = ws[m];
  }
  return AS_OK;
}

extern "C" int as_liif_layer1_fwd(int num_maps, const float* const* P, const float* const* Wr, const int* hs, const int* ws,
                                  const float* hr_coord, const float* b1, float* out, int B, int C, long long Q,
                                  as_stream_t stream) {
  if (!hr_coord || !b1 || !out || B <= 0 || Q <= 0 || C <= 0 || (C & 3) || C > 128) return AS_ERR_BAD_ARG;
  L1Maps mp{};
  int rc = l1_fill(mp, num_maps, P, Wr, hs, ws);
  if (rc != AS_OK) return rc;
  const long long total = (long long)B * Q * (C / 4);
  if (as_ceil_div_ll(total, 256) >= (1LL << 31)) return AS_ERR_INDEX_RANGE;
  liif_layer1_fwd_kernel<<<(unsigned)as_ceil_div_ll(total, 256), 256, 0, as_cu(stream)>>>(
      mp, hr_coord, b1, reinterpret_cast<float4*>(out), C / 4, Q, total);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_liif_layer1_bwd(int num_maps, const float* const* P, const float* const* Wr, const int* hs, const int* ws,
                                  const float* hr_coord, const float* h1, const float* g_out, float* const* g_P,
                                  float* const* g_Wr, float* g_b1, int B, int C, long long Q, as_stream_t stream) {
  if (!hr_coord || !h1 || !g_out || !g_P || !g_Wr || !g_b1 || B <= 0 || Q <= 0 || C <= 0 || (C & 3) || C > 128) return AS_ERR_BAD_ARG;
  L1Maps mp{};
  int rc = l1_fill(mp, num_maps, P, Wr, hs, ws);
  if (rc != AS_OK) return rc;
  cudaStream_t st = as_cu(stream);
  cudaError_t e = cudaMemsetAsync(g_b1, 0, sizeof(float) * C, st);
  for (int m = 0; m < num_maps && e == cudaSuccess; ++m) {
    if (!g_P[m] || !g_Wr[m]) return AS_ERR_BAD_ARG;
    mp.gP[m] = g_P[m]; mp.gWr[m] = g_Wr[m];
    e = cudaMemsetAsync(g_P[m], 0, sizeof(float) * (size_t)B * hs[m] * ws[m] * C, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(g_Wr[m], 0, sizeof(float) * 2 * C, st);
  }
  if (e != cudaSuccess) return (int)e;
  const long long BQ = (long long)B * Q;
  const long long blocks = as_ceil_div_ll(BQ, 8 * kQPT);
  if (blocks >= (1LL << 31)) return AS_ERR_INDEX_RANGE;
  liif_layer1_bwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(mp, hr_coord, reinterpret_cast<const float4*>(h1),
                                                         reinterpret_cast<const float4*>(g_out), g_b1, C / 4, Q, BQ);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_nearest_gather_fwd(const float* src, const float* coord, float* out, int B, int h, int w, int C, long long Q,
                                     as_stream_t stream) {
  if (!src || !coord || !out || B <= 0 || h <= 0 || w <= 0 || C <= 0 || Q <= 0) return AS_ERR_BAD_ARG;
  if ((C & 3) || !as_aligned16(src) || !as_aligned16(out)) return AS_ERR_ALIGNMENT;
  const long long total = (long long)B * Q * (C / 4);
  if (as_ceil_div_ll(total, 256) >= (1LL << 31)) return AS_ERR_INDEX_RANGE;
  nearest_gather_fwd_kernel<<<(unsigned)as_ceil_div_ll(total, 256), 256, 0, as_cu(stream)>>>(
      reinterpret_cast<const float4*>(src), coord, reinterpret_cast<float4*>(out), h, w, C / 4, Q, total);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_nearest_gather_bwd(const float* gout, const float* coord, float* gsrc, int B, int h, int w, int C, long long Q,
                                     as_stream_t stream) {
  if (!gout || !coord || !gsrc || B <= 0 || h <= 0 || w <= 0 || C <= 0 || Q <= 0) return AS_ERR_BAD_ARG;
  if ((C & 3) || !as_aligned16(gout) || !as_aligned16(gsrc)) return AS_ERR_ALIGNMENT;
  const long long total = (long long)B * Q * (C / 4);
  if (as_ceil_div_ll(total, 256) >= (1LL << 31)) return AS_ERR_INDEX_RANGE;
  cudaError_t e = cudaMemsetAsync(gsrc, 0, sizeof(float) * (size_t)B * h * w * C, as_cu(stream));
  if (e != cudaSuccess) return (int)e;
  nearest_gather_bwd_kernel<<<(unsigned)as_ceil_div_ll(total, 256), 256, 0, as_cu(stream)>>>(
      reinterpret_cast<const float4*>(gout), coord, gsrc, h, w, C / 4, Q, total);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_context_upsample_multiscale_bwd(const float* disp_low, const float* up_weights, const float* hr_coord,
                                                  const float* g_out, float* g_disp, float* g_weights, int B, int h, int w,
                                                  long long Q, as_stream_t stream) {
  if (!disp_low || !up_weights || !hr_coord || !g_out || (!g_disp && !g_weights)) return AS_ERR_BAD_ARG;
  if (B <= 0 || h <= 0 || w <= 0 || Q <= 0) return AS_ERR_BAD_ARG;
  const long long total = (long long)B * Q;
  if (as_ceil_div_ll(total, 256) >= (1LL << 31)) return AS_ERR_INDEX_RANGE;
  if (g_disp) {
    cudaError_t e = cudaMemsetAsync(g_disp, 0, sizeof(float) * (size_t)B * h * w, as_cu(stream));
    if (e != cudaSuccess) return (int)e;
  }
  context_upsample_bwd_kernel<<<(unsigned)as_ceil_div_ll(total, 256), 256, 0, as_cu(stream)>>>(
      disp_low, up_weights, hr_coord, g_out, g_disp, g_weights, h, w, Q, total);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}
