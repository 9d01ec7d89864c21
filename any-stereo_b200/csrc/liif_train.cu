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

}  // namespace

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
