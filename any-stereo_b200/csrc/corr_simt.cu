// a1 (exact-fp32 mode) + a13-iii: all-pairs row correlation and its adjoint on CUDA cores.
//
// corr[b,y,x1,x2] = sum_d f1[b,d,y,x1] * f2[b,d,y,x2]  (einsum 'aijk,aijh->ajkh',
// corePrune_RAFT/geometry.py:52, coreContinuous_IGEV/geometry.py:70).  This SIMT kernel is the
// bit-stable fp32 parity baseline; the production path is the tcgen05 kernel in corr_umma.cu.
// Both operands are read straight from NCHW (rows of W contiguous floats): no permute copies.
#include "simt_gemm.cuh"

namespace {

using CT = SimtTile<64, 64, 16, 4, 4>;

__global__ void __launch_bounds__(CT::kThreads) corr_fwd_simt_kernel(const float* __restrict__ f1,
                                                                     const float* __restrict__ f2,
                                                                     float* __restrict__ lvl0, int D, int H,
                                                                     int W1, int W2, int pitch) {
  __shared__ __align__(16) float smem[CT::kSmemFloats];
  float* sA = smem;
  float* sB = smem + 16 * CT::kSA;
  const int tid = threadIdx.x;
  const int ty = tid / 16, tx = tid % 16;
  const int n0 = blockIdx.x * 64, m0 = blockIdx.y * 64;
  const int by = blockIdx.z;               // b*H + y
  const int b = by / H, y = by - b * H;
  const long long HW1 = (long long)H * W1, HW2 = (long long)H * W2;
  const float* a_base = f1 + (long long)b * D * HW1 + (long long)y * W1;
  const float* b_base = f2 + (long long)b * D * HW2 + (long long)y * W2;
  float acc[4][4] = {};
  for (int d0 = 0; d0 < D; d0 += 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + 256 * i;
      const int m = e & 63, k = e >> 6;
      const int d = d0 + k;
      sA[k * CT::kSA + m] = (d < D && m0 + m < W1) ? __ldg(a_base + (long long)d * HW1 + m0 + m) : 0.f;
      sB[k * CT::kSB + m] = (d < D && n0 + m < W2) ? __ldg(b_base + (long long)d * HW2 + n0 + m) : 0.f;
    }
    __syncthreads();
    CT::mac(sA, sB, ty, tx, acc);
    __syncthreads();
  }
  const int n = n0 + tx * 4;
  if (n >= pitch) return;   // pitch % 4 == 0: the float4 is entirely inside or outside the padded row
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m < W1)
      *reinterpret_cast<float4*>(lvl0 + ((long long)by * W1 + m) * pitch + n) =
          make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
}

// SECOND = false: g_f1[b,d,y,x1] = sum_x2 gC[b,y,x1,x2] f2[b,d,y,x2]   (M = x1, K = x2)
// SECOND = true : g_f2[b,d,y,x2] = sum_x1 gC[b,y,x1,x2] f1[b,d,y,x1]   (M = x2, K = x1)
template <bool SECOND>
__global__ void __launch_bounds__(CT::kThreads) corr_bwd_simt_kernel(const float* __restrict__ gc, int pitch,
                                                                     const float* __restrict__ other,
                                                                     float* __restrict__ gout, int D, int H,
                                                                     int W1, int W2) {
  __shared__ __align__(16) float smem[CT::kSmemFloats];
  float* sA = smem;
  float* sB = smem + 16 * CT::kSA;
  const int tid = threadIdx.x;
  const int ty = tid / 16, tx = tid % 16;
  const int n0 = blockIdx.x * 64;          // d tile
  const int m0 = blockIdx.y * 64;          // x1 (or x2) tile
  const int by = blockIdx.z;
  const int b = by / H, y = by - b * H;
  const int Wm = SECOND ? W2 : W1;         // output width
  const int Wk = SECOND ? W1 : W2;         // reduction width
  const long long HWk = (long long)H * Wk, HWm = (long long)H * Wm;
  const float* g_base = gc + (long long)by * W1 * pitch;
  const float* o_base = other + (long long)b * D * HWk + (long long)y * Wk;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < Wk; k0 += 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + 256 * i;
      int m, k;
      if (SECOND) { m = e & 63; k = e >> 6; } else { k = e & 15; m = e >> 4; }
      const int kk = k0 + k, mm = m0 + m;
      float v = 0.f;
      if (kk < Wk && mm < Wm)
        v = SECOND ? __ldg(g_base + (long long)kk * pitch + mm) : __ldg(g_base + (long long)mm * pitch + kk);
      sA[k * CT::kSA + m] = v;
      const int kb = e & 15, nb = e >> 4;
      const int d = n0 + nb;
      sB[kb * CT::kSB + nb] = (k0 + kb < Wk && d < D) ? __ldg(o_base + (long long)d * HWk + k0 + kb) : 0.f;
    }
    __syncthreads();
    CT::mac(sA, sB, ty, tx, acc);
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int d = n0 + tx * 4 + j;
    if (d >= D) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i;
      if (m < Wm) gout[((long long)b * D + d) * HWm + (long long)y * Wm + m] = acc[i][j];
    }
  }
}

}  // namespace

int as_corr_fwd_simt_launch(const float* f1, const float* f2, float* lvl0, int B, int D, int H, int W1, int W2,
                            int pitch, cudaStream_t st) {
  if ((long long)B * H > 65535) return AS_ERR_UNSUPPORTED;
  dim3 grid(as_ceil_div(pitch, 64), as_ceil_div(W1, 64), B * H);
  corr_fwd_simt_kernel<<<grid, CT::kThreads, 0, st>>>(f1, f2, lvl0, D, H, W1, W2, pitch);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_corr1d_bwd(const float* g_corr, int pitch, const float* f1, const float* f2, float* g_f1,
                             float* g_f2, int B, int D, int H, int W1, int W2, as_stream_t stream) {
  if (!g_corr || !f1 || !f2 || !g_f1 || !g_f2) return AS_ERR_BAD_ARG;
  if (B <= 0 || D <= 0 || H <= 0 || W1 <= 0 || W2 <= 0 || pitch < W2) return AS_ERR_BAD_ARG;
  if ((long long)B * H > 65535) return AS_ERR_UNSUPPORTED;
  cudaStream_t st = as_cu(stream);
  {
    dim3 grid(as_ceil_div(D, 64), as_ceil_div(W1, 64), B * H);
    corr_bwd_simt_kernel<false><<<grid, CT::kThreads, 0, st>>>(g_corr, pitch, f2, g_f1, D, H, W1, W2);
    AS_RETURN_IF_LAUNCH_FAILED();
  }
  {
    dim3 grid(as_ceil_div(D, 64), as_ceil_div(W2, 64), B * H);
    corr_bwd_simt_kernel<true><<<grid, CT::kThreads, 0, st>>>(g_corr, pitch, f1, g_f2, D, H, W1, W2);
    AS_RETURN_IF_LAUNCH_FAILED();
  }
  return AS_OK;
}
