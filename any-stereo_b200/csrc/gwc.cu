// a7: build_gwc_volume -- group-wise correlation volume (models/coreContinuous_IGEV/submodule.py:253-271).
//
// The reference runs `maxdisp` slice-multiply-mean-assign passes, re-reading both feature maps 48x and
// memset-ing the volume first.  Here one CTA owns one (batch, row, group): the group's C/G channels of the
// left and right rows are staged in shared memory ONCE (right row left-padded with zeros so x-d < 0 reads 0),
// each thread accumulates a 4(x) x 16(d) register tile (6 x 128-bit smem loads per 64 FMAs) and the volume is
// written exactly once with 128-bit stores, coalesced along x.
#include "common.cuh"

namespace {

constexpr int kDT = 16;  // disparities per thread tile

__global__ void __launch_bounds__(256, 3) gwc_fwd_tiled_kernel(const float* __restrict__ left,
                                                            const float* __restrict__ right,
                                                            float* __restrict__ out, int C, int H, int W,
                                                            int maxdisp, int G, int padl) {
  extern __shared__ __align__(16) float s[];
  const int cpg = C / G;
  const int rp = padl + W;              // right-row pitch (multiple of 4)
  float* sL = s;                        // [cpg][W]
  float* sR = s + (size_t)cpg * W;      // [cpg][padl + W]
  const int g = blockIdx.x, y = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x;
  const long long HW = (long long)H * W;
  const float* lrow = left + ((long long)b * C + (long long)g * cpg) * HW + (long long)y * W;
  const float* rrow = right + ((long long)b * C + (long long)g * cpg) * HW + (long long)y * W;
  const int w4 = W >> 2;
  for (int i = tid; i < cpg * w4; i += 256) {
    const int c = i / w4, x4 = i - c * w4;
    reinterpret_cast<float4*>(sL + c * W)[x4] = __ldg(reinterpret_cast<const float4*>(lrow + c * HW) + x4);
    reinterpret_cast<float4*>(sR + c * rp + padl)[x4] = __ldg(reinterpret_cast<const float4*>(rrow + c * HW) + x4);
  }
  const int p4 = padl >> 2;
  for (int i = tid; i < cpg * p4; i += 256) {
    const int c = i / p4, x4 = i - c * p4;
    reinterpret_cast<float4*>(sR + c * rp)[x4] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();

  const float inv = 1.0f / (float)cpg;
  const int nd = as_ceil_div(maxdisp, kDT);
  float* obase = out + (((long long)b * G + g) * maxdisp) * HW + (long long)y * W;
  for (int item = tid; item < w4 * nd; item += 256) {
    const int dt = item / w4, x4 = item - dt * w4;
    const int x = x4 * 4, d0 = dt * kDT;
    float acc[kDT][4];
#pragma unroll
    for (int j = 0; j < kDT; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }
    // r[i] = R[x - d0 - 16 + i], i in [0,20): element (xi, dj) needs R[x+xi-d0-dj] = r[16 + xi - dj]
    const int rbase = padl + x - d0 - kDT;   // >= 0 because padl >= roundup(maxdisp,16)
    for (int c = 0; c < cpg; ++c) {
      const float4 lv = reinterpret_cast<const float4*>(sL + c * W)[x4];
      const float4* rp4 = reinterpret_cast<const float4*>(sR + c * rp + rbase);
      float r[20];
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const float4 t = rp4[q];
        r[4 * q] = t.x; r[4 * q + 1] = t.y; r[4 * q + 2] = t.z; r[4 * q + 3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < kDT; ++j) {
        acc[j][0] = fmaf(lv.x, r[16 - j], acc[j][0]);
        acc[j][1] = fmaf(lv.y, r[17 - j], acc[j][1]);
        acc[j][2] = fmaf(lv.z, r[18 - j], acc[j][2]);
        acc[j][3] = fmaf(lv.w, r[19 - j], acc[j][3]);
      }
    }
#pragma unroll
    for (int j = 0; j < kDT; ++j) {
      const int d = d0 + j;
      if (d < maxdisp)
        *reinterpret_cast<float4*>(obase + (long long)d * HW + x) =
            make_float4(acc[j][0] * inv, acc[j][1] * inv, acc[j][2] * inv, acc[j][3] * inv);
    }
  }
}

// any shape: one thread per output element
__global__ void gwc_fwd_generic_kernel(const float* __restrict__ left, const float* __restrict__ right,
                                       float* __restrict__ out, int C, int H, int W, int maxdisp, int G,
                                       long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = (int)(idx % W);
  long long t = idx / W;
  const int y = (int)(t % H); t /= H;
  const int d = (int)(t % maxdisp); t /= maxdisp;
  const int g = (int)(t % G);
  const int b = (int)(t / G);
  const int cpg = C / G;
  float acc = 0.f;
  if (x >= d) {
    const long long HW = (long long)H * W;
    const float* l = left + ((long long)b * C + (long long)g * cpg) * HW + (long long)y * W + x;
    const float* r = right + ((long long)b * C + (long long)g * cpg) * HW + (long long)y * W + x - d;
    for (int c = 0; c < cpg; ++c) acc = fmaf(l[c * HW], r[c * HW], acc);
    acc *= 1.0f / (float)cpg;
  }
  out[idx] = acc;
}

// adjoint: one thread per (b,c,y,x) for dL and dR
__global__ void gwc_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ left,
                               const float* __restrict__ right, float* __restrict__ gl, float* __restrict__ gr,
                               int C, int H, int W, int maxdisp, int G, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = (int)(idx % W);
  long long t = idx / W;
  const int y = (int)(t % H); t /= H;
  const int c = (int)(t % C);
  const int b = (int)(t / C);
  const int cpg = C / G;
  const int g = c / cpg;
  const long long HW = (long long)H * W;
  const float* gv = gout + (((long long)b * G + g) * maxdisp) * HW + (long long)y * W;
  const float* lrow = left + ((long long)b * C + c) * HW + (long long)y * W;
  const float* rrow = right + ((long long)b * C + c) * HW + (long long)y * W;
  float al = 0.f, ar = 0.f;
  for (int d = 0; d < maxdisp; ++d) {
    if (x >= d) al = fmaf(gv[(long long)d * HW + x], rrow[x - d], al);
    if (x + d < W) ar = fmaf(gv[(long long)d * HW + x + d], lrow[x + d], ar);
  }
  const float inv = 1.0f / (float)cpg;
  gl[idx] = al * inv;
  gr[idx] = ar * inv;
}

}  // namespace

extern "C" int as_gwc_build_fwd(const float* left, const float* right, float* out, int B, int C, int H, int W,
                                int maxdisp, int G, as_stream_t stream) {
  if (!left || !right || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0 || maxdisp <= 0 || G <= 0) return AS_ERR_BAD_ARG;
  if (C % G != 0) return AS_ERR_BAD_ARG;   // reference: assert C % num_groups == 0 (submodule.py:255)
  const long long total = (long long)B * G * maxdisp * H * W;
  cudaStream_t st = as_cu(stream);
  const int cpg = C / G;
  const int padl = as_ceil_div(maxdisp, kDT) * kDT;
  const size_t smem = sizeof(float) * (size_t)cpg * (2 * (size_t)W + padl);
  const bool fast = (W % 4 == 0) && as_aligned16(left) && as_aligned16(right) && as_aligned16(out) &&
                    smem <= 200 * 1024 && G <= 65535 && H <= 65535 && B <= 65535;
  if (fast) {
    cudaError_t e = cudaFuncSetAttribute(gwc_fwd_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid(G, H, B);
    gwc_fwd_tiled_kernel<<<grid, 256, smem, st>>>(left, right, out, C, H, W, maxdisp, G, padl);
  } else {
    gwc_fwd_generic_kernel<<<(unsigned)as_ceil_div_ll(total, 256), 256, 0, st>>>(left, right, out, C, H, W, maxdisp, G,
                                                                              total);
  }
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_gwc_build_bwd(const float* g_out, const float* left, const float* right, float* g_left,
                                float* g_right, int B, int C, int H, int W, int maxdisp, int G,
                                as_stream_t stream) {
  if (!g_out || !left || !right || !g_left || !g_right) return AS_ERR_BAD_ARG;
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || maxdisp <= 0 || G <= 0 || C % G != 0) return AS_ERR_BAD_ARG;
  const long long total = (long long)B * C * H * W;
  gwc_bwd_kernel<<<(unsigned)as_ceil_div_ll(total, 256), 256, 0, as_cu(stream)>>>(g_out, left, right, g_left, g_right,
                                                                              C, H, W, maxdisp, G, total);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}
