// a7: build_gwc_volume -- group-wise correlation volume (models/coreContinuous_IGEV/submodule.py:253-271).
//
// The reference runs `maxdisp` slice-multiply-mean-assign passes, re-reading both feature maps 48x and
// memset-ing the volume first.  Here one CTA owns one (batch, row, group): the group's C/G channels of the
// left and right rows are staged in shared memory ONCE (right row left-padded with zeros so x-d < 0 reads 0),
// each thread accumulates a 4(x) x 16(d) register tile (6 x 128-bit smem loads per 64 FMAs) and the volume is
// written exactly once with 128-bit stores, coalesced along x.
#include <cstdlib>
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


// Persistent, double-buffered variant of the tiled kernel: a CTA walks over (batch, row, group) units; while it
// accumulates and stores unit k, the 2 x cpg feature rows of unit k+1 are already streaming into the other shared-memory
// stage through cp.async (16-byte LDGSTS, fully coalesced rows), so HBM reads, the shared-memory-bound accumulation
// (96 B of LDS per 64 FMA) and the 128-bit volume stores overlap instead of alternating per CTA.
// r1 kernel: 170 us = 0.49 of the copy peak at config 2 (load -> barrier -> compute -> store, 3 CTAs/SM).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}

// Register tile 8(x) x 16(d) per thread (r1: 4 x 16): per channel 2 + 6 128-bit smem loads feed 128 FMAs (1 B/FMA instead
// of 1.5) -- ncu showed the 4 x 16 version bound by shared-memory wavefronts (L1 71 % busy, DRAM 40 %).
constexpr int kPT = 128;   // threads per CTA of the pipelined kernel
constexpr int kXT = 8;     // x positions per thread tile

template <bool SPLIT>
__global__ void __launch_bounds__(kPT, 3) gwc_fwd_pipe_kernel(const float* __restrict__ left,
                                                           const float* __restrict__ right,
                                                           float* __restrict__ out, int C, int H, int W,
                                                           int maxdisp, int G, int padl, int units) {
  extern __shared__ __align__(16) float s[];
  const int cpg = C / G;
  const int rp = padl + W;                         // right-row pitch (multiple of 4)
  const int stage_floats = cpg * (W + rp);
  const int tid = threadIdx.x;
  const long long HW = (long long)H * W;
  const int w4 = W >> 2, p4 = padl >> 2;
  // Rows are stored with their 16-byte quads split by parity (even quads first, then odd quads): a thread's 8-wide x tile
  // is one even and one odd quad, so the 32 lanes of every 128-bit shared-memory load read CONSECUTIVE quads (the
  // natural order put them 32 bytes apart: 2-way bank conflicts on all 8 loads per channel, 10.9 M of 20 M wavefronts).
  const int halfL = (w4 + 1) >> 1, halfR = ((rp >> 2) + 1) >> 1;
  auto physL = [&](int q) { return SPLIT ? (q & 1) * halfL + (q >> 1) : q; };
  auto physR = [&](int q) { return SPLIT ? (q & 1) * halfR + (q >> 1) : q; };
  // the left zero padding of both stages is written once and never overwritten
  for (int i = tid; i < 2 * cpg * p4; i += kPT) {
    const int st = i / (cpg * p4), r = i - st * (cpg * p4);
    const int c = r / p4, x4 = r - c * p4;
    reinterpret_cast<float4*>(s + st * stage_floats + cpg * W + c * rp)[physR(x4)] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  auto unit_of = [&](int u, int& b, int& y, int& g) {   // g fastest: the G units of one image row run back to back
    g = u % G;
    const int t = u / G;
    y = t % H;
    b = t / H;
  };
  auto issue = [&](int u, int st) {
    if (u < units) {
      int b, y, g;
      unit_of(u, b, y, g);
      const float* lrow = left + ((long long)b * C + (long long)g * cpg) * HW + (long long)y * W;
      const float* rrow = right + ((long long)b * C + (long long)g * cpg) * HW + (long long)y * W;
      float* sL = s + st * stage_floats;
      float* sR = sL + cpg * W;
      for (int i = tid; i < cpg * w4; i += kPT) {
        const int c = i / w4, x4 = i - c * w4;
        const int rq = p4 + x4;
        cp_async16(sL + c * W + physL(x4) * 4, lrow + c * HW + x4 * 4);
        cp_async16(sR + c * rp + physR(rq) * 4, rrow + c * HW + x4 * 4);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const float inv = 1.0f / (float)cpg;
  const int nd = as_ceil_div(maxdisp, kDT);
  const int w8 = W / kXT;                          // W % 8 == 0 on this path
  int k = 0;
  issue(blockIdx.x, 0);
  for (int u = blockIdx.x; u < units; u += gridDim.x, ++k) {
    issue(u + gridDim.x, (k + 1) & 1);
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    int b, y, g;
    unit_of(u, b, y, g);
    const float* sL = s + (k & 1) * stage_floats;
    const float* sR = sL + cpg * W;
    float* obase = out + (((long long)b * G + g) * maxdisp) * HW + (long long)y * W;
    for (int item = tid; item < w8 * nd; item += kPT) {
      const int dt = item / w8, x8 = item - dt * w8;
      const int x = x8 * kXT, d0 = dt * kDT;
      float acc[kDT][kXT];
#pragma unroll
      for (int j = 0; j < kDT; ++j)
#pragma unroll
        for (int i = 0; i < kXT; ++i) acc[j][i] = 0.f;
      // r[i] = R[x - d0 - 16 + i], i in [0,24): element (xi, dj) needs R[x+xi-d0-dj] = r[16 + xi - dj]
      const int rq0 = (padl + x - d0 - kDT) >> 3;   // first quad of the window, halved: the window starts on an EVEN quad
      for (int c = 0; c < cpg; ++c) {              // (padl, x, d0, kDT are multiples of 8) and >= 0 (padl >= roundup(maxdisp,16))
        const float4* lp4 = reinterpret_cast<const float4*>(sL + c * W);
        const float4 la = lp4[physL(2 * x8)];
        const float4 lb = lp4[physL(2 * x8 + 1)];
        const float lv[kXT] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
        const float4* rp4 = reinterpret_cast<const float4*>(sR + c * rp);
        float r[24];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const float4 t = rp4[physR(2 * rq0 + q)];
          r[4 * q] = t.x; r[4 * q + 1] = t.y; r[4 * q + 2] = t.z; r[4 * q + 3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < kDT; ++j)
#pragma unroll
          for (int i = 0; i < kXT; ++i) acc[j][i] = fmaf(lv[i], r[16 + i - j], acc[j][i]);
      }
#pragma unroll
      for (int j = 0; j < kDT; ++j) {
        const int d = d0 + j;
        if (d < maxdisp) {
          float* o = obase + (long long)d * HW + x;
          as_stg_stream4(reinterpret_cast<float4*>(o),
                         make_float4(acc[j][0] * inv, acc[j][1] * inv, acc[j][2] * inv, acc[j][3] * inv));
          as_stg_stream4(reinterpret_cast<float4*>(o + 4),
                         make_float4(acc[j][4] * inv, acc[j][5] * inv, acc[j][6] * inv, acc[j][7] * inv));
        }
      }
    }
    __syncthreads();                               // stage k&1 is refilled by the next iteration's issue
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
  const long long units_ll = (long long)B * H * G;
  static const bool use_pipe = !(getenv("AS_GWC_PIPE") && getenv("AS_GWC_PIPE")[0] == '0');   // A/B knob
  if (fast && use_pipe && (W % kXT == 0) && (padl % 8 == 0) && 2 * smem <= 72 * 1024 && units_ll < (1ll << 30)) {
    // two stages per CTA, 3 CTAs per SM, one persistent CTA per resident slot
    // parity-split rows remove the 2-way bank conflicts of the 8-wide tiles but measured SLOWER (157 vs 139 us at config 2,
    // same run): off unless AS_GWC_SPLIT=1
    static const bool split = getenv("AS_GWC_SPLIT") && getenv("AS_GWC_SPLIT")[0] == '1';
    cudaError_t e = cudaFuncSetAttribute(gwc_fwd_pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * smem));
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(gwc_fwd_pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * smem));
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int units = (int)units_ll;
    const int grid = units < 3 * sms ? units : 3 * sms;
    if (split) gwc_fwd_pipe_kernel<true><<<grid, kPT, 2 * smem, st>>>(left, right, out, C, H, W, maxdisp, G, padl, units);
    else gwc_fwd_pipe_kernel<false><<<grid, kPT, 2 * smem, st>>>(left, right, out, C, H, W, maxdisp, G, padl, units);
  } else if (fast) {
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
