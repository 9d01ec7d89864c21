// Shared helpers for the anystereo_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "anystereo_b200.h"

#define AS_RETURN_IF_LAUNCH_FAILED()                 \
  do {                                               \
    cudaError_t e__ = cudaGetLastError();            \
    if (e__ != cudaSuccess) return (int)e__;         \
  } while (0)

static inline cudaStream_t as_cu(as_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

__host__ __device__ static inline int as_ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ static inline long long as_ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

static inline bool as_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// floor division by 4 that is correct for negative numerators
__device__ __forceinline__ int as_floor4(int v) { return v >> 2; }

// streaming (evict-first) 128-bit global accesses for data that is touched once per launch
__device__ __forceinline__ float4 as_ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void as_stg_stream(float* p, float v) {
  asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v));
}
__device__ __forceinline__ void as_stg_stream4(float4* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}

// ----------------------------------------------------------------------------------------------
// 16-bit operand format of the tensor-core path: bf16 (default; split hi/lo = fp32 parity) or IEEE half
// (as_set_operand_format(AS_FMT_F16): single-MMA fast mode with 11-bit mantissas, the analogue of the reference's
// autocast mixed precision).  Process-wide, read by every launcher at call time.
// ----------------------------------------------------------------------------------------------
int as_operand_f16_internal();      // hi planes are IEEE half (AS_FMT_F16 or AS_FMT_F16F8)
int as_operand_fmt_internal();      // AS_FMT_BF16 / AS_FMT_F16 / AS_FMT_F16F8

// two floats -> packed 16-bit pair (lo in bits [0,16)), round to nearest even
__device__ __forceinline__ uint32_t as_cvt16x2(float lo, float hi, bool f16) {
  uint32_t r;
  if (f16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float as_widen_lo16(uint32_t pk, bool f16) {
  return f16 ? __half2float(__ushort_as_half((unsigned short)(pk & 0xFFFFu))) : __uint_as_float(pk << 16);
}
__device__ __forceinline__ float as_widen_hi16(uint32_t pk, bool f16) {
  return f16 ? __half2float(__ushort_as_half((unsigned short)(pk >> 16))) : __uint_as_float(pk & 0xFFFF0000u);
}
// hi = round(v), lo = round(v - hi) for a pair
__device__ __forceinline__ void as_split2(float v0, float v1, uint32_t& hi, uint32_t& lo, bool f16) {
  hi = as_cvt16x2(v0, v1, f16);
  lo = as_cvt16x2(v0 - as_widen_lo16(hi, f16), v1 - as_widen_hi16(hi, f16), f16);
}

// ----------------------------------------------------------------------------------------------
// AS_FMT_F16F8 -- the 2-pass fp32-parity format: x = hi + lo with hi in IEEE half; the "lo" plane (same bytes as a 16-bit
// plane) carries, per 64-channel chunk, 128 bytes [ e5m2(lo * 2^6) x 64 | e5m2(hi * 2^-8) x 64 ] for activations and
// [ e5m2(hi * 2^-6) x 64 | e5m2(lo * 2^8) x 64 ] for weights.  One kind::f16 MMA computes a_hi * w_hi and ONE kind::f8f6f4
// MMA over the 128-byte rows computes a_lo * w_hi + a_hi * w_lo (the scales cancel; fp8 runs at twice the f16 rate, so the
// doubled K costs one f16 pass): 2 pass-equivalents instead of the 3 of hi/lo split products.  The cross terms are
// 2^-12 of the result, so the 2-bit mantissas of e5m2 leave a 2^-15 relative error -- measured on the reference models:
// final-disparity drift 1.0e-4 px (IGEV) / 3.4e-4 px (RAFT) against 0.96e-4 / 1.9e-4 for the 3-pass bf16 split
// (tools/experiments/precision_sim.py).
// ----------------------------------------------------------------------------------------------
constexpr float kX8ActLoScale = 64.0f, kX8ActHiScale = 1.0f / 256.0f;     // activations: lo * 2^6, hi * 2^-8
constexpr float kX8WgtHiScale = 1.0f / 64.0f, kX8WgtLoScale = 256.0f;     // weights:     hi * 2^-6, lo * 2^8

// 4 floats -> 4 e5m2 bytes, `a` in the lowest byte
__device__ __forceinline__ uint32_t as_e5m2x4(float a, float b, float c, float d) {
  uint16_t lo, hi;
  asm("cvt.rn.satfinite.e5m2x2.f32 %0, %1, %2;" : "=h"(lo) : "f"(b), "f"(a));   // first source -> upper byte
  asm("cvt.rn.satfinite.e5m2x2.f32 %0, %1, %2;" : "=h"(hi) : "f"(d), "f"(c));
  return (uint32_t)lo | ((uint32_t)hi << 16);
}
// byte offset inside an x8 plane of the element at flat index `off` (channel pitch a multiple of 64)
__device__ __forceinline__ long long as_x8_byte(long long off) { return ((off >> 6) << 7) + (off & 63); }

// "lo" information of 4 consecutive channels (off % 4 == 0): v = exact values, h = their packed 16-bit hi halves
__device__ __forceinline__ void as_store_lo4(void* lo, long long off, float4 v, uint2 h, int fmt) {
  const bool f16 = fmt != 0;
  const float h0 = as_widen_lo16(h.x, f16), h1 = as_widen_hi16(h.x, f16), h2 = as_widen_lo16(h.y, f16), h3 = as_widen_hi16(h.y, f16);
  if (fmt != 2) {
    uint2 l;
    l.x = as_cvt16x2(v.x - h0, v.y - h1, f16);
    l.y = as_cvt16x2(v.z - h2, v.w - h3, f16);
    *reinterpret_cast<uint2*>(reinterpret_cast<unsigned short*>(lo) + off) = l;
  } else {
    uint8_t* b = reinterpret_cast<uint8_t*>(lo) + as_x8_byte(off);
    *reinterpret_cast<uint32_t*>(b) = as_e5m2x4((v.x - h0) * kX8ActLoScale, (v.y - h1) * kX8ActLoScale, (v.z - h2) * kX8ActLoScale,
                                                (v.w - h3) * kX8ActLoScale);
    *reinterpret_cast<uint32_t*>(b + 64) = as_e5m2x4(h0 * kX8ActHiScale, h1 * kX8ActHiScale, h2 * kX8ActHiScale, h3 * kX8ActHiScale);
  }
}
// 8 consecutive channels (off % 8 == 0): v[8] exact values, h[4] packed hi halves
__device__ __forceinline__ void as_store_lo8(void* lo, long long off, const float* v, const uint32_t* h, int fmt) {
  const bool f16 = fmt != 0;
  float hv[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) { hv[2 * i] = as_widen_lo16(h[i], f16); hv[2 * i + 1] = as_widen_hi16(h[i], f16); }
  if (fmt != 2) {
    uint4 l;
    l.x = as_cvt16x2(v[0] - hv[0], v[1] - hv[1], f16); l.y = as_cvt16x2(v[2] - hv[2], v[3] - hv[3], f16);
    l.z = as_cvt16x2(v[4] - hv[4], v[5] - hv[5], f16); l.w = as_cvt16x2(v[6] - hv[6], v[7] - hv[7], f16);
    *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(lo) + off) = l;
  } else {
    uint8_t* b = reinterpret_cast<uint8_t*>(lo) + as_x8_byte(off);
    uint2 a, c;
    a.x = as_e5m2x4((v[0] - hv[0]) * kX8ActLoScale, (v[1] - hv[1]) * kX8ActLoScale, (v[2] - hv[2]) * kX8ActLoScale, (v[3] - hv[3]) * kX8ActLoScale);
    a.y = as_e5m2x4((v[4] - hv[4]) * kX8ActLoScale, (v[5] - hv[5]) * kX8ActLoScale, (v[6] - hv[6]) * kX8ActLoScale, (v[7] - hv[7]) * kX8ActLoScale);
    c.x = as_e5m2x4(hv[0] * kX8ActHiScale, hv[1] * kX8ActHiScale, hv[2] * kX8ActHiScale, hv[3] * kX8ActHiScale);
    c.y = as_e5m2x4(hv[4] * kX8ActHiScale, hv[5] * kX8ActHiScale, hv[6] * kX8ActHiScale, hv[7] * kX8ActHiScale);
    *reinterpret_cast<uint2*>(b) = a;
    *reinterpret_cast<uint2*>(b + 64) = c;
  }
}

// ----------------------------------------------------------------------------------------------
// 256-bit global stores (sm_100: STG.E.ENL2.256).  Epilogues whose threads each own one pixel row write row-strided
// pieces; with 16 bytes per lane every 32-byte sector is written in two halves by two instructions and the L2 handles
// partial sectors (measured on the fused lookup kernel: 61 MB of output cost 16-20 us with 16-byte stores, 10 us with
// 32-byte stores).  The address must be 32-byte aligned.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void as_stg256(void* p, const uint32_t (&w)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
               "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
__device__ __forceinline__ void as_ldg256f(const float* p, float* v) {     // 8 consecutive floats, read-only path
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void as_ld256f(const float* p, float* v) {      // same through the coherent path
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p) : "memory");
}
__device__ __forceinline__ void as_stg256f(float* p, const float* v) {     // 8 consecutive floats
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
// 32 consecutive channels of one pixel (off = element offset of the first, a multiple of 32; 32-byte aligned rows) -> the
// 16-bit hi plane and the second plane of the format (16-bit lo, or the e5m2 pair plane of AS_FMT_F16F8), 32 bytes per store
__device__ __forceinline__ void as_store_split32_v8(const float* y, void* out_hi, void* out_lo, long long off, int fmt) {
  const bool f16 = fmt != 0;
  uint32_t h[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) h[i] = as_cvt16x2(y[2 * i], y[2 * i + 1], f16);
  {
    const uint32_t a[8] = {h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]};
    const uint32_t b[8] = {h[8], h[9], h[10], h[11], h[12], h[13], h[14], h[15]};
    as_stg256(reinterpret_cast<unsigned short*>(out_hi) + off, a);
    as_stg256(reinterpret_cast<unsigned short*>(out_hi) + off + 16, b);
  }
  if (!out_lo) return;
  float r[32], hv[32];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    hv[2 * i] = as_widen_lo16(h[i], f16);
    hv[2 * i + 1] = as_widen_hi16(h[i], f16);
    r[2 * i] = y[2 * i] - hv[2 * i];
    r[2 * i + 1] = y[2 * i + 1] - hv[2 * i + 1];
  }
  if (fmt != 2) {
    uint32_t l[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) l[i] = as_cvt16x2(r[2 * i], r[2 * i + 1], f16);
    const uint32_t a[8] = {l[0], l[1], l[2], l[3], l[4], l[5], l[6], l[7]};
    const uint32_t b[8] = {l[8], l[9], l[10], l[11], l[12], l[13], l[14], l[15]};
    as_stg256(reinterpret_cast<unsigned short*>(out_lo) + off, a);
    as_stg256(reinterpret_cast<unsigned short*>(out_lo) + off + 16, b);
  } else {
    uint8_t* bp = reinterpret_cast<uint8_t*>(out_lo) + as_x8_byte(off);
    uint32_t a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a[i] = as_e5m2x4(r[4 * i] * kX8ActLoScale, r[4 * i + 1] * kX8ActLoScale, r[4 * i + 2] * kX8ActLoScale, r[4 * i + 3] * kX8ActLoScale);
      b[i] = as_e5m2x4(hv[4 * i] * kX8ActHiScale, hv[4 * i + 1] * kX8ActHiScale, hv[4 * i + 2] * kX8ActHiScale, hv[4 * i + 3] * kX8ActHiScale);
    }
    as_stg256(bp, a);
    as_stg256(bp + 64, b);
  }
}
