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
int as_operand_f16_internal();

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
