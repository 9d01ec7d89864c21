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
