// Register-tiled fp32 CUDA-core GEMM micro-kernel shared by the exact-fp32 correlation and
// convolution paths (the parity baselines for the tcgen05 kernels).
#pragma once
#include "common.cuh"

template <int BM, int BN, int BK, int TM, int TN>
struct SimtTile {
  static constexpr int kThreads = (BM / TM) * (BN / TN);
  static constexpr int kSA = BM + 4;  // row strides (floats); multiples of 4 keep 128-bit LDS aligned
  static constexpr int kSB = BN + 4;
  static constexpr int kSmemFloats = BK * (kSA + kSB);
  static_assert(TM % 4 == 0 && TN % 4 == 0, "tiles are read as float4");

  // acc[i][j] += sum_k sA[k][ty*TM+i] * sB[k][tx*TN+j]
  __device__ static __forceinline__ void mac(const float* sA, const float* sB, int ty, int tx,
                                             float (&acc)[TM][TN]) {
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 t = *reinterpret_cast<const float4*>(sA + k * kSA + ty * TM + i);
        a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 t = *reinterpret_cast<const float4*>(sB + k * kSB + tx * TN + j);
        b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
};
