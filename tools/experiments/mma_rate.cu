// Experiment (not product code): how long does ONE SS-mode tcgen05.mma (kind::f16, K = 16, operands in 128B-swizzled shared
// memory) occupy the tensor core as a function of M, N and cta_group?  One CTA (or CTA pair) per SM issues R MMAs back to back
// into one accumulator, commits, waits; cycles / R is printed per shape.  Operand contents are whatever shared memory holds.
// Build (on the GPU box): nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I include -I any-stereo_b200/csrc
//                         -o /tmp/mma_rate tools/experiments/mma_rate.cu -lcuda
#include <cstdio>
#include <cuda.h>
#include "umma.cuh"

int as_operand_f16_internal() { return 0; }
int as_operand_fmt_internal() { return 0; }

// MODE 0: kind::f16 only; 1: kind::f8f6f4 (e5m2, K = 32) only; 2: alternating f16 / e5m2 like the 2-pass engine;
// 3: mode 2 with the bookkeeping of the convolution kernel's issue loop around every 8 MMAs (wait on a barrier that is already
//    complete, tcgen05.fence, tcgen05.commit to a scratch barrier)
template <bool TWO, int MODE>
__global__ void __launch_bounds__(128, 1) mma_rate(int M, int N, int R, int a_rows_distinct, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, ready, scratch;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = TWO ? umma::cluster_ctarank() : 0u;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    umma::mbar_init(&bar, 1); umma::mbar_init(&ready, 1); umma::mbar_init(&scratch, 1);
    umma::fence_barrier_init();
    umma::mbar_arrive(&ready);                 // phase 0 of `ready` is complete: waits on it return at once
  }
  if (warp == 0) {
    if (TWO) { umma::tmem_alloc_2sm(&slot, 512); umma::tmem_relinquish_2sm(); }
    else { umma::tmem_alloc(&slot, 512); umma::tmem_relinquish(); }
  }
  umma::fence_proxy_async();
  umma::tc_fence_before();
  __syncthreads();
  if (TWO) umma::cluster_sync_all();
  umma::tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1 && lane == 0 && crank == 0) {
    const uint32_t idesc = umma::idesc_16_f32(M, N, false), idesc8 = umma::idesc_e5m2_f32(M, N);
    const uint32_t a0 = umma::desc_lo_sw128(umma::smem_u32(smem)), b0 = umma::desc_lo_sw128(umma::smem_u32(smem + 64 * 1024));
    const long long t0 = clock64();
    for (int r = 0; r < R; ++r) {
      // walk over 4 K-steps of 8 different 16-KB operand tiles like a real K loop does
      const uint32_t ko = (uint32_t)(r & 3) * 2u + (uint32_t)((r >> 2) % a_rows_distinct) * (16384u >> 4);
      if ((MODE == 3 || MODE == 4) && (r & 7) == 0) {
        umma::mbar_wait(&ready, 0);
        umma::tc_fence_after();
      }
      if (MODE == 6 && (r & 7) == 0) umma::mbar_wait(&ready, 0);
      if (MODE == 7 && (r & 7) == 0) umma::tc_fence_after();
      if (MODE == 8 && (r & 7) == 0) { while (!umma::mbar_test_wait(&ready, 0)) {} }
      if (MODE == 0 || (MODE >= 2 && !(r & 1)))
        umma::mma_ss_lohi<TWO, false>(tmem, a0 + (ko & 0xfffu), b0 + (uint32_t)(r & 3) * 2u, idesc, r ? 1u : 0u);
      else
        umma::mma_ss_lohi<TWO, true>(tmem, a0 + (ko & 0xfffu), b0 + (uint32_t)(r & 3) * 2u, idesc8, r ? 1u : 0u);
      if ((MODE == 3 || MODE == 5) && (r & 7) == 7) {
        if (TWO) umma::mma_commit_2sm(&scratch, 1); else umma::mma_commit(&scratch);
      }
    }
    if (TWO) umma::mma_commit_2sm(&bar, 1); else umma::mma_commit(&bar);
    umma::mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  umma::tc_fence_before();
  __syncthreads();
  if (TWO) umma::cluster_sync_all();
  if (warp == 0) { if (TWO) umma::tmem_dealloc_2sm(tmem, 512); else umma::tmem_dealloc(tmem, 512); }
}

template <bool TWO, int MODE>
static void run(int M, int N, int R, int grid) {
  long long* d;
  cudaMalloc(&d, 8);
  cudaMemset(d, 0, 8);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(mma_rate<TWO, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = TWO ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, mma_rate<TWO, MODE>, M, N, R, 4, d);
    if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return; }
  }
  long long c = 0;
  cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  const double per = (double)c / R;
  const double macs = (double)M * N * 16 / (TWO ? 2 : 1);      // per SM and instruction
  printf("cta_group::%d  %s M=%3d N=%3d : %7.1f cycles / MMA   %6.0f MAC/clk/SM (K=16 equivalents)\n", TWO ? 2 : 1,
         MODE == 0 ? "f16      " : MODE == 1 ? "e5m2     " : MODE == 2 ? "f16+e5m2 " : MODE == 3 ? "conv-loop" : MODE == 4 ? "wait+fence" : MODE == 5 ? "commit   " : MODE == 6 ? "try_wait " : MODE == 7 ? "fence    " : "test_wait", M, N, per, macs / per);
  cudaFree(d);
}

int main() {
  const int R = 4096;
  for (int n : {32, 64, 128, 192, 256}) run<false, 0>(128, n, R, 148);
  for (int n : {64, 128, 256}) run<false, 0>(64, n, R, 148);
  for (int n : {64, 128, 256}) run<true, 0>(256, n, R, 148);
  for (int n : {64, 128, 256}) run<true, 0>(128, n, R, 148);
  for (int n : {64, 128, 256}) run<true, 1>(256, n, R, 148);
  for (int n : {64, 128, 256}) run<true, 2>(256, n, R, 148);
  for (int n : {64, 128, 256}) run<false, 1>(128, n, R, 148);
  for (int n : {64, 128, 256}) run<true, 3>(256, n, R, 148);
  for (int n : {128}) run<true, 4>(256, n, R, 148);
  for (int n : {128}) run<true, 5>(256, n, R, 148);
  for (int n : {128}) run<true, 6>(256, n, R, 148);
  for (int n : {128}) run<true, 7>(256, n, R, 148);
  for (int n : {128}) run<true, 8>(256, n, R, 148);
  for (int n : {128}) run<false, 3>(128, n, R, 148);
  for (int n : {128}) run<false, 4>(128, n, R, 148);
  for (int n : {128}) run<false, 5>(128, n, R, 148);
  return 0;
}
