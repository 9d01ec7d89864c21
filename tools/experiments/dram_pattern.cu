// Experiment (not product code): what does HBM deliver for the lookup's access pattern?
//  R: per pixel, one 320-byte window at a random 32-byte-aligned offset of a 1536-byte row (+768-byte row, + two 64-byte
//     windows of 1248/624-byte rows), lanes 4-per-64B like the real kernel; W: 162 output planes written in P-pixel tiles.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s\n",cudaGetErrorString(e)); exit(1);} }while(0)

__global__ void read_windows(const float4* __restrict__ a, int row_f4, int win_f4, const int* __restrict__ off, long long npix,
                             float* sink) {
  // each group of (win_f4) consecutive threads reads one pixel's window
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long pix = gid / win_f4; int k = gid % win_f4;
  if (pix >= npix) return;
  float4 v = a[pix * row_f4 + off[pix] + k];
  if (v.x == 12345.678f) sink[0] = v.y;
}
template <int P>
__global__ void write_planes(float* __restrict__ out, long long npix, int C) {
  long long p0 = (long long)blockIdx.x * P;
  for (int i = threadIdx.x; i < P * C; i += blockDim.x) {
    int c = i / P, p = i % P;
    if (p0 + p < npix) out[(long long)c * npix + p0 + p] = (float)i;
  }
}
__global__ void write_linear(float4* __restrict__ out, long long n4) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) out[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}
template <class F> float timeit(F f, int reps = 10) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  static char* flush = nullptr; if (!flush) CK(cudaMalloc(&flush, 256 << 20));
  float best = 1e9;
  for (int r = 0; r < reps; ++r) {
    cudaMemsetAsync(flush, r, 256 << 20);
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best * 1e3f;
}
int main() {
  const long long N = 239616;
  float* sink; CK(cudaMalloc(&sink, 4));
  struct Arr { int row_bytes, win_bytes, align; } arrs[] = {{1536, 320, 32}, {768, 320, 32}, {1248, 64, 16}, {624, 64, 16}, {1536, 1536, 1536}, {1536, 384, 64}, {1536, 640, 128}};
  for (auto& A : arrs) {
    float4* a; CK(cudaMalloc(&a, N * A.row_bytes)); CK(cudaMemset(a, 0, N * A.row_bytes));
    int* off; CK(cudaMalloc(&off, N * 4));
    int* h = (int*)malloc(N * 4);
    for (long long i = 0; i < N; ++i) { int slots = (A.row_bytes - A.win_bytes) / A.align + 1; h[i] = (rand() % slots) * A.align / 16; }
    CK(cudaMemcpy(off, h, N * 4, cudaMemcpyHostToDevice));
    int win_f4 = A.win_bytes / 16; long long threads = N * win_f4;
    float us = timeit([&] { read_windows<<<(unsigned)((threads + 255) / 256), 256>>>(a, A.row_bytes / 16, win_f4, off, N, sink); });
    printf("READ  row %4d B  window %4d B (align %4d): %7.1f us  useful %6.1f GB/s  (if 64B-granular fetch: %6.1f GB/s)\n", A.row_bytes,
           A.win_bytes, A.align, us, N * A.win_bytes / us / 1e3, N * (A.win_bytes + (A.align < 64 ? 32 : 0)) / us / 1e3);
    cudaFree(a); cudaFree(off); free(h);
  }
  float* out; CK(cudaMalloc(&out, N * 162 * 4));
  float us;
  us = timeit([&] { write_planes<32><<<(unsigned)((N + 31) / 32), 256>>>(out, N, 162); });  printf("WRITE 162 planes, tile  32 px: %7.1f us %6.1f GB/s\n", us, N * 648 / us / 1e3);
  us = timeit([&] { write_planes<64><<<(unsigned)((N + 63) / 64), 256>>>(out, N, 162); });  printf("WRITE 162 planes, tile  64 px: %7.1f us %6.1f GB/s\n", us, N * 648 / us / 1e3);
  us = timeit([&] { write_planes<128><<<(unsigned)((N + 127) / 128), 256>>>(out, N, 162); }); printf("WRITE 162 planes, tile 128 px: %7.1f us %6.1f GB/s\n", us, N * 648 / us / 1e3);
  us = timeit([&] { write_planes<256><<<(unsigned)((N + 255) / 256), 256>>>(out, N, 162); }); printf("WRITE 162 planes, tile 256 px: %7.1f us %6.1f GB/s\n", us, N * 648 / us / 1e3);
  long long n4 = N * 162 / 4;
  us = timeit([&] { write_linear<<<(unsigned)((n4 + 255) / 256), 256>>>((float4*)out, n4); }); printf("WRITE linear float4          : %7.1f us %6.1f GB/s\n", us, N * 648 / us / 1e3);
  return 0;
}
