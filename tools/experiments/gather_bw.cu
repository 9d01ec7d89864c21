// Experiment (not product code): what can a B200 deliver for the IGEV lookup's READ pattern when the loads are deep and
// asynchronous?  Per pixel: one 320-B window at a random 32-B-aligned offset of its 1536-B geometry row (level 0), one of its
// 768-B row (level 1), one 64-B window (16-B aligned) of its 1248-B correlation row and one of its 624-B row.
//   mode 0: cp.async.bulk (TMA 1-D), one thread per window, mbarrier completion, S stages of 64 pixels
//   mode 1: cp.async 16 B (LDGSTS), commit groups, S stages
//   mode 2: ld.global.nc.v4 with 4 lanes per 64 B, results summed (register path, like the round-1 kernel), occupancy-driven
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gather_bw.bin gather_bw.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s line %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}

struct Arrs { const float* geo0; const float* geo1; const float* cor0; const float* cor1; const int4* off; };   // off: element offsets (floats)
constexpr int P = 64;                  // pixels per stage
constexpr int PIXB = 336 * 2 + 80 * 2; // bytes per pixel in smem: 2 x (320+16 pad) + 2 x (64+16 pad)

template <int S>
__global__ void __launch_bounds__(128, 1) gather_bulk(Arrs a, int npix, int ntiles, float* sink) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t full[S], empty[S];
  const int tid = threadIdx.x;
  if (tid == 0) { for (int s = 0; s < S; ++s) { mbar_init(&full[s], P); mbar_init(&empty[s], 64); } asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  const int my = (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1;
  if (tid < P) {                        // producers: one pixel each
    for (int k = 0; k < my; ++k) {
      const int s = k % S; const uint32_t ph = (k / S) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      const long long n = (long long)(blockIdx.x + k * gridDim.x) * P + tid;
      uint8_t* dst = sm + s * (P * PIXB) + tid * PIXB;
      if (n < npix) {
        const int4 o = a.off[n];
        mbar_expect(&full[s], 320 + 320 + 64 + 64);
        bulk_g2s(dst, a.geo0 + n * 384 + o.x, 320, &full[s]);
        bulk_g2s(dst + 336, a.geo1 + n * 192 + o.y, 320, &full[s]);
        bulk_g2s(dst + 672, a.cor0 + n * 312 + o.z, 64, &full[s]);
        bulk_g2s(dst + 752, a.cor1 + n * 156 + o.w, 64, &full[s]);
      } else mbar_arrive(&full[s]);
    }
  } else {                              // consumers: touch the data lightly (one float4 per window), release
    const int c = tid - P;
    float acc = 0.f;
    for (int k = 0; k < my; ++k) {
      const int s = k % S; const uint32_t ph = (k / S) & 1;
      mbar_wait(&full[s], ph);
      const float4 v = *reinterpret_cast<const float4*>(sm + s * (P * PIXB) + c * PIXB);
      acc += v.x + v.w;
      mbar_arrive(&empty[s]);
    }
    if (acc == 12345.f) sink[0] = acc;
  }
}

template <int S>
__global__ void __launch_bounds__(256, 1) gather_ldgsts(Arrs a, int npix, int ntiles, float* sink) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int tid = threadIdx.x;
  const int my = (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1;
  // 48 16-B chunks per pixel (20 + 20 + 4 + 4); 64 px -> 3072 chunks -> 12 per thread
  auto issue = [&](int k) {
    if (k < my) {
      const long long n0 = (long long)(blockIdx.x + k * gridDim.x) * P;
      uint8_t* st = sm + (k % S) * (P * PIXB);
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        const int item = tid + i * 256;
        const int pix = item / 48, ch = item - pix * 48;
        const long long n = n0 + pix;
        if (n < npix) {
          const int4 o = a.off[n];
          const float* src; int doff;
          if (ch < 20) { src = a.geo0 + n * 384 + o.x + ch * 4; doff = ch * 16; }
          else if (ch < 40) { src = a.geo1 + n * 192 + o.y + (ch - 20) * 4; doff = 336 + (ch - 20) * 16; }
          else if (ch < 44) { src = a.cor0 + n * 312 + o.z + (ch - 40) * 4; doff = 672 + (ch - 40) * 16; }
          else { src = a.cor1 + n * 156 + o.w + (ch - 44) * 4; doff = 752 + (ch - 44) * 16; }
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(st + pix * PIXB + doff)), "l"(src) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int k = 0; k < S - 1; ++k) issue(k);
  float acc = 0.f;
  for (int k = 0; k < my; ++k) {
    issue(k + S - 1);
    asm volatile("cp.async.wait_group %0;" ::"n"(S - 1) : "memory");
    __syncthreads();
    if (tid < P) { const float4 v = *reinterpret_cast<const float4*>(sm + (k % S) * (P * PIXB) + tid * PIXB); acc += v.x + v.w; }
    __syncthreads();
  }
  if (acc == 12345.f) sink[0] = acc;
}

__global__ void __launch_bounds__(256) gather_ldg(Arrs a, int npix, float* sink) {
  // 4 lanes per 64 B: thread = (pixel, chunk of 16 B); 48 chunks per pixel
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = gid / 48; const int ch = (int)(gid - n * 48);
  if (n >= npix) return;
  const int4 o = a.off[n];
  const float* src;
  if (ch < 20) src = a.geo0 + n * 384 + o.x + ch * 4;
  else if (ch < 40) src = a.geo1 + n * 192 + o.y + (ch - 20) * 4;
  else if (ch < 44) src = a.cor0 + n * 312 + o.z + (ch - 40) * 4;
  else src = a.cor1 + n * 156 + o.w + (ch - 44) * 4;
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(src));
  if (v.x == 12345.678f) sink[0] = v.y;
}


// all 256 threads issue ONE bulk copy per tile (64 px x 4 windows), then every thread consumes
template <int S>
__global__ void __launch_bounds__(256) gather_bulk_all(Arrs a, int npix, int ntiles, float* sink) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t full[S];
  const int tid = threadIdx.x;
  if (tid == 0) { for (int s = 0; s < S; ++s) mbar_init(&full[s], 256); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  const int my = (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1;
  const int pix = tid >> 2, wnd = tid & 3;
  auto issue = [&](int k) {
    if (k >= my) return;
    const int s = k % S;
    const long long n = (long long)(blockIdx.x + k * gridDim.x) * P + pix;
    uint8_t* dst = sm + s * (P * PIXB) + pix * PIXB;
    if (n < npix) {
      const int4 o = a.off[n];
      const float* src; uint32_t bytes; int doff;
      if (wnd == 0) { src = a.geo0 + n * 384 + o.x; bytes = 320; doff = 0; }
      else if (wnd == 1) { src = a.geo1 + n * 192 + o.y; bytes = 320; doff = 336; }
      else if (wnd == 2) { src = a.cor0 + n * 312 + o.z; bytes = 64; doff = 672; }
      else { src = a.cor1 + n * 156 + o.w; bytes = 64; doff = 752; }
      mbar_expect(&full[s], bytes);
      bulk_g2s(dst + doff, src, bytes, &full[s]);
    } else mbar_arrive(&full[s]);
  };
  for (int k = 0; k < S - 1; ++k) issue(k);
  float acc = 0.f;
  for (int k = 0; k < my; ++k) {
    issue(k + S - 1);
    mbar_wait(&full[k % S], (k / S) & 1);
    const float4 v = *reinterpret_cast<const float4*>(sm + (k % S) * (P * PIXB) + pix * PIXB + wnd * 16);
    acc += v.x + v.w;
    __syncthreads();                 // stage k%S free again before anyone refills it (issue(k+S) happens next iteration)
  }
  if (acc == 12345.f) sink[0] = acc;
}

// register path with U independent 16-B loads in flight per thread (thread = chunk ch of pixels n, n+stride, ...)
template <int U>
__global__ void __launch_bounds__(256) gather_ldg_u(Arrs a, int npix, float* sink) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long grp = gid / 48; const int ch = (int)(gid - grp * 48);
  float4 v[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const long long n = grp * U + u;
    v[u] = make_float4(0, 0, 0, 0);
    if (n < npix) {
      const int4 o = a.off[n];
      const float* src;
      if (ch < 20) src = a.geo0 + n * 384 + o.x + ch * 4;
      else if (ch < 40) src = a.geo1 + n * 192 + o.y + (ch - 20) * 4;
      else if (ch < 44) src = a.cor0 + n * 312 + o.z + (ch - 40) * 4;
      else src = a.cor1 + n * 156 + o.w + (ch - 44) * 4;
      asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(src));
    }
  }
  float acc = 0.f;
#pragma unroll
  for (int u = 0; u < U; ++u) acc += v[u].x + v[u].w;
  if (acc == 12345.678f) sink[0] = acc;
}
// plain streaming read of the same four arrays, whole rows (upper bound of what the memory system gives a reader)
__global__ void __launch_bounds__(256) stream_read(const float4* __restrict__ p, long long n4, float* sink) {
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x * 4) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const long long j = i + (long long)u * gridDim.x * blockDim.x; v[u] = j < n4 ? p[j] : make_float4(0, 0, 0, 0); }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += v[u].x + v[u].w;
  }
  if (acc == 12345.678f) sink[0] = acc;
}

// write side: 162 planes [C][npix] fp32 written as float4 (4 pixels x 1 channel per thread), tile of TP pixels per CTA
template <int TP>
__global__ void __launch_bounds__(256) write_planes_v4(float* __restrict__ out, long long npix, int C) {
  const long long p0 = (long long)blockIdx.x * TP;
  for (int i = threadIdx.x; i < (TP / 4) * C; i += blockDim.x) {
    const int c = i / (TP / 4), q = i % (TP / 4);
    if (p0 + q * 4 < npix) {
      float4 v = make_float4((float)i, 1.f, 2.f, 3.f);
      asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(out + (long long)c * npix + p0 + q * 4), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
    }
  }
}

template <class F> float timeit(F f, int reps = 10) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  static char* flush = nullptr; if (!flush) CK(cudaMalloc(&flush, 512 << 20));
  float best = 1e9, sum = 0;
  for (int r = 0; r < reps; ++r) {
    cudaMemsetAsync(flush, r, 512 << 20);
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; sum += ms;
  }
  CK(cudaGetLastError());
  return best * 1e3f;
}

int main() {
  const int N = 239616;
  float *g0, *g1, *c0, *c1, *sink; int4* off;
  CK(cudaMalloc(&g0, (size_t)N * 384 * 4)); CK(cudaMalloc(&g1, (size_t)N * 192 * 4));
  CK(cudaMalloc(&c0, (size_t)N * 312 * 4)); CK(cudaMalloc(&c1, (size_t)N * 156 * 4));
  CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&off, (size_t)N * 16));
  CK(cudaMemset(g0, 0, (size_t)N * 384 * 4)); CK(cudaMemset(g1, 0, (size_t)N * 192 * 4));
  CK(cudaMemset(c0, 0, (size_t)N * 312 * 4)); CK(cudaMemset(c1, 0, (size_t)N * 156 * 4));
  int4* h = (int4*)malloc((size_t)N * 16);
  srand(1);
  for (int i = 0; i < N; ++i) {
    h[i].x = (rand() % 39) * 8;         // tap start 0..38 of 48 (x 8 groups)
    h[i].y = (rand() % 15) * 8;         // 0..14 of 24
    h[i].z = (rand() % 75) * 4;         // 16-B aligned start, 16 floats inside 312
    h[i].w = (rand() % 36) * 4;
  }
  CK(cudaMemcpy(off, h, (size_t)N * 16, cudaMemcpyHostToDevice));
  Arrs a{g0, g1, c0, c1, off};
  const int ntiles = (N + P - 1) / P;
  const double useful = (double)N * (320 + 320 + 64 + 64 + 16);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  auto report = [&](const char* name, float us) { printf("%-44s %7.1f us  useful %6.1f GB/s\n", name, us, useful / us / 1e3); };
#define BULK(S)                                                                                                   \
  {                                                                                                               \
    CK(cudaFuncSetAttribute(gather_bulk<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, S * P * PIXB));           \
    float us = timeit([&] { gather_bulk<S><<<sms, 128, S * P * PIXB>>>(a, N, ntiles, sink); });                    \
    report("bulk (TMA 1-D) stages=" #S " 1 CTA/SM", us);                                                          \
  }
  BULK(2) BULK(3) BULK(4)
#define LDGSTS(S)                                                                                                 \
  {                                                                                                               \
    CK(cudaFuncSetAttribute(gather_ldgsts<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, S * P * PIXB));         \
    float us = timeit([&] { gather_ldgsts<S><<<sms, 256, S * P * PIXB>>>(a, N, ntiles, sink); });                  \
    report("cp.async 16B stages=" #S " 1 CTA/SM", us);                                                            \
  }
  LDGSTS(2) LDGSTS(3) LDGSTS(4)
  {
    CK(cudaFuncSetAttribute(gather_bulk<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * P * PIXB));
    float us = timeit([&] { gather_bulk<2><<<sms * 2, 128, 2 * P * PIXB>>>(a, N, ntiles, sink); });
    report("bulk stages=2, 2 CTA/SM", us);
  }
  {
    const long long threads = (long long)N * 48;
    float us = timeit([&] { gather_ldg<<<(unsigned)((threads + 255) / 256), 256>>>(a, N, sink); });
    report("ld.global.v4 one chunk per thread", us);
  }

#define BULKALL(S, OCC)                                                                                            \
  {                                                                                                               \
    CK(cudaFuncSetAttribute(gather_bulk_all<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, S * P * PIXB));       \
    float us = timeit([&] { gather_bulk_all<S><<<sms * OCC, 256, S * P * PIXB>>>(a, N, ntiles, sink); });          \
    report("bulk all-threads-issue stages=" #S " CTAs/SM=" #OCC, us);                                              \
  }
  BULKALL(2, 1) BULKALL(4, 1) BULKALL(2, 2) BULKALL(2, 3) BULKALL(1, 4)
  { const long long threads = ((long long)N / 2 + 1) * 48; float us = timeit([&] { gather_ldg_u<2><<<(unsigned)((threads + 255) / 256), 256>>>(a, N, sink); }); report("ld.global.v4 U=2 per thread", us); }
  { const long long threads = ((long long)N / 4 + 1) * 48; float us = timeit([&] { gather_ldg_u<4><<<(unsigned)((threads + 255) / 256), 256>>>(a, N, sink); }); report("ld.global.v4 U=4 per thread", us); }
  { const long long threads = ((long long)N / 8 + 1) * 48; float us = timeit([&] { gather_ldg_u<8><<<(unsigned)((threads + 255) / 256), 256>>>(a, N, sink); }); report("ld.global.v4 U=8 per thread", us); }
  { const long long n4 = (long long)N * 384 / 4; float us = timeit([&] { stream_read<<<sms * 8, 256>>>((const float4*)g0, n4, sink); }); printf("stream read of geo0 (368 MB)                 %7.1f us  %6.1f GB/s\n", us, (double)n4 * 16 / us / 1e3); }
  float* out; CK(cudaMalloc(&out, (size_t)N * 162 * 4));
  { float us = timeit([&] { write_planes_v4<64><<<(N + 63) / 64, 256>>>(out, N, 162); }); printf("WRITE 162 planes float4, tile 64 px : %7.1f us %6.1f GB/s\n", us, (double)N * 648 / us / 1e3); }
  { float us = timeit([&] { write_planes_v4<128><<<(N + 127) / 128, 256>>>(out, N, 162); }); printf("WRITE 162 planes float4, tile 128 px: %7.1f us %6.1f GB/s\n", us, (double)N * 648 / us / 1e3); }
  { float us = timeit([&] { write_planes_v4<256><<<(N + 255) / 256, 256>>>(out, N, 162); }); printf("WRITE 162 planes float4, tile 256 px: %7.1f us %6.1f GB/s\n", us, (double)N * 648 / us / 1e3); }
  return 0;
}
