// a3 / a4: per-iteration radius lookups over the correlation pyramid (RAFT) and over the combined
// geometry-encoding + correlation pyramids (IGEV), forward and adjoint.
//
// Reference behaviour: CorrBlock1D.__call__ (models/corePrune_RAFT/geometry.py:24-43),
// Combined_Geo_Encoding_Volume.__call__ (models/coreContinuous_IGEV/geometry.py:34-60), both through
// bilinear_sampler -> F.grid_sample (models/*/utils/utils.py:59-72); index math as in
// sampler/sampler_kernel.cu:39-58.
//
// B200 design (HBM-bound kernels):
//  * ONE launch per iteration produces the final [B,C,H,W] tensor for all levels (the reference runs
//    ~20-40 small torch kernels and 1 host sync per level per sampler).
//  * Every pixel needs a private, data-dependent window of its own pyramid row, so lane=pixel loads are
//    32-lines-per-request.  Instead a CTA loads the windows of P consecutive pixels with 128-bit loads in
//    which neighbouring lanes cover the SAME pixel's window (sector-exact traffic), transposes them
//    through shared memory (conflict-free strides), and then lane=pixel threads interpolate and write
//    each output channel as full 128-byte lines.
//  * The geometry pyramid is stored [pixel][disparity][group] (see pyramid.cu) so the (2r+2) taps x 8
//    groups of one pixel are ONE contiguous, 32-byte aligned 320-byte run: zero sector waste.
#include <cstdlib>
#include "common.cuh"

namespace {

struct LevelSet {
  const float* ptr[AS_MAX_LEVELS];
  int width[AS_MAX_LEVELS];
  int pitch[AS_MAX_LEVELS];
};
struct LevelSetRW {
  float* ptr[AS_MAX_LEVELS];
  int width[AS_MAX_LEVELS];
  int pitch[AS_MAX_LEVELS];
};

__device__ __forceinline__ void split_pos(float x, int r, int& t0, float& f) {
  const float fl = floorf(x);
  f = x - fl;                                                   // sampler_kernel.cu:42
  t0 = (int)fminf(fmaxf(fl, -1.0e6f), 1.0e6f) - r;              // sampler_kernel.cu:47 (clamped: far OOB)
}

__device__ __forceinline__ float level_scale(int l) { return __int_as_float((127 - l) << 23); }  // 2^-l

// ------------------------------------------------------------------------------------------------
// RAFT: correlation pyramid only
// ------------------------------------------------------------------------------------------------
constexpr int kRP = 128;   // pixels per CTA
constexpr int kRSP = 130;  // smem row stride (floats): 4*kRSP % 32 == 8 -> conflict-free transposing stores

__global__ void __launch_bounds__(kRP) corr_lookup_fwd_kernel(LevelSet lv, int L, const float* __restrict__ disp,
                                                              const float* __restrict__ coords,
                                                              float* __restrict__ out, int HW, int W, int r,
                                                              int NV) {
  extern __shared__ __align__(16) float s_win[];  // [L][4*NV][kRSP]
  __shared__ int s_t0[AS_MAX_LEVELS][kRP];
  __shared__ float s_f[AS_MAX_LEVELS][kRP];
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * kRP;
  const long long nbase = (long long)b * HW;
  {
    const int p = p0 + tid;
    float d = 0.f, c = 0.f;
    if (p < HW) {
      d = disp[nbase + p];
      c = coords ? coords[nbase + p] : (float)(p % W);
    }
    for (int l = 0; l < L; ++l) {
      const float sc = level_scale(l);
      int t0; float f;
      split_pos(c * sc - d * sc, r, t0, f);   // geometry.py:35 (coords/2^i - disp/2^i)
      s_t0[l][tid] = t0;
      s_f[l][tid] = f;
    }
  }
  __syncthreads();

  // phase 1: 128-bit window loads; lanes = 8 pixels x 4 quads (NV==4) -> transposed smem
  const int items = L * kRP * NV;
  for (int i0 = tid; i0 < items; i0 += kRP * 4) {
    float4 v[4];
    int dst[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int item = i0 + u * kRP;
      dst[u] = -1;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (item < items) {
        const int q = item % NV;
        const int rest = item / NV;
        const int pix = rest % kRP;
        const int l = rest / kRP;
        const int p = p0 + pix;
        dst[u] = (l * 4 * NV + q * 4) * kRSP + pix;
        const int c0 = as_floor4(s_t0[l][pix]) * 4 + q * 4;
        if (p < HW && c0 >= 0 && c0 < lv.pitch[l])
          v[u] = __ldg(reinterpret_cast<const float4*>(lv.ptr[l] + (nbase + p) * lv.pitch[l] + c0));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (dst[u] >= 0) {
        s_win[dst[u]] = v[u].x;
        s_win[dst[u] + kRSP] = v[u].y;
        s_win[dst[u] + 2 * kRSP] = v[u].z;
        s_win[dst[u] + 3 * kRSP] = v[u].w;
      }
    }
  }
  __syncthreads();

  // phase 2: lane = pixel
  const int p = p0 + tid;
  if (p >= HW) return;
  const int K = 2 * r + 1;
  float* o = out + (long long)b * L * K * HW + p;
  for (int l = 0; l < L; ++l) {
    const int t0 = s_t0[l][tid];
    const float f = s_f[l][tid], omf = 1.0f - f;
    const int off = t0 - as_floor4(t0) * 4;
    const float* w = s_win + (l * 4 * NV + off) * kRSP + tid;
    const int Wl = lv.width[l];
    float prev = (t0 >= 0 && t0 < Wl) ? w[0] : 0.f;
    for (int k = 0; k < K; ++k) {
      const int x1 = t0 + k + 1;
      const float cur = (x1 >= 0 && x1 < Wl) ? w[(k + 1) * kRSP] : 0.f;
      o[(long long)(l * K + k) * HW] = prev * omf + cur * f;
      prev = cur;
    }
  }
}

// fast path: radius 4, L <= 4.  32 pixels per CTA (each output channel = one 128-byte line per warp),
// 4 warps; a warp-wide 128-bit load covers 8 pixels x 64 contiguous bytes.
constexpr int kFP = 32, kFT = 128, kFSP = 34;

template <int L>
__global__ void __launch_bounds__(kFT) corr_lookup_fwd_r4_kernel(LevelSet lv, const float* __restrict__ disp,
                                                                 const float* __restrict__ coords,
                                                                 float* __restrict__ out, int HW, int W) {
  __shared__ __align__(16) float s_win[L * 16 * kFSP];
  __shared__ int s_t0[L][kFP];
  __shared__ float s_f[L][kFP];
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * kFP;
  const long long nbase = (long long)b * HW;
  if (tid < kFP) {
    const int p = p0 + tid;
    float d = 0.f, c = 0.f;
    if (p < HW) {
      d = disp[nbase + p];
      c = coords ? coords[nbase + p] : (float)(p % W);
    }
#pragma unroll
    for (int l = 0; l < L; ++l) {
      const float sc = level_scale(l);
      int t0; float f;
      split_pos(c * sc - d * sc, 4, t0, f);
      s_t0[l][tid] = t0;
      s_f[l][tid] = f;
    }
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  const int pix8 = lane >> 2, q4 = lane & 3;
  constexpr int kJobs = L * (kFP / 8);
  constexpr int kPerWarp = (kJobs + 3) / 4;
  float4 v[kPerWarp];
  int dst[kPerWarp];
#pragma unroll
  for (int i = 0; i < kPerWarp; ++i) {
    const int job = warp + 4 * i;
    dst[i] = -1;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (job < kJobs) {
      const int l = job / (kFP / 8);
      const int pix = (job - l * (kFP / 8)) * 8 + pix8;
      const int p = p0 + pix;
      const int c0 = as_floor4(s_t0[l][pix]) * 4 + q4 * 4;
      dst[i] = (l * 16 + q4 * 4) * kFSP + pix;
      if (p < HW && c0 >= 0 && c0 < lv.pitch[l])
        v[i] = as_ldg_stream(reinterpret_cast<const float4*>(lv.ptr[l] + (nbase + p) * lv.pitch[l] + c0));
    }
  }
#pragma unroll
  for (int i = 0; i < kPerWarp; ++i) {
    if (dst[i] >= 0) {
      float* sp = s_win + dst[i];
      sp[0] = v[i].x; sp[kFSP] = v[i].y; sp[2 * kFSP] = v[i].z; sp[3 * kFSP] = v[i].w;
    }
  }
  __syncthreads();
  const int pix = lane, part = warp;
  const int p = p0 + pix;
  if (p >= HW) return;
  float* o = out + (long long)b * L * 9 * HW + p;
  for (int l = part; l < L; l += 4) {
    const int t0 = s_t0[l][pix];
    const float f = s_f[l][pix], omf = 1.0f - f;
    const int off = t0 - as_floor4(t0) * 4;
    const float* w = s_win + (l * 16 + off) * kFSP + pix;
    const int Wl = lv.width[l];
    float* oc = o + (long long)(l * 9) * HW;
    float prev = (t0 >= 0 && t0 < Wl) ? w[0] : 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int x1 = t0 + k + 1;
      const float cur = (x1 >= 0 && x1 < Wl) ? w[(k + 1) * kFSP] : 0.f;
      oc[(long long)k * HW] = prev * omf + cur * f;
      prev = cur;
    }
  }
}

// adjoint w.r.t. the levels; one thread per (pixel, level); rows are pixel-private -> no atomics
__global__ void corr_lookup_bwd_kernel(LevelSetRW lv, int L, const float* __restrict__ disp,
                                       const float* __restrict__ coords, const float* __restrict__ gout,
                                       int HW, int W, int r, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int l = (int)(idx / ((long long)total / L));
  const long long n = idx - (long long)l * (total / L);
  const int b = (int)(n / HW);
  const int p = (int)(n - (long long)b * HW);
  const float d = disp[n];
  const float c = coords ? coords[n] : (float)(p % W);
  const float sc = level_scale(l);
  int t0; float f;
  split_pos(c * sc - d * sc, r, t0, f);
  const float omf = 1.0f - f;
  const int K = 2 * r + 1;
  const float* g = gout + ((long long)b * L * K + (long long)l * K) * HW + p;
  float* row = lv.ptr[l] + n * lv.pitch[l];
  const int Wl = lv.width[l];
  float gm1 = 0.f;
  for (int j = 0; j <= K; ++j) {
    const float g0 = (j < K) ? g[(long long)j * HW] : 0.f;
    const int x1 = t0 + j;
    if (x1 >= 0 && x1 < Wl) row[x1] += gm1 * f + g0 * omf;   // sampler_kernel.cu:90-103
    gm1 = g0;
  }
}

// ------------------------------------------------------------------------------------------------
// IGEV: geometry volume (8 groups) + correlation, fast path G == 8, r == 4
// ------------------------------------------------------------------------------------------------
constexpr int kGP = 64;    // pixels per tile: each output channel is written as 256 contiguous bytes
constexpr int kGT = 192;   // 6 warps (4-byte emit)
constexpr int kGT4 = 192;  // 128-bit emit: the same 6 warps; L*(G+1)*16 = 288 (row, pixel-quad) tasks per tile at L = 2
#ifndef AS_GEO_MINB
#define AS_GEO_MINB 2
#endif
constexpr int kG = 8, kR = 4, kTaps = 10, kK = 9;

struct GeoTile {
  int b, p0;
};

// Persistent, software-pipelined: while tile i is interpolated and written (phase C), the 128-bit window loads
// of tile i+1 are already in flight in registers and the disparities of tile i+2 are being fetched, so every CTA
// keeps HBM reads and writes overlapped instead of alternating between a load phase and a store phase.
// V4 = 128-bit output stores: one thread interpolates 4 consecutive pixels of one (level, group) row and writes every
// channel as float4 (a warp instruction = 2 x 256 contiguous bytes).  B200 needs 16 bytes per lane to approach the copy
// rate on plane-strided writes: 162 planes written with 4-byte stores top out at 3.6 TB/s, with 16-byte stores at
// 5.8 TB/s (tools/experiments/gather_bw.cu).  Needs H*W % 4 == 0.
template <int L, int NT, bool V4>
__global__ void __launch_bounds__(NT, AS_GEO_MINB) geo_lookup_fwd_kernel(LevelSet geo, int Dg, LevelSet corr,
                                                                const float* __restrict__ disp,
                                                                const float* __restrict__ coords,
                                                                float* __restrict__ out, int HW, int W,
                                                                int tiles_per_img, int num_tiles) {
  constexpr int kGT = NT;
  constexpr int kGSP = 66;                               // smem row stride (floats): conflict-free transposing stores, 8-byte aligned rows
  extern __shared__ __align__(16) float smem[];
  float* s_geo = smem;                                   // [L][kTaps*kG = 80][kGSP]
  float* s_cor = smem + L * kTaps * kG * kGSP;           // [L][16][kGSP]
  __shared__ __align__(16) int s_tg[2][L][kGP], s_tc[2][L][kGP];
  __shared__ __align__(16) float s_fg[2][L][kGP], s_fc[2][L][kGP];

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int pix8 = lane >> 2, q4 = lane & 3;
  constexpr int kGeoJobs = L * (kGP / 8) * 5;    // (level, pixel-group, chunk-quad): 20 chunks of 16 B per pixel
  constexpr int kCorJobs = L * (kGP / 8);        // 4 chunks per pixel
  constexpr int kJobs = kGeoJobs + kCorJobs;
  constexpr int kWarps = kGT / 32;
  constexpr int kPerWarp = (kJobs + kWarps - 1) / kWarps;
  constexpr int C = L * (kG + 1) * kK;

  auto tile_of = [&](int t, GeoTile& g) {
    g.b = t / tiles_per_img;
    g.p0 = (t - g.b * tiles_per_img) * kGP;
  };
  // (0) disparity / coordinate of this thread's pixel in tile t (threads < kGP)
  auto fetch_dc = [&](int t, float& d, float& c) {
    d = 0.f; c = 0.f;
    if (t < num_tiles && tid < kGP) {
      GeoTile g; tile_of(t, g);
      const int p = g.p0 + tid;
      if (p < HW) {
        d = __ldg(disp + (long long)g.b * HW + p);
        c = coords ? __ldg(coords + (long long)g.b * HW + p) : (float)(p % W);
      }
    }
  };
  auto put_params = [&](int buf, float d, float c) {
    if (tid < kGP) {
#pragma unroll
      for (int l = 0; l < L; ++l) {
        const float sc = level_scale(l);
        int t0; float f;
        split_pos(d * sc, kR, t0, f);                 // geometry.py:43  x0 = dx + disp/2^i
        s_tg[buf][l][tid] = t0; s_fg[buf][l][tid] = f;
        split_pos(c * sc - d * sc, kR, t0, f);        // geometry.py:52
        s_tc[buf][l][tid] = t0; s_fc[buf][l][tid] = f;
      }
    }
  };
  float4 v[kPerWarp];
  // (B) issue the window loads of tile t (parameters in buffer `buf`); a warp instruction = 8 pixels x 64 B
  auto issue_loads = [&](int t, int buf) {
    GeoTile g; tile_of(t, g);
    const long long nbase = (long long)g.b * HW;
#pragma unroll
    for (int i = 0; i < kPerWarp; ++i) {
      const int job = warp + i * kWarps;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (job < kGeoJobs) {
        const int l = job / ((kGP / 8) * 5);
        const int rem = job - l * ((kGP / 8) * 5);
        const int pg = rem / 5, qq = rem - pg * 5;
        const int pix = pg * 8 + pix8;
        const int q = qq * 4 + q4;            // chunk 0..19 : tap j = q>>1, groups (q&1)*4..+3
        const int p = g.p0 + pix;
        const int Dl = Dg >> l;
        const int x1 = s_tg[buf][l][pix] + (q >> 1);
        if (p < HW && x1 >= 0 && x1 < Dl)
          v[i] = as_ldg_stream(reinterpret_cast<const float4*>(geo.ptr[l] + ((nbase + p) * Dl + x1) * kG + (q & 1) * 4));
      } else if (job < kJobs) {
        const int cj = job - kGeoJobs;
        const int l = cj / (kGP / 8);
        const int pix = (cj - l * (kGP / 8)) * 8 + pix8;
        const int p = g.p0 + pix;
        const int c0 = as_floor4(s_tc[buf][l][pix]) * 4 + q4 * 4;
        if (p < HW && c0 >= 0 && c0 < corr.pitch[l])
          v[i] = as_ldg_stream(reinterpret_cast<const float4*>(corr.ptr[l] + (nbase + p) * corr.pitch[l] + c0));
      }
    }
  };
  // (A) registers -> transposed shared-memory windows
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < kPerWarp; ++i) {
      const int job = warp + i * kWarps;
      float* sp = nullptr;
      if (job < kGeoJobs) {
        const int l = job / ((kGP / 8) * 5);
        const int rem = job - l * ((kGP / 8) * 5);
        const int pg = rem / 5, qq = rem - pg * 5;
        sp = s_geo + (l * kTaps * kG + (qq * 4 + q4) * 4) * kGSP + pg * 8 + pix8;
      } else if (job < kJobs) {
        const int cj = job - kGeoJobs;
        const int l = cj / (kGP / 8);
        sp = s_cor + (l * 16 + q4 * 4) * kGSP + (cj - l * (kGP / 8)) * 8 + pix8;
      }
      if (sp) { sp[0] = v[i].x; sp[kGSP] = v[i].y; sp[2 * kGSP] = v[i].z; sp[3 * kGSP] = v[i].w; }
    }
  };
  // (C) lane = pixel: interpolate and write every channel of tile t as full lines
  auto emit = [&](int t, int buf) {
    GeoTile g; tile_of(t, g);
    const int pix = tid & (kGP - 1);
    const int part = tid / kGP;          // 0..2
    const int p = g.p0 + pix;
    if (p >= HW) return;
    float* o = out + (long long)g.b * C * HW + p;
    for (int row = part; row < L * (kG + 1); row += kGT / kGP) {
      const int l = row / (kG + 1);
      const int gi = row - l * (kG + 1);
      if (gi < kG) {
        const int t0 = s_tg[buf][l][pix];
        const float f = s_fg[buf][l][pix], omf = 1.0f - f;
        const int Dl = Dg >> l;
        const float* w = s_geo + (l * kTaps * kG + gi) * kGSP + pix;
        float* oc = o + (long long)(l * (kG + 1) * kK + gi * kK) * HW;
        float prev = (t0 >= 0 && t0 < Dl) ? w[0] : 0.f;
#pragma unroll
        for (int k = 0; k < kK; ++k) {
          const int x1 = t0 + k + 1;
          const float cur = (x1 >= 0 && x1 < Dl) ? w[(k + 1) * kG * kGSP] : 0.f;
          oc[(long long)k * HW] = prev * omf + cur * f;
          prev = cur;
        }
      } else {
        const int t0 = s_tc[buf][l][pix];
        const float f = s_fc[buf][l][pix], omf = 1.0f - f;
        const int Wl = corr.width[l];
        const int off = t0 - as_floor4(t0) * 4;
        const float* w = s_cor + (l * 16 + off) * kGSP + pix;
        float* oc = o + (long long)(l * (kG + 1) * kK + kG * kK) * HW;
        float prev = (t0 >= 0 && t0 < Wl) ? w[0] : 0.f;
#pragma unroll
        for (int k = 0; k < kK; ++k) {
          const int x1 = t0 + k + 1;
          const float cur = (x1 >= 0 && x1 < Wl) ? w[(k + 1) * kGSP] : 0.f;
          oc[(long long)k * HW] = prev * omf + cur * f;
          prev = cur;
        }
      }
    }
  };

  // (C') 128-bit emit: thread = (row, 4 consecutive pixels)
  auto emit4 = [&](int t, int buf) {
    GeoTile g; tile_of(t, g);
    for (int task = tid; task < L * (kG + 1) * (kGP / 4); task += kGT) {
      const int row = task / (kGP / 4);
      const int pq = (task - row * (kGP / 4)) * 4;
      const int p = g.p0 + pq;
      if (p >= HW) continue;                               // HW % 4 == 0: a quad is inside or outside as a whole
      const int l = row / (kG + 1);
      const int gi = row - l * (kG + 1);
      float* o = out + (long long)g.b * C * HW + p;
      if (gi < kG) {
        const int4 t0 = *reinterpret_cast<const int4*>(&s_tg[buf][l][pq]);
        const float4 f = *reinterpret_cast<const float4*>(&s_fg[buf][l][pq]);
        const unsigned Dl = (unsigned)(Dg >> l);
        const float* w = s_geo + (l * kTaps * kG + gi) * kGSP + pq;
        float* oc = o + (long long)(l * (kG + 1) * kK + gi * kK) * HW;
        auto ld4 = [](const float* q) {                     // rows are 8-byte aligned (stride 66): two 64-bit reads
          const float2 a = *reinterpret_cast<const float2*>(q), b = *reinterpret_cast<const float2*>(q + 2);
          return make_float4(a.x, a.y, b.x, b.y);
        };
        float4 prev = ld4(w);
        prev.x = (unsigned)t0.x < Dl ? prev.x : 0.f; prev.y = (unsigned)t0.y < Dl ? prev.y : 0.f;
        prev.z = (unsigned)t0.z < Dl ? prev.z : 0.f; prev.w = (unsigned)t0.w < Dl ? prev.w : 0.f;
#pragma unroll
        for (int k = 0; k < kK; ++k) {
          float4 cur = ld4(w + (k + 1) * kG * kGSP);
          cur.x = (unsigned)(t0.x + k + 1) < Dl ? cur.x : 0.f; cur.y = (unsigned)(t0.y + k + 1) < Dl ? cur.y : 0.f;
          cur.z = (unsigned)(t0.z + k + 1) < Dl ? cur.z : 0.f; cur.w = (unsigned)(t0.w + k + 1) < Dl ? cur.w : 0.f;
          as_stg_stream4(reinterpret_cast<float4*>(oc + (long long)k * HW),
                         make_float4(prev.x * (1.0f - f.x) + cur.x * f.x, prev.y * (1.0f - f.y) + cur.y * f.y,
                                     prev.z * (1.0f - f.z) + cur.z * f.z, prev.w * (1.0f - f.w) + cur.w * f.w));
          prev = cur;
        }
      } else {
        const int4 t0v = *reinterpret_cast<const int4*>(&s_tc[buf][l][pq]);
        const float4 fv = *reinterpret_cast<const float4*>(&s_fc[buf][l][pq]);
        const int t0a[4] = {t0v.x, t0v.y, t0v.z, t0v.w};
        const float fa[4] = {fv.x, fv.y, fv.z, fv.w};
        const unsigned Wl = (unsigned)corr.width[l];
        float* oc = o + (long long)(l * (kG + 1) * kK + kG * kK) * HW;
        const float* wq[4];
        float prev[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {                       // every pixel has its own row offset inside the 16-float window
          const int off = t0a[i] - as_floor4(t0a[i]) * 4;
          wq[i] = s_cor + (l * 16 + off) * kGSP + pq + i;
          prev[i] = (unsigned)t0a[i] < Wl ? wq[i][0] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < kK; ++k) {
          float r[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float cur = (unsigned)(t0a[i] + k + 1) < Wl ? wq[i][(k + 1) * kGSP] : 0.f;
            r[i] = prev[i] * (1.0f - fa[i]) + cur * fa[i];
            prev[i] = cur;
          }
          as_stg_stream4(reinterpret_cast<float4*>(oc + (long long)k * HW), make_float4(r[0], r[1], r[2], r[3]));
        }
      }
    }
  };

  // ---- prologue: parameters + loads of the first tile, disparities of the second
  int t = blockIdx.x;
  if (t >= num_tiles) return;
  float d_nxt, c_nxt;
  fetch_dc(t, d_nxt, c_nxt);
  put_params(0, d_nxt, c_nxt);
  __syncthreads();
  issue_loads(t, 0);
  fetch_dc(t + gridDim.x, d_nxt, c_nxt);
  int buf = 0;
  for (; t < num_tiles; t += gridDim.x, buf ^= 1) {
    const int tn = t + gridDim.x;
    stash();                                  // tile t windows -> smem
    put_params(buf ^ 1, d_nxt, c_nxt);        // tile t+1 parameters
    __syncthreads();
    if (tn < num_tiles) issue_loads(tn, buf ^ 1);          // in flight during emit(t)
    fetch_dc(tn + gridDim.x, d_nxt, c_nxt);
    if (V4) emit4(t, buf); else emit(t, buf);
    __syncthreads();
  }
}

// generic (any G, radius, L): one thread per (pixel, level, row) with row in [0,G] (G = the corr row)
__global__ void geo_lookup_fwd_generic_kernel(LevelSet geo, int G, int Dg, LevelSet corr, int L,
                                              const float* __restrict__ disp, const float* __restrict__ coords,
                                              float* __restrict__ out, int HW, int W, int r, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long N = total / ((long long)L * (G + 1));
  const long long n = idx % N;
  const int rest = (int)(idx / N);
  const int g = rest % (G + 1);
  const int l = rest / (G + 1);
  const int b = (int)(n / HW);
  const int p = (int)(n - (long long)b * HW);
  const float d = disp[n];
  const float c = coords ? coords[n] : (float)(p % W);
  const float sc = level_scale(l);
  const int K = 2 * r + 1;
  int t0; float f;
  const float* base; int stride, limit;
  if (g < G) {
    split_pos(d * sc, r, t0, f);
    limit = Dg >> l;
    base = geo.ptr[l] + n * limit * G + g;
    stride = G;
  } else {
    split_pos(c * sc - d * sc, r, t0, f);
    limit = corr.width[l];
    base = corr.ptr[l] + n * corr.pitch[l];
    stride = 1;
  }
  const float omf = 1.0f - f;
  float* o = out + ((long long)b * L * (G + 1) * K + (long long)(l * (G + 1) + g) * K) * HW + p;
  float prev = (t0 >= 0 && t0 < limit) ? base[(long long)t0 * stride] : 0.f;
  for (int k = 0; k < K; ++k) {
    const int x1 = t0 + k + 1;
    const float cur = (x1 >= 0 && x1 < limit) ? base[(long long)x1 * stride] : 0.f;
    o[(long long)k * HW] = prev * omf + cur * f;
    prev = cur;
  }
}

__global__ void geo_lookup_bwd_kernel(LevelSetRW ggeo, int G, int Dg, LevelSetRW gcorr, int L,
                                      const float* __restrict__ disp, const float* __restrict__ coords,
                                      const float* __restrict__ gout, int HW, int W, int r, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long N = total / ((long long)L * (G + 1));
  const long long n = idx % N;
  const int rest = (int)(idx / N);
  const int g = rest % (G + 1);
  const int l = rest / (G + 1);
  const int b = (int)(n / HW);
  const int p = (int)(n - (long long)b * HW);
  const float d = disp[n];
  const float c = coords ? coords[n] : (float)(p % W);
  const float sc = level_scale(l);
  const int K = 2 * r + 1;
  int t0; float f;
  float* base; int stride, limit;
  if (g < G) {
    split_pos(d * sc, r, t0, f);
    limit = Dg >> l;
    base = ggeo.ptr[l] + n * limit * G + g;
    stride = G;
  } else {
    split_pos(c * sc - d * sc, r, t0, f);
    limit = gcorr.width[l];
    base = gcorr.ptr[l] + n * gcorr.pitch[l];
    stride = 1;
  }
  const float omf = 1.0f - f;
  const float* go = gout + ((long long)b * L * (G + 1) * K + (long long)(l * (G + 1) + g) * K) * HW + p;
  float gm1 = 0.f;
  for (int j = 0; j <= K; ++j) {
    const float g0 = (j < K) ? go[(long long)j * HW] : 0.f;
    const int x1 = t0 + j;
    if (x1 >= 0 && x1 < limit) base[(long long)x1 * stride] += gm1 * f + g0 * omf;
    gm1 = g0;
  }
}

// Same adjoint for the IGEV shape (G = 8, radius 4), coalesced on both sides: the generic kernel above walks a pixel's
// taps with one thread per (pixel, level, group), i.e. 1.5 KB between the lanes of every read-modify-write (408 us per
// call at config-5 size).  Here a CTA stages the 162 gradient rows of 32 pixels through shared memory (reads: lanes =
// pixels), then 8 lanes = the 8 groups of one (pixel, level) update one 32-byte sector per tap.
template <int L>
__global__ void __launch_bounds__(256) geo_lookup_bwd_tiled_kernel(LevelSetRW ggeo, int Dg, LevelSetRW gcorr,
                                                                   const float* __restrict__ disp,
                                                                   const float* __restrict__ coords,
                                                                   const float* __restrict__ gout, int HW, int W,
                                                                   int tiles_per_img) {
  constexpr int G = 8, R = 4, K = 9, C = L * (G + 1) * K, TP = 32;
  __shared__ float s[C][TP + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / tiles_per_img;
  const int p0 = (blockIdx.x - b * tiles_per_img) * TP;
  const float* go = gout + (long long)b * C * HW + p0;
  for (int c = warp; c < C; c += 8) s[c][lane] = (p0 + lane < HW) ? __ldg(go + (long long)c * HW + lane) : 0.f;
  __syncthreads();
  const long long nbase = (long long)b * HW + p0;
  // geometry part: item = (pixel, level), 8 lanes per item
  const int g = lane & 7;
  for (int item = warp * 4 + (lane >> 3); item < TP * L; item += 32) {
    const int l = item / TP, pp = item - l * TP;
    if (p0 + pp >= HW) continue;
    const long long n = nbase + pp;
    int t0; float f;
    split_pos(__ldg(disp + n) * level_scale(l), R, t0, f);
    const float omf = 1.0f - f;
    const int Dl = Dg >> l;
    float* base = ggeo.ptr[l] + n * Dl * G + g;
    const float* sg = &s[(l * (G + 1) + g) * K][pp];
    float gm1 = 0.f;
#pragma unroll
    for (int j = 0; j <= K; ++j) {
      const float g0 = (j < K) ? sg[j * (TP + 1)] : 0.f;
      const int x1 = t0 + j;
      if (x1 >= 0 && x1 < Dl) base[(long long)x1 * G] += gm1 * f + g0 * omf;
      gm1 = g0;
    }
  }
  // correlation part: thread = (pixel, level)
  if (tid < TP * L) {
    const int l = tid / TP, pp = tid - l * TP;
    if (p0 + pp < HW) {
      const long long n = nbase + pp;
      const float d = __ldg(disp + n);
      const float c = coords ? __ldg(coords + n) : (float)((p0 + pp) % W);
      const float sc = level_scale(l);
      int t0; float f;
      split_pos(c * sc - d * sc, R, t0, f);
      const float omf = 1.0f - f;
      const int limit = gcorr.width[l];
      float* base = gcorr.ptr[l] + n * gcorr.pitch[l];
      const float* sg = &s[(l * (G + 1) + G) * K][pp];
      float gm1 = 0.f;
#pragma unroll
      for (int j = 0; j <= K; ++j) {
        const float g0 = (j < K) ? sg[j * (TP + 1)] : 0.f;
        const int x1 = t0 + j;
        if (x1 >= 0 && x1 < limit) base[x1] += gm1 * f + g0 * omf;
        gm1 = g0;
      }
    }
  }
}

__global__ void lookup_taps_kernel(const float* __restrict__ disp, const float* __restrict__ coords, int HW,
                                   int W, int r, int level, int kind, int32_t* __restrict__ tap0,
                                   float* __restrict__ frac, long long N) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int p = (int)(n % HW);
  const float d = disp[n];
  const float c = coords ? coords[n] : (float)(p % W);
  const float sc = level_scale(level);
  int t0; float f;
  split_pos(kind == 0 ? d * sc : c * sc - d * sc, r, t0, f);
  tap0[n] = t0;
  frac[n] = f;
}

int fill_levels(LevelSet& ls, const float* const* ptrs, const int* widths, const int* pitches, int L) {
  for (int l = 0; l < L; ++l) {
    if (!ptrs[l] || widths[l] < 0 || pitches[l] < widths[l]) return AS_ERR_BAD_ARG;
    if ((pitches[l] & 3) || !as_aligned16(ptrs[l])) return AS_ERR_ALIGNMENT;
    ls.ptr[l] = ptrs[l];
    ls.width[l] = widths[l];
    ls.pitch[l] = pitches[l];
  }
  return AS_OK;
}

}  // namespace

extern "C" int as_corr_lookup_fwd(const float* const* levels, const int* widths, const int* pitches,
                                  int num_levels, const float* disp, const float* coords, float* out, int B,
                                  int H, int W, int radius, as_stream_t stream) {
  if (!levels || !widths || !pitches || !disp || !out) return AS_ERR_BAD_ARG;
  if (num_levels < 1 || num_levels > AS_MAX_LEVELS || B <= 0 || H <= 0 || W <= 0 || radius < 0) return AS_ERR_BAD_ARG;
  if (B > 65535 || radius > 15) return AS_ERR_UNSUPPORTED;
  LevelSet ls{};
  int rc = fill_levels(ls, levels, widths, pitches, num_levels);
  if (rc != AS_OK) return rc;
  const int HW = H * W;
  if (radius == 4 && num_levels <= 4) {
    dim3 grid(as_ceil_div(HW, kFP), B);
    cudaStream_t st = as_cu(stream);
    switch (num_levels) {
      case 1: corr_lookup_fwd_r4_kernel<1><<<grid, kFT, 0, st>>>(ls, disp, coords, out, HW, W); break;
      case 2: corr_lookup_fwd_r4_kernel<2><<<grid, kFT, 0, st>>>(ls, disp, coords, out, HW, W); break;
      case 3: corr_lookup_fwd_r4_kernel<3><<<grid, kFT, 0, st>>>(ls, disp, coords, out, HW, W); break;
      default: corr_lookup_fwd_r4_kernel<4><<<grid, kFT, 0, st>>>(ls, disp, coords, out, HW, W); break;
    }
    AS_RETURN_IF_LAUNCH_FAILED();
    return AS_OK;
  }
  const int NV = (2 * radius + 5 + 3) / 4;
  const size_t smem = sizeof(float) * num_levels * 4 * NV * kRSP;
  if (smem > 200 * 1024) return AS_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(corr_lookup_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid(as_ceil_div(HW, kRP), B);
  corr_lookup_fwd_kernel<<<grid, kRP, smem, as_cu(stream)>>>(ls, num_levels, disp, coords, out, HW, W, radius, NV);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_corr_lookup_bwd(float* const* g_levels, const int* widths, const int* pitches, int num_levels,
                                  const float* disp, const float* coords, const float* g_out, int B, int H,
                                  int W, int radius, as_stream_t stream) {
  if (!g_levels || !widths || !pitches || !disp || !g_out) return AS_ERR_BAD_ARG;
  if (num_levels < 1 || num_levels > AS_MAX_LEVELS || B <= 0 || H <= 0 || W <= 0 || radius < 0) return AS_ERR_BAD_ARG;
  LevelSetRW ls{};
  for (int l = 0; l < num_levels; ++l) {
    if (!g_levels[l] || pitches[l] < widths[l]) return AS_ERR_BAD_ARG;
    ls.ptr[l] = g_levels[l]; ls.width[l] = widths[l]; ls.pitch[l] = pitches[l];
  }
  const long long total = (long long)B * H * W * num_levels;
  corr_lookup_bwd_kernel<<<(unsigned)as_ceil_div_ll(total, 256), 256, 0, as_cu(stream)>>>(
      ls, num_levels, disp, coords, g_out, H * W, W, radius, total);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_geo_lookup_fwd(const float* const* geo_levels, int G, int Dg, const float* const* corr_levels,
                                 const int* corr_widths, const int* corr_pitches, int num_levels,
                                 const float* disp, const float* coords, float* out, int B, int H, int W,
                                 int radius, as_stream_t stream) {
  if (!geo_levels || !corr_levels || !corr_widths || !corr_pitches || !disp || !out) return AS_ERR_BAD_ARG;
  if (num_levels < 1 || num_levels > AS_MAX_LEVELS || B <= 0 || H <= 0 || W <= 0 || radius < 0 || G < 1 || Dg < 1)
    return AS_ERR_BAD_ARG;
  if (B > 65535) return AS_ERR_UNSUPPORTED;
  LevelSet cs{}, gs{};
  int rc = fill_levels(cs, corr_levels, corr_widths, corr_pitches, num_levels);
  if (rc != AS_OK) return rc;
  for (int l = 0; l < num_levels; ++l) {
    if (!geo_levels[l]) return AS_ERR_BAD_ARG;
    if (!as_aligned16(geo_levels[l])) return AS_ERR_ALIGNMENT;
    gs.ptr[l] = geo_levels[l]; gs.width[l] = Dg >> l; gs.pitch[l] = (Dg >> l) * G;
  }
  const int HW = H * W;
  cudaStream_t st = as_cu(stream);
  if (G == kG && radius == kR && num_levels <= 4) {
    const int tiles_per_img = as_ceil_div(HW, kGP);
    const long long nt = (long long)tiles_per_img * B;
    if (nt >= (1LL << 31)) return AS_ERR_INDEX_RANGE;
    const int num_tiles = (int)nt;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static const bool allow_v4 = !(getenv("AS_GEO_LOOKUP_V4") && getenv("AS_GEO_LOOKUP_V4")[0] == '0');   // A/B knob
    const bool v4 = allow_v4 && (HW % 4 == 0) && as_aligned16(out);
    const size_t smem = sizeof(float) * num_levels * (kTaps * kG + 16) * 66;
#define AS_LAUNCH_GEO_K(KERNEL, NT)                                                                            \
  {                                                                                                            \
    cudaError_t e = cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
    if (e != cudaSuccess) return (int)e;                                                                       \
    int occ = 1;                                                                                               \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, KERNEL, NT, smem);                                     \
    if (occ < 1) occ = 1;                                                                                      \
    const int grid = num_tiles < occ * sms ? num_tiles : occ * sms; /* persistent: all CTAs resident */        \
    KERNEL<<<grid, NT, smem, st>>>(gs, Dg, cs, disp, coords, out, HW, W, tiles_per_img, num_tiles);            \
  }
#define AS_LAUNCH_GEO(LV)                                                                                      \
  case LV:                                                                                                     \
    if (v4) AS_LAUNCH_GEO_K((geo_lookup_fwd_kernel<LV, kGT4, true>), kGT4)                                     \
    else AS_LAUNCH_GEO_K((geo_lookup_fwd_kernel<LV, kGT, false>), kGT)                                         \
    break;
    switch (num_levels) {
      AS_LAUNCH_GEO(1)
      AS_LAUNCH_GEO(2)
      AS_LAUNCH_GEO(3)
      AS_LAUNCH_GEO(4)
    }
#undef AS_LAUNCH_GEO
#undef AS_LAUNCH_GEO_K
  } else {
    const long long total = (long long)B * HW * num_levels * (G + 1);
    geo_lookup_fwd_generic_kernel<<<(unsigned)as_ceil_div_ll(total, 256), 256, 0, st>>>(
        gs, G, Dg, cs, num_levels, disp, coords, out, HW, W, radius, total);
  }
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_geo_lookup_bwd(float* const* g_geo_levels, int G, int Dg, float* const* g_corr_levels,
                                 const int* corr_widths, const int* corr_pitches, int num_levels,
                                 const float* disp, const float* coords, const float* g_out, int B, int H,
                                 int W, int radius, as_stream_t stream) {
  if (!g_geo_levels || !g_corr_levels || !corr_widths || !corr_pitches || !disp || !g_out) return AS_ERR_BAD_ARG;
  if (num_levels < 1 || num_levels > AS_MAX_LEVELS || B <= 0 || H <= 0 || W <= 0 || radius < 0 || G < 1 || Dg < 1)
    return AS_ERR_BAD_ARG;
  LevelSetRW cs{}, gs{};
  for (int l = 0; l < num_levels; ++l) {
    if (!g_geo_levels[l] || !g_corr_levels[l] || corr_pitches[l] < corr_widths[l]) return AS_ERR_BAD_ARG;
    cs.ptr[l] = g_corr_levels[l]; cs.width[l] = corr_widths[l]; cs.pitch[l] = corr_pitches[l];
    gs.ptr[l] = g_geo_levels[l]; gs.width[l] = Dg >> l; gs.pitch[l] = (Dg >> l) * G;
  }
  static const bool tiled_ok = !(getenv("AS_GEO_LOOKUP_BWD_TILED") && getenv("AS_GEO_LOOKUP_BWD_TILED")[0] == '0');   // A/B knob
  if (tiled_ok && G == 8 && radius == 4 && (num_levels == 1 || num_levels == 2) && (long long)B * as_ceil_div(H * W, 32) < (1LL << 31)) {
    const int tiles_per_img = as_ceil_div(H * W, 32);
    const unsigned grid = (unsigned)((long long)B * tiles_per_img);
    if (num_levels == 2)
      geo_lookup_bwd_tiled_kernel<2><<<grid, 256, 0, as_cu(stream)>>>(gs, Dg, cs, disp, coords, g_out, H * W, W, tiles_per_img);
    else
      geo_lookup_bwd_tiled_kernel<1><<<grid, 256, 0, as_cu(stream)>>>(gs, Dg, cs, disp, coords, g_out, H * W, W, tiles_per_img);
    AS_RETURN_IF_LAUNCH_FAILED();
    return AS_OK;
  }
  const long long total = (long long)B * H * W * num_levels * (G + 1);
  geo_lookup_bwd_kernel<<<(unsigned)as_ceil_div_ll(total, 256), 256, 0, as_cu(stream)>>>(
      gs, G, Dg, cs, num_levels, disp, coords, g_out, H * W, W, radius, total);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_lookup_taps(const float* disp, const float* coords, int B, int H, int W, int radius, int level,
                              int kind, int32_t* tap0, float* frac, as_stream_t stream) {
  if (!disp || !tap0 || !frac || B <= 0 || H <= 0 || W <= 0 || level < 0 || level >= 30) return AS_ERR_BAD_ARG;
  const long long N = (long long)B * H * W;
  lookup_taps_kernel<<<(unsigned)as_ceil_div_ll(N, 256), 256, 0, as_cu(stream)>>>(disp, coords, H * W, W, radius,
                                                                               level, kind, tap0, frac, N);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}
