// EXPERIMENT (not built into the library; kept for the record with its measurements).
// Outcome: bit-for-bit correct against the oracle (29 GPU tests), but SLOWER than the first fused kernel: 139-141 us vs
// 96-110 us at config 2.  ncu (--set full, source page): DRAM 27 %, L1 33 %, tensor 4 %; 24 % of all stall samples are
// the feature warps waiting on `raw_full`, i.e. on the gather.  cp.async.bulk is issued once per ACTIVE LANE from a
// divergent ELECT / R2UR.BROADCAST loop (~60 cycles per copy and warp), so 4 gather warps deliver a 32-pixel stage in
// ~2.8k cycles whatever the ring depth -- the same ~65 us floor the stand-alone gather experiment shows for bulk copies
// (tools/experiments/gather_bw.cu: 63-68 us) before any waiting is added.  The design therefore cannot beat the
// register-path kernel; the lean feature / MMA half (7 M L1 wavefronts instead of 15.7 M) would need an LDG-based gather
// to pay off.
//
// SURVEY 8(f) rank 1, second generation: Combined_Geo_Encoding_Volume.__call__ (models/coreContinuous_IGEV/geometry.py:34-60)
// fused with BasicMotionEncoder.convc1 + ReLU (update.py:78,85) for the IGEV shape (2 levels, 8 groups, radius 4).
//
// Why a second kernel.  ncu on the first one (lookup_c1_umma.cu): 15.7 M L1 wavefronts per launch -- 5.1 M from 4-byte
// window loads (8 lanes per pixel), 4.2 M from 4-byte operand stores (36 % bank-conflicted), 6.5 M from M = 128 MMAs over
// 64-pixel tiles -- the L1 / shared-memory pipe was 64 % busy and the kernel ran at 0.32-0.37 of the HBM roofline, the
// HBM itself at 50 %.  Measured separately (tools/experiments/gather_bw.cu): just READING the four windows of every
// pixel takes 57-65 us on a B200 however it is done, i.e. the layout-imposed floor of the whole kernel.
//
// This version moves the windows with the TMA unit and touches shared memory in 16-byte units only:
//   gather   4 warps: thread = (pixel, window); ONE cp.async.bulk per window (320 B geometry runs, 64 B correlation runs,
//            clipped to the row) straight into a 3-deep ring of 32-pixel raw stages, mbarrier completion -- no registers,
//            no L1 wavefronts, the copies of the next stages are always in flight
//   feature  8 warps: lane = (pixel, level, tap k): two 128-bit reads per tap give the 8 groups of taps k and k+1, the
//            interpolation runs in fp32, and the 8 results leave as ONE 128-bit store into the K-major, 128B-swizzled A
//            tile (K order level*96 + k*8 + g; correlation taps at level*96 + 72 + k) plus the second plane of the
//            arithmetic mode (16-bit lo, or the e5m2 pair encoding of AS_FMT_F16F8); 8 consecutive lanes = 8 consecutive
//            taps of one pixel: conflict-free reads and writes
//   MMA      M = 128 REAL pixels x N = 64 x K = 192, weights resident (TMA, once), 2 or 3 passes per K-step
//   epilogue 4 warps: TMEM -> bias + ReLU -> the hi / lo planes [N][64] convc2 consumes
// L1 wavefronts per launch: ~7 M (was 15.7 M); warp instructions ~8 M (was ~30 M).
#include <cstdlib>
#include "umma.cuh"

namespace {

constexpr int kTile = 128;                 // pixels per tile = rows of the MMA
constexpr int kStagePix = 32;              // pixels per raw stage
constexpr int kRawStages = 3;
constexpr int kPixBytes = 784;             // 320 + 320 + 64 + 64 + 16 pad: an ODD number of 16-byte chunks per pixel
constexpr int kRawStageBytes = kStagePix * kPixBytes;
constexpr int kGatherWarps = 4, kFeatWarps = 8, kEpiWarps = 4;
constexpr int kThreads = 32 * (kGatherWarps + kFeatWarps + 1 + kEpiWarps);     // 544
constexpr int kKB = 3;                     // K = 192 = 3 blocks of 64
constexpr int kABlock = kTile * 128;       // bytes of one [128 x 64] 16-bit K-block
constexpr int kBBlock = 64 * 128;
constexpr int kAPlane = kKB * kABlock;     // 49152
constexpr int kBPlane = kKB * kBBlock;     // 24576
constexpr int kG = 8, kR = 4, kTaps = 10;
constexpr int kRoundsPerStage = 20;        // 16 (k = 0..7 of 64 (pixel, level) pairs) + 2 (k = 8) + 2 (correlation)
constexpr int kParamWords = 8;             // per pixel: t0 of the 4 windows, fraction of the 4 windows
constexpr int kSmemBytes = 1024 + 2 * kAPlane + 2 * kBPlane + kRawStages * kRawStageBytes +
                           kRawStages * kStagePix * kParamWords * 4 + 256;
static_assert(kSmemBytes <= 227 * 1024, "lookup_c1_v2 shared memory");

struct V2Levels {
  const float* geo[2];
  const float* corr[2];
  int width[2];
  int pitch[2];
};

__device__ __forceinline__ void split_pos(float x, int& t0, float& f) {
  const float fl = floorf(x);
  f = x - fl;                                                           // sampler_kernel.cu:42
  t0 = (int)fminf(fmaxf(fl, -1.0e6f), 1.0e6f) - kR;                     // sampler_kernel.cu:47
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(umma::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void wait_backoff(uint64_t* bar, uint32_t parity) {
  while (!umma::mbar_try_wait(bar, parity)) __nanosleep(64);
}

// 8 consecutive K positions (kk % 8 == 0) of one A row: the 16-bit hi plane and the second plane of the mode
//   NS = 3: 16-bit lo = v - hi;   NS = 2: e5m2 pair encoding (common.cuh): [lo * 2^6 | hi * 2^-8] per 64-wide K block
template <int NS>
__device__ __forceinline__ void put8(uint32_t a_hi, int row, int kk, const float (&v)[8], bool f16) {
  const uint32_t blk = (uint32_t)(kk >> 6) * (uint32_t)kABlock + (uint32_t)row * 128u;
  const uint32_t sw = (uint32_t)(row & 7);
  uint32_t h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = as_cvt16x2(v[2 * i], v[2 * i + 1], f16);
  sts128(a_hi + blk + ((((uint32_t)(kk & 63) >> 3) ^ sw) << 4), h[0], h[1], h[2], h[3]);
  if (NS == 1) return;
  float hv[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) { hv[2 * i] = as_widen_lo16(h[i], f16); hv[2 * i + 1] = as_widen_hi16(h[i], f16); }
  const uint32_t a_lo = a_hi + (uint32_t)kAPlane;
  if (NS == 3) {
    uint32_t l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) l[i] = as_cvt16x2(v[2 * i] - hv[2 * i], v[2 * i + 1] - hv[2 * i + 1], f16);
    sts128(a_lo + blk + ((((uint32_t)(kk & 63) >> 3) ^ sw) << 4), l[0], l[1], l[2], l[3]);
  } else {
    const uint32_t b = (uint32_t)(kk & 63);                              // byte of the first e5m2 value inside the 128-B row
    const uint32_t lo0 = as_e5m2x4((v[0] - hv[0]) * kX8ActLoScale, (v[1] - hv[1]) * kX8ActLoScale, (v[2] - hv[2]) * kX8ActLoScale,
                                   (v[3] - hv[3]) * kX8ActLoScale);
    const uint32_t lo1 = as_e5m2x4((v[4] - hv[4]) * kX8ActLoScale, (v[5] - hv[5]) * kX8ActLoScale, (v[6] - hv[6]) * kX8ActLoScale,
                                   (v[7] - hv[7]) * kX8ActLoScale);
    const uint32_t hi0 = as_e5m2x4(hv[0] * kX8ActHiScale, hv[1] * kX8ActHiScale, hv[2] * kX8ActHiScale, hv[3] * kX8ActHiScale);
    const uint32_t hi1 = as_e5m2x4(hv[4] * kX8ActHiScale, hv[5] * kX8ActHiScale, hv[6] * kX8ActHiScale, hv[7] * kX8ActHiScale);
    sts64(a_lo + blk + ((((b >> 4)) ^ sw) << 4) + (b & 15u), lo0, lo1);
    sts64(a_lo + blk + ((((b >> 4) + 4u) ^ sw) << 4) + (b & 15u), hi0, hi1);
  }
}

template <int NS>
__global__ void __launch_bounds__(kThreads, 1)
lookup_c1_v2_kernel(const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo, const V2Levels lv,
                    int Dg, const float* __restrict__ disp, const float* __restrict__ coords, const float* __restrict__ bias,
                    __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, int HW, int W, int tiles_per_img,
                    int num_tiles, bool f16, int out_fmt) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_base = smem;                                  // hi plane | second plane
  uint8_t* b_hi = a_base + 2 * kAPlane;
  uint8_t* b_lo = b_hi + kBPlane;
  uint8_t* raw = b_lo + kBPlane;                           // [kRawStages][32 px][784 B]
  int* prm = reinterpret_cast<int*>(raw + kRawStages * kRawStageBytes);   // [kRawStages][8 words][32 px]
  uint64_t* bars = reinterpret_cast<uint64_t*>(prm + kRawStages * kStagePix * kParamWords);
  uint64_t* w_full = bars;
  uint64_t* raw_full = bars + 1;                           // [3] gather -> feature (128 arrivals + bytes)
  uint64_t* raw_empty = raw_full + kRawStages;             // [3] feature warps done with the stage (8 arrivals)
  uint64_t* a_full = raw_empty + kRawStages;               // feature warps -> MMA (8 arrivals)
  uint64_t* a_empty = a_full + 1;                          // MMAs of the tile have read A (commit)
  uint64_t* acc_full = a_empty + 1;                        // [2] MMA -> epilogue (commit)
  uint64_t* acc_empty = acc_full + 2;                      // [2] epilogue drained the accumulator (4 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  constexpr int kMmaWarp = kGatherWarps + kFeatWarps;

  if (tid == 0) {
    umma::prefetch_tmap(&tmW_hi);
    if (NS > 1) umma::prefetch_tmap(&tmW_lo);
    umma::mbar_init(w_full, 1);
    for (int i = 0; i < kRawStages; ++i) { umma::mbar_init(raw_full + i, kGatherWarps * 32); umma::mbar_init(raw_empty + i, kFeatWarps); }
    umma::mbar_init(a_full, kFeatWarps);
    umma::mbar_init(a_empty, 1);
    for (int i = 0; i < 2; ++i) { umma::mbar_init(acc_full + i, 1); umma::mbar_init(acc_empty + i, kEpiWarps); }
    umma::fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    umma::tmem_alloc(tmem_slot, 128);
    umma::tmem_relinquish();
  }
  // zero the operand tile once: K padding (72 + 16 used of every 96) and, for ragged images, never-written rows
  for (int i = tid; i < (2 * kAPlane) / 16; i += kThreads) reinterpret_cast<uint4*>(a_base)[i] = make_uint4(0, 0, 0, 0);
  umma::fence_proxy_async();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  const int my_tiles = blockIdx.x < num_tiles ? (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const uint32_t raw_s = umma::smem_u32(raw);

  if (warp < kGatherWarps) {
    // ================= gather: thread = (pixel of the stage, window) =================
    const int pix = tid >> 2, wnd = tid & 3;               // wnd: 0 geo L0, 1 geo L1, 2 corr L0, 3 corr L1
    const int l = wnd & 1;
    const bool is_geo = wnd < 2;
    const float sc = l ? 0.5f : 1.0f;
    for (int s = 0; s < my_tiles * 4; ++s) {
      const int slot = s % kRawStages;
      const uint32_t ph = (uint32_t)(s / kRawStages) & 1u;
      const int t = blockIdx.x + (s >> 2) * gridDim.x;
      const int b = t / tiles_per_img;
      const int p = (t - b * tiles_per_img) * kTile + (s & 3) * kStagePix + pix;
      const bool inside = p < HW;
      const long long n = (long long)b * HW + (inside ? p : 0);
      float d = 0.f, c = 0.f;
      if (inside) {
        d = __ldg(disp + n);
        c = coords ? __ldg(coords + n) : (float)(p % W);
      }
      int t0; float f;
      if (is_geo) split_pos(d * sc, t0, f);                 // geometry.py:43  x = disp / 2^l (+ dx)
      else split_pos(c * sc - d * sc, t0, f);               // geometry.py:52  x = coords / 2^l - disp / 2^l (+ dx)
      if (!inside) t0 = -(1 << 20);                         // every tap out of range -> zero features
      umma::mbar_wait(raw_empty + slot, ph ^ 1);
      int* pp = prm + slot * (kStagePix * kParamWords);
      pp[wnd * kStagePix + pix] = t0;
      pp[(4 + wnd) * kStagePix + pix] = __float_as_int(f);
      const uint32_t dst0 = raw_s + (uint32_t)(slot * kRawStageBytes + pix * kPixBytes);
      uint32_t bytes = 0;
      const float* src = nullptr;
      uint32_t dst = 0;
      if (is_geo) {
        const int Dl = Dg >> l;
        const int lo = max(t0, 0), hi = min(t0 + kTaps, Dl);
        if (hi > lo) {
          src = lv.geo[l] + (n * Dl + lo) * kG;
          bytes = (uint32_t)(hi - lo) * 32u;
          dst = dst0 + (uint32_t)(l * 320 + (lo - t0) * 32);
        }
      } else {
        const int c0 = as_floor4(t0) * 4;
        const int lo = max(c0, 0), hi = min(c0 + 16, lv.pitch[l]);
        if (hi > lo) {
          src = lv.corr[l] + n * lv.pitch[l] + lo;
          bytes = (uint32_t)(hi - lo) * 4u;
          dst = dst0 + (uint32_t)(640 + l * 64 + (lo - c0) * 4);
        }
      }
      if (bytes) {
        umma::mbar_expect_tx(raw_full + slot, bytes);       // counts as this thread's arrival
        bulk_g2s(dst, src, bytes, raw_full + slot);
      } else {
        umma::mbar_arrive(raw_full + slot);
      }
    }
  } else if (warp < kMmaWarp) {
    // ================= feature warps =================
    const int fw = warp - kGatherWarps;                    // 0..7
    const uint32_t a_hi = umma::smem_u32(a_base);
    for (int s = 0; s < my_tiles * 4; ++s) {
      const int slot = s % kRawStages;
      const uint32_t ph = (uint32_t)(s / kRawStages) & 1u;
      const int quarter = s & 3;
      if (quarter == 0) wait_backoff(a_empty, (uint32_t)((s >> 2) & 1) ^ 1u);   // the previous tile's MMAs have read A
      umma::mbar_wait(raw_full + slot, ph);
      const int* pp = prm + slot * (kStagePix * kParamWords);
      const uint32_t rs = raw_s + (uint32_t)(slot * kRawStageBytes);
      // rounds of this stage, dealt round-robin with a per-stage rotation so the 20 rounds spread evenly over 8 warps
      for (int r = (fw + s * 4) & 7; r < kRoundsPerStage; r += kFeatWarps) {
        if (r < 18) {
          int pix, l, k;
          if (r < 16) { const int qi = r * 4 + (lane >> 3); pix = qi >> 1; l = qi & 1; k = lane & 7; }
          else { pix = lane; l = r - 16; k = 8; }
          const int t0 = pp[l * kStagePix + pix];
          const float f = __int_as_float(pp[(4 + l) * kStagePix + pix]), omf = 1.0f - f;
          const unsigned Dl = (unsigned)(Dg >> l);
          const uint32_t base = rs + (uint32_t)(pix * kPixBytes + l * 320 + k * 32);
          // lanes 0-3 of a quarter warp read the low half of their tap first, lanes 4-7 the high half: the 8 lanes of
          // every 128-bit read hit 8 distinct 16-byte bank groups
          const uint32_t h0 = (uint32_t)((k >> 2) & 1) * 16u;
          const float4 a0 = lds128(base + h0), a1 = lds128(base + (h0 ^ 16u));
          const float4 b0 = lds128(base + 32u + h0), b1 = lds128(base + 32u + (h0 ^ 16u));
          const bool swap = h0 != 0;
          const float4 alo = swap ? a1 : a0, ahi = swap ? a0 : a1, blo = swap ? b1 : b0, bhi = swap ? b0 : b1;
          const bool va = (unsigned)(t0 + k) < Dl, vb = (unsigned)(t0 + k + 1) < Dl;
          const float av[8] = {alo.x, alo.y, alo.z, alo.w, ahi.x, ahi.y, ahi.z, ahi.w};
          const float bv[8] = {blo.x, blo.y, blo.z, blo.w, bhi.x, bhi.y, bhi.z, bhi.w};
          float v[8];
#pragma unroll
          for (int g = 0; g < 8; ++g) v[g] = (va ? av[g] : 0.f) * omf + (vb ? bv[g] : 0.f) * f;
          put8<NS>(a_hi, quarter * kStagePix + pix, l * 96 + k * 8, v, f16);
        } else {
          const int pix = lane, l = r - 18;
          const int t0 = pp[(2 + l) * kStagePix + pix];
          const float f = __int_as_float(pp[(6 + l) * kStagePix + pix]), omf = 1.0f - f;
          const unsigned Wl = (unsigned)lv.width[l];
          const int off = t0 - as_floor4(t0) * 4;
          const uint32_t base = rs + (uint32_t)(pix * kPixBytes + 640 + l * 64 + off * 4);
          float w[kTaps];
#pragma unroll
          for (int j = 0; j < kTaps; ++j) {
            const float x = lds32(base + (uint32_t)(j * 4));
            w[j] = (unsigned)(t0 + j) < Wl ? x : 0.f;
          }
          float v[8], v2[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = w[k] * omf + w[k + 1] * f;
          v2[0] = w[8] * omf + w[9] * f;
#pragma unroll
          for (int k = 1; k < 8; ++k) v2[k] = 0.f;
          put8<NS>(a_hi, quarter * kStagePix + pix, l * 96 + 72, v, f16);
          put8<NS>(a_hi, quarter * kStagePix + pix, l * 96 + 80, v2, f16);
        }
      }
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(raw_empty + slot);
      if (quarter == 3) {
        umma::fence_proxy_async();                           // generic-proxy writes of A -> visible to the tensor-core proxy
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(a_full);
      }
    }
  } else if (warp == kMmaWarp) {
    // ================= MMA issuer =================
    if (lane == 0) {
      umma::mbar_expect_tx(w_full, (uint32_t)kBPlane * (NS > 1 ? 2u : 1u));
      for (int kb = 0; kb < kKB; ++kb) {
        umma::tma_load_2d(b_hi + kb * kBBlock, &tmW_hi, w_full, kb * 64, 0);
        if (NS > 1) umma::tma_load_2d(b_lo + kb * kBBlock, &tmW_lo, w_full, kb * 64, 0);
      }
      umma::mbar_wait(w_full, 0);
      const uint32_t idesc = umma::idesc_16_f32(128, 64, f16);
      const uint32_t idesc8 = umma::idesc_e5m2_f32(128, 64);
      const uint32_t ah = umma::smem_u32(a_base), al = ah + (uint32_t)kAPlane;
      const uint32_t bh = umma::smem_u32(b_hi), bl = umma::smem_u32(b_lo);
      for (int i = 0; i < my_tiles; ++i) {
        const int acc = i & 1;
        wait_backoff(a_full, (uint32_t)i & 1u);
        wait_backoff(acc_empty + acc, (uint32_t)((i >> 1) & 1) ^ 1u);
        umma::tc_fence_after();
        const uint32_t td = tmem_d + (uint32_t)(acc * 64);
#pragma unroll
        for (int kb = 0; kb < kKB; ++kb) {
          const uint32_t dah0 = umma::desc_lo_sw128(ah + kb * kABlock), dal0 = umma::desc_lo_sw128(al + kb * kABlock);
          const uint32_t dbh0 = umma::desc_lo_sw128(bh + kb * kBBlock), dbl0 = umma::desc_lo_sw128(bl + kb * kBBlock);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t k2 = 2u * (uint32_t)k;
            umma::mma_ss_lohi<false, false>(td, dah0 + k2, dbh0 + k2, idesc, (kb | k) != 0 ? 1u : 0u);
            if (NS == 3) {
              umma::mma_ss_lohi<false, false>(td, dah0 + k2, dbl0 + k2, idesc, 1u);
              umma::mma_ss_lohi<false, false>(td, dal0 + k2, dbh0 + k2, idesc, 1u);
            } else if (NS == 2) {
              umma::mma_ss_lohi<false, true>(td, dal0 + k2, dbl0 + k2, idesc8, 1u);
            }
          }
        }
        umma::mma_commit(a_empty);
        umma::mma_commit(acc_full + acc);
      }
    }
  } else {
    // ================= epilogue: 4 warps <-> TMEM lanes; bias + ReLU -> hi / lo planes [N][64] =================
    const int q = warp & 3;                                 // TMEM lane quarter of this warp (warp % 4)
    for (int i = 0; i < my_tiles; ++i) {
      const int acc = i & 1;
      wait_backoff(acc_full + acc, (uint32_t)((i >> 1) & 1));
      umma::tc_fence_after();
      const int t = blockIdx.x + i * gridDim.x;
      const int b = t / tiles_per_img;
      const int p = (t - b * tiles_per_img) * kTile + q * 32 + lane;
      float v[2][32];
      const uint32_t ta = tmem_d + (uint32_t)(acc * 64) + ((uint32_t)(q * 32) << 16);
      umma::tmem_ld_32x32(ta, v[0]);
      umma::tmem_ld_32x32(ta + 32, v[1]);
      umma::tmem_ld_wait();
      umma::tc_fence_before();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(acc_empty + acc);
      if (p < HW) {
        const long long o = ((long long)b * HW + p) * 64;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
          for (int jj = 0; jj < 32; jj += 8) {
            uint32_t h[4];
            float y[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) y[e] = fmaxf(v[hf][jj + e] + __ldg(bias + hf * 32 + jj + e), 0.f);
#pragma unroll
            for (int e = 0; e < 4; ++e) h[e] = as_cvt16x2(y[2 * e], y[2 * e + 1], f16);
            *reinterpret_cast<uint4*>(out_hi + o + hf * 32 + jj) = make_uint4(h[0], h[1], h[2], h[3]);
            if (out_lo) as_store_lo8(out_lo, o + hf * 32 + jj, y, h, out_fmt);
          }
        }
      }
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) umma::tmem_dealloc(tmem_d, 128);
}

}  // namespace

// K order of the convc1 weights this kernel expects: channel (level l, group g, tap k) at K = l*96 + k*8 + g, correlation
// tap k of level l at K = l*96 + 72 + k; everything else zero.  (The first-generation kernel uses l*96 + g*10 + k.)
bool as_lookup_c1_v2_enabled() {
  static const bool on = !(getenv("AS_LOOKUP_C1_V2") && getenv("AS_LOOKUP_C1_V2")[0] == '0');
  return on;
}

int as_lookup_c1_v2_launch(const float* const* geo_levels, int Dg, const float* const* corr_levels, const int* corr_widths,
                           const int* corr_pitches, const float* disp, const float* coords, const void* w_hi, const void* w_lo,
                           const float* bias, int nsplit, void* out_hi, void* out_lo, int B, int H, int W, cudaStream_t st) {
  if (nsplit < 1 || nsplit > 3) return AS_ERR_BAD_ARG;
  if (nsplit > 1 && (!w_lo || !out_lo)) return AS_ERR_BAD_ARG;
  const int fmt = as_operand_fmt_internal();
  if ((nsplit == 2) != (fmt == AS_FMT_F16F8) && nsplit != 1) return AS_ERR_BAD_ARG;
  V2Levels lv{};
  for (int l = 0; l < 2; ++l) {
    if (!geo_levels[l] || !corr_levels[l] || corr_pitches[l] < corr_widths[l] || (corr_pitches[l] & 3)) return AS_ERR_BAD_ARG;
    if (!as_aligned16(geo_levels[l]) || !as_aligned16(corr_levels[l])) return AS_ERR_ALIGNMENT;
    lv.geo[l] = geo_levels[l]; lv.corr[l] = corr_levels[l]; lv.width[l] = corr_widths[l]; lv.pitch[l] = corr_pitches[l];
  }
  CUtensorMap tW_hi, tW_lo;
  const uint64_t dims[2] = {192, 64};
  const uint64_t str[1] = {192 * 2};
  const uint32_t box[2] = {64u, 64u};
  int rc;
  if ((rc = umma::make_tmap_bf16(&tW_hi, w_hi, 2, dims, str, box)) != AS_OK) return rc;
  if (nsplit > 1) {
    if ((rc = umma::make_tmap_bf16(&tW_lo, w_lo, 2, dims, str, box)) != AS_OK) return rc;
  } else {
    tW_lo = tW_hi;
  }
  const int HW = H * W;
  const int tiles_per_img = as_ceil_div(HW, kTile);
  const long long nt = (long long)tiles_per_img * B;
  if (nt >= (1LL << 29)) return AS_ERR_INDEX_RANGE;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = nt < sms ? (int)nt : sms;
  const bool f16 = as_operand_f16_internal() != 0;
  cudaError_t e;
#define AS_V2_LAUNCH(NS_)                                                                                               \
  do {                                                                                                                  \
    e = cudaFuncSetAttribute(lookup_c1_v2_kernel<NS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);        \
    if (e != cudaSuccess) return (int)e;                                                                                \
    lookup_c1_v2_kernel<NS_><<<grid, kThreads, kSmemBytes, st>>>(tW_hi, tW_lo, lv, Dg, disp, coords, bias,              \
                                                                 (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, HW, W, \
                                                                 tiles_per_img, (int)nt, f16, fmt);                     \
  } while (0)
  if (nsplit == 3) AS_V2_LAUNCH(3);
  else if (nsplit == 2) AS_V2_LAUNCH(2);
  else AS_V2_LAUNCH(1);
#undef AS_V2_LAUNCH
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}
