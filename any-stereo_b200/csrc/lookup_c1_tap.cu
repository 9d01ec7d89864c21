// SURVEY 8(f) rank 1, IGEV shape (2 levels, 8 groups, radius 4): Combined_Geo_Encoding_Volume.__call__
// (models/coreContinuous_IGEV/geometry.py:34-60) fused with BasicMotionEncoder.convc1 + ReLU (update.py:78,85), the
// "tap-major" generation of lookup_c1_umma.cu.
//
// What the first kernel paid for (ncu, profiles/ncu_full_lookup_convc1_r01_summary.csv): 15.7 M L1 wavefronts per launch,
// 5.1 M of them 4-byte window loads (8 lanes per pixel, one 32-byte sector per request and pixel), 4.2 M 4-byte operand
// stores (36 % bank-conflicted) and 6.5 M operand reads of M = 128 MMAs over 64-pixel tiles (half the rows don't-care).
// This one keeps the register path for the gather (the bulk-copy gather of tools/experiments/lookup_c1_v2.cu is
// issue-bound) but moves 16 bytes per lane everywhere:
//   gather   a warp round = 4 pixels x 2 levels = 8 windows; lane = (window, q): five ld.global.nc.v4 fetch chunks
//            4i+q of the window's 20 (a chunk = 4 of the 8 groups of one tap), i.e. 64 contiguous bytes per window and
//            instruction; one more v4 fetches the lane's quarter of the 64-byte aligned span of the correlation row.
//            The loads of the warp's NEXT tile are in flight while the current one is interpolated (two register
//            buffers).  ptxas puts every global load of the loop on one scoreboard, so the first use of a buffer waits
//            for ALL loads in flight: the loop therefore touches the current buffer first (one instruction that reads
//            a loaded register -- everything older has landed), only then issues the next tile's loads (their addresses
//            carry a data dependency on the touch, so they cannot be hoisted above it), then consumes from registers
//            without waiting.  (Issued the other way round -- prefetch, then consume -- a round took exactly load
//            latency + consumption, 3.4 us: no overlap at all.)
//   interp   tap k+1 of the same groups sits two lanes away: ONE shfl.xor(2) per value (lanes q < 2 forward their next
//            chunk, lanes q >= 2 their current one); the unaligned correlation taps go through 64 bytes of per-window
//            scratch.
//   operand  K order (level l, tap k, group g) -> l*96 + k*8 + g, correlation taps at l*96 + 72 + k: a lane's 4 results
//            are 8 contiguous bytes of the K-major 128B-swizzled A tile (st.shared.v2 per plane, conflict-free across the
//            rows {m, m+2, m+4, m+6} a round covers).
//   bias     K index 81 (a pad position of level 0) carries the constant 1 in the A tile and the bias (hi/lo split like
//            every weight) in the weight matrix: the epilogue has no bias loads (they missed L1 under the gather
//            traffic and made the epilogue, 7 us per tile, the bottleneck of the first version).
//   MMA      tiles of 80 pixels = 20 rounds, one per producer warp, in TWO 60-KB A stages (hi + lo planes): producers
//            never wait for the tensor core and drift apart, so loads, interpolation and MMAs of neighbouring tiles
//            overlap.  The M = 128 MMA reads 48 rows past each 80-row K-block (the next block, inside the allocation);
//            those rows land in TMEM lanes 80..127, which nobody reads.  An SS-mode MMA costs about the same whatever
//            N <= 256 is, so the split products are TWO instructions per K step instead of three:
//            A_hi x [W_hi | W_lo] with N = 128 (accumulator columns 0..63 and 64..127), then A_lo x W_hi into columns
//            0..63; the epilogue adds the halves.  Two TMEM accumulators (2 x 128 columns).
#include "umma.cuh"

namespace {

constexpr int kProdWarps = 20;             // one round (4 pixels) of every tile each
constexpr int kEpiWarps = 3;               // warps 20..22 <-> TMEM lanes 0..79 (warp & 3 == lane quarter)
constexpr int kMmaWarp = kProdWarps + 3;   // 24 warps = 6 per scheduler: 80 registers per thread
constexpr int kThreads = 32 * (kProdWarps + 4);
constexpr int kTile = 80;                  // pixels per tile; the MMA has M = 128 (48 don't-care rows)
constexpr int kStages = 2;
constexpr int kKB = 3;                     // K = 192 = 3 blocks of 64
constexpr int kNOut = 64;
constexpr int kABlock = kTile * 128;       // bytes of one [80 x 64] 16-bit K-block (10 swizzle atoms)
constexpr int kAPlane = kKB * kABlock;
constexpr int kBBlock = kNOut * 128;       // one plane of one K-block; a K-block holds [hi rows | lo rows]
constexpr int kBPlane = kKB * kBBlock;
constexpr int kAccCols = 128;              // per accumulator: hi-weight product | lo-weight product
constexpr int kScratch = kProdWarps * 512; // 8 windows x 64 B per producer warp
constexpr int kAStage = 2 * kAPlane;       // hi | lo
constexpr int kSmemBytes = 1024 + kStages * kAStage + 2 * kBPlane + kScratch + 256;
constexpr int kR = 4;
static_assert(kSmemBytes <= 227 * 1024, "lookup_c1_tap shared memory");

#ifdef AS_TAP_TRACE
// timeline instrumentation (experiment builds only: make EXTRA=-DAS_TAP_TRACE): CTA 0 records clock64 at the hand-over points
__device__ long long* g_tap_trace = nullptr;
#define TAP_TRACE(who, k, ev)                                                                                   \
  do {                                                                                                          \
    if (blockIdx.x == 0 && lane == 0 && (k) < 16 && g_tap_trace) g_tap_trace[((who) * 16 + (k)) * 8 + (ev)] = clock64(); \
  } while (0)
#else
#define TAP_TRACE(who, k, ev) do {} while (0)
#endif

struct TapLevels {
  const float* geo[2];
  const float* corr[2];
  int width[2];
  int pitch[2];
};

__device__ __forceinline__ void split_pos(float x, int& t0, float& f) {
  const float fl = floorf(x);
  f = x - fl;
  t0 = (int)fminf(fmaxf(fl, -1.0e6f), 1.0e6f) - kR;
}

__device__ __forceinline__ float4 ldg4_or_zero(const float4* p, bool valid) {
  float4 r;
#ifdef AS_TAP_NOLOAD                         // experiment: no gather traffic (addresses still computed)
  const float f = valid ? __int_as_float((int)(reinterpret_cast<uintptr_t>(p) & 0x3f800000u)) : 0.f;
  return make_float4(f, f, f, f);
#endif
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\tmov.f32 %2, 0f00000000;\n\tmov.f32 %3, 0f00000000;\n\t"
      "@p ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];\n\t}"     // every sector is used once, by this instruction
      : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
      : "l"(p), "r"((int)valid));
  return r;
}

struct PixelIn {
  float d, c;
  long long n;                             // flat pixel index (0 when !inside)
  bool inside;
};

// what one lane holds for one round: its chunks of the geometry window, its quarter of the correlation span
struct Round {
  float4 g[5];
  float4 cw;
  float fg, fc;
  int ca;                                  // offset (0..3) of the first correlation tap inside the 16 staged floats
  int crem;                                // row elements left from this lane's chunk start (masks the row end)
};

__device__ __forceinline__ void wait_backoff(uint64_t* bar, uint32_t parity) {
  while (!umma::mbar_try_wait(bar, parity)) __nanosleep(100);
}

template <bool kF16>
__device__ __forceinline__ void put2(uint32_t addr, float v0, float v1, bool split) {
  const uint32_t h = as_cvt16x2(v0, v1, kF16);
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(h) : "memory");
  if (split) {
    const uint32_t l = as_cvt16x2(v0 - as_widen_lo16(h, kF16), v1 - as_widen_hi16(h, kF16), kF16);
    asm volatile("st.shared.b32 [%0+%2], %1;" ::"r"(addr), "r"(l), "n"(kAPlane) : "memory");
  }
}
template <bool kF16>
__device__ __forceinline__ void put4(uint32_t addr, float v0, float v1, float v2, float v3, bool split) {
  const uint32_t h0 = as_cvt16x2(v0, v1, kF16), h1 = as_cvt16x2(v2, v3, kF16);
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(h0), "r"(h1) : "memory");
  if (split) {
    const uint32_t l0 = as_cvt16x2(v0 - as_widen_lo16(h0, kF16), v1 - as_widen_hi16(h0, kF16), kF16);
    const uint32_t l1 = as_cvt16x2(v2 - as_widen_lo16(h1, kF16), v3 - as_widen_hi16(h1, kF16), kF16);
    asm volatile("st.shared.v2.b32 [%0+%3], {%1, %2};" ::"r"(addr), "r"(l0), "r"(l1), "n"(kAPlane) : "memory");
  }
}

// byte offset of K index kidx inside row `row` of an A stage: K-block | row | swizzled byte inside the 128-byte row
__device__ __forceinline__ uint32_t a_offset(int kidx, int row) {
  return (uint32_t)((kidx >> 6) * kABlock + row * 128) + ((((uint32_t)(kidx & 63)) << 1) ^ ((uint32_t)(row & 7) << 4));
}

template <bool kF16, bool kSplit>
__global__ void __launch_bounds__(kThreads, 1)
geo_lookup_convc1_tap_kernel(const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                             const TapLevels lv, int Dg, const float* __restrict__ disp, const float* __restrict__ coords,
                             const float* __restrict__ bias, __nv_bfloat16* __restrict__ out_hi,
                             __nv_bfloat16* __restrict__ out_lo, int HW, int W, int tiles_per_img, int num_tiles, int out_fmt) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_base = smem;                              // kStages x (hi plane [3][80 x 128 B] | lo plane)
  uint8_t* b_w = smem + kStages * kAStage;             // [3 K-blocks][hi 64 rows | lo 64 rows] x 128 B
  float* scratch = reinterpret_cast<float*>(b_w + 2 * kBPlane);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(scratch) + kScratch);
  uint64_t* w_full = bars;                             // weights landed
  uint64_t* a_full = bars + 1;                         // [2] producers -> MMA
  uint64_t* a_empty = bars + 3;                        // [2] MMAs of the tile have read the stage -> producers
  uint64_t* acc_full = bars + 5;                       // [2] MMA -> epilogue
  uint64_t* acc_empty = bars + 7;                      // [2] epilogue drained TMEM -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  constexpr bool split = kSplit;                       // a template parameter: no run-time predicate around the MMAs

  if (tid == 0) {
    umma::prefetch_tmap(&tmW_hi);
    umma::mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(a_full + i, kProdWarps);
      umma::mbar_init(a_empty + i, 1);
      umma::mbar_init(acc_full + i, 1);
      umma::mbar_init(acc_empty + i, kEpiWarps * 32);
    }
    umma::fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    umma::tmem_alloc(tmem_slot, 2 * kAccCols);
    umma::tmem_relinquish();
  }
  // zero the A stages once: the pad channels are never written again and must read as 0
  for (int i = tid; i < (kStages * kAStage) / 16; i += kThreads) reinterpret_cast<uint4*>(a_base)[i] = make_uint4(0, 0, 0, 0);
  umma::fence_proxy_async();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  const int my_tiles = blockIdx.x < num_tiles ? (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp < kProdWarps) {
    // ================= producers =================
    const int q = lane & 3, lvl = (lane >> 2) & 1, pw = lane >> 3;
    const float sc = lvl ? 0.5f : 1.0f;
    const int Dl = Dg >> lvl;
    const float4* gbase = reinterpret_cast<const float4*>(lvl ? lv.geo[1] : lv.geo[0]);
    const float* cbase = lvl ? lv.corr[1] : lv.corr[0];
    const int Wl = lvl ? lv.width[1] : lv.width[0];
    const int pitch = lvl ? lv.pitch[1] : lv.pitch[0];
    const uint32_t a_s = umma::smem_u32(a_base);
    const uint32_t scr = umma::smem_u32(scratch) + warp * 512 + (lane >> 2) * 64;   // this lane's window
    const int kbase = lvl * 96 + (q >> 1) * 8 + (q & 1) * 4;                         // K index of the lane's results, i = 0
    const int row = 8 * (warp >> 1) + (warp & 1) + 2 * pw;                           // rows {m, m+2, m+4, m+6} of an 8-row group

    auto pixel_at = [&](int k) {                                                     // k-th tile of this CTA
      PixelIn px;
      const bool tv = k < my_tiles;
      const int t = blockIdx.x + (tv ? k : 0) * gridDim.x;
      const int b = t / tiles_per_img;
      const int p = (t - b * tiles_per_img) * kTile + row;
      px.inside = tv && p < HW;
      px.n = px.inside ? (long long)b * HW + p : 0;
      px.d = px.inside ? __ldg(disp + px.n) : 0.f;
      px.c = px.inside ? (coords ? __ldg(coords + px.n) : (float)(p % W)) : 0.f;
      return px;
    };
    auto issue = [&](Round& r, const PixelIn& px) {
      int t0;
      split_pos(px.d * sc, t0, r.fg);                                       // geometry.py:43
      const float4* gp = gbase + ((px.n * Dl + t0) * 2 + q);                 // never dereferenced outside the row
#pragma unroll
      for (int i = 0; i < 5; ++i)
        r.g[i] = ldg4_or_zero(gp + 4 * i, px.inside && (unsigned)(t0 + 2 * i + (q >> 1)) < (unsigned)Dl);
      int t0c;
      split_pos(px.c * sc - px.d * sc, t0c, r.fc);                           // geometry.py:52
      r.ca = t0c & 3;
      const int e0 = (t0c - r.ca) + 4 * q;                                   // multiple of 4: a chunk never straddles 0
      r.crem = Wl - e0;
      r.cw = ldg4_or_zero(reinterpret_cast<const float4*>(cbase + px.n * pitch + e0), px.inside && e0 >= 0 && e0 < Wl);
    };
    auto consume = [&](const Round& r, uint32_t a_s) {
      const float f = r.fg, omf = 1.0f - f;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        float4 snd = r.g[i];
        if (q < 2) snd = i < 4 ? r.g[i + 1] : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 nx;
        nx.x = __shfl_xor_sync(0xffffffffu, snd.x, 2);
        nx.y = __shfl_xor_sync(0xffffffffu, snd.y, 2);
        nx.z = __shfl_xor_sync(0xffffffffu, snd.z, 2);
        nx.w = __shfl_xor_sync(0xffffffffu, snd.w, 2);
        if (i < 4 || q < 2)                                                  // tap k = 2i + (q >> 1) <= 8
          put4<kF16>(a_s + a_offset(kbase + 16 * i, row), r.g[i].x * omf + nx.x * f, r.g[i].y * omf + nx.y * f,
                     r.g[i].z * omf + nx.z * f, r.g[i].w * omf + nx.w * f, split);
      }
      // correlation taps: the 16 staged floats start at element (t0 & ~3) of the row; tap j is staged[ca + j]
      float4 c = r.cw;
      if (r.crem < 2) c.y = 0.f;
      if (r.crem < 3) c.z = 0.f;
      if (r.crem < 4) c.w = 0.f;
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(scr + q * 16), "f"(c.x), "f"(c.y), "f"(c.z), "f"(c.w) : "memory");
      __syncwarp();
      const float fc = r.fc, omfc = 1.0f - fc;
      const uint32_t s = scr + (r.ca + 2 * q) * 4;
      float x0, x1, x2;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x0) : "r"(s) : "memory");
      asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(x1) : "r"(s) : "memory");
      asm volatile("ld.shared.f32 %0, [%1+8];" : "=f"(x2) : "r"(s) : "memory");
      put2<kF16>(a_s + a_offset(lvl * 96 + 72 + 2 * q, row), x0 * omfc + x1 * fc, x1 * omfc + x2 * fc, split);
      if (q == 0) {                                                          // tap 8; behind it the constant-1 column the bias rides on
        float x8, x9;
        asm volatile("ld.shared.f32 %0, [%1+32];" : "=f"(x8) : "r"(s) : "memory");
        asm volatile("ld.shared.f32 %0, [%1+36];" : "=f"(x9) : "r"(s) : "memory");
        put2<kF16>(a_s + a_offset(lvl * 96 + 80, row), x8 * omfc + x9 * fc, 1.0f, split);
      }
      __syncwarp();                                                          // scratch is rewritten by the next round
    };

    // one pipeline step: `cur` holds tile k (loads in flight), px is the pixel of tile k+1
    PixelIn px;
    auto step = [&](Round& cur, Round& nxt, int k) {
      TAP_TRACE(warp, k, 0);
      // touch: the first read of a loaded register waits for every load in flight; popc(x) >> 6 is always 0, which
      // ptxas cannot fold, so the next tile's addresses depend on it and its loads stay below this point
      int z;
      asm volatile("{\n\t.reg .b32 t;\n\tpopc.b32 t, %1;\n\tshr.u32 %0, t, 6;\n\t}" : "=r"(z) : "r"(__float_as_uint(cur.g[0].x)));
      TAP_TRACE(warp, k, 1);
      PixelIn pz = px;
      pz.n += z;
      issue(nxt, pz);
      px = pixel_at(k + 2);
      const int st = k & 1;
      TAP_TRACE(warp, k, 2);
      umma::mbar_wait(a_empty + st, ((k >> 1) & 1) ^ 1);                     // the MMAs two tiles back have read the stage
      TAP_TRACE(warp, k, 3);
      consume(cur, a_s + st * kAStage);
      umma::fence_proxy_async();                                             // generic-proxy writes -> tensor-core proxy
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(a_full + st);
      TAP_TRACE(warp, k, 4);
    };
    Round ra, rb;
    px = pixel_at(0);
    issue(ra, px);
    px = pixel_at(1);
    for (int k = 0; k < my_tiles; k += 2) {
      step(ra, rb, k);
      if (k + 1 < my_tiles) step(rb, ra, k + 1);
    }
  } else if (warp == kMmaWarp) {
    // ================= MMA issuer =================
    if (lane == 0) {
      umma::mbar_expect_tx(w_full, (uint32_t)kBPlane * (split ? 2u : 1u));
      for (int kb = 0; kb < kKB; ++kb) {
        umma::tma_load_2d(b_w + 2 * kb * kBBlock, &tmW_hi, w_full, kb * 64, 0);
        if (split) umma::tma_load_2d(b_w + (2 * kb + 1) * kBBlock, &tmW_lo, w_full, kb * 64, 0);
      }
      umma::mbar_wait(w_full, 0);
      const uint32_t idesc_w = umma::idesc_16_f32(128, split ? 2 * kNOut : kNOut, kF16);   // A_hi x [W_hi | W_lo]
      const uint32_t idesc_h = umma::idesc_16_f32(128, kNOut, kF16);                         // A_lo x W_hi
      const uint32_t bw = umma::desc_lo_sw128(umma::smem_u32(b_w));
      for (int t = 0; t < my_tiles; ++t) {
        const int buf = t & 1;                           // A stage and accumulator of this tile
        wait_backoff(a_full + buf, (t >> 1) & 1);
        TAP_TRACE(kMmaWarp, t, 0);
        wait_backoff(acc_empty + buf, ((t >> 1) & 1) ^ 1);
        TAP_TRACE(kMmaWarp, t, 1);
        umma::tc_fence_after();
        const uint32_t ah = umma::desc_lo_sw128(umma::smem_u32(a_base + buf * kAStage)), al = ah + (kAPlane >> 4);
        const uint32_t acc = tmem_d + (uint32_t)(buf * kAccCols);
        uint32_t accumulate = 0;
#pragma unroll 1
        for (int kb = 0; kb < kKB; ++kb) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t ao = (uint32_t)(kb * kABlock + k * 32) >> 4, bo = (uint32_t)(2 * kb * kBBlock + k * 32) >> 4;
            umma::mma_ss_lohi<false, false>(acc, ah + ao, bw + bo, idesc_w, accumulate);
            if (split) umma::mma_ss_lohi<false, false>(acc, al + ao, bw + bo, idesc_h, 1u);
            accumulate = 1u;
          }
        }
        umma::mma_commit(a_empty + buf);
        umma::mma_commit(acc_full + buf);
        TAP_TRACE(kMmaWarp, t, 2);
      }
    }
  } else if (warp < kProdWarps + kEpiWarps) {
    // ================= epilogue: 3 warps <-> TMEM lanes 0..79; bias + ReLU -> the two planes [N][64] convc2 reads =======
    const int qe = warp & 3;
    for (int t = 0; t < my_tiles; ++t) {
      const int buf = t & 1;
      TAP_TRACE(warp, t, 0);
      umma::mbar_wait(acc_full + buf, (t >> 1) & 1);
      TAP_TRACE(warp, t, 1);
      umma::tc_fence_after();
      const int tt = blockIdx.x + t * gridDim.x;
      const int b = tt / tiles_per_img;
      const int p = qe * 32 + lane < kTile ? (tt - b * tiles_per_img) * kTile + qe * 32 + lane : HW;   // rows 80..95: none
      const long long o = ((long long)b * HW + p) * kNOut;
      const uint32_t ta = tmem_d + (uint32_t)(buf * kAccCols) + ((uint32_t)(qe * 32) << 16);
#pragma unroll 1
      for (int hf = 0; hf < 2; ++hf) {
        float v[32];
        umma::tmem_ld_32x32(ta + hf * 32, v);
        if (split) {
          float u[32];
          umma::tmem_ld_32x32(ta + kNOut + hf * 32, u);
          umma::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += u[i];
        } else {
          umma::tmem_ld_wait();
        }
        if (hf == 1) {
          umma::tc_fence_before();
          umma::mbar_arrive(acc_empty + buf);
          TAP_TRACE(warp, t, 2);
        }
        if (p < HW) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);              // the bias is column 81 of the GEMM
#ifdef AS_TAP_NOEPI                          // experiment: no output traffic
          if (__float_as_uint(v[0]) == 0x12345678u && __float_as_uint(v[31]) == 0x9abcdef0u)
#endif
          as_store_split32_v8(v, out_hi, out_lo, o + hf * 32, out_fmt);   // 32-byte stores (common.cuh)
        }
      }
      TAP_TRACE(warp, t, 3);
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) umma::tmem_dealloc(tmem_d, 2 * kAccCols);
}

}  // namespace

#ifdef AS_TAP_TRACE
extern "C" int as_tap_trace_set(long long* buf) {      // [28 warps][16 tiles][8 events] clock64 values of CTA 0
  return (int)cudaMemcpyToSymbol(g_tap_trace, &buf, sizeof(buf));
}
#endif

// Same contract as as_geo_lookup_convc1 (lookup_c1_umma.cu) with the convc1 weights packed in the tap-major K order:
// channel (level l, group g, tap k) at K = l*96 + k*8 + g, correlation tap k of level l at K = l*96 + 72 + k, and the
// BIAS in column K = 81 (the kernel feeds a constant 1 there); the `bias` argument is accepted for signature parity only.
// Needs 2 levels, 16-byte aligned level buffers and correlation pitches that are multiples of 4 floats.
extern "C" int as_geo_lookup_convc1_tap(const float* const* geo_levels, int G, int Dg, const float* const* corr_levels,
                                        const int* corr_widths, const int* corr_pitches, int num_levels, const float* disp,
                                        const float* coords, const void* w_hi, const void* w_lo, const float* bias, int nsplit,
                                        void* out_hi, void* out_lo, int B, int H, int W, int radius, as_stream_t stream) {
  if (!geo_levels || !corr_levels || !corr_widths || !corr_pitches || !disp || !w_hi || !bias || !out_hi) return AS_ERR_BAD_ARG;
  if (B <= 0 || H <= 0 || W <= 0 || Dg <= 1) return AS_ERR_BAD_ARG;
  if (G != 8 || radius != kR || num_levels != 2) return AS_ERR_UNSUPPORTED;
  if (nsplit != 1 && nsplit != 3) return AS_ERR_BAD_ARG;
  if (nsplit == 3 && (!w_lo || !out_lo)) return AS_ERR_BAD_ARG;
  TapLevels lv{};
  for (int l = 0; l < 2; ++l) {
    if (!geo_levels[l] || !corr_levels[l] || corr_pitches[l] < corr_widths[l]) return AS_ERR_BAD_ARG;
    if (!as_aligned16(geo_levels[l]) || !as_aligned16(corr_levels[l]) || (corr_pitches[l] & 3)) return AS_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(out_hi) | reinterpret_cast<uintptr_t>(out_lo)) & 31u) return AS_ERR_UNSUPPORTED;   // 32-byte stores
    lv.geo[l] = geo_levels[l];
    lv.corr[l] = corr_levels[l]; lv.width[l] = corr_widths[l]; lv.pitch[l] = corr_pitches[l];
  }
  CUtensorMap tW_hi, tW_lo;
  const uint64_t dims[2] = {192u, (uint64_t)kNOut};
  const uint64_t str[1] = {192u * 2};
  const uint32_t box[2] = {64u, (uint32_t)kNOut};
  int rc;
  if ((rc = umma::make_tmap_bf16(&tW_hi, w_hi, 2, dims, str, box)) != AS_OK) return rc;
  if (nsplit == 3) {
    if ((rc = umma::make_tmap_bf16(&tW_lo, w_lo, 2, dims, str, box)) != AS_OK) return rc;
  } else {
    tW_lo = tW_hi;
  }
  const int HW = H * W;
  const int tiles_per_img = as_ceil_div(HW, kTile);
  const long long nt = (long long)tiles_per_img * B;
  if (nt >= (1LL << 31)) return AS_ERR_INDEX_RANGE;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = nt < sms ? (int)nt : sms;
  cudaStream_t st = as_cu(stream);
  cudaError_t e;
  const bool f16 = as_operand_f16_internal() != 0;
  const int out_fmt = as_operand_fmt_internal();
#define AS_TAP_LAUNCH(F, S)                                                                                               \
  do {                                                                                                                    \
    e = cudaFuncSetAttribute(geo_lookup_convc1_tap_kernel<F, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes); \
    if (e != cudaSuccess) return (int)e;                                                                                  \
    geo_lookup_convc1_tap_kernel<F, S><<<grid, kThreads, kSmemBytes, st>>>(                                               \
        tW_hi, tW_lo, lv, Dg, disp, coords, bias, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, HW, W, tiles_per_img,   \
        (int)nt, out_fmt);                                                                                                \
  } while (0)
  if (nsplit == 3) { if (f16) AS_TAP_LAUNCH(true, true); else AS_TAP_LAUNCH(false, true); }
  else { if (f16) AS_TAP_LAUNCH(true, false); else AS_TAP_LAUNCH(false, false); }
#undef AS_TAP_LAUNCH
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}
