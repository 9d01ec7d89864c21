// the IGEV lookup fused with its only consumer, BasicMotionEncoder.convc1 (1x1, 162 -> 64) + ReLU
// (models/coreContinuous_IGEV/geometry.py:34-60 feeding update.py:78,85).
//
// The 162-channel lookup tensor (648 B/pixel written, then re-read and re-laid-out for the tensor cores) never exists
// in HBM: per 128-pixel tile the interpolated taps are produced in registers, split into bf16 hi/lo and written
// straight into the K-major, 128B-swizzled shared-memory tiles a tcgen05 MMA reads as its A operand
// (K = 192 = 162 channels + zero padding); the convc1 weights (B operand, 48 KB hi+lo) are fetched once per CTA by
// TMA and stay resident; the 128 x 64 fp32 accumulator lives in TMEM and leaves through bias + ReLU as the bf16
// hi/lo planes [N][64] that convc2 consumes.  HBM traffic per pixel: 728 B read + 256 B written (was 1372 + 648 +
// 768 + 768 + 256 across three kernels).
//
// Thread mapping while producing features: 8 lanes per pixel, lane = geometry group g, so the 8 lanes read one
// 32-byte sector per tap ([pixel][disparity][group] layout) and each lane owns the 9 channels of its group;
// lanes 0..L-1 additionally own the correlation row of level g.
#include "umma.cuh"

namespace {

constexpr int kNG = 3;                     // producer groups == A stages == TMEM accumulators
constexpr int kGroupThreads = 128;         // 4 warps per producer group
constexpr int kThreads = kNG * kGroupThreads + 96;   // + 2 epilogue warps + 1 MMA warp
constexpr int kTile = 64;                  // pixels per tile (rows 64..127 of the M = 128 MMA are don't-care)
constexpr int kG = 8, kR = 4, kK = 9, kTaps = 10;
constexpr int kKPad = 192;                 // 162 channels padded to 3 K-blocks of 64
constexpr int kNOut = 64;                  // convc1 output channels == UMMA N
constexpr int kABlock = kTile * 128;       // bytes of one [64 x 64] bf16 K-block
constexpr int kBBlock = kNOut * 128;
// GEO = true: IGEV (8 geometry groups + correlation, K = 192 = 3 blocks); GEO = false: RAFT CorrBlock1D (correlation rows
// of up to 6 levels only, K = 10*L <= 64 = 1 block; corePrune_RAFT/geometry.py:24-43 feeding update.py:78,85)
__host__ __device__ constexpr int k_blocks(bool geo) { return geo ? 3 : 1; }
__host__ __device__ constexpr int a_stage_bytes(bool geo) { return 2 * k_blocks(geo) * kABlock; }   // hi + lo planes of a tile
// The M = 128 MMA reads 128 rows per K-block, i.e. 8 KB past each 64-row block: into the next block / stage / the
// weight tiles, all inside this allocation.  Those rows land in TMEM lanes 64..127, which nobody reads.
__host__ __device__ constexpr int smem_bytes(bool geo) {
  return 1024 + kNG * a_stage_bytes(geo) + 2 * k_blocks(geo) * kBBlock + 256 + (geo ? 0 : kABlock);   // RAFT: slack for the over-read
}

struct C1Levels {
  const float* geo[4];
  const float* corr[4];
  int width[4];
  int pitch[4];
};

__device__ __forceinline__ void split_pos(float x, int& t0, float& f) {
  const float fl = floorf(x);
  f = x - fl;
  t0 = (int)fminf(fmaxf(fl, -1.0e6f), 1.0e6f) - kR;
}
__device__ __forceinline__ float level_scale(int l) { return __int_as_float((127 - l) << 23); }

// K order of the fused GEMM: channel (level l, group g, tap k) sits at K = l*96 + g*10 + k (g == 8: the correlation taps),
// i.e. every 9-tap run starts on an even K and is padded to 10, so a lane emits aligned bf16 PAIRS (st.shared.b32);
// the pad positions carry zero weights.  2 levels x 96 = 192 = 3 K-blocks.
// Byte offset of K index kidx inside a row-0 tile: K-block (8 KB apart) | byte inside the 128-byte row.  Row r adds
// r*128 and XORs the 16-byte chunk bits [4,7) with r & 7 (the 128B swizzle the MMA descriptor expects).
__device__ __forceinline__ uint32_t k_offset(int kidx) { return ((uint32_t)(kidx >> 6) << 13) | ((uint32_t)(kidx & 63) << 1); }

template <bool kF16, int LO_OFF>
__device__ __forceinline__ void put_pair(uint32_t row_addr, uint32_t r7s, uint32_t koff, float v0, float v1, bool split) {
  const uint32_t addr = row_addr + (koff ^ r7s);
  const uint32_t h = as_cvt16x2(v0, v1, kF16);
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(h) : "memory");
  if (split) {
    const uint32_t l = as_cvt16x2(v0 - as_widen_lo16(h, kF16), v1 - as_widen_hi16(h, kF16), kF16);
    asm volatile("st.shared.b32 [%0+%2], %1;" ::"r"(addr), "r"(l), "n"(LO_OFF) : "memory");
  }
}

// tap j of a row: 0 outside [0, limit), else row[j*stride]; one compare + one predicated load with an immediate offset
template <int BYTE_OFF>
__device__ __forceinline__ float tap_or_zero(const float* row, unsigned idx, unsigned limit) {
  float v;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.lt.u32 p, %2, %3;\n\t"
      "mov.f32 %0, 0f00000000;\n\t"
      "@p ld.global.nc.f32 %0, [%1+%4];\n\t}"
      : "=f"(v)
      : "l"(row), "r"(idx), "r"(limit), "n"(BYTE_OFF));
  return v;
}

template <int L>
struct Taps {                              // the raw taps one lane needs for one pixel
  float wg[L][kTaps], fg[L];
  float wc[kTaps], fc;
};

template <int STRIDE_BYTES, int J = 0>
struct TapLoader {
  static __device__ __forceinline__ void run(float (&w)[kTaps], const float* row, int t0, unsigned limit) {
    w[J] = tap_or_zero<J * STRIDE_BYTES>(row, (unsigned)(t0 + J), limit);
    TapLoader<STRIDE_BYTES, J + 1>::run(w, row, t0, limit);
  }
};
template <int STRIDE_BYTES>
struct TapLoader<STRIDE_BYTES, kTaps> {
  static __device__ __forceinline__ void run(float (&)[kTaps], const float*, int, unsigned) {}
};

// disparity / x-coordinate of one pixel visit, fetched two visits ahead of the conversion
struct PixelIn {
  float d, c;
  long long n;                             // flat pixel index (any valid pixel when !inside)
  bool inside;
};

__device__ __forceinline__ PixelIn load_pixel(const float* __restrict__ disp, const float* __restrict__ coords, long long nbase,
                                              int p, int HW, int W, bool tile_valid) {
  PixelIn px;
  px.inside = tile_valid && p < HW;
  px.n = px.inside ? nbase + p : 0;
  px.d = px.inside ? __ldg(disp + px.n) : 0.f;
  px.c = px.inside ? (coords ? __ldg(coords + px.n) : (float)(p % W)) : 0.f;
  return px;
}

template <int L, bool GEO>
__device__ __forceinline__ void load_taps(Taps<GEO ? L : 1>& t, const C1Levels& lv, int Dg, const PixelIn& px, int g) {
  const float d = px.d, c = px.c;
  const bool inside = px.inside;
  const long long n = px.n;
#pragma unroll
  for (int l = 0; l < (GEO ? L : 0); ++l) {
    const float sc = level_scale(l);
    int t0;
    split_pos(d * sc, t0, t.fg[l]);                                  // geometry.py:43
    const int Dl = Dg >> l;
    const float* row = lv.geo[l] + (n * Dl + t0) * kG + g;            // may point before the row: never dereferenced there
    TapLoader<kG * 4>::run(t.wg[l], row, t0, inside ? (unsigned)Dl : 0u);
  }
  t.fc = 0.f;
#pragma unroll
  for (int j = 0; j < kTaps; ++j) t.wc[j] = 0.f;
  if (g < L) {                                                        // correlation row of level g
    const float sc = level_scale(g);
    int t0;
    split_pos(c * sc - d * sc, t0, t.fc);                             // geometry.py:52
    int Wl = lv.width[0], pitch = lv.pitch[0];
    const float* cbase = lv.corr[0];
#pragma unroll
    for (int l = 1; l < L; ++l)                                      // select, not an indexed parameter read
      if (g == l) { Wl = lv.width[l]; pitch = lv.pitch[l]; cbase = lv.corr[l]; }
    const float* row = cbase + n * pitch + t0;
    TapLoader<4>::run(t.wc, row, t0, inside ? (unsigned)Wl : 0u);
  }
}

template <int L, bool kF16, bool GEO>
__device__ __forceinline__ void emit_features(const Taps<GEO ? L : 1>& t, uint32_t a_s, int row, int g, bool split) {
  const uint32_t row_addr = a_s + row * 128, r7s = (uint32_t)(row & 7) << 4;
  constexpr int kLo = k_blocks(GEO) * kABlock;
#pragma unroll
  for (int l = 0; l < (GEO ? L : 0); ++l) {
    const float f = t.fg[l], omf = 1.0f - f;
    float v[kTaps];
#pragma unroll
    for (int k = 0; k < kK; ++k) v[k] = t.wg[l][k] * omf + t.wg[l][k + 1] * f;
    v[kK] = 0.f;
#pragma unroll
    for (int k = 0; k < kTaps; k += 2) put_pair<kF16, kLo>(row_addr, r7s, k_offset(l * 96 + g * kTaps + k), v[k], v[k + 1], split);
  }
  if (g < L) {
    const float omf = 1.0f - t.fc;
    float v[kTaps];
#pragma unroll
    for (int k = 0; k < kK; ++k) v[k] = t.wc[k] * omf + t.wc[k + 1] * t.fc;
    v[kK] = 0.f;
#pragma unroll
    for (int k = 0; k < kTaps; k += 2) put_pair<kF16, kLo>(row_addr, r7s, k_offset(GEO ? g * 96 + kG * kTaps + k : g * kTaps + k), v[k], v[k + 1], split);
  }
}

// consumers (MMA issuer, epilogue) wait for microseconds: back off so the spin does not steal issue slots
__device__ __forceinline__ void wait_backoff(uint64_t* bar, uint32_t parity) {
  while (!umma::mbar_try_wait(bar, parity)) __nanosleep(100);
}

template <int L, bool kF16, bool GEO>
__global__ void __launch_bounds__(kThreads, 1)
geo_lookup_convc1_kernel(const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                         const C1Levels lv, int Dg, const float* __restrict__ disp, const float* __restrict__ coords,
                         const float* __restrict__ bias, __nv_bfloat16* __restrict__ out_hi,
                         __nv_bfloat16* __restrict__ out_lo, int HW, int W, int tiles_per_img, int num_tiles, int nsplit,
                         int out_fmt, bool wide) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int kKB = k_blocks(GEO), kAStage = a_stage_bytes(GEO);
  uint8_t* a_base = smem;                              // kNG stages x (hi[kKB blocks] | lo[kKB blocks])
  uint8_t* b_hi = smem + kNG * kAStage;
  uint8_t* b_lo = b_hi + kKB * kBBlock;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_lo + kKB * kBBlock + (GEO ? 0 : kABlock));
  uint64_t* w_full = bars;                             // weights landed
  uint64_t* a_full = bars + 1;                         // [kNG] producers -> MMA
  uint64_t* a_empty = a_full + kNG;                    // [kNG] MMA done reading the stage -> producers
  uint64_t* acc_full = a_empty + kNG;                  // [kNG] MMA -> epilogue
  uint64_t* acc_empty = acc_full + kNG;                // [kNG] epilogue drained TMEM -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kNG);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform: role branches stay on the uniform datapath
  const bool split = nsplit == 3;
  constexpr int kMmaWarp = kNG * 4 + 2;

  if (tid == 0) {
    umma::prefetch_tmap(&tmW_hi);
    umma::mbar_init(w_full, 1);
    for (int i = 0; i < kNG; ++i) {
      umma::mbar_init(a_full + i, kGroupThreads);
      umma::mbar_init(a_empty + i, 1);
      umma::mbar_init(acc_full + i, 1);
      umma::mbar_init(acc_empty + i, 64);
    }
    umma::fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    umma::tmem_alloc(tmem_slot, kNG > 2 ? 256 : 128);
    umma::tmem_relinquish();
  }
  // zero the A stages once: the pad channels are never written again and must read as 0
  for (int i = tid; i < (kNG * kAStage) / 16; i += kThreads) reinterpret_cast<uint4*>(a_base)[i] = make_uint4(0, 0, 0, 0);
  umma::fence_proxy_async();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  const int my_tiles = blockIdx.x < num_tiles ? (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp < kNG * 4) {
    // ================= producers: one group of 4 warps per stage =================
    const int grp = warp >> 2, gt = tid - grp * kGroupThreads;
    const int g = gt & 7;                              // geometry group owned by this lane
    const int prow = gt >> 3;                          // 16 pixels per pass, 4 passes per tile
    const uint32_t a_s = umma::smem_u32(a_base + grp * kAStage);
    const int group_tiles = my_tiles > grp ? (my_tiles - 1 - grp) / kNG + 1 : 0;
    // Software pipeline, unrolled over the 4 passes of a tile so the two tap buffers swap roles without register copies:
    // raw taps are fetched one pass ahead (the last pass fetches pass 0 of the group's NEXT tile), disparities a whole
    // tile ahead.
    struct TileAt { long long nbase; int p0; bool valid; };
    auto tile_at = [&](int k) {                        // k-th tile of this group
      TileAt ta;
      ta.valid = k < group_tiles;
      const int t = blockIdx.x + (grp + (ta.valid ? k : 0) * kNG) * gridDim.x;
      const int b = t / tiles_per_img;
      ta.nbase = (long long)b * HW;
      ta.p0 = (t - b * tiles_per_img) * kTile + prow;
      return ta;
    };
    TileAt ta = tile_at(0);
    PixelIn px[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) px[i] = load_pixel(disp, coords, ta.nbase, ta.p0 + i * 16, HW, W, ta.valid);
    Taps<GEO ? L : 1> bufA, bufB;
    load_taps<L, GEO>(bufA, lv, Dg, px[0], g);
    for (int k = 0; k < group_tiles; ++k) {
      const TileAt tn = tile_at(k + 1);
      PixelIn pn[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pn[i] = load_pixel(disp, coords, tn.nbase, tn.p0 + i * 16, HW, W, tn.valid);
      load_taps<L, GEO>(bufB, lv, Dg, px[1], g);
      umma::mbar_wait(a_empty + grp, (k & 1) ^ 1);     // the MMAs of this stage's previous tile have read it
      emit_features<L, kF16, GEO>(bufA, a_s, prow, g, split);
      load_taps<L, GEO>(bufA, lv, Dg, px[2], g);
      emit_features<L, kF16, GEO>(bufB, a_s, 16 + prow, g, split);
      load_taps<L, GEO>(bufB, lv, Dg, px[3], g);
      emit_features<L, kF16, GEO>(bufA, a_s, 32 + prow, g, split);
      load_taps<L, GEO>(bufA, lv, Dg, pn[0], g);
      emit_features<L, kF16, GEO>(bufB, a_s, 48 + prow, g, split);
      umma::fence_proxy_async();                       // generic-proxy smem writes -> visible to the tensor-core proxy
      umma::mbar_arrive(a_full + grp);
#pragma unroll
      for (int i = 0; i < 4; ++i) px[i] = pn[i];
    }
  } else if (warp == kMmaWarp) {
    // ================= MMA issuer =================
    if (lane == 0) {
      umma::mbar_expect_tx(w_full, (uint32_t)(kKB * kBBlock) * (split ? 2u : 1u));     // resident weights
      for (int kb = 0; kb < kKB; ++kb) {
        umma::tma_load_2d(b_hi + kb * kBBlock, &tmW_hi, w_full, kb * 64, 0);
        if (split) umma::tma_load_2d(b_lo + kb * kBBlock, &tmW_lo, w_full, kb * 64, 0);
      }
      umma::mbar_wait(w_full, 0);
      const uint32_t idesc = umma::idesc_16_f32(128, kNOut, kF16);
      const uint32_t bh = umma::smem_u32(b_hi), bl = umma::smem_u32(b_lo);
      int used[kNG];                                   // tiles already issued per group
#pragma unroll
      for (int i = 0; i < kNG; ++i) used[i] = 0;
      for (int issued = 0; issued < my_tiles;) {
        // groups finish out of order (memory latency): take whichever stage is full instead of a fixed round-robin
        int grp = -1, use = 0;
#pragma unroll
        for (int i = 0; i < kNG; ++i) {
          if (grp < 0 && i + used[i] * kNG < my_tiles && umma::mbar_test_wait(a_full + i, used[i] & 1)) {
            grp = i;
            use = used[i];
            used[i] += 1;
          }
        }
        if (grp < 0) { __nanosleep(20); continue; }
        ++issued;
        wait_backoff(acc_empty + grp, (use & 1) ^ 1);
        umma::tc_fence_after();
        const uint32_t ah = umma::smem_u32(a_base + grp * kAStage), al = ah + kKB * kABlock;
        const uint32_t acc = tmem_d + (uint32_t)(grp * kNOut);
        uint32_t accumulate = 0;
#pragma unroll 1
        for (int kb = 0; kb < kKB; ++kb) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t ko = (uint32_t)k * 32u;
            const uint64_t dah = umma::smem_desc_k_sw128(ah + kb * kABlock + ko);
            const uint64_t dbh = umma::smem_desc_k_sw128(bh + kb * kBBlock + ko);
            umma::mma_bf16_ss(acc, dah, dbh, idesc, accumulate);
            accumulate = 1u;
            if (split) {
              umma::mma_bf16_ss(acc, dah, umma::smem_desc_k_sw128(bl + kb * kBBlock + ko), idesc, 1u);
              umma::mma_bf16_ss(acc, umma::smem_desc_k_sw128(al + kb * kABlock + ko), dbh, idesc, 1u);
            }
          }
        }
        umma::mma_commit(a_empty + grp);
        umma::mma_commit(acc_full + grp);
      }
    }
  } else {
    // ================= epilogue: 2 warps <-> TMEM lanes 0..63; bias + ReLU -> bf16 hi/lo planes [N][64] ==========
    const int q = warp & 3;                            // kNG*4 is a multiple of 4: q is 0 or 1
    int used[kNG];
#pragma unroll
    for (int i = 0; i < kNG; ++i) used[i] = 0;
    for (int drained = 0; drained < my_tiles;) {
      int grp = -1, use = 0;
#pragma unroll
      for (int i = 0; i < kNG; ++i) {
        if (grp < 0 && i + used[i] * kNG < my_tiles && umma::mbar_test_wait(acc_full + i, used[i] & 1)) {
          grp = i;
          use = used[i];
          used[i] += 1;
        }
      }
      grp = __shfl_sync(0xffffffffu, grp, 0);          // lanes may observe the barrier flip at different polls
      if (grp < 0) { __nanosleep(20); continue; }
      use = __shfl_sync(0xffffffffu, use, 0);
#pragma unroll
      for (int i = 0; i < kNG; ++i) used[i] = __shfl_sync(0xffffffffu, used[i], 0);
      ++drained;
      umma::mbar_wait(acc_full + grp, use & 1);        // every lane observes the completed phase itself
      const int j = grp + use * kNG;
      const int t = blockIdx.x + j * gridDim.x;
      const int b = t / tiles_per_img;
      const int p = (t - b * tiles_per_img) * kTile + q * 32 + lane;
      umma::tc_fence_after();
      float v[2][32];
      const uint32_t ta = tmem_d + (uint32_t)(grp * kNOut) + ((uint32_t)(q * 32) << 16);
      umma::tmem_ld_32x32(ta, v[0]);
      umma::tmem_ld_32x32(ta + 32, v[1]);
      umma::tmem_ld_wait();
      umma::tc_fence_before();
      umma::mbar_arrive(acc_empty + grp);
      if (p < HW) {
        const long long o = ((long long)b * HW + p) * kNOut;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          if (wide) {                              // 32-byte aligned rows: full-sector stores (common.cuh)
#pragma unroll
            for (int i = 0; i < 32; ++i) v[hf][i] = fmaxf(v[hf][i] + __ldg(bias + hf * 32 + i), 0.f);
            as_store_split32_v8(v[hf], out_hi, out_lo, o + hf * 32, out_fmt);
            continue;
          }
#pragma unroll
          for (int jj = 0; jj < 32; jj += 8) {
            uint32_t h[4];
            float y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = fmaxf(v[hf][jj + i] + __ldg(bias + hf * 32 + jj + i), 0.f);
#pragma unroll
            for (int i = 0; i < 4; ++i) h[i] = as_cvt16x2(y[2 * i], y[2 * i + 1], kF16);
            *reinterpret_cast<uint4*>(out_hi + o + hf * 32 + jj) = make_uint4(h[0], h[1], h[2], h[3]);
            if (out_lo) as_store_lo8(out_lo, o + hf * 32 + jj, y, h, out_fmt);     // 16-bit lo or the e5m2 pair plane
          }
        }
      }
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) umma::tmem_dealloc(tmem_d, kNG > 2 ? 256 : 128);
}

}  // namespace

// shared host path of the two entry points
static int launch_lookup_convc1(bool geo, const float* const* geo_levels, int Dg, const float* const* corr_levels,
                                const int* corr_widths, const int* corr_pitches, int num_levels, const float* disp,
                                const float* coords, const void* w_hi, const void* w_lo, const float* bias, int nsplit,
                                void* out_hi, void* out_lo, int B, int H, int W, as_stream_t stream) {
  if (nsplit != 1 && nsplit != 3) return AS_ERR_BAD_ARG;
  if (nsplit == 3 && (!w_lo || !out_lo)) return AS_ERR_BAD_ARG;
  C1Levels lv{};
  for (int l = 0; l < num_levels; ++l) {
    if ((geo && !geo_levels[l]) || !corr_levels[l] || corr_pitches[l] < corr_widths[l]) return AS_ERR_BAD_ARG;
    lv.geo[l] = geo ? geo_levels[l] : nullptr;
    lv.corr[l] = corr_levels[l]; lv.width[l] = corr_widths[l]; lv.pitch[l] = corr_pitches[l];
  }
  const int kpad = geo ? kKPad : 64;
  CUtensorMap tW_hi, tW_lo;
  const uint64_t dims[2] = {(uint64_t)kpad, (uint64_t)kNOut};
  const uint64_t str[1] = {(uint64_t)kpad * 2};
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
  // the kernel's own GEMM always runs on 16-bit hi(/lo) operands (it is bound by the L1 / shared-memory pipe, not by the
  // tensor pipe); under AS_FMT_F16F8 only its OUTPUT planes switch to the e5m2 pair encoding convc2 consumes
  const bool f16 = as_operand_f16_internal() != 0;
  const int out_fmt = as_operand_fmt_internal();
  const bool wide = !((reinterpret_cast<uintptr_t>(out_hi) | reinterpret_cast<uintptr_t>(out_lo)) & 31u);   // 256-bit stores
#define AS_C1_LAUNCH(LV, F, GEO)                                                                                          \
  do {                                                                                                                    \
    e = cudaFuncSetAttribute(geo_lookup_convc1_kernel<LV, F, GEO>, cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                             smem_bytes(GEO));                                                                            \
    if (e != cudaSuccess) return (int)e;                                                                                  \
    geo_lookup_convc1_kernel<LV, F, GEO><<<grid, kThreads, smem_bytes(GEO), st>>>(                                        \
        tW_hi, tW_lo, lv, Dg, disp, coords, bias, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, HW, W, tiles_per_img,   \
        (int)nt, nsplit, out_fmt, wide);                                                                                  \
  } while (0)
  if (geo) {
    if (num_levels == 2) { if (f16) AS_C1_LAUNCH(2, true, true); else AS_C1_LAUNCH(2, false, true); }
    else { if (f16) AS_C1_LAUNCH(1, true, true); else AS_C1_LAUNCH(1, false, true); }
  } else {
    if (num_levels == 4) { if (f16) AS_C1_LAUNCH(4, true, false); else AS_C1_LAUNCH(4, false, false); }
    else { if (f16) AS_C1_LAUNCH(2, true, false); else AS_C1_LAUNCH(2, false, false); }
  }
#undef AS_C1_LAUNCH
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_geo_lookup_convc1(const float* const* geo_levels, int G, int Dg, const float* const* corr_levels,
                                    const int* corr_widths, const int* corr_pitches, int num_levels, const float* disp,
                                    const float* coords, const void* w_hi, const void* w_lo, const float* bias, int nsplit,
                                    void* out_hi, void* out_lo, int B, int H, int W, int radius, as_stream_t stream) {
  if (!geo_levels || !corr_levels || !corr_widths || !corr_pitches || !disp || !w_hi || !bias || !out_hi) return AS_ERR_BAD_ARG;
  if (B <= 0 || H <= 0 || W <= 0 || Dg <= 0) return AS_ERR_BAD_ARG;
  if (G != kG || radius != kR || num_levels < 1 || num_levels > 2) return AS_ERR_UNSUPPORTED;   // 162 (or 81) channels
  return launch_lookup_convc1(true, geo_levels, Dg, corr_levels, corr_widths, corr_pitches, num_levels, disp, coords, w_hi, w_lo,
                              bias, nsplit, out_hi, out_lo, B, H, W, stream);
}

extern "C" int as_corr_lookup_convc1(const float* const* corr_levels, const int* corr_widths, const int* corr_pitches,
                                     int num_levels, const float* disp, const float* coords, const void* w_hi, const void* w_lo,
                                     const float* bias, int nsplit, void* out_hi, void* out_lo, int B, int H, int W, int radius,
                                     as_stream_t stream) {
  if (!corr_levels || !corr_widths || !corr_pitches || !disp || !w_hi || !bias || !out_hi) return AS_ERR_BAD_ARG;
  if (B <= 0 || H <= 0 || W <= 0) return AS_ERR_BAD_ARG;
  if (radius != kR || (num_levels != 2 && num_levels != 4)) return AS_ERR_UNSUPPORTED;          // 18 or 36 channels
  return launch_lookup_convc1(false, nullptr, 1, corr_levels, corr_widths, corr_pitches, num_levels, disp, coords, w_hi, w_lo,
                              bias, nsplit, out_hi, out_lo, B, H, W, stream);
}
