// a1 + a2 on tensor cores: all-pairs row correlation with the whole average-pool pyramid fused into the
// epilogue (reference: einsum at corePrune_RAFT/geometry.py:52 / coreContinuous_IGEV/geometry.py:70, then
// F.avg_pool2d chains at :18 / :28).
//
//   pre-pass   : NCHW fp32 features -> K-major bf16 [B*H][W][Dp] (hi and, in the fp32-parity mode, lo = x - hi)
//   main kernel: persistent, warp-specialised, one CTA per SM
//       warp 0   TMA producer   (cp.async.bulk.tensor, 128B swizzle, 2-stage mbarrier ring)
//       warp 1   MMA issuer     (tcgen05.mma kind::f16, M=128, N=BN<=192, fp32 accumulators in TMEM,
//                                3 MMAs per K-step in split mode: hi*hi + hi*lo + lo*hi)
//       warp 2   TMEM allocator (512 columns = 2 accumulator buffers, so the epilogue of tile i overlaps
//                                the MMAs of tile i+1 -- the kernel is store-bound)
//       warps4-11 epilogue      two groups of 4 warps, one per TMEM accumulator buffer, so two tiles drain
//                                concurrently (tcgen05.ld -> registers -> pairwise pooling for every level ->
//                                shared-memory transpose -> 128-bit row-contiguous global stores)
// The volume is written exactly once; pooled levels never re-read level 0 from HBM.
#include <cstdlib>
#include "umma.cuh"

namespace {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kStages = 2;
constexpr int kThreads = 384;
constexpr int kMaxBN = 160;
constexpr int kStageBytes = 2 * (kBM * 128) + 2 * (kMaxBN * 128);   // A hi/lo + B hi/lo
constexpr int kStgStride = 68;                                      // floats per epilogue staging row (manual-store path)
// per-warp epilogue staging, 1024-byte aligned.  TMA-store path: one swizzled box per level (level 0: 32 rows x 128 B at +0,
// level 1: 32 x 64 B at +4096, level 2: 32 x 32 B at +6144, level 3: 32 x 16 B at +7168 = 7680 B).  Manual path (L > 4,
// narrow or unaligned levels): [32][kStgStride] floats = 8704 B.
constexpr int kStgWarpBytes = 32 * kStgStride * 4;
constexpr int kStgWarpPitch = (kStgWarpBytes + 1023) / 1024 * 1024;
constexpr int kStgBytes = 8 * kStgWarpPitch;
constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kStgBytes + 256;
static_assert(kSmemBytes <= 227 * 1024, "corr_umma shared memory");
constexpr int kMaxPoolLevels = 6;                                   // 32 columns pool down to 1

struct CorrUmmaParams {
  int BH, W1, W2, D;
  int BN, MT, NT, num_tiles, nsplit, L;
  int tma_store;                        // epilogue leaves through cp.async.bulk.tensor stores (L <= 4, 16-byte pitches)
  float* lvl[AS_MAX_LEVELS];
  int pitch[AS_MAX_LEVELS];
};

// fp32 NCHW -> bf16 hi/lo, K-major [BH][W][Dp].  One CTA = 64 pixels of one image row x 64 channels: 16 coalesced
// loads per thread are issued back to back (memory-level parallelism), transposed through shared memory, and each
// pixel's 64 channels leave as one 128-byte line per plane (8 lanes x 16 B).
__global__ void __launch_bounds__(256) split_transpose_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi,
                                                              __nv_bfloat16* __restrict__ lo, int D, int Dp, int H,
                                                              int W) {
  __shared__ float t[64][65];
  const int by = blockIdx.y, b = by / H, y = by - b * H;
  const int x0 = blockIdx.x * 64, d0 = blockIdx.z * 64;
  const int tid = threadIdx.x;
  const long long HW = (long long)H * W;
  const float* src = in + ((long long)b * D) * HW + (long long)y * W;
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int e = tid + 256 * i;
    const int dd = e >> 6, xx = e & 63;
    const int d = d0 + dd, x = x0 + xx;
    v[i] = (d < D && x < W) ? __ldg(src + (long long)d * HW + x) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int e = tid + 256 * i;
    t[e >> 6][e & 63] = v[i];
  }
  __syncthreads();
  const int c = tid & 7;                 // 8-channel chunk
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int xx = (tid >> 3) + 32 * i;
    const int x = x0 + xx;
    if (x >= W) continue;
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a0 = t[8 * c + 2 * j][xx], a1 = t[8 * c + 2 * j + 1][xx];
      const __nv_bfloat16 h0 = __float2bfloat16_rn(a0), h1 = __float2bfloat16_rn(a1);
      __nv_bfloat162 hh = __halves2bfloat162(h0, h1);
      __nv_bfloat162 ll = __halves2bfloat162(__float2bfloat16_rn(a0 - __bfloat162float(h0)),
                                             __float2bfloat16_rn(a1 - __bfloat162float(h1)));
      h[j] = *reinterpret_cast<uint32_t*>(&hh);
      l[j] = *reinterpret_cast<uint32_t*>(&ll);
    }
    const long long o = ((long long)by * W + x) * Dp + d0 + 8 * c;
    *reinterpret_cast<uint4*>(hi + o) = make_uint4(h[0], h[1], h[2], h[3]);
    if (lo) *reinterpret_cast<uint4*>(lo + o) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

__device__ __forceinline__ void decode_tile(int tile, const CorrUmmaParams& p, int& by, int& mt, int& nt) {
  nt = tile % p.NT;
  const int r = tile / p.NT;
  mt = r % p.MT;
  by = r / p.MT;
}

template <int N>
__device__ __forceinline__ void pool_regs(const float (&in)[2 * N], float (&out)[N]) {
#pragma unroll
  for (int j = 0; j < N; ++j) out[j] = (in[2 * j] + in[2 * j + 1]) * 0.5f;   // == F.avg_pool2d([1,2]) bit for bit
}

template <int N>
__device__ __forceinline__ void stage_regs(float* dst, const float (&v)[N]) {
  if constexpr (N >= 4) {
#pragma unroll
    for (int j = 0; j < N; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < N; ++j) dst[j] = v[j];
  }
}

// write a [32 rows][w cols] block of one level from the warp's staging buffer, row-contiguous
__device__ __forceinline__ void flush_level(const float* stg, int off, int w, float* lvl, int pitch, long long row0,
                                            int rows_valid, int col0, int lane) {
  if (w >= 4) {
    const int lpr = w >> 2;            // lanes per row
    const int rpi = 32 / lpr;          // rows per instruction
    const int c4 = (lane % lpr) * 4;
    const int gcol = col0 + c4;
    for (int i = 0; i < lpr; ++i) {
      const int rr = i * rpi + lane / lpr;
      if (rr < rows_valid && gcol < pitch)
        *reinterpret_cast<float4*>(lvl + (row0 + rr) * pitch + gcol) =
            *reinterpret_cast<const float4*>(stg + rr * kStgStride + off + c4);
    }
  } else {
    for (int e = lane; e < 32 * w; e += 32) {
      const int rr = e / w, cc = e - rr * w;
      if (rr < rows_valid && col0 + cc < pitch) lvl[(row0 + rr) * pitch + col0 + cc] = stg[rr * kStgStride + off + cc];
    }
  }
}

struct StoreMaps {
  CUtensorMap lvl[4];
};

__global__ void __launch_bounds__(kThreads, 1)
corr_umma_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                 const __grid_constant__ StoreMaps tmS, const CorrUmmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stages = smem;
  uint8_t* stg_all = smem + kStages * kStageBytes;                   // 1024-byte aligned (kStageBytes is a multiple of 1024)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes + kStgBytes);
  uint64_t* full = bars;                 // [kStages]
  uint64_t* empty = bars + kStages;      // [kStages]
  uint64_t* tfull = bars + 2 * kStages;  // [2]
  uint64_t* tempty = tfull + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a_bytes = kBM * 128, b_bytes = p.BN * 128;

  if (warp == 0 && lane == 0) {
    umma::prefetch_tmap(&tmA_hi);
    umma::prefetch_tmap(&tmB_hi);
    if (p.nsplit == 3) { umma::prefetch_tmap(&tmA_lo); umma::prefetch_tmap(&tmB_lo); }
    if (p.tma_store) for (int l = 0; l < p.L; ++l) umma::prefetch_tmap(&tmS.lvl[l]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { umma::mbar_init(&full[s], 1); umma::mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { umma::mbar_init(&tfull[a], 1); umma::mbar_init(&tempty[a], 4); }
    umma::fence_barrier_init();
  }
  if (warp == 2) {
    umma::tmem_alloc(tmem_slot, 512);
    umma::tmem_relinquish();
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nkb = (p.D + kBK - 1) / kBK;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t tx = (uint32_t)(a_bytes + b_bytes) * (p.nsplit == 3 ? 2u : 1u);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int by, mt, nt;
        decode_tile(tile, p, by, mt, nt);
        for (int kb = 0; kb < nkb; ++kb) {
          umma::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* st = stages + stage * kStageBytes;
          umma::mbar_expect_tx(&full[stage], tx);
          umma::tma_load_3d(st, &tmA_hi, &full[stage], kb * kBK, mt * kBM, by);
          umma::tma_load_3d(st + 2 * a_bytes, &tmB_hi, &full[stage], kb * kBK, nt * p.BN, by);
          if (p.nsplit == 3) {
            umma::tma_load_3d(st + a_bytes, &tmA_lo, &full[stage], kb * kBK, mt * kBM, by);
            umma::tma_load_3d(st + 2 * a_bytes + kMaxBN * 128, &tmB_lo, &full[stage], kb * kBK, nt * p.BN, by);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t idesc = umma::idesc_bf16_f32(kBM, p.BN);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        umma::mbar_wait(&tempty[acc], acc_phase ^ 1);
        umma::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * 256u;
        for (int kb = 0; kb < nkb; ++kb) {
          umma::mbar_wait(&full[stage], phase);
          umma::tc_fence_after();
          const uint32_t st = umma::smem_u32(stages + stage * kStageBytes);
          const uint32_t a_hi = st, a_lo = st + a_bytes, b_hi = st + 2 * a_bytes, b_lo = b_hi + kMaxBN * 128;
          const int kext = min(kBK, p.D - kb * kBK);
          const int nk = (kext + 15) >> 4;
          for (int k = 0; k < nk; ++k) {
            const uint32_t ko = (uint32_t)k * 32u;   // 16 bf16 = 32 bytes inside the 128-byte swizzle row
            const uint64_t dah = umma::smem_desc_k_sw128(a_hi + ko), dbh = umma::smem_desc_k_sw128(b_hi + ko);
            umma::mma_bf16_ss(tmem_d, dah, dbh, idesc, (kb | k) != 0);
            if (p.nsplit == 3) {
              const uint64_t dal = umma::smem_desc_k_sw128(a_lo + ko), dbl = umma::smem_desc_k_sw128(b_lo + ko);
              umma::mma_bf16_ss(tmem_d, dah, dbl, idesc, 1u);
              umma::mma_bf16_ss(tmem_d, dal, dbh, idesc, 1u);
            }
          }
          umma::mma_commit(&empty[stage]);          // smem slot reusable once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma::mma_commit(&tfull[acc]);              // accumulator complete -> epilogue
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp & 3;                          // TMEM lane quarter == warp % 4
    const int group = (warp - 4) >> 2;               // which accumulator buffer this warp drains
    uint8_t* stg_b = stg_all + (warp - 4) * kStgWarpPitch;
    float* stg = reinterpret_cast<float*>(stg_b);
    float* mine = stg + lane * kStgStride;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      if ((it & 1) != group) continue;
      int by, mt, nt;
      decode_tile(tile, p, by, mt, nt);
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      umma::mbar_wait(&tfull[acc], acc_phase);
      umma::tc_fence_after();
      const int row_in_img = mt * kBM + q * 32;               // first x1 of this warp's 32 rows
      const int rows_valid = min(32, p.W1 - row_in_img);      // may be <= 0
      const long long row0 = (long long)by * p.W1 + row_in_img;
      for (int c = 0; c < p.BN / 32; ++c) {
        float r0[32];
        umma::tmem_ld_32x32(tmem_base + (uint32_t)acc * 256u + (uint32_t)c * 32u + ((uint32_t)(q * 32) << 16), r0);
        umma::tmem_ld_wait();
        const int ncol0 = nt * p.BN + c * 32;
        if (p.tma_store) {
          // TMEM -> registers -> (pool) -> one swizzled box per level in shared memory -> ONE TMA store per level.
          // The store engine streams full 128/64/32/16-byte rows and clips rows >= W1 / columns >= pitch itself; the warp
          // goes straight back to the next tcgen05.ld (the manual path below kept 8 warps per SM busy with
          // LDS -> STG chains and reached 13.5 B/clk/SM of stores: the kernel was bound by its own epilogue).
          if (rows_valid > 0 && ncol0 < p.pitch[0]) {
            float r1[16], r2[8], r3[4];
            if (p.L > 1) pool_regs<16>(r0, r1);
            if (p.L > 2) pool_regs<8>(r1, r2);
            if (p.L > 3) pool_regs<4>(r2, r3);
            if (lane == 0) umma::bulk_wait_group_read<0>();       // the previous chunk's stores have read the staging
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(stg_b + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                  make_float4(r0[4 * j], r0[4 * j + 1], r0[4 * j + 2], r0[4 * j + 3]);
            if (p.L > 1) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<float4*>(stg_b + 4096 + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
                    make_float4(r1[4 * j], r1[4 * j + 1], r1[4 * j + 2], r1[4 * j + 3]);
            }
            if (p.L > 2) {
#pragma unroll
              for (int j = 0; j < 2; ++j)
                *reinterpret_cast<float4*>(stg_b + 6144 + lane * 32 + ((j ^ ((lane >> 2) & 1)) << 4)) =
                    make_float4(r2[4 * j], r2[4 * j + 1], r2[4 * j + 2], r2[4 * j + 3]);
            }
            if (p.L > 3) *reinterpret_cast<float4*>(stg_b + 7168 + lane * 16) = make_float4(r3[0], r3[1], r3[2], r3[3]);
            umma::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              umma::tma_store_3d(&tmS.lvl[0], stg_b, ncol0, row_in_img, by);
              if (p.L > 1) umma::tma_store_3d(&tmS.lvl[1], stg_b + 4096, ncol0 >> 1, row_in_img, by);
              if (p.L > 2) umma::tma_store_3d(&tmS.lvl[2], stg_b + 6144, ncol0 >> 2, row_in_img, by);
              if (p.L > 3) umma::tma_store_3d(&tmS.lvl[3], stg_b + 7168, ncol0 >> 3, row_in_img, by);
              umma::bulk_commit_group();
            }
          }
        } else if (rows_valid > 0 && ncol0 < p.pitch[0]) {
          stage_regs<32>(mine, r0);
          if (p.L > 1) {
            float r1[16]; pool_regs<16>(r0, r1); stage_regs<16>(mine + 32, r1);
            if (p.L > 2) {
              float r2[8]; pool_regs<8>(r1, r2); stage_regs<8>(mine + 48, r2);
              if (p.L > 3) {
                float r3[4]; pool_regs<4>(r2, r3); stage_regs<4>(mine + 56, r3);
                if (p.L > 4) {
                  float r4[2]; pool_regs<2>(r3, r4); stage_regs<2>(mine + 60, r4);
                  if (p.L > 5) { float r5[1]; pool_regs<1>(r4, r5); mine[62] = r5[0]; }
                }
              }
            }
          }
          __syncwarp();
          int off = 0;
          for (int l = 0; l < p.L; ++l) {
            const int w = 32 >> l;
            flush_level(stg, off, w, p.lvl[l], p.pitch[l], row0, rows_valid, ncol0 >> l, lane);
            off += w;
          }
          __syncwarp();
        }
      }
      umma::tc_fence_before();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&tempty[acc]);
    }
  }
  if (p.tma_store && warp >= 4 && lane == 0) umma::bulk_wait_group<0>();   // every store of this thread has completed
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 2) umma::tmem_dealloc(tmem_base, 512);
}

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

size_t as_corr_umma_workspace_bytes(int B, int D, int H, int W1, int W2, int mode) {
  const size_t Dp = (size_t)((D + 63) / 64) * 64;
  const size_t planes = (mode == AS_CORR_BF16X3) ? 2 : 1;
  return planes * (align256((size_t)B * H * W1 * Dp * 2) + align256((size_t)B * H * W2 * Dp * 2)) + 256;
}

int as_corr_umma_launch(const float* f1, const float* f2, int B, int D, int H, int W1, int W2, int num_levels,
                        float* const* levels, const int* pitches, int mode, void* ws, size_t ws_bytes,
                        cudaStream_t st) {
  if (num_levels > kMaxPoolLevels) return AS_ERR_UNSUPPORTED;
  if (!ws || ws_bytes < as_corr_umma_workspace_bytes(B, D, H, W1, W2, mode)) return AS_ERR_BAD_ARG;
  if ((long long)B * H > 65535) return AS_ERR_UNSUPPORTED;
  const int BH = B * H;
  const int Dp = (D + 63) / 64 * 64;
  const bool split = (mode == AS_CORR_BF16X3);
  uint8_t* w = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  const size_t s1 = align256((size_t)BH * W1 * Dp * 2), s2 = align256((size_t)BH * W2 * Dp * 2);
  __nv_bfloat16* hi1 = reinterpret_cast<__nv_bfloat16*>(w);
  __nv_bfloat16* hi2 = reinterpret_cast<__nv_bfloat16*>(w + s1);
  __nv_bfloat16* lo1 = split ? reinterpret_cast<__nv_bfloat16*>(w + s1 + s2) : nullptr;
  __nv_bfloat16* lo2 = split ? reinterpret_cast<__nv_bfloat16*>(w + 2 * s1 + s2) : nullptr;

  split_transpose_kernel<<<dim3(as_ceil_div(W1, 64), BH, Dp / 64), 256, 0, st>>>(f1, hi1, lo1, D, Dp, H, W1);
  AS_RETURN_IF_LAUNCH_FAILED();
  split_transpose_kernel<<<dim3(as_ceil_div(W2, 64), BH, Dp / 64), 256, 0, st>>>(f2, hi2, lo2, D, Dp, H, W2);
  AS_RETURN_IF_LAUNCH_FAILED();

  CorrUmmaParams p{};
  p.BH = BH; p.W1 = W1; p.W2 = W2; p.D = D;
  p.NT = as_ceil_div(W2, kMaxBN);
  p.BN = as_ceil_div(as_ceil_div(W2, p.NT), 32) * 32;
  p.MT = as_ceil_div(W1, kBM);
  p.num_tiles = BH * p.MT * p.NT;
  p.nsplit = split ? 3 : 1;
  p.L = num_levels;
  for (int l = 0; l < num_levels; ++l) { p.lvl[l] = levels[l]; p.pitch[l] = pitches[l]; }

  CUtensorMap tA_hi, tA_lo, tB_hi, tB_lo;
  const uint64_t dA[3] = {(uint64_t)Dp, (uint64_t)W1, (uint64_t)BH};
  const uint64_t sA[2] = {(uint64_t)Dp * 2, (uint64_t)W1 * Dp * 2};
  const uint64_t dB[3] = {(uint64_t)Dp, (uint64_t)W2, (uint64_t)BH};
  const uint64_t sB[2] = {(uint64_t)Dp * 2, (uint64_t)W2 * Dp * 2};
  const uint32_t bA[3] = {(uint32_t)kBK, (uint32_t)kBM, 1u};
  const uint32_t bB[3] = {(uint32_t)kBK, (uint32_t)p.BN, 1u};
  int rc;
  if ((rc = umma::make_tmap_bf16(&tA_hi, hi1, 3, dA, sA, bA)) != AS_OK) return rc;
  if ((rc = umma::make_tmap_bf16(&tB_hi, hi2, 3, dB, sB, bB)) != AS_OK) return rc;
  if (split) {
    if ((rc = umma::make_tmap_bf16(&tA_lo, lo1, 3, dA, sA, bA)) != AS_OK) return rc;
    if ((rc = umma::make_tmap_bf16(&tB_lo, lo2, 3, dB, sB, bB)) != AS_OK) return rc;
  } else {
    tA_lo = tA_hi; tB_lo = tB_hi;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaError_t e = cudaFuncSetAttribute(corr_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (e != cudaSuccess) return (int)e;
  // TMA-store epilogue: one fp32 map per level, dims {pitch_l, W1, B*H}, box {32 >> l, 32, 1}; needs L <= 4 (a 16-byte
  // inner box at level 3) and 16-byte aligned level buffers (pitches are multiples of 4 floats by construction)
  StoreMaps tS;
  static const bool allow_tma_store = !(getenv("AS_CORR_TMA_STORE") && getenv("AS_CORR_TMA_STORE")[0] == '0');   // A/B knob
  p.tma_store = allow_tma_store && num_levels <= 4;
  for (int l = 0; l < num_levels && p.tma_store; ++l)
    if ((pitches[l] & 3) || !as_aligned16(levels[l]) || pitches[l] < (32 >> l)) p.tma_store = 0;
  for (int l = 0; l < 4; ++l) {
    const int ll = (p.tma_store && l < num_levels) ? l : 0;
    if (!p.tma_store) { tS.lvl[l] = tA_hi; continue; }
    const uint64_t dS[3] = {(uint64_t)pitches[ll], (uint64_t)W1, (uint64_t)BH};
    const uint64_t sS[2] = {(uint64_t)pitches[ll] * 4, (uint64_t)W1 * pitches[ll] * 4};
    const uint32_t bS[3] = {(uint32_t)(32 >> ll), 32u, 1u};
    if ((rc = umma::make_tmap_f32_store(&tS.lvl[l], levels[ll], 3, dS, sS, bS)) != AS_OK) return rc;
  }
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  corr_umma_kernel<<<grid, kThreads, kSmemBytes, st>>>(tA_hi, tA_lo, tB_hi, tB_lo, tS, p);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}
