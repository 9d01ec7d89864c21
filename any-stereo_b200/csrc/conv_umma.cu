// a8-a11 on tensor cores: 3x3 / 1x1 convolutions of the update block as implicit GEMMs on tcgen05.
//
// Reference: models/*/update.py -- ConvGRU :26-41, BasicMotionEncoder :73-92, DispHead :16-24.
//
// GEMM view: M = 128 output pixels (a TW x TH spatial patch), N = Cout (<= 256, z|r share one N = 256 pass),
// K = taps x Cin.  Activations are pixel-major bf16 planes [B][H][W][C] (hi, and lo = x - hi in the
// fp32-parity mode).  For every (tap, 64-channel chunk) the TMA producer fetches the patch SHIFTED by the tap
// as one 4-D box {64ch, TW, TH, 1}: out-of-image coordinates are zero-filled by the TMA unit, which IS the
// convolution's zero padding; the box lands in shared memory as a K-major, 128B-swizzled [128 x 64] UMMA
// operand -- no im2col buffer, no halo code.  torch.cat of update.py:35-36,39,90 is a loop over source tensors.
//   warp 0 TMA producer | warp 1 tcgen05.mma issuer (3 MMAs per K-step in split mode) | warp 2 TMEM allocator
//   warps 4-7 epilogue: tcgen05.ld -> gates (sigmoid / r*h / tanh / GRU blend / relu / disparity-head dot)
//   -> per-pixel contiguous stores (fp32 state and/or bf16 hi/lo planes for the next convolution).
// Two 256-column TMEM accumulators let the epilogue of tile i overlap the MMAs of tile i+1.
// Operand reuse: for a 3x3 kernel the producer loads ONE (TH+2)-row patch per (dx, channel chunk); the three dy taps
// are 1024-byte-aligned row offsets into that patch (TW*128 B = one or two swizzle atoms), i.e. three UMMA descriptors
// over the same shared memory.  L2->smem traffic of the activations drops from 9 to 3.75 tile loads per chunk.
#include <cstdio>
#include <cstdlib>
#include "umma.cuh"

namespace {

constexpr int kMaxStages = 4;
// warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 3 idle, 4-11 epilogue.  EIGHT epilogue warps: two per TMEM lane
// quarter.  With one sub-tile per CTA (SUB = 1) the two warps of a quarter split the accumulator columns (chunks of 32,
// alternating); with SUB = 2 each drains one sub-tile.  r1 had four epilogue warps (one per scheduler, every instruction
// of a 300-instruction chunk waiting for the previous one): ~4 us per 32-column chunk and tile, which is LONGER than the
// MMAs of a tile in the N <= 128 layers and, in the 2- and 1-pass engines, at N = 256 too (ncu: tensor pipe 23-49 % on
// convc2 / convd2 / motion / q, L2 33 %, DRAM 17 %: neither memory nor tensor bound).
constexpr int kEpiWarps = 8;
constexpr int kConvThreads = 128 + 32 * kEpiWarps;
constexpr int kUFloats = 128 * 9;                 // DISPHEAD: partial sums of the second column half (SUB = 1)
__host__ __device__ constexpr int conv_threads(int) { return kConvThreads; }
constexpr int kW2Floats = 9 * 256;
constexpr int kSmemBudget = 227 * 1024;

struct ConvMaps {
  CUtensorMap a_hi[3];
  CUtensorMap a_lo[3];
  CUtensorMap b_hi;
  CUtensorMap b_lo;
};

struct ConvUmmaParams {
  int B, H, W, TW, TH, tiles_x, tiles_y, num_tiles;
  int KH, KW, num_src, src_ch[3], cin_total;
  int N, nsplit, epilogue;
  int a_plane, a_stage, b_stage, nstA, nstB;   // bytes of one A patch plane / A stage (hi+lo) / B stage; ring depths
  const float* bias; const float* ctx; int ctx_pitch; const float* h; float* z;
  float* out_f32; __nv_bfloat16* out_hi; __nv_bfloat16* out_lo; int out_pitch, out_coff, cout_valid;
  const float* disp; const float* w2; float* u;
  bool f16;                                   // hi planes / weights are IEEE half instead of bf16
  int fmt;                                    // AS_FMT_* of the planes this launch reads and writes
  bool wide;                                  // output planes: 32-byte aligned rows -> 256-bit stores
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

#ifdef AS_CONV_TRACE
// experiment builds only (make EXTRA=-DAS_CONV_TRACE): where the MMA-issuing thread of CTA 0 spends its time.
// [0] waiting for an activation patch, [1] for a weight stage, [2] for a drained accumulator, [3] whole loop, [4] MMAs issued
__device__ long long g_conv_trace[8];
#define CONV_TRACE_WAIT(slot, stmt)                          \
  do {                                                       \
    const long long t__ = clock64();                         \
    stmt;                                                    \
    if (blockIdx.x == 0) tr[slot] += clock64() - t__;        \
  } while (0)
#else
#define CONV_TRACE_WAIT(slot, stmt) stmt
#endif

// store 32 consecutive channels of one pixel as 16-bit hi planes (+ the lo plane of the format: x - hi in 16 bits, or the
// e5m2 pair encoding of AS_FMT_F16F8, common.cuh)
__device__ __forceinline__ void store_split32(const float* v, __nv_bfloat16* hi, __nv_bfloat16* lo, long long off, int fmt,
                                              bool wide) {
  if (wide) {                                   // rows are 32-byte aligned: full-sector stores
    as_store_split32_v8(v, hi, lo, off, fmt);
    return;
  }
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    uint32_t h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = as_cvt16x2(v[j + 2 * i], v[j + 2 * i + 1], fmt != 0);
    *reinterpret_cast<uint4*>(hi + off + j) = make_uint4(h[0], h[1], h[2], h[3]);
    if (lo) as_store_lo8(lo, off + j, v + j, h, fmt);
  }
}

// TWO = CTA-pair mode (tcgen05 cta_group::2).  The SS-mode MMA is bound by shared-memory READ bandwidth (per K-step it
// reads 128 x 32 B of A and N x 32 B of B; at N = 256 that is 96 B/clk of the 128 B/clk port before the TMA fills are
// counted), so two CTAs of a cluster pair up: each loads its own 128-pixel patch and HALF of the weight rows, the
// leader issues M = 256 MMAs that read A from both SMs and half of B from each, and each CTA drains its own 128
// accumulator lanes.  Per SM the B reads and the B ring footprint halve (4 weight stages instead of 2).
//   full barriers live in the leader (both producers arm them remotely, TMA credits them with cta_group::2),
//   "slot free" / "accumulator ready" are multicast commits to both CTAs, "accumulator drained" arrives remotely.
// NS = tensor-core passes per K-step: 1 = hi*hi; 3 = hi*hi + hi*lo + lo*hi (16-bit hi/lo planes); 2 = hi*hi in kind::f16 plus
// ONE kind::f8f6f4 MMA over the e5m2 pair planes of AS_FMT_F16F8 (both cross terms, common.cuh); 4 = like 2 with only the
// weight-residual cross term hi_a*lo_w (half of the e5m2 row: 6 MMAs per K-block instead of 8) -- for layers whose
// activation rounding the result does not see.  Compile-time so that no
// MMA sits under a run-time branch (see the predicated-MMA lint in tests/test_cpu_boundary.py).
// SUB = 128-pixel sub-tiles per CTA and weight pass.  SUB = 2 ("tall" tile, 16 x 16 pixels): ONE (16+2)-row activation patch
// per (dx, channel chunk) and ONE weight stage feed two accumulators, so the L2 -> shared-memory weight traffic per
// output pixel halves and the patch halo shrinks from 10/8 to 18/16 rows.  The 2-pass and 1-pass engines are bound by that
// traffic (45.6 KB per K-block and CTA against 1024 / 512 tensor-core cycles: the 42 B/clk/SM that the L2 delivers), the
// 3-pass engine is tensor-bound and keeps SUB = 1 (no tile-count quantisation loss).  With N > 128 the two accumulators
// fill TMEM (2 x 256 columns): no double buffering, the epilogue (~5 % of a tile) is exposed.
// GRP = "group stages" (3x3 layers with N <= 128): one ring whose stage holds an activation patch AND the KH weight tiles it
// feeds, one full / one empty barrier per stage, i.e. ONE tcgen05.commit per 3 K-blocks instead of four per 3.  Measured
// (tools/experiments/mma_rate.cu): an SS-mode MMA occupies the tensor core for max(75, N/2) cycles whatever M and the
// operand kind are, and every tcgen05.commit between MMAs adds ~170 cycles during which the pipe drains; with a commit
// per K-block (8 MMAs) the N <= 128 layers ran at 110 cycles per MMA (tools/experiments/conv_trace.sh) although the issuing
// thread waited for data only 12 % of the time.
template <bool TWO, int NS, int SUB, bool GRP>
__global__ void __launch_bounds__(conv_threads(SUB), 1) conv_umma_kernel(const __grid_constant__ ConvMaps maps,
                                                                const ConvUmmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + p.nstA * p.a_stage;
  float* w2s = reinterpret_cast<float*>(GRP ? smem + p.nstA * (p.a_stage + p.KH * p.b_stage) : b_ring + p.nstB * p.b_stage);
  float* ubuf = w2s + kW2Floats;                 // [128][9]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ubuf + kUFloats);
  uint64_t* fullA = bars;
  uint64_t* emptyA = bars + kMaxStages;
  uint64_t* fullB = bars + 2 * kMaxStages;
  uint64_t* emptyB = bars + 3 * kMaxStages;
  uint64_t* tfull = bars + 4 * kMaxStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = TWO ? umma::cluster_ctarank() : 0u;
  const bool leader = crank == 0;
  const int nb_rows = TWO ? p.N / 2 : p.N;       // weight rows held by this CTA
  const int b_bytes = nb_rows * 128;             // one weight plane of a stage
  const int niter = (p.num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;   // identical for both CTAs of a pair

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.num_src; ++s) {
      umma::prefetch_tmap(&maps.a_hi[s]);
      if (NS > 1) umma::prefetch_tmap(&maps.a_lo[s]);
    }
    umma::prefetch_tmap(&maps.b_hi);
    if (NS > 1) umma::prefetch_tmap(&maps.b_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      umma::mbar_init(&fullA[s], TWO ? 2 : 1); umma::mbar_init(&emptyA[s], 1);
      umma::mbar_init(&fullB[s], TWO ? 2 : 1); umma::mbar_init(&emptyB[s], 1);
    }
    for (int a = 0; a < 2; ++a) { umma::mbar_init(&tfull[a], 1); umma::mbar_init(&tempty[a], (TWO ? 2 : 1) * kEpiWarps); }
    umma::fence_barrier_init();
  }
  if (warp == 2) {
    if (TWO) { umma::tmem_alloc_2sm(tmem_slot, 512); umma::tmem_relinquish_2sm(); }
    else { umma::tmem_alloc(tmem_slot, 512); umma::tmem_relinquish(); }
  }
  if (p.epilogue == AS_UEPI_DISPHEAD) {
    for (int i = threadIdx.x; i < kW2Floats; i += conv_threads(SUB)) w2s[i] = p.w2[i];
  }
  umma::tc_fence_before();
  __syncthreads();
  if (TWO) umma::cluster_sync_all();             // both CTAs' barriers exist before anyone signals them
  umma::tc_fence_after();
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, tensor-map prefetch, cluster sync)
  // touched no global memory and overlaps the tail of the previous kernel in the stream; from here on the kernel reads
  // what its predecessors wrote.  launch_dependents lets the NEXT kernel's prologue overlap this kernel's tail.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // The kernel allocates all 512 TMEM columns of its SM (1 CTA/SM), so the allocation starts at column 0, lane 0.  Using
  // the constant keeps the accumulator address of every MMA in a uniform register (a value read back from shared memory
  // is not provably uniform: ptxas wrapped each tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop).
  const uint32_t tmem_base = 0u;
  if (*tmem_slot != 0u) __trap();
  const int chunks = p.cin_total >> 6;           // 64-channel chunks per tap
  const int ngroups = p.KW * chunks;             // A patches per tile; each feeds KH taps
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int dy_bytes = p.TW * 128;               // row offset of one dy step inside a patch (multiple of 1024)
  // accumulators: (buffer, sub-tile) -> TMEM column.  SUB = 1: two buffers of <= 256 columns; SUB = 2: two buffers x two
  // sub-tiles of <= 128 columns, or ONE buffer x two sub-tiles of 256 columns
  const int nbuf = (SUB * p.N <= 256) ? 2 : 1;
  const uint32_t sub_cols = SUB == 1 ? 0u : (nbuf == 2 ? 128u : 256u);

  if (warp == 0) {
    if (lane == 0) {
      int sa = 0, sb = 0; uint32_t pha = 0, phb = 0;
      const uint32_t txA = (uint32_t)p.a_plane * (NS > 1 ? 2u : 1u);
      const uint32_t txB = (uint32_t)b_bytes * (NS > 1 ? 2u : 1u);
      const int ph = p.KH >> 1, pw = p.KW >> 1;
      const int brow0 = (int)crank * nb_rows;    // first weight row of my half
      for (int itn = 0; itn < niter; ++itn) {
        int tile = blockIdx.x + itn * gridDim.x;
        if (tile >= p.num_tiles) tile = 0;       // padding iteration: keeps the pair in lockstep, result discarded
        const int b = tile / tiles_per_img;
        const int r = tile - b * tiles_per_img;
        const int ty = r / p.tiles_x, txi = r - ty * p.tiles_x;
        const int x0 = txi * p.TW, y0 = ty * p.TH;
        for (int kx = 0; kx < p.KW; ++kx) {
          int coff = 0;
          for (int s = 0; s < p.num_src; ++s) {
            for (int c0 = 0; c0 < p.src_ch[s]; c0 += 64) {
              umma::mbar_wait(&emptyA[sa], pha ^ 1);
              uint8_t* sta = a_ring + sa * (GRP ? p.a_stage + p.KH * p.b_stage : p.a_stage);
              if (GRP) {                       // the patch and its KH weight tiles land in one stage, behind one barrier
                const uint32_t tx = txA + (uint32_t)p.KH * txB;
                const uint32_t bar = TWO ? umma::mapa(umma::smem_u32(&fullA[sa]), 0) : 0u;
                if (TWO) umma::mbar_expect_tx_cluster(bar, tx); else umma::mbar_expect_tx(&fullA[sa], tx);
                if (TWO) {
                  umma::tma_load_4d_2sm(sta, &maps.a_hi[s], bar, c0, x0 + kx - pw, y0 - ph, b);
                  if (NS > 1) umma::tma_load_4d_2sm(sta + p.a_plane, &maps.a_lo[s], bar, c0, x0 + kx - pw, y0 - ph, b);
                } else {
                  umma::tma_load_4d(sta, &maps.a_hi[s], &fullA[sa], c0, x0 + kx - pw, y0 - ph, b);
                  if (NS > 1) umma::tma_load_4d(sta + p.a_plane, &maps.a_lo[s], &fullA[sa], c0, x0 + kx - pw, y0 - ph, b);
                }
                for (int ky = 0; ky < p.KH; ++ky) {
                  uint8_t* stb = sta + p.a_stage + ky * p.b_stage;
                  const int kcoord = (ky * p.KW + kx) * p.cin_total + coff + c0;
                  if (TWO) {
                    umma::tma_load_2d_2sm(stb, &maps.b_hi, bar, kcoord, brow0);
                    if (NS > 1) umma::tma_load_2d_2sm(stb + b_bytes, &maps.b_lo, bar, kcoord, brow0);
                  } else {
                    umma::tma_load_2d(stb, &maps.b_hi, &fullA[sa], kcoord, 0);
                    if (NS > 1) umma::tma_load_2d(stb + b_bytes, &maps.b_lo, &fullA[sa], kcoord, 0);
                  }
                }
                if (++sa == p.nstA) { sa = 0; pha ^= 1; }
                continue;
              }
              if (!TWO) {
                umma::mbar_expect_tx(&fullA[sa], txA);
                umma::tma_load_4d(sta, &maps.a_hi[s], &fullA[sa], c0, x0 + kx - pw, y0 - ph, b);
                if (NS > 1) umma::tma_load_4d(sta + p.a_plane, &maps.a_lo[s], &fullA[sa], c0, x0 + kx - pw, y0 - ph, b);
              } else {
                const uint32_t bar = umma::mapa(umma::smem_u32(&fullA[sa]), 0);
                umma::mbar_expect_tx_cluster(bar, txA);
                umma::tma_load_4d_2sm(sta, &maps.a_hi[s], bar, c0, x0 + kx - pw, y0 - ph, b);
                if (NS > 1) umma::tma_load_4d_2sm(sta + p.a_plane, &maps.a_lo[s], bar, c0, x0 + kx - pw, y0 - ph, b);
              }
              if (++sa == p.nstA) { sa = 0; pha ^= 1; }
              for (int ky = 0; ky < p.KH; ++ky) {
                umma::mbar_wait(&emptyB[sb], phb ^ 1);
                uint8_t* stb = b_ring + sb * p.b_stage;
                const int kcoord = (ky * p.KW + kx) * p.cin_total + coff + c0;
                if (!TWO) {
                  umma::mbar_expect_tx(&fullB[sb], txB);
                  umma::tma_load_2d(stb, &maps.b_hi, &fullB[sb], kcoord, 0);
                  if (NS > 1) umma::tma_load_2d(stb + b_bytes, &maps.b_lo, &fullB[sb], kcoord, 0);
                } else {
                  const uint32_t bar = umma::mapa(umma::smem_u32(&fullB[sb]), 0);
                  umma::mbar_expect_tx_cluster(bar, txB);
                  umma::tma_load_2d_2sm(stb, &maps.b_hi, bar, kcoord, brow0);
                  if (NS > 1) umma::tma_load_2d_2sm(stb + b_bytes, &maps.b_lo, bar, kcoord, brow0);
                }
                if (++sb == p.nstB) { sb = 0; phb ^= 1; }
              }
            }
            coff += p.src_ch[s];
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      int sa = 0, sb = 0; uint32_t pha = 0, phb = 0;
      const uint32_t idesc = umma::idesc_16_f32(TWO ? 256 : 128, p.N, p.f16);
      const uint32_t idesc8 = umma::idesc_e5m2_f32(TWO ? 256 : 128, p.N);
#ifdef AS_CONV_TRACE
      long long tr[5] = {0, 0, 0, 0, 0};
      const long long t_loop = clock64();
#endif
      for (int it = 0; it < niter; ++it) {
        const int acc = it % nbuf;
        const uint32_t acc_phase = (uint32_t)(it / nbuf) & 1u;
        CONV_TRACE_WAIT(2, umma::mbar_wait(&tempty[acc], acc_phase ^ 1));
        umma::tc_fence_after();
        const uint32_t tmem_d0 = tmem_base + (uint32_t)acc * 256u;
        uint32_t accumulate = 0;
        for (int g = 0; g < ngroups; ++g) {
          CONV_TRACE_WAIT(0, umma::mbar_wait(&fullA[sa], pha));
          const uint32_t sta = umma::smem_u32(a_ring + sa * (GRP ? p.a_stage + p.KH * p.b_stage : p.a_stage));
          if (GRP) umma::tc_fence_after();
          for (int ky = 0; ky < p.KH; ++ky) {
            if (!GRP) {
              CONV_TRACE_WAIT(1, umma::mbar_wait(&fullB[sb], phb));
              umma::tc_fence_after();
            }
            const uint32_t a_hi = sta + (uint32_t)(ky * dy_bytes), a_lo = a_hi + (uint32_t)p.a_plane;
            const uint32_t b_hi = GRP ? sta + (uint32_t)(p.a_stage + ky * p.b_stage) : umma::smem_u32(b_ring + sb * p.b_stage);
            const uint32_t b_lo = b_hi + (uint32_t)b_bytes;
            // descriptor low words at K offset 0; a K-step of 32 bytes adds 2 (addresses are in 16-byte units)
            const uint32_t dbh0 = umma::desc_lo_sw128(b_hi), dbl0 = umma::desc_lo_sw128(b_lo);
#pragma unroll
            for (int t = 0; t < SUB; ++t) {          // sub-tile t = patch rows 8t .. 8t+7 (+ dy): the same weight stage
              const uint32_t tmem_d = tmem_d0 + (uint32_t)t * sub_cols;
              const uint32_t dah0 = umma::desc_lo_sw128(a_hi + (uint32_t)(t * 8 * dy_bytes));
              const uint32_t dal0 = umma::desc_lo_sw128(a_lo + (uint32_t)(t * 8 * dy_bytes));
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t k2 = 2u * (uint32_t)k;
                umma::mma_ss_lohi<TWO, false>(tmem_d, dah0 + k2, dbh0 + k2, idesc, k == 0 ? accumulate : 1u);
                if (NS == 3) {
                  umma::mma_ss_lohi<TWO, false>(tmem_d, dah0 + k2, dbl0 + k2, idesc, 1u);
                  umma::mma_ss_lohi<TWO, false>(tmem_d, dal0 + k2, dbh0 + k2, idesc, 1u);
                } else if (NS == 2) {  // [a_lo*2^6 | a_hi*2^-8] . [w_hi*2^-6 | w_lo*2^8] in e5m2: 32 of the 128 K-bytes per step
                  umma::mma_ss_lohi<TWO, true>(tmem_d, dal0 + k2, dbl0 + k2, idesc8, 1u);
                } else if (NS == 4) {  // the weight-residual half of that row only: a_hi*2^-8 . w_lo*2^8 (K-bytes 64..127)
                  if (k >= 2) umma::mma_ss_lohi<TWO, true>(tmem_d, dal0 + k2, dbl0 + k2, idesc8, 1u);
                }
              }
            }
            accumulate = 1u;
            if (!GRP) {
              if (TWO) umma::mma_commit_2sm(&emptyB[sb], 3); else umma::mma_commit(&emptyB[sb]);
              if (++sb == p.nstB) { sb = 0; phb ^= 1; }
            }
          }
          if (TWO) umma::mma_commit_2sm(&emptyA[sa], 3); else umma::mma_commit(&emptyA[sa]);
          if (++sa == p.nstA) { sa = 0; pha ^= 1; }
        }
        if (TWO) umma::mma_commit_2sm(&tfull[acc], 3); else umma::mma_commit(&tfull[acc]);
      }
#ifdef AS_CONV_TRACE
      if (blockIdx.x == 0) {
        tr[3] = clock64() - t_loop;
        tr[4] = (long long)niter * ngroups * p.KH * 4 * NS * SUB;
        for (int i = 0; i < 5; ++i) g_conv_trace[i] = tr[i];
      }
#endif
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int q = (warp - 4) & 3;                // TMEM lane quarter
    const int st = SUB > 1 ? (warp - 4) >> 2 : 0;        // sub-tile drained by this warp
    const int chalf = SUB > 1 ? 0 : (warp - 4) >> 2;     // SUB = 1: which half of the 32-column chunks
    constexpr int kCStep = SUB > 1 ? 32 : 64;
    const int m = q * 32 + lane;                 // accumulator row == TMEM lane == pixel within the sub-tile
    const int py = m / p.TW + (SUB > 1 ? st * 8 : 0), px = m % p.TW;
    for (int it = 0; it < niter; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const bool real_tile = tile < p.num_tiles;
      const int b = tile / tiles_per_img;
      const int r = tile - b * tiles_per_img;
      const int ty = r / p.tiles_x, txi = r - ty * p.tiles_x;
      const int x = txi * p.TW + px, y = ty * p.TH + py;
      const bool valid = real_tile && (x < p.W) && (y < p.H);
      const long long n = ((long long)b * p.H + y) * p.W + x;
      const int acc = it % nbuf;
      const uint32_t acc_phase = (uint32_t)(it / nbuf) & 1u;
      umma::mbar_wait(&tfull[acc], acc_phase);
      umma::tc_fence_after();
      const uint32_t trow = tmem_base + (uint32_t)acc * 256u + (uint32_t)st * sub_cols + ((uint32_t)(q * 32) << 16);
      float u[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) u[t] = 0.f;
      for (int c0 = chalf * 32; c0 < p.N; c0 += kCStep) {
        float v[32];
        umma::tmem_ld_32x32(trow + (uint32_t)c0, v);
        // the context row of the GRU epilogues is fetched while the TMEM load is in flight
        float cpre[32];
        const bool gru = p.epilogue == AS_UEPI_GRU_ZR || p.epilogue == AS_UEPI_GRU_Q;
        if (gru && valid) {
          const float* cx = p.ctx + n * p.ctx_pitch + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) as_ldg256f(cx + j, cpre + j);      // 32-byte loads: whole sectors per lane
        }
        umma::tmem_ld_wait();
        if (valid) {
        if (p.epilogue == AS_UEPI_GRU_ZR) {
          const int Hd = p.N >> 1;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = sigmoidf_(v[j] + cpre[j]);
          if (c0 < Hd) {                                     // z = sigmoid(convz + cz)      update.py:37
            float* zp = p.z + n * Hd + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) as_stg256f(zp + j, v + j);
          } else {                                           // r*h feeds convq                update.py:38-39
            const float* hp = p.h + n * Hd + (c0 - Hd);
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float h8[8];
              as_ldg256f(hp + j, h8);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[j + i] *= h8[i];
            }
            store_split32(v, p.out_hi, p.out_lo, n * p.out_pitch + p.out_coff + (c0 - Hd), p.fmt, p.wide);
          }
        } else if (p.epilogue == AS_UEPI_GRU_Q) {            // h' = (1-z) h + z tanh(convq + cq)   update.py:39-40
          const float* zp = p.z + n * p.N + c0;
          const float* hp = p.h + n * p.N + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float z8[8], h8[8];
            as_ld256f(zp + j, z8);              // z was written by the previous launch of this stream: coherent path
            as_ldg256f(hp + j, h8);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[j + i] = (1.0f - z8[i]) * h8[i] + z8[i] * tanhf(v[j + i] + cpre[j + i]);
          }
          float* op = p.out_f32 + n * p.N + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) as_stg256f(op + j, v + j);
          store_split32(v, p.out_hi, p.out_lo, n * p.out_pitch + p.out_coff + c0, p.fmt, p.wide);
        } else if (p.epilogue == AS_UEPI_DISPHEAD) {         // relu(conv1) dotted with conv2's 9 taps  update.py:23-24
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + j));
            const float y0 = fmaxf(v[j] + b4.x, 0.f), y1 = fmaxf(v[j + 1] + b4.y, 0.f);
            const float y2 = fmaxf(v[j + 2] + b4.z, 0.f), y3 = fmaxf(v[j + 3] + b4.w, 0.f);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              const float4 w4 = *reinterpret_cast<const float4*>(w2s + t * 256 + c0 + j);
              u[t] = fmaf(w4.x, y0, fmaf(w4.y, y1, fmaf(w4.z, y2, fmaf(w4.w, y3, u[t]))));
            }
          }
        } else if (p.epilogue == AS_UEPI_LINEAR_F32) {       // plain (1x1) linear map, fp32 NHWC out (LIIF first layer)
          // out_pitch / out_coff (0 = dense [n][N]) place the N columns inside a wider row: chunked data gradients
          float* op = p.out_f32 + n * (p.out_pitch ? p.out_pitch : p.N) + p.out_coff + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + j));
            *reinterpret_cast<float4*>(op + j) = make_float4(v[j] + b4.x, v[j + 1] + b4.y, v[j + 2] + b4.z, v[j + 3] + b4.w);
          }
        } else {                                             // relu(conv + bias) [+ disp in the last channel]
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + j));
            v[j] = fmaxf(v[j] + b4.x, 0.f); v[j + 1] = fmaxf(v[j + 1] + b4.y, 0.f);
            v[j + 2] = fmaxf(v[j + 2] + b4.z, 0.f); v[j + 3] = fmaxf(v[j + 3] + b4.w, 0.f);
          }
          if (p.epilogue == AS_UEPI_MOTION && c0 + 32 == p.N) v[31] = __ldg(p.disp + n);   // cat(out, disp) update.py:92
          store_split32(v, p.out_hi, p.out_lo, n * p.out_pitch + p.out_coff + c0, p.fmt, p.wide);
        }
        }
        __syncwarp();     // reconverge before the next warp-aligned tcgen05.ld
      }
      if (p.epilogue == AS_UEPI_DISPHEAD) {
        if (SUB == 1) {       // the two warps of a lane quarter each hold the dot products over half of the 256 channels
          if (chalf == 1) {
#pragma unroll
            for (int t = 0; t < 9; ++t) ubuf[m * 9 + t] = u[t];
          }
          asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
          if (chalf == 0) {
#pragma unroll
            for (int t = 0; t < 9; ++t) u[t] += ubuf[m * 9 + t];
          }
          asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");   // ubuf may be overwritten by the next tile
        }
        if (valid && chalf == 0) {
#pragma unroll
          for (int t = 0; t < 9; ++t) p.u[n * 9 + t] = u[t];
        }
      }
      umma::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (TWO) umma::mbar_arrive_cluster(umma::mapa(umma::smem_u32(&tempty[acc]), 0));   // the leader issues the MMAs
        else umma::mbar_arrive(&tempty[acc]);
      }
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (TWO) umma::cluster_sync_all();             // nobody exits while the pair may still signal / read its smem
  if (warp == 2) {
    if (TWO) umma::tmem_dealloc_2sm(tmem_base, 512); else umma::tmem_dealloc(tmem_base, 512);
  }
}

template <bool TWO, int NS, int SUB, bool GRP>
int launch_conv(const ConvMaps& maps, const ConvUmmaParams& p, int sms, int smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel<TWO, NS, SUB, GRP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  constexpr int CS = TWO ? 2 : 1;
  int grid = p.num_tiles < sms ? p.num_tiles : sms;
  grid = (grid / CS) * CS;
  if (grid < CS) grid = CS;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(conv_threads(SUB));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  static const bool pdl = !(getenv("AS_CONV_PDL") && getenv("AS_CONV_PDL")[0] == '0');   // A/B knob
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  e = cudaLaunchKernelEx(&cfg, conv_umma_kernel<TWO, NS, SUB, GRP>, maps, p);
  if (e != cudaSuccess) return (int)e;
#ifdef AS_CONV_TRACE
  if (getenv("AS_CONV_TRACE_PRINT")) {
    long long tr[8];
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(tr, g_conv_trace, sizeof(tr));
    fprintf(stderr, "conv N=%3d Cin=%3d k=%d epi=%d tiles=%d passes=%d: loop %lld cycles, %lld MMAs (%.1f cyc/MMA); waits: patch %.1f%% weights %.1f%% accumulator %.1f%%\n",
            p.N, p.cin_total, p.KH, p.epilogue, p.num_tiles, NS, tr[3], tr[4], (double)tr[3] / (double)(tr[4] ? tr[4] : 1),
            100.0 * tr[0] / tr[3], 100.0 * tr[1] / tr[3], 100.0 * tr[2] / tr[3]);
  }
#endif
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

// AS_CONV_2CTA=0|1 selects the CTA-pair (cta_group::2) kernel; default on.
bool two_cta_enabled() {
  const char* v = getenv("AS_CONV_2CTA");
  return !(v && v[0] == '0');
}

}  // namespace

extern "C" int as_conv2d_umma(const as_conv_umma_desc* d, as_stream_t stream) {
  if (!d || !d->w_hi) return AS_ERR_BAD_ARG;
  if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->num_src < 1 || d->num_src > 3) return AS_ERR_BAD_ARG;
  if (!((d->KH == 1 && d->KW == 1) || (d->KH == 3 && d->KW == 3))) return AS_ERR_UNSUPPORTED;
  if (d->Cout < 32 || d->Cout > 256 || (d->Cout & 31)) return AS_ERR_UNSUPPORTED;
  if (d->nsplit < 1 || d->nsplit > 4) return AS_ERR_BAD_ARG;
  if (d->nsplit > 1 && !d->w_lo) return AS_ERR_BAD_ARG;
  // 2 passes = the e5m2 pair planes of AS_FMT_F16F8; 3 passes = 16-bit hi/lo planes: the plane format must match
  if ((d->nsplit == 2 || d->nsplit == 4) != (as_operand_fmt_internal() == AS_FMT_F16F8) && d->nsplit != 1) return AS_ERR_BAD_ARG;
  ConvUmmaParams p{};
  p.B = d->B; p.H = d->H; p.W = d->W; p.KH = d->KH; p.KW = d->KW;
  // 128-pixel patch 16 (x) x 8 (y).  Measured on B200 at 312x96: the 8x16 orientation tiles the image exactly (2.5 %
  // fewer MMAs, smaller halo) but runs 2.5 % SLOWER (shorter TMA rows), so 16x8 stays the default.
  {
    p.TW = 16; p.TH = 8;
    if (d->W <= 8) { p.TW = 8; p.TH = 16; }
    const char* force = getenv("AS_CONV_TILE");       // tuning knob: "16x8" | "8x16"
    if (force && force[0] == '1' && d->W > 8) { p.TW = 16; p.TH = 8; }
    if (force && force[0] == '8') { p.TW = 8; p.TH = 16; }
  }
  // tall tiles (two 16x8 sub-tiles per weight pass, AS_CONV_TALL=1): halves the weight traffic per pixel, but at N = 256
  // the two accumulators fill TMEM (no epilogue overlap) and the tile count quantises worse -- measured slower
  // (gru04 z|r 692 -> 886 us, q 443 -> 451 us in the 2-pass engine), so it is an experiment knob, not the default.
  // Taken only when two weight stages still fit next to the two 72 KB activation stages.
  int sub = 1;
  {
    const char* tall = getenv("AS_CONV_TALL");
    // off by default; '1' = every 3x3 layer, '2' = only the N <= 128 layers (they keep accumulator double buffering)
    const bool want = tall && (tall[0] == '1' || (tall[0] == '2' && d->Cout <= 128));
    if (want && d->KH == 3 && p.TW == 16 && d->H > 8) {
      const int tiles2 = as_ceil_div(d->W, 16) * as_ceil_div(d->H, 16) * d->B;
      const bool two2 = two_cta_enabled() && tiles2 >= 4;
      const int a_stage2 = 2 * (16 + 2) * 16 * 128;
      const int b_stage2 = 2 * (two2 ? d->Cout / 2 : d->Cout) * 128;
      if ((kSmemBudget - (1024 + (kW2Floats + kUFloats) * 4 + 256) - 2 * a_stage2) / b_stage2 >= 2) { sub = 2; p.TH = 16; }
    }
  }
  p.tiles_x = as_ceil_div(d->W, p.TW); p.tiles_y = as_ceil_div(d->H, p.TH);
  p.num_tiles = p.tiles_x * p.tiles_y * d->B;
  p.num_src = d->num_src;
  int cin = 0;
  for (int s = 0; s < d->num_src; ++s) {
    if (!d->src[s].hi || (d->nsplit > 1 && !d->src[s].lo)) return AS_ERR_BAD_ARG;
    if (d->src[s].channels <= 0 || (d->src[s].channels & 63)) return AS_ERR_UNSUPPORTED;
    if (!as_aligned16(d->src[s].hi) || (d->src[s].lo && !as_aligned16(d->src[s].lo))) return AS_ERR_ALIGNMENT;
    p.src_ch[s] = d->src[s].channels;
    cin += d->src[s].channels;
  }
  p.cin_total = cin;
  p.N = d->Cout; p.nsplit = d->nsplit; p.epilogue = d->epilogue;
  p.f16 = as_operand_f16_internal() != 0;
  p.fmt = as_operand_fmt_internal();
  const bool two = two_cta_enabled() && p.num_tiles >= 4 && (p.N % 32) == 0;
  p.a_plane = (p.TH + d->KH - 1) * p.TW * 128;          // (TH+2)-row patch for 3x3, the tile itself for 1x1
  p.a_stage = 2 * p.a_plane;
  p.b_stage = 2 * (two ? p.N / 2 : p.N) * 128;           // CTA-pair mode: each CTA holds half of the weight rows
  const int fixed = 1024 + (kW2Floats + kUFloats) * 4 + 256;
  // Ring depths.  The activation patches stream from HBM (each conv input is read once, 60-370 MB), the weights from L2:
  // the ACTIVATION ring is the one that has to cover DRAM latency.  One patch stage feeds KH K-blocks, i.e. 3 x 12 MMAs
  // (2.8 us) in the 3-pass engine but only 3 x 8 / 3 x 4 (1.8 / 0.9 us, half of that at N = 128) in the 2- and 1-pass
  // engines: with 2 patch stages the MMA issuer starved on `fullA` (ncu: tensor pipe 65 % / 49 % on gru04 z|r / q with
  // L2 at 33 %, DRAM at 17 %).  So: 3 weight stages (2 when a stage is 64 KB), then as many patch stages as fit.
  // AS_CONV_RING=r1 restores the round-1 split (2 patch stages, up to 4 weight stages) for A/B runs.
  {
    const char* ring = getenv("AS_CONV_RING");
    const int avail = kSmemBudget - fixed;
    if (ring && ring[0] == 'r') {
      p.nstA = 2;
      p.nstB = (avail - p.nstA * p.a_stage) / p.b_stage;
      if (p.nstB > kMaxStages) p.nstB = kMaxStages;
      if (p.nstB >= 4 && p.nstA * p.a_stage + 4 * p.b_stage + p.a_stage <= avail) p.nstA = 3;
    } else {
      const int nb_min = p.b_stage <= 32 * 1024 ? 3 : 2;
      p.nstA = (avail - nb_min * p.b_stage) / p.a_stage;
      if (p.nstA > kMaxStages) p.nstA = kMaxStages;
      if (p.nstA < 2) p.nstA = 2;
      p.nstB = (avail - p.nstA * p.a_stage) / p.b_stage;
      if (p.nstB > kMaxStages) p.nstB = kMaxStages;
    }
  }
  if (p.nstB < 2) return AS_ERR_UNSUPPORTED;
  // group stages (see the kernel): 3x3 layers with N <= 128 when at least two {patch + 3 weight tiles} stages fit
  bool group = false;
  {
    static const bool allow = !(getenv("AS_CONV_GROUP") && getenv("AS_CONV_GROUP")[0] == '0');      // A/B knob
    const int gstage = p.a_stage + d->KH * p.b_stage;
    const int ng = (kSmemBudget - fixed) / gstage;
    if (allow && sub == 1 && d->KH == 3 && p.N <= 128 && ng >= 2) {
      group = true;
      p.nstA = ng > kMaxStages ? kMaxStages : ng;
      p.nstB = 0;
    }
  }
  p.bias = d->bias; p.ctx = d->ctx; p.ctx_pitch = d->ctx_pitch; p.h = d->h; p.z = d->z;
  p.out_f32 = d->out_f32; p.out_hi = (__nv_bfloat16*)d->out_hi; p.out_lo = (__nv_bfloat16*)d->out_lo;
  p.out_pitch = d->out_pitch; p.out_coff = d->out_coff; p.cout_valid = d->cout_valid;
  p.disp = d->disp; p.w2 = d->w2; p.u = d->u;
  switch (d->epilogue) {
    case AS_UEPI_RELU_SPLIT:
      if (!d->bias || !d->out_hi) return AS_ERR_BAD_ARG;
      break;
    case AS_UEPI_MOTION:
      if (!d->bias || !d->out_hi || !d->disp) return AS_ERR_BAD_ARG;
      break;
    case AS_UEPI_GRU_ZR:
      if (!d->ctx || !d->h || !d->z || !d->out_hi || (d->Cout & 63)) return AS_ERR_BAD_ARG;
      break;
    case AS_UEPI_GRU_Q:
      if (!d->ctx || !d->h || !d->z || !d->out_f32 || !d->out_hi) return AS_ERR_BAD_ARG;
      break;
    case AS_UEPI_DISPHEAD:
      if (!d->bias || !d->w2 || !d->u || d->Cout != 256) return AS_ERR_BAD_ARG;
      break;
    case AS_UEPI_LINEAR_F32:
      if (!d->out_f32 || (d->out_pitch && d->out_pitch < d->out_coff + d->Cout) || (d->out_pitch & 3) || (d->out_coff & 3))
        return AS_ERR_BAD_ARG;
      break;
    default: return AS_ERR_UNSUPPORTED;
  }
  if ((d->out_pitch & 7) || (d->out_coff & 7)) return AS_ERR_ALIGNMENT;
  static const bool allow_wide = !(getenv("AS_CONV_WIDE") && getenv("AS_CONV_WIDE")[0] == '0');       // A/B knob
  p.wide = allow_wide && !(d->out_pitch & 15) && !(d->out_coff & 15) && !(reinterpret_cast<uintptr_t>(d->out_hi) & 31) &&
           !(reinterpret_cast<uintptr_t>(d->out_lo) & 31);
  if (d->epilogue == AS_UEPI_GRU_ZR || d->epilogue == AS_UEPI_GRU_Q) {      // 256-bit accesses to the fp32 state rows
    const uintptr_t a = reinterpret_cast<uintptr_t>(d->z) | reinterpret_cast<uintptr_t>(d->h) | reinterpret_cast<uintptr_t>(d->ctx) |
                        (d->epilogue == AS_UEPI_GRU_Q ? reinterpret_cast<uintptr_t>(d->out_f32) : 0);
    if ((a & 31) || (d->ctx_pitch & 7) || (d->Cout & 15)) return AS_ERR_ALIGNMENT;
  }

  ConvMaps maps;
  int rc;
  for (int s = 0; s < d->num_src; ++s) {
    const uint64_t C = (uint64_t)d->src[s].channels;
    const uint64_t dims[4] = {C, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    const uint64_t str[3] = {C * 2, (uint64_t)d->W * C * 2, (uint64_t)d->H * d->W * C * 2};
    const uint32_t box[4] = {64u, (uint32_t)p.TW, (uint32_t)(p.TH + d->KH - 1), 1u};
    if ((rc = umma::make_tmap_bf16(&maps.a_hi[s], d->src[s].hi, 4, dims, str, box)) != AS_OK) return rc;
    if (d->nsplit > 1) {       // (the e5m2 pair plane has the same 128 bytes per 64-channel chunk: same map, TMA moves bytes)
      if ((rc = umma::make_tmap_bf16(&maps.a_lo[s], d->src[s].lo, 4, dims, str, box)) != AS_OK) return rc;
    } else {
      maps.a_lo[s] = maps.a_hi[s];
    }
  }
  for (int s = d->num_src; s < 3; ++s) { maps.a_hi[s] = maps.a_hi[0]; maps.a_lo[s] = maps.a_lo[0]; }
  {
    const uint64_t Kt = (uint64_t)d->KH * d->KW * cin;
    const uint64_t dims[2] = {Kt, (uint64_t)p.N};
    const uint64_t str[1] = {Kt * 2};
    const uint32_t box[2] = {64u, (uint32_t)p.N};
    if ((rc = umma::make_tmap_bf16(&maps.b_hi, d->w_hi, 2, dims, str, box)) != AS_OK) return rc;
    if (d->nsplit > 1) {
      if ((rc = umma::make_tmap_bf16(&maps.b_lo, d->w_lo, 2, dims, str, box)) != AS_OK) return rc;
    } else {
      maps.b_lo = maps.b_hi;
    }
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int smem = group ? fixed + p.nstA * (p.a_stage + d->KH * p.b_stage) : fixed + p.nstA * p.a_stage + p.nstB * p.b_stage;
  if (two) {   // the weight tensor map's box is one CTA's half of the rows
    const uint64_t Kt = (uint64_t)d->KH * d->KW * cin;
    const uint64_t dims[2] = {Kt, (uint64_t)p.N};
    const uint64_t str[1] = {Kt * 2};
    const uint32_t box[2] = {64u, (uint32_t)(p.N / 2)};
    if ((rc = umma::make_tmap_bf16(&maps.b_hi, d->w_hi, 2, dims, str, box)) != AS_OK) return rc;
    if (d->nsplit > 1) {
      if ((rc = umma::make_tmap_bf16(&maps.b_lo, d->w_lo, 2, dims, str, box)) != AS_OK) return rc;
    } else {
      maps.b_lo = maps.b_hi;
    }
#define AS_CONV_DISPATCH(TWO_)                                                                            \
  do {                                                                                                    \
    if (sub == 2) {                                                                                       \
      if (d->nsplit == 4) return launch_conv<TWO_, 2, 2, false>(maps, p, sms, smem, as_cu(stream));       \
      if (d->nsplit == 3) return launch_conv<TWO_, 3, 2, false>(maps, p, sms, smem, as_cu(stream));       \
      if (d->nsplit == 2) return launch_conv<TWO_, 2, 2, false>(maps, p, sms, smem, as_cu(stream));       \
      return launch_conv<TWO_, 1, 2, false>(maps, p, sms, smem, as_cu(stream));                           \
    }                                                                                                     \
    if (group) {                                                                                          \
      if (d->nsplit == 4) return launch_conv<TWO_, 2, 1, true>(maps, p, sms, smem, as_cu(stream));        \
      if (d->nsplit == 3) return launch_conv<TWO_, 3, 1, true>(maps, p, sms, smem, as_cu(stream));        \
      if (d->nsplit == 2) return launch_conv<TWO_, 2, 1, true>(maps, p, sms, smem, as_cu(stream));        \
      return launch_conv<TWO_, 1, 1, true>(maps, p, sms, smem, as_cu(stream));                            \
    }                                                                                                     \
    if (d->nsplit == 4) return launch_conv<TWO_, 4, 1, false>(maps, p, sms, smem, as_cu(stream));         \
    if (d->nsplit == 3) return launch_conv<TWO_, 3, 1, false>(maps, p, sms, smem, as_cu(stream));         \
    if (d->nsplit == 2) return launch_conv<TWO_, 2, 1, false>(maps, p, sms, smem, as_cu(stream));         \
    return launch_conv<TWO_, 1, 1, false>(maps, p, sms, smem, as_cu(stream));                             \
  } while (0)
    AS_CONV_DISPATCH(true);
  }
  AS_CONV_DISPATCH(false);
#undef AS_CONV_DISPATCH
}
