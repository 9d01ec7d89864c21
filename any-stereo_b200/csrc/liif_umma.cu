// the arbitrary-scale (LIIF) disparity upsampler that follows the iterative loop
// (models/coreContinuous_IGEV/liif.py:575-678 liif_out_multi_scale_Training, :417-449 AffinityFeature,
//  :108-137 liif_feat_multiscale_train; submodule.py:357-372 context_upsample_multiscale_train;
//  continuous_IGEVstereo.py:192-237 upsample_disp).
//
// B200-first restructuring.  The reference gathers a 228-d latent per QUERY (3-6.6 M queries per pair) and runs the
// whole MLP on it.  The gather is nearest-neighbour, so the first Linear commutes with it:
//     W1 . cat_i[feat_i(pix_i(q)), rel_i(q)] + b1  =  sum_i P_i[pix_i(q)] + Wc . rel(q) + b1,   P_i = W1_i . feat_i
// P_i is a 1x1 convolution at the SOURCE resolution (as_conv2d_umma, AS_UEPI_LINEAR_F32) - 25-100x fewer rows than
// queries - and the per-query kernel below starts from 128 channels: add the gathered P rows and the 2-D relative
// coordinates, ReLU, then layers 2..4 (128->64->64->9) as three chained tcgen05 MMAs whose activations never leave
// the SM (TMEM -> registers -> swizzled shared-memory operand of the next MMA), softmax and the 3x3 context
// upsample in the last epilogue.  fp32 parity via split-bf16 (3 MMAs per K-step), or single bf16.
#include "umma.cuh"

namespace {

// ------------------------------------------------------------------------------------------------------------------
// AffinityFeature (liif.py:434-449): cosine similarity with the 8 neighbours of a 3x3 window, clipped at 0
// ------------------------------------------------------------------------------------------------------------------
constexpr int kIsuT = 16;                      // interior tile edge
constexpr int kIsuH = kIsuT + 2;               // with halo
constexpr int kIsuThreads = 352;               // >= 18*18
constexpr int kIsuCh = 8;                      // channel planes staged per barrier pair

__global__ void __launch_bounds__(kIsuThreads)
isu_affinity_kernel(const float* __restrict__ feat, float* __restrict__ aff, __nv_bfloat16* __restrict__ hi,
                    __nv_bfloat16* __restrict__ lo, int C, int H, int W, int c_pad, int c_off) {
  __shared__ float tile[kIsuCh][kIsuH][kIsuH + 1];
  __shared__ float nrm[kIsuH][kIsuH + 1];
  const int b = blockIdx.z, y0 = blockIdx.y * kIsuT, x0 = blockIdx.x * kIsuT;
  const int t = threadIdx.x;
  const int hy = t / kIsuH, hx = t - hy * kIsuH;             // halo coordinates of this thread
  const bool in_halo = t < kIsuH * kIsuH;
  const int y = y0 - 1 + hy, x = x0 - 1 + hx;
  const bool in_img = in_halo && y >= 0 && y < H && x >= 0 && x < W;
  const bool interior = in_halo && hy >= 1 && hy <= kIsuT && hx >= 1 && hx <= kIsuT && y < H && x < W;
  const long long HW = (long long)H * W;
  const float* src = feat + (long long)b * C * HW + (long long)y * W + x;
  float n2 = 0.f, dot[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) dot[k] = 0.f;
  for (int c0 = 0; c0 < C; c0 += kIsuCh) {                   // kIsuCh channel planes per barrier pair
    float v[kIsuCh];
#pragma unroll
    for (int j = 0; j < kIsuCh; ++j) {
      v[j] = (in_img && c0 + j < C) ? __ldg(src + (long long)(c0 + j) * HW) : 0.f;
      n2 = fmaf(v[j], v[j], n2);
    }
    if (in_halo) {
#pragma unroll
      for (int j = 0; j < kIsuCh; ++j) tile[j][hy][hx] = v[j];
    }
    __syncthreads();
    if (interior) {
#pragma unroll
      for (int j = 0; j < kIsuCh; ++j) {
        int k = 0;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            if (dy == 0 && dx == 0) continue;
            dot[k] = fmaf(v[j], tile[j][hy + dy][hx + dx], dot[k]);
            ++k;
          }
      }
    }
    __syncthreads();
  }
  if (in_halo) nrm[hy][hx] = fmaxf(sqrtf(n2), 1e-12f);          // F.normalize eps
  __syncthreads();
  if (!interior) return;
  const float nc = nrm[hy][hx];
  float a[8];
  int k = 0;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      if (dy == 0 && dx == 0) continue;
      a[k] = fmaxf(dot[k] / (nc * nrm[hy + dy][hx + dx]), 0.f);    // affinity[affinity < 0] = 0  (:447)
      ++k;
    }
  const long long pix = ((long long)b * H + y) * W + x;
  if (aff) {
#pragma unroll
    for (int j = 0; j < 8; ++j) aff[((long long)b * 8 + j) * H * W + (long long)y * W + x] = a[j];
  }
  if (hi) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(a[2 * j]), h1 = __float2bfloat16_rn(a[2 * j + 1]);
      const __nv_bfloat162 hp = __halves2bfloat162(h0, h1);
      const __nv_bfloat162 lp = __halves2bfloat162(__float2bfloat16_rn(a[2 * j] - __bfloat162float(h0)),
                                                   __float2bfloat16_rn(a[2 * j + 1] - __bfloat162float(h1)));
      h[j] = *reinterpret_cast<const uint32_t*>(&hp);
      l[j] = *reinterpret_cast<const uint32_t*>(&lp);
    }
    *reinterpret_cast<uint4*>(hi + pix * c_pad + c_off) = make_uint4(h[0], h[1], h[2], h[3]);
    if (lo) *reinterpret_cast<uint4*>(lo + pix * c_pad + c_off) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// nearest-neighbour query arithmetic, bit-compatible with the reference's fp32 evaluation order
// ------------------------------------------------------------------------------------------------------------------
// F.grid_sample(mode='nearest', align_corners=False) after clamp_(-1+1e-6, 1-1e-6)  (liif.py:118-123)
__device__ __forceinline__ int nearest_index(float c, int n) {
  c = fminf(fmaxf(c, -0.999999f), 0.999999f);
  const float x = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(c, 1.0f), (float)n), -1.0f), 0.5f);   // no FMA contraction
  const int i = (int)rintf(x);                                                                 // round half to even
  return min(max(i, 0), n - 1);
}
// (coord - make_coord(n)[i]) * n with make_coord = fl(fl(-1 + r) + fl(2r) * i), r = 1/n  (liif.py:32-45, :129-131)
__device__ __forceinline__ float rel_coord(float c, int i, int n, float c0, float c1) {
  const float q = __fadd_rn(c0, __fmul_rn(c1, (float)i));
  return __fmul_rn(__fadd_rn(c, -q), (float)n);
}

#ifndef AS_LIIF_QT
#define AS_LIIF_QT 128
#endif
constexpr int kQT = AS_LIIF_QT;                // queries per tile: 128 (= UMMA M) or 64 (rows 64..127 are don't-care)
constexpr int kGroups = 256 / kQT;             // independent tile pipelines per CTA sharing the resident weights
constexpr int kGroupThreads = 2 * kQT;         // 16 queries per warp in layer 1
constexpr int kQThreads = kGroups * kGroupThreads;
constexpr int kBlk = kQT * 128;                // one [kQT rows x 64 K] bf16 block (the M = 128 MMA over-reads into the
                                               // next block / the weights when kQT = 64: lanes 64..127 of TMEM are never read)
constexpr int kH1 = 128, kH2 = 64, kH3 = 64, kOutPad = 16, kOut = 9;
// smem: per group activations (hi: 2 blocks, lo: 2 blocks) | W2 hi,lo (2 blocks of 64 rows) | W3 hi,lo | W4 hi,lo | Wc | bars
constexpr int kActBytes = 4 * kBlk;
constexpr int kW2Bytes = 2 * 2 * kH2 * 128, kW3Bytes = 2 * kH3 * 128, kW4Bytes = 2 * kOutPad * 128;
constexpr int kWcBytes = 7 * kH1 * 4;
constexpr int kQSmem = 1024 + kGroups * kActBytes + kW2Bytes + kW3Bytes + kW4Bytes + kWcBytes + 64;

struct QueryArgs {
  const float* P[3];                           // fp32 [B][h_i][w_i][128]
  int h[3], w[3];
  float cy0[3], cy1[3], cx0[3], cx1[3];         // make_coord constants per map and axis
  int n_in;
  const float* coords;                         // [B][Q][2] (y, x)
  const float* wc;                             // [1 + 2*n_in][128]: b1, then the rel-coordinate columns of W1
  const float *b2, *b3, *b4;                   // 64, 64, 16 (padded)
  const float* disp;                           // [B][hd][wd]
  const float* disp_scale;                     // [B] or null (1)
  int hd, wd;
  float* logits;                               // [B][9][Q] or null
  float* out;                                  // [B][Q] or null
  int B, Q, tiles_per_b, num_tiles, nsplit;
};

__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// K-major, 128B-swizzled operand tile: block kb (64 K values) is 16 KB, row r at r*128, 16-byte chunks XORed with r & 7
// four consecutive K values (k multiple of 4) of one row as one 8-byte store per plane
__device__ __forceinline__ void put_quad(uint32_t a_hi, uint32_t lo_off, int row, int k, float v0, float v1, float v2, float v3,
                                         bool split) {
  const uint32_t addr = a_hi + ((uint32_t)(k >> 6) * kBlk) + row * 128 + ((((uint32_t)(k & 63) << 1)) ^ ((uint32_t)(row & 7) << 4));
  const uint32_t h0 = cvt_bf16x2(v0, v1), h1 = cvt_bf16x2(v2, v3);
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(h0), "r"(h1) : "memory");
  if (split) {
    const uint32_t l0 = cvt_bf16x2(v0 - __uint_as_float(h0 << 16), v1 - __uint_as_float(h0 & 0xFFFF0000u));
    const uint32_t l1 = cvt_bf16x2(v2 - __uint_as_float(h1 << 16), v3 - __uint_as_float(h1 & 0xFFFF0000u));
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr + lo_off), "r"(l0), "r"(l1) : "memory");
  }
}

// eight consecutive K values (k multiple of 8) of one row: a whole 16-byte swizzle chunk per plane
__device__ __forceinline__ void put_oct(uint32_t a_hi, uint32_t lo_off, int row, int k, const float (&v)[8], bool split) {
  const uint32_t addr = a_hi + ((uint32_t)(k >> 6) * kBlk) + row * 128 + ((((uint32_t)(k & 63) << 1)) ^ ((uint32_t)(row & 7) << 4));
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = cvt_bf16x2(v[2 * i], v[2 * i + 1]);
    l[i] = cvt_bf16x2(v[2 * i] - __uint_as_float(h[i] << 16), v[2 * i + 1] - __uint_as_float(h[i] & 0xFFFF0000u));
  }
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
  if (split)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + lo_off), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
}

__device__ __forceinline__ void issue_layer(uint32_t acc, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, int kblocks,
                                            int b_blk_bytes, uint32_t idesc, bool split) {
  uint32_t accumulate = 0;
  for (int kb = 0; kb < kblocks; ++kb) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t ko = (uint32_t)k * 32u;
      const uint64_t dah = umma::smem_desc_k_sw128(a_hi + kb * kBlk + ko);
      const uint64_t dbh = umma::smem_desc_k_sw128(b_hi + kb * b_blk_bytes + ko);
      umma::mma_bf16_ss(acc, dah, dbh, idesc, accumulate);
      accumulate = 1u;
      if (split) {
        umma::mma_bf16_ss(acc, dah, umma::smem_desc_k_sw128(b_lo + kb * b_blk_bytes + ko), idesc, 1u);
        umma::mma_bf16_ss(acc, umma::smem_desc_k_sw128(a_lo + kb * kBlk + ko), dbh, idesc, 1u);
      }
    }
  }
}

__device__ __forceinline__ void group_sync(int group) {      // named barrier over the 256 threads of one pipeline
  asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(kGroupThreads) : "memory");
}

__global__ void __launch_bounds__(kQThreads, 1)
liif_query_kernel(const __grid_constant__ CUtensorMap tW2h, const __grid_constant__ CUtensorMap tW2l,
                  const __grid_constant__ CUtensorMap tW3h, const __grid_constant__ CUtensorMap tW3l,
                  const __grid_constant__ CUtensorMap tW4h, const __grid_constant__ CUtensorMap tW4l, const QueryArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* act0 = smem;                                  // per group: hi blocks 0,1 | lo blocks 0,1
  uint8_t* w2h = act0 + kGroups * kActBytes;
  uint8_t* w2l = w2h + kW2Bytes / 2;
  uint8_t* w3h = w2h + kW2Bytes;
  uint8_t* w3l = w3h + kW3Bytes / 2;
  uint8_t* w4h = w3h + kW3Bytes;
  uint8_t* w4l = w4h + kW4Bytes / 2;
  float* wc_s = reinterpret_cast<float*>(w4h + kW4Bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(wc_s) + kWcBytes);
  uint64_t* w_full = bars;
  uint64_t* mma_done = bars + 1;                         // [kGroups]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + kGroups);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  constexpr int kGroupWarps = kGroupThreads / 32;
  const int group = warp / kGroupWarps, gtid = tid - group * kGroupThreads, gwarp = warp - group * kGroupWarps;
  const bool split = a.nsplit == 3;
  const uint32_t lo_off = 2 * kBlk;

  if (tid == 0) {
    umma::prefetch_tmap(&tW2h);
    umma::mbar_init(w_full, 1);
    for (int g = 0; g < kGroups; ++g) umma::mbar_init(mma_done + g, 1);
    umma::fence_barrier_init();
  }
  if (warp == 1) {
    umma::tmem_alloc(tmem_slot, 128 * kGroups);
    umma::tmem_relinquish();
  }
  // b1 and the relative-coordinate columns of W1
  for (int i = tid; i < (1 + 2 * a.n_in) * kH1; i += kQThreads) {
    wc_s[i] = __ldg(a.wc + i);
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_d = *tmem_slot + (uint32_t)(group * 128);
  if (tid == 0) {                                        // resident weights of layers 2..4
    const uint32_t bytes = (uint32_t)(kW2Bytes + kW3Bytes + kW4Bytes) / (split ? 1u : 2u);
    umma::mbar_expect_tx(w_full, bytes);
    for (int kb = 0; kb < 2; ++kb) {
      umma::tma_load_2d(w2h + kb * kH2 * 128, &tW2h, w_full, kb * 64, 0);
      if (split) umma::tma_load_2d(w2l + kb * kH2 * 128, &tW2l, w_full, kb * 64, 0);
    }
    umma::tma_load_2d(w3h, &tW3h, w_full, 0, 0);
    if (split) umma::tma_load_2d(w3l, &tW3l, w_full, 0, 0);
    umma::tma_load_2d(w4h, &tW4h, w_full, 0, 0);
    if (split) umma::tma_load_2d(w4l, &tW4l, w_full, 0, 0);
  }
  const uint32_t act_s = umma::smem_u32(act0 + group * kActBytes);
  uint64_t* done = mma_done + group;
  uint32_t phase = 0;
  bool weights_ready = false;

  // relu(acc + bias) of a hidden layer -> K = 64 operand tile of the next one.  A warp reads the TMEM lane quarter
  // (warp % 4); with 128-query tiles warps w and w+4 split the 64 columns, with 64-query tiles warps 0,1 take all 64.
  auto hidden_epilogue = [&](uint32_t col0, const float* bias) {
    const int q = gwarp & 3;
    const int half0 = kQT == 128 ? (gwarp >> 2) : 0, nhalf = kQT == 128 ? 1 : 2;
    if (q * 32 < kQT) {
      const int row = q * 32 + lane;
      for (int hh = 0; hh < nhalf; ++hh) {
        const int half = half0 + hh;
        float v[32];
        umma::tmem_ld_32x32(tmem_d + col0 + (uint32_t)(half * 32) + ((uint32_t)(q * 32) << 16), v);
        umma::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 8) {                // one 16-byte chunk (8 channels) per store: conflict-free across rows
          const int c = half * 32 + j;
          const float4 ba = __ldg(reinterpret_cast<const float4*>(bias + c)), bb = __ldg(reinterpret_cast<const float4*>(bias + c + 4));
          float y[8];
          y[0] = fmaxf(v[j] + ba.x, 0.f); y[1] = fmaxf(v[j + 1] + ba.y, 0.f);
          y[2] = fmaxf(v[j + 2] + ba.z, 0.f); y[3] = fmaxf(v[j + 3] + ba.w, 0.f);
          y[4] = fmaxf(v[j + 4] + bb.x, 0.f); y[5] = fmaxf(v[j + 5] + bb.y, 0.f);
          y[6] = fmaxf(v[j + 6] + bb.z, 0.f); y[7] = fmaxf(v[j + 7] + bb.w, 0.f);
          put_oct(act_s, lo_off, row, c, y, split);
        }
      }
    }
    umma::tc_fence_before();
    umma::fence_proxy_async();
    group_sync(group);
  };
  auto run_layer = [&](uint32_t acc_col, uint32_t b_hi, uint32_t b_lo, int kblocks, int b_blk_bytes, int n) {
    if (gtid == 0) {
      if (!weights_ready) umma::mbar_wait(w_full, 0);
      umma::tc_fence_after();
      issue_layer(tmem_d + acc_col, act_s, act_s + lo_off, b_hi, b_lo, kblocks, b_blk_bytes, umma::idesc_bf16_f32(128, n), split);
      umma::mma_commit(done);
    }
    weights_ready = true;
    umma::mbar_wait(done, phase);
    phase ^= 1;
    umma::tc_fence_after();
  };

  // this lane's 4 channels of b1 and of the relative-coordinate columns of W1 (layer-1 mapping: lane = channel quad)
  const float4 bias4 = *reinterpret_cast<const float4*>(wc_s + lane * 4);
  float4 wy4[3], wx4[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    wy4[i] = wx4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < a.n_in) {
      wy4[i] = *reinterpret_cast<const float4*>(wc_s + (1 + 2 * i) * kH1 + lane * 4);
      wx4[i] = *reinterpret_cast<const float4*>(wc_s + (2 + 2 * i) * kH1 + lane * 4);
    }
  }

  // lanes 0..15 of a warp own one query each of the warp's 16 (lanes 16..31 mirror them)
  auto load_coords = [&](int tile) {
    float2 c = make_float2(0.f, 0.f);
    if (tile < a.num_tiles) {
      const int tb = tile / a.tiles_per_b;
      const int q = (tile - tb * a.tiles_per_b) * kQT + gwarp * 16 + (lane & 15);
      if (q < a.Q) c = __ldg(reinterpret_cast<const float2*>(a.coords + ((long long)tb * a.Q + q) * 2));
    }
    return c;
  };
  float2 next_c = load_coords(blockIdx.x * kGroups + group);

  for (int t = blockIdx.x * kGroups + group; t < a.num_tiles; t += gridDim.x * kGroups) {
    const int b = t / a.tiles_per_b;
    const int q0 = (t - b * a.tiles_per_b) * kQT;
    // ---- layer 1 (gathered): z1 = relu(b1 + sum_i P_i[pix_i] + Wc . rel)  -> operand tile, K = 128.
    // One warp per query, lane = 4 consecutive channels: the 512-byte P rows are read fully coalesced, the lane's
    // slice of b1 / Wc lives in registers, and the row of the operand tile is written as 8-byte pieces.
    {
      // lanes 0..15 each resolve one of the warp's 16 queries (nearest pixels, relative coordinates)
      const float cy = next_c.x, cx = next_c.y;          // fetched while the previous tile was in its MMA phases
      next_c = load_coords(t + gridDim.x * kGroups);
      long long off[3];
      float ry[3], rx[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        off[i] = 0; ry[i] = 0.f; rx[i] = 0.f;
        if (i < a.n_in) {
          const int iy = nearest_index(cy, a.h[i]), ix = nearest_index(cx, a.w[i]);
          ry[i] = rel_coord(cy, iy, a.h[i], a.cy0[i], a.cy1[i]);
          rx[i] = rel_coord(cx, ix, a.w[i], a.cx0[i], a.cx1[i]);
          off[i] = (((long long)b * a.h[i] + iy) * a.w[i] + ix) * (kH1 / 4);       // in float4 units
        }
      }
      // consecutive queries of a dense grid fall into the same source pixel (4*scale of them per 1/4-res pixel):
      // a P row is fetched only when the (warp-uniform) row offset changes, otherwise it is reused from registers
      float4 p4[3];
      long long held[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) { p4[i] = make_float4(0.f, 0.f, 0.f, 0.f); held[i] = -1; }
#pragma unroll 4
      for (int j = 0; j < 16; ++j) {
        const int row = gwarp * 16 + j;
        const bool valid = q0 + row < a.Q;
        float4 z = bias4;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          if (i < a.n_in) {
            const long long o = __shfl_sync(0xffffffffu, off[i], j);
            const float qy = __shfl_sync(0xffffffffu, ry[i], j), qx = __shfl_sync(0xffffffffu, rx[i], j);
            if (o != held[i]) {
              p4[i] = __ldg(reinterpret_cast<const float4*>(a.P[i]) + o + lane);
              held[i] = o;
            }
            z.x += p4[i].x + wy4[i].x * qy + wx4[i].x * qx;
            z.y += p4[i].y + wy4[i].y * qy + wx4[i].y * qx;
            z.z += p4[i].z + wy4[i].z * qy + wx4[i].z * qx;
            z.w += p4[i].w + wy4[i].w * qy + wx4[i].w * qx;
          }
        }
        if (!valid) z = make_float4(0.f, 0.f, 0.f, 0.f);
        put_quad(act_s, lo_off, row, lane * 4, fmaxf(z.x, 0.f), fmaxf(z.y, 0.f), fmaxf(z.z, 0.f), fmaxf(z.w, 0.f), split);
      }
    }
    umma::fence_proxy_async();
    group_sync(group);
    run_layer(0, umma::smem_u32(w2h), umma::smem_u32(w2l), 2, kH2 * 128, kH2);         // [128 x 128] . W2^T -> cols [0, 64)
    hidden_epilogue(0, a.b2);
    run_layer(64, umma::smem_u32(w3h), umma::smem_u32(w3l), 1, kH3 * 128, kH3);        // [128 x 64] . W3^T -> cols [64, 128)
    hidden_epilogue(64, a.b3);
    run_layer(0, umma::smem_u32(w4h), umma::smem_u32(w4l), 1, kOutPad * 128, kOutPad); // [128 x 64] . W4^T -> cols [0, 16)
    // ---- final epilogue: logits -> softmax -> 3x3 context upsample of the low-res disparity
    if (gwarp * 32 < kQT) {
      float v[32];
      umma::tmem_ld_32x32(tmem_d + ((uint32_t)(gwarp * 32) << 16), v);
      umma::tmem_ld_wait();
      const int gq = q0 + gwarp * 32 + lane;
      if (gq < a.Q) {
        float lg[kOut], m = -3.0e38f;
#pragma unroll
        for (int k = 0; k < kOut; ++k) {
          lg[k] = v[k] + __ldg(a.b4 + k);
          m = fmaxf(m, lg[k]);
        }
        if (a.logits) {
#pragma unroll
          for (int k = 0; k < kOut; ++k) a.logits[((long long)b * kOut + k) * a.Q + gq] = lg[k];
        }
        if (a.out) {
          const float2 c2 = __ldg(reinterpret_cast<const float2*>(a.coords + ((long long)b * a.Q + gq) * 2));
          const int iy = nearest_index(c2.x, a.hd), ix = nearest_index(c2.y, a.wd);
          const float sc = a.disp_scale ? __ldg(a.disp_scale + b) : 1.0f;
          const float* dm = a.disp + (long long)b * a.hd * a.wd;
          float s = 0.f, acc = 0.f;
#pragma unroll
          for (int k = 0; k < kOut; ++k) {
            const float e = __expf(lg[k] - m);
            const int yy = iy + k / 3 - 1, xx = ix + k % 3 - 1;                    // F.unfold(3, 1, 1) tap order
            const float d = (yy >= 0 && yy < a.hd && xx >= 0 && xx < a.wd) ? __ldg(dm + (long long)yy * a.wd + xx) * sc : 0.f;
            s += e;
            acc = fmaf(e, d, acc);
          }
          a.out[(long long)b * a.Q + gq] = acc / s;
        }
      }
    }
    umma::tc_fence_before();
    group_sync(group);
  }
  if (tid == 0 && !weights_ready) umma::mbar_wait(w_full, 0);
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(*tmem_slot, 128 * kGroups);
}

// standalone context_upsample_multiscale_train (submodule.py:357-372) for callers that supply their own weights
__global__ void context_upsample_kernel(const float* __restrict__ disp, const float* __restrict__ wts,
                                        const float* __restrict__ coords, float* __restrict__ out, int h, int w, int Q,
                                        long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int b = (int)(i / Q), q = (int)(i - (long long)b * Q);
  const float2 c2 = __ldg(reinterpret_cast<const float2*>(coords + i * 2));
  const int iy = nearest_index(c2.x, h), ix = nearest_index(c2.y, w);
  const float* dm = disp + (long long)b * h * w;
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int yy = iy + k / 3 - 1, xx = ix + k % 3 - 1;
    const float d = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(dm + (long long)yy * w + xx) : 0.f;
    acc += d * __ldg(wts + ((long long)b * 9 + k) * Q + q);
  }
  out[i] = acc;
}

}  // namespace

extern "C" int as_isu_affinity(const float* feat, int B, int C, int H, int W, float* aff, void* hi, void* lo, int c_pad,
                               int c_off, as_stream_t stream) {
  if (!feat || B <= 0 || C <= 0 || H <= 0 || W <= 0 || (!aff && !hi)) return AS_ERR_BAD_ARG;
  if (hi && ((c_pad & 7) || (c_off & 7) || c_off + 8 > c_pad)) return AS_ERR_ALIGNMENT;
  if (hi && (!as_aligned16(hi) || (lo && !as_aligned16(lo)))) return AS_ERR_ALIGNMENT;
  if (B > 65535 || as_ceil_div(H, kIsuT) > 65535) return AS_ERR_INDEX_RANGE;
  dim3 grid(as_ceil_div(W, kIsuT), as_ceil_div(H, kIsuT), B);
  isu_affinity_kernel<<<grid, kIsuThreads, 0, as_cu(stream)>>>(feat, aff, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, C, H, W, c_pad, c_off);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_liif_query(const as_liif_query_desc* d, as_stream_t stream) {
  if (!d || !d->coords || !d->wc || !d->b2 || !d->b3 || !d->b4 || !d->w2_hi || !d->w3_hi || !d->w4_hi) return AS_ERR_BAD_ARG;
  if (d->n_in < 1 || d->n_in > 3 || d->B <= 0 || d->Q <= 0) return AS_ERR_BAD_ARG;
  if (d->nsplit != 1 && d->nsplit != 3) return AS_ERR_BAD_ARG;
  if (d->nsplit == 3 && (!d->w2_lo || !d->w3_lo || !d->w4_lo)) return AS_ERR_BAD_ARG;
  if (!d->logits && !d->out) return AS_ERR_BAD_ARG;
  if (d->out && (!d->disp || d->hd <= 0 || d->wd <= 0)) return AS_ERR_BAD_ARG;
  QueryArgs a{};
  for (int i = 0; i < d->n_in; ++i) {
    if (!d->P[i] || d->h[i] <= 0 || d->w[i] <= 0 || !as_aligned16(d->P[i])) return AS_ERR_BAD_ARG;
    a.P[i] = d->P[i]; a.h[i] = d->h[i]; a.w[i] = d->w[i];
    const double ry = 1.0 / d->h[i], rx = 1.0 / d->w[i];           // make_coord: r = (v1 - v0) / (2 n)
    a.cy0[i] = (float)(-1.0 + ry); a.cy1[i] = (float)(2.0 * ry);
    a.cx0[i] = (float)(-1.0 + rx); a.cx1[i] = (float)(2.0 * rx);
  }
  a.n_in = d->n_in; a.coords = d->coords; a.wc = d->wc; a.b2 = d->b2; a.b3 = d->b3; a.b4 = d->b4;
  a.disp = d->disp; a.disp_scale = d->disp_scale; a.hd = d->hd; a.wd = d->wd; a.logits = d->logits; a.out = d->out;
  a.B = d->B; a.Q = d->Q; a.nsplit = d->nsplit;
  a.tiles_per_b = as_ceil_div(d->Q, kQT);
  const long long nt = (long long)a.tiles_per_b * d->B;
  if (nt >= (1LL << 31)) return AS_ERR_INDEX_RANGE;
  a.num_tiles = (int)nt;
  CUtensorMap m[6];
  int rc;
  {
    const uint64_t d2[2] = {(uint64_t)kH1, (uint64_t)kH2}, s2[1] = {(uint64_t)kH1 * 2};
    const uint32_t b2[2] = {64u, (uint32_t)kH2};
    const uint64_t d3[2] = {(uint64_t)kH2, (uint64_t)kH3}, s3[1] = {(uint64_t)kH2 * 2};
    const uint32_t b3[2] = {64u, (uint32_t)kH3};
    const uint64_t d4[2] = {(uint64_t)kH3, (uint64_t)kOutPad}, s4[1] = {(uint64_t)kH3 * 2};
    const uint32_t b4[2] = {64u, (uint32_t)kOutPad};
    if ((rc = umma::make_tmap_bf16(&m[0], d->w2_hi, 2, d2, s2, b2)) != AS_OK) return rc;
    if ((rc = umma::make_tmap_bf16(&m[2], d->w3_hi, 2, d3, s3, b3)) != AS_OK) return rc;
    if ((rc = umma::make_tmap_bf16(&m[4], d->w4_hi, 2, d4, s4, b4)) != AS_OK) return rc;
    if (d->nsplit == 3) {
      if ((rc = umma::make_tmap_bf16(&m[1], d->w2_lo, 2, d2, s2, b2)) != AS_OK) return rc;
      if ((rc = umma::make_tmap_bf16(&m[3], d->w3_lo, 2, d3, s3, b3)) != AS_OK) return rc;
      if ((rc = umma::make_tmap_bf16(&m[5], d->w4_lo, 2, d4, s4, b4)) != AS_OK) return rc;
    } else {
      m[1] = m[0]; m[3] = m[2]; m[5] = m[4];
    }
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaError_t e = cudaFuncSetAttribute(liif_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kQSmem);
  if (e != cudaSuccess) return (int)e;
  const long long want = (nt + kGroups - 1) / kGroups;
  const int grid = want < sms ? (int)want : sms;
  liif_query_kernel<<<grid, kQThreads, kQSmem, as_cu(stream)>>>(m[0], m[1], m[2], m[3], m[4], m[5], a);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_context_upsample_multiscale(const float* disp_low, const float* up_weights, const float* hr_coord, float* out,
                                              int B, int h, int w, int Q, as_stream_t stream) {
  if (!disp_low || !up_weights || !hr_coord || !out || B <= 0 || h <= 0 || w <= 0 || Q <= 0) return AS_ERR_BAD_ARG;
  const long long total = (long long)B * Q;
  const long long blocks = as_ceil_div_ll(total, 256);
  if (blocks >= (1LL << 31)) return AS_ERR_INDEX_RANGE;
  context_upsample_kernel<<<(unsigned)blocks, 256, 0, as_cu(stream)>>>(disp_low, up_weights, hr_coord, out, h, w, Q, total);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}
