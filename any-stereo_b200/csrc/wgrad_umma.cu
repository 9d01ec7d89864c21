// a13-vi on the tensor cores: weight gradient of the update block's 1x1 / 3x3 convolutions (training, BASELINE.json
// config 5).  The reference gets it from autograd (cuDNN wgrad) over models/*/update.py:16-136.
//
//   dW[co][ci][ky][kx] = sum over pixels n=(b,y,x) of dY[n][co] * X[(b, y+ky-KH/2, x+kx-KW/2)][ci]
//
// is a GEMM with K = pixels.  Both operands are first written CHANNEL-major ("transposed planes" [C][B][H][Wp], bf16
// hi/lo, as_transpose_split) so that a K-block = 64 consecutive pixels of one image row is a K-major, 128B-swizzled
// TMA box [128 channels][64 pixels] for dY and for X alike.  The vertical tap shift is a coordinate offset of X's box
// (rows above/below the image are TMA out-of-bounds fill, like channels past the tensor).  The horizontal shift
// cannot be one: a box must start on a 16-byte boundary of the innermost dimension (a +-1 pixel = 2-byte offset raises
// an illegal-instruction fault), so as_transpose_split writes X three times, pre-shifted by -1/0/+1 pixels with the
// zero padding baked in, and the tap selects the copy through a fifth tensor-map dimension.  One CTA = one [128 co] x [128 ci] tile of one tap over a slice of the pixels (split-K):
// TMA producer thread -> 3-stage ring -> tcgen05.mma M=128,N=128 (hi*hi + hi*lo + lo*hi in the fp32-parity mode),
// fp32 accumulator in TMEM -> registers -> smem transpose -> coalesced fp32 atomics into ws[tap][co][ci];
// wgrad_finish_kernel adds ws into the nn.Conv2d layout [co][ci][ky][kx].
#include "umma.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kTile = 128 * 128;                 // bytes of one operand tile: 128 rows x 64 x 16-bit
constexpr int kStages = 3;
constexpr int kStageBytes = 4 * kTile;           // dY hi | dY lo | X hi | X lo
constexpr int kSmem = 1024 + kStages * kStageBytes + 128;

struct WgMaps {
  CUtensorMap a_hi, a_lo;
  CUtensorMap b_hi[3], b_lo[3];
};

struct WgParams {
  int H, KH, KW, Cout, Cin, XC;                  // XC = 64-pixel chunks per image row
  int num_src;
  int src_c0[4];                                 // first concatenated input channel of each source; [num_src] = Cin
  int ntiles, ksplit, kblocks;
  int nsplit, f16;
  float* ws;                                     // [KH*KW][Cout][Cin]
};

__global__ void __launch_bounds__(kThreads, 1) wgrad_umma_kernel(const __grid_constant__ WgMaps maps, const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* done = bars + 2 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const bool split = p.nsplit == 3;

  if (tid == 0) {
    umma::prefetch_tmap(&maps.a_hi);
    for (int s = 0; s < kStages; ++s) {
      umma::mbar_init(&full[s], 1);
      umma::mbar_init(&empty[s], 1);
    }
    umma::mbar_init(done, 1);
    umma::fence_barrier_init();
  }
  if (warp == 1) {
    umma::tmem_alloc(tmem_slot, 128);
    umma::tmem_relinquish();
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  // ---- which tile / tap / pixel slice
  int bid = blockIdx.x;
  const int ks = bid % p.ksplit; bid /= p.ksplit;
  const int nt = bid % p.ntiles; bid /= p.ntiles;
  const int taps = p.KH * p.KW;
  const int tap = bid % taps;
  const int mt = bid / taps;
  const int ky = tap / p.KW;
  const int dy = ky - p.KH / 2, dx = (tap - ky * p.KW) - p.KW / 2;
  const int co0 = mt * 128, ci0 = nt * 128;
  int si = 0, c_first = 0, ci_end = p.src_c0[1];  // static indices only: the tensor maps must stay in param space
#pragma unroll
  for (int s = 1; s < 3; ++s)
    if (s < p.num_src && ci0 >= p.src_c0[s]) { si = s; c_first = p.src_c0[s]; ci_end = p.src_c0[s + 1]; }
  const int cs0 = ci0 - c_first;                 // channel offset inside the source
  const int kb0 = (int)((long long)ks * p.kblocks / p.ksplit);
  const int kb1 = (int)((long long)(ks + 1) * p.kblocks / p.ksplit);

  if (warp == 0 && lane == 0) {                  // ---- TMA producer
    const CUtensorMap* bh = si == 0 ? &maps.b_hi[0] : si == 1 ? &maps.b_hi[1] : &maps.b_hi[2];
    const CUtensorMap* bl = si == 0 ? &maps.b_lo[0] : si == 1 ? &maps.b_lo[1] : &maps.b_lo[2];
    for (int kb = kb0; kb < kb1; ++kb) {
      const int it = kb - kb0, s = it % kStages, r = it / kStages;
      if (r > 0) umma::mbar_wait(&empty[s], (uint32_t)((r - 1) & 1));
      const int row = kb / p.XC, x0 = (kb - row * p.XC) * 64;
      const int b = row / p.H, y = row - b * p.H;
      uint8_t* st = smem + s * kStageBytes;
      umma::mbar_expect_tx(&full[s], (uint32_t)kTile * (split ? 4u : 2u));
      umma::tma_load_4d(st, &maps.a_hi, &full[s], x0, y, b, co0);
      if (split) umma::tma_load_4d(st + kTile, &maps.a_lo, &full[s], x0, y, b, co0);
      umma::tma_load_5d(st + 2 * kTile, bh, &full[s], x0, y + dy, b, cs0, dx + p.KW / 2);
      if (split) umma::tma_load_5d(st + 3 * kTile, bl, &full[s], x0, y + dy, b, cs0, dx + p.KW / 2);
    }
  } else if (warp == 1 && lane == 0) {           // ---- MMA issuer
    const uint32_t idesc = umma::idesc_16_f32(128, 128, p.f16 != 0);
    for (int kb = kb0; kb < kb1; ++kb) {
      const int it = kb - kb0, s = it % kStages, r = it / kStages;
      umma::mbar_wait(&full[s], (uint32_t)(r & 1));
      umma::tc_fence_after();
      const uint32_t base = umma::smem_u32(smem + s * kStageBytes);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t ko = (uint32_t)k * 32u;
        const uint64_t dah = umma::smem_desc_k_sw128(base + ko), dbh = umma::smem_desc_k_sw128(base + 2 * kTile + ko);
        umma::mma_bf16_ss(tmem_d, dah, dbh, idesc, (it > 0 || k > 0) ? 1u : 0u);
        if (split) {
          umma::mma_bf16_ss(tmem_d, dah, umma::smem_desc_k_sw128(base + 3 * kTile + ko), idesc, 1u);
          umma::mma_bf16_ss(tmem_d, umma::smem_desc_k_sw128(base + kTile + ko), dbh, idesc, 1u);
        }
      }
      umma::mma_commit(&empty[s]);               // stage reusable once these MMAs have read it
    }
    umma::mma_commit(done);
  }
  __syncwarp();

  // ---- epilogue: TMEM -> registers -> per-warp smem transpose -> coalesced atomics
  umma::mbar_wait(done, 0);
  umma::tc_fence_after();
  float* sT = reinterpret_cast<float*>(smem) + warp * (32 * 33);     // the ring is idle now
  float* wsb = p.ws + (long long)tap * p.Cout * p.Cin;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    float v[32];
    umma::tmem_ld_32x32(tmem_d + (uint32_t)(c * 32) + ((uint32_t)(warp * 32) << 16), v);
    umma::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) sT[lane * 33 + j] = v[j];
    __syncwarp();
    const int ci = ci0 + c * 32 + lane;
    if (ci < ci_end) {
      for (int rr = 0; rr < 32; ++rr) {
        const int co = co0 + warp * 32 + rr;
        if (co < p.Cout) atomicAdd(wsb + (long long)co * p.Cin + ci, sT[rr * 33 + lane]);
      }
    }
    __syncwarp();
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem_d, 128);
}

// dw[co][ci][t] += ws[t][co][ci]
__global__ void wgrad_finish_kernel(const float* __restrict__ ws, float* __restrict__ dw, int taps, long long cc, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long e = idx / taps;                // co*Cin + ci
  const int t = (int)(idx - e * taps);
  dw[idx] += ws[(long long)t * cc + e];
}

// fp32 pixel-major [B*H*W][pitch] (channels coff..coff+C) -> 16-bit hi/lo planes [nshift][C][B*H][Wp];
// copy j holds the row shifted by dx = j - nshift/2 pixels: out_j[c][row][x] = in[row][x + dx][c] (0 outside the row)
__global__ void __launch_bounds__(256) transpose_split_kernel(const float* __restrict__ in, int pitch, int coff, int C, int W,
                                                              int Wp, long long rows, int nshift, uint16_t* __restrict__ hi,
                                                              uint16_t* __restrict__ lo, bool f16) {
  __shared__ float tile[66][33];                      // pixels x0-1 .. x0+64
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long row = blockIdx.y;
  const int x0 = blockIdx.x * 64, c0 = blockIdx.z * 32;
  for (int i = warp; i < 66; i += 8) {
    const int x = x0 - 1 + i, c = c0 + lane;
    tile[i][lane] = (x >= 0 && x < W && c < C) ? __ldg(in + (row * W + x) * pitch + coff + c) : 0.f;
  }
  __syncthreads();
  const long long plane = (long long)C * rows * Wp;
  for (int j = 0; j < nshift; ++j) {
    const int dx = j - nshift / 2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int cl = warp * 4 + i, c = c0 + cl, x = x0 + 2 * lane;
      if (c < C && x < Wp) {
        uint32_t h, l;
        as_split2(tile[2 * lane + 1 + dx][cl], tile[2 * lane + 2 + dx][cl], h, l, f16);
        const long long o = j * plane + ((long long)c * rows + row) * Wp + x;
        *reinterpret_cast<uint32_t*>(hi + o) = h;
        if (lo) *reinterpret_cast<uint32_t*>(lo + o) = l;
      }
    }
  }
}

}  // namespace

extern "C" int as_transpose_split(const float* in, int pitch, int coff, int C, int B, int H, int W, void* hi, void* lo, int Wp,
                                  int nshift, as_stream_t stream) {
  if (!in || !hi || C <= 0 || B <= 0 || H <= 0 || W <= 0 || pitch < coff + C || Wp < W) return AS_ERR_BAD_ARG;
  if (nshift != 1 && nshift != 3) return AS_ERR_UNSUPPORTED;
  if ((Wp & 7) || !as_aligned16(hi) || (lo && !as_aligned16(lo))) return AS_ERR_ALIGNMENT;
  const long long rows = (long long)B * H;
  if (rows > 65535) return AS_ERR_INDEX_RANGE;
  dim3 grid((unsigned)as_ceil_div(Wp, 64), (unsigned)rows, (unsigned)as_ceil_div(C, 32));
  transpose_split_kernel<<<grid, 256, 0, as_cu(stream)>>>(in, pitch, coff, C, W, Wp, rows, nshift, (uint16_t*)hi, (uint16_t*)lo,
                                                          as_operand_f16_internal() != 0);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_conv2d_wgrad_umma(const as_wgrad_umma_desc* d, as_stream_t stream) {
  if (!d || !d->dy_hi || !d->ws || !d->dw_acc) return AS_ERR_BAD_ARG;
  if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->Cout <= 0 || d->num_src < 1 || d->num_src > 3) return AS_ERR_BAD_ARG;
  if (!((d->KH == 1 && d->KW == 1) || (d->KH == 3 && d->KW == 3))) return AS_ERR_UNSUPPORTED;
  if (d->nsplit != 1 && d->nsplit != 3) return AS_ERR_BAD_ARG;
  if (d->nsplit == 3 && !d->dy_lo) return AS_ERR_BAD_ARG;
  if (d->Wp < d->W || (d->Wp & 7)) return AS_ERR_ALIGNMENT;
  WgParams p{};
  p.H = d->H; p.KH = d->KH; p.KW = d->KW; p.Cout = d->Cout;
  p.num_src = d->num_src;
  int cin = 0;
  for (int s = 0; s < d->num_src; ++s) {
    if (!d->src[s].hi || (d->nsplit == 3 && !d->src[s].lo) || d->src[s].channels <= 0) return AS_ERR_BAD_ARG;
    if (s + 1 < d->num_src && (d->src[s].channels & 127)) return AS_ERR_UNSUPPORTED;   // N tiles must not straddle sources
    if (!as_aligned16(d->src[s].hi) || (d->src[s].lo && !as_aligned16(d->src[s].lo))) return AS_ERR_ALIGNMENT;
    p.src_c0[s] = cin;
    cin += d->src[s].channels;
  }
  p.src_c0[d->num_src] = cin;
  p.Cin = cin;
  p.XC = as_ceil_div(d->W, 64);
  const long long kblocks = (long long)d->B * d->H * p.XC;
  if (kblocks >= (1LL << 30)) return AS_ERR_INDEX_RANGE;
  p.kblocks = (int)kblocks;
  p.ntiles = as_ceil_div(cin, 128);
  const int mtiles = as_ceil_div(d->Cout, 128);
  const int taps = d->KH * d->KW;
  const int tiles = mtiles * taps * p.ntiles;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int ksplit = as_ceil_div(2 * sms, tiles);
  if (ksplit > p.kblocks / 8) ksplit = p.kblocks / 8;          // at least 8 K-blocks per slice
  if (ksplit < 1) ksplit = 1;
  p.ksplit = ksplit;
  p.nsplit = d->nsplit;
  p.f16 = as_operand_f16_internal();
  p.ws = d->ws;

  WgMaps maps;
  const uint64_t rowb = (uint64_t)d->Wp * 2;
  const uint64_t gx = (uint64_t)d->W;                 // columns W..Wp-1 of a row are out of bounds = zero fill
  const uint32_t box[4] = {64u, 1u, 1u, 128u};
  int rc;
  {
    const uint64_t dims[4] = {gx, (uint64_t)d->H, (uint64_t)d->B, (uint64_t)d->Cout};
    const uint64_t str[3] = {rowb, rowb * d->H, rowb * d->H * d->B};
    if ((rc = umma::make_tmap_bf16(&maps.a_hi, d->dy_hi, 4, dims, str, box)) != AS_OK) return rc;
    if (d->nsplit == 3) {
      if ((rc = umma::make_tmap_bf16(&maps.a_lo, d->dy_lo, 4, dims, str, box)) != AS_OK) return rc;
    } else {
      maps.a_lo = maps.a_hi;
    }
  }
  for (int s = 0; s < d->num_src; ++s) {        // [KW shifted copies][channels][B][H][Wp]
    const uint64_t ch = (uint64_t)d->src[s].channels;
    const uint64_t dims[5] = {gx, (uint64_t)d->H, (uint64_t)d->B, ch, (uint64_t)d->KW};
    const uint64_t str[4] = {rowb, rowb * d->H, rowb * d->H * d->B, rowb * d->H * d->B * ch};
    const uint32_t box5[5] = {64u, 1u, 1u, 128u, 1u};
    if ((rc = umma::make_tmap_bf16(&maps.b_hi[s], d->src[s].hi, 5, dims, str, box5)) != AS_OK) return rc;
    if (d->nsplit == 3) {
      if ((rc = umma::make_tmap_bf16(&maps.b_lo[s], d->src[s].lo, 5, dims, str, box5)) != AS_OK) return rc;
    } else {
      maps.b_lo[s] = maps.b_hi[s];
    }
  }
  for (int s = d->num_src; s < 3; ++s) { maps.b_hi[s] = maps.b_hi[0]; maps.b_lo[s] = maps.b_lo[0]; }

  const long long cc = (long long)d->Cout * cin;
  cudaError_t e = cudaMemsetAsync(d->ws, 0, (size_t)(cc * taps) * sizeof(float), as_cu(stream));
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  if (e != cudaSuccess) return (int)e;
  wgrad_umma_kernel<<<(unsigned)(tiles * ksplit), kThreads, kSmem, as_cu(stream)>>>(maps, p);
  AS_RETURN_IF_LAUNCH_FAILED();
  const long long total = cc * taps;
  wgrad_finish_kernel<<<(unsigned)as_ceil_div_ll(total, 256), 256, 0, as_cu(stream)>>>(d->ws, d->dw_acc, taps, cc, total);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}
