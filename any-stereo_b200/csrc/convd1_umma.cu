// BasicMotionEncoder.convd1 (7x7, 1 -> 64 channels, update.py:80,87) + bias + ReLU on the tensor cores.
// The 49-tap neighbourhood of every pixel is laid out as one K = 64 row (im2col in shared memory, zero padded taps
// and image borders) of a 128-pixel operand tile; the [64 x 64] weights stay resident; one tcgen05 MMA group per tile
// (4 K-steps x 3 split products); epilogue = bias + ReLU -> bf16 hi/lo planes, like every other conv of the block.
// Replaces the CUDA-core kernel (as_convd1_split, kept as the exact-fp32 reference) on the tensor-core engines.
#include "umma.cuh"

namespace {

constexpr int kTW = 16, kTH = 8;                 // pixel tile = 128 rows of the MMA
constexpr int kThreads = 256;
constexpr int kBlk = 128 * 128;                  // [128 rows][64 K] bf16
constexpr int kWBlk = 64 * 128;                  // [64 rows][64 K] bf16
constexpr int kSmem = 1024 + 2 * kBlk + 2 * kWBlk + (kTH + 6) * (kTW + 6) * 4 + 64;

__device__ __forceinline__ void put_oct(uint32_t a_hi, uint32_t lo_off, int row, int k, const float (&v)[8], bool split, bool f16) {
  const uint32_t addr = a_hi + row * 128 + ((((uint32_t)(k & 63) << 1)) ^ ((uint32_t)(row & 7) << 4));
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    as_split2(v[2 * i], v[2 * i + 1], h[i], l[i], f16);
  }
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
  if (split)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + lo_off), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
}

__global__ void __launch_bounds__(kThreads, 3)
convd1_umma_kernel(const __grid_constant__ CUtensorMap tWh, const __grid_constant__ CUtensorMap tWl,
                   const float* __restrict__ disp, const float* __restrict__ bias, __nv_bfloat16* __restrict__ out_hi,
                   __nv_bfloat16* __restrict__ out_lo, int H, int W, int tiles_x, int tiles_y, int num_tiles, int pitch,
                   int coff, int nsplit, bool f16, int out_fmt, bool wide) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* act = smem;                                   // hi | lo
  uint8_t* wh = act + 2 * kBlk;
  uint8_t* wl = wh + kWBlk;
  float* patch = reinterpret_cast<float*>(wl + kWBlk);   // [kTH+6][kTW+6]
  uint64_t* bars = reinterpret_cast<uint64_t*>(patch + (kTH + 6) * (kTW + 6));
  uint64_t* w_full = bars;
  uint64_t* mma_done = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const bool split = nsplit == 3;
  if (tid == 0) {
    umma::prefetch_tmap(&tWh);
    umma::mbar_init(w_full, 1);
    umma::mbar_init(mma_done, 1);
    umma::fence_barrier_init();
  }
  if (warp == 1) {
    umma::tmem_alloc(tmem_slot, 64);
    umma::tmem_relinquish();
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  if (tid == 0) {
    umma::mbar_expect_tx(w_full, (uint32_t)kWBlk * (split ? 2u : 1u));
    umma::tma_load_2d(wh, &tWh, w_full, 0, 0);
    if (split) umma::tma_load_2d(wl, &tWl, w_full, 0, 0);
  }
  const uint32_t act_s = umma::smem_u32(act);
  const int tiles_per_img = tiles_x * tiles_y;
  const long long HW = (long long)H * W;
  uint32_t phase = 0;
  bool weights_ready = false;
  constexpr int PW = kTW + 6;
  for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
    const int b = t / tiles_per_img;
    const int r0 = t - b * tiles_per_img;
    const int ty = r0 / tiles_x, tx = r0 - ty * tiles_x;
    const int x0 = tx * kTW, y0 = ty * kTH;
    for (int i = tid; i < (kTH + 6) * PW; i += kThreads) {
      const int pr = i / PW, pc = i - pr * PW;
      const int yy = y0 + pr - 3, xx = x0 + pc - 3;
      patch[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(disp + (long long)b * HW + (long long)yy * W + xx) : 0.f;
    }
    __syncthreads();
    {   // im2col: row = pixel of the tile, K = tap (ky*7 + kx), taps 49..63 = 0
      const int row = tid >> 1, half = tid & 1;
      const int py = row >> 4, px = row & 15;
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int tap = half * 32 + o * 8 + i;
          const int ky = tap / 7, kx = tap - ky * 7;
          v[i] = tap < 49 ? patch[(py + ky) * PW + px + kx] : 0.f;
        }
        put_oct(act_s, kBlk, row, half * 32 + o * 8, v, split, f16);
      }
    }
    umma::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      if (!weights_ready) umma::mbar_wait(w_full, 0);
      umma::tc_fence_after();
      const uint32_t idesc = umma::idesc_16_f32(128, 64, f16);
      const uint32_t bh = umma::smem_u32(wh), bl = umma::smem_u32(wl);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t ko = (uint32_t)k * 32u;
        const uint64_t dah = umma::smem_desc_k_sw128(act_s + ko), dbh = umma::smem_desc_k_sw128(bh + ko);
        umma::mma_bf16_ss(tmem_d, dah, dbh, idesc, k ? 1u : 0u);
        if (split) {
          umma::mma_bf16_ss(tmem_d, dah, umma::smem_desc_k_sw128(bl + ko), idesc, 1u);
          umma::mma_bf16_ss(tmem_d, umma::smem_desc_k_sw128(act_s + kBlk + ko), dbh, idesc, 1u);
        }
      }
      umma::mma_commit(mma_done);
    }
    weights_ready = true;
    umma::mbar_wait(mma_done, phase);
    phase ^= 1;
    umma::tc_fence_after();
    {
      const int q = warp & 3, half = warp >> 2;
      float v[32];
      umma::tmem_ld_32x32(tmem_d + (uint32_t)(half * 32) + ((uint32_t)(q * 32) << 16), v);
      umma::tmem_ld_wait();
      const int row = q * 32 + lane;
      const int x = x0 + (row & 15), y = y0 + (row >> 4);
      if (x < W && y < H) {
        const long long o = ((long long)b * HW + (long long)y * W + x) * pitch + coff + half * 32;
        if (wide) {                               // 32-byte aligned rows: full-sector stores (common.cuh)
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + __ldg(bias + half * 32 + i), 0.f);
          as_store_split32_v8(v, out_hi, out_lo, o, out_fmt);
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t h[4];
            float yv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) yv[i] = fmaxf(v[j + i] + __ldg(bias + half * 32 + j + i), 0.f);
#pragma unroll
            for (int i = 0; i < 4; ++i) h[i] = as_cvt16x2(yv[2 * i], yv[2 * i + 1], f16);
            *reinterpret_cast<uint4*>(out_hi + o + j) = make_uint4(h[0], h[1], h[2], h[3]);
            if (out_lo) as_store_lo8(out_lo, o + j, yv, h, out_fmt);                 // 16-bit lo or the e5m2 pair plane
          }
        }
      }
    }
    umma::tc_fence_before();
    __syncthreads();
  }
  if (tid == 0 && !weights_ready) umma::mbar_wait(w_full, 0);
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem_d, 64);
}

}  // namespace

extern "C" int as_convd1_umma(const float* disp, const void* w_hi, const void* w_lo, const float* bias, void* out_hi, void* out_lo,
                              int B, int H, int W, int out_pitch, int out_coff, int nsplit, as_stream_t stream) {
  if (!disp || !w_hi || !bias || !out_hi || B <= 0 || H <= 0 || W <= 0 || out_pitch < out_coff + 64) return AS_ERR_BAD_ARG;
  if (nsplit != 1 && nsplit != 3) return AS_ERR_BAD_ARG;
  if (nsplit == 3 && (!w_lo || !out_lo)) return AS_ERR_BAD_ARG;
  if ((out_pitch & 7) || (out_coff & 7) || !as_aligned16(out_hi) || (out_lo && !as_aligned16(out_lo))) return AS_ERR_ALIGNMENT;
  CUtensorMap tWh, tWl;
  const uint64_t dims[2] = {64, 64};
  const uint64_t str[1] = {128};
  const uint32_t box[2] = {64u, 64u};
  int rc;
  if ((rc = umma::make_tmap_bf16(&tWh, w_hi, 2, dims, str, box)) != AS_OK) return rc;
  if (nsplit == 3) {
    if ((rc = umma::make_tmap_bf16(&tWl, w_lo, 2, dims, str, box)) != AS_OK) return rc;
  } else {
    tWl = tWh;
  }
  const int tiles_x = as_ceil_div(W, kTW), tiles_y = as_ceil_div(H, kTH);
  const long long nt = (long long)tiles_x * tiles_y * B;
  if (nt >= (1LL << 31)) return AS_ERR_INDEX_RANGE;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaError_t e = cudaFuncSetAttribute(convd1_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  if (e != cudaSuccess) return (int)e;
  const int grid = nt < 3LL * sms ? (int)nt : 3 * sms;
  const bool wide = !(out_pitch & 15) && !(out_coff & 15) && !(reinterpret_cast<uintptr_t>(out_hi) & 31) &&
                    !(reinterpret_cast<uintptr_t>(out_lo) & 31);                       // 256-bit stores need 32-byte aligned rows
  convd1_umma_kernel<<<grid, kThreads, kSmem, as_cu(stream)>>>(tWh, tWl, disp, bias, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo,
                                                               H, W, tiles_x, tiles_y, (int)nt, out_pitch, out_coff, nsplit,
                                                               as_operand_f16_internal() != 0, as_operand_fmt_internal(), wide);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}
