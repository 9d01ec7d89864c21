// Small memory-bound kernels around the tcgen05 convolutions: fp32 -> bf16 hi/lo plane conversion (with the
// layout move, pooling or bilinear resize fused in), weight packing, the 7x7 single-channel convd1 and the
// disparity-head gather.  Reference: models/*/update.py:80,87 (convd1), :94-95 (pool2x), :100-102 (interp),
// :23-24 (DispHead.conv2).
#include <cuda_bf16.h>
#include "common.cuh"

namespace {

// fmt: AS_FMT_BF16 / AS_FMT_F16 (16-bit hi + lo = x - hi) / AS_FMT_F16F8 (half hi + e5m2 pair plane, common.cuh)
__device__ __forceinline__ void store_split4(float4 v, __nv_bfloat16* hi, __nv_bfloat16* lo, long long off, int fmt) {
  uint2 h;
  h.x = as_cvt16x2(v.x, v.y, fmt != 0);
  h.y = as_cvt16x2(v.z, v.w, fmt != 0);
  *reinterpret_cast<uint2*>(hi + off) = h;
  if (lo) as_store_lo4(lo, off, v, h, fmt);
}

__global__ void split_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                             long long n4, int fmt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  store_split4(__ldg(reinterpret_cast<const float4*>(in) + i), hi, lo, i * 4, fmt);
}

// [B,C,HW] fp32 -> [B*HW][Cp] bf16 hi/lo, channels >= C zero-filled
__global__ void __launch_bounds__(256) nchw_to_nhwc_split_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi,
                                                                 __nv_bfloat16* __restrict__ lo, int C, long long HW, int Cp, int fmt) {
  __shared__ float t[32][33];
  const int b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
#pragma unroll
  for (int r = ly; r < 32; r += 8) {
    const int c = c0 + r;
    const long long pp = p0 + lx;
    t[r][lx] = (c < C && pp < HW) ? __ldg(in + ((long long)b * C + c) * HW + pp) : 0.f;
  }
  __syncthreads();
  // each thread writes 4 consecutive channels of one pixel
  const int pr = threadIdx.x >> 3, cq = (threadIdx.x & 7) * 4;
  const long long pp = p0 + pr;
  if (pp < HW && c0 + cq < Cp)
    store_split4(make_float4(t[cq][pr], t[cq + 1][pr], t[cq + 2][pr], t[cq + 3][pr]), hi, lo,
                 ((long long)b * HW + pp) * Cp + c0 + cq, fmt);
}

// grid = (ceil(Wo*C4 / 256), Ho, B): the row / image indices come from the block, one 32-bit division per thread
__global__ void pool2x_split_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                    int H, int W, int Ho, int Wo, int C4, int fmt) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= Wo * C4) return;
  const int xo = r / C4, c4 = r - xo * C4;
  const int yo = blockIdx.y, b = blockIdx.z;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int y = 2 * yo - 1 + i;
    if (y < 0 || y >= H) continue;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int x = 2 * xo - 1 + j;
      if (x < 0 || x >= W) continue;
      const float4 v = __ldg(reinterpret_cast<const float4*>(in) + (((long long)b * H + y) * W + x) * C4 + c4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  const float inv = 1.0f / 9.0f;
  const long long idx = (((long long)b * Ho + yo) * Wo + xo) * C4 + c4;
  store_split4(make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv), hi, lo, idx * 4, fmt);
}

__global__ void interp_split_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                    int Hi, int Wi, int Ho, int Wo, int C4, float sy, float sx, int fmt) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= Wo * C4) return;
  const int xo = r / C4, c4 = r - xo * C4;
  const int yo = blockIdx.y, b = blockIdx.z;
  const float fy = sy * yo, fx = sx * xo;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = min(y0 + 1, Hi - 1), x1 = min(x0 + 1, Wi - 1);
  const float ly = fy - y0, lx = fx - x0, hy = 1.0f - ly, hx = 1.0f - lx;
  const float4* base = reinterpret_cast<const float4*>(in) + (long long)b * Hi * Wi * C4 + c4;
  const float4 v00 = __ldg(base + ((long long)y0 * Wi + x0) * C4), v01 = __ldg(base + ((long long)y0 * Wi + x1) * C4);
  const float4 v10 = __ldg(base + ((long long)y1 * Wi + x0) * C4), v11 = __ldg(base + ((long long)y1 * Wi + x1) * C4);
  float4 o;
  o.x = hy * (hx * v00.x + lx * v01.x) + ly * (hx * v10.x + lx * v11.x);
  o.y = hy * (hx * v00.y + lx * v01.y) + ly * (hx * v10.y + lx * v11.y);
  o.z = hy * (hx * v00.z + lx * v01.z) + ly * (hx * v10.z + lx * v11.z);
  o.w = hy * (hx * v00.w + lx * v01.w) + ly * (hx * v10.w + lx * v11.w);
  const long long idx = (((long long)b * Ho + yo) * Wo + xo) * C4 + c4;
  store_split4(o, hi, lo, idx * 4, fmt);
}

// 8 channels per thread: 256-bit loads of the four neighbours, one 16-byte hi store + the second plane.  (The 4-channel
// kernel above ran at 1.5 TB/s: 7.7 M threads of ~80 dependent instructions each at config 2.)  Same arithmetic per channel.
__global__ void __launch_bounds__(256) interp_split8_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi,
                                                            __nv_bfloat16* __restrict__ lo, int Hi, int Wi, int Ho, int Wo, int C8,
                                                            float sy, float sx, int fmt) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= Wo * C8) return;
  const int xo = r / C8, c8 = r - xo * C8;
  const int yo = blockIdx.y, b = blockIdx.z;
  const float fy = sy * yo, fx = sx * xo;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = min(y0 + 1, Hi - 1), x1 = min(x0 + 1, Wi - 1);
  const float ly = fy - y0, lx = fx - x0, hy = 1.0f - ly, hx = 1.0f - lx;
  const int C = C8 * 8;
  const float* base = in + (long long)b * Hi * Wi * C + c8 * 8;
  float v00[8], v01[8], v10[8], v11[8], o[8];
  as_ldg256f(base + ((long long)y0 * Wi + x0) * C, v00);
  as_ldg256f(base + ((long long)y0 * Wi + x1) * C, v01);
  as_ldg256f(base + ((long long)y1 * Wi + x0) * C, v10);
  as_ldg256f(base + ((long long)y1 * Wi + x1) * C, v11);
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = hy * (hx * v00[i] + lx * v01[i]) + ly * (hx * v10[i] + lx * v11[i]);
  const long long off = ((((long long)b * Ho + yo) * Wo + xo) * C8 + c8) * 8;
  uint32_t h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = as_cvt16x2(o[2 * i], o[2 * i + 1], fmt != 0);
  *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
  if (lo) as_store_lo8(lo, off, o, h, fmt);
}

__global__ void pack_weight_bf16_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ hi,
                                        __nv_bfloat16* __restrict__ lo, int Cout, int Cin, int T, int n_pad, int cin_pad,
                                        long long total, int fmt) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  // idx = (n * T + tap) * cin_pad + c
  const int c = (int)(idx % cin_pad);
  const long long r = idx / cin_pad;
  const int tap = (int)(r % T);
  const int n = (int)(r / T);
  float v = 0.f;
  if (n < Cout && c < Cin) v = w[((long long)n * Cin + c) * T + tap];
  const bool f16 = fmt != 0;
  const uint32_t h = as_cvt16x2(v, 0.f, f16);
  reinterpret_cast<unsigned short*>(hi)[idx] = (unsigned short)(h & 0xFFFFu);
  if (lo) {
    const float hv = as_widen_lo16(h, f16);
    if (fmt != 2) {
      reinterpret_cast<unsigned short*>(lo)[idx] = (unsigned short)(as_cvt16x2(v - hv, 0.f, f16) & 0xFFFFu);
    } else {            // AS_FMT_F16F8 weights: [ e5m2(hi * 2^-6) | e5m2(lo * 2^8) ] per 64-wide K chunk (row length % 64 == 0)
      uint8_t* b = reinterpret_cast<uint8_t*>(lo) + as_x8_byte(idx);
      b[0] = (uint8_t)(as_e5m2x4(hv * kX8WgtHiScale, 0.f, 0.f, 0.f) & 0xFFu);
      b[64] = (uint8_t)(as_e5m2x4((v - hv) * kX8WgtLoScale, 0.f, 0.f, 0.f) & 0xFFu);
    }
  }
}

// 7x7 conv, 1 input channel -> 64, + bias, relu.  CTA = 32x8 pixel tile: the (8+6)x(32+6) disparity patch and the
// 64x49 weights live in shared memory.  One thread = 4 adjacent pixels x 16 output channels: per kernel row it reads
// 10 patch values and 7x16 weights (warp-uniform 128-bit broadcasts) for 448 FMAs -- ~12 FMAs per smem load.
constexpr int kD1TX = 32, kD1TY = 8;
// F32OUT: the same arithmetic with an fp32 pixel-major output (training keeps relu(convd1) for the backward pass)
template <bool F32OUT>
__global__ void __launch_bounds__(256) convd1_split_kernel(const float* __restrict__ disp, const float* __restrict__ w,
                                                           const float* __restrict__ bias, __nv_bfloat16* __restrict__ hi,
                                                           __nv_bfloat16* __restrict__ lo, float* __restrict__ out_f32, int H,
                                                           int W, int pitch, int coff, int fmt) {
  __shared__ __align__(16) float ws[49 * 64];     // [tap][channel]
  __shared__ float bs[64];
  __shared__ float patch[kD1TY + 6][kD1TX + 6 + 2];
  const int tid = threadIdx.x;
  const int b = blockIdx.z, x0 = blockIdx.x * kD1TX, y0 = blockIdx.y * kD1TY;
  const long long HW = (long long)H * W;
  for (int i = tid; i < 64 * 49; i += 256) { const int c = i / 49, t = i - c * 49; ws[t * 64 + c] = __ldg(w + i); }
  if (tid < 64) bs[tid] = __ldg(bias + tid);
  for (int i = tid; i < (kD1TY + 6) * (kD1TX + 6); i += 256) {
    const int r = i / (kD1TX + 6), c = i - r * (kD1TX + 6);
    const int yy = y0 + r - 3, xx = x0 + c - 3;
    patch[r][c] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(disp + (long long)b * HW + (long long)yy * W + xx) : 0.f;
  }
  __syncthreads();
  const int xq = tid & 7, ty = (tid >> 3) & 7, cg = (tid >> 6) * 16;
  float acc[4][16];
#pragma unroll
  for (int px = 0; px < 4; ++px)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[px][j] = bs[cg + j];
  for (int dy = 0; dy < 7; ++dy) {
    float pr[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) pr[i] = patch[ty + dy][xq * 4 + i];
#pragma unroll
    for (int dx = 0; dx < 7; ++dx) {
      const float4* wp = reinterpret_cast<const float4*>(ws + (dy * 7 + dx) * 64 + cg);
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const float4 w4 = wp[j4];
#pragma unroll
        for (int px = 0; px < 4; ++px) {
          const float d = pr[px + dx];
          acc[px][4 * j4] = fmaf(w4.x, d, acc[px][4 * j4]);
          acc[px][4 * j4 + 1] = fmaf(w4.y, d, acc[px][4 * j4 + 1]);
          acc[px][4 * j4 + 2] = fmaf(w4.z, d, acc[px][4 * j4 + 2]);
          acc[px][4 * j4 + 3] = fmaf(w4.w, d, acc[px][4 * j4 + 3]);
        }
      }
    }
  }
  const int y = y0 + ty;
  if (y >= H) return;
#pragma unroll
  for (int px = 0; px < 4; ++px) {
    const int x = x0 + xq * 4 + px;
    if (x >= W) continue;
    const long long n = (long long)b * HW + (long long)y * W + x;
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 y4 = make_float4(fmaxf(acc[px][j], 0.f), fmaxf(acc[px][j + 1], 0.f), fmaxf(acc[px][j + 2], 0.f),
                                    fmaxf(acc[px][j + 3], 0.f));
      if (F32OUT) *reinterpret_cast<float4*>(out_f32 + n * pitch + coff + cg + j) = y4;
      else store_split4(y4, hi, lo, n * pitch + coff + cg + j, fmt);
    }
  }
}

__global__ void disp_delta_kernel(const float* __restrict__ u, const float* __restrict__ bias2, float* __restrict__ delta,
                                  int H, int W, long long N) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const long long HW = (long long)H * W;
  const int rem = (int)(n % HW);
  const int y = rem / W, x = rem - y * W;
  float acc = __ldg(bias2);
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int dy = t / 3 - 1, dx = t % 3 - 1;
    const int yy = y + dy, xx = x + dx;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) acc += __ldg(u + (n + (long long)dy * W + dx) * 9 + t);
  }
  delta[n] = acc;
}

}  // namespace

extern "C" int as_split_f32(const float* in, void* hi, void* lo, long long n, as_stream_t stream) {
  if (!in || !hi || n <= 0) return AS_ERR_BAD_ARG;
  if ((n & 3) || !as_aligned16(in)) return AS_ERR_ALIGNMENT;
  if (lo && as_operand_fmt_internal() == AS_FMT_F16F8 && (n & 63)) return AS_ERR_UNSUPPORTED;   // whole 64-channel chunks
  split_kernel<<<(unsigned)as_ceil_div_ll(n / 4, 256), 256, 0, as_cu(stream)>>>(in, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, n / 4, as_operand_fmt_internal());
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_nchw_to_nhwc_split(const float* in, void* hi, void* lo, int B, int C, int H, int W, int c_pad,
                                     as_stream_t stream) {
  if (!in || !hi || B <= 0 || C <= 0 || H <= 0 || W <= 0 || c_pad < C) return AS_ERR_BAD_ARG;
  if ((c_pad & 31) || B > 65535) return AS_ERR_UNSUPPORTED;
  if (lo && as_operand_fmt_internal() == AS_FMT_F16F8 && (c_pad & 63)) return AS_ERR_UNSUPPORTED;
  const long long HW = (long long)H * W;
  dim3 grid((unsigned)as_ceil_div_ll(HW, 32), c_pad / 32, B);
  nchw_to_nhwc_split_kernel<<<grid, 256, 0, as_cu(stream)>>>(in, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, C, HW, c_pad, as_operand_fmt_internal());
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_pool2x_nhwc_split(const float* in, void* hi, void* lo, int B, int H, int W, int C, as_stream_t stream) {
  if (!in || !hi || B <= 0 || H <= 0 || W <= 0 || C <= 0) return AS_ERR_BAD_ARG;
  if ((C & 3) || !as_aligned16(in)) return AS_ERR_ALIGNMENT;
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  if (B > 65535 || Ho > 65535 || (long long)Wo * (C / 4) >= (1LL << 31)) return AS_ERR_INDEX_RANGE;
  dim3 grid(as_ceil_div(Wo * (C / 4), 256), Ho, B);
  pool2x_split_kernel<<<grid, 256, 0, as_cu(stream)>>>(in, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, H, W, Ho, Wo, C / 4,
                                                       as_operand_fmt_internal());
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_interp_bilinear_nhwc_split(const float* in, void* hi, void* lo, int B, int Hin, int Win, int Hout,
                                             int Wout, int C, as_stream_t stream) {
  if (!in || !hi || B <= 0 || Hin <= 0 || Win <= 0 || Hout <= 0 || Wout <= 0 || C <= 0) return AS_ERR_BAD_ARG;
  if ((C & 3) || !as_aligned16(in)) return AS_ERR_ALIGNMENT;
  const float sy = Hout > 1 ? (float)(Hin - 1) / (float)(Hout - 1) : 0.f;
  const float sx = Wout > 1 ? (float)(Win - 1) / (float)(Wout - 1) : 0.f;
  if (B > 65535 || Hout > 65535 || (long long)Wout * (C / 4) >= (1LL << 31)) return AS_ERR_INDEX_RANGE;
  if (!(C & 7) && !(reinterpret_cast<uintptr_t>(in) & 31)) {
    dim3 grid8(as_ceil_div(Wout * (C / 8), 256), Hout, B);
    interp_split8_kernel<<<grid8, 256, 0, as_cu(stream)>>>(in, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, Hin, Win, Hout, Wout,
                                                            C / 8, sy, sx, as_operand_fmt_internal());
    AS_RETURN_IF_LAUNCH_FAILED();
    return AS_OK;
  }
  dim3 grid(as_ceil_div(Wout * (C / 4), 256), Hout, B);
  interp_split_kernel<<<grid, 256, 0, as_cu(stream)>>>(in, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, Hin, Win, Hout, Wout, C / 4,
                                                       sy, sx, as_operand_fmt_internal());
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_pack_conv_weight_bf16(const float* w_oihw, void* w_hi, void* w_lo, int Cout, int Cin, int KH, int KW,
                                        int n_pad, int cin_pad, as_stream_t stream) {
  if (!w_oihw || !w_hi || Cout <= 0 || Cin <= 0 || KH <= 0 || KW <= 0 || n_pad < Cout || cin_pad < Cin) return AS_ERR_BAD_ARG;
  const long long total = (long long)n_pad * KH * KW * cin_pad;
  if (w_lo && as_operand_fmt_internal() == AS_FMT_F16F8 && (((long long)KH * KW * cin_pad) & 63)) return AS_ERR_UNSUPPORTED;
  pack_weight_bf16_kernel<<<(unsigned)as_ceil_div_ll(total, 256), 256, 0, as_cu(stream)>>>(
      w_oihw, (__nv_bfloat16*)w_hi, (__nv_bfloat16*)w_lo, Cout, Cin, KH * KW, n_pad, cin_pad, total, as_operand_fmt_internal());
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_convd1_split(const float* disp, const float* w, const float* bias, void* hi, void* lo, int B, int H,
                               int W, int out_pitch, int out_coff, as_stream_t stream) {
  if (!disp || !w || !bias || !hi || B <= 0 || H <= 0 || W <= 0 || out_pitch < out_coff + 64) return AS_ERR_BAD_ARG;
  if ((out_pitch & 3) || (out_coff & 3)) return AS_ERR_ALIGNMENT;
  if (B > 65535 || as_ceil_div(H, kD1TY) > 65535) return AS_ERR_UNSUPPORTED;
  dim3 grid(as_ceil_div(W, kD1TX), as_ceil_div(H, kD1TY), B);
  convd1_split_kernel<false><<<grid, 256, 0, as_cu(stream)>>>(disp, w, bias, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, nullptr,
                                                              H, W, out_pitch, out_coff, as_operand_fmt_internal());
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_convd1_fp32(const float* disp, const float* w, const float* bias, float* out, int B, int H, int W,
                              int out_pitch, int out_coff, as_stream_t stream) {
  if (!disp || !w || !bias || !out || B <= 0 || H <= 0 || W <= 0 || out_pitch < out_coff + 64) return AS_ERR_BAD_ARG;
  if ((out_pitch & 3) || (out_coff & 3) || !as_aligned16(out)) return AS_ERR_ALIGNMENT;
  if (B > 65535 || as_ceil_div(H, kD1TY) > 65535) return AS_ERR_UNSUPPORTED;
  dim3 grid(as_ceil_div(W, kD1TX), as_ceil_div(H, kD1TY), B);
  convd1_split_kernel<true><<<grid, 256, 0, as_cu(stream)>>>(disp, w, bias, nullptr, nullptr, out, H, W, out_pitch, out_coff,
                                                             0);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_disp_delta(const float* u, const float* bias2, float* delta, int B, int H, int W, as_stream_t stream) {
  if (!u || !bias2 || !delta || B <= 0 || H <= 0 || W <= 0) return AS_ERR_BAD_ARG;
  const long long N = (long long)B * H * W;
  disp_delta_kernel<<<(unsigned)as_ceil_div_ll(N, 256), 256, 0, as_cu(stream)>>>(u, bias2, delta, H, W, N);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}
