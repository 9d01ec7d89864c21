// a8-a11 (exact-fp32 mode): implicit-GEMM convolution on CUDA cores with the update block's consumers
// fused into the epilogue, plus the small helpers between the GRU scales.
//
// Reference: models/*/update.py -- ConvGRU :26-41, BasicMotionEncoder :73-92, DispHead :16-24,
// pool2x :94-95, interp :100-102.  GEMM view: M = pixels, N = Cout, K = taps x Cin where Cin is the
// concatenation of up to 4 source tensors (torch.cat of update.py:35-36,39,90 is never materialised).
// This kernel is the bit-stable fp32 baseline the tcgen05 path (conv_umma.cu) is checked against on
// the GPU at full size.
#include "simt_gemm.cuh"

namespace {

using VT = SimtTile<128, 64, 16, 8, 4>;   // 128 pixels x 64 channels per CTA, 256 threads

struct ConvParams {
  int B, H, W, KH, KW, Cout, Cin;
  int num_src;
  const float* src_ptr[AS_MAX_SRC];
  int src_ch[AS_MAX_SRC];
  int src_pitch[AS_MAX_SRC];
  const float* weight;   // [KH*KW*Cin][Cout]
  const float* bias;
  int epilogue;
  float* out; int out_pitch, out_coff, out_layout;
  const float* ctx; int ctx_pitch;
  const float* h;
  float* z;
  float* save;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

template <bool NCHW_IN>
__global__ void __launch_bounds__(VT::kThreads) conv_simt_kernel(ConvParams p) {
  __shared__ __align__(16) float smem[VT::kSmemFloats];
  float* sA = smem;
  float* sB = smem + 16 * VT::kSA;
  const int tid = threadIdx.x;
  const int ty = tid / 16, tx = tid % 16;
  const long long HW = (long long)p.H * p.W;
  const long long N = (long long)p.B * HW;
  const long long m0 = (long long)blockIdx.x * 128;
  const int n0 = blockIdx.y * 64;
  const int K = p.KH * p.KW * p.Cin;
  const int ph = p.KH / 2, pw = p.KW / 2;

  // pixel coordinates of the A-tile rows this thread loads
  // NHWC: 8 pixels (m = tid/16 + 16*i), channel lane kk = tid%16 ; NCHW: 1 pixel (m = tid%128), kk = tid/128 + 2*i
  constexpr int NP = NCHW_IN ? 1 : 8;
  int py[NP], px[NP];
  long long pbase[NP];   // pixel index (NHWC) or b*Cin_src*HW + y*W + x handled later (NCHW)
  int pb[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const long long n = m0 + (NCHW_IN ? (tid & 127) : (tid / 16 + 16 * i));
    if (n < N) {
      const int b = (int)(n / HW);
      const int rem = (int)(n - (long long)b * HW);
      py[i] = rem / p.W; px[i] = rem - py[i] * p.W; pb[i] = b; pbase[i] = n;
    } else {
      py[i] = -100000; px[i] = 0; pb[i] = 0; pbase[i] = 0;
    }
  }

  float acc[8][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    // ---- A tile
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int kk = NCHW_IN ? (tid / 128 + 2 * i) : (tid % 16);
      const int mi = NCHW_IN ? (tid & 127) : (tid / 16 + 16 * i);
      const int pi = NCHW_IN ? 0 : i;
      const int k = k0 + kk;
      float v = 0.f;
      if (k < K) {
        const int tap = k / p.Cin;
        int c = k - tap * p.Cin;
        const int dy = tap / p.KW - ph, dx = tap - (tap / p.KW) * p.KW - pw;
        const int yy = py[pi] + dy, xx = px[pi] + dx;
        if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) {
          int s = 0;
          while (s < p.num_src - 1 && c >= p.src_ch[s]) { c -= p.src_ch[s]; ++s; }
          if (NCHW_IN)
            v = __ldg(p.src_ptr[s] + ((long long)pb[pi] * p.src_ch[s] + c) * HW + (long long)yy * p.W + xx);
          else
            v = __ldg(p.src_ptr[s] + (pbase[pi] + (long long)dy * p.W + dx) * p.src_pitch[s] + c);
        }
      }
      sA[kk * VT::kSA + mi] = v;
    }
    // ---- B tile: rows k0..k0+15 of the packed [K][Cout] weight matrix
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + 256 * i;
      const int nn = e & 63, kk = e >> 6;
      const int k = k0 + kk, n = n0 + nn;
      sB[kk * VT::kSB + nn] = (k < K && n < p.Cout) ? __ldg(p.weight + (long long)k * p.Cout + n) : 0.f;
    }
    __syncthreads();
    VT::mac(sA, sB, ty, tx, acc);
    __syncthreads();
  }

  // ---- fused epilogue
  const int nb = n0 + tx * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long n = m0 + ty * 8 + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = nb + j;
      if (co >= p.Cout) continue;
      float v = acc[i][j] + (p.bias ? __ldg(p.bias + co) : 0.f);
      if (p.epilogue == AS_EPI_GRU_ZR) {
        const int Hd = p.Cout >> 1;
        v = sigmoidf_(v + __ldg(p.ctx + n * p.ctx_pitch + co));
        if (co < Hd) {
          p.z[n * Hd + co] = v;                                            // update.py:37
        } else {
          const int ch = co - Hd;
          p.out[n * p.out_pitch + p.out_coff + ch] = v * __ldg(p.h + n * Hd + ch);   // r*h, update.py:38-39
          if (p.save) p.save[n * Hd + ch] = v;                                        // r, kept for the backward pass
        }
        continue;
      }
      if (p.epilogue == AS_EPI_GRU_Q) {
        const float q = tanhf(v + __ldg(p.ctx + n * p.ctx_pitch + co));    // update.py:39
        if (p.save) p.save[n * p.Cout + co] = q;
        const float zz = p.z[n * p.Cout + co];
        const float hh = __ldg(p.h + n * p.Cout + co);
        v = (1.0f - zz) * hh + zz * q;                                      // update.py:40
      } else if (p.epilogue == AS_EPI_BIAS_RELU) {
        v = fmaxf(v, 0.f);
      }
      if (p.out_layout == AS_LAYOUT_NHWC) {
        p.out[n * p.out_pitch + p.out_coff + co] = v;
      } else {
        const int b = (int)(n / HW);
        const long long rem = n - (long long)b * HW;
        p.out[((long long)b * p.Cout + co) * HW + rem] = v;
      }
    }
  }
}

__global__ void pack_weight_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int T,
                                   long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  // out index = (tap*Cin + c)*Cout + n
  const int n = (int)(idx % Cout);
  const long long r = idx / Cout;
  const int c = (int)(r % Cin);
  const int tap = (int)(r / Cin);
  out[idx] = w[((long long)n * Cin + c) * T + tap];
}

// F.avg_pool2d(x, 3, stride=2, padding=1), count_include_pad -> always /9 (update.py:94-95)
__global__ void pool2x_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W, int Ho, int Wo,
                                   int C4, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c4 = (int)(idx % C4);
  long long t = idx / C4;
  const int xo = (int)(t % Wo); t /= Wo;
  const int yo = (int)(t % Ho);
  const int b = (int)(t / Ho);
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
  reinterpret_cast<float4*>(out)[idx] = make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv);
}

// F.interpolate(mode='bilinear', align_corners=True) (update.py:100-102)
__global__ void interp_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int Hi, int Wi, int Ho,
                                   int Wo, int C4, float sy, float sx, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c4 = (int)(idx % C4);
  long long t = idx / C4;
  const int xo = (int)(t % Wo); t /= Wo;
  const int yo = (int)(t % Ho);
  const int b = (int)(t / Ho);
  const float fy = sy * yo, fx = sx * xo;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = min(y0 + 1, Hi - 1), x1 = min(x0 + 1, Wi - 1);
  const float ly = fy - y0, lx = fx - x0;
  const float hy = 1.0f - ly, hx = 1.0f - lx;
  const float4* base = reinterpret_cast<const float4*>(in) + (long long)b * Hi * Wi * C4 + c4;
  const float4 v00 = __ldg(base + ((long long)y0 * Wi + x0) * C4);
  const float4 v01 = __ldg(base + ((long long)y0 * Wi + x1) * C4);
  const float4 v10 = __ldg(base + ((long long)y1 * Wi + x0) * C4);
  const float4 v11 = __ldg(base + ((long long)y1 * Wi + x1) * C4);
  float4 o;
  o.x = hy * (hx * v00.x + lx * v01.x) + ly * (hx * v10.x + lx * v11.x);
  o.y = hy * (hx * v00.y + lx * v01.y) + ly * (hx * v10.y + lx * v11.y);
  o.z = hy * (hx * v00.z + lx * v01.z) + ly * (hx * v10.z + lx * v11.z);
  o.w = hy * (hx * v00.w + lx * v01.w) + ly * (hx * v10.w + lx * v11.w);
  reinterpret_cast<float4*>(out)[idx] = o;
}

// [B,C,HW] -> [B*HW][pitch] (+coff): 32x32 smem transpose tiles
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int C,
                                                           long long HW, int pitch, int coff,
                                                           const float* __restrict__ bias = nullptr) {
  __shared__ float t[32][33];
  const int b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;   // 8 rows per pass
#pragma unroll
  for (int r = ly; r < 32; r += 8) {
    const int c = c0 + r;
    const long long pp = p0 + lx;
    t[r][lx] = (c < C && pp < HW) ? __ldg(in + ((long long)b * C + c) * HW + pp) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = ly; r < 32; r += 8) {
    const long long pp = p0 + r;
    const int c = c0 + lx;
    if (c < C && pp < HW) out[((long long)b * HW + pp) * pitch + coff + c] = t[lx][r] + (bias ? __ldg(bias + c) : 0.f);
  }
}

__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int C,
                                                           long long HW, int pitch, int coff) {
  __shared__ float t[32][33];
  const int b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
#pragma unroll
  for (int r = ly; r < 32; r += 8) {
    const long long pp = p0 + r;
    const int c = c0 + lx;
    t[r][lx] = (c < C && pp < HW) ? __ldg(in + ((long long)b * HW + pp) * pitch + coff + c) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = ly; r < 32; r += 8) {
    const int c = c0 + r;
    const long long pp = p0 + lx;
    if (c < C && pp < HW) out[((long long)b * C + c) * HW + pp] = t[lx][r];
  }
}

__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y,
                           long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a[i] + b[i];
}

}  // namespace

extern "C" int as_conv2d_fp32(const as_conv_desc* d, as_stream_t stream) {
  if (!d || !d->weight || !d->out) return AS_ERR_BAD_ARG;
  if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->Cout <= 0 || d->num_src < 1 || d->num_src > AS_MAX_SRC) return AS_ERR_BAD_ARG;
  if (d->KH < 1 || d->KW < 1 || !(d->KH & 1) || !(d->KW & 1)) return AS_ERR_UNSUPPORTED;
  ConvParams p{};
  p.B = d->B; p.H = d->H; p.W = d->W; p.KH = d->KH; p.KW = d->KW; p.Cout = d->Cout;
  p.num_src = d->num_src;
  int cin = 0;
  const int layout = d->src[0].layout;
  for (int s = 0; s < d->num_src; ++s) {
    if (!d->src[s].ptr || d->src[s].channels <= 0) return AS_ERR_BAD_ARG;
    if (d->src[s].layout != layout) return AS_ERR_UNSUPPORTED;
    if (layout == AS_LAYOUT_NHWC && d->src[s].pitch < d->src[s].channels) return AS_ERR_BAD_ARG;
    p.src_ptr[s] = d->src[s].ptr; p.src_ch[s] = d->src[s].channels; p.src_pitch[s] = d->src[s].pitch;
    cin += d->src[s].channels;
  }
  p.Cin = cin;
  p.weight = d->weight; p.bias = d->bias; p.epilogue = d->epilogue;
  p.out = d->out; p.out_pitch = d->out_pitch; p.out_coff = d->out_coff; p.out_layout = d->out_layout;
  p.ctx = d->ctx; p.ctx_pitch = d->ctx_pitch; p.h = d->h; p.z = d->z; p.save = d->save;
  if (d->epilogue == AS_EPI_GRU_ZR || d->epilogue == AS_EPI_GRU_Q) {
    if (!d->ctx || !d->h || !d->z) return AS_ERR_BAD_ARG;
    if (d->out_layout != AS_LAYOUT_NHWC) return AS_ERR_UNSUPPORTED;
    if (d->epilogue == AS_EPI_GRU_ZR && (d->Cout & 1)) return AS_ERR_BAD_ARG;
  } else if (d->epilogue != AS_EPI_BIAS && d->epilogue != AS_EPI_BIAS_RELU) {
    return AS_ERR_UNSUPPORTED;
  }
  if (d->out_layout == AS_LAYOUT_NHWC && d->out_pitch < d->out_coff + (d->epilogue == AS_EPI_GRU_ZR ? d->Cout / 2 : d->Cout))
    return AS_ERR_BAD_ARG;
  const long long N = (long long)d->B * d->H * d->W;
  dim3 grid((unsigned)as_ceil_div_ll(N, 128), as_ceil_div(d->Cout, 64));
  if (layout == AS_LAYOUT_NCHW)
    conv_simt_kernel<true><<<grid, VT::kThreads, 0, as_cu(stream)>>>(p);
  else
    conv_simt_kernel<false><<<grid, VT::kThreads, 0, as_cu(stream)>>>(p);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_pack_conv_weight(const float* w_oihw, float* w_packed, int Cout, int Cin, int KH, int KW,
                                   as_stream_t stream) {
  if (!w_oihw || !w_packed || Cout <= 0 || Cin <= 0 || KH <= 0 || KW <= 0) return AS_ERR_BAD_ARG;
  const long long total = (long long)Cout * Cin * KH * KW;
  pack_weight_kernel<<<(unsigned)as_ceil_div_ll(total, 256), 256, 0, as_cu(stream)>>>(w_oihw, w_packed, Cout, Cin,
                                                                                  KH * KW, total);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_pool2x_nhwc(const float* in, float* out, int B, int H, int W, int C, as_stream_t stream) {
  if (!in || !out || B <= 0 || H <= 0 || W <= 0 || C <= 0) return AS_ERR_BAD_ARG;
  if ((C & 3) || !as_aligned16(in) || !as_aligned16(out)) return AS_ERR_ALIGNMENT;
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const long long total = (long long)B * Ho * Wo * (C / 4);
  pool2x_nhwc_kernel<<<(unsigned)as_ceil_div_ll(total, 256), 256, 0, as_cu(stream)>>>(in, out, H, W, Ho, Wo, C / 4, total);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_interp_bilinear_nhwc(const float* in, float* out, int B, int Hin, int Win, int Hout, int Wout,
                                       int C, as_stream_t stream) {
  if (!in || !out || B <= 0 || Hin <= 0 || Win <= 0 || Hout <= 0 || Wout <= 0 || C <= 0) return AS_ERR_BAD_ARG;
  if ((C & 3) || !as_aligned16(in) || !as_aligned16(out)) return AS_ERR_ALIGNMENT;
  // align_corners=True scale: (in-1)/(out-1), 0 when out == 1 (ATen area_pixel_compute_scale)
  const float sy = Hout > 1 ? (float)(Hin - 1) / (float)(Hout - 1) : 0.f;
  const float sx = Wout > 1 ? (float)(Win - 1) / (float)(Wout - 1) : 0.f;
  const long long total = (long long)B * Hout * Wout * (C / 4);
  interp_nhwc_kernel<<<(unsigned)as_ceil_div_ll(total, 256), 256, 0, as_cu(stream)>>>(in, out, Hin, Win, Hout, Wout,
                                                                                  C / 4, sy, sx, total);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_nchw_to_nhwc(const float* in, float* out, int B, int C, int H, int W, int out_pitch, int out_coff,
                               as_stream_t stream) {
  if (!in || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0 || out_pitch < out_coff + C) return AS_ERR_BAD_ARG;
  if (B > 65535 || as_ceil_div(C, 32) > 65535) return AS_ERR_UNSUPPORTED;
  const long long HW = (long long)H * W;
  dim3 grid((unsigned)as_ceil_div_ll(HW, 32), as_ceil_div(C, 32), B);
  nchw_to_nhwc_kernel<<<grid, 256, 0, as_cu(stream)>>>(in, out, C, HW, out_pitch, out_coff);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_nchw_to_nhwc_bias(const float* in, const float* bias, float* out, int B, int C, int H, int W,
                                    int out_pitch, int out_coff, as_stream_t stream) {
  if (!in || !bias || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0 || out_pitch < out_coff + C) return AS_ERR_BAD_ARG;
  if (B > 65535 || as_ceil_div(C, 32) > 65535) return AS_ERR_UNSUPPORTED;
  const long long HW = (long long)H * W;
  dim3 grid((unsigned)as_ceil_div_ll(HW, 32), as_ceil_div(C, 32), B);
  nchw_to_nhwc_kernel<<<grid, 256, 0, as_cu(stream)>>>(in, out, C, HW, out_pitch, out_coff, bias);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_nhwc_to_nchw(const float* in, float* out, int B, int C, int H, int W, int in_pitch, int in_coff,
                               as_stream_t stream) {
  if (!in || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0 || in_pitch < in_coff + C) return AS_ERR_BAD_ARG;
  if (B > 65535 || as_ceil_div(C, 32) > 65535) return AS_ERR_UNSUPPORTED;
  const long long HW = (long long)H * W;
  dim3 grid((unsigned)as_ceil_div_ll(HW, 32), as_ceil_div(C, 32), B);
  nhwc_to_nchw_kernel<<<grid, 256, 0, as_cu(stream)>>>(in, out, C, HW, in_pitch, in_coff);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_add_f32(const float* a, const float* b, float* y, long long n, as_stream_t stream) {
  if (!a || !b || !y || n <= 0) return AS_ERR_BAD_ARG;
  add_kernel<<<(unsigned)as_ceil_div_ll(n, 256), 256, 0, as_cu(stream)>>>(a, b, y, n);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}
