// a13-vi: adjoints of the per-iteration update block (training, BASELINE.json config 5), exact fp32 on CUDA
// cores.  The reference gets these from autograd over models/*/update.py:16-136 (cuDNN dgrad/wgrad + ATen
// elementwise); here:
//   * data gradient  = the forward implicit-GEMM kernel run on dY with flipped/transposed weights
//                      (as_pack_conv_weight_dgrad + as_conv2d_fp32),
//   * weight gradient = implicit GEMM with K = pixels (this file), split over pixel chunks, fp32 atomics,
//   * gate / relu / pool2x / bilinear adjoints = small fused elementwise kernels.
#include "simt_gemm.cuh"

namespace {

using WT = SimtTile<64, 64, 16, 4, 4>;   // 64 (Cout) x 64 (Cin) tile, K = 16 pixels per step
constexpr int kPixChunk = 2048;

struct WgradParams {
  int B, H, W, KH, KW, Cout, Cin;
  int num_src;
  const float* src_ptr[AS_MAX_SRC];
  int src_ch[AS_MAX_SRC];
  int src_pitch[AS_MAX_SRC];
  const float* dy; int dy_pitch;
  float* dw; float* db;
};

// grid: x = pixel chunk, y = tap * co_tiles * ci_tiles
__global__ void __launch_bounds__(WT::kThreads) wgrad_kernel(WgradParams p) {
  __shared__ __align__(16) float smem[WT::kSmemFloats];
  float* sA = smem;                      // [16][64+4]  dY  (k = pixel, m = co)
  float* sB = smem + 16 * WT::kSA;       // [16][64+4]  X   (k = pixel, n = ci)
  const int tid = threadIdx.x;
  const int ty = tid / 16, tx = tid % 16;
  const int co_tiles = as_ceil_div(p.Cout, 64), ci_tiles = as_ceil_div(p.Cin, 64);
  int r = blockIdx.y;
  const int cit = r % ci_tiles; r /= ci_tiles;
  const int cot = r % co_tiles;
  const int tap = r / co_tiles;
  const int co0 = cot * 64, ci0 = cit * 64;
  const int dyk = tap / p.KW - p.KH / 2, dxk = tap % p.KW - p.KW / 2;
  const long long HW = (long long)p.H * p.W;
  const long long N = (long long)p.B * HW;
  const long long n_begin = (long long)blockIdx.x * kPixChunk;
  const long long n_end = n_begin + kPixChunk < N ? n_begin + kPixChunk : N;

  // each thread loads 4 elements of each tile per step: pixel kk = tid/16, channels (tid%16)*4..+3
  const int kk = tid / 16, c4 = (tid % 16) * 4;
  // resolve the source tensor that holds input channels ci0+c4.. (chunks of 4 never straddle sources when
  // every source has a multiple of 4 channels; otherwise fall back to per-element lookup)
  float acc[4][4] = {};
  for (long long n0 = n_begin; n0 < n_end; n0 += 16) {
    const long long n = n0 + kk;
    float a[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (n < n_end) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (co0 + c4 + j < p.Cout) a[j] = __ldg(p.dy + n * p.dy_pitch + co0 + c4 + j);
      const int b = (int)(n / HW);
      const int rem = (int)(n - (long long)b * HW);
      const int y = rem / p.W + dyk, x = rem % p.W + dxk;
      if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
        const long long ns = n + (long long)dyk * p.W + dxk;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int c = ci0 + c4 + j;
          if (c < p.Cin) {
            int s = 0;
            while (s < p.num_src - 1 && c >= p.src_ch[s]) { c -= p.src_ch[s]; ++s; }
            bv[j] = __ldg(p.src_ptr[s] + ns * p.src_pitch[s] + c);
          }
        }
      }
    }
    *reinterpret_cast<float4*>(sA + kk * WT::kSA + c4) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4*>(sB + kk * WT::kSB + c4) = make_float4(bv[0], bv[1], bv[2], bv[3]);
    __syncthreads();
    WT::mac(sA, sB, ty, tx, acc);
    __syncthreads();
  }
  const int T = p.KH * p.KW;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= p.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + tx * 4 + j;
      if (ci < p.Cin) atomicAdd(p.dw + ((long long)co * p.Cin + ci) * T + tap, acc[i][j]);
    }
  }
}

// db[c] += sum_n dy[n][c]; one block = kBiasRows pixels x 32 channels (1024 rows: ~900 blocks at config 5, all resident)
constexpr int kBiasRows = 1024;
__global__ void __launch_bounds__(256) bias_grad_kernel(const float* __restrict__ dy, int pitch, int C, long long N,
                                                        float* __restrict__ db) {
  const int c = blockIdx.y * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;                      // 8 row lanes
  const long long n_begin = (long long)blockIdx.x * kBiasRows;
  const long long n_end = n_begin + kBiasRows < N ? n_begin + kBiasRows : N;
  float acc = 0.f;
  if (c < C)
    for (long long n = n_begin + rl; n < n_end; n += 8) acc += __ldg(dy + n * pitch + c);
  __shared__ float red[8][33];
  red[rl][threadIdx.x & 31] = acc;
  __syncthreads();
  if (rl == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x & 31];
    atomicAdd(db + c, s);
  }
}

__global__ void pack_weight_dgrad_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int T,
                                         long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  // out index = (tapf*Cout + co)*Cin + ci  : a [K = T*Cout][N = Cin] GEMM operand for conv(dY)
  const int ci = (int)(idx % Cin);
  const long long r = idx / Cin;
  const int co = (int)(r % Cout);
  const int tapf = (int)(r / Cout);
  out[idx] = w[((long long)co * Cin + ci) * T + (T - 1 - tapf)];
}

__global__ void relu_bwd_kernel(const float* __restrict__ dy, int dyp, int dyo, const float* __restrict__ y, int yp, int yo,
                                float* __restrict__ dx, int dxp, int dxo, int C, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long n = idx / C;
  const int c = (int)(idx - n * C);
  const float g = dy[n * dyp + dyo + c];
  dx[n * dxp + dxo + c] = y[n * yp + yo + c] > 0.f ? g : 0.f;
}

__global__ void gru_gates1_kernel(const float* __restrict__ dhn, const float* __restrict__ z, const float* __restrict__ q,
                                  const float* __restrict__ h, float* __restrict__ dq, float* __restrict__ dzr,
                                  float* __restrict__ dh, int Hd, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long n = idx / Hd;
  const int c = (int)(idx - n * Hd);
  const float g = dhn[idx], zz = z[idx], qq = q[idx], hh = h[idx];
  dq[idx] = g * zz * (1.0f - qq * qq);                       // through tanh
  dzr[n * 2 * Hd + c] = g * (qq - hh) * zz * (1.0f - zz);    // through sigmoid(z)
  dh[idx] += g * (1.0f - zz);
}

__global__ void gru_gates2_kernel(const float* __restrict__ drh, int pitch, const float* __restrict__ h,
                                  const float* __restrict__ r, float* __restrict__ dzr, float* __restrict__ dh, int Hd,
                                  long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long n = idx / Hd;
  const int c = (int)(idx - n * Hd);
  const float g = drh[n * pitch + c], hh = h[idx], rr = r[idx];
  dzr[n * 2 * Hd + Hd + c] = g * hh * rr * (1.0f - rr);      // through sigmoid(r)
  dh[idx] += g * rr;
}

__global__ void add_slice_kernel(const float* __restrict__ src, int sp, int so, float* __restrict__ dst, int dp, int doff,
                                 int C, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long n = idx / C;
  const int c = (int)(idx - n * C);
  dst[n * dp + doff + c] += src[n * sp + so + c];
}

// adjoint of avg_pool2d(3, stride 2, pad 1): gather form, one thread per input element
__global__ void pool2x_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int H, int W, int Ho, int Wo, int C,
                                  long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  long long t = idx / C;
  const int x = (int)(t % W); t /= W;
  const int y = (int)(t % H);
  const int b = (int)(t / H);
  float acc = 0.f;
  // outputs (yo,xo) whose 3x3 window [2yo-1, 2yo+1] contains y: yo in {ceil((y-1)/2) .. floor((y+1)/2)}
  for (int yo = (y) / 2; yo <= (y + 1) / 2; ++yo) {
    if (yo < 0 || yo >= Ho || y < 2 * yo - 1 || y > 2 * yo + 1) continue;
    for (int xo = (x) / 2; xo <= (x + 1) / 2; ++xo) {
      if (xo < 0 || xo >= Wo || x < 2 * xo - 1 || x > 2 * xo + 1) continue;
      acc += dy[(((long long)b * Ho + yo) * Wo + xo) * C + c];
    }
  }
  dx[idx] += acc * (1.0f / 9.0f);
}

// adjoint of bilinear align_corners=True resize: scatter with fp32 atomics
__global__ void interp_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int Hi, int Wi, int Ho, int Wo, int C,
                                  float sy, float sx, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  long long t = idx / C;
  const int xo = (int)(t % Wo); t /= Wo;
  const int yo = (int)(t % Ho);
  const int b = (int)(t / Ho);
  const float fy = sy * yo, fx = sx * xo;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = min(y0 + 1, Hi - 1), x1 = min(x0 + 1, Wi - 1);
  const float ly = fy - y0, lx = fx - x0, hy = 1.0f - ly, hx = 1.0f - lx;
  const float g = dy[idx];
  float* base = dx + (long long)b * Hi * Wi * C + c;
  atomicAdd(base + ((long long)y0 * Wi + x0) * C, g * hy * hx);
  atomicAdd(base + ((long long)y0 * Wi + x1) * C, g * hy * lx);
  atomicAdd(base + ((long long)y1 * Wi + x0) * C, g * ly * hx);
  atomicAdd(base + ((long long)y1 * Wi + x1) * C, g * ly * lx);
}

// the epilogues of conv_simt_kernel on a raw (acc + bias) convolution output produced by the tensor-core kernel
__device__ __forceinline__ float sigmoid_(float x) { return 1.0f / (1.0f + expf(-x)); }
__global__ void conv_epilogue_kernel(const float* __restrict__ raw, int raw_pitch, int Cout, int epilogue,
                                     const float* __restrict__ ctx, int ctx_pitch, const float* __restrict__ h,
                                     float* __restrict__ z, float* __restrict__ save, float* __restrict__ out, int out_pitch,
                                     int out_coff, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long n = idx / Cout;
  const int co = (int)(idx - n * Cout);
  float v = raw[n * raw_pitch + co];
  if (epilogue == AS_EPI_GRU_ZR) {
    const int Hd = Cout >> 1;
    v = sigmoid_(v + __ldg(ctx + n * ctx_pitch + co));
    if (co < Hd) {
      z[n * Hd + co] = v;                                                  // update.py:37
    } else {
      const int ch = co - Hd;
      out[n * out_pitch + out_coff + ch] = v * __ldg(h + n * Hd + ch);      // r*h, update.py:38-39
      if (save) save[n * Hd + ch] = v;
    }
    return;
  }
  if (epilogue == AS_EPI_GRU_Q) {
    const float q = tanhf(v + __ldg(ctx + n * ctx_pitch + co));            // update.py:39
    if (save) save[n * Cout + co] = q;
    const float zz = z[n * Cout + co];
    const float hh = __ldg(h + n * Cout + co);
    v = (1.0f - zz) * hh + zz * q;                                          // update.py:40
  } else if (epilogue == AS_EPI_BIAS_RELU) {
    v = fmaxf(v, 0.f);
  }
  out[n * out_pitch + out_coff + co] = v;
}

// weight gradient of BasicMotionEncoder.convd1 (7x7, 1 -> 64, update.py:80): dW[co][t] += sum_px dY[px][co] * disp[px + t].
// The generic wgrad kernel wastes 63/64 of its 64x64 tile on the single input channel; here a block owns a
// 32x8 pixel tile, thread = (output channel, group of 13 taps), the disparity patch and one dY row live in smem.
__global__ void __launch_bounds__(256) convd1_wgrad_kernel(const float* __restrict__ disp, const float* __restrict__ dy,
                                                           int dy_pitch, int H, int W, float* __restrict__ dw) {
  __shared__ float patch[8 + 6][32 + 6];
  __shared__ float sdy[32][64];
  const int tid = threadIdx.x;
  const int b = blockIdx.z, x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
  const long long HW = (long long)H * W;
  for (int i = tid; i < 14 * 38; i += 256) {
    const int r = i / 38, c = i - r * 38;
    const int yy = y0 + r - 3, xx = x0 + c - 3;
    patch[r][c] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(disp + (long long)b * HW + (long long)yy * W + xx) : 0.f;
  }
  const int co = tid & 63, t0 = (tid >> 6) * 13;          // taps t0 .. min(49, t0 + 13)
  int ky[13], kx[13];
#pragma unroll
  for (int j = 0; j < 13; ++j) {
    const int t = min(t0 + j, 48);
    ky[j] = t / 7;
    kx[j] = t - ky[j] * 7;
  }
  float acc[13];
#pragma unroll
  for (int j = 0; j < 13; ++j) acc[j] = 0.f;
  for (int ty = 0; ty < 8; ++ty) {
    __syncthreads();                                       // patch ready (first pass) / previous row consumed
    const int y = y0 + ty;
    for (int i = tid; i < 32 * 64; i += 256) {
      const int px = i >> 6, c = i & 63, x = x0 + px;
      sdy[px][c] = (y < H && x < W) ? __ldg(dy + ((long long)b * HW + (long long)y * W + x) * dy_pitch + c) : 0.f;
    }
    __syncthreads();
    for (int px = 0; px < 32; ++px) {
      const float g = sdy[px][co];
#pragma unroll
      for (int j = 0; j < 13; ++j) acc[j] = fmaf(g, patch[ty + ky[j]][px + kx[j]], acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 13; ++j)
    if (t0 + j < 49) atomicAdd(dw + co * 49 + t0 + j, acc[j]);
}

inline unsigned blocks_for(long long total) { return (unsigned)as_ceil_div_ll(total, 256); }

}  // namespace

extern "C" int as_pack_conv_weight_dgrad(const float* w_oihw, float* w_packed, int Cout, int Cin, int KH, int KW,
                                         as_stream_t stream) {
  if (!w_oihw || !w_packed || Cout <= 0 || Cin <= 0 || KH <= 0 || KW <= 0) return AS_ERR_BAD_ARG;
  const long long total = (long long)Cout * Cin * KH * KW;
  pack_weight_dgrad_kernel<<<blocks_for(total), 256, 0, as_cu(stream)>>>(w_oihw, w_packed, Cout, Cin, KH * KW, total);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_conv2d_wgrad_fp32(const as_conv_desc* d, const float* dy, int dy_pitch, int Cout, float* dw_acc,
                                    float* db_acc, as_stream_t stream) {
  if (!d || !dy || !dw_acc || Cout <= 0 || dy_pitch < Cout) return AS_ERR_BAD_ARG;
  if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->num_src < 1 || d->num_src > AS_MAX_SRC) return AS_ERR_BAD_ARG;
  if (d->KH < 1 || d->KW < 1 || !(d->KH & 1) || !(d->KW & 1)) return AS_ERR_UNSUPPORTED;
  WgradParams p{};
  p.B = d->B; p.H = d->H; p.W = d->W; p.KH = d->KH; p.KW = d->KW; p.Cout = Cout;
  p.num_src = d->num_src;
  int cin = 0;
  for (int s = 0; s < d->num_src; ++s) {
    if (!d->src[s].ptr || d->src[s].channels <= 0 || d->src[s].layout != AS_LAYOUT_NHWC) return AS_ERR_UNSUPPORTED;
    p.src_ptr[s] = d->src[s].ptr; p.src_ch[s] = d->src[s].channels; p.src_pitch[s] = d->src[s].pitch;
    cin += d->src[s].channels;
  }
  p.Cin = cin; p.dy = dy; p.dy_pitch = dy_pitch; p.dw = dw_acc; p.db = db_acc;
  const long long N = (long long)d->B * d->H * d->W;
  const int tiles = d->KH * d->KW * as_ceil_div(Cout, 64) * as_ceil_div(cin, 64);
  if (tiles > 65535) return AS_ERR_UNSUPPORTED;
  dim3 grid((unsigned)as_ceil_div_ll(N, kPixChunk), tiles);
  wgrad_kernel<<<grid, WT::kThreads, 0, as_cu(stream)>>>(p);
  AS_RETURN_IF_LAUNCH_FAILED();
  if (db_acc) {
    dim3 g2((unsigned)as_ceil_div_ll(N, kBiasRows), as_ceil_div(Cout, 32));
    bias_grad_kernel<<<g2, 256, 0, as_cu(stream)>>>(dy, dy_pitch, Cout, N, db_acc);
    AS_RETURN_IF_LAUNCH_FAILED();
  }
  return AS_OK;
}

extern "C" int as_convd1_wgrad_fp32(const float* disp, const float* dy, int dy_pitch, int B, int H, int W, float* dw_acc,
                                    as_stream_t stream) {
  if (!disp || !dy || !dw_acc || B <= 0 || H <= 0 || W <= 0 || dy_pitch < 64) return AS_ERR_BAD_ARG;
  if (B > 65535 || as_ceil_div(H, 8) > 65535) return AS_ERR_UNSUPPORTED;
  dim3 grid(as_ceil_div(W, 32), as_ceil_div(H, 8), B);
  convd1_wgrad_kernel<<<grid, 256, 0, as_cu(stream)>>>(disp, dy, dy_pitch, H, W, dw_acc);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_bias_grad_fp32(const float* dy, int dy_pitch, int Cout, long long N, float* db_acc, as_stream_t stream) {
  if (!dy || !db_acc || Cout <= 0 || N <= 0 || dy_pitch < Cout) return AS_ERR_BAD_ARG;
  dim3 g2((unsigned)as_ceil_div_ll(N, kBiasRows), as_ceil_div(Cout, 32));
  bias_grad_kernel<<<g2, 256, 0, as_cu(stream)>>>(dy, dy_pitch, Cout, N, db_acc);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_relu_bwd(const float* dy, int dy_pitch, int dy_coff, const float* y, int y_pitch, int y_coff, float* dx,
                           int dx_pitch, int dx_coff, long long N, int C, as_stream_t stream) {
  if (!dy || !y || !dx || N <= 0 || C <= 0) return AS_ERR_BAD_ARG;
  relu_bwd_kernel<<<blocks_for(N * C), 256, 0, as_cu(stream)>>>(dy, dy_pitch, dy_coff, y, y_pitch, y_coff, dx, dx_pitch,
                                                                dx_coff, C, N * C);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_gru_bwd_gates1(const float* dhn, const float* z, const float* q, const float* h, float* dq_pre,
                                 float* dzr_pre, float* dh_acc, long long N, int Hd, as_stream_t stream) {
  if (!dhn || !z || !q || !h || !dq_pre || !dzr_pre || !dh_acc || N <= 0 || Hd <= 0) return AS_ERR_BAD_ARG;
  gru_gates1_kernel<<<blocks_for(N * Hd), 256, 0, as_cu(stream)>>>(dhn, z, q, h, dq_pre, dzr_pre, dh_acc, Hd, N * Hd);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_gru_bwd_gates2(const float* drh, int drh_pitch, const float* h, const float* r, float* dzr_pre,
                                 float* dh_acc, long long N, int Hd, as_stream_t stream) {
  if (!drh || !h || !r || !dzr_pre || !dh_acc || N <= 0 || Hd <= 0 || drh_pitch < Hd) return AS_ERR_BAD_ARG;
  gru_gates2_kernel<<<blocks_for(N * Hd), 256, 0, as_cu(stream)>>>(drh, drh_pitch, h, r, dzr_pre, dh_acc, Hd, N * Hd);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_conv_epilogue_fp32(const float* raw, int raw_pitch, long long N, int Cout, int epilogue, const float* ctx,
                                     int ctx_pitch, const float* h, float* z, float* save, float* out, int out_pitch,
                                     int out_coff, as_stream_t stream) {
  if (!raw || !out || N <= 0 || Cout <= 0 || raw_pitch < Cout) return AS_ERR_BAD_ARG;
  if (epilogue == AS_EPI_GRU_ZR) {
    if (!ctx || !h || !z || (Cout & 1) || ctx_pitch < Cout || out_pitch < out_coff + Cout / 2) return AS_ERR_BAD_ARG;
  } else if (epilogue == AS_EPI_GRU_Q) {
    if (!ctx || !h || !z || ctx_pitch < Cout || out_pitch < out_coff + Cout) return AS_ERR_BAD_ARG;
  } else if (epilogue == AS_EPI_BIAS || epilogue == AS_EPI_BIAS_RELU) {
    if (out_pitch < out_coff + Cout) return AS_ERR_BAD_ARG;
  } else {
    return AS_ERR_UNSUPPORTED;
  }
  conv_epilogue_kernel<<<blocks_for(N * Cout), 256, 0, as_cu(stream)>>>(raw, raw_pitch, Cout, epilogue, ctx, ctx_pitch, h, z,
                                                                       save, out, out_pitch, out_coff, N * Cout);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_add_slice(const float* src, int spitch, int scoff, float* dst, int dpitch, int dcoff, long long N, int C,
                            as_stream_t stream) {
  if (!src || !dst || N <= 0 || C <= 0) return AS_ERR_BAD_ARG;
  add_slice_kernel<<<blocks_for(N * C), 256, 0, as_cu(stream)>>>(src, spitch, scoff, dst, dpitch, dcoff, C, N * C);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_pool2x_nhwc_bwd(const float* dy, float* dx_acc, int B, int H, int W, int C, as_stream_t stream) {
  if (!dy || !dx_acc || B <= 0 || H <= 0 || W <= 0 || C <= 0) return AS_ERR_BAD_ARG;
  const long long total = (long long)B * H * W * C;
  pool2x_bwd_kernel<<<blocks_for(total), 256, 0, as_cu(stream)>>>(dy, dx_acc, H, W, (H + 1) / 2, (W + 1) / 2, C, total);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_interp_bilinear_nhwc_bwd(const float* dy, float* dx_acc, int B, int Hin, int Win, int Hout, int Wout,
                                           int C, as_stream_t stream) {
  if (!dy || !dx_acc || B <= 0 || Hin <= 0 || Win <= 0 || Hout <= 0 || Wout <= 0 || C <= 0) return AS_ERR_BAD_ARG;
  const float sy = Hout > 1 ? (float)(Hin - 1) / (float)(Hout - 1) : 0.f;
  const float sx = Wout > 1 ? (float)(Win - 1) / (float)(Wout - 1) : 0.f;
  const long long total = (long long)B * Hout * Wout * C;
  interp_bwd_kernel<<<blocks_for(total), 256, 0, as_cu(stream)>>>(dy, dx_acc, Hin, Win, Hout, Wout, C, sy, sx, total);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}
