// a6: corr_sampler.forward / backward for sm_100a.
//
// Replaces sampler/sampler_kernel.cu:19-166 of the reference.  Differences in HOW (not WHAT):
//  * forward: a CTA stages the (2r+2)-tap windows of 128 consecutive pixels in shared memory with
//    cooperative loads (consecutive lanes read consecutive taps of the same pixel), then each thread
//    interpolates its pixel from smem and writes the 2r+1 output planes coalesced.  No zero-filled
//    output + global read-modify-write as in the reference (sampler_kernel.cu:52-56,123).
//  * backward: one warp owns one volume row and writes it completely (window values, zeros elsewhere),
//    so the full-volume memset of sampler_kernel.cu:148 is fused away.
//  * launches on the caller's stream and reports launch errors.
#include "common.cuh"

namespace {

template <typename T> struct Acc { using type = float; };
template <> struct Acc<double> { using type = double; };

template <typename T> __device__ __forceinline__ typename Acc<T>::type to_acc(T v) { return (typename Acc<T>::type)v; }
template <> __device__ __forceinline__ float to_acc<__half>(__half v) { return __half2float(v); }
template <typename T, typename A> __device__ __forceinline__ T from_acc(A v) { return (T)v; }
template <> __device__ __forceinline__ __half from_acc<__half, float>(float v) { return __float2half_rn(v); }

constexpr int kPix = 128;  // pixels per CTA

// RT > 0: compile-time radius (the models use 4): the staging loop unrolls into independent loads with constant
// index arithmetic; RT = 0: runtime radius
template <typename T, int RT>
__global__ void __launch_bounds__(kPix) sampler_fwd_kernel(const T* __restrict__ vol,
                                                           const float* __restrict__ coords,
                                                           long long coords_bstride, T* __restrict__ out,
                                                           int HW, int W2, int r_) {
  const int r = RT > 0 ? RT : r_;
  using A = typename Acc<T>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  A* s_win = reinterpret_cast<A*>(smem_raw);  // [kPix][ntap+1]
  __shared__ int s_t0[kPix];
  __shared__ float s_f[kPix];

  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * kPix;
  const int ntap = 2 * r + 2;
  const int stride = ntap + 1;  // odd -> conflict-free when lane = pixel

  {
    const int p = p0 + tid;
    float x0 = 0.f;
    if (p < HW) x0 = coords[(long long)b * coords_bstride + p];
    const float fl = floorf(x0);
    s_t0[tid] = (int)fl - r;      // sampler_kernel.cu:47
    s_f[tid] = x0 - fl;           // sampler_kernel.cu:42
  }
  __syncthreads();

  const T* vrow = vol + (long long)b * HW * W2;
#pragma unroll
  for (int flat = tid; flat < kPix * ntap; flat += kPix) {
    const int pix = flat / ntap;
    const int j = flat - pix * ntap;
    const int p = p0 + pix;
    A v = (A)0;
    if (p < HW) {
      const int x1 = s_t0[pix] + j;
      if (x1 >= 0 && x1 < W2) v = to_acc<T>(vrow[(long long)p * W2 + x1]);  // sampler_kernel.cu:49-50
    }
    s_win[pix * stride + j] = v;
  }
  __syncthreads();

  const int p = p0 + tid;
  if (p >= HW) return;
  const A f = (A)s_f[tid];
  const A omf = (A)(1.0f - s_f[tid]);   // scalar_t(1.0f - dx), sampler_kernel.cu:56
  const A* w = s_win + tid * stride;
  T* o = out + (long long)b * (2 * r + 1) * HW + p;
  A prev = w[0];
#pragma unroll
  for (int k = 0; k < 2 * r + 1; ++k) {
    const A cur = w[k + 1];
    o[(long long)k * HW] = from_acc<T, A>(prev * omf + cur * f);
    prev = cur;
  }
}

// one warp per volume row (pixel); 8 rows per CTA
template <typename T>
__global__ void __launch_bounds__(256) sampler_bwd_kernel(const float* __restrict__ coords,
                                                          long long coords_bstride,
                                                          const T* __restrict__ gout, T* __restrict__ gvol,
                                                          int HW, int W2, int r, long long total_rows) {
  using A = typename Acc<T>::type;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= total_rows) return;
  const int b = (int)(row / HW);
  const int p = (int)(row - (long long)b * HW);
  const float x0 = coords[(long long)b * coords_bstride + p];
  const float fl = floorf(x0);
  const A f = (A)(x0 - fl);
  const A omf = (A)(1.0f - (x0 - fl));
  const int t0 = (int)fl - r;
  const int K = 2 * r + 1;
  // lanes 0..K-1 hold g[k]; K <= 32 is enforced on the host
  A g = (A)0;
  if (lane < K) g = to_acc<T>(gout[((long long)b * K + lane) * HW + p]);
  T* dst = gvol + row * W2;
  for (int base = 0; base < W2; base += 32) {  // warp-uniform trip count (shuffles inside)
    const int x1 = base + lane;
    const int j = x1 - t0;  // tap index in [0, 2r+1] when inside the window
    const A gm1 = __shfl_sync(0xffffffffu, g, (j - 1) & 31);
    const A g0 = __shfl_sync(0xffffffffu, g, j & 31);
    A v = (A)0;
    if (j >= 0 && j <= K) {
      if (j > 0) v += gm1 * f;      // sampler_kernel.cu:94-95
      if (j < K) v += g0 * omf;     // sampler_kernel.cu:97-98
    }
    if (x1 < W2) dst[x1] = from_acc<T, A>(v);
  }
}

// fp32, W2 % 4 == 0: one warp writes 8 consecutive volume rows.  The 8 sample positions and the 8 x (2r+1)
// output gradients are fetched first in two batched rounds (two DRAM latencies per 8 rows instead of per row),
// then every row is written with 128-bit streaming stores: window values where they fall, zeros elsewhere.
constexpr int kBwdRows = 8;
__global__ void __launch_bounds__(256) sampler_bwd_f32v_kernel(const float* __restrict__ coords, long long coords_bstride,
                                                               const float* __restrict__ gout, float* __restrict__ gvol,
                                                               int HW, int W2, int r, long long total_rows) {
  __shared__ float sg[8][kBwdRows][33];
  __shared__ int st0[8][kBwdRows];
  __shared__ float sf[8][kBwdRows];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long row0 = ((long long)blockIdx.x * 8 + warp) * kBwdRows;
  if (row0 >= total_rows) return;
  const int K = 2 * r + 1;
  if (lane < kBwdRows) {
    const long long row = row0 + lane;
    float x0 = 0.f;
    if (row < total_rows) {
      const int b = (int)(row / HW);
      x0 = __ldg(coords + (long long)b * coords_bstride + (row - (long long)b * HW));
    }
    const float fl = floorf(x0);
    st0[warp][lane] = (int)fl - r;
    sf[warp][lane] = x0 - fl;
  }
  for (int idx = lane; idx < kBwdRows * K; idx += 32) {
    const int rr = idx / K, k = idx - rr * K;
    const long long row = row0 + rr;
    float v = 0.f;
    if (row < total_rows) {
      const int b = (int)(row / HW);
      v = __ldg(gout + ((long long)b * K + k) * HW + (row - (long long)b * HW));
    }
    sg[warp][rr][k] = v;
  }
  __syncwarp();
  for (int rr = 0; rr < kBwdRows; ++rr) {
    const long long row = row0 + rr;
    if (row >= total_rows) break;
    const int t0 = st0[warp][rr];
    const float f = sf[warp][rr], omf = 1.0f - f;
    float* dst = gvol + row * W2;
    const float* g = sg[warp][rr];
    for (int x1 = lane * 4; x1 < W2; x1 += 128) {
      const int j = x1 - t0;
      float o[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int ju = j + u;                              // tap index, inside the window when 0 <= ju <= K
        float v = 0.f;
        if (ju >= 0 && ju <= K) {
          if (ju > 0) v += g[ju - 1] * f;                  // sampler_kernel.cu:94-95
          if (ju < K) v += g[ju] * omf;                    // sampler_kernel.cu:97-98
        }
        o[u] = v;
      }
      as_stg_stream4(reinterpret_cast<float4*>(dst + x1), make_float4(o[0], o[1], o[2], o[3]));
    }
  }
}

template <typename T>
int launch_fwd(const void* volume, const float* coords, int coords_ch, void* out, int B, int H, int W1,
               int W2, int radius, cudaStream_t st) {
  using A = typename Acc<T>::type;
  const int HW = H * W1;
  dim3 grid(as_ceil_div(HW, kPix), B);
  const size_t smem = sizeof(A) * kPix * (2 * radius + 3);
  if (radius == 4)
    sampler_fwd_kernel<T, 4><<<grid, kPix, smem, st>>>(static_cast<const T*>(volume), coords,
                                                        (long long)coords_ch * HW, static_cast<T*>(out), HW, W2, radius);
  else
    sampler_fwd_kernel<T, 0><<<grid, kPix, smem, st>>>(static_cast<const T*>(volume), coords,
                                                        (long long)coords_ch * HW, static_cast<T*>(out), HW, W2, radius);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

template <typename T>
int launch_bwd(const float* coords, int coords_ch, const void* gout, void* gvol, int B, int H, int W1,
               int W2, int radius, cudaStream_t st) {
  const int HW = H * W1;
  const long long rows = (long long)B * HW;
  if (sizeof(T) == 4 && (W2 & 3) == 0 && as_aligned16(gvol)) {
    const long long vblocks = as_ceil_div_ll(rows, 8 * kBwdRows);
    sampler_bwd_f32v_kernel<<<(unsigned)vblocks, 256, 0, st>>>(coords, (long long)coords_ch * HW,
                                                              static_cast<const float*>(gout),
                                                              static_cast<float*>(gvol), HW, W2, radius, rows);
    AS_RETURN_IF_LAUNCH_FAILED();
    return AS_OK;
  }
  const long long blocks = as_ceil_div_ll(rows, 8);
  sampler_bwd_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(coords, (long long)coords_ch * HW,
                                                          static_cast<const T*>(gout),
                                                          static_cast<T*>(gvol), HW, W2, radius, rows);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

int check_args(const void* a, const void* b, const void* c, int coords_ch, int B, int H, int W1, int W2,
               int radius) {
  if (!a || !b || !c) return AS_ERR_BAD_ARG;
  if (B <= 0 || H <= 0 || W1 <= 0 || W2 <= 0 || radius < 0 || coords_ch < 1) return AS_ERR_BAD_ARG;
  if (2 * radius + 1 > 32) return AS_ERR_UNSUPPORTED;
  if (B > 65535) return AS_ERR_UNSUPPORTED;
  // the reference uses 32-bit accessors (PackedTensorAccessor32): same numel limit
  if ((long long)B * H * W1 * W2 >= (1LL << 31)) return AS_ERR_INDEX_RANGE;
  return AS_OK;
}

}  // namespace

extern "C" int as_sampler_fwd(const void* volume, const float* coords, int coords_ch, void* out, int B,
                              int H, int W1, int W2, int radius, int dtype, as_stream_t stream) {
  int rc = check_args(volume, coords, out, coords_ch, B, H, W1, W2, radius);
  if (rc != AS_OK) return rc;
  cudaStream_t st = as_cu(stream);
  switch (dtype) {
    case AS_DTYPE_F32: return launch_fwd<float>(volume, coords, coords_ch, out, B, H, W1, W2, radius, st);
    case AS_DTYPE_F16: return launch_fwd<__half>(volume, coords, coords_ch, out, B, H, W1, W2, radius, st);
    case AS_DTYPE_F64: return launch_fwd<double>(volume, coords, coords_ch, out, B, H, W1, W2, radius, st);
    default: return AS_ERR_UNSUPPORTED;
  }
}

extern "C" int as_sampler_bwd(const float* coords, int coords_ch, const void* corr_grad,
                              void* volume_grad, int B, int H, int W1, int W2, int radius, int dtype,
                              as_stream_t stream) {
  int rc = check_args(coords, corr_grad, volume_grad, coords_ch, B, H, W1, W2, radius);
  if (rc != AS_OK) return rc;
  cudaStream_t st = as_cu(stream);
  switch (dtype) {
    case AS_DTYPE_F32: return launch_bwd<float>(coords, coords_ch, corr_grad, volume_grad, B, H, W1, W2, radius, st);
    case AS_DTYPE_F16: return launch_bwd<__half>(coords, coords_ch, corr_grad, volume_grad, B, H, W1, W2, radius, st);
    case AS_DTYPE_F64: return launch_bwd<double>(coords, coords_ch, corr_grad, volume_grad, B, H, W1, W2, radius, st);
    default: return AS_ERR_UNSUPPORTED;
  }
}
