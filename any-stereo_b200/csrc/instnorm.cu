// SURVEY 8(f)-4, RAFT family: the feature encoder `fnet` = BasicEncoder(norm_fn='instance')
// (models/corePrune_RAFT/extractor.py:126-201, ResidualBlock :9-58; call site prune_raft_stereo.py:108,252) normalises every
// convolution output with nn.InstanceNorm2d(affine=False, track_running_stats=False): per image and channel
//
//     y = (x - mean_hw(x)) / sqrt(var_hw(x) + eps)          (biased variance, eps = 1e-5)
//
// followed by ReLU, and at the end of a ResidualBlock by relu(x' + y).  ATen runs each of them as a training-mode batch norm
// over a reshaped NCHW tensor plus separate ReLU / add kernels: 15 normalisations of up to 245 MB each are what makes
// `fnet` the largest module of the RAFT forward once the loop is fast (11.9 of 32 ms for one 384x1248 pair).
//
// Here, on pixel-major (channels-last) fp32 tensors [B][H*W][C]:
//   1. instnorm_stats_kernel   one read of x: per-thread fp32 partial sums of x - x0 and (x - x0)^2 over <= 256 pixels (x0 = the
//                              image's first pixel of that channel: the shift removes the cancellation in E[x^2] - mean^2),
//                              block reduction through shared memory, one fp64 atomic pair per (block, channel);
//   2. instnorm_finalize_kernel mean / rstd per (image, channel) in fp64 -> fp32;
//   3. instnorm_apply_kernel   y = (x - mean) * rstd, optional ReLU, optional relu(resid + y): one read of x (and of the
//                              residual), one write -- the ReLU and the residual add never make their own passes.
// HBM-bound: 4 B/element read in (1), 4 (+4) read + 4 written in (3).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxPixPerBlock = 2048;   // pixels of one image a statistics block walks, at most
constexpr int kMinStatBlocks = 592;     // ... and at least 4 blocks per SM in flight: the small (1/4-resolution) tensors of the
                                        // encoder would otherwise run the reduction on 30 blocks (91 us instead of ~40)

__global__ void __launch_bounds__(kThreads)
instnorm_stats_kernel(const float* __restrict__ x, double* __restrict__ sums, long long HW, int C, int pix_per_block) {
  // thread = (pixel row r, channel quad q): consecutive threads read consecutive float4 of one pixel row -> coalesced
  const int c4 = C >> 2;
  const int rows = kThreads / c4;
  const int q = threadIdx.x % c4, r = threadIdx.x / c4;
  const int b = blockIdx.y;
  const float4* xb = reinterpret_cast<const float4*>(x) + (long long)b * HW * c4;
  __shared__ float4 red_s[kThreads], red_q[kThreads];
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), ss = s;
  if (r < rows) {
    const float4 x0 = __ldg(xb + q);                                   // the shift: pixel 0 of this image, this channel quad
    const long long p0 = (long long)blockIdx.x * pix_per_block;
    const long long p1 = min(p0 + (long long)pix_per_block, HW);
    auto acc = [&](const float4 v) {
      const float dx = v.x - x0.x, dy = v.y - x0.y, dz = v.z - x0.z, dw = v.w - x0.w;
      s.x += dx; s.y += dy; s.z += dz; s.w += dw;
      ss.x = fmaf(dx, dx, ss.x); ss.y = fmaf(dy, dy, ss.y); ss.z = fmaf(dz, dz, ss.z); ss.w = fmaf(dw, dw, ss.w);
    };
    long long p = p0 + r;
    const long long step = rows;
    for (; p + 3 * step < p1; p += 4 * step) {                          // four independent 16-byte loads in flight per thread
      const float4 v0 = __ldg(xb + p * c4 + q), v1 = __ldg(xb + (p + step) * c4 + q);
      const float4 v2 = __ldg(xb + (p + 2 * step) * c4 + q), v3 = __ldg(xb + (p + 3 * step) * c4 + q);
      acc(v0); acc(v1); acc(v2); acc(v3);
    }
    for (; p < p1; p += step) acc(__ldg(xb + p * c4 + q));
  }
  red_s[threadIdx.x] = s;
  red_q[threadIdx.x] = ss;
  __syncthreads();
  if (threadIdx.x < c4) {                                              // one thread per channel quad folds the pixel rows in fp64
    double a[4] = {0, 0, 0, 0}, e[4] = {0, 0, 0, 0};
    for (int i = 0; i < rows; ++i) {
      const float4 u = red_s[i * c4 + threadIdx.x], w = red_q[i * c4 + threadIdx.x];
      a[0] += u.x; a[1] += u.y; a[2] += u.z; a[3] += u.w;
      e[0] += w.x; e[1] += w.y; e[2] += w.z; e[3] += w.w;
    }
    double* dst = sums + ((long long)b * C + threadIdx.x * 4) * 2;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(dst + 2 * j, a[j]);
      atomicAdd(dst + 2 * j + 1, e[j]);
    }
  }
}

__global__ void instnorm_finalize_kernel(const float* __restrict__ x, const double* __restrict__ sums, float* __restrict__ mr,
                                         long long HW, int C, int BC, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BC) return;
  const int b = i / C, c = i - b * C;
  const double x0 = (double)__ldg(x + (long long)b * HW * C + c);
  const double m = sums[2 * i] / (double)HW;                            // mean of x - x0
  double var = sums[2 * i + 1] / (double)HW - m * m;
  if (var < 0.0) var = 0.0;
  mr[2 * i] = (float)(x0 + m);
  mr[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

template <bool RELU, bool RES>
__global__ void __launch_bounds__(kThreads)
instnorm_apply_kernel(const float* __restrict__ x, const float* __restrict__ resid, const float* __restrict__ mr,
                      float* __restrict__ out, long long HW, int C, long long n4) {
  const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n4) return;
  const int c4 = C >> 2;
  const long long pix = i / c4;
  const int q = (int)(i - pix * c4);
  const int b = (int)(pix / HW);
  const float4* st = reinterpret_cast<const float4*>(mr + ((long long)b * C + q * 4) * 2);
  const float4 m01 = __ldg(st), m23 = __ldg(st + 1);                    // (mean, rstd) x 4 channels
  const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
  float4 y;
  y.x = (v.x - m01.x) * m01.y; y.y = (v.y - m01.z) * m01.w;
  y.z = (v.z - m23.x) * m23.y; y.w = (v.w - m23.z) * m23.w;
  if (RELU) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
  if (RES) {                                                            // ResidualBlock.forward: relu(x' + y)  extractor.py:58
    const float4 rz = __ldg(reinterpret_cast<const float4*>(resid) + i);
    y.x = fmaxf(y.x + rz.x, 0.f); y.y = fmaxf(y.y + rz.y, 0.f); y.z = fmaxf(y.z + rz.z, 0.f); y.w = fmaxf(y.w + rz.w, 0.f);
  }
  reinterpret_cast<float4*>(out)[i] = y;
}

}  // namespace

extern "C" size_t as_instnorm_workspace_bytes(int B, int C) {
  if (B <= 0 || C <= 0) return 0;
  return (size_t)B * C * (2 * sizeof(double) + 2 * sizeof(float));
}

// x, out (and resid): fp32 [B][HW][C] (channels-last); out may alias x or resid (every element is read before it is
// written, by the same thread); resid may be NULL.
extern "C" int as_instnorm_nhwc(const float* x, const float* resid, float* out, void* workspace, size_t workspace_bytes, int B,
                                long long HW, int C, float eps, int relu, as_stream_t stream) {
  if (!x || !out || !workspace) return AS_ERR_BAD_ARG;
  if (B <= 0 || HW <= 0 || C <= 0 || !(eps > 0.f)) return AS_ERR_BAD_ARG;
  if (C % 4 != 0 || C > 4 * kThreads) return AS_ERR_UNSUPPORTED;
  if (workspace_bytes < as_instnorm_workspace_bytes(B, C)) return AS_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(resid) |
       reinterpret_cast<uintptr_t>(workspace)) & 15)
    return AS_ERR_ALIGNMENT;
  const long long n4 = (long long)B * HW * (C / 4);
  const long long blocks = as_ceil_div_ll(n4, (long long)kThreads);
  long long per_image = as_ceil_div_ll(HW, (long long)kMaxPixPerBlock);
  const long long want = as_ceil_div(kMinStatBlocks, B);
  if (per_image < want) per_image = want;
  long long ppb = as_ceil_div_ll(HW, per_image);
  if (ppb < 64) ppb = 64;
  const int pix_per_block = (int)ppb;
  const long long sblocks = as_ceil_div_ll(HW, ppb);
  if (B > 65535 || blocks >= (1LL << 31) || sblocks >= (1LL << 31)) return AS_ERR_INDEX_RANGE;
  double* sums = reinterpret_cast<double*>(workspace);
  float* mr = reinterpret_cast<float*>(sums + (size_t)B * C * 2);
  cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)B * C * 2 * sizeof(double), as_cu(stream));
  if (e != cudaSuccess) return (int)e;
  instnorm_stats_kernel<<<dim3((unsigned)sblocks, (unsigned)B), kThreads, 0, as_cu(stream)>>>(x, sums, HW, C, pix_per_block);
  AS_RETURN_IF_LAUNCH_FAILED();
  instnorm_finalize_kernel<<<as_ceil_div(B * C, 128), 128, 0, as_cu(stream)>>>(x, sums, mr, HW, C, B * C, eps);
  AS_RETURN_IF_LAUNCH_FAILED();
  const unsigned g = (unsigned)blocks;
  if (resid) {
    if (relu) instnorm_apply_kernel<true, true><<<g, kThreads, 0, as_cu(stream)>>>(x, resid, mr, out, HW, C, n4);
    else instnorm_apply_kernel<false, true><<<g, kThreads, 0, as_cu(stream)>>>(x, resid, mr, out, HW, C, n4);
  } else {
    if (relu) instnorm_apply_kernel<true, false><<<g, kThreads, 0, as_cu(stream)>>>(x, resid, mr, out, HW, C, n4);
    else instnorm_apply_kernel<false, false><<<g, kThreads, 0, as_cu(stream)>>>(x, resid, mr, out, HW, C, n4);
  }
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}
