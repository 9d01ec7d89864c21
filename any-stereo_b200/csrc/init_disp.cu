// SURVEY 8(f) rank 3 (first half): the initial-disparity head that precedes the loop
// (models/coreContinuous_IGEV/continuous_IGEVstereo.py:267-268, submodule.py:321-325):
//     prob = softmax_D( Conv3d(G -> 1, 3x3x3, pad 1, no bias)(geo_encoding_volume) );  init_disp = sum_d prob[d] * d
// One kernel: per channel the [D+2, 4+2 rows, 32+2 columns] neighbourhood of a 32x4 output patch is staged in shared
// memory, every thread owns one pixel and a chunk of 16 disparities whose accumulators stay in registers over the
// channel loop (18 shared loads feed 48 FMAs per (channel, dy, dx)); the cost column goes through shared memory to the
// softmax and the expectation; neither the cost volume nor the probability volume reaches HBM (the reference writes
// and re-reads both).  Exact fp32.
#include "common.cuh"

namespace {

constexpr int kTX = 32;                  // pixels per row segment
constexpr int kRY = 4;                   // output rows per CTA
constexpr int kDC = 16;                  // disparities per thread chunk
constexpr int kMaxD = 64;
constexpr int kXP = kTX + 2, kRP = kRY + 2;

// CTA = 32 x 4 output pixels x all D.  Thread = (pixel, row, 16-disparity chunk) with its 16 accumulators in
// registers across the channel loop; per channel the [D+2][6][34] neighbourhood (zero padded) is staged once.
__global__ void __launch_bounds__(kTX * kRY * (kMaxD / kDC))
init_disp_kernel(const float* __restrict__ geo, const float* __restrict__ wgt, float* __restrict__ disp_out,
                 float* __restrict__ prob_out, int G, int D, int H, int W) {
  extern __shared__ float sm[];
  const int DP = D + 2;
  float* vol = sm;                                        // [DP][kRP][kXP]
  float* cost = vol + (size_t)DP * kRP * kXP;             // [D][kRY][kTX]
  float* ws = cost + (size_t)D * kRY * kTX;               // [G*27]
  const int b = blockIdx.z, y0 = blockIdx.y * kRY, x0 = blockIdx.x * kTX;
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int i = tid; i < G * 27; i += nthr) ws[i] = __ldg(wgt + i);
  const long long HW = (long long)H * W;
  const int px = tid % kTX, row = (tid / kTX) % kRY, chunk = tid / (kTX * kRY);
  const int d0 = chunk * kDC;
  float acc[kDC];
#pragma unroll
  for (int j = 0; j < kDC; ++j) acc[j] = 0.f;
  const int plane = kRP * kXP;
  for (int c = 0; c < G; ++c) {
    __syncthreads();                                      // previous channel fully consumed (and ws visible)
    const float* gc = geo + ((long long)b * G + c) * D * HW;
    for (int i = tid; i < DP * plane; i += nthr) {
      const int dd = i / plane, rem = i - dd * plane;
      const int r = rem / kXP, xx = rem - r * kXP;
      const int gx = x0 - 1 + xx, gy = y0 - 1 + r, gd = dd - 1;
      float v = 0.f;
      if (gx >= 0 && gx < W && gy >= 0 && gy < H && gd >= 0 && gd < D) v = __ldg(gc + (long long)gd * HW + (long long)gy * W + gx);
      vol[i] = v;
    }
    __syncthreads();
    if (d0 < D) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          // Conv3d weight [1][G][kd][kh][kw]
          const float w0 = ws[c * 27 + r * 3 + dx], w1 = ws[c * 27 + 9 + r * 3 + dx], w2 = ws[c * 27 + 18 + r * 3 + dx];
          const float* col = vol + (size_t)d0 * plane + (row + r) * kXP + px + dx;       // padded index d0 == disparity d0-1
          float vm = col[0], v0 = col[plane];
#pragma unroll
          for (int j = 0; j < kDC; ++j) {
            const float vp = (d0 + j + 2 < DP) ? col[(size_t)(j + 2) * plane] : 0.f;
            acc[j] = fmaf(w0, vm, fmaf(w1, v0, fmaf(w2, vp, acc[j])));
            vm = v0; v0 = vp;
          }
        }
      }
    }
  }
  if (d0 < D) {
#pragma unroll
    for (int j = 0; j < kDC; ++j)
      if (d0 + j < D) cost[((d0 + j) * kRY + row) * kTX + px] = acc[j];
  }
  __syncthreads();
  if (tid < kTX * kRY) {
    const int y = y0 + row, x = x0 + px;                  // tid < 128: chunk == 0, (row, px) as above
    if (y < H && x < W) {
      const float* cc = cost + row * kTX + px;
      const int st = kRY * kTX;
      float m = -3.0e38f;
      for (int d = 0; d < D; ++d) m = fmaxf(m, cc[d * st]);
      float s = 0.f, e1 = 0.f;
      for (int d = 0; d < D; ++d) s += expf(cc[d * st] - m);
      const float inv = 1.0f / s;
      const long long pix = (long long)y * W + x;
      for (int d = 0; d < D; ++d) {
        const float p = expf(cc[d * st] - m) * inv;
        e1 = fmaf(p, (float)d, e1);                                    // disparity_regression, submodule.py:321-325
        if (prob_out) prob_out[((long long)b * D + d) * HW + pix] = p;
      }
      disp_out[(long long)b * HW + pix] = e1;
    }
  }
}

__global__ void disparity_regression_kernel(const float* __restrict__ prob, float* __restrict__ out, int D, long long HW,
                                            long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long b = i / HW, p = i - b * HW;
  float s = 0.f;
  for (int d = 0; d < D; ++d) s += __ldg(prob + (b * D + d) * HW + p) * (float)d;     // torch.sum(x * disp_values, 1)
  out[i] = s;
}

}  // namespace

extern "C" int as_init_disparity(const float* geo, const float* weight, float* disp_out, float* prob_out, int B, int G, int D,
                                 int H, int W, as_stream_t stream) {
  if (!geo || !weight || !disp_out || B <= 0 || G <= 0 || D <= 0 || H <= 0 || W <= 0) return AS_ERR_BAD_ARG;
  if (D > kMaxD) return AS_ERR_UNSUPPORTED;
  if (B > 65535 || H > 65535) return AS_ERR_INDEX_RANGE;
  const size_t smem = ((size_t)(D + 2) * kRP * kXP + (size_t)D * kRY * kTX + (size_t)G * 27) * sizeof(float);
  if (smem > 227 * 1024) return AS_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(init_disp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  if (as_ceil_div(H, kRY) > 65535) return AS_ERR_INDEX_RANGE;
  dim3 grid(as_ceil_div(W, kTX), as_ceil_div(H, kRY), B);
  const int threads = kTX * kRY * as_ceil_div(D, kDC);
  init_disp_kernel<<<grid, threads, smem, as_cu(stream)>>>(geo, weight, disp_out, prob_out, G, D, H, W);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}

extern "C" int as_disparity_regression(const float* prob, float* out, int B, int D, int H, int W, as_stream_t stream) {
  if (!prob || !out || B <= 0 || D <= 0 || H <= 0 || W <= 0) return AS_ERR_BAD_ARG;
  const long long HW = (long long)H * W, total = HW * B;
  const long long blocks = as_ceil_div_ll(total, 256);
  if (blocks >= (1LL << 31)) return AS_ERR_INDEX_RANGE;
  disparity_regression_kernel<<<(unsigned)blocks, 256, 0, as_cu(stream)>>>(prob, out, D, HW, total);
  AS_RETURN_IF_LAUNCH_FAILED();
  return AS_OK;
}
