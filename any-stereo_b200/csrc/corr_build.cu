// a1 + a2: as_corr1d_build -- all-pairs correlation level 0 (mode-selected kernel) and the pooled levels.
#include "common.cuh"

int as_corr_fwd_simt_launch(const float* f1, const float* f2, float* lvl0, int B, int D, int H, int W1, int W2,
                            int pitch, cudaStream_t st);
// tcgen05 path (corr_umma.cu)
size_t as_corr_umma_workspace_bytes(int B, int D, int H, int W1, int W2, int mode);
int as_corr_umma_launch(const float* f1, const float* f2, int B, int D, int H, int W1, int W2, int num_levels,
                        float* const* levels, const int* pitches, int mode, void* ws, size_t ws_bytes,
                        cudaStream_t st);

extern "C" size_t as_corr1d_workspace_bytes(int B, int D, int H, int W1, int W2, int mode) {
  if (mode == AS_CORR_FP32_SIMT) return 0;
  return as_corr_umma_workspace_bytes(B, D, H, W1, W2, mode);
}

extern "C" int as_corr1d_build(const float* f1, const float* f2, int B, int D, int H, int W1, int W2, int num_levels,
                               float* const* levels, const int* pitches, int mode, void* workspace,
                               size_t workspace_bytes, as_stream_t stream) {
  if (!f1 || !f2 || !levels || !pitches) return AS_ERR_BAD_ARG;
  if (B <= 0 || D <= 0 || H <= 0 || W1 <= 0 || W2 <= 0 || num_levels < 1 || num_levels > AS_MAX_LEVELS)
    return AS_ERR_BAD_ARG;
  for (int l = 0; l < num_levels; ++l) {
    if (!levels[l] || pitches[l] < (W2 >> l)) return AS_ERR_BAD_ARG;
    if ((pitches[l] & 3) || !as_aligned16(levels[l])) return AS_ERR_ALIGNMENT;
  }
  cudaStream_t st = as_cu(stream);
  const long long rows = (long long)B * H * W1;
  if (mode == AS_CORR_FP32_SIMT) {
    int rc = as_corr_fwd_simt_launch(f1, f2, levels[0], B, D, H, W1, W2, pitches[0], st);
    if (rc != AS_OK) return rc;
    for (int l = 1; l < num_levels; ++l) {
      rc = as_pool1d_halve(levels[l - 1], levels[l], rows, W2 >> (l - 1), pitches[l - 1], pitches[l], stream);
      if (rc != AS_OK) return rc;
    }
    return AS_OK;
  }
  if (mode == AS_CORR_BF16X3 || mode == AS_CORR_BF16)
    return as_corr_umma_launch(f1, f2, B, D, H, W1, W2, num_levels, levels, pitches, mode, workspace,
                               workspace_bytes, st);
  return AS_ERR_UNSUPPORTED;
}
