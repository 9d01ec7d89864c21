/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference's corr_sampler kernels.
 *
 * Follows sampler/sampler_kernel.cu of the reference: forward :19-60, backward :63-105 (one loop body
 * per CUDA thread), including the accumulate-into-zeroed-output order of :52-56 and :94-102.
 * Built by oracle/Makefile into oracle/_build/libsampler_oracle.so; used by tests and never by the product.
 * Parity pin: tests/test_oracle_golden.py checks it against the golden vectors made from the reference's
 * Python lookup (the CUDA extension itself cannot run in the build container).
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

static int within(int w, int W) { return w >= 0 && w < W; }

/* volume [B,H,W1,W2], coords [B,C,H,W1] (channel 0 used), corr [B,2r+1,H,W1] */
void sampler_forward_f32(const float* volume, const float* coords, int coords_ch, float* corr, int B, int H,
                         int W1, int W2, int r) {
  const int rd = 2 * r + 1;
  memset(corr, 0, sizeof(float) * (size_t)B * rd * H * W1);              /* torch::zeros, :123 */
  for (int n = 0; n < B; ++n)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W1; ++x) {
        const float x0 = coords[(((size_t)n * coords_ch + 0) * H + y) * W1 + x];   /* :39 */
        const float dx = x0 - floorf(x0);                                           /* :42 */
        for (int i = 0; i < rd + 1; ++i) {
          const int x1 = (int)floorf(x0) - r + i;                                   /* :47 */
          if (within(x1, W2)) {                                                      /* :49 */
            const float s = volume[(((size_t)n * H + y) * W1 + x) * W2 + x1];
            if (i > 0) corr[(((size_t)n * rd + (i - 1)) * H + y) * W1 + x] += s * dx;          /* :52-53 */
            if (i < rd) corr[(((size_t)n * rd + i) * H + y) * W1 + x] += s * (1.0f - dx);      /* :55-56 */
          }
        }
      }
}

void sampler_backward_f32(const float* coords, int coords_ch, const float* corr_grad, float* volume_grad, int B,
                          int H, int W1, int W2, int r) {
  const int rd = 2 * r + 1;
  memset(volume_grad, 0, sizeof(float) * (size_t)B * H * W1 * W2);        /* zeros_like, :148 */
  for (int n = 0; n < B; ++n)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W1; ++x) {
        const float x0 = coords[(((size_t)n * coords_ch + 0) * H + y) * W1 + x];
        const float dx = x0 - floorf(x0);
        for (int i = 0; i < rd + 1; ++i) {
          const int x1 = (int)floorf(x0) - r + i;
          if (within(x1, W2)) {
            float g = 0.0f;
            if (i > 0) g += corr_grad[(((size_t)n * rd + (i - 1)) * H + y) * W1 + x] * dx;     /* :94-95 */
            if (i < rd) g += corr_grad[(((size_t)n * rd + i) * H + y) * W1 + x] * (1.0f - dx); /* :97-98 */
            volume_grad[(((size_t)n * H + y) * W1 + x) * W2 + x1] += g;                        /* :100 */
          }
        }
      }
}

/* integer tap bookkeeping alone, for bit-exact index checks */
void sampler_taps_f32(const float* x0s, size_t n, int r, int* tap0, float* frac) {
  for (size_t i = 0; i < n; ++i) {
    tap0[i] = (int)floorf(x0s[i]) - r;
    frac[i] = x0s[i] - floorf(x0s[i]);
  }
}
