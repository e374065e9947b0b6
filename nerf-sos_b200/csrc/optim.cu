// Multi-tensor Adam with the learning rate of the step passed in (exponential schedule evaluated by the caller): one launch
// updates every trainable tensor.  Replaces torch.optim.Adam.step() of run_nerf.py:320 + engines/lr.py:20-23 on the training
// path (8 tensors under --fix_backbone, 56 for all parameters: 150+ ATen launches per step otherwise).
// Same update as torch.optim.Adam(amsgrad=False, weight_decay=0):
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
#include <math.h>

#include "internal.h"

namespace nsos {
namespace {
constexpr int kAdamMax = 64;
struct AdamBatch {
  float* p[kAdamMax];
  const float* g[kAdamMax];
  float* m[kAdamMax];
  float* v[kAdamMax];
  long long n[kAdamMax];
};
__global__ void __launch_bounds__(256) k_adam(const __grid_constant__ AdamBatch b, float b1, float b2, float step_size, float inv_bc2_sqrt, float eps) {
  const int t = blockIdx.y;
  float* __restrict__ p = b.p[t];
  const float* __restrict__ g = b.g[t];
  float* __restrict__ m = b.m[t];
  float* __restrict__ v = b.v[t];
  const long long n = b.n[t];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi; v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) * inv_bc2_sqrt + eps);
  }
}
}  // namespace

int adam_multi(const NsosAdamTensor* t, int n_tensors, float lr, float beta1, float beta2, float eps, int64_t step, cudaStream_t st) {
  NSOS_REQUIRE(step >= 1, NSOS_ERR_BAD_ARG, "nsos_adam_multi: step counts from 1");
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1), inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  for (int i0 = 0; i0 < n_tensors; i0 += kAdamMax) {
    AdamBatch b;
    memset(&b, 0, sizeof(b));
    const int nb = n_tensors - i0 < kAdamMax ? n_tensors - i0 : kAdamMax;
    long long mx = 0;
    for (int i = 0; i < nb; ++i) {
      const NsosAdamTensor& e = t[i0 + i];
      NSOS_REQUIRE(e.param && e.grad && e.exp_avg && e.exp_avg_sq && e.n >= 0, NSOS_ERR_BAD_ARG, "nsos_adam_multi: null tensor %d", i0 + i);
      b.p[i] = e.param; b.g[i] = e.grad; b.m[i] = e.exp_avg; b.v[i] = e.exp_avg_sq; b.n[i] = e.n;
      if (e.n > mx) mx = e.n;
    }
    if (mx == 0) continue;
    const int gx = (int)((mx + 255) / 256 < 296 ? (mx + 255) / 256 : 296);      // <= 2 waves of 148 SMs per tensor, grid-stride beyond
    k_adam<<<dim3(gx, nb), 256, 0, st>>>(b, beta1, beta2, step_size, inv_bc2_sqrt, eps);
  }
  NSOS_CHECK_CUDA(cudaGetLastError());
  return NSOS_OK;
}
}  // namespace nsos
