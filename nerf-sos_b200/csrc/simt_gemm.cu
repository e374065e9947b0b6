// Generic strided fp32 GEMM on CUDA cores.  Used only by the NSOS_MODE_SIMT_FP32 path (the
// same-device fp32 reference and the generic-configuration / backward path).  The throughput path is
// the tcgen05 kernel in tc_render.cu.
#include "common.cuh"

namespace nsos {

namespace {
constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

__device__ __forceinline__ float load_a(const GemmArgs& g, int m, int k) {
  if (k < g.K1) return __ldg(g.A1 + (int64_t)(m / g.a1_rowdiv) * g.a1_rs + (int64_t)k * g.a1_cs);
  k -= g.K1;
  return __ldg(g.A2 + (int64_t)(m / g.a2_rowdiv) * g.a2_rs + (int64_t)k * g.a2_cs);
}

__global__ void __launch_bounds__(NT) gemm_kernel(const GemmArgs g) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int t = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int K = g.K1 + g.K2;
  // K range of this split (multiple of BK so splits do not straddle tiles)
  int kt = (K + BK - 1) / BK;
  int per = (kt + g.split_k - 1) / g.split_k;
  int kbeg = blockIdx.z * per * BK, kend = min(K, kbeg + per * BK);
  if (kbeg >= kend) return;

  const bool a_kfast = (g.a1_cs == 1);
  const bool b_nfast = (g.b_cs == 1);
  const int tx = t % 16, ty = t / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int m, k;
      if (a_kfast) { k = t % BK; m = t / BK + 16 * j; } else { m = t % BM; k = t / BM + 4 * j; }
      float v = 0.f;
      if (m0 + m < g.M && k0 + k < kend) v = load_a(g, m0 + m, k0 + k);
      As[k][m] = v;
      int n, kb;
      if (b_nfast) { n = t % BN; kb = t / BN + 4 * j; } else { kb = t % BK; n = t / BK + 16 * j; }
      float w = 0.f;
      if (n0 + n < g.N && k0 + kb < kend) w = __ldg(g.B + (int64_t)((k0 + kb) / g.b_rowdiv) * g.b_rs + (int64_t)(n0 + n) * g.b_cs);
      Bs[kb][n] = w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.bias && blockIdx.z == 0) v += __ldg(g.bias + n);
      if (g.relu) v = fmaxf(v, 0.f);
      if (g.mask && !(__ldg(g.mask + (int64_t)m * g.mask_ld + n) > 0.f)) v = 0.f;
      float* c = g.C + (int64_t)m * g.c_rs + (int64_t)n * g.c_cs;
      if (g.split_k > 1) atomicAdd(c, v);
      else if (g.accumulate) *c += v;
      else *c = v;
    }
  }
}
}  // namespace

int launch_gemm(const GemmArgs& a, cudaStream_t s) {
  if (a.M <= 0 || a.N <= 0) return NSOS_OK;
  GemmArgs g = a;
  if (g.split_k < 1) g.split_k = 1;
  if (g.a1_rowdiv < 1) g.a1_rowdiv = 1;
  if (g.a2_rowdiv < 1) g.a2_rowdiv = 1;
  if (g.split_k > 1 && (g.relu || g.mask)) {
    set_error("launch_gemm: split_k cannot be combined with relu/mask epilogues");
    return NSOS_ERR_BAD_ARG;
  }
  if (g.b_rowdiv < 1) g.b_rowdiv = 1;
  dim3 grid((g.M + BM - 1) / BM, (g.N + BN - 1) / BN, g.split_k);
  gemm_kernel<<<grid, NT, 0, s>>>(g);
  NSOS_CHECK_CUDA(cudaGetLastError());
  return NSOS_OK;
}

}  // namespace nsos
