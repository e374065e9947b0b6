// NSOS_MODE_SIMT_FP32: layer-by-layer fp32 render path on CUDA cores.
//  * forward for ANY net geometry (the same-device fp32 reference for the tcgen05 kernel, and the path
//    for configurations the fused kernel does not cover, e.g. BASELINE config[0] D=4 W=64)
//  * backward (recompute-in-backward, chunked) for training: replaces autograd through
//    models/nerf_mlp.py:67-100 and models/renderer.py:21-85 (engines/trainer.py:201).
#include <cstdlib>
#include "render_device.cuh"
#include "internal.h"

namespace nsos {

namespace {

constexpr int kEncLd = 64;   // padded row length of the point encoding (63 -> 64)
constexpr int kEncVLd = 32;  // padded row length of the direction encoding (27 -> 32)

// ---- sampling / encoding kernels ---------------------------------------------------------------
__global__ void k_coarse_z(const float* __restrict__ near, const float* __restrict__ far, const float* __restrict__ t_rand,
                           float perturb, uint64_t seed, int64_t ray0, float* __restrict__ z, int64_t n_rays, int S) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rays * S) return;
  int64_t r = idx / S; int i = (int)(idx % S);
  bool pert = perturb > 0.f;
  float t = 0.f;
  if (pert) t = t_rand ? t_rand[idx] : rng_uniform(seed, ray0 + r, RNG_T_RAND, i);
  z[idx] = z_stratified(near[r], far[r], i, S, pert, t);
}

// enc[p, 0:63] = gamma(o + d*z) (embedder.py:34-48), enc[p,63] = 0
__global__ void k_encode_pts(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ z,
                             float* __restrict__ enc, int64_t n_pts, int S, int L) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pts) return;
  int64_t r = p / S;
  float zz = z[p];
  float x[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) x[a] = pt_coord(rays_o[r * 3 + a], rays_d[r * 3 + a], zz);
  float* e = enc + p * kEncLd;
  float buf[3 + 6 * 16];
  encode3(x, L, buf);
  int n = 3 + 6 * L;
  for (int c = 0; c < kEncLd; ++c) e[c] = (c < n) ? buf[c] : 0.f;
}

// raw points (mlp_query): enc from pts directly
__global__ void k_encode_raw(const float* __restrict__ x3, float* __restrict__ enc, int64_t n, int L, int ld) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  float x[3] = {x3[p * 3], x3[p * 3 + 1], x3[p * 3 + 2]};
  float buf[3 + 6 * 16];
  encode3(x, L, buf);
  int m = 3 + 6 * L;
  for (int c = 0; c < ld; ++c) enc[p * ld + c] = (c < m) ? buf[c] : 0.f;
}

// per-ray: viewdirs = d/||d|| (nerf_net.py:165), gamma_v(viewdirs) [N,32], dnorm [N]
__global__ void k_encode_dirs(const float* __restrict__ rays_d, float* __restrict__ encv, float* __restrict__ dnorm, int64_t n_rays,
                              int L) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  float d0 = rays_d[r * 3], d1 = rays_d[r * 3 + 1], d2 = rays_d[r * 3 + 2];
  float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)));
  dnorm[r] = nrm;
  float v[3] = {__fdiv_rn(d0, nrm), __fdiv_rn(d1, nrm), __fdiv_rn(d2, nrm)};
  float buf[3 + 6 * 16];
  encode3(v, L, buf);
  int m = 3 + 6 * L;
  for (int c = 0; c < kEncVLd; ++c) encv[r * kEncVLd + c] = (c < m) ? buf[c] : 0.f;
}

// ---- compositing / importance kernels: one warp per ray ----------------------------------------
__global__ void k_composite(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ dnorm,
                            const float* __restrict__ noise, float noise_std, uint64_t seed, int64_t ray0, int rng_stream,
                            int S, int C, int sem_dim, int white_bkgd, float* __restrict__ maps, int maps_ld, int maps_off,
                            float* __restrict__ weights, int64_t n_rays) {
  int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= n_rays) return;
  RayPass p;
  p.raw = raw + r * S * C; p.z = z + r * S; p.noise = noise ? noise + r * S : nullptr;
  p.noise_std = noise_std; p.seed = seed; p.ray = ray0 + r; p.rng_stream = rng_stream;
  p.dnorm = dnorm[r]; p.S = S; p.C = C; p.sem_dim = sem_dim; p.white_bkgd = white_bkgd;
  warp_composite(p, lane, maps + r * maps_ld + maps_off, weights ? weights + r * S : nullptr);
}

__global__ void k_composite_bwd(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ dnorm,
                                const float* __restrict__ noise, float noise_std, uint64_t seed, int64_t ray0, int rng_stream,
                                int S, int C, int sem_dim, int white_bkgd, const float* __restrict__ g_maps, int maps_ld,
                                int maps_off, float* __restrict__ g_raw, int64_t n_rays) {
  int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= n_rays) return;
  RayPass p;
  p.raw = raw + r * S * C; p.z = z + r * S; p.noise = noise ? noise + r * S : nullptr;
  p.noise_std = noise_std; p.seed = seed; p.ray = ray0 + r; p.rng_stream = rng_stream;
  p.dnorm = dnorm[r]; p.S = S; p.C = C; p.sem_dim = sem_dim; p.white_bkgd = white_bkgd;
  warp_composite_bwd(p, lane, g_maps + r * maps_ld + maps_off, g_raw + r * S * C);
}

// 4 warps / block; dynamic smem per warp: z0[Sc] w0[Sc] cdf[Sc] bins[Sc] zall[Sc+K]
__global__ void k_importance(const float* __restrict__ z0, const float* __restrict__ w0, const float* __restrict__ u,
                             const float* __restrict__ z_inject, float perturb,
                             uint64_t seed, int64_t ray0, int Sc, int K, float* __restrict__ z_fine, float* __restrict__ z_samples,
                             int64_t* __restrict__ inds, float* __restrict__ z_std, int zstd_ld, int64_t n_rays) {
  extern __shared__ float sm[];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (r >= n_rays) return;
  float* base = sm + (size_t)warp * (5 * Sc + K);
  float* sz = base; float* sw = base + Sc;
  for (int i = lane; i < Sc; i += 32) { sz[i] = z0[r * Sc + i]; sw[i] = w0[r * Sc + i]; }
  __syncwarp();
  ImportanceIO io;
  io.z0 = sz; io.w0 = sw; io.cdf = base + 2 * Sc; io.bins = base + 3 * Sc; io.zall = base + 4 * Sc;
  io.zsorted = z_fine + r * (Sc + K);
  io.u = u ? u + r * K : nullptr;
  io.z_inject = z_inject ? z_inject + r * K : nullptr;
  io.z_samples = z_samples ? z_samples + r * K : nullptr;
  io.inds = inds ? inds + r * K : nullptr;
  io.z_std = z_std ? z_std + r * zstd_ld : nullptr;
  io.Sc = Sc; io.K = K; io.det = !(perturb > 0.f); io.seed = seed; io.ray = ray0 + r;
  warp_importance(io, lane);
}

__global__ void k_invert_cdf(const float* __restrict__ bins, const float* __restrict__ cdf, const float* __restrict__ u,
                             float* __restrict__ samples, int64_t* __restrict__ inds, int64_t n_rays, int M, int K) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rays * K) return;
  int64_t r = idx / K;
  int ind;
  float s = invert_cdf_one(cdf + r * M, bins + r * M, M, u[idx], &ind);
  samples[idx] = s;
  inds[idx] = ind;
}

__global__ void k_fill(float* p, float v, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
// out[n] += sum_m X[m*ld + n]   (bias gradients)
__global__ void k_colsum(const float* __restrict__ X, int64_t ld, int64_t M, int N, float* __restrict__ out) {
  int n = blockIdx.y * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int64_t per = (M + gridDim.x - 1) / gridDim.x;
  int64_t m0 = blockIdx.x * per, m1 = min(M, m0 + per);
  float s = 0.f;
  for (int64_t m = m0; m < m1; ++m) s += X[m * ld + n];
  atomicAdd(out + n, s);
}
// G[m,n] = (H[m,n] > 0) ? G[m,n] : 0
__global__ void k_relu_mask(float* __restrict__ G, const float* __restrict__ H, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && !(H[i] > 0.f)) G[i] = 0.f;
}

inline dim3 grid1(int64_t n, int bs = 256) { return dim3((unsigned)((n + bs - 1) / bs)); }

// ---- narrow products of the head backward (the 1-, 3- and sem_dim-wide outputs): memory-bound fp32 kernels -----------------
// dX[p, k] (=|+=) [mask(p,k) > 0] * sum_{n<N} dY[p*ldy + n] * W[n*ldw + k],  N <= 4, K % 4 == 0
__global__ void k_small_dgrad(const float* __restrict__ dY, int64_t ldy, int N, const float* __restrict__ W, int ldw, int K,
                              float* __restrict__ dX, int64_t ldx, const float* __restrict__ mask, int accumulate, int64_t P) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int k4n = K / 4;
  if (idx >= P * k4n) return;
  const int64_t p = idx / k4n;
  const int k = (int)(idx % k4n) * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int n = 0; n < N; ++n) {
    const float g = __ldg(&dY[p * ldy + n]);
    const float* w = W + (int64_t)n * ldw + k;
    acc.x = fmaf(g, __ldg(w), acc.x); acc.y = fmaf(g, __ldg(w + 1), acc.y); acc.z = fmaf(g, __ldg(w + 2), acc.z); acc.w = fmaf(g, __ldg(w + 3), acc.w);
  }
  float4* dst = reinterpret_cast<float4*>(dX + p * ldx + k);
  if (mask) {
    const float4 m = *reinterpret_cast<const float4*>(mask + p * ldx + k);
    if (!(m.x > 0.f)) acc.x = 0.f;
    if (!(m.y > 0.f)) acc.y = 0.f;
    if (!(m.z > 0.f)) acc.z = 0.f;
    if (!(m.w > 0.f)) acc.w = 0.f;
  }
  if (accumulate) { const float4 o = *dst; acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w; }
  *dst = acc;
}
// dW[n*ldw + k] += sum_p dY[p*ldy + n] * X[p*ldx + k]  (N <= 4, K <= blockDim.x);  db[n] += sum_p dY[p*ldy + n]
__global__ void __launch_bounds__(256) k_small_wgrad(const float* __restrict__ dY, int64_t ldy, int N, const float* __restrict__ X, int64_t ldx,
                                                     int K, float* __restrict__ dW, int ldw, float* __restrict__ db, int64_t P) {
  const int k = threadIdx.x;
  const int64_t per = (P + gridDim.x - 1) / gridDim.x;
  const int64_t p0 = (int64_t)blockIdx.x * per, p1 = min(P, p0 + per);
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, bsum = 0.f;
  for (int64_t p = p0; p < p1; ++p) {
    const float x = (k < K) ? X[p * ldx + k] : 0.f;
#pragma unroll
    for (int n = 0; n < 4; ++n)
      if (n < N) acc[n] = fmaf(__ldg(&dY[p * ldy + n]), x, acc[n]);
    if (db && k < N) bsum += __ldg(&dY[p * ldy + k]);
  }
  if (k < K)
    for (int n = 0; n < N; ++n) atomicAdd(&dW[(int64_t)n * ldw + k], acc[n]);
  if (db && k < N) atomicAdd(&db[k], bsum);
}

// ---- tensor-core dispatch for the large GEMMs of the all-parameter backward ------------------------
// NSOS_MODE_TC_*: products with a 64..256-wide contraction and a 32..256-wide output run on tcgen05 (tc_rowgemm / tc_wgrad_gen:
// bf16 hi/lo operands, 3 MMAs per product, fp32 accumulate); everything narrow (the 1-, 3- and sem_dim-wide heads, the per-ray
// direction encoding) stays on the fp32 CUDA-core GEMM.
struct TcCtx {
  bool on = false;
  void* scratch = nullptr;       // packed B operand of tc_rowgemm
  size_t scratch_bytes = 0;
  void* part = nullptr;          // per-CTA partial sums of the weight-gradient kernels (tc_wgrad_part_bytes())
  size_t part_bytes = 0;
};
int gemm_dispatch(const GemmArgs& a, const TcCtx* tc, cudaStream_t st) {
  if (tc && tc->on && !a.A2 && a.a1_cs == 1 && a.a1_rowdiv <= 1 && a.b_rowdiv <= 1 && a.c_cs == 1 && a.split_k <= 1 && a.K1 % 64 == 0 &&
      tc_rowgemm_supported(a.K1, a.N, a.a1_rs, a.c_rs, a.mask ? a.mask_ld : 0) &&
      tc->scratch_bytes >= tc_rowgemm_scratch_bytes(a.K1, a.N))
    return tc_rowgemm(a.A1, a.a1_rs, a.K1, a.B, a.b_rs, a.b_cs, a.C, a.c_rs, a.N, a.mask, a.mask_ld, a.bias, a.relu, a.accumulate, a.M,
                      tc->scratch, tc->scratch_bytes, st);
  return launch_gemm(a, st);
}

// ---- MLP forward over P points -----------------------------------------------------------------
struct MlpBufs {
  float* h[16];   // per-layer post-ReLU activations [P,W]; forward-only mode aliases two buffers
  float* feat;    // [P,W]
  float* hv;      // [P,W/2]
  float* s0;      // [P,W/2]
};

GemmArgs lin(const float* A1, int64_t lda1, int K1, const float* A2, int64_t lda2, int K2, int a2_rowdiv, const float* Wt,
             const float* bias, float* C, int64_t ldc, int64_t M, int N, int relu) {
  GemmArgs g{};
  g.A1 = A1; g.a1_rs = lda1; g.a1_cs = 1; g.K1 = K1; g.a1_rowdiv = 1;
  g.A2 = A2; g.a2_rs = lda2; g.a2_cs = 1; g.K2 = K2; g.a2_rowdiv = a2_rowdiv;
  g.B = Wt; g.b_rs = 1; g.b_cs = K1 + K2; g.b_rowdiv = 1;   // W is [N, K] row-major: B(k,n) = W[n*K + k]
  g.bias = bias; g.C = C; g.c_rs = ldc; g.c_cs = 1; g.M = (int)M; g.N = N; g.relu = relu; g.split_k = 1;
  return g;
}

// MLP.forward (nerf_mlp.py:67-100).  enc [P,kEncLd], encv [P/S rows, kEncVLd] broadcast per ray via rowdiv=S.
int mlp_forward(const NetGeom& g, const float* prm, const float* enc, const float* encv, int S, int64_t P, const MlpBufs& b,
                float* raw, cudaStream_t st, const TcCtx* tc = nullptr) {
  const int W = g.W;
  const float* h = nullptr;
  for (int i = 0; i < g.D; ++i) {
    GemmArgs a;
    if (i == 0) a = lin(enc, kEncLd, g.enc, nullptr, 0, 0, 1, prm + g.w_pts[i], prm + g.b_pts[i], b.h[i], W, P, W, 1);
    else if (g.in_pts[i] == W) a = lin(h, W, W, nullptr, 0, 0, 1, prm + g.w_pts[i], prm + g.b_pts[i], b.h[i], W, P, W, 1);
    else a = lin(enc, kEncLd, g.enc, h, W, W, 1, prm + g.w_pts[i], prm + g.b_pts[i], b.h[i], W, P, W, 1);  // [enc, h] :74
    int rc = gemm_dispatch(a, tc, st); if (rc) return rc;
    h = b.h[i];
  }
  if (!g.use_viewdirs) {
    return launch_gemm(lin(h, W, W, nullptr, 0, 0, 1, prm + g.w_out, prm + g.b_out, raw, g.C, P, 4, 0), st);
  }
  int rc;
  rc = launch_gemm(lin(h, W, W, nullptr, 0, 0, 1, prm + g.w_alpha, prm + g.b_alpha, raw + 3, g.C, P, 1, 0), st);  // :77
  if (rc) return rc;
  if (g.use_sem) {                                                                                             // :79-80
    GemmArgs a = g.sem_coord ? lin(h, W, W, enc, kEncLd, g.enc, 1, prm + g.w_s0, prm + g.b_s0, b.s0, W / 2, P, W / 2, 1)
                             : lin(h, W, W, nullptr, 0, 0, 1, prm + g.w_s0, prm + g.b_s0, b.s0, W / 2, P, W / 2, 1);
    rc = gemm_dispatch(a, tc, st); if (rc) return rc;
    rc = launch_gemm(lin(b.s0, W / 2, W / 2, nullptr, 0, 0, 1, prm + g.w_s2, prm + g.b_s2, raw + 4, g.C, P, g.sem_dim, 0), st);
    if (rc) return rc;
  }
  rc = gemm_dispatch(lin(h, W, W, nullptr, 0, 0, 1, prm + g.w_feat, prm + g.b_feat, b.feat, W, P, W, 0), tc, st);       // :86
  if (rc) return rc;
  rc = launch_gemm(lin(b.feat, W, W, encv, kEncVLd, g.encv, S, prm + g.w_views, prm + g.b_views, b.hv, W / 2, P, W / 2, 1), st);
  if (rc) return rc;                                                                                           // :87-90
  return launch_gemm(lin(b.hv, W / 2, W / 2, nullptr, 0, 0, 1, prm + g.w_rgb, prm + g.b_rgb, raw, g.C, P, 3, 0), st);  // :92
}

// dW[N,K] += dY[P,N]^T . X[P,K]  (X possibly two sources)   -- split-K over P with atomics
// db != nullptr: also the bias gradient db[N] += column sums of dY (in the same tensor-core launch when possible)
int bgrad(const float* dY, int64_t ldy, int N, float* db, int64_t P, cudaStream_t st);
int wgrad(const float* dY, int64_t ldy, int N, const float* X1, int64_t ldx1, int K1, const float* X2, int64_t ldx2, int K2,
          int x2_rowdiv, float* dW, int64_t P, cudaStream_t st, const TcCtx* tc = nullptr, float* db = nullptr) {
  if (N <= 4 && X1 && !X2 && K1 <= 256) {          // narrow head: one memory-bound pass over X
    k_small_wgrad<<<(unsigned)std::min<int64_t>(1184, (P + 255) / 256), 256, 0, st>>>(dY, ldy, N, X1, ldx1, K1, dW, K1, db, P);
    NSOS_CHECK_CUDA(cudaGetLastError());
    return NSOS_OK;
  }
  // tensor cores: the 256-wide source is `main`, a <= 64-wide source of per-point rows is `aux` (gamma(x)); one launch for both
  bool done[2] = {false, false};
  if (tc && tc->on && tc->part && (N == 128 || N == 256) && ldy % 4 == 0) {
    const float* X[2] = {X1, X2}; const int64_t ld[2] = {ldx1, ldx2}; const int K[2] = {K1, K2}; const int col[2] = {0, K1};
    const int rd[2] = {1, x2_rowdiv};
    int im = -1, ia = -1;
    for (int part = 0; part < 2; ++part) {
      if (!X[part] || K[part] == 0 || rd[part] > 1) continue;
      if (K[part] == 256 && ld[part] % 4 == 0 && im < 0) im = part;
      else if (K[part] <= 64 && ia < 0) ia = part;
    }
    if (im >= 0 || ia >= 0) {
      float* db_tc = (ia < 0 || K[ia] <= 63) ? db : nullptr;
      int rc = tc_wgrad_gen(dY, ldy, N, im >= 0 ? X[im] : nullptr, im >= 0 ? ld[im] : 0, im >= 0 ? col[im] : 0, ia >= 0 ? X[ia] : nullptr,
                            ia >= 0 ? ld[ia] : 0, ia >= 0 ? K[ia] : 0, ia >= 0 ? col[ia] : 0, dW, K1 + K2, db_tc, P, tc->part, tc->part_bytes, st);
      if (rc) return rc;
      if (im >= 0) done[im] = true;
      if (ia >= 0) done[ia] = true;
      if (db_tc) db = nullptr;
    }
  }
  if (db) { int rc = bgrad(dY, ldy, N, db, P, st); if (rc) return rc; }
  // as a GEMM: M=N (rows of dW), N=K (cols), K=P.  A(m,k)=dY[k*ldy+m]; B(k,n)=X[k*ldx+n]
  for (int part = 0; part < 2; ++part) {
    const float* X = part ? X2 : X1; int64_t ldx = part ? ldx2 : ldx1; int Kp = part ? K2 : K1;
    if (!X || Kp == 0 || done[part]) continue;
    GemmArgs g{};
    g.A1 = dY; g.a1_rs = 1; g.a1_cs = ldy; g.K1 = (int)P; g.a1_rowdiv = 1; g.a2_rowdiv = 1;
    g.B = X; g.b_rs = ldx; g.b_cs = 1; g.b_rowdiv = part ? x2_rowdiv : 1;
    g.C = dW + (part ? K1 : 0); g.c_rs = K1 + K2; g.c_cs = 1; g.M = N; g.N = Kp; g.accumulate = 1;
    int64_t sk = P / 2048; g.split_k = (int)(sk < 2 ? 2 : (sk > 512 ? 512 : sk));
    int rc = launch_gemm(g, st); if (rc) return rc;
  }
  return NSOS_OK;
}
int bgrad(const float* dY, int64_t ldy, int N, float* db, int64_t P, cudaStream_t st) {
  dim3 grid((unsigned)std::min<int64_t>(256, (P + 255) / 256), (N + 63) / 64);
  k_colsum<<<grid, 64, 0, st>>>(dY, ldy, P, N, db);
  NSOS_CHECK_CUDA(cudaGetLastError());
  return NSOS_OK;
}
// dX[P,K] (=|+=) dY[P,N] . W[N, k0:k0+K]   with optional relu mask from H
int dgrad(const float* dY, int64_t ldy, int N, const float* Wt, int ldw, int k0, int K, float* dX, int64_t ldx, const float* mask,
          int accumulate, int64_t P, cudaStream_t st, const TcCtx* tc = nullptr) {
  if (N <= 4 && K % 4 == 0 && ldx % 4 == 0) {        // narrow head: elementwise outer product
    k_small_dgrad<<<grid1(P * (K / 4)), 256, 0, st>>>(dY, ldy, N, Wt + k0, ldw, K, dX, ldx, mask, accumulate, P);
    NSOS_CHECK_CUDA(cudaGetLastError());
    return NSOS_OK;
  }
  GemmArgs g{};
  g.A1 = dY; g.a1_rs = ldy; g.a1_cs = 1; g.K1 = N; g.a1_rowdiv = 1; g.a2_rowdiv = 1;
  g.B = Wt + k0; g.b_rs = ldw; g.b_cs = 1; g.b_rowdiv = 1;   // B(k=n_out, n=k_in) = W[n_out*ldw + k0 + k_in]
  g.C = dX; g.c_rs = ldx; g.c_cs = 1; g.M = (int)P; g.N = K; g.mask = mask; g.mask_ld = ldx; g.accumulate = accumulate;
  g.split_k = 1;
  return gemm_dispatch(g, tc, st);
}

struct BwdBufs {
  float* g_raw;  // [P,C]
  float* G[2];   // [P,W] ping-pong grads w.r.t. post-ReLU trunk activations
  float* g_half; // [P,W/2] grads w.r.t. hv / s0
  float* g_feat; // [P,W]
  TcCtx tc;      // tensor-core dispatch of the large products (all-parameter backward in a tcgen05 mode)
};

// Backward of mlp_forward given g_raw; accumulates into grads (flat layout).  trunk=0: semantic head only.
int mlp_backward(const NetGeom& g, const float* prm, float* grads, const float* enc, const float* encv, int S, int64_t P,
                 const MlpBufs& b, const BwdBufs& w, int trunk, cudaStream_t st) {
  const int W = g.W, H = g.W / 2, C = g.C;
  const float* hl = b.h[g.D - 1];
  int rc;
  if (!g.use_viewdirs) {
    if (!trunk) return NSOS_OK;
    if ((rc = wgrad(w.g_raw, C, 4, hl, W, W, nullptr, 0, 0, 1, grads + g.w_out, P, st, &w.tc, grads + g.b_out))) return rc;
    
    if ((rc = dgrad(w.g_raw, C, 4, prm + g.w_out, W, 0, W, w.G[0], W, hl, 0, P, st, &w.tc))) return rc;
  } else {
    bool have_G = false;
    if (g.use_sem) {
      // sem = W_s2 . s0 + b ; s0 = relu(W_s0 . [h, enc] + b)
      if ((rc = wgrad(w.g_raw + 4, C, g.sem_dim, b.s0, H, H, nullptr, 0, 0, 1, grads + g.w_s2, P, st, &w.tc, grads + g.b_s2))) return rc;
      
      if ((rc = dgrad(w.g_raw + 4, C, g.sem_dim, prm + g.w_s2, H, 0, H, w.g_half, H, b.s0, 0, P, st, &w.tc))) return rc;
      if ((rc = wgrad(w.g_half, H, H, hl, W, W, g.sem_coord ? enc : nullptr, kEncLd, g.sem_coord ? g.enc : 0, 1,
                      grads + g.w_s0, P, st, &w.tc, grads + g.b_s0))) return rc;
      
      if (trunk) {
        if ((rc = dgrad(w.g_half, H, H, prm + g.w_s0, g.sem_in, 0, W, w.G[0], W, nullptr, 0, P, st, &w.tc))) return rc;
        have_G = true;
      }
    }
    if (!trunk) return NSOS_OK;
    // rgb = W_rgb . hv + b ; hv = relu(W_v . [feat, encv] + b) ; feat = W_f . h + b
    if ((rc = wgrad(w.g_raw, C, 3, b.hv, H, H, nullptr, 0, 0, 1, grads + g.w_rgb, P, st, &w.tc, grads + g.b_rgb))) return rc;
    
    if ((rc = dgrad(w.g_raw, C, 3, prm + g.w_rgb, H, 0, H, w.g_half, H, b.hv, 0, P, st, &w.tc))) return rc;
    if ((rc = wgrad(w.g_half, H, H, b.feat, W, W, encv, kEncVLd, g.encv, S, grads + g.w_views, P, st, &w.tc, grads + g.b_views))) return rc;
    
    if ((rc = dgrad(w.g_half, H, H, prm + g.w_views, W + g.encv, 0, W, w.g_feat, W, nullptr, 0, P, st, &w.tc))) return rc;
    if ((rc = wgrad(w.g_feat, W, W, hl, W, W, nullptr, 0, 0, 1, grads + g.w_feat, P, st, &w.tc, grads + g.b_feat))) return rc;
    
    if ((rc = dgrad(w.g_feat, W, W, prm + g.w_feat, W, 0, W, w.G[0], W, nullptr, have_G ? 1 : 0, P, st, &w.tc))) return rc;
    // alpha = w_a . h + b
    if ((rc = wgrad(w.g_raw + 3, C, 1, hl, W, W, nullptr, 0, 0, 1, grads + g.w_alpha, P, st, &w.tc, grads + g.b_alpha))) return rc;
    
    if ((rc = dgrad(w.g_raw + 3, C, 1, prm + g.w_alpha, W, 0, W, w.G[0], W, nullptr, 1, P, st, &w.tc))) return rc;
    k_relu_mask<<<grid1(P * W), 256, 0, st>>>(w.G[0], hl, P * W);
    NSOS_CHECK_CUDA(cudaGetLastError());
  }
  // trunk: G[cur] = d(loss)/d(pre-activation of layer i), already masked
  int cur = 0;
  for (int i = g.D - 1; i >= 0; --i) {
    const float* dpre = w.G[cur];
    const float* hin = (i > 0) ? b.h[i - 1] : nullptr;
    if (i == 0) { if ((rc = wgrad(dpre, W, W, enc, kEncLd, g.enc, nullptr, 0, 0, 1, grads + g.w_pts[i], P, st, &w.tc, grads + g.b_pts[i]))) return rc; }
    else if (g.in_pts[i] == W) { if ((rc = wgrad(dpre, W, W, hin, W, W, nullptr, 0, 0, 1, grads + g.w_pts[i], P, st, &w.tc, grads + g.b_pts[i]))) return rc; }
    else { if ((rc = wgrad(dpre, W, W, enc, kEncLd, g.enc, hin, W, W, 1, grads + g.w_pts[i], P, st, &w.tc, grads + g.b_pts[i]))) return rc; }
    
    if (i > 0) {
      int k0 = (g.in_pts[i] == W) ? 0 : g.enc;   // h part of [enc, h]
      if ((rc = dgrad(dpre, W, W, prm + g.w_pts[i], g.in_pts[i], k0, W, w.G[cur ^ 1], W, hin, 0, P, st, &w.tc))) return rc;
      cur ^= 1;
    }
  }
  return NSOS_OK;
}

struct Carver {
  char* base; size_t off, cap;
  template <typename T> T* take(size_t n) {
    off = align_up(off, 256);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

int64_t fwd_chunk_rays(int64_t n_rays) { return n_rays < 8192 ? n_rays : 8192; }
int64_t bwd_chunk_rays(int64_t n_rays) { return n_rays < 2048 ? n_rays : 2048; }
// tensor-core all-parameter backward: 11 KB of saved activations / gradients per sample point -> 23 GB per 8192 rays x 256 points
int64_t bwd_all_chunk_rays(int64_t n_rays) {
  const char* e = getenv("NSOS_BWD_CHUNK");
  const int64_t c = e ? atoll(e) : 8192;
  return n_rays < c ? n_rays : (c < 64 ? 64 : c);
}

struct FwdWs {
  float *z0, *z1, *enc, *encv, *dnorm, *raw, *w0, *hA, *hB, *feat, *hv, *s0;
};
size_t carve_fwd(const NsosRenderCfg& cfg, const NetGeom& gc, const NetGeom& gf, int64_t R, char* base, FwdWs* ws) {
  int Sc = cfg.n_samples, Sf = cfg.n_samples + cfg.n_importance;
  int64_t P = R * (cfg.n_importance > 0 ? Sf : Sc);
  int Wm = std::max(gc.W, gf.W), Cm = std::max(gc.C, gf.C);
  Carver c{base, 0, 0};
  FwdWs w;
  w.z0 = c.take<float>(R * Sc); w.z1 = c.take<float>(R * Sf);
  w.enc = c.take<float>(P * kEncLd); w.encv = c.take<float>(R * kEncVLd); w.dnorm = c.take<float>(R);
  w.raw = c.take<float>(P * Cm); w.w0 = c.take<float>(R * Sc);
  w.hA = c.take<float>(P * Wm); w.hB = c.take<float>(P * Wm); w.feat = c.take<float>(P * Wm);
  w.hv = c.take<float>(P * (Wm / 2 + 1)); w.s0 = c.take<float>(P * (Wm / 2 + 1));
  if (ws) *ws = w;
  return align_up(c.off, 256);
}

}  // namespace

// ---- public (library-internal) entry points ------------------------------------------------------
size_t simt_render_workspace_bytes(const NsosRenderCfg& cfg, int64_t n_rays) {
  NetGeom gc, gf;
  if (!make_geom(cfg.coarse, gc)) return 0;
  if (cfg.n_importance > 0) { if (!make_geom(cfg.fine, gf)) return 0; } else gf = gc;
  return carve_fwd(cfg, gc, gf, fwd_chunk_rays(n_rays), nullptr, nullptr);
}

int simt_render_fwd(const NsosRenderCfg& cfg, const float* pc, const float* pf, const float* rays_o, const float* rays_d,
                    const float* near, const float* far, const NsosRandoms* rnd, uint64_t seed, const NsosRenderOut& out,
                    void* workspace, size_t workspace_bytes, int64_t n_rays, cudaStream_t st) {
  NetGeom gc, gf;
  NSOS_REQUIRE(make_geom(cfg.coarse, gc), NSOS_ERR_UNSUPPORTED, "invalid coarse net descriptor");
  const bool fine = cfg.n_importance > 0;
  if (fine) NSOS_REQUIRE(make_geom(cfg.fine, gf), NSOS_ERR_UNSUPPORTED, "invalid fine net descriptor"); else gf = gc;
  const int Sc = cfg.n_samples, K = cfg.n_importance, Sf = Sc + K;
  NSOS_REQUIRE(Sc >= 2 && Sc <= kMaxS && Sf <= kMaxS, NSOS_ERR_UNSUPPORTED, "n_samples/n_importance out of range (<=%d total)", kMaxS);
  NSOS_REQUIRE(gc.enc <= kEncLd && gc.encv <= kEncVLd && gf.enc <= kEncLd && gf.encv <= kEncVLd, NSOS_ERR_UNSUPPORTED,
               "positional encodings wider than %d / %d columns are not implemented (multires <= 10, multires_views <= 4)", kEncLd, kEncVLd);
  NSOS_REQUIRE(!fine || gc.C == gf.C, NSOS_ERR_UNSUPPORTED, "coarse and fine nets must have the same output channels");
  const int64_t R = fwd_chunk_rays(n_rays);
  FwdWs w;
  size_t need = carve_fwd(cfg, gc, gf, R, (char*)workspace, &w);
  NSOS_REQUIRE(workspace_bytes >= need, NSOS_ERR_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, need);
  const int C6 = 6 + gc.sem_dim, ML = 2 * C6 + 1;
  NsosRandoms rn{nullptr, nullptr, nullptr, nullptr, nullptr};
  if (rnd) rn = *rnd;
  NSOS_CHECK_CUDA(cudaMemsetAsync(out.maps, 0, sizeof(float) * n_rays * ML, st));

  for (int64_t r0 = 0; r0 < n_rays; r0 += R) {
    const int64_t n = std::min(R, n_rays - r0);
    const float* ro = rays_o + r0 * 3; const float* rd = rays_d + r0 * 3;
    float* maps = out.maps + r0 * ML;
    MlpBufs b{};
    // ---- coarse pass (nerf_net.py:93-99)
    k_coarse_z<<<grid1(n * Sc), 256, 0, st>>>(near + r0, far + r0, rn.t_rand ? rn.t_rand + r0 * Sc : nullptr, cfg.perturb, seed, r0,
                                               w.z0, n, Sc);
    k_encode_dirs<<<grid1(n), 256, 0, st>>>(rd, w.encv, w.dnorm, n, gc.Lv);
    k_encode_pts<<<grid1(n * Sc), 256, 0, st>>>(ro, rd, w.z0, w.enc, n * Sc, Sc, gc.Lp);
    NSOS_CHECK_CUDA(cudaGetLastError());
    for (int i = 0; i < gc.D; ++i) b.h[i] = (i & 1) ? w.hB : w.hA;
    b.feat = w.feat; b.hv = w.hv; b.s0 = w.s0;
    float* raw0 = out.raw0 ? out.raw0 + r0 * Sc * gc.C : w.raw;
    if (!fine && out.raw) raw0 = out.raw + r0 * Sc * gc.C;
    int rc = mlp_forward(gc, pc, w.enc, w.encv, Sc, n * Sc, b, raw0, st);
    if (rc) return rc;
    float* wts0 = fine ? (out.weights0 ? out.weights0 + r0 * Sc : w.w0) : (out.weights ? out.weights + r0 * Sc : w.w0);
    k_composite<<<grid1(n * 32, 128), 128, 0, st>>>(raw0, w.z0, w.dnorm, rn.noise0 ? rn.noise0 + r0 * Sc : nullptr, cfg.raw_noise_std,
                                                     seed, r0, RNG_NOISE0, Sc, gc.C, gc.sem_dim, cfg.white_bkgd, maps, ML,
                                                     fine ? C6 : 0, wts0, n);
    NSOS_CHECK_CUDA(cudaGetLastError());
    float* zout0 = fine ? out.z_vals0 : out.z_vals;
    if (zout0) NSOS_CHECK_CUDA(cudaMemcpyAsync(zout0 + r0 * Sc, w.z0, sizeof(float) * n * Sc, cudaMemcpyDeviceToDevice, st));
    if (!fine) continue;
    // ---- importance resampling + fine pass (nerf_net.py:104-128)
    float* z1 = out.z_vals ? out.z_vals + r0 * Sf : w.z1;
    size_t smem = 4 * sizeof(float) * (5 * Sc + K);
    k_importance<<<grid1(n * 32, 128), 128, smem, st>>>(w.z0, wts0, rn.u ? rn.u + r0 * K : nullptr, rn.z_samples ? rn.z_samples + r0 * K : nullptr,
                                                         cfg.perturb, seed, r0, Sc, K, z1,
                                                         out.z_samples ? out.z_samples + r0 * K : nullptr,
                                                         out.inds ? out.inds + r0 * K : nullptr, maps + 2 * C6, ML, n);
    k_encode_pts<<<grid1(n * Sf), 256, 0, st>>>(ro, rd, z1, w.enc, n * Sf, Sf, gf.Lp);
    NSOS_CHECK_CUDA(cudaGetLastError());
    for (int i = 0; i < gf.D; ++i) b.h[i] = (i & 1) ? w.hB : w.hA;
    float* raw1 = out.raw ? out.raw + r0 * Sf * gf.C : w.raw;
    rc = mlp_forward(gf, pf, w.enc, w.encv, Sf, n * Sf, b, raw1, st);
    if (rc) return rc;
    k_composite<<<grid1(n * 32, 128), 128, 0, st>>>(raw1, z1, w.dnorm, rn.noise1 ? rn.noise1 + r0 * Sf : nullptr, cfg.raw_noise_std, seed,
                                                     r0, RNG_NOISE1, Sf, gf.C, gf.sem_dim, cfg.white_bkgd, maps, ML, 0,
                                                     out.weights ? out.weights + r0 * Sf : nullptr, n);
    NSOS_CHECK_CUDA(cudaGetLastError());
  }
  return NSOS_OK;
}

// ---- backward --------------------------------------------------------------------------------------
namespace {
struct BwdWs {
  float *enc, *encv, *dnorm, *raw, *g_raw, *h[16], *feat, *hv, *s0, *G0, *G1, *g_half, *g_feat;
  uint8_t* tc_scratch;
};
constexpr size_t kTcScratch = 2 * 256 * 256 * 2 * 2 + 4096;      // tc_rowgemm_scratch_bytes(256, 256) and some
size_t carve_bwd(const NsosRenderCfg& cfg, const NetGeom& gc, const NetGeom& gf, int64_t R, char* base, BwdWs* ws) {
  int Sc = cfg.n_samples, Sf = cfg.n_samples + cfg.n_importance;
  int64_t P = R * (cfg.n_importance > 0 ? Sf : Sc);
  int Wm = std::max(gc.W, gf.W), Cm = std::max(gc.C, gf.C), Dm = std::max(gc.D, gf.D);
  Carver c{base, 0, 0};
  BwdWs w{};
  w.enc = c.take<float>(P * kEncLd); w.encv = c.take<float>(R * kEncVLd); w.dnorm = c.take<float>(R);
  w.raw = c.take<float>(P * Cm); w.g_raw = c.take<float>(P * Cm);
  for (int i = 0; i < Dm; ++i) w.h[i] = c.take<float>(P * Wm);
  w.feat = c.take<float>(P * Wm); w.hv = c.take<float>(P * (Wm / 2 + 1)); w.s0 = c.take<float>(P * (Wm / 2 + 1));
  w.G0 = c.take<float>(P * Wm); w.G1 = c.take<float>(P * Wm); w.g_half = c.take<float>(P * (Wm / 2 + 1));
  w.g_feat = c.take<float>(P * Wm);
  w.tc_scratch = c.take<uint8_t>(kTcScratch);
  if (ws) *ws = w;
  return align_up(c.off, 256);
}
}  // namespace

// All-parameter backward in a tcgen05 mode: the recompute is ONE replay launch of kernel A per chunk that saves every hidden
// layer of both nets (fp16 hi/lo arithmetic: the same activations, ReLU masks and raw outputs as the forward call), and the
// large dgrad / wgrad products run on tc_rowgemm / tc_wgrad_gen.
namespace {
struct BwdAllWs {
  float *enc, *encv, *dnorm, *g_raw, *G0, *G1, *g_half, *g_feat, *feat;
  float *raw[2], *h[2][16], *hv[2], *s0[2];
  uint8_t *scratch, *part;     // rowgemm operand image; partial sums of the weight-gradient kernels
};
size_t carve_bwd_all(const NsosRenderCfg& cfg, const NetGeom& gc, const NetGeom& gf, int64_t R, char* base, BwdAllWs* ws) {
  const bool fine = cfg.n_importance > 0;
  const int Sc = cfg.n_samples, Sf = cfg.n_samples + cfg.n_importance;
  const int64_t Pp[2] = {R * Sc, fine ? R * Sf : 0}, Pm = std::max(Pp[0], Pp[1]);
  const int Wm = std::max(gc.W, gf.W), Cm = std::max(gc.C, gf.C);
  Carver c{base, 0, 0};
  BwdAllWs w{};
  w.enc = c.take<float>(Pm * kEncLd); w.encv = c.take<float>(R * kEncVLd); w.dnorm = c.take<float>(R);
  w.g_raw = c.take<float>(Pm * Cm);
  w.G0 = c.take<float>(Pm * Wm); w.G1 = c.take<float>(Pm * Wm); w.g_half = c.take<float>(Pm * (Wm / 2 + 1));
  w.g_feat = c.take<float>(Pm * Wm); w.feat = c.take<float>(Pm * Wm);
  for (int p = 0; p < 2; ++p) {
    const NetGeom& g = p ? gf : gc;
    w.raw[p] = c.take<float>(Pp[p] * g.C + 1);
    for (int i = 0; i < g.D; ++i) w.h[p][i] = c.take<float>(Pp[p] * g.W + 1);
    w.hv[p] = c.take<float>(Pp[p] * (g.W / 2) + 1); w.s0[p] = c.take<float>(Pp[p] * (g.W / 2) + 1);
  }
  w.scratch = c.take<uint8_t>(kTcScratch);
  w.part = c.take<uint8_t>(tc_wgrad_part_bytes());
  if (ws) *ws = w;
  return align_up(c.off, 256);
}
bool bwd_all_uses_tc(const NsosRenderCfg& cfg, int trunk, const NetGeom& gc, const NetGeom& gf, const void* packed_c, const void* packed_f) {
  if (!trunk || !(cfg.mode == NSOS_MODE_TC_EXACT || cfg.mode == NSOS_MODE_TC_FAST) || getenv("NSOS_BWD_SIMT")) return false;
  if (!packed_c || (cfg.n_importance > 0 && !packed_f) || cfg.n_samples > 128) return false;
  if (gc.W != 256 || gf.W != 256 || !gc.use_viewdirs || !gf.use_viewdirs) return false;
  return tc_net_supported(cfg.coarse) && (cfg.n_importance == 0 || tc_net_supported(cfg.fine));
}
}  // namespace

// Semantic-head-only backward with the trunk recomputed on tensor cores (tc_render_replay): per chunk the replay kernel
// leaves h_last / s0 / raw per point, the GEMMs below only touch the 4 semantic_linear tensors.
namespace {
int64_t bwd_tc_chunk_rays(int64_t n_rays) { return n_rays < 4096 ? n_rays : 4096; }
struct BwdTcWs {
  float *enc, *encv, *dnorm, *raw[2], *g_raw, *h[2], *s0[2], *g_half;
  uint8_t* part;      // partial sums of the weight-gradient kernel
};
size_t carve_bwd_tc(const NsosRenderCfg& cfg, const NetGeom& gc, const NetGeom& gf, int64_t R, char* base, BwdTcWs* ws) {
  const bool fine = cfg.n_importance > 0;
  const int Sc = cfg.n_samples, Sf = cfg.n_samples + cfg.n_importance;
  const int64_t P0 = R * Sc, P1 = fine ? R * Sf : 0, Pm = std::max(P0, P1);
  Carver c{base, 0, 0};
  BwdTcWs w{};
  w.enc = c.take<float>(Pm * kEncLd); w.encv = c.take<float>(R * kEncVLd); w.dnorm = c.take<float>(R);
  w.raw[0] = c.take<float>(P0 * gc.C); w.raw[1] = c.take<float>(P1 * gf.C + 1);
  w.g_raw = c.take<float>(Pm * std::max(gc.C, gf.C));
  // (+32 points: the blocked layout of the replayed activations rounds the point count up to whole groups)
  w.h[0] = c.take<float>((P0 + 32) * gc.W); w.s0[0] = c.take<float>((P0 + 32) * (gc.W / 2));
  w.h[1] = c.take<float>((P1 + 32) * gf.W + 1); w.s0[1] = c.take<float>((P1 + 32) * (gf.W / 2) + 1);
  w.g_half = c.take<float>(Pm * (std::max(gc.W, gf.W) / 2 + 1));
  w.part = c.take<uint8_t>(tc_wgrad_part_bytes());
  if (ws) *ws = w;
  return align_up(c.off, 256);
}
bool bwd_uses_tc(const NsosRenderCfg& cfg, int trunk, const NetGeom& gc, const NetGeom& gf) {
  if (trunk || !(cfg.mode == NSOS_MODE_TC_EXACT || cfg.mode == NSOS_MODE_TC_FAST)) return false;
  if (getenv("NSOS_BWD_SIMT")) return false;   // parity tests: force the all-fp32 recompute
  if (!gc.use_sem || !gf.use_sem || cfg.n_samples > 128) return false;
  return tc_net_supported(cfg.coarse) && (cfg.n_importance == 0 || tc_net_supported(cfg.fine));
}
}  // namespace

size_t simt_render_bwd_workspace_bytes(const NsosRenderCfg& cfg, int64_t n_rays, int trunk) {
  NetGeom gc, gf;
  if (!make_geom(cfg.coarse, gc)) return 0;
  if (cfg.n_importance > 0) { if (!make_geom(cfg.fine, gf)) return 0; } else gf = gc;
  if (trunk && (cfg.mode == NSOS_MODE_TC_EXACT || cfg.mode == NSOS_MODE_TC_FAST))        // either all-parameter path may run
    return std::max(carve_bwd_all(cfg, gc, gf, bwd_all_chunk_rays(n_rays), nullptr, nullptr),
                    carve_bwd(cfg, gc, gf, bwd_chunk_rays(n_rays), nullptr, nullptr));
  if (bwd_uses_tc(cfg, trunk, gc, gf)) return carve_bwd_tc(cfg, gc, gf, bwd_tc_chunk_rays(n_rays), nullptr, nullptr);
  return carve_bwd(cfg, gc, gf, bwd_chunk_rays(n_rays), nullptr, nullptr);
}

int simt_render_bwd(const NsosRenderCfg& cfg, const float* pc, const float* pf, const float* rays_o, const float* rays_d,
                    const float* z_vals0, const float* z_vals, const NsosRandoms* rnd, uint64_t seed, const float* g_maps,
                    float* grads_c, float* grads_f, int trunk, const void* packed_c, const void* packed_f,
                    const NsosRenderOut* saved, void* workspace, size_t workspace_bytes, int64_t n_rays, cudaStream_t st) {
  NetGeom gc, gf;
  NSOS_REQUIRE(make_geom(cfg.coarse, gc), NSOS_ERR_UNSUPPORTED, "invalid coarse net descriptor");
  const bool fine = cfg.n_importance > 0;
  if (fine) NSOS_REQUIRE(make_geom(cfg.fine, gf), NSOS_ERR_UNSUPPORTED, "invalid fine net descriptor"); else gf = gc;
  const int Sc = cfg.n_samples, K = cfg.n_importance, Sf = Sc + K;
  NSOS_REQUIRE(Sc >= 2 && Sf <= kMaxS, NSOS_ERR_UNSUPPORTED, "n_samples/n_importance out of range");
  NSOS_REQUIRE(gc.enc <= kEncLd && gc.encv <= kEncVLd && gf.enc <= kEncLd && gf.encv <= kEncVLd, NSOS_ERR_UNSUPPORTED,
               "positional encodings wider than %d / %d columns are not implemented (multires <= 10, multires_views <= 4)", kEncLd, kEncVLd);
  if (!trunk && !gc.use_sem && !gf.use_sem) return NSOS_OK;   // nothing trainable outside the trunk
  if (bwd_uses_tc(cfg, trunk, gc, gf)) {
    // activations saved by the forward call make the trunk replay unnecessary
    const float* sv_raw[2] = {nullptr, nullptr}; const float* sv_h[2] = {nullptr, nullptr}; const float* sv_s[2] = {nullptr, nullptr};
    const float* sv_e[2] = {nullptr, nullptr};
    if (saved) {
      if (fine) { sv_raw[0] = saved->raw0; sv_h[0] = saved->h_last0; sv_s[0] = saved->s_hid0; sv_e[0] = saved->enc0;
                  sv_raw[1] = saved->raw; sv_h[1] = saved->h_last; sv_s[1] = saved->s_hid; sv_e[1] = saved->enc; }
      else { sv_raw[0] = saved->raw; sv_h[0] = saved->h_last; sv_s[0] = saved->s_hid; sv_e[0] = saved->enc; }
    }
    const bool have_saved = sv_raw[0] && sv_h[0] && sv_s[0] && (!fine || (sv_raw[1] && sv_h[1] && sv_s[1]));
    if (!have_saved)
      NSOS_REQUIRE(packed_c && (!fine || packed_f), NSOS_ERR_BAD_ARG, "nsos_render_bwd: tcgen05 modes need the packed weights");
    const int64_t R = bwd_tc_chunk_rays(n_rays);
    BwdTcWs w;
    size_t need = carve_bwd_tc(cfg, gc, gf, R, (char*)workspace, &w);
    NSOS_REQUIRE(workspace_bytes >= need, NSOS_ERR_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, need);
    const int C6 = 6 + gc.sem_dim, ML = 2 * C6 + 1;
    NsosRandoms rn{nullptr, nullptr, nullptr, nullptr, nullptr};
    if (rnd) rn = *rnd;
    for (int64_t r0 = 0; r0 < n_rays; r0 += R) {
      const int64_t n = std::min(R, n_rays - r0);
      const float* ro = rays_o + r0 * 3; const float* rd = rays_d + r0 * 3;
      const float* zc = (fine ? z_vals0 : z_vals) + r0 * Sc;
      const float* zf = fine ? z_vals + r0 * Sf : nullptr;
      int rc = NSOS_OK;
      const float* raw_p[2] = {w.raw[0], w.raw[1]}; const float* h_p[2] = {w.h[0], w.h[1]}; const float* s_p[2] = {w.s0[0], w.s0[1]};
      if (have_saved) {
        raw_p[0] = sv_raw[0] + r0 * Sc * gc.C; h_p[0] = sv_h[0] + r0 * Sc * gc.W; s_p[0] = sv_s[0] + r0 * Sc * (gc.W / 2);
        if (fine) { raw_p[1] = sv_raw[1] + r0 * Sf * gf.C; h_p[1] = sv_h[1] + r0 * Sf * gf.W; s_p[1] = sv_s[1] + r0 * Sf * (gf.W / 2); }
      } else {
        rc = tc_render_replay(cfg, packed_c, fine ? packed_f : packed_c, ro, rd, zc, zf, w.raw[0], w.raw[1], w.h[0], w.s0[0], w.h[1],
                              w.s0[1], n, st);
      }
      if (rc) return rc;
      k_encode_dirs<<<grid1(n), 256, 0, st>>>(rd, w.encv, w.dnorm, n, gc.Lv);
      for (int pass = 0; pass < (fine ? 2 : 1); ++pass) {
        const bool is_fine = fine && pass == 1;
        const NetGeom& g = is_fine ? gf : gc;
        const int S = is_fine ? Sf : Sc;
        const float* z = is_fine ? zf : zc;
        const float* noise = is_fine ? rn.noise1 : rn.noise0;
        const int moff = (fine && !is_fine) ? C6 : 0;
        const int64_t P = n * S;
        // gamma(x): saved by the training forward (same 64-float row pitch), else recomputed here
        const float* enc_p = (have_saved && sv_e[pass]) ? sv_e[pass] + r0 * S * kEncLd : w.enc;
        if (enc_p == w.enc && g.sem_coord) k_encode_pts<<<grid1(P), 256, 0, st>>>(ro, rd, z, w.enc, P, S, g.Lp);
        k_composite_bwd<<<grid1(n * 32, 128), 128, 0, st>>>(raw_p[pass], z, w.dnorm, noise ? noise + r0 * S : nullptr, cfg.raw_noise_std,
                                                             seed, r0, is_fine ? RNG_NOISE1 : RNG_NOISE0, S, g.C, g.sem_dim,
                                                             cfg.white_bkgd, g_maps + r0 * ML, ML, moff, w.g_raw, n);
        NSOS_CHECK_CUDA(cudaGetLastError());
        if (sem_saves_blocked(gc, gf)) {  // h / s0 (and a saved gamma) are in the blocked layout then
          rc = tc_sem_wgrad(g, is_fine ? pf : pc, is_fine ? grads_f : grads_c, h_p[pass], enc_p, kEncLd, enc_p != w.enc, s_p[pass],
                            w.g_raw, P, w.part, tc_wgrad_part_bytes(), st);
        } else {
          MlpBufs b{};
          b.h[g.D - 1] = const_cast<float*>(h_p[pass]); b.s0 = const_cast<float*>(s_p[pass]);
          BwdBufs bw{w.g_raw, {nullptr, nullptr}, w.g_half, nullptr, TcCtx{}};
          rc = mlp_backward(g, is_fine ? pf : pc, is_fine ? grads_f : grads_c, enc_p, w.encv, S, P, b, bw, 0, st);
        }
        if (rc) return rc;
      }
    }
    return NSOS_OK;
  }
  if (bwd_all_uses_tc(cfg, trunk, gc, gf, packed_c, packed_f)) {
    const int64_t R = bwd_all_chunk_rays(n_rays);
    BwdAllWs w;
    size_t need = carve_bwd_all(cfg, gc, gf, R, (char*)workspace, &w);
    NSOS_REQUIRE(workspace_bytes >= need, NSOS_ERR_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, need);
    const int C6 = 6 + gc.sem_dim, ML = 2 * C6 + 1;
    NsosRandoms rn{nullptr, nullptr, nullptr, nullptr, nullptr};
    if (rnd) rn = *rnd;
    TcCtx tcx;
    tcx.on = true; tcx.scratch = w.scratch; tcx.scratch_bytes = kTcScratch; tcx.part = w.part; tcx.part_bytes = tc_wgrad_part_bytes();
    for (int64_t r0 = 0; r0 < n_rays; r0 += R) {
      const int64_t n = std::min(R, n_rays - r0);
      const float* ro = rays_o + r0 * 3; const float* rd = rays_d + r0 * 3;
      const float* zc = (fine ? z_vals0 : z_vals) + r0 * Sc;
      const float* zf = fine ? z_vals + r0 * Sf : nullptr;
      float* const raw2[2] = {w.raw[0], w.raw[1]};
      float* const* const hall[2] = {w.h[0], w.h[1]};
      float* const hv2[2] = {w.hv[0], w.hv[1]};
      float* const s02[2] = {w.s0[0], w.s0[1]};
      int rc = tc_render_replay_all(cfg, packed_c, fine ? packed_f : packed_c, ro, rd, zc, zf, raw2, hall, hv2, s02, n, st);
      if (rc) return rc;
      k_encode_dirs<<<grid1(n), 256, 0, st>>>(rd, w.encv, w.dnorm, n, gc.Lv);
      for (int pass = 0; pass < (fine ? 2 : 1); ++pass) {
        const bool is_fine = fine && pass == 1;
        const NetGeom& g = is_fine ? gf : gc;
        const float* prm = is_fine ? pf : pc;
        float* grads = is_fine ? grads_f : grads_c;
        const int S = is_fine ? Sf : Sc;
        const float* z = is_fine ? zf : zc;
        const float* noise = is_fine ? rn.noise1 : rn.noise0;
        const int moff = (fine && !is_fine) ? C6 : 0;
        const int64_t P = n * S;
        k_encode_pts<<<grid1(P), 256, 0, st>>>(ro, rd, z, w.enc, P, S, g.Lp);
        k_composite_bwd<<<grid1(n * 32, 128), 128, 0, st>>>(w.raw[pass], z, w.dnorm, noise ? noise + r0 * S : nullptr, cfg.raw_noise_std,
                                                             seed, r0, is_fine ? RNG_NOISE1 : RNG_NOISE0, S, g.C, g.sem_dim,
                                                             cfg.white_bkgd, g_maps + r0 * ML, ML, moff, w.g_raw, n);
        NSOS_CHECK_CUDA(cudaGetLastError());
        MlpBufs b{};
        for (int i = 0; i < g.D; ++i) b.h[i] = w.h[pass][i];
        b.feat = w.feat; b.hv = w.hv[pass]; b.s0 = w.s0[pass];
        // feature_linear's output is folded away in the forward kernel (views fusion); the views weight gradient needs it
        rc = gemm_dispatch(lin(b.h[g.D - 1], g.W, g.W, nullptr, 0, 0, 1, prm + g.w_feat, prm + g.b_feat, b.feat, g.W, P, g.W, 0), &tcx, st);
        if (rc) return rc;
        BwdBufs bw{w.g_raw, {w.G0, w.G1}, w.g_half, w.g_feat, tcx};
        rc = mlp_backward(g, prm, grads, w.enc, w.encv, S, P, b, bw, trunk, st);
        if (rc) return rc;
      }
    }
    return NSOS_OK;
  }
  const int64_t R = bwd_chunk_rays(n_rays);
  BwdWs w;
  size_t need = carve_bwd(cfg, gc, gf, R, (char*)workspace, &w);
  NSOS_REQUIRE(workspace_bytes >= need, NSOS_ERR_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, need);
  const int C6 = 6 + gc.sem_dim, ML = 2 * C6 + 1;
  NsosRandoms rn{nullptr, nullptr, nullptr, nullptr, nullptr};
  if (rnd) rn = *rnd;
  for (int64_t r0 = 0; r0 < n_rays; r0 += R) {
    const int64_t n = std::min(R, n_rays - r0);
    const float* ro = rays_o + r0 * 3; const float* rd = rays_d + r0 * 3;
    k_encode_dirs<<<grid1(n), 256, 0, st>>>(rd, w.encv, w.dnorm, n, gc.Lv);
    for (int pass = 0; pass < (fine ? 2 : 1); ++pass) {
      const bool is_fine = fine && pass == 1;
      const NetGeom& g = is_fine ? gf : gc;
      const float* prm = is_fine ? pf : pc;
      float* grads = is_fine ? grads_f : grads_c;
      const int S = is_fine ? Sf : Sc;
      const float* z = (is_fine ? z_vals : (fine ? z_vals0 : z_vals)) + r0 * S;
      const float* noise = is_fine ? rn.noise1 : rn.noise0;
      const int moff = (fine && !is_fine) ? C6 : 0;
      const int64_t P = n * S;
      k_encode_pts<<<grid1(P), 256, 0, st>>>(ro, rd, z, w.enc, P, S, g.Lp);
      NSOS_CHECK_CUDA(cudaGetLastError());
      MlpBufs b{};
      for (int i = 0; i < g.D; ++i) b.h[i] = w.h[i];
      b.feat = w.feat; b.hv = w.hv; b.s0 = w.s0;
      // all-parameter backward in a tcgen05 mode: the large GEMMs of the recompute, of dgrad and of wgrad run on the tensor cores
      TcCtx tcx;
      tcx.on = false;                     // reached only in NSOS_MODE_SIMT_FP32 / NSOS_BWD_SIMT / unsupported geometry: all fp32
      tcx.scratch = w.tc_scratch; tcx.scratch_bytes = kTcScratch;
      int rc = mlp_forward(g, prm, w.enc, w.encv, S, P, b, w.raw, st, &tcx);
      if (rc) return rc;
      k_composite_bwd<<<grid1(n * 32, 128), 128, 0, st>>>(w.raw, z, w.dnorm, noise ? noise + r0 * S : nullptr, cfg.raw_noise_std, seed,
                                                           r0, is_fine ? RNG_NOISE1 : RNG_NOISE0, S, g.C, g.sem_dim, cfg.white_bkgd,
                                                           g_maps + r0 * ML, ML, moff, w.g_raw, n);
      NSOS_CHECK_CUDA(cudaGetLastError());
      BwdBufs bw{w.g_raw, {w.G0, w.G1}, w.g_half, w.g_feat, tcx};
      rc = mlp_backward(g, prm, grads, w.enc, w.encv, S, P, b, bw, trunk, st);
      if (rc) return rc;
    }
  }
  return NSOS_OK;
}

// ---- stage-wise and raw-MLP entry points ------------------------------------------------------------
int simt_invert_cdf(const float* bins, const float* cdf, const float* u, float* samples, int64_t* inds, int64_t n_rays, int M, int K,
                    cudaStream_t st) {
  k_invert_cdf<<<grid1(n_rays * K), 256, 0, st>>>(bins, cdf, u, samples, inds, n_rays, M, K);
  NSOS_CHECK_CUDA(cudaGetLastError());
  return NSOS_OK;
}

size_t simt_mlp_workspace_bytes(const NetGeom& g, int64_t P) {
  Carver c{nullptr, 0, 0};
  c.take<float>(P * kEncLd); c.take<float>(P * kEncVLd);
  c.take<float>(P * g.W); c.take<float>(P * g.W); c.take<float>(P * g.W);
  c.take<float>(P * (g.W / 2 + 1)); c.take<float>(P * (g.W / 2 + 1));
  return align_up(c.off, 256);
}

int simt_mlp_query(const NetGeom& g, const float* prm, const float* pts, const float* viewdirs, float* raw, void* workspace,
                   size_t workspace_bytes, int64_t P, cudaStream_t st) {
  NSOS_REQUIRE(g.enc <= kEncLd && g.encv <= kEncVLd, NSOS_ERR_UNSUPPORTED,
               "positional encodings wider than %d / %d columns are not implemented (multires <= 10, multires_views <= 4)", kEncLd, kEncVLd);
  NSOS_REQUIRE(workspace_bytes >= simt_mlp_workspace_bytes(g, P), NSOS_ERR_WORKSPACE, "workspace too small");
  Carver c{(char*)workspace, 0, 0};
  float* enc = c.take<float>(P * kEncLd); float* encv = c.take<float>(P * kEncVLd);
  float* hA = c.take<float>(P * g.W); float* hB = c.take<float>(P * g.W);
  MlpBufs b{};
  for (int i = 0; i < g.D; ++i) b.h[i] = (i & 1) ? hB : hA;
  b.feat = c.take<float>(P * g.W); b.hv = c.take<float>(P * (g.W / 2 + 1)); b.s0 = c.take<float>(P * (g.W / 2 + 1));
  k_encode_raw<<<grid1(P), 256, 0, st>>>(pts, enc, P, g.Lp, kEncLd);
  if (g.use_viewdirs) k_encode_raw<<<grid1(P), 256, 0, st>>>(viewdirs, encv, P, g.Lv, kEncVLd);
  NSOS_CHECK_CUDA(cudaGetLastError());
  return mlp_forward(g, prm, enc, encv, 1, P, b, raw, st);
}

}  // namespace nsos
