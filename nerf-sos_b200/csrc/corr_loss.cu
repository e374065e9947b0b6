// Kernel B: patch-wise correlation losses without materialising the pair matrices.
//
// Reference: utils/image.py:263-482 (CorrelationLoss / GeoCorrelationLoss).  For a pair of patches
// (first = n, second = n for the self term or neg_idx[n] for the negative term), every loss has the form
//     loss_h = mean_{n,p,q} [ -max(cd[n,p,q], 0) * (fd_c[n,p,q] - shift_h) ]
//     fd_c   = fd - rowmean_q(fd) - mean(fd - rowmean) + mean(fd)            (image.py:316-319 / :420-424)
// where mean(fd - rowmean) == 0 analytically, so fd_c = fd - rowmean[n,p] + old_mean.  fd carries no
// gradient (torch.no_grad), cd does.  The reference materialises fd/cd as [B, M, M] tensors (M = 4096 for the
// geometry loss: 537 MB each, 9.2 GB peak); here each (n, p) thread streams over q in shared-memory tiles:
//   pass 1  rowmean[h,n,p]                      (k_geo_rowmean / k_app_fd)
//   pass 2  loss + d/d(first operand)           (thread per p, loop over q)
//   pass 3  d/d(second operand)                 (thread per q, loop over p)
// so only O(B*M) floats ever touch HBM.  CUDA-core fp32 work (abs-diff / rcp / min chains), bounded by
// FP32 + MUFU issue rate, not by HBM or tensor cores.
#include "internal.h"

#include <algorithm>
#include <type_traits>

namespace nsos {
namespace {

constexpr int kT = 256;     // threads per block == tile width
constexpr int kCMax = 8;    // max code channels
constexpr float kMaxCorr = 15.f, kEpsCorr = 5e-2f;

struct GeoWs {
  float* chat;      // [B,C,M] normalised code
  float* invn;      // [B,M]   1/max(|code|,eps)  (0 where the norm is clamped: no gradient through eps)
  float* rowmean;   // [js][2,B,M] partial row means, one per split of the streamed pixels (summed in a fixed order by the readers)
  int js;           // number of partials
  float* g_chat;    // [B,C,M]
  double* acc;      // [0..1] sum of rowmeans per helper, [2..3] loss sums per helper
  float* oldmean;   // [2]
};

__device__ __forceinline__ float inv_l1(float s) { return fminf(kMaxCorr, 1.f / (s + kEpsCorr)); }   // image.py:404-413

// The pair kernels have one thread per pixel of a patch and stream the M pixels of the other patch: B*M*2 = 65 k threads for the
// shipped batch, 256 blocks for 148 SMs (1.7 waves of 8 warps per SM -- latency-bound at a sixth of the issue rate).  The streamed
// range is therefore split `js` ways over more blocks (partial sums meet in the atomics that were there already; row means are
// kept as `js` partials and summed in a fixed order where they are read, so the result does not depend on block scheduling).
constexpr int kGeoJsMax = 8, kGeoBlocksTarget = 1024;     // ~7 resident 256-thread blocks on each of the 148 SMs
__host__ inline int geo_jsplit(int blocks, int M) {
  const int ntile = (M + kT - 1) / kT;
  int js = 1;
  while (js < kGeoJsMax && blocks * js < kGeoBlocksTarget && js * 2 <= ntile) js *= 2;
  return js;
}
// [j_begin, j_end) of split jp
__device__ __forceinline__ void geo_jrange(int M, int js, int jp, int& jb, int& je) {
  const int ntile = (M + kT - 1) / kT, per = (ntile + js - 1) / js;
  jb = min(M, jp * per * kT);
  je = min(M, (jp + 1) * per * kT);
}
__device__ __forceinline__ float geo_rowmean_at(const GeoWs& w, int h, int B, int M, int n, int i) {
  float r = 0.f;
  for (int k = 0; k < w.js; ++k) r += w.rowmean[(((size_t)k * 2 + h) * B + n) * M + i];
  return r;
}

// F.normalize(code, dim=1, eps=1e-10) (image.py:300-301)
__global__ void k_normalize(const float* __restrict__ code, float* __restrict__ chat, float* __restrict__ invn, int B, int C, int M) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * M) return;
  int n = i / M, m = i % M;
  float s = 0.f;
  for (int c = 0; c < C; ++c) { float v = code[((size_t)n * C + c) * M + m]; s += v * v; }
  float nrm = sqrtf(s);
  float inv = 1.f / fmaxf(nrm, 1e-10f);
  for (int c = 0; c < C; ++c) chat[((size_t)n * C + c) * M + m] = code[((size_t)n * C + c) * M + m] * inv;
  invn[i] = (nrm > 1e-10f) ? inv : 0.f;
}

// pass 1: rowmean[h,n,p] = mean_q min(15, 1/(|X_p - X'_q|_1 + 0.05))
__global__ void __launch_bounds__(kT) k_geo_rowmean(const float* __restrict__ xyz, const int64_t* __restrict__ neg, GeoWs w, int B, int M, int q0) {
  __shared__ float sx[3][kT];
  const int n = q0 + blockIdx.y, h = blockIdx.z;
  const int nb = (h == 0) ? (int)neg[n] : n;
  const int nbi = (M + kT - 1) / kT, jp = blockIdx.x / nbi;
  const int p = (blockIdx.x % nbi) * kT + threadIdx.x;
  const bool act = p < M;
  int jb, je;
  geo_jrange(M, w.js, jp, jb, je);
  float x0 = 0, x1 = 0, x2 = 0;
  if (act) { x0 = xyz[((size_t)n * 3 + 0) * M + p]; x1 = xyz[((size_t)n * 3 + 1) * M + p]; x2 = xyz[((size_t)n * 3 + 2) * M + p]; }
  float sum = 0.f;
  for (int q0 = jb; q0 < je; q0 += kT) {
    int q = q0 + threadIdx.x;
    __syncthreads();
    for (int c = 0; c < 3; ++c) sx[c][threadIdx.x] = (q < M) ? xyz[((size_t)nb * 3 + c) * M + q] : 0.f;
    __syncthreads();
    int lim = min(kT, je - q0);
#pragma unroll 8
    for (int j = 0; j < lim; ++j) sum += inv_l1(fabsf(x0 - sx[0][j]) + fabsf(x1 - sx[1][j]) + fabsf(x2 - sx[2][j]));
  }
  float rm = sum / (float)M;
  if (act) w.rowmean[(((size_t)jp * 2 + h) * B + n) * M + p] = rm;        // this split's share of the mean
  // block partial of the sum of row means -> old_mean
  float v = act ? rm : 0.f;
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&w.acc[h], (double)v);
}

// old_mean[h] = (sum of all row means of the GLOBAL batch) / count; `sums` = w.acc on one GPU, the all-reduced partial sums
// of every rank when the query patches are sharded
__global__ void k_oldmean(GeoWs w, const double* sums, double count) {
  if (threadIdx.x < 2) w.oldmean[threadIdx.x] = (float)(sums[threadIdx.x] / count);
}
__global__ void k_export_sums(const double* acc, double* sums) {
  if (threadIdx.x < 2) sums[threadIdx.x] = acc[threadIdx.x];
}

// pass 2 (SECOND=false): thread per p of the first patch, loop over q of the second: loss + grad wrt chat[n,:,p]
// pass 3 (SECOND=true):  thread per q of the second patch, loop over p of the first: grad wrt chat[nb,:,q]
// CT = compile-time channel count (1..4, the shipped sem_dim is 2), or kCMax for the guarded generic loop
template <bool SECOND, int CT>
__global__ void __launch_bounds__(kT) k_geo_pairs(const float* __restrict__ xyz, const int64_t* __restrict__ neg, GeoWs w, int B, int C, int M,
                                                  float shift0, float shift1, float coef0, float coef1, int want_grad, int q0, int js) {
  __shared__ float sx[3][kT];
  __shared__ float sc[kCMax][kT];
  __shared__ float srm[kT];
  const int n = q0 + blockIdx.y, h = blockIdx.z;
  const int nb = (h == 0) ? (int)neg[n] : n;
  const float shift = h ? shift1 : shift0;
  const float coef = h ? coef1 : coef0;                  // weight_h / (B*M*M)
  const float om = w.oldmean[h];
  const int me = SECOND ? nb : n, other = SECOND ? n : nb;   // patch this thread's pixel / the streamed pixels belong to
  const int nbi = (M + kT - 1) / kT, jp = blockIdx.x / nbi;
  const int i = (blockIdx.x % nbi) * kT + threadIdx.x;
  const bool act = i < M;
  int jb, je;
  geo_jrange(M, js, jp, jb, je);
  float x[3] = {0, 0, 0}, c[CT], g[CT];
#pragma unroll
  for (int k = 0; k < CT; ++k) { c[k] = 0.f; g[k] = 0.f; }
  float my_rm = 0.f;
  if (act) {
#pragma unroll
    for (int k = 0; k < 3; ++k) x[k] = xyz[((size_t)me * 3 + k) * M + i];
#pragma unroll
    for (int k = 0; k < CT; ++k) if (k < C) c[k] = w.chat[((size_t)me * C + k) * M + i];
    if (!SECOND) my_rm = geo_rowmean_at(w, h, B, M, n, i);
  }
  float loss = 0.f;
  for (int j0 = jb; j0 < je; j0 += kT) {
    int j = j0 + threadIdx.x;
    __syncthreads();
    for (int k = 0; k < 3; ++k) sx[k][threadIdx.x] = (j < M) ? xyz[((size_t)other * 3 + k) * M + j] : 0.f;
#pragma unroll
    for (int k = 0; k < CT; ++k) if (k < C) sc[k][threadIdx.x] = (j < M) ? w.chat[((size_t)other * C + k) * M + j] : 0.f;
    // row means of the streamed pixels: pass 3 needs them (its pixel is the SECOND operand), and so does the self term of pass 2,
    // which folds its own pass 3 in: for the self pair fd and cd are symmetric, so d/d(chat_i) as second operand is the same sum
    // with t(q,i) = fd - rowmean[q] + ... in place of t(i,q)
    if (SECOND || h == 1) srm[threadIdx.x] = (j < M) ? geo_rowmean_at(w, h, B, M, n, j) : 0.f;
    __syncthreads();
    int lim = min(kT, je - j0);
#pragma unroll 4
    for (int jj = 0; jj < lim; ++jj) {
      float fd = inv_l1(fabsf(x[0] - sx[0][jj]) + fabsf(x[1] - sx[1][jj]) + fabsf(x[2] - sx[2][jj]));
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < CT; ++k)
        if (CT < kCMax || k < C) s += fabsf(c[k] - sc[k][jj]);
      float raw = 1.f / (s + kEpsCorr);
      float cd = fminf(kMaxCorr, raw);
      float t = fd - (SECOND ? srm[jj] : my_rm) + om - shift;          // fd_c - shift
      if (!SECOND) loss -= cd * t;                                      // -clamp(cd,0)*(fd_c-shift), cd > 0 always
      if (!SECOND && h == 1) t += fd - srm[jj] + om - shift;            // self term: + the role of this pixel as second operand
      if (want_grad && raw <= kMaxCorr) {                               // masked assignment blocks the gradient (:411)
        float a = t * cd * cd;                                          // d(-cd*t)/ds = t*cd^2
#pragma unroll
        for (int k = 0; k < CT; ++k)
          if (CT < kCMax || k < C) {
            float d = c[k] - sc[k][jj];
            float sg = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);      // d|u|/du, 0 at 0 like torch.abs
            g[k] += a * sg;                                             // same sign for both roles: d = mine - other
          }
      }
    }
  }
  if (want_grad && act) {
#pragma unroll
    for (int k = 0; k < CT; ++k) if (k < C) atomicAdd(&w.g_chat[((size_t)me * C + k) * M + i], g[k] * coef);
  }
  if (!SECOND) {
    float v = act ? loss : 0.f;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&w.acc[2 + h], (double)v);
  }
}

__global__ void k_finish_loss(const double* acc, float* loss, double coef0, double coef1) {
  if (threadIdx.x == 0) loss[0] = (float)(acc[2] * coef0 + acc[3] * coef1);
}

// backward of F.normalize: g_code = (g_chat - chat * <chat, g_chat>) / |code|
__global__ void k_normalize_bwd(const float* __restrict__ chat, const float* __restrict__ invn, const float* __restrict__ g_chat,
                                float* __restrict__ g_code, int B, int C, int M) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * M) return;
  int n = i / M, m = i % M;
  float dot = 0.f;
  for (int c = 0; c < C; ++c) dot += chat[((size_t)n * C + c) * M + m] * g_chat[((size_t)n * C + c) * M + m];
  float inv = invn[i];
  for (int c = 0; c < C; ++c) {
    size_t o = ((size_t)n * C + c) * M + m;
    g_code[o] = (g_chat[o] - chat[o] * dot) * inv;
  }
}

size_t carve_geo(char* base, int B, int C, int M, GeoWs* w) {
  size_t off = 0;
  auto take = [&](size_t bytes) { off = align_up(off, 256); size_t o = off; off += bytes; return o; };
  size_t o1 = take(sizeof(float) * B * C * M), o2 = take(sizeof(float) * B * M), o3 = take(sizeof(float) * kGeoJsMax * 2 * B * M),
         o4 = take(sizeof(float) * B * C * M), o5 = take(sizeof(double) * 4), o6 = take(sizeof(float) * 2);
  if (w && base) {
    w->chat = (float*)(base + o1); w->invn = (float*)(base + o2); w->rowmean = (float*)(base + o3);
    w->g_chat = (float*)(base + o4); w->acc = (double*)(base + o5); w->oldmean = (float*)(base + o6);
  }
  return align_up(off, 256);
}

// ---- appearance loss on sampled tensors (S = 121 samples per patch) ------------------------------------------
struct AppWs {
  float* fhat;     // [2,B,Cf,S] normalised feats / nfeats
  float* chat;     // [2,B,C,S]  normalised code / ncode
  float* invn;     // [2,B,S]
  float* fd;       // [2(h),B,S,S]
  float* rowmean;  // [2,B,S]
  float* g_chat;   // [2,B,C,S]
  double* acc;     // 4
  float* oldmean;  // 2
};

// normalise over channels: t [B,Cn,S] -> that, optional invn
__global__ void k_app_normalize(const float* __restrict__ t, float* __restrict__ that, float* __restrict__ invn, int B, int Cn, int S) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * S) return;
  int n = i / S, s = i % S;
  float acc = 0.f;
  for (int c = 0; c < Cn; ++c) { float v = t[((size_t)n * Cn + c) * S + s]; acc += v * v; }
  float nrm = sqrtf(acc), inv = 1.f / fmaxf(nrm, 1e-10f);
  for (int c = 0; c < Cn; ++c) that[((size_t)n * Cn + c) * S + s] = t[((size_t)n * Cn + c) * S + s] * inv;
  if (invn) invn[i] = (nrm > 1e-10f) ? inv : 0.f;
}
// fd[h,n,p,q] = <fhat1[n,:,p], fhat_h[n,:,q]>  (h=0: negatives, h=1: self); also row sums
__global__ void k_app_fd(AppWs w, int B, int Cf, int S) {
  const int n = blockIdx.y, h = blockIdx.z, p = blockIdx.x;
  const float* f1 = w.fhat + (size_t)n * Cf * S;                                  // feats
  const float* f2 = w.fhat + ((size_t)(h == 0 ? B : 0) + n) * Cf * S;             // nfeats (h=0) or feats (h=1)
  float rs = 0.f;
  for (int q = threadIdx.x; q < S; q += blockDim.x) {
    float d = 0.f;
    for (int c = 0; c < Cf; ++c) d = fmaf(f1[(size_t)c * S + p], f2[(size_t)c * S + q], d);
    w.fd[(((size_t)h * B + n) * S + p) * S + q] = d;
    rs += d;
  }
  __shared__ float red[32];
#pragma unroll
  for (int o = 16; o; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = rs;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (blockDim.x + 31) / 32; ++i) t += red[i];
    float rm = t / (float)S;
    w.rowmean[((size_t)h * B + n) * S + p] = rm;
    atomicAdd(&w.acc[h], (double)rm);
  }
}
// one block per (n,h): loss and gradients w.r.t. both normalised code operands
__global__ void k_app_pairs(AppWs w, int B, int C, int S, float shift0, float shift1, float coef0, float coef1, int want_grad) {
  const int n = blockIdx.x, h = blockIdx.y;
  const float shift = h ? shift1 : shift0, coef = h ? coef1 : coef0, om = w.oldmean[h];
  const float* c1 = w.chat + (size_t)n * C * S;
  const float* c2 = w.chat + ((size_t)(h == 0 ? B : 0) + n) * C * S;
  float* g1 = w.g_chat + (size_t)n * C * S;
  float* g2 = w.g_chat + ((size_t)(h == 0 ? B : 0) + n) * C * S;
  const float* fd = w.fd + ((size_t)h * B + n) * S * S;
  const float* rm = w.rowmean + ((size_t)h * B + n) * S;
  float loss = 0.f;
  for (int idx = threadIdx.x; idx < S * S; idx += blockDim.x) {
    int p = idx / S, q = idx % S;
    float cd = 0.f;
    for (int c = 0; c < C; ++c) cd = fmaf(c1[c * S + p], c2[c * S + q], cd);
    float t = fd[idx] - rm[p] + om - shift;
    if (cd > 0.f) {                                         // clamp(min=0): zero value and zero gradient below 0
      loss -= cd * t;
      if (want_grad)
        for (int c = 0; c < C; ++c) {
          atomicAdd(&g1[c * S + p], -t * coef * c2[c * S + q]);
          atomicAdd(&g2[c * S + q], -t * coef * c1[c * S + p]);
        }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&w.acc[2 + h], (double)loss);
}

size_t carve_app(char* base, int B, int Cf, int C, int S, AppWs* w) {
  size_t off = 0;
  auto take = [&](size_t bytes) { off = align_up(off, 256); size_t o = off; off += bytes; return o; };
  size_t o1 = take(sizeof(float) * 2 * B * Cf * S), o2 = take(sizeof(float) * 2 * B * C * S), o3 = take(sizeof(float) * 2 * B * S),
         o4 = take(sizeof(float) * 2 * B * S * S), o5 = take(sizeof(float) * 2 * B * S), o6 = take(sizeof(float) * 2 * B * C * S),
         o7 = take(sizeof(double) * 4), o8 = take(sizeof(float) * 2);
  if (w && base) {
    w->fhat = (float*)(base + o1); w->chat = (float*)(base + o2); w->invn = (float*)(base + o3); w->fd = (float*)(base + o4);
    w->rowmean = (float*)(base + o5); w->g_chat = (float*)(base + o6); w->acc = (double*)(base + o7); w->oldmean = (float*)(base + o8);
  }
  return align_up(off, 256);
}

}  // namespace

size_t geo_corr_workspace_bytes(int B, int C, int M) {
  if (B <= 0 || C <= 0 || C > kCMax || M <= 0) return 0;
  return carve_geo(nullptr, B, C, M, nullptr);
}

// Sharding (NsosLossShard, include/nerfsos.h): the arrays hold all B patches of the global batch, this call evaluates the query
// patches [q0, q0+nq); phase 1 leaves this rank's partial sums of the row means in sh->sums, the caller all-reduces them, phase 2
// turns the GLOBAL sums into old_mean and evaluates loss + gradients with the global denominators.  Row means and normalised
// codes stay in the workspace between the two phases.  sh == NULL: the whole batch in one call.
int geo_corr_loss(const float* xyz, const float* code, const int64_t* neg_idx, const float* params, float* loss, float* g_code,
                  int B, int C, int M, const NsosLossShard* sh, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  NSOS_REQUIRE(B > 0 && M > 0 && C > 0 && C <= kCMax, NSOS_ERR_UNSUPPORTED, "geo_corr_loss: need 1<=C<=%d", kCMax);
  const int q0 = sh ? sh->q0 : 0, nq = sh ? sh->nq : B, phase = sh ? sh->phase : 0, Bt = sh ? sh->B_total : B;
  NSOS_REQUIRE(q0 >= 0 && nq >= 0 && q0 + nq <= B && Bt >= B && phase >= 0 && phase <= 2, NSOS_ERR_BAD_ARG, "geo_corr_loss: bad shard descriptor");
  NSOS_REQUIRE(phase == 0 || (sh && sh->sums), NSOS_ERR_BAD_ARG, "geo_corr_loss: phases 1/2 need shard->sums");
  GeoWs w;
  size_t need = carve_geo((char*)workspace, B, C, M, &w);
  NSOS_REQUIRE(workspace && workspace_bytes >= need, NSOS_ERR_WORKSPACE, "geo_corr_loss: workspace too small (%zu < %zu)", workspace_bytes, need);
  const float self_shift = params[0], self_w = params[1], neg_shift = params[2], neg_w = params[3];   // HOST array
  const int nb = (B * M + 255) / 256;
  const int nbi = (M + kT - 1) / kT;
  // splits of the streamed pixel range (the same in phases 1 and 2: the row-mean partials stay in the workspace between them)
  const int js = geo_jsplit(nbi * std::max(nq, 1) * 2, M), js3 = geo_jsplit(nbi * std::max(nq, 1), M);
  w.js = js;
  dim3 grid(nbi * js, nq, 2);
  if (phase != 2) {
    NSOS_CHECK_CUDA(cudaMemsetAsync(w.acc, 0, sizeof(double) * 4, st));
    k_normalize<<<nb, 256, 0, st>>>(code, w.chat, w.invn, B, C, M);
    if (nq > 0) k_geo_rowmean<<<grid, kT, 0, st>>>(xyz, neg_idx, w, B, M, q0);
    if (phase == 1) { k_export_sums<<<1, 32, 0, st>>>(w.acc, sh->sums); NSOS_CHECK_CUDA(cudaGetLastError()); return NSOS_OK; }
  }
  NSOS_REQUIRE(loss, NSOS_ERR_BAD_ARG, "geo_corr_loss: loss pointer missing");
  if (g_code) NSOS_CHECK_CUDA(cudaMemsetAsync(w.g_chat, 0, sizeof(float) * B * C * M, st));
  k_oldmean<<<1, 32, 0, st>>>(w, phase == 2 ? sh->sums : w.acc, (double)Bt * M);
  const double denom = (double)Bt * M * M;
  // helper 0 = negative pair (neg_shift, neg_weight), helper 1 = self pair (image.py:476-482)
  const float coef0 = (float)(neg_w / denom), coef1 = (float)(self_w / denom);
  auto pairs = [&](auto second, dim3 gr, int want, int jsplit) {
    constexpr bool S2 = decltype(second)::value;
    switch (C) {
      case 1: k_geo_pairs<S2, 1><<<gr, kT, 0, st>>>(xyz, neg_idx, w, B, C, M, neg_shift, self_shift, coef0, coef1, want, q0, jsplit); break;
      case 2: k_geo_pairs<S2, 2><<<gr, kT, 0, st>>>(xyz, neg_idx, w, B, C, M, neg_shift, self_shift, coef0, coef1, want, q0, jsplit); break;
      case 3: k_geo_pairs<S2, 3><<<gr, kT, 0, st>>>(xyz, neg_idx, w, B, C, M, neg_shift, self_shift, coef0, coef1, want, q0, jsplit); break;
      case 4: k_geo_pairs<S2, 4><<<gr, kT, 0, st>>>(xyz, neg_idx, w, B, C, M, neg_shift, self_shift, coef0, coef1, want, q0, jsplit); break;
      default: k_geo_pairs<S2, kCMax><<<gr, kT, 0, st>>>(xyz, neg_idx, w, B, C, M, neg_shift, self_shift, coef0, coef1, want, q0, jsplit); break;
    }
  };
  if (nq > 0) pairs(std::false_type{}, grid, g_code != nullptr, js);
  if (g_code) {
    // pass 3 only for the negative pairs (blockIdx.z = 0): the self pairs were folded into pass 2
    if (nq > 0) pairs(std::true_type{}, dim3(nbi * js3, nq, 1), 1, js3);
    k_normalize_bwd<<<nb, 256, 0, st>>>(w.chat, w.invn, w.g_chat, g_code, B, C, M);
  }
  k_finish_loss<<<1, 32, 0, st>>>(w.acc, loss, neg_w / denom, self_w / denom);
  NSOS_CHECK_CUDA(cudaGetLastError());
  return NSOS_OK;
}

size_t app_corr_workspace_bytes(int B, int Cf, int C, int S) {
  if (B <= 0 || Cf <= 0 || C <= 0 || S <= 0) return 0;
  return carve_app(nullptr, B, Cf, C, S, nullptr);
}

// Sharded use: feats / nfeats / code / ncode hold this rank's B query patches (and their negatives, sampled by the caller from the
// gathered tensors); only the denominators (B_total) and old_mean (global sums of the row means, phases 1/2) are global.
int app_corr_loss(const float* feats, const float* nfeats, const float* code, const float* ncode, const float* params, float* loss,
                  float* g_code, float* g_ncode, int B, int Cf, int C, int S, const NsosLossShard* sh, void* workspace, size_t workspace_bytes,
                  cudaStream_t st) {
  NSOS_REQUIRE(B > 0 && Cf > 0 && C > 0 && S > 0, NSOS_ERR_BAD_ARG, "app_corr_loss: bad sizes");
  const int phase = sh ? sh->phase : 0, Bt = sh ? sh->B_total : B;
  NSOS_REQUIRE(Bt >= B && phase >= 0 && phase <= 2 && (phase == 0 || (sh && sh->sums)), NSOS_ERR_BAD_ARG, "app_corr_loss: bad shard descriptor");
  AppWs w;
  size_t need = carve_app((char*)workspace, B, Cf, C, S, &w);
  NSOS_REQUIRE(workspace && workspace_bytes >= need, NSOS_ERR_WORKSPACE, "app_corr_loss: workspace too small (%zu < %zu)", workspace_bytes, need);
  const float self_shift = params[0], self_w = params[1], neg_shift = params[2], neg_w = params[3];   // HOST array
  const bool want = g_code != nullptr && g_ncode != nullptr;
  const int nb = (B * S + 255) / 256;
  if (phase != 2) {
    NSOS_CHECK_CUDA(cudaMemsetAsync(w.acc, 0, sizeof(double) * 4, st));
    k_app_normalize<<<nb, 256, 0, st>>>(feats, w.fhat, nullptr, B, Cf, S);
    k_app_normalize<<<nb, 256, 0, st>>>(nfeats, w.fhat + (size_t)B * Cf * S, nullptr, B, Cf, S);
    k_app_normalize<<<nb, 256, 0, st>>>(code, w.chat, w.invn, B, C, S);
    k_app_normalize<<<nb, 256, 0, st>>>(ncode, w.chat + (size_t)B * C * S, w.invn + (size_t)B * S, B, C, S);
    k_app_fd<<<dim3(S, B, 2), 128, 0, st>>>(w, B, Cf, S);
    if (phase == 1) { k_export_sums<<<1, 32, 0, st>>>(w.acc, sh->sums); NSOS_CHECK_CUDA(cudaGetLastError()); return NSOS_OK; }
  }
  NSOS_REQUIRE(loss, NSOS_ERR_BAD_ARG, "app_corr_loss: loss pointer missing");
  NSOS_CHECK_CUDA(cudaMemsetAsync(w.g_chat, 0, sizeof(float) * 2 * B * C * S, st));
  k_oldmean<<<1, 32, 0, st>>>(GeoWs{nullptr, nullptr, nullptr, 1, nullptr, w.acc, w.oldmean}, phase == 2 ? sh->sums : w.acc, (double)Bt * S);
  const double denom = (double)Bt * S * S;
  const float coef0 = (float)(neg_w / denom), coef1 = (float)(self_w / denom);
  k_app_pairs<<<dim3(B, 2), 256, 0, st>>>(w, B, C, S, neg_shift, self_shift, coef0, coef1, want);
  if (want) {
    k_normalize_bwd<<<nb, 256, 0, st>>>(w.chat, w.invn, w.g_chat, g_code, B, C, S);
    k_normalize_bwd<<<nb, 256, 0, st>>>(w.chat + (size_t)B * C * S, w.invn + (size_t)B * S, w.g_chat + (size_t)B * C * S, g_ncode, B, C, S);
  }
  k_finish_loss<<<1, 32, 0, st>>>(w.acc, loss, neg_w / denom, self_w / denom);
  NSOS_CHECK_CUDA(cudaGetLastError());
  return NSOS_OK;
}

}  // namespace nsos
