// Warp-level alpha-compositing and importance resampling shared by the SIMT path (simt_render.cu) and
// the fused tcgen05 kernel (tc_render.cu).  One warp owns one ray; lane l owns the contiguous sample
// chunk [l*n, l*n+n), n = ceil(S/32) <= 8 (S <= 256).
#pragma once
#include "common.cuh"

namespace nsos {

constexpr int kMaxS = 256;          // max samples per ray in one pass
constexpr int kMaxChunk = kMaxS / 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct RayPass {
  const float* raw;    // [S, C]  (shared or global)
  const float* z;      // [S]
  const float* noise;  // [S] injected N(0,1) draws or nullptr
  float noise_std;     // raw_noise_std (0: no noise)
  uint64_t seed; int64_t ray; int rng_stream;  // Philox fallback when noise == nullptr && noise_std > 0
  float dnorm;         // ||rays_d||
  int S, C, sem_dim, white_bkgd;
};

__device__ __forceinline__ float pass_noise(const RayPass& p, int i) {
  if (p.noise_std <= 0.f) return 0.f;
  float n = p.noise ? p.noise[i] : rng_normal(p.seed, p.ray, p.rng_stream, i);
  return n * p.noise_std;                                                     // renderer.py:47
}

// VolumetricRenderer.forward (models/renderer.py:35-85) for one ray.
// maps_out (global, 6+sem_dim floats): rgb[3] disp acc depth sem[sem_dim]; weights_out: [S] or nullptr
// (may be shared memory).  All lanes must call.
__device__ inline void warp_composite(const RayPass& p, int lane, float* maps_out, float* weights_out) {
  const int n = (p.S + 31) >> 5;
  const int i0 = lane * n;
  float a[kMaxChunk];
  float prod = 1.f;
#pragma unroll
  for (int j = 0; j < kMaxChunk; ++j) {
    a[j] = 0.f;
    int i = i0 + j;
    if (j < n && i < p.S) {
      float zi = p.z[i];
      float dist = (i + 1 < p.S) ? __fsub_rn(p.z[i + 1], zi) : 1e10f;        // :35-37
      dist = __fmul_rn(dist, p.dnorm);                                        // :38
      float sig = __fadd_rn(p.raw[i * p.C + 3], pass_noise(p, i));           // :50
      float al = __fsub_rn(1.f, expf(-__fmul_rn(fmaxf(sig, 0.f), dist)));    // :52
      a[j] = al;
      prod = __fmul_rn(prod, __fadd_rn(__fsub_rn(1.f, al), 1e-10f));         // :57
    }
  }
  // exclusive multiplicative scan over lanes -> transmittance at the start of this lane's chunk
  float incl = prod;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl = __fmul_rn(incl, v);
  }
  float T = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) T = 1.f;
  float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f, dep = 0.f, acc = 0.f;
  float sem[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) sem[c] = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxChunk; ++j) {
    int i = i0 + j;
    if (j < n && i < p.S) {
      float w = __fmul_rn(a[j], T);                                           // :61
      const float* r = p.raw + i * p.C;
      rgb0 = fmaf(w, sigmoidf_(r[0]), rgb0);                                  // :41, :62
      rgb1 = fmaf(w, sigmoidf_(r[1]), rgb1);
      rgb2 = fmaf(w, sigmoidf_(r[2]), rgb2);
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c < p.sem_dim) sem[c] = fmaf(w, r[4 + c], sem[c]);               // :65-66 (logits)
      dep = fmaf(w, p.z[i], dep);                                             // :69
      acc += w;                                                               // :71
      if (weights_out) weights_out[i] = w;
      T = __fmul_rn(T, __fadd_rn(__fsub_rn(1.f, a[j]), 1e-10f));
    }
  }
  rgb0 = warp_sum(rgb0); rgb1 = warp_sum(rgb1); rgb2 = warp_sum(rgb2);
  dep = warp_sum(dep); acc = warp_sum(acc);
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (c < p.sem_dim) sem[c] = warp_sum(sem[c]);
  if (lane == 0) {
    if (acc <= 1e-10f) dep = 1e10f;                                           // :72
    float disp = 1.f / fmaxf(1e-10f, dep / acc);                              // :74
    float bg = p.white_bkgd ? (1.f - acc) : 0.f;                              // :77-81
    maps_out[0] = rgb0 + bg; maps_out[1] = rgb1 + bg; maps_out[2] = rgb2 + bg;
    maps_out[3] = disp; maps_out[4] = acc; maps_out[5] = dep;
    for (int c = 0; c < p.sem_dim; ++c) maps_out[6 + c] = sem[c] + bg;
  }
}

// Backward of warp_composite.  g_maps: upstream grads in the maps layout
// (rgb, disp, acc, depth, semantics).  g_raw_out: [S, C] (overwritten).
__device__ inline void warp_composite_bwd(const RayPass& p, int lane, const float* g_maps, float* g_raw_out) {
  const int n = (p.S + 31) >> 5;
  const int i0 = lane * n;
  float a[kMaxChunk], dist[kMaxChunk];
  bool pos[kMaxChunk];
  float prod = 1.f;
#pragma unroll
  for (int j = 0; j < kMaxChunk; ++j) {
    a[j] = 0.f; dist[j] = 0.f; pos[j] = false;
    int i = i0 + j;
    if (j < n && i < p.S) {
      float zi = p.z[i];
      float d = (i + 1 < p.S) ? __fsub_rn(p.z[i + 1], zi) : 1e10f;
      d = __fmul_rn(d, p.dnorm);
      float sig = __fadd_rn(p.raw[i * p.C + 3], pass_noise(p, i));
      a[j] = __fsub_rn(1.f, expf(-__fmul_rn(fmaxf(sig, 0.f), d)));
      dist[j] = d; pos[j] = sig > 0.f;
      prod = __fmul_rn(prod, __fadd_rn(__fsub_rn(1.f, a[j]), 1e-10f));
    }
  }
  float incl = prod;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl = __fmul_rn(incl, v);
  }
  float T0 = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) T0 = 1.f;
  // acc decides whether depth carries gradient (masked overwrite, renderer.py:72)
  float T = T0, accl = 0.f, depl = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxChunk; ++j) {
    int i = i0 + j;
    if (j < n && i < p.S) { accl += a[j] * T; depl = fmaf(a[j] * T, p.z[i], depl); T *= (1.f - a[j] + 1e-10f); }
  }
  float acc = warp_sum(accl);
  const float dep = warp_sum(depl);
  float gr0 = g_maps[0], gr1 = g_maps[1], gr2 = g_maps[2];
  float g_acc = g_maps[4];
  float g_dep = (acc <= 1e-10f) ? 0.f : g_maps[5];
  // disp = 1 / max(1e-10, depth / acc) (renderer.py:74): gradient where neither the depth overwrite nor the max clamps
  const float g_disp = g_maps[3];
  if (g_disp != 0.f && acc > 1e-10f && dep / acc > 1e-10f) {
    const float disp = acc / dep;
    g_dep = fmaf(-g_disp, disp * disp / acc, g_dep);                 // d disp / d depth = -acc / depth^2
    g_acc = fmaf(g_disp, disp * disp * dep / (acc * acc), g_acc);    // d disp / d acc   =  1 / depth
  }
  float gs[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) gs[c] = (c < p.sem_dim) ? g_maps[6 + c] : 0.f;
  if (p.white_bkgd) {  // rgb += 1-acc ; sem += 1-acc
    float s = gr0 + gr1 + gr2;
    for (int c = 0; c < p.sem_dim; ++c) s += gs[c];
    g_acc -= s;
  }
  // G_i and local suffix sums of w_i G_i
  float G[kMaxChunk], w[kMaxChunk], Tj[kMaxChunk];
  float local = 0.f;
  T = T0;
#pragma unroll
  for (int j = 0; j < kMaxChunk; ++j) {
    G[j] = 0.f; w[j] = 0.f; Tj[j] = 0.f;
    int i = i0 + j;
    if (j < n && i < p.S) {
      const float* r = p.raw + i * p.C;
      float g = gr0 * sigmoidf_(r[0]) + gr1 * sigmoidf_(r[1]) + gr2 * sigmoidf_(r[2]) + g_dep * p.z[i] + g_acc;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c < p.sem_dim) g += gs[c] * r[4 + c];
      G[j] = g; Tj[j] = T; w[j] = a[j] * T;
      local += w[j] * g;
      T *= (1.f - a[j] + 1e-10f);
    }
  }
  // exclusive suffix sum over lanes (sum of `local` for lanes > lane)
  float suf = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float v = __shfl_down_sync(0xffffffffu, suf, o);
    if (lane + o < 32) suf += v;
  }
  float after = suf - local;  // sum over later lanes
#pragma unroll
  for (int j = kMaxChunk - 1; j >= 0; --j) {
    int i = i0 + j;
    if (j < n && i < p.S) {
      const float* r = p.raw + i * p.C;
      float* g = g_raw_out + i * p.C;
      float om = 1.f - a[j] + 1e-10f;
      float dalpha = Tj[j] * G[j] - after / om;
      g[3] = pos[j] ? dalpha * dist[j] * (1.f - a[j]) : 0.f;
      float c0 = sigmoidf_(r[0]), c1 = sigmoidf_(r[1]), c2 = sigmoidf_(r[2]);
      g[0] = w[j] * gr0 * c0 * (1.f - c0);
      g[1] = w[j] * gr1 * c1 * (1.f - c1);
      g[2] = w[j] * gr2 * c2 * (1.f - c2);
      for (int c = 0; c < p.sem_dim; ++c) g[4 + c] = w[j] * gs[c];
      after += w[j] * G[j];
    }
  }
}

// Inverse-CDF lookup (sampler.py:117-132) on a shared/global cdf & bins of length M.  Exact stage.
__device__ __forceinline__ float invert_cdf_one(const float* cdf, const float* bins, int M, float u, int* ind_out) {
  // inds = #{k : cdf[k] <= u}  == searchsorted(cdf, u, right=True)
  int lo = 0, hi = M;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
  }
  int ind = lo;
  int below = max(0, ind - 1), above = min(M - 1, ind);
  float cb = cdf[below], ca = cdf[above];
  float denom = __fsub_rn(ca, cb);
  if (denom < 1e-5f) denom = 1.f;
  float t = __fdiv_rn(__fsub_rn(u, cb), denom);
  float bb = bins[below], ba = bins[above];
  *ind_out = ind;
  return __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
}

struct ImportanceIO {
  const float* z0;   // [Sc] coarse z (shared)
  const float* w0;   // [Sc] coarse weights (shared)
  float* cdf;        // [Sc-1] scratch (shared)
  float* bins;       // [Sc-1] scratch (shared)
  float* zall;       // [Sc+K] scratch (shared)
  float* zsorted;    // [Sc+K] out (shared or global)
  const float* u;    // [K] injected draws or nullptr
  const float* z_inject = nullptr;   // [K] injected importance samples (stage-wise parity hook) or nullptr
  float* z_samples;  // [K] out (global) or nullptr
  int64_t* inds;     // [K] out (global) or nullptr
  float* z_std;      // out (global, 1 float) or nullptr
  int Sc, K;
  bool det;          // perturb == 0
  uint64_t seed; int64_t ray;
  long long* dbg = nullptr;   // optional clock64 stamps (lane 0): [0] start [1] cdf done [2] search done [3] std done [4] sort done
};

// ImportanceSampler.forward (sampler.py:136-170) for one ray.  All lanes must call.
__device__ inline void warp_importance(const ImportanceIO& io, int lane) {
  const int M = io.Sc - 1;   // bins (z_vals_mid)
  if (io.dbg && lane == 0) io.dbg[0] = clock64();
  const int Mw = io.Sc - 2;  // weights[..., 1:-1]
  for (int i = lane; i < M; i += 32) io.bins[i] = __fmul_rn(0.5f, __fadd_rn(io.z0[i + 1], io.z0[i]));  // :157
  for (int i = lane; i < io.Sc; i += 32) io.zall[i] = io.z0[i];
  // pdf/cdf (:93-96): fp64-accumulated like ATen's CPU cumsum (see oracle.pdf_cdf)
  double part = 0.0;
  for (int i = lane; i < Mw; i += 32) part += (double)__fadd_rn(io.w0[i + 1], 1e-5f);
  float tot = (float)warp_sum(part);
  __syncwarp();
  if (lane == 0) {
    double c = 0.0;
    io.cdf[0] = 0.f;
    for (int i = 0; i < Mw; ++i) {
      float pdf = __fdiv_rn(__fadd_rn(io.w0[i + 1], 1e-5f), tot);
      c += (double)pdf;
      io.cdf[i + 1] = (float)c;
    }
  }
  __syncwarp();
  if (io.dbg && lane == 0) io.dbg[1] = clock64();
  double s1 = 0.0;
  for (int j = lane; j < io.K; j += 32) {
    float u = io.det ? lin01(j, io.K) : (io.u ? io.u[j] : rng_uniform(io.seed, io.ray, RNG_U, j));  // :98-103
    int ind;
    float zs = invert_cdf_one(io.cdf, io.bins, M, u, &ind);
    if (io.z_inject) zs = io.z_inject[j];
    io.zall[io.Sc + j] = zs;
    if (io.z_samples) io.z_samples[j] = zs;
    if (io.inds) io.inds[j] = ind;
    s1 += (double)zs;
  }
  if (io.dbg && lane == 0) io.dbg[2] = clock64();
  // z_std = population std of the K new samples (nerf_net.py:124)
  double mean = warp_sum(s1) / (double)io.K;
  __syncwarp();
  double s2 = 0.0;
  for (int j = lane; j < io.K; j += 32) {
    double d = (double)io.zall[io.Sc + j] - mean;
    s2 += d * d;
  }
  s2 = warp_sum(s2);
  if (lane == 0 && io.z_std) *io.z_std = (float)sqrt(s2 / (double)io.K);
  if (io.dbg && lane == 0) io.dbg[3] = clock64();
  // sort(cat([z, z_samples])) (:161) by stable rank
  const int n = io.Sc + io.K;
  for (int e = lane; e < n; e += 32) {
    float v = io.zall[e];
    int rank = 0;
    for (int i = 0; i < n; ++i) {
      float x = io.zall[i];
      rank += (x < v) || (x == v && i < e);
    }
    io.zsorted[rank] = v;
  }
  __syncwarp();
  if (io.dbg && lane == 0) io.dbg[4] = clock64();
}

}  // namespace nsos
