// 128-thread ("ray group") versions of alpha-compositing and importance resampling for the fused tcgen05 kernel.
// Same arithmetic as the warp versions in render_device.cuh (models/renderer.py:35-85, models/sampler.py:91-170),
// spread over 4 warps per ray with one named barrier per ray: the per-pair serial section between the MLP tiles
// shrinks from ~55k cycles (one warp, NSOS_TRACE timeline) to a few thousand.
#pragma once
#include "render_device.cuh"

namespace nsos {

constexpr int kGroup = 128;   // threads per ray

struct GroupScratch {   // shared memory, per ray
  float* f;             // >= 64 floats: cross-warp partials
  double* d;            // >= 8 doubles
};

__device__ __forceinline__ void group_bar(int bar_id) { asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(kGroup) : "memory"); }

// block-of-4-warps sum; result valid in every thread.  `slot` selects a distinct scratch row (0..7).
__device__ __forceinline__ float group_sum(float v, int tid, int bar_id, float* scr) {
  v = warp_sum(v);
  if ((tid & 31) == 0) scr[tid >> 5] = v;
  group_bar(bar_id);
  float r = scr[0] + scr[1] + scr[2] + scr[3];
  group_bar(bar_id);
  return r;
}
__device__ __forceinline__ double group_sum(double v, int tid, int bar_id, double* scr) {
  v = warp_sum(v);
  if ((tid & 31) == 0) scr[tid >> 5] = v;
  group_bar(bar_id);
  double r = scr[0] + scr[1] + scr[2] + scr[3];
  group_bar(bar_id);
  return r;
}

// VolumetricRenderer.forward for one ray with 128 threads; thread t owns samples [t*ns, t*ns+ns), ns = ceil(S/128) <= 2.
// maps_out may be nullptr (padding ray of an odd batch): the maps are then not written.
__device__ inline void group_composite(const RayPass& p, int tid, int bar_id, const GroupScratch& gs, float* maps_out, float* weights_out,
                                       long long* dbg = nullptr) {   // dbg: optional clock64 stamps (thread 0 of the CTA)
  const int ns = (p.S + kGroup - 1) / kGroup;
  const int i0 = tid * ns;
  const int lane = tid & 31, w = tid >> 5;
  float a[2] = {0.f, 0.f};
  float prod = 1.f;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int i = i0 + j;
    if (j < ns && i < p.S) {
      float zi = p.z[i];
      float dist = (i + 1 < p.S) ? __fsub_rn(p.z[i + 1], zi) : 1e10f;        // renderer.py:35-37
      dist = __fmul_rn(dist, p.dnorm);                                        // :38
      float sig = __fadd_rn(p.raw[i * p.C + 3], pass_noise(p, i));           // :50
      float al = __fsub_rn(1.f, expf(-__fmul_rn(fmaxf(sig, 0.f), dist)));    // :52
      a[j] = al;
      prod = __fmul_rn(prod, __fadd_rn(__fsub_rn(1.f, al), 1e-10f));         // :57
    }
  }
  if (dbg) dbg[0] = clock64();                      // alpha computed
  // exclusive multiplicative scan over the 128 threads
  float incl = prod;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl = __fmul_rn(incl, v);
  }
  if (lane == 31) gs.f[w] = incl;
  float T = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) T = 1.f;
  group_bar(bar_id);
  for (int k = 0; k < w; ++k) T = __fmul_rn(gs.f[k], T);
  group_bar(bar_id);
  if (dbg) dbg[1] = clock64();                      // transmittance scan done
  float acc[9];   // rgb0 rgb1 rgb2 depth acc sem0..3
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = 0.f;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int i = i0 + j;
    if (j < ns && i < p.S) {
      float wt = __fmul_rn(a[j], T);                                          // :61
      const float* r = p.raw + i * p.C;
      acc[0] = fmaf(wt, sigmoidf_(r[0]), acc[0]);                             // :41, :62
      acc[1] = fmaf(wt, sigmoidf_(r[1]), acc[1]);
      acc[2] = fmaf(wt, sigmoidf_(r[2]), acc[2]);
      acc[3] = fmaf(wt, p.z[i], acc[3]);                                      // :69
      acc[4] += wt;                                                           // :71
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < p.sem_dim) acc[5 + c] = fmaf(wt, r[4 + c], acc[5 + c]);       // :65-66 (logits)
      if (weights_out) weights_out[i] = wt;
      T = __fmul_rn(T, __fadd_rn(__fsub_rn(1.f, a[j]), 1e-10f));
    }
  }
  if (dbg) dbg[2] = clock64();                      // per-sample products done
  // nine independent butterfly reductions, interleaved (channels >= 5 + sem_dim are zero)
#pragma unroll
  for (int o = 16; o; o >>= 1) {
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
  }
  if (lane < 9) {
    float v = acc[0];
#pragma unroll
    for (int k = 1; k < 9; ++k) v = (lane == k) ? acc[k] : v;
    gs.f[8 + w * 9 + lane] = v;
  }
  group_bar(bar_id);
  if (dbg) dbg[3] = clock64();                      // warp sums done
  // threads 0..8 each finish one channel (4 warp partials) and write one map entry; depth and acc also feed disp
  if (tid < 9 && maps_out) {
    const float* f = gs.f + 8;
    const float mine = f[tid] + f[9 + tid] + f[18 + tid] + f[27 + tid];
    const float ac = f[4] + f[9 + 4] + f[18 + 4] + f[27 + 4];
    const float bg = p.white_bkgd ? (1.f - ac) : 0.f;                           // renderer.py:77-81
    if (tid < 3) maps_out[tid] = mine + bg;                                     // rgb
    else if (tid == 3) {
      float dep = mine;
      if (ac <= 1e-10f) dep = 1e10f;                                            // :72
      maps_out[5] = dep;
      maps_out[3] = 1.f / fmaxf(1e-10f, dep / ac);                              // :74 disp
    } else if (tid == 4) maps_out[4] = ac;
    else if (tid - 5 < p.sem_dim) maps_out[6 + tid - 5] = mine + bg;            // semantic logits (:65-66)
  }
  group_bar(bar_id);
}

// ImportanceSampler.forward for one ray with 128 threads (Sc <= 128, Sc + K <= 256).
__device__ inline void group_importance(const ImportanceIO& io, int tid, int bar_id, const GroupScratch& gs) {
  const int M = io.Sc - 1, Mw = io.Sc - 2;
  const int lane = tid & 31, w = tid >> 5;
  if (io.dbg) io.dbg[0] = clock64();
  if (tid < M) io.bins[tid] = __fmul_rn(0.5f, __fadd_rn(io.z0[tid + 1], io.z0[tid]));   // sampler.py:157
  if (tid < io.Sc) io.zall[tid] = io.z0[tid];
  if (w == 0) {
    // pdf / cdf (:93-96) by one warp: fp64 accumulation like ATen's CPU cumsum, as a two-level prefix sum
    const int per = (Mw + 31) / 32;                       // <= 4 for Sc <= 128
    double part = 0.0;
    for (int k = 0; k < per; ++k) { int i = lane * per + k; if (i < Mw) part += (double)__fadd_rn(io.w0[i + 1], 1e-5f); }
    const float tot = (float)warp_sum(part);
    double loc[4], run = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      loc[k] = 0.0;
      int i = lane * per + k;
      if (k < per && i < Mw) { run += (double)__fdiv_rn(__fadd_rn(io.w0[i + 1], 1e-5f), tot); loc[k] = run; }
    }
    double incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const double off = incl - run;
    if (lane == 0) io.cdf[0] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) { int i = lane * per + k; if (k < per && i < Mw) io.cdf[i + 1] = (float)(off + loc[k]); }
  }
  group_bar(bar_id);
  if (io.dbg) io.dbg[1] = clock64();                // cdf done
  double s1 = 0.0;
  for (int j = tid; j < io.K; j += kGroup) {
    float u = io.det ? lin01(j, io.K) : (io.u ? io.u[j] : rng_uniform(io.seed, io.ray, RNG_U, j));   // :98-103
    int ind;
    float zs = invert_cdf_one(io.cdf, io.bins, M, u, &ind);
    if (io.z_inject) zs = io.z_inject[j];
    io.zall[io.Sc + j] = zs;
    if (io.z_samples) io.z_samples[j] = zs;
    if (io.inds) io.inds[j] = ind;
    s1 += (double)zs;
  }
  if (io.dbg) io.dbg[2] = clock64();                // inverse cdf done
  // z_std = population std of the K new samples (nerf_net.py:124)
  const double mean = group_sum(s1, tid, bar_id, gs.d) / (double)io.K;      // (also publishes zall)
  double s2 = 0.0;
  for (int j = tid; j < io.K; j += kGroup) { double d = (double)io.zall[io.Sc + j] - mean; s2 += d * d; }
  s2 = group_sum(s2, tid, bar_id, gs.d);
  if (tid == 0 && io.z_std) *io.z_std = (float)sqrt(s2 / (double)io.K);
  if (io.dbg) io.dbg[3] = clock64();                // z_std done
  // sort(cat([z, z_samples])) (:161) by stable rank.
  const int n = io.Sc + io.K;
  if (io.det) {
    // eval mode: u ascends, the inverse CDF is monotone, so both lists are already sorted -> merge by binary search:
    //   rank(coarse i) = i + #{samples <  z_i}      rank(sample j) = j + #{coarse <= s_j}
    const float* smp = io.zall + io.Sc;
    for (int e = tid; e < n; e += kGroup) {
      const float v = io.zall[e];
      int lo = 0, hi, rank;
      if (e < io.Sc) { hi = io.K; while (lo < hi) { int m = (lo + hi) >> 1; if (smp[m] < v) lo = m + 1; else hi = m; } rank = e + lo; }
      else { hi = io.Sc; while (lo < hi) { int m = (lo + hi) >> 1; if (io.zall[m] <= v) lo = m + 1; else hi = m; } rank = (e - io.Sc) + lo; }
      io.zsorted[rank] = v;
    }
  } else if (io.K <= kGroup && n >= kGroup) {            // (zsorted doubles as the 128-float exchange buffer)
    // train mode: the K drawn samples are unsorted.  Bitonic sort with one element per thread (padded with +inf): strides below 32
    // exchange by shuffle, the three wider ones through shared memory (zsorted is free until the final write); then the same
    // merge of two sorted lists as above.  (Round 1 ranked every element against all n: 11.6 k cycles per ray pair, now ~2.5 k.)
    group_bar(bar_id);                                   // every thread has read zall for z_std
    float v = (tid < io.K) ? io.zall[io.Sc + tid] : __int_as_float(0x7f800000);
    for (int k = 2; k <= kGroup; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        float o;
        if (j >= 32) {
          io.zsorted[tid] = v;
          group_bar(bar_id);
          o = io.zsorted[tid ^ j];
          group_bar(bar_id);
        } else {
          o = __shfl_xor_sync(0xffffffffu, v, j);
        }
        const bool up = (tid & k) == 0, lower = (tid & j) == 0;
        v = (lower == up) ? fminf(v, o) : fmaxf(v, o);
      }
    if (tid < io.K) io.zall[io.Sc + tid] = v;            // sorted samples replace the drawn order (outputs were written above)
    group_bar(bar_id);
    const float* smp = io.zall + io.Sc;
    for (int e = tid; e < n; e += kGroup) {
      const float x = io.zall[e];
      int lo = 0, hi, rank;
      if (e < io.Sc) { hi = io.K; while (lo < hi) { int m = (lo + hi) >> 1; if (smp[m] < x) lo = m + 1; else hi = m; } rank = e + lo; }
      else { hi = io.Sc; while (lo < hi) { int m = (lo + hi) >> 1; if (io.zall[m] <= x) lo = m + 1; else hi = m; } rank = (e - io.Sc) + lo; }
      io.zsorted[rank] = x;
    }
  } else {
    for (int e = tid; e < n; e += kGroup) {
      const float v = io.zall[e];
      int rank = 0;
#pragma unroll 8
      for (int i = 0; i < n; ++i) {
        float x = io.zall[i];
        rank += (x < v) || (x == v && i < e);
      }
      io.zsorted[rank] = v;
    }
  }
  group_bar(bar_id);
}

}  // namespace nsos
