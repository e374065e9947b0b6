// Shared device/host helpers for libnerfsos (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/nerfsos.h"

namespace nsos {

// ------------------------------------------------------------------------------------------------
// error plumbing (thread-local message behind nsos_last_error())
// ------------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define NSOS_CHECK_CUDA(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      nsos::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return NSOS_ERR_CUDA;                                                                \
    }                                                                                      \
  } while (0)

#define NSOS_REQUIRE(cond, code, ...)  \
  do {                                 \
    if (!(cond)) {                     \
      nsos::set_error(__VA_ARGS__);    \
      return code;                     \
    }                                  \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------------
// network geometry (mirrors models/nerf_mlp.py:24-65)
// ------------------------------------------------------------------------------------------------
struct NetGeom {
  int D, W, skip, Lp, Lv, enc, encv;  // enc = 3+6*Lp (63), encv = 3+6*Lv (27)
  int use_viewdirs, use_sem, sem_dim, sem_coord;
  int C;                               // raw channels: 4 + sem_dim
  // flat-parameter offsets (floats)
  int64_t w_pts[16], b_pts[16];
  int in_pts[16];                      // input width of pts_linears[i]
  int64_t w_alpha, b_alpha, w_feat, b_feat, w_views, b_views, w_rgb, b_rgb, w_s0, b_s0, w_s2, b_s2;
  int64_t w_out, b_out;                // output_linear (use_viewdirs == 0)
  int sem_in;                          // W (+enc)
  int64_t n_params;
};

// Returns false when the descriptor is invalid (e.g. D==skip+1: the reference itself cannot run it).
__host__ __device__ inline bool make_geom(const NsosNetDesc& d, NetGeom& g) {
  if (d.D < 1 || d.D > 16 || d.W < 1 || d.multires < 0 || d.multires > 16 || d.multires_views < 0 ||
      d.multires_views > 16)
    return false;
  g.D = d.D; g.W = d.W; g.Lp = d.multires; g.Lv = d.multires_views;
  g.enc = 3 + 6 * d.multires; g.encv = 3 + 6 * d.multires_views;
  g.skip = (d.skip >= 0 && d.skip < d.D - 1) ? d.skip : -1;
  if (d.skip >= 0 && d.skip == d.D - 1) return false;  // cat after the last layer breaks alpha_linear (ref. too)
  g.use_viewdirs = d.use_viewdirs != 0;
  g.use_sem = g.use_viewdirs && d.use_semantics != 0;
  g.sem_dim = g.use_sem ? d.sem_dim : 0;
  if (g.use_sem && (d.sem_dim < 1 || d.sem_dim > 8)) return false;
  g.sem_coord = d.sem_with_coord != 0;
  g.C = 4 + g.sem_dim;
  int64_t off = 0;
  for (int i = 0; i < g.D; ++i) {
    int in = (i == 0) ? g.enc : ((i - 1 == g.skip) ? g.W + g.enc : g.W);
    g.in_pts[i] = in;
    g.w_pts[i] = off; off += (int64_t)g.W * in;
    g.b_pts[i] = off; off += g.W;
  }
  g.w_out = g.b_out = -1;
  g.w_alpha = g.b_alpha = g.w_feat = g.b_feat = g.w_views = g.b_views = g.w_rgb = g.b_rgb = -1;
  g.w_s0 = g.b_s0 = g.w_s2 = g.b_s2 = -1;
  g.sem_in = g.W + (g.sem_coord ? g.enc : 0);
  if (g.use_viewdirs) {
    g.w_alpha = off; off += g.W;  g.b_alpha = off; off += 1;
    g.w_feat = off; off += (int64_t)g.W * g.W;  g.b_feat = off; off += g.W;
    g.w_views = off; off += (int64_t)(g.W / 2) * (g.W + g.encv);  g.b_views = off; off += g.W / 2;
    g.w_rgb = off; off += 3 * (g.W / 2);  g.b_rgb = off; off += 3;
    if (g.use_sem) {
      g.w_s0 = off; off += (int64_t)(g.W / 2) * g.sem_in;  g.b_s0 = off; off += g.W / 2;
      g.w_s2 = off; off += (int64_t)g.sem_dim * (g.W / 2);  g.b_s2 = off; off += g.sem_dim;
    }
  } else {
    g.w_out = off; off += 4 * g.W;  g.b_out = off; off += 4;
  }
  g.n_params = off;
  return true;
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG (production train mode; parity tests inject the reference's draws)
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
enum { RNG_T_RAND = 0, RNG_NOISE0 = 1, RNG_U = 2, RNG_NOISE1 = 3 };
// U[0,1) with 24 random bits, indexed by (ray, stream, sample)
__device__ inline float rng_uniform(uint64_t seed, int64_t ray, int stream, int idx) {
  uint32_t o[4];
  philox4x32_10((uint32_t)ray, (uint32_t)((uint64_t)ray >> 32), (uint32_t)stream, (uint32_t)(idx >> 2), (uint32_t)seed,
                (uint32_t)(seed >> 32), o);
  return (float)(o[idx & 3] >> 8) * (1.0f / 16777216.0f);
}
// N(0,1) via Box-Muller on two lanes of the same Philox block
__device__ inline float rng_normal(uint64_t seed, int64_t ray, int stream, int idx) {
  uint32_t o[4];
  philox4x32_10((uint32_t)ray, (uint32_t)((uint64_t)ray >> 32), (uint32_t)stream, (uint32_t)(idx >> 1), (uint32_t)seed,
                (uint32_t)(seed >> 32), o);
  float u1 = ((float)(o[0] >> 8) + 1.0f) * (1.0f / 16777216.0f);  // (0,1]
  float u2 = (float)(o[1] >> 8) * (1.0f / 16777216.0f);
  float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  return (idx & 1) ? r * s : r * c;
}

// ------------------------------------------------------------------------------------------------
// bit-faithful pieces of the sampler (models/sampler.py) -- explicit _rn ops: no FMA contraction
// ------------------------------------------------------------------------------------------------
// torch.linspace(0,1,steps)[i]: lower half step*i, upper half fma(-step, steps-1-i, 1) (see oracle)
__device__ inline float lin01(int i, int steps) {
  if (steps <= 1) return 0.0f;
  float step = __fdiv_rn(1.0f, (float)(steps - 1));
  return (i < steps / 2) ? __fmul_rn(step, (float)i) : __fmaf_rn(-step, (float)(steps - 1 - i), 1.0f);
}
// sampler.py:48  z = near*(1-t) + far*t
__device__ inline float z_lin(float near, float far, int i, int steps) {
  float t = lin01(i, steps);
  return __fadd_rn(__fmul_rn(near, __fsub_rn(1.0f, t)), __fmul_rn(far, t));
}
// sampler.py:46-69
__device__ inline float z_stratified(float near, float far, int i, int steps, bool perturb, float t_rand) {
  float z = z_lin(near, far, i, steps);
  if (!perturb) return z;
  float lower = (i == 0) ? z : __fmul_rn(0.5f, __fadd_rn(z, z_lin(near, far, i - 1, steps)));
  float upper = (i == steps - 1) ? z : __fmul_rn(0.5f, __fadd_rn(z_lin(near, far, i + 1, steps), z));
  return __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t_rand));
}
// sampler.py:71  pts = o + d*z
__device__ inline float pt_coord(float o, float d, float z) { return __fadd_rn(o, __fmul_rn(d, z)); }

// sin and cos of one fp32 argument, |x| < 2^20: three-term Cody-Waite reduction by pi/2 (FMA) and degree-9 / degree-8
// minimax polynomials on [-pi/4, pi/4].  Branch-free, ~22 instructions, no local memory: the positional encoding calls it
// 15 times per sample point, and libdevice's sincosf (range check + Payne-Hanek slow path) costs 2-3x the instructions
// and code size.  Max error over the encoder's argument range (|2^k x| <= 8192): 1.5 ulp / 7.4e-8 absolute, the same as
// a correctly rounded fp32 sine of the fp32 argument (checked against float64 on 2e6 arguments per octave).
__device__ __forceinline__ void sincos_cw(float x, float* s, float* c) {
  const float t = __fmaf_rn(x, 0.636619772f, 12582912.f);            // 1.5 * 2^23: the low mantissa bits hold rint(x * 2/pi)
  const int i = __float_as_int(t);
  const float q = __fsub_rn(t, 12582912.f);
  float r = __fmaf_rn(q, -1.5707963705062866f, x);
  r = __fmaf_rn(q, 4.371138828673793e-08f, r);
  r = __fmaf_rn(q, 1.7763568394002505e-15f, r);
  const float r2 = __fmul_rn(r, r);
  float ps = __fmaf_rn(2.7172509362571873e-06f, r2, -0.00019839218293782324f);
  ps = __fmaf_rn(ps, r2, 0.008333329111337662f);
  ps = __fmaf_rn(ps, r2, -0.1666666716337204f);
  const float sn = __fmaf_rn(ps, __fmul_rn(r2, r), r);
  float pc = __fmaf_rn(2.438097908452619e-05f, r2, -0.001388666103594005f);
  pc = __fmaf_rn(pc, r2, 0.04166661947965622f);
  pc = __fmaf_rn(pc, r2, -0.5f);
  const float cs = __fmaf_rn(pc, r2, 1.0f);
  const bool swap = (i & 1) != 0;
  const float ss = swap ? cs : sn, cc = swap ? sn : cs;
  *s = __uint_as_float(__float_as_uint(ss) ^ (((uint32_t)i & 2u) << 30));
  *c = __uint_as_float(__float_as_uint(cc) ^ (((uint32_t)(i + 1) & 2u) << 30));
}

// embedder.py:34-48, column c of gamma(x) for c in [0, 3+6L): [x, sin(2^0 x), cos(2^0 x), ...]
__device__ inline void encode3(const float x[3], int L, float* out) {
  out[0] = x[0]; out[1] = x[1]; out[2] = x[2];
  float f = 1.0f;
  for (int k = 0; k < L; ++k) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float s, c;
      sincos_cw(__fmul_rn(x[a], f), &s, &c);
      out[3 + 6 * k + a] = s;
      out[3 + 6 * k + 3 + a] = c;
    }
    f *= 2.0f;
  }
}

__device__ inline float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ------------------------------------------------------------------------------------------------
// generic fp32 GEMM used by the SIMT reference path (simt_gemm.cu)
// ------------------------------------------------------------------------------------------------
// C[M,N] = epi( alpha * sum_k A(m,k) * B(k,n) ) with arbitrary strides; K may be split in two sources
// (A1|A2 concatenated along K) to express the reference's torch.cat inputs without materialising them.
struct GemmArgs {
  const float* A1; int64_t a1_rs, a1_cs; int K1; int a1_rowdiv;  // element (m,k) = A1[(m/rowdiv)*rs + k*cs]
  const float* A2; int64_t a2_rs, a2_cs; int K2; int a2_rowdiv;
  const float* B;  int64_t b_rs, b_cs; int b_rowdiv;               // element (k,n) = B[(k/rowdiv)*rs + n*cs], k over K1+K2
  const float* bias;                                                // [N] or null
  const float* mask; int64_t mask_ld;                               // relu-backward mask: keep where mask(m,n) > 0
  float* C; int64_t c_rs, c_cs;
  int M, N;
  int relu;       // apply max(0,.) after bias
  int accumulate; // C += result (atomicAdd when split_k > 1)
  int split_k;    // number of K splits (grid.z)
};
int launch_gemm(const GemmArgs& a, cudaStream_t s);

}  // namespace nsos
