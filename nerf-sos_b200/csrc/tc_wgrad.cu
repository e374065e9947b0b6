// Kernel C: weight gradients of the semantic head on the tcgen05 tensor cores (the --fix_backbone training recipe,
// run_nerf.py:307-318 -- the only parameters every shipped config trains are semantic_linear.{0,2} of both nets).
//
// Replaces, for one net and one chunk of P sample points, autograd through nerf_mlp.py:79-80
//     s0  = relu(W0 . [h, gamma] + b0)            (semantic_linear.0, 319 -> 128)
//     sem = W2 . s0 + b2                          (semantic_linear.2, 128 -> sem_dim)
// given g_sem = d(loss)/d(sem) per point (from the compositing backward):
//     g_s0 = (W2^T g_sem) * [s0 > 0]
//     dW0 += g_s0^T . [h, gamma]      db0 += sum_p g_s0        dW2 += g_sem^T . s0       db2 += sum_p g_sem
//
// The contraction runs over POINTS, so both operands are transposed on the way into shared memory: per slab of 64
// points the 8 worker warps read h / gamma / s0 / g_raw rows (one point per lane), split every value into bf16 hi + lo,
// pair neighbouring points with one shuffle and store K-major SWIZZLE_128B tiles
//     A  [128 units   x 64 pts] = g_s0^T      B  [320 feats x 64 pts] = [h(256), gamma(63), 1]^T
//     A2 [128 units   x 64 pts] = s0^T        B2 [ 16       x 64 pts] = g_sem^T (rows >= sem_dim zero)
// and one elected thread issues 3 x 4 x 3 tcgen05.mma (hi.hi + lo.hi + hi.lo; M=128, N=256/64/16, K=16) that accumulate
// in TMEM for the whole life of the CTA:  D[:, 0:319] = dW0, D[:, 319] = db0 (the constant-one feature), D[:, 320:324] =
// dW2^T.  At the end every CTA stores its partial sums column-major into scratch memory and k_wgrad_reduce adds them to the flat
// gradient buffer in a fixed order (round 1 used atomics: 148 CTAs x 41 k atomics on the same addresses per launch).
// bf16 hi+lo carries 16 mantissa bits per operand with fp32 range (gradients can be far below the fp16 range).
#include <cuda_bf16.h>

#include "internal.h"
#include "tc_ptx.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace nsos {
namespace {
using namespace ptx;

constexpr int kWgWorkers = 256, kWgThreads = 288, kWgMmaWarp = 8;   // k_wgrad_gen: 8 fill warps + an issuer warp
// k_sem_wgrad: 8 fill warps + the issuer's warpgroup (3 of its warps idle).  12 warps are launched at 168 registers; the issuer's group
// drops to 24 and the fill warps grow to 240 with setmaxnreg (the pool is what the CTA's own warps release).  History: 8 warps at 255
// registers with the issue duty inside a fill warp issued one instruction per 8 cycles (ncu: 75 % of cycles without an eligible warp);
// 16 fill warps at 112 registers next to the issuer group measured the same as this split.
constexpr int kSemWarps = 8, kSemIssuer = kSemWarps, kSemThreads = 32 * (kSemWarps + 4);
constexpr int kHalfPts = 32;                       // points per fill / MMA unit: half of a 64-point tile
constexpr int kSlabPts = 64;
constexpr int kRowsA = 128, kRowsB = 320, kRowsB2 = 16;
constexpr int kColD1 = 0, kColD1b = 256, kColD2 = 320;
constexpr uint32_t kWgTmemCols = 512;

// ---- second stage of the weight gradients ----------------------------------------------------------------------
// Every CTA of a weight-gradient kernel holds a [128 x 336] partial sum in TMEM.  Adding those to the gradient buffer with
// atomics (round 1) cost more than the contraction: 148 CTAs x 41 k atomics on the same addresses per launch, 16 launches per
// training step = 4.8 of the kernel's 7 ms.  The CTAs now store their partials column-major ([cta][col][128 rows]: a warp
// store is one line) into scratch memory and k_wgrad_reduce adds them up in a fixed order -- coalesced, deterministic.
constexpr int kPartCols = 336, kPartRows = 128, kPartMaxCtas = 160;
constexpr size_t kPartFloatsPerCta = (size_t)kPartCols * kPartRows;
struct WgSeg {          // columns [c0, c1) of the partial -> dst[row * rs + (col - c0) * cs], rows < nrows
  int c0, c1, nrows;
  float* dst;
  long long rs, cs;
};
struct WgReduceParams {
  const float* part;    // [ny][ncta][kPartCols][kPartRows]
  int ncta;
  WgSeg seg[2][4];      // [blockIdx.y]
};
__global__ void __launch_bounds__(256) k_wgrad_reduce(const __grid_constant__ WgReduceParams R) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= kPartCols * kPartRows) return;
  const int col = idx / kPartRows, row = idx % kPartRows, y = blockIdx.y;
  const WgSeg* sg = nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (col >= R.seg[y][i].c0 && col < R.seg[y][i].c1 && row < R.seg[y][i].nrows) sg = &R.seg[y][i];
  if (!sg) return;
  const float* p = R.part + ((size_t)y * R.ncta * kPartCols + col) * kPartRows + row;
  float s0 = 0.f, s1 = 0.f;
  int b = 0;
  for (; b + 1 < R.ncta; b += 2) { s0 += p[(size_t)b * kPartFloatsPerCta]; s1 += p[(size_t)(b + 1) * kPartFloatsPerCta]; }
  if (b < R.ncta) s0 += p[(size_t)b * kPartFloatsPerCta];
  float* d = sg->dst + row * sg->rs + (col - sg->c0) * sg->cs;
  *d += s0 + s1;
}

struct WgParams {
  const float *h, *enc, *s0, *g_raw, *w_s2;
  float *gW0, *gb0, *gW2, *gb2;
  float* part;       // [gridDim.x][kPartCols][128] partial sums of this launch (k_wgrad_reduce adds them to the gradients)
  long long P;
  int C, enc_dim, enc_ld, sem_dim, sem_coord, ld0;
  int trace;         // NSOS_WG_TRACE=1: CTA 0 prints a timeline of its first halves (debug)
  int enc_blocked;   // gamma rows: blocked like h / s0 (saved by the training forward) or row-major with pitch enc_ld (k_encode_pts)
};

struct WgSmem {
  uint8_t *a[2], *a2[2], *b[2], *b2[2];   // [plane]: 0 = hi, 1 = lo
  float* w2;                              // [4][128]
  uint64_t *ready, *done;                 // [2] each: one pair per tile half (k_sem_wgrad counts arrivals in `cnt` instead of ready)
  uint32_t* cnt;                          // [2] fill warps that have finished the half
  uint32_t* tmem_ptr;
};
__host__ __device__ inline size_t wg_carve(uint8_t* base, WgSmem* s) {
  size_t off = 0;
  auto take = [&](size_t bytes, size_t align) { off = (off + align - 1) / align * align; size_t o = off; off += bytes; return o; };
  size_t oa[2], oa2[2], ob[2], ob2[2];
  for (int p = 0; p < 2; ++p) oa[p] = take(kRowsA * 128, 1024);
  for (int p = 0; p < 2; ++p) oa2[p] = take(kRowsA * 128, 1024);
  for (int p = 0; p < 2; ++p) ob[p] = take(kRowsB * 128, 1024);
  for (int p = 0; p < 2; ++p) ob2[p] = take(kRowsB2 * 128, 1024);
  size_t ow = take(sizeof(float) * 4 * 128, 16), obar = take(32, 8), otp = take(16, 16), ocnt = take(8, 8);
  if (s) {
    for (int p = 0; p < 2; ++p) { s->a[p] = base + oa[p]; s->a2[p] = base + oa2[p]; s->b[p] = base + ob[p]; s->b2[p] = base + ob2[p]; }
    s->w2 = (float*)(base + ow); s->ready = (uint64_t*)(base + obar); s->done = s->ready + 2; s->tmem_ptr = (uint32_t*)(base + otp); s->cnt = (uint32_t*)(base + ocnt);
  }
  return off;
}

__host__ __device__ __forceinline__ uint32_t make_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // D=f32, A=B=bf16, K-major
}

// bf16 hi in the low half, bf16 lo (= residual) in the high half
__device__ __forceinline__ uint32_t split_bf16(float x) {
  __nv_bfloat16 hi = __float2bfloat16_rn(x);
  __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
  return (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
}

// Values xa (tile row ra) and xb (tile row ra+4) of THIS lane's point; lanes 2k / 2k+1 hold neighbouring points.
// Even lanes store row ra, odd lanes row ra+4 (their swizzle images use disjoint bank halves): one 32-bit word
// (2 points) per plane.
__device__ __forceinline__ void put_pair(uint8_t* hi_tile, uint8_t* lo_tile, int ra, float xa, float xb, int lane, int wd) {
  const uint32_t pa = split_bf16(xa), pb = split_bf16(xb);
  const bool even = (lane & 1) == 0;
  const uint32_t recv = __shfl_xor_sync(0xffffffffu, even ? pb : pa, 1);
  const uint32_t first = even ? pa : recv, second = even ? recv : pb;     // first = even point, second = odd point
  const int R = even ? ra : ra + 4;
  const size_t off = (size_t)(R >> 3) * 1024 + (size_t)(R & 7) * 128 + (size_t)((((wd >> 2) ^ (R & 7)) << 4) + ((wd & 3) << 2));
  *reinterpret_cast<uint32_t*>(hi_tile + off) = __byte_perm(first, second, 0x5410);
  *reinterpret_cast<uint32_t*>(lo_tile + off) = __byte_perm(first, second, 0x7632);
}

__device__ __forceinline__ void put8(uint8_t* hi_tile, uint8_t* lo_tile, int r0, const float (&v)[8], int lane, int wd) {
#pragma unroll
  for (int i = 0; i < 4; ++i) put_pair(hi_tile, lo_tile, r0 + i, v[i], v[4 + i], lane, wd);
}

__device__ __forceinline__ void load8(const float* __restrict__ src, bool valid, float (&v)[8]) {
  if (valid) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
  }
}

constexpr int kSemH = 128 / kSemWarps, kSemE = 32 / kSemWarps, kSemS = 64 / kSemWarps;   // loads per lane: h, gamma, s0 (2 features per instruction)
// One lane's share of a 32-point half slab: TWO neighbouring points (2m, 2m+1; m = lane & 15) of the features
// f = base + 8*(k >> 2) + (k & 3) + 4*(lane >> 4) (h, s0; gamma: base + 2k + (lane >> 4)): 256/NW h features, 64/NW gamma features,
// 128/NW s0 units of fill warp e (14 independent 8-byte loads), and the
// two points' semantic-logit gradients.  A neighbouring pair is one 32-bit word of a K-major bf16 tile row: no shuffles.
struct SemLoads {
  float hx[kSemH], hy[kSemH], ex[kSemE], ey[kSemE], sx[kSemS], sy[kSemS];     // .x = point 2m, .y = point 2m+1
  float gs0[4], gs1[4];
  bool valid0, valid1;
};
// fp32 pair -> bf16x2 hi word (low half = first point) and bf16x2 word of the residuals
__device__ __forceinline__ void split_pair_bf16(float2 v, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v.y), "f"(v.x));
  const float rx = v.x - __uint_as_float(hi << 16), ry = v.y - __uint_as_float(hi & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(ry), "f"(rx));
}

// Round 2: the 64-point tiles are filled and consumed as two 32-point halves (K-steps 0-1 / 2-3 of the same SWIZZLE_128B rows), each
// with its own ready / done barrier pair: the eight fill warps convert half h+1 while the issuer warp's MMAs read half h, and every
// lane keeps the loads of the NEXT half in flight (second register set) while it converts the current one.  Lane = point, warp =
// feature eighth.  Before: one 64-point buffer, loads -> wait -> fill -> MMA in sequence, 2.0 TB/s.
__global__ void __launch_bounds__(kSemThreads, 1) k_sem_wgrad(const __grid_constant__ WgParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  WgSmem sm;
  wg_carve(base, &sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;
  const long long nhalf = (P.P + kHalfPts - 1) / kHalfPts;
  const long long my_n = (nhalf - blockIdx.x + gridDim.x - 1) / gridDim.x;     // >= 1 (grid <= nhalf)
  if (t == 0) {
    for (int h = 0; h < 2; ++h) { mbar_init(smem_u32(&sm.ready[h]), kSemWarps); mbar_init(smem_u32(&sm.done[h]), 1); }
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(smem_u32(sm.tmem_ptr), kWgTmemCols); tmem_relinquish(); }
  // B2 rows >= 8 are never written by the fill: zero the whole small tile once;  W2 -> smem
  for (int i = t; i < kRowsB2 * 128 / 4; i += kSemThreads) {
    reinterpret_cast<uint32_t*>(sm.b2[0])[i] = 0u;
    reinterpret_cast<uint32_t*>(sm.b2[1])[i] = 0u;
  }
  for (int i = t; i < 4 * 128; i += kSemThreads) sm.w2[i] = (i / 128 < P.sem_dim) ? __ldg(&P.w_s2[i]) : 0.f;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *sm.tmem_ptr;
  __shared__ long long tr[4][24];      // [issue begin, issue end, warp-0 done-wait begin, end][half]
  __shared__ long long tw[kSemWarps][4][4];   // [warp][half 8..11][fill start, done seen, h stored, fill end]
  const bool trace = P.trace && blockIdx.x == 0;

  const uint32_t a[2] = {smem_u32(sm.a[0]), smem_u32(sm.a[1])}, a2[2] = {smem_u32(sm.a2[0]), smem_u32(sm.a2[1])};
  const uint32_t b[2] = {smem_u32(sm.b[0]), smem_u32(sm.b[1])}, b2[2] = {smem_u32(sm.b2[0]), smem_u32(sm.b2[1])};
  const uint32_t id256 = make_idesc_bf16(256), id64 = make_idesc_bf16(64), id16 = make_idesc_bf16(16);
  // 3 x 2 x 3 MMAs per half (hi.hi, lo.hi, hi.lo; M=128, N=256/64/16, K=16), issued by the half's designated warp once all have arrived
  auto issue = [&](int half, long long it) {
    tc_fence_after();
    if (elect_one()) {
      if (trace && it < 24) tr[0][it] = clock64();
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
        const int pa = (pass == 1) ? 1 : 0, pb = (pass == 2) ? 1 : 0;     // hi.hi, lo.hi, hi.lo
#pragma unroll
        for (int k2 = 0; k2 < 2; ++k2) {
          const int ks = 2 * half + k2;
          const uint32_t acc = (it > 0 || pass > 0 || k2 > 0) ? 1u : 0u;
          const uint64_t da = make_sw128_desc(a[pa] + ks * 32), db = make_sw128_desc(b[pb] + ks * 32);
          umma_ss(tm + kColD1, da, db, id256, acc);
          umma_ss(tm + kColD1b, da, make_sw128_desc(b[pb] + 256 * 128 + ks * 32), id64, acc);
          umma_ss(tm + kColD2, make_sw128_desc(a2[pa] + ks * 32), make_sw128_desc(b2[pb] + ks * 32), id16, acc);
        }
      }
      umma_commit(smem_u32(&sm.done[half]));
      if (trace && it < 24) tr[1][it] = clock64();
    }
    __syncwarp();
  };
  if (warp >= kSemWarps) {
    // ================= issuer warpgroup: warp 8 issues every half's MMAs, warps 9-11 only give their registers away ==========
    // 12 warps launch at 168 registers; this group drops to 24 (frees 128 x 144), the 8 fill warps grow to 240 (take 256 x 72).
    // A fill warp that also issues -- ~1.1 k cycles per half -- falls behind the others and becomes what every half waits for
    // (per-warp timelines in profiles/r02_notes.md).
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (warp == kSemIssuer)
      for (long long it = 0; it < my_n; ++it) {
        const int half = (int)(it & 1);
        mbar_wait(smem_u32(&sm.ready[half]), (uint32_t)((it >> 1) & 1), 700 + half);
        issue(half, it);
      }
  } else {
    // ================= fill warps =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");
    const int e = warp;                                  // feature eighth: h 32e.., gamma 8e.., s0 units 16e..
    float gb2_acc[4] = {0.f, 0.f, 0.f, 0.f};
    // h / s0 (and gamma when the forward pass saved it) arrive in the blocked layout (internal.h: sem_saves_blocked): the half is
    // one group of 32 points and a feature's 32 values are one 128-byte line.  Lanes 0-15 read point pairs of feature f, lanes
    // 16-31 of feature f+1: every load instruction of the warp covers two whole lines.
    const int m = lane & 15, fb = lane >> 4;             // point pair, feature parity
    auto load = [&](SemLoads& L, long long it) {
      const long long grp = blockIdx.x + it * gridDim.x;
      const long long p = grp * kHalfPts + 2 * m;
      const bool v0 = p < P.P, v1 = p + 1 < P.P;
      L.valid0 = v0; L.valid1 = v1;
      // NOTHING here may consume a loaded value (a warp issues in order: one dependent select behind every load serialises the
      // loads -- that, not bandwidth, held rounds 1-2 of this kernel at 2 TB/s): predicated loads into zeroed registers; the odd
      // tail point (valid0 && !valid1, last pair of an odd point count) is masked when the values are used
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        L.gs0[c] = 0.f; L.gs1[c] = 0.f;
        if (v0 && c < P.sem_dim) L.gs0[c] = __ldg(&P.g_raw[p * P.C + 4 + c]);
        if (v1 && c < P.sem_dim) L.gs1[c] = __ldg(&P.g_raw[(p + 1) * P.C + 4 + c]);
      }
      // both points or none: groups are whole inside the buffers
#define NSOS_LD2(X, Y, PTR)                                                              \
  {                                                                                      \
    float2 v_ = make_float2(0.f, 0.f);                                                   \
    if (v0) v_ = __ldg(reinterpret_cast<const float2*>(PTR));                            \
    X = v_.x; Y = v_.y;                                                                  \
  }
      // lanes 16-31 take the feature FOUR rows further: tile rows r and r+4 sit in different bank halves (conflict-free stores)
      const float* hb = P.h + (grp * 256 + e * (2 * kSemH) + 4 * fb) * kHalfPts + 2 * m;
#pragma unroll
      for (int k = 0; k < kSemH; ++k) NSOS_LD2(L.hx[k], L.hy[k], hb + (8 * (k >> 2) + (k & 3)) * kHalfPts)
      if (!P.sem_coord) {
#pragma unroll
        for (int k = 0; k < kSemE; ++k) { L.ex[k] = 0.f; L.ey[k] = 0.f; }
      } else if (P.enc_blocked) {
        const float* eb = P.enc + (grp * P.enc_ld + e * (2 * kSemE) + fb) * kHalfPts + 2 * m;
#pragma unroll
        for (int k = 0; k < kSemE; ++k) NSOS_LD2(L.ex[k], L.ey[k], eb + 2 * k * kHalfPts)
      } else {                                           // row-major gamma (k_encode_pts, replay path)
#pragma unroll
        for (int k = 0; k < kSemE; ++k) {
          const float* er = P.enc + p * P.enc_ld + e * (2 * kSemE) + fb + 2 * k;
          L.ex[k] = 0.f; L.ey[k] = 0.f;
          if (v0) L.ex[k] = __ldg(er);
          if (v1) L.ey[k] = __ldg(er + P.enc_ld);
        }
      }
      const float* sb = P.s0 + (grp * 128 + e * (2 * kSemS) + 4 * fb) * kHalfPts + 2 * m;
#pragma unroll
      for (int k = 0; k < kSemS; ++k) NSOS_LD2(L.sx[k], L.sy[k], sb + (8 * (k >> 2) + (k & 3)) * kHalfPts)
#undef NSOS_LD2
    };
    // Tile row R = 8*r8 + 2*(k & 3) + fb of a K-major SWIZZLE_128B tile; this lane's word is point pair 16*half + m:
    // byte offset = r8*1024 + (R & 7)*128 + (((4*half + (m >> 2)) ^ (R & 7)) << 4) + (m & 3)*4
    auto fill = [&](SemLoads& L, int half, long long it) {
      // the MMAs that read this half two iterations ago have completed
      if (trace && t == 0 && it < 24) tr[2][it] = clock64();
      const bool tw_on = trace && lane == 0 && it >= 8 && it < 12;
      if (tw_on) tw[warp][it - 8][0] = clock64();
      if (it >= 2) {                                     // (polling warps back off: they share schedulers with the warps still converting)
        const uint32_t bar = smem_u32(&sm.done[half]), par = (uint32_t)(((it >> 1) - 1) & 1);
        for (uint32_t spins = 0; !mbar_try_wait(bar, par); ++spins) {
          __nanosleep(40);
          if (spins > (1u << 24)) { printf("[nerfsos] k_sem_wgrad: done[%d] timed out\n", half); __trap(); }
        }
        tc_fence_after();
      }
      if (trace && t == 0 && it < 24) tr[3][it] = clock64();
      if (tw_on) tw[warp][it - 8][1] = clock64();
      const int xk = 4 * half + (m >> 2);               // 16-byte chunk of this lane's word before the swizzle
      auto offs = [&](int r7) { return (uint32_t)(r7 * 128 + ((xk ^ r7) << 4) + (m & 3) * 4); };
      uint32_t off[4];                                   // per (k & 3), rows (k & 3) + 4*fb: everything of the offset except r8*1024
#pragma unroll
      for (int j = 0; j < 4; ++j) off[j] = offs(j + 4 * fb);
      const bool v1 = L.valid1;
      // 32-bit shared-window addresses: tile base + 8-row group + swizzled offset (generic 64-bit pointer arithmetic per store
      // was a third of the fill's instructions)
      auto put = [&](const uint32_t (&tile)[2], int r8, uint32_t o, float x, float y) {
        uint32_t hi, lo;
        split_pair_bf16(make_float2(x, v1 ? y : 0.f), hi, lo);          // (odd tail point: its slot in the padded group holds garbage)
        st_shared_u32(tile[0] + (uint32_t)r8 * 1024u + o, hi);
        st_shared_u32(tile[1] + (uint32_t)r8 * 1024u + o, lo);
      };
      // ---- B rows 0..255: h (feature 2*kSemH*e + 8*(k >> 2) + (k & 3) + 4*fb)
#pragma unroll
      for (int k = 0; k < kSemH; ++k) put(b, (2 * kSemH / 8) * e + (k >> 2), off[k & 3], L.hx[k], L.hy[k]);
      if (tw_on) tw[warp][it - 8][2] = clock64();
      // ---- B rows 256..319: gamma (63) and the constant-one feature (-> db0); feature 2*kSemE*e + 2k + fb
#pragma unroll
      for (int k = 0; k < kSemE; ++k) {
        const int f = e * (2 * kSemE) + 2 * k + fb;
        float x = L.ex[k], y = L.ey[k];
        if (f >= P.enc_dim) { x = 0.f; y = 0.f; }
        if (f == 63) { x = L.valid0 ? 1.f : 0.f; y = L.valid1 ? 1.f : 0.f; }
        put(b, 32 + (f >> 3), offs(f & 7), x, y);
      }
      // ---- A2 = s0^T and A = g_s0^T, unit 2*kSemS*e + 8*(k >> 2) + (k & 3) + 4*fb
#pragma unroll
      for (int k = 0; k < kSemS; ++k) {
        const int u = e * (2 * kSemS) + 8 * (k >> 2) + (k & 3) + 4 * fb, r8 = (2 * kSemS / 8) * e + (k >> 2);
        put(a2, r8, off[k & 3], L.sx[k], L.sy[k]);
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {                    // zero rows beyond sem_dim
          const float wv = sm.w2[c * 128 + u];
          d0 = fmaf(L.gs0[c], wv, d0); d1 = fmaf(L.gs1[c], wv, d1);
        }
        put(a, r8, off[k & 3], L.sx[k] > 0.f ? d0 : 0.f, (v1 && L.sy[k] > 0.f) ? d1 : 0.f);
      }
      // ---- B2 = g_sem^T (rows 0..3; rows >= sem_dim are zero, rows 4..15 were cleared once)
      if (e == kSemWarps - 1) {
#pragma unroll
        for (int k = 0; k < 2; ++k)                     // row c = 2k + fb
          put(b2, 0, offs(2 * k + fb), fb ? L.gs0[2 * k + 1] : L.gs0[2 * k], fb ? L.gs1[2 * k + 1] : L.gs1[2 * k]);
        if (fb == 0) {
#pragma unroll
          for (int c = 0; c < 4; ++c) gb2_acc[c] += L.gs0[c] + L.gs1[c];
        }
      }
      // ---- one arrival per warp; the half's MMAs are issued by warp (it mod 16): the duty ROTATES.  (With "whoever arrives last
      // issues", the ~1.1 k cycles of issuing made the same warp the last one again and again -- per-warp timeline, NSOS_WG_TRACE=1,
      // profiles/r02_notes.md: one warp 7 k cycles behind the other fifteen.)
      if (tw_on) tw[warp][it - 8][3] = clock64();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&sm.ready[half]));
    };
    SemLoads L0, L1;
    load(L0, 0);
    for (long long it = 0; it < my_n; it += 2) {
      if (it + 1 < my_n) load(L1, it + 1);
      fill(L0, 0, it);
      if (it + 2 < my_n) load(L0, it + 2);
      if (it + 1 < my_n) fill(L1, 1, it + 1);
    }
    // ---- epilogue: TMEM partial sums -> this CTA's slice of the scratch buffer, column-major (one line per warp store):
    // cols 0..318 dW0, 319 db0, 320..323 dW2^T, 324..327 db2 (row 0)
    mbar_wait(smem_u32(&sm.done[(my_n - 1) & 1]), (uint32_t)(((my_n - 1) >> 1) & 1), 720);     // the last commit covers every earlier MMA
    tc_fence_after();
    constexpr int kColShare = 20 / (kSemWarps / 4);     // D1 chunks (16 columns) per warp of a lane quarter
    const int q4 = warp & 3, hf = warp >> 2;
    const int u = q4 * 32 + lane;
    const uint32_t tm_lane = tm + ((uint32_t)(q4 * 32) << 16);
    float* part = P.part + (size_t)blockIdx.x * kPartFloatsPerCta + u;
    for (int c = hf * kColShare; c < hf * kColShare + kColShare; ++c) {
      uint32_t r[16];
      tmem_ld16(tm_lane + kColD1 + c * 16, r);
      tmem_wait_ld_fence16(r);
#pragma unroll
      for (int j = 0; j < 16; ++j) part[(size_t)(c * 16 + j) * kPartRows] = __uint_as_float(r[j]);
    }
    if (hf == kSemWarps / 4 - 1) {
      uint32_t r[16];
      tmem_ld16(tm_lane + kColD2, r);
      tmem_wait_ld_fence16(r);
#pragma unroll
      for (int c = 0; c < 4; ++c) part[(size_t)(320 + c) * kPartRows] = __uint_as_float(r[c]);
    }
    if (e == kSemWarps - 1) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float s = gb2_acc[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) P.part[(size_t)blockIdx.x * kPartFloatsPerCta + (size_t)(324 + c) * kPartRows] = s;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (trace && t == 0 && my_n >= 12)
    for (int w = 0; w < kSemWarps; ++w)                 // per-warp timeline of halves 8..11: fill start, done seen, h stored, fill end
      printf("[ww] warp %2d: h8 %lld %lld %lld %lld | h9 %lld %lld %lld %lld | h10 %lld %lld %lld %lld | h11 %lld %lld %lld %lld\n", w,
             tw[w][0][0] - tr[2][8], tw[w][0][1] - tr[2][8], tw[w][0][2] - tr[2][8], tw[w][0][3] - tr[2][8],
             tw[w][1][0] - tr[2][8], tw[w][1][1] - tr[2][8], tw[w][1][2] - tr[2][8], tw[w][1][3] - tr[2][8],
             tw[w][2][0] - tr[2][8], tw[w][2][1] - tr[2][8], tw[w][2][2] - tr[2][8], tw[w][2][3] - tr[2][8],
             tw[w][3][0] - tr[2][8], tw[w][3][1] - tr[2][8], tw[w][3][2] - tr[2][8], tw[w][3][3] - tr[2][8]);
  if (trace && t == 0)
    for (int i = 0; i < 24 && i < my_n; ++i)
      printf("[wg] half %2d: fill starts %7lld, done(it-2) seen %7lld, issue %7lld .. %7lld\n", i, tr[2][i] - tr[2][0], tr[3][i] - tr[2][0],
             tr[0][i] - tr[2][0], tr[1][i] - tr[2][0]);
  if (warp == 0) tmem_dealloc(tm, kWgTmemCols);
}


// ---- general weight gradient: dW[Mo, Ki] += dY[P, Mo]^T . X[P, Ki]  (the all-parameter backward, trainer.py:201 with every parameter
// trainable).  Same transposing fill and bf16 hi/lo contraction over points as k_sem_wgrad, for any layer:
//   A  [128 x 64 pts] = dY[:, m0:m0+128]^T (one 128-row block of dW per blockIdx.y)
//   B  [320 x 64 pts] = [main (256 wide, e.g. the previous layer's activations) ; aux (<= 64 wide, e.g. gamma(x))]^T
// D[:, 0:256] accumulates the `main` columns of dW, D[:, 256:320] the `aux` columns, for the whole life of the CTA.
struct WgGenParams {
  const float* dY; long long ldy; int Mo;
  const float* main; long long ld_main; int main_col;      // may be null
  const float* aux; long long ld_aux; int aux_w, aux_col;  // may be null; aux_w <= 64
  float* dW; long long ldw;
  float* db;                 // optional: db[Mo] += sum_p dY[p, :]  (a constant-one feature in aux row 63; needs aux_w <= 63)
  float* part;               // [gridDim.y][gridDim.x][kPartCols][128] partial sums (k_wgrad_reduce)
  long long P;
};
struct WgGenSmem {
  uint8_t *a[2], *b[2];
  uint64_t *ready, *done;
  uint32_t* tmem_ptr;
};
__host__ __device__ inline size_t wgg_carve(uint8_t* base, WgGenSmem* s) {
  size_t off = 0;
  auto take = [&](size_t bytes, size_t align) { off = (off + align - 1) / align * align; size_t o = off; off += bytes; return o; };
  size_t oa[2], ob[2];
  for (int p = 0; p < 2; ++p) oa[p] = take(kRowsA * 128, 1024);
  for (int p = 0; p < 2; ++p) ob[p] = take(kRowsB * 128, 1024);
  size_t obar = take(16, 8), otp = take(16, 16);
  if (s) {
    for (int p = 0; p < 2; ++p) { s->a[p] = base + oa[p]; s->b[p] = base + ob[p]; }
    s->ready = (uint64_t*)(base + obar); s->done = s->ready + 1; s->tmem_ptr = (uint32_t*)(base + otp);
  }
  return off;
}

__global__ void __launch_bounds__(kWgThreads, 1) k_wgrad_gen(const __grid_constant__ WgGenParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  WgGenSmem sm;
  wgg_carve(base, &sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;
  const int m0 = blockIdx.y * 128;
  const long long nslabs = (P.P + kSlabPts - 1) / kSlabPts;
  const long long my_slabs = (nslabs - blockIdx.x + gridDim.x - 1) / gridDim.x;     // >= 1 (grid.x <= nslabs)
  const bool has_main = P.main != nullptr, has_aux = P.aux != nullptr || P.db != nullptr;     // the bias column lives in the aux block
  if (t == 0) {
    mbar_init(smem_u32(sm.ready), kWgWorkers);
    mbar_init(smem_u32(sm.done), 1);
    fence_mbar_init();
  }
  if (warp == kWgMmaWarp) { tmem_alloc(smem_u32(sm.tmem_ptr), kWgTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *sm.tmem_ptr;

  if (warp == kWgMmaWarp) {
    const uint32_t a[2] = {smem_u32(sm.a[0]), smem_u32(sm.a[1])}, b[2] = {smem_u32(sm.b[0]), smem_u32(sm.b[1])};
    const uint32_t id256 = make_idesc_bf16(256), id64 = make_idesc_bf16(64);
    for (long long it = 0; it < my_slabs; ++it) {
      mbar_wait(smem_u32(sm.ready), (uint32_t)(it & 1), 730);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const int pa = (pass == 1) ? 1 : 0, pb = (pass == 2) ? 1 : 0;     // hi.hi, lo.hi, hi.lo
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t acc = (it > 0 || pass > 0 || ks > 0) ? 1u : 0u;
            const uint64_t da = make_sw128_desc(a[pa] + ks * 32);
            if (has_main) umma_ss(tm + kColD1, da, make_sw128_desc(b[pb] + ks * 32), id256, acc);
            if (has_aux) umma_ss(tm + kColD1b, da, make_sw128_desc(b[pb] + 256 * 128 + ks * 32), id64, acc);
          }
        }
        umma_commit(smem_u32(sm.done));
      }
      __syncwarp();
    }
  } else {
    const int pgp = warp & 1, q = warp >> 1;            // point group (32 points), feature quarter
    const int pl = 32 * pgp + lane, wd = pl >> 1;
    for (long long it = 0; it < my_slabs; ++it) {
      const long long slab = blockIdx.x + it * gridDim.x;
      const long long p = slab * kSlabPts + pl;
      const bool valid = p < P.P;
      float hv[4][8];                                   // rolling window over this thread's 64 `main` values
      const float* hrow = has_main ? P.main + p * P.ld_main + q * 64 : nullptr;
      if (has_main) {
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) load8(hrow + 8 * g8, valid, hv[g8]);
      }
      float gv[4][8];                                   // this thread's 32 dY values (units m0 + q*32 ..)
      {
        const float* grow = P.dY + p * P.ldy + m0 + q * 32;
        const bool gval = valid && (m0 + q * 32 < P.Mo);
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) load8(grow + 8 * g8, gval, gv[g8]);
      }
      if (it > 0) { mbar_wait(smem_u32(sm.done), (uint32_t)((it - 1) & 1), 740); tc_fence_after(); }
      if (has_main) {
#pragma unroll
        for (int g8 = 0; g8 < 8; ++g8) {
          put8(sm.b[0], sm.b[1], q * 64 + 8 * g8, hv[g8 & 3], lane, wd);
          if (g8 + 4 < 8) load8(hrow + 8 * (g8 + 4), valid, hv[g8 & 3]);
        }
      }
      if (has_aux) {
        // aux rows (ld_aux may be odd-sized: scalar loads), zero beyond aux_w
#pragma unroll
        for (int g8 = 0; g8 < 2; ++g8) {
          const int e0 = q * 16 + 8 * g8;
          float ev[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            ev[i] = (valid && P.aux && e0 + i < P.aux_w) ? __ldg(&P.aux[p * P.ld_aux + e0 + i]) : 0.f;
            if (P.db && e0 + i == 63) ev[i] = valid ? 1.f : 0.f;
          }
          put8(sm.b[0], sm.b[1], 256 + e0, ev, lane, wd);
        }
      }
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) put8(sm.a[0], sm.a[1], q * 32 + 8 * g8, gv[g8], lane, wd);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(smem_u32(sm.ready));
    }
    // ---- epilogue: TMEM partial sums -> this CTA's slice of the scratch buffer (column-major; rows >= Mo - m0 are never read)
    mbar_wait(smem_u32(sm.done), (uint32_t)((my_slabs - 1) & 1), 750);
    tc_fence_after();
    const int q4 = warp & 3, hf = warp >> 2;
    const uint32_t tm_lane = tm + ((uint32_t)(q4 * 32) << 16);
    for (int c = hf * 10; c < hf * 10 + 10; ++c) {
      if (c < 16 ? !has_main : !has_aux) continue;
      uint32_t r[16];
      tmem_ld16(tm_lane + kColD1 + c * 16, r);
      tmem_wait_ld_fence16(r);
      float* part = P.part + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * kPartFloatsPerCta + (q4 * 32 + lane);
#pragma unroll
      for (int j = 0; j < 16; ++j) part[(size_t)(c * 16 + j) * kPartRows] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWgMmaWarp) tmem_dealloc(tm, kWgTmemCols);
}


// ---- k_wgrad_gen, second version: rows staged through shared memory with cp.async ---------------------------------------------
// k_wgrad_gen's fill threads own a POINT (a row of the row-major operands), so each of their 16-byte loads touches 32 different
// 128-byte lines per warp instruction: 6 k L1 tag cycles per 64-point slab against 2 k cycles of MMAs, and the loads of a slab are
// only requested once the previous slab has been converted.  Here the operands of a 32-point HALF slab (main 1 KB + dY 512 B +
// aux 256 B per point) are copied by all 256 fill threads with coalesced 16-byte cp.async (a warp instruction = 512 contiguous
// bytes of one row, zero-filled beyond P) into a staging image of pitch kStRow = 1808 bytes (113 16-byte units: unit c of row r
// sits in bank group (r + c) mod 8, so lanes that read the same unit of consecutive rows do not conflict).  Two images: while
// half u is converted, half u+1 is in flight; the request for half u+2 goes out as soon as every thread has picked its values up.
// A lane converts TWO points of four features at a time (cvt.rn.bf16x2: one 32-bit word of a K-major tile row per plane, no
// shuffles).  The contraction runs over points, so their order inside a half is free as long as A and B agree: K position 2m
// holds staging row m, 2m+1 row m+16 (rows m / m+16 of consecutive lanes are conflict-free; 2m / 2m+1 would be 2-way).
// Lanes 0-15 write tile rows 8k..8k+3, lanes 16-31 rows 8k+4..8k+7 of the same 16 words: their swizzled chunks differ in bit 2,
// all 32 banks.  Tiles are filled and consumed per half (K-steps 0-1 / 2-3) with a ready / done barrier pair each, like k_sem_wgrad.
constexpr int kStRow = 1808, kStDy = 1024, kStAux = 1536;
constexpr int kStHalf = 32 * kStRow;
struct WgGen2Smem {
  uint8_t *a[2], *b[2], *st[2];
  uint64_t *ready, *done;       // [2] each: one pair per half
  uint64_t *full, *empty;       // [2] each: staging images (loader-warp variant)
  uint32_t* tmem_ptr;
};
__host__ __device__ inline size_t wgg2_carve(uint8_t* base, WgGen2Smem* s) {
  size_t off = 0;
  auto take = [&](size_t bytes, size_t align) { off = (off + align - 1) / align * align; size_t o = off; off += bytes; return o; };
  size_t oa[2], ob[2], os[2];
  for (int p = 0; p < 2; ++p) oa[p] = take(kRowsA * 128, 1024);
  for (int p = 0; p < 2; ++p) ob[p] = take(kRowsB * 128, 1024);
  for (int p = 0; p < 2; ++p) os[p] = take(kStHalf, 16);
  size_t obar = take(64, 8), otp = take(16, 16);
  if (s) {
    for (int p = 0; p < 2; ++p) { s->a[p] = base + oa[p]; s->b[p] = base + ob[p]; s->st[p] = base + os[p]; }
    s->ready = (uint64_t*)(base + obar); s->done = s->ready + 2; s->full = s->ready + 4; s->empty = s->ready + 6;
    s->tmem_ptr = (uint32_t*)(base + otp);
  }
  return off;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {     // src_bytes 0: 16 zero bytes
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {       // the arrival is one of the barrier's expected count (noinc)
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// four features x two points -> four words per plane; toff[i] = byte offset of tile row (8k + 4g + i)'s word inside 8-row group 0
__device__ __forceinline__ void put_quad(uint8_t* hi_tile, uint8_t* lo_tile, uint32_t grp_off, const uint32_t (&toff)[4], float4 xa, float4 xb) {
  uint32_t hi, lo;
  split_pair_bf16(make_float2(xa.x, xb.x), hi, lo);
  *reinterpret_cast<uint32_t*>(hi_tile + grp_off + toff[0]) = hi; *reinterpret_cast<uint32_t*>(lo_tile + grp_off + toff[0]) = lo;
  split_pair_bf16(make_float2(xa.y, xb.y), hi, lo);
  *reinterpret_cast<uint32_t*>(hi_tile + grp_off + toff[1]) = hi; *reinterpret_cast<uint32_t*>(lo_tile + grp_off + toff[1]) = lo;
  split_pair_bf16(make_float2(xa.z, xb.z), hi, lo);
  *reinterpret_cast<uint32_t*>(hi_tile + grp_off + toff[2]) = hi; *reinterpret_cast<uint32_t*>(lo_tile + grp_off + toff[2]) = lo;
  split_pair_bf16(make_float2(xa.w, xb.w), hi, lo);
  *reinterpret_cast<uint32_t*>(hi_tile + grp_off + toff[3]) = hi; *reinterpret_cast<uint32_t*>(lo_tile + grp_off + toff[3]) = lo;
}

// NW fill warps (8 or 16: warp e converts the feature groups [e*32/NW, (e+1)*32/NW) of main, e*16/NW.. of dY and, for e < 8, aux
// group e) + the issuer warp.  With 16 warps every scheduler has four fill warps to choose from instead of two.
// NL = 0: the fill threads request the next halves themselves (cp.async groups + named barriers).  NL = 4: four LOADER warps own
// the requests and hand the images over through mbarriers (full: cp.async.mbarrier.arrive of the loader threads; empty: the fill
// threads after their reads) -- no fill thread then has copies in flight when it executes fence.proxy.async (MEMBAR.ALL.CTA in
// SASS) at the end of a half.
template <int NW, int NL>
__global__ void __launch_bounds__(32 * (NW + 1 + NL), 1) k_wgrad_gen2(const __grid_constant__ WgGenParams P) {
  constexpr int NT = 32 * NW, GM = 32 / NW, GD = 16 / NW;
  constexpr int RT = NL ? 32 * NL : NT;                      // threads that issue the requests
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  WgGen2Smem sm;
  wgg2_carve(base, &sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;
  const int m0 = blockIdx.y * 128;
  const long long nslabs = (P.P + kSlabPts - 1) / kSlabPts;
  const long long my_slabs = (nslabs - blockIdx.x + gridDim.x - 1) / gridDim.x;     // >= 1 (grid.x <= nslabs)
  const long long nunits = 2 * my_slabs;                                            // 32-point halves
  const bool has_main = P.main != nullptr, has_aux = P.aux != nullptr || P.db != nullptr;     // the bias column lives in the aux block
  if (t == 0) {
    for (int h = 0; h < 2; ++h) {
      mbar_init(smem_u32(sm.ready + h), NT); mbar_init(smem_u32(sm.done + h), 1);
      mbar_init(smem_u32(sm.full + h), RT); mbar_init(smem_u32(sm.empty + h), NT);
    }
    fence_mbar_init();
  }
  if (warp == NW) { tmem_alloc(smem_u32(sm.tmem_ptr), kWgTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *sm.tmem_ptr;

  if (warp == NW) {
    const uint32_t a[2] = {smem_u32(sm.a[0]), smem_u32(sm.a[1])}, b[2] = {smem_u32(sm.b[0]), smem_u32(sm.b[1])};
    const uint32_t id256 = make_idesc_bf16(256), id64 = make_idesc_bf16(64);
    for (long long u = 0; u < nunits; ++u) {
      const int h = (int)(u & 1);
      mbar_wait(smem_u32(sm.ready + h), (uint32_t)((u >> 1) & 1), 731);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const int pa = (pass == 1) ? 1 : 0, pb = (pass == 2) ? 1 : 0;     // hi.hi, lo.hi, hi.lo
#pragma unroll
          for (int k2 = 0; k2 < 2; ++k2) {
            const int ks = 2 * h + k2;                                      // K-steps of this half
            const uint32_t acc = (u > 0 || pass > 0 || k2 > 0) ? 1u : 0u;
            const uint64_t da = make_sw128_desc(a[pa] + ks * 32);
            if (has_main) umma_ss(tm + kColD1, da, make_sw128_desc(b[pb] + ks * 32), id256, acc);
            if (has_aux) umma_ss(tm + kColD1b, da, make_sw128_desc(b[pb] + 256 * 128 + ks * 32), id64, acc);
          }
        }
        umma_commit(smem_u32(sm.done + h));
      }
      __syncwarp();
    }
  } else {
    const bool loader = warp > NW;
    const int rt = NL ? t - 32 * (NW + 1) : t;               // index among the requesting threads
    const int m = lane & 15, g = lane >> 4, e = warp;        // point pair (staging rows m, m+16), tile-row half, feature share
    const bool aux_rows = P.aux != nullptr, aux_warp = has_aux && e < 8;
    // Requests of half u: requesting thread rt moves unit (rt + RT i) of the main / dY / aux block (64 / 32 / 16 units per row): a fixed unit
    // column and rows RT/64 (RT/32, RT/16) apart, so the addresses are one pointer per block that advances by a constant.
    const int rm = rt >> 6, rd = rt >> 5, rx = rt >> 4;         // first row of this thread in each block
    const uint32_t dm = (uint32_t)(rm * kStRow + 16 * (rt & 63)), dd = (uint32_t)(rd * kStRow + kStDy + 16 * (rt & 31)),
                   dx = (uint32_t)(rx * kStRow + kStAux + 16 * (rt & 15));
    const float* const bm = has_main ? P.main + 4 * (rt & 63) : nullptr;
    const float* const bd = P.dY + m0 + 4 * (rt & 31);
    const float* const bx = aux_rows ? P.aux + 4 * (rt & 15) : nullptr;
    const uint32_t S0 = smem_u32(sm.st[0]);
    // halves are requested in order, so the request state is carried along: next half, its first point, one source pointer per
    // block.  Rows beyond P are zero-filled (source size 0; the pointer is parked on row 0 of the block, which always exists).
    long long rq_u = 0, rq_p0 = (long long)blockIdx.x * kSlabPts;
    const int step_odd = (int)gridDim.x * kSlabPts - 32;     // points from an odd half to the next even one (pitches < 2^16: products fit 32 bits)
    const int em[2] = {32 * (int)P.ld_main, step_odd * (int)P.ld_main}, ed[2] = {32 * (int)P.ldy, step_odd * (int)P.ldy},
              ex[2] = {32 * (int)P.ld_aux, step_odd * (int)P.ld_aux};
    const float* pm = bm + (rq_p0 + rm) * P.ld_main;
    const float* pd = bd + (rq_p0 + rd) * P.ldy;
    const float* px = bx + (rq_p0 + rx) * P.ld_aux;
    auto request = [&]() {
      if (rq_u < nunits) {
        const long long left = P.P - rq_p0;
        const int nv = left < 32 ? (left < 0 ? 0 : (int)left) : 32;                   // valid rows of this half
        const uint32_t S = S0 + (uint32_t)(rq_u & 1) * kStHalf;                       // the two images are adjacent
        if (has_main) {
          const float* src = pm;
#pragma unroll
          for (int i = 0; i < 2048 / RT; ++i) {
            const bool v = rm + (RT / 64) * i < nv;
            cp_async16(S + dm + (uint32_t)((RT / 64) * i * kStRow), v ? src : bm, v ? 16u : 0u);
            src += (RT / 64) * P.ld_main;
          }
        }
        {
          const float* src = pd;
#pragma unroll
          for (int i = 0; i < 1024 / RT; ++i) {
            const bool v = rd + (RT / 32) * i < nv;
            cp_async16(S + dd + (uint32_t)((RT / 32) * i * kStRow), v ? src : bd, v ? 16u : 0u);
            src += (RT / 32) * P.ldy;
          }
        }
        if (aux_rows) {
          const float* src = px;
#pragma unroll
          for (int i = 0; i < 512 / RT; ++i) {
            const bool v = rx + (RT / 16) * i < nv;
            cp_async16(S + dx + (uint32_t)((RT / 16) * i * kStRow), v ? src : bx, v ? 16u : 0u);
            src += (RT / 16) * P.ld_aux;
          }
        }
      }
      if (NL) cp_async_arrive(smem_u32(sm.full + (rq_u & 1)));     // arrives once this thread's copies above have landed
      else cp_async_commit();        // (an empty group past the end keeps the group count uniform)
      const bool odd = (rq_u & 1) != 0;
      rq_p0 += odd ? step_odd : 32; pm += odd ? em[1] : em[0]; pd += odd ? ed[1] : ed[0]; px += odd ? ex[1] : ex[0];
      ++rq_u;
    };
    if (NL && loader) {
      for (long long u = 0; u < nunits; ++u) {
        if (u >= 2) mbar_wait(smem_u32(sm.empty + (u & 1)), (uint32_t)(((u >> 1) - 1) & 1), 761);
        request();
      }
    } else {
    if (!NL) { request(); request(); }
    for (long long u = 0; u < nunits; ++u) {
      const int h = (int)(u & 1);
      const long long it = u >> 1;
      if (NL) mbar_wait(smem_u32(sm.full + h), (uint32_t)(it & 1), 771);     // the loaders' copies of half u have landed
      else {
        cp_async_wait_but_one();                             // this thread's pieces of half u have landed
        named_bar_sync(1, NT);                               // ... and everybody else's
      }
      const uint8_t* ra = sm.st[0] + h * kStHalf + m * kStRow + 16 * g;    // rows m / m+16; the lane's 16-byte unit of 8-feature group k: 2k + g
      const uint8_t* rb = ra + 16 * kStRow;
      float4 xa[GM + GD + 1], xb[GM + GD + 1];               // GM main groups, GD dY groups, one aux group
#pragma unroll
      for (int k = 0; k < GM; ++k) {
        xa[k] = *reinterpret_cast<const float4*>(ra + 32 * (GM * e + k));
        xb[k] = *reinterpret_cast<const float4*>(rb + 32 * (GM * e + k));
      }
#pragma unroll
      for (int k = 0; k < GD; ++k) {
        xa[GM + k] = *reinterpret_cast<const float4*>(ra + kStDy + 32 * (GD * e + k));
        xb[GM + k] = *reinterpret_cast<const float4*>(rb + kStDy + 32 * (GD * e + k));
      }
      if (aux_warp) {
        xa[GM + GD] = *reinterpret_cast<const float4*>(ra + kStAux + 32 * e);
        xb[GM + GD] = *reinterpret_cast<const float4*>(rb + kStAux + 32 * e);
      }
      if (NL) mbar_arrive(smem_u32(sm.empty + h));           // the image is free again
      else {
        named_bar_sync(1, NT);
        request();                                           // half u + 2
      }
      if (aux_warp) {                                        // aux features 8e + 4g + i: zero beyond aux_w, constant one in feature 63
        const long long p0 = ((long long)blockIdx.x + it * gridDim.x) * kSlabPts + 32 * h;
        const int f0 = 8 * e + 4 * g;
        float va[4] = {xa[GM + GD].x, xa[GM + GD].y, xa[GM + GD].z, xa[GM + GD].w};
        float vb[4] = {xb[GM + GD].x, xb[GM + GD].y, xb[GM + GD].z, xb[GM + GD].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (!aux_rows || f0 + i >= P.aux_w) { va[i] = 0.f; vb[i] = 0.f; }
          if (P.db && f0 + i == 63) { va[i] = (p0 + m < P.P) ? 1.f : 0.f; vb[i] = (p0 + m + 16 < P.P) ? 1.f : 0.f; }
        }
        xa[GM + GD] = make_float4(va[0], va[1], va[2], va[3]);
        xb[GM + GD] = make_float4(vb[0], vb[1], vb[2], vb[3]);
      }
      if (it > 0) { mbar_wait(smem_u32(sm.done + h), (uint32_t)((it - 1) & 1), 741); tc_fence_after(); }
      uint32_t toff[4];                                      // word (16 h + m) of tile rows 4g + i inside an 8-row group
#pragma unroll
      for (int i = 0; i < 4; ++i) toff[i] = (uint32_t)((4 * g + i) * 128 + ((((4 * h + (m >> 2)) ^ (4 * g + i)) << 4) + ((m & 3) << 2)));
      if (has_main) {
#pragma unroll
        for (int k = 0; k < GM; ++k) put_quad(sm.b[0], sm.b[1], (uint32_t)(GM * e + k) * 1024u, toff, xa[k], xb[k]);
      }
      if (aux_warp) put_quad(sm.b[0], sm.b[1], (uint32_t)(32 + e) * 1024u, toff, xa[GM + GD], xb[GM + GD]);
#pragma unroll
      for (int k = 0; k < GD; ++k) put_quad(sm.a[0], sm.a[1], (uint32_t)(GD * e + k) * 1024u, toff, xa[GM + k], xb[GM + k]);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(smem_u32(sm.ready + h));
    }
    // ---- epilogue: TMEM partial sums -> this CTA's slice of the scratch buffer (column-major; rows >= Mo - m0 are never read)
    for (int h = 0; h < 2; ++h) mbar_wait(smem_u32(sm.done + h), (uint32_t)((my_slabs - 1) & 1), 751);
    tc_fence_after();
    constexpr int CPW = 20 / (NW / 4);                       // 16-column chunks per warp of a lane quarter
    const int q4 = warp & 3, hf = warp >> 2;
    const uint32_t tm_lane = tm + ((uint32_t)(q4 * 32) << 16);
    for (int c = hf * CPW; c < hf * CPW + CPW; ++c) {
      if (c < 16 ? !has_main : !has_aux) continue;
      uint32_t r[16];
      tmem_ld16(tm_lane + kColD1 + c * 16, r);
      tmem_wait_ld_fence16(r);
      float* part = P.part + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * kPartFloatsPerCta + (q4 * 32 + lane);
#pragma unroll
      for (int j = 0; j < 16; ++j) part[(size_t)(c * 16 + j) * kPartRows] = __uint_as_float(r[j]);
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NW) tmem_dealloc(tm, kWgTmemCols);
}

}  // namespace

bool sem_saves_blocked(const NetGeom& gc, const NetGeom& gf) {
  return tc_sem_wgrad_supported(gc) && tc_sem_wgrad_supported(gf) && !getenv("NSOS_WGRAD_SIMT");
}
bool tc_sem_wgrad_supported(const NetGeom& g) {
  return g.use_viewdirs && g.use_sem && g.W == 256 && g.enc <= 63 && g.sem_dim >= 1 && g.sem_dim <= 4;
}

// h (256 features) and s0 (128, post-ReLU) in the BLOCKED layout (sem_saves_blocked), starting at a multiple of 32 points;
// enc = gamma(x), 64 features, only read with sem_with_coord: blocked (enc_blocked) or row-major [P,enc_ld]; g_raw [P,C] row-major
// (sem gradients in columns 4..).  Accumulates into the flat gradient buffer `grads` (layout of nsos_param_layout).
size_t tc_wgrad_part_bytes() { return sizeof(float) * kPartMaxCtas * kPartFloatsPerCta; }     // 27.5 MB

int tc_sem_wgrad(const NetGeom& g, const float* prm, float* grads, const float* h, const float* enc, int enc_ld, int enc_blocked,
                 const float* s0, const float* g_raw, int64_t P, void* part, size_t part_bytes, cudaStream_t st) {
  NSOS_REQUIRE(tc_sem_wgrad_supported(g), NSOS_ERR_UNSUPPORTED, "semantic-head wgrad kernel needs W=256, sem_dim<=4");
  NSOS_REQUIRE(part && part_bytes >= tc_wgrad_part_bytes(), NSOS_ERR_WORKSPACE, "tc_sem_wgrad: partial-sum scratch too small");
  if (P <= 0) return NSOS_OK;
  WgParams p;
  memset(&p, 0, sizeof(p));
  p.h = h; p.enc = enc; p.s0 = s0; p.g_raw = g_raw; p.w_s2 = prm + g.w_s2;
  p.gW0 = grads + g.w_s0; p.gb0 = grads + g.b_s0; p.gW2 = grads + g.w_s2; p.gb2 = grads + g.b_s2;
  p.P = P; p.C = g.C; p.enc_dim = g.enc; p.enc_ld = enc_ld; p.enc_blocked = enc_blocked; p.sem_dim = g.sem_dim; p.sem_coord = g.sem_coord; p.ld0 = g.sem_in;
  int dev = 0, sms = 0;
  NSOS_CHECK_CUDA(cudaGetDevice(&dev));
  NSOS_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long nhalf = (P + kHalfPts - 1) / kHalfPts;
  const int grid = (int)std::min<long long>(nhalf, std::min(sms, kPartMaxCtas));
  const size_t need = wg_carve(nullptr, nullptr) + 1024;
  p.part = (float*)part;
  if (const char* e = getenv("NSOS_WG_TRACE")) p.trace = atoi(e);
  NSOS_CHECK_CUDA(cudaFuncSetAttribute(k_sem_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
  k_sem_wgrad<<<grid, kSemThreads, need, st>>>(p);
  WgReduceParams r;
  memset(&r, 0, sizeof(r));
  r.part = p.part; r.ncta = grid;
  const int nfeat = g.sem_coord ? 256 + g.enc : 256;
  r.seg[0][0] = WgSeg{0, nfeat, 128, p.gW0, (long long)p.ld0, 1};
  r.seg[0][1] = WgSeg{319, 320, 128, p.gb0, 1, 0};
  r.seg[0][2] = WgSeg{320, 320 + g.sem_dim, 128, p.gW2, 1, 128};
  r.seg[0][3] = WgSeg{324, 324 + g.sem_dim, 1, p.gb2, 0, 1};
  k_wgrad_reduce<<<dim3((kPartCols * kPartRows + 255) / 256, 1), 256, 0, st>>>(r);
  NSOS_CHECK_CUDA(cudaGetLastError());
  return NSOS_OK;
}

}  // namespace nsos

namespace nsos {
// dW[Mo, .] += dY[P, Mo]^T . [main (256 wide) | aux (aux_w <= 64 wide)]:  columns main_col.. / aux_col.. of dW (row pitch ldw).
bool tc_wgrad_gen_supported(int Mo, int64_t ldy, int main_w, int64_t ld_main, int aux_w) {
  return (Mo == 128 || Mo == 256) && ldy % 4 == 0 && (main_w == 0 || (main_w == 256 && ld_main % 4 == 0)) && aux_w >= 0 && aux_w <= 64 &&
         (main_w > 0 || aux_w > 0);
}
int tc_wgrad_gen(const float* dY, int64_t ldy, int Mo, const float* main, int64_t ld_main, int main_col, const float* aux, int64_t ld_aux,
                 int aux_w, int aux_col, float* dW, int64_t ldw, float* db, int64_t P, void* part, size_t part_bytes, cudaStream_t st) {
  NSOS_REQUIRE(tc_wgrad_gen_supported(Mo, ldy, main ? 256 : 0, ld_main, aux ? aux_w : 0), NSOS_ERR_UNSUPPORTED, "tc_wgrad_gen: unsupported shape");
  NSOS_REQUIRE(part && part_bytes >= tc_wgrad_part_bytes(), NSOS_ERR_WORKSPACE, "tc_wgrad_gen: partial-sum scratch too small");
  if (P <= 0) return NSOS_OK;
  WgGenParams p;
  memset(&p, 0, sizeof(p));
  p.dY = dY; p.ldy = ldy; p.Mo = Mo; p.main = main; p.ld_main = ld_main; p.main_col = main_col;
  p.aux = aux; p.ld_aux = ld_aux; p.aux_w = aux ? aux_w : 0; p.aux_col = aux_col; p.dW = dW; p.ldw = ldw; p.P = P;
  p.db = (p.aux_w <= 63) ? db : nullptr;
  NSOS_REQUIRE(!db || p.db, NSOS_ERR_UNSUPPORTED, "tc_wgrad_gen: the bias column needs aux_w <= 63");
  int dev = 0, sms = 0;
  NSOS_CHECK_CUDA(cudaGetDevice(&dev));
  NSOS_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long nslabs = (P + kSlabPts - 1) / kSlabPts;
  const int mb = Mo / 128;
  const int gx = (int)std::min<long long>(nslabs, std::max(1, std::min(sms, kPartMaxCtas) / mb));
  const size_t need = wgg_carve(nullptr, nullptr) + 1024;
  p.part = (float*)part;
  NSOS_REQUIRE(mb >= 1 && mb <= 2, NSOS_ERR_UNSUPPORTED, "tc_wgrad_gen: at most 256 output rows");
  // second version (rows staged with cp.async): needs 16-byte aligned rows of every operand and full 64-float aux rows
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  const bool v2 = !getenv("NSOS_WGRAD_V1") && ldy < 65536 && ld_main < 65536 && ld_aux < 65536 && al16(dY) && (!main || al16(main)) && (!p.aux || (al16(p.aux) && ld_aux % 4 == 0 && ld_aux >= 64));
  if (v2) {
    const size_t need2 = wgg2_carve(nullptr, nullptr) + 1024;
    auto go = [&](auto kern, int threads) -> int {
      NSOS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need2));
      kern<<<dim3(gx, mb), threads, need2, st>>>(p);
      return NSOS_OK;
    };
    const bool w8 = getenv("NSOS_WGRAD_W8") != nullptr, noloader = getenv("NSOS_WGRAD_NOLOADER") != nullptr;
    int rc = w8 ? (noloader ? go(k_wgrad_gen2<8, 0>, 32 * 9) : go(k_wgrad_gen2<8, 4>, 32 * 13))
                : (noloader ? go(k_wgrad_gen2<16, 0>, 32 * 17) : go(k_wgrad_gen2<16, 4>, 32 * 21));
    if (rc) return rc;
  } else {
    NSOS_CHECK_CUDA(cudaFuncSetAttribute(k_wgrad_gen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    k_wgrad_gen<<<dim3(gx, mb), kWgThreads, need, st>>>(p);
  }
  WgReduceParams r;
  memset(&r, 0, sizeof(r));
  r.part = p.part; r.ncta = gx;
  for (int y = 0; y < mb; ++y) {
    const int rows = std::min(128, Mo - 128 * y);
    float* dwy = dW + (size_t)128 * y * ldw;
    if (main) r.seg[y][0] = WgSeg{0, 256, rows, dwy + main_col, (long long)ldw, 1};
    if (p.aux_w > 0) r.seg[y][1] = WgSeg{256, 256 + p.aux_w, rows, dwy + aux_col, (long long)ldw, 1};
    if (p.db) r.seg[y][2] = WgSeg{319, 320, rows, p.db + 128 * y, 1, 0};
  }
  k_wgrad_reduce<<<dim3((kPartCols * kPartRows + 255) / 256, mb), 256, 0, st>>>(r);
  NSOS_CHECK_CUDA(cudaGetLastError());
  return NSOS_OK;
}
}  // namespace nsos
