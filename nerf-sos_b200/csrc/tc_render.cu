// Kernel A on the 5th-gen tensor cores: fused hierarchical render for sm_100a.
//
// One persistent CTA per SM walks ray PAIRS.  Per pair: coarse tile(s) -> composite + inverse-CDF
// resampling + sort (4 warps per ray, in shared memory) -> fine tiles -> composite -> one write of the
// per-ray maps.  A tile is 128 sample points = the 128 TMEM lanes; the whole MLP of a tile runs without
// leaving the SM:
//   * TENSOR MEMORY holds two 256-column buffers that alternate per stage: stage s accumulates D (fp32) into buffer s&1
//     while reading its A operand from the other one.  The epilogue of a hidden layer converts D IN PLACE into the next
//     layer's A operand: the 16 fp32 columns of K-step c become 8 columns of fp16 hi pairs + 8 columns of fp16 lo pairs
//     (same bytes), so a layer's output never needs a third region -- and because the next stage writes the OTHER buffer,
//     its MMAs start as soon as the first 64-column K slab has been converted (per-slab a_ready barriers): the epilogue
//     of layer l runs under the MMAs of layer l+1.  tcgen05.mma kind::f16, M=128, N<=256, K=16;
//   * weights stream from L2 through a ring of 32 KB shared-memory slots with cp.async.bulk (TMA engine,
//     mbarrier complete_tx), pre-swizzled at pack time into the canonical K-major SWIZZLE_128B slabs;
//   * NSOS_MODE_TC_EXACT splits activations and weights into fp16 hi+lo and issues A_hi.W_hi + A_lo.W_hi
//     + A_hi.W_lo (fp32-equivalent: ~22 mantissa bits, measured 1e-6 on composited maps);
//     NSOS_MODE_TC_FAST issues the hi.hi product only;
//   * the positional encoding gamma(x) is written by the row threads straight into a swizzled smem tile
//     and used as an SMEM A operand (layer 0, the skip layer, the semantic head);
//   * nine stages per tile for D=8: the trunk, then ONE head stage whose accumulator holds the semantic hidden layer
//     (cols 0..W/2) and the views hidden layer (cols W/2..W); feature_linear is folded into views_linears.0 at pack
//     time (see build_prog);
//   * the N<=3 heads (sigma, rgb, semantic logits) and the per-ray view-direction half of views_linears
//     are evaluated in fp32 on CUDA cores inside the epilogues, directly on the TMEM read-out.
// Warp roles: warps 0-7 = row workers (setup, epilogues, compositing; warp w owns TMEM lane quarter w%4 and column
// half w/4), warp 8 = MMA issuer (one elected lane), warps 9-10 = bulk-copy producers (one elected lane each).
// What bounds it (DESIGN.md sections 3 and 6): the MMA phases run at the tensor pipe's floor (129 cycles per
// M128 x N256 x K16), the epilogues at the TMEM read port's (64 B/clk: 2048 cycles per 256-column accumulator).
//
// Reference semantics: models/nerf_net.py:71-130 and the modules it calls (see include/nerfsos.h).
#include <algorithm>
#include <cstdlib>

#include <cuda_bf16.h>

#include "internal.h"
#include "render_device.cuh"
#include "render_group.cuh"
#include "tc_ptx.cuh"

namespace nsos {
namespace {
using namespace ptx;

constexpr int kSlotBytes = 32768;      // one weight slab plane: N(<=256) rows x 64 fp16
constexpr int kMaxSlots = 6, kDefaultSlots = 3;
constexpr int kWorkers = 256;          // 8 row-worker warps: warp w -> TMEM lane quarter w%4, column half w/4
constexpr int kMmaWarp = 8, kProducerWarp = 9;   // producer warps: kProducerWarp .. kProducerWarp + kNumProducers - 1
constexpr int kNumProducers = 2;                // one thread can issue only ~1 bulk copy per 680 cycles (tools/bulk_rate.cu)
#ifdef NSOS_AB_REGS
// 12 warps at 168 registers; the 8 worker warps grow to 216, the other 4 (MMA issuer, producers, one idle) shrink to 72:
// 8*32*216 + 4*32*72 = 12*32*168.
constexpr int kThreads = 384;
__device__ __forceinline__ void regs_worker() { asm volatile("setmaxnreg.inc.sync.aligned.u32 216;"); }
__device__ __forceinline__ void regs_other() { asm volatile("setmaxnreg.dec.sync.aligned.u32 72;"); }
#else
constexpr int kThreads = 32 * (kProducerWarp + kNumProducers);
__device__ __forceinline__ void regs_worker() {}
__device__ __forceinline__ void regs_other() {}
#endif
constexpr float kActScale = 16.f;      // activations are stored as fp16(16*a): keeps the lo plane out of fp16 subnormals
constexpr int kMaxStages = 13;
constexpr int kMaxSlabs = 5;
constexpr int kCtabMax = kMaxStages * kMaxSlabs * 2;
constexpr uint32_t kBufCols = 256, kTmemCols = 512;   // two D / A buffers (see the header comment); K-step c of a buffer: hi pairs
constexpr uint32_t kAStep = 16, kALo = 8;              // at columns [16c, 16c+8), lo pairs at [16c+8, 16c+16)
constexpr int kMaxASlabs = 5;                          // a_ready barriers: [0], [1] = K-steps 0-1 / 2-3 of slab 0 (early start of the
                                                       // next layer), [1+j] = slab j >= 1 (W/64 slabs of 64 columns)
constexpr int kGBytes = 128 * 128;     // gamma tile plane: 128 rows x 64 fp16, SWIZZLE_128B
constexpr int kHalfMax = 128;          // W/2 max
constexpr int kSemMax = 4;
constexpr int kHeadFloats = 256 + 4 + kSemMax * kHalfMax + 4 + 3 * kHalfMax + 4;

enum EpiKind { EPI_HIDDEN = 0, EPI_HIDDEN_SIGMA = 1, EPI_SEM = 2, EPI_SEM_WIDE = 3, EPI_RGB = 4, EPI_RAW = 5, EPI_SEM_RGB = 6 };
constexpr int A_GAMMA = -1;

// fp32 side data at the head of the packed image
struct TcAux {
  float inv_scale[16][2];       // per stage and row group: 1 / (weight scale * kActScale)
  uint32_t absmax_bits[16][2];  // per stage and row group max |w| (float bits), scratch of the pack pass
  float bias[16][256];          // per stage bias, by accumulator column (zero padded)
  float heads[kHeadFloats];     // w_alpha[256] b_alpha[4] w_s2[4][128] b_s2[4] w_rgb[3][128] b_rgb[4]
  float w_vdir[kHalfMax][28];   // views_linears.0.weight[:, W:W+encv]
};
constexpr size_t kAuxBytes = (sizeof(TcAux) + 1023) / 1024 * 1024;
constexpr int kHeadWAlpha = 0, kHeadBAlpha = 256, kHeadWS2 = 260, kHeadBS2 = 260 + kSemMax * kHalfMax,
              kHeadWRgb = kHeadBS2 + 4, kHeadBRgb = kHeadWRgb + 3 * kHalfMax;

struct TcStage {
  int n;                 // accumulator columns of the stage (multiple of 16)
  int nslab;             // K slabs of 64
  int epi;               // EpiKind
  int asrc[kMaxSlabs];   // A_GAMMA or TMEM K-slab index
  int sn[kMaxSlabs];     // MMA N of the slab (<= n; the gamma slab of the merged head stage only feeds the semantic half)
};
struct TcProg {
  int nst, W, Lp, Lv, enc, encv, sem_dim, H2;
  int ub_off;            // column of the last stage's bias row where the (fused) views bias starts
  TcStage st[kMaxStages];
};

// host-side plan used by the pack kernels.  A stage's accumulator columns come from up to two ROW GROUPS (weight
// matrices stacked along N): group 0 = columns [0, rows[0]), group 1 = the next rows[1].  src 0 = the flat parameter
// buffer, src 1 = the fused views matrix W_v[:, :W] . W_feat computed into the image by k_fuse_views.
struct PackGroup { int src; int64_t w_off; int ld, rows; int bsrc; int64_t b_off; };
struct PackSlab { int stage, n; int col0[2], kvalid[2]; int64_t dst_hi, dst_lo; };
struct PackPlan {
  int nslab, nst;
  PackSlab s[kMaxStages * kMaxSlabs];
  PackGroup grp[kMaxStages][2];
  int64_t fused_off;     // byte offset of the fused fp32 block in the image: W_uf[H2][W], then b_uf[H2]
  int64_t total_bytes;
};

bool tc_supported(const NetGeom& g, const char** why) {
  static const char* w;
  auto fail = [&](const char* m) { w = m; if (why) *why = w; return false; };
  if (!g.use_viewdirs) return fail("use_viewdirs=0 is only implemented by NSOS_MODE_SIMT_FP32");
  if (g.W % 64 != 0 || g.W < 64 || g.W > 256) return fail("tcgen05 path needs W in {64,128,192,256}");
  if (g.D > 10) return fail("tcgen05 path needs D <= 10");
  if (g.enc > 63 || g.encv > 27) return fail("tcgen05 path needs multires<=10, multires_views<=4");
  if (g.sem_dim > kSemMax) return fail("tcgen05 path needs sem_dim <= 4");
  return true;
}

// Stage program.  Two algebraic fusions keep work off the tensor pipe (both exact up to fp32 rounding order):
//  * feature_linear has no non-linearity (nerf_mlp.py:86-89), so views_linears.0([feature_linear(h), enc_dirs]) =
//    (W_v[:, :W] W_f) h + (W_v[:, :W] b_f + b_v) + W_v[:, W:] enc_dirs: the W x W feature stage disappears, the product
//    matrix is formed once at pack time (k_fuse_views, fp64 accumulation);
//  * semantic_linear.0 and the fused views layer both read h_last, so they share ONE stage: accumulator columns
//    [0, W/2) = semantic hidden layer, [W/2, W) = views hidden layer; the gamma slab (sem_with_coord) is issued with
//    N = W/2 so that it only feeds the semantic half.
void build_prog(const NetGeom& g, bool exact, TcProg& p, PackPlan& plan) {
  memset(&p, 0, sizeof(p)); memset(&plan, 0, sizeof(plan));
  p.W = g.W; p.Lp = g.Lp; p.Lv = g.Lv; p.enc = g.enc; p.encv = g.encv; p.sem_dim = g.sem_dim; p.H2 = g.W / 2;
  const int ks = g.W / 64, H2 = g.W / 2;
  int64_t off = (int64_t)kAuxBytes;
  auto add_slab = [&](int asrc, int n, int col0_0, int kv0, int col0_1, int kv1) {
    TcStage& s = p.st[p.nst];
    s.sn[s.nslab] = n;
    s.asrc[s.nslab++] = asrc;
    PackSlab& ps = plan.s[plan.nslab++];
    ps.stage = p.nst; ps.n = n; ps.col0[0] = col0_0; ps.kvalid[0] = kv0; ps.col0[1] = col0_1; ps.kvalid[1] = kv1;
    ps.dst_hi = off; off += (int64_t)n * 128;
    ps.dst_lo = -1;
    if (exact) { ps.dst_lo = off; off += (int64_t)n * 128; }
  };
  auto group = [&](int gi, int src, int64_t w_off, int ld, int rows, int bsrc, int64_t b_off) {
    PackGroup& pgp = plan.grp[p.nst][gi];
    pgp.src = src; pgp.w_off = w_off; pgp.ld = ld; pgp.rows = rows; pgp.bsrc = bsrc; pgp.b_off = b_off;
  };
  for (int i = 0; i < g.D; ++i) {
    TcStage& s = p.st[p.nst];
    s.n = g.W; s.epi = (i == g.D - 1) ? EPI_HIDDEN_SIGMA : EPI_HIDDEN;
    group(0, 0, g.w_pts[i], g.in_pts[i], g.W, 0, g.b_pts[i]);
    if (i == 0) add_slab(A_GAMMA, g.W, 0, g.enc, 0, 0);
    else if (g.in_pts[i] == g.W) for (int j = 0; j < ks; ++j) add_slab(j, g.W, 64 * j, 64, 0, 0);
    else {                                                           // [enc, h] (nerf_mlp.py:73-74)
      add_slab(A_GAMMA, g.W, 0, g.enc, 0, 0);
      for (int j = 0; j < ks; ++j) add_slab(j, g.W, g.enc + 64 * j, 64, 0, 0);
    }
    ++p.nst;
  }
  {
    TcStage& s = p.st[p.nst];
    if (g.use_sem) {                                                 // semantic_linear.0 on [h, enc] | fused views on h
      s.n = 2 * H2; s.epi = EPI_SEM_RGB; p.ub_off = H2;
      group(0, 0, g.w_s0, g.sem_in, H2, 0, g.b_s0);
      group(1, 1, 0, g.W, H2, 1, 0);
      for (int j = 0; j < ks; ++j) add_slab(j, 2 * H2, 64 * j, 64, 64 * j, 64);
      if (g.sem_coord) add_slab(A_GAMMA, H2, g.W, g.enc, 0, 0);
    } else {
      s.n = H2; s.epi = EPI_RGB; p.ub_off = 0;
      group(0, 1, 0, g.W, H2, 1, 0);
      for (int j = 0; j < ks; ++j) add_slab(j, H2, 64 * j, 64, 0, 0);
    }
    ++p.nst;
  }
  plan.nst = p.nst;
  plan.fused_off = off;
  off += (int64_t)sizeof(float) * ((int64_t)H2 * g.W + H2);
  plan.total_bytes = (off + 255) / 256 * 256;
}

// ---- pack kernels -----------------------------------------------------------------------------------
// W_uf[j][k] = sum_m W_v[j][m] W_f[m][k],  b_uf[j] = b_v[j] + sum_m W_v[j][m] b_f[m]   (fp64 accumulation, fp32 result)
__global__ void k_fuse_views(const float* __restrict__ prm, float* __restrict__ fused, const NetGeom g) {
  const int j = blockIdx.x, H2 = g.W / 2;
  const float* wv = prm + g.w_views + (int64_t)j * (g.W + g.encv);
  for (int k = threadIdx.x; k <= g.W; k += blockDim.x) {
    double acc = 0.0;
    if (k < g.W) {
      for (int m = 0; m < g.W; ++m) acc += (double)wv[m] * (double)prm[g.w_feat + (int64_t)m * g.W + k];
      fused[(int64_t)j * g.W + k] = (float)acc;
    } else {
      for (int m = 0; m < g.W; ++m) acc += (double)wv[m] * (double)prm[g.b_feat + m];
      fused[(int64_t)H2 * g.W + j] = (float)(acc + (double)prm[g.b_views + j]);
    }
  }
}
__device__ __forceinline__ const float* group_src(const PackGroup& pg, const float* prm, const float* fused) {
  return (pg.src ? fused : prm) + pg.w_off;
}
__global__ void k_pack_absmax(const float* __restrict__ prm, uint8_t* __restrict__ img, const PackPlan plan) {
  const int st = blockIdx.y, gi = blockIdx.z;
  const PackGroup& pg = plan.grp[st][gi];
  TcAux* aux = reinterpret_cast<TcAux*>(img);
  const int64_t n = (int64_t)pg.rows * pg.ld;
  const float* w = group_src(pg, prm, reinterpret_cast<const float*>(img + plan.fused_off));
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(w[i]));
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && n > 0) atomicMax(&aux->absmax_bits[st][gi], __float_as_uint(m));
}
__device__ __forceinline__ float stage_scale(const TcAux* aux, int st, int gi) {
  float amax = __uint_as_float(aux->absmax_bits[st][gi]);
  if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
  int e;
  frexpf(amax, &e);                      // amax = m * 2^e, m in [0.5, 1)
  return ldexpf(1.f, 15 - e);            // amax*scale in [2^14, 2^15)
}
__global__ void k_pack_aux(const float* __restrict__ prm, uint8_t* __restrict__ img, const PackPlan plan, const NetGeom g) {
  TcAux* aux = reinterpret_cast<TcAux*>(img);
  const float* fused = reinterpret_cast<const float*>(img + plan.fused_off);
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int nt = gridDim.x * blockDim.x;
  for (int i = t; i < 32; i += nt) {
    const int st = i >> 1, gi = i & 1;
    float v = 0.f;
    if (st < plan.nst) v = 1.f / stage_scale(aux, st, plan.grp[st][gi].rows > 0 ? gi : 0);   // acc -> 16*x
    aux->inv_scale[st][gi] = v;
  }
  const int H = g.W / 2;
  for (int i = t; i < 16 * 256; i += nt) {
    int st = i / 256, c = i % 256;
    float v = 0.f;
    if (st < plan.nst) {
      const int r0 = plan.grp[st][0].rows;
      const int gi = (c < r0) ? 0 : 1, cc = c - (gi ? r0 : 0);
      const PackGroup& pg = plan.grp[st][gi];
      if (cc < pg.rows) v = (pg.bsrc ? fused[(int64_t)H * g.W + cc] : prm[pg.b_off + cc]) * kActScale;   // 16*b
    }
    aux->bias[st][c] = v;
  }
  for (int i = t; i < kHeadFloats; i += nt) {
    float v = 0.f;
    const float hs = 1.f / kActScale;          // head weights act on 16*x
    if (i < kHeadBAlpha) { if (i < g.W) v = prm[g.w_alpha + i] * hs; }
    else if (i < kHeadWS2) { if (i == kHeadBAlpha) v = prm[g.b_alpha]; }
    else if (i < kHeadBS2) { int s = (i - kHeadWS2) / kHalfMax, c = (i - kHeadWS2) % kHalfMax; if (s < g.sem_dim && c < H) v = prm[g.w_s2 + (int64_t)s * H + c] * hs; }
    else if (i < kHeadWRgb) { int s = i - kHeadBS2; if (s < g.sem_dim) v = prm[g.b_s2 + s]; }
    else if (i < kHeadBRgb) { int k = (i - kHeadWRgb) / kHalfMax, c = (i - kHeadWRgb) % kHalfMax; if (c < H) v = prm[g.w_rgb + (int64_t)k * H + c] * hs; }
    else { int k = i - kHeadBRgb; if (k < 3) v = prm[g.b_rgb + k]; }
    aux->heads[i] = v;
  }
  for (int i = t; i < kHalfMax * 28; i += nt) {
    int r = i / 28, c = i % 28;
    aux->w_vdir[r][c] = (r < H && c < g.encv) ? prm[g.w_views + (int64_t)r * (g.W + g.encv) + g.W + c] : 0.f;
  }
}
__global__ void k_fix_scale(TcAux* aux) { aux->inv_scale[0][0] = aux->inv_scale[0][1] = 1.f / stage_scale(aux, 0, 0); }
// one block per (slab, 64-row group): writes the hi (and lo) SWIZZLE_128B K-major image of the slab
__global__ void k_pack_slabs(const float* __restrict__ prm, uint8_t* __restrict__ img, const PackPlan plan) {
  const PackSlab ps = plan.s[blockIdx.x];
  const TcAux* aux = reinterpret_cast<const TcAux*>(img);
  const float* fused = reinterpret_cast<const float*>(img + plan.fused_off);
  const PackGroup g0 = plan.grp[ps.stage][0], g1 = plan.grp[ps.stage][1];
  const float* src[2] = {group_src(g0, prm, fused), group_src(g1, prm, fused)};
  const float scale[2] = {stage_scale(aux, ps.stage, 0), g1.rows > 0 ? stage_scale(aux, ps.stage, 1) : 1.f};
  for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < ps.n * 64; e += gridDim.y * blockDim.x) {
    int n = e >> 6, k = e & 63;
    const int gi = (n < g0.rows) ? 0 : 1, r = n - (gi ? g0.rows : 0);
    const PackGroup& pg = gi ? g1 : g0;
    float w = 0.f;
    if (r < pg.rows && k < ps.kvalid[gi]) w = src[gi][(int64_t)r * pg.ld + ps.col0[gi] + k] * scale[gi];
    __half h = __float2half_rn(w);
    size_t byte = (size_t)(n >> 3) * 1024 + (size_t)(n & 7) * 128 + (size_t)(((k >> 3) ^ (n & 7)) << 4) + (size_t)(k & 7) * 2;
    *reinterpret_cast<__half*>(img + ps.dst_hi + byte) = h;
    if (ps.dst_lo >= 0) *reinterpret_cast<__half*>(img + ps.dst_lo + byte) = __float2half_rn(w - __half2float(h));
  }
}

// ---- render kernel ------------------------------------------------------------------------------------
struct TcParams {
  TcProg prog[2];
  uint32_t ctab[2][kCtabMax];   // per-net chunk program of the MMA issuer (build_ctab)
  int nch[2];
  const uint8_t* packed[2];
  const float *rays_o, *rays_d, *near, *far;
  NsosRandoms rnd;
  unsigned long long seed;
  NsosRenderOut out;
  long long n_rays;
  int Sc, K, Sf;
  float perturb, noise_std;
  int white_bkgd, exact, fine, C, sem_dim, C6, ML, nslots;
  long long* trace;   // optional timeline buffer (NSOS_TRACE=1): clock64 stamps of CTA 0, see tools/trace_report.py
  // replay (backward recompute on given samples): sample depths in, last trunk activation / semantic hidden layer out
  const float* z_in[2];
  float* dump_h[2];    // [n_rays*S, W]   relu(pts_linears[D-1])
  float* dump_s0[2];   // [n_rays*S, W/2] relu(semantic_linear.0)
  // MODE 3 (replay for the all-parameter backward): every hidden layer and the views hidden layer as well
  float* dump_all[2][kMaxStages];   // [layer][n_rays*S, W]  relu(pts_linears[layer]); entries may be null
  float* dump_hv[2];                // [n_rays*S, W/2] relu(views_linears.0)
  float* dump_enc[2];               // [n_rays*S, 64]  gamma(x) (training forward: the semantic-head weight gradients read it)
  // Point query (nsos_mlp_query_dir, MODE 2 only): row q of a tile is point `ray*S + i` of pts_in [n_pts,3], evaluated with ONE
  // view direction qdir for all points (used as given, not normalised); no rays, no depths.
  const float* pts_in;
  long long n_pts;
  float qdir[3];
  // Layout of dump_h / dump_s0 / dump_enc per pass.  0: row-major [point][feature].  1: blocked -- groups of 32 consecutive points,
  // feature-major inside a group: element (pt, f) of an F-wide tensor sits at ((pt >> 5) * F + f) * 32 + (pt & 31).  The 32 lanes of
  // an epilogue warp hold 32 consecutive points, so one store instruction covers one 128-byte line (row-major: 32 lines), and
  // kernel C (tc_wgrad.cu), which contracts over points, reads 32 points of a feature with one coalesced load.
  int dump_blocked[2];
};
constexpr int kTraceTiles = 16, kTraceStamps = 12;  // [tile][stage][stamp]; 5..8: a_ready[j] seen by the MMA lane, 9..11: worker 0 hands over slab 0..2

struct Smem {
  uint8_t* ring; uint8_t* g_hi; uint8_t* g_lo;
  float *rawbuf, *hpart, *zc, *w0, *cdf, *bins, *zall, *zf, *dirbias, *sbias, *heads, *rayp, *encv;
  uint64_t *full, *empty, *acc_full, *g_ready, *a_ready;   // a_ready[kMaxASlabs]
  uint32_t* tmem_ptr;
  uint32_t* ctab;      // [2][kCtabMax] chunk program of the MMA issuer (copy of TcParams::ctab)
};
constexpr int kRayP = 16;  // per ray: o[3] d[3] near far |d| valid v[3] = d/|d| (+3 pad)

__host__ __device__ inline size_t carve_smem(uint8_t* base, int nslots, int Sc, int Sf, int C, Smem* s) {
  size_t off = 0;
  auto take = [&](size_t bytes, size_t align) { off = (off + align - 1) / align * align; size_t o = off; off += bytes; return o; };
  size_t o_ring = take((size_t)nslots * kSlotBytes, 1024);
  size_t o_ghi = take(kGBytes, 1024), o_glo = take(kGBytes, 1024);
  size_t o_raw = take(sizeof(float) * 2 * Sf * C, 16);
  size_t o_hp = take(sizeof(float) * 128 * 8, 16);
  size_t o_zc = take(sizeof(float) * 2 * Sc, 16), o_w0 = take(sizeof(float) * 2 * Sc, 16), o_cdf = take(sizeof(float) * 2 * Sc, 16),
         o_bins = take(sizeof(float) * 2 * Sc, 16);
  size_t o_zall = take(sizeof(float) * 2 * Sf, 16), o_zf = take(sizeof(float) * 2 * Sf, 16);
  size_t o_db = take(sizeof(float) * 2 * 2 * kHalfMax, 16), o_sb = take(sizeof(float) * 2 * 256, 16);
  size_t o_heads = take(sizeof(float) * 2 * kHeadFloats, 16), o_rayp = take(sizeof(float) * 2 * kRayP, 16),
         o_encv = take(sizeof(float) * 2 * 28, 16);
  size_t o_bar = take(sizeof(uint64_t) * (2 * kMaxSlots + 2 + kMaxASlabs), 8), o_tp = take(16, 16);
  size_t o_ct = take(sizeof(uint32_t) * 2 * kCtabMax, 16);
  if (s) {
    s->ring = base + o_ring; s->g_hi = base + o_ghi; s->g_lo = base + o_glo;
    s->rawbuf = (float*)(base + o_raw); s->hpart = (float*)(base + o_hp); s->zc = (float*)(base + o_zc); s->w0 = (float*)(base + o_w0);
    s->cdf = (float*)(base + o_cdf); s->bins = (float*)(base + o_bins); s->zall = (float*)(base + o_zall);
    s->zf = (float*)(base + o_zf); s->dirbias = (float*)(base + o_db); s->sbias = (float*)(base + o_sb);
    s->heads = (float*)(base + o_heads); s->rayp = (float*)(base + o_rayp); s->encv = (float*)(base + o_encv);
    s->full = (uint64_t*)(base + o_bar); s->empty = s->full + kMaxSlots; s->acc_full = s->empty + kMaxSlots;
    s->g_ready = s->acc_full + 1; s->a_ready = s->g_ready + 1; s->tmem_ptr = (uint32_t*)(base + o_tp);
    s->ctab = (uint32_t*)(base + o_ct);
  }
  return off;
}

// ---- producer: stream one tile's worth of weight planes through the ring -------------------------------
// Called by all 32 lanes of the producer warp; the elected lane issues the bulk copies.  In a cluster of
// `csize` CTAs every CTA fetches 1/csize of each plane and multicasts it into the same ring slot of all CTAs,
// so each weight byte crosses L2 -> SM once per cluster instead of once per CTA.
__device__ __forceinline__ void producer_tile(const TcProg& pg, const uint8_t* img, bool exact, const Smem& sm, int nslots,
                                              uint32_t& chunk, uint32_t crank, uint32_t csize, uint32_t pidx) {
  const uint8_t* src = img + kAuxBytes;
  const uint16_t mask = (uint16_t)((1u << csize) - 1u);
  const int nplanes = exact ? 2 : 1;
  for (int st = 0; st < pg.nst; ++st) {
    const int planes = pg.st[st].nslab * nplanes;
    for (int c = 0; c < planes; ++c) {
      const uint32_t bytes = (uint32_t)pg.st[st].sn[c / nplanes] * 128u;
      const uint32_t share = bytes / csize;
      if (chunk % kNumProducers != pidx) { src += bytes; ++chunk; continue; }   // chunks alternate between the producer warps
      const uint32_t slot = chunk % nslots, par = (chunk / nslots) & 1u;
      mbar_wait(smem_u32(&sm.empty[slot]), par ^ 1u, 100 + (int)slot);     // released by every CTA of the cluster
      if (elect_one()) {
        mbar_arrive_expect_tx(smem_u32(&sm.full[slot]), bytes);
        const uint32_t dst = smem_u32(sm.ring + (size_t)slot * kSlotBytes) + crank * share;
        if (csize == 1) bulk_g2s(dst, src, bytes, smem_u32(&sm.full[slot]));
        else bulk_g2s_multicast(dst, src + (size_t)crank * share, share, smem_u32(&sm.full[slot]), mask);
      }
      __syncwarp();
      src += bytes;
      ++chunk;
    }
  }
}

// ---- MMA issuer: all stages of one tile ------------------------------------------------------------------
// Called by all 32 lanes of the MMA warp (uniform control flow); the elected lane issues every tcgen05.mma and
// tcgen05.commit so that the commits track that lane's MMAs.
//
// Issue-side latency matters: the warp leaves an elect block only when the block's MMAs have been dispatched to
// the tensor pipe (the reconvergence point waits on their scoreboards), so whatever runs between two blocks is a
// pipe bubble unless it is shorter than one MMA (~130 cycles).  Therefore (measured with tools/umma_queue.cu and
// the NSOS_TRACE timeline): the ring position is tracked incrementally (no division), and the mbarrier wait for
// the NEXT weight chunk is issued by the elected lane inside the current block, before its last MMA, where it
// overlaps with the MMAs already queued (the queue holds ~10).
struct RingPos {
  uint32_t slot, par;
  __device__ __forceinline__ void advance(uint32_t nslots) { if (++slot == nslots) { slot = 0; par ^= 1u; } }
};

// The issuer's program for one tile, flattened to one 32-bit word per weight chunk (= one K slab plane in the ring), built on
// the host (kernel parameters) and copied to shared memory: the elected lane runs ONE small loop for the whole kernel.  Why:
// whatever the issuer executes between the last MMA of layer l and the first MMA of layer l+1 sits on the critical path of
// the layer pipeline, and an instruction of this warp costs ~10 cycles while the eight epilogue warps are busy (NSOS_TRACE
// stamps: the old per-stage prologue -- stage descriptor reads, elect.sync, reconvergence -- took 2.5k cycles, the whole
// overlap window; tools/pipe_overlap.cu: ten extra instructions per MMA double the time per MMA).  Measured alternative
// (profiles/r02_notes.md): all 32 lanes in uniform control flow with an elect region per chunk gives textbook SASS (UTCHMMA
// operands stay in uniform registers, no R2UR) but was 9 % slower end to end -- 32 lanes polling the barriers.
//   bits 0-2 A source (0..3 = TMEM K slab, 7 = gamma tile in shared memory)   bits 3-8 N/8   bit 9 W_lo plane
//   bit 10 first chunk of a stage   bit 11 last chunk of a stage   bit 12 also issue the A_lo pass (exact mode, W_hi plane)
//   bits 13-15 / 16-18: a_ready barrier (+1; 0 = none) to wait for before K-step 0 / before K-step 2 of this chunk
constexpr uint32_t kCtGamma = 7u, kCtLoPlane = 1u << 9, kCtFirst = 1u << 10, kCtLast = 1u << 11, kCtTwo = 1u << 12;
constexpr int kCtWait0 = 13, kCtWait2 = 16;
__host__ __device__ inline int build_ctab(const TcProg& pg, bool exact, uint32_t* tab) {
  int n = 0;
  const int nplanes = exact ? 2 : 1;
  for (int st = 0; st < pg.nst; ++st) {
    const TcStage& S = pg.st[st];
    for (int j = 0; j < S.nslab; ++j)
      for (int plane = 0; plane < nplanes; ++plane) {
        uint32_t w = (S.asrc[j] == A_GAMMA ? kCtGamma : (uint32_t)S.asrc[j]) | ((uint32_t)(S.sn[j] >> 3) << 3);
        if (plane) w |= kCtLoPlane;
        if (j == 0 && plane == 0) w |= kCtFirst;
        if (j == S.nslab - 1 && plane == nplanes - 1) w |= kCtLast;
        if (exact && plane == 0) w |= kCtTwo;
        if (S.asrc[j] != A_GAMMA && plane == 0) {                  // first use of this TMEM slab in the stage
          if (S.asrc[j] == 0) w |= (1u << kCtWait0) | (2u << kCtWait2);
          else w |= (uint32_t)(S.asrc[j] + 2) << kCtWait0;
        }
        tab[n++] = w;
      }
  }
  return n;
}

// Issuer-side pipeline state that lives across tiles (registers of the elected lane): ring position, global stage counter
// (selects the D buffer), phase parities of the gamma-tile barrier and of the per-slab A barriers (bit j = a_ready[j]).
struct MmaState {
  RingPos pos;
  uint32_t gs, gpar, apar;
  uint32_t fmt = 0;      // OR-ed into the instruction descriptor: 0 = fp16 operands, kIdescBf16 = bf16 operands (row GEMM)
};
constexpr uint32_t kIdescBf16 = (1u << 7) | (1u << 10);

// All stages of one tile.  Called by the ELECTED LANE ONLY (the caller holds the elect block around the whole kernel loop).
// Issue-side latency matters (tools/umma_queue.cu): the ring position is tracked incrementally, and the mbarrier wait for the
// NEXT weight chunk is issued before the last MMA of the current one, where it overlaps with the MMAs already queued.
// A TMEM K slab is consumed as soon as the previous layer's epilogue has converted it in place (a_ready[slab]) while the later
// slabs of that buffer are still fp32 accumulator columns; the stage accumulates into the OTHER buffer, so epilogue(l) and
// MMA(l+1) overlap.
__device__ __forceinline__ void mma_tile(const uint32_t* __restrict__ tab, int nch, const Smem& sm, uint32_t nslots, uint32_t tm,
                                         MmaState& ms, uint32_t csize, bool last_tile, long long* trace = nullptr) {
  const uint32_t g_hi = smem_u32(sm.g_hi), g_lo = smem_u32(sm.g_lo);
  const uint32_t ring = smem_u32(sm.ring), full0 = smem_u32(&sm.full[0]), empty0 = smem_u32(&sm.empty[0]);
  const uint32_t a_ready0 = smem_u32(&sm.a_ready[0]), acc_full = smem_u32(sm.acc_full);
  const uint16_t mask = (uint16_t)((1u << csize) - 1u);
  // the gamma tile of this 128-point tile has been written (first stage, skip stage and semantic head read it)
  mbar_wait(smem_u32(sm.g_ready), ms.gpar, 200);
  ms.gpar ^= 1u;
  tc_fence_after();
  uint32_t accum = 0, dcol = 0, acol = 0;
  int st = 0;
#pragma unroll 1
  for (int c = 0; c < nch; ++c) {
    const uint32_t w = tab[c];
    const uint32_t asrc = w & 7u;
    if (w & kCtFirst) {
      dcol = tm + (ms.gs & 1u) * kBufCols;
      acol = tm + ((ms.gs & 1u) ^ 1u) * kBufCols;
      accum = 0;
      if (trace) trace[st * kTraceStamps + 3] = clock64();      // stage issue starts
    }
    const uint32_t wait0 = (w >> kCtWait0) & 7u, wait2 = (w >> kCtWait2) & 7u;
    if (wait0) {
      mbar_wait(a_ready0 + (wait0 - 1u) * 8u, (ms.apar >> (wait0 - 1u)) & 1u, 210 + (int)wait0);
      ms.apar ^= 1u << (wait0 - 1u);
      tc_fence_after();
      if (trace) trace[st * kTraceStamps + 5 + asrc] = clock64();
    }
    const uint32_t idesc = make_idesc_f16((int)((w >> 3) & 63u) << 3) | ms.fmt;
    const uint32_t a_hi = acol + asrc * 64;
    // full[pos.slot] of THIS chunk was already waited for (inside the previous chunk, or before the first tile)
    const uint32_t b = ring + ms.pos.slot * kSlotBytes;
    const uint32_t my_empty = empty0 + ms.pos.slot * 8u;
    ms.pos.advance(nslots);
    const bool has_next = !(last_tile && c == nch - 1);
    const bool two = (w & kCtTwo) != 0;        // W_hi plane in exact mode: A_hi.W_hi + A_lo.W_hi;  W_lo plane: A_hi.W_lo
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint64_t bd = make_sw128_desc(b + ks * 32);
      const uint32_t acc = (ks == 0) ? accum : 1u;
      if (ks == 3 && !two && has_next) { mbar_wait(full0 + ms.pos.slot * 8u, ms.pos.par, 300 + (int)ms.pos.slot); tc_fence_after(); }
      if (ks == 2 && wait2) {                   // second half of slab 0
        mbar_wait(a_ready0 + (wait2 - 1u) * 8u, (ms.apar >> (wait2 - 1u)) & 1u, 220);
        ms.apar ^= 1u << (wait2 - 1u);
        tc_fence_after();
      }
      if (asrc == kCtGamma) umma_ss(dcol, make_sw128_desc(g_hi + ks * 32), bd, idesc, acc);
      else umma_ts(dcol, a_hi + ks * kAStep, bd, idesc, acc);
      if (two) {
        if (ks == 3 && has_next) { mbar_wait(full0 + ms.pos.slot * 8u, ms.pos.par, 300 + (int)ms.pos.slot); tc_fence_after(); }
        if (asrc == kCtGamma) umma_ss(dcol, make_sw128_desc(g_lo + ks * 32), bd, idesc, 1);
        else umma_ts(dcol, a_hi + ks * kAStep + kALo, bd, idesc, 1);
      }
    }
    if (csize == 1) umma_commit(my_empty);
    else umma_commit_multicast(my_empty, mask);
    accum = 1;
    if (w & kCtLast) {
      umma_commit(acc_full);
      ++ms.gs;
      if (trace) trace[st * kTraceStamps + 4] = clock64();      // all MMAs of the stage issued
      ++st;
    }
  }
}

// ---- worker helpers -----------------------------------------------------------------------------------------
// Four 16-byte chunks (32 fp16) of one row of a K-major SWIZZLE_128B tile: columns [32*half, 32*half+32).
// v holds the 32 fp32 values (already x kActScale).
__device__ __forceinline__ void store_halfrow_sw128(uint8_t* hi_tile, uint8_t* lo_tile, int row, int half, const float* v, bool exact) {
  const size_t rbase = (size_t)(row >> 3) * 1024 + (size_t)(row & 7) * 128;
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    const int j = half * 4 + jj;
    uint32_t h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float a0 = v[8 * jj + 2 * q], a1 = v[8 * jj + 2 * q + 1];
      __half2 hh = __floats2half2_rn(a0, a1);
      float2 hf = __half22float2(hh);
      __half2 ll = __floats2half2_rn(a0 - hf.x, a1 - hf.y);
      h[q] = *reinterpret_cast<uint32_t*>(&hh);
      l[q] = *reinterpret_cast<uint32_t*>(&ll);
    }
    const size_t off = rbase + (size_t)((j ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(hi_tile + off) = make_uint4(h[0], h[1], h[2], h[3]);
    if (exact) *reinterpret_cast<uint4*>(lo_tile + off) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// gamma(x) columns [32*HALF, 32*HALF+32) (embedder.py:34-48), scaled by kActScale, zero beyond `enc`
template <int HALF>
__device__ __forceinline__ void encode_half(const float x[3], int L, int enc, bool valid, float* e) {
#pragma unroll
  for (int c = 0; c < 32; ++c) e[c] = 0.f;
  if (HALF == 0) { e[0] = x[0]; e[1] = x[1]; e[2] = x[2]; }
  constexpr int lo = 32 * HALF, hi = lo + 32;
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    // frequency k occupies columns [3+6k, 9+6k)
    if (3 + 6 * k < hi && 9 + 6 * k > lo && k < L) {
      const float f = (float)(1 << k);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        float s, c;
        sincos_cw(__fmul_rn(x[a], f), &s, &c);
        const int cs = 3 + 6 * k + a, cc = cs + 3;
        if (cs >= lo && cs < hi) e[cs - lo] = s;
        if (cc >= lo && cc < hi) e[cc - lo] = c;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 32; ++c) e[c] = (valid && lo + c < enc) ? e[c] * kActScale : 0.f;
}

// One 16-column chunk of an epilogue.  v = D read-out (already waited for).  x16 = 16*x = acc*inv16 + bias16
// (+relu) feeds the fp32 head accumulators (head weights are pre-divided by 16) and/or becomes the next A operand.
constexpr int kCW = 16;   // epilogue chunk width in columns (one tcgen05.ld) = one K-step of the next layer
constexpr int kSW = 16;   // arithmetic sub-block of a chunk
template <int KIND, bool EXACT, bool DUMP, bool DALL = false, bool BLK = false>
__device__ __forceinline__ void epi_sub(const uint32_t* v, uint32_t tm_lane, int c0, float inv16, const float* __restrict__ bias,
                                        const float* __restrict__ hw, int sem_dim, float (&hacc)[4], float (&hodd)[4],
                                        float* __restrict__ gout, float& vmax) {
  uint32_t hi[kSW / 2], lo[kSW / 2];
  // bias and head weights of this sub-block as 16-byte shared loads (c0 is a multiple of 16 floats; all bases 16 B aligned)
  float bz[kSW], h0[kSW], h1[kSW], h2[kSW];
  {
    const float4* b4 = reinterpret_cast<const float4*>(bias + c0);
#pragma unroll
    for (int q = 0; q < kSW / 4; ++q) { float4 t4 = b4[q]; bz[4 * q] = t4.x; bz[4 * q + 1] = t4.y; bz[4 * q + 2] = t4.z; bz[4 * q + 3] = t4.w; }
    constexpr bool SEM = (KIND == EPI_SEM || KIND == EPI_SEM_WIDE);
    if (KIND == EPI_HIDDEN_SIGMA || SEM || KIND == EPI_RGB) {
      const int o0 = (KIND == EPI_HIDDEN_SIGMA) ? kHeadWAlpha : SEM ? kHeadWS2 : kHeadWRgb;
      const float4* w0 = reinterpret_cast<const float4*>(hw + o0 + c0);
      const float4* w1 = reinterpret_cast<const float4*>(hw + o0 + kHalfMax + c0);
      const float4* w2 = reinterpret_cast<const float4*>(hw + o0 + 2 * kHalfMax + c0);
#pragma unroll
      for (int q = 0; q < kSW / 4; ++q) {
        float4 t4 = w0[q]; h0[4 * q] = t4.x; h0[4 * q + 1] = t4.y; h0[4 * q + 2] = t4.z; h0[4 * q + 3] = t4.w;
        if (KIND != EPI_HIDDEN_SIGMA) { t4 = w1[q]; h1[4 * q] = t4.x; h1[4 * q + 1] = t4.y; h1[4 * q + 2] = t4.z; h1[4 * q + 3] = t4.w; }
        if (KIND == EPI_RGB) { t4 = w2[q]; h2[4 * q] = t4.x; h2[4 * q + 1] = t4.y; h2[4 * q + 2] = t4.z; h2[4 * q + 3] = t4.w; }
      }
    }
  }
  float dq0 = 0.f, dq1 = 0.f;
#pragma unroll
  for (int j = 0; j < kSW; j += 2) {
    float x0 = fmaf(__uint_as_float(v[j]), inv16, bz[j]);
    float x1 = fmaf(__uint_as_float(v[j + 1]), inv16, bz[j + 1]);
    if (KIND != EPI_RAW) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
    // head partial sums: two independent FMA chains (even / odd columns) per output, all in registers
    if (KIND == EPI_HIDDEN_SIGMA) {
      hacc[0] = fmaf(h0[j], x0, hacc[0]);
      hodd[0] = fmaf(h0[j + 1], x1, hodd[0]);
    } else if (KIND == EPI_SEM || KIND == EPI_SEM_WIDE) {
      hacc[0] = fmaf(h0[j], x0, hacc[0]); hodd[0] = fmaf(h0[j + 1], x1, hodd[0]);
      if (sem_dim > 1) { hacc[1] = fmaf(h1[j], x0, hacc[1]); hodd[1] = fmaf(h1[j + 1], x1, hodd[1]); }
    } else if (KIND == EPI_RGB) {
      hacc[0] = fmaf(h0[j], x0, hacc[0]); hodd[0] = fmaf(h0[j + 1], x1, hodd[0]);
      hacc[1] = fmaf(h1[j], x0, hacc[1]); hodd[1] = fmaf(h1[j + 1], x1, hodd[1]);
      hacc[2] = fmaf(h2[j], x0, hacc[2]); hodd[2] = fmaf(h2[j + 1], x1, hodd[2]);
    } else if (KIND == EPI_RAW) {
      gout[c0 + j] = x0 * (1.f / kActScale); gout[c0 + j + 1] = x1 * (1.f / kActScale);
    }
    if (KIND == EPI_SEM_WIDE) {                 // sem_dim 3..4: own kind, so that the common case carries no branches

#pragma unroll
      for (int s = 2; s < kSemMax; ++s)
        if (s < sem_dim) {
          hacc[s] = fmaf(hw[kHeadWS2 + s * kHalfMax + c0 + j], x0, hacc[s]);
          hodd[s] = fmaf(hw[kHeadWS2 + s * kHalfMax + c0 + j + 1], x1, hodd[s]);
        }
    }
    if (((DUMP && (KIND == EPI_HIDDEN_SIGMA || KIND == EPI_SEM || KIND == EPI_SEM_WIDE)) || (DALL && (KIND == EPI_HIDDEN || KIND == EPI_RGB))) && gout) {
      if (BLK) {
        // blocked layout (gout = this point's slot in its group of 32): one 128-byte line per warp store
        gout[(size_t)(c0 + j) * 32] = x0 * (1.f / kActScale);
        gout[(size_t)(c0 + j + 1) * 32] = x1 * (1.f / kActScale);
      } else if (j & 2) {
        // row-major: 16-byte stores (rows are W resp. W/2 floats, c0 + j a multiple of 4 on the odd pair)
        *reinterpret_cast<float4*>(gout + c0 + j - 2) = make_float4(dq0, dq1, x0 * (1.f / kActScale), x1 * (1.f / kActScale));
      } else { dq0 = x0 * (1.f / kActScale); dq1 = x1 * (1.f / kActScale); }
    }
    if (KIND == EPI_HIDDEN || KIND == EPI_HIDDEN_SIGMA) {
      vmax = fmaxf(vmax, fmaxf(x0, x1));          // range guard: 16*a must stay below the fp16 maximum (checked once per tile)
      __half2 hh = __floats2half2_rn(x0, x1);
      hi[j >> 1] = *reinterpret_cast<uint32_t*>(&hh);
      if (EXACT) {
        float2 hf = __half22float2(hh);
        __half2 ll = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
        lo[j >> 1] = *reinterpret_cast<uint32_t*>(&ll);
      }
    }
  }
  if (KIND == EPI_HIDDEN || KIND == EPI_HIDDEN_SIGMA) {
    // in place: the 16 fp32 accumulator columns just read become the fp16 hi / lo pairs of the same K-step
    tmem_st8(tm_lane + c0, hi);
    if (EXACT) tmem_st8(tm_lane + c0 + kALo, lo);
  }
}

template <int KIND, bool EXACT, bool DUMP, bool DALL = false, bool BLK = false>
__device__ __forceinline__ void epi_chunk(const uint32_t (&v)[kCW], uint32_t tm_lane, int c0, float inv16, const float* __restrict__ bias,
                                          const float* __restrict__ hw, int sem_dim, float (&hacc)[4], float (&hodd)[4],
                                          float* __restrict__ gout, float& vmax) {
#pragma unroll
  for (int sb = 0; sb < kCW / kSW; ++sb)
    epi_sub<KIND, EXACT, DUMP, DALL, BLK>(&v[kSW * sb], tm_lane, c0 + kSW * sb, inv16, bias, hw, sem_dim, hacc, hodd, gout, vmax);
}
__device__ __forceinline__ void tmem_ldc(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld16(taddr, r); }
__device__ __forceinline__ void tmem_wait_ldc(uint32_t (&r)[16]) { tmem_wait_ld_fence16(r); }

// tm_lane = TMEM address of this warp's lane quarter inside the stage's D buffer.
// Head kinds: 16-column chunks [cb, ce).  Hidden kinds (the output becomes the next layer's A operand, in place): both warps
// of a lane quarter work on the SAME 64-column K slab -- half hf takes two of the four chunks of slab j = 0 .. ce/2-1 -- and
// hand the slab to the MMA warp once their part is stored, so the next stage's MMAs start after a fraction of the epilogue
// instead of all of it.  Slab 0 is handed over in two halves (K-steps 0-1 after the first chunk of every warp, 2-3 after the
// second): a_ready[0], a_ready[1]; slab j >= 1 uses a_ready[1+j].
template <int KIND, bool EXACT, bool DUMP, bool DALL = false, bool BLK = false>
__device__ __forceinline__ void epi_kind(int cb, int ce, int hf, uint32_t tm_lane, float inv16, const float* bias, const float* hw,
                                         int sem_dim, float* hacc, float* gout, float& vmax, uint32_t a_ready0, long long* trs = nullptr) {
  constexpr bool SLAB = (KIND == EPI_HIDDEN || KIND == EPI_HIDDEN_SIGMA);
  // software pipeline: the tcgen05.ld of chunk i+1 is in flight while chunk i is processed
  uint32_t va[kCW], vb[kCW];
  const int cnt = ce - cb;
  if (cnt <= 0) return;
  auto col = [&](int i) { return (SLAB ? (i < 2 ? 2 * i + hf : 4 * (i >> 1) + 2 * hf + (i & 1)) : (cb + i)) * kCW; };
  auto hand_over = [&](uint32_t bar) { tmem_wait_st(); tc_fence_before(); mbar_arrive(a_ready0 + 8u * bar); };
  float* const hout = hacc;
  float he[4] = {0.f, 0.f, 0.f, 0.f}, ho[4] = {0.f, 0.f, 0.f, 0.f};
  tmem_ldc(tm_lane + col(0), va);
  tmem_wait_ldc(va);
  for (int i = 0; i < cnt; i += 2) {
    if (i + 1 < cnt) tmem_ldc(tm_lane + col(i + 1), vb);
    epi_chunk<KIND, EXACT, DUMP, DALL, BLK>(va, tm_lane, col(i), inv16, bias, hw, sem_dim, he, ho, gout, vmax);
    if (SLAB && i == 0) { hand_over(0u); if (trs) trs[9] = clock64(); }
    if (i + 1 < cnt) {
      tmem_wait_ldc(vb);
      if (i + 2 < cnt) tmem_ldc(tm_lane + col(i + 2), va);
      epi_chunk<KIND, EXACT, DUMP, DALL, BLK>(vb, tm_lane, col(i + 1), inv16, bias, hw, sem_dim, he, ho, gout, vmax);
      if (SLAB) {                       // my chunks of slab i/2 are stored
        hand_over(1u + (uint32_t)(i >> 1));
        if (trs && (i >> 1) < 2) trs[10 + (i >> 1)] = clock64();
      }
      if (i + 2 < cnt) tmem_wait_ldc(va);
    }
  }
  if (KIND == EPI_HIDDEN_SIGMA) hout[0] += he[0] + ho[0];
  if (KIND == EPI_RGB) { hout[0] += he[0] + ho[0]; hout[1] += he[1] + ho[1]; hout[2] += he[2] + ho[2]; }
  if (KIND == EPI_SEM) { hout[0] += he[0] + ho[0]; if (sem_dim > 1) hout[1] += he[1] + ho[1]; }
  if (KIND == EPI_SEM_WIDE) {
#pragma unroll
    for (int k = 0; k < kSemMax; ++k) if (k < sem_dim) hout[k] += he[k] + ho[k];
  }
}

template <bool EXACT, bool DUMP = false, bool DALL = false, bool BLK = false>
__device__ __forceinline__ void epilogue(int kind, int cb, int ce, int hf, uint32_t tm_lane, float inv16, const float* bias,
                                         const float* hw, int sem_dim, float* hacc, float* gout, float& vmax, uint32_t a_ready0,
                                         long long* trs = nullptr) {
#define NSOS_EPI(K) epi_kind<K, EXACT, DUMP, DALL, BLK>(cb, ce, hf, tm_lane, inv16, bias, hw, sem_dim, hacc, gout, vmax, a_ready0, trs)
  switch (kind) {
    case EPI_HIDDEN: NSOS_EPI(EPI_HIDDEN); break;
    case EPI_HIDDEN_SIGMA: NSOS_EPI(EPI_HIDDEN_SIGMA); break;
    case EPI_SEM:
      if (sem_dim <= 2) NSOS_EPI(EPI_SEM);
      else NSOS_EPI(EPI_SEM_WIDE);
      break;
    case EPI_RGB: NSOS_EPI(EPI_RGB); break;
    default: NSOS_EPI(EPI_RAW); break;
  }
#undef NSOS_EPI
}

// chunk range (units of kCW columns) of worker half `hf` for a stage of n columns
__device__ __forceinline__ void chunk_range(int n, int hf, int& cb, int& ce) {
  const int nch = n / kCW, mid = (nch + 1) >> 1;
  cb = hf ? mid : 0;
  ce = hf ? nch : mid;
}

__device__ __forceinline__ void init_pipeline(const Smem& sm, int nslots, int warp, uint32_t csize) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < nslots; ++i) { mbar_init(smem_u32(&sm.full[i]), 1); mbar_init(smem_u32(&sm.empty[i]), csize); }
    mbar_init(smem_u32(sm.acc_full), 1);
    mbar_init(smem_u32(sm.g_ready), kWorkers);
    for (int i = 0; i < kMaxASlabs; ++i) mbar_init(smem_u32(&sm.a_ready[i]), kWorkers);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) { tmem_alloc(smem_u32(sm.tmem_ptr), kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync();          // peers' barriers must be initialised before any multicast signal
  tc_fence_after();
}

// MODE 0: forward.  MODE 1: forward that also saves h_last / s_hid per point (training).  MODE 2: replay at given sample
// depths (backward recompute: no sampling, no compositing) with the same saves.  MODE 3: replay that saves EVERY hidden layer and
// the views hidden layer (recompute of the all-parameter backward).
// BLK (MODE 1 / 2): dump_h / dump_s0 / dump_enc in the blocked layout of TcParams::dump_blocked (both passes alike)
template <bool EXACT, int MODE, bool BLK = false>
__global__ void __launch_bounds__(kThreads, 1) k_render_tc(const __grid_constant__ TcParams P) {
  constexpr bool REPLAY = (MODE >= 2), DUMP = (MODE >= 1), DALL = (MODE == 3);
  static_assert(!BLK || (DUMP && !DALL), "the blocked layout belongs to the semantic-head saves");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS)
  Smem sm;
  carve_smem(base, P.nslots, P.Sc, P.Sf, P.C, &sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;
  const uint32_t csize = cluster_nctarank(), crank = cluster_ctarank();
  init_pipeline(sm, P.nslots, warp, csize);
  const uint32_t tm = *sm.tmem_ptr;
  const long long n_pairs = (P.n_rays + 1) / 2;
  // every CTA runs the same number of pair iterations (cluster peers stream the weights in lock step);
  // iterations past the end render a dummy pair whose outputs are suppressed
  const long long iters = (n_pairs + gridDim.x - 1) / gridDim.x;
  const int npass = P.fine ? 2 : 1;

  if (warp >= kProducerWarp + kNumProducers) {
    regs_other();                                                // spare warp (register redistribution build only)
  } else if (warp >= kProducerWarp) {
    regs_other();
    uint32_t chunk = 0;
    for (long long itp = 0; itp < iters; ++itp)
      for (int pass = 0; pass < npass; ++pass) {
        const int S = pass ? P.Sf : P.Sc, ntiles = (2 * S + 127) / 128;
        for (int tile = 0; tile < ntiles; ++tile)
          producer_tile(P.prog[pass], P.packed[pass], EXACT, sm, P.nslots, chunk, crank, csize, warp - kProducerWarp);
      }
  } else if (warp == kMmaWarp) {
    regs_other();
    for (int i = lane; i < 2 * kCtabMax; i += 32) sm.ctab[i] = P.ctab[i / kCtabMax][i % kCtabMax];
    __syncwarp();
    if (elect_one()) {
      // ONE elect block around the whole kernel: the issuing lane never reconverges between stages or tiles
      MmaState ms{RingPos{0u, 0u}, 0u, 0u, 0u};
      int ntile_seen = 0;
      mbar_wait(smem_u32(&sm.full[0]), 0u, 299);                // first weight chunk; later ones are pre-waited inside mma_tile
      tc_fence_after();
      for (long long itp = 0; itp < iters; ++itp)
        for (int pass = 0; pass < npass; ++pass) {
          const int S = pass ? P.Sf : P.Sc, ntiles = (2 * S + 127) / 128;
          for (int tile = 0; tile < ntiles; ++tile) {
            long long* tr = nullptr;
            if (P.trace && blockIdx.x == 0 && ntile_seen < kTraceTiles - 1) tr = P.trace + (size_t)ntile_seen * 16 * kTraceStamps;
            ++ntile_seen;
            const bool last_tile = (itp == iters - 1) && (pass == npass - 1) && (tile == ntiles - 1);
            mma_tile(sm.ctab + pass * kCtabMax, P.nch[pass], sm, (uint32_t)P.nslots, tm, ms, csize, last_tile, tr);
          }
        }
    }
    __syncwarp();
  } else {
    // ================= row workers: 8 warps; warp w owns TMEM lanes 32*(w%4).. and column half w/4 =================
    regs_worker();
    const int q4 = warp & 3, hf = warp >> 2;
    const int row = q4 * 32 + lane;
    const uint32_t tm_lane = tm + ((uint32_t)(q4 * 32) << 16);
    uint32_t it_acc = 0, it_bias = 0, gs = 0;      // gs: global stage counter, selects the D buffer exactly as the MMA warp does
    const uint32_t a_ready0 = smem_u32(&sm.a_ready[0]);
    int ntile_seen = 0;
    for (int net = 0; net < npass; ++net) {
      const TcAux* aux = reinterpret_cast<const TcAux*>(P.packed[net]);
      for (int i = t; i < kHeadFloats; i += kWorkers) sm.heads[net * kHeadFloats + i] = __ldg(&aux->heads[i]);
    }
    for (long long itp = 0; itp < iters; ++itp) {
      const long long pair_raw = (long long)blockIdx.x + itp * gridDim.x;
      const bool pair_valid = pair_raw < n_pairs;
      const long long pair = pair_valid ? pair_raw : 0;
      named_bar_sync(1, kWorkers);
      long long* pdbg = (P.trace && blockIdx.x == 0 && itp == 1 && t == 0) ? P.trace + (size_t)(kTraceTiles - 1) * 16 * kTraceStamps + 40 : nullptr;
      if (pdbg) pdbg[0] = clock64();
      // ---- per-pair ray setup: o, d, near, far, |d|, then gamma_v(d/|d|) with one (ray, octave, axis) per thread
      if (t < 2) {
        long long r = pair * 2 + t;
        bool valid = pair_valid && r < P.n_rays;
        long long rr = (r < P.n_rays) ? r : pair * 2;
        float* rp = sm.rayp + t * kRayP;
        const bool query = REPLAY && P.pts_in != nullptr;                 // point query: no rays, the direction is given
        float d0 = query ? P.qdir[0] : P.rays_d[rr * 3], d1 = query ? P.qdir[1] : P.rays_d[rr * 3 + 1], d2 = query ? P.qdir[2] : P.rays_d[rr * 3 + 2];
        rp[0] = query ? 0.f : P.rays_o[rr * 3]; rp[1] = query ? 0.f : P.rays_o[rr * 3 + 1]; rp[2] = query ? 0.f : P.rays_o[rr * 3 + 2];
        rp[3] = d0; rp[4] = d1; rp[5] = d2;
        rp[6] = REPLAY ? 0.f : P.near[rr]; rp[7] = REPLAY ? 1.f : P.far[rr];
        float nrm = query ? 1.f : sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)));
        rp[8] = nrm; rp[9] = valid ? 1.f : 0.f;
        rp[10] = __fdiv_rn(d0, nrm); rp[11] = __fdiv_rn(d1, nrm); rp[12] = __fdiv_rn(d2, nrm);   // nerf_net.py:165
      }
      named_bar_sync(1, kWorkers);
      if (t < 64) {
        // thread = (ray, octave slot, axis): slot 0 copies v, slot k+1 writes sin/cos(2^k v) (embedder.py:34-48); columns >= encv are zero
        const int rl = t >> 5, k = (t & 31) / 3, a = (t & 31) % 3;
        if ((t & 31) < 15) {
          const float v = sm.rayp[rl * kRayP + 10 + a];
          float* ev = sm.encv + rl * 28;
          if (k == 0) ev[a] = v;
          else if (k - 1 < P.prog[0].Lv) {
            float sn, cs;
            sincos_cw(__fmul_rn(v, (float)(1 << (k - 1))), &sn, &cs);
            ev[3 + 6 * (k - 1) + a] = sn; ev[6 + 6 * (k - 1) + a] = cs;
          } else { ev[3 + 6 * (k - 1) + a] = 0.f; ev[6 + 6 * (k - 1) + a] = 0.f; }
          if (t == 32 * rl) ev[27] = 0.f;
        }
      }
      named_bar_sync(1, kWorkers);
      if (pdbg) pdbg[1] = clock64();
      // view-direction half of views_linears.0 folded into a per-ray bias (fp32, x16): dirbias[net][ray][j].
      // One (net, j) per thread: the weight row is read once (7 x 16 B) and used for both rays of the pair.
      {
        const int net = t >> 7, j = t & (kHalfMax - 1);
        if (net < npass) {
          const TcAux* aux = reinterpret_cast<const TcAux*>(P.packed[net]);
          const TcProg& pg = P.prog[net];
          float acc0 = 0.f, acc1 = 0.f;
          if (j < pg.H2) {
            const float b = __ldg(&aux->bias[pg.nst - 1][pg.ub_off + j]);   // 16 * (b_views + W_v[:, :W] b_feat)
            const float4* wr = reinterpret_cast<const float4*>(&aux->w_vdir[j][0]);
            float d0 = 0.f, d1 = 0.f;
#pragma unroll
            for (int c4 = 0; c4 < 7; ++c4) {
              const float4 w = __ldg(wr + c4);
              const float4 e0 = *reinterpret_cast<const float4*>(sm.encv + 4 * c4), e1 = *reinterpret_cast<const float4*>(sm.encv + 28 + 4 * c4);
              d0 = fmaf(w.x, e0.x, d0); d0 = fmaf(w.y, e0.y, d0); d0 = fmaf(w.z, e0.z, d0); d0 = fmaf(w.w, e0.w, d0);
              d1 = fmaf(w.x, e1.x, d1); d1 = fmaf(w.y, e1.y, d1); d1 = fmaf(w.z, e1.z, d1); d1 = fmaf(w.w, e1.w, d1);
            }
            acc0 = fmaf(d0, kActScale, b); acc1 = fmaf(d1, kActScale, b);
          }
          sm.dirbias[(net * 2 + 0) * kHalfMax + j] = acc0;
          sm.dirbias[(net * 2 + 1) * kHalfMax + j] = acc1;
        }
      }
      named_bar_sync(1, kWorkers);
      if (pdbg) pdbg[2] = clock64();

#pragma unroll 1
      for (int pass = 0; pass < npass; ++pass) {
        const TcProg& pg = P.prog[pass];
        const TcAux* aux = reinterpret_cast<const TcAux*>(P.packed[pass]);
        const float* hw = sm.heads + pass * kHeadFloats;
        const int S = pass ? P.Sf : P.Sc, ntiles = (2 * S + 127) / 128;
#pragma unroll 1
        for (int tile = 0; tile < ntiles; ++tile) {
          // ---- tile setup: sample position, point, gamma(x) -> swizzled smem A tile (each half-warp-group writes 32 columns)
          long long* tr = nullptr;
          if (P.trace && blockIdx.x == 0 && t == 0 && ntile_seen < kTraceTiles - 1) tr = P.trace + (size_t)ntile_seen * 16 * kTraceStamps;
          ++ntile_seen;
          if (tr) tr[15 * kTraceStamps + 0] = clock64();                   // tile setup starts
          const int q = tile * 128 + row;
          const bool rowvalid = q < 2 * S;
          const int rl = rowvalid ? q / S : 0, i = rowvalid ? q % S : 0;
          const float* rp = sm.rayp + rl * kRayP;
          const long long ray = pair * 2 + rl;
          {
            float z;
            if (REPLAY) {
              z = (rowvalid && rp[9] > 0.f && !P.pts_in) ? P.z_in[pass][ray * S + i] : 1.f;
            } else if (pass == 0) {
              const bool pert = P.perturb > 0.f;
              float tr = 0.f;
              if (pert) tr = (P.rnd.t_rand && rp[9] > 0.f) ? P.rnd.t_rand[ray * P.Sc + i] : rng_uniform(P.seed, ray, RNG_T_RAND, i);
              z = z_stratified(rp[6], rp[7], i, S, pert, tr);
              if (rowvalid && hf == 0) sm.zc[rl * P.Sc + i] = z;
            } else {
              z = sm.zf[rl * P.Sf + i];
            }
            float x[3] = {pt_coord(rp[0], rp[3], z), pt_coord(rp[1], rp[4], z), pt_coord(rp[2], rp[5], z)};
            if (REPLAY && P.pts_in) {                                        // point query: the row's point is read, not sampled
              const long long pt = ray * S + i;
              const bool pv = rowvalid && pt < P.n_pts;
#pragma unroll
              for (int k = 0; k < 3; ++k) x[k] = pv ? P.pts_in[pt * 3 + k] : 0.f;
            }
            float e[32];
            if (hf == 0) encode_half<0>(x, pg.Lp, pg.enc, rowvalid, e); else encode_half<1>(x, pg.Lp, pg.enc, rowvalid, e);
            store_halfrow_sw128(sm.g_hi, sm.g_lo, row, hf, e, EXACT);
            if (DUMP && P.dump_enc[pass] && rowvalid && rp[9] > 0.f) {       // e = 16 * gamma(x): exact power-of-two rescale
              const size_t pt = (size_t)ray * S + i;
              if (BLK) {
                float* ge = P.dump_enc[pass] + ((pt >> 5) * 64 + 32 * hf) * 32 + (pt & 31);
#pragma unroll
                for (int c = 0; c < 32; ++c) ge[c * 32] = e[c] * (1.f / kActScale);
              } else {
                float4* ge = reinterpret_cast<float4*>(P.dump_enc[pass] + pt * 64 + 32 * hf);
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4)
                  ge[c4] = make_float4(e[4 * c4] * (1.f / kActScale), e[4 * c4 + 1] * (1.f / kActScale), e[4 * c4 + 2] * (1.f / kActScale),
                                       e[4 * c4 + 3] * (1.f / kActScale));
              }
            }
          }
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(smem_u32(sm.g_ready));
          if (tr) tr[15 * kTraceStamps + 1] = clock64();                   // gamma tile written, g_ready signalled

          float hacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // [0] sigma, [1..3] rgb, [4..7] sem (partial over my columns)
          float vmax = 0.f;
#pragma unroll 1
          for (int st = 0; st < pg.nst; ++st) {
            const TcStage& Sg = pg.st[st];
            float* sb = sm.sbias + (it_bias & 1u) * 256;
            ++it_bias;
            sb[t] = __ldg(&aux->bias[st][t]);                            // kWorkers == 256 == bias row length
            const bool merged = Sg.epi == EPI_SEM_RGB;                   // columns [0,H2) semantic hidden | [H2,2*H2) views hidden
            const float inv16 = __ldg(&aux->inv_scale[st][merged ? hf : 0]);
            named_bar_sync(1, kWorkers);
            const bool rgb_part = Sg.epi == EPI_RGB || (merged && hf == 1);
            const bool sem_part = merged && hf == 0;
            const float* bias = rgb_part ? (sm.dirbias + (pass * 2 + rl) * kHalfMax) : sb;
            const bool hidden = Sg.epi == EPI_HIDDEN || Sg.epi == EPI_HIDDEN_SIGMA;
            int cb, ce;
            if (merged) { cb = 0; ce = pg.H2 / kCW; }
            else if (hidden) { cb = 0; ce = 2 * (Sg.n / 64); }             // two chunks of every 64-column K slab (see epi_kind)
            else chunk_range(Sg.n, hf, cb, ce);
            const uint32_t tm_d = tm_lane + (gs & 1u) * kBufCols;
            ++gs;
            if (tr) tr[st * kTraceStamps + 0] = clock64();                 // worker starts waiting for the accumulator
            mbar_wait(smem_u32(sm.acc_full), it_acc & 1u, 500 + st);
            ++it_acc;
            tc_fence_after();
            if (tr) tr[st * kTraceStamps + 1] = clock64();                 // accumulator ready
            float* ha = (Sg.epi == EPI_HIDDEN_SIGMA) ? &hacc[0] : sem_part ? &hacc[4] : &hacc[1];
            float* gout = nullptr;
            constexpr bool blk = BLK;
            if (DUMP && rowvalid && rp[9] > 0.f) {
              const size_t pt = (size_t)ray * S + i;
              if (Sg.epi == EPI_HIDDEN_SIGMA && P.dump_h[pass]) gout = P.dump_h[pass] + (blk ? (pt >> 5) * 32 * pg.W + (pt & 31) : pt * pg.W);
              if (sem_part && P.dump_s0[pass]) gout = P.dump_s0[pass] + (blk ? (pt >> 5) * 32 * pg.H2 + (pt & 31) : pt * pg.H2);
              if (DALL) {
                if (Sg.epi == EPI_HIDDEN && P.dump_all[pass][st]) gout = P.dump_all[pass][st] + pt * pg.W;
                if (rgb_part && P.dump_hv[pass]) gout = P.dump_hv[pass] + pt * pg.H2;
              }
            }
            const int kind = merged ? (hf ? EPI_RGB : EPI_SEM) : Sg.epi;
            epilogue<EXACT, DUMP, DALL, BLK>(kind, cb, ce, hf, tm_d + ((merged && hf) ? (uint32_t)pg.H2 : 0u), inv16, bias, hw, P.sem_dim, ha, gout, vmax, a_ready0,
                                  tr ? tr + st * kTraceStamps : nullptr);
            if (Sg.epi == EPI_HIDDEN_SIGMA && hf == 1) sm.hpart[row * 8] = hacc[0];   // sigma share of the upper column half
            if (tr) tr[st * kTraceStamps + 2] = clock64();                 // epilogue done (this thread)
          }
          // fp16(16*a) saturates at |a| > 4094 (inf hi plane, NaN lo plane -- which the next ReLU's fmaxf would quietly turn
          // into 0): poison this point's colour / semantic outputs so the ray's maps are NaN, and name the cause in the
          // sticky status word
          if (!(vmax <= 65504.f)) {
            hacc[0] = hacc[1] = hacc[4] = __int_as_float(0x7fc00000);
            if (P.out.status && rowvalid && rp[9] > 0.f) atomicOr(P.out.status, 1u);
          }
          // ---- emit the raw outputs of this row: [rgb(3), sigma, sem...] (nerf_mlp.py:94)
          float* graw = (pass == 0 && P.fine) ? P.out.raw0 : P.out.raw;
          float* gr_row = (graw && rp[9] > 0.f) ? graw + ((size_t)ray * S + i) * P.C : nullptr;
          if (REPLAY && P.pts_in && ray * S + i >= P.n_pts) gr_row = nullptr;        // tail of a point query
          if (pg.st[pg.nst - 1].epi == EPI_SEM_RGB) {
            // merged head stage: worker half 1 owns the rgb sums, half 0 the semantic sums; the sigma share of half 1 was
            // exchanged after the last trunk layer (at least one named barrier ago), so no barrier is needed here and the
            // next tile's setup starts at once
            if (rowvalid) {
              float* rw = sm.rawbuf + (size_t)q * P.C;
              if (hf == 0) {
                rw[3] = hacc[0] + sm.hpart[row * 8] + hw[kHeadBAlpha];
                for (int s = 0; s < P.sem_dim; ++s) rw[4 + s] = hacc[4 + s] + hw[kHeadBS2 + s];
                if (gr_row) for (int c = 3; c < P.C; ++c) gr_row[c] = rw[c];
              } else {
                rw[0] = hacc[1] + hw[kHeadBRgb]; rw[1] = hacc[2] + hw[kHeadBRgb + 1]; rw[2] = hacc[3] + hw[kHeadBRgb + 2];
                if (gr_row) { gr_row[0] = rw[0]; gr_row[1] = rw[1]; gr_row[2] = rw[2]; }
              }
            }
          } else {
            if (hf == 1) {
#pragma unroll
              for (int c = 0; c < 8; ++c) sm.hpart[row * 8 + c] = hacc[c];
            }
            named_bar_sync(1, kWorkers);
            if (hf == 0 && rowvalid) {
              const float* hp = sm.hpart + row * 8;
              float* rw = sm.rawbuf + (size_t)q * P.C;
              rw[0] = hacc[1] + hp[1] + hw[kHeadBRgb]; rw[1] = hacc[2] + hp[2] + hw[kHeadBRgb + 1]; rw[2] = hacc[3] + hp[3] + hw[kHeadBRgb + 2];
              rw[3] = hacc[0] + hp[0] + hw[kHeadBAlpha];
              for (int s = 0; s < P.sem_dim; ++s) rw[4 + s] = hacc[4 + s] + hp[4 + s] + hw[kHeadBS2 + s];
              if (gr_row) for (int c = 0; c < P.C; ++c) gr_row[c] = rw[c];
            }
          }
          if (tr) tr[15 * kTraceStamps + 2] = clock64();                   // raw outputs of the tile emitted
        }
        // ---- compositing (+ resampling after the coarse pass): 4 warps (128 threads) per ray of the pair.
        // The gamma tiles are idle between tiles (all MMAs of the tile have completed): their first bytes serve as
        // reduction scratch here.
        named_bar_sync(1, kWorkers);
        if (!REPLAY) {
          const int gr = warp >> 2, gt = t & (kGroup - 1), gbar = 2 + gr;
          long long* gdbg = (P.trace && blockIdx.x == 0 && itp == 0 && t == 0) ? P.trace + (size_t)(kTraceTiles - 1) * 16 * kTraceStamps + pass * 16 : nullptr;
          if (gdbg) gdbg[8] = clock64();
          GroupScratch gsc;
          gsc.f = reinterpret_cast<float*>(sm.g_hi + gr * 512);
          gsc.d = reinterpret_cast<double*>(sm.g_hi + 1024 + gr * 64);
          const float* rp = sm.rayp + gr * kRayP;
          const long long ray = pair * 2 + gr;
          const bool valid = rp[9] > 0.f;
          const bool coarse_of_two = (pass == 0 && P.fine);
          RayPass rpx;
          rpx.raw = sm.rawbuf + (size_t)gr * S * P.C;
          rpx.z = pass ? sm.zf + gr * P.Sf : sm.zc + gr * P.Sc;
          const float* nz = pass ? P.rnd.noise1 : P.rnd.noise0;
          rpx.noise = (nz && valid) ? nz + ray * S : nullptr;
          rpx.noise_std = P.noise_std; rpx.seed = P.seed; rpx.ray = ray; rpx.rng_stream = pass ? RNG_NOISE1 : RNG_NOISE0;
          rpx.dnorm = rp[8]; rpx.S = S; rpx.C = P.C; rpx.sem_dim = P.sem_dim; rpx.white_bkgd = P.white_bkgd;
          float* maps = valid ? P.out.maps + (size_t)ray * P.ML + (coarse_of_two ? P.C6 : 0) : nullptr;
          float* wsm = sm.w0 + gr * P.Sc;     // coarse weights stay in smem for the resampling
          float* gw = coarse_of_two ? P.out.weights0 : P.out.weights;
          float* wout = (pass == 0) ? wsm : ((gw && valid) ? gw + ray * S : nullptr);
          group_composite(rpx, gt, gbar, gsc, maps, wout, gdbg ? gdbg + 11 : nullptr);      // maps == nullptr: dummy ray, nothing is written
          if (gdbg) gdbg[9] = clock64();
          if (pass == 0) {
            if (gw && valid) for (int k = gt; k < S; k += kGroup) gw[ray * S + k] = wsm[k];
            float* gz = coarse_of_two ? P.out.z_vals0 : P.out.z_vals;
            if (gz && valid) for (int k = gt; k < S; k += kGroup) gz[ray * S + k] = rpx.z[k];
            if (P.fine) {
              ImportanceIO io;
              io.z0 = rpx.z; io.w0 = wsm; io.cdf = sm.cdf + gr * P.Sc; io.bins = sm.bins + gr * P.Sc;
              io.zall = sm.zall + gr * P.Sf; io.zsorted = sm.zf + gr * P.Sf;
              io.u = (P.rnd.u && valid) ? P.rnd.u + ray * P.K : nullptr;
              io.z_inject = (P.rnd.z_samples && valid) ? P.rnd.z_samples + ray * P.K : nullptr;
              io.z_samples = (P.out.z_samples && valid) ? P.out.z_samples + ray * P.K : nullptr;
              io.inds = (P.out.inds && valid) ? P.out.inds + ray * P.K : nullptr;
              io.z_std = valid ? P.out.maps + (size_t)ray * P.ML + 2 * P.C6 : nullptr;
              io.Sc = P.Sc; io.K = P.K; io.det = !(P.perturb > 0.f); io.seed = P.seed; io.ray = ray;
              io.dbg = (gdbg && pass == 0) ? gdbg + 32 + 24 : nullptr;      // trace slots 56..59 of the last trace tile
              group_importance(io, gt, gbar, gsc);
              if (P.out.z_vals && valid) for (int k = gt; k < P.Sf; k += kGroup) P.out.z_vals[ray * P.Sf + k] = io.zsorted[k];
            }
          }
          if (gdbg) gdbg[10] = clock64();
        }
        named_bar_sync(1, kWorkers);
      }
    }
  }
  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync();           // no CTA may exit while a peer can still signal its barriers
  if (warp == kMmaWarp) tmem_dealloc(tm, kTmemCols);
}

// ---- self test: D[128,N] = A[128,K] . W[N,K]^T through the same producer / MMA / epilogue code ----------------
struct SelfParams {
  TcProg prog;
  uint32_t ctab[kCtabMax];
  int nch;
  const uint8_t* packed;
  const float* a;
  float* d;
  int K, N, a_in_tmem, nslots;
};
template <bool EXACT>
__global__ void __launch_bounds__(kThreads, 1) k_selftest(const __grid_constant__ SelfParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS)
  Smem sm;
  carve_smem(base, P.nslots, 2, 2, 8, &sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;
  init_pipeline(sm, P.nslots, warp, 1);
  const uint32_t tm = *sm.tmem_ptr;
  if (warp >= kProducerWarp + kNumProducers) {
  } else if (warp >= kProducerWarp) {
    uint32_t chunk = 0;
    producer_tile(P.prog, P.packed, EXACT, sm, P.nslots, chunk, 0, 1, warp - kProducerWarp);
  } else if (warp == kMmaWarp) {
    for (int i = lane; i < kCtabMax; i += 32) sm.ctab[i] = P.ctab[i];
    __syncwarp();
    if (elect_one()) {
      MmaState ms{RingPos{0u, 0u}, 0u, 0u, 0u};
      mbar_wait(smem_u32(&sm.full[0]), 0u, 299);
      tc_fence_after();
      mma_tile(sm.ctab, P.nch, sm, (uint32_t)P.nslots, tm, ms, 1, true);
    }
    __syncwarp();
  } else {
    const int q4 = warp & 3, hf = warp >> 2, row = q4 * 32 + lane;
    const uint32_t tm_lane = tm + ((uint32_t)(q4 * 32) << 16);
    const TcAux* aux = reinterpret_cast<const TcAux*>(P.packed);
    if (P.a_in_tmem) {
      // each worker half writes its half of the K columns (16-column chunks) of the A planes
      int cb, ce;
      chunk_range(P.K, hf, cb, ce);
      for (int c = cb * (kCW / kSW); c < ce * (kCW / kSW); ++c) {
        const int c0 = c * kSW;
        uint32_t hi[kSW / 2], lo[kSW / 2];
#pragma unroll
        for (int j = 0; j < kSW; j += 2) {
          float a0 = P.a[(size_t)row * P.K + c0 + j] * kActScale, a1 = P.a[(size_t)row * P.K + c0 + j + 1] * kActScale;
          __half2 hh = __floats2half2_rn(a0, a1);
          float2 hf2 = __half22float2(hh);
          __half2 ll = __floats2half2_rn(a0 - hf2.x, a1 - hf2.y);
          hi[j >> 1] = *reinterpret_cast<uint32_t*>(&hh);
          lo[j >> 1] = *reinterpret_cast<uint32_t*>(&ll);
        }
        tmem_st8(tm_lane + kBufCols + c0, hi);              // stage 0 reads its A operand from buffer 1, in-place K-step layout
        if (EXACT) tmem_st8(tm_lane + kBufCols + c0 + kALo, lo);
      }
      tmem_wait_st();
    } else {
      float e[32];
      for (int c = 0; c < 32; ++c) { int col = hf * 32 + c; e[c] = (col < P.K) ? P.a[(size_t)row * P.K + col] * kActScale : 0.f; }
      store_halfrow_sw128(sm.g_hi, sm.g_lo, row, hf, e, EXACT);
      fence_proxy_async_smem();
    }
    sm.sbias[t] = 0.f;
    named_bar_sync(1, kWorkers);
    tc_fence_before();
    mbar_arrive(smem_u32(sm.g_ready));
    for (int j = 0; j < kMaxASlabs; ++j) mbar_arrive(smem_u32(&sm.a_ready[j]));
    mbar_wait(smem_u32(sm.acc_full), 0, 600);
    tc_fence_after();
    float dummy[4];
    int cb, ce;
    chunk_range(P.N, hf, cb, ce);
    float vm = 0.f;
    epilogue<EXACT>(EPI_RAW, cb, ce, hf, tm_lane, __ldg(&aux->inv_scale[0][0]), sm.sbias, nullptr, 0, dummy, P.d + (size_t)row * P.N, vm, 0u);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tm, kTmemCols);
}

// ---- row GEMM: C[P,N] (=|+=) epi(A[P,K] . B[K,N]) for the all-parameter backward (dgrad products, feature_linear) ----------
// The same pipeline as a one-stage tile of the render kernel: the eight worker warps read 128 rows of A (fp32, global),
// split every value into bf16 hi + lo (fp32 range: gradients sit far below the fp16 range) and store them as the TMEM A
// operand; B (a weight matrix or its transpose, <= 256 x 256) is packed once per call into bf16 hi/lo SWIZZLE_128B slabs and
// streamed through the bulk-copy ring; 3 MMAs per product (hi.hi + lo.hi + hi.lo, fp32 accumulate); the epilogue applies
// bias / ReLU / ReLU-mask / accumulate and writes fp32 rows.  Persistent CTAs over 128-row tiles.
struct RowGemmParams {
  TcProg prog;
  uint32_t ctab[kCtabMax];
  int nch;
  const uint8_t* packed;      // slab image; producer_tile reads from packed + kAuxBytes
  const float* A; long long lda; int K;      // K in {64,128,192,256}
  float* C; long long ldc; int N;            // N multiple of 32, <= 256
  const float* mask; long long mask_ld;      // keep where mask(m,n) > 0
  const float* bias;
  int accumulate, relu;
  long long P;
  int nslots;
  int mask_rows;              // 1 (default): the row-owning threads read their mask rows themselves under the MMAs; 0 (NSOS_RG_MASK_BITS=1):
                              // whole rows per half warp, sign bits by ballot through the C image
};
// B(k,n) = B[k*b_rs + n*b_cs]  ->  per 64-wide K slab: bf16 hi plane [N x 64] then lo plane, K-major SWIZZLE_128B
__global__ void k_pack_b_bf16(const float* __restrict__ B, long long b_rs, long long b_cs, int K, int N, uint8_t* __restrict__ out) {
  const int j = blockIdx.x;                                   // K slab
  uint8_t* hi = out + (size_t)j * 2 * N * 128;
  uint8_t* lo = hi + (size_t)N * 128;
  for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < N * 64; e += gridDim.y * blockDim.x) {
    const int n = e >> 6, k = e & 63, kk = 64 * j + k;
    const float w = kk < K ? B[(long long)kk * b_rs + (long long)n * b_cs] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(w);
    const __nv_bfloat16 l = __float2bfloat16_rn(w - __bfloat162float(h));
    const size_t byte = (size_t)(n >> 3) * 1024 + (size_t)(n & 7) * 128 + (size_t)(((k >> 3) ^ (n & 7)) << 4) + (size_t)(k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(hi + byte) = h;
    *reinterpret_cast<__nv_bfloat16*>(lo + byte) = l;
  }
}
// Shared memory of k_rowgemm: the weight ring, FOUR 32 KB staging images (three rotate over the A slabs requested ahead with cp.async,
// the fourth carries the C slabs of the epilogue), barriers, the issuer's chunk table.
constexpr int kRgImgBytes = 32768, kRgSlots = 3;
__host__ __device__ inline size_t carve_rowgemm(uint8_t* base, int nslots, Smem* s, uint8_t** img) {
  size_t off = 0;
  auto take = [&](size_t bytes, size_t align) { off = (off + align - 1) / align * align; size_t o = off; off += bytes; return o; };
  size_t o_ring = take((size_t)nslots * kSlotBytes, 1024);
  size_t o_img = take((size_t)4 * kRgImgBytes, 1024);
  size_t o_bar = take(sizeof(uint64_t) * (2 * kMaxSlots + 2 + kMaxASlabs), 8), o_tp = take(16, 16);
  size_t o_ct = take(sizeof(uint32_t) * 2 * kCtabMax, 16);
  if (s) {
    *s = Smem{};
    s->ring = base + o_ring; s->g_hi = base + o_img + 3 * kRgImgBytes; s->g_lo = s->g_hi + kGBytes;
    s->full = (uint64_t*)(base + o_bar); s->empty = s->full + kMaxSlots; s->acc_full = s->empty + kMaxSlots;
    s->g_ready = s->acc_full + 1; s->a_ready = s->g_ready + 1; s->tmem_ptr = (uint32_t*)(base + o_tp);
    s->ctab = (uint32_t*)(base + o_ct);
    *img = base + o_img;
  }
  return off;
}
__device__ __forceinline__ void rg_cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {     // src_bytes 0: 16 zero bytes
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b)) << 16);
}
__global__ void __launch_bounds__(kThreads, 1) k_rowgemm(const __grid_constant__ RowGemmParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Smem sm;
  uint8_t* img0;
  carve_rowgemm(base, P.nslots, &sm, &img0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  init_pipeline(sm, P.nslots, warp, 1);
  const uint32_t tm = *sm.tmem_ptr;
  const long long ntiles = (P.P + 127) / 128;
  const long long my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;      // >= 1 (grid <= ntiles)
  if (warp >= kProducerWarp + kNumProducers) {
  } else if (warp >= kProducerWarp) {
    uint32_t chunk = 0;
    for (long long it = 0; it < my_tiles; ++it) producer_tile(P.prog, P.packed, true, sm, P.nslots, chunk, 0, 1, warp - kProducerWarp);
  } else if (warp == kMmaWarp) {
    for (int i = lane; i < kCtabMax; i += 32) sm.ctab[i] = P.ctab[i];
    __syncwarp();
    if (elect_one()) {
      MmaState ms{RingPos{0u, 0u}, 0u, 0u, 0u, kIdescBf16};
      mbar_wait(smem_u32(&sm.full[0]), 0u, 299);
      tc_fence_after();
      for (long long it = 0; it < my_tiles; ++it) mma_tile(sm.ctab, P.nch, sm, (uint32_t)P.nslots, tm, ms, 1, it == my_tiles - 1);
    }
    __syncwarp();
  } else {
    const int q4 = warp & 3, hf = warp >> 2, row = q4 * 32 + lane;
    const uint32_t tm_lane = tm + ((uint32_t)(q4 * 32) << 16);
    uint32_t it_acc = 0;
    // Row-major operands, row-owning threads: a thread owns a ROW of the tile (its TMEM lane), so direct 16-byte accesses of a warp
    // touch 32 different 128-byte lines each (8192 L1 tag cycles per operand and tile where 1024 do).  Every 64-column slab of A and
    // of C therefore passes through a 32 KB staging image in shared memory: warps move whole rows between global and shared memory
    // (one row = 256 bytes per 16 lanes), lanes pick up / drop their own row from the image.
    // Image: row r at r*256 bytes, its 16-byte unit u at ((u ^ (r & 15)) << 4): both access patterns are conflict-free.
    // A slabs are requested THREE ahead with cp.async into three rotating images (one commit group per slab, zero fill beyond P):
    // 96 KB per SM are in flight under the MMAs and the epilogue of the previous tile, and a slab's image is requested again as soon
    // as it has been converted.  (One slab ahead in registers left ~32 KB in flight per SM: the A fill waited for memory.)
    const int t = threadIdx.x;                                   // 0..255: the worker warps are warps 0..7
    uint8_t* S = sm.g_hi;                                        // the C image
    const int nsa = P.K / 64, nsc = P.N / 64;
    const long long tile0 = (long long)blockIdx.x * 128;
    const long long nslabs_total = my_tiles * nsa;
    // request state: next slab to request (tile rq_it, slab rq_j), rotating image index
    long long rq = 0, rq_it = 0;
    int rq_j = 0, rq_img = 0;
    const uint32_t img_u32 = smem_u32(img0);
    auto request = [&]() {
      if (rq < nslabs_total) {
        const long long trow = tile0 + rq_it * (long long)gridDim.x * 128;
        const uint32_t dst0 = img_u32 + (uint32_t)rq_img * kRgImgBytes;
        const float* src0 = P.A + 64 * rq_j + 4 * (t & 15);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = (t + 256 * i) >> 4;
          const long long gr = trow + rr;
          const bool v = gr < P.P;
          rg_cp_async16(dst0 + (uint32_t)(rr * 256 + (((t & 15) ^ (rr & 15)) << 4)), src0 + (v ? gr : 0) * P.lda, v ? 16u : 0u);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");        // (an empty group past the end keeps the group count uniform)
      ++rq;
      if (++rq_j == nsa) { rq_j = 0; ++rq_it; }
      if (++rq_img == 3) rq_img = 0;
    };
    request(); request(); request();
    int cur_img = 0;
    for (long long it = 0; it < my_tiles; ++it) {
      const long long trow0 = tile0 + it * (long long)gridDim.x * 128;
      const long long r = trow0 + row;
      const bool valid = r < P.P;
      const uint32_t dbuf = (uint32_t)(it & 1) * kBufCols, abuf = dbuf ^ kBufCols;
      // ---- A slabs: staging image -> this thread's row, columns [32 hf, 32 hf + 32) of the slab -> bf16 hi/lo -> TMEM
      for (int j = 0; j < nsa; ++j) {
        asm volatile("cp.async.wait_group 2;" ::: "memory");     // this thread's pieces of the slab have landed
        named_bar_sync(1, kWorkers);                             // ... and everybody else's
        const uint8_t* Srow = img0 + cur_img * kRgImgBytes + row * 256;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          float v[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 t4 = *reinterpret_cast<const float4*>(Srow + (((8 * hf + 4 * cc + q) ^ (row & 15)) << 4));
            v[4 * q] = t4.x; v[4 * q + 1] = t4.y; v[4 * q + 2] = t4.z; v[4 * q + 3] = t4.w;
          }
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int k2 = 0; k2 < 16; k2 += 2) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(v[k2]), h1 = __float2bfloat16_rn(v[k2 + 1]);
            hi[k2 >> 1] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            lo[k2 >> 1] = pack_bf16x2(v[k2] - __bfloat162float(h0), v[k2 + 1] - __bfloat162float(h1));
          }
          const int c = 4 * j + 2 * hf + cc;                     // 16-column K-step chunk of A
          tmem_st8(tm_lane + abuf + 16 * c, hi);
          tmem_st8(tm_lane + abuf + 16 * c + kALo, lo);
        }
        named_bar_sync(1, kWorkers);                             // the image is free again
        request();                                               // three slabs ahead, into the image just read
        if (++cur_img == 3) cur_img = 0;
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(smem_u32(sm.g_ready));
      for (int j = 0; j < kMaxASlabs; ++j) mbar_arrive(smem_u32(&sm.a_ready[j]));
      // ---- under the MMAs: the ReLU mask of this thread's C columns (slab j: [64 j + 32 hf, +32)) as bits, the next tile's slab 0
      const float* mrow = P.mask ? P.mask + r * P.mask_ld : nullptr;
      uint32_t mbits[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};     // word j: the 32 columns of slab j
      if (P.mask && !P.mask_rows) {
        // Whole rows per half warp (a warp instruction = two rows x 256 bytes), one ballot per vector component: the 16 sign bits of
        // a row's units land in the C image (free until the epilogue), 8 bytes per row and slab, and every thread picks its row up.
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {                         // two batches of 16 loads (slabs 2 hb, 2 hb + 1)
          float4 mb[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            mb[i] = make_float4(1.f, 1.f, 1.f, 1.f);
            const int j = 2 * hb + (i >> 3);
            const long long gr = trow0 + ((t + 256 * (i & 7)) >> 4);
            if (j < nsc && gr < P.P) mb[i] = __ldg(reinterpret_cast<const float4*>(P.mask + gr * P.mask_ld + 64 * j) + (t & 15));
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int j = 2 * hb + (i >> 3), rr = (t + 256 * (i & 7)) >> 4;
            const uint32_t b0 = __ballot_sync(0xffffffffu, mb[i].x > 0.f), b1 = __ballot_sync(0xffffffffu, mb[i].y > 0.f),
                           b2 = __ballot_sync(0xffffffffu, mb[i].z > 0.f), b3 = __ballot_sync(0xffffffffu, mb[i].w > 0.f);
            if ((lane & 15) == 0 && j < nsc) {                   // lane 0: row rr (low halves), lane 16: its row rr (high halves)
              const int sh = lane;                               // 0 or 16
              const uint32_t w0 = ((b0 >> sh) & 0xffffu) | (((b1 >> sh) & 0xffffu) << 16), w1 = ((b2 >> sh) & 0xffffu) | (((b3 >> sh) & 0xffffu) << 16);
              *reinterpret_cast<uint2*>(S + (size_t)(j * 128 + rr) * 8) = make_uint2(w0, w1);
            }
          }
        }
        named_bar_sync(1, kWorkers);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j < nsc) {
            const uint2 w = *reinterpret_cast<const uint2*>(S + (size_t)(j * 128 + row) * 8);
            uint32_t bits = 0u;                                  // bit 16 cc + 4 q + i  <-  component i of unit 8 hf + 4 cc + q
#pragma unroll
            for (int cc = 0; cc < 2; ++cc)
#pragma unroll
              for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  bits |= ((((i < 2) ? w.x : w.y) >> (16 * (i & 1) + 8 * hf + 4 * cc + q)) & 1u) << (16 * cc + 4 * q + i);
            mbits[j] = bits;
          }
        }
        named_bar_sync(1, kWorkers);                             // everybody has its bits before the epilogue reuses the image
      } else if (mrow && valid) {
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {                         // two batches of 16 loads (slabs 2 hb, 2 hb + 1)
          float4 mb[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            mb[i] = make_float4(1.f, 1.f, 1.f, 1.f);
            const int j = 2 * hb + (i >> 3);
            if (j < nsc) mb[i] = __ldg(reinterpret_cast<const float4*>(mrow + 64 * j + 32 * hf) + (i & 7));
          }
#pragma unroll
          for (int w2 = 0; w2 < 2; ++w2) {
            uint32_t bits = 0u;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 m4 = mb[8 * w2 + i];
              bits |= (m4.x > 0.f ? 1u : 0u) << (4 * i) | (m4.y > 0.f ? 1u : 0u) << (4 * i + 1) | (m4.z > 0.f ? 1u : 0u) << (4 * i + 2) |
                      (m4.w > 0.f ? 1u : 0u) << (4 * i + 3);
            }
            mbits[2 * hb + w2] = bits;
          }
        }
      }
      mbar_wait(smem_u32(sm.acc_full), it_acc & 1u, 600);
      ++it_acc;
      tc_fence_after();
      // ---- epilogue per 64-column slab: D -> (+bias) (relu) (mask) -> staging image -> whole rows (+C) -> global
      for (int j = 0; j < nsc; ++j) {
        const uint32_t mword = j == 0 ? mbits[0] : j == 1 ? mbits[1] : j == 2 ? mbits[2] : mbits[3];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c = 4 * j + 2 * hf + cc;                     // 16-column chunk of D
          uint32_t vr[16];
          tmem_ld16(tm_lane + dbuf + 16 * c, vr);
          tmem_wait_ld_fence16(vr);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float x[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              x[i] = __uint_as_float(vr[4 * q + i]);
              if (P.bias) x[i] += __ldg(&P.bias[16 * c + 4 * q + i]);
              if (P.relu) x[i] = fmaxf(x[i], 0.f);
              if (!((mword >> (16 * cc + 4 * q + i)) & 1u)) x[i] = 0.f;
            }
            *reinterpret_cast<float4*>(S + row * 256 + (((8 * hf + 4 * cc + q) ^ (row & 15)) << 4)) = make_float4(x[0], x[1], x[2], x[3]);
          }
        }
        named_bar_sync(1, kWorkers);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = (t + 256 * i) >> 4;
          const long long gr = trow0 + rr;
          if (gr < P.P) {
            float4 o = *reinterpret_cast<const float4*>(S + rr * 256 + (((t & 15) ^ (rr & 15)) << 4));
            float4* dst = reinterpret_cast<float4*>(P.C + gr * P.ldc + 64 * j) + (t & 15);
            if (P.accumulate) { const float4 old = *dst; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
            *dst = o;
          }
        }
        named_bar_sync(1, kWorkers);                             // the image is free again
      }
      tc_fence_before();
      named_bar_sync(1, kWorkers);          // every worker is done with D(it) before anyone refills that buffer as A(it+1)
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tm, kTmemCols);
}

int run_pack(const PackPlan& plan, const NetGeom* g, const float* params, void* packed, cudaStream_t st) {
  TcAux* aux = reinterpret_cast<TcAux*>(packed);
  uint8_t* img = reinterpret_cast<uint8_t*>(packed);
  NSOS_CHECK_CUDA(cudaMemsetAsync(aux->absmax_bits, 0, sizeof(aux->absmax_bits), st));
  if (g) k_fuse_views<<<g->W / 2, 128, 0, st>>>(params, reinterpret_cast<float*>(img + plan.fused_off), *g);
  k_pack_absmax<<<dim3(32, plan.nst, 2), 256, 0, st>>>(params, img, plan);
  if (g) k_pack_aux<<<16, 256, 0, st>>>(params, img, plan, *g);
  k_pack_slabs<<<dim3(plan.nslab, 8), 256, 0, st>>>(params, img, plan);
  NSOS_CHECK_CUDA(cudaGetLastError());
  return NSOS_OK;
}

int max_optin_smem() {
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return 0;
  return v;
}

}  // namespace

// ---- library-internal entry points ---------------------------------------------------------------------------
size_t tc_packed_bytes(const NsosNetDesc& net, int mode) {
  NetGeom g;
  if (!make_geom(net, g) || !tc_supported(g, nullptr)) return 0;
  TcProg p; PackPlan plan;
  build_prog(g, mode == NSOS_MODE_TC_EXACT, p, plan);
  return (size_t)plan.total_bytes;
}

int tc_pack_weights(const NsosNetDesc& net, const float* params, void* packed, int mode, cudaStream_t st) {
  NetGeom g;
  NSOS_REQUIRE(make_geom(net, g), NSOS_ERR_UNSUPPORTED, "invalid net descriptor");
  const char* why = "";
  NSOS_REQUIRE(tc_supported(g, &why), NSOS_ERR_UNSUPPORTED, "%s", why);
  TcProg p; PackPlan plan;
  build_prog(g, mode == NSOS_MODE_TC_EXACT, p, plan);
  return run_pack(plan, &g, params, packed, st);
}

size_t tc_render_workspace_bytes(const NsosRenderCfg&, int64_t) { return sizeof(long long) * kTraceTiles * 16 * kTraceStamps + 256; }

namespace {
struct ReplayIO {
  const float* z_in[2];
  float* dump_h[2];
  float* dump_s0[2];
  float* const* dump_all[2] = {nullptr, nullptr};   // [layer] per pass (all-parameter backward) or null
  float* dump_hv[2] = {nullptr, nullptr};
  const float* pts_in = nullptr;                    // point query (tc_mlp_query_dir): points, count, the one view direction
  long long n_pts = 0;
  float qdir[3] = {0.f, 0.f, 0.f};
};

// common launcher of k_render_tc: forward (replay == nullptr) or backward recompute on given sample depths
int tc_launch(const NsosRenderCfg& cfg, const void* packed_c, const void* packed_f, const float* rays_o, const float* rays_d,
              const float* near, const float* far, const NsosRandoms* rnd, uint64_t seed, const NsosRenderOut& out,
              const ReplayIO* replay, void* workspace, size_t workspace_bytes, int64_t n_rays, cudaStream_t st) {
  NetGeom gc, gf;
  NSOS_REQUIRE(make_geom(cfg.coarse, gc), NSOS_ERR_UNSUPPORTED, "invalid coarse net descriptor");
  const bool fine = cfg.n_importance > 0;
  if (fine) NSOS_REQUIRE(make_geom(cfg.fine, gf), NSOS_ERR_UNSUPPORTED, "invalid fine net descriptor"); else gf = gc;
  const char* why = "";
  NSOS_REQUIRE(tc_supported(gc, &why) && tc_supported(gf, &why), NSOS_ERR_UNSUPPORTED, "%s", why);
  NSOS_REQUIRE(gc.C == gf.C && gc.Lv == gf.Lv, NSOS_ERR_UNSUPPORTED, "coarse and fine nets must share output channels and multires_views");
  const int Sc = cfg.n_samples, K = cfg.n_importance, Sf = Sc + K;
  NSOS_REQUIRE(Sc >= 2 && Sc <= 128 && Sf <= kMaxS, NSOS_ERR_UNSUPPORTED, "tcgen05 path needs 2<=n_samples<=128 and n_samples+n_importance<=256");
  const bool exact = cfg.mode == NSOS_MODE_TC_EXACT;
  TcParams P;
  memset(&P, 0, sizeof(P));
  PackPlan plan;
  build_prog(gc, exact, P.prog[0], plan);
  build_prog(gf, exact, P.prog[1], plan);
  for (int i = 0; i < 2; ++i) P.nch[i] = build_ctab(P.prog[i], exact, P.ctab[i]);
  P.packed[0] = (const uint8_t*)packed_c; P.packed[1] = (const uint8_t*)packed_f;
  P.rays_o = rays_o; P.rays_d = rays_d; P.near = near; P.far = far;
  if (rnd) P.rnd = *rnd;
  P.seed = seed; P.out = out; P.n_rays = n_rays; P.Sc = Sc; P.K = K; P.Sf = fine ? Sf : Sc;
  P.perturb = cfg.perturb; P.noise_std = cfg.raw_noise_std; P.white_bkgd = cfg.white_bkgd; P.exact = exact; P.fine = fine;
  P.C = gc.C; P.sem_dim = gc.sem_dim; P.C6 = 6 + gc.sem_dim; P.ML = 2 * P.C6 + 1;
  if (replay) {
    for (int i = 0; i < 2; ++i) {
      P.z_in[i] = replay->z_in[i]; P.dump_h[i] = replay->dump_h[i]; P.dump_s0[i] = replay->dump_s0[i]; P.dump_hv[i] = replay->dump_hv[i];
      P.pts_in = replay->pts_in; P.n_pts = replay->n_pts;
      for (int k = 0; k < 3; ++k) P.qdir[k] = replay->qdir[k];
      if (replay->dump_all[i]) for (int l = 0; l < kMaxStages; ++l) P.dump_all[i][l] = replay->dump_all[i][l];
    }
  } else if (fine) {
    P.dump_h[0] = out.h_last0; P.dump_s0[0] = out.s_hid0; P.dump_h[1] = out.h_last; P.dump_s0[1] = out.s_hid;
    P.dump_enc[0] = out.enc0; P.dump_enc[1] = out.enc;
  } else {
    P.dump_h[0] = out.h_last; P.dump_s0[0] = out.s_hid; P.dump_enc[0] = out.enc;
  }
  const bool dump = !replay && (P.dump_h[0] || P.dump_h[1] || P.dump_s0[0] || P.dump_s0[1]);
  // what the semantic-head weight-gradient kernel reads is written in its blocked layout (the all-parameter replay feeds row GEMMs)
  if (!(replay && (replay->dump_all[0] || replay->dump_all[1]))) {
    P.dump_blocked[0] = P.dump_blocked[1] = sem_saves_blocked(gc, gf) ? 1 : 0;
  }
  const bool blk = P.dump_blocked[0] != 0;
  if (!replay && getenv("NSOS_TRACE") && workspace && workspace_bytes >= tc_render_workspace_bytes(cfg, n_rays)) {
    P.trace = reinterpret_cast<long long*>(workspace);
    NSOS_CHECK_CUDA(cudaMemsetAsync(workspace, 0, sizeof(long long) * kTraceTiles * 16 * kTraceStamps, st));
  }
  const int smem_max = max_optin_smem();
  NSOS_REQUIRE(smem_max >= 200 * 1024, NSOS_ERR_DEVICE, "device offers only %d B of opt-in shared memory", smem_max);
  // ring depth: three 32 KB slots sustain the MMA rate; deeper rings measured 1-2 % slower (profiles/r02_notes.md).  NSOS_NSLOTS overrides (experiments)
  int nslots = kDefaultSlots;
  if (const char* e = getenv("NSOS_NSLOTS")) nslots = std::max(2, std::min(kMaxSlots, atoi(e)));
  size_t need = 0;
  for (; nslots >= 2; --nslots) {
    need = carve_smem(nullptr, nslots, Sc, P.Sf, P.C, nullptr) + 1024;
    if ((int)need <= smem_max) break;
  }
  NSOS_REQUIRE(nslots >= 2, NSOS_ERR_UNSUPPORTED, "shared memory budget exceeded");
  P.nslots = nslots;
  int dev = 0, sms = 0;
  NSOS_CHECK_CUDA(cudaGetDevice(&dev));
  NSOS_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long n_pairs = (n_rays + 1) / 2;
  // Thread-block clusters can share the weight stream (multicast bulk copies): NSOS_CLUSTER = 1 (default), 2 or 4.
  // Measured on B200 (profiles/r01_notes.md): no gain -- the stream is bound by the per-thread bulk-copy issue rate and
  // the per-SM ingest (~96 B/clk), not by L2, and multicast does not lower the bytes landing in each SM.
  int csize = 1;
  if (const char* e = getenv("NSOS_CLUSTER")) csize = atoi(e);
  if (csize != 1 && csize != 2 && csize != 4) csize = 1;
  while (csize > 1 && n_pairs < csize) csize >>= 1;
  long long g = std::min<long long>((n_pairs + csize - 1) / csize * csize, (long long)(sms / csize) * csize);
  const int grid = (int)g;
  // with a fine pass every entry of a ray's row is written by the kernel (fine block, coarse block, z_std); coarse-only
  // renders leave the second block and z_std untouched, so the row is cleared first
  if (!replay && !fine) NSOS_CHECK_CUDA(cudaMemsetAsync(out.maps, 0, sizeof(float) * n_rays * P.ML, st));
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3(grid); lc.blockDim = dim3(kThreads); lc.dynamicSmemBytes = need; lc.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  lc.attrs = attr; lc.numAttrs = 1;
  auto launch = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
    if (e != cudaSuccess) return e;
    return cudaLaunchKernelEx(&lc, kern, P);
  };
  const bool replay_all = replay && (replay->dump_all[0] || replay->dump_all[1]);
  if (replay_all) NSOS_CHECK_CUDA(exact ? launch(k_render_tc<true, 3>) : launch(k_render_tc<false, 3>));
  else if (replay && blk) NSOS_CHECK_CUDA(exact ? launch(k_render_tc<true, 2, true>) : launch(k_render_tc<false, 2, true>));
  else if (replay) NSOS_CHECK_CUDA(exact ? launch(k_render_tc<true, 2>) : launch(k_render_tc<false, 2>));
  else if (dump && blk) NSOS_CHECK_CUDA(exact ? launch(k_render_tc<true, 1, true>) : launch(k_render_tc<false, 1, true>));
  else if (dump) NSOS_CHECK_CUDA(exact ? launch(k_render_tc<true, 1>) : launch(k_render_tc<false, 1>));
  else NSOS_CHECK_CUDA(exact ? launch(k_render_tc<true, 0>) : launch(k_render_tc<false, 0>));
  NSOS_CHECK_CUDA(cudaGetLastError());
  return NSOS_OK;
}
}  // namespace

int tc_render_fwd(const NsosRenderCfg& cfg, const float* /*pc*/, const float* /*pf*/, const void* packed_c, const void* packed_f,
                  const float* rays_o, const float* rays_d, const float* near, const float* far, const NsosRandoms* rnd,
                  uint64_t seed, const NsosRenderOut& out, void* workspace, size_t workspace_bytes, int64_t n_rays,
                  cudaStream_t st) {
  return tc_launch(cfg, packed_c, packed_f, rays_o, rays_d, near, far, rnd, seed, out, nullptr, workspace, workspace_bytes, n_rays, st);
}

bool tc_net_supported(const NsosNetDesc& net) {
  NetGeom g;
  return make_geom(net, g) && tc_supported(g, nullptr);
}

// Backward recompute: evaluates both nets at the SAVED sample depths (no sampling, no compositing) and writes, per point, the
// raw outputs plus the last trunk activation and the semantic hidden layer -- everything the semantic-head gradients need.
int tc_render_replay(const NsosRenderCfg& cfg, const void* packed_c, const void* packed_f, const float* rays_o, const float* rays_d,
                     const float* z0, const float* z1, float* raw0, float* raw1, float* h0, float* s00, float* h1, float* s01,
                     int64_t n_rays, cudaStream_t st) {
  ReplayIO io{{z0, z1}, {h0, h1}, {s00, s01}};
  NsosRenderOut out;
  memset(&out, 0, sizeof(out));
  const bool fine = cfg.n_importance > 0;
  if (fine) { out.raw0 = raw0; out.raw = raw1; } else { out.raw = raw0; }
  return tc_launch(cfg, packed_c, packed_f, rays_o, rays_d, nullptr, nullptr, nullptr, 0, out, &io, nullptr, 0, n_rays, st);
}

// NeRFMLP.forward (nerf_mlp.py:179-215) for points that share ONE view direction (export_density, eval.py:290-297: zeros) on the
// tensor cores: the replay mode of the render kernel with the tile rows read from `pts` instead of sampled along rays.
// `dir` (host, 3 floats) is used as given.  raw [n_pts, 4+sem_dim].
int tc_mlp_query_dir(const NsosNetDesc& net, const void* packed, const float* pts, const float* dir, float* raw, int mode, int64_t n_pts,
                     cudaStream_t st) {
  if (n_pts <= 0) return NSOS_OK;
  NsosRenderCfg cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.coarse = net; cfg.fine = net; cfg.n_samples = 64; cfg.n_importance = 0; cfg.mode = mode;
  ReplayIO io{{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
  io.pts_in = pts; io.n_pts = n_pts;
  for (int k = 0; k < 3; ++k) io.qdir[k] = dir[k];
  NsosRenderOut out;
  memset(&out, 0, sizeof(out));
  out.raw = raw;
  const int64_t n_rays = (n_pts + cfg.n_samples - 1) / cfg.n_samples;      // 64 consecutive points per "ray", two per tile
  return tc_launch(cfg, packed, packed, nullptr, nullptr, nullptr, nullptr, nullptr, 0, out, &io, nullptr, 0, n_rays, st);
}

// C[P,N] (=|+=) epi(A[P,K] . B),  B(k,n) = B[k*b_rs + n*b_cs].  scratch: >= tc_rowgemm_scratch_bytes(K, N).
size_t tc_rowgemm_scratch_bytes(int K, int N) { return (size_t)((K + 63) / 64) * 2 * N * 128 + 1024; }
bool tc_rowgemm_supported(int K, int N, int64_t lda, int64_t ldc, int64_t mask_ld) {
  return K >= 64 && K <= 256 && K % 64 == 0 && N >= 64 && N <= 256 && N % 64 == 0 && lda % 4 == 0 && ldc % 4 == 0 && mask_ld % 4 == 0;
}
int tc_rowgemm(const float* A, int64_t lda, int K, const float* B, int64_t b_rs, int64_t b_cs, float* C, int64_t ldc, int N,
               const float* mask, int64_t mask_ld, const float* bias, int relu, int accumulate, int64_t P, void* scratch,
               size_t scratch_bytes, cudaStream_t st) {
  NSOS_REQUIRE(tc_rowgemm_supported(K, N, lda, ldc, mask ? mask_ld : 0), NSOS_ERR_UNSUPPORTED, "tc_rowgemm: unsupported shape K=%d N=%d", K, N);
  NSOS_REQUIRE(scratch && scratch_bytes >= tc_rowgemm_scratch_bytes(K, N), NSOS_ERR_WORKSPACE, "tc_rowgemm: scratch too small");
  if (P <= 0) return NSOS_OK;
  RowGemmParams p;
  memset(&p, 0, sizeof(p));
  const int nsl = (K + 63) / 64;
  p.prog.nst = 1; p.prog.W = N;
  TcStage& S = p.prog.st[0];
  S.n = N; S.epi = EPI_RAW; S.nslab = nsl;
  for (int j = 0; j < nsl; ++j) { S.sn[j] = N; S.asrc[j] = j; }
  p.nch = build_ctab(p.prog, true, p.ctab);
  uint8_t* img = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(scratch) + 1023) / 1024 * 1024);
  k_pack_b_bf16<<<dim3(nsl, 8), 256, 0, st>>>(B, b_rs, b_cs, K, N, img);
  p.packed = img - kAuxBytes;          // producer_tile skips the aux header
  p.A = A; p.lda = lda; p.K = nsl * 64; p.C = C; p.ldc = ldc; p.N = N; p.mask = mask; p.mask_ld = mask_ld; p.bias = bias;
  p.accumulate = accumulate; p.relu = relu; p.P = P; p.nslots = kRgSlots;
  p.mask_rows = getenv("NSOS_RG_MASK_BITS") == nullptr;     // the ballot variant measured slower (2.18 vs 1.75 ms on 1.5 M rows): opt-in
  // a K that is not a multiple of 64 (e.g. 96) is zero-padded in B; A columns beyond K must not be read: require K % 64 == 0 there
  NSOS_REQUIRE(K % 64 == 0, NSOS_ERR_UNSUPPORTED, "tc_rowgemm: K must be a multiple of 64");
  int dev = 0, sms = 0;
  NSOS_CHECK_CUDA(cudaGetDevice(&dev));
  NSOS_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long ntiles = (P + 127) / 128;
  const int grid = (int)std::min<long long>(ntiles, sms);
  const size_t need = carve_rowgemm(nullptr, p.nslots, nullptr, nullptr) + 1024;
  NSOS_CHECK_CUDA(cudaFuncSetAttribute(k_rowgemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
  k_rowgemm<<<grid, kThreads, need, st>>>(p);
  NSOS_CHECK_CUDA(cudaGetLastError());
  return NSOS_OK;
}

// Recompute of the all-parameter backward: both nets at the saved sample depths, every hidden layer saved.
// h_all[pass][layer] -> [n_rays*S, W] (layer < D), hv / s0 [n_rays*S, W/2], raw [n_rays*S, C].  Entries of pass 1 are ignored
// without a fine pass.
int tc_render_replay_all(const NsosRenderCfg& cfg, const void* packed_c, const void* packed_f, const float* rays_o, const float* rays_d,
                         const float* z0, const float* z1, float* const raw[2], float* const* const h_all[2], float* const hv[2],
                         float* const s0[2], int64_t n_rays, cudaStream_t st) {
  NetGeom gc, gf;
  NSOS_REQUIRE(make_geom(cfg.coarse, gc), NSOS_ERR_UNSUPPORTED, "invalid coarse net descriptor");
  const bool fine = cfg.n_importance > 0;
  if (fine) NSOS_REQUIRE(make_geom(cfg.fine, gf), NSOS_ERR_UNSUPPORTED, "invalid fine net descriptor"); else gf = gc;
  ReplayIO io;
  io.z_in[0] = z0; io.z_in[1] = z1;
  for (int p = 0; p < 2; ++p) {
    const NetGeom& g = p ? gf : gc;
    io.dump_all[p] = h_all[p];
    io.dump_h[p] = h_all[p] ? h_all[p][g.D - 1] : nullptr;       // the last trunk layer goes through the EPI_HIDDEN_SIGMA save
    io.dump_s0[p] = s0[p]; io.dump_hv[p] = hv[p];
  }
  NsosRenderOut out;
  memset(&out, 0, sizeof(out));
  if (fine) { out.raw0 = raw[0]; out.raw = raw[1]; } else { out.raw = raw[0]; }
  return tc_launch(cfg, packed_c, packed_f, rays_o, rays_d, nullptr, nullptr, nullptr, 0, out, &io, nullptr, 0, n_rays, st);
}

int tc_selftest(const float* a, const float* w, float* d, int N, int K, int a_in_tmem, int mode, void* scratch, size_t scratch_bytes,
                cudaStream_t st) {
  NSOS_REQUIRE(N % 32 == 0 && N >= 32 && N <= 256, NSOS_ERR_BAD_ARG, "selftest: N must be a multiple of 32 in [32,256]");
  NSOS_REQUIRE(a_in_tmem ? (K % 64 == 0 && K >= 64 && K <= 256) : (K >= 1 && K <= 64), NSOS_ERR_BAD_ARG,
               "selftest: K must be 64..256 step 64 (TMEM A) or <=64 (SMEM A)");
  const bool exact = mode == NSOS_MODE_TC_EXACT;
  SelfParams P;
  memset(&P, 0, sizeof(P));
  PackPlan plan;
  memset(&plan, 0, sizeof(plan));
  P.prog.nst = 1; P.prog.W = N;
  TcStage& S = P.prog.st[0];
  S.n = N; S.epi = EPI_RAW; S.nslab = 0;
  int64_t off = (int64_t)kAuxBytes;
  const int nsl = a_in_tmem ? K / 64 : 1;
  for (int j = 0; j < nsl; ++j) {
    S.sn[S.nslab] = N;
    S.asrc[S.nslab++] = a_in_tmem ? j : A_GAMMA;
    PackSlab& ps = plan.s[plan.nslab++];
    ps.stage = 0; ps.n = N; ps.col0[0] = 64 * j; ps.kvalid[0] = a_in_tmem ? 64 : K; ps.col0[1] = 0; ps.kvalid[1] = 0;
    ps.dst_hi = off; off += (int64_t)N * 128;
    ps.dst_lo = -1;
    if (exact) { ps.dst_lo = off; off += (int64_t)N * 128; }
  }
  plan.nst = 1;
  plan.grp[0][0].src = 0; plan.grp[0][0].w_off = 0; plan.grp[0][0].ld = K; plan.grp[0][0].rows = N;
  plan.fused_off = off;
  plan.total_bytes = off;
  NSOS_REQUIRE(scratch_bytes >= (size_t)off, NSOS_ERR_WORKSPACE, "selftest scratch too small: need %lld", (long long)off);
  NSOS_CHECK_CUDA(cudaMemsetAsync(scratch, 0, kAuxBytes, st));
  int rc = run_pack(plan, nullptr, w, scratch, st);
  if (rc) return rc;
  // inv_scale for the single stage (k_pack_aux is skipped: no net geometry here)
  k_fix_scale<<<1, 1, 0, st>>>(reinterpret_cast<TcAux*>(scratch));
  P.nch = build_ctab(P.prog, exact, P.ctab);
  P.packed = (const uint8_t*)scratch; P.a = a; P.d = d; P.K = K; P.N = N; P.a_in_tmem = a_in_tmem; P.nslots = 4;
  size_t need = carve_smem(nullptr, P.nslots, 2, 2, 8, nullptr) + 1024;
  if (exact) {
    NSOS_CHECK_CUDA(cudaFuncSetAttribute(k_selftest<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    k_selftest<true><<<1, kThreads, need, st>>>(P);
  } else {
    NSOS_CHECK_CUDA(cudaFuncSetAttribute(k_selftest<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    k_selftest<false><<<1, kThreads, need, st>>>(P);
  }
  NSOS_CHECK_CUDA(cudaGetLastError());
  return NSOS_OK;
}

}  // namespace nsos
