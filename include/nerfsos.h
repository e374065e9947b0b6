/* nerfsos.h -- C ABI of the B200-native NeRF-SOS render hot path (libnerfsos.so).
 *
 * The reference (VITA-Group/NeRF-SOS) has no FFI: its boundary for this path is the Python nn.Module
 * surface  models/nerf_net.py:132 (NeRFNet.forward) -> :71 (render_rays)  and the two loss modules
 * utils/image.py:263 (CorrelationLoss) / :373 (GeoCorrelationLoss).  The drop-in Python classes in
 * nerf-sos_b200/ keep those names and signatures and bind the entry points below through ctypes
 * (see INTEGRATION.md).  Every entry point cites the reference code it replaces.
 *
 * Conventions: plain C structs; raw DEVICE pointers (fp32 unless stated); the caller owns all memory,
 * provides the CUDA stream (cudaStream_t passed as void*) and the workspace; no allocation, no
 * synchronisation and no cudaSetDevice inside; re-entrant across streams.  Returns NSOS_OK (0) or an
 * NsosStatus; nsos_last_error() gives a thread-local message.
 */
#ifndef NERFSOS_H
#define NERFSOS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSOS_ABI_VERSION 1

typedef enum NsosStatus {
  NSOS_OK = 0,
  NSOS_ERR_BAD_ARG = 1,      /* null pointer / inconsistent sizes                                   */
  NSOS_ERR_UNSUPPORTED = 2,  /* configuration outside what the selected mode implements             */
  NSOS_ERR_CUDA = 3,         /* a CUDA runtime call or kernel launch failed                         */
  NSOS_ERR_WORKSPACE = 4,    /* workspace_bytes smaller than nsos_*_workspace_bytes()               */
  NSOS_ERR_DEVICE = 5        /* not an sm_100 device                                                */
} NsosStatus;

/* Arithmetic mode of the MLP contraction. */
typedef enum NsosMode {
  NSOS_MODE_SIMT_FP32 = 0, /* fp32 FMA on CUDA cores; any D/W; the same-device fp32 reference path        */
  NSOS_MODE_TC_EXACT = 1,  /* tcgen05 kind::f16, activations and weights split fp16 hi+lo, 3 MMAs/tile:
                              fp32-equivalent (meets the 1e-4 parity contract)                           */
  NSOS_MODE_TC_FAST = 2    /* tcgen05 kind::f16, single fp16 pass (throughput mode, judged on PSNR)       */
} NsosMode;

/* One NeRF MLP.  Mirrors the ctor of models/nerf_mlp.py:24-65 (MLP) + :132-177 (NeRFMLP). */
typedef struct NsosNetDesc {
  int32_t D;              /* netdepth                                                                    */
  int32_t W;              /* netwidth                                                                    */
  int32_t skip;           /* index i of the layer after which h=cat([enc,h]) (skips=[4]); <0: none.
                             Only takes effect when skip < D-1 (nerf_mlp.py:40-41, :73-74).               */
  int32_t multires;       /* L for points (10 -> 63 channels)                                            */
  int32_t multires_views; /* L for view dirs (4 -> 27 channels)                                          */
  int32_t use_viewdirs;   /* 1: alpha/feature/views/rgb heads; 0: output_linear (no semantics)           */
  int32_t use_semantics;  /* semantic_linear = Linear(sem_in, W/2) - ReLU - Linear(W/2, sem_dim)         */
  int32_t sem_dim;        /* number of semantic logits (2)                                               */
  int32_t sem_with_coord; /* sem_in = cat([h, enc]) (nerf_mlp.py:79)                                     */
} NsosNetDesc;

/* Render configuration.  Mirrors NeRFNet.__init__ / render_kwargs (models/nerf_net.py:22-69). */
typedef struct NsosRenderCfg {
  NsosNetDesc coarse;
  NsosNetDesc fine;      /* ignored when n_importance == 0 (nerf_fine is nerf, nerf_net.py:49)           */
  int32_t n_samples;     /* N_samples (64)                                                               */
  int32_t n_importance;  /* N_importance (128); 0 disables the fine pass                                 */
  float perturb;         /* >0: stratified jitter + random u (train); 0: deterministic (eval)            */
  float raw_noise_std;   /* >0: sigma noise N(0, std^2) (renderer.py:46-47)                              */
  int32_t white_bkgd;    /* renderer.py:77-81                                                            */
  int32_t mode;          /* NsosMode                                                                     */
} NsosRenderCfg;

/* Optional injected random draws (device pointers, each may be NULL).  Same four draws, same order
 * and shapes as the reference: sampler.py:61 rand [N,Sc]; renderer.py:47 randn [N,Sc];
 * sampler.py:103 rand [N,K]; renderer.py:47 randn [N,Sc+K].  NULL + perturb/noise>0 => in-kernel
 * Philox4x32-10 keyed by (seed, ray index, stream id, sample index). */
typedef struct NsosRandoms {
  const float* t_rand;
  const float* noise0;
  const float* u;
  const float* noise1;
  /* Optional stage-wise hook [N,K]: use these importance samples (sampler.py:132 `samples`) instead of inverting the
   * kernel's own cdf; inds / z_std are still computed from the kernel's cdf.  The inverse cdf is ill-conditioned in
   * low-mass bins (d sample / d cdf = bin width / denom, up to 1e4), so fine-pass parity -- outputs and gradients -- is
   * checked stage-wise on the reference's own samples, exactly like the index contract on the reference's own cdf. */
  const float* z_samples;
} NsosRandoms;

/* Outputs of nsos_render_fwd (device pointers).  `maps` is required, the rest may be NULL.
 * maps row layout, C6 = 6 + sem_dim (sem_dim = 0 without semantics), stride 2*C6+1:
 *   [0,3) rgb  [3] disp  [4] acc  [5] depth  [6,C6) semantics            -- fine (or only) pass
 *   [C6, 2*C6)  the same for the coarse pass ('0'-suffixed keys, nerf_net.py:127-128)
 *   [2*C6]      z_std (nerf_net.py:124)
 * With n_importance == 0 the first block holds the coarse pass and the rest is zero. */
typedef struct NsosRenderOut {
  float* maps;       /* [N, 2*C6+1]                  */
  float* weights0;   /* [N, Sc]                      */
  float* weights;    /* [N, Sc+K]                    */
  float* raw0;       /* [N, Sc, 4+sem_dim]           */
  float* raw;        /* [N, Sc+K, 4+sem_dim]         */
  float* z_vals0;    /* [N, Sc]                      */
  float* z_vals;     /* [N, Sc+K] sorted             */
  float* z_samples;  /* [N, K] (unsorted, as drawn)  */
  int64_t* inds;     /* [N, K] searchsorted indices  */
  /* Optional activations saved for nsos_render_bwd (tcgen05 modes, W=256; ignored by NSOS_MODE_SIMT_FP32):
   * last trunk activation relu(pts_linears[D-1]) and semantic hidden layer relu(semantic_linear.0), per point.
   * OPAQUE to the caller: hand the same pointers back to nsos_render_bwd.  The library writes them in the layout its
   * weight-gradient kernel contracts over -- groups of 32 consecutive points, feature-major inside a group: element
   * (pt, f) of an F-wide array at ((pt >> 5) * F + f) * 32 + (pt & 31) -- so every buffer must hold
   * ceil(points / 32) * 32 * F floats (points = N*Sc resp. N*(Sc+K)). */
  float* h_last0;    /* N*Sc points,     F = W       */
  float* s_hid0;     /* N*Sc points,     F = W/2     */
  float* h_last;     /* N*(Sc+K) points, F = W       */
  float* s_hid;      /* N*(Sc+K) points, F = W/2     */
  float* enc0;       /* N*Sc points,     F = 64: gamma(x) of the coarse points (63 columns + one zero), only needed with sem_with_coord: */
  float* enc;        /* N*(Sc+K) points, F = 64: saves the backward pass a separate positional-encoding kernel                          */
  /* Optional sticky status word (caller zero-initialises, reads and clears it; never reset by the library).
   * bit 0: NSOS_MODE_TC_EXACT/FAST only -- a hidden activation exceeded the fp16 range of the activation planes
   *        (|a| > 4094): the rgb / semantics maps of the affected rays are NaN.  Render such nets with NSOS_MODE_SIMT_FP32. */
  uint32_t* status;
} NsosRenderOut;

/* ---- introspection ------------------------------------------------------------------------- */
int nsos_abi_version(void);
const char* nsos_last_error(void);
/* Number of fp32 parameters of one net and, per tensor in state_dict order
 * (pts_linears.i.{weight,bias}, alpha_linear, feature_linear, views_linears.0, rgb_linear,
 *  semantic_linear.0, semantic_linear.2  |  output_linear when !use_viewdirs), its offset in the flat
 * buffer plus rows/cols ([out,in]; bias: rows=out, cols=1).  Returns the tensor count (<= cap
 * entries are written) or a negative NsosStatus. */
int64_t nsos_param_count(const NsosNetDesc* net);
int nsos_param_layout(const NsosNetDesc* net, int64_t* offsets, int32_t* rows, int32_t* cols, int cap);

/* ---- weight staging for the tcgen05 path ---------------------------------------------------- */
/* Replaces the implicit "weights live in nn.Linear" of nerf_mlp.py:40-64: converts the flat fp32
 * parameter buffer into the kernel's stream image (per-layer power-of-two scale, fp16 hi [+lo]
 * planes, 128B-swizzled K-major slabs in consumption order).  Must be re-run after the parameters
 * change.  nsos_packed_bytes returns 0 for NSOS_MODE_SIMT_FP32 or an unsupported net. */
size_t nsos_packed_bytes(const NsosNetDesc* net, int mode);
int nsos_pack_weights(const NsosNetDesc* net, const float* params, void* packed, int mode, void* stream);

/* ---- kernel A: fused hierarchical render ---------------------------------------------------- */
/* Replaces NeRFNet.forward's per-chunk body = render_rays (nerf_net.py:71-130): StratifiedSampler
 * (sampler.py:25-74), PositionEncoder (embedder.py:34-48), NeRFMLP/MLP (nerf_mlp.py:179-215,
 * 67-100), VolumetricRenderer (renderer.py:21-85), ImportanceSampler (sampler.py:91-170).
 * rays_o/rays_d [N,3] (d un-normalised; viewdirs = d/|d| is formed inside, nerf_net.py:163-166),
 * near/far [N].  packed_* may be NULL in NSOS_MODE_SIMT_FP32. */
size_t nsos_render_workspace_bytes(const NsosRenderCfg* cfg, int64_t n_rays);
int nsos_render_fwd(const NsosRenderCfg* cfg, const float* params_coarse, const float* params_fine,
                    const void* packed_coarse, const void* packed_fine, const float* rays_o,
                    const float* rays_d, const float* near, const float* far, const NsosRandoms* rnd,
                    uint64_t seed, const NsosRenderOut* out, void* workspace, size_t workspace_bytes,
                    int64_t n_rays, void* stream);

/* Backward of nsos_render_fwd (replaces autograd through the modules above, trainer.py:201).
 * g_maps [N, 2*C6+1] holds d(loss)/d(maps) (rgb, disp, acc, depth, semantics of both passes; the z_std column is
 * ignored: the importance samples are detached).  z_vals0/z_vals are the sample positions saved by the forward call (the
 * importance samples are detached, sampler.py:159, so the two passes are independent); the randoms /
 * seed must equal the forward call's.  Gradients are ACCUMULATED into grads_* (flat layout of
 * nsos_param_layout); trunk_grads=0 computes only semantic_linear.{0,2} (--fix_backbone,
 * run_nerf.py:307-318).  With trunk_grads=0 and a tcgen05 cfg->mode the trunk is recomputed on the tensor
 * cores from packed_* (the images nsos_render_fwd used; may be NULL in NSOS_MODE_SIMT_FP32) and only the
 * semantic-head weight gradients run on the tensor cores too (bf16 hi/lo, fp32 accumulate); otherwise everything
 * runs in fp32 on CUDA cores.  `saved` (may be NULL) points at the NsosRenderOut of the forward call: when it holds
 * raw0/raw and the h_last / s_hid arrays the trunk is not recomputed at all. */
size_t nsos_render_bwd_workspace_bytes(const NsosRenderCfg* cfg, int64_t n_rays, int trunk_grads);
int nsos_render_bwd(const NsosRenderCfg* cfg, const float* params_coarse, const float* params_fine,
                    const void* packed_coarse, const void* packed_fine,
                    const float* rays_o, const float* rays_d, const float* z_vals0, const float* z_vals,
                    const NsosRandoms* rnd, uint64_t seed, const float* g_maps, float* grads_coarse,
                    float* grads_fine, int trunk_grads, const NsosRenderOut* saved, void* workspace,
                    size_t workspace_bytes, int64_t n_rays, void* stream);

/* Stage-wise entry for the 'exact sample indices' contract: ImportanceSampler.sample_pdf
 * (sampler.py:117-132) on caller-supplied cdf.  bins/cdf [N,M], u [N,K] -> samples [N,K], inds [N,K]. */
int nsos_invert_cdf(const float* bins, const float* cdf, const float* u, float* samples, int64_t* inds,
                    int64_t n_rays, int32_t n_bins, int32_t n_u, void* stream);

/* Raw MLP query, replaces NeRFMLP.forward (nerf_mlp.py:179-215) as called by export_density
 * (engines/eval.py:297): pts [P,3], viewdirs [P,3] (already normalised) -> raw [P, 4+sem_dim]. */
size_t nsos_mlp_workspace_bytes(const NsosNetDesc* net, int64_t n_pts);
int nsos_mlp_query(const NsosNetDesc* net, const float* params, const float* pts, const float* viewdirs,
                   float* raw, void* workspace, size_t workspace_bytes, int64_t n_pts, void* stream);

/* The same query on the tensor cores for points that share ONE view direction -- export_density's case (engines/eval.py:290-297:
 * a regular grid, viewdirs = zeros): pts [P,3] (device), viewdir = HOST array of 3 floats used as given (not normalised),
 * packed = nsos_pack_weights image of the net for `mode` (NSOS_MODE_TC_EXACT / _FAST) -> raw [P, 4+sem_dim].  No workspace. */
int nsos_mlp_query_dir(const NsosNetDesc* net, const void* packed, const float* pts, const float* viewdir, float* raw,
                       int32_t mode, int64_t n_pts, void* stream);

/* ---- kernel B: patch-wise correlation losses ------------------------------------------------- */
/* GeoCorrelationLoss.forward (utils/image.py:448-482) without materialising the P^2 x P^2 pair
 * matrices.  xyz [B,3,M] (= ray_o + ray_d*depth after the caller's depth clip, :455/:443),
 * code [B,C,M], neg_idx [B] (argmin of the similarity matrix, :354), params = HOST array (self_shift,
 * self_weight, neg_shift, neg_weight) in the order of --geo_corr_params.  Writes loss[0] and, if g_code != NULL, d(loss)/d(code)
 * [B,C,M] (overwritten).  workspace: nsos_geo_corr_workspace_bytes(). */
size_t nsos_geo_corr_workspace_bytes(int32_t B, int32_t C, int32_t M);
int nsos_geo_corr_loss(const float* xyz, const float* code, const int64_t* neg_idx, const float* params,
                       float* loss, float* g_code, int32_t B, int32_t C, int32_t M, void* workspace,
                       size_t workspace_bytes, void* stream);
/* CorrelationLoss.forward (utils/image.py:335-370) on already-sampled tensors: feats [B,Cf,S],
 * nfeats [B,Cf,S] (negatives sampled at coords2), code/ncode [B,C,S] (S = 11*11), params = HOST array
 * as above (--app_corr_params); g_code/g_ncode [B,C,S] may both be NULL.  The bilinear
 * grid_sample (:303-304) stays in the caller (torch) so autograd routes g_code/g_ncode back. */
size_t nsos_app_corr_workspace_bytes(int32_t B, int32_t Cf, int32_t C, int32_t S);
int nsos_app_corr_loss(const float* feats, const float* nfeats, const float* code, const float* ncode,
                       const float* params, float* loss, float* g_code, float* g_ncode, int32_t B,
                       int32_t Cf, int32_t C, int32_t S, void* workspace, size_t workspace_bytes,
                       void* stream);

/* Sharded evaluation of the two losses (data-parallel training: every rank owns B_total/G whole patches; negatives may live on
 * another rank, utils/image.py:354,359,473).  phase 1 computes the row means of this rank's query patches and writes the partial
 * sums of the two helpers (negative, self) to sums[0..1]; the caller all-reduces them (one scalar pair per loss call -- the
 * batch-wide `old_mean` of image.py:316-319 / :420-424) and calls phase 2 with the SAME workspace, which evaluates this rank's
 * share of the loss (already divided by the GLOBAL pair count) and the code gradients.  Summing loss and g_code over the ranks
 * gives exactly the single-GPU result on the global batch.
 *   geometry loss:   xyz / code / neg_idx / g_code describe all B = B_total patches (gathered); queries are [q0, q0+nq)
 *   appearance loss: the sampled tensors describe this rank's B query patches; q0 / nq are ignored */
typedef struct NsosLossShard {
  int32_t q0, nq;
  int32_t B_total;
  int32_t phase;      /* 0: everything in one call; 1: row means + partial sums; 2: loss + gradients from the global sums */
  double* sums;       /* [2] DEVICE */
} NsosLossShard;
int nsos_geo_corr_loss_sharded(const float* xyz, const float* code, const int64_t* neg_idx, const float* params, float* loss,
                               float* g_code, int32_t B, int32_t C, int32_t M, const NsosLossShard* shard, void* workspace,
                               size_t workspace_bytes, void* stream);
int nsos_app_corr_loss_sharded(const float* feats, const float* nfeats, const float* code, const float* ncode, const float* params,
                               float* loss, float* g_code, float* g_ncode, int32_t B, int32_t Cf, int32_t C, int32_t S,
                               const NsosLossShard* shard, void* workspace, size_t workspace_bytes, void* stream);

/* ---- optimiser step ------------------------------------------------------------------------------ */
/* torch.optim.Adam(params, lr, betas=(0.9, 0.999)).step() of run_nerf.py:320 for a list of fp32 tensors in ONE launch, with the
 * learning rate of this step (engines/lr.py:20-23: lr * decay_rate ** (step / decay_steps), evaluated by the caller).  `tensors`
 * is a HOST array of device pointers; exp_avg / exp_avg_sq are the optimiser's state tensors (torch's state_dict names);
 * `step` counts from 1 (bias corrections 1 - beta^step). */
typedef struct NsosAdamTensor {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t n;
} NsosAdamTensor;
int nsos_adam_multi(const NsosAdamTensor* tensors, int32_t n_tensors, float lr, float beta1, float beta2, float eps, int64_t step,
                    void* stream);

/* ---- self test of the tcgen05 building blocks ------------------------------------------------ */
/* One CTA: D[128,N] = A[128,K] * W[N,K]^T through the same pack / bulk-copy / UMMA / TMEM code the
 * render kernel uses.  a [128,K], w [N,K] fp32 device; d [128,N] fp32 device.  a_in_tmem selects the
 * TMEM-A (1) or SMEM-A (0) operand form; mode is NSOS_MODE_TC_EXACT or NSOS_MODE_TC_FAST. */
int nsos_selftest_umma(const float* a, const float* w, float* d, int32_t N, int32_t K, int a_in_tmem,
                       int mode, void* scratch, size_t scratch_bytes, void* stream);

/* Self tests of the two tcgen05 GEMM kernels behind the all-parameter backward (bf16 hi/lo operands, 3 MMAs per product, fp32
 * accumulate), checked against fp64 matmuls in tests/test_gpu_render.py:
 *   rowgemm: C[P,N] (=|+=) epi(A[P,K] . B), B(k,n) = b[k*b_rs + n*b_cs]   (dgrad through a layer: B = W; feature_linear: B = W^T)
 *            K multiple of 64 <= 256, N multiple of 32 <= 256; epi = +bias, ReLU, keep where mask > 0; scratch >= 2*K*N*2 + 2048 B
 *   wgrad:   dW[Mo,.] += dY[P,Mo]^T . [main (256 wide) | aux (<= 64 wide)]   (contraction over points; Mo = 128 or 256);
 *            db (optional, aux_w <= 63): db[Mo] += column sums of dY (bias gradient) from a constant-one feature in the same launch;
 *            scratch >= nsos_selftest_wgrad_scratch_bytes(): per-CTA partial sums, reduced by a second launch (no atomics) */
int nsos_selftest_rowgemm(const float* a, int64_t lda, int32_t K, const float* b, int64_t b_rs, int64_t b_cs, float* c, int64_t ldc,
                          int32_t N, const float* mask, int64_t mask_ld, const float* bias, int relu, int accumulate, int64_t P,
                          void* scratch, size_t scratch_bytes, void* stream);
int nsos_selftest_wgrad(const float* dY, int64_t ldy, int32_t Mo, const float* main, int64_t ld_main, int32_t main_col, const float* aux,
                        int64_t ld_aux, int32_t aux_w, int32_t aux_col, float* dW, int64_t ldw, float* db, int64_t P, void* scratch,
                        size_t scratch_bytes, void* stream);
size_t nsos_selftest_wgrad_scratch_bytes(void);

#ifdef __cplusplus
}
#endif
#endif /* NERFSOS_H */
