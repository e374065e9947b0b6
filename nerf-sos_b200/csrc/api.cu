// C-ABI entry points of libnerfsos.so (declared in include/nerfsos.h).
#include <stdarg.h>

#include "common.cuh"
#include "internal.h"

namespace nsos {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace nsos

using namespace nsos;

extern "C" {

int nsos_abi_version(void) { return NSOS_ABI_VERSION; }
const char* nsos_last_error(void) { return g_err; }

int64_t nsos_param_count(const NsosNetDesc* net) {
  NetGeom g;
  if (!net || !make_geom(*net, g)) { set_error("nsos_param_count: invalid net descriptor"); return -NSOS_ERR_BAD_ARG; }
  return g.n_params;
}

int nsos_param_layout(const NsosNetDesc* net, int64_t* offsets, int32_t* rows, int32_t* cols, int cap) {
  NetGeom g;
  if (!net || !make_geom(*net, g)) { set_error("nsos_param_layout: invalid net descriptor"); return -NSOS_ERR_BAD_ARG; }
  int n = 0;
  auto put = [&](int64_t off, int r, int c) {
    if (n < cap) { if (offsets) offsets[n] = off; if (rows) rows[n] = r; if (cols) cols[n] = c; }
    ++n;
  };
  for (int i = 0; i < g.D; ++i) { put(g.w_pts[i], g.W, g.in_pts[i]); put(g.b_pts[i], g.W, 1); }
  if (g.use_viewdirs) {
    put(g.w_alpha, 1, g.W); put(g.b_alpha, 1, 1);
    put(g.w_feat, g.W, g.W); put(g.b_feat, g.W, 1);
    put(g.w_views, g.W / 2, g.W + g.encv); put(g.b_views, g.W / 2, 1);
    put(g.w_rgb, 3, g.W / 2); put(g.b_rgb, 3, 1);
    if (g.use_sem) {
      put(g.w_s0, g.W / 2, g.sem_in); put(g.b_s0, g.W / 2, 1);
      put(g.w_s2, g.sem_dim, g.W / 2); put(g.b_s2, g.sem_dim, 1);
    }
  } else {
    put(g.w_out, 4, g.W); put(g.b_out, 4, 1);
  }
  return n;
}

size_t nsos_packed_bytes(const NsosNetDesc* net, int mode) {
  if (!net || mode == NSOS_MODE_SIMT_FP32) return 0;
  return tc_packed_bytes(*net, mode);
}

int nsos_pack_weights(const NsosNetDesc* net, const float* params, void* packed, int mode, void* stream) {
  NSOS_REQUIRE(net && params && packed, NSOS_ERR_BAD_ARG, "nsos_pack_weights: null argument");
  NSOS_REQUIRE(mode == NSOS_MODE_TC_EXACT || mode == NSOS_MODE_TC_FAST, NSOS_ERR_BAD_ARG, "nsos_pack_weights: mode %d has no packed image", mode);
  return tc_pack_weights(*net, params, packed, mode, (cudaStream_t)stream);
}

size_t nsos_render_workspace_bytes(const NsosRenderCfg* cfg, int64_t n_rays) {
  if (!cfg || n_rays <= 0) return 0;
  if (cfg->mode == NSOS_MODE_SIMT_FP32) return simt_render_workspace_bytes(*cfg, n_rays);
  return tc_render_workspace_bytes(*cfg, n_rays);
}

int nsos_render_fwd(const NsosRenderCfg* cfg, const float* params_coarse, const float* params_fine, const void* packed_coarse,
                    const void* packed_fine, const float* rays_o, const float* rays_d, const float* near, const float* far,
                    const NsosRandoms* rnd, uint64_t seed, const NsosRenderOut* out, void* workspace, size_t workspace_bytes,
                    int64_t n_rays, void* stream) {
  NSOS_REQUIRE(cfg && params_coarse && rays_o && rays_d && near && far && out && out->maps, NSOS_ERR_BAD_ARG,
               "nsos_render_fwd: null argument");
  NSOS_REQUIRE(n_rays >= 0, NSOS_ERR_BAD_ARG, "nsos_render_fwd: negative ray count");
  if (n_rays == 0) return NSOS_OK;
  if (cfg->n_importance > 0) NSOS_REQUIRE(params_fine, NSOS_ERR_BAD_ARG, "nsos_render_fwd: fine parameters missing");
  else params_fine = params_coarse;
  if (cfg->mode == NSOS_MODE_SIMT_FP32)
    return simt_render_fwd(*cfg, params_coarse, params_fine, rays_o, rays_d, near, far, rnd, seed, *out, workspace, workspace_bytes,
                           n_rays, (cudaStream_t)stream);
  NSOS_REQUIRE(cfg->mode == NSOS_MODE_TC_EXACT || cfg->mode == NSOS_MODE_TC_FAST, NSOS_ERR_BAD_ARG, "unknown mode %d", cfg->mode);
  NSOS_REQUIRE(packed_coarse && (cfg->n_importance == 0 || packed_fine), NSOS_ERR_BAD_ARG,
               "nsos_render_fwd: tcgen05 modes need packed weights (nsos_pack_weights)");
  if (cfg->n_importance == 0) packed_fine = packed_coarse;
  return tc_render_fwd(*cfg, params_coarse, params_fine, packed_coarse, packed_fine, rays_o, rays_d, near, far, rnd, seed, *out,
                       workspace, workspace_bytes, n_rays, (cudaStream_t)stream);
}

size_t nsos_render_bwd_workspace_bytes(const NsosRenderCfg* cfg, int64_t n_rays, int trunk_grads) {
  if (!cfg || n_rays <= 0) return 0;
  return simt_render_bwd_workspace_bytes(*cfg, n_rays, trunk_grads);
}

int nsos_render_bwd(const NsosRenderCfg* cfg, const float* params_coarse, const float* params_fine, const void* packed_coarse,
                    const void* packed_fine, const float* rays_o,
                    const float* rays_d, const float* z_vals0, const float* z_vals, const NsosRandoms* rnd, uint64_t seed,
                    const float* g_maps, float* grads_coarse, float* grads_fine, int trunk_grads, const NsosRenderOut* saved,
                    void* workspace, size_t workspace_bytes, int64_t n_rays, void* stream) {
  NSOS_REQUIRE(cfg && params_coarse && rays_o && rays_d && z_vals && g_maps && grads_coarse, NSOS_ERR_BAD_ARG,
               "nsos_render_bwd: null argument");
  if (n_rays <= 0) return NSOS_OK;
  if (cfg->n_importance > 0) NSOS_REQUIRE(params_fine && grads_fine && z_vals0, NSOS_ERR_BAD_ARG, "nsos_render_bwd: fine-pass argument missing");
  else { params_fine = params_coarse; grads_fine = grads_coarse; }
  return simt_render_bwd(*cfg, params_coarse, params_fine, rays_o, rays_d, z_vals0, z_vals, rnd, seed, g_maps, grads_coarse,
                         grads_fine, trunk_grads, packed_coarse, packed_fine, saved, workspace, workspace_bytes, n_rays,
                         (cudaStream_t)stream);
}

int nsos_invert_cdf(const float* bins, const float* cdf, const float* u, float* samples, int64_t* inds, int64_t n_rays,
                    int32_t n_bins, int32_t n_u, void* stream) {
  NSOS_REQUIRE(bins && cdf && u && samples && inds, NSOS_ERR_BAD_ARG, "nsos_invert_cdf: null argument");
  NSOS_REQUIRE(n_bins >= 1 && n_u >= 1, NSOS_ERR_BAD_ARG, "nsos_invert_cdf: bad sizes");
  if (n_rays <= 0) return NSOS_OK;
  return simt_invert_cdf(bins, cdf, u, samples, inds, n_rays, n_bins, n_u, (cudaStream_t)stream);
}

size_t nsos_mlp_workspace_bytes(const NsosNetDesc* net, int64_t n_pts) {
  NetGeom g;
  if (!net || n_pts <= 0 || !make_geom(*net, g)) return 0;
  return simt_mlp_workspace_bytes(g, n_pts);
}

int nsos_mlp_query(const NsosNetDesc* net, const float* params, const float* pts, const float* viewdirs, float* raw,
                   void* workspace, size_t workspace_bytes, int64_t n_pts, void* stream) {
  NetGeom g;
  NSOS_REQUIRE(net && make_geom(*net, g), NSOS_ERR_UNSUPPORTED, "nsos_mlp_query: invalid net descriptor");
  NSOS_REQUIRE(params && pts && raw && (viewdirs || !g.use_viewdirs), NSOS_ERR_BAD_ARG, "nsos_mlp_query: null argument");
  if (n_pts <= 0) return NSOS_OK;
  return simt_mlp_query(g, params, pts, viewdirs, raw, workspace, workspace_bytes, n_pts, (cudaStream_t)stream);
}

size_t nsos_geo_corr_workspace_bytes(int32_t B, int32_t C, int32_t M) { return geo_corr_workspace_bytes(B, C, M); }
int nsos_geo_corr_loss(const float* xyz, const float* code, const int64_t* neg_idx, const float* params, float* loss, float* g_code,
                       int32_t B, int32_t C, int32_t M, void* workspace, size_t workspace_bytes, void* stream) {
  NSOS_REQUIRE(xyz && code && neg_idx && params && loss, NSOS_ERR_BAD_ARG, "nsos_geo_corr_loss: null argument");
  return geo_corr_loss(xyz, code, neg_idx, params, loss, g_code, B, C, M, nullptr, workspace, workspace_bytes, (cudaStream_t)stream);
}
int nsos_geo_corr_loss_sharded(const float* xyz, const float* code, const int64_t* neg_idx, const float* params, float* loss,
                               float* g_code, int32_t B, int32_t C, int32_t M, const NsosLossShard* shard, void* workspace,
                               size_t workspace_bytes, void* stream) {
  NSOS_REQUIRE(xyz && code && neg_idx && params && shard, NSOS_ERR_BAD_ARG, "nsos_geo_corr_loss_sharded: null argument");
  return geo_corr_loss(xyz, code, neg_idx, params, loss, g_code, B, C, M, shard, workspace, workspace_bytes, (cudaStream_t)stream);
}
size_t nsos_app_corr_workspace_bytes(int32_t B, int32_t Cf, int32_t C, int32_t S) { return app_corr_workspace_bytes(B, Cf, C, S); }
int nsos_app_corr_loss(const float* feats, const float* nfeats, const float* code, const float* ncode, const float* params,
                       float* loss, float* g_code, float* g_ncode, int32_t B, int32_t Cf, int32_t C, int32_t S, void* workspace,
                       size_t workspace_bytes, void* stream) {
  NSOS_REQUIRE(feats && nfeats && code && ncode && params && loss, NSOS_ERR_BAD_ARG, "nsos_app_corr_loss: null argument");
  return app_corr_loss(feats, nfeats, code, ncode, params, loss, g_code, g_ncode, B, Cf, C, S, nullptr, workspace, workspace_bytes,
                       (cudaStream_t)stream);
}
int nsos_app_corr_loss_sharded(const float* feats, const float* nfeats, const float* code, const float* ncode, const float* params,
                               float* loss, float* g_code, float* g_ncode, int32_t B, int32_t Cf, int32_t C, int32_t S,
                               const NsosLossShard* shard, void* workspace, size_t workspace_bytes, void* stream) {
  NSOS_REQUIRE(feats && nfeats && code && ncode && params && shard, NSOS_ERR_BAD_ARG, "nsos_app_corr_loss_sharded: null argument");
  return app_corr_loss(feats, nfeats, code, ncode, params, loss, g_code, g_ncode, B, Cf, C, S, shard, workspace, workspace_bytes,
                       (cudaStream_t)stream);
}

int nsos_adam_multi(const NsosAdamTensor* tensors, int32_t n_tensors, float lr, float beta1, float beta2, float eps, int64_t step,
                    void* stream) {
  NSOS_REQUIRE(tensors || n_tensors == 0, NSOS_ERR_BAD_ARG, "nsos_adam_multi: null tensor list");
  if (n_tensors <= 0) return NSOS_OK;
  return adam_multi(tensors, n_tensors, lr, beta1, beta2, eps, step, (cudaStream_t)stream);
}

int nsos_mlp_query_dir(const NsosNetDesc* net, const void* packed, const float* pts, const float* viewdir, float* raw, int32_t mode,
                       int64_t n_pts, void* stream) {
  NSOS_REQUIRE(net && packed && pts && viewdir && raw, NSOS_ERR_BAD_ARG, "nsos_mlp_query_dir: null argument");
  NSOS_REQUIRE(mode == NSOS_MODE_TC_EXACT || mode == NSOS_MODE_TC_FAST, NSOS_ERR_UNSUPPORTED, "nsos_mlp_query_dir: tcgen05 modes only");
  NSOS_REQUIRE(tc_net_supported(*net), NSOS_ERR_UNSUPPORTED, "nsos_mlp_query_dir: net geometry not supported by the tcgen05 path");
  NSOS_REQUIRE(net->use_viewdirs, NSOS_ERR_UNSUPPORTED, "nsos_mlp_query_dir: the net takes no view direction");
  return tc_mlp_query_dir(*net, packed, pts, viewdir, raw, mode, n_pts, (cudaStream_t)stream);
}

int nsos_selftest_umma(const float* a, const float* w, float* d, int32_t N, int32_t K, int a_in_tmem, int mode, void* scratch,
                       size_t scratch_bytes, void* stream) {
  NSOS_REQUIRE(a && w && d && scratch, NSOS_ERR_BAD_ARG, "nsos_selftest_umma: null argument");
  return tc_selftest(a, w, d, N, K, a_in_tmem, mode, scratch, scratch_bytes, (cudaStream_t)stream);
}

int nsos_selftest_rowgemm(const float* a, int64_t lda, int32_t K, const float* b, int64_t b_rs, int64_t b_cs, float* c, int64_t ldc,
                          int32_t N, const float* mask, int64_t mask_ld, const float* bias, int relu, int accumulate, int64_t P,
                          void* scratch, size_t scratch_bytes, void* stream) {
  NSOS_REQUIRE(a && b && c && scratch, NSOS_ERR_BAD_ARG, "nsos_selftest_rowgemm: null argument");
  return tc_rowgemm(a, lda, K, b, b_rs, b_cs, c, ldc, N, mask, mask_ld, bias, relu, accumulate, P, scratch, scratch_bytes, (cudaStream_t)stream);
}
int nsos_selftest_wgrad(const float* dY, int64_t ldy, int32_t Mo, const float* main, int64_t ld_main, int32_t main_col, const float* aux,
                        int64_t ld_aux, int32_t aux_w, int32_t aux_col, float* dW, int64_t ldw, float* db, int64_t P, void* scratch,
                        size_t scratch_bytes, void* stream) {
  NSOS_REQUIRE(dY && dW && (main || aux) && scratch, NSOS_ERR_BAD_ARG, "nsos_selftest_wgrad: null argument");
  return tc_wgrad_gen(dY, ldy, Mo, main, ld_main, main_col, aux, ld_aux, aux_w, aux_col, dW, ldw, db, P, scratch, scratch_bytes,
                      (cudaStream_t)stream);
}
size_t nsos_selftest_wgrad_scratch_bytes(void) { return tc_wgrad_part_bytes();
}

}  // extern "C"
