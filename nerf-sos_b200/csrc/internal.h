// Library-internal interfaces between the translation units of libnerfsos.so.
#pragma once
#include "common.cuh"

namespace nsos {

// simt_render.cu
size_t simt_render_workspace_bytes(const NsosRenderCfg& cfg, int64_t n_rays);
int simt_render_fwd(const NsosRenderCfg& cfg, const float* pc, const float* pf, const float* rays_o, const float* rays_d,
                    const float* near, const float* far, const NsosRandoms* rnd, uint64_t seed, const NsosRenderOut& out,
                    void* workspace, size_t workspace_bytes, int64_t n_rays, cudaStream_t st);
size_t simt_render_bwd_workspace_bytes(const NsosRenderCfg& cfg, int64_t n_rays, int trunk);
int simt_render_bwd(const NsosRenderCfg& cfg, const float* pc, const float* pf, const float* rays_o, const float* rays_d,
                    const float* z_vals0, const float* z_vals, const NsosRandoms* rnd, uint64_t seed, const float* g_maps,
                    float* grads_c, float* grads_f, int trunk, const void* packed_c, const void* packed_f,
                    const NsosRenderOut* saved, void* workspace, size_t workspace_bytes, int64_t n_rays, cudaStream_t st);
int simt_invert_cdf(const float* bins, const float* cdf, const float* u, float* samples, int64_t* inds, int64_t n_rays, int M, int K,
                    cudaStream_t st);
size_t simt_mlp_workspace_bytes(const NetGeom& g, int64_t P);
int simt_mlp_query(const NetGeom& g, const float* prm, const float* pts, const float* viewdirs, float* raw, void* workspace,
                   size_t workspace_bytes, int64_t P, cudaStream_t st);

// tc_render.cu (tcgen05 / TMEM / bulk-copy path)
size_t tc_packed_bytes(const NsosNetDesc& net, int mode);
int tc_pack_weights(const NsosNetDesc& net, const float* params, void* packed, int mode, cudaStream_t st);
size_t tc_render_workspace_bytes(const NsosRenderCfg& cfg, int64_t n_rays);
int tc_render_fwd(const NsosRenderCfg& cfg, const float* pc, const float* pf, const void* packed_c, const void* packed_f,
                  const float* rays_o, const float* rays_d, const float* near, const float* far, const NsosRandoms* rnd,
                  uint64_t seed, const NsosRenderOut& out, void* workspace, size_t workspace_bytes, int64_t n_rays,
                  cudaStream_t st);
bool tc_net_supported(const NsosNetDesc& net);
int tc_render_replay(const NsosRenderCfg& cfg, const void* packed_c, const void* packed_f, const float* rays_o, const float* rays_d,
                     const float* z0, const float* z1, float* raw0, float* raw1, float* h0, float* s00, float* h1, float* s01,
                     int64_t n_rays, cudaStream_t st);
int tc_mlp_query_dir(const NsosNetDesc& net, const void* packed, const float* pts, const float* dir, float* raw, int mode, int64_t n_pts,
                     cudaStream_t st);
int tc_render_replay_all(const NsosRenderCfg& cfg, const void* packed_c, const void* packed_f, const float* rays_o, const float* rays_d,
                         const float* z0, const float* z1, float* const raw[2], float* const* const h_all[2], float* const hv[2],
                         float* const s0[2], int64_t n_rays, cudaStream_t st);
// row GEMM on tcgen05 (bf16 hi/lo, fp32 accumulate) for the all-parameter backward: C[P,N] (=|+=) epi(A[P,K] . B), B(k,n) = B[k*b_rs + n*b_cs]
size_t tc_rowgemm_scratch_bytes(int K, int N);
bool tc_rowgemm_supported(int K, int N, int64_t lda, int64_t ldc, int64_t mask_ld);
int tc_rowgemm(const float* A, int64_t lda, int K, const float* B, int64_t b_rs, int64_t b_cs, float* C, int64_t ldc, int N,
               const float* mask, int64_t mask_ld, const float* bias, int relu, int accumulate, int64_t P, void* scratch,
               size_t scratch_bytes, cudaStream_t st);
// tc_wgrad.cu (semantic-head weight gradients on tcgen05)
bool tc_sem_wgrad_supported(const NetGeom& g);
// true: the activations saved for the semantic-head backward (h_last, s_hid, gamma) use the blocked layout -- groups of 32
// consecutive points, feature-major inside a group: (pt, f) of an F-wide tensor at ((pt >> 5) * F + f) * 32 + (pt & 31);
// buffers hold ceil(points / 32) * 32 * F floats.  false: row-major [point][feature] (fp32 fallback of the weight gradients).
bool sem_saves_blocked(const NetGeom& coarse, const NetGeom& fine);   // one answer for both nets of a render call
// both weight-gradient kernels leave per-CTA partial sums in `part` (>= tc_wgrad_part_bytes()) and reduce them in a second launch
size_t tc_wgrad_part_bytes();
int tc_sem_wgrad(const NetGeom& g, const float* prm, float* grads, const float* h, const float* enc, int enc_ld, int enc_blocked,
                 const float* s0, const float* g_raw, int64_t P, void* part, size_t part_bytes, cudaStream_t st);
bool tc_wgrad_gen_supported(int Mo, int64_t ldy, int main_w, int64_t ld_main, int aux_w);
int tc_wgrad_gen(const float* dY, int64_t ldy, int Mo, const float* main, int64_t ld_main, int main_col, const float* aux, int64_t ld_aux,
                 int aux_w, int aux_col, float* dW, int64_t ldw, float* db, int64_t P, void* part, size_t part_bytes, cudaStream_t st);
int tc_selftest(const float* a, const float* w, float* d, int N, int K, int a_in_tmem, int mode, void* scratch, size_t scratch_bytes,
                cudaStream_t st);

// optim.cu
int adam_multi(const NsosAdamTensor* t, int n_tensors, float lr, float beta1, float beta2, float eps, int64_t step, cudaStream_t st);

// corr_loss.cu
size_t geo_corr_workspace_bytes(int B, int C, int M);
int geo_corr_loss(const float* xyz, const float* code, const int64_t* neg_idx, const float* params, float* loss, float* g_code,
                  int B, int C, int M, const NsosLossShard* sh, void* workspace, size_t workspace_bytes, cudaStream_t st);
size_t app_corr_workspace_bytes(int B, int Cf, int C, int S);
int app_corr_loss(const float* feats, const float* nfeats, const float* code, const float* ncode, const float* params, float* loss,
                  float* g_code, float* g_ncode, int B, int Cf, int C, int S, const NsosLossShard* sh, void* workspace,
                  size_t workspace_bytes, cudaStream_t st);

}  // namespace nsos
