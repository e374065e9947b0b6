// Kernel B placeholder translation unit (filled in below by corr_loss kernels).
#include "internal.h"
namespace nsos {
size_t geo_corr_workspace_bytes(int, int, int) { return 0; }
int geo_corr_loss(const float*, const float*, const int64_t*, const float*, float*, float*, int, int, int, void*, size_t, cudaStream_t) {
  set_error("geo_corr_loss: not built yet"); return NSOS_ERR_UNSUPPORTED; }
size_t app_corr_workspace_bytes(int, int, int, int) { return 0; }
int app_corr_loss(const float*, const float*, const float*, const float*, const float*, float*, float*, float*, int, int, int, int, void*, size_t, cudaStream_t) {
  set_error("app_corr_loss: not built yet"); return NSOS_ERR_UNSUPPORTED; }
}
