// glm_ols.cu -- instantiates the fused kernel for FamOls, K = 1..32.
#include "glm_kernel.cuh"
namespace apb {
cudaError_t glm_launch_ols(int k, bool has_w, const GlmLaunch& L, const GlmParams& prm) {
    if (FamOls::USES_W && has_w) return GlmTable<FamOls, FamOls::USES_W, MAX_K_REG>::launch(k, L, prm);
    return GlmTable<FamOls, false, MAX_K_REG>::launch(k, L, prm);
}
int glm_blocks_per_sm_ols(int k, bool has_w) {
    if (FamOls::USES_W && has_w) return GlmTable<FamOls, FamOls::USES_W, MAX_K_REG>::blocks_per_sm(k);
    return GlmTable<FamOls, false, MAX_K_REG>::blocks_per_sm(k);
}
}  // namespace apb
