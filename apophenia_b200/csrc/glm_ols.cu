// glm_ols.cu -- instantiates the fused kernel for FamOls, K = 1..32.
#include "glm_kernel.cuh"
// has_w: the resident tiles carry a weights column (tile stride k + 2).  Only OLS reads it; the other families
// skip it (the reference's probit / logit never look at d->weights), the load is dead code there.
namespace apb {
cudaError_t glm_launch_ols(int k, bool has_w, const GlmLaunch& L, const GlmParams& prm) {
    if (has_w) return GlmTable<FamOls, true, GLM_K_REG>::launch(k, L, prm);
    return GlmTable<FamOls, false, GLM_K_REG>::launch(k, L, prm);
}
int glm_blocks_per_sm_ols(int k, bool has_w) {
    if (has_w) return GlmTable<FamOls, true, GLM_K_REG>::blocks_per_sm(k);
    return GlmTable<FamOls, false, GLM_K_REG>::blocks_per_sm(k);
}
}  // namespace apb
