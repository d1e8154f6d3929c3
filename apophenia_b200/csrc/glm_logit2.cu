// glm_logit2.cu -- instantiates the fused kernel for FamLogit2, K = 1..32.
#include "glm_kernel.cuh"
namespace apb {
cudaError_t glm_launch_logit2(int k, bool has_w, const GlmLaunch& L, const GlmParams& prm) {
    if (FamLogit2::USES_W && has_w) return GlmTable<FamLogit2, FamLogit2::USES_W, MAX_K_REG>::launch(k, L, prm);
    return GlmTable<FamLogit2, false, MAX_K_REG>::launch(k, L, prm);
}
int glm_blocks_per_sm_logit2(int k, bool has_w) {
    if (FamLogit2::USES_W && has_w) return GlmTable<FamLogit2, FamLogit2::USES_W, MAX_K_REG>::blocks_per_sm(k);
    return GlmTable<FamLogit2, false, MAX_K_REG>::blocks_per_sm(k);
}
}  // namespace apb
