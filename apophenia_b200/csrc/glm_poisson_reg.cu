// glm_poisson_reg.cu -- instantiates the fused kernel for FamPoissonReg, K = 1..32.
#include "glm_kernel.cuh"
namespace apb {
cudaError_t glm_launch_poisson_reg(int k, bool has_w, const GlmLaunch& L, const GlmParams& prm) {
    if (FamPoissonReg::USES_W && has_w) return GlmTable<FamPoissonReg, FamPoissonReg::USES_W, MAX_K_REG>::launch(k, L, prm);
    return GlmTable<FamPoissonReg, false, MAX_K_REG>::launch(k, L, prm);
}
int glm_blocks_per_sm_poisson_reg(int k, bool has_w) {
    if (FamPoissonReg::USES_W && has_w) return GlmTable<FamPoissonReg, FamPoissonReg::USES_W, MAX_K_REG>::blocks_per_sm(k);
    return GlmTable<FamPoissonReg, false, MAX_K_REG>::blocks_per_sm(k);
}
}  // namespace apb
