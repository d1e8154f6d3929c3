// glm_probit.cu -- instantiates the fused kernel for FamProbit, K = 1..32.
#include "glm_kernel.cuh"
namespace apb {
cudaError_t glm_launch_probit(int k, bool has_w, const GlmLaunch& L, const GlmParams& prm) {
    if (FamProbit::USES_W && has_w) return GlmTable<FamProbit, FamProbit::USES_W, MAX_K_REG>::launch(k, L, prm);
    return GlmTable<FamProbit, false, MAX_K_REG>::launch(k, L, prm);
}
int glm_blocks_per_sm_probit(int k, bool has_w) {
    if (FamProbit::USES_W && has_w) return GlmTable<FamProbit, FamProbit::USES_W, MAX_K_REG>::blocks_per_sm(k);
    return GlmTable<FamProbit, false, MAX_K_REG>::blocks_per_sm(k);
}
}  // namespace apb
