// glm_dispatch.cu -- family dispatch for the fused GLM kernels
#include "glm_kernel.cuh"
namespace apb {
#define APB_DECL(name)                                                                        \
    cudaError_t glm_launch_##name(int k, bool has_w, const GlmLaunch& L, const GlmParams& prm); \
    int glm_blocks_per_sm_##name(int k, bool has_w);
APB_DECL(probit)
APB_DECL(logit2)
APB_DECL(poisson_reg)
APB_DECL(ols)
#undef APB_DECL

cudaError_t glm_launch(GlmFam fam, int k, bool has_w, const GlmLaunch& L, const GlmParams& prm) {
    switch (fam) {
    case GLM_PROBIT: return glm_launch_probit(k, has_w, L, prm);
    case GLM_LOGIT2: return glm_launch_logit2(k, has_w, L, prm);
    case GLM_POISSON_REG: return glm_launch_poisson_reg(k, has_w, L, prm);
    case GLM_OLS: return glm_launch_ols(k, has_w, L, prm);
    default: return cudaErrorInvalidValue;
    }
}
int glm_blocks_per_sm(GlmFam fam, int k, bool has_w) {
    switch (fam) {
    case GLM_PROBIT: return glm_blocks_per_sm_probit(k, has_w);
    case GLM_LOGIT2: return glm_blocks_per_sm_logit2(k, has_w);
    case GLM_POISSON_REG: return glm_blocks_per_sm_poisson_reg(k, has_w);
    case GLM_OLS: return glm_blocks_per_sm_ols(k, has_w);
    default: return 0;
    }
}
}  // namespace apb
