// logit_kernel.cu -- multinomial logit: dispatch between the kernels of logit_hybrid.cu (default)
// and logit_mma.cu.
//
// LL row: one_logit_row (model/apop_probit.c:212-229): num - (max + log(sum_c
// exp(xb_c - max) + exp(-max))), the numeraire category carrying beta_0 = 0.
// Score (the reference keeps it commented out, :244-289; new vtable entry):
//   dLL/dB[j][c] = sum_i x_ij (1{idx_i == c+1} - p_ic),  output index j*(C-1)+c.
//
// Two earlier kernels lived here in the first half of round 1 and were removed once the hybrid
// kernel replaced them: an FP64-pipe kernel (phase B lane = column through a shared-memory slab,
// 1.40 ms on 20M x 32, C = 8: shared-memory bound at 254 wavefronts per tile) and an all-DMMA kernel
// with 8 warps per SM (1.40 ms: latency bound).  profiles/r01_ncu_full_logit8_4Mx32_* is the former.
#include "logit_kernel.cuh"
#include "fastmath.cuh"
#include <stdlib.h>
#include <string.h>

namespace apb {

int logit_kb(int k) { return k <= 8 ? 8 : k <= 16 ? 16 : k <= 24 ? 24 : 32; }

// APOP_B200_LOGIT_VARIANT=hybrid (default, logit_hybrid.cu) | mma (logit_mma.cu), kept for A/B runs
static int logit_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("APOP_B200_LOGIT_VARIANT");
        v = (e && !strcmp(e, "mma")) ? 1 : 0;
    }
    return v;
}
cudaError_t logit_launch(int c1, const LogitLaunch& L, const LogitParams& prm) {
    return logit_variant() == 1 ? logit_mma_launch(c1, L, prm) : logit_hybrid_launch(c1, L, prm);
}
int logit_blocks_per_sm(int k, int c1) {
    return logit_variant() == 1 ? logit_mma_blocks_per_sm(k, c1) : logit_hybrid_blocks_per_sm(k, c1);
}

}  // namespace apb
