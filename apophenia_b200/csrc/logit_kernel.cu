// logit_kernel.cu -- multinomial logit: entry points of the kernel in logit_hybrid.cu.
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

// (An all-DMMA variant, 1.14 ms against the hybrid's 1.12-1.165, and a TMA-streamed probit kernel, slower than
// the register-resident one, are kept out of the build under tools/experiments/.)
// Default: logit_stream.cu (swizzled resident tiles, one bulk copy per tile).  APOP_B200_LOGIT_KERNEL=hybrid selects
// the cp.async kernel of logit_hybrid.cu (plain tiles) for A/B measurements.
static bool use_hybrid() {
    static const int v = [] { const char* e = getenv("APOP_B200_LOGIT_KERNEL"); return (e && !strcmp(e, "hybrid")) ? 1 : 0; }();
    return v != 0;
}
cudaError_t logit_launch(int c1, const LogitLaunch& L, const LogitParams& prm) {
    return use_hybrid() ? logit_hybrid_launch(c1, L, prm) : logit_stream_launch(c1, L, prm);
}
int logit_blocks_per_sm(int k, int c1) { return use_hybrid() ? logit_hybrid_blocks_per_sm(k, c1) : logit_stream_blocks_per_sm(k, c1); }
bool logit_wants_swizzled() { return !use_hybrid(); }
int logit_warps_per_block() { return use_hybrid() ? LOGIT_THREADS / 32 : logit_stream_warps_per_block(); }

}  // namespace apb
