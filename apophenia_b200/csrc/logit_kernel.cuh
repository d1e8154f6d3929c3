// logit_kernel.cuh -- multinomial logit fused kernel (see logit_kernel.cu)
#pragma once
#include "common.cuh"
#include <stddef.h>

namespace apb {
constexpr int LOGIT_THREADS = 128;
constexpr int LOGIT_MAX_C1 = 7;  // up to 8 outcome categories
struct LogitParams {
    double B[MAX_K_REG * LOGIT_MAX_C1];  // k x (C-1), row-major (parameters->matrix), zero padded to KB rows
    double cats[8];                      // the first C-1 sorted category values (factor page vector)
};
struct LogitLaunch {
    const double* tiles;
    unsigned long long n_rows, n_tiles;
    int k;
    double* partials;      // [grid][1 + KB*(C-1)]
    unsigned int* ticket;
    double* out;           // [1 + KB*(C-1)]: LL, then G[j][c] at 1 + j*(C-1) + c
    int grid;
    cudaStream_t stream;
    FinalizeCtx fin;
};
int logit_kb(int k);
// logit_hybrid.cu: the default kernel (lane = row X.B and soft-max, DMMA score, cp.async ring)
cudaError_t logit_hybrid_launch(int c1, const LogitLaunch& L, const LogitParams& prm);
int logit_hybrid_blocks_per_sm(int k, int c1);
// logit_mma.cu: both contractions as DMMA, lane = row soft-max in between
cudaError_t logit_mma_launch(int c1, const LogitLaunch& L, const LogitParams& prm);
int logit_mma_blocks_per_sm(int k, int c1);
cudaError_t logit_launch(int c1, const LogitLaunch& L, const LogitParams& prm);
int logit_blocks_per_sm(int k, int c1);
}  // namespace apb
