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
    int k, cols;           // cols = resident columns per tile (k + y [+ w, ignored])
    const unsigned char* counts = nullptr;   // resample multiplicities of one bootstrap replicate (NULL: every row once)
    int m_pad = 0, replicate = 0;
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
// logit_info.cu: observed information, LOGIT_INFO_NP category pairs (a <= b; a < 0 = padding) per launch;
// out = NP blocks of KB x KB (full, symmetric), block i = X' diag([a==b] p_a - p_a p_b) X
constexpr int LOGIT_INFO_THREADS = 128;
constexpr int LOGIT_INFO_NP = 4;
struct LogitInfoPairs { int a[LOGIT_INFO_NP], b[LOGIT_INFO_NP]; };
cudaError_t logit_info_launch(int c1, const LogitLaunch& L, const LogitParams& prm, const LogitInfoPairs& pairs);
int logit_info_blocks_per_sm(int k, int c1);
cudaError_t logit_launch(int c1, const LogitLaunch& L, const LogitParams& prm);
int logit_blocks_per_sm(int k, int c1);
}  // namespace apb
