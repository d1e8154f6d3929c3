// xtwx_kernel.cuh -- X' diag(w) [X | y] in one pass (see xtwx_kernel.cu)
#pragma once
#include "common.cuh"
#include <stddef.h>

namespace apb {
constexpr int XTWX_THREADS = 128;
enum XtwxMode {
    XTWX_GRAM = 0,         // w_i = weights or 1: X'X and X'y (model/apop_ols.c:333-334)
    XTWX_PROBIT = 1,       // w_i = lambda_i (lambda_i + x_i.beta): observed information of the probit LL
    XTWX_LOGIT2 = 2,       // w_i = p_i (1 - p_i)
    XTWX_POISSON_REG = 3,  // w_i = exp(x_i.beta)
    XTWX_MODE_COUNT = 4
};
struct XtwxParams {
    double beta[MAX_K_REG];
    double aux[4];
};
struct XtwxLaunch {
    const double* tiles;
    unsigned long long n_rows, n_tiles;
    int k, cols, has_w;
    double* partials;      // [grid][KB*(KB+1)]
    unsigned int* ticket;
    double* out;           // [KB*(KB+1)]: H[j][c] at j*(KB+1)+c, c = KB is the y column
    int grid;
    cudaStream_t stream;
    FinalizeCtx fin;
};
int xtwx_kb(int k);
cudaError_t xtwx_launch(int mode, const XtwxLaunch& L, const XtwxParams& prm);
int xtwx_blocks_per_sm(int mode, int k);
}  // namespace apb
