// cells.cuh -- per-cell map/sum reductions (see cells.cu)
#pragma once
#include "common.cuh"
#include <stddef.h>

namespace apb {
enum CellsMode { CELLS_IDENTITY = 0, CELLS_NORMAL = 1, CELLS_POISSON = 2, CELLS_LOG = 3, CELLS_EXP = 4,
                 CELLS_GAMMA = 5,       // f0 = x, f1 = ln x over non-zero cells; slot 4 counts the ZERO cells
                 CELLS_LOGNORMAL = 6,   // f0 = ln x, f1 = (ln x - p0)^2
                 CELLS_NONZERO = 7,     // f0 = (x != 0)                                   (bernoulli: bernie_ll / nonzero)
                 CELLS_BETA = 8,        // f0 = ln x, f1 = ln(1-x) over cells in [0,1]; slot 4 counts the cells outside (betamap)
                 CELLS_BETA_RAW = 9,    // f0 = ln x, f1 = ln(1-x), no range test            (beta_dlog_likelihood)
                 CELLS_T = 10,          // f0 = ln(1 + z^2/p2), z = (x - p0)/p1             (one_t)
                 CELLS_YULE = 11 };     // f0 = (x>=1 ? lgamma x : 0) - lgamma(x+p0), f1 = -psi(x+p0)  (yule apply_me / dapply_me)
constexpr int CELLS_THREADS = 256;
constexpr int CELLS_NOUT = 5;  // matrix s0, s1; vector s0, s1; count of -inf cells
cudaError_t cells_launch(int mode, const double* tiles, size_t n_rows, size_t n_tiles, int k,
                         int ycol, int cols, double p0, double* partials, unsigned int* ticket,
                         double* out, int grid, cudaStream_t st, const FinalizeCtx& fin, double p1 = 0.0, double p2 = 0.0);
// per-column sums (mode 0: ln|x| over non-zero cells, 1: x, 2: x^2, 3: ln x); out has 32 slots
cudaError_t colsums_launch(const double* tiles, size_t n_rows, size_t n_tiles, int k, int cols, int mode,
                           double* partials, unsigned int* ticket, double* out, int grid, cudaStream_t st, const FinalizeCtx& fin);
// a device vector of n doubles summed across ranks by the fused tail alone
cudaError_t exchange_launch(const double* vals, int n, unsigned int* ticket, double* out, cudaStream_t st, const FinalizeCtx& fin);
}  // namespace apb
