// cells.cu -- apop_map_sum / apop_matrix_sum shaped reductions over every cell
// of a resident data set (apop_mapply.m4.c:485-517, apop_stats.m4.c:15,310).
//
// One pass returns, separately for the matrix cells and the vector cells,
//   s0 = sum f0(x)   s1 = sum f1(x)   and a count of cells mapped to -inf,
// which is what the iid models need: Normal wants sum (x-mu) and sum (x-mu)^2
// (model/apop_normal.c:38-50,108-118), Poisson wants sum x and sum lgamma(x+1)
// with its validity test (model/apop_poisson.c:14-28,66-73).
// Bound: HBM (8 bytes per cell, read once).
#include "cells.cuh"
#include "fastmath.cuh"

namespace apb {

// psi(x), x > 0: recurrence up to x >= 12, then the asymptotic series (the stand-in for gsl_sf_psi
// that oracle/oracle.c uses, here in double)
__device__ __forceinline__ double digamma_pos(double x) {
    double r = 0.0;
    while (x < 12.0) { r -= 1.0 / x; x += 1.0; }
    const double f = 1.0 / (x * x);
    return r + log(x) - 0.5 / x - f * (1.0 / 12 - f * (1.0 / 120 - f * (1.0 / 252 - f * (1.0 / 240 - f * (1.0 / 132 - f * (691.0 / 32760 - f / 12))))));
}

template <int MODE>
__device__ __forceinline__ void cell_eval(double x, double p0, double p1, double p2, double& f0, double& f1, int& bad) {
    if (MODE == CELLS_NORMAL) {
        const double d = x - p0;
        f0 = d;
        f1 = d * d;
    } else if (MODE == CELLS_POISSON) {
        // apply_me: -inf for negative or non-integer cells (fractional part > 1e-4)
        const bool invalid = (x < 0.0) || ((x - (double)(int)x) > 1e-4);
        bad += invalid ? 1 : 0;
        f0 = x;                                   // ln(lambda) * x summed as ln(lambda) * sum x
        f1 = (x == 0.0 || invalid) ? 0.0 : lgamma(x + 1.0);
    } else if (MODE == CELLS_IDENTITY) {
        f0 = x; f1 = 0.0;
    } else if (MODE == CELLS_GAMMA) {
        bad += (x == 0.0) ? 1 : 0;                 // here: the number of zero cells
        f0 = x;
        f1 = (x != 0.0) ? fm_log(x) : 0.0;
    } else if (MODE == CELLS_LOGNORMAL) {
        const double l = log(x);                   // ln 0 = -inf propagates, as in the reference
        f0 = l;
        f1 = (l - p0) * (l - p0);
    } else if (MODE == CELLS_LOG) {
        f0 = log(x); f1 = 0.0;
    } else if (MODE == CELLS_NONZERO) {
        f0 = (x != 0.0) ? 1.0 : 0.0; f1 = 0.0;
    } else if (MODE == CELLS_BETA) {
        const bool in = !(x < 0.0 || x > 1.0);     // betamap's test, NaN cells pass it as they do there
        bad += in ? 0 : 1;
        f0 = in ? log(x) : 0.0;
        f1 = in ? log(1.0 - x) : 0.0;
    } else if (MODE == CELLS_BETA_RAW) {
        f0 = log(x); f1 = log(1.0 - x);
    } else if (MODE == CELLS_T) {
        const double z = (x - p0) / p1;
        f0 = log(1.0 + z * z / p2); f1 = 0.0;
    } else if (MODE == CELLS_YULE) {
        f0 = ((x >= 1.0) ? lgamma(x) : 0.0) - lgamma(x + p0);
        f1 = -digamma_pos(x + p0);
    } else {  // CELLS_EXP
        f0 = exp(x); f1 = 0.0;
    }
}

template <int MODE>
__global__ void __launch_bounds__(CELLS_THREADS)
cells_kernel(const double* __restrict__ tiles, unsigned long long n_rows, unsigned long long n_tiles,
             int k, int ycol, int cols, double p0, double p1, double p2, double* __restrict__ partials,
             unsigned int* __restrict__ ticket, double* __restrict__ out, const __grid_constant__ FinalizeCtx fin) {
    constexpr int WARPS = CELLS_THREADS / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long gwarp = (unsigned long long)blockIdx.x * WARPS + warp;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * WARPS;
    KSum m0, m1, v0, v1;
    int bad = 0;
    for (unsigned long long t = gwarp; t < n_tiles; t += nwarps) {
        const double* base = tiles + t * (unsigned long long)cols * TILE_ROWS + lane;
        const bool valid = t * TILE_ROWS + lane < n_rows;
        // plain sums within the tile (k terms), one compensated add per tile: the FP64 pipe
        // must stay under 8 bytes per cell of HBM time
        double s0 = 0.0, s1 = 0.0;
        int b = 0;
        int j = 0;
        for (; j + 16 <= k; j += 16) {
            double x[16];
#pragma unroll
            for (int u = 0; u < 16; u++) x[u] = ld_stream(base + (j + u) * TILE_ROWS);
#pragma unroll
            for (int u = 0; u < 16; u++) {
                double f0, f1;
                cell_eval<MODE>(x[u], p0, p1, p2, f0, f1, b);
                s0 += f0; s1 += f1;
            }
        }
        for (; j + 4 <= k; j += 4) {
            double x[4];
#pragma unroll
            for (int u = 0; u < 4; u++) x[u] = ld_stream(base + (j + u) * TILE_ROWS);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                double f0, f1;
                cell_eval<MODE>(x[u], p0, p1, p2, f0, f1, b);
                s0 += f0; s1 += f1;
            }
        }
        for (; j < k; j++) {
            double f0, f1;
            cell_eval<MODE>(ld_stream(base + j * TILE_ROWS), p0, p1, p2, f0, f1, b);
            s0 += f0; s1 += f1;
        }
        if (valid) { m0.add(s0); m1.add(s1); bad += b; }
        if (ycol >= 0) {
            double f0, f1; int by = 0;
            cell_eval<MODE>(ld_stream(base + ycol * TILE_ROWS), p0, p1, p2, f0, f1, by);
            if (valid) { v0.add(f0); v1.add(f1); bad += by; }
        }
    }
    __shared__ double red[WARPS][CELLS_NOUT];
    double vals[CELLS_NOUT] = {m0.value(), m1.value(), v0.value(), v1.value(), (double)bad};
#pragma unroll
    for (int a = 0; a < CELLS_NOUT; a++) {
        const double v = warp_sum(vals[a]);
        if (lane == 0) red[warp][a] = v;
    }
    __syncthreads();
    if (threadIdx.x < CELLS_NOUT) {
        double s = 0.0;
#pragma unroll
        for (int wv = 0; wv < WARPS; wv++) s += red[wv][threadIdx.x];
        partials[(size_t)blockIdx.x * CELLS_NOUT + threadIdx.x] = s;
    }
    grid_finalize(partials, CELLS_NOUT, ticket, out, fin, &red[0][0]);
}

cudaError_t cells_launch(int mode, const double* tiles, size_t n_rows, size_t n_tiles, int k,
                         int ycol, int cols, double p0, double* partials, unsigned int* ticket,
                         double* out, int grid, cudaStream_t st, const FinalizeCtx& fin, double p1, double p2) {
#define APB_CELLS(M)                                                                          \
    case M:                                                                                   \
        cells_kernel<M><<<grid, CELLS_THREADS, 0, st>>>(tiles, n_rows, n_tiles, k, ycol,         \
                                                        cols, p0, p1, p2, partials, ticket, out, fin); \
        break;
    switch (mode) {
        APB_CELLS(CELLS_IDENTITY)
        APB_CELLS(CELLS_NORMAL)
        APB_CELLS(CELLS_POISSON)
        APB_CELLS(CELLS_LOG)
        APB_CELLS(CELLS_EXP)
        APB_CELLS(CELLS_GAMMA)
        APB_CELLS(CELLS_LOGNORMAL)
        APB_CELLS(CELLS_NONZERO)
        APB_CELLS(CELLS_BETA)
        APB_CELLS(CELLS_BETA_RAW)
        APB_CELLS(CELLS_T)
        APB_CELLS(CELLS_YULE)
    default: return cudaErrorInvalidValue;
    }
#undef APB_CELLS
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// Per-column sums over the rows: the O(N k) prep reductions either side of the path
// (SURVEY 8(f) rank 2).  mode 0: sum ln|x| over non-zero cells -- probit_prep's start point
// exp(mean ln|x_ij|) (model/apop_probit.c:65-80); mode 1: sum x; mode 2: sum x^2; mode 3: sum ln x
// (the column sums of dirichlet_dlog_likelihood, model/apop_dirichlet.c:28-38).
// ---------------------------------------------------------------------------------------
template <int KB>
__global__ void __launch_bounds__(CELLS_THREADS)
colsums_kernel(const double* __restrict__ tiles, unsigned long long n_rows, unsigned long long n_tiles, int K,
               int cols, int mode, double* __restrict__ partials, unsigned int* __restrict__ ticket,
               double* __restrict__ out, const __grid_constant__ FinalizeCtx fin) {
    constexpr int WARPS = CELLS_THREADS / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long gwarp = (unsigned long long)blockIdx.x * WARPS + warp;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * WARPS;
    double acc[KB];
#pragma unroll
    for (int j = 0; j < KB; j++) acc[j] = 0.0;
    for (unsigned long long t = gwarp; t < n_tiles; t += nwarps) {
        const double* base = tiles + t * (unsigned long long)cols * TILE_ROWS + lane;
        const bool valid = t * TILE_ROWS + lane < n_rows;
        double x[KB];
#pragma unroll
        for (int j = 0; j < KB; j++) x[j] = (j < K) ? ld_stream(base + j * TILE_ROWS) : 0.0;
#pragma unroll
        for (int j = 0; j < KB; j++) {
            double f;
            if (mode == 0) f = (x[j] != 0.0) ? fm_log(fabs(x[j])) : 0.0;
            else if (mode == 1) f = x[j];
            else if (mode == 2) f = x[j] * x[j];
            else f = (valid && j < K) ? log(x[j]) : 0.0;   // plain ln x (dirichlet); padding rows and columns hold zeros
            acc[j] += valid ? f : 0.0;
        }
    }
    __shared__ double red[WARPS][KB];
#pragma unroll
    for (int j = 0; j < KB; j++) {
        const double v = warp_sum(acc[j]);
        if (lane == 0) red[warp][j] = v;
    }
    __syncthreads();
    if (threadIdx.x < KB) {
        double s = 0.0;
#pragma unroll
        for (int wv = 0; wv < WARPS; wv++) s += red[wv][threadIdx.x];
        partials[(size_t)blockIdx.x * KB + threadIdx.x] = s;
    }
    grid_finalize(partials, KB, ticket, out, fin, &red[0][0]);
}

cudaError_t colsums_launch(const double* tiles, size_t n_rows, size_t n_tiles, int k, int cols, int mode,
                           double* partials, unsigned int* ticket, double* out, int grid, cudaStream_t st, const FinalizeCtx& fin) {
    if (k <= 8) colsums_kernel<8><<<grid, CELLS_THREADS, 0, st>>>(tiles, n_rows, n_tiles, k, cols, mode, partials, ticket, out, fin);
    else if (k <= 16) colsums_kernel<16><<<grid, CELLS_THREADS, 0, st>>>(tiles, n_rows, n_tiles, k, cols, mode, partials, ticket, out, fin);
    else if (k <= 32) colsums_kernel<32><<<grid, CELLS_THREADS, 0, st>>>(tiles, n_rows, n_tiles, k, cols, mode, partials, ticket, out, fin);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

// a device vector summed across ranks with the same tail (for host-side scalars)
__global__ void __launch_bounds__(128)
exchange_kernel(const double* __restrict__ vals, int n, unsigned int* __restrict__ ticket, double* __restrict__ out,
                const __grid_constant__ FinalizeCtx fin) {
    extern __shared__ double sc[];
    grid_finalize(vals, n, ticket, out, fin, sc);
}
cudaError_t exchange_launch(const double* vals, int n, unsigned int* ticket, double* out, cudaStream_t st, const FinalizeCtx& fin) {
    exchange_kernel<<<1, 128, n * sizeof(double), st>>>(vals, n, ticket, out, fin);
    return cudaGetLastError();
}

}  // namespace apb
