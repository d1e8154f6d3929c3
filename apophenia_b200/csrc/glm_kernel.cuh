// glm_kernel.cuh -- the fused  X.beta -> link/density -> LL + score  row reduction.
//
// One launch reads every resident tile exactly once (8(k+1) bytes per row, +8
// with weights) and produces [s_0..s_2 | g_0..g_{k-1}]:
//   * lane = row inside a 32-row tile; the k loads of a row are issued
//     back-to-back (k x 256 B per warp in flight) and stay in registers for the
//     score, so X is never re-read (the reference computes X.beta twice,
//     model/apop_probit.c:92 and :140);
//   * beta lives in the kernel parameter (constant bank), so x.beta is k DFMAs
//     with a constant operand and no shared memory;
//   * reduction: per-thread Neumaier sums for the scalar terms and plain FMA
//     accumulators for the score, then a shuffle tree per warp, a fixed-order
//     sum across the CTA's warps, per-CTA partials in global memory and a
//     fixed-order two-sum pass over the partials by the last CTA to finish
//     (ticket counter).  The result is bitwise reproducible for a given grid.
#pragma once
#include "families.cuh"

namespace apb {

struct GlmParams {
    double beta[GLM_K_REG];
    double aux[4];
};

struct GlmLaunch {
    const double* tiles;
    unsigned long long n_rows;
    unsigned long long n_tiles;
    double* partials;        // [grid][MAX_ACC + K]
    unsigned int* ticket;    // zero on entry, zero on exit
    double* out;             // [MAX_ACC + K]
    int grid;
    cudaStream_t stream;
    FinalizeCtx fin;
};

template <int K, class Fam, bool HAS_W>
__global__ void __launch_bounds__(GLM_THREADS)
glm_fused_kernel(const double* __restrict__ tiles, const unsigned long long n_rows,
                 const unsigned long long n_tiles, const __grid_constant__ GlmParams prm,
                 double* __restrict__ partials, unsigned int* __restrict__ ticket,
                 double* __restrict__ out, const __grid_constant__ FinalizeCtx fin) {
    constexpr int COLS = K + 1 + (HAS_W ? 1 : 0);
    constexpr int NOUT = MAX_ACC + K;
    constexpr int WARPS = GLM_THREADS / 32;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned long long gwarp = (unsigned long long)blockIdx.x * WARPS + warp;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * WARPS;

    double g[K];
#pragma unroll
    for (int j = 0; j < K; j++) g[j] = 0.0;
    KSum acc[MAX_ACC];

    for (unsigned long long t = gwarp; t < n_tiles; t += nwarps) {
        const double* base = tiles + t * (unsigned long long)(COLS * TILE_ROWS) + lane;
        double x[K];
#pragma unroll
        for (int j = 0; j < K; j++) x[j] = ld_stream(base + j * TILE_ROWS);
        const double y = ld_stream(base + K * TILE_ROWS);
        const double w = HAS_W ? ld_stream(base + (K + 1) * TILE_ROWS) : 1.0;

        double xb = 0.0;
#pragma unroll
        for (int j = 0; j < K; j++) xb = fma(x[j], prm.beta[j], xb);

        RowOut o = Fam::eval(xb, y, w, prm.aux);
        const bool valid = t * TILE_ROWS + lane < n_rows;  // last tile is zero padded
#pragma unroll
        for (int a = 0; a < Fam::NACC; a++) acc[a].add(valid ? o.s[a] : 0.0);
        const double d = valid ? o.d : 0.0;
#pragma unroll
        for (int j = 0; j < K; j++) g[j] = fma(x[j], d, g[j]);
    }

    // ---- CTA reduction -------------------------------------------------
    __shared__ double red[WARPS][NOUT];
#pragma unroll
    for (int a = 0; a < MAX_ACC; a++) {
        const double v = warp_sum(a < Fam::NACC ? acc[a].value() : 0.0);
        if (lane == 0) red[warp][a] = v;
    }
#pragma unroll
    for (int j = 0; j < K; j++) {
        const double v = warp_sum(g[j]);
        if (lane == 0) red[warp][MAX_ACC + j] = v;
    }
    __syncthreads();
    if (threadIdx.x < NOUT) {
        double s = 0.0;
#pragma unroll
        for (int wv = 0; wv < WARPS; wv++) s += red[wv][threadIdx.x];
        partials[(size_t)blockIdx.x * NOUT + threadIdx.x] = s;
    }

    // ---- grid reduction (+ cross-GPU sum over NVLink) by the last CTA --------
    grid_finalize(partials, NOUT, ticket, out, fin, &red[0][0]);
}

template <int K, class Fam, bool HAS_W>
cudaError_t glm_launch_one(const GlmLaunch& L, const GlmParams& prm) {
    glm_fused_kernel<K, Fam, HAS_W><<<L.grid, GLM_THREADS, 0, L.stream>>>(
        L.tiles, L.n_rows, L.n_tiles, prm, L.partials, L.ticket, L.out, L.fin);
    return cudaGetLastError();
}

template <int K, class Fam, bool HAS_W>
int glm_blocks_per_sm_one() {
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, glm_fused_kernel<K, Fam, HAS_W>, GLM_THREADS, 0);
    return nb;
}

// K-dispatch tables, instantiated once per family in glm_<family>.cu
template <class Fam, bool HAS_W, int K>
struct GlmTable {
    static cudaError_t launch(int k, const GlmLaunch& L, const GlmParams& prm) {
        if (k == K) return glm_launch_one<K, Fam, HAS_W>(L, prm);
        return GlmTable<Fam, HAS_W, K - 1>::launch(k, L, prm);
    }
    static int blocks_per_sm(int k) {
        if (k == K) return glm_blocks_per_sm_one<K, Fam, HAS_W>();
        return GlmTable<Fam, HAS_W, K - 1>::blocks_per_sm(k);
    }
};
template <class Fam, bool HAS_W>
struct GlmTable<Fam, HAS_W, 0> {
    static cudaError_t launch(int, const GlmLaunch&, const GlmParams&) { return cudaErrorInvalidValue; }
    static int blocks_per_sm(int) { return 0; }
};

// wide data (32 < k <= GLM_WIDE_MAX_K): glm_wide.cu; beta is a device pointer there
constexpr int GLM_WIDE_MAX_K = 1024;
cudaError_t glm_wide_launch(GlmFam fam, int k, int cols, bool has_w, const GlmLaunch& L, const double* beta, const double* aux);
int glm_wide_blocks_per_sm(GlmFam fam, int k);
int glm_wide_tiles_per_block(int k);   // tiles a CTA works on at a time (1: all its warps share a tile)

// defined in glm_<family>.cu
cudaError_t glm_launch(GlmFam fam, int k, bool has_w, const GlmLaunch& L, const GlmParams& prm);
int glm_blocks_per_sm(GlmFam fam, int k, bool has_w);

}  // namespace apb
