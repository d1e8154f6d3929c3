// glm_wide.cu -- the fused LL + score reduction for data wider than the register-resident kernel
// covers (GLM_K_REG < k <= GLM_WIDE_MAX_K).  Same tile layout, same families, same reduction tail;
// what changes is that a row no longer fits in one thread's registers, so the lanes of a warp tile a
// (rows x columns) patch instead of lane = row.
//
// The warps of a CTA walk one tile RH rows at a time (the whole tile, RH = 32, up to k = 256; half a
// tile beyond).  Inside a warp LPC = RH/4 neighbouring lanes share a column, each reading 4 of its rows
// with one 32-byte load (LDG.256), so the lanes of a column cover one contiguous run of 8*RH bytes --
// whole 128-byte lines -- and every sector of the tile is fetched exactly once; one load instruction of
// the warp spans 32/LPC columns, a lane holds 8 such columns (32 doubles) and a warp 8 * 32/LPC
// consecutive columns.  A row's k values never sit in one thread:
//   * x.beta: 4 partial sums per lane, a butterfly reduce-scatter over the lanes of different columns
//     (3-4 shuffle-adds), then one shared-memory exchange between the warps (one barrier per step,
//     ping-pong buffers);
//   * EVERY warp evaluates the family for the RH rows of the step, lane and lane + RH on the same row
//     (redundant: ~100 issue slots against the 64 DFMA of the step, well under the HBM time of
//     RH x k x 8 bytes; all 32 lanes take part because Fam::eval votes warp-wide) and hands the four d_i
//     a lane needs back by shuffle;
//   * the score accumulators are 8 registers per lane (its columns, its rows), summed over the LPC lanes
//     of a column once, at the end.
// HBM traffic is 8(k+1) bytes per row, nothing is read twice, nothing is transposed through shared memory.
//
// History: round 1 had a cooperative lane = row kernel for k <= 256 (5.5-6.8 TB/s)
// and a two-sweep kernel beyond that re-read every tile from L2 and transposed it through shared memory
// (1.6-1.9 TB/s); a first lane = column version of this kernel, 32 different lines per load instruction
// with one 32-byte piece each, ran at 4.3-4.7 TB/s whatever k -- the L1 wavefront rate.
#include "glm_kernel.cuh"

namespace apb {

constexpr int COLS_NCJ = 8;          // columns per lane
__device__ __forceinline__ void ld_stream_v4(const double* p, double& a, double& b, double& c, double& d) {
    asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

template <int RH, bool BIG, class Fam>
__global__ void __launch_bounds__(BIG ? 512 : 256, BIG ? 1 : 2)
glm_cols_kernel(const double* __restrict__ tiles, const unsigned long long n_rows, const unsigned long long n_tiles,
                const int K, const int cols, const int has_w, const double* __restrict__ beta,
                const double4 aux4, double* __restrict__ partials, unsigned int* __restrict__ ticket,
                double* __restrict__ out, const __grid_constant__ FinalizeCtx fin) {
    static_assert(RH == 16 || RH == 32, "rows per step");
    constexpr int LPC = RH / 4;               // lanes per column
    constexpr int CPI = 32 / LPC;             // columns per load instruction
    constexpr int CPW = CPI * COLS_NCJ;       // columns per warp
    extern __shared__ double smem[];
    const int NW = blockDim.x >> 5;
    double* part = smem;                      // [2][NW][RH] partial dot products, ping-pong per step
    double* red = smem + 2 * NW * RH;         // MAX_ACC + K doubles for the reduction tail
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rq = lane & (LPC - 1), cg = lane / LPC;
    const double aux[4] = {aux4.x, aux4.y, aux4.z, aux4.w};
    const int NOUT = MAX_ACC + K;

    const int c0 = warp * CPW + cg;           // this lane's columns: c0 + j * CPI
    const int nown = (K - c0 + CPI - 1) / CPI;   // how many of them exist (may be <= 0 or > NCJ)
    double b[COLS_NCJ], g[COLS_NCJ];
#pragma unroll
    for (int j = 0; j < COLS_NCJ; j++) {
        b[j] = (j < nown) ? __ldg(beta + c0 + j * CPI) : 0.0;
        g[j] = 0.0;
    }
    // the row of the step whose x.beta this lane holds after the reduce-scatter; lanes that differ only in
    // the all-reduce bits hold the same one, the lane with those bits clear publishes it
    const int myrow = rq * 4 + ((lane & 16) ? 2 : 0) + ((lane & 8) ? 1 : 0);
    const bool publisher = (lane & (7 & ~(LPC - 1))) == 0;

    KSum acc[MAX_ACC];
    int pp = 0;
    for (unsigned long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const double* tb = tiles + t * (unsigned long long)cols * TILE_ROWS;
#pragma unroll 1
        for (int r0 = 0; r0 < TILE_ROWS; r0 += RH) {
            double x[COLS_NCJ][4];
            const double* cp = tb + (size_t)c0 * TILE_ROWS + r0 + rq * 4;
#pragma unroll
            for (int j = 0; j < COLS_NCJ; j++) {
                if (j < nown) ld_stream_v4(cp + (size_t)j * CPI * TILE_ROWS, x[j][0], x[j][1], x[j][2], x[j][3]);
                else x[j][0] = x[j][1] = x[j][2] = x[j][3] = 0.0;
            }
            // every lane takes part in the family evaluation (Fam::eval votes warp-wide): lane and lane + RH both
            // work on row lane of the step, the copy's sums are dropped
            const int er = lane & (RH - 1);
            const double y = ld_stream(tb + (size_t)K * TILE_ROWS + r0 + er);
            const double w = has_w ? ld_stream(tb + (size_t)(K + 1) * TILE_ROWS + r0 + er) : 1.0;
            // partial x.beta over this lane's columns, for its 4 rows
            double p[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                p[i] = x[0][i] * b[0];
#pragma unroll
                for (int j = 1; j < COLS_NCJ; j++) p[i] = fma(x[j][i], b[j], p[i]);
            }
            // sum over the lanes of different columns (lane bits log2(LPC)..4): 4 -> 2 -> 1 values, then all-reduce
            {
                const bool u16 = (lane & 16) != 0, u8 = (lane & 8) != 0;
                const double s0 = u16 ? p[0] : p[2], k0 = u16 ? p[2] : p[0];
                const double s1 = u16 ? p[1] : p[3], k1 = u16 ? p[3] : p[1];
                const double q0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
                const double q1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
                const double s2 = u8 ? q0 : q1, k2 = u8 ? q1 : q0;
                p[0] = k2 + __shfl_xor_sync(0xffffffffu, s2, 8);
                if (LPC <= 4) p[0] += __shfl_xor_sync(0xffffffffu, p[0], 4);
            }
            if (publisher) part[(pp * NW + warp) * RH + myrow] = p[0];
            __syncthreads();
            double xb = 0.0;
            for (int wv = 0; wv < NW; wv++) xb += part[(pp * NW + wv) * RH + er];   // fixed order: the same x.beta in every warp
            const RowOut o = Fam::eval(xb, y, w, aux);
            const bool valid = t * TILE_ROWS + r0 + er < n_rows;
            if (warp == 0) {
#pragma unroll
                for (int a = 0; a < Fam::NACC; a++) acc[a].add(valid && lane < RH ? o.s[a] : 0.0);
            }
            const double d = valid ? o.d : 0.0;
            pp ^= 1;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const double di = __shfl_sync(0xffffffffu, d, rq * 4 + i);
#pragma unroll
                for (int j = 0; j < COLS_NCJ; j++) g[j] = fma(x[j][i], di, g[j]);
            }
        }
    }
    // ---- CTA reduction: add the LPC row groups of each column
#pragma unroll
    for (int j = 0; j < COLS_NCJ; j++) {
        double v = g[j];
#pragma unroll
        for (int m = 1; m < LPC; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        if (rq == 0 && j < nown) red[MAX_ACC + c0 + j * CPI] = v;
    }
    if (warp == 0) {
#pragma unroll
        for (int a = 0; a < MAX_ACC; a++) {
            const double v = warp_sum(a < Fam::NACC ? acc[a].value() : 0.0);
            if (lane == 0) red[a] = v;
        }
    }
    __syncthreads();
    for (int v = threadIdx.x; v < NOUT; v += blockDim.x) partials[(size_t)blockIdx.x * NOUT + v] = red[v];
    grid_finalize(partials, NOUT, ticket, out, fin, red);
}
// rows per step and CTA shape: the whole tile (32 columns per warp, up to 8 warps) to k = 256; half a tile
// (64 columns per warp) beyond -- two CTAs of up to 8 warps per SM to k = 512, one CTA of up to 16 warps to 1024
static int cols_rh(int k) { return k <= 256 ? 32 : 16; }
static int cols_warps(int k, int rh) { const int cpw = 32 / (rh / 4) * COLS_NCJ; return (k + cpw - 1) / cpw; }
static size_t cols_smem(int k, int rh) { return ((size_t)2 * cols_warps(k, rh) * rh + MAX_ACC + k) * sizeof(double); }
template <int RH, bool BIG, class Fam>
static cudaError_t cols_launch_one(int k, int cols, bool has_w, const GlmLaunch& L, const double* beta, const double* aux) {
    glm_cols_kernel<RH, BIG, Fam><<<L.grid, cols_warps(k, RH) * 32, cols_smem(k, RH), L.stream>>>(L.tiles, L.n_rows, L.n_tiles, k, cols, has_w ? 1 : 0, beta,
                                                                                                make_double4(aux[0], aux[1], aux[2], aux[3]), L.partials,
                                                                                                L.ticket, L.out, L.fin);
    return cudaGetLastError();
}
template <int RH, bool BIG, class Fam>
static int cols_bps_one(int k) {
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, glm_cols_kernel<RH, BIG, Fam>, cols_warps(k, RH) * 32, cols_smem(k, RH));
    return nb;
}
template <class Fam>
static cudaError_t cols_launch_fam(int k, int cols, bool has_w, const GlmLaunch& L, const double* beta, const double* aux) {
    if (cols_rh(k) == 32) return cols_launch_one<32, false, Fam>(k, cols, has_w, L, beta, aux);
    if (cols_warps(k, 16) <= 8) return cols_launch_one<16, false, Fam>(k, cols, has_w, L, beta, aux);
    return cols_launch_one<16, true, Fam>(k, cols, has_w, L, beta, aux);
}
template <class Fam>
static int cols_bps_fam(int k) {
    if (cols_rh(k) == 32) return cols_bps_one<32, false, Fam>(k);
    if (cols_warps(k, 16) <= 8) return cols_bps_one<16, false, Fam>(k);
    return cols_bps_one<16, true, Fam>(k);
}

// a CTA works on one tile at a time
int glm_wide_tiles_per_block(int) { return 1; }

cudaError_t glm_wide_launch(GlmFam fam, int k, int cols, bool has_w, const GlmLaunch& L, const double* beta, const double* aux) {
    switch (fam) {
    case GLM_PROBIT: return cols_launch_fam<FamProbit>(k, cols, has_w, L, beta, aux);
    case GLM_LOGIT2: return cols_launch_fam<FamLogit2>(k, cols, has_w, L, beta, aux);
    case GLM_POISSON_REG: return cols_launch_fam<FamPoissonReg>(k, cols, has_w, L, beta, aux);
    case GLM_OLS: return cols_launch_fam<FamOls>(k, cols, has_w, L, beta, aux);
    default: return cudaErrorInvalidValue;
    }
}
int glm_wide_blocks_per_sm(GlmFam fam, int k) {
    switch (fam) {
    case GLM_PROBIT: return cols_bps_fam<FamProbit>(k);
    case GLM_LOGIT2: return cols_bps_fam<FamLogit2>(k);
    case GLM_POISSON_REG: return cols_bps_fam<FamPoissonReg>(k);
    case GLM_OLS: return cols_bps_fam<FamOls>(k);
    default: return 0;
    }
}

}  // namespace apb
