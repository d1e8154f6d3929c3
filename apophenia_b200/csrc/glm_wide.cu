// glm_wide.cu -- the fused LL + score reduction for data wider than the register-resident
// kernels cover (32 < k <= GLM_WIDE_MAX_K).  Same tile layout, same families, same reduction
// tail; what changes is that a row no longer fits in registers:
//   sweep 1  lane = row streams the tile's columns once (HBM, coalesced) accumulating x.beta
//            (beta from global memory, broadcast through L1), then evaluates the family;
//   sweep 2  re-reads the tile in chunks of 32 columns (L1/L2 hits), parks each chunk
//            transposed in the warp's shared-memory slab and lets lane = column add
//            sum_i x_ij d_i into the warp's accumulator row in shared memory.
// HBM traffic stays 8(k+1) bytes per row; the second touch of each tile is served by L2.
//
// For 32 < k <= 256 a cooperative variant keeps everything in registers instead: the NW = ceil(k/32)
// warps of a CTA work on the SAME tile, warp w owning ceil(k/NW) consecutive columns exactly like the
// register-resident kernel owns its 32 (lane = row, coalesced 256-byte loads, x kept for the score).
// The partial dot products meet in shared memory (one barrier per tile, ping-pong buffers), every warp
// then evaluates the family on the full x.beta -- NW-fold redundant FP64 work that stays far below the
// HBM time of a (k+1) x 256-byte tile -- and adds its own columns' score terms.  No re-read, no transpose.
#include "glm_kernel.cuh"

namespace apb {

constexpr int WIDE_XS = 33;   // row stride of the parked 32 x 32 chunk (odd: conflict-free both ways)

template <class Fam>
__global__ void __launch_bounds__(GLM_THREADS)
glm_wide_kernel(const double* __restrict__ tiles, const unsigned long long n_rows, const unsigned long long n_tiles,
                const int K, const int cols, const int has_w, const double* __restrict__ beta,
                const double4 aux4, double* __restrict__ partials, unsigned int* __restrict__ ticket,
                double* __restrict__ out, const __grid_constant__ FinalizeCtx fin) {
    constexpr int WARPS = GLM_THREADS / 32;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Kp = (K + 31) / 32 * 32;             // accumulator row padded to whole chunks
    double* sx = smem + warp * (32 * WIDE_XS + 32 + Kp);
    double* sd = sx + 32 * WIDE_XS;
    double* accg = sd + 32;
    const unsigned long long gwarp = (unsigned long long)blockIdx.x * WARPS + warp;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * WARPS;
    const double aux[4] = {aux4.x, aux4.y, aux4.z, aux4.w};
    const int NOUT = MAX_ACC + K;

    for (int j = lane; j < Kp; j += 32) accg[j] = 0.0;
    KSum acc[MAX_ACC];
    __syncwarp();

    for (unsigned long long t = gwarp; t < n_tiles; t += nwarps) {
        const double* base = tiles + t * (unsigned long long)cols * TILE_ROWS + lane;
        // ---- sweep 1: x.beta, 8 columns in flight at a time
        double xb = 0.0;
        int j = 0;
        for (; j + 8 <= K; j += 8) {
            double x[8];
#pragma unroll
            for (int u = 0; u < 8; u++) x[u] = __ldg(base + (j + u) * TILE_ROWS);
#pragma unroll
            for (int u = 0; u < 8; u++) xb = fma(x[u], __ldg(beta + j + u), xb);
        }
        for (; j < K; j++) xb = fma(__ldg(base + j * TILE_ROWS), __ldg(beta + j), xb);
        const double y = __ldg(base + K * TILE_ROWS);
        const double w = has_w ? __ldg(base + (K + 1) * TILE_ROWS) : 1.0;
        RowOut o = Fam::eval(xb, y, w, aux);
        const bool valid = t * TILE_ROWS + lane < n_rows;
#pragma unroll
        for (int a = 0; a < Fam::NACC; a++) acc[a].add(valid ? o.s[a] : 0.0);
        __syncwarp();
        sd[lane] = valid ? o.d : 0.0;
        // ---- sweep 2: score, one 32-column chunk at a time
        for (int c0 = 0; c0 < K; c0 += 32) {
#pragma unroll 8
            for (int u = 0; u < 32; u++) sx[lane * WIDE_XS + u] = (c0 + u < K) ? __ldg(base + (c0 + u) * TILE_ROWS) : 0.0;
            __syncwarp();
            double gsum = 0.0;   // lane = column c0 + lane
#pragma unroll 8
            for (int i = 0; i < 32; i += 2) {
                const double2 d2 = *reinterpret_cast<const double2*>(sd + i);
                gsum = fma(sx[i * WIDE_XS + lane], d2.x, gsum);
                gsum = fma(sx[(i + 1) * WIDE_XS + lane], d2.y, gsum);
            }
            accg[c0 + lane] += gsum;
            __syncwarp();
        }
    }

    // ---- CTA reduction
    __syncthreads();
    double* red = smem + WARPS * (32 * WIDE_XS + 32 + Kp);   // NOUT doubles
#pragma unroll
    for (int a = 0; a < MAX_ACC; a++) {
        const double v = warp_sum(a < Fam::NACC ? acc[a].value() : 0.0);
        if (lane == 0) sd[a] = v;   // per-warp slot (sd is free now)
    }
    __syncthreads();
    for (int v = threadIdx.x; v < NOUT; v += GLM_THREADS) {
        double s = 0.0;
        for (int wv = 0; wv < WARPS; wv++) {
            const double* wb = smem + wv * (32 * WIDE_XS + 32 + Kp);
            s += (v < MAX_ACC) ? wb[32 * WIDE_XS + v] : wb[32 * WIDE_XS + 32 + (v - MAX_ACC)];
        }
        partials[(size_t)blockIdx.x * NOUT + v] = s;
    }
    __syncthreads();
    grid_finalize(partials, NOUT, ticket, out, fin, red);
}

template <int NW, class Fam>
__global__ void __launch_bounds__(NW * 32, (12 / NW > 0) ? 12 / NW : 1)   // about 12 warps per SM, as the k = 32 kernel runs
glm_coop_kernel(const double* __restrict__ tiles, const unsigned long long n_rows, const unsigned long long n_tiles,
                const int K, const int cols, const int has_w, const double* __restrict__ beta,
                const double4 aux4, double* __restrict__ partials, unsigned int* __restrict__ ticket,
                double* __restrict__ out, const __grid_constant__ FinalizeCtx fin) {
    extern __shared__ double smem[];
    double* sbeta = smem;                     // NW*32 doubles, zero beyond K
    double* part = smem + NW * 32;            // [2][NW][32] partial dot products, ping-pong per tile
    double* red = part + 2 * NW * 32;         // MAX_ACC + NW*32 doubles for the reduction tail
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kw = (K + NW - 1) / NW;         // columns per warp, balanced (k = 40: 20 + 20, not 32 + 8)
    const int c0 = warp * kw;                 // first column of this warp
    const double aux[4] = {aux4.x, aux4.y, aux4.z, aux4.w};
    const int NOUT = MAX_ACC + K;
    for (int j = threadIdx.x; j < NW * 32; j += NW * 32) sbeta[j] = (j < K) ? beta[j] : 0.0;
    __syncthreads();

    double g[32];
#pragma unroll
    for (int j = 0; j < 32; j++) g[j] = 0.0;
    KSum acc[MAX_ACC];
    int pp = 0;
    for (unsigned long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const double* base = tiles + t * (unsigned long long)cols * TILE_ROWS + lane;
        double x[32];
#pragma unroll
        for (int j = 0; j < 32; j++) x[j] = (j < kw && c0 + j < K) ? ld_stream(base + (c0 + j) * TILE_ROWS) : 0.0;   // warp-uniform predicate
        const double y = ld_stream(base + K * TILE_ROWS);
        const double w = has_w ? ld_stream(base + (K + 1) * TILE_ROWS) : 1.0;
        double xbp = 0.0;
#pragma unroll
        for (int j = 0; j < 32; j++) xbp = fma(x[j], (j < kw) ? sbeta[c0 + j] : 0.0, xbp);
        part[(pp * NW + warp) * 32 + lane] = xbp;
        __syncthreads();
        double xb = 0.0;
#pragma unroll
        for (int wv = 0; wv < NW; wv++) xb += part[(pp * NW + wv) * 32 + lane];   // fixed order: the same x.beta in every warp
        pp ^= 1;
        RowOut o = Fam::eval(xb, y, w, aux);
        const bool valid = t * TILE_ROWS + lane < n_rows;
        if (warp == 0) {
#pragma unroll
            for (int a = 0; a < Fam::NACC; a++) acc[a].add(valid ? o.s[a] : 0.0);
        }
        const double d = valid ? o.d : 0.0;
#pragma unroll
        for (int j = 0; j < 32; j++) g[j] = fma(x[j], d, g[j]);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j++) {
        const double v = warp_sum(g[j]);
        if (lane == 0 && j < kw && c0 + j < K) red[MAX_ACC + c0 + j] = v;
    }
    if (warp == 0) {
#pragma unroll
        for (int a = 0; a < MAX_ACC; a++) {
            const double v = warp_sum(a < Fam::NACC ? acc[a].value() : 0.0);
            if (lane == 0) red[a] = v;
        }
    }
    __syncthreads();
    for (int v = threadIdx.x; v < NOUT; v += NW * 32) partials[(size_t)blockIdx.x * NOUT + v] = red[v];
    grid_finalize(partials, NOUT, ticket, out, fin, red);
}
constexpr int COOP_MAX_K = 256;
static size_t coop_smem(int nw) { return ((size_t)nw * 32 + 2 * nw * 32 + MAX_ACC + nw * 32) * sizeof(double); }
template <int NW, class Fam>
static cudaError_t coop_launch_one(int k, int cols, bool has_w, const GlmLaunch& L, const double* beta, const double* aux) {
    glm_coop_kernel<NW, Fam><<<L.grid, NW * 32, coop_smem(NW), L.stream>>>(L.tiles, L.n_rows, L.n_tiles, k, cols, has_w ? 1 : 0, beta,
                                                                          make_double4(aux[0], aux[1], aux[2], aux[3]), L.partials,
                                                                          L.ticket, L.out, L.fin);
    return cudaGetLastError();
}
template <int NW, class Fam>
static int coop_bps_one() {
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, glm_coop_kernel<NW, Fam>, NW * 32, coop_smem(NW));
    return nb;
}
template <class Fam>
static cudaError_t coop_launch_fam(int k, int cols, bool has_w, const GlmLaunch& L, const double* beta, const double* aux) {
    switch ((k + 31) / 32) {
    case 2: return coop_launch_one<2, Fam>(k, cols, has_w, L, beta, aux);
    case 3: return coop_launch_one<3, Fam>(k, cols, has_w, L, beta, aux);
    case 4: return coop_launch_one<4, Fam>(k, cols, has_w, L, beta, aux);
    case 5: return coop_launch_one<5, Fam>(k, cols, has_w, L, beta, aux);
    case 6: return coop_launch_one<6, Fam>(k, cols, has_w, L, beta, aux);
    case 7: return coop_launch_one<7, Fam>(k, cols, has_w, L, beta, aux);
    case 8: return coop_launch_one<8, Fam>(k, cols, has_w, L, beta, aux);
    }
    return cudaErrorInvalidValue;
}
template <class Fam>
static int coop_bps_fam(int k) {
    switch ((k + 31) / 32) {
    case 2: return coop_bps_one<2, Fam>();
    case 3: return coop_bps_one<3, Fam>();
    case 4: return coop_bps_one<4, Fam>();
    case 5: return coop_bps_one<5, Fam>();
    case 6: return coop_bps_one<6, Fam>();
    case 7: return coop_bps_one<7, Fam>();
    case 8: return coop_bps_one<8, Fam>();
    }
    return 0;
}
// tiles a CTA works on at a time: the cooperative kernel walks the tiles block by block
int glm_wide_tiles_per_block(int k) { return (k <= COOP_MAX_K) ? 1 : GLM_THREADS / 32; }

static size_t wide_smem(int K) {
    const int Kp = (K + 31) / 32 * 32;
    return ((size_t)(GLM_THREADS / 32) * (32 * WIDE_XS + 32 + Kp) + MAX_ACC + K) * sizeof(double);
}

template <class Fam>
static cudaError_t wide_launch_fam(int k, int cols, bool has_w, const GlmLaunch& L, const double* beta, const double* aux) {
    const size_t sm = wide_smem(k);
    cudaFuncSetAttribute(glm_wide_kernel<Fam>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    glm_wide_kernel<Fam><<<L.grid, GLM_THREADS, sm, L.stream>>>(L.tiles, L.n_rows, L.n_tiles, k, cols, has_w ? 1 : 0, beta,
                                                              make_double4(aux[0], aux[1], aux[2], aux[3]), L.partials,
                                                              L.ticket, L.out, L.fin);
    return cudaGetLastError();
}
template <class Fam>
static int wide_bps_fam(int k) {
    const size_t sm = wide_smem(k);
    cudaFuncSetAttribute(glm_wide_kernel<Fam>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, glm_wide_kernel<Fam>, GLM_THREADS, sm);
    return nb;
}

cudaError_t glm_wide_launch(GlmFam fam, int k, int cols, bool has_w, const GlmLaunch& L, const double* beta, const double* aux) {
    if (k <= COOP_MAX_K) {
        switch (fam) {
        case GLM_PROBIT: return coop_launch_fam<FamProbit>(k, cols, has_w, L, beta, aux);
        case GLM_LOGIT2: return coop_launch_fam<FamLogit2>(k, cols, has_w, L, beta, aux);
        case GLM_POISSON_REG: return coop_launch_fam<FamPoissonReg>(k, cols, has_w, L, beta, aux);
        case GLM_OLS: return coop_launch_fam<FamOls>(k, cols, has_w, L, beta, aux);
        default: return cudaErrorInvalidValue;
        }
    }
    switch (fam) {
    case GLM_PROBIT: return wide_launch_fam<FamProbit>(k, cols, has_w, L, beta, aux);
    case GLM_LOGIT2: return wide_launch_fam<FamLogit2>(k, cols, has_w, L, beta, aux);
    case GLM_POISSON_REG: return wide_launch_fam<FamPoissonReg>(k, cols, has_w, L, beta, aux);
    case GLM_OLS: return wide_launch_fam<FamOls>(k, cols, has_w, L, beta, aux);
    default: return cudaErrorInvalidValue;
    }
}
int glm_wide_blocks_per_sm(GlmFam fam, int k) {
    if (k <= COOP_MAX_K) {
        switch (fam) {
        case GLM_PROBIT: return coop_bps_fam<FamProbit>(k);
        case GLM_LOGIT2: return coop_bps_fam<FamLogit2>(k);
        case GLM_POISSON_REG: return coop_bps_fam<FamPoissonReg>(k);
        case GLM_OLS: return coop_bps_fam<FamOls>(k);
        default: return 0;
        }
    }
    switch (fam) {
    case GLM_PROBIT: return wide_bps_fam<FamProbit>(k);
    case GLM_LOGIT2: return wide_bps_fam<FamLogit2>(k);
    case GLM_POISSON_REG: return wide_bps_fam<FamPoissonReg>(k);
    case GLM_OLS: return wide_bps_fam<FamOls>(k);
    default: return 0;
    }
}

}  // namespace apb
