// glm_wide.cu -- the fused LL + score reduction for data wider than the register-resident
// kernels cover (32 < k <= GLM_WIDE_MAX_K).  Same tile layout, same families, same reduction
// tail; what changes is that a row no longer fits in registers:
//   sweep 1  lane = row streams the tile's columns once (HBM, coalesced) accumulating x.beta
//            (beta from global memory, broadcast through L1), then evaluates the family;
//   sweep 2  re-reads the tile in chunks of 32 columns (L1/L2 hits), parks each chunk
//            transposed in the warp's shared-memory slab and lets lane = column add
//            sum_i x_ij d_i into the warp's accumulator row in shared memory.
// HBM traffic stays 8(k+1) bytes per row; the second touch of each tile is served by L2.
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
    switch (fam) {
    case GLM_PROBIT: return wide_launch_fam<FamProbit>(k, cols, has_w, L, beta, aux);
    case GLM_LOGIT2: return wide_launch_fam<FamLogit2>(k, cols, has_w, L, beta, aux);
    case GLM_POISSON_REG: return wide_launch_fam<FamPoissonReg>(k, cols, has_w, L, beta, aux);
    case GLM_OLS: return wide_launch_fam<FamOls>(k, cols, has_w, L, beta, aux);
    default: return cudaErrorInvalidValue;
    }
}
int glm_wide_blocks_per_sm(GlmFam fam, int k) {
    switch (fam) {
    case GLM_PROBIT: return wide_bps_fam<FamProbit>(k);
    case GLM_LOGIT2: return wide_bps_fam<FamLogit2>(k);
    case GLM_POISSON_REG: return wide_bps_fam<FamPoissonReg>(k);
    case GLM_OLS: return wide_bps_fam<FamOls>(k);
    default: return 0;
    }
}

}  // namespace apb
