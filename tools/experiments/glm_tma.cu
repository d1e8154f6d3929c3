// glm_tma.cu -- the fused LL + score pass with the tiles streamed by the TMA engine.
//
// A resident tile is one contiguous (k+1) x 256-byte chunk, so one cp.async.bulk (SASS UBLKCP)
// per tile moves it global -> shared; each warp owns a ring of GLM_TMA_STAGES tile buffers with
// one mbarrier per stage, keeps GLM_TMA_STAGES-1 tiles in flight ahead of the one it is
// consuming, and reads its row (lane = row, column-major tile: conflict free) into registers.
// Memory-level parallelism no longer depends on registers or occupancy.  Same arithmetic,
// same reduction tail and the same results bit for bit as glm_fused_kernel for a given grid.
#include "glm_kernel.cuh"

namespace apb {

constexpr int GLM_TMA_STAGES = 3;

template <int K, class Fam, bool HAS_W>
__global__ void __launch_bounds__(GLM_THREADS)
glm_tma_kernel(const double* __restrict__ tiles, const unsigned long long n_rows, const unsigned long long n_tiles,
               const __grid_constant__ GlmParams prm, double* __restrict__ partials, unsigned int* __restrict__ ticket,
               double* __restrict__ out, const __grid_constant__ FinalizeCtx fin) {
    constexpr int COLS = K + 1 + (HAS_W ? 1 : 0);
    constexpr int NOUT = MAX_ACC + K;
    constexpr int WARPS = GLM_THREADS / 32;
    constexpr int TILE_D = COLS * TILE_ROWS;
    constexpr unsigned TILE_BYTES = TILE_D * sizeof(double);
    extern __shared__ __align__(128) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* ring = smem + warp * (GLM_TMA_STAGES * TILE_D);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + WARPS * GLM_TMA_STAGES * TILE_D) + warp * GLM_TMA_STAGES;
    const unsigned long long gwarp = (unsigned long long)blockIdx.x * WARPS + warp;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * WARPS;

    if (lane == 0)
        for (int s = 0; s < GLM_TMA_STAGES; s++) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    auto issue = [&](unsigned long long tt, int stage) {
        if (lane == 0) {
            fence_proxy_async();
            mbar_expect_tx(bars + stage, TILE_BYTES);
            tma_load_1d(ring + stage * TILE_D, tiles + tt * (unsigned long long)TILE_D, TILE_BYTES, bars + stage);
        }
    };
#pragma unroll
    for (int s = 0; s < GLM_TMA_STAGES - 1; s++)
        if (gwarp + s * nwarps < n_tiles) issue(gwarp + s * nwarps, s);

    double g[K];
#pragma unroll
    for (int j = 0; j < K; j++) g[j] = 0.0;
    KSum acc[MAX_ACC];

    unsigned it = 0;
    for (unsigned long long t = gwarp; t < n_tiles; t += nwarps, it++) {
        const int stage = it % GLM_TMA_STAGES;
        const unsigned long long ahead = t + (unsigned long long)(GLM_TMA_STAGES - 1) * nwarps;
        if (ahead < n_tiles) issue(ahead, (it + GLM_TMA_STAGES - 1) % GLM_TMA_STAGES);   // the stage consumed last iteration
        mbar_wait(bars + stage, (it / GLM_TMA_STAGES) & 1u);
        const double* base = ring + stage * TILE_D + lane;
        double x[K];
#pragma unroll
        for (int j = 0; j < K; j++) x[j] = base[j * TILE_ROWS];
        const double y = base[K * TILE_ROWS];
        const double w = HAS_W ? base[(K + 1) * TILE_ROWS] : 1.0;
        __syncwarp();   // every lane has its row in registers: the stage may be refilled

        double xb = 0.0;
#pragma unroll
        for (int j = 0; j < K; j++) xb = fma(x[j], prm.beta[j], xb);
        RowOut o = Fam::eval(xb, y, w, prm.aux);
        const bool valid = t * TILE_ROWS + lane < n_rows;
#pragma unroll
        for (int a = 0; a < Fam::NACC; a++) acc[a].add(valid ? o.s[a] : 0.0);
        const double d = valid ? o.d : 0.0;
#pragma unroll
        for (int j = 0; j < K; j++) g[j] = fma(x[j], d, g[j]);
    }

    __syncthreads();
    double (*red)[NOUT] = reinterpret_cast<double (*)[NOUT]>(smem);   // the rings are free now
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
    double* scratch = smem + WARPS * NOUT;
    if (threadIdx.x < NOUT) {
        double s = 0.0;
#pragma unroll
        for (int wv = 0; wv < WARPS; wv++) s += red[wv][threadIdx.x];
        partials[(size_t)blockIdx.x * NOUT + threadIdx.x] = s;
    }
    grid_finalize(partials, NOUT, ticket, out, fin, scratch);
}

template <int K, class Fam, bool HAS_W>
static size_t tma_smem() {
    const size_t ring = (size_t)(GLM_THREADS / 32) * GLM_TMA_STAGES * (K + 1 + (HAS_W ? 1 : 0)) * TILE_ROWS * sizeof(double);
    return ring + (GLM_THREADS / 32) * GLM_TMA_STAGES * sizeof(unsigned long long) + 64;
}
template <int K, class Fam, bool HAS_W>
static cudaError_t tma_launch_one(const GlmLaunch& L, const GlmParams& prm) {
    const size_t sm = tma_smem<K, Fam, HAS_W>();
    static bool done = false;
    if (!done) { cudaFuncSetAttribute(glm_tma_kernel<K, Fam, HAS_W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); done = true; }
    glm_tma_kernel<K, Fam, HAS_W><<<L.grid, GLM_THREADS, sm, L.stream>>>(L.tiles, L.n_rows, L.n_tiles, prm, L.partials, L.ticket, L.out, L.fin);
    return cudaGetLastError();
}
template <int K, class Fam, bool HAS_W>
static int tma_bps_one() {
    const size_t sm = tma_smem<K, Fam, HAS_W>();
    cudaFuncSetAttribute(glm_tma_kernel<K, Fam, HAS_W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, glm_tma_kernel<K, Fam, HAS_W>, GLM_THREADS, sm);
    return nb;
}

// probit, the wide-row shapes where the register-resident kernel runs at 12 warps/SM
#define APB_TMA_K(OP)        \
    switch (k) {             \
    case 16: return OP(16);  \
    case 20: return OP(20);  \
    case 24: return OP(24);  \
    case 32: return OP(32);  \
    default: break;          \
    }
cudaError_t glm_tma_launch_probit(int k, const GlmLaunch& L, const GlmParams& prm) {
#define APB_OP(KK) tma_launch_one<KK, FamProbit, false>(L, prm)
    APB_TMA_K(APB_OP)
#undef APB_OP
    return cudaErrorInvalidValue;
}
int glm_tma_blocks_per_sm_probit(int k) {
#define APB_OP(KK) tma_bps_one<KK, FamProbit, false>()
    APB_TMA_K(APB_OP)
#undef APB_OP
    return 0;
}

}  // namespace apb
