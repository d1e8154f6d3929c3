// logit_mma.cu -- multinomial logit with BOTH contractions on the FP64 MMA path and the soft-max
// lane = row in between (APOP_B200_LOGIT_VARIANT=mma).  Same ring, same soft-max and same score
// stage as logit_hybrid.cu; stage 1 differs:
//   stage 1   S^T[c][row] = B^T[c][j] . X^T[j][row] as 32 DMMA.8x8x4 (A fragments = B^T, loop
//             invariant in registers; B fragments = 8-byte shared-memory loads of the staged tile),
//             the C fragments go through the 8 x 32 slab to the lane = row soft-max and the
//             residuals come back through the same slab cells (same lane reads and writes a cell:
//             no barrier in between).
// DMMA and DFMA share one datapath (tools/micro/pipes.cu), so this costs 500 instead of 470
// datapath cycles per tile for stage 1 -- but 36 instructions instead of ~430 (224 DFMA, 118
// uniform loads of B, 33 shared loads, address arithmetic).  16-byte chunks of the staged columns
// are XOR-swizzled by the two low bits of the column index so that all three access patterns
// (stage-1 fragments, stage-3 fragments, lane = row) are bank-conflict free without padding.
#include "logit_kernel.cuh"
#include "fastmath.cuh"
#include <stdlib.h>

namespace apb {

// volatile: ptxas otherwise re-serialises the independent accumulator chains (dependent DMMAs back to
// back, each waiting out the 26-cycle latency of the previous one) to save registers
__device__ __forceinline__ void dmma_884m(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1)
        : "d"(a), "d"(b));
}


template <int KB, int C1, int HYB_STAGES>
__global__ void __launch_bounds__(LOGIT_THREADS, HYB_STAGES == 2 ? 3 : 4)
logit_mma_kernel(const double* __restrict__ tiles, const unsigned long long n_rows,
                    const unsigned long long n_tiles, const int K, const __grid_constant__ LogitParams prm,
                    double* __restrict__ partials, unsigned int* __restrict__ ticket,
                    double* __restrict__ out, const __grid_constant__ FinalizeCtx fin) {
    constexpr int WARPS = LOGIT_THREADS / 32;
    constexpr int JB = KB / 8;
    constexpr int TILE_D = (KB + 1) * TILE_ROWS;          // staged tile: KB columns (>= K: zero) + y, unpadded
    constexpr int WARP_D = HYB_STAGES * TILE_D + 8 * TILE_ROWS;
    constexpr int NOUT = 1 + KB * C1;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    double* ring = smem + warp * WARP_D;
    double* slab = ring + HYB_STAGES * TILE_D;            // R^T[c][row], 8 rows of 32, same swizzle
    const unsigned long long gwarp = (unsigned long long)blockIdx.x * WARPS + warp;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * WARPS;
    const int cols = K + 1;

    // columns K..KB-1 of both stages and the slab rows of the padding categories are zero for the
    // whole kernel (the prefetch never writes them): no predicate in the loops below
    for (int i = lane; i < HYB_STAGES * TILE_D; i += 32) {
        const int col = (i % TILE_D) / TILE_ROWS;
        if (col >= K && col < KB) ring[i] = 0.0;
    }
    for (int i = C1 * TILE_ROWS + lane; i < 8 * TILE_ROWS; i += 32) slab[i] = 0.0;
    __syncwarp();

    // swizzle: 16-byte chunk c of column j is stored at chunk c ^ sw(j), sw(j) = ((j & 1) << 2) | (((j >> 1) & 1) << 1).
    // 16-byte chunk `lane + 32 i` of a resident tile is chunk (lane & 15) of column (lane >> 4) + 2 i, so
    // sw = (par << 2) | ((i & 1) << 1): two per-lane constants.  y (resident column K) is staged at the
    // fixed position KB (a multiple of 8: unswizzled)
    const int par = lane >> 4;
    const int pf_e = par * TILE_ROWS + (((lane & 15) ^ (par << 2)) << 1);
    const int pf_o = par * TILE_ROWS + (((lane & 15) ^ (par << 2) ^ 2) << 1);
    const int pf_y = KB * TILE_ROWS + ((lane & 15) << 1);
    auto prefetch = [&](unsigned long long tt, int stage) {
        const double* src = tiles + tt * (unsigned long long)(cols * TILE_ROWS) + lane * 2;
        const unsigned se = smem_u32(ring + stage * TILE_D + pf_e);
        const unsigned so_ = smem_u32(ring + stage * TILE_D + pf_o);
        const unsigned sy = smem_u32(ring + stage * TILE_D + pf_y);
#pragma unroll
        for (int i = 0; i < (KB + 2) / 2; i++) {
            const int col = par + 2 * i;
            if (col < K)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(((i & 1) ? so_ : se) + i * 2 * TILE_ROWS * 8), "l"(src + i * 64));
            else if (col == K)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sy), "l"(src + i * 64));
        }
        asm volatile("cp.async.commit_group;");
    };
    // lane = row view of slab row c: row `lane` sits at c*32 + ro[c & 3]
    int ro[4];
#pragma unroll
    for (int s = 0; s < 4; s++) ro[s] = (((lane >> 1) ^ (((s & 1) << 2) | ((s >> 1) << 1))) << 1) | (lane & 1);
    // stage-3 / slab fragment view: rows nb*8 + 2q + {0,1} of column (category) n = .. + g are chunk nb*4 + q
    const int swg = ((g & 1) << 2) | (((g >> 1) & 1) << 1);
    int fo[4];
#pragma unroll
    for (int nb = 0; nb < 4; nb++) fo[nb] = g * TILE_ROWS + ((((nb << 2) + q) ^ swg) << 1);
    // stage-1 fragment view: row nb*8 + g of column ks*4 + q (sw depends on q only)
    const int swq = ((q & 1) << 2) | ((q >> 1) << 1);
    int xo[4];
#pragma unroll
    for (int nb = 0; nb < 4; nb++) xo[nb] = q * TILE_ROWS + (((((nb << 2) + (g >> 1)) ^ swq) << 1) | (g & 1));
    constexpr int KS = KB / 4;
    double bfrag[KS];              // B^T[c = g][j = ks*4 + q], zero for the padding category / columns
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
        const int j = ks * 4 + q;
        bfrag[ks] = (g < C1 && j < K) ? prm.B[j * C1 + g] : 0.0;
    }

    double G[JB][2], G2[JB][2];
#pragma unroll
    for (int jb = 0; jb < JB; jb++) G[jb][0] = G[jb][1] = G2[jb][0] = G2[jb][1] = 0.0;
    KSum ll;

    int cur = 0;
    if (gwarp < n_tiles) prefetch(gwarp, 0);
    for (unsigned long long t = gwarp; t < n_tiles; t += nwarps) {
        if (HYB_STAGES == 2 && t + nwarps < n_tiles) {
            prefetch(t + nwarps, cur ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        const double* tile = ring + cur * TILE_D;

        // ---- stage 1 on the MMA path: 4 accumulator chains (one per row block), consecutive DMMAs of a
        // chain are 4 issue slots = 64 cycles apart (latency 26)
        double S[4][2];
#pragma unroll
        for (int nb = 0; nb < 4; nb++) S[nb][0] = S[nb][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
            double xa[4];
#pragma unroll
            for (int nb = 0; nb < 4; nb++) xa[nb] = tile[ks * 4 * TILE_ROWS + xo[nb]];
#pragma unroll
            for (int nb = 0; nb < 4; nb++) dmma_884m(S[nb][0], S[nb][1], bfrag[ks], xa[nb]);
            __syncwarp();   // scheduling fence: keeps ptxas from gathering each chain's DMMAs back to back
        }
        // C fragments S^T[c = g][rows nb*8 + 2q + {0,1}] -> slab (the padding rows receive zeros)
#pragma unroll
        for (int nb = 0; nb < 4; nb++) *reinterpret_cast<double2*>(slab + fo[nb]) = make_double2(S[nb][0], S[nb][1]);
        __syncwarp();
        double xb[C1];
#pragma unroll
        for (int c = 0; c < C1; c++) xb[c] = slab[c * TILE_ROWS + ro[c & 3]];
        const double y = tile[KB * TILE_ROWS + lane];
        // find_index (model/apop_probit.c:206-210): first category equal to y, at most C1 steps
        int idx = C1;
#pragma unroll
        for (int c = C1 - 1; c >= 0; c--) idx = (y == prm.cats[c]) ? c : idx;
        const bool valid = t * TILE_ROWS + lane < n_rows;

        // ---- soft-max
        bool fast = true;
#pragma unroll
        for (int c = 0; c < C1; c++) fast = fast && fm_exp_fast_ok(xb[c]);
        // the log of the LL term is not needed by the score: only its argument is fixed here, and the
        // (long, serial) fm_log chain is issued after the barrier, among the DMMAs of stage 3
        double e[C1], num = 0.0, inv, lsum, lbase = 0.0;
        {
            double sum = 1.0;                      // the numeraire's exp(0)
#pragma unroll
            for (int c = 0; c < C1; c++) {
                e[c] = fm_exp_fast(xb[c]);         // garbage (never a trap) where the guard fails
                sum += e[c];
                num = (idx == c + 1) ? xb[c] : num;
            }
            lsum = sum;
            inv = fm_rcp(sum);
        }
        if (!__all_sync(0xffffffffu, fast)) {      // rare: the reference-shaped shifted form for the whole warp
            double mx = xb[0];
#pragma unroll
            for (int c = 1; c < C1; c++) mx = fmax(mx, xb[c]);
            double sum = 0.0;
#pragma unroll
            for (int c = 0; c < C1; c++) {
                e[c] = fm_exp(xb[c] - mx);
                sum += e[c];
            }
            const double e0 = fm_exp(-mx);         // the numeraire's exp(0 - max)
            // when exp(-max) overflows the reference takes log-sum-exp = max - max:
            // lbase + log(lsum) = 0 with lsum = 1
            const bool over = mx < -700.0;
            lsum = over ? 1.0 : sum + e0; lbase = over ? 0.0 : mx;
            // probabilities for the score: overflow of e0 means p_c = 0 for every free category
            inv = over ? 0.0 : fm_rcp(sum + e0);
        }
        inv = valid ? inv : 0.0;
#pragma unroll
        for (int c = 0; c < C1; c++) {
            const double ind = (valid && idx == c + 1) ? 1.0 : 0.0;
            slab[c * TILE_ROWS + ro[c & 3]] = fma(-e[c], inv, ind);   // the cell this lane read x.b_c from
        }
        __syncwarp();

        // ---- stage 3: A fragments R^T[c = g][rows nb*8 + 2q + {0,1}]; X fragments with the same rows
        double R[4][2];
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
            const double2 r2 = *reinterpret_cast<const double2*>(slab + fo[nb]);
            R[nb][0] = r2.x; R[nb][1] = r2.y;
        }
        ll.add(valid ? num - (lbase + fm_log_normal(lsum)) : 0.0);
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {           // JB chains; consecutive DMMAs of a chain are JB slots apart
            double2 x2[JB];
#pragma unroll
            for (int jb = 0; jb < JB; jb++) x2[jb] = *reinterpret_cast<const double2*>(tile + jb * 8 * TILE_ROWS + fo[nb]);
#pragma unroll
            for (int jb = 0; jb < JB; jb++) dmma_884m(G[jb][0], G[jb][1], R[nb][0], x2[jb].x);
#pragma unroll
            for (int jb = 0; jb < JB; jb++) dmma_884m((JB >= 2 ? G[jb][0] : G2[jb][0]), (JB >= 2 ? G[jb][1] : G2[jb][1]), R[nb][1], x2[jb].y);
            __syncwarp();   // scheduling fence, as in stage 1
        }
        __syncwarp();                              // every lane is done with this stage before it is refilled
        if (HYB_STAGES == 2) cur ^= 1;
        else if (t + nwarps < n_tiles) prefetch(t + nwarps, 0);   // single buffer: latency hidden across warps only
    }

    // ---- CTA reduction: lane (g, q) holds G^T[c = g][j = jb*8 + 2q + {0,1}]
#pragma unroll
    for (int jb = 0; jb < JB; jb++) { G[jb][0] += G2[jb][0]; G[jb][1] += G2[jb][1]; }
    __syncthreads();                               // the rings are free: reuse as WARPS x NOUT scratch
    const double llw = warp_sum(ll.value());
    if (lane == 0) smem[warp * NOUT] = llw;
    if (g < C1) {
#pragma unroll
        for (int jb = 0; jb < JB; jb++) {
            smem[warp * NOUT + 1 + (jb * 8 + 2 * q) * C1 + g] = G[jb][0];
            smem[warp * NOUT + 1 + (jb * 8 + 2 * q + 1) * C1 + g] = G[jb][1];
        }
    }
    __syncthreads();
    double* scratch = smem + WARPS * NOUT;
    for (int v = threadIdx.x; v < NOUT; v += LOGIT_THREADS) {
        double s = 0.0;
#pragma unroll
        for (int wv = 0; wv < WARPS; wv++) s += smem[wv * NOUT + v];
        partials[(size_t)blockIdx.x * NOUT + v] = s;
    }
    grid_finalize(partials, NOUT, ticket, out, fin, scratch);
}

// two ring stages per warp; a single stage with 16 warps per SM measured no better (1.17 vs 1.15 ms on
// 20M x 32, C = 8), three stages with 8 warps worse (1.20 ms): the template parameter stays for such runs
constexpr int MMA_STAGES = 2;
template <int KB, int ST>
static constexpr size_t mma_smem() {
    return (size_t)(LOGIT_THREADS / 32) * (ST * (KB + 1) * TILE_ROWS + 8 * TILE_ROWS) * sizeof(double);
}
template <int KB, int C1, int ST>
static cudaError_t mma_launch_st(const LogitLaunch& L, const LogitParams& prm) {
    static_assert((size_t)(LOGIT_THREADS / 32) * (1 + KB * C1) + (1 + KB * C1) <= mma_smem<KB, ST>() / sizeof(double), "reduction scratch must fit the rings");
    static bool done = false;
    if (!done) { cudaFuncSetAttribute(logit_mma_kernel<KB, C1, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mma_smem<KB, ST>()); done = true; }
    logit_mma_kernel<KB, C1, ST><<<L.grid, LOGIT_THREADS, mma_smem<KB, ST>(), L.stream>>>(L.tiles, L.n_rows, L.n_tiles, L.k, prm, L.partials, L.ticket, L.out, L.fin);
    return cudaGetLastError();
}
template <int KB, int C1, int ST>
static int mma_bps_st() {
    cudaFuncSetAttribute(logit_mma_kernel<KB, C1, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mma_smem<KB, ST>());
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, logit_mma_kernel<KB, C1, ST>, LOGIT_THREADS, mma_smem<KB, ST>());
    return nb;
}
template <int KB, int C1>
static cudaError_t mma_launch_one(const LogitLaunch& L, const LogitParams& prm) {
    return mma_launch_st<KB, C1, MMA_STAGES>(L, prm);
}
template <int KB, int C1>
static int mma_bps_one() { return mma_bps_st<KB, C1, MMA_STAGES>(); }

#define APB_MMA_C(KB)                          \
    switch (c1) {                              \
    case 2: return APB_OP(KB, 2);              \
    case 3: return APB_OP(KB, 3);              \
    case 4: return APB_OP(KB, 4);              \
    case 5: return APB_OP(KB, 5);              \
    case 6: return APB_OP(KB, 6);              \
    case 7: return APB_OP(KB, 7);              \
    default: break;                            \
    }
cudaError_t logit_mma_launch(int c1, const LogitLaunch& L, const LogitParams& prm) {
#define APB_OP(KB, C) mma_launch_one<KB, C>(L, prm)
    switch (logit_kb(L.k)) {
    case 8: APB_MMA_C(8) break;
    case 16: APB_MMA_C(16) break;
    case 24: APB_MMA_C(24) break;
    case 32: APB_MMA_C(32) break;
    }
#undef APB_OP
    return cudaErrorInvalidValue;
}
int logit_mma_blocks_per_sm(int k, int c1) {
#define APB_OP(KB, C) mma_bps_one<KB, C>()
    switch (logit_kb(k)) {
    case 8: APB_MMA_C(8) break;
    case 16: APB_MMA_C(16) break;
    case 24: APB_MMA_C(24) break;
    case 32: APB_MMA_C(32) break;
    }
#undef APB_OP
    return 0;
}

}  // namespace apb
