// logit_hybrid.cu -- multinomial logit, default kernel: lane = row for X.B and the soft-max,
// FP64 DMMA for the score contraction, tiles streamed through a per-warp shared-memory ring.
//
// Measured on the B200 (tools/micro/pipes.cu): DFMA and DMMA.8x8x4 issue to the SAME FP64
// datapath (34 and 37 TFLOP/s alone, 32 TFLOP/s interleaved), so the tensor path buys no
// extra flops here -- what it buys is shared-memory traffic and issue slots.  Per 32-row tile:
//   prefetch  the next tile of this warp arrives with 16-byte cp.async (LDGSTS, L1 bypassed)
//             while the current one is processed; 16-byte chunks of odd columns are XOR-
//             swizzled (chunk ^ 4) so that BOTH access patterns below are bank-conflict free
//             without padding;
//   stage 1   lane = row: x_ij from shared memory (one 256-byte row of banks per column),
//             C-1 DFMA chains against B in the constant bank -- no padded category, no
//             cross-lane traffic;
//   soft-max  lane = row, one_logit_row (model/apop_probit.c:212-229).  Fast path when every
//             |x.b_c| < 708 (tested on the high words, ALU pipe): log(1 + sum_c exp(xb_c))
//             directly, which is the reference's max + log(sum exp(xb_c - max) + exp(-max))
//             without the shift (equal to rounding), with the 17-operation fm_exp_fast; any
//             other row sends its warp through the reference-shaped shifted form;
//   stage 3   G^T[c][j] += R^T[c][row] . X[row][j] as DMMA.8x8x4 (M = categories, N = 8
//             columns, K = 4 rows): the residuals go through an 8 x 32 slab (written lane =
//             row, read back as A fragments), X fragments are 16-byte shared-memory loads
//             (rows 2q, 2q+1 of a column), 32 DMMA per tile in 8 independent chains.
// X and y are read from HBM exactly once: 8(k+1) bytes per row.
#include "logit_kernel.cuh"
#include "fastmath.cuh"
#include <stdlib.h>

namespace apb {

// WEIGHTED: every row enters with the multiplicity a bootstrap replicate gives it (the resident resample counts of
// apop_b200_bootstrap_counts, one byte per (row, replicate), layout of batched_kernel.cu): LL and score are those of
// the resampled data set without a single row being copied (apop_bootstrap.m4.c:143-149 copies all N).
template <int KB, int C1, int HYB_STAGES, int VAR, bool WEIGHTED>
__global__ void __launch_bounds__(LOGIT_THREADS, HYB_STAGES == 2 ? 3 : 4)
logit_hybrid_kernel(const double* __restrict__ tiles, const unsigned long long n_rows,
                    const unsigned long long n_tiles, const int K, const int cols, const unsigned char* __restrict__ counts,
                    const int m_pad, const int replicate, const __grid_constant__ LogitParams prm,
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

    // columns K..KB-1 of both stages and the slab rows of the padding categories are zero for the
    // whole kernel (the prefetch never writes them): no predicate in the loops below
    for (int i = lane; i < HYB_STAGES * TILE_D; i += 32) {
        const int col = (i % TILE_D) / TILE_ROWS;
        if (col >= K && col < KB) ring[i] = 0.0;
    }
    for (int i = C1 * TILE_ROWS + lane; i < 8 * TILE_ROWS; i += 32) slab[i] = 0.0;
    __syncwarp();

    // 16-byte chunk `lane + 32 i` of a resident tile is chunk (lane & 15) of column (lane >> 4) + 2 i:
    // the column parity, hence the swizzle, is a per-lane constant.  y (resident column K) is staged
    // at the fixed position KB (even: unswizzled)
    const int par = lane >> 4;
    const int pf_x = par * TILE_ROWS + (((lane & 15) ^ (par << 2)) << 1);
    const int pf_y = KB * TILE_ROWS + (((lane & 15) ^ ((KB & 1) << 2)) << 1);
    auto prefetch = [&](unsigned long long tt, int stage) {
        const double* src = tiles + tt * (unsigned long long)(cols * TILE_ROWS) + lane * 2;
        const unsigned sx = smem_u32(ring + stage * TILE_D + pf_x);
        const unsigned sy = smem_u32(ring + stage * TILE_D + pf_y);
#pragma unroll
        for (int i = 0; i < (KB + 2) / 2; i++) {
            // (a branch-free, predicated form of this loop costs 48 more registers and measured 11 % slower.  Round 2:
            // ncu's source view puts 22 % of the stall samples on this loop -- ~12 issue slots per copy, VIADD / ISETP /
            // BRA chains -- yet a K == KB specialisation without any column test, 17 copies back to back or spread in
            // three portions over stage 1, measured 1.21-1.22 ms against 1.085-1.14 for this form on the same boxes:
            // 168 registers and spills instead of 136.  It stays as it is.)
            const int col = par + 2 * i;
            if (col < K)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sx + i * 2 * TILE_ROWS * 8), "l"(src + i * 64));
            else if (col == K)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sy), "l"(src + i * 64));
        }
        asm volatile("cp.async.commit_group;");
    };
    // lane = row view: element (row = lane, column j) sits at j*32 + row_off[j & 1]
    const int row_off0 = lane, row_off1 = (((lane >> 1) ^ 4) << 1) | (lane & 1);
    // fragment view: rows nb*8 + 2q + {0,1} of column (or category) n are chunk nb*4 + q of that
    // column, stored at chunk (nb*4 + q) ^ 4s = (nb ^ s)*4 + q with s = n & 1 (= g & 1 here)
    int fo[4];
#pragma unroll
    for (int nb = 0; nb < 4; nb++) fo[nb] = g * TILE_ROWS + ((((nb ^ (g & 1)) << 2) + q) << 1);

    double G[JB][2], G2[JB][2];
#pragma unroll
    for (int jb = 0; jb < JB; jb++) G[jb][0] = G[jb][1] = G2[jb][0] = G2[jb][1] = 0.0;
    KSum ll;
    // VAR & 1: sum_i log(lsum_i) as the log of a running product: mantissa in prod (kept in [1, 2) by moving its
    // exponent into prod_e with integer operations), one fm_log per lane at the end instead of one per row
    double prod = 1.0;
    long long prod_e = 0;
    constexpr bool PROD = (VAR & 1) != 0 && !WEIGHTED, TAB = (VAR & 2) != 0, NB_OUTER = (VAR & 4) != 0;
    // this lane's byte inside a replicate's 32-byte count record (count_pos of batched_kernel.cu)
    const int cpos = ((lane & 7) >> 1) * 8 + (lane >> 3) * 2 + (lane & 1);
    __shared__ double exp_tab[16];
    if (TAB) {
        if (threadIdx.x < 16) exp_tab[threadIdx.x] = exp2((double)threadIdx.x * 0.0625);
        __syncthreads();
    }

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
        const double* p0 = tile + row_off0;
        const double* p1 = tile + row_off1;

        // ---- stage 1: lane = row; x in register blocks of 16 columns so that the shared-memory
        // latency is paid per block, not per column
        double xb[C1];
#pragma unroll
        for (int c = 0; c < C1; c++) xb[c] = 0.0;
        // (XBLK must divide KB: a 16-wide block over KB = 24 read columns 24..31 -- the y column and the other ring
        // stage, times the zero padding of B -- and turned the LL into NaN whenever that stage had never been filled
        // and held a NaN bit pattern: found by tools/fuzz_parity.py at n = 4097, k = 18)
        constexpr int XBLK = (KB % 16 == 0) ? 16 : 8;
        static_assert(KB % XBLK == 0, "register block must tile the staged columns exactly");
#pragma unroll
        for (int j0 = 0; j0 < KB; j0 += XBLK) {
            double xv[XBLK];
#pragma unroll
            for (int u = 0; u < XBLK; u++) xv[u] = (((j0 + u) & 1) ? p1 : p0)[(j0 + u) * TILE_ROWS];
#pragma unroll
            for (int u = 0; u < XBLK; u++)
#pragma unroll
                for (int c = 0; c < C1; c++) xb[c] = fma(xv[u], prm.B[(j0 + u) * C1 + c], xb[c]);
        }
        const double y = ((KB & 1) ? p1 : p0)[KB * TILE_ROWS];
        // find_index (model/apop_probit.c:206-210): first category equal to y, at most C1 steps
        int idx = C1;
#pragma unroll
        for (int c = C1 - 1; c >= 0; c--) idx = (y == prm.cats[c]) ? c : idx;
        const bool in_range = t * TILE_ROWS + lane < n_rows;
        double wrow = 1.0;
        if (WEIGHTED) wrow = (double)__ldg(counts + (t * (unsigned long long)m_pad + replicate) * TILE_ROWS + cpos);
        const bool valid = in_range && (!WEIGHTED || wrow != 0.0);

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
                e[c] = TAB ? exp_tab16(xb[c], exp_tab) : fm_exp_fast(xb[c]);   // garbage (never a trap) where the guard fails
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
        // (the slab's previous contents were read before the barrier that ended the last iteration)
        inv = valid ? inv * wrow : 0.0;
#pragma unroll
        for (int c = 0; c < C1; c++) {
            const double ind = (valid && idx == c + 1) ? wrow : 0.0;
            slab[c * TILE_ROWS + ((c & 1) ? row_off1 : row_off0)] = fma(-e[c], inv, ind);
        }
        __syncwarp();

        // ---- stage 3: A fragments R^T[c = g][rows nb*8 + 2q + {0,1}]; X fragments with the same rows
        double R[4][2];
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
            const double2 r2 = *reinterpret_cast<const double2*>(slab + fo[nb]);
            R[nb][0] = r2.x; R[nb][1] = r2.y;
        }
        if (PROD) {
            ll.add(valid ? num - lbase : 0.0);
            // lsum in [1, 8 e^708): split off its exponent first so that the product cannot overflow
            const double ls = valid ? lsum : 1.0;
            const int eh = __double2hiint(ls);
            prod *= __hiloint2double((eh & 0x000fffff) | 0x3ff00000, __double2loint(ls));   // mantissa of lsum, in [1, 2)
            const int ph = __double2hiint(prod);                                              // prod in [1, 4)
            prod_e += (long long)((eh >> 20) - 1023) + (long long)((ph >> 20) - 1023);
            prod = __hiloint2double((ph & 0x000fffff) | 0x3ff00000, __double2loint(prod));
        } else
            ll.add(valid ? wrow * (num - (lbase + fm_log_normal(lsum))) : 0.0);
        if (NB_OUTER) {
#pragma unroll
            for (int nb = 0; nb < 4; nb++) {
#pragma unroll
                for (int jb = 0; jb < JB; jb++) {
                    const double2 x2 = *reinterpret_cast<const double2*>(tile + jb * 8 * TILE_ROWS + fo[nb]);
                    dmma_884(G[jb][0], G[jb][1], R[nb][0], x2.x);
                    dmma_884(G2[jb][0], G2[jb][1], R[nb][1], x2.y);
                }
            }
        } else {
#pragma unroll
            for (int jb = 0; jb < JB; jb++) {
#pragma unroll
                for (int nb = 0; nb < 4; nb++) {
                    const double2 x2 = *reinterpret_cast<const double2*>(tile + jb * 8 * TILE_ROWS + fo[nb]);
                    dmma_884(G[jb][0], G[jb][1], R[nb][0], x2.x);
                    dmma_884(G2[jb][0], G2[jb][1], R[nb][1], x2.y);
                }
            }
        }
        __syncwarp();                              // every lane is done with this stage before it is refilled
        if (HYB_STAGES == 2) cur ^= 1;
        else if (t + nwarps < n_tiles) prefetch(t + nwarps, 0);   // single buffer: latency hidden across warps only
    }

    // ---- CTA reduction: lane (g, q) holds G^T[c = g][j = jb*8 + 2q + {0,1}]
#pragma unroll
    for (int jb = 0; jb < JB; jb++) { G[jb][0] += G2[jb][0]; G[jb][1] += G2[jb][1]; }
    __syncthreads();                               // the rings are free: reuse as WARPS x NOUT scratch
    if (PROD) ll.add(-((double)prod_e * 6.93147180369123816490e-01 + ((double)prod_e * 1.90821492927058770002e-10 + fm_log_normal(prod))));
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

// product-form LL and table exponential: 1.145 -> 1.098 ms on 20M x 32, C = 8 (each alone: 1.115, 1.134; the nb-outer
// DMMA order made no difference: 1.146); APOP_B200_LOGIT_VAR selects among them in a -DAPB_LOGIT_DEV_VARIANTS build
constexpr int HYB_DEFAULT_VAR = 3;
// two ring stages per warp; a single stage with 16 warps per SM measured no better (1.17 vs 1.15 ms on
// 20M x 32, C = 8), three stages with 8 warps worse (1.20 ms): the template parameter stays for such runs
constexpr int HYB_STAGES = 2;
template <int KB, int ST>
static constexpr size_t hyb_smem() {
    return (size_t)(LOGIT_THREADS / 32) * (ST * (KB + 1) * TILE_ROWS + 8 * TILE_ROWS) * sizeof(double);
}
#ifdef APB_LOGIT_DEV_VARIANTS
static int hyb_var() {   // development switch: APOP_B200_LOGIT_VAR = bit mask (1 product LL, 2 table exp, 4 nb-outer DMMA order)
    const char* e = getenv("APOP_B200_LOGIT_VAR");
    return e ? atoi(e) : HYB_DEFAULT_VAR;
}
#endif
template <int KB, int C1, int ST, int VAR>
static cudaError_t hyb_launch_st(const LogitLaunch& L, const LogitParams& prm) {
    static_assert((size_t)(LOGIT_THREADS / 32) * (1 + KB * C1) + (1 + KB * C1) <= hyb_smem<KB, ST>() / sizeof(double), "reduction scratch must fit the rings");
    static bool done = false;
    if (!done) {
        cudaFuncSetAttribute(logit_hybrid_kernel<KB, C1, ST, VAR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hyb_smem<KB, ST>());
        cudaFuncSetAttribute(logit_hybrid_kernel<KB, C1, ST, VAR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hyb_smem<KB, ST>());
        done = true;
    }
    if (L.counts)
        logit_hybrid_kernel<KB, C1, ST, VAR, true><<<L.grid, LOGIT_THREADS, hyb_smem<KB, ST>(), L.stream>>>(L.tiles, L.n_rows, L.n_tiles, L.k, L.cols, L.counts, L.m_pad, L.replicate, prm, L.partials, L.ticket, L.out, L.fin);
    else
        logit_hybrid_kernel<KB, C1, ST, VAR, false><<<L.grid, LOGIT_THREADS, hyb_smem<KB, ST>(), L.stream>>>(L.tiles, L.n_rows, L.n_tiles, L.k, L.cols, nullptr, 0, 0, prm, L.partials, L.ticket, L.out, L.fin);
    return cudaGetLastError();
}
template <int KB, int C1, int ST, int VAR>
static int hyb_bps_st() {
    cudaFuncSetAttribute(logit_hybrid_kernel<KB, C1, ST, VAR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hyb_smem<KB, ST>());
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, logit_hybrid_kernel<KB, C1, ST, VAR, false>, LOGIT_THREADS, hyb_smem<KB, ST>());
    return nb;
}
template <int KB, int C1>
static cudaError_t hyb_launch_one(const LogitLaunch& L, const LogitParams& prm) {
#ifdef APB_LOGIT_DEV_VARIANTS
    if (KB == 32 && C1 == 7) {
        switch (hyb_var()) {
        case 1: return hyb_launch_st<KB, C1, HYB_STAGES, 1>(L, prm);
        case 2: return hyb_launch_st<KB, C1, HYB_STAGES, 2>(L, prm);
        case 3: return hyb_launch_st<KB, C1, HYB_STAGES, 3>(L, prm);
        case 4: return hyb_launch_st<KB, C1, HYB_STAGES, 4>(L, prm);
        case 7: return hyb_launch_st<KB, C1, HYB_STAGES, 7>(L, prm);
        case 0: return hyb_launch_st<KB, C1, HYB_STAGES, 0>(L, prm);
        default: break;
        }
    }
#endif
    return hyb_launch_st<KB, C1, HYB_STAGES, HYB_DEFAULT_VAR>(L, prm);
}
template <int KB, int C1>
static int hyb_bps_one() { return hyb_bps_st<KB, C1, HYB_STAGES, HYB_DEFAULT_VAR>(); }

#define APB_HYB_C(KB)                          \
    switch (c1) {                              \
    case 2: return APB_OP(KB, 2);              \
    case 3: return APB_OP(KB, 3);              \
    case 4: return APB_OP(KB, 4);              \
    case 5: return APB_OP(KB, 5);              \
    case 6: return APB_OP(KB, 6);              \
    case 7: return APB_OP(KB, 7);              \
    default: break;                            \
    }
cudaError_t logit_hybrid_launch(int c1, const LogitLaunch& L, const LogitParams& prm) {
#define APB_OP(KB, C) hyb_launch_one<KB, C>(L, prm)
    switch (logit_kb(L.k)) {
    case 8: APB_HYB_C(8) break;
    case 16: APB_HYB_C(16) break;
    case 24: APB_HYB_C(24) break;
    case 32: APB_HYB_C(32) break;
    }
#undef APB_OP
    return cudaErrorInvalidValue;
}
int logit_hybrid_blocks_per_sm(int k, int c1) {
#define APB_OP(KB, C) hyb_bps_one<KB, C>()
    switch (logit_kb(k)) {
    case 8: APB_HYB_C(8) break;
    case 16: APB_HYB_C(16) break;
    case 24: APB_HYB_C(24) break;
    case 32: APB_HYB_C(32) break;
    }
#undef APB_OP
    return 0;
}

}  // namespace apb
