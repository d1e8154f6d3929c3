// logit_stream.cu -- multinomial logit, default kernel: one TMA bulk copy per tile, both contractions on the
// FP64 MMA path, the soft-max lane = row in between.
//
// What the cp.async kernel of logit_hybrid.cu spends its time on (ncu source view, profiles/r02_ncu_logit8_source_sass_top.csv): 30 % of a
// warp's cycles go to issuing the 17 LDGSTS of a tile and their column tests, and stage 1 needs ~430
// instructions (224 DFMA, 112 uniform loads of B, 33 shared loads).  Neither is arithmetic.  Here
//   prefetch  ONE cp.async.bulk (UBLKCP) per tile, issued by one lane and completing on an mbarrier: the
//             resident tile is contiguous -- (K+1) x 256 bytes -- and is kept in HBM in the SWIZZLED form of
//             layout.cu (16-byte chunk c of column j stored at c ^ sw(j), sw(j) = ((j & 1) << 2) | (j & 2)), so
//             the bytes land in shared memory ready for all three access patterns below without a conflict;
//   stage 1   S^T[c][row] = B^T[c][j] . X^T[j][row] as 32 DMMA.8x8x4: the A fragments (B^T) are loop invariant
//             in 2 KB/8 = 16 registers, the B fragments 8-byte shared loads; 36 instructions;
//   soft-max  lane = row through an 8 x 32 slab: one_logit_row (model/apop_probit.c:212-229) with the table
//             exponential and the product-form LL of logit_hybrid.cu;
//   stage 3   G^T[c][j] += R^T[c][row] . X[row][j] as 32 DMMA.8x8x4, X fragments 16-byte shared loads.
// Occupancy.  A warp with a private two-slot ring needs 19 KB of shared memory: 12 warps per SM, 3 per scheduler, and
// ncu shows the shared FP64 datapath 77 % busy (0.97 ms) -- the rest is latency no third warp was there to cover.
// The PAIRED form lets two warps share a THREE-slot ring (the pair's tiles alternate between them; whoever finishes
// a slot refills it for its partner) and keeps the slab at 7 rows: 28.9 KB per pair, 16 warps per SM as ONE CTA of
// 512 threads at 128 registers.
// Datapath: 64 DMMA x 16 + ~110 FP64 x 2 cycles = ~1250 cycles per tile and scheduler, against 1330 cycles of
// HBM time at 7.3 TB/s.  X and y are read from HBM exactly once: 8(k+1) bytes per row.
#include "logit_kernel.cuh"
#include "fastmath.cuh"
#include <stdlib.h>
#include <string.h>

namespace apb {

// volatile: ptxas otherwise gathers the DMMAs of one accumulator chain back to back (each waiting out the
// 26-cycle latency of the previous one) to save registers
__device__ __forceinline__ void dmma_884v(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// PAIRED: 16 warps in 8 pairs, 3 slots per pair, one CTA per SM.  Otherwise 4 warps with 2 private slots each, 3 CTAs per SM.
template <bool PAIRED> struct StreamCfg {
    static constexpr int WARPS = PAIRED ? 16 : 4;      // per CTA
    static constexpr int GROUP = PAIRED ? 2 : 1;       // warps sharing a ring
    static constexpr int SLOTS = PAIRED ? 3 : 2;       // tiles in a ring
    static constexpr int SLAB_ROWS = PAIRED ? 7 : 8;   // category rows of a warp's slab
    static constexpr int CTAS = PAIRED ? 1 : 3;        // per SM
};

template <int KB, int C1, bool WEIGHTED, bool PAIRED>
__global__ void __launch_bounds__(StreamCfg<PAIRED>::WARPS * 32, StreamCfg<PAIRED>::CTAS)
logit_stream_kernel(const double* __restrict__ tiles, const unsigned long long n_rows,
                    const unsigned long long n_tiles, const int K, const int cols, const unsigned char* __restrict__ counts,
                    const int m_pad, const int replicate, const __grid_constant__ LogitParams prm,
                    double* __restrict__ partials, unsigned int* __restrict__ ticket,
                    double* __restrict__ out, const __grid_constant__ FinalizeCtx fin) {
    using Cfg = StreamCfg<PAIRED>;
    constexpr int WARPS = Cfg::WARPS, GROUP = Cfg::GROUP, SLOTS = Cfg::SLOTS, SLAB_ROWS = Cfg::SLAB_ROWS;
    constexpr int THREADS = WARPS * 32, NGROUPS = WARPS / GROUP;
    constexpr int JB = KB / 8, KS = KB / 4;
    constexpr int TILE_D = (KB + 1) * TILE_ROWS;          // staged tile: the resident K + 1 columns, byte for byte
    constexpr int SLAB_D = SLAB_ROWS * TILE_ROWS;
    constexpr int GROUP_D = SLOTS * TILE_D + GROUP * SLAB_D;
    constexpr int NOUT = 1 + KB * C1;
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) unsigned long long bars[NGROUPS * SLOTS];
    __shared__ double exp_tab[16];
    __shared__ unsigned seen[NGROUPS * SLOTS];            // PAIRED: fills of each slot seen complete by their consumers
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = warp / GROUP, wg = warp % GROUP;      // ring, and this warp's turn inside it
    const int g = lane >> 2, q = lane & 3;
    double* ring = smem + grp * GROUP_D;
    double* slab = ring + SLOTS * TILE_D + wg * SLAB_D;   // [category][row], swizzled like a tile column
    // the ring's tiles: item i is tile ggrp + i * ngrps; warp wg of the ring takes the items i = wg (mod GROUP);
    // item i lives in slot i % SLOTS and is the (i / SLOTS)-th fill of that slot
    const unsigned ggrp = blockIdx.x * NGROUPS + grp, ngrps = gridDim.x * NGROUPS;

    // the bulk copy writes columns 0..K (X and y); K+1..KB stay zero for the whole kernel, and so do the slab
    // rows of the padding categories
    for (int i = wg * 32 + lane; i < SLOTS * TILE_D; i += GROUP * 32)
        if ((i % TILE_D) / TILE_ROWS > K) ring[i] = 0.0;
    for (int i = C1 * TILE_ROWS + lane; i < SLAB_D; i += 32) slab[i] = 0.0;
    if (wg == 0 && lane < SLOTS) { mbar_init(bars + grp * SLOTS + lane, 1); seen[grp * SLOTS + lane] = 0; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (threadIdx.x < 16) exp_tab[threadIdx.x] = exp2((double)threadIdx.x * 0.0625);
    __syncthreads();

    const unsigned bar_u32 = smem_u32(bars + grp * SLOTS), ring_u32 = smem_u32(ring), seen_u32 = smem_u32(seen + grp * SLOTS);
    const unsigned tile_bytes = (unsigned)(K + 1) * TILE_ROWS * sizeof(double);
    // fill slot `slot` with tile tt: one lane, one UBLKCP; completion (all bytes landed) flips the slot's barrier phase
    auto fill = [&](unsigned long long tt, int slot) {
        if (lane == 0) {
            const unsigned bar = bar_u32 + slot * 8, dst = ring_u32 + slot * (TILE_D * 8);
            const double* src = tiles + tt * (unsigned long long)(cols * TILE_ROWS);
            fence_proxy_async();   // the slot's reads (generic proxy, before the last __syncwarp) precede the async writes
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tile_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(src), "r"(tile_bytes), "r"(bar) : "memory");
        }
    };
    auto wait_fill = [&](int slot, unsigned parity) {
        const unsigned bar = bar_u32 + slot * 8;
        asm volatile("{\n\t.reg .pred p;\n\tWAIT_TILE:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra TILE_READY;\n\tbra WAIT_TILE;\n\tTILE_READY:\n\t}"
                     ::"r"(bar), "r"(parity) : "memory");
    };

    // swizzle of column (or slab row) n: chunk ^ sw(n); as a row offset, row ^ (sw(n) << 1)
    // lane = row view: row `lane` of column n sits at n*32 + ro[n & 3]
    int ro[4];
#pragma unroll
    for (int s = 0; s < 4; s++) ro[s] = lane ^ ((((s & 1) << 2) | (s & 2)) << 1);
    // stage-3 / slab fragment view: rows nb*8 + 2q + {0,1} of column (category) .. + g are chunk nb*4 + q
    const int swg = ((g & 1) << 2) | (g & 2);
    int fo[4];
#pragma unroll
    for (int nb = 0; nb < 4; nb++) fo[nb] = g * TILE_ROWS + ((((nb << 2) + q) ^ swg) << 1);
    // stage-1 fragment view: row nb*8 + g of column ks*4 + q (the swizzle depends on q only)
    const int swq = ((q & 1) << 2) | (q & 2);
    int xo[4];
#pragma unroll
    for (int nb = 0; nb < 4; nb++) xo[nb] = q * TILE_ROWS + (((((nb << 2) + (g >> 1)) ^ swq) << 1) | (g & 1));
    const int y_off = K * TILE_ROWS + (lane ^ ((((K & 1) << 2) | (K & 2)) << 1));
    double bfrag[KS];              // B^T[c = g][j = ks*4 + q], zero for the padding category / columns
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
        const int j = ks * 4 + q;
        bfrag[ks] = (g < C1 && j < K) ? prm.B[j * C1 + g] : 0.0;
    }
    // this lane's byte inside a replicate's 32-byte count record (count_pos of batched_kernel.cu)
    const int cpos = ((lane & 7) >> 1) * 8 + (lane >> 3) * 2 + (lane & 1);

    double G[JB][2];
#pragma unroll
    for (int jb = 0; jb < JB; jb++) G[jb][0] = G[jb][1] = 0.0;
    KSum ll;
    // sum_i log(lsum_i) as the log of a running product: mantissa in prod (kept in [1, 2) by moving its exponent
    // into prod_e with integer operations), one fm_log per lane at the end instead of one per row
    constexpr bool PROD = !WEIGHTED;
    double prod = 1.0;
    int prod_e = 0;   // (a lane's rows x 1025 stay far below 2^31: the grid spreads 10^11 rows before it would not)

    // the first SLOTS items of the ring
    if (wg == 0) {
#pragma unroll
        for (int i = 0; i < SLOTS; i++)
            if (ggrp + (unsigned long long)i * ngrps < n_tiles) fill(ggrp + (unsigned long long)i * ngrps, i);
    }
    for (unsigned i = wg;; i += GROUP) {
        const unsigned long long t = ggrp + (unsigned long long)i * ngrps;
        if (t >= n_tiles) break;
        const int slot = (int)(i % SLOTS);
        const unsigned nfill = i / SLOTS;
        // PAIRED: the previous fill of this slot was the partner's, so this warp has not seen it complete, and a parity
        // wait cannot tell "fill n-1 still in flight" from "fill n done" (same parity two phases apart).  The consumer
        // of every fill therefore publishes the number of fills it has seen land; the barrier is then known to be in
        // phase n or n+1 when the parity wait below starts.
        if (GROUP > 1 && nfill > 0) {
            unsigned have;
            do asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(have) : "r"(seen_u32 + slot * 4) : "memory");
            while (have < nfill);
        }
        wait_fill(slot, nfill & 1u);
        if (GROUP > 1 && lane == 0) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(seen_u32 + slot * 4), "r"(nfill + 1) : "memory");
        const double* tile = ring + slot * TILE_D;

        // ---- stage 1 on the MMA path: 4 accumulator chains (one per row block), consecutive DMMAs of a chain are
        // 4 issue slots = 64 cycles apart (latency 26)
        double S[4][2];
#pragma unroll
        for (int nb = 0; nb < 4; nb++) S[nb][0] = S[nb][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
            double xa[4];
#pragma unroll
            for (int nb = 0; nb < 4; nb++) xa[nb] = tile[ks * 4 * TILE_ROWS + xo[nb]];
#pragma unroll
            for (int nb = 0; nb < 4; nb++) dmma_884v(S[nb][0], S[nb][1], bfrag[ks], xa[nb]);
            __syncwarp();   // scheduling fence: keeps ptxas from gathering each chain's DMMAs back to back
        }
        // C fragments S^T[c = g][rows nb*8 + 2q + {0,1}] -> slab (the padding rows receive zeros)
#pragma unroll
        for (int nb = 0; nb < 4; nb++)
            if (g < SLAB_ROWS) *reinterpret_cast<double2*>(slab + fo[nb]) = make_double2(S[nb][0], S[nb][1]);
        __syncwarp();
        double xb[C1];
#pragma unroll
        for (int c = 0; c < C1; c++) xb[c] = slab[c * TILE_ROWS + ro[c & 3]];
        const double y = tile[y_off];
        // find_index (model/apop_probit.c:206-210): first category equal to y, at most C1 steps
        int idx = C1;
#pragma unroll
        for (int c = C1 - 1; c >= 0; c--) idx = (y == prm.cats[c]) ? c : idx;
        const bool in_range = t * TILE_ROWS + lane < n_rows;
        double wrow = 1.0;
        if (WEIGHTED) wrow = (double)__ldg(counts + (t * (unsigned long long)m_pad + replicate) * TILE_ROWS + cpos);
        const bool valid = in_range && (!WEIGHTED || wrow != 0.0);

        // ---- soft-max.  Fast path when every |x.b_c| < 708 (tested on the high words): log(1 + sum_c exp(xb_c))
        // directly, which is the reference's max + log(sum exp(xb_c - max) + exp(-max)) without the shift (equal to
        // rounding); any other row sends its warp through the reference-shaped shifted form
        bool fast = true;
#pragma unroll
        for (int c = 0; c < C1; c++) fast = fast && fm_exp_fast_ok(xb[c]);
        double e[C1], num = 0.0, inv, lsum, lbase = 0.0;
        {
            double sum = 1.0;                      // the numeraire's exp(0)
#pragma unroll
            for (int c = 0; c < C1; c++) {
                e[c] = exp_tab16(xb[c], exp_tab);  // garbage (never a trap) where the guard fails
                sum += e[c];
                num = (idx == c + 1) ? xb[c] : num;
            }
            lsum = sum;
            inv = fm_rcp(sum);
        }
        if (!__all_sync(0xffffffffu, fast)) {      // rare
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
            // when exp(-max) overflows the reference takes log-sum-exp = max - max: lbase + log(lsum) = 0 with lsum = 1
            const bool over = mx < -700.0;
            lsum = over ? 1.0 : sum + e0; lbase = over ? 0.0 : mx;
            // probabilities for the score: overflow of e0 means p_c = 0 for every free category
            inv = over ? 0.0 : fm_rcp(sum + e0);
        }
        inv = valid ? inv * wrow : 0.0;
#pragma unroll
        for (int c = 0; c < C1; c++) {
            const double ind = (valid && idx == c + 1) ? wrow : 0.0;
            slab[c * TILE_ROWS + ro[c & 3]] = fma(-e[c], inv, ind);   // the cell this lane read x.b_c from
        }
        __syncwarp();

        // ---- stage 3: A fragments R^T[c = g][rows nb*8 + 2q + {0,1}]; X fragments with the same rows
        double R[4][2];
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
            double2 r2 = make_double2(0.0, 0.0);
            if (g < SLAB_ROWS) r2 = *reinterpret_cast<const double2*>(slab + fo[nb]);
            R[nb][0] = r2.x; R[nb][1] = r2.y;
        }
        if (PROD) {
            ll.add(valid ? num - lbase : 0.0);
            // lsum in [1, 8 e^708): split off its exponent first so that the product cannot overflow
            const double ls = valid ? lsum : 1.0;
            const int eh = __double2hiint(ls);
            prod *= __hiloint2double((eh & 0x000fffff) | 0x3ff00000, __double2loint(ls));   // mantissa of lsum, in [1, 2)
            const int ph = __double2hiint(prod);                                              // prod in [1, 4)
            prod_e += ((eh >> 20) - 1023) + ((ph >> 20) - 1023);
            prod = __hiloint2double((ph & 0x000fffff) | 0x3ff00000, __double2loint(prod));
        } else
            ll.add(valid ? wrow * (num - (lbase + fm_log_normal(lsum))) : 0.0);
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {           // JB chains; consecutive DMMAs of a chain are JB slots (64 cycles) apart
            double2 x2[JB];
#pragma unroll
            for (int jb = 0; jb < JB; jb++) x2[jb] = *reinterpret_cast<const double2*>(tile + jb * 8 * TILE_ROWS + fo[nb]);
#pragma unroll
            for (int jb = 0; jb < JB; jb++) dmma_884v(G[jb][0], G[jb][1], R[nb][0], x2[jb].x);
#pragma unroll
            for (int jb = 0; jb < JB; jb++) dmma_884v(G[jb][0], G[jb][1], R[nb][1], x2[jb].y);
            __syncwarp();   // scheduling fence, as in stage 1
        }
        __syncwarp();                              // every lane is done with this slot before it is refilled
        if (t + (unsigned long long)SLOTS * ngrps < n_tiles) fill(t + (unsigned long long)SLOTS * ngrps, slot);   // item i + SLOTS: the partner's, when PAIRED
    }

    // ---- CTA reduction: lane (g, q) holds G^T[c = g][j = jb*8 + 2q + {0,1}]
    __syncthreads();                               // the rings are free: reuse as WARPS x NOUT scratch
    if (PROD) ll.add(-((double)prod_e * 6.93147180369123816490e-01 + ((double)prod_e * 1.90821492927058770002e-10 + fm_log_normal(prod))));
    const double llw = warp_sum(ll.value());
    if (lane == 0) smem[warp * NOUT] = llw;
    if (g < C1) {
#pragma unroll
        for (int jb = 0; jb < JB; jb++) {
            // (the y column of a K < KB tile went through the contraction as column K: dropped here)
            smem[warp * NOUT + 1 + (jb * 8 + 2 * q) * C1 + g] = (jb * 8 + 2 * q < K) ? G[jb][0] : 0.0;
            smem[warp * NOUT + 1 + (jb * 8 + 2 * q + 1) * C1 + g] = (jb * 8 + 2 * q + 1 < K) ? G[jb][1] : 0.0;
        }
    }
    __syncthreads();
    double* scratch = smem + WARPS * NOUT;
    for (int v = threadIdx.x; v < NOUT; v += THREADS) {
        double s = 0.0;
#pragma unroll
        for (int wv = 0; wv < WARPS; wv++) s += smem[wv * NOUT + v];
        partials[(size_t)blockIdx.x * NOUT + v] = s;
    }
    grid_finalize(partials, NOUT, ticket, out, fin, scratch);
}

template <int KB, bool PAIRED>
static constexpr size_t stream_smem() {
    using Cfg = StreamCfg<PAIRED>;
    return (size_t)(Cfg::WARPS / Cfg::GROUP) * (Cfg::SLOTS * (KB + 1) * TILE_ROWS + Cfg::GROUP * Cfg::SLAB_ROWS * TILE_ROWS) * sizeof(double);
}
// APOP_B200_LOGIT_STREAM=quad selects the 4-warp CTAs (12 warps per SM) for A/B runs
static bool stream_paired() {
    static const int v = [] { const char* e = getenv("APOP_B200_LOGIT_STREAM"); return (e && !strcmp(e, "quad")) ? 0 : 1; }();
    return v != 0;
}
int logit_stream_warps_per_block() { return stream_paired() ? StreamCfg<true>::WARPS : StreamCfg<false>::WARPS; }
template <int KB, int C1, bool PAIRED>
static cudaError_t stream_launch_p(const LogitLaunch& L, const LogitParams& prm) {
    using Cfg = StreamCfg<PAIRED>;
    static_assert((size_t)Cfg::WARPS * (1 + KB * C1) + (1 + KB * C1) <= stream_smem<KB, PAIRED>() / sizeof(double), "reduction scratch must fit the rings");
    static bool done = false;
    if (!done) {
        cudaFuncSetAttribute(logit_stream_kernel<KB, C1, false, PAIRED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stream_smem<KB, PAIRED>());
        cudaFuncSetAttribute(logit_stream_kernel<KB, C1, true, PAIRED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stream_smem<KB, PAIRED>());
        done = true;
    }
    if (L.counts)
        logit_stream_kernel<KB, C1, true, PAIRED><<<L.grid, Cfg::WARPS * 32, stream_smem<KB, PAIRED>(), L.stream>>>(L.tiles, L.n_rows, L.n_tiles, L.k, L.cols, L.counts, L.m_pad, L.replicate, prm, L.partials, L.ticket, L.out, L.fin);
    else
        logit_stream_kernel<KB, C1, false, PAIRED><<<L.grid, Cfg::WARPS * 32, stream_smem<KB, PAIRED>(), L.stream>>>(L.tiles, L.n_rows, L.n_tiles, L.k, L.cols, nullptr, 0, 0, prm, L.partials, L.ticket, L.out, L.fin);
    return cudaGetLastError();
}
template <int KB, int C1>
static cudaError_t stream_launch_one(const LogitLaunch& L, const LogitParams& prm) {
    return stream_paired() ? stream_launch_p<KB, C1, true>(L, prm) : stream_launch_p<KB, C1, false>(L, prm);
}
template <int KB, int C1, bool PAIRED>
static int stream_bps_p() {
    cudaFuncSetAttribute(logit_stream_kernel<KB, C1, false, PAIRED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stream_smem<KB, PAIRED>());
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, logit_stream_kernel<KB, C1, false, PAIRED>, StreamCfg<PAIRED>::WARPS * 32, stream_smem<KB, PAIRED>());
    return nb;
}
template <int KB, int C1>
static int stream_bps_one() { return stream_paired() ? stream_bps_p<KB, C1, true>() : stream_bps_p<KB, C1, false>(); }

#define APB_STREAM_C(KB)                       \
    switch (c1) {                              \
    case 2: return APB_OP(KB, 2);              \
    case 3: return APB_OP(KB, 3);              \
    case 4: return APB_OP(KB, 4);              \
    case 5: return APB_OP(KB, 5);              \
    case 6: return APB_OP(KB, 6);              \
    case 7: return APB_OP(KB, 7);              \
    default: break;                            \
    }
cudaError_t logit_stream_launch(int c1, const LogitLaunch& L, const LogitParams& prm) {
    if (!L.swz) return cudaErrorInvalidValue;   // this kernel reads the swizzled resident form only
#define APB_OP(KB, C) stream_launch_one<KB, C>(L, prm)
    switch (logit_kb(L.k)) {
    case 8: APB_STREAM_C(8) break;
    case 16: APB_STREAM_C(16) break;
    case 24: APB_STREAM_C(24) break;
    case 32: APB_STREAM_C(32) break;
    }
#undef APB_OP
    return cudaErrorInvalidValue;
}
int logit_stream_blocks_per_sm(int k, int c1) {
#define APB_OP(KB, C) stream_bps_one<KB, C>()
    switch (logit_kb(k)) {
    case 8: APB_STREAM_C(8) break;
    case 16: APB_STREAM_C(16) break;
    case 24: APB_STREAM_C(24) break;
    case 32: APB_STREAM_C(32) break;
    }
#undef APB_OP
    return 0;
}

}  // namespace apb
