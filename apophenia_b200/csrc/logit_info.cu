// logit_info.cu -- observed information of the multinomial logit,
//     -d2LL/dB dB' = sum_i (x_i x_i') (x) (diag p_i - p_i p_i'),      P x P with P = k (C-1),
// in place of the reference's numeric Hessian (apop_model_hessian, apop_mle.m4.c:187-223: the numeric
// gradient of every score component, >= 4 P score passes -- and the multinomial logit has no analytic
// score there either, model/apop_probit.c:244-289, so >= 16 P^2 likelihood passes).
//
// Block form: for a category pair (a, b) the k x k block  H_ab = X' diag(w_ab) X  with
// w_ab,i = [a == b] p_ia - p_ia p_ib  is a weighted Gram matrix, and H_ab = H_ba, H_ab = H_ab'.  So the
// (C-1) C / 2 pairs with a <= b, upper block triangle each, carry everything: 28 x 528 multiply-adds per
// row at k = 32, C = 8 instead of 224 x 225 / 2.  One launch handles NP = 4 pairs (4 x 10 x 2 accumulator
// registers per lane), ceil(pairs / 4) launches per matrix.  Per 32-row tile and warp:
//   stage   the tile arrives in shared memory by 16-byte cp.async (padded column stride, as xtwx_kernel.cu);
//   weights lane = row: x.B_c against the constant bank, the shifted soft-max of one_logit_row
//           (model/apop_probit.c:212-229), then the NP pair weights into a 4 x 32 slab;
//   DMMA    fragment (jb, ks) = X[rows ks*4 + q][columns jb*8 + g] is the B operand and, scaled by the pair
//           weight of its row, the A operand: 8 k-steps x NP x JB (JB+1) / 2 DMMA.8x8x4 in NP x 10
//           independent chains.
// Bound: the FP64 tensor pipe (DMMA); X and y are read once per launch.
#include "logit_kernel.cuh"
#include "fastmath.cuh"

namespace apb {

__device__ __forceinline__ void dmma_info(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1)
        : "d"(a), "d"(b));
}
constexpr int INFO_CS = 36;   // padded column stride of a staged tile (doubles)

// BULK: the resident tiles are in the swizzled form (layout.cu) and arrive by one TMA bulk copy per tile into a
// two-slot ring per warp, the next tile in flight while this one is contracted; the swizzle makes the 8-byte fragment
// loads below conflict free without padding (the four columns g = 0..3 of a half-warp land in different quarters of a
// bank row).  Otherwise (plain tiles): 16-byte cp.async into a padded buffer, no overlap inside the warp.
template <int KB, int C1, bool BULK>
__global__ void __launch_bounds__(LOGIT_INFO_THREADS, 2)
logit_info_kernel(const double* __restrict__ tiles, const unsigned long long n_rows, const unsigned long long n_tiles,
                  const int K, const int cols, const int swz, const __grid_constant__ LogitParams prm, const __grid_constant__ LogitInfoPairs pairs,
                  double* __restrict__ partials, unsigned int* __restrict__ ticket, double* __restrict__ out,
                  const __grid_constant__ FinalizeCtx fin) {
    constexpr int WARPS = LOGIT_INFO_THREADS / 32;
    constexpr int NP = LOGIT_INFO_NP;
    constexpr int JB = KB / 8;
    constexpr int NTRI = JB * (JB + 1) / 2;
    constexpr int NOUT = NP * KB * KB;
    constexpr int CS = BULK ? TILE_ROWS : INFO_CS;                       // column stride of a staged tile
    constexpr int NSLOT = BULK ? 2 : 1;
    constexpr int TILE_D = BULK ? (KB + 1) * TILE_ROWS : KB * INFO_CS;   // bulk: X and y as resident; else X only
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const unsigned long long gwarp = (unsigned long long)blockIdx.x * WARPS + warp;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * WARPS;
    double* ring = smem + warp * (NSLOT * TILE_D + NP * 32 + 8 * 32);
    double* wbuf = ring + NSLOT * TILE_D;
    double* pbuf = wbuf + NP * 32;                 // p[c][row]: the pair weights pick their two categories by address

    double acc[NP][NTRI][2];
#pragma unroll
    for (int i = 0; i < NP; i++)
#pragma unroll
        for (int b = 0; b < NTRI; b++) acc[i][b][0] = acc[i][b][1] = 0.0;
    // columns K..KB-1 are never staged (bulk: K+1..KB; column K receives y and contracts into entries nobody reads): zero them once
    for (int i = lane; i < NSLOT * TILE_D; i += 32)
        if (!BULK || (i % TILE_D) / TILE_ROWS > K) ring[i] = 0.0;
    __shared__ __align__(8) unsigned long long bars[WARPS * 2];
    if (BULK) {
        if (lane < 2) mbar_init(bars + warp * 2 + lane, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const unsigned bar_u32 = smem_u32(bars + warp * 2), ring_u32 = smem_u32(ring);
    const unsigned tile_bytes = (unsigned)(K + 1) * TILE_ROWS * sizeof(double);
    auto fill = [&](unsigned long long tt, int slot) {
        if (lane == 0) {
            const unsigned bar = bar_u32 + slot * 8, dst = ring_u32 + slot * (TILE_D * 8);
            const double* src = tiles + tt * (unsigned long long)(cols * TILE_ROWS);
            fence_proxy_async();
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tile_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(src), "r"(tile_bytes), "r"(bar) : "memory");
        }
    };
    // swizzled positions: row r of column j at j*32 + (r ^ rx(j)), rx(j) = 8 (j & 1) + 4 ((j >> 1) & 1); rx(.. + g) = rxg
    const int rxg = BULK ? (((g & 1) << 3) | ((g & 2) << 1)) : 0;

    unsigned it = 0;
    if (BULK && gwarp < n_tiles) fill(gwarp, 0);
    for (unsigned long long t = gwarp; t < n_tiles; t += nwarps, it++) {
        double* tile = ring + (BULK ? (int)(it & 1u) : 0) * TILE_D;
        if (BULK) {
            if (t + nwarps < n_tiles) fill(t + nwarps, (int)((it + 1) & 1u));   // that slot was consumed in the last iteration
            const unsigned bar = bar_u32 + (it & 1u) * 8, parity = (it >> 1) & 1u;
            asm volatile("{\n\t.reg .pred p;\n\tWAIT_TILE:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra TILE_READY;\n\tbra WAIT_TILE;\n\tTILE_READY:\n\t}"
                         ::"r"(bar), "r"(parity) : "memory");
        } else {   // stage the K columns of X: 16-byte chunk `lane + 32 i` lies in resident column (lane >> 4) + 2 i
            const double* src = tiles + t * (unsigned long long)(cols * TILE_ROWS) + lane * 2;
            // (swizzled resident form, layout.cu: chunk c of column j holds the rows of chunk c ^ sw(j) -- undone on the
            // way in; the column of copy i is (lane >> 4) + 2 i, so sw = ((lane >> 4) << 2) | ((i & 1) << 1))
            const int ch = (lane & 15) ^ ((swz & (lane >> 4)) << 2);
            const unsigned sa0 = smem_u32(tile + (lane >> 4) * INFO_CS + ch * 2);
            const unsigned sa1 = smem_u32(tile + (lane >> 4) * INFO_CS + (ch ^ (swz << 1)) * 2);
#pragma unroll
            for (int i = 0; i < KB / 2; i++) {
                const int col = (lane >> 4) + 2 * i;
                if (col < K)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(((i & 1) ? sa1 : sa0) + i * 2 * INFO_CS * 8), "l"(src + i * 64));
            }
            asm volatile("cp.async.commit_group;");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
        }
        // ---- pair weights, lane = row
        {
            double xb[C1];
#pragma unroll
            for (int c = 0; c < C1; c++) xb[c] = 0.0;
#pragma unroll
            for (int j = 0; j < KB; j++) {
                const double x = tile[j * CS + (BULK ? (lane ^ (((j & 1) << 3) | ((j & 2) << 1))) : lane)];
#pragma unroll
                for (int c = 0; c < C1; c++) xb[c] = fma(x, prm.B[j * C1 + c], xb[c]);
            }
            double mx = 0.0;   // the numeraire's x.b = 0 takes part in the shift
#pragma unroll
            for (int c = 0; c < C1; c++) mx = fmax(mx, xb[c]);
            double den = fm_exp(-mx), p[C1];
#pragma unroll
            for (int c = 0; c < C1; c++) { p[c] = fm_exp(xb[c] - mx); den += p[c]; }
            const bool valid = t * TILE_ROWS + lane < n_rows;
            const double inv = valid ? fm_rcp(den) : 0.0;
#pragma unroll
            for (int c = 0; c < C1; c++) pbuf[c * 32 + lane] = p[c] * inv;
#pragma unroll
            for (int i = 0; i < NP; i++) {
                // a padding pair (a < 0) contributes nothing; each lane reads back only what it wrote
                const int a = pairs.a[i] < 0 ? 0 : pairs.a[i], b = pairs.b[i] < 0 ? 0 : pairs.b[i];
                const double pa = pbuf[a * 32 + lane], pb = pbuf[b * 32 + lane];
                wbuf[i * 32 + lane] = (pairs.a[i] < 0) ? 0.0 : ((a == b) ? pa * (1.0 - pa) : -pa * pb);
            }
        }
        __syncwarp();
        // ---- DMMA: rows are the contraction dimension
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
            double xf[JB];
#pragma unroll
            for (int jb = 0; jb < JB; jb++) xf[jb] = tile[(jb * 8 + g) * CS + ((ks * 4 + q) ^ rxg)];
#pragma unroll
            for (int i = 0; i < NP; i++) {
                const double wv = wbuf[i * 32 + ks * 4 + q];
                int b = 0;
#pragma unroll
                for (int ab = 0; ab < JB; ab++) {
                    const double a = xf[ab] * wv;
#pragma unroll
                    for (int cb = ab; cb < JB; cb++, b++) dmma_info(acc[i][b][0], acc[i][b][1], a, xf[cb]);
                }
            }
        }
        __syncwarp();   // the tile buffer is about to be overwritten
    }

    // ---- CTA reduction: expand the block triangles into NP full KB x KB blocks
    __syncthreads();
    double* red = smem;
    for (int v = threadIdx.x; v < NOUT; v += LOGIT_INFO_THREADS) red[v] = 0.0;
    __syncthreads();
    for (int wv_ = 0; wv_ < WARPS; wv_++) {
        if (warp == wv_) {
#pragma unroll
            for (int i = 0; i < NP; i++) {
                int b = 0;
#pragma unroll
                for (int ab = 0; ab < JB; ab++)
#pragma unroll
                    for (int cb = ab; cb < JB; cb++, b++)
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int a = ab * 8 + g, c = cb * 8 + 2 * q + e;
                            red[i * KB * KB + a * KB + c] += acc[i][b][e];
                            if (cb != ab) red[i * KB * KB + c * KB + a] += acc[i][b][e];
                        }
            }
        }
        __syncthreads();
    }
    for (int v = threadIdx.x; v < NOUT; v += LOGIT_INFO_THREADS) partials[(size_t)blockIdx.x * NOUT + v] = red[v];
    __syncthreads();
    grid_finalize(partials, NOUT, ticket, out, fin, red);
}

template <int KB, bool BULK>
static constexpr size_t info_smem() {
    const size_t ring = (size_t)(LOGIT_INFO_THREADS / 32) * ((BULK ? 2 * (KB + 1) * TILE_ROWS : KB * INFO_CS) + LOGIT_INFO_NP * 32 + 8 * 32) * sizeof(double);
    const size_t red = (size_t)LOGIT_INFO_NP * KB * KB * sizeof(double);
    return ring > red ? ring : red;
}
template <int KB, int C1, bool BULK>
static cudaError_t info_launch_b(const LogitLaunch& L, const LogitParams& prm, const LogitInfoPairs& pairs) {
    static bool done = false;
    if (!done) { cudaFuncSetAttribute(logit_info_kernel<KB, C1, BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)info_smem<KB, BULK>()); done = true; }
    logit_info_kernel<KB, C1, BULK><<<L.grid, LOGIT_INFO_THREADS, info_smem<KB, BULK>(), L.stream>>>(L.tiles, L.n_rows, L.n_tiles, L.k, L.cols, L.swz, prm, pairs,
                                                                                                 L.partials, L.ticket, L.out, L.fin);
    return cudaGetLastError();
}
// swizzled tiles take the bulk-copy ring, plain tiles the cp.async form
template <int KB, int C1>
static cudaError_t info_launch_one(const LogitLaunch& L, const LogitParams& prm, const LogitInfoPairs& pairs) {
    return L.swz ? info_launch_b<KB, C1, true>(L, prm, pairs) : info_launch_b<KB, C1, false>(L, prm, pairs);
}
template <int KB, int C1>
static int info_bps_one() {
    int nb = 0, nb2 = 0;
    cudaFuncSetAttribute(logit_info_kernel<KB, C1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)info_smem<KB, true>());
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, logit_info_kernel<KB, C1, true>, LOGIT_INFO_THREADS, info_smem<KB, true>());
    cudaFuncSetAttribute(logit_info_kernel<KB, C1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)info_smem<KB, false>());
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, logit_info_kernel<KB, C1, false>, LOGIT_INFO_THREADS, info_smem<KB, false>());
    return nb < nb2 ? nb : nb2;
}

#define APB_INFO_C(KB)                         \
    switch (c1) {                              \
    case 2: return APB_OP(KB, 2);              \
    case 3: return APB_OP(KB, 3);              \
    case 4: return APB_OP(KB, 4);              \
    case 5: return APB_OP(KB, 5);              \
    case 6: return APB_OP(KB, 6);              \
    case 7: return APB_OP(KB, 7);              \
    default: break;                            \
    }
cudaError_t logit_info_launch(int c1, const LogitLaunch& L, const LogitParams& prm, const LogitInfoPairs& pairs) {
#define APB_OP(KB, C) info_launch_one<KB, C>(L, prm, pairs)
    switch (logit_kb(L.k)) {
    case 8: APB_INFO_C(8) break;
    case 16: APB_INFO_C(16) break;
    case 24: APB_INFO_C(24) break;
    case 32: APB_INFO_C(32) break;
    }
#undef APB_OP
    return cudaErrorInvalidValue;
}
int logit_info_blocks_per_sm(int k, int c1) {
#define APB_OP(KB, C) info_bps_one<KB, C>()
    switch (logit_kb(k)) {
    case 8: APB_INFO_C(8) break;
    case 16: APB_INFO_C(16) break;
    case 24: APB_INFO_C(24) break;
    case 32: APB_INFO_C(32) break;
    }
#undef APB_OP
    return 0;
}

}  // namespace apb
