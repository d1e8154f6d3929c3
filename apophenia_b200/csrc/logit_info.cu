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

template <int KB, int C1>
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
    constexpr int TILE_D = KB * INFO_CS;           // X only: the information does not depend on the outcomes
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const unsigned long long gwarp = (unsigned long long)blockIdx.x * WARPS + warp;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * WARPS;
    double* tile = smem + warp * (TILE_D + NP * 32 + 8 * 32);
    double* wbuf = tile + TILE_D;
    double* pbuf = wbuf + NP * 32;                 // p[c][row]: the pair weights pick their two categories by address

    double acc[NP][NTRI][2];
#pragma unroll
    for (int i = 0; i < NP; i++)
#pragma unroll
        for (int b = 0; b < NTRI; b++) acc[i][b][0] = acc[i][b][1] = 0.0;
    // columns K..KB-1 are never staged: zero them once
    for (int i = lane; i < TILE_D; i += 32) tile[i] = 0.0;
    __syncwarp();

    for (unsigned long long t = gwarp; t < n_tiles; t += nwarps) {
        {   // stage the K columns of X: 16-byte chunk `lane + 32 i` lies in resident column (lane >> 4) + 2 i
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
                const double x = tile[j * INFO_CS + lane];
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
            for (int jb = 0; jb < JB; jb++) xf[jb] = tile[(jb * 8 + g) * INFO_CS + ks * 4 + q];
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

template <int KB>
static constexpr size_t info_smem() {
    const size_t ring = (size_t)(LOGIT_INFO_THREADS / 32) * (KB * INFO_CS + LOGIT_INFO_NP * 32 + 8 * 32) * sizeof(double);
    const size_t red = (size_t)LOGIT_INFO_NP * KB * KB * sizeof(double);
    return ring > red ? ring : red;
}
template <int KB, int C1>
static cudaError_t info_launch_one(const LogitLaunch& L, const LogitParams& prm, const LogitInfoPairs& pairs) {
    static bool done = false;
    if (!done) { cudaFuncSetAttribute(logit_info_kernel<KB, C1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)info_smem<KB>()); done = true; }
    logit_info_kernel<KB, C1><<<L.grid, LOGIT_INFO_THREADS, info_smem<KB>(), L.stream>>>(L.tiles, L.n_rows, L.n_tiles, L.k, L.cols, L.swz, prm, pairs,
                                                                                      L.partials, L.ticket, L.out, L.fin);
    return cudaGetLastError();
}
template <int KB, int C1>
static int info_bps_one() {
    cudaFuncSetAttribute(logit_info_kernel<KB, C1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)info_smem<KB>());
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, logit_info_kernel<KB, C1>, LOGIT_INFO_THREADS, info_smem<KB>());
    return nb;
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
