// logit_kernel.cu -- multinomial logit: fused X.B -> log-sum-exp -> LL + analytic score.
//
// LL row: one_logit_row (model/apop_probit.c:212-229): num - (max + log(sum_c
// exp(xb_c - max) + exp(-max))), the numeraire category carrying beta_0 = 0.
// Score (the reference keeps it commented out, :244-289; new vtable entry):
//   dLL/dB[j][c] = sum_i x_ij (1{idx_i == c+1} - p_ic),  output index j*(C-1)+c.
//
// P = k (C-1) accumulators (224 in BASELINE config 2) do not fit a lane=row
// thread, so each 32-row tile is processed in two phases by its warp:
//   A  lane = row:    x row in registers, C-1 dot products against B in the
//                     constant bank, soft-max, residuals r_ic;  x and r are
//                     parked in the warp's shared-memory slab (conflict-free:
//                     row stride KB+1 doubles; r stored category-major);
//   B  lane = column: G[lane][c] += sum_i x_i,lane r_ic, x read conflict-free,
//                     r read as 16-byte broadcasts (two rows per load).
// X and y are read from HBM exactly once: 8(k+1) bytes per row.  Per row the
// kernel issues ~2 k (C-1) DFMA + (C-1) exp + 1 log, so at k = 32, C = 8 the
// FP64 pipe is about level with HBM.
#include "logit_kernel.cuh"
#include "fastmath.cuh"

namespace apb {

template <int KB, int C1>
__global__ void __launch_bounds__(LOGIT_THREADS)
logit_fused_kernel(const double* __restrict__ tiles, const unsigned long long n_rows,
                   const unsigned long long n_tiles, const int K, const __grid_constant__ LogitParams prm,
                   double* __restrict__ partials, unsigned int* __restrict__ ticket,
                   double* __restrict__ out) {
    constexpr int WARPS = LOGIT_THREADS / 32;
    constexpr int XS = KB + 1;  // row stride of the parked x rows (odd: conflict-free)
    constexpr int NOUT = 1 + KB * C1;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* sx = smem + warp * (32 * XS + C1 * 32);
    double* sr = sx + 32 * XS;
    const unsigned long long gwarp = (unsigned long long)blockIdx.x * WARPS + warp;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * WARPS;
    const int cols = K + 1;

    double G[C1];
#pragma unroll
    for (int c = 0; c < C1; c++) G[c] = 0.0;
    KSum ll;

    for (unsigned long long t = gwarp; t < n_tiles; t += nwarps) {
        const double* base = tiles + t * (unsigned long long)(cols * TILE_ROWS) + lane;
        // ---- phase A: lane = row
        double x[KB];
#pragma unroll
        for (int j = 0; j < KB; j++) x[j] = (j < K) ? ld_stream(base + j * TILE_ROWS) : 0.0;
        const double y = ld_stream(base + K * TILE_ROWS);
        double xb[C1];
#pragma unroll
        for (int c = 0; c < C1; c++) xb[c] = 0.0;
#pragma unroll
        for (int j = 0; j < KB; j++)
#pragma unroll
            for (int c = 0; c < C1; c++) xb[c] = fma(x[j], prm.B[j * C1 + c], xb[c]);
        // find_index (model/apop_probit.c:206-210): first category equal to y, at most C1 steps
        int idx = C1;
#pragma unroll
        for (int c = C1 - 1; c >= 0; c--) idx = (y == prm.cats[c]) ? c : idx;
        double mx = xb[0];
#pragma unroll
        for (int c = 1; c < C1; c++) mx = fmax(mx, xb[c]);
        double e[C1], sum = 0.0, num = 0.0;
#pragma unroll
        for (int c = 0; c < C1; c++) {
            e[c] = fm_exp(xb[c] - mx);
            sum += e[c];
            num = (idx == c + 1) ? xb[c] : num;
        }
        const double e0 = fm_exp(-mx);  // the numeraire's exp(0 - max)
        const bool valid = t * TILE_ROWS + lane < n_rows;
        // when exp(-max) overflows the reference takes log-sum-exp = max - max
        const double lse = (mx < -700.0) ? 0.0 : mx + fm_log(sum + e0);
        ll.add(valid ? num - lse : 0.0);
        // probabilities for the score: overflow of e0 means p_c = 0 for every free category
        const double inv = (mx < -700.0) ? 0.0 : fm_rcp(sum + e0);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < KB; j++) sx[lane * XS + j] = x[j];
#pragma unroll
        for (int c = 0; c < C1; c++) {
            const double r = ((idx == c + 1) ? 1.0 : 0.0) - e[c] * inv;
            sr[c * 32 + lane] = valid ? r : 0.0;
        }
        __syncwarp();
        // ---- phase B: lane = column
#pragma unroll 4
        for (int i = 0; i < 32; i += 2) {
            const double xa = sx[i * XS + lane];
            const double xc = sx[(i + 1) * XS + lane];
#pragma unroll
            for (int c = 0; c < C1; c++) {
                const double2 r = *reinterpret_cast<const double2*>(sr + c * 32 + i);
                G[c] = fma(xa, r.x, G[c]);
                G[c] = fma(xc, r.y, G[c]);
            }
        }
    }

    // ---- CTA reduction: G is already per column; sum over the CTA's warps
    __syncthreads();
    double* red = smem;  // reuse: WARPS x NOUT doubles (fits: NOUT <= 32*XS + 32*C1)
    const double llw = warp_sum(ll.value());
    if (lane == 0) red[warp * NOUT] = llw;
    if (lane < KB) {
#pragma unroll
        for (int c = 0; c < C1; c++) red[warp * NOUT + 1 + lane * C1 + c] = G[c];
    }
    __syncthreads();
    for (int v = threadIdx.x; v < NOUT; v += LOGIT_THREADS) {
        double s = 0.0;
#pragma unroll
        for (int wv = 0; wv < WARPS; wv++) s += red[wv * NOUT + v];
        partials[(size_t)blockIdx.x * NOUT + v] = s;
    }
    __shared__ unsigned int is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int v = threadIdx.x; v < NOUT; v += LOGIT_THREADS) {
        double s = 0.0, c = 0.0;
        for (unsigned b = 0; b < gridDim.x; b++) {
            double er;
            two_sum(s, __ldcg(partials + (size_t)b * NOUT + v), s, er);
            c += er;
        }
        out[v] = s + c;
    }
    if (threadIdx.x == 0) *ticket = 0;
}

template <int KB, int C1>
static size_t logit_smem() { return (size_t)(LOGIT_THREADS / 32) * (32 * (KB + 1) + C1 * 32) * sizeof(double); }

template <int KB, int C1>
static cudaError_t launch_one(const LogitLaunch& L, const LogitParams& prm) {
    static bool attr_done = false;
    const size_t sm = logit_smem<KB, C1>();
    if (!attr_done) {
        cudaFuncSetAttribute(logit_fused_kernel<KB, C1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        attr_done = true;
    }
    logit_fused_kernel<KB, C1><<<L.grid, LOGIT_THREADS, sm, L.stream>>>(L.tiles, L.n_rows, L.n_tiles, L.k, prm,
                                                                        L.partials, L.ticket, L.out);
    return cudaGetLastError();
}
template <int KB, int C1>
static int bps_one() {
    const size_t sm = logit_smem<KB, C1>();
    cudaFuncSetAttribute(logit_fused_kernel<KB, C1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, logit_fused_kernel<KB, C1>, LOGIT_THREADS, sm);
    return nb;
}

#define APB_LOGIT_C(KB)                                   \
    switch (c1) {                                         \
    case 2: return APB_OP(KB, 2);                         \
    case 3: return APB_OP(KB, 3);                         \
    case 4: return APB_OP(KB, 4);                         \
    case 5: return APB_OP(KB, 5);                         \
    case 6: return APB_OP(KB, 6);                         \
    case 7: return APB_OP(KB, 7);                         \
    default: break;                                       \
    }
int logit_kb(int k) { return k <= 8 ? 8 : k <= 16 ? 16 : k <= 24 ? 24 : 32; }

cudaError_t logit_launch(int c1, const LogitLaunch& L, const LogitParams& prm) {
#define APB_OP(KB, C) launch_one<KB, C>(L, prm)
    switch (logit_kb(L.k)) {
    case 8: APB_LOGIT_C(8) break;
    case 16: APB_LOGIT_C(16) break;
    case 24: APB_LOGIT_C(24) break;
    case 32: APB_LOGIT_C(32) break;
    }
#undef APB_OP
    return cudaErrorInvalidValue;
}
int logit_blocks_per_sm(int k, int c1) {
#define APB_OP(KB, C) bps_one<KB, C>()
    switch (logit_kb(k)) {
    case 8: APB_LOGIT_C(8) break;
    case 16: APB_LOGIT_C(16) break;
    case 24: APB_LOGIT_C(24) break;
    case 32: APB_LOGIT_C(32) break;
    }
#undef APB_OP
    return 0;
}

}  // namespace apb
