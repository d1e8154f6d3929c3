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
#include <stdlib.h>
#include <string.h>

namespace apb {

template <int KB, int C1>
__global__ void __launch_bounds__(LOGIT_THREADS)
logit_fused_kernel(const double* __restrict__ tiles, const unsigned long long n_rows,
                   const unsigned long long n_tiles, const int K, const __grid_constant__ LogitParams prm,
                   double* __restrict__ partials, unsigned int* __restrict__ ticket,
                   double* __restrict__ out, const __grid_constant__ FinalizeCtx fin) {
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
    grid_finalize(partials, NOUT, ticket, out, fin, red);
}


// ---------------------------------------------------------------------------------------
// Tensor-core variant (APOP_B200_LOGIT_DMMA=1; round-1 measurement at 20M x 32, C = 8: 1.38-1.40 ms,
// level with the FP64-pipe kernel's 1.40 ms -- DMMA pipe ~40 % busy, bounded by the latency of its
// short dependent MMA chains at 8-16 warps per SM): the two contractions of a tile run as FP64 DMMA
// (mma.sync m8n8k4), which executes on the tensor pipe next to the FP64 pipe, leaving the
// latter to the soft-max.  The (C-1 <= 7) free categories are the M = 8 rows of the MMA:
//   stage 1  S^T[c][row] = B^T[c][j] . X^T[j][row]      (K = columns)
//   soft-max with lane = row: the 8 x 32 block of x.beta goes through a per-warp shared-memory
//            slab, so each row's exps, log and reciprocal are evaluated exactly once (FP64
//            pipe), and the residuals return through the slab as stage-3 fragments
//   stage 3  G^T[c][j] += R^T[c][row] . X[row][j]       (K = rows, slots = rows 2q+e)
// Same fragment trick as batched_kernel.cu: the C fragment of stage 1 is already the A
// fragment of stage 3.  X is read once from HBM (stage-1 fragments, 64-byte segments of the
// tile columns) and once more from L1 (stage-3 fragments).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ double group_max(double v) {   // over the 8 lanes sharing q = lane%4
    v = fmax(v, shfl_xor_d(v, 4)); v = fmax(v, shfl_xor_d(v, 8)); return fmax(v, shfl_xor_d(v, 16));
}
__device__ __forceinline__ double group_sum(double v) {
    v += shfl_xor_d(v, 4); v += shfl_xor_d(v, 8); return v + shfl_xor_d(v, 16);
}

template <int KB, int NBUF>
__global__ void __launch_bounds__(LOGIT_THREADS, NBUF == 1 ? 4 : 2)
logit_dmma_kernel(const double* __restrict__ tiles, const unsigned long long n_rows,
                  const unsigned long long n_tiles, const int K, const int C1,
                  const __grid_constant__ LogitParams prm, double* __restrict__ partials,
                  unsigned int* __restrict__ ticket, double* __restrict__ out,
                  const __grid_constant__ FinalizeCtx fin) {
    constexpr int WARPS = LOGIT_THREADS / 32;
    constexpr int KS = KB / 4, JB = KB / 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const unsigned long long gwarp = (unsigned long long)blockIdx.x * WARPS + warp;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * WARPS;
    const int cols = K + 1;
    const bool cat_ok = g < C1;   // this lane's category slot is a real category (g+1)
    // per-warp double buffer: the next tile streams in with 16-byte cp.async while this one
    // is being processed; columns are padded to CS doubles so that the stage-1 fragment reads
    // (4 columns x 8 rows per request) are bank-conflict free
    extern __shared__ __align__(16) double smem[];
    double* ring = smem + warp * (NBUF * LOGIT_TILE_DOUBLES + 8 * LOGIT_XS);
    double* xs = ring + NBUF * LOGIT_TILE_DOUBLES;   // 8 x LOGIT_XS exchange slab
    // chunk c = lane + 32 i of a tile lies in column (lane >> 4) + 2 i at 16-byte slot lane & 15
    const int pf_dst = (lane >> 4) * LOGIT_CS + (lane & 15) * 2, pf_src = lane * 2;
    auto prefetch = [&](unsigned long long tt, int buf) {
        const double* src = tiles + tt * (unsigned long long)(cols * TILE_ROWS) + pf_src;
        const unsigned sa = (unsigned)__cvta_generic_to_shared(ring + buf * LOGIT_TILE_DOUBLES + pf_dst);
#pragma unroll
        for (int i = 0; i < (KB + 2) / 2; i++) {
            const int col = (lane >> 4) + 2 * i;
            if (col < cols)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa + i * 2 * LOGIT_CS * 8), "l"(src + i * 64));
        }
        asm volatile("cp.async.commit_group;");
    };

    double bfrag[KS];              // B^T[c = g][j = ks*4 + q]
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
        const int j = ks * 4 + q;
        bfrag[ks] = (cat_ok && j < K) ? prm.B[j * C1 + g] : 0.0;
    }
    double cat[LOGIT_MAX_C1];
#pragma unroll
    for (int c = 0; c < LOGIT_MAX_C1; c++) cat[c] = prm.cats[c];

    double G[JB][2], G2[JB][2];   // even / odd contraction slots accumulate separately
#pragma unroll
    for (int jb = 0; jb < JB; jb++) G[jb][0] = G[jb][1] = G2[jb][0] = G2[jb][1] = 0.0;
    KSum ll;

    int cur = 0;
    if (NBUF == 2 && gwarp < n_tiles) prefetch(gwarp, 0);
    for (unsigned long long t = gwarp; t < n_tiles; t += nwarps) {
        if (NBUF == 1) {   // single buffer, twice the warps: latency is hidden across warps
            prefetch(t, 0);
            asm volatile("cp.async.wait_group 0;");
        } else if (t + nwarps < n_tiles) {
            prefetch(t + nwarps, cur ^ 1);
            asm volatile("cp.async.wait_group 1;");
        } else
            asm volatile("cp.async.wait_group 0;");
        __syncwarp();
        const double* tile = ring + cur * LOGIT_TILE_DOUBLES;
        // ---- stage 1: all fragment loads of the tile are issued before the first MMA
        // (KS*4 x 256 B per warp in flight), like the lane = row kernels
        double xf[KS][4];
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
            const int j = ks * 4 + q;
#pragma unroll
            for (int nb = 0; nb < 4; nb++) xf[ks][nb] = (j < K) ? tile[j * LOGIT_CS + nb * 8 + g] : 0.0;
        }
        double S[4][2], S2[4][2];   // two accumulator sets: 8 independent MMA chains of depth KS/2
#pragma unroll
        for (int nb = 0; nb < 4; nb++) S[nb][0] = S[nb][1] = S2[nb][0] = S2[nb][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ks += 2)
#pragma unroll
            for (int nb = 0; nb < 4; nb++) {
                dmma_8x8x4(S[nb][0], S[nb][1], bfrag[ks], xf[ks][nb]);
                dmma_8x8x4(S2[nb][0], S2[nb][1], bfrag[ks + 1], xf[ks + 1][nb]);
            }
#pragma unroll
        for (int nb = 0; nb < 4; nb++) { S[nb][0] += S2[nb][0]; S[nb][1] += S2[nb][1]; }
        // ---- soft-max with lane = row: the C fragments go through the warp's exchange slab
        // (category-major, stride LOGIT_XS), every row's exp / log / reciprocal is evaluated
        // exactly once, and the residuals come back as the A fragments of stage 3
        __syncwarp();
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
            xs[g * LOGIT_XS + nb * 8 + 2 * q] = S[nb][0];
            xs[g * LOGIT_XS + nb * 8 + 2 * q + 1] = S[nb][1];
        }
        __syncwarp();
        {
            const double y = tile[K * LOGIT_CS + lane];
            double xbv[LOGIT_MAX_C1];
#pragma unroll
            for (int c = 0; c < LOGIT_MAX_C1; c++) xbv[c] = (c < C1) ? xs[c * LOGIT_XS + lane] : -1e300;
            // find_index (model/apop_probit.c:206-210): first category value equal to y, else C1
            int idx = C1;
#pragma unroll
            for (int c = LOGIT_MAX_C1 - 1; c >= 0; c--) idx = (c < C1 && y == cat[c]) ? c : idx;
            double mx = xbv[0];
#pragma unroll
            for (int c = 1; c < LOGIT_MAX_C1; c++) mx = fmax(mx, xbv[c]);
            double ev[LOGIT_MAX_C1], sum = 0.0, num = 0.0;
#pragma unroll
            for (int c = 0; c < LOGIT_MAX_C1; c++) {
                ev[c] = fm_exp(xbv[c] - mx);       // padded categories: exp(-1e300) ~ 1e-308, vanishes in the sum
                sum += ev[c];
                num = (idx == c + 1) ? xbv[c] : num;
            }
            const double e0 = fm_exp(-mx);         // the numeraire's exp(0 - max)
            const bool valid = t * TILE_ROWS + lane < n_rows;
            // when exp(-max) overflows the reference takes log-sum-exp = max - max
            const double lse = (mx < -700.0) ? 0.0 : mx + fm_log(sum + e0);
            ll.add(valid ? num - lse : 0.0);
            const double inv = (mx < -700.0) ? 0.0 : fm_rcp(sum + e0);
            __syncwarp();
#pragma unroll
            for (int c = 0; c < LOGIT_MAX_C1; c++) {
                const double r = ((idx == c + 1) ? 1.0 : 0.0) - ev[c] * inv;
                xs[c * LOGIT_XS + lane] = (valid && c < C1) ? r : 0.0;
            }
            xs[LOGIT_MAX_C1 * LOGIT_XS + lane] = 0.0;   // the padding category (M = 8)
        }
        __syncwarp();
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
            const double2 r2 = *reinterpret_cast<const double2*>(xs + g * LOGIT_XS + nb * 8 + 2 * q);
            S[nb][0] = r2.x; S[nb][1] = r2.y;           // S now holds R^T
        }
        // ---- stage 3
#pragma unroll
        for (int jb = 0; jb < JB; jb++) {
            const int j = jb * 8 + g;
#pragma unroll
            for (int nb = 0; nb < 4; nb++) {
                double2 x2 = make_double2(0.0, 0.0);
                if (j < K) x2 = *reinterpret_cast<const double2*>(tile + j * LOGIT_CS + nb * 8 + 2 * q);
                dmma_8x8x4(G[jb][0], G[jb][1], S[nb][0], x2.x);
                dmma_8x8x4(G2[jb][0], G2[jb][1], S[nb][1], x2.y);
            }
        }
        __syncwarp();   // every lane is done with this buffer before the next prefetch overwrites it
        if (NBUF == 2) cur ^= 1;
    }

    // ---- CTA reduction: lane (g, q) holds G^T[c = g][j = jb*8 + 2q + {0,1}]
#pragma unroll
    for (int jb = 0; jb < JB; jb++) { G[jb][0] += G2[jb][0]; G[jb][1] += G2[jb][1]; }
    const int NOUT = 1 + KB * C1;
    __syncthreads();                   // the ring is reused as the WARPS x NOUT reduction scratch
    const double llw = warp_sum(ll.value());
    if (lane == 0) smem[warp * NOUT] = llw;
    if (cat_ok) {
#pragma unroll
        for (int jb = 0; jb < JB; jb++) {
            smem[warp * NOUT + 1 + (jb * 8 + 2 * q) * C1 + g] = G[jb][0];
            smem[warp * NOUT + 1 + (jb * 8 + 2 * q + 1) * C1 + g] = G[jb][1];
        }
    }
    __syncthreads();
    for (int v = threadIdx.x; v < NOUT; v += LOGIT_THREADS) {
        double s = 0.0;
#pragma unroll
        for (int wv = 0; wv < WARPS; wv++) s += smem[wv * NOUT + v];
        partials[(size_t)blockIdx.x * NOUT + v] = s;
    }
    grid_finalize(partials, NOUT, ticket, out, fin, smem);
}

static int dmma_nbuf() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("APOP_B200_LOGIT_NBUF"); v = (e && e[0] == '2') ? 2 : 1; }
    return v;
}
static size_t dmma_smem() { return (size_t)(LOGIT_THREADS / 32) * (dmma_nbuf() * LOGIT_TILE_DOUBLES + 8 * LOGIT_XS) * sizeof(double); }
template <int KB>
static cudaError_t dmma_launch_one(int c1, const LogitLaunch& L, const LogitParams& prm) {
    const size_t sm = dmma_smem();
    if (dmma_nbuf() == 1) {
        static bool done = false;
        if (!done) { cudaFuncSetAttribute(logit_dmma_kernel<KB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); done = true; }
        logit_dmma_kernel<KB, 1><<<L.grid, LOGIT_THREADS, sm, L.stream>>>(L.tiles, L.n_rows, L.n_tiles, L.k, c1, prm, L.partials, L.ticket, L.out, L.fin);
    } else {
        static bool done = false;
        if (!done) { cudaFuncSetAttribute(logit_dmma_kernel<KB, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); done = true; }
        logit_dmma_kernel<KB, 2><<<L.grid, LOGIT_THREADS, sm, L.stream>>>(L.tiles, L.n_rows, L.n_tiles, L.k, c1, prm, L.partials, L.ticket, L.out, L.fin);
    }
    return cudaGetLastError();
}
template <int KB>
static int dmma_bps_one(int c1) {
    const size_t sm = dmma_smem();
    int nb = 0;
    if (dmma_nbuf() == 1) {
        cudaFuncSetAttribute(logit_dmma_kernel<KB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, logit_dmma_kernel<KB, 1>, LOGIT_THREADS, sm);
    } else {
        cudaFuncSetAttribute(logit_dmma_kernel<KB, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, logit_dmma_kernel<KB, 2>, LOGIT_THREADS, sm);
    }
    return nb;
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
                                                                        L.partials, L.ticket, L.out, L.fin);
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

// 0 = hybrid (default, logit_hybrid.cu), 1 = FP64-pipe kernel above, 2 = all-DMMA kernel above;
// APOP_B200_LOGIT_VARIANT=hybrid|fp64|dmma selects (kept for A/B measurements)
static int logit_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("APOP_B200_LOGIT_VARIANT");
        v = (e && !strcmp(e, "fp64")) ? 1 : (e && !strcmp(e, "dmma")) ? 2 : 0;
        const char* d = getenv("APOP_B200_LOGIT_DMMA");
        if (!e && d && d[0] == '1') v = 2;
    }
    return v;
}
static bool use_smem_variant() { return logit_variant() == 1; }
cudaError_t logit_launch(int c1, const LogitLaunch& L, const LogitParams& prm) {
    if (logit_variant() == 0) return logit_hybrid_launch(c1, L, prm);
    if (!use_smem_variant() && c1 >= 1 && c1 <= LOGIT_MAX_C1) {
        switch (logit_kb(L.k)) {
        case 8: return dmma_launch_one<8>(c1, L, prm);
        case 16: return dmma_launch_one<16>(c1, L, prm);
        case 24: return dmma_launch_one<24>(c1, L, prm);
        case 32: return dmma_launch_one<32>(c1, L, prm);
        }
    }
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
    if (logit_variant() == 0) return logit_hybrid_blocks_per_sm(k, c1);
    if (!use_smem_variant() && c1 >= 1 && c1 <= LOGIT_MAX_C1) {
        switch (logit_kb(k)) {
        case 8: return dmma_bps_one<8>(c1);
        case 16: return dmma_bps_one<16>(c1);
        case 24: return dmma_bps_one<24>(c1);
        case 32: return dmma_bps_one<32>(c1);
        }
    }
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
