// xtwx_kernel.cu -- H = X' diag(w) [X | y]: the k x k observed information of the GLM
// families (replacing the reference's numeric Hessian, apop_mle.m4.c:187-223, which costs
// >= 4 P score passes) and the OLS normal equations X'X, X'y (model/apop_ols.c:333-334).
//
// Same two-phase tile scheme as the multinomial logit kernel: lane = row computes the
// row weight and parks x (row-major, odd stride) and w*[x | y] (column-major) in the warp's
// shared-memory slab; lane = column then accumulates its row of H with 16-byte broadcast
// loads.  X and y are read from HBM once (8(k+1) bytes per row); per row the kernel issues
// k (k+1) DFMA, so for k = 32 it is FP64-pipe bound (~1056 DFMA per lane per 32 rows).
#include "xtwx_kernel.cuh"
#include "families.cuh"
#include <stdlib.h>

namespace apb {

template <int MODE>
__device__ __forceinline__ double row_weight(double xb, double y, double w, const double* aux) {
    if (MODE == XTWX_GRAM) return w;
    if (MODE == XTWX_PROBIT) {
        const RowOut o = FamProbit::eval(xb, y, w, aux);   // same clamps as the score
        return o.d * (o.d + xb);
    }
    if (MODE == XTWX_LOGIT2) {
        const double p = fm_rcp(1.0 + fmin(fm_exp(-xb), 1e300));
        return p * (1.0 - p);
    }
    return fm_exp(xb);
}

template <int KB, int MODE>
__global__ void __launch_bounds__(XTWX_THREADS)
xtwx_kernel(const double* __restrict__ tiles, const unsigned long long n_rows, const unsigned long long n_tiles,
            const int K, const int cols, const int has_w, const __grid_constant__ XtwxParams prm,
            double* __restrict__ partials, unsigned int* __restrict__ ticket, double* __restrict__ out,
            const __grid_constant__ FinalizeCtx fin) {
    constexpr int WARPS = XTWX_THREADS / 32;
    constexpr int XS = KB + 1;       // parked x rows: odd stride, conflict-free both ways
    constexpr int NC = KB + 1;       // columns of [X | y]
    constexpr int NOUT = KB * NC;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* sx = smem + warp * (32 * XS + NC * 32);
    double* sz = sx + 32 * XS;
    const unsigned long long gwarp = (unsigned long long)blockIdx.x * WARPS + warp;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * WARPS;

    double H[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) H[c] = 0.0;

    for (unsigned long long t = gwarp; t < n_tiles; t += nwarps) {
        const double* base = tiles + t * (unsigned long long)(cols * TILE_ROWS) + lane;
        double x[KB];
#pragma unroll
        for (int j = 0; j < KB; j++) x[j] = (j < K) ? ld_stream(base + j * TILE_ROWS) : 0.0;
        const double y = ld_stream(base + K * TILE_ROWS);
        const double w = has_w ? ld_stream(base + (K + 1) * TILE_ROWS) : 1.0;
        double xb = 0.0;
        if (MODE != XTWX_GRAM) {
#pragma unroll
            for (int j = 0; j < KB; j++) xb = fma(x[j], prm.beta[j], xb);
        }
        const bool valid = t * TILE_ROWS + lane < n_rows;
        const double wgt = valid ? row_weight<MODE>(xb, y, w, prm.aux) : 0.0;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < KB; j++) {
            sx[lane * XS + j] = x[j];
            sz[j * 32 + lane] = wgt * x[j];
        }
        sz[KB * 32 + lane] = wgt * y;
        __syncwarp();
#pragma unroll 2
        for (int i = 0; i < 32; i += 2) {
            const double xa = sx[i * XS + lane];
            const double xc = sx[(i + 1) * XS + lane];
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const double2 r = *reinterpret_cast<const double2*>(sz + c * 32 + i);
                H[c] = fma(xa, r.x, H[c]);
                H[c] = fma(xc, r.y, H[c]);
            }
        }
    }
    __syncthreads();
    double* red = smem;  // WARPS x NOUT doubles fits in the slabs (NOUT <= 32*XS + NC*32)
    if (lane < KB) {
#pragma unroll
        for (int c = 0; c < NC; c++) red[warp * NOUT + lane * NC + c] = H[c];
    }
    __syncthreads();
    for (int v = threadIdx.x; v < NOUT; v += XTWX_THREADS) {
        double s = 0.0;
#pragma unroll
        for (int wv = 0; wv < WARPS; wv++) s += red[wv * NOUT + v];
        partials[(size_t)blockIdx.x * NOUT + v] = s;
    }
    grid_finalize(partials, NOUT, ticket, out, fin, red);
}


// ---------------------------------------------------------------------------------------
// Tensor-pipe variant (default): H = (W X)' [X | y] as FP64 DMMA (mma.sync m8n8k4) with the
// rows of the tile as the contraction dimension.  Each warp stages its tile in shared memory
// (16-byte cp.async into padded columns; a TMA bulk-copy + mbarrier pipeline is selectable), computes the row weights with lane = row (x.beta against the
// constant bank, then the family's weight), and keeps the X fragments of the tile in registers:
// fragment (jb, ks) = X[rows ks*4 + q][columns jb*8 + g] serves both as the A operand (scaled by
// the row weight) of column block jb and as the B operand of column block jb.  Only the upper
// block triangle is accumulated (H is symmetric); X'y is one extra block column in GRAM mode.
// Per 32-row tile: 8 k-steps x (JB (JB+1)/2 [+ JB]) DMMA instead of k (k+1) DFMA per lane.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1)
        : "d"(a), "d"(b));
}
constexpr int XTWX_CS = 36;   // padded column stride of a staged tile (doubles)

template <int KB, int MODE, bool USE_TMA>
__global__ void __launch_bounds__(XTWX_THREADS, 3)
xtwx_dmma_kernel(const double* __restrict__ tiles, const unsigned long long n_rows, const unsigned long long n_tiles,
                 const int K, const int cols, const int has_w, const __grid_constant__ XtwxParams prm,
                 double* __restrict__ partials, unsigned int* __restrict__ ticket, double* __restrict__ out,
                 const __grid_constant__ FinalizeCtx fin) {
    constexpr int WARPS = XTWX_THREADS / 32;
    constexpr int JB = KB / 8;
    constexpr int NTRI = JB * (JB + 1) / 2;
    constexpr bool WITH_Y = (MODE == XTWX_GRAM);
    constexpr int NC = KB + 1;
    constexpr int NOUT = KB * NC;
    constexpr int TILE_D = (MAX_K_REG + 2) * XTWX_CS;   // X columns, y, w
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const unsigned long long gwarp = (unsigned long long)blockIdx.x * WARPS + warp;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * WARPS;
    constexpr int NBUF = USE_TMA ? 2 : 1;   // TMA: the next tile streams in while this one is consumed
    double* slab = smem + warp * (NBUF * TILE_D + 36);
    double* wbuf = slab + NBUF * TILE_D;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(wbuf + 32);
    if (USE_TMA) {
        if (lane == 0) { mbar_init(bars, 1); mbar_init(bars + 1, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned phase[2] = {0u, 0u};
    int cur = 0;
    auto tma_issue = [&](unsigned long long tt, int buf) {   // one bulk copy per (padded) column
        if (lane == 0) {
            fence_proxy_async();   // earlier generic reads of this buffer precede the async writes
            mbar_expect_tx(bars + buf, (unsigned)(cols * TILE_ROWS * sizeof(double)));
            const double* src = tiles + tt * (unsigned long long)(cols * TILE_ROWS);
            double* dst = slab + buf * TILE_D;
            for (int c = 0; c < cols; c++) tma_load_1d(dst + c * XTWX_CS, src + c * TILE_ROWS, TILE_ROWS * sizeof(double), bars + buf);
        }
    };
    if (USE_TMA && gwarp < n_tiles) tma_issue(gwarp, 0);

    double acc[NTRI][2], accy[JB][2];
#pragma unroll
    for (int b = 0; b < NTRI; b++) acc[b][0] = acc[b][1] = 0.0;
#pragma unroll
    for (int b = 0; b < JB; b++) accy[b][0] = accy[b][1] = 0.0;

    for (unsigned long long t = gwarp; t < n_tiles; t += nwarps) {
        double* tile = slab + (USE_TMA ? cur : 0) * TILE_D;
        if (USE_TMA) {
            if (t + nwarps < n_tiles) tma_issue(t + nwarps, cur ^ 1);
            mbar_wait(bars + cur, phase[cur]);
            phase[cur] ^= 1u;
        } else {
            // stage the tile with 16-byte cp.async: chunk c = lane + 32 i lies in column (lane >> 4) + 2 i
            const double* src = tiles + t * (unsigned long long)(cols * TILE_ROWS) + lane * 2;
            const unsigned sa = smem_u32(tile + (lane >> 4) * XTWX_CS + (lane & 15) * 2);
#pragma unroll
            for (int i = 0; i < (KB + 3) / 2; i++) {
                const int col = (lane >> 4) + 2 * i;
                if (col < cols)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa + i * 2 * XTWX_CS * 8), "l"(src + i * 64));
            }
            asm volatile("cp.async.commit_group;");
            asm volatile("cp.async.wait_group 0;");
            __syncwarp();
        }
        // ---- row weights, lane = row
        {
            double xb = 0.0;
            if (MODE != XTWX_GRAM) {
#pragma unroll
                for (int j = 0; j < KB; j++) xb = fma((j < K) ? tile[j * XTWX_CS + lane] : 0.0, prm.beta[j], xb);
            }
            const double y = tile[K * XTWX_CS + lane];
            const double w = has_w ? tile[(K + 1) * XTWX_CS + lane] : 1.0;
            const bool valid = t * TILE_ROWS + lane < n_rows;
            wbuf[lane] = valid ? row_weight<MODE>(xb, y, w, prm.aux) : 0.0;
        }
        __syncwarp();
        // ---- fragments: X[rows ks*4+q][columns jb*8+g], and the weights of the same rows
        double wv[8];
#pragma unroll
        for (int ks = 0; ks < 8; ks++) wv[ks] = wbuf[ks * 4 + q];
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
            double xf[JB];
#pragma unroll
            for (int jb = 0; jb < JB; jb++) xf[jb] = (jb * 8 + g < K) ? tile[(jb * 8 + g) * XTWX_CS + ks * 4 + q] : 0.0;
            int b = 0;
#pragma unroll
            for (int ab = 0; ab < JB; ab++) {
                const double a = xf[ab] * wv[ks];
#pragma unroll
                for (int cb = ab; cb < JB; cb++, b++) dmma_8x8x4(acc[b][0], acc[b][1], a, xf[cb]);
                if (WITH_Y) {   // the y column sits in slot n = 0 of an otherwise empty block
                    const double yb = (g == 0) ? tile[K * XTWX_CS + ks * 4 + q] : 0.0;
                    dmma_8x8x4(accy[ab][0], accy[ab][1], a, yb);
                }
            }
        }
        __syncwarp();   // the tile buffer is about to be overwritten
        if (USE_TMA) cur ^= 1;
    }

    // ---- CTA reduction: expand the block triangle into the full H[j][c] layout
    __syncthreads();
    double* red = smem;   // NOUT doubles, accumulated warp by warp in fixed order
    for (int v = threadIdx.x; v < NOUT; v += XTWX_THREADS) red[v] = 0.0;
    __syncthreads();
    for (int wv_ = 0; wv_ < WARPS; wv_++) {
        if (warp == wv_) {
            int b = 0;
#pragma unroll
            for (int ab = 0; ab < JB; ab++) {
#pragma unroll
                for (int cb = ab; cb < JB; cb++, b++) {
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int a = ab * 8 + g, c = cb * 8 + 2 * q + e;
                        // diagonal blocks hold both triangles already; off-diagonal ones are mirrored
                        red[a * NC + c] += acc[b][e];
                        if (cb != ab) red[c * NC + a] += acc[b][e];
                    }
                }
                if (WITH_Y && q == 0) red[(ab * 8 + g) * NC + KB] += accy[ab][0];
            }
        }
        __syncthreads();
    }
    for (int v = threadIdx.x; v < NOUT; v += XTWX_THREADS) partials[(size_t)blockIdx.x * NOUT + v] = red[v];
    __syncthreads();
    grid_finalize(partials, NOUT, ticket, out, fin, red);
}

// Staging choice, measured at 100M x 32 (round 1): 16-byte cp.async, single buffer, 12 warps/SM: 5.99 ms;
// TMA bulk copies (one lane, one mbarrier per warp) single buffer: 6.38 ms; double-buffered TMA
// (8 warps/SM): 7.7 ms -- the kernel is bound by the latency of its DMMA chains, so warps beat
// buffers.  cp.async is the default; APOP_B200_XTWX_TMA=1 selects the TMA pipeline.
static bool xtwx_use_tma() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("APOP_B200_XTWX_TMA"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}
template <int KB>
static size_t xtwx_dmma_smem() {
    const size_t ring = (size_t)(XTWX_THREADS / 32) * ((xtwx_use_tma() ? 2 : 1) * (MAX_K_REG + 2) * XTWX_CS + 36) * sizeof(double);
    const size_t red = (size_t)KB * (KB + 1) * sizeof(double);
    return ring > red ? ring : red;
}
template <int KB, int MODE>
static cudaError_t dmma_launch_one(const XtwxLaunch& L, const XtwxParams& prm) {
    const size_t sm = xtwx_dmma_smem<KB>();
    if (xtwx_use_tma()) {
        static bool done = false;
        if (!done) { cudaFuncSetAttribute(xtwx_dmma_kernel<KB, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); done = true; }
        xtwx_dmma_kernel<KB, MODE, true><<<L.grid, XTWX_THREADS, sm, L.stream>>>(L.tiles, L.n_rows, L.n_tiles, L.k, L.cols, L.has_w, prm,
                                                                                L.partials, L.ticket, L.out, L.fin);
    } else {
        static bool done = false;
        if (!done) { cudaFuncSetAttribute(xtwx_dmma_kernel<KB, MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); done = true; }
        xtwx_dmma_kernel<KB, MODE, false><<<L.grid, XTWX_THREADS, sm, L.stream>>>(L.tiles, L.n_rows, L.n_tiles, L.k, L.cols, L.has_w, prm,
                                                                                 L.partials, L.ticket, L.out, L.fin);
    }
    return cudaGetLastError();
}
template <int KB, int MODE>
static int dmma_bps_one() {
    const size_t sm = xtwx_dmma_smem<KB>();
    int nb = 0;
    if (xtwx_use_tma()) {
        cudaFuncSetAttribute(xtwx_dmma_kernel<KB, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, xtwx_dmma_kernel<KB, MODE, true>, XTWX_THREADS, sm);
    } else {
        cudaFuncSetAttribute(xtwx_dmma_kernel<KB, MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, xtwx_dmma_kernel<KB, MODE, false>, XTWX_THREADS, sm);
    }
    return nb;
}
static bool xtwx_use_fp64_variant() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("APOP_B200_XTWX_FP64"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

template <int KB>
static size_t xtwx_smem() { return (size_t)(XTWX_THREADS / 32) * (32 * (KB + 1) + (KB + 1) * 32) * sizeof(double); }

template <int KB, int MODE>
static cudaError_t launch_one(const XtwxLaunch& L, const XtwxParams& prm) {
    const size_t sm = xtwx_smem<KB>();
    static bool done = false;
    if (!done) { cudaFuncSetAttribute(xtwx_kernel<KB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); done = true; }
    xtwx_kernel<KB, MODE><<<L.grid, XTWX_THREADS, sm, L.stream>>>(L.tiles, L.n_rows, L.n_tiles, L.k, L.cols, L.has_w, prm,
                                                                 L.partials, L.ticket, L.out, L.fin);
    return cudaGetLastError();
}
template <int KB, int MODE>
static int bps_one() {
    const size_t sm = xtwx_smem<KB>();
    cudaFuncSetAttribute(xtwx_kernel<KB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, xtwx_kernel<KB, MODE>, XTWX_THREADS, sm);
    return nb;
}
int xtwx_kb(int k) { return k <= 8 ? 8 : k <= 16 ? 16 : k <= 24 ? 24 : 32; }

#define APB_MODES(KB)                              \
    switch (mode) {                                \
    case XTWX_GRAM: return APB_OP(KB, XTWX_GRAM);  \
    case XTWX_PROBIT: return APB_OP(KB, XTWX_PROBIT); \
    case XTWX_LOGIT2: return APB_OP(KB, XTWX_LOGIT2); \
    case XTWX_POISSON_REG: return APB_OP(KB, XTWX_POISSON_REG); \
    default: break;                                \
    }
cudaError_t xtwx_launch(int mode, const XtwxLaunch& L, const XtwxParams& prm) {
    if (!xtwx_use_fp64_variant()) {
#define APB_OP(KB, M) dmma_launch_one<KB, M>(L, prm)
        switch (xtwx_kb(L.k)) {
        case 8: APB_MODES(8) break;
        case 16: APB_MODES(16) break;
        case 24: APB_MODES(24) break;
        case 32: APB_MODES(32) break;
        }
#undef APB_OP
        return cudaErrorInvalidValue;
    }
#define APB_OP(KB, M) launch_one<KB, M>(L, prm)
    switch (xtwx_kb(L.k)) {
    case 8: APB_MODES(8) break;
    case 16: APB_MODES(16) break;
    case 24: APB_MODES(24) break;
    case 32: APB_MODES(32) break;
    }
#undef APB_OP
    return cudaErrorInvalidValue;
}
int xtwx_blocks_per_sm(int mode, int k) {
    if (!xtwx_use_fp64_variant()) {
#define APB_OP(KB, M) dmma_bps_one<KB, M>()
        switch (xtwx_kb(k)) {
        case 8: APB_MODES(8) break;
        case 16: APB_MODES(16) break;
        case 24: APB_MODES(24) break;
        case 32: APB_MODES(32) break;
        }
#undef APB_OP
        return 0;
    }
#define APB_OP(KB, M) bps_one<KB, M>()
    switch (xtwx_kb(k)) {
    case 8: APB_MODES(8) break;
    case 16: APB_MODES(16) break;
    case 24: APB_MODES(24) break;
    case 32: APB_MODES(32) break;
    }
#undef APB_OP
    return 0;
}
}  // namespace apb
