// layout.cu -- resident tiled layout: pack (row-major -> tiles), unpack, and the
// device-side synthetic data generator (recipe of tests/test_apop.c:889-907).
#include "layout.cuh"

namespace apb {

// ---------------------------------------------------------------------------
// pack: staging buffers hold a chunk of rows in the reference's layout (X
// row-major with leading dimension k, y and w contiguous); one CTA of 32x8
// threads transposes one 32-row tile through shared memory.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tile_pack_kernel(const double* __restrict__ Xs, const double* __restrict__ ys,
                 const double* __restrict__ ws, unsigned long long rows, int k, int cols,
                 double* __restrict__ tiles, int y_col0) {
    __shared__ double sm[32][33];
    const unsigned long long tile = blockIdx.x;
    const unsigned long long row0 = tile * TILE_ROWS;
    double* dst = tiles + tile * (unsigned long long)cols * TILE_ROWS;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int cb = 0; cb < k; cb += 32) {
        for (int r = ty; r < 32; r += 8) {
            const unsigned long long row = row0 + r;
            sm[r][tx] = (row < rows && cb + tx < k) ? Xs[row * k + cb + tx] : 0.0;
        }
        __syncthreads();
        // y_col0: the device side of ols_shuffle (model/apop_ols.c:97-108) -- column 0 of the
        // caller's matrix is the outcome: it becomes the y column and column 0 becomes the constant 1
        if (y_col0 && cb == 0 && ty == 0) dst[k * TILE_ROWS + tx] = sm[tx][0];
        for (int c = ty; c < 32; c += 8)
            if (cb + c < k) dst[(cb + c) * TILE_ROWS + tx] = (y_col0 && cb + c == 0) ? ((row0 + tx < rows) ? 1.0 : 0.0) : sm[tx][c];
        __syncthreads();
    }
    int col = k;
    if (y_col0) col++;
    else if (ys) {
        if (ty == 0) dst[col * TILE_ROWS + tx] = (row0 + tx < rows) ? ys[row0 + tx] : 0.0;
        col++;
    }
    if (ws && ty == 1) dst[col * TILE_ROWS + tx] = (row0 + tx < rows) ? ws[row0 + tx] : 0.0;
}

__global__ void __launch_bounds__(256)
tile_unpack_kernel(const double* __restrict__ tiles, unsigned long long rows, int k, int cols,
                   int has_y, int has_w, double* __restrict__ Xs, double* __restrict__ ys,
                   double* __restrict__ ws) {
    __shared__ double sm[32][33];
    const unsigned long long tile = blockIdx.x;
    const unsigned long long row0 = tile * TILE_ROWS;
    const double* src = tiles + tile * (unsigned long long)cols * TILE_ROWS;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int cb = 0; Xs && cb < k; cb += 32) {   // Xs == NULL: only the outcome / weights are wanted
        for (int c = ty; c < 32; c += 8) sm[c][tx] = (cb + c < k) ? src[(cb + c) * TILE_ROWS + tx] : 0.0;
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const unsigned long long row = row0 + r;
            if (row < rows && cb + tx < k) Xs[row * k + cb + tx] = sm[tx][r];
        }
        __syncthreads();
    }
    int col = k;
    if (has_y) {
        if (ty == 0 && row0 + tx < rows && ys) ys[row0 + tx] = src[col * TILE_ROWS + tx];
        col++;
    }
    if (has_w && ty == 1 && row0 + tx < rows && ws) ws[row0 + tx] = src[col * TILE_ROWS + tx];
}

// ---------------------------------------------------------------------------
// Swizzled form of the resident layout (the multinomial logit kernels'): the 16-byte chunk c of column j of a
// tile is stored at chunk c ^ sw(j), sw(j) = ((j & 1) << 2) | (j & 2) -- i.e. row r at position
// r ^ (8 (j & 1) + 4 ((j >> 1) & 1)).  A tile brought into shared memory by ONE bulk copy is then bank-conflict
// free for all three access patterns of logit_stream.cu: lane = row (any permutation inside a 256-byte column
// is), the 16-byte DMMA fragments of the score (columns g, g+1 of a fragment land in different halves of a
// 128-byte bank row) and the 8-byte fragments of X.B (the four columns of a k-step land in different quarters).
// The map is an involution: the same kernel converts in either direction, in place, one warp per column.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tile_swizzle_kernel(double* __restrict__ tiles, unsigned long long n_tiles, int cols) {
    const int lane = threadIdx.x & 31;
    const unsigned long long total = n_tiles * (unsigned long long)cols;
    const unsigned long long nw = (unsigned long long)gridDim.x * (blockDim.x >> 5);
    for (unsigned long long w = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < total; w += nw) {
        const int c = (int)(w % (unsigned long long)cols);
        const int rx = ((c & 1) << 3) | ((c & 2) << 1);
        if (!rx) continue;                       // every fourth column stays as it is
        double* p = tiles + w * TILE_ROWS;      // column c of tile w / cols
        const double v = p[lane ^ rx];
        __syncwarp();
        p[lane] = v;
    }
}
cudaError_t tile_swizzle(double* tiles, size_t n_tiles, int cols, cudaStream_t st) {
    if (!n_tiles || cols < 2) return cudaSuccess;
    const unsigned long long total = (unsigned long long)n_tiles * cols;
    const unsigned long long want = (total + 7) / 8;
    const unsigned grid = (unsigned)(want < 148ull * 32 ? (want ? want : 1) : 148ull * 32);
    tile_swizzle_kernel<<<grid, 256, 0, st>>>(tiles, n_tiles, cols);
    return cudaGetLastError();
}

cudaError_t tile_pack(const double* Xs, const double* ys, const double* ws, size_t rows, int k,
                      int cols, double* tiles, cudaStream_t st, bool y_col0) {
    const size_t nt = (rows + TILE_ROWS - 1) / TILE_ROWS;
    if (!nt) return cudaSuccess;
    tile_pack_kernel<<<(unsigned)nt, dim3(32, 8), 0, st>>>(Xs, ys, ws, rows, k, cols, tiles, y_col0 ? 1 : 0);
    return cudaGetLastError();
}
cudaError_t tile_unpack(const double* tiles, size_t rows, int k, int cols, bool has_y, bool has_w,
                        double* Xs, double* ys, double* ws, cudaStream_t st) {
    const size_t nt = (rows + TILE_ROWS - 1) / TILE_ROWS;
    if (!nt) return cudaSuccess;
    tile_unpack_kernel<<<(unsigned)nt, dim3(32, 8), 0, st>>>(tiles, rows, k, cols, has_y, has_w, Xs, ys, ws);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Philox4x32-10 counter RNG (Salmon et al. 2011), written out so that the same
// stream is defined by (seed, global row, column pair) on any rank / any grid.
// ---------------------------------------------------------------------------
struct U4 { uint32_t x, y, z, w; };
__device__ __forceinline__ U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = U4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}
// two uniforms in (0,1) with 53 random bits each
__device__ __forceinline__ void uniform2(unsigned long long row, uint32_t idx, uint32_t stream,
                                         unsigned long long seed, double& u0, double& u1) {
    const U4 r = philox4x32_10(U4{(uint32_t)row, (uint32_t)(row >> 32), idx, stream},
                               (uint32_t)seed, (uint32_t)(seed >> 32));
    const unsigned long long a = (((unsigned long long)r.x << 32) | r.y) >> 11;
    const unsigned long long b = (((unsigned long long)r.z << 32) | r.w) >> 11;
    u0 = ((double)a + 0.5) * 0x1.0p-53;
    u1 = ((double)b + 0.5) * 0x1.0p-53;
}
__device__ __forceinline__ double std_normal(double u0, double u1) {
    double s, c;
    sincospi(2.0 * u1, &s, &c);
    return sqrt(-2.0 * log(u0)) * c;
}
__device__ double poisson_draw(double mu, unsigned long long row, uint32_t idx, unsigned long long seed) {
    // inversion by sequential search; mu is small (<= ~30) in every recipe here
    double u0, u1;
    uniform2(row, idx, 2u, seed, u0, u1);
    double p = exp(-mu), cdf = p;
    int x = 0;
    while (u0 > cdf && x < 2000) {
        x++;
        p *= mu / x;
        cdf += p;
    }
    return (double)x;
}

// family codes follow apop_b200_family in include/apop_b200.h
__global__ void __launch_bounds__(128)
synth_kernel(int family, unsigned long long n_rows, int k, int cols, int ncat, SynthParams prm,
             unsigned long long seed, unsigned long long row_offset, double* __restrict__ tiles) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long n_pad = (n_rows + TILE_ROWS - 1) / TILE_ROWS * TILE_ROWS;
    if (i >= n_pad) return;
    double* dst = tiles + (i / TILE_ROWS) * (unsigned long long)cols * TILE_ROWS + (i % TILE_ROWS);
    if (i >= n_rows) {
        for (int c = 0; c < cols; c++) dst[c * TILE_ROWS] = 0.0;
        return;
    }
    const unsigned long long row = row_offset + i;
    if (family == 2 || family == 3) {  // iid cells in the matrix, no vector
        for (int j = 0; j < k; j++) {
            double v;
            if (family == 2) v = poisson_draw(prm.beta[0], row, (uint32_t)j, seed);
            else {
                double u0, u1;
                uniform2(row, (uint32_t)j, 1u, seed, u0, u1);
                v = prm.beta[0] + prm.beta[1] * std_normal(u0, u1);
            }
            dst[j * TILE_ROWS] = v;
        }
        return;
    }
    // regression families: column 0 = 1, columns 1.. ~ U(-1,1)
    const int C1 = (family == 1) ? ncat - 1 : 1;
    double xb[SYNTH_MAX_C1];
    for (int c = 0; c < C1; c++) xb[c] = 0.0;
    for (int jb = 0; jb < k; jb += 2) {
        double u0, u1;
        uniform2(row, (uint32_t)(jb >> 1), 0u, seed, u0, u1);
        const double x0 = (jb == 0) ? 1.0 : 2.0 * u0 - 1.0;
        const double x1 = 2.0 * u1 - 1.0;
        dst[jb * TILE_ROWS] = x0;
        for (int c = 0; c < C1; c++) xb[c] += x0 * prm.beta[jb * C1 + c];
        if (jb + 1 < k) {
            dst[(jb + 1) * TILE_ROWS] = x1;
            for (int c = 0; c < C1; c++) xb[c] += x1 * prm.beta[(jb + 1) * C1 + c];
        }
    }
    double u0, u1, y;
    uniform2(row, 0u, 1u, seed, u0, u1);
    if (family == 0) {  // probit: y = 1{N(0,1) > -xb}
        y = (std_normal(u0, u1) > -xb[0]) ? 1.0 : 0.0;
    } else if (family == 1) {  // (multinomial) logit: softmax over [0, xb_1..]
        double m = 0.0;
        for (int c = 0; c < C1; c++) m = fmax(m, xb[c]);
        double den = exp(-m);
        for (int c = 0; c < C1; c++) den += exp(xb[c] - m);
        double cum = exp(-m) / den;
        int cat = 0;
        while (u0 > cum && cat < C1) {
            cum += exp(xb[cat] - m) / den;
            cat++;
        }
        y = (double)cat;
    } else if (family == 5) {  // poisson regression
        y = poisson_draw(exp(xb[0]), row, 0u, seed);
    } else {  // OLS: y = xb + N(0,1)
        y = xb[0] + std_normal(u0, u1);
    }
    dst[k * TILE_ROWS] = y;
}

cudaError_t synth_fill(int family, size_t n_rows, int k, int cols, int ncat, const SynthParams& prm,
                       unsigned long long seed, unsigned long long row_offset, double* tiles,
                       cudaStream_t st) {
    const size_t n_pad = (n_rows + TILE_ROWS - 1) / TILE_ROWS * TILE_ROWS;
    if (!n_pad) return cudaSuccess;
    synth_kernel<<<(unsigned)((n_pad + 127) / 128), 128, 0, st>>>(family, n_rows, k, cols, ncat, prm,
                                                                  seed, row_offset, tiles);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// apop_dot(d, parameters) (apop_linear_algebra.m4.c:424-448): the materialised N x c
// product X.B that the reference forms before every map (and that ols_predict /
// multilogit_expected return, model/apop_ols.c:359-366, model/apop_probit.c:166-199).  The
// fused kernels never need it; it exists for callers that want the fitted index itself.
// ---------------------------------------------------------------------------------------
template <int KB>
__global__ void __launch_bounds__(128)
xbeta_kernel(const double* __restrict__ tiles, unsigned long long row0, unsigned long long rows, int K, int cols, int C1,
             const double* __restrict__ B, double* __restrict__ out) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const unsigned long long r = row0 + i;
    const double* base = tiles + (r / TILE_ROWS) * (unsigned long long)cols * TILE_ROWS + (r % TILE_ROWS);
    double x[KB];
#pragma unroll
    for (int j = 0; j < KB; j++) x[j] = (j < K) ? __ldcs(base + j * TILE_ROWS) : 0.0;
    for (int c = 0; c < C1; c++) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < KB; j++) s = fma(x[j], (j < K) ? __ldg(B + j * C1 + c) : 0.0, s);
        out[i * C1 + c] = s;
    }
}
cudaError_t xbeta_launch(const double* tiles, size_t row0, size_t rows, int k, int cols, int c1, const double* B,
                         double* out, cudaStream_t st) {
    if (!rows) return cudaSuccess;
    const unsigned grid = (unsigned)((rows + 127) / 128);
    if (k <= 8) xbeta_kernel<8><<<grid, 128, 0, st>>>(tiles, row0, rows, k, cols, c1, B, out);
    else if (k <= 16) xbeta_kernel<16><<<grid, 128, 0, st>>>(tiles, row0, rows, k, cols, c1, B, out);
    else if (k <= 32) xbeta_kernel<32><<<grid, 128, 0, st>>>(tiles, row0, rows, k, cols, c1, B, out);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

}  // namespace apb
