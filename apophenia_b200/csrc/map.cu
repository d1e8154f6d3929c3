// map.cu -- apop_map with a materialised result (apop_mapply.m4.c:119-217).
//
// The reference applies a host function pointer per cell (fn_d / fn_dp, parts 'a' 'm' 'v'), per row
// (fn_v / fn_vp / fn_r / fn_rp, part 'r') or per column (part 'c') and returns a new apop_data of the
// matching shape (:300-313).  A host function cannot run on the device, so the device side offers the
// functions the models on this path map with (and the elementary ones users write by hand) as an enum:
//   cell functions  -> an output of the input's shape (written in the resident tile layout, then
//                      re-laid row-major by tile_unpack on its way to the caller's buffers);
//   row functions   -> one double per row, lane = row, coalesced stores.  The per-row log-likelihood
//                      terms are the same Fam::eval the fused kernels use (biprobit_ll_row,
//                      model/apop_probit.c:83-88; one_logit_row with one free category, :212-229).
// HBM bound: one read of the tile and one write of the result.
#include "map.cuh"
#include "families.cuh"

namespace apb {

__device__ __forceinline__ double map_cell(int fn, double x, double p0) {
    switch (fn) {
    case MAPC_IDENTITY: return x;
    case MAPC_DEV: return x - p0;
    case MAPC_DEV2: return (x - p0) * (x - p0);
    case MAPC_POISSON: {   // poisson apply_me (model/apop_poisson.c:14-19); p0 = ln(lambda)
        if (x < 0.0 || (x - (double)(int)x) > 1e-4) return -INFINITY;
        return (x == 0.0) ? 0.0 : p0 * x - lgamma(x + 1.0);
    }
    case MAPC_LOG: return log(x);
    case MAPC_EXP: return exp(x);
    case MAPC_ABS: return fabs(x);
    case MAPC_SQRT: return sqrt(x);
    case MAPC_SCALE: return x * p0;
    case MAPC_POW: return pow(x, p0);
    default: return NAN;
    }
}

// out_tiles has the same tile geometry as the input chunk; the weights column (if any) is copied
__global__ void __launch_bounds__(256)
map_cells_kernel(const double* __restrict__ tiles, unsigned long long n_tiles, int k, int ycol, int cols, int fn, double p0,
                 int do_matrix, int do_vector, double* __restrict__ out_tiles) {
    const unsigned long long total = n_tiles * (unsigned long long)cols * TILE_ROWS;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const int col = (int)((i / TILE_ROWS) % (unsigned long long)cols);
        const double x = __ldcs(tiles + i);
        const bool apply = (col < k) ? do_matrix != 0 : (col == ycol ? do_vector != 0 : false);
        out_tiles[i] = apply ? map_cell(fn, x, p0) : x;
    }
}

template <class Fam>
__device__ __forceinline__ double row_ll(double xb, double y, const double* aux) {
    return Fam::eval(xb, y, 1.0, aux).s[0];
}

// one value per row; param (device) holds k doubles for the functions that take a vector, aux 4 doubles
__global__ void __launch_bounds__(256)
map_rows_kernel(const double* __restrict__ tiles, unsigned long long row0, unsigned long long rows, unsigned long long n_rows_total,
                int k, int ycol, int cols, int fn, const double* __restrict__ param, double4 aux4, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const unsigned long long gwarp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    const unsigned long long t0 = row0 / TILE_ROWS, t1 = (row0 + rows + TILE_ROWS - 1) / TILE_ROWS;
    const double aux[4] = {aux4.x, aux4.y, aux4.z, aux4.w};
    for (unsigned long long t = t0 + gwarp; t < t1; t += nwarps) {
        const double* base = tiles + t * (unsigned long long)cols * TILE_ROWS + lane;
        const unsigned long long row = t * TILE_ROWS + lane;
        double acc = (fn == MAPR_MAX) ? -INFINITY : (fn == MAPR_MIN ? INFINITY : 0.0);
        for (int j = 0; j < k; j++) {
            const double x = __ldcs(base + j * TILE_ROWS);
            switch (fn) {
            case MAPR_SUM: case MAPR_MEAN: acc += x; break;
            case MAPR_SUMSQ: acc = fma(x, x, acc); break;
            case MAPR_MAX: acc = fmax(acc, x); break;
            case MAPR_MIN: acc = fmin(acc, x); break;
            default: acc = fma(x, __ldg(param + j), acc); break;   // x . param for MAPR_DOT and the row likelihoods
            }
        }
        const double y = (ycol >= 0) ? __ldcs(base + ycol * TILE_ROWS) : 0.0;
        double v = acc;
        if (fn == MAPR_MEAN) v = acc / (double)k;
        else if (fn == MAPR_PROBIT_LL) v = row_ll<FamProbit>(acc, y, aux);
        else if (fn == MAPR_LOGIT2_LL) v = row_ll<FamLogit2>(acc, y, aux);
        else if (fn == MAPR_POISSON_REG_LL) v = row_ll<FamPoissonReg>(acc, y, aux) - lgamma(y + 1.0);
        else if (fn == MAPR_OLS_ERROR) v = acc - y;
        if (row >= row0 && row < row0 + rows && row < n_rows_total) out[row - row0] = v;
    }
}

cudaError_t map_cells_launch(const double* tiles, size_t n_tiles, int k, int ycol, int cols, int fn, double p0, bool do_matrix,
                             bool do_vector, double* out_tiles, cudaStream_t st) {
    if (!n_tiles) return cudaSuccess;
    const unsigned long long total = (unsigned long long)n_tiles * cols * TILE_ROWS;
    unsigned long long grid = (total + 255) / 256;
    if (grid > 148ull * 32) grid = 148ull * 32;
    map_cells_kernel<<<(unsigned)grid, 256, 0, st>>>(tiles, n_tiles, k, ycol, cols, fn, p0, do_matrix, do_vector, out_tiles);
    return cudaGetLastError();
}
cudaError_t map_rows_launch(const double* tiles, size_t row0, size_t rows, size_t n_rows_total, int k, int ycol, int cols, int fn,
                            const double* param, const double* aux, double* out, cudaStream_t st) {
    if (!rows) return cudaSuccess;
    const unsigned long long tiles_n = (rows + 2 * TILE_ROWS - 1) / TILE_ROWS;
    unsigned long long grid = (tiles_n + 7) / 8;
    if (grid > 148ull * 16) grid = 148ull * 16;
    map_rows_kernel<<<(unsigned)grid, 256, 0, st>>>(tiles, row0, rows, n_rows_total, k, ycol, cols, fn, param,
                                                    make_double4(aux[0], aux[1], aux[2], aux[3]), out);
    return cudaGetLastError();
}

}  // namespace apb
