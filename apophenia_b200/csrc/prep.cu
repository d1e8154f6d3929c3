// prep.cu -- the O(N) steps of the models' prep routines on the resident data.
//
// get_category_table / apop_data_to_factors (model/apop_probit.c:32-40, apop_regression.m4.c:62-78,
// 423-450) sort the N outcome values on the host to find the C distinct ones and then rewrite
// every outcome as its index.  With C tiny and the outcome column already in HBM this is C + 1
// streaming passes over 8 bytes per row (the y segment of each tile): pass c finds the smallest
// value above category c-1 (an ordered-bits atomicMin), and one more pass recodes the column in
// place.  100M rows, C = 2: 3 x 0.8 GB of reads instead of a 100M-element qsort.
#include "layout.cuh"

namespace apb {

// doubles ordered as unsigned integers (NaNs are skipped by the caller)
__device__ __forceinline__ unsigned long long ordered_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(256)
next_distinct_kernel(const double* __restrict__ tiles, unsigned long long n_rows, unsigned long long n_tiles, int ycol, int cols,
                     double above, int first, unsigned long long* __restrict__ key) {
    const int lane = threadIdx.x & 31;
    const unsigned long long gwarp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    unsigned long long best = ~0ull;
    for (unsigned long long t = gwarp; t < n_tiles; t += nwarps) {
        const double y = __ldcs(tiles + t * (unsigned long long)cols * TILE_ROWS + (unsigned long long)ycol * TILE_ROWS + lane) + 0.0;   // -0 -> +0: one category
        const bool take = (t * TILE_ROWS + lane < n_rows) && (y == y) && (first || y > above);
        const unsigned long long k = take ? ordered_key(y) : ~0ull;
        best = k < best ? k : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    if (lane == 0 && best != ~0ull) atomicMin(key, best);
}

__global__ void __launch_bounds__(256)
recode_kernel(double* __restrict__ tiles, unsigned long long n_rows, unsigned long long n_tiles, int ycol, int cols,
              const __grid_constant__ CatTable cats) {
    const int lane = threadIdx.x & 31;
    const unsigned long long gwarp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    for (unsigned long long t = gwarp; t < n_tiles; t += nwarps) {
        double* p = tiles + t * (unsigned long long)cols * TILE_ROWS + (unsigned long long)ycol * TILE_ROWS + lane;
        const double y = *p;
        // index of the last category value <= y (the host mirror's bisection; exact matches by construction)
        int idx = 0;
        for (int c = 1; c < cats.n; c++) idx = (cats.v[c] <= y) ? c : idx;
        if (t * TILE_ROWS + lane < n_rows) *p = (double)idx;
    }
}

cudaError_t next_distinct_launch(const double* tiles, size_t n_rows, size_t n_tiles, int ycol, int cols, double above, int first,
                                 unsigned long long* d_key, int grid, cudaStream_t st) {
    next_distinct_kernel<<<grid, 256, 0, st>>>(tiles, n_rows, n_tiles, ycol, cols, above, first, d_key);
    return cudaGetLastError();
}
cudaError_t recode_launch(double* tiles, size_t n_rows, size_t n_tiles, int ycol, int cols, const CatTable& cats, cudaStream_t st) {
    const size_t warps = n_tiles ? n_tiles : 1;
    size_t grid = (warps + 7) / 8;
    if (grid > 148 * 16) grid = 148 * 16;
    recode_kernel<<<(unsigned)grid, 256, 0, st>>>(tiles, n_rows, n_tiles, ycol, cols, cats);
    return cudaGetLastError();
}

}  // namespace apb
