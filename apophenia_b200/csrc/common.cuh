// common.cuh -- shared device helpers for the likelihood/score kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace apb {

// Resident data layout (see DESIGN.md "Data layout in HBM"):
// rows are grouped in tiles of TILE_ROWS; inside a tile the data is
// column-major, so that a warp (lane = row) reads one 256-byte line pair per
// column.  Tile t occupies [t*cols*32, (t+1)*cols*32) doubles, column c of row
// r at offset c*32 + r.  Columns: X_0..X_{k-1}, then y (if the data set has a
// vector), then w (if it has weights).
constexpr int TILE_ROWS = 32;
constexpr int MAX_K_REG = 32;      // widest X for the register-resident kernels
constexpr int GLM_THREADS = 128;   // threads per CTA of the fused kernels
constexpr int MAX_ACC = 3;         // scalar sums a family may produce besides the score

// Compensated (Kahan-Babuska / Neumaier) running sum: the per-thread leg of the
// block-then-grid compensated reduction.
struct KSum {
    double s, c;
    __device__ __forceinline__ KSum() : s(0.0), c(0.0) {}
    __device__ __forceinline__ void add(double x) {
        const double t = s + x;
        // Neumaier: pick the branch-free two-sum (exact for any ordering)
        const double bp = t - s;
        c += (s - (t - bp)) + (x - bp);
        s = t;
    }
    __device__ __forceinline__ double value() const { return s + c; }
};

// error-free a+b = s + e
__device__ __forceinline__ void two_sum(double a, double b, double& s, double& e) {
    s = a + b;
    const double bp = s - a;
    e = (a - (s - bp)) + (b - bp);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 8-byte load: read once, do not keep in L1
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }

}  // namespace apb
