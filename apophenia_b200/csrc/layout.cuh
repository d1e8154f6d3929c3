// layout.cuh -- resident tiled layout helpers (see layout.cu)
#pragma once
#include "common.cuh"
#include <stddef.h>

namespace apb {

constexpr int SYNTH_MAX_C1 = 15;
constexpr int SYNTH_MAX_P = 480;
struct SynthParams {
    double beta[SYNTH_MAX_P];  // k x (ncat-1) row-major, or (lambda) / (mu, sigma)
};

cudaError_t tile_pack(const double* Xs, const double* ys, const double* ws, size_t rows, int k,
                      int cols, double* tiles, cudaStream_t st, bool y_col0 = false);
// device prep (prep.cu): the next distinct outcome value above a threshold; recode y to category indices
cudaError_t next_distinct_launch(const double* tiles, size_t n_rows, size_t n_tiles, int ycol, int cols, double above, int first,
                                 unsigned long long* d_key, int grid, cudaStream_t st);
constexpr int PREP_MAX_CATS = 64;
struct CatTable { double v[PREP_MAX_CATS]; int n; };
cudaError_t recode_launch(double* tiles, size_t n_rows, size_t n_tiles, int ycol, int cols, const CatTable& cats, cudaStream_t st);
cudaError_t tile_unpack(const double* tiles, size_t rows, int k, int cols, bool has_y, bool has_w,
                        double* Xs, double* ys, double* ws, cudaStream_t st);
// in-place conversion between the plain and the swizzled resident layout (an involution; see layout.cu)
cudaError_t tile_swizzle(double* tiles, size_t n_tiles, int cols, cudaStream_t st);
cudaError_t synth_fill(int family, size_t n_rows, int k, int cols, int ncat, const SynthParams& prm,
                       unsigned long long seed, unsigned long long row_offset, double* tiles,
                       cudaStream_t st);

cudaError_t xbeta_launch(const double* tiles, size_t row0, size_t rows, int k, int cols, int c1, const double* B,
                         double* out, cudaStream_t st);
}  // namespace apb
