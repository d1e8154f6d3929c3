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
                      int cols, double* tiles, cudaStream_t st);
cudaError_t tile_unpack(const double* tiles, size_t rows, int k, int cols, bool has_y, bool has_w,
                        double* Xs, double* ys, double* ws, cudaStream_t st);
cudaError_t synth_fill(int family, size_t n_rows, int k, int cols, int ncat, const SynthParams& prm,
                       unsigned long long seed, unsigned long long row_offset, double* tiles,
                       cudaStream_t st);

cudaError_t xbeta_launch(const double* tiles, size_t row0, size_t rows, int k, int cols, int c1, const double* B,
                         double* out, cudaStream_t st);
}  // namespace apb
