// batched_kernel.cuh -- batched-beta LL/score on the FP64 tensor-core path (see batched_kernel.cu)
#pragma once
#include "common.cuh"
#include <stddef.h>

namespace apb {
constexpr int BATCH_THREADS = 256;  // 8 warps, one replicate group each, all on the same tile
constexpr int BATCH_RC_MAX = 16;
constexpr int BATCH_TILE_CS = 40;   // column stride (doubles) of a tile staged in shared memory by the shared-centre kernel    // replicates per warp: 8 (two CTAs per SM) or 16 (one)
struct BatchedLaunch {
    const double* tiles;
    unsigned long long n_rows, n_tiles;
    int k, cols;
    const double* B;              // device, replicate r at B + r*P
    int P, m;
    const unsigned char* counts;  // NULL, or [tile][m_pad][32] bytes (see count_pos)
    int m_pad;
    double aux[4];
    int n_rowgroups;
    double* partials;             // [n_rowgroups][m][1 + KB]
    double* out;                  // [m][1 + KB]: LL, then the score of that replicate
    int grid;
    int mb;                       // replicate blocks of 8 per warp: 1 or 2
    bool want_score;
    cudaStream_t stream;
    const double* centre = nullptr;   // device, P doubles: the shared expansion point (NULL: exact evaluation of every pair)
    double delta_max = 0.0;           // |x.beta_r - x.centre| up to which the Taylor form is used
    unsigned long long* exact_slots = nullptr;   // device counter: warp steps that needed the exact evaluation
};
int batched_kb(int k);
cudaError_t batched_launch(int glm_fam, const BatchedLaunch& L);
cudaError_t bootstrap_counts_launch(unsigned char* counts, size_t n_local, size_t row_offset, size_t n_global, int m,
                                    int m_pad, unsigned long long seed, cudaStream_t st);
}  // namespace apb
