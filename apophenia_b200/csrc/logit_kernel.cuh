// logit_kernel.cuh -- multinomial logit fused kernel (see logit_kernel.cu)
#pragma once
#include "common.cuh"
#include <stddef.h>
#include <cuda_runtime.h>

namespace apb {
constexpr int LOGIT_THREADS = 128;
constexpr int LOGIT_MAX_C1 = 7;  // up to 8 outcome categories
struct LogitParams {
    double B[MAX_K_REG * LOGIT_MAX_C1];  // k x (C-1), row-major (parameters->matrix), zero padded to KB rows
    double cats[8];                      // the first C-1 sorted category values (factor page vector)
};
struct LogitLaunch {
    const double* tiles;
    unsigned long long n_rows, n_tiles;
    int k, cols;           // cols = resident columns per tile (k + y [+ w, ignored])
    int swz = 0;           // the tiles are in the swizzled resident form (layout.cu: rows r ^ 8 in odd columns)
    const unsigned char* counts = nullptr;   // resample multiplicities of one bootstrap replicate (NULL: every row once)
    int m_pad = 0, replicate = 0;
    double* partials;      // [grid][1 + KB*(C-1)]
    unsigned int* ticket;
    double* out;           // [1 + KB*(C-1)]: LL, then G[j][c] at 1 + j*(C-1) + c
    int grid;
    cudaStream_t stream;
    FinalizeCtx fin;
};
int logit_kb(int k);
// logit_stream.cu: the default kernel (bulk-copy ring, both contractions on the FP64 MMA path, lane = row soft-max)
cudaError_t logit_stream_launch(int c1, const LogitLaunch& L, const LogitParams& prm);
int logit_stream_blocks_per_sm(int k, int c1);
int logit_stream_warps_per_block();
// logit_hybrid.cu: the round-1/2 kernel (lane = row X.B and soft-max, DMMA score, cp.async ring; plain tiles)
cudaError_t logit_hybrid_launch(int c1, const LogitLaunch& L, const LogitParams& prm);
int logit_hybrid_blocks_per_sm(int k, int c1);
// logit_info.cu: observed information, LOGIT_INFO_NP category pairs (a <= b; a < 0 = padding) per launch;
// out = NP blocks of KB x KB (full, symmetric), block i = X' diag([a==b] p_a - p_a p_b) X
constexpr int LOGIT_INFO_THREADS = 128;
constexpr int LOGIT_INFO_NP = 4;
struct LogitInfoPairs { int a[LOGIT_INFO_NP], b[LOGIT_INFO_NP]; };
cudaError_t logit_info_launch(int c1, const LogitLaunch& L, const LogitParams& prm, const LogitInfoPairs& pairs);
int logit_info_blocks_per_sm(int k, int c1);
cudaError_t logit_launch(int c1, const LogitLaunch& L, const LogitParams& prm);
int logit_blocks_per_sm(int k, int c1);
bool logit_wants_swizzled();
int logit_warps_per_block();   // of the kernel logit_launch picks (grid sizing)

#ifdef __CUDACC__
__device__ __forceinline__ void dmma_884(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1)
        : "d"(a), "d"(b));
}


// exp(x) for |x| <= 708 with a 16-entry table of 2^(j/16) in shared memory (one 128-byte line: 16 entries on
// 32 banks, so any pattern of lanes is conflict free): k = rint(16 x / ln 2), x = k ln2/16 + r, |r| <= ln2/32,
// exp(x) = 2^(k >> 4) * T[k & 15] * e^r with e^r - 1 by a degree-7 polynomial (r^8/8! < 1.2e-18): 12 FP64-pipe
// operations instead of the 17 of fm_exp_fast, the rest on the ALU and shared-memory pipes.
__device__ __forceinline__ double exp_tab16(double x, const double* __restrict__ tab) {
    const double t = fma(x, 23.083120654223414, 6755399441055744.0);       // 16 / ln 2
    const double kf = t - 6755399441055744.0;
    const int ki = __double2loint(t);
    double r = fma(kf, -4.33216987730702385306e-02, x);                      // ln2/16 hi (the fdlibm ln2 hi / 16: exact)
    r = fma(kf, -1.19263433079411731251e-11, r);                             // ln2/16 lo
    double p = 1.9841269841269841e-04;                                       // 1/7!
    p = fma(p, r, 1.3888888888888889e-03);
    p = fma(p, r, 8.3333333333333332e-03);
    p = fma(p, r, 4.1666666666666664e-02);
    p = fma(p, r, 1.6666666666666666e-01);
    p = fma(p, r, 5.0000000000000000e-01);
    p = fma(p, r, 1.0);
    const double T = tab[ki & 15];
    const double e = fma(T * r, p, T);                                       // T (1 + r p)
    return __hiloint2double(__double2hiint(e) + ((ki >> 4) << 20), __double2loint(e));
}
#endif
}  // namespace apb
