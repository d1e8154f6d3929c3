// mix.cu -- cost of mixing DFMA and DMMA.8x8x4 on one SM sub-partition (they share a datapath, pipes.cu).
// Every warp runs `iters` rounds; in a round an "F" warp issues NF independent-chain DFMAs, an "M" warp NM DMMAs.
//   case 1: all warps F            case 2: all warps M
//   case 3: even warps F, odd warps M (fine-grained inter-warp mixing, what the scheduler does in the logit kernel)
//   case 4: every warp alternates a block of F and a block of M (intra-warp phases, all warps in the same phase order)
// If the datapath switched for free, t3 = t4 = (t1 + t2) / 2 for the same total work.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int CASE>
__global__ void __launch_bounds__(128) k(double* out, int iters, double a, double b) {
    double f[8], m0[8], m1[8];
    for (int i = 0; i < 8; i++) { f[i] = threadIdx.x * 1e-3 + i; m0[i] = i; m1[i] = -i; }
    const int warp = threadIdx.x >> 5;
    const bool isF = (CASE == 1) || (CASE == 3 && (warp & 1) == 0);
    const bool isM = (CASE == 2) || (CASE == 3 && (warp & 1) == 1);
    for (int it = 0; it < iters; it++) {
        if (isF || CASE == 4) {
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) f[i] = fma(f[i], a, b);          // 64 DFMA = 128 issue cycles
        }
        if (isM || CASE == 4) {
#pragma unroll
            for (int i = 0; i < 8; i++) dmma(m0[i], m1[i], a, b);             // 8 DMMA = 128 issue cycles
        }
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += f[i] + m0[i] + m1[i];
    if (s == 12345.678) out[0] = s;
}
template <int CASE>
float run(int warps_per_sm) {
    int dev; cudaGetDevice(&dev); cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    double* out; cudaMalloc(&out, 8);
    const int iters = 20000, blocks = p.multiProcessorCount * (warps_per_sm / 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<CASE><<<blocks, 128>>>(out, 100, 0.999, 1e-3);
    cudaEventRecord(e0);
    k<CASE><<<blocks, 128>>>(out, iters, 0.999, 1e-3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaFree(out);
    return ms;
}
int main() {
    for (int w : {8, 16}) {   // 2 and 4 warps per sub-partition
        const float t1 = run<1>(w), t2 = run<2>(w), t3 = run<3>(w), t4 = run<4>(w);
        printf("warps/SM=%2d  all-DFMA %.3f ms  all-DMMA %.3f ms  | half/half warps %.3f ms (free switching would give %.3f)  | every warp alternates blocks %.3f ms (free: %.3f)\n",
               w, t1, t2, t3, (t1 + t2) / 2, t4, t1 + t2);
    }
    return 0;
}
