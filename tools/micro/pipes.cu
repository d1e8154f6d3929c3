// pipes.cu -- do the FP64 FMA pipe and the FP64 tensor (DMMA) pipe run concurrently on sm_100a?
// Times (a) DFMA only, (b) DMMA m8n8k4 only, (c) both interleaved, (d) DMMA m16n8k16, per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double* d, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void dmma1684(double* d, const double* a, double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}

template <int MODE, int CH>
__global__ void __launch_bounds__(128) k(double* out, int iters, double a, double b) {
    double f[CH], m0[CH], m1[CH];
    double big[4][4];
    for (int i = 0; i < CH; i++) { f[i] = threadIdx.x * 1e-3 + i; m0[i] = i; m1[i] = -i; }
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) big[i][j] = i + j;
    double av[8] = {a, a, a, a, a, a, a, a}, bv[4] = {b, b, b, b};
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (MODE == 0 || MODE == 2) { f[i] = fma(f[i], a, b); f[i] = fma(f[i], a, b); f[i] = fma(f[i], a, b); f[i] = fma(f[i], a, b);
                                          f[i] = fma(f[i], a, b); f[i] = fma(f[i], a, b); f[i] = fma(f[i], a, b); f[i] = fma(f[i], a, b); }   // 8 DFMA = 16 issue clk/SMSP
            if (MODE == 1 || MODE == 2) dmma884(m0[i], m1[i], a, b);                                                                             // 1 DMMA = 256 FMA
            if (MODE == 3) if (i < 4) dmma16816(big[i], av, bv);
            if (MODE == 4) if (i < 4) dmma1684(big[i], av, b);
        }
    }
    double s = 0;
    for (int i = 0; i < CH; i++) s += f[i] + m0[i] + m1[i];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) s += big[i][j];
    if (s == 12345.678) out[0] = s;
}

template <int MODE, int CH>
void run(const char* name, int warps_per_sm, double fma_per_iter_per_warp) {
    int dev; cudaGetDevice(&dev); cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    double* out; cudaMalloc(&out, 8);
    const int iters = 20000;
    const int threads = 128, blocks = p.multiProcessorCount * (warps_per_sm * 32 / threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE, CH><<<blocks, threads>>>(out, 100, 0.999, 1e-3);
    cudaEventRecord(e0);
    k<MODE, CH><<<blocks, threads>>>(out, iters, 0.999, 1e-3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = fma_per_iter_per_warp * iters * (double)blocks * threads / 32;
    const double clk_per_iter = ms * 1e-3 * 1.9e9 / iters;
    printf("[%7.1f clk/iter @1.9GHz] ", clk_per_iter);
    printf("%-34s warps/SM=%2d chains=%d  %8.3f ms  %7.2f TFLOP/s (2*FMA)  err=%s\n", name, warps_per_sm, CH, ms, 2 * fmas / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    for (int w : {8, 16, 32}) {
        run<0, 8>("DFMA only", w, 8.0 * 8 * 32);
        run<1, 8>("DMMA m8n8k4 only", w, 8.0 * 256);
        run<2, 8>("DFMA + DMMA interleaved", w, 8.0 * (8 * 32 + 256));
        run<3, 8>("DMMA m16n8k16 only (4 chains)", w, 4.0 * 2048);
        run<4, 8>("DMMA m16n8k4 only (4 chains)", w, 4.0 * 512);
    }
    run<1, 2>("DMMA m8n8k4, 2 chains", 8, 2.0 * 256);
    run<1, 4>("DMMA m8n8k4, 4 chains", 8, 4.0 * 256);
    run<1, 1>("DMMA m8n8k4, 1 chain (latency)", 4, 1.0 * 256);
    run<0, 1>("DFMA, 1 chain (latency)", 4, 8.0 * 32);
    return 0;
}
