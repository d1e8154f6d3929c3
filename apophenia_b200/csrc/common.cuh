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
constexpr int MAX_K_REG = 32;      // widest X for the shared-memory / tensor kernels (logit, X'WX, batched)
constexpr int GLM_K_REG = 48;      // widest X the fused GLM kernel keeps in registers (x[K] and g[K]: 2 CTAs of 4 warps per SM)
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


// ---------------------------------------------------------------------------------------
// TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier: one elected lane
// moves whole tile columns global -> shared without touching registers; the warp waits on the
// barrier's phase parity.  Used by the kernels that stage tiles in shared memory.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// order earlier generic-proxy accesses to shared memory before later async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------
// Grid reduction tail, fused with the cross-GPU sum.
//
// Every reduction kernel of the library ends the same way: per-CTA partials in global
// memory, a ticket counter, and the last CTA to finish sums the partials in block order
// (two-sum compensated) -- bitwise reproducible for a given grid.  With several ranks (one
// process per GPU) the same CTA then performs the allreduce itself over NVLink: it stores its
// vector into a slot of EVERY rank's inbox (peer memory mapped through CUDA IPC), publishes a
// sequence number with a system-scope release, waits for the other ranks' sequence numbers in
// its own inbox, and adds the slots in rank order -- so all ranks obtain bitwise identical
// sums without a separate NCCL kernel, and the result lands directly in mapped pinned host
// memory.  Inboxes are double buffered on the parity of the sequence number: a rank can be at
// most one pass ahead of its slowest peer.
// ---------------------------------------------------------------------------------------
constexpr int P2P_MAX_RANKS = 8;
struct FinalizeCtx {
    int nranks, rank;                       // nranks <= 1: no exchange
    unsigned long long seq;                 // > 0, incremented by the host every launch
    int slot;                               // doubles per rank slot in an inbox
    double* inbox[P2P_MAX_RANKS];           // inbox of rank p as mapped on this GPU: [2][nranks][slot]
    unsigned long long* flags[P2P_MAX_RANKS];  // flags of rank p: [2][nranks]
    double* host_out;                       // mapped pinned host buffer (may be null)
    unsigned long long spin_limit;
    int* err;                               // set to 1 if a peer never arrived
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// Call with all threads of every CTA after the CTA's partials[blockIdx.x * nout + v] are
// written.  `scratch` is shared memory with at least nout doubles (used by the last CTA only).
__device__ __forceinline__ void grid_finalize(const double* __restrict__ partials, int nout,
                                              unsigned int* __restrict__ ticket, double* __restrict__ out,
                                              const FinalizeCtx& ctx, double* scratch) {
    __shared__ unsigned int is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // partials are summed in a fixed order: G = blockDim/nout thread groups take the blocks
    // b = g, g+G, ... (two-sum compensated), then the G subtotals are added in group order
    __shared__ double fin_group[256];
    const int G = (nout <= 64) ? (int)blockDim.x / nout : 1;
    if (G > 1 && G * nout <= 256) {
        if ((int)threadIdx.x < G * nout) {
            const int v = threadIdx.x % nout, gi = threadIdx.x / nout;
            double s = 0.0, c = 0.0;
            unsigned b = gi;
            // eight loads in flight per dependent add chain: a plain loop pays one L2 round trip per partial
            // (~20 us for 444 CTAs, most of the kernel's fixed cost); the order of the additions is unchanged
            for (; b + 7 * G < gridDim.x; b += 8 * G) {
                double x[8], e;
#pragma unroll
                for (int u = 0; u < 8; u++) x[u] = __ldcg(partials + (size_t)(b + u * G) * nout + v);
#pragma unroll
                for (int u = 0; u < 8; u++) { two_sum(s, x[u], s, e); c += e; }
            }
            for (; b < gridDim.x; b += G) {
                double e;
                two_sum(s, __ldcg(partials + (size_t)b * nout + v), s, e);
                c += e;
            }
            fin_group[gi * nout + v] = s + c;
        }
        __syncthreads();
        if ((int)threadIdx.x < nout) {
            double s = 0.0, c = 0.0;
            for (int gi = 0; gi < G; gi++) {
                double e;
                two_sum(s, fin_group[gi * nout + threadIdx.x], s, e);
                c += e;
            }
            scratch[threadIdx.x] = s + c;
        }
    } else {
        for (int v = threadIdx.x; v < nout; v += blockDim.x) {
            double s = 0.0, c = 0.0;
            unsigned b = 0;
            for (; b + 4 <= gridDim.x; b += 4) {   // four loads in flight per dependent add chain (same order of additions)
                double x[4], e;
#pragma unroll
                for (int u = 0; u < 4; u++) x[u] = __ldcg(partials + (size_t)(b + u) * nout + v);
#pragma unroll
                for (int u = 0; u < 4; u++) { two_sum(s, x[u], s, e); c += e; }
            }
            for (; b < gridDim.x; b++) {
                double e;
                two_sum(s, __ldcg(partials + (size_t)b * nout + v), s, e);
                c += e;
            }
            scratch[v] = s + c;
        }
    }
    __syncthreads();
    if (ctx.nranks > 1) {
        const int par = (int)(ctx.seq & 1ull);
        const size_t my_slot = ((size_t)par * ctx.nranks + ctx.rank) * ctx.slot;
        for (int p = 0; p < ctx.nranks; p++)
            for (int v = threadIdx.x; v < nout; v += blockDim.x) ctx.inbox[p][my_slot + v] = scratch[v];
        __threadfence_system();
        __syncthreads();
        if ((int)threadIdx.x < ctx.nranks)   // thread p tells rank p that this rank's slot is complete
            st_release_sys(ctx.flags[threadIdx.x] + par * ctx.nranks + ctx.rank, ctx.seq);
        if ((int)threadIdx.x < ctx.nranks) {  // and waits for rank threadIdx.x in its own inbox
            const unsigned long long* f = ctx.flags[ctx.rank] + par * ctx.nranks + threadIdx.x;
            unsigned long long spins = 0;
            while (ld_acquire_sys(f) < ctx.seq) {
                if (++spins > ctx.spin_limit) { *ctx.err = 1; break; }
                __nanosleep(64);
            }
        }
        __syncthreads();
        const double* mine = ctx.inbox[ctx.rank] + (size_t)par * ctx.nranks * ctx.slot;
        for (int v = threadIdx.x; v < nout; v += blockDim.x) {
            double s = 0.0;
            for (int r = 0; r < ctx.nranks; r++) s += ld_relaxed_sys(mine + (size_t)r * ctx.slot + v);  // rank order: same bits everywhere
            scratch[v] = s;
        }
        __syncthreads();
    }
    for (int v = threadIdx.x; v < nout; v += blockDim.x) {
        out[v] = scratch[v];
        if (ctx.host_out) ctx.host_out[v] = scratch[v];
    }
    if (ctx.host_out) __threadfence_system();
    if (threadIdx.x == 0) *ticket = 0;
}

}  // namespace apb
