// batched_kernel.cu -- many parameter vectors per pass (bootstrap replicates, parallel
// MCMC chains): X.[beta_1 .. beta_m] is a real dense FP64 contraction and runs on the
// tensor-core path, mma.sync.aligned.m8n8k4 f64 (SASS: DMMA.8x8x4; tcgen05 has no FP64 kind).
//
// For a 32-row tile and a group of 16 replicates per warp:
//   stage 1  S^T[rep][row] = B^T[rep][j] . X^T[j][row]       (DMMA, K = columns)
//            A fragments (B^T) are loop invariant and live in registers;
//            B fragments (X^T) come straight from the resident tile: lane (g = lane/4,
//            q = lane%4) of block (kb, nb) reads column kb*4+q, row nb*8+g.
//   stage 2  per element: link/density and score multiplier of the family, times the
//            resample multiplicity c[row][rep] (bootstrap) -- the FP64-pipe bulk of the
//            kernel (N*m special-function evaluations);
//   stage 3  G^T[rep][j] += D^T[rep][row] . X[row][j]          (DMMA, K = rows)
//            The C fragment of stage 1 holds, per lane, rows nb*8 + 2q + {0,1} of replicate
//            g: exactly an A fragment if the four contraction slots are taken to be rows
//            {2q+e}, so the X operand is loaded with the same row permutation (one 16-byte
//            load gives e = 0,1) and no shuffle or shared-memory transpose is needed.
// All 8 warps of a CTA work on the same tile (L1-resident, 8(k+1)*32 bytes) for 8
// different replicate groups, so X is re-read from L2/HBM only m/128 times per pass.
// Reference: the serial loops this replaces are apop_bootstrap.m4.c:143-159 (resample
// + full fit per replicate) and apop_mcmc.m4.c:198-210 (one LL per step, one chain).
#include "batched_kernel.cuh"
#include "families.cuh"

namespace apb {

__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// position of row r (0..31) of a tile inside a replicate's 32-byte count record:
// lane q = (r%8)/2 owns bytes q*8 .. q*8+7, ordered (nb = r/8, e = r%2)
__host__ __device__ __forceinline__ int count_pos(int r) { return ((r & 7) >> 1) * 8 + (r >> 3) * 2 + (r & 1); }

// the per-row constant a family leaves out of s[0] (Poisson regression: -lgamma(y+1)); with resample counts it
// is weighted by the replicate's multiplicity, so it has to be part of the row term
template <class Fam, bool ON>
__device__ __forceinline__ double row_const_of(double y) {
    if constexpr (ON && Fam::ROW_CONST) return Fam::row_const(y);
    else return 0.0;
}

template <int KB, class Fam, bool WANT_SCORE, bool USE_COUNTS, int MB>
__global__ void __launch_bounds__(BATCH_THREADS, (MB == 1) ? 2 : 1)
batched_kernel(const double* __restrict__ tiles, const unsigned long long n_rows, const unsigned long long n_tiles,
               const int K, const int cols, const double* __restrict__ Bmat, const int P, const int m,
               const unsigned char* __restrict__ counts, const int m_pad, const double4 aux4,
               const int n_rowgroups, double* __restrict__ partials) {
    constexpr int WARPS = BATCH_THREADS / 32;
    constexpr int RC = MB * 8;         // replicates per warp (MB blocks of 8)
    constexpr int KS = KB / 4;         // stage-1 k-steps
    constexpr int JB = KB / 8;         // stage-3 column blocks
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int n_repgroups = (m + WARPS * RC - 1) / (WARPS * RC);
    const int repgroup = blockIdx.x % n_repgroups;
    const int rowgroup = blockIdx.x / n_repgroups;
    if (rowgroup >= n_rowgroups) return;
    const int rep0 = (repgroup * WARPS + warp) * RC;   // first replicate of this warp
    const double aux[4] = {aux4.x, aux4.y, aux4.z, aux4.w};

    // loop-invariant A fragments of stage 1: B^T[rep0 + mb*8 + g][ks*4 + q]
    double bfrag[MB][KS];
#pragma unroll
    for (int mb = 0; mb < MB; mb++)
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
            const int rep = rep0 + mb * 8 + g, j = ks * 4 + q;
            bfrag[mb][ks] = (rep < m && j < K) ? Bmat[(size_t)rep * P + j] : 0.0;
        }

    double G[MB][JB][2];   // G^T[rep = mb*8+g][j = jb*8 + 2q + {0,1}]
#pragma unroll
    for (int mb = 0; mb < MB; mb++)
#pragma unroll
        for (int jb = 0; jb < JB; jb++) G[mb][jb][0] = G[mb][jb][1] = 0.0;
    KSum ll[MB];

    for (unsigned long long t = rowgroup; t < n_tiles; t += n_rowgroups) {
        const double* tile = tiles + t * (unsigned long long)(cols * TILE_ROWS);
        // ---- stage 1: S^T = B^T X^T
        double S[MB][4][2];
#pragma unroll
        for (int mb = 0; mb < MB; mb++)
#pragma unroll
            for (int nb = 0; nb < 4; nb++) S[mb][nb][0] = S[mb][nb][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
            const int j = ks * 4 + q;
            double xf[4];
#pragma unroll
            for (int nb = 0; nb < 4; nb++) xf[nb] = (j < K) ? __ldg(tile + j * TILE_ROWS + nb * 8 + g) : 0.0;
#pragma unroll
            for (int mb = 0; mb < MB; mb++)
#pragma unroll
                for (int nb = 0; nb < 4; nb++) dmma_8x8x4(S[mb][nb][0], S[mb][nb][1], bfrag[mb][ks], xf[nb]);
        }
        // ---- stage 2: link/density per (row, replicate); this lane owns rows nb*8+2q+e, replicates mb*8+g
        double yv[4][2];
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
            const double2 y2 = __ldg(reinterpret_cast<const double2*>(tile + K * TILE_ROWS + nb * 8 + 2 * q));
            yv[nb][0] = y2.x; yv[nb][1] = y2.y;
        }
#pragma unroll
        for (int mb = 0; mb < MB; mb++) {
            const int rep = rep0 + mb * 8 + g;
            unsigned long long cbits = 0x0101010101010101ull;
            if (USE_COUNTS)
                cbits = (rep < m) ? __ldg(reinterpret_cast<const unsigned long long*>(
                                        counts + (t * (unsigned long long)m_pad + rep) * TILE_ROWS + q * 8))
                                  : 0ull;
            // LL-only passes (Metropolis chains) add the 8 terms of this lane plainly and make one compensated
            // add per tile (-3.4 % on 50M x 24, m = 256); with the score stage live the extra register tips the
            // 128-register kernel into spills (+3.8 % on 5M x 20, m = 1000), so there every term is added compensated
            double lsum = 0.0;
#pragma unroll
            for (int nb = 0; nb < 4; nb++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const unsigned long long row = t * TILE_ROWS + nb * 8 + 2 * q + e;
                    const double c = (row < n_rows) ? (double)((cbits >> (8 * (nb * 2 + e))) & 0xffull) : 0.0;
                    const RowOut o = Fam::eval(S[mb][nb][e], yv[nb][e], 1.0, aux);
                    const double term = o.s[0] + row_const_of<Fam, USE_COUNTS>(yv[nb][e]);
                    if (WANT_SCORE) ll[mb].add(c != 0.0 ? c * term : 0.0);
                    else lsum += c != 0.0 ? c * term : 0.0;
                    S[mb][nb][e] = c != 0.0 ? c * o.d : 0.0;   // S now holds D^T
                }
            if (!WANT_SCORE) ll[mb].add(lsum);
        }
        // ---- stage 3: G^T += D^T X, contraction slots = rows nb*8 + 2q + e
        if (WANT_SCORE) {
#pragma unroll
            for (int jb = 0; jb < JB; jb++) {
                const int j = jb * 8 + g;
#pragma unroll
                for (int nb = 0; nb < 4; nb++) {
                    double2 x2 = make_double2(0.0, 0.0);
                    if (j < K) x2 = __ldg(reinterpret_cast<const double2*>(tile + j * TILE_ROWS + nb * 8 + 2 * q));
#pragma unroll
                    for (int mb = 0; mb < MB; mb++) {
                        dmma_8x8x4(G[mb][jb][0], G[mb][jb][1], S[mb][nb][0], x2.x);
                        dmma_8x8x4(G[mb][jb][0], G[mb][jb][1], S[mb][nb][1], x2.y);
                    }
                }
            }
        }
    }

    // ---- per-CTA partials: [rowgroup][rep][1 + KB]
    const int NOUT = 1 + KB;
#pragma unroll
    for (int mb = 0; mb < MB; mb++) {
        const int rep = rep0 + mb * 8 + g;
        double v = ll[mb].value();
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (rep < m) {
            double* dst = partials + ((size_t)rowgroup * m + rep) * NOUT;
            if (q == 0) dst[0] = v;
            if (WANT_SCORE) {
#pragma unroll
                for (int jb = 0; jb < JB; jb++) {
                    dst[1 + jb * 8 + 2 * q] = G[mb][jb][0];
                    dst[1 + jb * 8 + 2 * q + 1] = G[mb][jb][1];
                }
            }
        }
    }
}


// ---------------------------------------------------------------------------------------
// Shared-centre variant (the default when the caller's m parameter vectors lie close together, as the
// replicates of a bootstrap and the chains of a sampler do): stage 2 evaluates the family's row term by its
// degree-6 Taylor polynomial around x0_i = x_i . centre (families.cuh: Taylor<Fam>), whose coefficients are
// computed once per row and shared by every replicate -- ~15 FP64 operations per (row, replicate) pair instead of
// the ~85 of a probit link evaluation.  A pair with |x_i.beta_r - x0_i| > delta_max, or a row whose x0 lies in
// the clamp region of the reference formula, takes the exact evaluation (a divergent branch, rare by
// construction), so the result never depends on the closeness assumption.
//   phase A  the 8 warps of the CTA take the next 8 tiles of the CTA's row group, one each, lane = row:
//            x0, the exact Fam::eval at x0, the coefficients -> a shared slab [8 tiles][NC][32 rows];
//   phase B  all 8 warps walk those 8 tiles as before (stage 1 DMMA, stage 2 Horner, stage 3 DMMA), reading
//            the coefficients of their rows back as 16-byte pairs (rows 2q, 2q+1).
// ---------------------------------------------------------------------------------------
template <int KB, class Fam, bool WANT_SCORE, bool USE_COUNTS>
__global__ void __launch_bounds__(BATCH_THREADS, 2)
batched_taylor_kernel(const double* __restrict__ tiles, const unsigned long long n_rows, const unsigned long long n_tiles,
                      const int K, const int cols, const double* __restrict__ Bmat, const int P, const int m,
                      const unsigned char* __restrict__ counts, const int m_pad, const double4 aux4,
                      const int n_rowgroups, double* __restrict__ partials, const double* __restrict__ centre,
                      const double delta_max, unsigned long long* __restrict__ exact_slots) {
    constexpr int WARPS = BATCH_THREADS / 32;
    constexpr int KS = KB / 4, JB = KB / 8;
    constexpr int NC = 8;                       // x0, a0..a6 (the derivative comes out of the same Horner recurrence)
    constexpr int CS = BATCH_TILE_CS;           // padded column stride of a staged tile: both fragment patterns conflict free
    extern __shared__ __align__(16) double dsm[];
    typedef double (*SlabT)[NC][TILE_ROWS];
    SlabT slab = reinterpret_cast<SlabT>(dsm);                       // [WARPS][NC][32]
    double* stile = dsm + WARPS * NC * TILE_ROWS;                    // [WARPS][(K + 1) * CS]: the 8 tiles of this round
    const int tile_d = (K + 1) * CS;
    __shared__ double s_centre[MAX_K_REG];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int n_repgroups = (m + WARPS * 8 - 1) / (WARPS * 8);
    const int repgroup = blockIdx.x % n_repgroups;
    const int rowgroup = blockIdx.x / n_repgroups;
    if (rowgroup >= n_rowgroups) return;
    const int rep0 = (repgroup * WARPS + warp) * 8;
    const int rep = rep0 + g;
    const double aux[4] = {aux4.x, aux4.y, aux4.z, aux4.w};
    if (threadIdx.x < MAX_K_REG) s_centre[threadIdx.x] = ((int)threadIdx.x < K) ? centre[threadIdx.x] : 0.0;

    double bfrag[KS];
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
        const int j = ks * 4 + q;
        bfrag[ks] = (rep < m && j < K) ? Bmat[(size_t)rep * P + j] : 0.0;
    }
    double G[JB][2];
#pragma unroll
    for (int jb = 0; jb < JB; jb++) G[jb][0] = G[jb][1] = 0.0;
    KSum ll;
    unsigned int n_exact = 0;     // (tile, row-slot) steps on which this warp ran the exact evaluation
    __syncthreads();

    for (unsigned long long t0 = rowgroup; t0 < n_tiles; t0 += (unsigned long long)n_rowgroups * WARPS) {
        // ---- phase A: this warp prepares tile t0 + warp * n_rowgroups
        {
            const unsigned long long t = t0 + (unsigned long long)warp * n_rowgroups;
            if (t < n_tiles) {
                // stage the tile (K columns + y) in shared memory for all 8 warps: 16-byte chunk c = lane + 32 i is chunk
                // (lane & 15) of column (lane >> 4) + 2 i
                double* st = stile + warp * tile_d;
                {
                    const double* src = tiles + t * (unsigned long long)(cols * TILE_ROWS) + lane * 2;
                    const unsigned sa = smem_u32(st + (lane >> 4) * CS + (lane & 15) * 2);
                    for (int i = 0; 2 * i <= K; i++) {
                        const int col = (lane >> 4) + 2 * i;
                        if (col <= K)
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa + i * 2 * CS * 8), "l"(src + i * 64));
                    }
                    asm volatile("cp.async.commit_group;");
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    __syncwarp();
                }
                double x0 = 0.0;
                for (int j = 0; j < K; j++) x0 = fma(st[j * CS + lane], s_centre[j], x0);
                const double y = st[K * CS + lane];
                const RowOut o = Fam::eval(x0, y, 1.0, aux);
                double a[TAYLOR_D + 1];
                const bool ok = Taylor<Fam>::coefs(x0, y, o, aux, a) && (t * TILE_ROWS + lane < n_rows);
                a[0] += row_const_of<Fam, USE_COUNTS>(y);
                slab[warp][0][lane] = ok ? x0 : __longlong_as_double(0x7ff8000000000000ll);   // NaN: every pair of the row goes exact
#pragma unroll
                for (int n = 0; n <= TAYLOR_D; n++) slab[warp][1 + n][lane] = a[n];
            }
        }
        __syncthreads();
        // ---- phase B: every warp walks the 8 prepared tiles
        for (int i = 0; i < WARPS; i++) {
            const unsigned long long t = t0 + (unsigned long long)i * n_rowgroups;
            if (t >= n_tiles || rep0 >= m) break;          // (a warp beyond the last replicate only helps with phase A)
            const double* tile = stile + i * tile_d;
            double S[4][2];
#pragma unroll
            for (int nb = 0; nb < 4; nb++) S[nb][0] = S[nb][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ks++) {
                const int j = ks * 4 + q;
                double xf[4];
#pragma unroll
                for (int nb = 0; nb < 4; nb++) xf[nb] = (j < K) ? tile[j * CS + nb * 8 + g] : 0.0;
#pragma unroll
                for (int nb = 0; nb < 4; nb++) dmma_8x8x4(S[nb][0], S[nb][1], bfrag[ks], xf[nb]);
            }
            unsigned long long cbits = (rep < m) ? 0x0101010101010101ull : 0ull;   // a lane without a replicate contributes nothing
            if (USE_COUNTS)
                cbits = (rep < m) ? __ldg(reinterpret_cast<const unsigned long long*>(
                                        counts + (t * (unsigned long long)m_pad + rep) * TILE_ROWS + q * 8))
                                  : 0ull;
            double lsum = 0.0;
#pragma unroll
            for (int nb = 0; nb < 4; nb++) {
                double2 c[NC];
#pragma unroll
                for (int n = 0; n < NC; n++) c[n] = *reinterpret_cast<const double2*>(&slab[i][n][nb * 8 + 2 * q]);
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const unsigned long long row = t * TILE_ROWS + nb * 8 + 2 * q + e;
                    const double cnt = (row < n_rows) ? (double)((cbits >> (8 * (nb * 2 + e))) & 0xffull) : 0.0;
#define APB_C(n) (e ? c[n].y : c[n].x)
                    const double xb = S[nb][e];
                    const double dl = xb - APB_C(0);
                    // Horner for the polynomial and, from the same partial results, for its derivative
                    // (dv_k = dv_{k-1} dl + p_k): 11 FMA as two separate recurrences would take, without the
                    // five extra coefficients n a_n to read back per row
                    double llv = fma(dl, APB_C(7), APB_C(6));
                    double dv = APB_C(7);
                    if (WANT_SCORE) dv = fma(dl, dv, llv);
                    llv = fma(dl, llv, APB_C(5));
                    if (WANT_SCORE) dv = fma(dl, dv, llv);
                    llv = fma(dl, llv, APB_C(4));
                    if (WANT_SCORE) dv = fma(dl, dv, llv);
                    llv = fma(dl, llv, APB_C(3));
                    if (WANT_SCORE) dv = fma(dl, dv, llv);
                    llv = fma(dl, llv, APB_C(2));
                    if (WANT_SCORE) dv = fma(dl, dv, llv);
                    llv = fma(dl, llv, APB_C(1));
#undef APB_C
                    // outside the radius (also: the NaN of a row outside the Taylor range): the exact evaluation.  The
                    // branch is taken by the whole warp (Fam::eval votes warp-wide for its fast exponential), each
                    // lane keeps what it needs
                    const bool exact = !(fabs(dl) <= delta_max) && cnt != 0.0;
                    if (__any_sync(0xffffffffu, exact)) {
                        n_exact++;
                        const double2 y2 = *reinterpret_cast<const double2*>(tile + K * CS + nb * 8 + 2 * q);
                        const RowOut o = Fam::eval(xb, e ? y2.y : y2.x, 1.0, aux);
                        llv = exact ? o.s[0] + row_const_of<Fam, USE_COUNTS>(e ? y2.y : y2.x) : llv;
                        dv = exact ? o.d : dv;
                    }
                    lsum += cnt != 0.0 ? cnt * llv : 0.0;
                    S[nb][e] = cnt != 0.0 ? cnt * dv : 0.0;            // S now holds D^T
                }
            }
            ll.add(lsum);
            if (WANT_SCORE) {
#pragma unroll
                for (int jb = 0; jb < JB; jb++) {
                    const int j = jb * 8 + g;
#pragma unroll
                    for (int nb = 0; nb < 4; nb++) {
                        double2 x2 = make_double2(0.0, 0.0);
                        if (j < K) x2 = *reinterpret_cast<const double2*>(tile + j * CS + nb * 8 + 2 * q);
                        dmma_8x8x4(G[jb][0], G[jb][1], S[nb][0], x2.x);
                        dmma_8x8x4(G[jb][0], G[jb][1], S[nb][1], x2.y);
                    }
                }
            }
        }
        __syncthreads();      // the slab is rewritten by the next phase A
    }

    const int NOUT = 1 + KB;
    double v = ll.value();
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (rep < m) {
        double* dst = partials + ((size_t)rowgroup * m + rep) * NOUT;
        if (q == 0) dst[0] = v;
        if (WANT_SCORE) {
#pragma unroll
            for (int jb = 0; jb < JB; jb++) {
                dst[1 + jb * 8 + 2 * q] = G[jb][0];
                dst[1 + jb * 8 + 2 * q + 1] = G[jb][1];
            }
        }
    }
    if (lane == 0 && n_exact) atomicAdd(exact_slots, (unsigned long long)n_exact);   // the host's choice of kernel for the next passes
}

// sum the row-group partials in fixed order: out[rep][0..KB] (compensated)
__global__ void batched_reduce_kernel(const double* __restrict__ partials, int n_rowgroups, int m, int nout,
                                      double* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)m * nout) return;
    double s = 0.0, c = 0.0;
    for (int rg = 0; rg < n_rowgroups; rg++) {
        double e;
        two_sum(s, partials[(size_t)rg * m * nout + idx], s, e);
        c += e;
    }
    out[idx] = s + c;
}

// exact multinomial resampling: replicate r draws n_global rows with replacement
// (apop_bootstrap.m4.c:144-147 draws gsl_rng_uniform_int(rng, N) per row); the count of
// each resident row lands in the layout stage 2 reads.  One thread per draw.
struct U4c { uint32_t x, y, z, w; };
__device__ __forceinline__ U4c philox_c(U4c c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = U4c{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c;
}
__global__ void __launch_bounds__(256)
bootstrap_counts_kernel(unsigned int* __restrict__ words, unsigned long long n_local, unsigned long long row_offset,
                        unsigned long long n_global, int m, int m_pad, unsigned long long seed) {
    const unsigned long long draw = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int r0 = blockIdx.y * 4;
    if (draw >= n_global) return;
    // 4 replicates per Philox call
    const U4c u = philox_c(U4c{(uint32_t)draw, (uint32_t)(draw >> 32), (uint32_t)blockIdx.y, 0xB007u},
                           (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint32_t rnd[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const int rep = r0 + e;
        if (rep >= m) break;
        // uniform row in [0, n_global): 32 random bits extended with the draw's low bits for n > 2^32
        const unsigned long long wide = ((unsigned long long)rnd[e] << 32) | (rnd[(e + 1) & 3] ^ (uint32_t)draw * 0x9E3779B9u);
        const unsigned long long row = (unsigned long long)(((unsigned __int128)wide * n_global) >> 64);
        if (row < row_offset || row >= row_offset + n_local) continue;   // another rank's row
        const unsigned long long lr = row - row_offset;
        const unsigned long long byte = ((lr >> 5) * (unsigned long long)m_pad + rep) * TILE_ROWS + count_pos((int)(lr & 31));
        atomicAdd(words + (byte >> 2), 1u << (8 * (byte & 3)));
    }
}

template <int KB, class Fam>
static cudaError_t launch_fam(const BatchedLaunch& L) {
    const double4 aux = make_double4(L.aux[0], L.aux[1], L.aux[2], L.aux[3]);
    if (L.centre) {
        const size_t tsm = (size_t)(BATCH_THREADS / 32) * (8 * TILE_ROWS + (size_t)(L.k + 1) * BATCH_TILE_CS) * sizeof(double);
#define APB_TGO(S, C) do { cudaFuncSetAttribute(batched_taylor_kernel<KB, Fam, S, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm); \
        batched_taylor_kernel<KB, Fam, S, C><<<L.grid, BATCH_THREADS, tsm, L.stream>>>(L.tiles, L.n_rows, L.n_tiles, L.k, L.cols, \
        L.B, L.P, L.m, L.counts, L.m_pad, aux, L.n_rowgroups, L.partials, L.centre, L.delta_max, L.exact_slots); } while (0)
        if (L.want_score) { if (L.counts) APB_TGO(true, true); else APB_TGO(true, false); }
        else { if (L.counts) APB_TGO(false, true); else APB_TGO(false, false); }
#undef APB_TGO
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        const int nout = 1 + KB;
        const size_t total = (size_t)L.m * nout;
        batched_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, L.stream>>>(L.partials, L.n_rowgroups, L.m, nout, L.out);
        return cudaGetLastError();
    }
#define APB_GO(S, C)                                                                                          \
    do { if (L.mb == 1) batched_kernel<KB, Fam, S, C, 1><<<L.grid, BATCH_THREADS, 0, L.stream>>>(L.tiles, L.n_rows, L.n_tiles, L.k, L.cols, \
        L.B, L.P, L.m, L.counts, L.m_pad, aux, L.n_rowgroups, L.partials); \
    else batched_kernel<KB, Fam, S, C, 2><<<L.grid, BATCH_THREADS, 0, L.stream>>>(L.tiles, L.n_rows, L.n_tiles, L.k, L.cols, \
        L.B, L.P, L.m, L.counts, L.m_pad, aux, L.n_rowgroups, L.partials); } while (0)
    if (L.want_score) { if (L.counts) APB_GO(true, true); else APB_GO(true, false); }
    else { if (L.counts) APB_GO(false, true); else APB_GO(false, false); }
#undef APB_GO
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int nout = 1 + KB;
    const size_t total = (size_t)L.m * nout;
    batched_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, L.stream>>>(L.partials, L.n_rowgroups, L.m, nout, L.out);
    return cudaGetLastError();
}
template <int KB>
static cudaError_t launch_kb(int fam, const BatchedLaunch& L) {
    switch (fam) {
    case GLM_PROBIT: return launch_fam<KB, FamProbit>(L);
    case GLM_LOGIT2: return launch_fam<KB, FamLogit2>(L);
    case GLM_POISSON_REG: return launch_fam<KB, FamPoissonReg>(L);
    default: return cudaErrorInvalidValue;
    }
}
int batched_kb(int k) { return k <= 8 ? 8 : k <= 16 ? 16 : k <= 24 ? 24 : 32; }
cudaError_t batched_launch(int fam, const BatchedLaunch& L) {
    switch (batched_kb(L.k)) {
    case 8: return launch_kb<8>(fam, L);
    case 16: return launch_kb<16>(fam, L);
    case 24: return launch_kb<24>(fam, L);
    case 32: return launch_kb<32>(fam, L);
    }
    return cudaErrorInvalidValue;
}
cudaError_t bootstrap_counts_launch(unsigned char* counts, size_t n_local, size_t row_offset, size_t n_global, int m,
                                    int m_pad, unsigned long long seed, cudaStream_t st) {
    if (!n_global || !m) return cudaSuccess;
    dim3 grid((unsigned)((n_global + 255) / 256), (unsigned)((m + 3) / 4));
    bootstrap_counts_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<unsigned int*>(counts), n_local, row_offset, n_global, m, m_pad, seed);
    return cudaGetLastError();
}

}  // namespace apb
