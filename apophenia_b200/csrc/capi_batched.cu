// capi_batched.cu -- the k x k passes (information, normal equations) and the batched-beta path
// with its resident resample counts.
#include "capi_internal.cuh"

using namespace apb;

// H = X' diag(w) [X | y] in one pass; out has K*(K+1) doubles, H[j][c] at j*(K+1)+c
bool eval_xtwx(Resident* r, int mode, const double* beta, const double* aux, std::vector<double>& out) {
    Runtime& R = rt();
    const int K = r->k;
    if (K < 1 || K > MAX_K_REG) return set_error("k = %d columns: the fused kernels cover 1..%d", K, MAX_K_REG);
    if (!r->has_y) return set_error("the data set has no outcome vector (run the model's prep first)");
    static int bps_cache[XTWX_MODE_COUNT][MAX_K_REG + 1];
    int& bps = bps_cache[mode][K];
    if (!bps) bps = xtwx_blocks_per_sm(mode, K);
    if (bps <= 0) return set_error("X'WX kernel unavailable for k=%d", K);
    const int KB = xtwx_kb(K), NC = KB + 1, NOUT = KB * NC;
    XtwxLaunch L;
    L.tiles = r->tiles; L.n_rows = r->n_rows; L.n_tiles = r->n_tiles; L.k = K; L.cols = r->cols;
    L.has_w = (r->has_w && mode == XTWX_GRAM) ? 1 : 0;
    L.grid = grid_for(bps, r->n_tiles, XTWX_THREADS / 32);
    if (!ensure_partials((size_t)L.grid * NOUT)) return false;
    L.partials = R.d_partials; L.ticket = R.d_ticket; L.out = R.d_out; L.stream = R.stream;
    L.fin = make_fin(NOUT);
    XtwxParams prm;
    memset(&prm, 0, sizeof prm);
    if (beta) memcpy(prm.beta, beta, K * sizeof(double));
    if (aux) memcpy(prm.aux, aux, 4 * sizeof(double));
    time_begin();
    if (!cuda_ok(xtwx_launch(mode, L, prm), "X'WX kernel launch")) return false;
    time_end_record();
    R.launches++;
    if (!reduce_and_fetch(NOUT, L.fin)) return false;
    time_collect();
    out.assign((size_t)K * (K + 1), 0.0);
    for (int j = 0; j < K; j++) {
        for (int c = 0; c < K; c++) out[(size_t)j * (K + 1) + c] = R.h_out[j * NC + c];
        out[(size_t)j * (K + 1) + K] = R.h_out[j * NC + KB];
    }
    return true;
}

extern "C" int apop_b200_ols_gram(apop_data* d, double* xtx, double* xty) {
    LOCK;
    if (!d || !xtx || !xty) { set_error("apop_b200_ols_gram: NULL argument"); return 1; }
    Resident* r = acquire(d);
    if (!r) return 1;
    std::vector<double> out;
    const bool ok = eval_xtwx(r, XTWX_GRAM, nullptr, nullptr, out);
    if (ok) {
        const int K = r->k;
        for (int j = 0; j < K; j++) {
            for (int c = 0; c < K; c++) xtx[j * K + c] = out[(size_t)j * (K + 1) + c];
            xty[j] = out[(size_t)j * (K + 1) + K];
        }
    }
    release(r);
    return ok ? 0 : 1;
}

// Observed information of the multinomial logit, P x P in pack order (index j*(C-1)+c), from
// ceil((C-1) C / 8) launches of the pair-block kernel (logit_info.cu).
static bool eval_logit_info(Resident* r, const double* B, size_t C1, double* H) {
    Runtime& R = rt();
    const int K = r->k;
    if (K < 1 || K > MAX_K_REG) return set_error("k = %d columns: the information kernel covers 1..%d", K, MAX_K_REG);
    if (C1 < 2 || C1 > (size_t)LOGIT_MAX_C1) return set_error("%zu outcome categories: the multinomial kernels cover 3..%d", C1 + 1, LOGIT_MAX_C1 + 1);
    static int bps_cache[MAX_K_REG + 1][LOGIT_MAX_C1 + 1];
    int& bps = bps_cache[K][C1];
    if (!bps) bps = logit_info_blocks_per_sm(K, (int)C1);
    if (bps <= 0) return set_error("multinomial information kernel unavailable for k=%d, C=%zu", K, C1 + 1);
    const int KB = logit_kb(K), NOUT = LOGIT_INFO_NP * KB * KB;
    const size_t P = (size_t)K * C1;
    // the information follows a multinomial fit, whose kernel keeps the tiles swizzled; the bulk-copy form of the
    // information kernel reads that layout (APOP_B200_INFO_PLAIN=1: convert back and take the cp.async form, for A/B runs)
    static const bool plain = getenv("APOP_B200_INFO_PLAIN") != nullptr;
    if (!ensure_layout(r, !plain)) return false;
    LogitParams prm;
    memset(&prm, 0, sizeof prm);
    memcpy(prm.B, B, P * sizeof(double));
    std::vector<std::pair<int, int>> pairs;
    for (int a = 0; a < (int)C1; a++)
        for (int b = a; b < (int)C1; b++) pairs.push_back({a, b});
    for (size_t p0 = 0; p0 < pairs.size(); p0 += LOGIT_INFO_NP) {
        LogitInfoPairs pr;
        for (int i = 0; i < LOGIT_INFO_NP; i++) {
            const bool live = p0 + i < pairs.size();
            pr.a[i] = live ? pairs[p0 + i].first : -1;
            pr.b[i] = live ? pairs[p0 + i].second : -1;
        }
        LogitLaunch L;
        L.tiles = r->tiles; L.n_rows = r->n_rows; L.n_tiles = r->n_tiles; L.k = K; L.cols = r->cols;
        L.swz = r->swz ? 1 : 0;   // the kernel stages either resident form
        L.grid = grid_for(bps, r->n_tiles, LOGIT_INFO_THREADS / 32);
        if (!ensure_partials((size_t)L.grid * NOUT)) return false;
        L.partials = R.d_partials; L.ticket = R.d_ticket; L.out = R.d_out; L.stream = R.stream;
        L.fin = make_fin(NOUT);
        time_begin();
        if (!cuda_ok(logit_info_launch((int)C1, L, prm, pr), "multinomial information launch")) return false;
        time_end_record();
        R.launches++;
        if (!reduce_and_fetch(NOUT, L.fin)) return false;
        time_collect();
        for (int i = 0; i < LOGIT_INFO_NP && p0 + i < pairs.size(); i++) {
            const int a = pr.a[i], b = pr.b[i];
            const double* blk = R.h_out + (size_t)i * KB * KB;
            for (int j = 0; j < K; j++)
                for (int j2 = 0; j2 < K; j2++) {
                    // X' diag(w) X is symmetric; the two triangles of a diagonal 8 x 8 block come from different
                    // products (fl(w x_j) x_j2 vs fl(w x_j2) x_j): take the upper one for both, H == H' bitwise
                    const double v = (j <= j2) ? blk[j * KB + j2] : blk[j2 * KB + j];
                    H[((size_t)j * C1 + a) * P + (size_t)j2 * C1 + b] = v;
                    H[((size_t)j * C1 + b) * P + (size_t)j2 * C1 + a] = v;   // W_ab = W_ba
                }
        }
    }
    return true;
}

extern "C" int apop_b200_information(apop_b200_family family, apop_data* d, const double* beta, size_t P, double* H) {
    LOCK;
    if (!d || !beta || !H) { set_error("apop_b200_information: NULL argument"); return 1; }
    Resident* r = acquire(d, family == APOP_B200_LOGIT ? LAYOUT_ANY : LAYOUT_PLAIN);
    if (!r) return 1;
    bool ok = false;
    std::vector<double> out;
    double scale = 1.0;
    if (family == APOP_B200_LOGIT && r->k > 0 && P > (size_t)r->k && P % (size_t)r->k == 0) {
        const bool okm = eval_logit_info(r, beta, P / (size_t)r->k, H);   // multinomial: P = k (C-1)
        release(r);
        return okm ? 0 : 1;
    }
    if (!ensure_layout(r, false)) { release(r); return 1; }
    if ((size_t)r->k != P) set_error("apop_b200_information: P = %zu but the data has %d columns", P, r->k);
    else if (family == APOP_B200_PROBIT) ok = eval_xtwx(r, XTWX_PROBIT, beta, nullptr, out);
    else if (family == APOP_B200_LOGIT) ok = eval_xtwx(r, XTWX_LOGIT2, beta, nullptr, out);
    else if (family == APOP_B200_POISSON_REG) ok = eval_xtwx(r, XTWX_POISSON_REG, beta, nullptr, out);
    else if (family == APOP_B200_OLS) {
        // -d2LL/dbeta2 at fixed sigma: X'WX / sigma^2, sigma^2 = sample variance of the errors
        std::vector<double> mom;
        double n, mc, vc;
        if (eval_glm(r, GLM_OLS, beta, P, nullptr, mom) && global_counts(r, n, mc, vc) && eval_xtwx(r, XTWX_GRAM, nullptr, nullptr, out)) {
            const long double var = ((long double)mom[1] - (long double)mom[0] * mom[0] / n) / (n - 1);
            scale = (double)(1.0L / var);
            ok = true;
        }
    } else set_error("apop_b200_information: family %d has no k x k information here", (int)family);
    if (ok)
        for (size_t j = 0; j < P; j++)
            for (size_t c = 0; c < P; c++) H[j * P + c] = out[j * (P + 1) + c] * scale;
    release(r);
    return ok ? 0 : 1;
}
// m parameter vectors in one pass on the DMMA path.  ll[m]; score[P*m] replicate-major.
extern "C" int apop_b200_ll_score_batched(apop_b200_family family, apop_data* d, const double* B, size_t P, size_t m,
                                          int use_counts, double* ll, double* score) {
    LOCK;
    if (!d || !B || !ll || !m) { set_error("apop_b200_ll_score_batched: NULL argument"); return 1; }
    GlmFam fam;
    if (family == APOP_B200_PROBIT) fam = GLM_PROBIT;
    else if (family == APOP_B200_LOGIT) fam = GLM_LOGIT2;
    else if (family == APOP_B200_POISSON_REG) fam = GLM_POISSON_REG;
    else { set_error("batched path: family %d is not a binary-outcome / GLM family", (int)family); return 1; }
    Resident* r = acquire(d);
    if (!r) return 1;
    Runtime& R = rt();
    bool ok = false;
    double *dB = nullptr, *dOut = nullptr;
    do {
        const int K = r->k;
        if ((size_t)K != P) { set_error("batched path: P = %zu but the data has %d columns", P, K); break; }
        if (K < 1 || K > MAX_K_REG || !r->has_y) { set_error("batched path: needs 1..%d columns and an outcome vector", MAX_K_REG); break; }   // a weights column is ignored, as these families do
        if (use_counts && (!r->counts || r->counts_m < m)) { set_error("batched path: no resident resample counts for %zu replicates (call apop_b200_bootstrap_counts)", m); break; }
        const int KB = batched_kb(K), NOUT = 1 + KB;
        static int mb_env = -1;
        if (mb_env < 0) { const char* e = getenv("APOP_B200_BATCH_MB"); mb_env = (e && e[0] == '2') ? 2 : 1; }
        const int mb = mb_env, ctas_per_sm = (mb == 1) ? 2 : 1;   // matches the kernels' launch bounds
        const int reps_per_cta = (BATCH_THREADS / 32) * mb * 8;
        const int n_repgroups = (int)((m + reps_per_cta - 1) / reps_per_cta);
        long long n_rowgroups = (long long)R.num_sms * ctas_per_sm / n_repgroups;
        if (n_rowgroups < 1) n_rowgroups = 1;
        if ((unsigned long long)n_rowgroups > r->n_tiles) n_rowgroups = r->n_tiles ? (long long)r->n_tiles : 1;
        BatchedLaunch L;
        L.tiles = r->tiles; L.n_rows = r->n_rows; L.n_tiles = r->n_tiles; L.k = K; L.cols = r->cols;
        L.P = (int)P; L.m = (int)m; L.counts = use_counts ? r->counts : nullptr; L.m_pad = (int)r->counts_m;
        // binary logit: the two category values (defaults 0, 1)
        L.aux[0] = 0.0; L.aux[1] = 1.0; L.aux[2] = L.aux[3] = 0.0;
        const apop_data* pg = category_page(d);
        if (fam == GLM_LOGIT2 && pg && pg->vector && pg->vector->size >= 2) { L.aux[0] = pg->vector->data[0]; L.aux[1] = pg->vector->data[pg->vector->stride]; }
        L.n_rowgroups = (int)n_rowgroups; L.grid = (int)(n_rowgroups * n_repgroups);
        L.want_score = score != nullptr; L.stream = R.stream; L.mb = mb;
        if (!ensure_partials((size_t)n_rowgroups * m * NOUT)) break;
        L.partials = R.d_partials;
        // The shared expansion point of the Taylor form of stage 2 (batched_kernel.cu): the mean of the m parameter
        // vectors.  APOP_B200_BATCH_TAYLOR=0 evaluates every (row, replicate) pair exactly instead;
        // APOP_B200_BATCH_DELTA overrides the radius (default 1.5e-2: truncation error < 1e-16 relative).
        static int taylor_env = -1;
        static double delta_env = 1.5e-2;
        if (taylor_env < 0) {
            const char* e = getenv("APOP_B200_BATCH_TAYLOR"); taylor_env = (e && e[0] == '0') ? 0 : 1;
            if (const char* d2 = getenv("APOP_B200_BATCH_DELTA")) { const double v = atof(d2); if (v > 0) delta_env = v; }
        }
        bool taylor = taylor_env == 1 && mb == 1;
        if (taylor && r->batched_exact_passes > 0) { r->batched_exact_passes--; taylor = false; }   // the vectors were too far apart last time
        std::vector<double> Bc(B, B + P * m);
        if (taylor) {
            std::vector<long double> c(P, 0.0L);
            for (size_t q = 0; q < m; q++) for (size_t j = 0; j < P; j++) c[j] += B[q * P + j];
            for (size_t j = 0; j < P; j++) Bc.push_back((double)(c[j] / (long double)m));
        }
        if (!cuda_ok(cudaMalloc(&dB, Bc.size() * sizeof(double)), "cudaMalloc(B)")) break;
        if (!cuda_ok(cudaMalloc(&dOut, m * NOUT * sizeof(double)), "cudaMalloc(out)")) break;
        if (!cuda_ok(cudaMemcpyAsync(dB, Bc.data(), Bc.size() * sizeof(double), cudaMemcpyHostToDevice, R.stream), "H2D B")) break;
        L.B = dB; L.out = dOut;
        unsigned long long* d_exact = reinterpret_cast<unsigned long long*>(R.d_vals);
        if (taylor) {
            L.centre = dB + P * m; L.delta_max = delta_env; L.exact_slots = d_exact;
            if (!cuda_ok(cudaMemsetAsync(d_exact, 0, sizeof(unsigned long long), R.stream), "memset")) break;
        }
        time_begin();
        if (!cuda_ok(batched_launch((int)fam, L), "batched kernel launch")) break;
        time_end_record();
        R.launches += 2;
        if (R.comm && R.nranks > 1 && !R.local_only &&
            !nccl_ok(R.nccl.AllReduce(dOut, dOut, m * NOUT, 8, 0, R.comm, R.stream), "ncclAllReduce")) break;
        std::vector<double> h(m * NOUT);
        if (!cuda_ok(cudaMemcpyAsync(h.data(), dOut, m * NOUT * sizeof(double), cudaMemcpyDeviceToHost, R.stream), "D2H")) break;
        unsigned long long n_exact = 0;
        if (taylor && !cuda_ok(cudaMemcpyAsync(&n_exact, d_exact, sizeof n_exact, cudaMemcpyDeviceToHost, R.stream), "D2H")) break;
        if (!cuda_ok(cudaStreamSynchronize(R.stream), "stream sync")) break;
        time_collect();
        if (taylor) {
            // every warp makes 8 row-slot steps per tile for its 8 replicates; when more than a quarter of them needed the
            // exact evaluation the shared-centre form costs more than it saves: the next 32 passes go exact, then re-probe
            const double steps = (double)r->n_tiles * 8.0 * (double)((m + 7) / 8);
            r->batched_last_exact_fraction = steps > 0 ? (double)n_exact / steps : 0.0;
            if (r->batched_last_exact_fraction > 0.25) r->batched_exact_passes = 32;
        }
        double lg = 0.0;
        // Poisson regression: -sum lgamma(y+1) is one constant for a plain pass; with resample counts it is weighted per
        // replicate and the kernels carry it per row
        if (fam == GLM_POISSON_REG && !use_counts && !lgamma_const(r, lg)) break;
        for (size_t q = 0; q < m; q++) {
            ll[q] = h[q * NOUT] - lg;
            if (score) memcpy(score + q * P, h.data() + q * NOUT + 1, P * sizeof(double));
        }
        ok = true;
    } while (0);
    if (dB) cudaFree(dB);
    if (dOut) cudaFree(dOut);
    release(r);
    return ok ? 0 : 1;
}

// share of the warp steps of the last shared-centre batched pass over d that needed the exact evaluation
// (0: every pair was within the Taylor radius; > 0.25 sends the next passes to the exact kernel); -1 if unknown
extern "C" double apop_b200_batched_exact_fraction(const apop_data* d) {
    LOCK;
    auto it = rt().reg.find(d);
    return it == rt().reg.end() ? -1.0 : it->second->batched_last_exact_fraction;
}

// Select one bootstrap replicate on a resident data set with resample counts: the single-beta multinomial logit passes
// (apop_b200_logit_ll / _score through any unmodified optimizer) then weight every row by that replicate's multiplicity,
// i.e. evaluate the resampled data set of apop_bootstrap.m4.c:143-149 without copying a row.  replicate < 0: off.
extern "C" int apop_b200_set_replicate(apop_data* d, long replicate) {
    LOCK;
    auto it = rt().reg.find(d);
    if (it == rt().reg.end()) { set_error("apop_b200_set_replicate: the data set is not resident"); return 1; }
    Resident* r = it->second;
    if (replicate >= 0 && (!r->counts || (size_t)replicate >= r->counts_m)) { set_error("apop_b200_set_replicate: no resident counts for replicate %ld", replicate); return 1; }
    r->active_replicate = replicate < 0 ? -1 : replicate;
    r->memo_valid = false;
    return 0;
}

// resident resample counts back to the host as an n_rows x m row-major uint8 matrix (tests)
extern "C" int apop_b200_counts_download(const apop_data* d, size_t m, unsigned char* out) {
    LOCK;
    Runtime& R = rt();
    auto it = R.reg.find(d);
    if (it == R.reg.end() || !it->second->counts || it->second->counts_m < m) { set_error("no resident counts"); return 1; }
    Resident* r = it->second;
    cudaSetDevice(R.device);
    const size_t m_pad = r->counts_m, bytes = r->n_tiles * m_pad * TILE_ROWS;
    std::vector<unsigned char> h(bytes);
    if (!cuda_ok(cudaMemcpy(h.data(), r->counts, bytes, cudaMemcpyDeviceToHost), "D2H counts")) return 1;
    for (size_t i = 0; i < r->n_rows; i++) {
        const int rr = (int)(i & 31);
        const int pos = ((rr & 7) >> 1) * 8 + (rr >> 3) * 2 + (rr & 1);
        for (size_t q = 0; q < m; q++) out[i * m + q] = h[((i >> 5) * m_pad + q) * TILE_ROWS + pos];
    }
    return 0;
}

// m exact multinomial resamples of the (global) rows, kept resident next to the data
extern "C" int apop_b200_bootstrap_counts(apop_data* d, size_t m, uint64_t seed) {
    LOCK;
    if (!d || !m) { set_error("apop_b200_bootstrap_counts: NULL argument"); return 1; }
    if (!require_ready()) return 1;
    Runtime& R = rt();
    auto it = R.reg.find(d);
    if (it == R.reg.end()) { set_error("apop_b200_bootstrap_counts: pin the data set first"); return 1; }
    Resident* r = it->second;
    cudaSetDevice(R.device);
    if (r->counts) { cudaFree(r->counts); r->counts = nullptr; r->counts_m = 0; }
    const size_t m_pad = (m + 3) / 4 * 4;
    const size_t bytes = r->n_tiles * m_pad * TILE_ROWS;
    if (!bytes) return 0;
    if (!cuda_ok(cudaMalloc(&r->counts, bytes), "cudaMalloc(counts)")) return 1;
    cudaMemsetAsync(r->counts, 0, bytes, R.stream);
    // the resample is of the global data set: ranks share seed, each keeps its own rows
    double n_global = (double)r->n_rows, offset = 0.0;
    if (R.nranks > 1 && !R.local_only) {
        // exclusive prefix of the shard sizes = this rank's first global row
        std::vector<double> sizes(R.nranks, 0.0);
        sizes[R.rank] = (double)r->n_rows;
        if (!allreduce_host(sizes.data(), R.nranks)) return 1;
        n_global = 0;
        for (int q = 0; q < R.nranks; q++) { if (q < R.rank) offset += sizes[q]; n_global += sizes[q]; }
    }
    if (!cuda_ok(bootstrap_counts_launch(r->counts, r->n_rows, (size_t)offset, (size_t)n_global, (int)m, (int)m_pad, seed, R.stream), "bootstrap counts") ||
        !cuda_ok(cudaStreamSynchronize(R.stream), "counts sync")) return 1;
    r->counts_m = m_pad;
    R.launches++;
    return 0;
}

