// capi_models.cu -- the reference-signature log-likelihood / score entry points of every model
// family, on top of the fused-pass evaluators (eval_glm, eval_logit, eval_cells) with their
// memoisation of the last evaluation.
#include "capi_internal.cuh"

using namespace apb;

// --------------------------------------------------------------------------
//  evaluation
// --------------------------------------------------------------------------
int grid_for(int blocks_per_sm, size_t n_tiles, int warps_per_block) {
    Runtime& R = rt();
    long long g = (long long)R.num_sms * (blocks_per_sm > 0 ? blocks_per_sm : 1);
    const long long need = (long long)((n_tiles + warps_per_block - 1) / warps_per_block);
    if (g > need) g = need;
    if (g < 1) g = 1;
    return (int)g;
}

// Per-pass CUDA-event timing is opt-in (apop_b200_set_timing / APOP_B200_TIMING=1): the two event records
// and the elapsed-time query cost ~25 us per pass, 5 % of a strong-scaling pass of 0.5 ms.
void time_begin() { if (rt().timing) cudaEventRecord(rt().ev0, rt().stream); }
void time_end_record() { if (rt().timing) cudaEventRecord(rt().ev1, rt().stream); }
void time_collect() {
    Runtime& R = rt();
    if (!R.timing) { R.n_passes++; return; }
    float ms = 0;
    if (cudaEventElapsedTime(&ms, R.ev0, R.ev1) == cudaSuccess) {
        R.last_ms = ms;
        R.total_ms += ms;
        R.n_passes++;
    }
}

// One fused pass of a GLM family; out receives MAX_ACC + k doubles, summed over
// all ranks.  Served from the memo when (family, beta, aux) repeat.
bool eval_glm(Resident* r, GlmFam fam, const double* beta, size_t P, const double* aux, std::vector<double>& out) {
    Runtime& R = rt();
    const int K = r->k;
    if ((size_t)K != P) return set_error("parameter count %zu does not match the %d data columns", P, K);
    if (K < 1 || K > GLM_WIDE_MAX_K) return set_error("k = %d columns: the fused kernels cover 1..%d", K, GLM_WIDE_MAX_K);
    if (!r->has_y) return set_error("the data set has no outcome vector (run the model's prep first)");
    if (r->active_replicate >= 0) return set_error("a bootstrap replicate is selected on this data set (apop_b200_set_replicate): the binary families take their replicates through apop_b200_ll_score_batched");
    std::vector<double> key(beta, beta + P);
    for (int a = 0; a < 4; a++) key.push_back(aux ? aux[a] : 0.0);
    if (r->memo_valid && r->memo_fam == (int)fam && r->memo_beta.size() == key.size() &&
        memcmp(r->memo_beta.data(), key.data(), key.size() * sizeof(double)) == 0) {
        out = r->memo_out;
        return true;
    }
    // a weights column sets the tile stride for every family, but only OLS reads it: probit, logit and the
    // Poisson regression ignore d->weights exactly as the reference's do (model/apop_probit.c never touches them)
    const bool reads_w = r->has_w && fam == GLM_OLS;
    const bool has_w = r->has_w;
    const bool wide = K > GLM_K_REG;
    static int bps_cache[GLM_FAM_COUNT][GLM_K_REG + 1][2];
    int bps_wide = 0;
    int& bps = wide ? bps_wide : bps_cache[fam][K][has_w ? 1 : 0];
    int& bps_ref = bps;
    if (!bps_ref) bps_ref = wide ? glm_wide_blocks_per_sm(fam, K) : glm_blocks_per_sm(fam, K, has_w);
    if (bps_ref <= 0) return set_error("kernel for family %d, k=%d is not available on this device", (int)fam, K);
    const int NOUT = MAX_ACC + K;
    GlmLaunch L;
    L.tiles = r->tiles; L.n_rows = r->n_rows; L.n_tiles = r->n_tiles;
    L.grid = grid_for(bps_ref, r->n_tiles, wide ? glm_wide_tiles_per_block(K) : GLM_THREADS / 32);
    if (!ensure_partials((size_t)L.grid * NOUT)) return false;
    L.partials = R.d_partials; L.ticket = R.d_ticket; L.out = R.d_out; L.stream = R.stream;
    L.fin = make_fin(NOUT);
    GlmParams prm;
    memset(&prm, 0, sizeof prm);
    if (!wide) memcpy(prm.beta, beta, P * sizeof(double));
    if (aux) memcpy(prm.aux, aux, 4 * sizeof(double));
    if (wide) {   // beta does not fit the kernel parameter: stage it in device memory
        if (!R.d_beta && !cuda_ok(cudaMalloc(&R.d_beta, GLM_WIDE_MAX_K * sizeof(double)), "cudaMalloc(beta)")) return false;
        if (!cuda_ok(cudaMemcpyAsync(R.d_beta, beta, P * sizeof(double), cudaMemcpyHostToDevice, R.stream), "H2D beta")) return false;
    }
    time_begin();
    if (!cuda_ok(wide ? glm_wide_launch(fam, K, r->cols, reads_w, L, R.d_beta, prm.aux) : glm_launch(fam, K, has_w, L, prm),
                 "glm kernel launch")) return false;
    time_end_record();
    R.launches++;
    if (!reduce_and_fetch(NOUT, L.fin)) return false;
    time_collect();
    out.assign(R.h_out, R.h_out + NOUT);
    r->memo_valid = true; r->memo_fam = (int)fam; r->memo_beta = key; r->memo_out = out;
    return true;
}

bool eval_cells(Resident* r, int mode, double p0, double* out5, bool matrix, double p1, double p2) {
    Runtime& R = rt();
    const int grid = grid_for(8, r->n_tiles, CELLS_THREADS / 32);
    if (!ensure_partials((size_t)grid * CELLS_NOUT)) return false;
    const FinalizeCtx fin = make_fin(CELLS_NOUT);
    time_begin();
    if (!cuda_ok(cells_launch(mode, r->tiles, r->n_rows, r->n_tiles, matrix ? r->k : 0, r->has_y ? r->k : -1, r->cols, p0, R.d_partials,
                              R.d_ticket, R.d_out, grid, R.stream, fin, p1, p2), "cells kernel launch")) return false;
    time_end_record();
    R.launches++;
    if (!reduce_and_fetch(CELLS_NOUT, fin)) return false;
    time_collect();
    memcpy(out5, R.h_out, CELLS_NOUT * sizeof(double));
    return true;
}

// global (all ranks) cell counts of a resident data set
bool global_counts(Resident* r, double& n_rows, double& m_cells, double& v_cells) {
    Runtime& R = rt();
    if (r->n_global_epoch != R.comm_epoch) {   // one exchange per (data set, communicator), then cached
        double n = (double)r->n_rows;
        if (R.nranks > 1 && !R.local_only && !allreduce_scalar(&n)) return false;
        r->n_global = n;
        r->n_global_epoch = R.comm_epoch;
    }
    n_rows = r->n_global;
    m_cells = n_rows * r->k;
    v_cells = r->has_y ? n_rows : 0.0;
    return true;
}

// parameters in apop_data_pack order: vector, matrix rows, weights
// (apop_conversions.m4.c:762-806)
bool pack_parameters(const apop_data* p, std::vector<double>& out) {
    out.clear();
    if (!p) return false;
    if (p->vector)
        for (size_t i = 0; i < p->vector->size; i++) out.push_back(p->vector->data[i * p->vector->stride]);
    if (p->matrix)
        for (size_t i = 0; i < p->matrix->size1; i++)
            for (size_t j = 0; j < p->matrix->size2; j++) out.push_back(p->matrix->data[i * p->matrix->tda + j]);
    if (p->weights)
        for (size_t i = 0; i < p->weights->size; i++) out.push_back(p->weights->data[i * p->weights->stride]);
    return !out.empty();
}

// the factor page apop_data_to_factors leaves behind ("<categories for ...>",
// apop_regression.m4.c:135-163); NULL if the data carries none
const apop_data* category_page(const apop_data* d) {
    for (const apop_data* pg = d ? d->more : nullptr; pg; pg = pg->more)
        if (pg->names && pg->names->title && !strncmp(pg->names->title, "<categories for", 15)) return pg;
    return nullptr;
}
const apop_data* named_page(const apop_data* d, const char* title) {
    for (const apop_data* pg = d ? d->more : nullptr; pg; pg = pg->more)
        if (pg->names && pg->names->title && !strcasecmp(pg->names->title, title)) return pg;
    return nullptr;
}

// apop_data_get_page (apop_data.m4.c:1298-1304) also matches the data set itself
const apop_data* named_page_or_self(const apop_data* d, const char* title) {
    for (const apop_data* pg = d; pg; pg = pg->more)
        if (pg->names && pg->names->title && !strcasecmp(pg->names->title, title)) return pg;
    return nullptr;
}

void fill_nan(gsl_vector* g) {
    if (g && g->data)
        for (size_t i = 0; i < g->size; i++) g->data[i * g->stride] = NAN;
}
void store_score(gsl_vector* g, const double* src, size_t n, double scale) {
    for (size_t i = 0; i < n && i < g->size; i++) g->data[i * g->stride] = src[i] * scale;
}

// ---- GLM-shaped families: probit, binary logit, poisson regression ---------
struct GlmEval {
    bool ok = false;
    std::vector<double> out;  // MAX_ACC + k
    std::vector<double> beta;
    double n_rows = 0;        // global
    Resident* r = nullptr;
};

static GlmEval glm_entry(GlmFam fam, apop_data* d, apop_model* m, const double* aux) {
    GlmEval e;
    if (!m || !m->parameters || !d) { set_error("NULL model, parameters or data"); return e; }
    if (!pack_parameters(m->parameters, e.beta)) { set_error("empty parameters"); return e; }
    Resident* r = acquire(d);
    if (!r) { d->error = 'a'; return e; }
    e.ok = eval_glm(r, fam, e.beta.data(), e.beta.size(), aux, e.out);
    if (!e.ok && m) m->error = 'd';
    e.r = r;
    return e;
}

extern "C" long double apop_b200_probit_ll(apop_data* d, apop_model* m) {
    LOCK;
    const apop_data* cats = category_page(d);
    if (cats && cats->vector && cats->vector->size != 2) {
        set_error("multinomial probit (%zu categories) is outside this library's path (SURVEY F4)", cats->vector->size);
        if (m) m->error = 'd';
        return NAN;
    }
    GlmEval e = glm_entry(GLM_PROBIT, d, m, nullptr);
    const long double ll = e.ok ? (long double)e.out[0] : (long double)NAN;
    release(e.r);
    return ll;
}
extern "C" void apop_b200_probit_score(apop_data* d, gsl_vector* g, apop_model* m) {
    LOCK;
    if (!g) return;
    GlmEval e = glm_entry(GLM_PROBIT, d, m, nullptr);
    if (e.ok) store_score(g, e.out.data() + MAX_ACC, e.beta.size(), 1.0);
    else fill_nan(g);
    release(e.r);
}

// One fused multinomial-logit pass; out = [LL | G in pack order (j*(C-1)+c)], all ranks summed.
bool eval_logit(Resident* r, const std::vector<double>& B, size_t C1, const std::vector<double>& cats,
                       std::vector<double>& out) {
    Runtime& R = rt();
    const int K = r->k;
    if (B.size() != (size_t)K * C1) return set_error("parameters are %zu doubles; the data has %d columns x %zu free categories", B.size(), K, C1);
    if (K < 1 || K > MAX_K_REG) return set_error("k = %d columns: the fused kernels cover 1..%d", K, MAX_K_REG);
    if (C1 < 2 || C1 > (size_t)LOGIT_MAX_C1) return set_error("%zu outcome categories: the multinomial kernel covers 3..%d", C1 + 1, LOGIT_MAX_C1 + 1);
    if (!r->has_y) return set_error("logit needs an outcome vector (run the model's prep first)");
    std::vector<double> key(B);
    key.insert(key.end(), cats.begin(), cats.end());
    key.push_back((double)r->active_replicate);
    if (r->active_replicate >= 0 && (!r->counts || (size_t)r->active_replicate >= r->counts_m))
        return set_error("replicate %ld selected but the data set has resample counts for %zu replicates (apop_b200_bootstrap_counts)", r->active_replicate, r->counts_m);
    if (r->memo_valid && r->memo_fam == 200 + (int)C1 && r->memo_beta.size() == key.size() &&
        memcmp(r->memo_beta.data(), key.data(), key.size() * sizeof(double)) == 0) {
        out = r->memo_out;
        return true;
    }
    static int bps_cache[MAX_K_REG + 1][LOGIT_MAX_C1 + 1];
    int& bps = bps_cache[K][C1];
    if (!bps) bps = logit_blocks_per_sm(K, (int)C1);
    if (bps <= 0) return set_error("multinomial logit kernel unavailable for k=%d, C=%zu", K, C1 + 1);
    const int KB = logit_kb(K);
    const int NOUT = 1 + KB * (int)C1;
    if (!ensure_layout(r, logit_wants_swizzled())) return false;
    LogitLaunch L;
    L.tiles = r->tiles; L.n_rows = r->n_rows; L.n_tiles = r->n_tiles; L.k = K; L.cols = r->cols;   // a weights column only widens the tile stride
    L.swz = r->swz ? 1 : 0;
    L.grid = grid_for(bps, r->n_tiles, logit_warps_per_block());
    if (r->active_replicate >= 0) { L.counts = r->counts; L.m_pad = (int)r->counts_m; L.replicate = (int)r->active_replicate; }
    if (!ensure_partials((size_t)L.grid * NOUT)) return false;
    L.partials = R.d_partials; L.ticket = R.d_ticket; L.out = R.d_out; L.stream = R.stream;
    L.fin = make_fin(NOUT);
    LogitParams prm;
    memset(&prm, 0, sizeof prm);
    memcpy(prm.B, B.data(), B.size() * sizeof(double));
    for (size_t c = 0; c < C1 && c < 8; c++) prm.cats[c] = cats[c];
    time_begin();
    if (!cuda_ok(logit_launch((int)C1, L, prm), "logit kernel launch")) return false;
    time_end_record();
    R.launches++;
    if (!reduce_and_fetch(NOUT, L.fin)) return false;
    time_collect();
    out.assign(R.h_out, R.h_out + 1 + (size_t)K * C1);
    r->memo_valid = true; r->memo_fam = 200 + (int)C1; r->memo_beta = key; r->memo_out = out;
    return true;
}

static bool logit_categories(apop_data* d, apop_model* m, size_t C, std::vector<double>& cats) {
    const apop_data* pg = category_page(m && m->data ? m->data : d);
    if (!pg) pg = category_page(d);
    cats.resize(C);
    for (size_t c = 0; c < C; c++) cats[c] = (double)c;
    if (pg && pg->vector && pg->vector->size >= C)
        for (size_t c = 0; c < C; c++) cats[c] = pg->vector->data[c * pg->vector->stride];
    return true;
}

extern "C" long double apop_b200_logit_ll(apop_data* d, apop_model* m) {
    LOCK;
    if (!m || !m->parameters || !d || !d->matrix) { set_error("NULL model, parameters or data"); return NAN; }
    const size_t C1 = m->parameters->matrix ? m->parameters->matrix->size2 : 1;
    std::vector<double> cats;
    logit_categories(d, m, C1 + 1, cats);
    if (C1 == 1) {
        const double aux[4] = {cats[0], cats[1], 0, 0};
        GlmEval e = glm_entry(GLM_LOGIT2, d, m, aux);
        const long double ll = e.ok ? (long double)e.out[0] : (long double)NAN;
        release(e.r);
        return ll;
    }
    std::vector<double> B, out;
    pack_parameters(m->parameters, B);
    Resident* r = acquire(d, LAYOUT_ANY);   // eval_logit converts to the layout its kernel reads
    if (!r) { d->error = 'a'; return NAN; }
    const bool ok = eval_logit(r, B, C1, cats, out);
    release(r);
    if (!ok) { m->error = 'd'; return NAN; }
    return (long double)out[0];
}
extern "C" void apop_b200_logit_score(apop_data* d, gsl_vector* g, apop_model* m) {
    LOCK;
    if (!g) return;
    if (!m || !m->parameters || !d || !d->matrix) { fill_nan(g); return; }
    const size_t C1 = m->parameters->matrix ? m->parameters->matrix->size2 : 1;
    std::vector<double> cats;
    logit_categories(d, m, C1 + 1, cats);
    if (C1 == 1) {
        const double aux[4] = {cats[0], cats[1], 0, 0};
        GlmEval e = glm_entry(GLM_LOGIT2, d, m, aux);
        if (e.ok) store_score(g, e.out.data() + MAX_ACC, e.beta.size(), 1.0);
        else fill_nan(g);
        release(e.r);
        return;
    }
    std::vector<double> B, out;
    pack_parameters(m->parameters, B);
    Resident* r = acquire(d, LAYOUT_ANY);   // eval_logit converts to the layout its kernel reads
    if (!r) { fill_nan(g); return; }
    const bool ok = eval_logit(r, B, C1, cats, out);
    release(r);
    if (ok) store_score(g, out.data() + 1, B.size(), 1.0);
    else { m->error = 'd'; fill_nan(g); }
}

// sum_i lgamma(y_i + 1): beta-independent, one pass per residency
bool lgamma_const(Resident* r, double& v) {
    if (!r->have_lgamma_y) {
        double o[CELLS_NOUT];
        if (!eval_cells(r, CELLS_POISSON, 0.0, o, /*matrix=*/false)) return false;
        r->lgamma_y = o[3];
        r->have_lgamma_y = true;
    }
    v = r->lgamma_y;
    return true;
}
extern "C" long double apop_b200_poisson_reg_ll(apop_data* d, apop_model* m) {
    LOCK;
    GlmEval e = glm_entry(GLM_POISSON_REG, d, m, nullptr);
    long double ll = NAN;
    double lg;
    if (e.ok && lgamma_const(e.r, lg)) ll = (long double)e.out[0] - (long double)lg;
    release(e.r);
    return ll;
}
extern "C" void apop_b200_poisson_reg_score(apop_data* d, gsl_vector* g, apop_model* m) {
    LOCK;
    if (!g) return;
    GlmEval e = glm_entry(GLM_POISSON_REG, d, m, nullptr);
    if (e.ok) store_score(g, e.out.data() + MAX_ACC, e.beta.size(), 1.0);
    else fill_nan(g);
    release(e.r);
}

// ---- OLS -------------------------------------------------------------------
// ols_log_likelihood (model/apop_ols.c:129-172) has four branches besides the plain one; all are served:
//   un-prepped data (:149-153)   no outcome vector yet: column 0 of the matrix is the outcome and
//                                expected += beta_0 (1 - actual), i.e. exactly x.beta with column 0 := 1.  The
//                                upload shuffles the column out on the device (tile_pack's y_col0).
//   <Predicted> page (:139-145)  when p->info carries it and d == p->data, the errors are column 2 of that
//                                page (zero padded to the rows of d), whatever p->parameters say.
//   <Error variance> page (:157) sigma from the page instead of the sample variance of the errors.
//   input_distribution (:163-166) every row's density is multiplied by P(x_i): LL += sum_i log P(x_i).
// ols_score (:175-205) has only the un-prepped branch; there the multiplier of column 0 is the RAW
// column 0 (the outcome), as coded: g_0 = sum w y e / sigma^2.
struct OlsEval {
    bool ok = false;
    long double ll = NAN, sigma2 = NAN;
    std::vector<double> g;
};
static void* settings_group(const apop_model* m, const char* name) {
    if (!m || !m->settings) return nullptr;
    for (const apop_settings_type* s = m->settings; s->name[0]; s++)
        if (!strcmp(s->name, name)) return s->setting_group;
    return nullptr;
}
// the resident copy OLS works on: prepped data as usual; un-prepped data with the outcome shuffled out
static Resident* acquire_ols(apop_data* d, bool& unprepped) {
    unprepped = d && !d->vector && d->matrix;
    if (!unprepped) return acquire(d);
    if (!require_ready()) return nullptr;
    Runtime& R = rt();
    cudaSetDevice(R.device);
    auto it = R.reg.find(d);
    if (it != R.reg.end()) {
        Resident* r = acquire(d);                 // refreshes a dirty or stale entry
        if (!r) return nullptr;
        if (r->has_y) { unprepped = !r->synthetic && r->y_col0; return r; }   // pinned through the device prep: already shuffled
        // pinned raw (k columns, no outcome): leave that copy alone and work on a shuffled one
    }
    Resident* r = new Resident();
    r->y_col0 = true;
    if (!upload(r, d)) { free_resident(r); return nullptr; }
    if (R.auto_pin && it == R.reg.end()) R.reg[d] = r;
    else r->transient = true;
    return r;
}
// sum_i log P(x_i) for the input distribution of apop_lm_settings; 0 for the improper uniform default
static bool ols_input_term(apop_data* d, Resident* r, apop_model* id, long double& term) {
    term = 0;
    if (!id || (!id->log_likelihood && !id->p)) return true;
    if (!strncasecmp(id->name, "Improper uniform", 16)) return true;       // P(x) = 1
    if (id->log_likelihood && is_device_entry((void*)id->log_likelihood)) {
        // the device families are sums over cells: one call over the matrix part of the resident copy
        // (justarow->vector = NULL, :163), through a header-only view registered for the call
        Runtime& R = rt();
        apop_data view;
        memset(&view, 0, sizeof view);
        view.matrix = d->matrix;
        view.more = d->more;
        Resident tmp = *r;
        tmp.has_y = false; tmp.transient = false; tmp.synthetic = true;    // shares the tiles: never uploaded, never freed
        tmp.memo_valid = false; tmp.have_lgamma_y = tmp.have_poisson = tmp.have_gamma = false; tmp.consts.clear();
        tmp.n_global_epoch = ~0ull; tmp.counts = nullptr; tmp.lgy = nullptr;
        R.reg[&view] = &tmp;
        const long double v = id->log_likelihood(&view, id);
        R.reg.erase(&view);
        r->swz = tmp.swz;   // the view shares the tiles: a layout change made through it is the owner's too
        term = v;
        return !std::isnan((double)v);
    }
    // a host model: the reference's own loop, one apop_p per row view (needs the host buffers)
    if (!d->matrix || !d->matrix->data) return set_error("OLS input_distribution: a host model needs host data");
    const gsl_matrix* M = d->matrix;
    for (size_t i = 0; i < M->size1; i++) {
        gsl_matrix row = {1, M->size2, M->tda, M->data + i * M->tda, nullptr, 0};
        apop_data just;
        memset(&just, 0, sizeof just);
        just.matrix = &row;
        const long double lp = id->log_likelihood ? id->log_likelihood(&just, id) : logl(id->p(&just, id));
        term += lp;
    }
    return true;
}
static OlsEval ols_entry(apop_data* d, apop_model* m, bool for_score) {
    OlsEval o;
    if (!m || !m->parameters || !d || !d->matrix) { set_error("NULL model, parameters or data"); return o; }
    std::vector<double> beta, out;
    if (!pack_parameters(m->parameters, beta)) { set_error("empty parameters"); return o; }
    bool unprepped = false;
    Resident* r = acquire_ols(d, unprepped);
    if (!r) { d->error = 'a'; return o; }
    double n, mc, vc;
    // the un-prepped score needs sum w e^2 (for g_0 = sum w y e); it takes the place of sum log w
    const double aux[4] = {(for_score && unprepped) ? 1.0 : 0.0, 0, 0, 0};
    if (!eval_glm(r, GLM_OLS, beta.data(), beta.size(), aux, out) || !global_counts(r, n, mc, vc)) {
        m->error = 'd';
        release(r);
        return o;
    }
    long double S1 = out[0], S2 = out[1];
    const long double S3 = (r->has_w && !for_score) ? out[2] : 0.0L;
    // the score always takes the sample variance of the errors at these parameters (:192)
    const long double var_score = (S2 - S1 * S1 / n) / (n - 1);
    o.sigma2 = var_score;
    if (for_score) {
        o.g.assign(out.begin() + MAX_ACC, out.end());
        if (unprepped) {
            // column 0 of the caller's matrix is y itself: sum w y e = beta . (sum w x e) - sum w e^2
            long double bg = 0;
            for (size_t j = 0; j < beta.size(); j++) bg += (long double)beta[j] * out[MAX_ACC + j];
            const long double swe2 = r->has_w ? (long double)out[2] : S2;
            o.g[0] = (double)(bg - swe2);
        }
        for (auto& v : o.g) v = (double)(v / var_score);
        o.ok = true;
        release(r);
        return o;
    }
    // ---- LL: errors from the <Predicted> page when the model carries one for this very data set
    const apop_data* pred = m->info ? named_page_or_self(m->info, "<Predicted>") : nullptr;
    if (pred && d == m->data && pred->matrix && pred->matrix->size2 >= 3) {
        // column 2 of the page as a strided vector; rows beyond it count as zero errors (:141-144)
        gsl_vector ev = {pred->matrix->size1 < r->n_rows ? pred->matrix->size1 : r->n_rows, pred->matrix->tda,
                         pred->matrix->data + 2, nullptr, 0};
        apop_data view;
        memset(&view, 0, sizeof view);
        view.vector = &ev;
        Resident* rv = acquire(&view);
        double c5[CELLS_NOUT];
        const bool okv = rv && eval_cells(rv, CELLS_NORMAL, 0.0, c5, /*matrix=*/false);
        release(rv);
        if (!okv) { release(r); m->error = 'd'; return o; }
        S1 = c5[2]; S2 = c5[3];
        if (rt().nranks > 1) {   // the page holds this rank's rows only
            double v2[2] = {(double)S1, (double)S2};
            if (!allreduce_host(v2, 2)) { release(r); return o; }
            S1 = v2[0]; S2 = v2[1];
        }
    }
    long double var_ll = (S2 - S1 * S1 / n) / (n - 1);
    const apop_data* ev = named_page(m->parameters, "<Error variance>");
    if (ev) var_ll = ev->matrix ? ev->matrix->data[0] : (ev->vector ? ev->vector->data[0] : var_ll);
    const long double sigma = sqrtl(var_ll);
    // sum_i log( phi(e_i; sigma) w_i P(x_i) ) with phi the Gaussian density
    o.ll = -S2 / (2 * var_ll) - n * logl(sigma * 2.50662827463100050241576528481104525L) + S3;
    long double xterm = 0;
    struct lm_group { int destroy_data; apop_data* instruments; char want_cov, want_expected_value; apop_model* input_distribution; };
    const lm_group* lms = (const lm_group*)settings_group(m, "apop_lm");
    if (lms && lms->input_distribution && !ols_input_term(d, r, lms->input_distribution, xterm)) { release(r); m->error = 'd'; return o; }
    o.ll += xterm;
    o.ok = true;
    release(r);
    return o;
}
extern "C" long double apop_b200_ols_ll(apop_data* d, apop_model* m) {
    LOCK;
    OlsEval o = ols_entry(d, m, false);
    return o.ok ? o.ll : (long double)NAN;
}
extern "C" void apop_b200_ols_score(apop_data* d, gsl_vector* g, apop_model* m) {
    LOCK;
    if (!g) return;
    OlsEval o = ols_entry(d, m, true);
    if (o.ok) store_score(g, o.g.data(), o.g.size(), 1.0);
    else fill_nan(g);
}

// ---- iid Poisson / Normal ----------------------------------------------------
static bool poisson_consts(Resident* r, double* o) {
    if (!r->have_poisson) {
        if (!eval_cells(r, CELLS_POISSON, 0.0, r->pois)) return false;
        r->have_poisson = true;
    }
    memcpy(o, r->pois, sizeof r->pois);
    return true;
}
extern "C" long double apop_b200_poisson_ll(apop_data* d, apop_model* m) {
    LOCK;
    if (!m || !m->parameters || !d) { set_error("NULL model, parameters or data"); return NAN; }
    std::vector<double> p;
    if (!pack_parameters(m->parameters, p)) return NAN;
    Resident* r = acquire(d);
    if (!r) return NAN;
    double o[CELLS_NOUT], n, mc, vc;
    long double ll = NAN;
    if (poisson_consts(r, o) && global_counts(r, n, mc, vc)) {
        const long double lambda = p[0], ln_l = logl(lambda);
        if (o[4] > 0) ll = -INFINITY;
        else ll = ln_l * ((long double)o[0] + o[2]) - ((long double)o[1] + o[3]) - ((long double)mc + vc) * lambda;
    }
    release(r);
    return ll;
}
extern "C" void apop_b200_poisson_score(apop_data* d, gsl_vector* g, apop_model* m) {
    LOCK;
    if (!g) return;
    std::vector<double> p;
    if (!m || !m->parameters || !d || !pack_parameters(m->parameters, p)) { fill_nan(g); return; }
    Resident* r = acquire(d);
    if (!r) { fill_nan(g); return; }
    double o[CELLS_NOUT], n, mc, vc;
    if (poisson_consts(r, o) && global_counts(r, n, mc, vc))
        g->data[0] = (double)((long double)o[0] / p[0] - ((long double)mc + vc));  // matrix cells only in the sum
    else fill_nan(g);
    release(r);
}

static bool normal_sums(Resident* r, double mu, double* o) {
    // memo on mu: LL and score at the same parameters share one pass
    const double key[5] = {mu, 0, 0, 0, 0};
    if (r->memo_valid && r->memo_fam == 100 && r->memo_beta.size() == 5 && !memcmp(r->memo_beta.data(), key, sizeof key)) {
        memcpy(o, r->memo_out.data(), CELLS_NOUT * sizeof(double));
        return true;
    }
    if (!eval_cells(r, CELLS_NORMAL, mu, o)) return false;
    r->memo_valid = true; r->memo_fam = 100;
    r->memo_beta.assign(key, key + 5);
    r->memo_out.assign(o, o + CELLS_NOUT);
    return true;
}
extern "C" long double apop_b200_normal_ll(apop_data* d, apop_model* m) {
    LOCK;
    std::vector<double> p;
    if (!m || !m->parameters || !d || !pack_parameters(m->parameters, p) || p.size() < 2) { set_error("NULL model, parameters or data"); return NAN; }
    Resident* r = acquire(d);
    if (!r) return NAN;
    double o[CELLS_NOUT], n, mc, vc;
    long double ll = NAN;
    if (normal_sums(r, p[0], o) && global_counts(r, n, mc, vc)) {
        const long double sd = p[1], lnpi = 1.14472988584940017414342735135305871L, ln2 = 0.693147180559945309417232121458176568L;
        ll = -((long double)o[1] + o[3]) / (2 * sd * sd) - ((long double)mc + vc) * ((lnpi + ln2) / 2 + logl(sd));
    }
    release(r);
    return ll;
}
extern "C" void apop_b200_normal_score(apop_data* d, gsl_vector* g, apop_model* m) {
    LOCK;
    if (!g) return;
    std::vector<double> p;
    if (!m || !m->parameters || !d || !pack_parameters(m->parameters, p) || p.size() < 2) { fill_nan(g); return; }
    Resident* r = acquire(d);
    if (!r) { fill_nan(g); return; }
    double o[CELLS_NOUT], n, mc, vc;
    if (normal_sums(r, p[0], o) && global_counts(r, n, mc, vc) && g->size >= 2) {
        const long double sd = p[1];
        g->data[0] = (double)(((long double)o[0] + o[2]) / (sd * sd));
        g->data[g->stride] = (double)(((long double)o[1] + o[3]) / (sd * sd * sd) - ((long double)mc + vc) / sd);
    } else fill_nan(g);
    release(r);
}

// ---- exponential, gamma, lognormal ----------------------------------------
// sum x, sum_{x != 0} ln x and the number of zero cells do not depend on the parameters:
// one pass per residency, like the Poisson sums
static bool gamma_consts(Resident* r, double* o) {
    if (!r->have_gamma) {
        if (!eval_cells(r, CELLS_GAMMA, 0.0, r->gam)) return false;
        r->have_gamma = true;
    }
    memcpy(o, r->gam, sizeof r->gam);
    return true;
}
static long double digamma_l(long double x) {   // gsl_sf_psi: recurrence to x >= 12, then the asymptotic series
    long double r = 0;
    while (x < 12) { r -= 1 / x; x += 1; }
    const long double f = 1 / (x * x);
    return r + logl(x) - 0.5L / x - f * (1.0L / 12 - f * (1.0L / 120 - f * (1.0L / 252 - f * (1.0L / 240 - f * (1.0L / 132 - f * (691.0L / 32760 - f / 12))))));
}
struct IidEval { bool ok = false; double o[CELLS_NOUT]; double cells = 0; std::vector<double> p; };
static IidEval iid_entry(apop_data* d, apop_model* m, size_t nparams, int mode, double p0_index) {
    IidEval e;
    if (!m || !m->parameters || !d || !pack_parameters(m->parameters, e.p) || e.p.size() < nparams) { set_error("NULL model, parameters or data"); return e; }
    Resident* r = acquire(d);
    if (!r) return e;
    double n, mc, vc;
    bool ok = global_counts(r, n, mc, vc);
    if (ok && mode == CELLS_GAMMA) ok = gamma_consts(r, e.o);
    else if (ok) {   // lognormal: memo on mu, like the Normal
        const double mu = e.p[(size_t)p0_index];
        const double key[5] = {mu, 1, 0, 0, 0};
        if (r->memo_valid && r->memo_fam == 101 && r->memo_beta.size() == 5 && !memcmp(r->memo_beta.data(), key, sizeof key))
            memcpy(e.o, r->memo_out.data(), sizeof e.o);
        else if ((ok = eval_cells(r, mode, mu, e.o))) {
            r->memo_valid = true; r->memo_fam = 101;
            r->memo_beta.assign(key, key + 5);
            r->memo_out.assign(e.o, e.o + CELLS_NOUT);
        }
    }
    e.cells = mc + vc;
    e.ok = ok;
    release(r);
    return e;
}
extern "C" long double apop_b200_exponential_ll(apop_data* d, apop_model* m) {
    LOCK;
    IidEval e = iid_entry(d, m, 1, CELLS_GAMMA, 0);
    if (!e.ok) return NAN;
    const long double mu = e.p[0];
    return -((long double)e.o[0] + e.o[2]) / mu - (long double)e.cells * logl(mu);
}
extern "C" void apop_b200_exponential_score(apop_data* d, gsl_vector* g, apop_model* m) {
    LOCK;
    if (!g) return;
    IidEval e = iid_entry(d, m, 1, CELLS_GAMMA, 0);
    if (!e.ok) { fill_nan(g); return; }
    const long double mu = e.p[0];
    g->data[0] = (double)(((long double)e.o[0] + e.o[2]) / (mu * mu) - (long double)e.cells / mu);
}
extern "C" long double apop_b200_gamma_ll(apop_data* d, apop_model* m) {
    LOCK;
    IidEval e = iid_entry(d, m, 2, CELLS_GAMMA, 0);
    if (!e.ok) return NAN;
    const long double a = e.p[0], b = e.p[1];
    const long double nz = (long double)e.cells - e.o[4];   // cells with x != 0
    return (a - 1) * ((long double)e.o[1] + e.o[3]) - ((long double)e.o[0] + e.o[2]) / b - nz * (lgammal(a) + a * logl(b));
}
extern "C" void apop_b200_gamma_score(apop_data* d, gsl_vector* g, apop_model* m) {
    LOCK;
    if (!g) return;
    IidEval e = iid_entry(d, m, 2, CELLS_GAMMA, 0);
    if (!e.ok || g->size < 2) { fill_nan(g); return; }
    const long double a = e.p[0], b = e.p[1];
    // a_callback has no x != 0 guard: a zero cell sends dLL/da to -inf (log 0)
    g->data[0] = (e.o[4] > 0) ? -INFINITY
                              : (double)(((long double)e.o[1] + e.o[3]) - (long double)e.cells * (digamma_l(a) + logl(b)));
    g->data[g->stride] = (double)(((long double)e.o[0] + e.o[2]) / (b * b) - (long double)e.cells * a / b);
}
extern "C" long double apop_b200_lognormal_ll(apop_data* d, apop_model* m) {
    LOCK;
    IidEval e = iid_entry(d, m, 2, CELLS_LOGNORMAL, 0);
    if (!e.ok) return NAN;
    const long double sd = e.p[1], lnpi = 1.14472988584940017414342735135305871L, ln2 = 0.693147180559945309417232121458176568L;
    return -((long double)e.o[1] + e.o[3]) / (2 * sd * sd) - ((long double)e.o[0] + e.o[2]) - (long double)e.cells * ((lnpi + ln2) / 2 + logl(sd));
}
extern "C" void apop_b200_lognormal_score(apop_data* d, gsl_vector* g, apop_model* m) {
    LOCK;
    if (!g) return;
    IidEval e = iid_entry(d, m, 2, CELLS_LOGNORMAL, 0);
    if (!e.ok || g->size < 2) { fill_nan(g); return; }
    const long double mu = e.p[0], sd = e.p[1];
    g->data[0] = (double)((((long double)e.o[0] + e.o[2]) - mu * (long double)e.cells) / (sd * sd));
    g->data[g->stride] = (double)(((long double)e.o[1] + e.o[3]) / (sd * sd * sd) - (long double)e.cells / sd);
}

extern "C" double apop_b200_map_sum(apop_data* d, apop_b200_map_fn fn, double p0, char part) {
    LOCK;
    if (!d) { set_error("NULL data"); return 0; }  // apop_map_sum returns 0 on NULL input
    Resident* r = acquire(d);
    if (!r) return NAN;
    int mode, slot = 0;
    switch (fn) {
    case APOP_B200_MAP_IDENTITY: mode = CELLS_IDENTITY; break;
    case APOP_B200_MAP_DEV: mode = CELLS_NORMAL; break;
    case APOP_B200_MAP_DEV2: mode = CELLS_NORMAL; slot = 1; break;
    case APOP_B200_MAP_POISSON: mode = CELLS_POISSON; break;
    case APOP_B200_MAP_LOG: mode = CELLS_LOG; break;
    case APOP_B200_MAP_EXP: mode = CELLS_EXP; break;
    default: release(r); set_error("unknown map function"); return NAN;
    }
    double o[CELLS_NOUT];
    double res = NAN;
    if (eval_cells(r, mode, p0, o)) {
        const bool um = (part != 'v'), uv = (part != 'm');
        if (fn == APOP_B200_MAP_POISSON) {
            // sum over cells of  x==0 ? 0 : p0*x - lgamma(x+1), -inf if any cell is invalid
            res = (o[4] > 0) ? -INFINITY
                             : (double)((long double)p0 * ((um ? o[0] : 0.0L) + (uv ? o[2] : 0.0L)) -
                                        ((um ? (long double)o[1] : 0.0L) + (uv ? (long double)o[3] : 0.0L)));
        } else
            res = (double)((um ? (long double)o[slot] : 0.0L) + (uv ? (long double)o[2 + slot] : 0.0L));
    }
    release(r);
    return res;
}

// ---- apop_dot(d, parameters): the N x c index X.B, copied back in chunks ----------------------
extern "C" int apop_b200_dot(apop_data* d, const double* B, size_t k, size_t c, double* out) {
    LOCK;
    if (!d || !B || !out || !c) { set_error("apop_b200_dot: NULL argument"); return 1; }
    Resident* r = acquire(d);
    if (!r) return 1;
    Runtime& R = rt();
    bool ok = false;
    double *dB = nullptr, *dOut = nullptr;
    do {
        if ((size_t)r->k != k || r->k < 1 || r->k > MAX_K_REG) {
            // the reference's dimension error: error char 'd' (apop_linear_algebra.m4.c:347-349)
            set_error("apop_b200_dot: the data has %d columns, the parameters %zu rows", r->k, k);
            d->error = 'd';
            break;
        }
        size_t chunk = ((size_t)16 << 20) / c;   // <= 128 MB of output per chunk
        if (chunk > r->n_rows) chunk = r->n_rows;
        if (!chunk) { ok = true; break; }
        if (!cuda_ok(cudaMalloc(&dB, k * c * sizeof(double)), "cudaMalloc") || !cuda_ok(cudaMalloc(&dOut, chunk * c * sizeof(double)), "cudaMalloc")) break;
        if (!cuda_ok(cudaMemcpyAsync(dB, B, k * c * sizeof(double), cudaMemcpyHostToDevice, R.stream), "H2D B")) break;
        ok = true;
        for (size_t r0 = 0; ok && r0 < r->n_rows; r0 += chunk) {
            const size_t rows = (r->n_rows - r0 < chunk) ? r->n_rows - r0 : chunk;
            ok = cuda_ok(xbeta_launch(r->tiles, r0, rows, r->k, r->cols, (int)c, dB, dOut, R.stream), "xbeta launch") &&
                 cuda_ok(cudaMemcpyAsync(out + r0 * c, dOut, rows * c * sizeof(double), cudaMemcpyDeviceToHost, R.stream), "D2H") &&
                 cuda_ok(cudaStreamSynchronize(R.stream), "sync");
            R.launches++;
        }
    } while (0);
    if (dB) cudaFree(dB);
    if (dOut) cudaFree(dOut);
    release(r);
    return ok ? 0 : 1;
}

// ---- prep reductions on the device --------------------------------------
extern "C" int apop_b200_col_sums(apop_data* d, int mode, double* out) {
    LOCK;
    if (!d || !out || mode < 0 || mode > 3) { set_error("apop_b200_col_sums: bad argument"); return 1; }
    Resident* r = acquire(d);
    if (!r) return 1;
    Runtime& R = rt();
    bool ok = false;
    if (r->k < 1 || r->k > MAX_K_REG) set_error("apop_b200_col_sums: 1..%d columns", MAX_K_REG);
    else {
        const int grid = grid_for(4, r->n_tiles, CELLS_THREADS / 32);
        if (ensure_partials((size_t)grid * MAX_K_REG)) {
            const FinalizeCtx fin = make_fin(MAX_K_REG);
            time_begin();
            ok = cuda_ok(colsums_launch(r->tiles, r->n_rows, r->n_tiles, r->k, r->cols, mode, R.d_partials, R.d_ticket,
                                        R.d_out, grid, R.stream, fin), "column sums launch");
            time_end_record();
            R.launches++;
            ok = ok && reduce_and_fetch(MAX_K_REG, fin);
            if (ok) { time_collect(); memcpy(out, R.h_out, r->k * sizeof(double)); }
        }
    }
    release(r);
    return ok ? 0 : 1;
}
// probit_prep's start point (model/apop_probit.c:65-80): beta0_j = exp(mean_i ln|x_ij|), zeros skipped
extern "C" int apop_b200_probit_start(apop_data* d, double* beta0) {
    LOCK;
    if (!d || !beta0) { set_error("apop_b200_probit_start: NULL argument"); return 1; }
    Resident* r = acquire(d);
    if (!r) return 1;
    double n, mc, vc;
    const int k = r->k;
    const bool okc = global_counts(r, n, mc, vc);
    const bool keep = !r->transient;
    release(r);
    if (!okc) return 1;
    if (apop_b200_col_sums(d, 0, beta0)) return 1;
    (void)keep;
    for (int j = 0; j < k; j++) beta0[j] = (double)expl((long double)beta0[j] / n);
    return 0;
}

// ---- generic fused value + gradient ------------------------------------------
extern "C" double apop_b200_ll_score(apop_b200_family family, apop_data* d, const double* beta, size_t P, double* score) {
    LOCK;
    if (!d || !beta) { set_error("NULL data or beta"); return NAN; }
    // a throw-away model header around beta: reuse the reference-signature entries
    apop_model m;
    memset(&m, 0, sizeof m);
    apop_data prm;
    memset(&prm, 0, sizeof prm);
    gsl_vector v = {P, 1, const_cast<double*>(beta), nullptr, 0};
    gsl_matrix mat = {0, 0, 0, const_cast<double*>(beta), nullptr, 0};
    gsl_vector g = {P, 1, score, nullptr, 0};
    m.parameters = &prm;
    m.data = d;
    prm.vector = &v;
    long double ll = NAN;
    switch (family) {
    case APOP_B200_PROBIT: ll = apop_b200_probit_ll(d, &m); if (score) apop_b200_probit_score(d, &g, &m); break;
    case APOP_B200_LOGIT: {
        const size_t k = d->matrix ? d->matrix->size2 : 0;
        if (!k || P % k) { set_error("logit: P = %zu is not a multiple of k = %zu", P, k); return NAN; }
        prm.vector = nullptr; prm.matrix = &mat;
        mat.size1 = k; mat.size2 = P / k; mat.tda = P / k;
        ll = apop_b200_logit_ll(d, &m); if (score) apop_b200_logit_score(d, &g, &m); break;
    }
    case APOP_B200_POISSON: ll = apop_b200_poisson_ll(d, &m); if (score) apop_b200_poisson_score(d, &g, &m); break;
    case APOP_B200_NORMAL: ll = apop_b200_normal_ll(d, &m); if (score) apop_b200_normal_score(d, &g, &m); break;
    case APOP_B200_OLS: ll = apop_b200_ols_ll(d, &m); if (score) apop_b200_ols_score(d, &g, &m); break;
    case APOP_B200_POISSON_REG: ll = apop_b200_poisson_reg_ll(d, &m); if (score) apop_b200_poisson_reg_score(d, &g, &m); break;
    case APOP_B200_EXPONENTIAL: ll = apop_b200_exponential_ll(d, &m); if (score) apop_b200_exponential_score(d, &g, &m); break;
    case APOP_B200_GAMMA: ll = apop_b200_gamma_ll(d, &m); if (score) apop_b200_gamma_score(d, &g, &m); break;
    case APOP_B200_LOGNORMAL: ll = apop_b200_lognormal_ll(d, &m); if (score) apop_b200_lognormal_score(d, &g, &m); break;
    case APOP_B200_BERNOULLI: ll = apop_b200_bernoulli_ll(d, &m); if (score) apop_b200_bernoulli_score(d, &g, &m); break;
    case APOP_B200_BETA: ll = apop_b200_beta_ll(d, &m); if (score) apop_b200_beta_score(d, &g, &m); break;
    case APOP_B200_DIRICHLET: ll = apop_b200_dirichlet_ll(d, &m); if (score) apop_b200_dirichlet_score(d, &g, &m); break;
    case APOP_B200_YULE: ll = apop_b200_yule_ll(d, &m); if (score) apop_b200_yule_score(d, &g, &m); break;
    case APOP_B200_ZIPF:
    case APOP_B200_T: {
        // no analytic score in the reference either: apop_score falls back to apop_numerical_gradient
        // (apop_mle.m4.c:92-143, gsl_deriv_central's 5-point rule at delta = 1e-3) over the device LL
        long double (*f)(apop_data*, apop_model*) = family == APOP_B200_ZIPF ? apop_b200_zipf_ll : apop_b200_t_ll;
        ll = f(d, &m);
        if (score) {
            std::vector<double> b(beta, beta + P);
            v.data = b.data();
            const double h = 1e-3, at[4] = {-h, -h / 2, h / 2, h};
            for (size_t j = 0; j < P; j++) {
                double fv[4];
                for (int q = 0; q < 4; q++) { b[j] = beta[j] + at[q]; fv[q] = (double)f(d, &m); }
                b[j] = beta[j];
                const double r3 = 0.5 * (fv[3] - fv[0]), r5 = (4.0 / 3.0) * (fv[2] - fv[1]) - (1.0 / 3.0) * r3;
                score[j] = r5 / h;
            }
        }
        break;
    }
    default: set_error("unknown family %d", (int)family);
    }
    return (double)ll;
}

