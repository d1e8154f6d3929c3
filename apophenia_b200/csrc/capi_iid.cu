// capi_iid.cu -- the remaining iid distributions of SURVEY 8(f) rank 3 behind the reference
// signatures: Bernoulli, Beta, Zipf, Dirichlet, Student t, Yule.  Like the Normal / Poisson /
// Gamma entries they are apop_map_sum shapes (apop_mapply.m4.c:485-517): one pass of the
// cells kernel (cells.cu) over the resident data, and a closed form in long double on the
// host.  Where the per-cell terms do not depend on the parameters (Bernoulli, Beta, Zipf,
// Dirichlet) the sums are reduced once per residency and cached, so an optimizer iteration
// costs no kernel at all; t and Yule stream the data once per evaluation.
#include "capi_internal.cuh"

using namespace apb;

typedef long double ld;

struct IidCall {
    bool ok = false;
    Resident* r = nullptr;
    std::vector<double> p;
    double n = 0, mc = 0, vc = 0;   // global rows, matrix cells, vector cells
    ld cells() const { return (ld)mc + vc; }
};
static IidCall iid_begin(apop_data* d, apop_model* m, size_t min_params) {
    IidCall c;
    if (!m || !m->parameters || !d || !pack_parameters(m->parameters, c.p) || c.p.size() < min_params) {
        set_error("NULL model, parameters or data");   // Nullcheck_mpd: the entries return GSL_NAN
        return c;
    }
    c.r = acquire(d);
    if (!c.r) return c;
    c.ok = global_counts(c.r, c.n, c.mc, c.vc);
    if (!c.ok) { release(c.r); c.r = nullptr; }
    return c;
}
// parameter-independent cell sums: one pass per residency (apop_b200_invalidate drops them)
static bool cached_cells(Resident* r, int mode, double* o) {
    auto it = r->consts.find(mode);
    if (it == r->consts.end()) {
        double t[CELLS_NOUT];
        if (!eval_cells(r, mode, 0.0, t)) return false;
        it = r->consts.emplace(mode, std::vector<double>(t, t + CELLS_NOUT)).first;
    }
    memcpy(o, it->second.data(), CELLS_NOUT * sizeof(double));
    return true;
}
static bool cached_colsums(Resident* r, int mode, std::vector<double>& out) {
    const int key = 1000 + mode;
    auto it = r->consts.find(key);
    if (it == r->consts.end()) {
        Runtime& R = rt();
        if (r->k < 1 || r->k > MAX_K_REG) return set_error("column sums cover 1..%d columns, the data has %d", MAX_K_REG, r->k);
        const int grid = grid_for(4, r->n_tiles, CELLS_THREADS / 32);
        if (!ensure_partials((size_t)grid * MAX_K_REG)) return false;
        const FinalizeCtx fin = make_fin(MAX_K_REG);
        time_begin();
        if (!cuda_ok(colsums_launch(r->tiles, r->n_rows, r->n_tiles, r->k, r->cols, mode, R.d_partials, R.d_ticket, R.d_out, grid,
                                    R.stream, fin), "column sums launch")) return false;
        time_end_record();
        R.launches++;
        if (!reduce_and_fetch(MAX_K_REG, fin)) return false;
        time_collect();
        it = r->consts.emplace(key, std::vector<double>(R.h_out, R.h_out + r->k)).first;
    }
    out = it->second;
    return true;
}
static ld digamma_ld(ld x) {   // gsl_sf_psi stand-in: recurrence to x >= 12, then the asymptotic series
    ld r = 0;
    while (x < 12) { r -= 1 / x; x += 1; }
    const ld f = 1 / (x * x);
    return r + logl(x) - 0.5L / x - f * (1.0L / 12 - f * (1.0L / 120 - f * (1.0L / 252 - f * (1.0L / 240 - f * (1.0L / 132 - f * (691.0L / 32760 - f / 12))))));
}
// Riemann zeta(s), s > 1 (gsl_sf_zeta stand-in): N leading terms + Euler-Maclaurin tail
static ld zeta_ld(ld s) {
    const int N = 24;
    ld sum = 0;
    for (int n = 1; n < N; n++) sum += powl((ld)n, -s);
    const ld a = N, as = powl(a, -s);
    sum += as * a / (s - 1) + as / 2;
    static const ld B2k[] = {1.0L / 6, -1.0L / 30, 1.0L / 42, -1.0L / 30, 5.0L / 66, -691.0L / 2730, 7.0L / 6, -3617.0L / 510, 43867.0L / 798};
    ld term = s * as / a, fact = 2;   // s a^{-s-1} / 2!
    for (int k = 1; k <= 9; k++) {
        sum += B2k[k - 1] * term / fact;
        term *= (s + 2 * k - 1) * (s + 2 * k) / (a * a);
        fact *= (2 * k + 1) * (2 * k + 2);
    }
    return sum;
}

// ---- Bernoulli: bernoulli_log_likelihood / bernie_ll (model/apop_bernoulli.c:13-23) ----------
// LL = n1 ln p + n0 ln(1-p), n1 the number of non-zero cells.  The reference registers no score
// (numeric gradient); the closed form d/dp = n1/p - n0/(1-p) is a new vtable entry.
extern "C" long double apop_b200_bernoulli_ll(apop_data* d, apop_model* m) {
    LOCK;
    IidCall c = iid_begin(d, m, 1);
    if (!c.ok) return NAN;
    double o[CELLS_NOUT];
    ld ll = NAN;
    if (cached_cells(c.r, CELLS_NONZERO, o)) {
        const ld p = c.p[0], n1 = (ld)o[0] + o[2], n0 = c.cells() - n1;
        // x ? log(p) : log(1-p) per cell: a class with no cells contributes nothing
        ll = (n1 > 0 ? n1 * (ld)log((double)p) : 0) + (n0 > 0 ? n0 * (ld)log((double)(1 - p)) : 0);
    }
    release(c.r);
    return ll;
}
extern "C" void apop_b200_bernoulli_score(apop_data* d, gsl_vector* g, apop_model* m) {
    LOCK;
    if (!g) return;
    IidCall c = iid_begin(d, m, 1);
    if (!c.ok) { fill_nan(g); return; }
    double o[CELLS_NOUT];
    if (cached_cells(c.r, CELLS_NONZERO, o)) {
        const ld p = c.p[0], n1 = (ld)o[0] + o[2], n0 = c.cells() - n1;
        g->data[0] = (double)((n1 > 0 ? n1 / p : 0) - (n0 > 0 ? n0 / (1 - p) : 0));
    } else fill_nan(g);
    release(c.r);
}

// ---- Beta: beta_log_likelihood / betamap, beta_dlog_likelihood (model/apop_beta.c:47-75) ------
// LL = (a-1) S_lnx + (b-1) S_ln(1-x) - tsize lnB(a,b), the sums over cells inside [0,1] only;
// the score sums ln x and ln(1-x) over EVERY cell (no range test there).
extern "C" long double apop_b200_beta_ll(apop_data* d, apop_model* m) {
    LOCK;
    IidCall c = iid_begin(d, m, 2);
    if (!c.ok) return NAN;
    double o[CELLS_NOUT];
    ld ll = NAN;
    const ld a = c.p[0], b = c.p[1];
    if (isnan((double)(a + b))) set_error("NaN alpha or beta input");
    else if (cached_cells(c.r, CELLS_BETA, o))
        ll = (a - 1) * ((ld)o[0] + o[2]) + (b - 1) * ((ld)o[1] + o[3]) - (lgammal(a) + lgammal(b) - lgammal(a + b)) * c.cells();
    release(c.r);
    return ll;
}
extern "C" void apop_b200_beta_score(apop_data* d, gsl_vector* g, apop_model* m) {
    LOCK;
    if (!g) return;
    IidCall c = iid_begin(d, m, 2);
    if (!c.ok || g->size < 2) { fill_nan(g); if (c.r) release(c.r); return; }
    double o[CELLS_NOUT];
    bool ok = cached_cells(c.r, CELLS_BETA, o);
    if (ok && o[4] > 0) ok = cached_cells(c.r, CELLS_BETA_RAW, o);   // cells outside [0,1]: the unguarded sums
    if (ok) {
        const ld a = c.p[0], b = c.p[1], t = c.cells();
        g->data[0] = (double)(((ld)o[0] + o[2]) + (digamma_ld(a + b) - digamma_ld(a)) * t);
        g->data[g->stride] = (double)(((ld)o[1] + o[3]) + (digamma_ld(a + b) - digamma_ld(b)) * t);
    } else fill_nan(g);
    release(c.r);
}

// ---- Zipf: zipf_log_likelihood (model/apop_zipf.c:31-40): LL = -bb S_lnx - tsize ln zeta(bb) ----
extern "C" long double apop_b200_zipf_ll(apop_data* d, apop_model* m) {
    LOCK;
    IidCall c = iid_begin(d, m, 1);
    if (!c.ok) return NAN;
    double o[CELLS_NOUT];
    ld ll = NAN;
    const ld bb = c.p[0];
    if (isnan((double)bb) || bb < 1) set_error("Zipf needs a parameter >= 1; got %Lg", bb);
    else if (cached_cells(c.r, CELLS_LOG, o)) {
        double like = (double)(-((ld)o[0] + o[2]) * bb);          // `double like` in the reference
        like -= log((double)zeta_ld(bb)) * (double)c.cells();
        ll = like;
    }
    release(c.r);
    return ll;
}

// ---- Dirichlet: dirichlet_log_likelihood / dirichlet_dlog_likelihood (model/apop_dirichlet.c:13-38)
// per row gsl_ran_dirichlet_lnpdf = sum_j (alpha_j - 1) ln x_ij + lnGamma(sum alpha) - sum lnGamma(alpha_j):
// the data enter through the column sums of ln x only.
extern "C" long double apop_b200_dirichlet_ll(apop_data* d, apop_model* m) {
    LOCK;
    if (m && m->parameters && !m->parameters->vector) { set_error("parameters should be in inmodel->parameters->vector"); return NAN; }
    IidCall c = iid_begin(d, m, 1);
    if (!c.ok) return NAN;
    ld ll = NAN;
    std::vector<double> S;
    ld asum = 0;
    for (double a : c.p) asum += a;
    if (fabsl(asum) < 1e-5) set_error("Parameter total is too close to zero");
    else if (isnan((double)asum)) set_error("NaN parameter");
    else if ((size_t)c.r->k != c.p.size()) { set_error("dirichlet: %zu parameters for %d columns", c.p.size(), c.r->k); m->error = 'd'; }
    else if (cached_colsums(c.r, 3, S)) {
        ld s = 0, lg = 0;
        for (size_t j = 0; j < c.p.size(); j++) { s += ((ld)c.p[j] - 1) * S[j]; lg += lgammal(c.p[j]); }
        ll = s + (ld)c.n * (lgammal(asum) - lg);
    }
    release(c.r);
    return ll;
}
extern "C" void apop_b200_dirichlet_score(apop_data* d, gsl_vector* g, apop_model* m) {
    LOCK;
    if (!g) return;
    IidCall c = iid_begin(d, m, 1);
    if (!c.ok) { fill_nan(g); return; }
    std::vector<double> S;
    if ((size_t)c.r->k == c.p.size() && g->size >= c.p.size() && cached_colsums(c.r, 3, S)) {
        ld asum = 0;
        for (double a : c.p) asum += a;
        for (size_t j = 0; j < c.p.size(); j++)
            g->data[j * g->stride] = (double)((ld)S[j] + (ld)c.n * digamma_ld(asum) - (ld)c.n * digamma_ld(c.p[j]));
    } else fill_nan(g);
    release(c.r);
}

// ---- Student t: apop_tdist_llike / one_t (model/apop_t.c:25-38), parameters (mu, sigma, df) ----
// ln gsl_ran_tdist_pdf(z, nu) = lnGamma((nu+1)/2) - lnGamma(nu/2) - ln(pi nu)/2 - (nu+1)/2 ln(1 + z^2/nu);
// the reference takes log(pdf), which is -inf once the pdf underflows (|z| beyond ~1e150/nu):
// this entry returns the finite value there (documented deviation, like the OLS moments form).
extern "C" long double apop_b200_t_ll(apop_data* d, apop_model* m) {
    LOCK;
    IidCall c = iid_begin(d, m, 3);
    if (!c.ok) return NAN;
    double o[CELLS_NOUT];
    ld ll = NAN;
    const double mu = c.p[0], sigma = c.p[1], nu = c.p[2];
    const double key[5] = {mu, sigma, nu, 7, 0};
    Resident* r = c.r;
    bool ok = true;
    if (r->memo_valid && r->memo_fam == 102 && r->memo_beta.size() == 5 && !memcmp(r->memo_beta.data(), key, sizeof key))
        memcpy(o, r->memo_out.data(), sizeof o);
    else if ((ok = eval_cells(r, CELLS_T, mu, o, true, sigma, nu))) {
        r->memo_valid = true; r->memo_fam = 102;
        r->memo_beta.assign(key, key + 5);
        r->memo_out.assign(o, o + CELLS_NOUT);
    }
    if (ok) {
        const ld lnpi = 1.14472988584940017414342735135305871L;
        const ld cst = lgammal(((ld)nu + 1) / 2) - lgammal((ld)nu / 2) - (lnpi + logl(nu)) / 2;
        ll = c.cells() * cst - ((ld)nu + 1) / 2 * ((ld)o[0] + o[2]) - c.cells() * (ld)log(sigma);
    }
    release(c.r);
    return ll;
}

// ---- Yule: yule_log_likelihood / yule_dlog_likelihood (model/apop_yule.c:30-60) -----------------
static bool yule_sums(Resident* r, double bb, double* o) {
    const double key[5] = {bb, 8, 0, 0, 0};
    if (r->memo_valid && r->memo_fam == 103 && r->memo_beta.size() == 5 && !memcmp(r->memo_beta.data(), key, sizeof key)) {
        memcpy(o, r->memo_out.data(), CELLS_NOUT * sizeof(double));
        return true;
    }
    if (!eval_cells(r, CELLS_YULE, bb, o)) return false;
    r->memo_valid = true; r->memo_fam = 103;
    r->memo_beta.assign(key, key + 5);
    r->memo_out.assign(o, o + CELLS_NOUT);
    return true;
}
extern "C" long double apop_b200_yule_ll(apop_data* d, apop_model* m) {
    LOCK;
    IidCall c = iid_begin(d, m, 1);
    if (!c.ok) return NAN;
    double o[CELLS_NOUT];
    ld ll = NAN;
    const double bb = c.p[0];
    if (yule_sums(c.r, bb, o)) {
        const ld ln_bb = lgammal(bb), ln_bb_less_1 = log(bb - 1);
        const double likelihood = (double)((ld)o[0] + o[2]);
        ll = likelihood + (ln_bb_less_1 + ln_bb) * c.cells();
    }
    release(c.r);
    return ll;
}
extern "C" void apop_b200_yule_score(apop_data* d, gsl_vector* g, apop_model* m) {
    LOCK;
    if (!g) return;
    IidCall c = iid_begin(d, m, 1);
    if (!c.ok) { fill_nan(g); return; }
    double o[CELLS_NOUT];
    const double bb = c.p[0];
    if (yule_sums(c.r, bb, o)) {
        const ld bb_minus_one_inv = 1 / ((ld)bb - 1), psi_bb = digamma_ld(bb);
        double d_bb = (double)((ld)o[1] + o[3]);
        d_bb += (double)((bb_minus_one_inv + psi_bb) * c.cells());
        g->data[0] = d_bb;
    } else fill_nan(g);
    release(c.r);
}
