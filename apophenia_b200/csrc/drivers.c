/* drivers.c -- the callers either side of the hot path that loop over it thousands of
 * times (SURVEY 8(a) rows a15/a17/a18 estimates, a19 covariance, a21 bootstrap, a22 MCMC),
 * mirrored on the host over the device entries.
 *
 *   apop_bootstrap_cov        apop_bootstrap.m4.c:122-182.  The reference copies N resampled
 *                             rows into a scratch set and runs a full apop_estimate per
 *                             replicate.  Here the resamples are multiplicity counts resident
 *                             next to the data and all replicates are fitted in lock-step, one
 *                             batched device pass per optimizer step (DMMA path).  The multinomial
 *                             logit takes the reference's own loop -- one fit per replicate -- over the
 *                             same resident data, each row weighted by its resample count
 *                             (apop_b200_set_replicate); anything else resamples on the host
 *                             (apop_b200_invalidate, apop_estimate).
 *   apop_model_metropolis     apop_mcmc.m4.c:281-340: one chain, one LL pass per step, MVN
 *                             proposal centred on the last accepted draw, variance scaled by
 *                             sigma_adapt (:18-32) on every acceptance.
 *   apop_b200_metropolis_chains  the same sampler for many chains at once (new; F6).
 *   closed-form estimates     poisson_estimate (model/apop_poisson.c:42-47), normal_estimate
 *                             (model/apop_normal.c:78-97), apop_estimate_OLS (model/apop_ols.c:307-353)
 *   apop_model_numerical_covariance  apop_mle.m4.c:239-265: -H^-1, with H from one
 *                             information pass instead of >= 4 P score passes.
 * RNG streams are not GSL's (taus2 / gsl_ran_gaussian): "parity unpinned" (SURVEY 8(c)).
 */
#include "hostmini.h"
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- the device library, found in the running process --------------------------- */
typedef int (*batched_fn)(int, apop_data *, const double *, size_t, size_t, int, double *, double *);
typedef int (*counts_fn)(apop_data *, size_t, uint64_t);
typedef int (*info_fn)(int, apop_data *, const double *, size_t, double *);
typedef int (*gram_fn)(apop_data *, double *, double *);
typedef double (*mapsum_fn)(apop_data *, int, double, char);
typedef int (*data_fn)(apop_data *);

static void *dev(const char *name) {
    void *p = dlsym(RTLD_DEFAULT, name);
    if (!p) fprintf(stderr, "%s not found: load libapop_b200.so (RTLD_GLOBAL) first; there is no CPU fallback\n", name);
    return p;
}
int apop_b200_family_of(apop_model *m) {
    static const char *names[] = {"apop_b200_probit_ll", "apop_b200_logit_ll", "apop_b200_poisson_ll", "apop_b200_normal_ll",
                                  "apop_b200_ols_ll", "apop_b200_poisson_reg_ll", "apop_b200_exponential_ll",
                                  "apop_b200_gamma_ll", "apop_b200_lognormal_ll", "apop_b200_bernoulli_ll", "apop_b200_beta_ll",
                                  "apop_b200_zipf_ll", "apop_b200_dirichlet_ll", "apop_b200_t_ll", "apop_b200_yule_ll"};
    if (!m || !m->log_likelihood) return -1;
    for (int f = 0; f < 15; f++) { void *p = dlsym(RTLD_DEFAULT, names[f]); if (p && p == (void *)m->log_likelihood) return f; }
    return -1;
}

/* ---- small host RNG (xoshiro256**, Box-Muller) ------------------------------------ */
typedef struct { uint64_t s[4]; int have; double spare; } hrng;
static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t hrng_next(hrng *r) {
    const uint64_t res = rotl(r->s[1] * 5, 7) * 9, t = r->s[1] << 17;
    r->s[2] ^= r->s[0]; r->s[3] ^= r->s[1]; r->s[1] ^= r->s[2]; r->s[0] ^= r->s[3]; r->s[2] ^= t; r->s[3] = rotl(r->s[3], 45);
    return res;
}
static void hrng_seed(hrng *r, uint64_t seed) {
    for (int i = 0; i < 4; i++) { seed += 0x9E3779B97F4A7C15ull; uint64_t z = seed; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; r->s[i] = z ^ (z >> 31); }
    r->have = 0;
}
static double hrng_uniform(hrng *r) { return ((hrng_next(r) >> 11) + 0.5) * 0x1.0p-53; }
static double hrng_normal(hrng *r) {
    if (r->have) { r->have = 0; return r->spare; }
    const double u = hrng_uniform(r), v = hrng_uniform(r), rad = sqrt(-2 * log(u));
    r->spare = rad * sin(2 * M_PI * v); r->have = 1;
    return rad * cos(2 * M_PI * v);
}

/* ---- dense helpers ------------------------------------------------------------------ */
static int chol_factor(double *A, size_t n) { /* lower triangle in place; 1 if not PD */
    for (size_t j = 0; j < n; j++) {
        long double s = A[j * n + j];
        for (size_t k = 0; k < j; k++) s -= (long double)A[j * n + k] * A[j * n + k];
        if (!(s > 0)) return 1;
        A[j * n + j] = sqrtl(s);
        for (size_t i = j + 1; i < n; i++) {
            long double t = A[i * n + j];
            for (size_t k = 0; k < j; k++) t -= (long double)A[i * n + k] * A[j * n + k];
            A[i * n + j] = t / A[j * n + j];
        }
    }
    return 0;
}
static void chol_solve_vec(const double *L, double *b, size_t n) {
    for (size_t i = 0; i < n; i++) { long double t = b[i]; for (size_t k = 0; k < i; k++) t -= (long double)L[i * n + k] * b[k]; b[i] = t / L[i * n + i]; }
    for (size_t i = n; i-- > 0;) { long double t = b[i]; for (size_t k = i + 1; k < n; k++) t -= (long double)L[k * n + i] * b[k]; b[i] = t / L[i * n + i]; }
}
static int spd_inverse(const double *A, double *inv, size_t n) {
    double *L = malloc(sizeof(double) * n * n);
    memcpy(L, A, sizeof(double) * n * n);
    if (chol_factor(L, n)) { free(L); return 1; }
    for (size_t c = 0; c < n; c++) {
        double *col = calloc(n, sizeof(double));
        col[c] = 1;
        chol_solve_vec(L, col, n);
        for (size_t r = 0; r < n; r++) inv[r * n + c] = col[r];
        free(col);
    }
    free(L);
    return 0;
}
/* apop_data_covariance (apop_stats.m4.c:578-593): sample covariance of the columns */
apop_data *apop_data_covariance(const apop_data *in) {
    if (!in || !in->matrix) return NULL;
    const size_t n = in->matrix->size1, p = in->matrix->size2;
    apop_data *out = apop_data_calloc(p, p, 0);
    long double *mean = calloc(p, sizeof(long double));
    for (size_t i = 0; i < n; i++) for (size_t j = 0; j < p; j++) mean[j] += in->matrix->data[i * in->matrix->tda + j];
    for (size_t j = 0; j < p; j++) mean[j] /= n;
    for (size_t a = 0; a < p; a++)
        for (size_t b = a; b < p; b++) {
            long double s = 0;
            for (size_t i = 0; i < n; i++) s += (in->matrix->data[i * in->matrix->tda + a] - mean[a]) * (in->matrix->data[i * in->matrix->tda + b] - mean[b]);
            out->matrix->data[a * p + b] = out->matrix->data[b * p + a] = (double)(s / (n - 1));
        }
    free(mean);
    return out;
}

/* general inverse (Gauss-Jordan with partial pivoting), for Hessians that are not positive definite */
static int general_inverse(const double *A, double *inv, size_t n) {
    long double *W = malloc(sizeof(long double) * n * 2 * n);
    for (size_t i = 0; i < n; i++) for (size_t j = 0; j < 2 * n; j++) W[i * 2 * n + j] = j < n ? A[i * n + j] : (j - n == i ? 1 : 0);
    for (size_t c = 0; c < n; c++) {
        size_t piv = c;
        for (size_t i = c + 1; i < n; i++) if (fabsl(W[i * 2 * n + c]) > fabsl(W[piv * 2 * n + c])) piv = i;
        if (!(fabsl(W[piv * 2 * n + c]) > 0)) { free(W); return 1; }
        if (piv != c) for (size_t j = 0; j < 2 * n; j++) { long double t = W[c * 2 * n + j]; W[c * 2 * n + j] = W[piv * 2 * n + j]; W[piv * 2 * n + j] = t; }
        const long double d = W[c * 2 * n + c];
        for (size_t j = 0; j < 2 * n; j++) W[c * 2 * n + j] /= d;
        for (size_t i = 0; i < n; i++) {
            if (i == c) continue;
            const long double f = W[i * 2 * n + c];
            if (f != 0) for (size_t j = 0; j < 2 * n; j++) W[i * 2 * n + j] -= f * W[c * 2 * n + j];
        }
    }
    for (size_t i = 0; i < n; i++) for (size_t j = 0; j < n; j++) inv[i * n + j] = (double)W[i * 2 * n + n + j];
    free(W);
    return 0;
}

/* ---- second derivatives ----------------------------------------------------------------
 * apop_model_hessian (apop_mle.m4.c:187-223): the numeric gradient (gsl_deriv_central's 5-point rule
 * at delta, :92-108) of every score component, the two estimates of each off-diagonal element averaged.
 * The reference differentiates one component at a time (4 P score passes each); one score pass returns
 * all P components, so perturbing parameter j serves column j of every component: 4 P passes in all.
 * The score comes from the vtable (device entry) or, without one, from apop_numerical_gradient. */
apop_data *apop_model_hessian(apop_data *data, apop_model *model, double delta) {
    if (!model || !model->parameters) return NULL;
    if (!delta) { apop_mle_settings *mp = apop_settings_get_grp(model, "apop_mle"); delta = mp ? mp->delta : 1e-3; }
    gsl_vector *beta = apop_data_pack(model->parameters, NULL);
    if (!beta) return NULL;
    const size_t P = beta->size;
    apop_data *out = apop_data_calloc(P, P, 0);
    gsl_vector *work = gsl_vector_alloc(P), *sc[4];
    for (int q = 0; q < 4; q++) sc[q] = gsl_vector_alloc(P);
    const double at[4] = {-delta, -delta / 2, delta / 2, delta};
    for (size_t j = 0; j < P; j++) {
        for (int q = 0; q < 4; q++) {
            memcpy(work->data, beta->data, sizeof(double) * P);
            work->data[j] += at[q];
            apop_data_unpack(work, model->parameters);
            apop_score(data, sc[q], model);
        }
        for (size_t k = 0; k < P; k++) {   /* d score_k / d beta_j */
            const double r3 = 0.5 * (sc[3]->data[k] - sc[0]->data[k]);
            const double r5 = (4.0 / 3.0) * (sc[2]->data[k] - sc[1]->data[k]) - (1.0 / 3.0) * r3;
            const double dv = r5 / delta;
            out->matrix->data[k * P + j] += dv / 2;
            out->matrix->data[j * P + k] += dv / 2;
        }
    }
    apop_data_unpack(beta, model->parameters);
    for (int q = 0; q < 4; q++) gsl_vector_free(sc[q]);
    gsl_vector_free(work); gsl_vector_free(beta);
    return out;
}

/* Observed information -d2LL/db db' at the model's parameters, P x P row-major: one device pass where the
 * family has an analytic form (probit, binary and multinomial logit, Poisson regression, OLS), else the
 * negated numeric Hessian above.  Returns 0 on success; *analytic tells which one it was. */
int apop_b200_observed_information(apop_data *data, apop_model *m, double *H, int *analytic) {
    if (!m || !m->parameters || !H) return 1;
    info_fn information = (info_fn)dev("apop_b200_information");
    const int fam = apop_b200_family_of(m);
    gsl_vector *b = apop_data_pack(m->parameters, NULL);
    if (!b) return 1;
    const size_t P = b->size;
    int rc = 1;
    if (analytic) *analytic = 0;
    if (information && (fam == 0 || fam == 1 || fam == 4 || fam == 5)) {
        rc = information(fam, data, b->data, P, H);
        if (rc == 0 && analytic) *analytic = 1;
    }
    if (rc != 0) {
        apop_data *hess = apop_model_hessian(data, m, 0);
        if (hess) {
            for (size_t i = 0; i < P * P; i++) H[i] = -hess->matrix->data[i];
            apop_data_free(hess);
            rc = 0;
        }
    }
    gsl_vector_free(b);
    return rc;
}

/* ---- covariance of an estimate: inverse observed information (apop_mle.m4.c:239-265) ------------ */
apop_data *apop_model_numerical_covariance(apop_data *data, apop_model *m) {
    if (!m || !m->parameters) return NULL;
    gsl_vector *b = apop_data_pack(m->parameters, NULL);
    if (!b) return NULL;
    const size_t P = b->size;
    gsl_vector_free(b);
    double *H = malloc(sizeof(double) * P * P);
    apop_data *cov = NULL;
    if (apop_b200_observed_information(data, m, H, NULL) == 0) {
        cov = apop_data_calloc(P, P, 0);
        /* -(hessian)^-1 = (information)^-1; Cholesky when it is positive definite, else a general inverse */
        if (spd_inverse(H, cov->matrix->data, P) && general_inverse(H, cov->matrix->data, P)) cov->error = 'm';
        if (!apop_data_get_page(m->parameters, "<Covariance>")) apop_data_add_page(m->parameters, cov, "<Covariance>");
    }
    free(H);
    return cov;
}

/* ---- closed-form estimates over device reductions ------------------------------------ */
static double tsize_of(const apop_data *d) {
    return (d->vector ? (double)d->vector->size : 0) + (d->matrix ? (double)d->matrix->size1 * d->matrix->size2 : 0);
}
static void poisson_estimate(apop_data *d, apop_model *est) { /* lambda-hat = mean of every cell */
    mapsum_fn map_sum = (mapsum_fn)dev("apop_b200_map_sum");
    if (!map_sum) { est->error = 'm'; return; }
    if (!est->parameters) apop_model_clear(d, est);
    est->parameters->vector->data[0] = map_sum(d, 0 /*identity*/, 0, 'a') / tsize_of(d);
    apop_data_free(est->info);
    est->info = apop_data_alloc(1, 0, 0);
    est->info->vector->data[0] = apop_log_likelihood(d, est);
}
static void normal_estimate(apop_data *d, apop_model *est) { /* mu-hat = mean, sigma-hat = sqrt(var), get_mu_var */
    mapsum_fn map_sum = (mapsum_fn)dev("apop_b200_map_sum");
    if (!map_sum) { est->error = 'm'; return; }
    if (!est->parameters) apop_model_clear(d, est);
    const double n = tsize_of(d), mu = map_sum(d, 0, 0, 'a') / n;
    const double ss = map_sum(d, 2 /*(x-mu)^2*/, mu, 'a');
    /* one part present: sample variance (n-1) for a vector, E[x^2]-E[x]^2 for a matrix; both: /n */
    const double var = (d->vector && !d->matrix) ? ss / (n - 1) : ss / n;
    est->parameters->vector->data[0] = mu;
    est->parameters->vector->data[1] = sqrt(var);
    est->data = d;
    apop_data_free(est->info);
    est->info = apop_data_alloc(1, 0, 0);
    est->info->vector->data[0] = apop_log_likelihood(d, est);
}
static void ols_estimate(apop_data *d, apop_model *est) { /* beta = (X'X)^-1 X'y from one device pass */
    gram_fn gram = (gram_fn)dev("apop_b200_ols_gram");
    if (!gram || !d->matrix) { est->error = 'm'; return; }
    const size_t k = d->matrix->size2;
    if (!est->parameters || !est->parameters->vector || est->parameters->vector->size != k) apop_model_clear(d, est);
    double *A = malloc(sizeof(double) * k * k), *b = malloc(sizeof(double) * k), *s = malloc(sizeof(double) * k);
    if (gram(d, A, b)) { est->error = 'm'; free(A); free(b); free(s); return; }
    /* equilibrate (the columns of the NIST problems span 13 orders of magnitude), then Cholesky */
    for (size_t i = 0; i < k; i++) s[i] = 1 / sqrt(A[i * k + i]);
    for (size_t i = 0; i < k; i++) { for (size_t j = 0; j < k; j++) A[i * k + j] *= s[i] * s[j]; b[i] *= s[i]; }
    if (chol_factor(A, k)) est->error = 'm';
    else { chol_solve_vec(A, b, k); for (size_t i = 0; i < k; i++) est->parameters->vector->data[i] = b[i] * s[i]; }
    est->data = d;
    apop_data_free(est->info);
    est->info = apop_data_alloc(1, 0, 0);
    est->info->vector->data[0] = est->error ? NAN : apop_log_likelihood(d, est);
    free(A); free(b); free(s);
}
void apop_b200_install_estimates(void) {
    apop_poisson->estimate = poisson_estimate;
    apop_normal->estimate = normal_estimate;
    apop_ols->estimate = ols_estimate;
}

/* ---- bootstrap ------------------------------------------------------------------------ */
static apop_data *bootstrap_host_loop(apop_data *data, apop_model *model, uint64_t seed, int iterations, apop_data **boot_store) {
    /* the reference's own loop, for models outside the batched path: same pointer, new
     * contents every replicate -> the resident copy must be invalidated */
    data_fn invalidate = (data_fn)dev("apop_b200_invalidate"), pin = (data_fn)dev("apop_b200_pin"), unpin = (data_fn)dev("apop_b200_unpin");
    if (!invalidate || !pin || !data || (data->matrix && !data->matrix->data)) return NULL;
    const size_t n = data->matrix ? data->matrix->size1 : data->vector->size, k = data->matrix ? data->matrix->size2 : 0;
    apop_data *subset = apop_data_alloc(data->vector ? n : 0, data->matrix ? n : 0, (int)k);
    hrng rng; hrng_seed(&rng, seed);
    apop_data *boots = NULL;
    for (int i = 0; i < iterations; i++) {
        for (size_t j = 0; j < n; j++) {
            const size_t r = (size_t)(hrng_uniform(&rng) * n);
            if (data->matrix) memcpy(subset->matrix->data + j * k, data->matrix->data + r * data->matrix->tda, sizeof(double) * k);
            if (data->vector) subset->vector->data[j] = data->vector->data[r * data->vector->stride];
        }
        if (i == 0) pin(subset); else invalidate(subset);
        apop_model *e = apop_model_copy(model);
        e->data = subset;   /* already prepped data: keep prep a no-op */
        if (e->estimate) e->estimate(subset, e); else apop_maximum_likelihood(subset, e);
        gsl_vector *p = apop_data_pack(e->parameters, NULL);
        if (!boots) boots = apop_data_calloc(iterations, p->size, 0);
        memcpy(boots->matrix->data + (size_t)i * p->size, p->data, sizeof(double) * p->size);
        gsl_vector_free(p);
        apop_model_free(e);
    }
    unpin(subset);
    apop_data_free(subset);
    apop_data *cov = apop_data_covariance(boots);
    if (boot_store) *boot_store = boots; else apop_data_free(boots);
    return cov;
}

/* Families the batched kernel does not cover but whose single-beta kernel takes resample counts (the multinomial
 * logit): the reference's own loop -- one full fit per replicate (apop_bootstrap.m4.c:143-159) -- over the SAME resident
 * data: replicate r is selected with apop_b200_set_replicate, which weights every row by its multiplicity in that
 * resample, so no row is copied and nothing is re-uploaded.  Every fit is warm-started at the full-sample estimate. */
static apop_data *bootstrap_replicate_loop(apop_data *data, apop_model *model, uint64_t seed, int iterations, apop_data **boot_store) {
    counts_fn counts = (counts_fn)dev("apop_b200_bootstrap_counts");
    int (*set_rep)(apop_data *, long) = (int (*)(apop_data *, long))dev("apop_b200_set_replicate");
    if (!counts || !set_rep) return NULL;
    if (counts(data, (size_t)iterations, seed)) return NULL;
    gsl_vector *b0 = apop_data_pack(model->parameters, NULL);
    const size_t P = b0->size;
    apop_data *boots = apop_data_calloc(iterations, P, 0);
    int ok = 1;
    for (int i = 0; ok && i < iterations; i++) {
        if (set_rep(data, i)) { ok = 0; break; }
        apop_model *e = apop_model_copy(model);
        e->data = data;
        apop_mle_settings *mp = apop_settings_get_grp(e, "apop_mle");
        if (!mp) { apop_mle_settings s = apop_mle_settings_defaults(); mp = apop_mle_settings_add(e, s); }
        mp->starting_pt = b0->data;                                  /* warm start: the full-sample estimate */
        if (!mp->method[0]) snprintf(mp->method, sizeof mp->method, "BFGS cg");
        if (mp->tolerance >= 1e-5 && !mp->relative_tolerance) { mp->tolerance = 1e-9; mp->relative_tolerance = 1; }
        apop_maximum_likelihood(data, e);
        gsl_vector *p = apop_data_pack(e->parameters, NULL);
        memcpy(boots->matrix->data + (size_t)i * P, p->data, sizeof(double) * P);
        if (e->error) ok = 0;
        gsl_vector_free(p);
        apop_model_free(e);
    }
    set_rep(data, -1);
    gsl_vector_free(b0);
    apop_data *cov = ok ? apop_data_covariance(boots) : NULL;
    if (boot_store && ok) *boot_store = boots; else apop_data_free(boots);
    return cov;
}

/* Bootstrap covariance of the parameter estimates.  `model` must be estimated already
 * (its parameters are the warm start of every replicate) and carry a registered device
 * likelihood; `data` must be pinned. */
apop_data *apop_bootstrap_cov(apop_data *data, apop_model *model, uint64_t seed, int iterations, apop_data **boot_store) {
    if (!data || !model || !model->parameters) return NULL;
    if (iterations <= 0) iterations = 1000;   /* the reference's default */
    const int fam = apop_b200_family_of(model);
    batched_fn batched = (batched_fn)dev("apop_b200_ll_score_batched");
    counts_fn counts = (counts_fn)dev("apop_b200_bootstrap_counts");
    info_fn information = (info_fn)dev("apop_b200_information");
    gsl_vector *b0 = apop_data_pack(model->parameters, NULL);
    const size_t P = b0->size, m = (size_t)iterations;
    const int binary = data->matrix && P == data->matrix->size2;
    if (fam == 1 && !binary && data->matrix && P % data->matrix->size2 == 0) {       /* multinomial logit */
        gsl_vector_free(b0);
        return bootstrap_replicate_loop(data, model, seed, iterations, boot_store);
    }
    if (!batched || !counts || !((fam == 0 && binary) || (fam == 1 && binary) || (fam == 5 && binary))) {
        gsl_vector_free(b0);
        return bootstrap_host_loop(data, model, seed, iterations, boot_store);
    }
    if (counts(data, m, seed)) { gsl_vector_free(b0); return NULL; }
    /* every replicate starts at the full-sample estimate with H0 = (full-sample information)^-1:
     * the resampled Hessians differ from it by O(N^-1/2), so the first steps are Newton-like */
    double *H0 = malloc(sizeof(double) * P * P), *Hinfo = malloc(sizeof(double) * P * P);
    int have_h0 = information && information(fam, data, b0->data, P, Hinfo) == 0 && spd_inverse(Hinfo, H0, P) == 0;
    double *X = malloc(sizeof(double) * m * P), *XT = malloc(sizeof(double) * m * P), *G = malloc(sizeof(double) * m * P),
           *GT = malloc(sizeof(double) * m * P), *Pd = malloc(sizeof(double) * m * P), *Hs = malloc(sizeof(double) * m * P * P),
           *F = malloc(sizeof(double) * m), *FT = malloc(sizeof(double) * m), *alpha = malloc(sizeof(double) * m);
    int *state = calloc(m, sizeof(int)); /* 0 = needs a direction, 1 = trial pending, 2 = done */
    for (size_t r = 0; r < m; r++) {
        memcpy(X + r * P, b0->data, sizeof(double) * P);
        if (have_h0) memcpy(Hs + r * P * P, H0, sizeof(double) * P * P);
        else { memset(Hs + r * P * P, 0, sizeof(double) * P * P); for (size_t i = 0; i < P; i++) Hs[r * P * P + i * P + i] = 1; }
    }
    int ok = batched(fam, data, X, P, m, 1, F, G) == 0;
    for (size_t r = 0; r < m && ok; r++) { F[r] = -F[r]; for (size_t i = 0; i < P; i++) G[r * P + i] = -G[r * P + i]; }
    size_t n_done = 0;
    for (int pass = 0; ok && pass < 200 && n_done < m; pass++) {
        for (size_t r = 0; r < m; r++) {
            double *x = X + r * P, *g = G + r * P, *p = Pd + r * P, *H = Hs + r * P * P, *xt = XT + r * P;
            if (state[r] == 0) {
                double gn = 0; for (size_t i = 0; i < P; i++) gn += g[i] * g[i];
                if (sqrt(gn) < 1e-9 * fmax(1.0, fabs(F[r]))) { state[r] = 2; n_done++; }
                else {
                    double pg = 0;
                    for (size_t i = 0; i < P; i++) { double t = 0; for (size_t j = 0; j < P; j++) t += H[i * P + j] * g[j]; p[i] = -t; pg += p[i] * g[i]; }
                    if (!(pg < 0)) { const double sc = have_h0 ? 1.0 : 1.0 / sqrt(gn); for (size_t i = 0; i < P; i++) p[i] = -sc * g[i]; }
                    alpha[r] = 1.0;
                    state[r] = 1;
                }
            }
            if (state[r] == 1) for (size_t i = 0; i < P; i++) xt[i] = x[i] + alpha[r] * p[i];
            else memcpy(xt, x, sizeof(double) * P);
        }
        if (n_done == m) break;
        ok = batched(fam, data, XT, P, m, 1, FT, GT) == 0;
        if (!ok) break;
        for (size_t r = 0; r < m; r++) {
            if (state[r] != 1) continue;
            double *x = X + r * P, *g = G + r * P, *p = Pd + r * P, *H = Hs + r * P * P, *xt = XT + r * P, *gt = GT + r * P;
            const double ft = -FT[r];
            double pg = 0; for (size_t i = 0; i < P; i++) pg += p[i] * g[i];
            /* the achievable decrease is bounded below by the rounding noise of the sums */
            const double noise = 1e-14 * fabs(F[r]);
            if (isfinite(ft) && ft <= F[r] + 1e-4 * alpha[r] * pg + noise) {
                double sy = 0, yy = 0, s[64], y[64], Hy[64];   /* P <= 32 on the batched path */
                for (size_t i = 0; i < P; i++) { s[i] = xt[i] - x[i]; y[i] = -gt[i] - g[i]; sy += s[i] * y[i]; yy += y[i] * y[i]; }
                if (sy > 1e-12 * sqrt(yy)) {
                    double yHy = 0;
                    for (size_t i = 0; i < P; i++) { double t = 0; for (size_t j = 0; j < P; j++) t += H[i * P + j] * y[j]; Hy[i] = t; yHy += t * y[i]; }
                    const double rho = 1 / sy;
                    for (size_t i = 0; i < P; i++) for (size_t j = 0; j < P; j++) H[i * P + j] += -rho * (Hy[i] * s[j] + s[i] * Hy[j]) + rho * rho * (sy + yHy) * s[i] * s[j];
                }
                memcpy(x, xt, sizeof(double) * P);
                for (size_t i = 0; i < P; i++) g[i] = -gt[i];
                F[r] = ft;
                state[r] = 0;
            } else {
                alpha[r] *= 0.5;
                /* no further measurable progress along this direction: the replicate is done */
                if (alpha[r] < 1e-12 || alpha[r] * fabs(pg) < noise) { state[r] = 2; n_done++; }
            }
        }
    }
    apop_data *boots = apop_data_alloc(m, P, 0), *cov = NULL;
    memcpy(boots->matrix->data, X, sizeof(double) * m * P);
    if (ok) cov = apop_data_covariance(boots);
    if (boot_store && ok) *boot_store = boots; else apop_data_free(boots);
    free(H0); free(Hinfo); free(X); free(XT); free(G); free(GT); free(Pd); free(Hs); free(F); free(FT); free(alpha); free(state);
    gsl_vector_free(b0);
    return cov;
}

/* ---- Metropolis ------------------------------------------------------------------------- */
apop_mcmc_settings apop_mcmc_settings_defaults(void) { /* apop_mcmc.m4.c:35-43 */
    return (apop_mcmc_settings){.periods = 6000, .burnin = 0.05, .target_accept_rate = 0.35, .initial_scale = 1.0, .seed = 8, .n_chains = 1};
}
/* sigma_adapt (apop_mcmc.m4.c:18-32): after an acceptance the proposal variance is scaled by
 * 1 + (ar/target - 1)/100, ar smoothed by 1% of the periods */
static double sigma_adapt(double var, long acc, long rej, const apop_mcmc_settings *s) {
    const double ar = (acc + .01 * s->periods * s->target_accept_rate) / (acc + rej + .01 * s->periods);
    return var * (1 + (ar / s->target_accept_rate - 1) / 100.);
}

/* one chain, exactly the unmodified caller: one model->log_likelihood pass per step */
apop_data *apop_model_metropolis(apop_data *d, apop_model *m, apop_mcmc_settings s) {
    if (!m || !m->parameters) return NULL;
    gsl_vector *cur = apop_data_pack(m->parameters, NULL), *prop = gsl_vector_alloc(cur->size);
    const size_t P = cur->size;
    const int burn = (int)(s.burnin * s.periods), keep = s.periods - burn;
    apop_data *out = apop_data_calloc(keep, P, 0);
    hrng rng; hrng_seed(&rng, s.seed);
    double var = s.initial_scale * s.initial_scale, last_ll = apop_log_likelihood(d, m);
    long acc = 0, rej = 0;
    for (int t = 0; t < s.periods; t++) {
        double ll = NAN;
        for (int tries = 0; tries < 100 && !isfinite(ll); tries++) { /* NaN/inf: throw out and draw again (:118) */
            for (size_t i = 0; i < P; i++) prop->data[i] = cur->data[i] + sqrt(var) * hrng_normal(&rng);
            apop_data_unpack(prop, m->parameters);
            if (m->constraint && m->constraint(d, m)) continue;
            ll = apop_log_likelihood(d, m);
        }
        const double ratio = ll - last_ll;
        if (isfinite(ll) && (ratio >= 0 || log(hrng_uniform(&rng)) < ratio)) {
            memcpy(cur->data, prop->data, sizeof(double) * P);
            last_ll = ll; acc++;
            var = sigma_adapt(var, acc, rej, &s);
        } else { rej++; apop_data_unpack(cur, m->parameters); }
        if (t >= burn) memcpy(out->matrix->data + (size_t)(t - burn) * P, cur->data, sizeof(double) * P);
    }
    apop_data_unpack(cur, m->parameters);
    gsl_vector_free(cur); gsl_vector_free(prop);
    return out;
}

/* n_chains chains in lock-step: one batched LL pass per step for all of them.  The draws
 * come back as (n_chains * kept periods) x P, chain-major. */
apop_data *apop_b200_metropolis_chains(apop_data *d, apop_model *m, apop_mcmc_settings s) {
    batched_fn batched = (batched_fn)dev("apop_b200_ll_score_batched");
    const int fam = apop_b200_family_of(m);
    if (!batched || fam < 0 || !m->parameters || s.n_chains < 1) return NULL;
    gsl_vector *start = apop_data_pack(m->parameters, NULL);
    const size_t P = start->size, C = (size_t)s.n_chains;
    const int burn = (int)(s.burnin * s.periods), keep = s.periods - burn;
    apop_data *out = apop_data_calloc(C * keep, P, 0);
    double *cur = malloc(sizeof(double) * C * P), *prop = malloc(sizeof(double) * C * P), *ll = malloc(sizeof(double) * C),
           *last = malloc(sizeof(double) * C), *var = malloc(sizeof(double) * C);
    long *acc = calloc(C, sizeof(long)), *rej = calloc(C, sizeof(long));
    hrng rng; hrng_seed(&rng, s.seed);
    for (size_t c = 0; c < C; c++) { memcpy(cur + c * P, start->data, sizeof(double) * P); var[c] = s.initial_scale * s.initial_scale; }
    int ok = batched(fam, d, cur, P, C, 0, last, NULL) == 0;
    for (int t = 0; ok && t < s.periods; t++) {
        for (size_t c = 0; c < C; c++) for (size_t i = 0; i < P; i++) prop[c * P + i] = cur[c * P + i] + sqrt(var[c]) * hrng_normal(&rng);
        ok = batched(fam, d, prop, P, C, 0, ll, NULL) == 0;
        for (size_t c = 0; c < C && ok; c++) {
            const double ratio = ll[c] - last[c];
            if (isfinite(ll[c]) && (ratio >= 0 || log(hrng_uniform(&rng)) < ratio)) {
                memcpy(cur + c * P, prop + c * P, sizeof(double) * P);
                last[c] = ll[c]; acc[c]++;
                var[c] = sigma_adapt(var[c], acc[c], rej[c], &s);
            } else rej[c]++;
            if (t >= burn) memcpy(out->matrix->data + (c * keep + (size_t)(t - burn)) * P, cur + c * P, sizeof(double) * P);
        }
    }
    free(cur); free(prop); free(ll); free(last); free(var); free(acc); free(rej);
    gsl_vector_free(start);
    if (!ok) { apop_data_free(out); return NULL; }
    return out;
}
