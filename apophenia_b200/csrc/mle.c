/* mle.c -- host optimizer around the device likelihood/score (GSL multimin is not in
 * this image).  Mirrors apop_maximum_likelihood (apop_mle.m4.c:635-660) and
 * its shells:
 *   negshell  (:296-330)  unpack beta -> parameters, constraint penalty, f = penalty - LL,
 *                         NaN ends the search with status -1
 *   dnegshell (:332-359)  score through the apop_score vtable (numeric fallback), negated
 *   default method (:638-642): score in the vtable ? "FR cg" : "NM simplex"
 *   stop rule (:458): |gradient| < tolerance (absolute), max_iterations, "no progress"
 *                     ends the search with status 0 like GSL_ENOPROG does there (:455-457)
 * f and df at the same beta cost one fused device pass (the library memoises the last
 * evaluation), which is what fdf_shell (:362-365) wants.
 *
 * The minimisers are written from the textbook algorithms (Nocedal & Wright: BFGS with a
 * strong-Wolfe bracketing/zoom line search; Fletcher-Reeves CG with restarts; Nelder-Mead;
 * Newton on the device-computed observed information).  Trajectories are not GSL's --
 * optimizer paths are "parity unpinned" (SURVEY 8(c)); fitted parameters are compared
 * between device-driven and oracle-driven runs of THIS optimizer (tests/).
 */
#include "hostmini.h"
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct {
    apop_data *data;
    apop_model *model;
    size_t n;
    int f_evals, df_evals, nan_seen;
    double best_ll;
    gsl_vector *bv, *gv;
} mle_info;

static __thread apop_mle_report last_report;
apop_mle_report apop_mle_last_report(void) { return last_report; }

static double negshell(const double *beta, mle_info *I) {
    memcpy(I->bv->data, beta, sizeof(double) * I->n);
    apop_data_unpack(I->bv, I->model->parameters);
    double penalty = 0;
    if (I->model->constraint) penalty = (double)I->model->constraint(I->data, I->model);
    const double f = apop_log_likelihood(I->data, I->model);
    I->f_evals++;
    const double out = penalty - f;
    if (isnan(out)) { I->nan_seen = 1; return NAN; }
    if (-out > I->best_ll) I->best_ll = -out;
    return out;
}
static void dnegshell(const double *beta, mle_info *I, double *g) {
    memcpy(I->bv->data, beta, sizeof(double) * I->n);
    apop_data_unpack(I->bv, I->model->parameters);
    apop_score(I->data, I->gv, I->model);
    I->df_evals++;
    for (size_t j = 0; j < I->n; j++) g[j] = -I->gv->data[j];
}
static double fdf(const double *x, mle_info *I, double *g) {
    const double f = negshell(x, I);
    if (!isnan(f)) dnegshell(x, I, g);
    return f;
}
static double dot(const double *a, const double *b, size_t n) { double s = 0; for (size_t i = 0; i < n; i++) s += a[i] * b[i]; return s; }
static double nrm(const double *a, size_t n) { return sqrt(dot(a, a, n)); }

/* Strong-Wolfe line search along p from x (Nocedal & Wright alg. 3.5 / 3.6).
 * On success x,f,g are replaced by the accepted point. Returns 0 ok, 1 no progress, -1 NaN everywhere. */
typedef struct { double a, f, d; } ls_pt;
static int ls_eval(mle_info *I, size_t n, const double *x, const double *p, double a, double *xt, double *gt, ls_pt *out) {
    for (size_t i = 0; i < n; i++) xt[i] = x[i] + a * p[i];
    const double f = fdf(xt, I, gt);
    out->a = a; out->f = f; out->d = isnan(f) ? NAN : dot(gt, p, n);
    return isnan(f) || !isfinite(f);
}
static int line_search(mle_info *I, size_t n, double *x, double *f, double *g, const double *p, double a0, double c2) {
    const double c1 = 1e-4, f0 = *f, d0 = dot(g, p, n);
    if (!(d0 < 0)) return 1;
    double *xt = malloc(sizeof(double) * n), *gt = malloc(sizeof(double) * n);
    ls_pt lo = {0, f0, d0}, hi = {0, 0, 0}, cur;
    int have_hi = 0, rc = 1, any_finite = 0;
    double a = a0;
    for (int it = 0; it < 40; it++) {
        const int bad = ls_eval(I, n, x, p, a, xt, gt, &cur);
        if (!bad) any_finite = 1;
        if (bad) { /* NaN or inf: treat as too long a step (a search whose trials are ALL NaN ends with -1 below) */
            hi = (ls_pt){a, INFINITY, 0}; have_hi = 1;
            a = lo.a + 0.25 * (a - lo.a);
            if (a - lo.a < 1e-16 * fmax(1.0, lo.a)) break;
            continue;
        }
        /* near the optimum the decrease of f drops below its rounding noise (|f| ~ 1e8 at 1e8 rows):
         * there the sufficient-decrease test is replaced by Hager & Zhang's approximate Wolfe
         * condition on the directional derivative, (2 delta - 1) d0 >= d(a) >= sigma d0 */
        const double noise = 1e-14 * fabs(f0);
        if (fabs(cur.f - f0) <= noise && cur.d >= c2 * d0 && cur.d <= (2 * 0.1 - 1) * d0) { rc = 0; break; }
        const int armijo_fail = cur.f > f0 + c1 * cur.a * d0 + noise || (lo.a > 0 && cur.f >= lo.f + noise);
        if (armijo_fail) { hi = cur; have_hi = 1; }
        else if (fabs(cur.d) <= -c2 * d0) { rc = 0; break; }
        else if (have_hi) { if (cur.d * (hi.a - lo.a) >= 0) hi = lo; lo = cur; }
        else if (cur.d >= 0) { hi = lo; lo = cur; have_hi = 1; }
        else { lo = cur; a *= 2.5; continue; } /* still descending: extend the bracket */
        /* zoom: safeguarded cubic (or bisection when the bracket end is infinite) */
        double an;
        if (isfinite(hi.f) && hi.a != lo.a) {
            const double da = hi.a - lo.a;
            if (hi.d != 0) {
                const double d1 = lo.d + hi.d - 3 * (lo.f - hi.f) / (lo.a - hi.a);
                const double rad = d1 * d1 - lo.d * hi.d;
                if (rad >= 0) {
                    const double d2 = (da > 0 ? 1 : -1) * sqrt(rad);
                    an = hi.a - da * (hi.d + d2 - d1) / (hi.d - lo.d + 2 * d2);
                } else an = lo.a + 0.5 * da;
            } else { /* quadratic through (lo.f, lo.d, hi.f) */
                const double den = 2 * (hi.f - lo.f - lo.d * da);
                an = den > 0 ? lo.a - lo.d * da * da / den : lo.a + 0.5 * da;
            }
            const double amin = fmin(lo.a, hi.a), amax = fmax(lo.a, hi.a), w = amax - amin;
            if (!(an > amin + 0.05 * w && an < amax - 0.05 * w)) an = lo.a + 0.5 * da;
        } else an = lo.a + 0.5 * (hi.a - lo.a);
        if (fabs(an - lo.a) < 1e-15 * fmax(1.0, fabs(lo.a))) { /* bracket collapsed */
            if (lo.a > 0 && lo.f < f0) { ls_eval(I, n, x, p, lo.a, xt, gt, &cur); rc = 0; }
            break;
        }
        a = an;
    }
    if (rc != 0 && lo.a > 0 && lo.f < f0) { /* take the best Armijo point we have */
        ls_eval(I, n, x, p, lo.a, xt, gt, &cur);
        rc = 0;
    }
    if (rc == 0) { memcpy(x, xt, sizeof(double) * n); memcpy(g, gt, sizeof(double) * n); *f = cur.f; }
    free(xt); free(gt);
    /* NaN trials next to finite ones are over-long steps; NaN everywhere is the reference's
     * negshell longjmp: the fit ends with status -1 (apop_mle.m4.c:310,451-453) */
    if (any_finite) I->nan_seen = 0;
    else if (I->nan_seen) return -1;
    return rc;
}

static int converged(const apop_mle_settings *mp, double gnorm, double f) {
    const double tol = mp->relative_tolerance ? mp->tolerance * fmax(1.0, fabs(f)) : mp->tolerance;
    return gnorm < tol;
}

/* "BFGS cg" and "FR cg" */
static int minimize_gradient(mle_info *I, const apop_mle_settings *mp, double *x, int bfgs, int *iters, double *gnorm_out) {
    const size_t n = I->n;
    double *g = malloc(sizeof(double) * n), *p = malloc(sizeof(double) * n), *xo = malloc(sizeof(double) * n),
           *go = malloc(sizeof(double) * n), *H = NULL, *s = malloc(sizeof(double) * n), *y = malloc(sizeof(double) * n),
           *Hy = malloc(sizeof(double) * n);
    if (bfgs) { H = calloc(n * n, sizeof(double)); for (size_t i = 0; i < n; i++) H[i * n + i] = 1; }
    double f = fdf(x, I, g);
    int status = 0, it = 0;
    if (isnan(f)) { status = -1; goto done; }
    for (size_t i = 0; i < n; i++) p[i] = -g[i];
    double gnorm = nrm(g, n);
    int first = 1;
    while (it < mp->max_iterations) {
        if (converged(mp, gnorm, f)) break;
        it++;
        memcpy(xo, x, sizeof(double) * n); memcpy(go, g, sizeof(double) * n);
        /* first step has length step_size, as gsl_multimin_fdfminimizer_set's step_size asks */
        const double a0 = first ? mp->step_size / fmax(nrm(p, n), 1e-300) : 1.0;
        int rc = line_search(I, n, x, &f, g, p, bfgs && !first ? 1.0 : a0, bfgs ? 0.9 : 0.1);
        if (rc != 0 && !first) { /* restart along steepest descent before giving up */
            for (size_t i = 0; i < n; i++) p[i] = -g[i];
            if (bfgs) { memset(H, 0, sizeof(double) * n * n); for (size_t i = 0; i < n; i++) H[i * n + i] = 1; }
            rc = line_search(I, n, x, &f, g, p, mp->step_size / fmax(nrm(p, n), 1e-300), 0.9);
        }
        if (I->nan_seen) { status = -1; break; }
        if (rc != 0) break; /* no further progress: the reference ends with status 0 on GSL_ENOPROG */
        first = 0;
        const double gnew = nrm(g, n);
        if (mp->verbose) printf("%5i f()=%.10g |gradient|=%.3g\n", it, f, gnew);
        if (bfgs) {
            for (size_t i = 0; i < n; i++) { s[i] = x[i] - xo[i]; y[i] = g[i] - go[i]; }
            const double sy = dot(s, y, n);
            if (sy > 1e-12 * nrm(s, n) * nrm(y, n)) {
                if (it == 1) { const double sc = sy / dot(y, y, n); for (size_t i = 0; i < n; i++) H[i * n + i] = sc; }
                for (size_t i = 0; i < n; i++) { double t = 0; for (size_t j = 0; j < n; j++) t += H[i * n + j] * y[j]; Hy[i] = t; }
                const double yHy = dot(y, Hy, n), rho = 1 / sy;
                for (size_t i = 0; i < n; i++)
                    for (size_t j = 0; j < n; j++)
                        H[i * n + j] += -rho * (Hy[i] * s[j] + s[i] * Hy[j]) + rho * rho * (sy + yHy) * s[i] * s[j];
            }
            for (size_t i = 0; i < n; i++) { double t = 0; for (size_t j = 0; j < n; j++) t += H[i * n + j] * g[j]; p[i] = -t; }
            if (!(dot(p, g, n) < 0)) for (size_t i = 0; i < n; i++) p[i] = -g[i];
        } else {
            const double beta = (it % (int)(n + 1) == 0) ? 0 : (gnew * gnew) / fmax(gnorm * gnorm, 1e-300);
            for (size_t i = 0; i < n; i++) p[i] = -g[i] + beta * p[i];
            if (!(dot(p, g, n) < 0)) for (size_t i = 0; i < n; i++) p[i] = -g[i];
            /* CG has no natural unit step: start from the previous accepted length */
            double steplen = 0; for (size_t i = 0; i < n; i++) steplen += (x[i] - xo[i]) * (x[i] - xo[i]);
            steplen = sqrt(steplen);
            const double pn = fmax(nrm(p, n), 1e-300);
            for (size_t i = 0; i < n; i++) p[i] *= fmax(steplen, 1e-12) / pn; /* so that a=1 repeats the last length */
        }
        gnorm = gnew;
    }
    if (it >= mp->max_iterations) status = -1;
    *gnorm_out = gnorm;
done:
    *iters = it;
    free(g); free(p); free(xo); free(go); free(H); free(s); free(y); free(Hy);
    return status;
}

/* "NM simplex": the reference's default when no score is registered (:475-527) */
static int minimize_simplex(mle_info *I, const apop_mle_settings *mp, double *x, int *iters) {
    const size_t n = I->n, m = n + 1;
    double *S = malloc(sizeof(double) * m * n), *fv = malloc(sizeof(double) * m), *c = malloc(sizeof(double) * n),
           *xr = malloc(sizeof(double) * n), *xe = malloc(sizeof(double) * n);
    for (size_t i = 0; i < m; i++) { memcpy(S + i * n, x, sizeof(double) * n); if (i) S[i * n + i - 1] += mp->step_size; fv[i] = negshell(S + i * n, I); }
    int it = 0, status = 0;
    for (; it < mp->max_iterations && !I->nan_seen; it++) {
        size_t lo = 0, hi = 0, nh = 0;
        for (size_t i = 0; i < m; i++) { if (fv[i] < fv[lo]) lo = i; if (fv[i] > fv[hi]) hi = i; }
        nh = lo; for (size_t i = 0; i < m; i++) if (i != hi && fv[i] > fv[nh]) nh = i;
        double size = 0; /* gsl_multimin_test_size: mean distance to the centre */
        memset(c, 0, sizeof(double) * n);
        for (size_t i = 0; i < m; i++) for (size_t j = 0; j < n; j++) c[j] += S[i * n + j] / m;
        for (size_t i = 0; i < m; i++) { double d = 0; for (size_t j = 0; j < n; j++) d += (S[i * n + j] - c[j]) * (S[i * n + j] - c[j]); size += sqrt(d) / m; }
        if (size < mp->tolerance) break;
        memset(c, 0, sizeof(double) * n);
        for (size_t i = 0; i < m; i++) if (i != hi) for (size_t j = 0; j < n; j++) c[j] += S[i * n + j] / n;
        for (size_t j = 0; j < n; j++) xr[j] = c[j] + (c[j] - S[hi * n + j]);
        const double fr = negshell(xr, I);
        if (fr < fv[lo]) {
            for (size_t j = 0; j < n; j++) xe[j] = c[j] + 2 * (c[j] - S[hi * n + j]);
            const double fe = negshell(xe, I);
            if (fe < fr) { memcpy(S + hi * n, xe, sizeof(double) * n); fv[hi] = fe; } else { memcpy(S + hi * n, xr, sizeof(double) * n); fv[hi] = fr; }
        } else if (fr < fv[nh]) { memcpy(S + hi * n, xr, sizeof(double) * n); fv[hi] = fr; }
        else {
            const int outside = fr < fv[hi];
            for (size_t j = 0; j < n; j++) xe[j] = c[j] + 0.5 * ((outside ? xr[j] : S[hi * n + j]) - c[j]);
            const double fc = negshell(xe, I);
            if (fc < (outside ? fr : fv[hi])) { memcpy(S + hi * n, xe, sizeof(double) * n); fv[hi] = fc; }
            else for (size_t i = 0; i < m; i++) if (i != lo) { for (size_t j = 0; j < n; j++) S[i * n + j] = S[lo * n + j] + 0.5 * (S[i * n + j] - S[lo * n + j]); fv[i] = negshell(S + i * n, I); }
        }
    }
    size_t lo = 0; for (size_t i = 0; i < m; i++) if (fv[i] < fv[lo]) lo = i;
    memcpy(x, S + lo * n, sizeof(double) * n);
    if (it >= mp->max_iterations || I->nan_seen) status = -1;
    *iters = it;
    free(S); free(fv); free(c); free(xr); free(xe);
    return status;
}

/* "Newton": damped Newton on the score with the observed information computed by one
 * device pass (apop_b200_information), in place of the reference's finite-difference
 * Jacobian root finder (apop_mle.m4.c:885-916). */
static int chol_solve(double *A, double *b, size_t n) { /* A = L L', solves in place; 1 if not PD */
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
    for (size_t i = 0; i < n; i++) { long double t = b[i]; for (size_t k = 0; k < i; k++) t -= (long double)A[i * n + k] * b[k]; b[i] = t / A[i * n + i]; }
    for (size_t i = n; i-- > 0;) { long double t = b[i]; for (size_t k = i + 1; k < n; k++) t -= (long double)A[k * n + i] * b[k]; b[i] = t / A[i * n + i]; }
    return 0;
}
static int minimize_newton(mle_info *I, const apop_mle_settings *mp, double *x, int *iters, double *gnorm_out) {
    const size_t n = I->n;
    const int fam = apop_b200_family_of(I->model);
    if (fam < 0) { fprintf(stderr, "Newton needs a registered device model\n"); return -1; }
    double *g = malloc(sizeof(double) * n), *p = malloc(sizeof(double) * n), *H = malloc(sizeof(double) * n * n);
    double f = fdf(x, I, g), gnorm = nrm(g, n);
    int it = 0, status = isnan(f) ? -1 : 0;
    while (status == 0 && it < mp->max_iterations && !converged(mp, gnorm, f)) {
        it++;
        /* the information at x: one device pass (probit, logit incl. multinomial, Poisson regression, OLS),
         * else the reference's numeric Hessian of the score (apop_mle.m4.c:885-916 takes a finite-difference
         * Jacobian for every model) */
        memcpy(I->bv->data, x, sizeof(double) * n);
        apop_data_unpack(I->bv, I->model->parameters);
        if (apop_b200_observed_information(I->data, I->model, H, NULL)) { status = -1; break; }
        for (size_t i = 0; i < n; i++) p[i] = -g[i];
        double ridge = 0;
        int solved = 0;
        for (int tries = 0; tries < 8; tries++) { /* observed information may be indefinite far from the optimum */
            double *A = malloc(sizeof(double) * n * n);
            memcpy(A, H, sizeof(double) * n * n);
            for (size_t i = 0; i < n; i++) { A[i * n + i] += ridge; p[i] = -g[i]; }
            const int bad = chol_solve(A, p, n);
            free(A);
            if (!bad) { solved = 1; break; }
            double tr = 0; for (size_t i = 0; i < n; i++) tr += fabs(H[i * n + i]);
            ridge = ridge ? ridge * 10 : 1e-6 * tr / n;
            if (!(ridge > 0)) break;              /* a vanishing information matrix: every probability saturated */
        }
        if (!solved) { /* no usable curvature: a steepest-descent step of the first-step length the gradient
                        * methods use (GSL: step_size along the normalised gradient), found by the line search */
            const double xn = nrm(x, n), len = (mp->step_size > 0 ? mp->step_size : 0.05) * (1 + xn) * 20;
            for (size_t i = 0; i < n; i++) p[i] = gnorm > 0 ? -g[i] * len / gnorm : 0;
        }
        { /* a nearly singular information matrix (saturated probabilities) gives an astronomically long
           * step: cap its length, the line search does the rest */
            const double pn = nrm(p, n), cap = 10 * (1 + nrm(x, n));
            if (!(pn <= cap)) for (size_t i = 0; i < n; i++) p[i] = isfinite(pn) && pn > 0 ? p[i] * cap / pn : -g[i] * cap / (gnorm > 0 ? gnorm : 1);
        }
        const int lrc = line_search(I, n, x, &f, g, p, 1.0, 0.9);
        if (I->nan_seen) { status = -1; break; }
        if (lrc) break;
        gnorm = nrm(g, n);
        if (mp->verbose) printf("%5i f()=%.10g |gradient|=%.3g\n", it, f, gnorm);
    }
    if (it >= mp->max_iterations) status = -1;
    *iters = it; *gnorm_out = gnorm;
    free(g); free(p); free(H);
    return status;
}

void apop_maximum_likelihood(apop_data *data, apop_model *dist) { /* apop_mle.m4.c:635-660 */
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    memset(&last_report, 0, sizeof last_report);
    if (!dist->log_likelihood && !dist->p) {
        fprintf(stderr, "apop_maximum_likelihood: the model has neither log_likelihood nor p. "
                        "This build has no CPU likelihoods: install the device entries with apop_b200_register.\n");
        dist->error = 'm'; last_report.status = -1;
        return;
    }
    apop_mle_settings defaults = apop_mle_settings_defaults();
    apop_mle_settings *mp = apop_settings_get_grp(dist, "apop_mle");
    if (!mp) mp = apop_mle_settings_add(dist, defaults);
    if (!dist->parameters) apop_model_clear(data, dist);
    if (!dist->data) dist->data = data;
    gsl_vector *packed = apop_data_pack(dist->parameters, NULL);
    if (!packed) { dist->error = 'p'; last_report.status = -1; return; }
    mle_info I = {.data = data, .model = dist, .n = packed->size, .best_ll = -INFINITY};
    I.bv = gsl_vector_alloc(I.n); I.gv = gsl_vector_alloc(I.n);
    double *x = malloc(sizeof(double) * I.n);
    for (size_t i = 0; i < I.n; i++) x[i] = mp->starting_pt ? mp->starting_pt[i] : 1.0; /* default start: all ones (:371) */
    const int have_score = apop_vtable_get("apop_score", apop_score_hash(dist)) != NULL;
    const char *method = mp->method[0] ? mp->method : (have_score ? "FR cg" : "NM simplex");
    int iters = 0, status;
    double gnorm = NAN;
    if (!strcasecmp(method, "NM simplex")) status = minimize_simplex(&I, mp, x, &iters);
    else if (!strcasecmp(method, "Newton")) status = minimize_newton(&I, mp, x, &iters, &gnorm);
    else status = minimize_gradient(&I, mp, x, !strcasecmp(method, "BFGS cg"), &iters, &gnorm);
    memcpy(I.bv->data, x, sizeof(double) * I.n);
    apop_data_unpack(I.bv, dist->parameters);
    const double ll = apop_log_likelihood(data, dist);
    /* info page: status, log likelihood, AIC, BIC (auxinfo / add_info_criteria, :378-409) */
    apop_data_free(dist->info);
    dist->info = apop_data_alloc(4, 0, 0);
    apop_data_add_page(NULL, dist->info, "<Info>");
    double nobs = data ? (data->matrix ? data->matrix->size1 : (data->vector ? data->vector->size : 0)) : 0;
    { /* row-sharded data: the information criteria count the rows of ALL ranks */
        double (*global_rows)(apop_data *) = (double (*)(apop_data *))dlsym(RTLD_DEFAULT, "apop_b200_global_rows");
        const double g = (global_rows && data) ? global_rows(data) : 0;
        if (g > 0) nobs = g;
    }
    dist->info->vector->data[0] = status;
    dist->info->vector->data[1] = ll;
    dist->info->vector->data[2] = 2 * (double)I.n - 2 * ll;
    dist->info->vector->data[3] = log(fmax(nobs, 1)) * (double)I.n - 2 * ll;
    clock_gettime(CLOCK_MONOTONIC, &t1);
    last_report = (apop_mle_report){.status = status, .iterations = iters, .f_evals = I.f_evals, .df_evals = I.df_evals,
                                    .ll = ll, .gradient_norm = gnorm, .seconds = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec)};
    gsl_vector_free(packed); gsl_vector_free(I.bv); gsl_vector_free(I.gv); free(x);
}
