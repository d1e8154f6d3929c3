/* oracle.c -- CPU restatement of Apophenia's likelihood/score path.
 * TEST INFRASTRUCTURE ONLY; see oracle.h for the scope, the GSL stand-ins and
 * the parity-pin statement.  Build: oracle/Makefile (gcc -O3 -fopenmp
 * -ffp-contract=off: the reference is plain C without fused multiply-adds).
 *
 * Every function names the reference lines (under /root/reference) it follows.
 * No reference source text is reproduced: bodies are written from the formulas
 * in SURVEY.md section 8(a).
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>
#include <time.h>

typedef long double ld;

static const ld SQRT2_L = 1.41421356237309504880168872420969808L;
static const ld SQRT2PI_L = 2.50662827463100050241576528481104525L;

/* ---- GSL stand-ins -------------------------------------------------- */

/* gsl_cdf_gaussian_P(x, 1): Phi(x).  GSL's cdf/gauss.c saturates to exactly 0
 * below GAUSS_XLOWER = -37.519 and to exactly 1 above GAUSS_XUPPER = 8.572; in
 * between it is a Cody rational approximation good to ~1e-15 relative, which
 * libm's erfc matches within tolerance. */
double oracle_gauss_cdf(double x) {
    if (x < -37.519) return 0.0;
    if (x > 8.572) return 1.0;
    return 0.5 * erfc(-x / M_SQRT2);
}
static ld gauss_cdf_l(ld x) {
    if (x < -37.519L) return 0.0L;
    if (x > 8.572L) return 1.0L;
    return 0.5L * erfcl(-x / SQRT2_L);
}
/* gsl_ran_gaussian_pdf(x, sigma) (randist/gauss.c): u = x/|sigma|;
 * p = (1/(sqrt(2 pi)|sigma|)) exp(-u u/2) */
double oracle_gauss_pdf(double x, double sigma) {
    double u = x / fabs(sigma);
    return (1 / (sqrt(2 * M_PI) * fabs(sigma))) * exp(-u * u / 2);
}
static ld gauss_pdf_l(ld x, ld sigma) {
    ld u = x / fabsl(sigma);
    return (1 / (SQRT2PI_L * fabsl(sigma))) * expl(-u * u / 2);
}

/* gsl_blas_dgemm(NoTrans, NoTrans, 1, A, B, 0, C) as the reference CBLAS does it
 * for row-major data (cblas/source_gemm_r.h): zero C, then for each (i,l) with
 * A[i][l] != 0 add A[i][l]*B[l][:] to C[i][:].  Single thread, like apop_dot
 * (apop_linear_algebra.m4.c:424-432). */
static void ref_dgemm(const double *A, size_t lda, size_t N, size_t K, const double *B, size_t ldb,
                      size_t J, double *C) {
    memset(C, 0, sizeof(double) * N * J);
    for (size_t i = 0; i < N; i++)
        for (size_t l = 0; l < K; l++) {
            const double t = A[i * lda + l];
            if (t != 0.0)
                for (size_t j = 0; j < J; j++) C[i * J + j] += t * B[l * ldb + j];
        }
}
/* gsl_blas_ddot: plain left-to-right double accumulation */
static double ref_ddot(const double *a, const double *b, size_t n) {
    double r = 0;
    for (size_t i = 0; i < n; i++) r += a[i] * b[i];
    return r;
}
/* gsl_stats_mean / gsl_stats_variance (statistics/mean_source.c, variance_source.c):
 * running recurrences in long double, variance scaled by n/(n-1). */
static double ref_stats_mean(const double *d, size_t n) {
    ld mean = 0;
    for (size_t i = 0; i < n; i++) mean += (d[i] - mean) / (i + 1);
    return mean;
}
static double ref_stats_variance(const double *d, size_t n) {
    const double mean = ref_stats_mean(d, n);
    ld var = 0;
    for (size_t i = 0; i < n; i++) {
        const ld delta = d[i] - mean;
        var += (delta * delta - var) / (i + 1);
    }
    return (double)var * ((double)n / (double)(n - 1));
}

/* apop_vector_sum (apop_stats.m4.c:15-21): serial, long double accumulator */
static ld ref_vector_sum(const double *v, size_t n) {
    ld out = 0;
    for (size_t i = 0; i < n; i++) out += v[i];
    return out;
}

/* ---- binary probit --------------------------------------------------- */

/* biprobit_ll_row, model/apop_probit.c:83-88 */
static double probit_row_faithful(double xb, double y) {
    ld n = oracle_gauss_cdf(-xb);
    n = n ? n : 1e-10;
    n = n < 1 ? n : 1 - 1e-10;
    return y ? log((double)(1 - n)) : log((double)n);
}
static ld probit_row_exact(ld xb, double y) {
    ld n = gauss_cdf_l(-xb);
    const double nd = (double)n; /* the reference clamps on the rounded double */
    if (nd == 0.0) n = 1e-10;
    if (!(nd < 1.0)) n = 1 - 1e-10;
    return y ? logl(1 - n) : logl(n);
}
static ld dot_l(const double *x, const double *b, size_t stride_b, size_t k) {
    ld s = 0;
    for (size_t j = 0; j < k; j++) s += (ld)x[j] * (ld)b[j * stride_b];
    return s;
}

/* biprobit_log_likelihood, model/apop_probit.c:91-98: apop_dot, apop_map_sum(.fn_r) */
void oracle_probit_ll(const double *X, size_t tda, const double *y, size_t N, size_t k,
                      const double *beta, int mode, long double *ll_out) {
    ld total = 0;
    if (mode == ORACLE_FAITHFUL) {
        double *xb = malloc(sizeof(double) * (N ? N : 1));
        double *mapped = malloc(sizeof(double) * (N ? N : 1));
        ref_dgemm(X, tda, N, k, beta, 1, 1, xb);
        for (size_t i = 0; i < N; i++) mapped[i] = probit_row_faithful(xb[i], y[i]);
        total = (double)ref_vector_sum(mapped, N); /* apop_map_sum returns double */
        free(xb);
        free(mapped);
    } else {
        for (size_t i = 0; i < N; i++) total += probit_row_exact(dot_l(X + i * tda, beta, 1, k), y[i]);
    }
    *ll_out = total;
}

/* probit_dlog_likelihood, model/apop_probit.c:130-154 */
void oracle_probit_score(const double *X, size_t tda, const double *y, size_t N, size_t k,
                         const double *beta, int mode, long double *g_out, long double *abs_out) {
    ld *ab = calloc(k ? k : 1, sizeof(ld));
    if (mode == ORACLE_FAITHFUL) {
        double *xb = malloc(sizeof(double) * (N ? N : 1));
        double *g = calloc(k ? k : 1, sizeof(double));
        ref_dgemm(X, tda, N, k, beta, 1, 1, xb);
        for (size_t i = 0; i < N; i++) {
            ld betax = xb[i];
            ld cdf = oracle_gauss_cdf((double)-betax);
            cdf = cdf ? cdf : 1e-10;
            cdf = cdf < 1 ? cdf : 1 - 1e-10;
            ld deriv = y[i] ? oracle_gauss_pdf((double)-betax, 1) / (1 - cdf)
                            : -oracle_gauss_pdf((double)-betax, 1) / cdf;
            for (size_t j = 0; j < k; j++) {
                g[j] += X[i * tda + j] * deriv; /* long double product, one rounding per += */
                ab[j] += fabsl(X[i * tda + j] * deriv);
            }
        }
        for (size_t j = 0; j < k; j++) g_out[j] = g[j];
        free(xb);
        free(g);
    } else {
        for (size_t j = 0; j < k; j++) g_out[j] = 0;
        for (size_t i = 0; i < N; i++) {
            const ld betax = dot_l(X + i * tda, beta, 1, k);
            ld cdf = gauss_cdf_l(-betax);
            const double cd = (double)cdf;
            if (cd == 0.0) cdf = 1e-10;
            if (!(cd < 1.0)) cdf = 1 - 1e-10;
            const ld pdf = gauss_pdf_l(-betax, 1);
            const ld deriv = y[i] ? pdf / (1 - cdf) : -pdf / cdf;
            for (size_t j = 0; j < k; j++) {
                g_out[j] += X[i * tda + j] * deriv;
                ab[j] += fabsl(X[i * tda + j] * deriv);
            }
        }
    }
    if (abs_out)
        for (size_t j = 0; j < k; j++) abs_out[j] = ab[j];
    free(ab);
}

/* probit_prep start point, model/apop_probit.c:65-80 */
void oracle_probit_start(const double *X, size_t tda, size_t N, size_t k, double *beta0) {
    for (size_t j = 0; j < k; j++) {
        ld logtotal = 0;
        for (size_t i = 0; i < N; i++) {
            const double val = X[i * tda + j];
            logtotal += val ? logl(fabs(val)) : 0;
        }
        logtotal /= N;
        beta0[j] = expl(logtotal);
    }
}

/* ---- multinomial logit ------------------------------------------------ */

/* find_index, model/apop_probit.c:206-210: position of y among the category
 * values, scanning at most `max` steps */
static size_t find_index(double in, const double *m, size_t max) {
    size_t i = 0;
    while (in != m[i] && i < max) i++;
    return i;
}

/* one_logit_row, model/apop_probit.c:212-229 */
static double logit_row_faithful(double *row, size_t C1, double yv, const double *cats) {
    const size_t index = find_index(yv, cats, C1);
    const double num = (index == 0) ? 0 : row[index - 1];
    double max = row[0];
    for (size_t c = 1; c < C1; c++)
        if (row[c] > max) max = row[c];
    for (size_t c = 0; c < C1; c++) row[c] = exp(row[c] - max);
    const ld expmax = expl(-max);
    return num - (max + (isfinite(expmax) ? logl(ref_vector_sum(row, C1) + expmax) : -max));
}
static ld logit_row_exact(const ld *row, size_t C1, double yv, const double *cats) {
    const size_t index = find_index(yv, cats, C1);
    const ld num = (index == 0) ? 0 : row[index - 1];
    ld max = row[0];
    for (size_t c = 1; c < C1; c++)
        if (row[c] > max) max = row[c];
    ld sum = 0;
    for (size_t c = 0; c < C1; c++) sum += expl(row[c] - max);
    const ld expmax = expl(-max);
    return num - (max + (isfinite(expmax) ? logl(sum + expmax) : -max));
}

/* multilogit_log_likelihood, model/apop_probit.c:231-242 */
void oracle_logit_ll(const double *X, size_t tda, const double *y, size_t N, size_t k,
                     const double *B, size_t C, const double *cats, int mode, long double *ll_out) {
    const size_t C1 = C - 1;
    ld total = 0;
    if (mode == ORACLE_FAITHFUL) {
        double *xb = malloc(sizeof(double) * (N ? N : 1) * C1);
        double *mapped = malloc(sizeof(double) * (N ? N : 1));
        ref_dgemm(X, tda, N, k, B, C1, C1, xb);
        for (size_t i = 0; i < N; i++) mapped[i] = logit_row_faithful(xb + i * C1, C1, y[i], cats);
        total = (double)ref_vector_sum(mapped, N);
        free(xb);
        free(mapped);
    } else {
        ld *row = malloc(sizeof(ld) * C1);
        for (size_t i = 0; i < N; i++) {
            for (size_t c = 0; c < C1; c++) row[c] = dot_l(X + i * tda, B + c, C1, k);
            total += logit_row_exact(row, C1, y[i], cats);
        }
        free(row);
    }
    *ll_out = total;
}

/* Analytic score (reference: commented out, model/apop_probit.c:244-289):
 * dLL/dB[j][c] = sum_i x_ij (1{idx_i == c+1} - p_ic),
 * p_ic = exp(xb_c)/(1 + sum_c' exp(xb_c')).  Output index j*(C-1)+c. */
void oracle_logit_score(const double *X, size_t tda, const double *y, size_t N, size_t k,
                        const double *B, size_t C, const double *cats, int mode,
                        long double *g_out, long double *abs_out) {
    const size_t C1 = C - 1, P = k * C1;
    ld *ab = calloc(P ? P : 1, sizeof(ld));
    ld *row = malloc(sizeof(ld) * C1), *r = malloc(sizeof(ld) * C1);
    double *gd = calloc(P ? P : 1, sizeof(double));
    for (size_t q = 0; q < P; q++) g_out[q] = 0;
    for (size_t i = 0; i < N; i++) {
        const double *x = X + i * tda;
        if (mode == ORACLE_FAITHFUL)
            for (size_t c = 0; c < C1; c++) {
                double s = 0;
                for (size_t j = 0; j < k; j++) s += x[j] * B[j * C1 + c];
                row[c] = s;
            }
        else
            for (size_t c = 0; c < C1; c++) row[c] = dot_l(x, B + c, C1, k);
        const size_t index = find_index(y[i], cats, C1);
        ld max = 0; /* the numeraire's x.beta_0 = 0 takes part in the max here */
        for (size_t c = 0; c < C1; c++)
            if (row[c] > max) max = row[c];
        ld den = expl(-max);
        for (size_t c = 0; c < C1; c++) den += expl(row[c] - max);
        for (size_t c = 0; c < C1; c++) {
            const ld p = expl(row[c] - max) / den;
            r[c] = ((index == c + 1) ? 1.0L : 0.0L) - p;
            if (mode == ORACLE_FAITHFUL) r[c] = (double)r[c];
        }
        for (size_t j = 0; j < k; j++)
            for (size_t c = 0; c < C1; c++) {
                if (mode == ORACLE_FAITHFUL) gd[j * C1 + c] += x[j] * (double)r[c];
                else g_out[j * C1 + c] += x[j] * r[c];
                ab[j * C1 + c] += fabsl(x[j] * r[c]);
            }
    }
    if (mode == ORACLE_FAITHFUL)
        for (size_t q = 0; q < P; q++) g_out[q] = gd[q];
    if (abs_out)
        for (size_t q = 0; q < P; q++) abs_out[q] = ab[q];
    free(ab); free(row); free(r); free(gd);
}

/* ---- apop_map_sum over cells ----------------------------------------- */

/* the callbacks: normal apply_me / apply_me2 (model/apop_normal.c:38-40),
 * poisson apply_me (model/apop_poisson.c:14-19) */
static double map_cell(int fn, double x, double p0) {
    switch (fn) {
    case 0: return x;
    case 1: return x - p0;
    case 2: { double d = x - p0; return d * d; }
    case 3:
        if (x < 0) return -INFINITY;
        if ((x - (int)x) > 1e-4) return -INFINITY;
        return x == 0 ? 0 : p0 * x - lgamma(x + 1);
    case 4: return log(x);
    case 5: return exp(x);
    }
    return NAN;
}
static ld map_cell_l(int fn, double x, double p0) {
    switch (fn) {
    case 0: return x;
    case 1: return (ld)x - p0;
    case 2: { ld d = (ld)x - p0; return d * d; }
    case 3:
        if (x < 0) return -INFINITY;
        if ((x - (int)x) > 1e-4) return -INFINITY;
        return x == 0 ? 0 : (ld)p0 * x - lgammal((ld)x + 1);
    case 4: return logl(x);
    case 5: return expl(x);
    }
    return NAN;
}

/* apop_map_sum (apop_mapply.m4.c:485-517): map every cell into a fresh copy,
 * then apop_sum(vector) + apop_matrix_sum(matrix), each serial long double
 * (apop_stats.m4.c:15-21, :310-317); the result is returned as double. */
void oracle_map_sum(const double *v, size_t vsize, const double *M, size_t tda, size_t m1,
                    size_t m2, int fn, double p0, int mode, long double *out) {
    ld sv = 0, sm = 0;
    if (mode == ORACLE_FAITHFUL) {
        for (size_t i = 0; i < vsize; i++) sv += map_cell(fn, v[i], p0);
        for (size_t i = 0; i < m1; i++)
            for (size_t j = 0; j < m2; j++) sm += map_cell(fn, M[i * tda + j], p0);
        *out = (double)(sv + sm);
    } else {
        for (size_t i = 0; i < vsize; i++) sv += map_cell_l(fn, v[i], p0);
        for (size_t i = 0; i < m1; i++)
            for (size_t j = 0; j < m2; j++) sm += map_cell_l(fn, M[i * tda + j], p0);
        *out = sv + sm;
    }
}

/* ---- iid Poisson ------------------------------------------------------ */

/* poisson_log_likelihood, model/apop_poisson.c:21-28 */
void oracle_poisson_ll(const double *v, size_t vsize, const double *M, size_t tda, size_t m1,
                       size_t m2, double lambda, int mode, long double *ll_out) {
    const int tsize = (int)(vsize + m1 * m2);
    ld s;
    if (mode == ORACLE_FAITHFUL) {
        const double ln_l = log(lambda);
        oracle_map_sum(v, vsize, M, tda, m1, m2, 3, ln_l, mode, &s);
        const double ll = (double)s;
        *ll_out = ll - tsize * lambda;
    } else {
        /* the map callback sees ln(lambda) as a double in the reference; keep the
         * extra digits here by summing x and lgamma separately */
        const ld ln_l = logl((ld)lambda);
        ld sx = 0, sg = 0;
        int bad = 0;
        for (size_t i = 0; i < vsize + m1 * m2; i++) {
            const double x = i < vsize ? v[i] : M[((i - vsize) / m2) * tda + (i - vsize) % m2];
            if (x < 0 || (x - (int)x) > 1e-4) bad = 1;
            if (x != 0) { sx += x; sg += lgammal((ld)x + 1); }
        }
        *ll_out = bad ? -INFINITY : ln_l * sx - sg - (ld)(vsize + m1 * m2) * lambda;
    }
}
/* poisson_dlog_likelihood, model/apop_poisson.c:66-73: matrix cells only in the sum */
void oracle_poisson_score(const double *v, size_t vsize, const double *M, size_t tda, size_t m1,
                          size_t m2, double lambda, int mode, long double *g_out) {
    (void)v;
    const int tsize = (int)(vsize + m1 * m2);
    ld sm = 0;
    for (size_t i = 0; i < m1; i++)
        for (size_t j = 0; j < m2; j++) sm += M[i * tda + j];
    if (mode == ORACLE_FAITHFUL) {
        const double d_a = (double)(sm / lambda - tsize);
        g_out[0] = d_a;
    } else
        g_out[0] = sm / (ld)lambda - (ld)tsize;
}

/* ---- iid Normal ------------------------------------------------------- */

/* normal_log_likelihood, model/apop_normal.c:42-50 */
void oracle_normal_ll(const double *v, size_t vsize, const double *M, size_t tda, size_t m1,
                      size_t m2, double mu, double sd, int mode, long double *ll_out) {
    const int tsize = (int)(vsize + m1 * m2);
    ld s;
    oracle_map_sum(v, vsize, M, tda, m1, m2, 2, mu, mode, &s);
    if (mode == ORACLE_FAITHFUL) {
        ld ll = -(double)s / (2 * (sd * sd));
        ll -= tsize * ((1.14472988584940017414 + M_LN2) / 2 + log(sd)); /* M_LNPI */
        *ll_out = ll;
    } else {
        const ld lnpi = 1.14472988584940017414342735135305871L, ln2 = 0.693147180559945309417232121458176568L;
        *ll_out = -s / (2 * ((ld)sd * sd)) - (ld)tsize * ((lnpi + ln2) / 2 + logl(sd));
    }
}
/* normal_dlog_likelihood, model/apop_normal.c:108-118 */
void oracle_normal_score(const double *v, size_t vsize, const double *M, size_t tda, size_t m1,
                         size_t m2, double mu, double sd, int mode, long double *g_out) {
    const int tsize = (int)(vsize + m1 * m2);
    ld dll, sll;
    oracle_map_sum(v, vsize, M, tda, m1, m2, 1, mu, mode, &dll);
    oracle_map_sum(v, vsize, M, tda, m1, m2, 2, mu, mode, &sll);
    if (mode == ORACLE_FAITHFUL) {
        g_out[0] = (double)dll / (sd * sd);
        g_out[1] = (double)sll / (sd * sd * sd) - tsize / sd;
    } else {
        g_out[0] = dll / ((ld)sd * sd);
        g_out[1] = sll / ((ld)sd * sd * sd) - (ld)tsize / sd;
    }
}

/* ---- OLS -------------------------------------------------------------- */

/* ols_log_likelihood, model/apop_ols.c:129-172 (prepped data, sigma from the
 * sample variance of the errors, P(x) = 1) */
void oracle_ols_ll(const double *X, size_t tda, const double *y, const double *w, size_t N,
                   size_t k, const double *beta, int mode, long double *ll_out) {
    ld ll = 0;
    if (mode == ORACLE_FAITHFUL) {
        double *e = malloc(sizeof(double) * (N ? N : 1));
        for (size_t i = 0; i < N; i++) {
            const double expected = ref_ddot(beta, X + i * tda, k);
            const ld actual = y[i];
            e[i] = expected - actual;
        }
        const ld sigma = sqrt(ref_stats_variance(e, N));
        for (size_t i = 0; i < N; i++) {
            const ld weight = w ? w[i] : 1;
            ll += logl(oracle_gauss_pdf(e[i], (double)sigma) * weight * 1.0);
        }
        free(e);
    } else {
        ld *e = malloc(sizeof(ld) * (N ? N : 1));
        ld mean = 0;
        for (size_t i = 0; i < N; i++) { e[i] = dot_l(X + i * tda, beta, 1, k) - y[i]; mean += e[i]; }
        mean /= N;
        ld ss = 0;
        for (size_t i = 0; i < N; i++) ss += (e[i] - mean) * (e[i] - mean);
        const ld sigma = sqrtl(ss / (N - 1));
        const ld lnorm = logl(SQRT2PI_L * sigma);
        /* log(pdf*w) written without the exp/log round trip */
        for (size_t i = 0; i < N; i++) {
            const ld u = e[i] / sigma;
            ll += -u * u / 2 - lnorm + (w ? logl(w[i]) : 0);
        }
        free(e);
    }
    *ll_out = ll;
}

/* ols_score, model/apop_ols.c:175-205: g_j = sum_i w_i X_ij e_i / sigma^2 with
 * e = x.beta - y.  (Sign as coded by the reference: it is the Normal model's
 * d/dmu at x = e, mu = 0, which is the negative of dLL/dbeta) */
void oracle_ols_score(const double *X, size_t tda, const double *y, const double *w, size_t N,
                      size_t k, const double *beta, int mode, long double *g_out,
                      long double *abs_out) {
    ld *ab = calloc(k ? k : 1, sizeof(ld));
    if (mode == ORACLE_FAITHFUL) {
        double *e = malloc(sizeof(double) * (N ? N : 1));
        double *g = calloc(k ? k : 1, sizeof(double));
        for (size_t i = 0; i < N; i++) {
            const double expected = ref_ddot(beta, X + i * tda, k);
            const ld actual = y[i];
            e[i] = expected - actual;
        }
        const ld sigma = sqrt(ref_stats_variance(e, N));
        const double sd = (double)sigma; /* apop_model_set_parameters stores a double */
        for (size_t i = 0; i < N; i++) {
            /* normal_dlog_likelihood on the 1x1 data set {e_i} with mu = 0 */
            const double ns0 = (double)(ld)(e[i] - 0.0) / (sd * sd);
            const ld weight = w ? w[i] : 1;
            for (size_t j = 0; j < k; j++) {
                g[j] += weight * X[i * tda + j] * ns0;
                ab[j] += fabsl(weight * X[i * tda + j] * ns0);
            }
        }
        for (size_t j = 0; j < k; j++) g_out[j] = g[j];
        free(e); free(g);
    } else {
        ld *e = malloc(sizeof(ld) * (N ? N : 1));
        ld mean = 0;
        for (size_t i = 0; i < N; i++) { e[i] = dot_l(X + i * tda, beta, 1, k) - y[i]; mean += e[i]; }
        mean /= N;
        ld ss = 0;
        for (size_t i = 0; i < N; i++) ss += (e[i] - mean) * (e[i] - mean);
        const ld var = ss / (N - 1);
        for (size_t j = 0; j < k; j++) g_out[j] = 0;
        for (size_t i = 0; i < N; i++) {
            const ld s = (w ? (ld)w[i] : 1) * e[i] / var;
            for (size_t j = 0; j < k; j++) {
                g_out[j] += X[i * tda + j] * s;
                ab[j] += fabsl(X[i * tda + j] * s);
            }
        }
        free(e);
    }
    if (abs_out)
        for (size_t j = 0; j < k; j++) abs_out[j] = ab[j];
    free(ab);
}

/* X'X and X'y: apop_dot(set, set, .form1='t') -> dgemm(Trans, NoTrans) and
 * dgemv(Trans) (model/apop_ols.c:333-334, apop_linear_algebra.m4.c:431,328) */
void oracle_ols_gram(const double *X, size_t tda, const double *y, size_t N, size_t k, int mode,
                     long double *xtx, long double *xty) {
    if (mode == ORACLE_FAITHFUL) {
        double *A = calloc(k * k > 0 ? k * k : 1, sizeof(double)), *b = calloc(k ? k : 1, sizeof(double));
        for (size_t i = 0; i < N; i++)
            for (size_t a = 0; a < k; a++) {
                const double t = X[i * tda + a];
                if (t != 0.0)
                    for (size_t c = 0; c < k; c++) A[a * k + c] += t * X[i * tda + c];
                b[a] += t * y[i];
            }
        for (size_t q = 0; q < k * k; q++) xtx[q] = A[q];
        for (size_t q = 0; q < k; q++) xty[q] = b[q];
        free(A); free(b);
    } else {
        for (size_t q = 0; q < k * k; q++) xtx[q] = 0;
        for (size_t q = 0; q < k; q++) xty[q] = 0;
        for (size_t i = 0; i < N; i++)
            for (size_t a = 0; a < k; a++) {
                const ld t = X[i * tda + a];
                for (size_t c = 0; c < k; c++) xtx[a * k + c] += t * X[i * tda + c];
                xty[a] += t * y[i];
            }
    }
}

/* ---- Poisson regression (new model) ----------------------- */
void oracle_poisson_reg_ll(const double *X, size_t tda, const double *y, size_t N, size_t k,
                           const double *beta, int mode, long double *ll_out) {
    ld total = 0;
    for (size_t i = 0; i < N; i++) {
        if (mode == ORACLE_FAITHFUL) {
            const double xb = ref_ddot(X + i * tda, beta, k);
            total += (double)(y[i] * xb - exp(xb) - lgamma(y[i] + 1));
        } else {
            const ld xb = dot_l(X + i * tda, beta, 1, k);
            total += y[i] * xb - expl(xb) - lgammal((ld)y[i] + 1);
        }
    }
    *ll_out = total;
}
void oracle_poisson_reg_score(const double *X, size_t tda, const double *y, size_t N, size_t k,
                              const double *beta, int mode, long double *g_out,
                              long double *abs_out) {
    ld *ab = calloc(k ? k : 1, sizeof(ld));
    double *gd = calloc(k ? k : 1, sizeof(double));
    for (size_t j = 0; j < k; j++) g_out[j] = 0;
    for (size_t i = 0; i < N; i++) {
        ld d;
        if (mode == ORACLE_FAITHFUL) d = (double)(y[i] - exp(ref_ddot(X + i * tda, beta, k)));
        else d = y[i] - expl(dot_l(X + i * tda, beta, 1, k));
        for (size_t j = 0; j < k; j++) {
            if (mode == ORACLE_FAITHFUL) gd[j] += X[i * tda + j] * (double)d;
            else g_out[j] += X[i * tda + j] * d;
            ab[j] += fabsl(X[i * tda + j] * d);
        }
    }
    if (mode == ORACLE_FAITHFUL)
        for (size_t j = 0; j < k; j++) g_out[j] = gd[j];
    if (abs_out)
        for (size_t j = 0; j < k; j++) abs_out[j] = ab[j];
    free(ab); free(gd);
}

/* ---- observed information (no reference counterpart; exact arithmetic) - */
void oracle_information(int family, const double *X, size_t tda, const double *y, size_t N,
                        size_t k, const double *beta, long double *H) {
    for (size_t q = 0; q < k * k; q++) H[q] = 0;
    for (size_t i = 0; i < N; i++) {
        const double *x = X + i * tda;
        const ld xb = dot_l(x, beta, 1, k);
        ld wgt;
        if (family == 0) { /* probit: lambda (lambda + xb), lambda = q phi(xb)/Phi(q xb) */
            ld cdf = gauss_cdf_l(-xb);
            const double cd = (double)cdf;
            if (cd == 0.0) cdf = 1e-10;
            if (!(cd < 1.0)) cdf = 1 - 1e-10;
            const ld pdf = gauss_pdf_l(-xb, 1);
            const ld lam = y[i] ? pdf / (1 - cdf) : -pdf / cdf;
            wgt = lam * (lam + xb);
        } else if (family == 1) { /* binary logit: p (1-p) */
            const ld p = 1 / (1 + expl(-xb));
            wgt = p * (1 - p);
        } else if (family == 5) { /* poisson regression: exp(xb) */
            wgt = expl(xb);
        } else
            wgt = NAN;
        for (size_t a = 0; a < k; a++)
            for (size_t c = 0; c < k; c++) H[a * k + c] += wgt * x[a] * x[c];
    }
}


/* ---- more iid distributions on the same apop_map_sum shape ---- */
static ld digamma_l(ld x) { /* gsl_sf_psi stand-in: recurrence up to x >= 12, then the asymptotic series */
    ld r = 0;
    while (x < 12) { r -= 1 / x; x += 1; }
    const ld f = 1 / (x * x);
    return r + logl(x) - 0.5L / x - f * (1.0L / 12 - f * (1.0L / 120 - f * (1.0L / 252 - f * (1.0L / 240 - f * (1.0L / 132 - f * (691.0L / 32760 - f / 12))))));
}
double oracle_digamma(double x) { return (double)digamma_l(x); }

/* exponential_log_likelihood / _dlog_likelihood, model/apop_exponential.c:33-52:
 * LL = -(sum x)/mu - tsize ln mu;  dLL/dmu = (sum x)/mu^2 - tsize/mu */
void oracle_exponential(const double *v, size_t vsize, const double *M, size_t tda, size_t m1, size_t m2,
                        double mu, int mode, long double *ll_out, long double *g_out) {
    const int tsize = (int)(vsize + m1 * m2);
    ld sv = 0, sm = 0;
    for (size_t i = 0; i < vsize; i++) sv += v[i];
    for (size_t i = 0; i < m1; i++) for (size_t j = 0; j < m2; j++) sm += M[i * tda + j];
    if (mode == ORACLE_FAITHFUL) {
        double ll = -(double)(sm + sv) / mu;   /* the two long double sums are added, then rounded */
        ll -= tsize * log(mu);
        double dl = (double)(sm + sv);
        dl /= mu * mu;
        dl -= tsize / mu;
        *ll_out = ll; g_out[0] = dl;
    } else {
        *ll_out = -(sm + sv) / (ld)mu - (ld)tsize * logl(mu);
        g_out[0] = (sm + sv) / ((ld)mu * mu) - (ld)tsize / mu;
    }
}

/* gamma_log_likelihood / gamma_dlog_likelihood, model/apop_gamma.c:31-62:
 * LL = sum_{x != 0} [(a-1) ln x - x/b - (lnGamma(a) + a ln b)];
 * dLL/da = sum (ln x - (psi(a) + ln b)) over ALL cells,  dLL/db = sum (x/b^2 - a/b) */
void oracle_gamma(const double *v, size_t vsize, const double *M, size_t tda, size_t m1, size_t m2,
                  double a, double b, int mode, long double *ll_out, long double *g_out) {
    const size_t n = vsize + m1 * m2;
    ld ll = 0, g0 = 0, g1 = 0;
    if (mode == ORACLE_FAITHFUL) {
        const double c = lgamma(a) + a * log(b), psi_ln = (double)digamma_l(a) + log(b), ab = a / b;
        ld llv = 0, llm = 0, g0v = 0, g0m = 0, g1v = 0, g1m = 0;
        for (size_t i = 0; i < n; i++) {
            const double x = i < vsize ? v[i] : M[((i - vsize) / m2) * tda + (i - vsize) % m2];
            const double t = x ? ((a - 1) * log(x) - x / b - c) : 0;
            const double t0 = log(x) - psi_ln, t1 = x / (b * b) - ab;
            if (i < vsize) { llv += t; g0v += t0; g1v += t1; } else { llm += t; g0m += t0; g1m += t1; }
        }
        ll = (double)(llv + llm); g0 = (double)(g0v + g0m); g1 = (double)(g1v + g1m);
    } else {
        const ld c = lgammal(a) + (ld)a * logl(b), psi_ln = digamma_l(a) + logl(b);
        for (size_t i = 0; i < n; i++) {
            const double x = i < vsize ? v[i] : M[((i - vsize) / m2) * tda + (i - vsize) % m2];
            if (x) ll += ((ld)a - 1) * logl(x) - x / (ld)b - c;
            g0 += logl(x) - psi_ln;
            g1 += x / ((ld)b * b) - (ld)a / b;
        }
    }
    *ll_out = ll; g_out[0] = g0; g_out[1] = g1;
}

/* lognormal_log_likelihood / _dlog_likelihood, model/apop_normal.c:166-237:
 * LL = -sum (ln x - mu)^2/(2 sd^2) - sum ln x - tsize ((ln pi + ln 2)/2 + ln sd) */
void oracle_lognormal(const double *v, size_t vsize, const double *M, size_t tda, size_t m1, size_t m2,
                      double mu, double sd, int mode, long double *ll_out, long double *g_out) {
    const size_t n = vsize + m1 * m2;
    const int tsize = (int)n;
    ld s1v = 0, s1m = 0, s2v = 0, s2m = 0;
    for (size_t i = 0; i < n; i++) {
        const double x = i < vsize ? v[i] : M[((i - vsize) / m2) * tda + (i - vsize) % m2];
        if (mode == ORACLE_FAITHFUL) {
            const double l = log(x), d = l - mu;
            if (i < vsize) { s1v += l; s2v += d * d; } else { s1m += l; s2m += d * d; }
        } else {
            const ld l = logl(x), d = l - mu;
            s1v += l; s2v += d * d;
        }
    }
    if (mode == ORACLE_FAITHFUL) {
        const double S1 = (double)(s1v + s1m), S2 = (double)(s2v + s2m);
        ld ll = -S2;
        ll /= (2 * (sd * sd));
        ll -= S1;
        ll -= tsize * ((1.14472988584940017414 + M_LN2) / 2 + log(sd));
        *ll_out = ll;
        g_out[0] = (S1 - mu * tsize) / (sd * sd);
        g_out[1] = S2 / (sd * sd * sd) - tsize / sd;
    } else {
        const ld lnpi = 1.14472988584940017414342735135305871L, ln2 = 0.693147180559945309417232121458176568L;
        *ll_out = -(s2v) / (2 * (ld)sd * sd) - s1v - (ld)tsize * ((lnpi + ln2) / 2 + logl(sd));
        g_out[0] = (s1v - (ld)mu * tsize) / ((ld)sd * sd);
        g_out[1] = s2v / ((ld)sd * sd * sd) - (ld)tsize / sd;
    }
}

/* ---- Bernoulli, Beta, Zipf, Dirichlet, Student t, Yule ----
 * Each follows the reference's own per-cell callback and apop_map_sum order (vector cells, then
 * matrix cells, each summed serially in long double, the two sums added: apop_mapply.m4.c:485-517).
 * FAITHFUL keeps the reference's double callbacks; EXACT evaluates every cell in long double. */
#define CELL(i) ((i) < vsize ? v[i] : M[(((i) - vsize) / m2) * tda + ((i) - vsize) % m2])

/* bernie_ll, model/apop_bernoulli.c:13-23: x ? log(p) : log(1-p).  The closed-form score
 * n1/p - n0/(1-p) is not in the reference (numeric gradient there). */
void oracle_bernoulli(const double *v, size_t vsize, const double *M, size_t tda, size_t m1, size_t m2,
                      double p, int mode, long double *ll_out, long double *g_out) {
    const size_t n = vsize + m1 * m2;
    ld sv = 0, sm = 0, n1 = 0;
    for (size_t i = 0; i < n; i++) {
        const double x = CELL(i);
        const ld t = (mode == ORACLE_FAITHFUL) ? (ld)(x ? log(p) : log(1 - p)) : (x ? logl(p) : log1pl(-(ld)p));
        if (i < vsize) sv += t; else sm += t;
        n1 += (x != 0);
    }
    *ll_out = (mode == ORACLE_FAITHFUL) ? (ld)(double)(sv + sm) : sv + sm;
    g_out[0] = n1 / (ld)p - ((ld)n - n1) / (1 - (ld)p);
}

/* betamap / beta_log_likelihood / beta_dlog_likelihood, model/apop_beta.c:47-75.
 * gsl_sf_lnbeta(a,b) -> lgamma(a) + lgamma(b) - lgamma(a+b) */
void oracle_beta(const double *v, size_t vsize, const double *M, size_t tda, size_t m1, size_t m2,
                 double a, double b, int mode, long double *ll_out, long double *g_out) {
    const size_t n = vsize + m1 * m2;
    ld sv = 0, sm = 0, lv = 0, lm = 0, l1v = 0, l1m = 0;
    for (size_t i = 0; i < n; i++) {
        const double x = CELL(i);
        ld t, lx, l1x;
        if (mode == ORACLE_FAITHFUL) {
            t = (x < 0 || x > 1) ? 0 : (a - 1) * log(x) + (b - 1) * log(1 - x);
            lx = log(x); l1x = log(1 - x);
        } else {
            t = (x < 0 || x > 1) ? 0 : ((ld)a - 1) * logl(x) + ((ld)b - 1) * log1pl(-(ld)x);
            lx = logl(x); l1x = log1pl(-(ld)x);
        }
        if (i < vsize) { sv += t; lv += lx; l1v += l1x; } else { sm += t; lm += lx; l1m += l1x; }
    }
    if (mode == ORACLE_FAITHFUL) {
        const double lnbeta = lgamma(a) + lgamma(b) - lgamma(a + b);
        *ll_out = (double)(sv + sm) - lnbeta * (double)n;
        g_out[0] = (double)(lv + lm) + ((double)digamma_l(a + b) - (double)digamma_l(a)) * (double)n;
        g_out[1] = (double)(l1v + l1m) + ((double)digamma_l(a + b) - (double)digamma_l(b)) * (double)n;
    } else {
        const ld lnbeta = lgammal(a) + lgammal(b) - lgammal((ld)a + b);
        *ll_out = (sv + sm) - lnbeta * (ld)n;
        g_out[0] = (lv + lm) + (digamma_l((ld)a + b) - digamma_l(a)) * (ld)n;
        g_out[1] = (l1v + l1m) + (digamma_l((ld)a + b) - digamma_l(b)) * (ld)n;
    }
}

/* gsl_sf_zeta stand-in for s > 1: direct sum to N, integral tail and five Euler-Maclaurin terms */
static ld zeta_l(ld s) {
    const int N = 40;
    ld sum = 0;
    for (int k = N - 1; k >= 1; k--) sum += powl(k, -s);
    const ld n = N;
    sum += powl(n, 1 - s) / (s - 1) + 0.5L * powl(n, -s);
    const ld b[5] = {1.0L / 6, -1.0L / 30, 1.0L / 42, -1.0L / 30, 5.0L / 66};
    ld rising = s, fact = 2, npow = powl(n, -s - 1);
    for (int j = 1; j <= 5; j++) {
        sum += b[j - 1] / fact * rising * npow;
        rising *= (s + 2 * j - 1) * (s + 2 * j);
        fact *= (2 * j + 1) * (2 * j + 2);
        npow /= n * n;
    }
    return sum;
}
double oracle_zeta(double s) { return (double)zeta_l(s); }

/* zipf_log_likelihood, model/apop_zipf.c:31-40: like = -map_sum(log) * bb; like -= log(zeta(bb)) * tsize */
void oracle_zipf(const double *v, size_t vsize, const double *M, size_t tda, size_t m1, size_t m2,
                 double bb, int mode, long double *ll_out, long double *g_out) {
    const size_t n = vsize + m1 * m2;
    ld sv = 0, sm = 0;
    for (size_t i = 0; i < n; i++) {
        const double x = CELL(i);
        const ld t = (mode == ORACLE_FAITHFUL) ? (ld)log(x) : logl(x);
        if (i < vsize) sv += t; else sm += t;
    }
    if (mode == ORACLE_FAITHFUL) {
        double like = -(double)(sv + sm) * bb;
        like -= log((double)zeta_l(bb)) * (double)n;
        *ll_out = like;
    } else
        *ll_out = -(sv + sm) * (ld)bb - logl(zeta_l(bb)) * (ld)n;
    g_out[0] = NAN;   /* no score in the reference */
}

/* dirichletlnmap / dirichlet_log_likelihood / dirichlet_dlog_likelihood, model/apop_dirichlet.c:13-38;
 * gsl_ran_dirichlet_lnpdf(K, alpha, theta) = sum (alpha_i - 1) ln theta_i + lnGamma(sum alpha) - sum lnGamma(alpha_i).
 * Matrix rows only (part = 'r'); g has m2 entries. */
void oracle_dirichlet(const double *M, size_t tda, size_t m1, size_t m2, const double *alpha, int mode,
                      long double *ll_out, long double *g_out) {
    ld total = 0, asum = 0;
    for (size_t j = 0; j < m2; j++) asum += alpha[j];
    for (size_t i = 0; i < m1; i++) {
        if (mode == ORACLE_FAITHFUL) {
            double lp = 0, sa = 0;
            for (size_t j = 0; j < m2; j++) lp += (alpha[j] - 1) * log(M[i * tda + j]);
            for (size_t j = 0; j < m2; j++) sa += alpha[j];
            lp += lgamma(sa);
            for (size_t j = 0; j < m2; j++) lp -= lgamma(alpha[j]);
            total += lp;
        } else {
            ld lp = 0;
            for (size_t j = 0; j < m2; j++) lp += ((ld)alpha[j] - 1) * logl(M[i * tda + j]);
            lp += lgammal(asum);
            for (size_t j = 0; j < m2; j++) lp -= lgammal(alpha[j]);
            total += lp;
        }
    }
    *ll_out = total;
    for (size_t j = 0; j < m2; j++) {
        ld s = 0;
        for (size_t i = 0; i < m1; i++) s += (mode == ORACLE_FAITHFUL) ? (ld)log(M[i * tda + j]) : logl(M[i * tda + j]);
        g_out[j] = (mode == ORACLE_FAITHFUL)
                       ? (ld)((double)s + (double)m1 * (double)digamma_l(asum) - (double)m1 * (double)digamma_l(alpha[j]))
                       : s + (ld)m1 * digamma_l(asum) - (ld)m1 * digamma_l(alpha[j]);
    }
}

/* one_t / apop_tdist_llike, model/apop_t.c:25-38: sum log(gsl_ran_tdist_pdf((x-mu)/sigma, df)) - tsize log(sigma);
 * gsl_ran_tdist_pdf(x, nu) = exp(lnGamma((nu+1)/2) - lnGamma(nu/2)) / sqrt(pi nu) * pow(1 + x^2/nu, -(nu+1)/2) */
void oracle_tdist(const double *v, size_t vsize, const double *M, size_t tda, size_t m1, size_t m2,
                  double mu, double sigma, double df, int mode, long double *ll_out, long double *g_out) {
    const size_t n = vsize + m1 * m2;
    ld sv = 0, sm = 0;
    const double lg1 = lgamma(df / 2), lg2 = lgamma((df + 1) / 2);
    const ld cst = lgammal(((ld)df + 1) / 2) - lgammal((ld)df / 2) - 0.5L * logl(3.14159265358979323846264338327950288L * df);
    for (size_t i = 0; i < n; i++) {
        const double x = CELL(i);
        ld t;
        if (mode == ORACLE_FAITHFUL) {
            const double z = (x - mu) / sigma;
            t = log((exp(lg2 - lg1) / sqrt(M_PI * df)) * pow((1 + z * z / df), -((df + 1) / 2)));
        } else {
            const ld z = ((ld)x - mu) / sigma;
            t = cst - ((ld)df + 1) / 2 * log1pl(z * z / df);
        }
        if (i < vsize) sv += t; else sm += t;
    }
    *ll_out = (mode == ORACLE_FAITHFUL) ? (ld)((double)(sv + sm) - (double)n * log(sigma)) : (sv + sm) - (ld)n * logl(sigma);
    g_out[0] = NAN;   /* no score in the reference */
}

/* yule apply_me / dapply_me / yule_log_likelihood / yule_dlog_likelihood, model/apop_yule.c:30-60 */
void oracle_yule(const double *v, size_t vsize, const double *M, size_t tda, size_t m1, size_t m2,
                 double bb, int mode, long double *ll_out, long double *g_out) {
    const size_t n = vsize + m1 * m2;
    ld sv = 0, sm = 0, dv = 0, dm = 0;
    for (size_t i = 0; i < n; i++) {
        const double x = CELL(i);
        ld t, dt;
        if (mode == ORACLE_FAITHFUL) {
            const double ln_k = (x >= 1) ? lgamma(x) : 0, ln_bb_k = lgamma(x + bb);
            t = ln_k - ln_bb_k;
            dt = -(double)digamma_l(x + bb);
        } else {
            t = ((x >= 1) ? lgammal(x) : 0) - lgammal((ld)x + bb);
            dt = -digamma_l((ld)x + bb);
        }
        if (i < vsize) { sv += t; dv += dt; } else { sm += t; dm += dt; }
    }
    if (mode == ORACLE_FAITHFUL) {
        const ld ln_bb = lgamma(bb), ln_bb_less_1 = log(bb - 1);
        const double likelihood = (double)(sv + sm);
        *ll_out = likelihood + (ln_bb_less_1 + ln_bb) * (ld)n;
        double d_bb = (double)(dv + dm);
        d_bb += (double)((1 / ((ld)bb - 1) + (ld)(double)digamma_l(bb)) * (ld)n);
        g_out[0] = d_bb;
    } else {
        *ll_out = (sv + sm) + (logl((ld)bb - 1) + lgammal(bb)) * (ld)n;
        g_out[0] = (dv + dm) + (1 / ((ld)bb - 1) + digamma_l(bb)) * (ld)n;
    }
}
#undef CELL

/* ---- structure-faithful timed path ------------ */
static double now_s(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec + 1e-9 * t.tv_nsec;
}
int oracle_max_threads(void) { return omp_get_num_procs(); }

typedef double (*row_fn)(double xb, double y);
static row_fn volatile probit_fn_ptr = probit_row_faithful;

/* the LL half alone, threaded as apop_map threads it (OMP_for over rows, apop_mapply.m4.c:240) */
void oracle_probit_ll_omp(const double *X, size_t tda, const double *y, size_t N, size_t k,
                          const double *beta, int threads, long double *ll_out) {
    omp_set_num_threads(threads > 0 ? threads : omp_get_num_procs());
    double *xb = malloc(sizeof(double) * (N ? N : 1));
    ref_dgemm(X, tda, N, k, beta, 1, 1, xb);
    double *mapped = malloc(sizeof(double) * (N ? N : 1));
    row_fn f = probit_fn_ptr;
#pragma omp parallel for
    for (size_t i = 0; i < N; i++) mapped[i] = f(xb[i], y[i]);
    *ll_out = (double)ref_vector_sum(mapped, N);
    free(mapped);
    free(xb);
}

double oracle_time_probit_ll_score(const double *X, size_t tda, const double *y, size_t N,
                                   size_t k, const double *beta, int threads, double *ll_out,
                                   double *g_out, double *sec_ll, double *sec_score) {
    omp_set_num_threads(threads > 0 ? threads : omp_get_num_procs());
    const double t0 = now_s();
    /* LL: apop_dot (single-thread dgemm, fresh N-vector), apop_map (OMP_for over
     * rows, one indirect call per row, fresh N-vector), serial long double sum */
    double *xb = malloc(sizeof(double) * N);
    ref_dgemm(X, tda, N, k, beta, 1, 1, xb);
    double *mapped = malloc(sizeof(double) * N);
    row_fn f = probit_fn_ptr;
#pragma omp parallel for
    for (size_t i = 0; i < N; i++) mapped[i] = f(xb[i], y[i]);
    const double ll = (double)ref_vector_sum(mapped, N);
    free(mapped);
    free(xb);
    const double t1 = now_s();
    /* score: apop_dot again, then the serial double loop */
    long double *g = malloc(sizeof(long double) * k);
    oracle_probit_score(X, tda, y, N, k, beta, ORACLE_FAITHFUL, g, NULL);
    for (size_t j = 0; j < k; j++) g_out[j] = (double)g[j];
    free(g);
    const double t2 = now_s();
    *ll_out = ll;
    if (sec_ll) *sec_ll = t1 - t0;
    if (sec_score) *sec_score = t2 - t1;
    return t2 - t0;
}

double oracle_time_logit_ll(const double *X, size_t tda, const double *y, size_t N, size_t k,
                            const double *B, size_t C, int threads, double *ll_out) {
    omp_set_num_threads(threads > 0 ? threads : omp_get_num_procs());
    const size_t C1 = C - 1;
    double *cats = malloc(sizeof(double) * C);
    for (size_t c = 0; c < C; c++) cats[c] = (double)c;
    const double t0 = now_s();
    double *xb = malloc(sizeof(double) * N * C1);
    ref_dgemm(X, tda, N, k, B, C1, C1, xb);
    double *mapped = malloc(sizeof(double) * N);
#pragma omp parallel for
    for (size_t i = 0; i < N; i++) mapped[i] = logit_row_faithful(xb + i * C1, C1, y[i], cats);
    *ll_out = (double)ref_vector_sum(mapped, N);
    free(mapped);
    free(xb);
    const double t1 = now_s();
    free(cats);
    return t1 - t0;
}

/* ---- round 2 additions --------------------------------------------------------------------- */

/* Observed information of the multinomial logit, -d2LL/dB dB' = sum_i (x_i x_i') (x) (diag p_i - p_i p_i'),
 * P x P in pack order (index j*(C-1)+c), all in long double.  The reference has no analytic form (its
 * Hessian is numeric, apop_mle.m4.c:187-223); this is the derivative of the commented-out score
 * (model/apop_probit.c:244-289) and is what the device kernel is checked against. */
void oracle_logit_information(const double *X, size_t tda, size_t N, size_t k, const double *B, size_t C,
                              long double *H) {
    const size_t C1 = C - 1, P = k * C1;
    for (size_t q = 0; q < P * P; q++) H[q] = 0;
    ld *row = malloc(sizeof(ld) * C1), *p = malloc(sizeof(ld) * C1), *W = malloc(sizeof(ld) * C1 * C1);
    for (size_t i = 0; i < N; i++) {
        const double *x = X + i * tda;
        for (size_t c = 0; c < C1; c++) row[c] = dot_l(x, B + c, C1, k);
        ld max = 0;
        for (size_t c = 0; c < C1; c++) if (row[c] > max) max = row[c];
        ld den = expl(-max);
        for (size_t c = 0; c < C1; c++) den += expl(row[c] - max);
        for (size_t c = 0; c < C1; c++) p[c] = expl(row[c] - max) / den;
        for (size_t a = 0; a < C1; a++)
            for (size_t b = 0; b < C1; b++) W[a * C1 + b] = (a == b ? p[a] : 0) - p[a] * p[b];
        for (size_t j = 0; j < k; j++)
            for (size_t j2 = 0; j2 < k; j2++) {
                const ld xx = (ld)x[j] * x[j2];
                for (size_t a = 0; a < C1; a++)
                    for (size_t b = 0; b < C1; b++) H[(j * C1 + a) * P + j2 * C1 + b] += xx * W[a * C1 + b];
            }
    }
    free(row); free(p); free(W);
}

/* Probit LL and score in EXACT mode over row blocks in parallel: every thread keeps long double
 * partials of its contiguous block, the blocks are added in block order.  Same arithmetic per row as
 * oracle_probit_ll / oracle_probit_score(ORACLE_EXACT); only the grouping of the additions differs
 * (long double: 64-bit mantissa, so the grouping is far below the 1e-12 tolerance).  Used to check the
 * 1e-8 fitted-parameter target at the full 100M x 32 size. */
void oracle_probit_ll_score_exact_omp(const double *X, size_t tda, const double *y, size_t N, size_t k,
                                      const double *beta, int threads, long double *ll_out, long double *g_out,
                                      long double *abs_out) {
    const int T = threads > 0 ? threads : omp_get_num_procs();
    ld *pl = calloc((size_t)T, sizeof(ld)), *pg = calloc((size_t)T * k, sizeof(ld)), *pa = calloc((size_t)T * k, sizeof(ld));
#pragma omp parallel num_threads(T)
    {
        const int t = omp_get_thread_num(), nt = omp_get_num_threads();
        const size_t lo = N * (size_t)t / nt, hi = N * (size_t)(t + 1) / nt;
        ld l = 0;
        ld *g = pg + (size_t)t * k, *a = pa + (size_t)t * k;
        oracle_probit_ll(X + lo * tda, tda, y + lo, hi - lo, k, beta, ORACLE_EXACT, &l);
        oracle_probit_score(X + lo * tda, tda, y + lo, hi - lo, k, beta, ORACLE_EXACT, g, a);
        pl[t] = l;
    }
    *ll_out = 0;
    for (size_t j = 0; j < k; j++) { g_out[j] = 0; if (abs_out) abs_out[j] = 0; }
    for (int t = 0; t < T; t++) {
        *ll_out += pl[t];
        for (size_t j = 0; j < k; j++) { g_out[j] += pg[(size_t)t * k + j]; if (abs_out) abs_out[j] += pa[(size_t)t * k + j]; }
    }
    free(pl); free(pg); free(pa);
}

/* The reference's bootstrap resample (apop_bootstrap.m4.c:143-149): for each of the N rows of the scratch
 * set one gsl_rng_uniform_int draw and an apop_data_memcpy of a one-row view (matrix row + vector element),
 * serial.  taus2 is replaced by xorshift (streams are "parity unpinned"); the memory traffic
 * and the per-row call structure are what is timed.  Returns seconds. */
static void copy_row(double *Xo, double *yo, const double *X, size_t tda, const double *y, size_t k, size_t dst, size_t src) {
    memcpy(Xo + dst * k, X + src * tda, sizeof(double) * k);
    yo[dst] = y[src];
}
double oracle_time_bootstrap_resample(const double *X, size_t tda, const double *y, size_t N, size_t k,
                                      unsigned long long seed, double *Xo, double *yo) {
    void (*volatile cp)(double *, double *, const double *, size_t, const double *, size_t, size_t, size_t) = copy_row;
    unsigned long long s = seed * 0x9E3779B97F4A7C15ull + 1;
    const double t0 = now_s();
    for (size_t j = 0; j < N; j++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        const size_t r = (size_t)(((unsigned __int128)s * N) >> 64);
        cp(Xo, yo, X, tda, y, k, j, r);
    }
    return now_s() - t0;
}
