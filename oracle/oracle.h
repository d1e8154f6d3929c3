/* oracle.h -- CPU restatement of Apophenia's likelihood/score path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call it, and only as the checker or the timed CPU arm.
 *
 * The reference (Apophenia 1.0) cannot be built in this image (needs m4,
 * autotools, GSL, sqlite3), so this file restates its arithmetic,
 * function by function, citing the reference lines each one follows.  GSL
 * (unpinned, >= 1.12; install/configure.ac:35) is absent as well; its calls on
 * the path are replaced by libm equivalents:
 *     gsl_cdf_gaussian_P(x,1)  -> 0.5*erfc(-x/sqrt 2) with GSL's saturation
 *                                 cut-offs (cdf/gauss.c: 0 below -37.519, 1 above 8.572)
 *     gsl_ran_gaussian_pdf     -> exp(-u*u/2)/(sqrt(2 pi)|sigma|)   (randist/gauss.c)
 *     gsl_sf_lngamma(x+1)      -> lgamma(x+1)
 *     gsl_blas_dgemm/dgemv     -> the reference CBLAS loops (row-major, no FMA)
 *     gsl_stats_mean/variance  -> the running-mean recurrences in long double
 * PARITY PIN: the reference holds no test of LL/score values at fixed beta for
 * probit/logit/Poisson.  What the reference's tests do pin -- the
 * eg/fake_logit.c fit, the NIST StRD OLS coefficients, apop_map_sum / apop_dot
 * small exact cases, pack order -- is checked in tests/test_oracle_golden.py;
 * the special functions are additionally checked against mpmath at 50 digits.
 *
 * Two accumulation modes:
 *   ORACLE_FAITHFUL  the reference's own types and operation order
 *                    (materialised X.beta, per-row map, serial x87 long double
 *                    sum; serial double score accumulation)
 *   ORACLE_EXACT     every operation in x87 long double (64-bit mantissa),
 *                    clamps applied where the double-rounded value would clamp
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stddef.h>

#define ORACLE_FAITHFUL 0
#define ORACLE_EXACT 1

#ifdef __cplusplus
extern "C" {
#endif

/* X is row-major N x k with row stride tda (gsl_matrix); y has N entries;
 * w is NULL or N weights.  Results are written through pointers so that
 * ctypes never has to return an x87 value. */

double oracle_gauss_cdf(double x);               /* gsl_cdf_gaussian_P(x, 1) stand-in */
double oracle_gauss_pdf(double x, double sigma); /* gsl_ran_gaussian_pdf stand-in */

/* model/apop_probit.c:83-98 (LL) and :130-154 (score).  abs_out (may be NULL)
 * receives sum_i |X_ij d_i|, the scale of the score tolerance. */
void oracle_probit_ll(const double *X, size_t tda, const double *y, size_t N, size_t k,
                      const double *beta, int mode, long double *ll_out);
void oracle_probit_score(const double *X, size_t tda, const double *y, size_t N, size_t k,
                         const double *beta, int mode, long double *g_out, long double *abs_out);
/* model/apop_probit.c:65-80: start point beta_j = exp(mean_i ln|x_ij|) */
void oracle_probit_start(const double *X, size_t tda, size_t N, size_t k, double *beta0);

/* model/apop_probit.c:206-242.  B is k x (C-1) row-major (parameters->matrix);
 * cats are the C sorted category values (factor page vector). */
void oracle_logit_ll(const double *X, size_t tda, const double *y, size_t N, size_t k,
                     const double *B, size_t C, const double *cats, int mode, long double *ll_out);
/* the analytic score the reference has commented out (model/apop_probit.c:244-289),
 * output index j*(C-1)+c (pack order, apop_conversions.m4.c:762-806) */
void oracle_logit_score(const double *X, size_t tda, const double *y, size_t N, size_t k,
                        const double *B, size_t C, const double *cats, int mode,
                        long double *g_out, long double *abs_out);

/* model/apop_poisson.c:14-28 and :66-73; v (vsize entries) and/or M (msize1 x msize2) */
void oracle_poisson_ll(const double *v, size_t vsize, const double *M, size_t tda, size_t m1,
                       size_t m2, double lambda, int mode, long double *ll_out);
void oracle_poisson_score(const double *v, size_t vsize, const double *M, size_t tda, size_t m1,
                          size_t m2, double lambda, int mode, long double *g_out);
/* model/apop_normal.c:38-50 and :108-118 */
void oracle_normal_ll(const double *v, size_t vsize, const double *M, size_t tda, size_t m1,
                      size_t m2, double mu, double sigma, int mode, long double *ll_out);
void oracle_normal_score(const double *v, size_t vsize, const double *M, size_t tda, size_t m1,
                         size_t m2, double mu, double sigma, int mode, long double *g_out);
/* model/apop_ols.c:129-172 and :175-205 (prepped data: y in the vector; no
 * <Predicted>/<Error variance> pages, improper-uniform input distribution) */
void oracle_ols_ll(const double *X, size_t tda, const double *y, const double *w, size_t N,
                   size_t k, const double *beta, int mode, long double *ll_out);
void oracle_ols_score(const double *X, size_t tda, const double *y, const double *w, size_t N,
                      size_t k, const double *beta, int mode, long double *g_out,
                      long double *abs_out);
/* X'X and X'y as apop_dot(.form1='t') computes them (model/apop_ols.c:333-334) */
void oracle_ols_gram(const double *X, size_t tda, const double *y, size_t N, size_t k, int mode,
                     long double *xtx, long double *xty);
/* Poisson regression, the new model of SURVEY a16 (no reference counterpart) */
void oracle_poisson_reg_ll(const double *X, size_t tda, const double *y, size_t N, size_t k,
                           const double *beta, int mode, long double *ll_out);
void oracle_poisson_reg_score(const double *X, size_t tda, const double *y, size_t N, size_t k,
                              const double *beta, int mode, long double *g_out,
                              long double *abs_out);

/* apop_map_sum over every cell (apop_mapply.m4.c:485-517 with .fn_dp, part 'a'):
 * fn ids match apop_b200_map_fn in include/apop_b200.h */
void oracle_map_sum(const double *v, size_t vsize, const double *M, size_t tda, size_t m1,
                    size_t m2, int fn, double p0, int mode, long double *out);

/* more iid distributions on the apop_map_sum shape: exponential
 * (model/apop_exponential.c:33-52), gamma (model/apop_gamma.c:31-62; gsl_sf_psi -> recurrence +
 * asymptotic series), lognormal (model/apop_normal.c:166-237).  g_out: 1, 2, 2 entries. */
double oracle_digamma(double x);
void oracle_bernoulli(const double *v, size_t vsize, const double *M, size_t tda, size_t m1, size_t m2,
                      double p, int mode, long double *ll_out, long double *g_out);
void oracle_beta(const double *v, size_t vsize, const double *M, size_t tda, size_t m1, size_t m2,
                 double a, double b, int mode, long double *ll_out, long double *g_out);
void oracle_zipf(const double *v, size_t vsize, const double *M, size_t tda, size_t m1, size_t m2,
                 double bb, int mode, long double *ll_out, long double *g_out);
void oracle_dirichlet(const double *M, size_t tda, size_t m1, size_t m2, const double *alpha, int mode,
                      long double *ll_out, long double *g_out);
void oracle_tdist(const double *v, size_t vsize, const double *M, size_t tda, size_t m1, size_t m2,
                  double mu, double sigma, double df, int mode, long double *ll_out, long double *g_out);
void oracle_yule(const double *v, size_t vsize, const double *M, size_t tda, size_t m1, size_t m2,
                 double bb, int mode, long double *ll_out, long double *g_out);
double oracle_zeta(double s);
void oracle_exponential(const double *v, size_t vsize, const double *M, size_t tda, size_t m1, size_t m2,
                        double mu, int mode, long double *ll_out, long double *g_out);
void oracle_gamma(const double *v, size_t vsize, const double *M, size_t tda, size_t m1, size_t m2,
                  double a, double b, int mode, long double *ll_out, long double *g_out);
void oracle_lognormal(const double *v, size_t vsize, const double *M, size_t tda, size_t m1, size_t m2,
                      double mu, double sd, int mode, long double *ll_out, long double *g_out);

/* observed information -d2LL/db db' for probit / binary logit / poisson_reg
 * (exact mode only; no reference counterpart) */
void oracle_information(int family, const double *X, size_t tda, const double *y, size_t N,
                        size_t k, const double *beta, long double *H);

/* Structure-faithful timed path for the CPU baseline:
 * LL = materialised dgemm + omp row map through a function pointer + serial
 * long double sum; score = dgemm again + serial loop.  Returns seconds. */
void oracle_probit_ll_omp(const double *X, size_t tda, const double *y, size_t N, size_t k,
                          const double *beta, int threads, long double *ll_out);
double oracle_time_probit_ll_score(const double *X, size_t tda, const double *y, size_t N,
                                   size_t k, const double *beta, int threads, double *ll_out,
                                   double *g_out, double *sec_ll, double *sec_score);
double oracle_time_logit_ll(const double *X, size_t tda, const double *y, size_t N, size_t k,
                            const double *B, size_t C, int threads, double *ll_out);
int oracle_max_threads(void);

/* round 2: multinomial-logit observed information (exact), the OpenMP EXACT probit pass used at the full
 * 100M x 32 size, and the bootstrap row-copy loop of apop_bootstrap.m4.c:143-149 for the CPU baseline */
void oracle_logit_information(const double *X, size_t tda, size_t N, size_t k, const double *B, size_t C,
                              long double *H);
void oracle_probit_ll_score_exact_omp(const double *X, size_t tda, const double *y, size_t N, size_t k,
                                      const double *beta, int threads, long double *ll_out, long double *g_out,
                                      long double *abs_out);
double oracle_time_bootstrap_resample(const double *X, size_t tda, const double *y, size_t N, size_t k,
                                      unsigned long long seed, double *Xo, double *yo);

#ifdef __cplusplus
}
#endif
#endif
