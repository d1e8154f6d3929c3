// families.cuh -- per-row link/density and score multiplier of each model family.
//
// A family turns (x.beta, y, w) of one row into up to MAX_ACC scalar terms that
// are summed over rows, and the multiplier d_i of the score  g_j = sum_i X_ij d_i.
// The formula shapes follow the reference (file:line under /root/reference) so
// that clamps and branches trigger on the same rows; see SURVEY.md 8(a).
#pragma once
#include "fastmath.cuh"
#ifdef __CUDACC__
#include "common.cuh"
#else
namespace apb { constexpr int MAX_ACC = 3; }
#endif

namespace apb {

// gsl_cdf_gaussian_P(x, 1): Phi(x) = erfc(-x/sqrt2)/2 with GSL's saturation
// cut-offs (cdf/gauss.c GAUSS_XLOWER / GAUSS_XUPPER), as in oracle/oracle.c.
APB_HD double gauss_cdf(double x) {
    double r = 0.5 * erfc(-x * 0.70710678118654752440);
    r = (x < -37.519) ? 0.0 : r;
    r = (x > 8.572) ? 1.0 : r;
    return r;
}
// gsl_ran_gaussian_pdf(x, 1)
APB_HD double gauss_pdf1(double x) {
    return 0.39894228040143267794 * exp(-0.5 * x * x);
}

struct RowOut {
    double s[MAX_ACC];  // scalar terms to sum over rows
    double d;           // score multiplier
};

// Binary probit: biprobit_ll_row (model/apop_probit.c:83-88) and the row body
// of probit_dlog_likelihood (:142-151), fused.
struct FamProbit {
    static constexpr int NACC = 1;
    static constexpr bool USES_W = false;
    APB_HD static RowOut eval(double xb, double y, double w, const double* aux) {
        RowOut o;
#ifdef APB_USE_LIBM_DEVICE
        double n = gauss_cdf(-xb);
        const double pdf = gauss_pdf1(-xb);
#else
        // Phi(-|xb|) and phi(xb) from one exponential (fastmath.cuh); then the reference's
        // own shape: n = Phi(-xb) as a rounded double, GSL's saturation, the two clamps.
        double E, T;
        fm_gauss_tail(fabs(xb), E, T);
        double n = (xb >= 0.0) ? T : 1.0 - T;
        n = (-xb < -37.519) ? 0.0 : n;
        n = (-xb > 8.572) ? 1.0 : n;
        const double pdf = 0.39894228040143267794 * E;
#endif
        n = (n != 0.0) ? n : 1e-10;          // prevent -inf in the log
        n = (n < 1.0) ? n : 1.0 - 1e-10;
        const bool yes = (y != 0.0);
        const double arg = yes ? 1.0 - n : n;  // P(observed outcome)
#ifdef APB_USE_LIBM_DEVICE
        o.s[0] = log(arg);
#else
        o.s[0] = fm_log_normal(arg);             // arg in [1e-10, 1): never subnormal
#endif
#ifdef APB_USE_LIBM_DEVICE
        o.d = (yes ? pdf : -pdf) / arg;
#else
        o.d = (yes ? pdf : -pdf) * fm_rcp(arg);  // <= 1.5 ulp; the exact quotient costs 3 more FP64 operations per row
#endif
        return o;
    }
};

// Binary logit: one_logit_row (model/apop_probit.c:212-229) with one free
// category, plus the analytic score the reference leaves commented out (:244-289).
// aux[0], aux[1] are the two category values (factor page vector).
struct FamLogit2 {
    static constexpr int NACC = 1;
    static constexpr bool USES_W = false;
    APB_HD static RowOut eval(double xb, double y, double w, const double* aux) {
        RowOut o;
        const bool chosen = (y != aux[0]);  // find_index: 0 = numeraire, else 1
        const double num = chosen ? xb : 0.0;
        // max + log(exp(xb-max) + exp(-max)) with max = xb:  xb + log(1 + exp(-xb));
        // when exp(-max) overflows the reference returns num - (max - max).
#ifdef APB_USE_LIBM_DEVICE
        const double em = exp(-xb);
        const double lse = (xb < -700.0) ? 0.0 : xb + log(1.0 + em);
#else
#ifdef __CUDA_ARCH__
        // every caller evaluates rows with the whole warp active: one vote picks the unclamped
        // 17-operation exponential when all 32 rows have |xb| < 708 (they practically always do)
        double em;
        if (__all_sync(0xffffffffu, fm_exp_fast_ok(xb))) em = fm_exp_fast(-xb);
        else em = fm_exp(-xb);
#else
        const double em = fm_exp(-xb);
#endif
        const double lse = (xb < -700.0) ? 0.0 : xb + fm_log_normal(1.0 + em);   // 1 + em >= 1 (inf only where xb < -709)
#endif
        o.s[0] = num - lse;
#ifdef APB_USE_LIBM_DEVICE
        const double p = 1.0 / (1.0 + em);   // P(category 1); em=inf -> 0
#else
        const double p = fm_rcp(1.0 + fmin(em, 1e300));   // P(category 1); em = inf -> 0
#endif
        o.d = (chosen ? 1.0 : 0.0) - p;
        return o;
    }
};

// Poisson regression (new model): l_i = y xb - exp(xb) - lgamma(y+1);
// the lgamma term does not depend on beta and is summed once at pin time.
struct FamPoissonReg {
    static constexpr int NACC = 1;
    static constexpr bool USES_W = false;
    APB_HD static RowOut eval(double xb, double y, double w, const double* aux) {
        RowOut o;
#ifdef APB_USE_LIBM_DEVICE
        const double mu = exp(xb);
#else
        const double mu = fm_exp(xb);
#endif
        o.s[0] = y * xb - mu;
        o.d = y - mu;
        return o;
    }
};

// OLS: e = x.beta - y (model/apop_ols.c:146-155).  The kernel returns the
// moments sum e, sum e^2, sum log w and the unnormalised score sum w X e; the
// host turns them into sigma (sample variance, :159), the Gaussian LL (:161-170)
// and the score / sigma^2 (:189-203).
struct FamOls {
    static constexpr int NACC = 3;
    static constexpr bool USES_W = true;
    APB_HD static RowOut eval(double xb, double y, double w, const double* aux) {
        RowOut o;
        const double e = xb - y;
        o.s[0] = e;
        o.s[1] = e * e;
        o.s[2] = (aux[0] != 0.0) ? w * e * e : log(w);   // aux[0]: the un-prepped score's sum w e^2 (capi_models.cu)
        o.d = w * e;
        return o;
    }
};

enum GlmFam { GLM_PROBIT = 0, GLM_LOGIT2 = 1, GLM_POISSON_REG = 2, GLM_OLS = 3, GLM_FAM_COUNT = 4 };

}  // namespace apb
