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
    static constexpr bool ROW_CONST = false;   // no beta-independent per-row term left out of s[0]
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
    static constexpr bool ROW_CONST = false;
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
    // s[0] leaves out -lgamma(y+1): summed once per residency for a plain pass; with resample counts it is
    // weighted per replicate, so the batched kernels add it per row (row_const)
    static constexpr bool ROW_CONST = true;
    APB_HD static double row_const(double y) { return -lgamma(y + 1.0); }
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
    static constexpr bool ROW_CONST = false;
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

// ---------------------------------------------------------------------------------------------
// Taylor coefficients of a family's row term around a centre x0, for the batched-beta kernel: the m
// parameter vectors of a bootstrap or of parallel Metropolis chains lie within a few standard errors of
// each other, so x_i.beta_r = x0_i + delta with |delta| ~ 1e-3, and
//     ll_i(x0 + delta) = sum_n a_n delta^n,   d_i = dll_i/dxb = sum_n n a_n delta^(n-1)
// to degree TAYLOR_D = 6 is exact to < 1e-16 relative for |delta| <= 1.5e-2 (|a_7| delta^7 <= 5e-5 * 1.7e-13;
// for the score 7 |a_7| delta^6 <= 4e-15)
// -- computed once per row and shared by all m replicates instead of one full link evaluation per
// (row, replicate) pair.  a_0 and a_1 come from the exact evaluation at x0 (Fam::eval: the reference's
// formula shape with its clamps); the higher ones are polynomials in the row's (x0, lambda) or (p).  A row
// whose x0 lies near a clamp / saturation of the reference formula returns false and takes the exact path.
// ---------------------------------------------------------------------------------------------
constexpr int TAYLOR_D = 6;
template <class Fam> struct Taylor;
template <> struct Taylor<FamProbit> {
    // ll(x) = g(s x), g = log Phi, s = +-1 by the outcome; l = phi(z)/Phi(z) = |d|, z = s x0; with l' = -l(l+z):
    //   g2 = -l(l+z), g3 = l(l(2l+3z)+z^2-1), g4 = l(l(l(-6l-12z)-7z^2+4)-z^3+3z),
    //   g5 = l(l(l(l(24l+60z)+50z^2-20)+15z^3-25z)+z^4-6z^2+3)        (derived with sympy, checked against mpmath)
    APB_HD static bool coefs(double x0, double y, const RowOut& o, const double*, double* a) {
        const double s = (y != 0.0) ? 1.0 : -1.0, z = s * x0, l = fabs(o.d), z2 = z * z;
        a[0] = o.s[0];
        a[1] = o.d;
        a[2] = -0.5 * l * (l + z);
        a[3] = s * l * (l * (2.0 * l + 3.0 * z) + z2 - 1.0) * (1.0 / 6.0);
        a[4] = l * (l * (l * (-6.0 * l - 12.0 * z) - 7.0 * z2 + 4.0) - z * (z2 - 3.0)) * (1.0 / 24.0);
        a[5] = s * l * (l * (l * (l * (24.0 * l + 60.0 * z) + 50.0 * z2 - 20.0) + z * (15.0 * z2 - 25.0)) + (z2 * (z2 - 6.0) + 3.0)) * (1.0 / 120.0);
        // g6 = l(l(l(l(l(-120l-360z)-390z^2+120)-180z^3+210z)-31z^4+101z^2-28)-z^5+10z^3-15z)
        a[6] = l * (l * (l * (l * (l * (-120.0 * l - 360.0 * z) - 390.0 * z2 + 120.0) + z * (210.0 - 180.0 * z2)) + (z2 * (101.0 - 31.0 * z2) - 28.0)) - z * (z2 * (z2 - 10.0) + 15.0)) * (1.0 / 720.0);
        // On the unlikely side (z < 0) the reference forms P(observed) = 1 - n from the ROUNDED n = Phi(-x.b): its row
        // term carries a rounding noise of 2^-53 / Phi(z) (8e-4 at z = -7.3, the knife edge of DESIGN.md section 2).  An
        // expansion is smooth where the formula is not, so such rows keep the formula itself: expanded only while that
        // noise is below 5e-13 (z > -3.5); on the likely side up to the saturation / clamp region.
        return z > -3.5 && z < 7.4;
    }
};
template <> struct Taylor<FamLogit2> {
    // ll1 = chosen - p, ll2 = -v with v = p(1-p), ll3 = -v(1-2p), ll4 = -v(1-6v), ll5 = -v(1-2p)(1-12v), ll6 = -v(1-30v+120v^2)
    APB_HD static bool coefs(double x0, double y, const RowOut& o, const double* aux, double* a) {
        const double p = ((y != aux[0]) ? 1.0 : 0.0) - o.d, v = p * (1.0 - p), w = 1.0 - 2.0 * p;
        a[0] = o.s[0];
        a[1] = o.d;
        a[2] = -0.5 * v;
        a[3] = -v * w * (1.0 / 6.0);
        a[4] = -v * (1.0 - 6.0 * v) * (1.0 / 24.0);
        a[5] = -v * w * (1.0 - 12.0 * v) * (1.0 / 120.0);
        a[6] = -v * (1.0 + v * (120.0 * v - 30.0)) * (1.0 / 720.0);     // ll6 = -v(1 - 30v + 120v^2)
        return fabs(x0) < 30.0;
    }
};
template <> struct Taylor<FamPoissonReg> {
    // ll = y x - e^x: every derivative from the second on is -mu, mu = y - d
    APB_HD static bool coefs(double x0, double y, const RowOut& o, const double*, double* a) {
        const double mu = y - o.d;
        a[0] = o.s[0];
        a[1] = o.d;
        a[2] = -mu * 0.5;
        a[3] = -mu * (1.0 / 6.0);
        a[4] = -mu * (1.0 / 24.0);
        a[5] = -mu * (1.0 / 120.0);
        a[6] = -mu * (1.0 / 720.0);
        return fabs(x0) < 30.0;
    }
};

enum GlmFam { GLM_PROBIT = 0, GLM_LOGIT2 = 1, GLM_POISSON_REG = 2, GLM_OLS = 3, GLM_FAM_COUNT = 4 };

}  // namespace apb
