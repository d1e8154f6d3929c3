// Host build of apophenia_b200/csrc/fastmath.cuh for the CPU accuracy tests
// (tests/test_fastmath_host.py).  Test infrastructure only.
#include "../apophenia_b200/csrc/families.cuh"
extern "C" {
double t_fm_exp(double x) { return apb::fm_exp(x); }
double t_fm_exp_fast(double x) { return apb::fm_exp_fast(x); }
int t_fm_exp_fast_ok(double x) { return apb::fm_exp_fast_ok(x) ? 1 : 0; }
double t_fm_log(double z) { return apb::fm_log(z); }
double t_fm_erfcx(double x) { return apb::fm_erfcx_pos(x); }
void t_probit_row(double xb, double y, double* ll, double* d) { apb::RowOut o = apb::FamProbit::eval(xb, y, 1.0, nullptr); *ll = o.s[0]; *d = o.d; }
void t_logit2_row(double xb, double y, double* ll, double* d) { const double aux[4] = {0, 1, 0, 0}; apb::RowOut o = apb::FamLogit2::eval(xb, y, 1.0, aux); *ll = o.s[0]; *d = o.d; }
void t_fm_gauss_tail(double a, double* E, double* T) { apb::fm_gauss_tail(a, *E, *T); }
}
