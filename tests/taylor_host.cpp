// host build of the Taylor coefficient functions of families.cuh (the device compiles the same source)
#include "../apophenia_b200/csrc/families.cuh"
using namespace apb;
template <class Fam>
static int coefs(double x0, double y, double* a, double* ll0, double* d0) {
    const double aux[4] = {0, 1, 0, 0};
    const RowOut o = Fam::eval(x0, y, 1.0, aux);
    *ll0 = o.s[0]; *d0 = o.d;
    return Taylor<Fam>::coefs(x0, y, o, aux, a) ? 1 : 0;
}
extern "C" {
int t_degree(void) { return TAYLOR_D; }
int t_probit_coefs(double x0, double y, double* a, double* ll0, double* d0) { return coefs<FamProbit>(x0, y, a, ll0, d0); }
int t_logit2_coefs(double x0, double y, double* a, double* ll0, double* d0) { return coefs<FamLogit2>(x0, y, a, ll0, d0); }
int t_poisson_coefs(double x0, double y, double* a, double* ll0, double* d0) { return coefs<FamPoissonReg>(x0, y, a, ll0, d0); }
}
