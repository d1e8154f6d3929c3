/* cpu_models.c -- the oracle behind the reference's model interface.
 * TEST / BASELINE INFRASTRUCTURE ONLY (see oracle.h): log_likelihood and score entries with
 * the reference's signatures that evaluate the structure-faithful CPU restatement on the HOST
 * copy of the data.  Installed into a model exactly where apop_b200_register installs the
 * device entries, they let the same host optimizer (csrc/mle.c) produce the "restated
 * reference CPU path" fit that BASELINE.md section 3 asks for (config 1: probit BFGS on
 * 1M x 10), next to the device-driven fit. */
#include "oracle.h"
#include "../include/apop_b200.h"
#include <math.h>

#define MAXK 256
static int threads = 0;   /* 0 = all host cores */
void oracle_models_set_threads(int t) { threads = t; }

static int beta_of(const apop_data *d, const apop_model *m, double *b) {
    if (!d || !m || !m->parameters || !d->matrix || !d->vector) return -1;
    const size_t k = d->matrix->size2;
    const apop_data *p = m->parameters;
    if (k > MAXK || (!p->matrix && !p->vector)) return -1;
    for (size_t j = 0; j < k; j++) b[j] = p->matrix ? p->matrix->data[j * p->matrix->tda] : p->vector->data[j * p->vector->stride];
    return (int)k;
}

/* multiprobit_log_likelihood -> biprobit_log_likelihood (model/apop_probit.c:91-98,104): apop_dot
 * (single-thread dgemm), apop_map_sum (OpenMP row map through a function pointer, serial long
 * double sum) */
long double oracle_model_probit_ll(apop_data *d, apop_model *m) {
    double b[MAXK];
    const int k = beta_of(d, m, b);
    if (k < 0) return NAN;
    long double out;
    oracle_probit_ll_omp(d->matrix->data, d->matrix->tda, d->vector->data, d->matrix->size1, (size_t)k, b, threads, &out);
    return out;
}
/* probit_dlog_likelihood (model/apop_probit.c:130-154): apop_dot again, then the serial loop */
void oracle_model_probit_score(apop_data *d, gsl_vector *grad, apop_model *m) {
    double b[MAXK];
    long double g[MAXK];
    const int k = beta_of(d, m, b);
    if (k < 0 || !grad) return;
    oracle_probit_score(d->matrix->data, d->matrix->tda, d->vector->data, d->matrix->size1, (size_t)k, b, ORACLE_FAITHFUL, g, NULL);
    for (int j = 0; j < k && (size_t)j < grad->size; j++) grad->data[j * grad->stride] = (double)g[j];
}
