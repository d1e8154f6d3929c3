/* dropin_probit.c -- the INTEGRATION.md flow as a plain C program: no Python, no torch.
 * Links the host-side mirror of the reference interface (libapop_hostmini) and the device
 * library (libapop_b200) exactly as a program would link libapophenia + libapop_b200.
 *
 * Generates the reference's own test sample (tests/test_apop.c:889-907, 80 000 rows), runs
 *     register -> prep -> pin -> apop_estimate -> numerical covariance -> bootstrap
 * and checks what test_probit_and_logit checks (distance to the true parameters < 0.07),
 * plus agreement of the bootstrap variances with the inverse information.
 * Exit status 0 on success.  Build/run: tests/test_gpu_c_dropin.py. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../apophenia_b200/csrc/hostmini.h"

static double urand(unsigned long long *s) { /* splitmix64 -> (0,1) */
    unsigned long long z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    return ((z >> 11) + 0.5) * 0x1.0p-53;
}
static double nrand(unsigned long long *s) { return sqrt(-2 * log(urand(s))) * cos(2 * M_PI * urand(s)); }

int main(void) {
    const size_t n = 80000, k = 5;
    unsigned long long seed = 8;
    double beta[5];
    for (size_t j = 0; j < k; j++) beta[j] = (urand(&seed) - 0.5) * 2;
    apop_data *d = apop_data_alloc(0, n, (int)k);       /* outcome in column 0, like the reference */
    for (size_t i = 0; i < n; i++) {
        double *row = d->matrix->data + i * k, xb = beta[0];
        for (size_t j = 1; j < k; j++) { row[j] = (urand(&seed) - 0.5) * 2; xb += row[j] * beta[j]; }
        row[0] = nrand(&seed) > -xb;
    }
    if (apop_b200_init(0)) { fprintf(stderr, "init: %s\n", apop_b200_last_error()); return 2; }
    apop_model *m = apop_model_copy(apop_probit);
    if (apop_b200_register(m, APOP_B200_PROBIT)) { fprintf(stderr, "register: %s\n", apop_b200_last_error()); return 2; }
    apop_mle_settings s = apop_mle_settings_defaults();
    snprintf(s.method, sizeof s.method, "BFGS cg");
    s.tolerance = 1e-6;
    apop_mle_settings_add(m, s);
    apop_prep(d, m);                 /* factor page, ols_shuffle, start point */
    if (apop_b200_pin(d)) { fprintf(stderr, "pin: %s\n", apop_b200_last_error()); return 2; }
    apop_model *est = apop_estimate(d, m);
    apop_mle_report rep = apop_mle_last_report();
    double dist = 0;
    for (size_t j = 0; j < k; j++) { const double e = est->parameters->matrix->data[j] - beta[j]; dist += e * e; }
    dist = sqrt(dist);
    printf("status %d, %d iterations, %d passes, LL %.6f, |beta_hat - beta| = %.4f\n", rep.status, rep.iterations, rep.f_evals, rep.ll, dist);
    apop_data *cov = apop_model_numerical_covariance(d, est);
    apop_data *boots = NULL;
    apop_data *bcov = apop_bootstrap_cov(d, est, 3, 300, &boots);
    if (!cov || !bcov) { fprintf(stderr, "covariance: %s\n", apop_b200_last_error()); return 2; }
    double worst = 0;
    for (size_t j = 0; j < k; j++) {
        const double r = bcov->matrix->data[j * k + j] / cov->matrix->data[j * k + j];
        printf("  beta[%zu] = % .5f  (true % .5f)  se(info) %.5f  se(bootstrap) %.5f\n", j, est->parameters->matrix->data[j], beta[j],
               sqrt(cov->matrix->data[j * k + j]), sqrt(bcov->matrix->data[j * k + j]));
        if (fabs(r - 1) > worst) worst = fabs(r - 1);
    }
    printf("kernels launched: %llu\n", apop_b200_launch_count());
    const int ok = rep.status == 0 && dist < 0.07 && worst < 0.35;
    apop_b200_unpin(d);
    apop_b200_finalize();
    printf(ok ? "OK\n" : "FAILED\n");
    return ok ? 0 : 1;
}
