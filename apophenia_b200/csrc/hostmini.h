/* hostmini.h -- host-side mirror of the reference interface around the hot path
 * (see hostmini.c).  Names and argument meaning follow the reference; variadic
 * "designated" wrappers (apop.m4.h:43-52) are replaced by their plain base forms. */
#ifndef APOP_HOSTMINI_H
#define APOP_HOSTMINI_H
#include "../../include/apop_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

gsl_vector *gsl_vector_alloc(size_t n);
gsl_vector *gsl_vector_calloc(size_t n);
void gsl_vector_free(gsl_vector *v);
gsl_matrix *gsl_matrix_alloc(size_t n1, size_t n2);
gsl_matrix *gsl_matrix_calloc(size_t n1, size_t n2);
void gsl_matrix_free(gsl_matrix *m);

apop_data *apop_data_alloc(size_t size1, size_t size2, int size3);
apop_data *apop_data_calloc(size_t size1, size_t size2, int size3);
void apop_data_free(apop_data *d);
apop_data *apop_data_add_page(apop_data *dataset, apop_data *newpage, const char *title);
apop_data *apop_data_get_page(const apop_data *data, const char *title);
gsl_vector *apop_data_pack(const apop_data *in, gsl_vector *out);
void apop_data_unpack(const gsl_vector *in, apop_data *d);

int apop_vtable_add(char const *tabname, void *fn_in, unsigned long hash);
void *apop_vtable_get(char const *tabname, unsigned long hash);
int apop_vtable_drop(char const *tabname, unsigned long hash);

/* apop_mle_settings (apop.m4.h:1200-1240; defaults apop_mle.m4.c:52-70), the fields this
 * path uses.  relative_tolerance is an addition: at 1e8 rows the reference's absolute
 * |gradient| < tolerance test is unreachable (last bullet); when set, the
 * search stops at |gradient| < tolerance * max(1, |LL|). */
typedef struct {
    double *starting_pt;
    double tolerance;
    int max_iterations;
    char method[32];   /* "BFGS cg" | "FR cg" | "NM simplex" | "Newton" | "" (default) */
    int verbose;
    double step_size;
    double delta;
    int relative_tolerance;
} apop_mle_settings;
apop_mle_settings apop_mle_settings_defaults(void);
apop_mle_settings *apop_mle_settings_add(apop_model *m, apop_mle_settings s);
void *apop_settings_get_grp(apop_model *m, const char *type);
void *apop_settings_group_alloc(apop_model *m, const char *type, void *group, size_t bytes);

apop_model *apop_model_copy(apop_model *in);
void apop_model_free(apop_model *m);
void apop_model_clear(apop_data *data, apop_model *m);
void apop_prep(apop_data *d, apop_model *m);
double apop_log_likelihood(apop_data *d, apop_model *m);
void apop_score(apop_data *d, gsl_vector *out, apop_model *m);
gsl_vector *apop_numerical_gradient(apop_data *data, apop_model *m, double delta);
apop_model *apop_estimate(apop_data *d, apop_model *m);
void apop_maximum_likelihood(apop_data *data, apop_model *dist);

apop_data *apop_b200_category_table(apop_data *d);
void apop_b200_ols_shuffle(apop_data *d);

/* fit diagnostics of the last apop_maximum_likelihood call on this thread */
typedef struct {
    int status;          /* 0 optimum found / no further progress, -1 NaN or max iterations */
    int iterations;
    int f_evals, df_evals;
    double ll;
    double gradient_norm;
    double seconds;
} apop_mle_report;
apop_mle_report apop_mle_last_report(void);

/* ---- drivers.c ---- */
int apop_b200_family_of(apop_model *m);      /* which device family is installed in m (-1: none) */
void apop_b200_install_estimates(void);      /* closed-form estimate methods of poisson / normal / ols */
apop_data *apop_data_covariance(const apop_data *in);
apop_data *apop_model_hessian(apop_data *data, apop_model *model, double delta);   /* numeric: apop_mle.m4.c:187-223 */
int apop_b200_observed_information(apop_data *data, apop_model *m, double *H, int *analytic);   /* device pass, else numeric */
apop_data *apop_model_numerical_covariance(apop_data *data, apop_model *m);   /* (observed information)^-1 */
/* apop_bootstrap_cov (apop_bootstrap.m4.c:122): gsl_rng* replaced by a seed; iterations <= 0 means 1000 */
apop_data *apop_bootstrap_cov(apop_data *data, apop_model *model, uint64_t seed, int iterations, apop_data **boot_store);
/* apop_mcmc_settings (apop.m4.h, defaults apop_mcmc.m4.c:35-43) + initial proposal scale, seed, chain count */
typedef struct {
    int periods;
    double burnin;
    double target_accept_rate;
    double initial_scale;
    uint64_t seed;
    int n_chains;
} apop_mcmc_settings;
apop_mcmc_settings apop_mcmc_settings_defaults(void);
apop_data *apop_model_metropolis(apop_data *d, apop_model *m, apop_mcmc_settings s);
apop_data *apop_b200_metropolis_chains(apop_data *d, apop_model *m, apop_mcmc_settings s);

extern apop_model *apop_exponential, *apop_gamma, *apop_lognormal;
extern apop_model *apop_probit, *apop_logit, *apop_poisson, *apop_normal, *apop_ols, *apop_poisson_regression;

#ifdef __cplusplus
}
#endif
#endif
