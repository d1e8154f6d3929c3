/* apop_b200.h -- C-ABI of the B200-native likelihood/score hot path.
 *
 * This is the drop-in boundary: every entry point below has the exact signature
 * the reference's own dispatch would bind for this path, so that a maintainer
 * can install it into an unmodified libapophenia.
 *
 *   model->log_likelihood   long double (*)(apop_data*, apop_model*)      apop.m4.h:106
 *   apop_score vtable entry void (*)(apop_data*, gsl_vector*, apop_model*) apop.m4.h:531
 *
 * No C++ or torch types cross this boundary: plain pointers and sizes only.
 *
 * When the real <apop.h> / GSL headers are available, compile with
 * -DAPOP_B200_USE_SYSTEM_HEADERS and the structs come from there.  Otherwise
 * the layouts are re-declared here bit-for-bit (public, ABI-stable layouts;
 * apop.m4.h:61-115 and GSL's gsl_block/gsl_vector/gsl_matrix) and checked by
 * static asserts in capi_runtime.cu.
 */
#ifndef APOP_B200_H
#define APOP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifdef APOP_B200_USE_SYSTEM_HEADERS
#include <apop.h>
#else
/* ---- GSL containers (gsl_block_double.h, gsl_vector_double.h, gsl_matrix_double.h) */
typedef struct { size_t size; double *data; } gsl_block;
typedef struct { size_t size; size_t stride; double *data; gsl_block *block; int owner; } gsl_vector;
typedef struct { size_t size1; size_t size2; size_t tda; double *data; gsl_block *block; int owner; } gsl_matrix;
typedef struct gsl_rng_s gsl_rng; /* opaque here */

/* ---- apop_name (apop.m4.h:61-69) */
typedef struct {
    char *title;
    char *vector;
    char **col;
    char **row;
    char **text;
    int colct, rowct, textct;
    unsigned long *colhash, *rowhash, *texthash;
} apop_name;

/* ---- apop_data (apop.m4.h:72-81) */
typedef struct apop_data {
    gsl_vector *vector;
    gsl_matrix *matrix;
    apop_name *names;
    char ***text;
    size_t textsize[2];
    gsl_vector *weights;
    struct apop_data *more;
    char error;
} apop_data;

/* ---- settings (apop.m4.h:85-91) */
typedef struct {
    char name[101];
    unsigned long name_hash;
    void *setting_group;
    void *copy;
    void *free;
} apop_settings_type;

/* ---- apop_model (apop.m4.h:94-115) */
typedef struct apop_model apop_model;
struct apop_model {
    char name[101];
    int vsize, msize1, msize2, dsize;
    apop_data *data;
    apop_data *parameters;
    apop_data *info;
    void (*estimate)(apop_data *data, apop_model *params);
    long double (*p)(apop_data *d, apop_model *params);
    long double (*log_likelihood)(apop_data *d, apop_model *params);
    long double (*cdf)(apop_data *d, apop_model *params);
    long double (*constraint)(apop_data *data, apop_model *params);
    int (*draw)(double *out, gsl_rng *r, apop_model *params);
    void (*prep)(apop_data *data, apop_model *params);
    apop_settings_type *settings;
    void *more;
    size_t more_size;
    char error;
};

/* ---- apop_lm_settings (apop.m4.h:1243-1249): the "apop_lm" settings group of apop_ols; the OLS entry
 * reads input_distribution from it (model/apop_ols.c:135-136,163-166) */
typedef struct {
    int destroy_data;
    apop_data *instruments;
    char want_cov;
    char want_expected_value;
    apop_model *input_distribution;
} apop_lm_settings;

/* vtable signature for the score (apop.m4.h:531-533) */
typedef void (*apop_score_type)(apop_data *d, gsl_vector *gradient, apop_model *params);
#define apop_score_hash(m1) ((size_t)((m1)->log_likelihood ? (m1)->log_likelihood : (m1)->p))
#endif /* APOP_B200_USE_SYSTEM_HEADERS */

/* ------------------------------------------------------------------ *
 *  Model families served by this library
 * ------------------------------------------------------------------ */
typedef enum {
    APOP_B200_PROBIT = 0,       /* binary probit            model/apop_probit.c:83-154  */
    APOP_B200_LOGIT = 1,        /* multinomial logit        model/apop_probit.c:206-242 (+ analytic score, :244-289) */
    APOP_B200_POISSON = 2,      /* iid Poisson(lambda)      model/apop_poisson.c:14-28,66-73 */
    APOP_B200_NORMAL = 3,       /* iid Normal(mu,sigma)     model/apop_normal.c:38-50,108-118 */
    APOP_B200_OLS = 4,          /* Gaussian-error OLS       model/apop_ols.c:129-205 */
    APOP_B200_POISSON_REG = 5,  /* Poisson regression (new model, no reference counterpart) */
    APOP_B200_EXPONENTIAL = 6,  /* iid Exponential(mu)      model/apop_exponential.c:33-52 */
    APOP_B200_GAMMA = 7,        /* iid Gamma(a, b)          model/apop_gamma.c:31-62 */
    APOP_B200_LOGNORMAL = 8,    /* iid Lognormal(mu, sigma) model/apop_normal.c:166-237 */
    APOP_B200_BERNOULLI = 9,    /* iid Bernoulli(p)         model/apop_bernoulli.c:13-23 */
    APOP_B200_BETA = 10,        /* iid Beta(alpha, beta)    model/apop_beta.c:47-75 */
    APOP_B200_ZIPF = 11,        /* iid Zipf(a)              model/apop_zipf.c:31-40 */
    APOP_B200_DIRICHLET = 12,   /* rows ~ Dirichlet(alpha)  model/apop_dirichlet.c:13-38 */
    APOP_B200_T = 13,           /* iid Student t(mu, sigma, df)  model/apop_t.c:25-38 */
    APOP_B200_YULE = 14,        /* iid Yule(a)              model/apop_yule.c:30-60 */
    APOP_B200_MODEL_COUNT = 15
} apop_b200_family;

/* ------------------------------------------------------------------ *
 *  Runtime
 * ------------------------------------------------------------------ */
/* Select the CUDA device this process drives (one process per GPU).  Returns 0
 * on success; nonzero (and a message through apop_b200_last_error) if no
 * sm_100 device is available.  There is no CPU fallback. */
int apop_b200_init(int device);
void apop_b200_finalize(void);
const char *apop_b200_last_error(void);
/* number of likelihood/score kernels launched since init (bench's gpu_launches) */
unsigned long long apop_b200_launch_count(void);
/* device milliseconds (CUDA events on the library stream) spent in the most
 * recent fused pass, and the accumulated total and count since the last reset */
double apop_b200_last_pass_ms(void);
void apop_b200_pass_timer_reset(void);
double apop_b200_pass_timer_total_ms(unsigned long long *n_passes);
/* Per-pass CUDA-event timing feeds the three calls above.  It is off by default (two event records and a
 * query cost ~25 us per pass); switch it on here or with APOP_B200_TIMING=1.  With timing off the pass
 * counter still runs and the millisecond figures stay 0. */
void apop_b200_set_timing(int on);

/* Multi-GPU: rows are sharded, one rank per process.  The 128-byte NCCL unique
 * id is produced on rank 0 and carried to the other ranks by the host program
 * (torch.distributed / MPI / a file); afterwards every ll/score call ends with
 * one ncclAllReduce(sum, f64) of [LL | score] on the library stream.
 * (No reference counterpart: the reference is single-process) */
int apop_b200_comm_unique_id(void *id128);
int apop_b200_comm_init(int nranks, int rank, const void *id128);
int apop_b200_comm_finalize(void);
int apop_b200_comm_size(void);
/* Fused allreduce over NVLink peer memory: after apop_b200_comm_init every rank exports a
 * 64-byte CUDA IPC handle of its inbox (apop_b200_comm_p2p_handle), the host program gathers
 * the nranks handles (rank order) and hands them to apop_b200_comm_p2p_init.  The reduction
 * kernels then sum [LL | score] across GPUs in their own last CTA (peer stores + system-scope
 * flags) instead of a separate ncclAllReduce; NCCL remains for the large batched outputs. */
int apop_b200_comm_p2p_handle(int nranks, void *handle64);
int apop_b200_comm_p2p_init(int nranks, int rank, const void *handles);
int apop_b200_comm_p2p_active(void);
/* Inside a multi-rank job, evaluate this rank's rows only (no exchange) while on: the parity probe of
 * bench.py compares the fused cross-GPU sum with an independent torch.distributed all-reduce of these
 * local results, and rank 0 times the single-GPU pass of the fixed-size problem with it. */
void apop_b200_set_local_only(int on);

/* ------------------------------------------------------------------ *
 *  Residency: X, y (and weights) live in HBM between optimizer iterations
 * ------------------------------------------------------------------ */
/* Copy d->matrix / d->vector / d->weights to the device in the tiled layout
 * described in DESIGN.md and remember it under the address of d.  Call after
 * the model's prep (prep mutates the data: ols_shuffle, model/apop_ols.c:97), or let
 * apop_b200_prep_outcome do both.  Pinning a data set that is already resident and has
 * not been invalidated is a no-op. */
int apop_b200_pin(apop_data *d);
int apop_b200_unpin(apop_data *d);
/* The host copy changed behind the same pointer (apop_bootstrap.m4.c:143-149):
 * re-upload on next use. */
int apop_b200_invalidate(apop_data *d);
int apop_b200_is_pinned(const apop_data *d);
/* How the most recent upload crossed PCIe:
 *   0  staged: host threads gather the caller's (pageable, possibly strided) buffers into pinned staging,
 *      one cudaMemcpyAsync per chunk, tiling on the device -- the default for pageable memory;
 *   2  zero copy from a page-locked source: the caller's matrix was allocated with apop_b200_host_alloc /
 *      cudaHostAlloc (or registered by the caller), so the DMA engine reads it in place, one cudaMemcpyAsync per
 *      128 MB window -- chosen automatically, and the way to feed several GPUs from one host;
 *   1  zero copy from pageable memory page-locked by the library window by window (APOP_B200_PIN_MODE=register
 *      only; the pages stay locked while the data set is resident).  Measured slower than staging on the pool
 *      boxes, see DESIGN.md.
 * register_GBps (may be NULL) receives the page-locking rate measured in mode 1. */
int apop_b200_last_pin_mode(double *register_GBps);
/* Page-locked host memory for the data blocks of an apop_data (gsl_block.data): what gsl_matrix_alloc would
 * malloc, allocated with cudaHostAlloc instead, so that apop_b200_pin needs no host-side copy at all. */
void *apop_b200_host_alloc(size_t bytes);
void apop_b200_host_free(void *p);
/* unpinned data policy: 0 = upload per call and drop (always correct, default),
 * 1 = pin on first sight and trust the caller to invalidate. */
void apop_b200_set_auto_pin(int on);
/* Rows of the whole data set behind d over all ranks (= its own rows on one GPU); 0 if d is not
 * resident.  The information criteria of a sharded fit count these (apop_mle.m4.c:378-409). */
double apop_b200_global_rows(apop_data *d);

/* Synthetic, device-generated data sets in the resident layout (SURVEY 8(d),
 * recipe of tests/test_apop.c:889-907).  The returned apop_data is a host
 * header whose matrix/vector carry sizes but NULL data pointers; it is pinned.
 * beta_true has k*(ncat-1) doubles in pack order.  Rows are numbered from
 * row_offset so that shards of one global data set can be generated per rank.
 * Free with apop_b200_synth_free. */
apop_data *apop_b200_synth(apop_b200_family family, size_t n_rows, size_t k, int ncat,
                           const double *beta_true, uint64_t seed, uint64_t row_offset);
void apop_b200_synth_free(apop_data *d);
/* Copy a resident data set back into caller-provided host buffers in the
 * reference's layout (row-major X with tda = k; y; w may be NULL). */
int apop_b200_download(const apop_data *d, double *X_rowmajor, double *y, double *w);

/* ------------------------------------------------------------------ *
 *  The path: log likelihood and score, reference signatures
 *  Column counts: probit / binary logit / Poisson regression / OLS take 1..1024 columns
 *  (1..48 on the register-resident kernel, wider data on a kernel whose lanes tile rows x columns);
 *  multinomial logit, the information matrix and the batched path take 1..32 and up to
 *  8 outcome categories.  Anything else is refused with NaN and model->error = 'd'.
 *  A weights column is read by OLS only; probit, logit and the Poisson regression ignore it,
 *  as the reference's do.  OLS serves every branch of ols_log_likelihood (model/apop_ols.c:129-172):
 *  un-prepped data (outcome still in column 0), the <Predicted> and <Error variance> pages, and
 *  the input_distribution of apop_lm_settings.
 * ------------------------------------------------------------------ */
long double apop_b200_probit_ll(apop_data *d, apop_model *m);       /* replaces multiprobit_log_likelihood, model/apop_probit.c:104 */
void apop_b200_probit_score(apop_data *d, gsl_vector *g, apop_model *m); /* replaces probit_dlog_likelihood, :130 */
long double apop_b200_logit_ll(apop_data *d, apop_model *m);        /* replaces multilogit_log_likelihood, :231 */
void apop_b200_logit_score(apop_data *d, gsl_vector *g, apop_model *m);  /* new (reference has it commented out, :244-289) */
long double apop_b200_poisson_ll(apop_data *d, apop_model *m);      /* replaces poisson_log_likelihood, model/apop_poisson.c:21 */
void apop_b200_poisson_score(apop_data *d, gsl_vector *g, apop_model *m); /* replaces poisson_dlog_likelihood, :66 */
long double apop_b200_normal_ll(apop_data *d, apop_model *m);       /* replaces normal_log_likelihood, model/apop_normal.c:42 */
void apop_b200_normal_score(apop_data *d, gsl_vector *g, apop_model *m);  /* replaces normal_dlog_likelihood, :108 */
long double apop_b200_ols_ll(apop_data *d, apop_model *m);          /* replaces ols_log_likelihood, model/apop_ols.c:129 */
void apop_b200_ols_score(apop_data *d, gsl_vector *g, apop_model *m);     /* replaces ols_score, :175 */
/* the next iid distributions on the same apop_map_sum shape */
long double apop_b200_exponential_ll(apop_data *d, apop_model *m);      /* replaces exponential_log_likelihood, model/apop_exponential.c:33 */
void apop_b200_exponential_score(apop_data *d, gsl_vector *g, apop_model *m);  /* :42 */
long double apop_b200_gamma_ll(apop_data *d, apop_model *m);            /* replaces gamma_log_likelihood, model/apop_gamma.c:36 */
void apop_b200_gamma_score(apop_data *d, gsl_vector *g, apop_model *m);        /* :56 */
long double apop_b200_lognormal_ll(apop_data *d, apop_model *m);        /* replaces lognormal_log_likelihood, model/apop_normal.c:170 */
void apop_b200_lognormal_score(apop_data *d, gsl_vector *g, apop_model *m);    /* :228 */
long double apop_b200_bernoulli_ll(apop_data *d, apop_model *m);        /* replaces bernoulli_log_likelihood, model/apop_bernoulli.c:18 */
void apop_b200_bernoulli_score(apop_data *d, gsl_vector *g, apop_model *m);    /* new: the reference registers none (numeric gradient) */
long double apop_b200_beta_ll(apop_data *d, apop_model *m);             /* replaces beta_log_likelihood, model/apop_beta.c:57 */
void apop_b200_beta_score(apop_data *d, gsl_vector *g, apop_model *m);         /* replaces beta_dlog_likelihood, :67 */
long double apop_b200_zipf_ll(apop_data *d, apop_model *m);             /* replaces zipf_log_likelihood, model/apop_zipf.c:31 (no score in the reference) */
long double apop_b200_dirichlet_ll(apop_data *d, apop_model *m);        /* replaces dirichlet_log_likelihood, model/apop_dirichlet.c:19 */
void apop_b200_dirichlet_score(apop_data *d, gsl_vector *g, apop_model *m);    /* replaces dirichlet_dlog_likelihood, :28 */
long double apop_b200_t_ll(apop_data *d, apop_model *m);                /* replaces apop_tdist_llike, model/apop_t.c:32 (no score in the reference) */
long double apop_b200_yule_ll(apop_data *d, apop_model *m);             /* replaces yule_log_likelihood, model/apop_yule.c:40 */
void apop_b200_yule_score(apop_data *d, gsl_vector *g, apop_model *m);         /* replaces yule_dlog_likelihood, :50 */
long double apop_b200_poisson_reg_ll(apop_data *d, apop_model *m);  /* new model behind the same API */
void apop_b200_poisson_reg_score(apop_data *d, gsl_vector *g, apop_model *m);

/* The O(N) steps of the models' prep routines on the device, in one call:
 *  - makes d resident; if d->vector is NULL, column 0 of d->matrix is the outcome and the tiling kernel
 *    performs ols_shuffle (model/apop_ols.c:97-108): resident column 0 := 1, outcome -> the y column;
 *  - if cats/ncat are given: get_category_table / apop_data_to_factors (model/apop_probit.c:32-40,
 *    apop_regression.m4.c:423-450) -- the sorted distinct outcome values (at most max_cats <= 64, all
 *    ranks) and the resident outcomes recoded to 0..C-1;
 *  - if start is given: probit_prep's start point exp(mean ln|x_ij|) (model/apop_probit.c:65-80), k doubles;
 *  - sync_host != 0 mirrors the reference's in-place mutation of the caller's data (vector allocated
 *    the way gsl_vector_alloc does, filled with the recoded outcomes; column 0 := 1).
 * Returns 0 on success. */
int apop_b200_prep_outcome(apop_data *d, double *cats, size_t max_cats, size_t *ncat, double *start, int sync_host);
/* multilogit_expected (model/apop_probit.c:166-199): category probabilities (n x C, row-major; may be
 * NULL) and the most likely category index per row (n; may be NULL); X.B is formed on the device. */
int apop_b200_logit_expected(apop_data *d, apop_model *m, double *probs, double *best);

/* Ingest: a delimited numeric text file straight into the resident layout,
 * replacing apop_text_to_data (apop_conversions.m4.c:600-690) + apop_b200_pin without the apop_data
 * detour.  The file is mmap'ed, cut at line boundaries, parsed by the host threads into pinned staging
 * and tiled on the device group by group.  Rules of the reference's text format
 * (apop_conversions.m4.c:339-403) that are honoured: delimiter set (NULL = "|,\t"), '#' comments, blank
 * lines, white space around fields, has_col_names (1/'y': the first non-blank line is skipped),
 * has_row_names (1/'y': the first field of every row is skipped), an empty field = NaN unless the
 * delimiter closing it is a space or tab, strtod number syntax, NaN for a field strtod cannot convert (what the
 * reference's code does, :659-666; its format page says zeros, :363), quoted fields ("..." read verbatim with the
 * marks stripped) and backslash escapes.  Fixed-width layouts: the _fixed forms below.  outcome_col0 != 0: column 0 is the outcome
 * (ols_shuffle while tiling).  Returns a device-only header (sizes set, data pointers NULL) that is
 * resident; free it with apop_b200_synth_free.  NULL on error. */
apop_data *apop_b200_text_to_resident(const char *path, int has_row_names, int has_col_names,
                                      const char *delimiters, int outcome_col0);
/* The same reader into host memory (what apop_text_to_data leaves in ->matrix): *out is a malloc'ed
 * row-major n x k buffer owned by the caller.  Needs no device. */
int apop_b200_text_to_host(const char *path, int has_row_names, int has_col_names, const char *delimiters,
                           double **out, size_t *n_rows, size_t *k);
/* Fixed-width files: apop_text_to_data(.field_ends = ...), parse_a_fixed_line (apop_conversions.m4.c:434-464).
 * field_ends[i] is the 1-based position of the last character of field i; characters after the last given end form
 * one more field; no comments, quotes or delimiters; a line without characters is blank; each field converts like
 * any other (strtod skips leading white space; empty or unconvertible = NaN). */
apop_data *apop_b200_text_to_resident_fixed(const char *path, int has_row_names, int has_col_names,
                                            const int *field_ends, int n_field_ends, int outcome_col0);
int apop_b200_text_to_host_fixed(const char *path, int has_row_names, int has_col_names, const int *field_ends,
                                 int n_field_ends, double **out, size_t *n_rows, size_t *k);

/* Row reductions the models are built from (apop_map_sum / apop_matrix_sum,
 * apop_mapply.m4.c:485, apop_stats.m4.c:15,310) over a resident data set:
 * sum over every cell of vector+matrix of f(x; param). */
typedef enum {
    APOP_B200_MAP_IDENTITY = 0,    /* x                         (apop_matrix_sum)        */
    APOP_B200_MAP_DEV = 1,         /* x - p0                    (normal apply_me)        */
    APOP_B200_MAP_DEV2 = 2,        /* (x - p0)^2                (normal apply_me2)       */
    APOP_B200_MAP_POISSON = 3,     /* x==0?0: p0*x - lgamma(x+1), -inf if invalid (poisson apply_me) */
    APOP_B200_MAP_LOG = 4,         /* log(x)                                              */
    APOP_B200_MAP_EXP = 5,         /* exp(x)                                              */
    APOP_B200_MAP_ABS = 6,         /* |x|                       (apop_map only)           */
    APOP_B200_MAP_SQRT = 7,        /* sqrt(x)                   (apop_map only)           */
    APOP_B200_MAP_SCALE = 8,       /* p0 * x                    (apop_map only)           */
    APOP_B200_MAP_POW = 9          /* x^p0                      (apop_map only)           */
} apop_b200_map_fn;
/* part: 'a' all (vector and matrix), 'v' vector only, 'm' matrix only */
double apop_b200_map_sum(apop_data *d, apop_b200_map_fn fn, double p0, char part);

/* apop_map with a materialised result (apop_mapply.m4.c:119-217, output shapes :300-313).  The reference
 * applies a HOST function pointer per cell / row; a device library cannot, so the functions the models on
 * this path map with -- and the elementary ones -- are offered as enums.
 *   apop_b200_map       fn_d / fn_dp shapes, part 'a' | 'v' | 'm': the result has the shape of the input;
 *                       out_vector (n_rows) and out_matrix (n_rows x k row-major) are caller buffers, either
 *                       may be NULL; a part that is not mapped is copied through.
 *   apop_b200_map_rows  fn_v / fn_vp / fn_r / fn_rp shapes, part 'r': one double per row into out (n_rows).
 *                       param = k doubles for ROW_DOT and the per-row log-likelihood terms (the model's beta);
 *                       aux = NULL or the two category values of the binary logit. */
typedef enum {
    APOP_B200_ROW_SUM = 0,             /* sum_j x_ij                                              */
    APOP_B200_ROW_MEAN = 1,            /* mean_j x_ij                                             */
    APOP_B200_ROW_SUMSQ = 2,           /* sum_j x_ij^2                                            */
    APOP_B200_ROW_MAX = 3,
    APOP_B200_ROW_MIN = 4,
    APOP_B200_ROW_DOT = 5,             /* x_i . param              (apop_dot, one column)         */
    APOP_B200_ROW_PROBIT_LL = 6,       /* biprobit_ll_row          model/apop_probit.c:83-88      */
    APOP_B200_ROW_LOGIT2_LL = 7,       /* one_logit_row, C = 2     model/apop_probit.c:212-229    */
    APOP_B200_ROW_POISSON_REG_LL = 8,  /* y x.b - exp(x.b) - lgamma(y+1)              */
    APOP_B200_ROW_OLS_ERROR = 9        /* x_i . param - y_i        model/apop_ols.c:146-155       */
} apop_b200_row_fn;
int apop_b200_map(apop_data *d, apop_b200_map_fn fn, double p0, char part, double *out_vector, double *out_matrix);
int apop_b200_map_rows(apop_data *d, apop_b200_row_fn fn, const double *param, size_t nparam, const double *aux, double *out);

/* apop_dot(d, parameters) (apop_linear_algebra.m4.c:424-448): the N x c index X.B, row-major
 * into a caller buffer; B is k x c row-major (parameters->matrix).  On a dimension mismatch
 * d->error = 'd' like the reference.  This is also ols_predict / the index of
 * multilogit_expected (model/apop_ols.c:359-366, model/apop_probit.c:166-199). */
int apop_b200_dot(apop_data *d, const double *B, size_t k, size_t c, double *out);

/* Prep reductions on the resident data (the O(N k) steps either side of the path):
 * per-column sums, mode 0 = ln|x| over non-zero cells, 1 = x, 2 = x^2; out has k entries.
 * apop_b200_probit_start is probit_prep's start point exp(mean ln|x_ij|)
 * (model/apop_probit.c:65-80). */
int apop_b200_col_sums(apop_data *d, int mode, double *out);
int apop_b200_probit_start(apop_data *d, double *beta0);

/* OLS normal equations X'X (k*k, row-major) and X'y (k) in one pass
 * (apop_estimate_OLS, model/apop_ols.c:333-334). */
int apop_b200_ols_gram(apop_data *d, double *xtx, double *xty);

/* Fused value+gradient in one call (what fdf_shell, apop_mle.m4.c:362, wants):
 * beta in pack order (P doubles), score may be NULL. Returns LL. */
double apop_b200_ll_score(apop_b200_family family, apop_data *d, const double *beta, size_t P,
                          double *score);

/* Observed information  -d2LL/dbeta dbeta' (P*P row-major, pack order) on the FP64 tensor path
 * (replaces the numeric Hessian, apop_mle.m4.c:187-223).  PROBIT, LOGIT, POISSON_REG, OLS: one pass with
 * P = k; multinomial LOGIT with P = k (C-1): sum_i x_i x_i' (x) (diag p_i - p_i p_i') from ceil((C-1) C / 8)
 * passes.  Families without an analytic form get the reference's own numeric Hessian over the device
 * score from the host side (apop_model_hessian in the host mirror).  Returns 0 on success. */
int apop_b200_information(apop_b200_family family, apop_data *d, const double *beta, size_t P,
                          double *H);

/* Batched parameter vectors (bootstrap replicates, parallel chains): B is
 * P x m column-major-by-replicate (replicate r at B + r*P); counts is NULL or an
 * n_rows x m row-major uint8 matrix of resample multiplicities already resident
 * (see apop_b200_bootstrap_counts); ll has m, score has P*m (NULL to skip).
 * PROBIT, binary LOGIT and POISSON_REG (with counts, the Poisson constant sum c_i lgamma(y_i+1) is carried per row).
 * (New entry; the unmodified callers are serial: apop_bootstrap.m4.c:143,
 * apop_mcmc.m4.c:198.) */
int apop_b200_ll_score_batched(apop_b200_family family, apop_data *d, const double *B, size_t P,
                               size_t m, int use_counts, double *ll, double *score);
/* The m vectors of a bootstrap or of parallel chains lie within a few standard errors of each other.  The batched
 * kernel exploits that: the family's row term is expanded to degree 6 around x_i . mean(B) once per row and every
 * (row, replicate) pair with |x_i . (beta_r - mean)| <= 1.5e-2 costs a Horner evaluation (truncation < 1e-16
 * relative) instead of a link evaluation; pairs outside the radius, and rows near a clamp of the reference formula,
 * are evaluated exactly, so results never depend on the closeness.  This returns the share of warp steps of the last
 * pass over d that needed the exact evaluation (above 0.25 the next 32 passes use the exact kernel outright);
 * APOP_B200_BATCH_TAYLOR=0 switches the expansion off, APOP_B200_BATCH_DELTA sets the radius. */
double apop_b200_batched_exact_fraction(const apop_data *d);
/* Families outside the batched kernel (the multinomial logit) take their bootstrap replicates one at a time: with a
 * replicate selected, apop_b200_logit_ll / _score weight every row by that replicate's resident resample count -- the
 * LL and score of the resampled data set of apop_bootstrap.m4.c:143-149 without copying a row -- so the unmodified
 * optimizer fits replicate after replicate on the same resident data.  replicate < 0 switches it off. */
int apop_b200_set_replicate(apop_data *d, long replicate);
/* draw m multinomial resamples of the rows into a resident n x m count matrix */
int apop_b200_bootstrap_counts(apop_data *d, size_t m, uint64_t seed);
/* the resident counts as an n_rows x m row-major uint8 matrix (for checking and for host-side drivers) */
int apop_b200_counts_download(const apop_data *d, size_t m, unsigned char *out);

/* ------------------------------------------------------------------ *
 *  Installing into the reference's dispatch
 * ------------------------------------------------------------------ */
/* Sets model->log_likelihood to the device entry of `family` and registers the
 * device score through apop_vtable_add("apop_score", fn, hash) -- the symbol is
 * looked up in the running process (libapophenia, or this repo's host mirror).
 * The model's own prep then finds the hash present and leaves it
 * (apop_vtables.c:98). */
int apop_b200_register(apop_model *model, apop_b200_family family);
int apop_b200_unregister(apop_model *model);

#ifdef __cplusplus
}
#endif
#endif
