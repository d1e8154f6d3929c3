/* hostmini.c -- host-side mirror of the reference interface around the hot path.
 *
 * libapophenia cannot be built in this image, so the callers either
 * side of the device boundary are mirrored here with the reference's names,
 * argument meaning and error behaviour, so that the shim can be driven exactly
 * as the reference drives it:
 *
 *   apop_vtable_add/get/drop        apop_vtables.c:51-118 (hash-keyed registry)
 *   apop_data_pack / apop_data_unpack   apop_conversions.m4.c:691-806
 *   apop_log_likelihood, apop_score, apop_prep, apop_estimate, apop_model_copy
 *                                   apop_model.m4.c:155,229,257,290,447
 *   probit_prep / logit_prep / ols_shuffle / factor page
 *                                   model/apop_probit.c:32-81, model/apop_ols.c:97-122
 *   apop_numerical_gradient         apop_mle.m4.c:92-143 (gsl_deriv_central's 5-point rule)
 *
 * The model objects exported here (apop_probit, apop_logit, apop_poisson,
 * apop_normal, apop_ols) carry NO CPU log_likelihood: until apop_b200_register
 * installs the device entries they cannot be evaluated, and apop_estimate says so.
 * Everything O(N) on the path runs on the B200; what stays here is O(k) host
 * logic plus the one-off prep passes.
 *
 * Written from the behaviour described in SURVEY.md; no reference source text.
 */
#include "hostmini.h"
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- GSL-shaped containers ------------------------------------------------ */
gsl_vector *gsl_vector_alloc(size_t n) {
    gsl_vector *v = calloc(1, sizeof *v);
    gsl_block *b = calloc(1, sizeof *b);
    b->size = n;
    b->data = malloc(sizeof(double) * (n ? n : 1));
    v->size = n; v->stride = 1; v->data = b->data; v->block = b; v->owner = 1;
    return v;
}
gsl_vector *gsl_vector_calloc(size_t n) {
    gsl_vector *v = gsl_vector_alloc(n);
    memset(v->data, 0, sizeof(double) * n);
    return v;
}
void gsl_vector_free(gsl_vector *v) {
    if (!v) return;
    if (v->owner && v->block) { free(v->block->data); free(v->block); }
    free(v);
}
gsl_matrix *gsl_matrix_alloc(size_t n1, size_t n2) {
    gsl_matrix *m = calloc(1, sizeof *m);
    gsl_block *b = calloc(1, sizeof *b);
    b->size = n1 * n2;
    b->data = malloc(sizeof(double) * (b->size > 0 ? b->size : 1));
    m->size1 = n1; m->size2 = n2; m->tda = n2; m->data = b->data; m->block = b; m->owner = 1;
    return m;
}
gsl_matrix *gsl_matrix_calloc(size_t n1, size_t n2) {
    gsl_matrix *m = gsl_matrix_alloc(n1, n2);
    memset(m->data, 0, sizeof(double) * n1 * n2);
    return m;
}
void gsl_matrix_free(gsl_matrix *m) {
    if (!m) return;
    if (m->owner && m->block) { free(m->block->data); free(m->block); }
    free(m);
}

/* ---- apop_data -------------------------------------------------------------- */
/* apop_data_alloc(size1,size2,size3): (v), (m1,m2) or (v,m1,m2); apop_data.m4.c:32-68 */
apop_data *apop_data_alloc(size_t size1, size_t size2, int size3) {
    size_t vsize = 0, m1 = 0; int m2 = 0;
    if (size3) { vsize = size1; m1 = size2; m2 = size3; }
    else if (size2) { m1 = size1; m2 = (int)size2; }
    else vsize = size1;
    apop_data *d = calloc(1, sizeof *d);
    if (m1 > 0 && m2 > 0) d->matrix = gsl_matrix_alloc(m1, m2);
    if (vsize) d->vector = gsl_vector_alloc(vsize);
    d->names = calloc(1, sizeof(apop_name));
    return d;
}
apop_data *apop_data_calloc(size_t size1, size_t size2, int size3) {
    apop_data *d = apop_data_alloc(size1, size2, size3);
    if (d->matrix) memset(d->matrix->data, 0, sizeof(double) * d->matrix->size1 * d->matrix->size2);
    if (d->vector) memset(d->vector->data, 0, sizeof(double) * d->vector->size);
    return d;
}
void apop_data_free(apop_data *d) {
    if (!d) return;
    { /* a resident copy keyed by this address must die with it: the next apop_data malloc'ed at the same
       * address would otherwise be evaluated against stale HBM contents (and the copy would leak) */
        static int (*unpin)(apop_data *) = NULL;
        static int looked = 0;
        if (!looked) { unpin = (int (*)(apop_data *))dlsym(RTLD_DEFAULT, "apop_b200_unpin"); looked = unpin != NULL; }
        if (unpin) unpin(d);
    }
    apop_data_free(d->more);
    gsl_vector_free(d->vector);
    gsl_matrix_free(d->matrix);
    gsl_vector_free(d->weights);
    if (d->names) { free(d->names->title); free(d->names); }
    free(d);
}
apop_data *apop_data_add_page(apop_data *dataset, apop_data *newpage, const char *title) {
    if (!newpage) return NULL;
    if (!newpage->names) newpage->names = calloc(1, sizeof(apop_name));
    free(newpage->names->title);
    newpage->names->title = strdup(title);
    if (!dataset) return newpage;
    apop_data *t = dataset;
    while (t->more) t = t->more;
    t->more = newpage;
    return newpage;
}
apop_data *apop_data_get_page(const apop_data *data, const char *title) {
    /* exact (case-insensitive) title match down the ->more list, the data set itself included */
    for (const apop_data *p = data; p; p = p->more)
        if (p->names && p->names->title && !strcasecmp(p->names->title, title)) return (apop_data *)p;
    return NULL;
}

/* ---- vtables (apop_vtables.c) ------------------------------------------------ */
typedef struct { size_t hash; void *fn; } vt_elmt;
typedef struct { unsigned long name_hash; int ct; vt_elmt *e; } vt_table;
static vt_table *vt_list;
static int vt_ct;
static unsigned long djb(const char *s) { /* Bernstein's string hash, apop_vtables.c:35-40 */
    unsigned long h = 5381;
    for (char c; (c = *s++);) h = h * 33 + c;
    return h;
}
static vt_table *vt_find(const char *tab, int create) {
    const unsigned long h = djb(tab);
    for (int i = 0; i < vt_ct; i++)
        if (vt_list[i].name_hash == h) return vt_list + i;
    if (!create) return NULL;
    vt_list = realloc(vt_list, sizeof(vt_table) * (vt_ct + 1));
    vt_list[vt_ct] = (vt_table){.name_hash = h};
    return vt_list + vt_ct++;
}
int apop_vtable_add(char const *tabname, void *fn_in, unsigned long hash) {
    vt_table *v;
#pragma omp critical(apop_b200_vtable)
    {
        v = vt_find(tabname, 1);
        int present = 0;
        for (int i = 0; i < v->ct; i++)
            if (v->e[i].hash == hash) present = 1; /* already there: no-op (apop_vtables.c:98) */
        if (!present) {
            v->e = realloc(v->e, sizeof(vt_elmt) * (v->ct + 1));
            v->e[v->ct++] = (vt_elmt){.hash = hash, .fn = fn_in};
        }
    }
    return 0;
}
void *apop_vtable_get(char const *tabname, unsigned long hash) {
    void *out = NULL;
#pragma omp critical(apop_b200_vtable)
    {
        vt_table *v = vt_find(tabname, 0);
        if (v)
            for (int i = 0; i < v->ct; i++)
                if (v->e[i].hash == hash) { out = v->e[i].fn; break; }
    }
    return out;
}
int apop_vtable_drop(char const *tabname, unsigned long hash) { /* 0 = removed, 1 = not found */
    int rc = 1;
#pragma omp critical(apop_b200_vtable)
    {
        vt_table *v = vt_find(tabname, 0);
        if (v)
            for (int i = 0; i < v->ct; i++)
                if (v->e[i].hash == hash) {
                    memmove(v->e + i, v->e + i + 1, sizeof(vt_elmt) * (v->ct - i - 1));
                    v->ct--;
                    rc = 0;
                    break;
                }
    }
    return rc;
}

/* ---- pack / unpack (apop_conversions.m4.c:691-806): vector, matrix rows, weights,
 * then the non-<info> pages ------------------------------------------------- */
static int is_info_page(const apop_data *d) {
    const char *t = d && d->names ? d->names->title : NULL;
    size_t n = t ? strlen(t) : 0;
    return n >= 2 && t[0] == '<' && t[n - 1] == '>';
}
static size_t pack_size(const apop_data *in) {
    size_t n = 0;
    for (const apop_data *p = in; p; p = p->more) {
        if (p != in && is_info_page(p)) continue;
        n += (p->vector ? p->vector->size : 0) + (p->matrix ? p->matrix->size1 * p->matrix->size2 : 0) +
             (p->weights ? p->weights->size : 0);
    }
    return n;
}
gsl_vector *apop_data_pack(const apop_data *in, gsl_vector *out) {
    if (!in) return NULL;
    const size_t total = pack_size(in);
    if (!total) return NULL;
    if (out && out->size != total) return NULL;
    if (!out) out = gsl_vector_alloc(total);
    size_t o = 0;
    for (const apop_data *p = in; p; p = p->more) {
        if (p != in && is_info_page(p)) continue;
        if (p->vector) for (size_t i = 0; i < p->vector->size; i++) out->data[(o++) * out->stride] = p->vector->data[i * p->vector->stride];
        if (p->matrix)
            for (size_t i = 0; i < p->matrix->size1; i++)
                for (size_t j = 0; j < p->matrix->size2; j++) out->data[(o++) * out->stride] = p->matrix->data[i * p->matrix->tda + j];
        if (p->weights) for (size_t i = 0; i < p->weights->size; i++) out->data[(o++) * out->stride] = p->weights->data[i * p->weights->stride];
    }
    return out;
}
void apop_data_unpack(const gsl_vector *in, apop_data *d) {
    if (!d || !in) return;
    size_t o = 0;
    for (apop_data *p = d; p && o < in->size; p = p->more) {
        if (p != d && is_info_page(p)) continue;
        if (p->vector) for (size_t i = 0; i < p->vector->size && o < in->size; i++) p->vector->data[i * p->vector->stride] = in->data[(o++) * in->stride];
        if (p->matrix)
            for (size_t i = 0; i < p->matrix->size1; i++)
                for (size_t j = 0; j < p->matrix->size2 && o < in->size; j++) p->matrix->data[i * p->matrix->tda + j] = in->data[(o++) * in->stride];
        if (p->weights) for (size_t i = 0; i < p->weights->size && o < in->size; i++) p->weights->data[i * p->weights->stride] = in->data[(o++) * in->stride];
    }
}

/* ---- settings (a single typed group list per model; apop_settings.c) ----------- */
void *apop_settings_get_grp(apop_model *m, const char *type) {
    if (!m || !m->settings) return NULL;
    for (apop_settings_type *s = m->settings; s->name[0]; s++)
        if (!strcmp(s->name, type)) return s->setting_group;
    return NULL;
}
void *apop_settings_group_alloc(apop_model *m, const char *type, void *group, size_t bytes) {
    int n = 0;
    if (m->settings) for (apop_settings_type *s = m->settings; s->name[0]; s++) {
        if (!strcmp(s->name, type)) { free(s->setting_group); s->setting_group = malloc(bytes); memcpy(s->setting_group, group, bytes); return s->setting_group; }
        n++;
    }
    m->settings = realloc(m->settings, sizeof(apop_settings_type) * (n + 2));
    memset(m->settings + n, 0, sizeof(apop_settings_type) * 2);
    snprintf(m->settings[n].name, 101, "%s", type);
    m->settings[n].name_hash = bytes; /* the mirror keeps the group size here for copies */
    m->settings[n].setting_group = malloc(bytes);
    memcpy(m->settings[n].setting_group, group, bytes);
    return m->settings[n].setting_group;
}
apop_mle_settings apop_mle_settings_defaults(void) { /* apop_mle.m4.c:52-70 */
    return (apop_mle_settings){.starting_pt = NULL, .tolerance = 1e-5, .max_iterations = 5000, .method = "",
                               .verbose = 0, .step_size = 0.05, .delta = 1e-3, .relative_tolerance = 0};
}
apop_mle_settings *apop_mle_settings_add(apop_model *m, apop_mle_settings s) {
    return apop_settings_group_alloc(m, "apop_mle", &s, sizeof s);
}

/* ---- model dispatch (apop_model.m4.c) --------------------------------------------- */
apop_model *apop_model_copy(apop_model *in) { /* :155-178: struct copy keeps fn pointers, hence vtable hashes */
    if (!in) return NULL;
    apop_model *out = malloc(sizeof *out);
    *out = *in;
    out->settings = NULL;
    if (in->settings)
        for (apop_settings_type *s = in->settings; s->name[0]; s++)
            apop_settings_group_alloc(out, s->name, s->setting_group, s->name_hash);
    if (in->parameters) {
        const apop_data *p = in->parameters;
        out->parameters = apop_data_alloc(p->vector ? p->vector->size : 0, p->matrix ? p->matrix->size1 : 0, p->matrix ? (int)p->matrix->size2 : 0);
        gsl_vector *packed = apop_data_pack(p, NULL);
        if (packed) { apop_data_unpack(packed, out->parameters); gsl_vector_free(packed); }
    }
    out->info = NULL;
    return out;
}
void apop_model_free(apop_model *m) {
    if (!m) return;
    apop_data_free(m->parameters);
    apop_data_free(m->info);
    if (m->settings) { for (apop_settings_type *s = m->settings; s->name[0]; s++) free(s->setting_group); free(m->settings); }
    free(m);
}
/* apop_model_clear (default prep): attach data, size the parameters from vsize/msize (-1 = data columns) */
void apop_model_clear(apop_data *data, apop_model *m) {
    m->data = data;
    const int cols = data && data->matrix ? (int)data->matrix->size2 : 0;
    const int v = m->vsize == -1 ? cols : m->vsize, m1 = m->msize1 == -1 ? cols : m->msize1, m2 = m->msize2 == -1 ? cols : m->msize2;
    apop_data_free(m->parameters);
    m->parameters = apop_data_calloc(v, m1, m2);
    apop_data_free(m->info);
    m->info = apop_data_alloc(0, 0, 0);
    apop_data_add_page(NULL, m->info, "<Info>");
    m->error = 0;
}
void apop_prep(apop_data *d, apop_model *m) {
    if (m->prep) m->prep(d, m);
    else apop_model_clear(d, m);
}
double apop_log_likelihood(apop_data *d, apop_model *m) { /* :257-265 */
    if (!m) return NAN;
    if (m->log_likelihood) return m->log_likelihood(d, m);
    if (m->p) return log(m->p(d, m));
    fprintf(stderr, "apop_log_likelihood: the model has neither log_likelihood nor p (is the B200 entry registered?)\n");
    return NAN;
}
void apop_score(apop_data *d, gsl_vector *out, apop_model *m) { /* :290-301 */
    if (!m || !out) return;
    apop_score_type ms = (apop_score_type)apop_vtable_get("apop_score", apop_score_hash(m));
    if (ms) { ms(d, out, m); return; }
    gsl_vector *num = apop_numerical_gradient(d, m, 0);
    if (num) { for (size_t i = 0; i < out->size && i < num->size; i++) out->data[i * out->stride] = num->data[i]; gsl_vector_free(num); }
}

/* apop_numerical_gradient (apop_mle.m4.c:92-143): per parameter, gsl_deriv_central's
 * 5-point rule at step delta: r3 = (f(+h)-f(-h))/2, r5 = 4/3 (f(+h/2)-f(-h/2)) - r3/3, df = r5/h */
gsl_vector *apop_numerical_gradient(apop_data *data, apop_model *m, double delta) {
    if (!m || !m->parameters) return NULL;
    if (!delta) { apop_mle_settings *mp = apop_settings_get_grp(m, "apop_mle"); delta = mp ? mp->delta : 1e-3; }
    gsl_vector *beta = apop_data_pack(m->parameters, NULL);
    gsl_vector *out = gsl_vector_calloc(beta->size);
    gsl_vector *work = gsl_vector_alloc(beta->size);
    for (size_t j = 0; j < beta->size; j++) {
        double f[4]; const double at[4] = {-delta, -delta / 2, delta / 2, delta};
        for (int q = 0; q < 4; q++) {
            memcpy(work->data, beta->data, sizeof(double) * beta->size);
            work->data[j] += at[q];
            apop_data_unpack(work, m->parameters);
            f[q] = apop_log_likelihood(data, m) + (m->constraint ? m->constraint(data, m) : 0);
        }
        const double r3 = 0.5 * (f[3] - f[0]), r5 = (4.0 / 3.0) * (f[2] - f[1]) - (1.0 / 3.0) * r3;
        out->data[j] = r5 / delta;
    }
    apop_data_unpack(beta, m->parameters);
    gsl_vector_free(beta); gsl_vector_free(work);
    return out;
}

/* ---- prep routines ---------------------------------------------------------------- */
static int cmp_double(const void *a, const void *b) { double x = *(const double *)a, y = *(const double *)b; return (x > y) - (x < y); }

/* get_category_table (model/apop_probit.c:32-40): find the factor page, or turn the outcome
 * column (vector if present, else matrix column 0) into factor indices and record the sorted
 * distinct values on a "<categories for ...>" page (apop_data_to_factors). */
apop_data *apop_b200_category_table(apop_data *d) {
    for (apop_data *p = d->more; p; p = p->more)
        if (p->names && p->names->title && !strncmp(p->names->title, "<categories for", 15)) return p;
    const size_t n = d->vector ? d->vector->size : d->matrix->size1;
    double *vals = malloc(sizeof(double) * (n ? n : 1));
    double *col = d->vector ? d->vector->data : d->matrix->data;
    const size_t stride = d->vector ? d->vector->stride : d->matrix->tda;
    for (size_t i = 0; i < n; i++) vals[i] = col[i * stride];
    qsort(vals, n, sizeof(double), cmp_double);
    size_t u = 0;
    for (size_t i = 0; i < n; i++) if (!u || vals[i] != vals[u - 1]) vals[u++] = vals[i];
    apop_data *page = apop_data_alloc(u, 0, 0);
    memcpy(page->vector->data, vals, sizeof(double) * u);
    for (size_t i = 0; i < n; i++) { /* replace each value by its index among the sorted distinct values */
        size_t lo = 0, hi = u;
        const double v = col[i * stride];
        while (lo + 1 < hi) { size_t mid = (lo + hi) / 2; if (vals[mid] <= v) lo = mid; else hi = mid; }
        col[i * stride] = (double)lo;
    }
    for (size_t c = 0; c < u; c++) page->vector->data[c] = (double)c; /* the column now holds 0..C-1 */
    free(vals);
    apop_data_add_page(d, page, d->vector ? "<categories for vector>" : "<categories for column 0>");
    return page;
}
/* ols_shuffle (model/apop_ols.c:97-108): outcome column -> vector, column 0 := 1 */
void apop_b200_ols_shuffle(apop_data *d) {
    if (!d || d->vector || !d->matrix) return;
    d->vector = gsl_vector_alloc(d->matrix->size1);
    for (size_t i = 0; i < d->matrix->size1; i++) { d->vector->data[i] = d->matrix->data[i * d->matrix->tda]; d->matrix->data[i * d->matrix->tda] = 1.0; }
}
static void ols_prep(apop_data *d, apop_model *m) { /* model/apop_ols.c:110-122 */
    if (m->data && m->info) return;
    if (!d || (!d->vector && !d->matrix)) { m->error = 'd'; return; }
    apop_b200_ols_shuffle(d);
    apop_model_clear(d, m);
}
/* The O(N) steps of probit_prep on the device when libapop_b200 is in the process (SURVEY 8(f)
 * rank 2): ols_shuffle, the category table, the recoding and the start point come back from one
 * call; the host keeps the small things.  APOP_B200_HOST_PREP=1 forces the host path below. */
typedef int (*prep_outcome_fn)(apop_data *, double *, size_t, size_t *, double *, int);
static int probit_prep_on_device(apop_data *d, apop_model *m) {
    if (getenv("APOP_B200_HOST_PREP") || !d || !d->matrix || d->weights) return 0;
    for (apop_data *p = d->more; p; p = p->more)
        if (p->names && p->names->title && !strncmp(p->names->title, "<categories for", 15)) return 0;   /* already factorised */
    if (d->vector && d->vector->stride != 1) return 0;
    prep_outcome_fn dev = (prep_outcome_fn)dlsym(RTLD_DEFAULT, "apop_b200_prep_outcome");
    if (!dev) return 0;
    const int had_vector = d->vector != NULL;
    const size_t k = d->matrix->size2;
    double cats[64], *start = malloc(sizeof(double) * (k ? k : 1));
    size_t count = 0;
    if (dev(d, cats, 64, &count, start, 1) != 0) { free(start); return 0; }   /* e.g. no device: fall through to the host path */
    apop_data *page = apop_data_alloc(count, 0, 0);
    for (size_t c = 0; c < count; c++) page->vector->data[c] = (double)c;   /* the column now holds 0..C-1 */
    apop_data_add_page(d, page, had_vector ? "<categories for vector>" : "<categories for column 0>");
    apop_model_clear(d, m);
    apop_data_free(m->parameters);
    m->parameters = apop_data_alloc(k, count > 1 ? count - 1 : 1, 0);
    apop_mle_settings *sets = apop_settings_get_grp(m, "apop_mle");
    if (!(sets && sets->starting_pt)) {
        for (size_t j = 0; j < k; j++) {
            if (!isfinite(start[j])) { m->error = 'd'; free(start); return 1; }
            for (size_t c = 0; c < m->parameters->matrix->size2; c++) m->parameters->matrix->data[j * m->parameters->matrix->tda + c] = start[j];
        }
        if (!sets) { apop_mle_settings s = apop_mle_settings_defaults(); sets = apop_mle_settings_add(m, s); }
        gsl_vector *sp = apop_data_pack(m->parameters, NULL);
        sets->starting_pt = sp->data; /* same small leak as the reference */
    }
    free(start);
    return 1;
}
static void probit_prep(apop_data *d, apop_model *m) { /* model/apop_probit.c:42-81 */
    if (m->data && m->parameters) return;
    if (probit_prep_on_device(d, m)) return;
    apop_data *factors = apop_b200_category_table(d);
    ols_prep(d, m);
    const size_t count = factors->vector->size, k = d->matrix->size2;
    apop_data_free(m->parameters);
    m->parameters = apop_data_alloc(k, count > 1 ? count - 1 : 1, 0);
    apop_mle_settings *sets = apop_settings_get_grp(m, "apop_mle");
    if (sets && sets->starting_pt) return;
    /* start point of the same order of magnitude as the data: exp(mean ln|x|), zeros skipped */
    for (size_t j = 0; j < k; j++) {
        long double lt = 0;
        for (size_t i = 0; i < d->matrix->size1; i++) { const double v = d->matrix->data[i * d->matrix->tda + j]; lt += v ? logl(fabs(v)) : 0; }
        lt /= d->matrix->size1;
        if (!isfinite((double)lt)) { m->error = 'd'; return; }
        for (size_t c = 0; c < m->parameters->matrix->size2; c++) m->parameters->matrix->data[j * m->parameters->matrix->tda + c] = (double)expl(lt);
    }
    if (!sets) { apop_mle_settings s = apop_mle_settings_defaults(); sets = apop_mle_settings_add(m, s); }
    gsl_vector *start = apop_data_pack(m->parameters, NULL);
    sets->starting_pt = start->data; /* same small leak as the reference */
}

/* the exported model objects: prep + shapes only; log_likelihood comes from apop_b200_register */
static apop_model probit_obj = {.name = "Probit", .dsize = -1, .prep = probit_prep};
static apop_model logit_obj = {.name = "Logit", .dsize = -1, .prep = probit_prep};
static apop_model poisson_obj = {.name = "Poisson distribution", .vsize = 1, .dsize = 1};
static apop_model normal_obj = {.name = "Normal distribution", .vsize = 2, .dsize = 1};
static apop_model ols_obj = {.name = "Ordinary Least Squares", .vsize = -1, .dsize = -1, .prep = ols_prep};
static apop_model poisson_reg_obj = {.name = "Poisson regression", .vsize = -1, .dsize = -1, .prep = ols_prep};
static apop_model exponential_obj = {.name = "Exponential distribution", .vsize = 1, .dsize = 1};
static apop_model gamma_obj = {.name = "Gamma distribution", .vsize = 2, .dsize = 1};
static apop_model lognormal_obj = {.name = "Lognormal distribution", .vsize = 2, .dsize = 1};
static apop_model bernoulli_obj = {.name = "Bernoulli distribution", .vsize = 1, .dsize = 1};
static apop_model beta_obj = {.name = "Beta distribution", .vsize = 2, .dsize = 1};
static apop_model zipf_obj = {.name = "Zipf distribution", .vsize = 1, .dsize = 1};
static apop_model dirichlet_obj = {.name = "Dirichlet distribution", .vsize = -1, .dsize = -1};
static apop_model t_obj = {.name = "t distribution", .vsize = 3, .dsize = 1};
static apop_model yule_obj = {.name = "Yule distribution", .vsize = 1, .dsize = 1};
apop_model *apop_bernoulli = &bernoulli_obj, *apop_beta = &beta_obj, *apop_zipf = &zipf_obj, *apop_dirichlet = &dirichlet_obj,
           *apop_t_distribution = &t_obj, *apop_yule = &yule_obj;
apop_model *apop_exponential = &exponential_obj, *apop_gamma = &gamma_obj, *apop_lognormal = &lognormal_obj;
apop_model *apop_probit = &probit_obj, *apop_logit = &logit_obj, *apop_poisson = &poisson_obj,
           *apop_normal = &normal_obj, *apop_ols = &ols_obj, *apop_poisson_regression = &poisson_reg_obj;

apop_model *apop_estimate(apop_data *d, apop_model *m) { /* apop_model.m4.c:229-235 */
    apop_model *out = apop_model_copy(m);
    apop_prep(d, out);
    if (out->error) return out;
    if (out->estimate) out->estimate(d, out);
    else apop_maximum_likelihood(d, out);
    return out;
}
