// capi_prep.cu -- the O(N) steps of probit_prep / ols_prep on the device:
// ols_shuffle while the data are being tiled, the category table from the resident outcome
// column, the recoding of the outcomes, and the start point.  One call, after which the host
// prep only has the small things left to do (pages, parameter shapes, settings).
#include "capi_internal.cuh"
#include "host_pool.h"
#include <algorithm>

using namespace apb;

static double key_to_double(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    double v;
    memcpy(&v, &b, 8);
    return v;
}

// sorted distinct values of the resident outcome column, all ranks; false if more than max_cats
static bool device_categories(Resident* r, size_t max_cats, std::vector<double>& cats) {
    Runtime& R = rt();
    cats.clear();
    unsigned long long* d_key = reinterpret_cast<unsigned long long*>(R.d_vals);
    const int grid = grid_for(8, r->n_tiles, 8);
    const int ycol = r->k;
    double above = 0.0;
    for (size_t c = 0; c <= max_cats; c++) {
        unsigned long long key = ~0ull;
        if (!cuda_ok(cudaMemcpyAsync(d_key, &key, 8, cudaMemcpyHostToDevice, R.stream), "H2D key")) return false;
        if (r->n_tiles &&
            !cuda_ok(next_distinct_launch(r->tiles, r->n_rows, r->n_tiles, ycol, r->cols, above, c == 0, d_key, grid, R.stream), "next_distinct launch")) return false;
        R.launches++;
        if (!cuda_ok(cudaMemcpyAsync(&key, d_key, 8, cudaMemcpyDeviceToHost, R.stream), "D2H key") ||
            !cuda_ok(cudaStreamSynchronize(R.stream), "prep sync")) {
            // leaving a loop of collectives on a local error would desynchronise the ranks: poison the
            // communicator so that this rank fails loudly from here on (its peers time out and do the same)
            if (R.nranks > 1) R.comm_failed = true;
            return false;
        }
        // the smallest candidate over the ranks: every rank publishes its own in its slot of a sum
        double mine = (key == ~0ull) ? INFINITY : key_to_double(key);
        if (R.nranks > 1) {
            std::vector<double> slots(2 * (size_t)R.nranks, 0.0);
            slots[2 * R.rank] = std::isinf(mine) ? 0.0 : mine;
            slots[2 * R.rank + 1] = std::isinf(mine) ? 0.0 : 1.0;
            if (!allreduce_host(slots.data(), (int)slots.size())) return false;
            mine = INFINITY;
            for (int p = 0; p < R.nranks; p++)
                if (slots[2 * p + 1] != 0.0 && slots[2 * p] < mine) mine = slots[2 * p];
        }
        if (std::isinf(mine)) return true;          // no value above the last category: done
        if (c == max_cats) return set_error("the outcome column has more than %zu distinct values", max_cats);
        cats.push_back(mine);
        above = mine;
    }
    return true;
}

static int prep_outcome_locked(apop_data* d, double* cats_out, size_t max_cats, size_t* ncat, double* start, int sync_host);
extern "C" int apop_b200_prep_outcome(apop_data* d, double* cats_out, size_t max_cats, size_t* ncat, double* start, int sync_host) {
    LOCK;
    const int rc = prep_outcome_locked(d, cats_out, max_cats, ncat, start, sync_host);
    if (rc != 0 && d) {                     // never leave a half-prepared resident copy of HOST data behind
        auto it = rt().reg.find(d);
        if (it != rt().reg.end() && !it->second->synthetic) apop_b200_unpin(d);
    }
    return rc;
}
static int prep_outcome_locked(apop_data* d, double* cats_out, size_t max_cats, size_t* ncat, double* start, int sync_host) {
    static const bool trace = getenv("APOP_B200_PIN_TRACE") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!trace) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[apop_b200 prep] %-28s %.3f s\n", what, std::chrono::duration<double>(now - t_prev).count());
        t_prev = now;
    };
    if (!d || !d->matrix) { set_error("apop_b200_prep_outcome: NULL data or no matrix"); return 1; }
    if (sync_host && d->vector && d->vector->stride != 1) { set_error("apop_b200_prep_outcome: sync_host needs a contiguous outcome vector (stride 1)"); return 1; }
    if (!require_ready()) return 1;
    Runtime& R = rt();
    cudaSetDevice(R.device);
    if (max_cats > (size_t)PREP_MAX_CATS) max_cats = PREP_MAX_CATS;
    // ---- residency, with ols_shuffle (model/apop_ols.c:97-108) done by the tiling kernel
    auto it = R.reg.find(d);
    Resident* r = (it != R.reg.end()) ? it->second : new Resident();
    if (r->synthetic) {
        // device-only data (apop_b200_text_to_resident with the outcome in column 0, apop_b200_synth):
        // already tiled and shuffled, and there is no host copy to mirror anything into
        if (!r->has_y) { set_error("apop_b200_prep_outcome: the device-only data set has no outcome column (ingest it with outcome_col0)"); return 1; }
        sync_host = 0;
        if (!ensure_layout(r, false)) return 1;   // the category kernels mask rows by position
    } else {
        r->y_col0 = (d->vector == nullptr);
        r->recoded = false;   // a fresh prep reads the raw outcomes of the host copy
        if (!upload(r, d)) {
            free_resident(r);
            if (it != R.reg.end()) R.reg.erase(it);
            d->error = 'a';
            return 1;
        }
        R.reg[d] = r;
    }
    lap("upload + tiling");
    const bool shuffled = r->y_col0;
    // ---- category table (get_category_table / apop_data_to_factors) and recoding
    std::vector<double> cats;
    if (cats_out || ncat) {
        if (!device_categories(r, max_cats ? max_cats : PREP_MAX_CATS, cats)) return 1;
        CatTable tab;
        memset(&tab, 0, sizeof tab);
        tab.n = (int)cats.size();
        for (size_t c = 0; c < cats.size(); c++) tab.v[c] = cats[c];
        if (!cuda_ok(recode_launch(r->tiles, r->n_rows, r->n_tiles, r->k, r->cols, tab, R.stream), "recode launch")) return 1;
        R.launches++;
        r->recoded = true; r->recode_pending = !sync_host; r->cats = tab;
        r->memo_valid = false;
        if (ncat) *ncat = cats.size();
        if (cats_out) for (size_t c = 0; c < cats.size(); c++) cats_out[c] = cats[c];
    }
    lap("category table + recode");
    // ---- start point of probit_prep (model/apop_probit.c:65-80) from the shuffled columns
    if (start && apop_b200_probit_start(d, start) != 0) return 1;
    lap("start point");
    // ---- the reference's prep mutates the caller's data: mirror that on request
    if (sync_host) {
        const size_t n = r->n_rows;
        if (shuffled) {
            gsl_block* b = (gsl_block*)malloc(sizeof(gsl_block));          // as gsl_vector_alloc lays it out
            gsl_vector* v = (gsl_vector*)malloc(sizeof(gsl_vector));
            if (!b || !v) { set_error("out of host memory"); return 1; }
            b->size = n; b->data = (double*)malloc(sizeof(double) * (n ? n : 1));
            v->size = n; v->stride = 1; v->data = b->data; v->block = b; v->owner = 1;
            d->vector = v;
            if (b->data) {   // first touch by the host pool: the driver's pageable copy below is slow on untouched pages
                const size_t pieces = (size_t)HostPool::get().size() * 4;
                double* y = b->data;
                HostPool::get().run(pieces, [&](size_t p) { memset(y + n * p / pieces, 0, sizeof(double) * (n * (p + 1) / pieces - n * p / pieces)); });
            }
        }
        lap("host vector allocation");
        if (apop_b200_download(d, nullptr, d->vector->data, nullptr) != 0) return 1;   // recoded outcomes (stride 1 only)
        lap("outcome download");
        if (shuffled) {   // column 0 := 1 is a strided write over the whole matrix: spread it over the host pool
            const size_t pieces = (size_t)HostPool::get().size() * 4;
            double* M = d->matrix->data;
            const size_t tda = d->matrix->tda;
            HostPool::get().run(pieces, [&](size_t p) { for (size_t i = n * p / pieces; i < n * (p + 1) / pieces; i++) M[i * tda] = 1.0; });
        }
        lap("host column 0 := 1");
        r->y_col0 = false;   // the host copy now has the reference's post-prep shape
        r->sig = host_sig(d);  // ... and that shape is what the resident copy corresponds to
    }
    return 0;
}

// multilogit_expected (model/apop_probit.c:166-199): per row the C category probabilities of the fitted
// model and the most likely category.  probs is n x C row-major (may be NULL), best has n entries (may be NULL).
extern "C" int apop_b200_logit_expected(apop_data* d, apop_model* m, double* probs, double* best) {
    LOCK;
    if (!d || !m || !m->parameters || !m->parameters->matrix) { set_error("apop_b200_logit_expected: NULL argument"); return 1; }
    const gsl_matrix* B = m->parameters->matrix;
    const size_t k = B->size1, c1 = B->size2;
    std::vector<double> Bp;
    if (!pack_parameters(m->parameters, Bp)) return 1;
    Resident* r = acquire(d);
    if (!r) return 1;
    const size_t n = r->n_rows;
    release(r);
    std::vector<double> xb(n * c1);
    if (apop_b200_dot(d, Bp.data(), k, c1, xb.data()) != 0) return 1;   // X.B on the device
    for (size_t i = 0; i < n; i++) {       // the soft-max over [0, xb_1..]: N x C work, written where the caller wants it
        long double mx = 0;
        for (size_t c = 0; c < c1; c++) mx = std::max<long double>(mx, xb[i * c1 + c]);
        long double den = expl(-mx);
        for (size_t c = 0; c < c1; c++) den += expl(xb[i * c1 + c] - mx);
        size_t arg = 0;
        long double pbest = expl(-mx) / den;
        if (probs) probs[i * (c1 + 1)] = (double)pbest;
        for (size_t c = 0; c < c1; c++) {
            const long double p = expl(xb[i * c1 + c] - mx) / den;
            if (probs) probs[i * (c1 + 1) + c + 1] = (double)p;
            if (p > pbest) { pbest = p; arg = c + 1; }
        }
        if (best) best[i] = (double)arg;
    }
    return 0;
}
