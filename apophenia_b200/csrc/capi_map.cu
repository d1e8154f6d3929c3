// capi_map.cu -- apop_map with a materialised output (apop_mapply.m4.c:119-217) over a resident data set.
#include "capi_internal.cuh"
#include "map.cuh"

using namespace apb;

// Cell-wise map: the result has the shape of the input (the reference's mapply_core returns a new
// apop_data of the same shape for fn_d / fn_dp with part 'a', :300-313).  part 'a' = vector and matrix,
// 'v' = vector only, 'm' = matrix only; a part that is not mapped is copied through unchanged.
extern "C" int apop_b200_map(apop_data* d, apop_b200_map_fn fn, double p0, char part, double* out_vector, double* out_matrix) {
    LOCK;
    if (!d || (!out_vector && !out_matrix)) { set_error("apop_b200_map: NULL argument"); return 1; }
    if ((int)fn < 0 || (int)fn >= MAPC_COUNT) { set_error("apop_b200_map: unknown cell function %d", (int)fn); return 1; }
    if (part != 'a' && part != 'v' && part != 'm') { set_error("apop_b200_map: part must be 'a', 'v' or 'm' (rows: apop_b200_map_rows)"); return 1; }
    Resident* r = acquire(d);
    if (!r) return 1;
    Runtime& R = rt();
    bool ok = true;
    double *dt = nullptr, *Xs = nullptr, *ys = nullptr;
    const bool want_m = out_matrix && r->k > 0, want_v = out_vector && r->has_y;
    size_t chunk = ((size_t)8 << 20) / (size_t)(r->cols ? r->cols : 1) / TILE_ROWS * TILE_ROWS;   // ~64 MB of doubles
    if (chunk < TILE_ROWS) chunk = TILE_ROWS;
    const size_t ctiles = chunk / TILE_ROWS;
    if (r->n_rows) {
        ok = cuda_ok(cudaMalloc(&dt, ctiles * (size_t)r->cols * TILE_ROWS * sizeof(double)), "cudaMalloc(map tiles)");
        if (ok && want_m) ok = cuda_ok(cudaMalloc(&Xs, chunk * (size_t)r->k * sizeof(double)), "cudaMalloc(map staging)");
        if (ok && want_v) ok = cuda_ok(cudaMalloc(&ys, chunk * sizeof(double)), "cudaMalloc(map staging)");
    }
    for (size_t r0 = 0; ok && r0 < r->n_rows; r0 += chunk) {
        const size_t rows = (r->n_rows - r0 < chunk) ? r->n_rows - r0 : chunk;
        const size_t nt = (rows + TILE_ROWS - 1) / TILE_ROWS;
        const double* src = r->tiles + (r0 / TILE_ROWS) * (size_t)r->cols * TILE_ROWS;
        ok = cuda_ok(map_cells_launch(src, nt, r->k, r->has_y ? r->k : -1, r->cols, (int)fn, p0, part != 'v', part != 'm', dt, R.stream), "map launch") &&
             cuda_ok(tile_unpack(dt, rows, r->k, r->cols, r->has_y, r->has_w, Xs, ys, nullptr, R.stream), "tile_unpack");
        R.launches += 2;
        if (ok && Xs) ok = cuda_ok(cudaMemcpyAsync(out_matrix + r0 * r->k, Xs, rows * r->k * sizeof(double), cudaMemcpyDeviceToHost, R.stream), "D2H");
        if (ok && ys) ok = cuda_ok(cudaMemcpyAsync(out_vector + r0, ys, rows * sizeof(double), cudaMemcpyDeviceToHost, R.stream), "D2H");
        ok = ok && cuda_ok(cudaStreamSynchronize(R.stream), "map sync");
    }
    if (dt) cudaFree(dt);
    if (Xs) cudaFree(Xs);
    if (ys) cudaFree(ys);
    release(r);
    return ok ? 0 : 1;
}

// Row-wise map (part 'r' with fn_v / fn_vp / fn_r / fn_rp): one double per row into out[n_rows].
// param: k doubles for APOP_B200_ROW_DOT and the per-row log-likelihood terms (the model's beta);
// aux: NULL or the two category values of the binary logit.
extern "C" int apop_b200_map_rows(apop_data* d, apop_b200_row_fn fn, const double* param, size_t nparam, const double* aux, double* out) {
    LOCK;
    if (!d || !out) { set_error("apop_b200_map_rows: NULL argument"); return 1; }
    if ((int)fn < 0 || (int)fn >= MAPR_COUNT) { set_error("apop_b200_map_rows: unknown row function %d", (int)fn); return 1; }
    Resident* r = acquire(d);
    if (!r) return 1;
    Runtime& R = rt();
    bool ok = false;
    double *dp = nullptr, *dout = nullptr;
    do {
        const bool needs_param = fn >= APOP_B200_ROW_DOT;
        const bool needs_y = fn >= APOP_B200_ROW_PROBIT_LL;
        if (needs_param && (!param || nparam != (size_t)r->k)) {
            set_error("apop_b200_map_rows: the data has %d columns, the parameter vector %zu entries", r->k, nparam);
            d->error = 'd';
            break;
        }
        if (needs_y && !r->has_y) { set_error("apop_b200_map_rows: this row function needs an outcome vector"); break; }
        if (!r->n_rows) { ok = true; break; }
        size_t chunk = (size_t)16 << 20;
        if (chunk > r->n_rows) chunk = (r->n_rows + TILE_ROWS - 1) / TILE_ROWS * TILE_ROWS;
        if (!cuda_ok(cudaMalloc(&dout, chunk * sizeof(double)), "cudaMalloc(map rows)")) break;
        if (needs_param) {
            if (!cuda_ok(cudaMalloc(&dp, r->k * sizeof(double)), "cudaMalloc(param)")) break;
            if (!cuda_ok(cudaMemcpyAsync(dp, param, r->k * sizeof(double), cudaMemcpyHostToDevice, R.stream), "H2D param")) break;
        }
        const double a4[4] = {aux ? aux[0] : 0.0, aux ? aux[1] : 1.0, 0.0, 0.0};
        ok = true;
        for (size_t r0 = 0; ok && r0 < r->n_rows; r0 += chunk) {
            const size_t rows = (r->n_rows - r0 < chunk) ? r->n_rows - r0 : chunk;
            ok = cuda_ok(map_rows_launch(r->tiles, r0, rows, r->n_rows, r->k, r->has_y ? r->k : -1, r->cols, (int)fn, dp, a4, dout, R.stream), "map rows launch") &&
                 cuda_ok(cudaMemcpyAsync(out + r0, dout, rows * sizeof(double), cudaMemcpyDeviceToHost, R.stream), "D2H") &&
                 cuda_ok(cudaStreamSynchronize(R.stream), "map sync");
            R.launches++;
        }
    } while (0);
    if (dp) cudaFree(dp);
    if (dout) cudaFree(dout);
    release(r);
    return ok ? 0 : 1;
}

// rows of the whole (all ranks) data set behind d; 0 if d is not resident (the host mirror's BIC uses it)
extern "C" double apop_b200_global_rows(apop_data* d) {
    LOCK;
    Runtime& R = rt();
    if (!R.ready || !d) return 0.0;
    auto it = R.reg.find(d);
    if (it == R.reg.end()) return 0.0;
    cudaSetDevice(R.device);
    double n, mc, vc;
    return global_counts(it->second, n, mc, vc) ? n : 0.0;
}
