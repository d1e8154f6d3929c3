// capi_residency.cu -- keeping X, y, w in HBM between optimizer iterations: the pipelined
// upload + re-tiling (apop_b200_pin), the registry keyed by the apop_data address, synthetic
// device-generated data sets and their download.
#include "capi_internal.cuh"
#include "host_pool.h"
#include <condition_variable>
#include <algorithm>

using namespace apb;

// --------------------------------------------------------------------------
//  residency
// --------------------------------------------------------------------------
static void data_shape(const apop_data* d, size_t& n, int& k, bool& has_y, bool& has_w) {
    n = d->matrix ? d->matrix->size1 : (d->vector ? d->vector->size : 0);
    k = d->matrix ? (int)d->matrix->size2 : 0;
    has_y = d->vector != nullptr;
    has_w = d->weights != nullptr;
}

HostSig host_sig(const apop_data* d) {
    HostSig g;
    if (d->matrix) { g.m = d->matrix->data; g.m1 = d->matrix->size1; g.m2 = d->matrix->size2; g.tda = d->matrix->tda; }
    if (d->vector) { g.v = d->vector->data; g.vn = d->vector->size; g.vs = d->vector->stride; }
    if (d->weights) { g.w = d->weights->data; g.wn = d->weights->size; }
    return g;
}
// a registry entry whose host header no longer looks like what was uploaded (a new apop_data allocated at
// the address of a freed one, or the same one reshaped by a prep) must not be served from the stale copy
static bool sig_stale(const Resident* r, const apop_data* d) {
    return !r->synthetic && !(r->sig == host_sig(d));
}

// --------------------------------------------------------------------------
//  Zero-copy upload: registered windows of the caller's own matrix
// --------------------------------------------------------------------------
// The staged upload below copies every byte twice on the host (pageable -> pinned staging, then DMA reads
// the staging buffer): 3-4x host-memory traffic per byte, which is what limits 8 ranks pinning through one
// host (7.6 GB/s per GPU in round 1).  When the matrix is contiguous (tda == size2) the DMA engine can read
// the caller's buffer itself: a helper thread page-locks it window by window (cudaHostRegister, 128 MB
// windows on 2 MB boundaries, a few windows ahead of the copy), every window crosses PCIe with ONE
// cudaMemcpyAsync straight from the caller's memory into a device landing buffer, and the tiling kernel
// re-lays it (rows that straddle a window boundary are carried over device-side).  The outcome and weight
// vectors (3 % of the bytes) still go through a small pinned staging buffer.
// Two cases:
//  * the caller's matrix is page-locked already (allocated with apop_b200_host_alloc / cudaHostAlloc, or
//    registered by the caller): always taken, no page-locking here.  This is the multi-GPU deployment path.
//  * pageable memory, APOP_B200_PIN_MODE=register only: a helper thread page-locks the windows ahead of the copy
//    and they stay page-locked while the data set is resident (a registration cache, as MPI / UCX keep one;
//    released by unpin / invalidate / re-upload / finalize).  It is NOT the default, by measurement on the pool
//    boxes (tools/pin_probe2.py, pin_probe3.py, pin_modes.py; profiles/README.md): cudaHostRegister runs at 9-12
//    GB/s on 4 KB pages and 38-53 GB/s on transparent huge pages first touched by the CPU, cudaHostUnregister at
//    36 GB/s, both serialise with the copy submissions on the driver's lock, and -- decisive -- two ranks
//    page-locking at once drop to 14 GB/s each (7 GB/s on a second registration of the same pages), below what
//    the staged path gives them (26 GB/s each), while one rank alone stages at 41-48 GB/s against 38.
struct RegWindow { uintptr_t lo, hi; int state; bool ours; };   // state: 0 pending, 1 registered, 2 failed
struct Registrar {
    std::mutex m;
    std::condition_variable cv;
    std::vector<RegWindow> w;
    size_t consumed = 0;          // windows the main thread has issued copies for
    bool stop = false;
    int device = 0;
    int ahead = 3;
    double first_rate = 0.0;
    double t_register = 0.0;      // seconds inside cudaHostRegister (trace)
    void run() {
        cudaSetDevice(device);
        for (size_t next = 0; next < w.size(); next++) {
            {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [&] { return stop || next < consumed + ahead; });
                if (stop) break;
            }
            RegWindow& x = w[next];
            const auto t0 = std::chrono::steady_clock::now();
            cudaError_t e = cudaHostRegister((void*)x.lo, x.hi - x.lo, cudaHostRegisterDefault);
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            t_register += dt;
            int st = 1;
            bool ours = true;
            // (a range somebody else registered, possibly a stale registration of freed memory, is not trusted: staged path)
            if (e != cudaSuccess) { st = 2; cudaGetLastError(); }
            std::lock_guard<std::mutex> lk(m);
            x.state = st; x.ours = ours;
            // the rate the decision is taken on: the best of the first two windows (the very first call also pays
            // the driver's one-off set-up)
            if (next < 2) { const double rate = (st == 1 && ours && dt > 0) ? (x.hi - x.lo) * 1e-9 / dt : (st == 1 ? 1e9 : 0.0); if (rate > first_rate) first_rate = rate; }
            if (st == 2) for (size_t j = next + 1; j < w.size(); j++) w[j].state = 2;   // nothing beyond this window
            cv.notify_all();
            if (st == 2) break;
        }
    }
};

static int pin_mode_env() {   // 0 auto, 1 staged, 2 register (read per upload: tests switch it)
    const char* e = getenv("APOP_B200_PIN_MODE");
    return (e && !strcmp(e, "staged")) ? 1 : (e && !strcmp(e, "register")) ? 2 : 0;
}

// Uploads rows [0, done) through registered windows and returns `done`: all n rows, or a multiple of
// TILE_ROWS (possibly 0) from which the staged path carries on.  ok = false on a CUDA error.
static size_t upload_registered(Resident* r, const apop_data* d, bool col0, int cols, bool& ok) {
    Runtime& R = rt();
    ok = true;
    const int mode = pin_mode_env();
    const size_t n = r->n_rows;
    const int k = r->k;
    const bool hy = d->vector != nullptr, hw = d->weights != nullptr;
    if (mode == 1 || !d->matrix || k < 1 || d->matrix->tda != (size_t)k) return 0;
    if ((hy && d->vector->stride != 1) || (hw && d->weights->stride != 1)) return 0;
    const size_t rb = (size_t)k * sizeof(double), total = n * rb;
    if (total < ((size_t)8 << 20) && mode != 2) return 0;            // small data: the staged path is simpler and as fast
    // Is the caller's matrix page-locked already (cudaHostAlloc / apop_b200_host_alloc / its own cudaHostRegister)?
    // Then the DMA engine reads it as it is: no page-locking here, no helper thread, no host copy at all.
    auto pinned_ptr = [](const void* p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeHost;
    };
    const size_t PAGE = 4096, ALIGN = (size_t)2 << 20;
    size_t WB = (size_t)128 << 20;                                    // window: a multiple of the 2 MB huge-page size
    if (const char* e = getenv("APOP_B200_PIN_WINDOW_MB")) { const long mb = atol(e); if (mb >= 2) WB = (size_t)(mb / 2 * 2) << 20; }
    const uintptr_t base = (uintptr_t)d->matrix->data;
    // Only pages that lie entirely inside the matrix are page-locked / read in place: a first or last page shared
    // with a neighbouring allocation (two arrays side by side on the heap) would otherwise be registered twice.
    // The partial head and tail (< one page each) travel through a small pinned scratch.
    const uintptr_t lo = (base + PAGE - 1) & ~(uintptr_t)(PAGE - 1), hi = (base + total) & ~(uintptr_t)(PAGE - 1);
    if (hi <= lo + PAGE) return 0;
    bool pinned_src = pinned_ptr((const void*)base) && pinned_ptr((const void*)(base + total - 1));
    for (uintptr_t a = lo; pinned_src && a < hi; a += WB) pinned_src = pinned_ptr((const void*)a);   // every window start, too
    // pageable memory is page-locked here only on request (APOP_B200_PIN_MODE=register): measured on the pool
    // boxes it loses to the staged path for one rank and does not scale over the ranks of one host (see above)
    if (!pinned_src && mode != 2) return 0;
    Registrar G;
    G.device = R.device;
    if (const char* e = getenv("APOP_B200_PIN_AHEAD")) { const int a = atoi(e); if (a >= 2) G.ahead = a; }
    double t_wait_reg = 0.0, t_wait_dev = 0.0, t_issue = 0.0;
    auto since = [](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    for (uintptr_t a = lo; a < hi;) {
        uintptr_t b = (a == lo) ? ((lo & ~(uintptr_t)(ALIGN - 1)) + WB) : a + WB;
        if (b > hi) b = hi;
        G.w.push_back(RegWindow{a, b, 0, true});
        a = b;
    }
    const size_t nw = G.w.size();
    // landing buffers: [carry: up to 32 rows + a partial one][window][y][w]
    const size_t carry_cap = ((33 * rb) + 255) / 256 * 256;
    const size_t max_rows = (WB + 2 * PAGE) / rb + 66;
    const size_t yw_bytes = max_rows * sizeof(double) * ((hy ? 1 : 0) + (hw ? 1 : 0));
    const size_t dev_need = carry_cap + PAGE + WB + 2 * PAGE + yw_bytes + 512;
    if (!R.reg_edge && !cuda_ok(cudaHostAlloc(&R.reg_edge, 2 * PAGE, cudaHostAllocDefault), "cudaHostAlloc(edge scratch)")) { ok = false; return 0; }
    const size_t head_len = lo - base, tail_len = base + total - hi;
    memcpy(R.reg_edge, (const void*)base, head_len);
    memcpy(R.reg_edge + PAGE, (const void*)hi, tail_len);
    if (R.reg_dev_cap < dev_need) {
        for (int b2 = 0; b2 < 2; b2++) { if (R.reg_dev[b2]) cudaFree(R.reg_dev[b2]); R.reg_dev[b2] = nullptr; }
        R.reg_dev_cap = 0;
        for (int b2 = 0; b2 < 2; b2++)
            if (!cuda_ok(cudaMalloc(&R.reg_dev[b2], dev_need), "cudaMalloc(landing buffer)")) { ok = false; return 0; }
        R.reg_dev_cap = dev_need;
    }
    if (yw_bytes && R.reg_hy_cap < yw_bytes) {
        for (int b2 = 0; b2 < 2; b2++) { if (R.reg_hy[b2]) cudaFreeHost(R.reg_hy[b2]); R.reg_hy[b2] = nullptr; }
        R.reg_hy_cap = 0;
        for (int b2 = 0; b2 < 2; b2++)
            if (!cuda_ok(cudaHostAlloc(&R.reg_hy[b2], yw_bytes, cudaHostAllocDefault), "cudaHostAlloc(outcome staging)")) { ok = false; return 0; }
        R.reg_hy_cap = yw_bytes;
    }
    if (!R.reg_copied[0])
        for (int b2 = 0; b2 < 2; b2++) { cudaEventCreateWithFlags(&R.reg_copied[b2], cudaEventDisableTiming); cudaEventCreateWithFlags(&R.reg_packed[b2], cudaEventDisableTiming); }

    static const bool trace = getenv("APOP_B200_PIN_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    if (pinned_src) for (auto& x : G.w) { x.state = 1; x.ours = false; }   // nothing to page-lock
    std::thread worker;
    if (!pinned_src) worker = std::thread([&G] { G.run(); });
    // stops the helper; keep = the page-locked windows stay with the resident copy, else they are released now
    auto finish = [&](bool keep) {
        { std::lock_guard<std::mutex> lk(G.m); G.stop = true; G.cv.notify_all(); }
        if (worker.joinable()) worker.join();
        for (auto& x : G.w)
            if (x.state == 1 && x.ours) {
                if (keep) r->host_registered.push_back({(void*)x.lo, (size_t)(x.hi - x.lo)});
                else cudaHostUnregister((void*)x.lo);
            }
    };
    // ---- the first two windows decide (auto mode): is page-locking this memory fast enough to pay off?
    {
        std::unique_lock<std::mutex> lk(G.m);
        const size_t probe = nw > 1 ? 1 : 0;
        G.cv.wait(lk, [&] { return G.w[probe].state != 0; });
    }
    R.last_register_GBps = pinned_src ? 0.0 : G.first_rate;
    const double need_rate = 0.0;
    if (G.w[0].state != 1) {
        if (trace) fprintf(stderr, "[apop_b200 pin] registered windows not used: first window %s at %.1f GB/s (need %.0f)\n",
                           G.w[0].state == 1 ? "page-locked" : "could not be page-locked", G.first_rate, need_rate);
        finish(false);
        return 0;
    }
    size_t row_cursor = 0, carry_len = head_len, prev_rows = 0, prev_carry = 0;   // the partial first page counts as carried-over bytes
    size_t wi = 0;
    for (; ok && wi < nw; wi++) {
        {
            const auto tw = std::chrono::steady_clock::now();
            std::unique_lock<std::mutex> lk(G.m);
            G.cv.wait(lk, [&] { return G.w[wi].state != 0; });
            t_wait_reg += since(tw);
            if (G.w[wi].state != 1) break;                   // registration failed from here on: the staged path takes over
        }
        const int slot = (int)(wi & 1);
        if (wi >= 2) {
            const auto tw = std::chrono::steady_clock::now();
            ok = cuda_ok(cudaEventSynchronize(R.reg_packed[slot]), "landing buffer wait");   // window wi-2 is tiled: its buffers are free,
            t_wait_dev += since(tw);
        }
        { std::lock_guard<std::mutex> lk(G.m); G.consumed = wi; G.cv.notify_all(); }   // the helper stays `ahead` windows in front
        const auto ti = std::chrono::steady_clock::now();
        const uintptr_t hs = G.w[wi].lo, he = G.w[wi].hi;
        const bool last = (wi + 1 == nw);
        const size_t wlen = he - hs, len = wlen + (last ? tail_len : 0);   // the last window is followed by the partial last page
        char* win = R.reg_dev[slot] + carry_cap + PAGE;
        const size_t avail = carry_len + len;
        size_t rows = avail / rb;
        if (!last) rows = rows / TILE_ROWS * TILE_ROWS;
        // outcome / weights of these rows through the pinned staging slot
        double* hyv = R.reg_hy[slot];
        double* hwv = hyv ? hyv + (hy ? rows : 0) : nullptr;
        if (hy && rows) memcpy(hyv, d->vector->data + row_cursor, rows * sizeof(double));
        if (hw && rows) memcpy(hwv, d->weights->data + row_cursor, rows * sizeof(double));
        if (wi >= 2) ok = ok && cuda_ok(cudaStreamWaitEvent(R.copy_stream, R.reg_packed[slot], 0), "stream wait");
        ok = ok && cuda_ok(cudaMemcpyAsync(win, (const void*)hs, wlen, cudaMemcpyHostToDevice, R.copy_stream), "H2D window");
        if (last && tail_len)
            ok = ok && cuda_ok(cudaMemcpyAsync(win + wlen, R.reg_edge + PAGE, tail_len, cudaMemcpyHostToDevice, R.copy_stream), "H2D tail");
        if (wi == 0) {     // the partial first page, from the pinned scratch
            if (carry_len) ok = ok && cuda_ok(cudaMemcpyAsync(win - carry_len, R.reg_edge, carry_len, cudaMemcpyHostToDevice, R.copy_stream), "H2D head");
        } else if (carry_len) {   // the rows the previous window left incomplete, placed right before this window's bytes
            const char* prev_win = R.reg_dev[slot ^ 1] + carry_cap + PAGE;
            const char* left = prev_win - prev_carry + prev_rows * rb;
            ok = ok && cuda_ok(cudaMemcpyAsync(win - carry_len, left, carry_len, cudaMemcpyDeviceToDevice, R.copy_stream), "carry");
        }
        ok = ok && cuda_ok(cudaEventRecord(R.reg_copied[slot], R.copy_stream), "event");
        ok = ok && cuda_ok(cudaStreamWaitEvent(R.stream, R.reg_copied[slot], 0), "stream wait");
        double* dy = (double*)(win + ((WB + 2 * PAGE + 255) / 256 * 256));
        double* dw = dy + (hy ? rows : 0);
        if (yw_bytes && rows)
            ok = ok && cuda_ok(cudaMemcpyAsync(dy, hyv, rows * sizeof(double) * ((hy ? 1 : 0) + (hw ? 1 : 0)), cudaMemcpyHostToDevice, R.stream), "H2D outcome");
        if (rows)
            ok = ok && cuda_ok(tile_pack((const double*)(win - carry_len), hy ? dy : nullptr, hw ? dw : nullptr, rows, k, cols,
                                         r->tiles + (row_cursor / TILE_ROWS) * (size_t)cols * TILE_ROWS, R.stream, col0), "tile_pack");
        ok = ok && cuda_ok(cudaEventRecord(R.reg_packed[slot], R.stream), "event");
        prev_rows = rows; prev_carry = carry_len;
        carry_len = avail - rows * rb;
        row_cursor += rows;
        t_issue += since(ti);
    }
    ok = ok && cuda_ok(cudaStreamSynchronize(R.stream), "pin sync") && cuda_ok(cudaStreamSynchronize(R.copy_stream), "pin sync");
    finish(true);
    R.last_pin_mode = pinned_src ? 2 : 1;
    if (trace)
        fprintf(stderr, "[apop_b200 pin] zero copy (%s): %zu of %zu rows from %zu of %zu windows in %.3f s = %.1f GB/s; main thread: waiting for "
                        "page-lock %.3f s, for the device %.3f s, issuing %.3f s; helper: page-locking %.3f s (best of the first two windows %.1f GB/s)\n",
                pinned_src ? "source already page-locked" : "windows page-locked here, kept while resident", row_cursor, n, wi, nw, since(t_begin),
                row_cursor * rb * 1e-9 / since(t_begin), t_wait_reg, t_wait_dev, t_issue, G.t_register, G.first_rate);
    return row_cursor;
}

bool upload(Resident* r, const apop_data* d) {
    Runtime& R = rt();
    size_t n;
    int k;
    bool hy, hw;
    data_shape(d, n, k, hy, hw);
    // device-side ols_shuffle: no vector on the host, column 0 of the matrix is the outcome
    const bool col0 = r->y_col0 && !d->vector && d->matrix && k >= 1;
    r->y_col0 = col0;
    if (d->matrix && d->vector && d->vector->size != d->matrix->size1)
        return set_error("vector has %zu entries but the matrix has %zu rows", d->vector->size, d->matrix->size1);
    if (d->weights && d->weights->size != n) return set_error("weights have %zu entries for %zu rows", d->weights->size, n);
    if ((d->matrix && n && !d->matrix->data) || (d->vector && n && !d->vector->data))
        return set_error("data set has no host buffers (device-only synthetic data cannot be re-uploaded)");
    const int cols = k + ((hy || col0) ? 1 : 0) + (hw ? 1 : 0);          // resident columns
    const int scols = k + (hy ? 1 : 0) + (hw ? 1 : 0);                    // columns in the staging buffers
    const size_t n_tiles = (n + TILE_ROWS - 1) / TILE_ROWS;
    const size_t need = n_tiles * (size_t)cols * TILE_ROWS;
    if (r->tiles && (r->n_tiles * (size_t)r->cols != n_tiles * (size_t)cols)) {
        cudaFree(r->tiles);
        r->tiles = nullptr;
    }
    if (!r->tiles && need)
        if (!cuda_ok(cudaMalloc(&r->tiles, need * sizeof(double)), "cudaMalloc(resident tiles)")) return false;
    r->n_rows = n; r->k = k; r->has_y = hy || col0; r->has_w = hw; r->cols = cols; r->n_tiles = n_tiles;
    r->swz = false;   // tile_pack writes the plain layout
    r->memo_valid = false; r->have_lgamma_y = false; r->have_poisson = false; r->have_gamma = false; r->consts.clear(); r->dirty = false;
    r->n_global_epoch = ~0ull;
    r->sig = host_sig(d);
    // resample counts and the per-row lgamma column belong to the rows that were resident before
    if (r->counts) { cudaFree(r->counts); r->counts = nullptr; }
    r->counts_m = 0;
    r->active_replicate = -1;
    if (r->lgy) { cudaFree(r->lgy); r->lgy = nullptr; }
    if (!need) return true;

    R.last_pin_mode = 0;
    release_host_registration(r);   // a refresh means the memory behind the pointer may be other pages now
    size_t first_row = 0;
    {
        bool okr = true;
        first_row = upload_registered(r, d, col0, cols, okr);
        if (!okr) return false;
    }
    // Staged, pipelined upload.  Chunks of rows are gathered from the caller's (pageable,
    // possibly strided) buffers into one of two pinned staging buffers by a few host threads,
    // sent with one asynchronous copy each, and re-tiled on the device; the gather of chunk
    // i+1 overlaps the copy and the transpose of chunk i.
    size_t chunk = ((size_t)8 << 20) / (size_t)(scols ? scols : 1);   // ~64 MB of doubles per chunk
    chunk = chunk / TILE_ROWS * TILE_ROWS;
    if (chunk < TILE_ROWS) chunk = TILE_ROWS;
    if (chunk > n_tiles * TILE_ROWS) chunk = n_tiles * TILE_ROWS;
    const size_t chunk_doubles = chunk * (size_t)scols;
    if (R.pin_cap < chunk_doubles) {
        for (int b2 = 0; b2 < 2; b2++) {
            if (R.pin_host[b2]) cudaFreeHost(R.pin_host[b2]);
            if (R.pin_dev[b2]) cudaFree(R.pin_dev[b2]);
            R.pin_host[b2] = nullptr; R.pin_dev[b2] = nullptr;
        }
        R.pin_cap = 0;
        bool okb = true;
        for (int b2 = 0; b2 < 2 && okb; b2++)
            okb = cuda_ok(cudaHostAlloc(&R.pin_host[b2], chunk_doubles * sizeof(double), cudaHostAllocDefault), "cudaHostAlloc(staging)") &&
                  cuda_ok(cudaMalloc(&R.pin_dev[b2], chunk_doubles * sizeof(double)), "cudaMalloc(staging)");
        if (!okb) return false;
        if (!R.pin_ev[0]) {
            for (int b2 = 0; b2 < 2; b2++) { cudaEventCreateWithFlags(&R.pin_ev[b2], cudaEventDisableTiming); cudaEventCreateWithFlags(&R.pin_copied[b2], cudaEventDisableTiming); }
        }
        R.pin_cap = chunk_doubles;
    }
    const int n_threads = HostPool::get().size();
    bool ok = true;
    int ci = 0;
    static const bool trace = getenv("APOP_B200_PIN_TRACE") != nullptr;
    double t_wait = 0.0, t_gather = 0.0;
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    for (size_t r0 = first_row; ok && r0 < n; r0 += chunk, ci++) {
        const size_t rows = (n - r0 < chunk) ? n - r0 : chunk;
        const int buf = ci & 1;
        auto tw = std::chrono::steady_clock::now();
        if (ci >= 2) ok = cuda_ok(cudaEventSynchronize(R.pin_ev[buf]), "pin staging wait");   // chunk ci-2 has left this buffer
        t_wait += since(tw);
        auto tg = std::chrono::steady_clock::now();
        double* hX = R.pin_host[buf];
        double* hy_ = hX + rows * (size_t)k;
        double* hw_ = hy_ + (hy ? rows : 0);
        auto gather = [&](size_t a0, size_t a1) {   // rows [a0, a1) of this chunk
            if (k) {
                if (d->matrix->tda == (size_t)k) memcpy(hX + a0 * k, d->matrix->data + (r0 + a0) * k, (a1 - a0) * k * sizeof(double));
                else for (size_t i = a0; i < a1; i++) memcpy(hX + i * k, d->matrix->data + (r0 + i) * d->matrix->tda, k * sizeof(double));
            }
            if (hy) {
                if (d->vector->stride == 1) memcpy(hy_ + a0, d->vector->data + r0 + a0, (a1 - a0) * sizeof(double));
                else for (size_t i = a0; i < a1; i++) hy_[i] = d->vector->data[(r0 + i) * d->vector->stride];
            }
            if (hw) {
                if (d->weights->stride == 1) memcpy(hw_ + a0, d->weights->data + r0 + a0, (a1 - a0) * sizeof(double));
                else for (size_t i = a0; i < a1; i++) hw_[i] = d->weights->data[(r0 + i) * d->weights->stride];
            }
        };
        if (rows * (size_t)scols < ((size_t)1 << 17)) gather(0, rows);   // small chunks: not worth waking the pool
        else {
            const size_t pieces = (size_t)n_threads * 4;                 // finer than the thread count: evens out stragglers
            HostPool::get().run(pieces, [&](size_t p) { gather(rows * p / pieces, rows * (p + 1) / pieces); });
        }
        t_gather += since(tg);
        double* dX = R.pin_dev[buf];
        double* dy = dX + rows * (size_t)k;
        double* dw = dy + (hy ? rows : 0);
        // copies on their own stream, back to back; the tiling kernel of chunk i runs on the main stream while
        // chunk i+1 is crossing PCIe (pin_ev[buf], waited for above, also covers the device staging buffer)
        ok = ok && cuda_ok(cudaMemcpyAsync(dX, hX, rows * (size_t)scols * sizeof(double), cudaMemcpyHostToDevice, R.copy_stream), "H2D chunk");
        ok = ok && cuda_ok(cudaEventRecord(R.pin_copied[buf], R.copy_stream), "event");
        ok = ok && cuda_ok(cudaStreamWaitEvent(R.stream, R.pin_copied[buf], 0), "stream wait");
        ok = ok && cuda_ok(tile_pack(k ? dX : nullptr, hy ? dy : nullptr, hw ? dw : nullptr, rows, k, cols,
                                     r->tiles + (r0 / TILE_ROWS) * (size_t)cols * TILE_ROWS, R.stream, col0), "tile_pack");
        ok = ok && cuda_ok(cudaEventRecord(R.pin_ev[buf], R.stream), "event");
    }
    // the outcome was recoded on the device while the host copy kept its raw values (prep_outcome with
    // sync_host = 0): a refresh from that host copy applies the same category table again
    if (ok && r->recoded && r->recode_pending && r->has_y)
        ok = cuda_ok(recode_launch(r->tiles, r->n_rows, r->n_tiles, r->k, r->cols, r->cats, R.stream), "recode launch");
    ok = ok && cuda_ok(cudaStreamSynchronize(R.stream), "pin sync");
    if (trace)
        fprintf(stderr, "[apop_b200 pin] %.3f GB in %.3f s (%d chunks of %zu rows, %d threads): gather %.3f s, waiting for the copy engine %.3f s\n",
                need * 8e-9, since(t_begin), ci, chunk, n_threads, t_gather, t_wait);
    return ok;
}

extern "C" int apop_b200_pin(apop_data* d) {
    LOCK;
    if (!d) { set_error("apop_b200_pin: NULL data"); return 1; }
    if (!require_ready()) return 1;
    Runtime& R = rt();
    cudaSetDevice(R.device);
    auto it = R.reg.find(d);
    Resident* r = (it != R.reg.end()) ? it->second : new Resident();
    if (r->synthetic) return 0;
    if (it != R.reg.end() && sig_stale(r, d)) { r->dirty = true; r->recoded = false; r->y_col0 = false; }
    if (it != R.reg.end() && !r->dirty) return 0;   // already resident (e.g. by the device prep); apop_b200_invalidate forces a refresh
    if (!upload(r, d)) {
        if (it == R.reg.end()) free_resident(r);
        else { free_resident(r); R.reg.erase(it); }
        d->error = 'a';
        return 1;
    }
    R.reg[d] = r;
    return 0;
}
extern "C" int apop_b200_unpin(apop_data* d) {
    LOCK;
    Runtime& R = rt();
    auto it = R.reg.find(d);
    if (it == R.reg.end()) return 1;
    cudaSetDevice(R.device);
    free_resident(it->second);
    R.reg.erase(it);
    return 0;
}
extern "C" int apop_b200_invalidate(apop_data* d) {
    LOCK;
    auto it = rt().reg.find(d);
    if (it == rt().reg.end()) return 1;
    if (it->second->synthetic) return 1;
    it->second->dirty = true;
    it->second->memo_valid = false;
    cudaSetDevice(rt().device);
    release_host_registration(it->second);   // the caller is about to change (or has changed) that memory
    return 0;
}
extern "C" int apop_b200_is_pinned(const apop_data* d) {
    LOCK;
    return rt().reg.count(d) ? 1 : 0;
}

// find the resident copy of d; upload it for this call if it is not pinned
// the resident copy is kept in one of two layouts; a conversion is one in-place pass over half the columns
// (2.6 GB read + written for 20M x 33: ~1 ms), paid only when a data set changes hands between the multinomial
// logit kernels and any other kernel
bool ensure_layout(Resident* r, bool swz) {
    if (!r || r->swz == swz || !r->tiles) return true;
    if (!cuda_ok(tile_swizzle(r->tiles, r->n_tiles, r->cols, rt().stream), "tile_swizzle")) return false;
    rt().launches++;
    r->swz = swz;
    return true;
}
Resident* acquire(apop_data* d, int layout) {
    if (!require_ready()) return nullptr;
    Runtime& R = rt();
    cudaSetDevice(R.device);
    auto it = R.reg.find(d);
    if (it != R.reg.end()) {
        Resident* r = it->second;
        if (sig_stale(r, d)) { r->dirty = true; r->recoded = false; r->y_col0 = false; }
        if (r->dirty && !upload(r, d)) return nullptr;
        if (r->consts_epoch != R.comm_epoch) {   // cached sums are sums over the ranks of one communicator
            r->memo_valid = false; r->have_lgamma_y = false; r->have_poisson = false; r->have_gamma = false; r->consts.clear();
            r->consts_epoch = R.comm_epoch;
        }
        if (layout != LAYOUT_ANY && !ensure_layout(r, layout == LAYOUT_SWZ)) return nullptr;
        return r;
    }
    Resident* r = new Resident();
    if (!upload(r, d)) { free_resident(r); return nullptr; }
    if (R.auto_pin) R.reg[d] = r;
    else r->transient = true;
    if (layout == LAYOUT_SWZ && !ensure_layout(r, true)) { if (r->transient) free_resident(r); return nullptr; }
    return r;
}
void release(Resident* r) {
    if (r && r->transient) free_resident(r);
}

extern "C" apop_data* apop_b200_synth(apop_b200_family family, size_t n_rows, size_t k, int ncat,
                                      const double* beta_true, uint64_t seed, uint64_t row_offset) {
    LOCK;
    if (!require_ready()) return nullptr;
    Runtime& R = rt();
    cudaSetDevice(R.device);
    const bool iid = (family == APOP_B200_POISSON || family == APOP_B200_NORMAL);
    if (family == APOP_B200_LOGIT && ncat < 2) ncat = 2;
    const int C1 = (family == APOP_B200_LOGIT) ? ncat - 1 : 1;
    const size_t P = iid ? (family == APOP_B200_NORMAL ? 2 : 1) : k * C1;
    if (P > SYNTH_MAX_P || C1 > SYNTH_MAX_C1 || k == 0) { set_error("apop_b200_synth: shape out of range"); return nullptr; }
    SynthParams prm;
    memset(&prm, 0, sizeof prm);
    memcpy(prm.beta, beta_true, P * sizeof(double));
    Resident* r = new Resident();
    r->n_rows = n_rows; r->k = (int)k; r->has_y = !iid; r->has_w = false;
    r->cols = (int)k + (iid ? 0 : 1);
    r->n_tiles = (n_rows + TILE_ROWS - 1) / TILE_ROWS;
    r->synthetic = true;
    const size_t need = r->n_tiles * (size_t)r->cols * TILE_ROWS;
    if (need && !cuda_ok(cudaMalloc(&r->tiles, need * sizeof(double)), "cudaMalloc(synthetic tiles)")) { delete r; return nullptr; }
    if (!cuda_ok(synth_fill((int)family, n_rows, (int)k, r->cols, ncat, prm, seed, row_offset, r->tiles, R.stream), "synth") ||
        !cuda_ok(cudaStreamSynchronize(R.stream), "synth sync")) { free_resident(r); return nullptr; }
    apop_data* d = (apop_data*)calloc(1, sizeof(apop_data));
    d->matrix = (gsl_matrix*)calloc(1, sizeof(gsl_matrix));
    d->matrix->size1 = n_rows; d->matrix->size2 = k; d->matrix->tda = k;
    if (!iid) {
        d->vector = (gsl_vector*)calloc(1, sizeof(gsl_vector));
        d->vector->size = n_rows; d->vector->stride = 1;
    }
    R.reg[d] = r;
    return d;
}
extern "C" void apop_b200_synth_free(apop_data* d) {
    if (!d) return;
    apop_b200_unpin(d);
    free(d->matrix);
    free(d->vector);
    free(d);
}
extern "C" int apop_b200_download(const apop_data* d, double* X, double* y, double* w) {
    LOCK;
    Runtime& R = rt();
    auto it = R.reg.find(d);
    if (it == R.reg.end()) { set_error("apop_b200_download: data set is not resident"); return 1; }
    Resident* r = it->second;
    cudaSetDevice(R.device);
    if (!r->n_rows) return 0;
    if (!ensure_layout(r, false)) return 1;
    size_t chunk = ((size_t)32 << 20) / (size_t)(r->cols ? r->cols : 1) / TILE_ROWS * TILE_ROWS;
    if (chunk < TILE_ROWS) chunk = TILE_ROWS;
    double *Xs = nullptr, *ys = nullptr, *ws = nullptr;
    bool ok = true;
    if (r->k && X) ok = ok && cuda_ok(cudaMalloc(&Xs, chunk * r->k * sizeof(double)), "cudaMalloc(staging)");
    if (r->has_y && y) ok = ok && cuda_ok(cudaMalloc(&ys, chunk * sizeof(double)), "cudaMalloc(staging)");
    if (r->has_w && w) ok = ok && cuda_ok(cudaMalloc(&ws, chunk * sizeof(double)), "cudaMalloc(staging)");
    for (size_t r0 = 0; ok && r0 < r->n_rows; r0 += chunk) {
        const size_t rows = (r->n_rows - r0 < chunk) ? r->n_rows - r0 : chunk;
        ok = ok && cuda_ok(tile_unpack(r->tiles + (r0 / TILE_ROWS) * (size_t)r->cols * TILE_ROWS, rows, r->k, r->cols,
                                       r->has_y, r->has_w, Xs, ys, ws, R.stream), "tile_unpack");
        if (Xs) ok = ok && cuda_ok(cudaMemcpyAsync(X + r0 * r->k, Xs, rows * r->k * sizeof(double), cudaMemcpyDeviceToHost, R.stream), "D2H X");
        if (ys) ok = ok && cuda_ok(cudaMemcpyAsync(y + r0, ys, rows * sizeof(double), cudaMemcpyDeviceToHost, R.stream), "D2H y");
        if (ws) ok = ok && cuda_ok(cudaMemcpyAsync(w + r0, ws, rows * sizeof(double), cudaMemcpyDeviceToHost, R.stream), "D2H w");
        ok = ok && cuda_ok(cudaStreamSynchronize(R.stream), "download sync");
    }
    if (Xs) cudaFree(Xs);
    if (ys) cudaFree(ys);
    if (ws) cudaFree(ws);
    return ok ? 0 : 1;
}

