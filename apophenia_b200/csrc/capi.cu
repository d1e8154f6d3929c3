// capi.cu -- the C-ABI of libapop_b200 (include/apop_b200.h): runtime, residency,
// and the reference-signature log-likelihood / score entry points.
//
// There is no CPU fallback anywhere in this file: without a B200 every entry
// point fails loudly (NaN + apop_b200_last_error).
#include "runtime.cuh"
#include "logit_kernel.cuh"
#include "xtwx_kernel.cuh"
#include "batched_kernel.cuh"
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <thread>

using namespace apb;

// ---- struct layouts are part of the ABI: pin them ------------------------
static_assert(sizeof(gsl_vector) == 40 && offsetof(gsl_vector, data) == 16, "gsl_vector layout");
static_assert(sizeof(gsl_matrix) == 48 && offsetof(gsl_matrix, data) == 24, "gsl_matrix layout");
static_assert(offsetof(apop_data, weights) == 48 && offsetof(apop_data, more) == 56 &&
                  offsetof(apop_data, error) == 64, "apop_data layout (apop.m4.h:72-81)");
static_assert(offsetof(apop_model, data) == 120 && offsetof(apop_model, parameters) == 128 &&
                  offsetof(apop_model, log_likelihood) == 160 && offsetof(apop_model, prep) == 192 &&
                  offsetof(apop_model, error) == 224, "apop_model layout (apop.m4.h:98-115)");

namespace apb {
Runtime& rt() {
    static Runtime r;
    return r;
}
std::recursive_mutex& rt_mutex() {
    static std::recursive_mutex m;
    return m;
}
bool set_error(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    rt().err = buf;
    if (getenv("APOP_B200_VERBOSE")) fprintf(stderr, "apop_b200: %s\n", buf);
    return false;
}
bool cuda_ok(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    return set_error("%s: %s", what, cudaGetErrorString(e));
}
bool ensure_partials(size_t doubles) {
    Runtime& R = rt();
    if (doubles <= R.partials_cap) return true;
    if (R.d_partials) cudaFree(R.d_partials);
    R.partials_cap = 0;
    if (!cuda_ok(cudaMalloc(&R.d_partials, doubles * sizeof(double)), "cudaMalloc(partials)")) return false;
    R.partials_cap = doubles;
    return true;
}
}  // namespace apb

#define LOCK std::lock_guard<std::recursive_mutex> lock_(rt_mutex())

static bool require_ready() {
    if (rt().ready) return true;
    // lazy init on device 0 so that a bare registered model works out of the box
    return apop_b200_init(-1) == 0;
}

// --------------------------------------------------------------------------
//  runtime
// --------------------------------------------------------------------------
extern "C" int apop_b200_init(int device) {
    LOCK;
    Runtime& R = rt();
    if (R.ready && (device < 0 || device == R.device)) return 0;
    if (R.ready) apop_b200_finalize();
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        set_error("no CUDA device: libapop_b200 has no CPU fallback");
        return 1;
    }
    if (device < 0) {
        const char* lr = getenv("LOCAL_RANK");
        device = lr ? atoi(lr) % n : 0;
    }
    if (!cuda_ok(cudaSetDevice(device), "cudaSetDevice")) return 1;
    cudaDeviceProp prop;
    if (!cuda_ok(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) return 1;
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return 1;
    }
    R.device = device;
    R.num_sms = prop.multiProcessorCount;
    if (!cuda_ok(cudaStreamCreateWithFlags(&R.stream, cudaStreamNonBlocking), "cudaStreamCreate")) return 1;
    cudaEventCreate(&R.ev0);
    cudaEventCreate(&R.ev1);
    if (!cuda_ok(cudaMalloc(&R.d_ticket, 64), "cudaMalloc")) return 1;
    cudaMemset(R.d_ticket, 0, 64);
    if (!cuda_ok(cudaMalloc(&R.d_out, OUT_CAP * sizeof(double)), "cudaMalloc")) return 1;
    if (!cuda_ok(cudaHostAlloc(&R.h_out, OUT_CAP * sizeof(double), cudaHostAllocMapped), "cudaHostAlloc")) return 1;
    if (!cuda_ok(cudaHostGetDevicePointer(&R.h_out_dev, R.h_out, 0), "cudaHostGetDevicePointer")) return 1;
    if (!cuda_ok(cudaHostAlloc(&R.h_err, 64, cudaHostAllocMapped), "cudaHostAlloc")) return 1;
    *R.h_err = 0;
    if (!cuda_ok(cudaHostGetDevicePointer(&R.h_err_dev, R.h_err, 0), "cudaHostGetDevicePointer")) return 1;
    if (!cuda_ok(cudaMalloc(&R.d_vals, 64 * sizeof(double)), "cudaMalloc")) return 1;
    R.ready = true;
    R.err.clear();
    return 0;
}

static void free_resident(Resident* r) {
    if (!r) return;
    if (r->tiles) cudaFree(r->tiles);
    if (r->counts) cudaFree(r->counts);
    delete r;
}

extern "C" void apop_b200_finalize(void) {
    LOCK;
    Runtime& R = rt();
    if (!R.ready) return;
    cudaSetDevice(R.device);
    apop_b200_comm_finalize();
    for (auto& kv : R.reg) free_resident(kv.second);
    R.reg.clear();
    if (R.d_partials) cudaFree(R.d_partials);
    R.d_partials = nullptr;
    R.partials_cap = 0;
    cudaFree(R.d_ticket);
    cudaFree(R.d_out);
    cudaFreeHost(R.h_out);
    cudaFreeHost(R.h_err);
    for (int b2 = 0; b2 < 2; b2++) { if (R.pin_host[b2]) cudaFreeHost(R.pin_host[b2]); if (R.pin_dev[b2]) cudaFree(R.pin_dev[b2]); R.pin_host[b2] = nullptr; R.pin_dev[b2] = nullptr; }
    R.pin_cap = 0;
    cudaFree(R.d_vals);
    if (R.d_beta) { cudaFree(R.d_beta); R.d_beta = nullptr; }
    cudaEventDestroy(R.ev0);
    cudaEventDestroy(R.ev1);
    cudaStreamDestroy(R.stream);
    R.ready = false;
}

extern "C" const char* apop_b200_last_error(void) { return rt().err.c_str(); }
extern "C" unsigned long long apop_b200_launch_count(void) { return rt().launches; }
extern "C" double apop_b200_last_pass_ms(void) { return rt().last_ms; }
extern "C" void apop_b200_pass_timer_reset(void) {
    rt().total_ms = 0;
    rt().n_passes = 0;
}
extern "C" double apop_b200_pass_timer_total_ms(unsigned long long* n) {
    if (n) *n = rt().n_passes;
    return rt().total_ms;
}
extern "C" void apop_b200_set_auto_pin(int on) { rt().auto_pin = on != 0; }

// --------------------------------------------------------------------------
//  NCCL, bound at run time (torch's bundled libnccl if it is already loaded in
//  the process, else the system one)
// --------------------------------------------------------------------------
static bool nccl_load() {
    NcclApi& N = rt().nccl;
    if (N.handle) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) return set_error("cannot load libnccl.so.2: %s", dlerror());
    N.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
    N.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(h, "ncclCommInitRank");
    N.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
    N.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
    N.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    if (!N.GetUniqueId || !N.CommInitRank || !N.AllReduce || !N.CommDestroy)
        return set_error("libnccl.so.2 lacks the expected symbols");
    N.handle = h;
    return true;
}
static bool nccl_ok(int rc, const char* what) {
    if (rc == 0) return true;
    NcclApi& N = rt().nccl;
    return set_error("%s: %s", what, N.GetErrorString ? N.GetErrorString(rc) : "NCCL error");
}
extern "C" int apop_b200_comm_unique_id(void* id128) {
    LOCK;
    if (!require_ready() || !nccl_load()) return 1;
    return nccl_ok(rt().nccl.GetUniqueId(id128), "ncclGetUniqueId") ? 0 : 1;
}
extern "C" int apop_b200_comm_init(int nranks, int rank, const void* id128) {
    LOCK;
    if (!require_ready() || !nccl_load()) return 1;
    Runtime& R = rt();
    cudaSetDevice(R.device);
    NcclId id;
    memcpy(&id, id128, sizeof id);
    if (!nccl_ok(R.nccl.CommInitRank(&R.comm, nranks, id, rank), "ncclCommInitRank")) return 1;
    R.nranks = nranks;
    R.rank = rank;
    R.comm_epoch++;
    return 0;
}
extern "C" int apop_b200_comm_finalize(void) {
    LOCK;
    Runtime& R = rt();
    if (R.comm) {
        cudaStreamSynchronize(R.stream);
        R.nccl.CommDestroy(R.comm);
        R.comm = nullptr;
    }
    if (R.p2p_active) {
        for (int p = 0; p < R.nranks; p++)
            if (p != R.rank && R.p2p_inbox[p]) cudaIpcCloseMemHandle(R.p2p_inbox[p]);
        R.p2p_active = false;
    }
    if (R.p2p_local) { cudaFree(R.p2p_local); R.p2p_local = nullptr; }
    R.nranks = 1;
    R.rank = 0;
    R.comm_epoch++;
    return 0;
}
extern "C" int apop_b200_comm_size(void) { return rt().nranks; }

// ---- fused allreduce over NVLink peer memory ------------------------------------------------
// Each rank allocates one buffer [2][nranks][slot] doubles + [2][nranks] flags, exports it with
// cudaIpcGetMemHandle (64 bytes, carried to the peers by the host program like the NCCL id), and
// maps every peer's buffer.  From then on the reduction kernels exchange their result vectors
// themselves (common.cuh: grid_finalize); NCCL stays for the large batched outputs.
static const int P2P_SLOT = 1088;   // >= the widest fused result (32 x 33 information matrix)
static size_t p2p_bytes(int nranks) { return (size_t)2 * nranks * P2P_SLOT * sizeof(double) + (size_t)2 * nranks * sizeof(unsigned long long); }
extern "C" int apop_b200_comm_p2p_handle(int nranks, void* handle64) {
    LOCK;
    if (!require_ready()) return 1;
    Runtime& R = rt();
    if (nranks < 2 || nranks > P2P_MAX_RANKS) { set_error("fused allreduce: 2..%d ranks", P2P_MAX_RANKS); return 1; }
    cudaSetDevice(R.device);
    if (R.p2p_local) { cudaFree(R.p2p_local); R.p2p_local = nullptr; }
    if (!cuda_ok(cudaMalloc(&R.p2p_local, p2p_bytes(nranks)), "cudaMalloc(p2p inbox)")) return 1;
    cudaMemset(R.p2p_local, 0, p2p_bytes(nranks));
    cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (!cuda_ok(cudaIpcGetMemHandle(&h, R.p2p_local), "cudaIpcGetMemHandle")) return 1;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle64, &h, 64);
    return 0;
}
extern "C" int apop_b200_comm_p2p_init(int nranks, int rank, const void* handles) {
    LOCK;
    Runtime& R = rt();
    if (!R.p2p_local || nranks != R.nranks || rank != R.rank) { set_error("fused allreduce: call apop_b200_comm_init and apop_b200_comm_p2p_handle first"); return 1; }
    cudaSetDevice(R.device);
    for (int p = 0; p < nranks; p++) {
        void* base = R.p2p_local;
        if (p != rank) {
            cudaIpcMemHandle_t h;
            memcpy(&h, (const char*)handles + 64 * p, 64);
            if (!cuda_ok(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle")) return 1;
        }
        R.p2p_inbox[p] = (double*)base;
        R.p2p_flags[p] = (unsigned long long*)((char*)base + (size_t)2 * nranks * P2P_SLOT * sizeof(double));
    }
    R.p2p_slot = P2P_SLOT;
    R.p2p_seq = 0;
    R.p2p_active = true;
    return 0;
}
extern "C" int apop_b200_comm_p2p_active(void) { return rt().p2p_active ? 1 : 0;}

// The context the last CTA of a reduction kernel gets: where to publish, and -- with the fused
// allreduce active -- the peers' inboxes.  Without it on several ranks the kernel only writes
// d_out and NCCL sums afterwards.
static FinalizeCtx make_fin(int nout) {
    Runtime& R = rt();
    FinalizeCtx f;
    memset(&f, 0, sizeof f);
    f.nranks = 1;
    f.spin_limit = 40000000ull;   // ~10 s of 64 ns sleeps + loads, then the error flag
    f.err = R.h_err_dev;
    if (R.nranks > 1 && R.p2p_active && nout <= R.p2p_slot) {
        f.nranks = R.nranks; f.rank = R.rank; f.seq = ++R.p2p_seq; f.slot = R.p2p_slot;
        for (int p = 0; p < R.nranks; p++) { f.inbox[p] = R.p2p_inbox[p]; f.flags[p] = R.p2p_flags[p]; }
        f.host_out = R.h_out_dev;
    } else if (R.nranks <= 1)
        f.host_out = R.h_out_dev;
    return f;
}
// complete a reduction launched with `fin`: `n` doubles, summed over all ranks, end up in h_out
static bool reduce_and_fetch(size_t n, const FinalizeCtx& fin) {
    Runtime& R = rt();
    if (fin.host_out) {   // the kernel published (and, with peers, exchanged) the result itself
        if (!cuda_ok(cudaStreamSynchronize(R.stream), "stream sync")) return false;
        if (*R.h_err) { *R.h_err = 0; return set_error("fused allreduce: a peer rank never arrived (sequence %llu)", fin.seq); }
        return true;
    }
    if (R.comm && R.nranks > 1)
        if (!nccl_ok(R.nccl.AllReduce(R.d_out, R.d_out, n, /*ncclDouble*/ 8, /*ncclSum*/ 0, R.comm, R.stream),
                     "ncclAllReduce"))
            return false;
    if (!cuda_ok(cudaMemcpyAsync(R.h_out, R.d_out, n * sizeof(double), cudaMemcpyDeviceToHost, R.stream), "D2H")) return false;
    return cuda_ok(cudaStreamSynchronize(R.stream), "stream sync");
}
// sum host doubles across ranks (n <= 64); result back in v
static bool allreduce_host(double* v, int n) {
    Runtime& R = rt();
    if (R.nranks <= 1) return true;
    if (!cuda_ok(cudaMemcpyAsync(R.d_vals, v, n * sizeof(double), cudaMemcpyHostToDevice, R.stream), "H2D")) return false;
    const FinalizeCtx fin = make_fin(n);
    if (!cuda_ok(exchange_launch(R.d_vals, n, R.d_ticket, R.d_out, R.stream, fin), "exchange launch")) return false;
    if (!reduce_and_fetch(n, fin)) return false;
    memcpy(v, R.h_out, n * sizeof(double));
    return true;
}
static bool allreduce_scalar(double* v) { return allreduce_host(v, 1); }

// --------------------------------------------------------------------------
//  residency
// --------------------------------------------------------------------------
static void data_shape(const apop_data* d, size_t& n, int& k, bool& has_y, bool& has_w) {
    n = d->matrix ? d->matrix->size1 : (d->vector ? d->vector->size : 0);
    k = d->matrix ? (int)d->matrix->size2 : 0;
    has_y = d->vector != nullptr;
    has_w = d->weights != nullptr;
}

static bool upload(Resident* r, const apop_data* d) {
    Runtime& R = rt();
    size_t n;
    int k;
    bool hy, hw;
    data_shape(d, n, k, hy, hw);
    if (d->matrix && d->vector && d->vector->size != d->matrix->size1)
        return set_error("vector has %zu entries but the matrix has %zu rows", d->vector->size, d->matrix->size1);
    if (d->weights && d->weights->size != n) return set_error("weights have %zu entries for %zu rows", d->weights->size, n);
    if ((d->matrix && n && !d->matrix->data) || (d->vector && n && !d->vector->data))
        return set_error("data set has no host buffers (device-only synthetic data cannot be re-uploaded)");
    const int cols = k + (hy ? 1 : 0) + (hw ? 1 : 0);
    const size_t n_tiles = (n + TILE_ROWS - 1) / TILE_ROWS;
    const size_t need = n_tiles * (size_t)cols * TILE_ROWS;
    if (r->tiles && (r->n_tiles * (size_t)r->cols != n_tiles * (size_t)cols)) {
        cudaFree(r->tiles);
        r->tiles = nullptr;
    }
    if (!r->tiles && need)
        if (!cuda_ok(cudaMalloc(&r->tiles, need * sizeof(double)), "cudaMalloc(resident tiles)")) return false;
    r->n_rows = n; r->k = k; r->has_y = hy; r->has_w = hw; r->cols = cols; r->n_tiles = n_tiles;
    r->memo_valid = false; r->have_lgamma_y = false; r->have_poisson = false; r->have_gamma = false; r->dirty = false;
    r->n_global_epoch = ~0ull;
    if (!need) return true;

    // Staged, pipelined upload.  Chunks of rows are gathered from the caller's (pageable,
    // possibly strided) buffers into one of two pinned staging buffers by a few host threads,
    // sent with one asynchronous copy each, and re-tiled on the device; the gather of chunk
    // i+1 overlaps the copy and the transpose of chunk i.
    size_t chunk = ((size_t)8 << 20) / (size_t)(cols ? cols : 1);   // ~64 MB of doubles per chunk
    chunk = chunk / TILE_ROWS * TILE_ROWS;
    if (chunk < TILE_ROWS) chunk = TILE_ROWS;
    if (chunk > n_tiles * TILE_ROWS) chunk = n_tiles * TILE_ROWS;
    const size_t chunk_doubles = chunk * (size_t)cols;
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
        if (!R.pin_ev[0]) { cudaEventCreateWithFlags(&R.pin_ev[0], cudaEventDisableTiming); cudaEventCreateWithFlags(&R.pin_ev[1], cudaEventDisableTiming); }
        R.pin_cap = chunk_doubles;
    }
    static int n_threads = 0;
    if (!n_threads) {
        const char* e = getenv("APOP_B200_PIN_THREADS");
        n_threads = e ? atoi(e) : 16;
        const int hw = (int)std::thread::hardware_concurrency();
        if (hw > 0 && n_threads > hw) n_threads = hw;
        if (n_threads < 1) n_threads = 1;
    }
    bool ok = true;
    int ci = 0;
    for (size_t r0 = 0; ok && r0 < n; r0 += chunk, ci++) {
        const size_t rows = (n - r0 < chunk) ? n - r0 : chunk;
        const int buf = ci & 1;
        if (ci >= 2) ok = cuda_ok(cudaEventSynchronize(R.pin_ev[buf]), "pin staging wait");   // chunk ci-2 has left this buffer
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
        const int nt = (rows * (size_t)cols < ((size_t)1 << 17)) ? 1 : n_threads;   // small chunks: not worth a thread
        if (nt == 1) gather(0, rows);
        else {
            std::vector<std::thread> pool;
            for (int t = 0; t < nt; t++) pool.emplace_back(gather, rows * t / nt, rows * (t + 1) / nt);
            for (auto& th : pool) th.join();
        }
        double* dX = R.pin_dev[buf];
        double* dy = dX + rows * (size_t)k;
        double* dw = dy + (hy ? rows : 0);
        ok = ok && cuda_ok(cudaMemcpyAsync(dX, hX, rows * (size_t)cols * sizeof(double), cudaMemcpyHostToDevice, R.stream), "H2D chunk");
        ok = ok && cuda_ok(tile_pack(k ? dX : nullptr, hy ? dy : nullptr, hw ? dw : nullptr, rows, k, cols,
                                     r->tiles + (r0 / TILE_ROWS) * (size_t)cols * TILE_ROWS, R.stream), "tile_pack");
        ok = ok && cuda_ok(cudaEventRecord(R.pin_ev[buf], R.stream), "event");
    }
    ok = ok && cuda_ok(cudaStreamSynchronize(R.stream), "pin sync");
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
    return 0;
}
extern "C" int apop_b200_is_pinned(const apop_data* d) {
    LOCK;
    return rt().reg.count(d) ? 1 : 0;
}

// find the resident copy of d; upload it for this call if it is not pinned
static Resident* acquire(apop_data* d) {
    if (!require_ready()) return nullptr;
    Runtime& R = rt();
    cudaSetDevice(R.device);
    auto it = R.reg.find(d);
    if (it != R.reg.end()) {
        Resident* r = it->second;
        if (r->dirty && !upload(r, d)) return nullptr;
        if (r->consts_epoch != R.comm_epoch) {   // cached sums are sums over the ranks of one communicator
            r->memo_valid = false; r->have_lgamma_y = false; r->have_poisson = false; r->have_gamma = false;
            r->consts_epoch = R.comm_epoch;
        }
        return r;
    }
    Resident* r = new Resident();
    if (!upload(r, d)) { free_resident(r); return nullptr; }
    if (R.auto_pin) R.reg[d] = r;
    else r->transient = true;
    return r;
}
static void release(Resident* r) {
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

// --------------------------------------------------------------------------
//  evaluation
// --------------------------------------------------------------------------
static int grid_for(int blocks_per_sm, size_t n_tiles, int warps_per_block) {
    Runtime& R = rt();
    long long g = (long long)R.num_sms * (blocks_per_sm > 0 ? blocks_per_sm : 1);
    const long long need = (long long)((n_tiles + warps_per_block - 1) / warps_per_block);
    if (g > need) g = need;
    if (g < 1) g = 1;
    return (int)g;
}

static void time_begin() { cudaEventRecord(rt().ev0, rt().stream); }
static void time_end_record() { cudaEventRecord(rt().ev1, rt().stream); }
static void time_collect() {
    Runtime& R = rt();
    float ms = 0;
    if (cudaEventElapsedTime(&ms, R.ev0, R.ev1) == cudaSuccess) {
        R.last_ms = ms;
        R.total_ms += ms;
        R.n_passes++;
    }
}

// One fused pass of a GLM family; out receives MAX_ACC + k doubles, summed over
// all ranks.  Served from the memo when (family, beta, aux) repeat.
static bool eval_glm(Resident* r, GlmFam fam, const double* beta, size_t P, const double* aux, std::vector<double>& out) {
    Runtime& R = rt();
    const int K = r->k;
    if ((size_t)K != P) return set_error("parameter count %zu does not match the %d data columns", P, K);
    if (K < 1 || K > GLM_WIDE_MAX_K) return set_error("k = %d columns: the fused kernels cover 1..%d", K, GLM_WIDE_MAX_K);
    if (!r->has_y) return set_error("the data set has no outcome vector (run the model's prep first)");
    std::vector<double> key(beta, beta + P);
    for (int a = 0; a < 4; a++) key.push_back(aux ? aux[a] : 0.0);
    if (r->memo_valid && r->memo_fam == (int)fam && r->memo_beta.size() == key.size() &&
        memcmp(r->memo_beta.data(), key.data(), key.size() * sizeof(double)) == 0) {
        out = r->memo_out;
        return true;
    }
    const bool has_w = r->has_w && fam == GLM_OLS;
    // a weights column the family ignores still sits in the tile: the kernels
    // index columns by COLS, so only OLS may carry weights
    if (r->has_w && fam != GLM_OLS) return set_error("weights are only used by the OLS family; drop them before pinning");
    const bool wide = K > MAX_K_REG;
    static int bps_cache[GLM_FAM_COUNT][MAX_K_REG + 1][2];
    int bps_wide = 0;
    int& bps = wide ? bps_wide : bps_cache[fam][K][has_w ? 1 : 0];
    if (!bps) bps = wide ? glm_wide_blocks_per_sm(fam, K) : glm_blocks_per_sm(fam, K, has_w);
    if (bps <= 0) return set_error("kernel for family %d, k=%d is not available on this device", (int)fam, K);
    const int NOUT = MAX_ACC + K;
    GlmLaunch L;
    L.tiles = r->tiles; L.n_rows = r->n_rows; L.n_tiles = r->n_tiles;
    L.grid = grid_for(bps, r->n_tiles, GLM_THREADS / 32);
    if (!ensure_partials((size_t)L.grid * NOUT)) return false;
    L.partials = R.d_partials; L.ticket = R.d_ticket; L.out = R.d_out; L.stream = R.stream;
    L.fin = make_fin(NOUT);
    GlmParams prm;
    memset(&prm, 0, sizeof prm);
    if (!wide) memcpy(prm.beta, beta, P * sizeof(double));
    if (aux) memcpy(prm.aux, aux, 4 * sizeof(double));
    if (wide) {   // beta does not fit the kernel parameter: stage it in device memory
        if (!R.d_beta && !cuda_ok(cudaMalloc(&R.d_beta, GLM_WIDE_MAX_K * sizeof(double)), "cudaMalloc(beta)")) return false;
        if (!cuda_ok(cudaMemcpyAsync(R.d_beta, beta, P * sizeof(double), cudaMemcpyHostToDevice, R.stream), "H2D beta")) return false;
    }
    time_begin();
    if (!cuda_ok(wide ? glm_wide_launch(fam, K, r->cols, has_w, L, R.d_beta, prm.aux) : glm_launch(fam, K, has_w, L, prm),
                 "glm kernel launch")) return false;
    time_end_record();
    R.launches++;
    if (!reduce_and_fetch(NOUT, L.fin)) return false;
    time_collect();
    out.assign(R.h_out, R.h_out + NOUT);
    r->memo_valid = true; r->memo_fam = (int)fam; r->memo_beta = key; r->memo_out = out;
    return true;
}

static bool eval_cells(Resident* r, int mode, double p0, double* out5, bool matrix = true) {
    Runtime& R = rt();
    const int grid = grid_for(8, r->n_tiles, CELLS_THREADS / 32);
    if (!ensure_partials((size_t)grid * CELLS_NOUT)) return false;
    const FinalizeCtx fin = make_fin(CELLS_NOUT);
    time_begin();
    if (!cuda_ok(cells_launch(mode, r->tiles, r->n_rows, r->n_tiles, matrix ? r->k : 0, r->has_y ? r->k : -1, r->cols, p0, R.d_partials,
                              R.d_ticket, R.d_out, grid, R.stream, fin), "cells kernel launch")) return false;
    time_end_record();
    R.launches++;
    if (!reduce_and_fetch(CELLS_NOUT, fin)) return false;
    time_collect();
    memcpy(out5, R.h_out, CELLS_NOUT * sizeof(double));
    return true;
}

// global (all ranks) cell counts of a resident data set
static bool global_counts(Resident* r, double& n_rows, double& m_cells, double& v_cells) {
    Runtime& R = rt();
    if (r->n_global_epoch != R.comm_epoch) {   // one exchange per (data set, communicator), then cached
        double n = (double)r->n_rows;
        if (R.nranks > 1 && !allreduce_scalar(&n)) return false;
        r->n_global = n;
        r->n_global_epoch = R.comm_epoch;
    }
    n_rows = r->n_global;
    m_cells = n_rows * r->k;
    v_cells = r->has_y ? n_rows : 0.0;
    return true;
}

// parameters in apop_data_pack order: vector, matrix rows, weights
// (apop_conversions.m4.c:762-806)
static bool pack_parameters(const apop_data* p, std::vector<double>& out) {
    out.clear();
    if (!p) return false;
    if (p->vector)
        for (size_t i = 0; i < p->vector->size; i++) out.push_back(p->vector->data[i * p->vector->stride]);
    if (p->matrix)
        for (size_t i = 0; i < p->matrix->size1; i++)
            for (size_t j = 0; j < p->matrix->size2; j++) out.push_back(p->matrix->data[i * p->matrix->tda + j]);
    if (p->weights)
        for (size_t i = 0; i < p->weights->size; i++) out.push_back(p->weights->data[i * p->weights->stride]);
    return !out.empty();
}

// the factor page apop_data_to_factors leaves behind ("<categories for ...>",
// apop_regression.m4.c:135-163); NULL if the data carries none
static const apop_data* category_page(const apop_data* d) {
    for (const apop_data* pg = d ? d->more : nullptr; pg; pg = pg->more)
        if (pg->names && pg->names->title && !strncmp(pg->names->title, "<categories for", 15)) return pg;
    return nullptr;
}
static const apop_data* named_page(const apop_data* d, const char* title) {
    for (const apop_data* pg = d ? d->more : nullptr; pg; pg = pg->more)
        if (pg->names && pg->names->title && !strcasecmp(pg->names->title, title)) return pg;
    return nullptr;
}

static void fill_nan(gsl_vector* g) {
    if (g && g->data)
        for (size_t i = 0; i < g->size; i++) g->data[i * g->stride] = NAN;
}
static void store_score(gsl_vector* g, const double* src, size_t n, double scale) {
    for (size_t i = 0; i < n && i < g->size; i++) g->data[i * g->stride] = src[i] * scale;
}

// ---- GLM-shaped families: probit, binary logit, poisson regression ---------
struct GlmEval {
    bool ok = false;
    std::vector<double> out;  // MAX_ACC + k
    std::vector<double> beta;
    double n_rows = 0;        // global
    Resident* r = nullptr;
};

static GlmEval glm_entry(GlmFam fam, apop_data* d, apop_model* m, const double* aux) {
    GlmEval e;
    if (!m || !m->parameters || !d) { set_error("NULL model, parameters or data"); return e; }
    if (!pack_parameters(m->parameters, e.beta)) { set_error("empty parameters"); return e; }
    Resident* r = acquire(d);
    if (!r) { d->error = 'a'; return e; }
    e.ok = eval_glm(r, fam, e.beta.data(), e.beta.size(), aux, e.out);
    if (!e.ok && m) m->error = 'd';
    e.r = r;
    return e;
}

extern "C" long double apop_b200_probit_ll(apop_data* d, apop_model* m) {
    LOCK;
    const apop_data* cats = category_page(d);
    if (cats && cats->vector && cats->vector->size != 2) {
        set_error("multinomial probit (%zu categories) is outside this library's path (SURVEY F4)", cats->vector->size);
        if (m) m->error = 'd';
        return NAN;
    }
    GlmEval e = glm_entry(GLM_PROBIT, d, m, nullptr);
    const long double ll = e.ok ? (long double)e.out[0] : (long double)NAN;
    release(e.r);
    return ll;
}
extern "C" void apop_b200_probit_score(apop_data* d, gsl_vector* g, apop_model* m) {
    LOCK;
    if (!g) return;
    GlmEval e = glm_entry(GLM_PROBIT, d, m, nullptr);
    if (e.ok) store_score(g, e.out.data() + MAX_ACC, e.beta.size(), 1.0);
    else fill_nan(g);
    release(e.r);
}

// One fused multinomial-logit pass; out = [LL | G in pack order (j*(C-1)+c)], all ranks summed.
static bool eval_logit(Resident* r, const std::vector<double>& B, size_t C1, const std::vector<double>& cats,
                       std::vector<double>& out) {
    Runtime& R = rt();
    const int K = r->k;
    if (B.size() != (size_t)K * C1) return set_error("parameters are %zu doubles; the data has %d columns x %zu free categories", B.size(), K, C1);
    if (K < 1 || K > MAX_K_REG) return set_error("k = %d columns: the fused kernels cover 1..%d", K, MAX_K_REG);
    if (C1 < 2 || C1 > (size_t)LOGIT_MAX_C1) return set_error("%zu outcome categories: the multinomial kernel covers 3..%d", C1 + 1, LOGIT_MAX_C1 + 1);
    if (!r->has_y || r->has_w) return set_error("logit needs an outcome vector and no weights column");
    std::vector<double> key(B);
    key.insert(key.end(), cats.begin(), cats.end());
    if (r->memo_valid && r->memo_fam == 200 + (int)C1 && r->memo_beta.size() == key.size() &&
        memcmp(r->memo_beta.data(), key.data(), key.size() * sizeof(double)) == 0) {
        out = r->memo_out;
        return true;
    }
    static int bps_cache[MAX_K_REG + 1][LOGIT_MAX_C1 + 1];
    int& bps = bps_cache[K][C1];
    if (!bps) bps = logit_blocks_per_sm(K, (int)C1);
    if (bps <= 0) return set_error("multinomial logit kernel unavailable for k=%d, C=%zu", K, C1 + 1);
    const int KB = logit_kb(K);
    const int NOUT = 1 + KB * (int)C1;
    LogitLaunch L;
    L.tiles = r->tiles; L.n_rows = r->n_rows; L.n_tiles = r->n_tiles; L.k = K;
    L.grid = grid_for(bps, r->n_tiles, LOGIT_THREADS / 32);
    if (!ensure_partials((size_t)L.grid * NOUT)) return false;
    L.partials = R.d_partials; L.ticket = R.d_ticket; L.out = R.d_out; L.stream = R.stream;
    L.fin = make_fin(NOUT);
    LogitParams prm;
    memset(&prm, 0, sizeof prm);
    memcpy(prm.B, B.data(), B.size() * sizeof(double));
    for (size_t c = 0; c < C1 && c < 8; c++) prm.cats[c] = cats[c];
    time_begin();
    if (!cuda_ok(logit_launch((int)C1, L, prm), "logit kernel launch")) return false;
    time_end_record();
    R.launches++;
    if (!reduce_and_fetch(NOUT, L.fin)) return false;
    time_collect();
    out.assign(R.h_out, R.h_out + 1 + (size_t)K * C1);
    r->memo_valid = true; r->memo_fam = 200 + (int)C1; r->memo_beta = key; r->memo_out = out;
    return true;
}

static bool logit_categories(apop_data* d, apop_model* m, size_t C, std::vector<double>& cats) {
    const apop_data* pg = category_page(m && m->data ? m->data : d);
    if (!pg) pg = category_page(d);
    cats.resize(C);
    for (size_t c = 0; c < C; c++) cats[c] = (double)c;
    if (pg && pg->vector && pg->vector->size >= C)
        for (size_t c = 0; c < C; c++) cats[c] = pg->vector->data[c * pg->vector->stride];
    return true;
}

extern "C" long double apop_b200_logit_ll(apop_data* d, apop_model* m) {
    LOCK;
    if (!m || !m->parameters || !d || !d->matrix) { set_error("NULL model, parameters or data"); return NAN; }
    const size_t C1 = m->parameters->matrix ? m->parameters->matrix->size2 : 1;
    std::vector<double> cats;
    logit_categories(d, m, C1 + 1, cats);
    if (C1 == 1) {
        const double aux[4] = {cats[0], cats[1], 0, 0};
        GlmEval e = glm_entry(GLM_LOGIT2, d, m, aux);
        const long double ll = e.ok ? (long double)e.out[0] : (long double)NAN;
        release(e.r);
        return ll;
    }
    std::vector<double> B, out;
    pack_parameters(m->parameters, B);
    Resident* r = acquire(d);
    if (!r) { d->error = 'a'; return NAN; }
    const bool ok = eval_logit(r, B, C1, cats, out);
    release(r);
    if (!ok) { m->error = 'd'; return NAN; }
    return (long double)out[0];
}
extern "C" void apop_b200_logit_score(apop_data* d, gsl_vector* g, apop_model* m) {
    LOCK;
    if (!g) return;
    if (!m || !m->parameters || !d || !d->matrix) { fill_nan(g); return; }
    const size_t C1 = m->parameters->matrix ? m->parameters->matrix->size2 : 1;
    std::vector<double> cats;
    logit_categories(d, m, C1 + 1, cats);
    if (C1 == 1) {
        const double aux[4] = {cats[0], cats[1], 0, 0};
        GlmEval e = glm_entry(GLM_LOGIT2, d, m, aux);
        if (e.ok) store_score(g, e.out.data() + MAX_ACC, e.beta.size(), 1.0);
        else fill_nan(g);
        release(e.r);
        return;
    }
    std::vector<double> B, out;
    pack_parameters(m->parameters, B);
    Resident* r = acquire(d);
    if (!r) { fill_nan(g); return; }
    const bool ok = eval_logit(r, B, C1, cats, out);
    release(r);
    if (ok) store_score(g, out.data() + 1, B.size(), 1.0);
    else { m->error = 'd'; fill_nan(g); }
}

// sum_i lgamma(y_i + 1): beta-independent, one pass per residency
static bool lgamma_const(Resident* r, double& v) {
    if (!r->have_lgamma_y) {
        double o[CELLS_NOUT];
        if (!eval_cells(r, CELLS_POISSON, 0.0, o, /*matrix=*/false)) return false;
        r->lgamma_y = o[3];
        r->have_lgamma_y = true;
    }
    v = r->lgamma_y;
    return true;
}
extern "C" long double apop_b200_poisson_reg_ll(apop_data* d, apop_model* m) {
    LOCK;
    GlmEval e = glm_entry(GLM_POISSON_REG, d, m, nullptr);
    long double ll = NAN;
    double lg;
    if (e.ok && lgamma_const(e.r, lg)) ll = (long double)e.out[0] - (long double)lg;
    release(e.r);
    return ll;
}
extern "C" void apop_b200_poisson_reg_score(apop_data* d, gsl_vector* g, apop_model* m) {
    LOCK;
    if (!g) return;
    GlmEval e = glm_entry(GLM_POISSON_REG, d, m, nullptr);
    if (e.ok) store_score(g, e.out.data() + MAX_ACC, e.beta.size(), 1.0);
    else fill_nan(g);
    release(e.r);
}

// ---- OLS -------------------------------------------------------------------
struct OlsEval {
    bool ok = false;
    long double ll = NAN, sigma2 = NAN;
    std::vector<double> g;
};
static OlsEval ols_entry(apop_data* d, apop_model* m) {
    OlsEval o;
    if (d && !d->vector) {
        set_error("OLS data is not prepped (no outcome vector): run ols_shuffle / the model's prep first");
        if (m) m->error = 'd';
        return o;
    }
    GlmEval e = glm_entry(GLM_OLS, d, m, nullptr);
    if (!e.ok) { release(e.r); return o; }
    double n, mc, vc;
    if (!global_counts(e.r, n, mc, vc)) { release(e.r); return o; }
    const long double S1 = e.out[0], S2 = e.out[1], S3 = e.r->has_w ? e.out[2] : 0.0L;
    // sample variance of the errors (gsl_stats_variance: n-1 in the denominator)
    long double var = (S2 - S1 * S1 / n) / (n - 1);
    o.sigma2 = var;
    const apop_data* ev = named_page(m->parameters, "<Error variance>");
    long double var_ll = var;
    if (ev) var_ll = ev->matrix ? ev->matrix->data[0] : (ev->vector ? ev->vector->data[0] : var);
    const long double sigma = sqrtl(var_ll);
    // sum_i log( phi(e_i; sigma) w_i ) with phi the Gaussian density
    o.ll = -S2 / (2 * var_ll) - n * logl(sigma * 2.50662827463100050241576528481104525L) + S3;
    o.g.assign(e.out.begin() + MAX_ACC, e.out.end());
    for (auto& v : o.g) v = (double)(v / var);
    o.ok = true;
    release(e.r);
    return o;
}
extern "C" long double apop_b200_ols_ll(apop_data* d, apop_model* m) {
    LOCK;
    OlsEval o = ols_entry(d, m);
    return o.ok ? o.ll : (long double)NAN;
}
extern "C" void apop_b200_ols_score(apop_data* d, gsl_vector* g, apop_model* m) {
    LOCK;
    if (!g) return;
    OlsEval o = ols_entry(d, m);
    if (o.ok) store_score(g, o.g.data(), o.g.size(), 1.0);
    else fill_nan(g);
}

// ---- iid Poisson / Normal ----------------------------------------------------
static bool poisson_consts(Resident* r, double* o) {
    if (!r->have_poisson) {
        if (!eval_cells(r, CELLS_POISSON, 0.0, r->pois)) return false;
        r->have_poisson = true;
    }
    memcpy(o, r->pois, sizeof r->pois);
    return true;
}
extern "C" long double apop_b200_poisson_ll(apop_data* d, apop_model* m) {
    LOCK;
    if (!m || !m->parameters || !d) { set_error("NULL model, parameters or data"); return NAN; }
    std::vector<double> p;
    if (!pack_parameters(m->parameters, p)) return NAN;
    Resident* r = acquire(d);
    if (!r) return NAN;
    double o[CELLS_NOUT], n, mc, vc;
    long double ll = NAN;
    if (poisson_consts(r, o) && global_counts(r, n, mc, vc)) {
        const long double lambda = p[0], ln_l = logl(lambda);
        if (o[4] > 0) ll = -INFINITY;
        else ll = ln_l * ((long double)o[0] + o[2]) - ((long double)o[1] + o[3]) - ((long double)mc + vc) * lambda;
    }
    release(r);
    return ll;
}
extern "C" void apop_b200_poisson_score(apop_data* d, gsl_vector* g, apop_model* m) {
    LOCK;
    if (!g) return;
    std::vector<double> p;
    if (!m || !m->parameters || !d || !pack_parameters(m->parameters, p)) { fill_nan(g); return; }
    Resident* r = acquire(d);
    if (!r) { fill_nan(g); return; }
    double o[CELLS_NOUT], n, mc, vc;
    if (poisson_consts(r, o) && global_counts(r, n, mc, vc))
        g->data[0] = (double)((long double)o[0] / p[0] - ((long double)mc + vc));  // matrix cells only in the sum
    else fill_nan(g);
    release(r);
}

static bool normal_sums(Resident* r, double mu, double* o) {
    // memo on mu: LL and score at the same parameters share one pass
    const double key[5] = {mu, 0, 0, 0, 0};
    if (r->memo_valid && r->memo_fam == 100 && r->memo_beta.size() == 5 && !memcmp(r->memo_beta.data(), key, sizeof key)) {
        memcpy(o, r->memo_out.data(), CELLS_NOUT * sizeof(double));
        return true;
    }
    if (!eval_cells(r, CELLS_NORMAL, mu, o)) return false;
    r->memo_valid = true; r->memo_fam = 100;
    r->memo_beta.assign(key, key + 5);
    r->memo_out.assign(o, o + CELLS_NOUT);
    return true;
}
extern "C" long double apop_b200_normal_ll(apop_data* d, apop_model* m) {
    LOCK;
    std::vector<double> p;
    if (!m || !m->parameters || !d || !pack_parameters(m->parameters, p) || p.size() < 2) { set_error("NULL model, parameters or data"); return NAN; }
    Resident* r = acquire(d);
    if (!r) return NAN;
    double o[CELLS_NOUT], n, mc, vc;
    long double ll = NAN;
    if (normal_sums(r, p[0], o) && global_counts(r, n, mc, vc)) {
        const long double sd = p[1], lnpi = 1.14472988584940017414342735135305871L, ln2 = 0.693147180559945309417232121458176568L;
        ll = -((long double)o[1] + o[3]) / (2 * sd * sd) - ((long double)mc + vc) * ((lnpi + ln2) / 2 + logl(sd));
    }
    release(r);
    return ll;
}
extern "C" void apop_b200_normal_score(apop_data* d, gsl_vector* g, apop_model* m) {
    LOCK;
    if (!g) return;
    std::vector<double> p;
    if (!m || !m->parameters || !d || !pack_parameters(m->parameters, p) || p.size() < 2) { fill_nan(g); return; }
    Resident* r = acquire(d);
    if (!r) { fill_nan(g); return; }
    double o[CELLS_NOUT], n, mc, vc;
    if (normal_sums(r, p[0], o) && global_counts(r, n, mc, vc) && g->size >= 2) {
        const long double sd = p[1];
        g->data[0] = (double)(((long double)o[0] + o[2]) / (sd * sd));
        g->data[g->stride] = (double)(((long double)o[1] + o[3]) / (sd * sd * sd) - ((long double)mc + vc) / sd);
    } else fill_nan(g);
    release(r);
}

// ---- exponential, gamma, lognormal ----------------------------------------
// sum x, sum_{x != 0} ln x and the number of zero cells do not depend on the parameters:
// one pass per residency, like the Poisson sums
static bool gamma_consts(Resident* r, double* o) {
    if (!r->have_gamma) {
        if (!eval_cells(r, CELLS_GAMMA, 0.0, r->gam)) return false;
        r->have_gamma = true;
    }
    memcpy(o, r->gam, sizeof r->gam);
    return true;
}
static long double digamma_l(long double x) {   // gsl_sf_psi: recurrence to x >= 12, then the asymptotic series
    long double r = 0;
    while (x < 12) { r -= 1 / x; x += 1; }
    const long double f = 1 / (x * x);
    return r + logl(x) - 0.5L / x - f * (1.0L / 12 - f * (1.0L / 120 - f * (1.0L / 252 - f * (1.0L / 240 - f * (1.0L / 132 - f * (691.0L / 32760 - f / 12))))));
}
struct IidEval { bool ok = false; double o[CELLS_NOUT]; double cells = 0; std::vector<double> p; };
static IidEval iid_entry(apop_data* d, apop_model* m, size_t nparams, int mode, double p0_index) {
    IidEval e;
    if (!m || !m->parameters || !d || !pack_parameters(m->parameters, e.p) || e.p.size() < nparams) { set_error("NULL model, parameters or data"); return e; }
    Resident* r = acquire(d);
    if (!r) return e;
    double n, mc, vc;
    bool ok = global_counts(r, n, mc, vc);
    if (ok && mode == CELLS_GAMMA) ok = gamma_consts(r, e.o);
    else if (ok) {   // lognormal: memo on mu, like the Normal
        const double mu = e.p[(size_t)p0_index];
        const double key[5] = {mu, 1, 0, 0, 0};
        if (r->memo_valid && r->memo_fam == 101 && r->memo_beta.size() == 5 && !memcmp(r->memo_beta.data(), key, sizeof key))
            memcpy(e.o, r->memo_out.data(), sizeof e.o);
        else if ((ok = eval_cells(r, mode, mu, e.o))) {
            r->memo_valid = true; r->memo_fam = 101;
            r->memo_beta.assign(key, key + 5);
            r->memo_out.assign(e.o, e.o + CELLS_NOUT);
        }
    }
    e.cells = mc + vc;
    e.ok = ok;
    release(r);
    return e;
}
extern "C" long double apop_b200_exponential_ll(apop_data* d, apop_model* m) {
    LOCK;
    IidEval e = iid_entry(d, m, 1, CELLS_GAMMA, 0);
    if (!e.ok) return NAN;
    const long double mu = e.p[0];
    return -((long double)e.o[0] + e.o[2]) / mu - (long double)e.cells * logl(mu);
}
extern "C" void apop_b200_exponential_score(apop_data* d, gsl_vector* g, apop_model* m) {
    LOCK;
    if (!g) return;
    IidEval e = iid_entry(d, m, 1, CELLS_GAMMA, 0);
    if (!e.ok) { fill_nan(g); return; }
    const long double mu = e.p[0];
    g->data[0] = (double)(((long double)e.o[0] + e.o[2]) / (mu * mu) - (long double)e.cells / mu);
}
extern "C" long double apop_b200_gamma_ll(apop_data* d, apop_model* m) {
    LOCK;
    IidEval e = iid_entry(d, m, 2, CELLS_GAMMA, 0);
    if (!e.ok) return NAN;
    const long double a = e.p[0], b = e.p[1];
    const long double nz = (long double)e.cells - e.o[4];   // cells with x != 0
    return (a - 1) * ((long double)e.o[1] + e.o[3]) - ((long double)e.o[0] + e.o[2]) / b - nz * (lgammal(a) + a * logl(b));
}
extern "C" void apop_b200_gamma_score(apop_data* d, gsl_vector* g, apop_model* m) {
    LOCK;
    if (!g) return;
    IidEval e = iid_entry(d, m, 2, CELLS_GAMMA, 0);
    if (!e.ok || g->size < 2) { fill_nan(g); return; }
    const long double a = e.p[0], b = e.p[1];
    // a_callback has no x != 0 guard: a zero cell sends dLL/da to -inf (log 0)
    g->data[0] = (e.o[4] > 0) ? -INFINITY
                              : (double)(((long double)e.o[1] + e.o[3]) - (long double)e.cells * (digamma_l(a) + logl(b)));
    g->data[g->stride] = (double)(((long double)e.o[0] + e.o[2]) / (b * b) - (long double)e.cells * a / b);
}
extern "C" long double apop_b200_lognormal_ll(apop_data* d, apop_model* m) {
    LOCK;
    IidEval e = iid_entry(d, m, 2, CELLS_LOGNORMAL, 0);
    if (!e.ok) return NAN;
    const long double sd = e.p[1], lnpi = 1.14472988584940017414342735135305871L, ln2 = 0.693147180559945309417232121458176568L;
    return -((long double)e.o[1] + e.o[3]) / (2 * sd * sd) - ((long double)e.o[0] + e.o[2]) - (long double)e.cells * ((lnpi + ln2) / 2 + logl(sd));
}
extern "C" void apop_b200_lognormal_score(apop_data* d, gsl_vector* g, apop_model* m) {
    LOCK;
    if (!g) return;
    IidEval e = iid_entry(d, m, 2, CELLS_LOGNORMAL, 0);
    if (!e.ok || g->size < 2) { fill_nan(g); return; }
    const long double mu = e.p[0], sd = e.p[1];
    g->data[0] = (double)((((long double)e.o[0] + e.o[2]) - mu * (long double)e.cells) / (sd * sd));
    g->data[g->stride] = (double)(((long double)e.o[1] + e.o[3]) / (sd * sd * sd) - (long double)e.cells / sd);
}

extern "C" double apop_b200_map_sum(apop_data* d, apop_b200_map_fn fn, double p0, char part) {
    LOCK;
    if (!d) { set_error("NULL data"); return 0; }  // apop_map_sum returns 0 on NULL input
    Resident* r = acquire(d);
    if (!r) return NAN;
    int mode, slot = 0;
    switch (fn) {
    case APOP_B200_MAP_IDENTITY: mode = CELLS_IDENTITY; break;
    case APOP_B200_MAP_DEV: mode = CELLS_NORMAL; break;
    case APOP_B200_MAP_DEV2: mode = CELLS_NORMAL; slot = 1; break;
    case APOP_B200_MAP_POISSON: mode = CELLS_POISSON; break;
    case APOP_B200_MAP_LOG: mode = CELLS_LOG; break;
    case APOP_B200_MAP_EXP: mode = CELLS_EXP; break;
    default: release(r); set_error("unknown map function"); return NAN;
    }
    double o[CELLS_NOUT];
    double res = NAN;
    if (eval_cells(r, mode, p0, o)) {
        const bool um = (part != 'v'), uv = (part != 'm');
        if (fn == APOP_B200_MAP_POISSON) {
            // sum over cells of  x==0 ? 0 : p0*x - lgamma(x+1), -inf if any cell is invalid
            res = (o[4] > 0) ? -INFINITY
                             : (double)((long double)p0 * ((um ? o[0] : 0.0L) + (uv ? o[2] : 0.0L)) -
                                        ((um ? (long double)o[1] : 0.0L) + (uv ? (long double)o[3] : 0.0L)));
        } else
            res = (double)((um ? (long double)o[slot] : 0.0L) + (uv ? (long double)o[2 + slot] : 0.0L));
    }
    release(r);
    return res;
}

// ---- apop_dot(d, parameters): the N x c index X.B, copied back in chunks ----------------------
extern "C" int apop_b200_dot(apop_data* d, const double* B, size_t k, size_t c, double* out) {
    LOCK;
    if (!d || !B || !out || !c) { set_error("apop_b200_dot: NULL argument"); return 1; }
    Resident* r = acquire(d);
    if (!r) return 1;
    Runtime& R = rt();
    bool ok = false;
    double *dB = nullptr, *dOut = nullptr;
    do {
        if ((size_t)r->k != k || r->k < 1 || r->k > MAX_K_REG) {
            // the reference's dimension error: error char 'd' (apop_linear_algebra.m4.c:347-349)
            set_error("apop_b200_dot: the data has %d columns, the parameters %zu rows", r->k, k);
            d->error = 'd';
            break;
        }
        size_t chunk = ((size_t)16 << 20) / c;   // <= 128 MB of output per chunk
        if (chunk > r->n_rows) chunk = r->n_rows;
        if (!chunk) { ok = true; break; }
        if (!cuda_ok(cudaMalloc(&dB, k * c * sizeof(double)), "cudaMalloc") || !cuda_ok(cudaMalloc(&dOut, chunk * c * sizeof(double)), "cudaMalloc")) break;
        if (!cuda_ok(cudaMemcpyAsync(dB, B, k * c * sizeof(double), cudaMemcpyHostToDevice, R.stream), "H2D B")) break;
        ok = true;
        for (size_t r0 = 0; ok && r0 < r->n_rows; r0 += chunk) {
            const size_t rows = (r->n_rows - r0 < chunk) ? r->n_rows - r0 : chunk;
            ok = cuda_ok(xbeta_launch(r->tiles, r0, rows, r->k, r->cols, (int)c, dB, dOut, R.stream), "xbeta launch") &&
                 cuda_ok(cudaMemcpyAsync(out + r0 * c, dOut, rows * c * sizeof(double), cudaMemcpyDeviceToHost, R.stream), "D2H") &&
                 cuda_ok(cudaStreamSynchronize(R.stream), "sync");
            R.launches++;
        }
    } while (0);
    if (dB) cudaFree(dB);
    if (dOut) cudaFree(dOut);
    release(r);
    return ok ? 0 : 1;
}

// ---- prep reductions on the device --------------------------------------
extern "C" int apop_b200_col_sums(apop_data* d, int mode, double* out) {
    LOCK;
    if (!d || !out || mode < 0 || mode > 2) { set_error("apop_b200_col_sums: bad argument"); return 1; }
    Resident* r = acquire(d);
    if (!r) return 1;
    Runtime& R = rt();
    bool ok = false;
    if (r->k < 1 || r->k > MAX_K_REG) set_error("apop_b200_col_sums: 1..%d columns", MAX_K_REG);
    else {
        const int grid = grid_for(4, r->n_tiles, CELLS_THREADS / 32);
        if (ensure_partials((size_t)grid * MAX_K_REG)) {
            const FinalizeCtx fin = make_fin(MAX_K_REG);
            time_begin();
            ok = cuda_ok(colsums_launch(r->tiles, r->n_rows, r->n_tiles, r->k, r->cols, mode, R.d_partials, R.d_ticket,
                                        R.d_out, grid, R.stream, fin), "column sums launch");
            time_end_record();
            R.launches++;
            ok = ok && reduce_and_fetch(MAX_K_REG, fin);
            if (ok) { time_collect(); memcpy(out, R.h_out, r->k * sizeof(double)); }
        }
    }
    release(r);
    return ok ? 0 : 1;
}
// probit_prep's start point (model/apop_probit.c:65-80): beta0_j = exp(mean_i ln|x_ij|), zeros skipped
extern "C" int apop_b200_probit_start(apop_data* d, double* beta0) {
    LOCK;
    if (!d || !beta0) { set_error("apop_b200_probit_start: NULL argument"); return 1; }
    Resident* r = acquire(d);
    if (!r) return 1;
    double n, mc, vc;
    const int k = r->k;
    const bool okc = global_counts(r, n, mc, vc);
    const bool keep = !r->transient;
    release(r);
    if (!okc) return 1;
    if (apop_b200_col_sums(d, 0, beta0)) return 1;
    (void)keep;
    for (int j = 0; j < k; j++) beta0[j] = (double)expl((long double)beta0[j] / n);
    return 0;
}

// ---- generic fused value + gradient ------------------------------------------
extern "C" double apop_b200_ll_score(apop_b200_family family, apop_data* d, const double* beta, size_t P, double* score) {
    LOCK;
    if (!d || !beta) { set_error("NULL data or beta"); return NAN; }
    // a throw-away model header around beta: reuse the reference-signature entries
    apop_model m;
    memset(&m, 0, sizeof m);
    apop_data prm;
    memset(&prm, 0, sizeof prm);
    gsl_vector v = {P, 1, const_cast<double*>(beta), nullptr, 0};
    gsl_matrix mat = {0, 0, 0, const_cast<double*>(beta), nullptr, 0};
    gsl_vector g = {P, 1, score, nullptr, 0};
    m.parameters = &prm;
    m.data = d;
    prm.vector = &v;
    long double ll = NAN;
    switch (family) {
    case APOP_B200_PROBIT: ll = apop_b200_probit_ll(d, &m); if (score) apop_b200_probit_score(d, &g, &m); break;
    case APOP_B200_LOGIT: {
        const size_t k = d->matrix ? d->matrix->size2 : 0;
        if (!k || P % k) { set_error("logit: P = %zu is not a multiple of k = %zu", P, k); return NAN; }
        prm.vector = nullptr; prm.matrix = &mat;
        mat.size1 = k; mat.size2 = P / k; mat.tda = P / k;
        ll = apop_b200_logit_ll(d, &m); if (score) apop_b200_logit_score(d, &g, &m); break;
    }
    case APOP_B200_POISSON: ll = apop_b200_poisson_ll(d, &m); if (score) apop_b200_poisson_score(d, &g, &m); break;
    case APOP_B200_NORMAL: ll = apop_b200_normal_ll(d, &m); if (score) apop_b200_normal_score(d, &g, &m); break;
    case APOP_B200_OLS: ll = apop_b200_ols_ll(d, &m); if (score) apop_b200_ols_score(d, &g, &m); break;
    case APOP_B200_POISSON_REG: ll = apop_b200_poisson_reg_ll(d, &m); if (score) apop_b200_poisson_reg_score(d, &g, &m); break;
    case APOP_B200_EXPONENTIAL: ll = apop_b200_exponential_ll(d, &m); if (score) apop_b200_exponential_score(d, &g, &m); break;
    case APOP_B200_GAMMA: ll = apop_b200_gamma_ll(d, &m); if (score) apop_b200_gamma_score(d, &g, &m); break;
    case APOP_B200_LOGNORMAL: ll = apop_b200_lognormal_ll(d, &m); if (score) apop_b200_lognormal_score(d, &g, &m); break;
    default: set_error("unknown family %d", (int)family);
    }
    return (double)ll;
}

// ---- not built yet -------------------------------------------------------------
// H = X' diag(w) [X | y] in one pass; out has K*(K+1) doubles, H[j][c] at j*(K+1)+c
static bool eval_xtwx(Resident* r, int mode, const double* beta, const double* aux, std::vector<double>& out) {
    Runtime& R = rt();
    const int K = r->k;
    if (K < 1 || K > MAX_K_REG) return set_error("k = %d columns: the fused kernels cover 1..%d", K, MAX_K_REG);
    if (!r->has_y) return set_error("the data set has no outcome vector (run the model's prep first)");
    static int bps_cache[XTWX_MODE_COUNT][MAX_K_REG + 1];
    int& bps = bps_cache[mode][K];
    if (!bps) bps = xtwx_blocks_per_sm(mode, K);
    if (bps <= 0) return set_error("X'WX kernel unavailable for k=%d", K);
    const int KB = xtwx_kb(K), NC = KB + 1, NOUT = KB * NC;
    XtwxLaunch L;
    L.tiles = r->tiles; L.n_rows = r->n_rows; L.n_tiles = r->n_tiles; L.k = K; L.cols = r->cols;
    L.has_w = (r->has_w && mode == XTWX_GRAM) ? 1 : 0;
    L.grid = grid_for(bps, r->n_tiles, XTWX_THREADS / 32);
    if (!ensure_partials((size_t)L.grid * NOUT)) return false;
    L.partials = R.d_partials; L.ticket = R.d_ticket; L.out = R.d_out; L.stream = R.stream;
    L.fin = make_fin(NOUT);
    XtwxParams prm;
    memset(&prm, 0, sizeof prm);
    if (beta) memcpy(prm.beta, beta, K * sizeof(double));
    if (aux) memcpy(prm.aux, aux, 4 * sizeof(double));
    time_begin();
    if (!cuda_ok(xtwx_launch(mode, L, prm), "X'WX kernel launch")) return false;
    time_end_record();
    R.launches++;
    if (!reduce_and_fetch(NOUT, L.fin)) return false;
    time_collect();
    out.assign((size_t)K * (K + 1), 0.0);
    for (int j = 0; j < K; j++) {
        for (int c = 0; c < K; c++) out[(size_t)j * (K + 1) + c] = R.h_out[j * NC + c];
        out[(size_t)j * (K + 1) + K] = R.h_out[j * NC + KB];
    }
    return true;
}

extern "C" int apop_b200_ols_gram(apop_data* d, double* xtx, double* xty) {
    LOCK;
    if (!d || !xtx || !xty) { set_error("apop_b200_ols_gram: NULL argument"); return 1; }
    Resident* r = acquire(d);
    if (!r) return 1;
    std::vector<double> out;
    const bool ok = eval_xtwx(r, XTWX_GRAM, nullptr, nullptr, out);
    if (ok) {
        const int K = r->k;
        for (int j = 0; j < K; j++) {
            for (int c = 0; c < K; c++) xtx[j * K + c] = out[(size_t)j * (K + 1) + c];
            xty[j] = out[(size_t)j * (K + 1) + K];
        }
    }
    release(r);
    return ok ? 0 : 1;
}

extern "C" int apop_b200_information(apop_b200_family family, apop_data* d, const double* beta, size_t P, double* H) {
    LOCK;
    if (!d || !beta || !H) { set_error("apop_b200_information: NULL argument"); return 1; }
    Resident* r = acquire(d);
    if (!r) return 1;
    bool ok = false;
    std::vector<double> out;
    double scale = 1.0;
    if ((size_t)r->k != P) set_error("apop_b200_information: P = %zu but the data has %d columns (binary-outcome families only)", P, r->k);
    else if (family == APOP_B200_PROBIT) ok = eval_xtwx(r, XTWX_PROBIT, beta, nullptr, out);
    else if (family == APOP_B200_LOGIT) ok = eval_xtwx(r, XTWX_LOGIT2, beta, nullptr, out);
    else if (family == APOP_B200_POISSON_REG) ok = eval_xtwx(r, XTWX_POISSON_REG, beta, nullptr, out);
    else if (family == APOP_B200_OLS) {
        // -d2LL/dbeta2 at fixed sigma: X'WX / sigma^2, sigma^2 = sample variance of the errors
        std::vector<double> mom;
        double n, mc, vc;
        if (eval_glm(r, GLM_OLS, beta, P, nullptr, mom) && global_counts(r, n, mc, vc) && eval_xtwx(r, XTWX_GRAM, nullptr, nullptr, out)) {
            const long double var = ((long double)mom[1] - (long double)mom[0] * mom[0] / n) / (n - 1);
            scale = (double)(1.0L / var);
            ok = true;
        }
    } else set_error("apop_b200_information: family %d has no k x k information here", (int)family);
    if (ok)
        for (size_t j = 0; j < P; j++)
            for (size_t c = 0; c < P; c++) H[j * P + c] = out[j * (P + 1) + c] * scale;
    release(r);
    return ok ? 0 : 1;
}
// m parameter vectors in one pass on the DMMA path.  ll[m]; score[P*m] replicate-major.
extern "C" int apop_b200_ll_score_batched(apop_b200_family family, apop_data* d, const double* B, size_t P, size_t m,
                                          int use_counts, double* ll, double* score) {
    LOCK;
    if (!d || !B || !ll || !m) { set_error("apop_b200_ll_score_batched: NULL argument"); return 1; }
    GlmFam fam;
    if (family == APOP_B200_PROBIT) fam = GLM_PROBIT;
    else if (family == APOP_B200_LOGIT) fam = GLM_LOGIT2;
    else if (family == APOP_B200_POISSON_REG) fam = GLM_POISSON_REG;
    else { set_error("batched path: family %d is not a binary-outcome / GLM family", (int)family); return 1; }
    Resident* r = acquire(d);
    if (!r) return 1;
    Runtime& R = rt();
    bool ok = false;
    double *dB = nullptr, *dOut = nullptr;
    do {
        const int K = r->k;
        if ((size_t)K != P) { set_error("batched path: P = %zu but the data has %d columns", P, K); break; }
        if (K < 1 || K > MAX_K_REG || !r->has_y || r->has_w) { set_error("batched path: needs 1..%d columns, an outcome vector and no weights", MAX_K_REG); break; }
        if (use_counts && (!r->counts || r->counts_m < m)) { set_error("batched path: no resident resample counts for %zu replicates (call apop_b200_bootstrap_counts)", m); break; }
        if (use_counts && fam == GLM_POISSON_REG) { set_error("batched path: resample counts are not supported for the Poisson regression constant"); break; }
        const int KB = batched_kb(K), NOUT = 1 + KB;
        static int mb_env = -1;
        if (mb_env < 0) { const char* e = getenv("APOP_B200_BATCH_MB"); mb_env = (e && e[0] == '2') ? 2 : 1; }
        const int mb = mb_env, ctas_per_sm = (mb == 1) ? 2 : 1;   // matches the kernels' launch bounds
        const int reps_per_cta = (BATCH_THREADS / 32) * mb * 8;
        const int n_repgroups = (int)((m + reps_per_cta - 1) / reps_per_cta);
        long long n_rowgroups = (long long)R.num_sms * ctas_per_sm / n_repgroups;
        if (n_rowgroups < 1) n_rowgroups = 1;
        if ((unsigned long long)n_rowgroups > r->n_tiles) n_rowgroups = r->n_tiles ? (long long)r->n_tiles : 1;
        BatchedLaunch L;
        L.tiles = r->tiles; L.n_rows = r->n_rows; L.n_tiles = r->n_tiles; L.k = K; L.cols = r->cols;
        L.P = (int)P; L.m = (int)m; L.counts = use_counts ? r->counts : nullptr; L.m_pad = (int)r->counts_m;
        // binary logit: the two category values (defaults 0, 1)
        L.aux[0] = 0.0; L.aux[1] = 1.0; L.aux[2] = L.aux[3] = 0.0;
        const apop_data* pg = category_page(d);
        if (fam == GLM_LOGIT2 && pg && pg->vector && pg->vector->size >= 2) { L.aux[0] = pg->vector->data[0]; L.aux[1] = pg->vector->data[pg->vector->stride]; }
        L.n_rowgroups = (int)n_rowgroups; L.grid = (int)(n_rowgroups * n_repgroups);
        L.want_score = score != nullptr; L.stream = R.stream; L.mb = mb;
        if (!ensure_partials((size_t)n_rowgroups * m * NOUT)) break;
        L.partials = R.d_partials;
        if (!cuda_ok(cudaMalloc(&dB, P * m * sizeof(double)), "cudaMalloc(B)")) break;
        if (!cuda_ok(cudaMalloc(&dOut, m * NOUT * sizeof(double)), "cudaMalloc(out)")) break;
        if (!cuda_ok(cudaMemcpyAsync(dB, B, P * m * sizeof(double), cudaMemcpyHostToDevice, R.stream), "H2D B")) break;
        L.B = dB; L.out = dOut;
        time_begin();
        if (!cuda_ok(batched_launch((int)fam, L), "batched kernel launch")) break;
        time_end_record();
        R.launches += 2;
        if (R.comm && R.nranks > 1 &&
            !nccl_ok(R.nccl.AllReduce(dOut, dOut, m * NOUT, 8, 0, R.comm, R.stream), "ncclAllReduce")) break;
        std::vector<double> h(m * NOUT);
        if (!cuda_ok(cudaMemcpyAsync(h.data(), dOut, m * NOUT * sizeof(double), cudaMemcpyDeviceToHost, R.stream), "D2H")) break;
        if (!cuda_ok(cudaStreamSynchronize(R.stream), "stream sync")) break;
        time_collect();
        double lg = 0.0;
        if (fam == GLM_POISSON_REG && !lgamma_const(r, lg)) break;
        for (size_t q = 0; q < m; q++) {
            ll[q] = h[q * NOUT] - lg;
            if (score) memcpy(score + q * P, h.data() + q * NOUT + 1, P * sizeof(double));
        }
        ok = true;
    } while (0);
    if (dB) cudaFree(dB);
    if (dOut) cudaFree(dOut);
    release(r);
    return ok ? 0 : 1;
}

// resident resample counts back to the host as an n_rows x m row-major uint8 matrix (tests)
extern "C" int apop_b200_counts_download(const apop_data* d, size_t m, unsigned char* out) {
    LOCK;
    Runtime& R = rt();
    auto it = R.reg.find(d);
    if (it == R.reg.end() || !it->second->counts || it->second->counts_m < m) { set_error("no resident counts"); return 1; }
    Resident* r = it->second;
    cudaSetDevice(R.device);
    const size_t m_pad = r->counts_m, bytes = r->n_tiles * m_pad * TILE_ROWS;
    std::vector<unsigned char> h(bytes);
    if (!cuda_ok(cudaMemcpy(h.data(), r->counts, bytes, cudaMemcpyDeviceToHost), "D2H counts")) return 1;
    for (size_t i = 0; i < r->n_rows; i++) {
        const int rr = (int)(i & 31);
        const int pos = ((rr & 7) >> 1) * 8 + (rr >> 3) * 2 + (rr & 1);
        for (size_t q = 0; q < m; q++) out[i * m + q] = h[((i >> 5) * m_pad + q) * TILE_ROWS + pos];
    }
    return 0;
}

// m exact multinomial resamples of the (global) rows, kept resident next to the data
extern "C" int apop_b200_bootstrap_counts(apop_data* d, size_t m, uint64_t seed) {
    LOCK;
    if (!d || !m) { set_error("apop_b200_bootstrap_counts: NULL argument"); return 1; }
    if (!require_ready()) return 1;
    Runtime& R = rt();
    auto it = R.reg.find(d);
    if (it == R.reg.end()) { set_error("apop_b200_bootstrap_counts: pin the data set first"); return 1; }
    Resident* r = it->second;
    cudaSetDevice(R.device);
    if (r->counts) { cudaFree(r->counts); r->counts = nullptr; r->counts_m = 0; }
    const size_t m_pad = (m + 3) / 4 * 4;
    const size_t bytes = r->n_tiles * m_pad * TILE_ROWS;
    if (!bytes) return 0;
    if (!cuda_ok(cudaMalloc(&r->counts, bytes), "cudaMalloc(counts)")) return 1;
    cudaMemsetAsync(r->counts, 0, bytes, R.stream);
    // the resample is of the global data set: ranks share seed, each keeps its own rows
    double n_global = (double)r->n_rows, offset = 0.0;
    if (R.nranks > 1) {
        // exclusive prefix of the shard sizes = this rank's first global row
        std::vector<double> sizes(R.nranks, 0.0);
        sizes[R.rank] = (double)r->n_rows;
        if (!allreduce_host(sizes.data(), R.nranks)) return 1;
        n_global = 0;
        for (int q = 0; q < R.nranks; q++) { if (q < R.rank) offset += sizes[q]; n_global += sizes[q]; }
    }
    if (!cuda_ok(bootstrap_counts_launch(r->counts, r->n_rows, (size_t)offset, (size_t)n_global, (int)m, (int)m_pad, seed, R.stream), "bootstrap counts") ||
        !cuda_ok(cudaStreamSynchronize(R.stream), "counts sync")) return 1;
    r->counts_m = m_pad;
    R.launches++;
    return 0;
}

// --------------------------------------------------------------------------
//  installing into the reference's dispatch
// --------------------------------------------------------------------------
typedef long double (*ll_fn)(apop_data*, apop_model*);
struct Entry { ll_fn ll; apop_score_type score; };
static Entry entry_of(apop_b200_family f) {
    switch (f) {
    case APOP_B200_PROBIT: return {apop_b200_probit_ll, apop_b200_probit_score};
    case APOP_B200_LOGIT: return {apop_b200_logit_ll, apop_b200_logit_score};
    case APOP_B200_POISSON: return {apop_b200_poisson_ll, apop_b200_poisson_score};
    case APOP_B200_NORMAL: return {apop_b200_normal_ll, apop_b200_normal_score};
    case APOP_B200_OLS: return {apop_b200_ols_ll, apop_b200_ols_score};
    case APOP_B200_POISSON_REG: return {apop_b200_poisson_reg_ll, apop_b200_poisson_reg_score};
    case APOP_B200_EXPONENTIAL: return {apop_b200_exponential_ll, apop_b200_exponential_score};
    case APOP_B200_GAMMA: return {apop_b200_gamma_ll, apop_b200_gamma_score};
    case APOP_B200_LOGNORMAL: return {apop_b200_lognormal_ll, apop_b200_lognormal_score};
    default: return {nullptr, nullptr};
    }
}
static std::unordered_map<apop_model*, ll_fn>& saved_ll() {
    static std::unordered_map<apop_model*, ll_fn> m;
    return m;
}
extern "C" int apop_b200_register(apop_model* model, apop_b200_family family) {
    LOCK;
    if (!model) { set_error("apop_b200_register: NULL model"); return 1; }
    Entry e = entry_of(family);
    if (!e.ll) { set_error("apop_b200_register: unknown family"); return 1; }
    // the registry lives in libapophenia (apop_vtables.c:68) or in this repo's host mirror
    typedef int (*vtable_add_fn)(char const*, void*, unsigned long);
    vtable_add_fn add = (vtable_add_fn)dlsym(RTLD_DEFAULT, "apop_vtable_add");
    if (!add) { set_error("apop_vtable_add not found in this process: load libapophenia (or the host mirror) first"); return 1; }
    if (!saved_ll().count(model)) saved_ll()[model] = model->log_likelihood;
    model->log_likelihood = e.ll;
    // hash = address of log_likelihood (apop.m4.h:532)
    return add("apop_score", (void*)e.score, (unsigned long)apop_score_hash(model));
}
extern "C" int apop_b200_unregister(apop_model* model) {
    LOCK;
    if (!model) return 1;
    typedef int (*vtable_drop_fn)(char const*, unsigned long);
    vtable_drop_fn drop = (vtable_drop_fn)dlsym(RTLD_DEFAULT, "apop_vtable_drop");
    if (drop) drop("apop_score", (unsigned long)apop_score_hash(model));
    auto it = saved_ll().find(model);
    if (it != saved_ll().end()) {
        model->log_likelihood = it->second;
        saved_ll().erase(it);
    }
    return 0;
}
