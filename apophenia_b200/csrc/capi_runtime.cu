// capi_runtime.cu -- process-wide device context of libapop_b200: init/finalize, error text,
// pass timers, NCCL binding, the CUDA-IPC inboxes of the fused allreduce, and the helpers that
// complete a reduction launch (make_fin / reduce_and_fetch).
//
// There is no CPU fallback anywhere in this library: without a B200 every entry point fails
// loudly (NaN + apop_b200_last_error).
#include "capi_internal.cuh"

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

bool require_ready() {
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
    if (!cuda_ok(cudaStreamCreateWithFlags(&R.copy_stream, cudaStreamNonBlocking), "cudaStreamCreate")) return 1;
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
    { const char* e = getenv("APOP_B200_TIMING"); R.timing = e && e[0] == '1'; }
    R.ready = true;
    R.err.clear();
    return 0;
}

void release_host_registration(Resident* r) {
    for (auto& w : r->host_registered) cudaHostUnregister(w.first);
    if (!r->host_registered.empty()) cudaGetLastError();   // a range the caller already unmapped is not an error worth keeping
    r->host_registered.clear();
}
void free_resident(Resident* r) {
    if (!r) return;
    release_host_registration(r);
    if (r->tiles) cudaFree(r->tiles);
    if (r->counts) cudaFree(r->counts);
    if (r->lgy) cudaFree(r->lgy);
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
    for (int b2 = 0; b2 < 2; b2++) {
        if (R.pin_host[b2]) cudaFreeHost(R.pin_host[b2]);
        if (R.pin_dev[b2]) cudaFree(R.pin_dev[b2]);
        R.pin_host[b2] = nullptr; R.pin_dev[b2] = nullptr;
        // events belong to the device they were created on: a later init on another device makes its own
        if (R.pin_ev[b2]) { cudaEventDestroy(R.pin_ev[b2]); R.pin_ev[b2] = nullptr; }
        if (R.pin_copied[b2]) { cudaEventDestroy(R.pin_copied[b2]); R.pin_copied[b2] = nullptr; }
    }
    R.pin_cap = 0;
    for (int b2 = 0; b2 < 2; b2++) {
        if (R.reg_dev[b2]) { cudaFree(R.reg_dev[b2]); R.reg_dev[b2] = nullptr; }
        if (R.reg_hy[b2]) { cudaFreeHost(R.reg_hy[b2]); R.reg_hy[b2] = nullptr; }
        if (R.reg_copied[b2]) { cudaEventDestroy(R.reg_copied[b2]); R.reg_copied[b2] = nullptr; }
        if (R.reg_packed[b2]) { cudaEventDestroy(R.reg_packed[b2]); R.reg_packed[b2] = nullptr; }
    }
    if (R.reg_edge) { cudaFreeHost(R.reg_edge); R.reg_edge = nullptr; }
    R.reg_dev_cap = R.reg_hy_cap = 0;
    cudaFree(R.d_vals);
    if (R.d_beta) { cudaFree(R.d_beta); R.d_beta = nullptr; }
    cudaEventDestroy(R.ev0);
    cudaEventDestroy(R.ev1);
    cudaStreamDestroy(R.stream);
    if (R.copy_stream) { cudaStreamDestroy(R.copy_stream); R.copy_stream = nullptr; }
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
extern "C" void apop_b200_set_timing(int on) { rt().timing = on != 0; }
// Page-locked host memory for the caller's matrix / vector blocks: the upload then needs no host copy
extern "C" void* apop_b200_host_alloc(size_t bytes) {
    LOCK;
    if (!require_ready()) return nullptr;
    cudaSetDevice(rt().device);
    void* p = nullptr;
    if (!cuda_ok(cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable), "cudaHostAlloc")) return nullptr;
    return p;
}
extern "C" void apop_b200_host_free(void* p) {
    if (!p) return;
    LOCK;
    if (rt().ready) cudaSetDevice(rt().device);
    cudaFreeHost(p);
}
extern "C" int apop_b200_last_pin_mode(double* register_GBps) {
    if (register_GBps) *register_GBps = rt().last_register_GBps;
    return rt().last_pin_mode;
}
extern "C" void apop_b200_set_local_only(int on) {
    LOCK;
    Runtime& R = rt();
    if (R.local_only == (on != 0)) return;
    R.local_only = on != 0;
    // memoised evaluations and cached constants are sums over a set of ranks: drop them with the set
    R.comm_epoch++;
}

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
bool nccl_ok(int rc, const char* what) {
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
    if (R.comm || R.p2p_active || R.p2p_local) apop_b200_comm_finalize();   // never leak an earlier communicator
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
    R.comm_failed = false;
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
FinalizeCtx make_fin(int nout) {
    Runtime& R = rt();
    FinalizeCtx f;
    memset(&f, 0, sizeof f);
    f.nranks = 1;
    static unsigned long long spin_env = 0;   // ~10 s of 64 ns sleeps + loads, then the error flag (tests shorten it)
    if (!spin_env) { const char* e = getenv("APOP_B200_P2P_SPIN_LIMIT"); spin_env = e ? strtoull(e, nullptr, 10) : 0; if (!spin_env) spin_env = 40000000ull; }
    f.spin_limit = spin_env;
    f.err = R.h_err_dev;
    // local-only mode, or a communicator that already failed (reduce_and_fetch then reports it): no exchange
    if (R.local_only || R.comm_failed) { f.host_out = R.h_out_dev; return f; }
    if (R.nranks > 1 && R.p2p_active && nout <= R.p2p_slot) {
        f.nranks = R.nranks; f.rank = R.rank; f.seq = ++R.p2p_seq; f.slot = R.p2p_slot;
        for (int p = 0; p < R.nranks; p++) { f.inbox[p] = R.p2p_inbox[p]; f.flags[p] = R.p2p_flags[p]; }
        f.host_out = R.h_out_dev;
    } else if (R.nranks <= 1)
        f.host_out = R.h_out_dev;
    return f;
}
// complete a reduction launched with `fin`: `n` doubles, summed over all ranks, end up in h_out
bool reduce_and_fetch(size_t n, const FinalizeCtx& fin) {
    Runtime& R = rt();
    if (R.comm_failed && R.nranks > 1 && !R.local_only) {
        // sticky: after a timeout the ranks' sequence counters are out of step, and a later pass could add a
        // slot its peer has already overwritten -- fail every reduction until the communicator is rebuilt
        cudaStreamSynchronize(R.stream);
        return set_error("the communicator failed earlier (a peer never arrived); call apop_b200_comm_finalize and re-initialise it");
    }
    if (fin.host_out) {   // the kernel published (and, with peers, exchanged) the result itself
        if (!cuda_ok(cudaStreamSynchronize(R.stream), "stream sync")) return false;
        if (*R.h_err) {
            *R.h_err = 0;
            R.comm_failed = true;
            return set_error("fused allreduce: a peer rank never arrived (sequence %llu)", fin.seq);
        }
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
bool allreduce_host(double* v, int n) {
    Runtime& R = rt();
    if (R.nranks <= 1 || R.local_only) return true;
    if (!cuda_ok(cudaMemcpyAsync(R.d_vals, v, n * sizeof(double), cudaMemcpyHostToDevice, R.stream), "H2D")) return false;
    const FinalizeCtx fin = make_fin(n);
    if (!cuda_ok(exchange_launch(R.d_vals, n, R.d_ticket, R.d_out, R.stream, fin), "exchange launch")) return false;
    if (!reduce_and_fetch(n, fin)) return false;
    memcpy(v, R.h_out, n * sizeof(double));
    return true;
}
bool allreduce_scalar(double* v) { return allreduce_host(v, 1); }

