// runtime.cuh -- per-process device context of libapop_b200 (one process per GPU).
#pragma once
#include <cuda_runtime.h>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../include/apop_b200.h"
#include "glm_kernel.cuh"
#include "cells.cuh"
#include "layout.cuh"

namespace apb {

// what a host apop_data looked like when it was uploaded
struct HostSig {
    const void *m = nullptr, *v = nullptr, *w = nullptr;
    size_t m1 = 0, m2 = 0, tda = 0, vn = 0, vs = 0, wn = 0;
    bool operator==(const HostSig& o) const {
        return m == o.m && v == o.v && w == o.w && m1 == o.m1 && m2 == o.m2 && tda == o.tda && vn == o.vn && vs == o.vs && wn == o.wn;
    }
};
// A data set kept in HBM between optimizer iterations.
struct Resident {
    size_t n_rows = 0;
    int k = 0;
    bool has_y = false, has_w = false;
    int cols = 0;           // k + has_y + has_w
    size_t n_tiles = 0;
    double* tiles = nullptr;
    bool swz = false;        // tiles are in the swizzled form (rows r ^ 8 in odd columns): only the multinomial logit kernels read it
    bool synthetic = false;
    bool y_col0 = false;     // pinned by the device prep: column 0 of the host matrix is the outcome (ols_shuffle done here)
    bool dirty = false;      // host copy changed: re-upload before next use
    // rows of the whole (all ranks) data set, exchanged once per communicator
    double n_global = 0.0;
    unsigned long long n_global_epoch = ~0ull;
    unsigned long long consts_epoch = 0;   // communicator the cached sums below belong to
    bool transient = false;  // uploaded for one call only
    // beta-independent constants, computed on first use
    bool have_lgamma_y = false;
    double lgamma_y = 0.0;   // sum_i lgamma(y_i + 1)
    bool have_poisson = false;
    double pois[CELLS_NOUT];
    bool have_gamma = false;   // sum x, sum ln x over non-zero cells, zero-cell count (exponential / gamma)
    double gam[CELLS_NOUT];
    // parameter-independent cell / column sums of the other iid families, keyed by kernel mode
    std::unordered_map<int, std::vector<double>> consts;
    // last evaluation (f then df at the same beta costs one pass)
    bool memo_valid = false;
    int memo_fam = -1;
    std::vector<double> memo_beta, memo_out;
    // resample multiplicities for the batched path
    unsigned char* counts = nullptr;
    size_t counts_m = 0;
    // per-row lgamma(y+1) in tile order (Poisson regression with resample counts), built on first use
    double* lgy = nullptr;
    // identity of the host buffers this copy was made from: a different apop_data allocated at the same
    // address (or the same one reshaped) is re-uploaded instead of being served from the stale copy
    HostSig sig;
    // the outcome column was recoded on the device (apop_b200_prep_outcome): a dirty re-upload of a host
    // copy that still holds the raw outcomes re-applies the table
    bool recoded = false, recode_pending = false;
    CatTable cats;
    // windows of the caller's matrix this library page-locked for a zero-copy upload and keeps page-locked
    // while the copy is resident (a registration cache: released by unpin / invalidate / re-upload / finalize)
    std::vector<std::pair<void*, size_t>> host_registered;
    // batched path: passes left for which the exact kernel is used because the last shared-centre pass found the
    // parameter vectors too far apart (then the Taylor form only adds work); re-probed when it reaches 0
    long active_replicate = -1;   // >= 0: single-beta multinomial passes weight every row by this replicate's resample count
    int batched_exact_passes = 0;
    double batched_last_exact_fraction = 0.0;
};

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, struct NcclId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
struct NcclId { char internal[128]; };

struct Runtime {
    bool ready = false;
    int device = -1;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timing = false;      // CUDA-event timing of every pass: opt-in (APOP_B200_TIMING=1 / apop_b200_set_timing(1))
    double* d_partials = nullptr;
    size_t partials_cap = 0;  // doubles
    unsigned int* d_ticket = nullptr;
    double* d_out = nullptr;  // OUT_CAP doubles
    double* h_out = nullptr;  // pinned + mapped: the last CTA writes results straight into it
    double* h_out_dev = nullptr;  // device address of h_out
    int* h_err = nullptr;     // mapped flag: a peer never arrived in the fused allreduce
    int* h_err_dev = nullptr;
    double* d_vals = nullptr; // staging for host scalars to be summed across ranks
    double* d_beta = nullptr; // parameters of the wide (k > 32) kernels
    bool auto_pin = false;
    // pinned / device staging pair for the pipelined upload in apop_b200_pin
    double* pin_host[2] = {nullptr, nullptr};
    double* pin_dev[2] = {nullptr, nullptr};
    cudaEvent_t pin_ev[2] = {nullptr, nullptr};
    cudaEvent_t pin_copied[2] = {nullptr, nullptr};
    cudaStream_t copy_stream = nullptr;   // H2D copies of the upload run here, back to back, while the main stream tiles
    size_t pin_cap = 0;   // doubles per staging buffer
    // zero-copy upload (registered windows of the caller's matrix): two device landing buffers, two pinned
    // staging buffers for the outcome / weights of a window, and their events
    char* reg_dev[2] = {nullptr, nullptr};
    double* reg_hy[2] = {nullptr, nullptr};
    char* reg_edge = nullptr;                 // pinned, two pages: the partial first / last page of a zero-copy source
    size_t reg_dev_cap = 0, reg_hy_cap = 0;   // bytes
    cudaEvent_t reg_copied[2] = {nullptr, nullptr}, reg_packed[2] = {nullptr, nullptr};
    int last_pin_mode = 0;                    // 0 staged, 1 registered windows (apop_b200_last_pin_mode)
    double last_register_GBps = 0.0;
    std::unordered_map<const apop_data*, Resident*> reg;
    // stats
    unsigned long long launches = 0, n_passes = 0;
    double last_ms = 0.0, total_ms = 0.0;
    // multi-GPU
    NcclApi nccl;
    void* comm = nullptr;
    int nranks = 1, rank = 0;
    unsigned long long comm_epoch = 0;   // bumped whenever the set of ranks changes
    // fused allreduce over NVLink peer memory (CUDA IPC), see common.cuh grid_finalize
    bool p2p_active = false;
    bool local_only = false;      // evaluate this rank's rows only (checks and single-rank baselines inside a multi-rank job)
    bool comm_failed = false;     // a collective timed out: every later reduction fails until the communicator is rebuilt
    int p2p_slot = 0;
    unsigned long long p2p_seq = 0;
    double* p2p_local = nullptr;                 // this rank's inbox + flags allocation
    double* p2p_inbox[P2P_MAX_RANKS] = {nullptr};
    unsigned long long* p2p_flags[P2P_MAX_RANKS] = {nullptr};
    std::string err;
};
constexpr size_t OUT_CAP = 4096;

Runtime& rt();
std::recursive_mutex& rt_mutex();
bool set_error(const char* fmt, ...);
bool cuda_ok(cudaError_t e, const char* what);
bool ensure_partials(size_t doubles);

}  // namespace apb
