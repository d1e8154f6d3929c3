// capi_internal.cuh -- declarations shared by the capi_*.cu translation units (not installed).
#pragma once
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
#include <chrono>

#define LOCK std::lock_guard<std::recursive_mutex> lock_(apb::rt_mutex())

// runtime / communication (capi_runtime.cu)
bool require_ready();
void free_resident(apb::Resident* r);
void release_host_registration(apb::Resident* r);
bool nccl_ok(int rc, const char* what);
apb::FinalizeCtx make_fin(int nout);
bool reduce_and_fetch(size_t n, const apb::FinalizeCtx& fin);
bool allreduce_host(double* v, int n);
bool allreduce_scalar(double* v);
// residency (capi_residency.cu)
bool upload(apb::Resident* r, const apop_data* d);
apb::HostSig host_sig(const apop_data* d);
// layout the caller's kernels read: every kernel but the multinomial logit's takes the plain tiles; LAYOUT_ANY leaves
// the copy as it is (the caller converts with ensure_layout once it knows)
enum { LAYOUT_PLAIN = 0, LAYOUT_SWZ = 1, LAYOUT_ANY = 2 };
apb::Resident* acquire(apop_data* d, int layout = LAYOUT_PLAIN);
bool ensure_layout(apb::Resident* r, bool swz);
void release(apb::Resident* r);
// evaluators (capi_models.cu, capi_batched.cu)
int grid_for(int blocks_per_sm, size_t n_tiles, int warps_per_block);
void time_begin();
void time_end_record();
void time_collect();
bool eval_glm(apb::Resident* r, apb::GlmFam fam, const double* beta, size_t P, const double* aux, std::vector<double>& out);
bool eval_cells(apb::Resident* r, int mode, double p0, double* out5, bool matrix = true, double p1 = 0.0, double p2 = 0.0);
bool eval_logit(apb::Resident* r, const std::vector<double>& B, size_t C1, const std::vector<double>& cats, std::vector<double>& out);
bool eval_xtwx(apb::Resident* r, int mode, const double* beta, const double* aux, std::vector<double>& out);
bool global_counts(apb::Resident* r, double& n_rows, double& m_cells, double& v_cells);
bool lgamma_const(apb::Resident* r, double& v);
bool pack_parameters(const apop_data* p, std::vector<double>& out);
const apop_data* category_page(const apop_data* d);
const apop_data* named_page(const apop_data* d, const char* title);
const apop_data* named_page_or_self(const apop_data* d, const char* title);
bool is_device_entry(void* log_likelihood_fn);   // capi_register.cu: one of this library's log-likelihood entries
void fill_nan(gsl_vector* g);
void store_score(gsl_vector* g, const double* src, size_t n, double scale);
