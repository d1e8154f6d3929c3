// capi_register.cu -- installing the device entries into the reference's dispatch
// (model->log_likelihood + the apop_score vtable).
#include "capi_internal.cuh"

using namespace apb;

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
    case APOP_B200_BERNOULLI: return {apop_b200_bernoulli_ll, apop_b200_bernoulli_score};
    case APOP_B200_BETA: return {apop_b200_beta_ll, apop_b200_beta_score};
    case APOP_B200_ZIPF: return {apop_b200_zipf_ll, nullptr};         // no analytic score in the reference: numeric gradient over the device LL
    case APOP_B200_DIRICHLET: return {apop_b200_dirichlet_ll, apop_b200_dirichlet_score};
    case APOP_B200_T: return {apop_b200_t_ll, nullptr};
    case APOP_B200_YULE: return {apop_b200_yule_ll, apop_b200_yule_score};
    default: return {nullptr, nullptr};
    }
}
bool is_device_entry(void* fn) {
    for (int f = 0; f < (int)APOP_B200_MODEL_COUNT; f++)
        if (fn && (void*)entry_of((apop_b200_family)f).ll == fn) return true;
    return false;
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
    if (!e.score) return 0;   // apop_score then falls back to the numeric gradient (apop_model.m4.c:296-299)
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

