"""ctypes front end of libapop_hostmini.so: the host-side mirror of the reference's
callers (apop_estimate -> apop_prep -> apop_maximum_likelihood, the apop_score vtable,
apop_data_pack/unpack; see csrc/hostmini.c and csrc/mle.c).

`estimate()` is the reference's `apop_estimate(data, model)` flow with the model's
log_likelihood / score replaced by the device entries through `apop_b200_register`
-- the same installation a maintainer performs on a real libapophenia
(INTEGRATION.md).
"""
import ctypes as C
import os

import numpy as np

from . import abi
from .abi import apop_data, apop_model, gsl_vector

_H = None
_PD, _PM, _PV = C.POINTER(apop_data), C.POINTER(apop_model), C.POINTER(gsl_vector)


class apop_mle_settings(C.Structure):
    _fields_ = [("starting_pt", C.POINTER(C.c_double)), ("tolerance", C.c_double), ("max_iterations", C.c_int),
                ("method", C.c_char * 32), ("verbose", C.c_int), ("step_size", C.c_double), ("delta", C.c_double),
                ("relative_tolerance", C.c_int)]


class apop_mle_report(C.Structure):
    _fields_ = [("status", C.c_int), ("iterations", C.c_int), ("f_evals", C.c_int), ("df_evals", C.c_int),
                ("ll", C.c_double), ("gradient_norm", C.c_double), ("seconds", C.c_double)]


class apop_mcmc_settings(C.Structure):
    _fields_ = [("periods", C.c_int), ("burnin", C.c_double), ("target_accept_rate", C.c_double),
                ("initial_scale", C.c_double), ("seed", C.c_uint64), ("n_chains", C.c_int)]


def lib():
    """dlopen the host mirror with RTLD_GLOBAL so apop_b200_register finds apop_vtable_add"""
    global _H
    if _H is None:
        if not os.path.exists(abi.HOSTMINI_PATH):
            raise abi.B200Error(f"{abi.HOSTMINI_PATH} is not built (make -C apophenia_b200/csrc)")
        h = C.CDLL(abi.HOSTMINI_PATH, mode=C.RTLD_GLOBAL)
        h.apop_data_alloc.restype = _PD
        h.apop_data_alloc.argtypes = [C.c_size_t, C.c_size_t, C.c_int]
        h.apop_data_free.argtypes = [_PD]
        h.apop_data_add_page.restype = _PD
        h.apop_data_add_page.argtypes = [_PD, _PD, C.c_char_p]
        h.apop_data_get_page.restype = _PD
        h.apop_data_get_page.argtypes = [_PD, C.c_char_p]
        h.apop_data_pack.restype = _PV
        h.apop_data_pack.argtypes = [_PD, _PV]
        h.apop_data_unpack.argtypes = [_PV, _PD]
        h.gsl_vector_alloc.restype = _PV
        h.gsl_vector_alloc.argtypes = [C.c_size_t]
        h.gsl_vector_free.argtypes = [_PV]
        h.apop_vtable_add.argtypes = [C.c_char_p, C.c_void_p, C.c_ulong]
        h.apop_vtable_get.restype = C.c_void_p
        h.apop_vtable_get.argtypes = [C.c_char_p, C.c_ulong]
        h.apop_vtable_drop.argtypes = [C.c_char_p, C.c_ulong]
        h.apop_model_copy.restype = _PM
        h.apop_model_copy.argtypes = [_PM]
        h.apop_model_free.argtypes = [_PM]
        h.apop_prep.argtypes = [_PD, _PM]
        h.apop_estimate.restype = _PM
        h.apop_estimate.argtypes = [_PD, _PM]
        h.apop_maximum_likelihood.argtypes = [_PD, _PM]
        h.apop_log_likelihood.restype = C.c_double
        h.apop_log_likelihood.argtypes = [_PD, _PM]
        h.apop_score.argtypes = [_PD, _PV, _PM]
        h.apop_numerical_gradient.restype = _PV
        h.apop_numerical_gradient.argtypes = [_PD, _PM, C.c_double]
        h.apop_mle_settings_defaults.restype = apop_mle_settings
        h.apop_mle_settings_add.restype = C.POINTER(apop_mle_settings)
        h.apop_mle_settings_add.argtypes = [_PM, apop_mle_settings]
        h.apop_mle_last_report.restype = apop_mle_report
        h.apop_b200_category_table.restype = _PD
        h.apop_b200_category_table.argtypes = [_PD]
        h.apop_b200_ols_shuffle.argtypes = [_PD]
        h.apop_b200_family_of.argtypes = [_PM]
        h.apop_data_covariance.restype = _PD
        h.apop_data_covariance.argtypes = [_PD]
        h.apop_model_numerical_covariance.restype = _PD
        h.apop_model_numerical_covariance.argtypes = [_PD, _PM]
        h.apop_model_hessian.restype = _PD
        h.apop_model_hessian.argtypes = [_PD, _PM, C.c_double]
        h.apop_b200_observed_information.argtypes = [_PD, _PM, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        h.apop_settings_group_alloc.restype = C.c_void_p
        h.apop_settings_group_alloc.argtypes = [_PM, C.c_char_p, C.c_void_p, C.c_size_t]
        h.apop_bootstrap_cov.restype = _PD
        h.apop_bootstrap_cov.argtypes = [_PD, _PM, C.c_uint64, C.c_int, C.POINTER(_PD)]
        h.apop_mcmc_settings_defaults.restype = apop_mcmc_settings
        h.apop_model_metropolis.restype = _PD
        h.apop_model_metropolis.argtypes = [_PD, _PM, apop_mcmc_settings]
        h.apop_b200_metropolis_chains.restype = _PD
        h.apop_b200_metropolis_chains.argtypes = [_PD, _PM, apop_mcmc_settings]
        h.apop_b200_install_estimates()
        _H = h
    return _H


def matrix_of(data_ptr):
    """copy of an apop_data's matrix as numpy"""
    m = data_ptr.contents.matrix.contents
    a = np.ctypeslib.as_array(m.data, shape=(m.size1, m.tda))
    return np.array(a[:, :m.size2])


def bootstrap_cov(data_ptr, est_ptr, iterations=1000, seed=8):
    """apop_bootstrap_cov on a fitted, registered model; returns (cov, replicate estimates)"""
    h = lib()
    store = _PD()
    cov = h.apop_bootstrap_cov(data_ptr, est_ptr, int(seed), int(iterations), C.byref(store))
    if not cov:
        raise abi.B200Error("apop_bootstrap_cov failed: " + abi.last_error())
    return matrix_of(cov), matrix_of(store)


def metropolis(data_ptr, model_ptr, periods=6000, burnin=0.05, target_accept_rate=0.35, initial_scale=1.0, seed=8,
               n_chains=1):
    """apop_model_metropolis (n_chains == 1: the unmodified one-LL-per-step caller) or the
    lock-step batched chains; returns draws as (n_chains, kept, P)"""
    h = lib()
    s = apop_mcmc_settings(int(periods), float(burnin), float(target_accept_rate), float(initial_scale), int(seed), int(n_chains))
    out = h.apop_model_metropolis(data_ptr, model_ptr, s) if n_chains == 1 else h.apop_b200_metropolis_chains(data_ptr, model_ptr, s)
    if not out:
        raise abi.B200Error("metropolis failed: " + abi.last_error())
    a = matrix_of(out)
    return a.reshape(n_chains, -1, a.shape[1])


_MODEL_SYMS = {abi.PROBIT: "apop_probit", abi.LOGIT: "apop_logit", abi.POISSON: "apop_poisson",
               abi.NORMAL: "apop_normal", abi.OLS: "apop_ols", abi.POISSON_REG: "apop_poisson_regression",
               abi.EXPONENTIAL: "apop_exponential", abi.GAMMA: "apop_gamma", abi.LOGNORMAL: "apop_lognormal",
               abi.BERNOULLI: "apop_bernoulli", abi.BETA: "apop_beta", abi.ZIPF: "apop_zipf",
               abi.DIRICHLET: "apop_dirichlet", abi.TDIST: "apop_t_distribution", abi.YULE: "apop_yule"}


def model_object(family):
    """the exported `apop_model *apop_probit` etc. of the mirror"""
    return _PM.in_dll(lib(), _MODEL_SYMS[family])


def data_from_numpy(matrix, vector=None):
    """apop_data_alloc + copy: the data set as a reference user would hold it.
    With vector=None the outcome is expected in matrix column 0 (prep shuffles it out)."""
    h = lib()
    m = np.ascontiguousarray(matrix, dtype=np.float64)
    d = h.apop_data_alloc(0 if vector is None else len(vector), m.shape[0], m.shape[1])
    C.memmove(d.contents.matrix.contents.data, m.ctypes.data, m.nbytes)
    if vector is not None:
        v = np.ascontiguousarray(vector, dtype=np.float64)
        C.memmove(d.contents.vector.contents.data, v.ctypes.data, v.nbytes)
    return d


def parameters_of(model_ptr):
    h = lib()
    pv = h.apop_data_pack(model_ptr.contents.parameters, None)
    out = np.array([pv.contents.data[i] for i in range(pv.contents.size)])
    h.gsl_vector_free(pv)
    return out


def mle_settings(model_ptr, method="BFGS cg", tolerance=1e-5, max_iterations=5000, step_size=0.05, verbose=0,
                 relative_tolerance=False, starting_pt=None, keep=None):
    h = lib()
    s = h.apop_mle_settings_defaults()
    s.method = method.encode()
    s.tolerance, s.max_iterations, s.step_size, s.verbose = tolerance, max_iterations, step_size, verbose
    s.relative_tolerance = int(relative_tolerance)
    if starting_pt is not None:
        sp = np.ascontiguousarray(starting_pt, dtype=np.float64)
        if keep is not None:
            keep.append(sp)
        s.starting_pt = sp.ctypes.data_as(C.POINTER(C.c_double))
    return h.apop_mle_settings_add(model_ptr, s)


def estimate(family, data_ptr, pin=True, **settings):
    """apop_estimate(data, model) with the device entries installed.  Returns
    (fitted model pointer, parameters in pack order, report dict)."""
    h = lib()
    b = abi.load_library()
    base = model_object(family)
    m = h.apop_model_copy(base)
    abi.check(b.apop_b200_register(m, int(family)), "apop_b200_register")
    keep = []
    if settings:
        mle_settings(m, keep=keep, **settings)
    # prep first (it reshapes the data: factors, ols_shuffle), then make the result resident
    h.apop_prep(data_ptr, m)
    if pin:
        abi.check(b.apop_b200_pin(data_ptr), "apop_b200_pin")
    est = h.apop_estimate(data_ptr, m)
    rep = h.apop_mle_last_report()
    report = {f: getattr(rep, f) for f, _ in apop_mle_report._fields_}
    params = parameters_of(est)
    h.apop_model_free(m)
    return est, params, report
