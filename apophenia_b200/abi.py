"""ctypes view of include/apop_b200.h: the reference's public structs and the C-ABI.

This is the Python host side of the drop-in boundary.  It builds `apop_data` /
`apop_model` structs with the reference's exact layout (apop.m4.h:61-115, GSL's
gsl_vector / gsl_matrix) around numpy buffers and calls libapop_b200.so through
plain pointers.  The product path is the CUDA library; if it cannot be loaded
or no B200 is present every call fails loudly -- there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libapop_b200.so")
HOSTMINI_PATH = os.path.join(_PKG, "libapop_hostmini.so")

PROBIT, LOGIT, POISSON, NORMAL, OLS, POISSON_REG, EXPONENTIAL, GAMMA, LOGNORMAL = range(9)
BERNOULLI, BETA, ZIPF, DIRICHLET, TDIST, YULE = range(9, 15)
FAMILY_NAMES = {PROBIT: "probit", LOGIT: "logit", POISSON: "poisson", NORMAL: "normal", OLS: "ols",
                POISSON_REG: "poisson_reg", EXPONENTIAL: "exponential", GAMMA: "gamma", LOGNORMAL: "lognormal",
                BERNOULLI: "bernoulli", BETA: "beta", ZIPF: "zipf", DIRICHLET: "dirichlet", TDIST: "t", YULE: "yule"}
MAP_IDENTITY, MAP_DEV, MAP_DEV2, MAP_POISSON, MAP_LOG, MAP_EXP, MAP_ABS, MAP_SQRT, MAP_SCALE, MAP_POW = range(10)
(ROW_SUM, ROW_MEAN, ROW_SUMSQ, ROW_MAX, ROW_MIN, ROW_DOT, ROW_PROBIT_LL, ROW_LOGIT2_LL, ROW_POISSON_REG_LL,
 ROW_OLS_ERROR) = range(10)


class gsl_block(C.Structure):
    _fields_ = [("size", C.c_size_t), ("data", C.POINTER(C.c_double))]


class gsl_vector(C.Structure):
    _fields_ = [("size", C.c_size_t), ("stride", C.c_size_t), ("data", C.POINTER(C.c_double)),
                ("block", C.POINTER(gsl_block)), ("owner", C.c_int)]


class gsl_matrix(C.Structure):
    _fields_ = [("size1", C.c_size_t), ("size2", C.c_size_t), ("tda", C.c_size_t),
                ("data", C.POINTER(C.c_double)), ("block", C.POINTER(gsl_block)), ("owner", C.c_int)]


class apop_name(C.Structure):
    _fields_ = [("title", C.c_char_p), ("vector", C.c_char_p), ("col", C.POINTER(C.c_char_p)),
                ("row", C.POINTER(C.c_char_p)), ("text", C.POINTER(C.c_char_p)),
                ("colct", C.c_int), ("rowct", C.c_int), ("textct", C.c_int),
                ("colhash", C.POINTER(C.c_ulong)), ("rowhash", C.POINTER(C.c_ulong)),
                ("texthash", C.POINTER(C.c_ulong))]


class apop_data(C.Structure):
    pass


apop_data._fields_ = [("vector", C.POINTER(gsl_vector)), ("matrix", C.POINTER(gsl_matrix)),
                      ("names", C.POINTER(apop_name)), ("text", C.c_void_p),
                      ("textsize", C.c_size_t * 2), ("weights", C.POINTER(gsl_vector)),
                      ("more", C.POINTER(apop_data)), ("error", C.c_char)]


class apop_settings_type(C.Structure):
    _fields_ = [("name", C.c_char * 101), ("name_hash", C.c_ulong), ("setting_group", C.c_void_p),
                ("copy", C.c_void_p), ("free", C.c_void_p)]


class apop_model(C.Structure):
    pass


LL_FN = C.CFUNCTYPE(C.c_longdouble, C.POINTER(apop_data), C.POINTER(apop_model))
SCORE_FN = C.CFUNCTYPE(None, C.POINTER(apop_data), C.POINTER(gsl_vector), C.POINTER(apop_model))

apop_model._fields_ = [("name", C.c_char * 101), ("vsize", C.c_int), ("msize1", C.c_int),
                       ("msize2", C.c_int), ("dsize", C.c_int),
                       ("data", C.POINTER(apop_data)), ("parameters", C.POINTER(apop_data)),
                       ("info", C.POINTER(apop_data)),
                       ("estimate", C.c_void_p), ("p", C.c_void_p), ("log_likelihood", C.c_void_p),
                       ("cdf", C.c_void_p), ("constraint", C.c_void_p), ("draw", C.c_void_p),
                       ("prep", C.c_void_p), ("settings", C.POINTER(apop_settings_type)),
                       ("more", C.c_void_p), ("more_size", C.c_size_t), ("error", C.c_char)]

assert C.sizeof(gsl_vector) == 40 and C.sizeof(gsl_matrix) == 48
assert apop_data.weights.offset == 48 and apop_data.error.offset == 64
assert apop_model.log_likelihood.offset == 160 and apop_model.error.offset == 224

_dp = C.POINTER(C.c_double)

# every symbol include/apop_b200.h declares: (restype, argtypes)
_PD, _PM, _PV = C.POINTER(apop_data), C.POINTER(apop_model), C.POINTER(gsl_vector)
SIGNATURES = {
    "apop_b200_init": (C.c_int, [C.c_int]),
    "apop_b200_finalize": (None, []),
    "apop_b200_last_error": (C.c_char_p, []),
    "apop_b200_launch_count": (C.c_ulonglong, []),
    "apop_b200_last_pass_ms": (C.c_double, []),
    "apop_b200_pass_timer_reset": (None, []),
    "apop_b200_pass_timer_total_ms": (C.c_double, [C.POINTER(C.c_ulonglong)]),
    "apop_b200_comm_unique_id": (C.c_int, [C.c_void_p]),
    "apop_b200_comm_init": (C.c_int, [C.c_int, C.c_int, C.c_void_p]),
    "apop_b200_comm_finalize": (C.c_int, []),
    "apop_b200_comm_size": (C.c_int, []),
    "apop_b200_comm_p2p_handle": (C.c_int, [C.c_int, C.c_void_p]),
    "apop_b200_comm_p2p_init": (C.c_int, [C.c_int, C.c_int, C.c_void_p]),
    "apop_b200_comm_p2p_active": (C.c_int, []),
    "apop_b200_pin": (C.c_int, [_PD]),
    "apop_b200_unpin": (C.c_int, [_PD]),
    "apop_b200_invalidate": (C.c_int, [_PD]),
    "apop_b200_is_pinned": (C.c_int, [_PD]),
    "apop_b200_set_auto_pin": (None, [C.c_int]),
    "apop_b200_last_pin_mode": (C.c_int, [C.POINTER(C.c_double)]),
    "apop_b200_host_alloc": (C.c_void_p, [C.c_size_t]),
    "apop_b200_host_free": (None, [C.c_void_p]),
    "apop_b200_set_timing": (None, [C.c_int]),
    "apop_b200_set_local_only": (None, [C.c_int]),
    "apop_b200_global_rows": (C.c_double, [_PD]),
    "apop_b200_map": (C.c_int, [_PD, C.c_int, C.c_double, C.c_char, _dp, _dp]),
    "apop_b200_map_rows": (C.c_int, [_PD, C.c_int, _dp, C.c_size_t, _dp, _dp]),
    "apop_b200_synth": (_PD, [C.c_int, C.c_size_t, C.c_size_t, C.c_int, _dp, C.c_uint64, C.c_uint64]),
    "apop_b200_synth_free": (None, [_PD]),
    "apop_b200_download": (C.c_int, [_PD, _dp, _dp, _dp]),
    "apop_b200_probit_ll": (C.c_longdouble, [_PD, _PM]),
    "apop_b200_probit_score": (None, [_PD, _PV, _PM]),
    "apop_b200_logit_ll": (C.c_longdouble, [_PD, _PM]),
    "apop_b200_logit_score": (None, [_PD, _PV, _PM]),
    "apop_b200_poisson_ll": (C.c_longdouble, [_PD, _PM]),
    "apop_b200_poisson_score": (None, [_PD, _PV, _PM]),
    "apop_b200_normal_ll": (C.c_longdouble, [_PD, _PM]),
    "apop_b200_normal_score": (None, [_PD, _PV, _PM]),
    "apop_b200_ols_ll": (C.c_longdouble, [_PD, _PM]),
    "apop_b200_ols_score": (None, [_PD, _PV, _PM]),
    "apop_b200_exponential_ll": (C.c_longdouble, [_PD, _PM]),
    "apop_b200_exponential_score": (None, [_PD, _PV, _PM]),
    "apop_b200_gamma_ll": (C.c_longdouble, [_PD, _PM]),
    "apop_b200_gamma_score": (None, [_PD, _PV, _PM]),
    "apop_b200_lognormal_ll": (C.c_longdouble, [_PD, _PM]),
    "apop_b200_lognormal_score": (None, [_PD, _PV, _PM]),
    "apop_b200_bernoulli_ll": (C.c_longdouble, [_PD, _PM]),
    "apop_b200_bernoulli_score": (None, [_PD, _PV, _PM]),
    "apop_b200_beta_ll": (C.c_longdouble, [_PD, _PM]),
    "apop_b200_beta_score": (None, [_PD, _PV, _PM]),
    "apop_b200_zipf_ll": (C.c_longdouble, [_PD, _PM]),
    "apop_b200_dirichlet_ll": (C.c_longdouble, [_PD, _PM]),
    "apop_b200_dirichlet_score": (None, [_PD, _PV, _PM]),
    "apop_b200_t_ll": (C.c_longdouble, [_PD, _PM]),
    "apop_b200_yule_ll": (C.c_longdouble, [_PD, _PM]),
    "apop_b200_yule_score": (None, [_PD, _PV, _PM]),
    "apop_b200_poisson_reg_ll": (C.c_longdouble, [_PD, _PM]),
    "apop_b200_poisson_reg_score": (None, [_PD, _PV, _PM]),
    "apop_b200_text_to_resident": (_PD, [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int]),
    "apop_b200_text_to_host": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.POINTER(_dp), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "apop_b200_text_to_resident_fixed": (_PD, [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int]),
    "apop_b200_text_to_host_fixed": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(_dp), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "apop_b200_prep_outcome": (C.c_int, [_PD, _dp, C.c_size_t, C.POINTER(C.c_size_t), _dp, C.c_int]),
    "apop_b200_logit_expected": (C.c_int, [_PD, _PM, _dp, _dp]),
    "apop_b200_map_sum": (C.c_double, [_PD, C.c_int, C.c_double, C.c_char]),
    "apop_b200_ols_gram": (C.c_int, [_PD, _dp, _dp]),
    "apop_b200_dot": (C.c_int, [_PD, _dp, C.c_size_t, C.c_size_t, _dp]),
    "apop_b200_col_sums": (C.c_int, [_PD, C.c_int, _dp]),
    "apop_b200_probit_start": (C.c_int, [_PD, _dp]),
    "apop_b200_ll_score": (C.c_double, [C.c_int, _PD, _dp, C.c_size_t, _dp]),
    "apop_b200_information": (C.c_int, [C.c_int, _PD, _dp, C.c_size_t, _dp]),
    "apop_b200_ll_score_batched": (C.c_int, [C.c_int, _PD, _dp, C.c_size_t, C.c_size_t, C.c_int, _dp, _dp]),
    "apop_b200_bootstrap_counts": (C.c_int, [_PD, C.c_size_t, C.c_uint64]),
    "apop_b200_batched_exact_fraction": (C.c_double, [_PD]),
    "apop_b200_set_replicate": (C.c_int, [_PD, C.c_long]),
    "apop_b200_counts_download": (C.c_int, [_PD, C.c_size_t, C.POINTER(C.c_ubyte)]),
    "apop_b200_register": (C.c_int, [_PM, C.c_int]),
    "apop_b200_unregister": (C.c_int, [_PM]),
}

_LIB = None


class B200Error(RuntimeError):
    pass


def load_library(path=None):
    """dlopen libapop_b200.so and type every entry point.  Raises if it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise B200Error(f"{path} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(make -C apophenia_b200/csrc).  There is no CPU fallback.")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def last_error():
    return load_library().apop_b200_last_error().decode()


def check(rc, what):
    if rc != 0:
        raise B200Error(f"{what} failed: {last_error()}")


# ---------------------------------------------------------------------------
# building the reference's structs around numpy buffers
# ---------------------------------------------------------------------------
def make_vector(arr):
    """gsl_vector view of a 1-d float64 array (no copy); returns (struct, keepalive)."""
    a = np.asarray(arr, dtype=np.float64)
    if a.ndim != 1:
        raise ValueError("vector must be 1-d")
    if a.size > 1 and a.strides[0] % 8:
        a = np.ascontiguousarray(a)
    stride = a.strides[0] // 8 if a.size > 1 else 1
    v = gsl_vector(a.size, max(stride, 1), a.ctypes.data_as(_dp), None, 0)
    return v, a


def make_matrix(arr):
    """gsl_matrix view of a 2-d float64 array whose rows are contiguous (tda = row stride)."""
    a = np.asarray(arr, dtype=np.float64)
    if a.ndim != 2:
        raise ValueError("matrix must be 2-d")
    if a.shape[1] > 1 and a.strides[1] != 8 or (a.shape[0] > 1 and a.strides[0] % 8):
        a = np.ascontiguousarray(a)
    tda = a.strides[0] // 8 if a.shape[0] > 1 else max(a.shape[1], 1)
    m = gsl_matrix(a.shape[0], a.shape[1], tda, a.ctypes.data_as(_dp), None, 0)
    return m, a


class Data:
    """An `apop_data` (vector, matrix, weights, pages) over numpy arrays."""

    def __init__(self, matrix=None, vector=None, weights=None, title=None):
        self._keep = []
        self.struct = apop_data()
        if vector is not None:
            self._v, a = make_vector(vector)
            self._keep.append(a)
            self.vector = a
            self.struct.vector = C.pointer(self._v)
        else:
            self.vector = None
        if matrix is not None:
            self._m, a = make_matrix(matrix)
            self._keep.append(a)
            self.matrix = a
            self.struct.matrix = C.pointer(self._m)
        else:
            self.matrix = None
        if weights is not None:
            self._w, a = make_vector(weights)
            self._keep.append(a)
            self.weights = a
            self.struct.weights = C.pointer(self._w)
        else:
            self.weights = None
        if title is not None:
            self._title = C.create_string_buffer(title.encode())
            self._names = apop_name()
            self._names.title = C.cast(self._title, C.c_char_p)
            self.struct.names = C.pointer(self._names)
        self._pages = []

    @property
    def ptr(self):
        return C.pointer(self.struct)

    def __del__(self):
        # the resident copy is keyed by this struct's address: release it with the owner (an apop_data
        # allocated later at the same address must never meet a stale registry entry)
        try:
            if _LIB is not None:
                _LIB.apop_b200_unpin(C.pointer(self.struct))
        except Exception:
            pass

    def add_page(self, page):
        """append a page to the ->more list (apop_data_add_page)"""
        tail = self.struct
        while tail.more:
            tail = tail.more.contents
        tail.more = page.ptr
        self._pages.append(page)
        return page


class Model:
    """An `apop_model` header with parameters; log_likelihood can be installed by register()."""

    def __init__(self, name, parameters: Data = None, data: Data = None):
        self.struct = apop_model()
        self.struct.name = name.encode()[:100]
        self.parameters = parameters
        self.data = data
        if parameters is not None:
            self.struct.parameters = parameters.ptr
        if data is not None:
            self.struct.data = data.ptr

    @property
    def ptr(self):
        return C.pointer(self.struct)

    @property
    def error(self):
        return self.struct.error
