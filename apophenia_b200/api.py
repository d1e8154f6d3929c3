"""Host-side mirror of the reference's operators for this path.

Names follow the reference: `log_likelihood(d, m)` is apop_log_likelihood
(apop_model.m4.c:257), `score(d, m)` is apop_score (:290), `pin` / `unpin` /
`invalidate` manage HBM residency.  Everything computes on the B200 through the
C-ABI in include/apop_b200.h.
"""
import ctypes as C

import numpy as np

from . import abi
from .abi import (B200Error, Data, Model, PROBIT, LOGIT, POISSON, NORMAL, OLS, POISSON_REG,  # noqa: F401
                  EXPONENTIAL, GAMMA, LOGNORMAL, BERNOULLI, BETA, ZIPF, DIRICHLET, TDIST, YULE,
                  MAP_IDENTITY, MAP_DEV, MAP_DEV2, MAP_POISSON, MAP_LOG, MAP_EXP, MAP_ABS, MAP_SQRT, MAP_SCALE, MAP_POW,
                  ROW_SUM, ROW_MEAN, ROW_SUMSQ, ROW_MAX, ROW_MIN, ROW_DOT, ROW_PROBIT_LL, ROW_LOGIT2_LL,
                  ROW_POISSON_REG_LL, ROW_OLS_ERROR)

__all__ = ["init", "finalize", "pin", "unpin", "invalidate", "is_pinned", "last_pin_mode", "host_empty", "log_likelihood", "score",
           "ll_score", "map_sum", "synth", "SynthData", "launch_count", "last_pass_ms", "pass_timer_reset", "set_timing", "set_local_only",
           "pass_timer", "probit_model", "logit_model", "vector_model", "Data", "Model", "B200Error",
           "PROBIT", "LOGIT", "POISSON", "NORMAL", "OLS", "POISSON_REG", "EXPONENTIAL", "GAMMA", "LOGNORMAL",
           "BERNOULLI", "BETA", "ZIPF", "DIRICHLET", "TDIST", "YULE", "comm_init_from_torch",
           "comm_finalize", "set_auto_pin", "last_error", "information", "ols_gram",
           "dot", "col_sums", "probit_start", "prep_outcome", "logit_expected", "text_to_host", "text_to_resident", "ll_score_batched", "batched_exact_fraction", "set_replicate", "bootstrap_counts", "counts_download",
           "MAP_IDENTITY", "MAP_DEV", "MAP_DEV2", "MAP_POISSON", "MAP_LOG", "MAP_EXP", "MAP_ABS", "MAP_SQRT",
           "MAP_SCALE", "MAP_POW", "ROW_SUM", "ROW_MEAN", "ROW_SUMSQ", "ROW_MAX", "ROW_MIN", "ROW_DOT",
           "ROW_PROBIT_LL", "ROW_LOGIT2_LL", "ROW_POISSON_REG_LL", "ROW_OLS_ERROR", "map", "map_rows", "global_rows"]

_dp = C.POINTER(C.c_double)


def _lib():
    return abi.load_library()


def init(device=-1):
    """Bind this process to one B200 (LOCAL_RANK if device < 0)."""
    abi.check(_lib().apop_b200_init(int(device)), "apop_b200_init")


def finalize():
    _lib().apop_b200_finalize()


def last_error():
    return abi.last_error()


def set_auto_pin(on):
    _lib().apop_b200_set_auto_pin(int(bool(on)))


def pin(d: Data):
    abi.check(_lib().apop_b200_pin(d.ptr), "apop_b200_pin")
    return d


def unpin(d):
    return _lib().apop_b200_unpin(d.ptr)


def invalidate(d):
    return _lib().apop_b200_invalidate(d.ptr)


def last_pin_mode():
    """('staged' | 'registered', page-locking GB/s measured on the first window) of the most recent upload"""
    r = C.c_double(0.0)
    m = _lib().apop_b200_last_pin_mode(C.byref(r))
    return {0: "staged", 1: "registered", 2: "zero-copy (page-locked source)"}.get(m, str(m)), float(r.value)


def host_empty(shape, hugepages=True, pinned=False):
    """A page-aligned float64 host array.  pinned: page-locked memory from apop_b200_host_alloc (cudaHostAlloc) --
    apop_b200_pin then reads it in place (zero copy).  Otherwise an anonymous mapping, with hugepages advised to
    transparent huge pages (what GLIBC_TUNABLES=glibc.malloc.hugetlb=1 does for a C caller's gsl_matrix_alloc)."""
    import mmap
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape)
    if pinned:
        import weakref
        lib = _lib()
        p = lib.apop_b200_host_alloc(n * 8)
        if not p:
            raise B200Error("apop_b200_host_alloc failed: " + abi.last_error())
        raw = (C.c_double * n).from_address(p)
        a = np.frombuffer(raw, dtype=np.float64, count=n).reshape(shape)
        weakref.finalize(raw, lib.apop_b200_host_free, p)
        return a
    buf = mmap.mmap(-1, n * 8, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    a = np.frombuffer(buf, dtype=np.float64, count=n).reshape(shape)
    if hugepages:
        try:
            C.CDLL(None).madvise(C.c_void_p(a.ctypes.data), C.c_size_t(n * 8), 14)   # MADV_HUGEPAGE
        except Exception:
            pass
    return a


def is_pinned(d):
    return bool(_lib().apop_b200_is_pinned(d.ptr))


def launch_count():
    return int(_lib().apop_b200_launch_count())


def last_pass_ms():
    return float(_lib().apop_b200_last_pass_ms())


def set_timing(on=True):
    """per-pass CUDA-event timing (off by default: it costs ~25 us per pass)"""
    _lib().apop_b200_set_timing(int(bool(on)))


def set_local_only(on=True):
    """evaluate this rank's rows only (no cross-GPU exchange) while on"""
    _lib().apop_b200_set_local_only(int(bool(on)))


def pass_timer_reset():
    """zero the accumulated kernel time; asking for the timer switches per-pass timing on"""
    _lib().apop_b200_set_timing(1)
    _lib().apop_b200_pass_timer_reset()


def pass_timer():
    n = C.c_ulonglong(0)
    ms = _lib().apop_b200_pass_timer_total_ms(C.byref(n))
    return float(ms), int(n.value)


_LL = {PROBIT: "apop_b200_probit_ll", LOGIT: "apop_b200_logit_ll", POISSON: "apop_b200_poisson_ll",
       NORMAL: "apop_b200_normal_ll", OLS: "apop_b200_ols_ll", POISSON_REG: "apop_b200_poisson_reg_ll",
       EXPONENTIAL: "apop_b200_exponential_ll", GAMMA: "apop_b200_gamma_ll", LOGNORMAL: "apop_b200_lognormal_ll",
       BERNOULLI: "apop_b200_bernoulli_ll", BETA: "apop_b200_beta_ll", ZIPF: "apop_b200_zipf_ll",
       DIRICHLET: "apop_b200_dirichlet_ll", TDIST: "apop_b200_t_ll", YULE: "apop_b200_yule_ll"}
_SC = {PROBIT: "apop_b200_probit_score", LOGIT: "apop_b200_logit_score", POISSON: "apop_b200_poisson_score",
       NORMAL: "apop_b200_normal_score", OLS: "apop_b200_ols_score",
       POISSON_REG: "apop_b200_poisson_reg_score", EXPONENTIAL: "apop_b200_exponential_score",
       GAMMA: "apop_b200_gamma_score", LOGNORMAL: "apop_b200_lognormal_score",
       BERNOULLI: "apop_b200_bernoulli_score", BETA: "apop_b200_beta_score", DIRICHLET: "apop_b200_dirichlet_score",
       YULE: "apop_b200_yule_score"}


def _ptr(obj):
    return obj.ptr if obj is not None else None


def log_likelihood(family, d, m):
    """model->log_likelihood(d, m) on the device; NaN on error (reference convention)."""
    return float(getattr(_lib(), _LL[family])(_ptr(d), _ptr(m)))


def score(family, d, m, size=None):
    """the apop_score vtable entry: fills a caller-allocated gsl_vector in pack order"""
    if size is None:
        p = m.parameters
        size = (p.vector.size if p.vector is not None else 0) + (p.matrix.size if p.matrix is not None else 0)
    out = np.zeros(size)
    g, keep = abi.make_vector(out)
    getattr(_lib(), _SC[family])(_ptr(d), C.byref(g), _ptr(m))
    return out


def ll_score(family, d, beta, want_score=True):
    beta = np.ascontiguousarray(beta, dtype=np.float64).ravel()
    g = np.zeros(beta.size) if want_score else None
    ll = _lib().apop_b200_ll_score(int(family), _ptr(d), beta.ctypes.data_as(_dp), beta.size,
                                   g.ctypes.data_as(_dp) if want_score else None)
    return float(ll), g


def information(family, d, beta):
    beta = np.ascontiguousarray(beta, dtype=np.float64).ravel()
    H = np.zeros((beta.size, beta.size))
    abi.check(_lib().apop_b200_information(int(family), _ptr(d), beta.ctypes.data_as(_dp), beta.size,
                                           H.ctypes.data_as(_dp)), "apop_b200_information")
    return H


def ll_score_batched(family, d, B, use_counts=False, want_score=True):
    """m parameter vectors in one pass (bootstrap replicates, parallel chains).
    B is m x P (one parameter vector per row).  Returns (ll[m], score[m x P] or None)."""
    B = np.ascontiguousarray(B, dtype=np.float64)
    m, P = B.shape
    ll = np.zeros(m)
    g = np.zeros((m, P)) if want_score else None
    abi.check(_lib().apop_b200_ll_score_batched(int(family), _ptr(d), B.ctypes.data_as(_dp), P, m, int(use_counts),
                                                ll.ctypes.data_as(_dp),
                                                g.ctypes.data_as(_dp) if want_score else None),
              "apop_b200_ll_score_batched")
    return ll, g


def batched_exact_fraction(d):
    """share of the last shared-centre batched pass over d that needed the exact link evaluation"""
    return float(_lib().apop_b200_batched_exact_fraction(_ptr(d)))


def set_replicate(d, replicate):
    """weight the single-beta multinomial logit passes over d by one bootstrap replicate's counts (-1: off)"""
    abi.check(_lib().apop_b200_set_replicate(_ptr(d), int(replicate)), "apop_b200_set_replicate")


def bootstrap_counts(d, m, seed=8):
    """draw m multinomial resamples of the rows into a resident count matrix"""
    abi.check(_lib().apop_b200_bootstrap_counts(_ptr(d), int(m), int(seed)), "apop_b200_bootstrap_counts")


def counts_download(d, m, n_rows):
    out = np.zeros((n_rows, m), dtype=np.uint8)
    abi.check(_lib().apop_b200_counts_download(_ptr(d), int(m), out.ctypes.data_as(C.POINTER(C.c_ubyte))),
              "apop_b200_counts_download")
    return out


def dot(d, B, n_rows):
    """apop_dot(d, parameters): X.B as an (n_rows, c) array"""
    B = np.ascontiguousarray(B, dtype=np.float64)
    if B.ndim == 1:
        B = B.reshape(-1, 1)
    out = np.zeros((n_rows, B.shape[1]))
    abi.check(_lib().apop_b200_dot(_ptr(d), B.ctypes.data_as(_dp), B.shape[0], B.shape[1], out.ctypes.data_as(_dp)),
              "apop_b200_dot")
    return out


def col_sums(d, mode, k):
    out = np.zeros(k)
    abi.check(_lib().apop_b200_col_sums(_ptr(d), int(mode), out.ctypes.data_as(_dp)), "apop_b200_col_sums")
    return out


def probit_start(d, k):
    """probit_prep's start point exp(mean ln|x_ij|) computed on the device"""
    out = np.zeros(k)
    abi.check(_lib().apop_b200_probit_start(_ptr(d), out.ctypes.data_as(_dp)), "apop_b200_probit_start")
    return out


def text_to_host(path, has_row_names=False, has_col_names=True, delimiters=None, field_ends=None):
    """apop_text_to_data's matrix for a numeric file, parsed by the library's multi-threaded reader;
    field_ends (1-based end positions) selects the fixed-width layout"""
    out = _dp(); n = C.c_size_t(0); k = C.c_size_t(0)
    if field_ends is not None:
        fe = (C.c_int * len(field_ends))(*[int(v) for v in field_ends])
        abi.check(_lib().apop_b200_text_to_host_fixed(str(path).encode(), int(has_row_names), int(has_col_names), fe, len(field_ends),
                                                      C.byref(out), C.byref(n), C.byref(k)), "apop_b200_text_to_host_fixed")
    else:
        abi.check(_lib().apop_b200_text_to_host(str(path).encode(), int(has_row_names), int(has_col_names),
                                                delimiters.encode() if delimiters else None, C.byref(out), C.byref(n), C.byref(k)),
                  "apop_b200_text_to_host")
    a = np.ctypeslib.as_array(out, shape=(n.value, k.value)).copy() if n.value else np.zeros((0, k.value))
    C.CDLL(None).free(out)
    return a


def prep_outcome(d, max_cats=64, want_start=True, sync_host=False):
    """device prep (ols_shuffle + category table + recode + start point); returns (cats, start)"""
    cats = np.zeros(max_cats)
    ncat = C.c_size_t(0)
    k = d.matrix.shape[1] if getattr(d, "matrix", None) is not None else d.k
    start = np.zeros(k) if want_start else None
    abi.check(_lib().apop_b200_prep_outcome(_ptr(d), cats.ctypes.data_as(_dp), max_cats, C.byref(ncat),
                                            start.ctypes.data_as(_dp) if want_start else None, int(sync_host)),
              "apop_b200_prep_outcome")
    return cats[:ncat.value].copy(), start


def logit_expected(d, m, n_rows, ncat):
    """multilogit_expected: (probabilities n x C, most likely category n)"""
    probs = np.zeros((n_rows, ncat)); best = np.zeros(n_rows)
    abi.check(_lib().apop_b200_logit_expected(_ptr(d), _ptr(m), probs.ctypes.data_as(_dp), best.ctypes.data_as(_dp)),
              "apop_b200_logit_expected")
    return probs, best


def ols_gram(d):
    k = d.matrix.shape[1] if getattr(d, "matrix", None) is not None else d.k
    A = np.zeros((k, k)); b = np.zeros(k)
    abi.check(_lib().apop_b200_ols_gram(_ptr(d), A.ctypes.data_as(_dp), b.ctypes.data_as(_dp)), "apop_b200_ols_gram")
    return A, b


def map_sum(d, fn, p0=0.0, part="a"):
    return float(_lib().apop_b200_map_sum(_ptr(d), int(fn), float(p0), part.encode()))


def map(d, fn, p0=0.0, part="a", n_rows=None, k=None):  # noqa: A001 (the reference's name)
    """apop_map with a materialised result: (vector or None, matrix or None) of the input's shape"""
    n = n_rows if n_rows is not None else (d.matrix.shape[0] if getattr(d, "matrix", None) is not None else d.vector.shape[0])
    kk = k if k is not None else (d.matrix.shape[1] if getattr(d, "matrix", None) is not None else 0)
    has_v = getattr(d, "vector", None) is not None or getattr(d, "has_y", False)
    ov = np.zeros(n) if has_v else None
    om = np.zeros((n, kk)) if kk else None
    abi.check(_lib().apop_b200_map(_ptr(d), int(fn), float(p0), part.encode(),
                                   ov.ctypes.data_as(_dp) if ov is not None else None,
                                   om.ctypes.data_as(_dp) if om is not None else None), "apop_b200_map")
    return ov, om


def map_rows(d, fn, n_rows, param=None, aux=None):
    """apop_map(.part='r') with one of the built-in row functions: one double per row"""
    out = np.zeros(n_rows)
    p = np.ascontiguousarray(param, dtype=np.float64).ravel() if param is not None else None
    a = np.ascontiguousarray(aux, dtype=np.float64).ravel() if aux is not None else None
    abi.check(_lib().apop_b200_map_rows(_ptr(d), int(fn), p.ctypes.data_as(_dp) if p is not None else None,
                                        p.size if p is not None else 0, a.ctypes.data_as(_dp) if a is not None else None,
                                        out.ctypes.data_as(_dp)), "apop_b200_map_rows")
    return out


def global_rows(d):
    return float(_lib().apop_b200_global_rows(_ptr(d)))


# ---- parameter containers shaped like the reference's --------------------------
def probit_model(beta, data=None, name="Probit"):
    """parameters = k x 1 matrix, as probit_prep allocates them (model/apop_probit.c:49)"""
    b = np.ascontiguousarray(beta, dtype=np.float64).reshape(-1, 1)
    return Model(name, Data(matrix=b), data)


def logit_model(B, data=None, name="Logit"):
    """parameters = k x (C-1) matrix"""
    B = np.ascontiguousarray(B, dtype=np.float64)
    if B.ndim == 1:
        B = B.reshape(-1, 1)
    return Model(name, Data(matrix=B), data)


def vector_model(params, data=None, name="model"):
    """parameters in the vector: Poisson (lambda), Normal (mu, sigma), OLS (beta)"""
    return Model(name, Data(vector=np.ascontiguousarray(params, dtype=np.float64)), data)


# ---- device-generated synthetic data -----------------------------------------------
class SynthData:
    """A resident, device-generated data set; `.ptr` is the apop_data* header."""

    def __init__(self, family, n_rows, k, beta_true, ncat=2, seed=8, row_offset=0):
        beta_true = np.ascontiguousarray(beta_true, dtype=np.float64).ravel()
        self.family, self.n_rows, self.k, self.ncat = family, int(n_rows), int(k), int(ncat)
        self.has_y = family not in (POISSON, NORMAL)
        self._p = _lib().apop_b200_synth(int(family), self.n_rows, self.k, self.ncat,
                                         beta_true.ctypes.data_as(_dp), int(seed), int(row_offset))
        if not self._p:
            raise B200Error("apop_b200_synth failed: " + abi.last_error())

    @property
    def ptr(self):
        return self._p

    def download(self, hugepages=False):
        """(X row-major, y) back on the host, in the reference's layout"""
        X = host_empty((self.n_rows, self.k)) if hugepages else np.zeros((self.n_rows, self.k))
        y = np.zeros(self.n_rows) if self.has_y else None
        abi.check(_lib().apop_b200_download(self._p, X.ctypes.data_as(_dp),
                                            y.ctypes.data_as(_dp) if y is not None else None, None),
                  "apop_b200_download")
        return X, y

    def free(self):
        if self._p:
            _lib().apop_b200_synth_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def synth(family, n_rows, k, beta_true, ncat=2, seed=8, row_offset=0):
    return SynthData(family, n_rows, k, beta_true, ncat, seed, row_offset)


class TextData(SynthData):
    """A numeric text file ingested straight into the resident layout (apop_b200_text_to_resident)."""

    def __init__(self, path, has_row_names=False, has_col_names=True, delimiters=None, outcome_col0=False, field_ends=None):
        if field_ends is not None:
            fe = (C.c_int * len(field_ends))(*[int(v) for v in field_ends])
            self._p = _lib().apop_b200_text_to_resident_fixed(str(path).encode(), int(has_row_names), int(has_col_names), fe,
                                                              len(field_ends), int(outcome_col0))
        else:
            self._p = _lib().apop_b200_text_to_resident(str(path).encode(), int(has_row_names), int(has_col_names),
                                                        delimiters.encode() if delimiters else None, int(outcome_col0))
        if not self._p:
            raise B200Error("apop_b200_text_to_resident failed: " + abi.last_error())
        m = self._p.contents.matrix.contents
        self.n_rows, self.k, self.has_y = int(m.size1), int(m.size2), bool(outcome_col0)
        self.family, self.ncat = None, 0


def text_to_resident(path, has_row_names=False, has_col_names=True, delimiters=None, outcome_col0=False, field_ends=None):
    return TextData(path, has_row_names, has_col_names, delimiters, outcome_col0, field_ends)


# ---- multi-GPU plumbing: torch.distributed carries the NCCL id ----------------------
def comm_init_from_torch():
    """Create the library's NCCL communicator over the ranks of the default
    torch.distributed process group (one process per GPU)."""
    import torch
    import torch.distributed as dist
    lib = _lib()
    rank, world = dist.get_rank(), dist.get_world_size()
    buf = (C.c_char * 128)()
    if rank == 0:
        abi.check(lib.apop_b200_comm_unique_id(C.cast(buf, C.c_void_p)), "apop_b200_comm_unique_id")
    t = torch.tensor(list(bytes(buf)), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().tolist())
    buf2 = (C.c_char * 128).from_buffer_copy(raw)
    abi.check(lib.apop_b200_comm_init(world, rank, C.cast(buf2, C.c_void_p)), "apop_b200_comm_init")
    # fused allreduce over NVLink peer memory (default; APOP_B200_ALLREDUCE=nccl keeps NCCL only)
    import os
    if os.environ.get("APOP_B200_ALLREDUCE", "p2p") != "nccl" and 2 <= world <= 8:
        hbuf = (C.c_char * 64)()
        abi.check(lib.apop_b200_comm_p2p_handle(world, C.cast(hbuf, C.c_void_p)), "apop_b200_comm_p2p_handle")
        mine = torch.tensor(list(bytes(hbuf)), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            mine = mine.cuda()
        allh = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        raw = b"".join(bytes(t.cpu().tolist()) for t in allh)
        hall = (C.c_char * (64 * world)).from_buffer_copy(raw)
        abi.check(lib.apop_b200_comm_p2p_init(world, rank, C.cast(hall, C.c_void_p)), "apop_b200_comm_p2p_init")
        dist.barrier()


def comm_finalize():
    _lib().apop_b200_comm_finalize()
