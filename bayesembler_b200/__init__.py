"""bayesembler_b200 -- B200-native per-locus Gibbs sampler of Bayesembler behind a C ABI.

This Python layer is only a ctypes binding of include/bayesembler_b200.h (used by tests/ and bench.py); the product
is the shared library bayesembler_b200/libbayesembler_b200.so (CUDA, sm_100a).  There is no CPU fallback: loading
fails loudly when the library is missing, and creating a context fails when no B200-class device is visible.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbayesembler_b200.so")

EXPORTED_SYMBOLS = [
    "bayes_gibbs_create", "bayes_gibbs_destroy", "bayes_gibbs_set_stream", "bayes_gibbs_get_stream",
    "bayes_gibbs_set_host_threads", "bayes_last_error", "bayes_gibbs_submit", "bayes_gibbs_upload", "bayes_gibbs_run",
    "bayes_gibbs_download", "bayes_gibbs_sync", "bayes_gibbs_run_batch", "bayes_gibbs_clear", "bayes_gibbs_fetch",
    "bayes_gibbs_fetch_all",
    "bayes_gibbs_locus_info", "bayes_gibbs_chain_info", "bayes_gibbs_num_loci", "bayes_gibbs_num_jobs", "bayes_gibbs_unit_stats", "bayes_gibbs_phase_seconds", "bayes_gibbs_last_launches",
    "bayes_gibbs_set_trace", "bayes_gibbs_fetch_trace", "bayes_top_marginals", "bayes_step_generate_assignment",
    "bayes_step_init_assignment", "bayes_step_generate_expression", "bayes_step_variates", "bayes_min_read_cover",
    "bayes_simplex_table", "bayes_chain_key", "bayes_locus_cost", "bayes_partition_lpt", "bayes_philox4x32",
]


class BayesError(RuntimeError):
    pass


class LocusDesc(C.Structure):
    """bayes_locus_desc"""
    _fields_ = [("T", C.c_int32), ("R", C.c_int32), ("row_ptr", C.c_void_p), ("col_idx", C.c_void_p), ("val", C.c_void_p),
                ("row_count", C.c_void_p), ("efflen", C.c_void_p), ("total_num_fragments", C.c_double), ("gamma", C.c_double),
                ("pi", C.c_double), ("burn_in", C.c_int32), ("samples", C.c_int32), ("seed", C.c_uint64),
                ("n_chains", C.c_int32), ("reserved", C.c_int32)]


class Locus(object):
    """One CollapsedMap (reference utils.h:129-136) in CSR form plus the sampler arguments of its call site."""

    def __init__(self, T, row_ptr, col_idx, val, row_count, efflen, n_total, gamma=1.0):
        self.T = int(T)
        self.row_ptr = np.ascontiguousarray(row_ptr, np.int32)
        self.col_idx = np.ascontiguousarray(col_idx, np.int32)
        self.val = np.ascontiguousarray(val, np.float64)
        self.row_count = np.ascontiguousarray(row_count, np.int32)
        self.efflen = np.ascontiguousarray(efflen, np.float64)
        self.n_total = float(n_total)
        self.gamma = float(gamma)
        self.R = len(self.row_count)

    @property
    def nnz(self):
        return int(self.row_ptr[-1])

    @property
    def n_frag(self):
        return int(self.row_count.sum())


_lib = None


def load_library(path=None):
    """Load the CUDA extension; raise (never fall back) when it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise BayesError("CUDA extension %s is missing; build it with `python -m bayesembler_b200._build` "
                         "(there is no CPU fallback)" % p)
    L = C.CDLL(p)
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
    sig = {
        "bayes_gibbs_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "bayes_gibbs_destroy": (None, [vp]),
        "bayes_gibbs_set_stream": (C.c_int, [vp, vp]),
        "bayes_gibbs_get_stream": (vp, [vp]),
        "bayes_gibbs_set_host_threads": (C.c_int, [vp, C.c_int]),
        "bayes_last_error": (C.c_char_p, []),
        "bayes_gibbs_submit": (C.c_int, [vp, i32, vp, C.POINTER(i64)]),
        "bayes_gibbs_upload": (C.c_int, [vp]),
        "bayes_gibbs_run": (C.c_int, [vp]),
        "bayes_gibbs_download": (C.c_int, [vp]),
        "bayes_gibbs_sync": (C.c_int, [vp]),
        "bayes_gibbs_run_batch": (C.c_int, [vp, i32, vp, C.POINTER(i64)]),
        "bayes_gibbs_clear": (C.c_int, [vp]),
        "bayes_gibbs_fetch": (C.c_int, [vp, i64, i32, vp, vp, vp, vp]),
        "bayes_gibbs_fetch_all": (C.c_int, [vp, vp, vp, vp, vp]),
        "bayes_gibbs_locus_info": (C.c_int, [vp, i64, C.POINTER(dbl), C.POINTER(i32), C.POINTER(i32)]),
        "bayes_gibbs_chain_info": (C.c_int, [vp, i64, i32, C.POINTER(dbl), C.POINTER(i32)]),
        "bayes_gibbs_num_loci": (i64, [vp]),
        "bayes_gibbs_num_jobs": (i64, [vp]),
        "bayes_gibbs_unit_stats": (i64, [vp, vp, vp]),
        "bayes_gibbs_phase_seconds": (C.c_int, [vp, vp]),
        "bayes_gibbs_last_launches": (i32, [vp]),
        "bayes_gibbs_set_trace": (C.c_int, [vp, i64, i32, i32]),
        "bayes_gibbs_fetch_trace": (C.c_int, [vp, vp, vp, vp, vp]),
        "bayes_top_marginals": (C.c_int, [i32, i32, vp, vp, vp, dbl, vp, vp, vp]),
        "bayes_step_generate_assignment": (C.c_int, [C.c_int, C.POINTER(LocusDesc), vp, u64, C.c_uint32, vp]),
        "bayes_step_init_assignment": (C.c_int, [C.c_int, C.POINTER(LocusDesc), u64, vp]),
        "bayes_step_generate_expression": (C.c_int, [C.c_int, i32, vp, i32, dbl, u64, C.c_uint32, vp]),
        "bayes_step_variates": (C.c_int, [C.c_int, C.c_int, dbl, dbl, u64, i32, vp]),
        "bayes_min_read_cover": (C.c_int, [C.POINTER(LocusDesc), u64, vp]),
        "bayes_simplex_table": (i64, [dbl, dbl, i32, i32, vp, vp]),
        "bayes_chain_key": (u64, [u64, i32]),
        "bayes_locus_cost": (dbl, [C.POINTER(LocusDesc)]),
        "bayes_partition_lpt": (C.c_int, [i64, vp, i32, vp]),
        "bayes_philox4x32": (None, [u64, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = L
    return L


def _check(rc, what):
    if rc < 0:
        raise BayesError("%s failed (%d): %s" % (what, rc, load_library().bayes_last_error().decode()))
    return rc


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_desc(loc, burn_in, samples, seed, pi=-1.0, n_chains=1):
    """bayes_locus_desc for a Locus-like object; keeps the arrays alive through the returned struct's _keep."""
    d = LocusDesc(loc.T, loc.R, loc.row_ptr.ctypes.data, loc.col_idx.ctypes.data, loc.val.ctypes.data,
                  loc.row_count.ctypes.data, loc.efflen.ctypes.data, loc.n_total, loc.gamma, pi, int(burn_in), int(samples),
                  int(seed), int(n_chains), 0)
    d._keep = loc
    return d


def make_desc_array(loci, burn_in, samples, seeds, pi=None, n_chains=1):
    n = len(loci)
    arr = (LocusDesc * n)()
    for i, loc in enumerate(loci):
        arr[i] = make_desc(loc, burn_in[i], samples[i], seeds[i], -1.0 if pi is None else pi[i],
                           n_chains if np.isscalar(n_chains) else n_chains[i])
    arr._keep = list(loci)
    return arr


class GibbsContext(object):
    """One GPU's sampler context (bayes_gibbs_ctx)."""

    def __init__(self, device=0, stream=None, host_threads=0):
        self.lib = load_library()
        h = C.c_void_p()
        _check(self.lib.bayes_gibbs_create(device, C.byref(h)), "bayes_gibbs_create")
        self.h = h
        self._T = []
        self._chains = []
        if stream is not None:
            _check(self.lib.bayes_gibbs_set_stream(self.h, C.c_void_p(stream)), "bayes_gibbs_set_stream")
        if host_threads:
            _check(self.lib.bayes_gibbs_set_host_threads(self.h, host_threads), "bayes_gibbs_set_host_threads")

    def close(self):
        if self.h:
            self.lib.bayes_gibbs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _note(self, desc_array):
        if isinstance(desc_array, np.ndarray):  # structured array laid out as bayes_locus_desc[] (synth.DESC_DTYPE)
            self._T.extend(int(t) for t in desc_array["T"])
            self._chains.extend(int(c) for c in desc_array["n_chains"])
            return C.c_void_p(desc_array.ctypes.data)
        for d in desc_array:
            self._T.append(d.T)
            self._chains.append(d.n_chains)
        return C.cast(desc_array, C.c_void_p)

    def submit(self, desc_array):
        first = C.c_int64(0)
        ptr = self._note(desc_array)
        _check(self.lib.bayes_gibbs_submit(self.h, len(desc_array), ptr, C.byref(first)), "bayes_gibbs_submit")
        return first.value

    def upload(self):
        _check(self.lib.bayes_gibbs_upload(self.h), "bayes_gibbs_upload")

    def run(self):
        _check(self.lib.bayes_gibbs_run(self.h), "bayes_gibbs_run")

    def download(self):
        _check(self.lib.bayes_gibbs_download(self.h), "bayes_gibbs_download")

    def sync(self):
        _check(self.lib.bayes_gibbs_sync(self.h), "bayes_gibbs_sync")

    def run_batch(self, desc_array):
        first = C.c_int64(0)
        ptr = self._note(desc_array)
        _check(self.lib.bayes_gibbs_run_batch(self.h, len(desc_array), ptr, C.byref(first)), "bayes_gibbs_run_batch")
        return first.value

    def fetch_all(self):
        """Accumulators of every (locus, chain 0..) as flat arrays in submission order."""
        n = sum(t * c for t, c in zip(self._T, self._chains))
        occ = np.zeros(n, np.int32)
        mean = np.zeros(n)
        m2 = np.zeros(n)
        counts = np.zeros(n, np.int32)
        _check(self.lib.bayes_gibbs_fetch_all(self.h, _ptr(occ), _ptr(mean), _ptr(m2), _ptr(counts)), "bayes_gibbs_fetch_all")
        return dict(occ=occ, mean=mean, m2=m2, final_counts=counts)

    def clear(self):
        _check(self.lib.bayes_gibbs_clear(self.h), "bayes_gibbs_clear")
        self._T = []
        self._chains = []

    def fetch(self, locus_id, chain=0):
        T = self._T[locus_id]
        occ = np.zeros(T, np.int32)
        mean = np.zeros(T)
        m2 = np.zeros(T)
        counts = np.zeros(T, np.int32)
        _check(self.lib.bayes_gibbs_fetch(self.h, locus_id, chain, _ptr(occ), _ptr(mean), _ptr(m2), _ptr(counts)),
               "bayes_gibbs_fetch")
        return dict(occ=occ, mean=mean, m2=m2, final_counts=counts)

    def locus_info(self, locus_id):
        pi = C.c_double(0)
        cover = C.c_int32(0)
        folded = C.c_int32(0)
        _check(self.lib.bayes_gibbs_locus_info(self.h, locus_id, C.byref(pi), C.byref(cover), C.byref(folded)),
               "bayes_gibbs_locus_info")
        return dict(pi=pi.value, cover=cover.value, folded=folded.value)

    def chain_info(self, locus_id, chain):
        pi = C.c_double(0)
        cover = C.c_int32(0)
        _check(self.lib.bayes_gibbs_chain_info(self.h, locus_id, chain, C.byref(pi), C.byref(cover)), "bayes_gibbs_chain_info")
        return dict(pi=pi.value, cover=cover.value)

    def num_jobs(self):
        return self.lib.bayes_gibbs_num_jobs(self.h)

    def unit_stats(self):
        """Per-unit shapes and (start ns, end ns, SM) of the last run."""
        n = self.lib.bayes_gibbs_unit_stats(self.h, None, None)
        shape = np.zeros((n, 8), np.int32)
        clock = np.zeros((n, 3), np.uint64)
        _check(self.lib.bayes_gibbs_unit_stats(self.h, _ptr(shape), _ptr(clock)), "bayes_gibbs_unit_stats")
        return shape, clock

    def phase_seconds(self):
        out = np.zeros(6)
        _check(self.lib.bayes_gibbs_phase_seconds(self.h, _ptr(out)), "bayes_gibbs_phase_seconds")
        return dict(zip(["prepare", "pack", "h2d_enqueue", "launch", "d2h_enqueue", "sync"], out.tolist()))

    def last_launches(self):
        return self.lib.bayes_gibbs_last_launches(self.h)

    def set_trace(self, locus_id, chain, trace_iters):
        _check(self.lib.bayes_gibbs_set_trace(self.h, locus_id, chain, trace_iters), "bayes_gibbs_set_trace")
        self._trace = (self._T[locus_id], trace_iters)

    def fetch_trace(self):
        T, n = self._trace
        theta = np.zeros((n, T))
        counts = np.zeros((n, T), np.int32)
        b = np.zeros(n, np.int32)
        init = np.zeros(T, np.int32)
        _check(self.lib.bayes_gibbs_fetch_trace(self.h, _ptr(theta), _ptr(counts), _ptr(b), _ptr(init)),
               "bayes_gibbs_fetch_trace")
        return dict(theta_trace=theta, counts_trace=counts, b_trace=b, init_counts=init)


# ---- stateless entry points -------------------------------------------------------------------------------------------
def chain_key(seed, chain=0):
    return int(load_library().bayes_chain_key(int(seed), int(chain)))


def min_read_cover(loc, key):
    d = make_desc(loc, 0, 1, 0)
    flags = np.zeros(loc.T, np.int32)
    n = _check(load_library().bayes_min_read_cover(C.byref(d), int(key), _ptr(flags)), "bayes_min_read_cover")
    return n, flags


def simplex_table(pi, gamma, T, n_frag):
    L = load_library()
    n = _check(L.bayes_simplex_table(pi, gamma, T, n_frag, None, None), "bayes_simplex_table")
    row_off = np.zeros(T + 1, np.int64)
    table = np.zeros(n)
    _check(L.bayes_simplex_table(pi, gamma, T, n_frag, _ptr(row_off), _ptr(table)), "bayes_simplex_table")
    return row_off, table


def top_marginals(samples, occ, mean, m2, threshold):
    T = len(occ)
    occ = np.ascontiguousarray(occ, np.int32)
    mean = np.ascontiguousarray(mean, np.float64)
    m2 = np.ascontiguousarray(m2, np.float64)
    idx = np.zeros(T, np.int32)
    mo = np.zeros(T)
    sd = np.zeros(T)
    n = load_library().bayes_top_marginals(T, samples, _ptr(occ), _ptr(mean), _ptr(m2), threshold, _ptr(idx), _ptr(mo), _ptr(sd))
    return idx[:n], mo[:n], sd[:n]


def partition_lpt(cost, n_parts):
    cost = np.ascontiguousarray(cost, np.float64)
    part = np.zeros(len(cost), np.int32)
    _check(load_library().bayes_partition_lpt(len(cost), _ptr(cost), n_parts, _ptr(part)), "bayes_partition_lpt")
    return part


def locus_cost(loc, burn_in, samples, n_chains=1):
    d = make_desc(loc, burn_in, samples, 0, n_chains=n_chains)
    return load_library().bayes_locus_cost(C.byref(d))


def philox4x32(key, ctr):
    ctr = np.ascontiguousarray(ctr, np.uint32)
    out = np.zeros(4, np.uint32)
    load_library().bayes_philox4x32(int(key), _ptr(ctr), _ptr(out))
    return out


def step_generate_assignment(loc, theta, key, iteration=0, device=0):
    d = make_desc(loc, 0, 1, 0, pi=0.5)
    theta = np.ascontiguousarray(theta, np.float64)
    counts = np.zeros(loc.T, np.int32)
    _check(load_library().bayes_step_generate_assignment(device, C.byref(d), _ptr(theta), int(key), iteration, _ptr(counts)),
           "bayes_step_generate_assignment")
    return counts


def step_init_assignment(loc, key, device=0):
    d = make_desc(loc, 0, 1, 0, pi=0.5)
    counts = np.zeros(loc.T, np.int32)
    _check(load_library().bayes_step_init_assignment(device, C.byref(d), int(key), _ptr(counts)), "bayes_step_init_assignment")
    return counts


def step_generate_expression(counts, b, gamma, key, iteration=0, device=0):
    counts = np.ascontiguousarray(counts, np.int32)
    theta = np.zeros(len(counts))
    _check(load_library().bayes_step_generate_expression(device, len(counts), _ptr(counts), b, gamma, int(key), iteration,
                                                         _ptr(theta)), "bayes_step_generate_expression")
    return theta


def step_variates(kind, a, p, key, n, device=0):
    out = np.zeros(n)
    k = {"gamma": 0, "binomial": 1, "uniform_int": 2}[kind]
    _check(load_library().bayes_step_variates(device, k, float(a), float(p), int(key), n, _ptr(out)), "bayes_step_variates")
    return out
