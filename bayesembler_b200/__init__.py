"""bayesembler_b200 -- B200-native per-locus Gibbs sampler of Bayesembler behind a C ABI.

This Python layer is only a ctypes binding of include/bayesembler_b200.h (used by tests/ and bench.py); the product
is the shared library bayesembler_b200/libbayesembler_b200.so (CUDA, sm_100a).  There is no CPU fallback: loading
fails loudly when the library is missing, and creating a context fails when no B200-class device is visible.
"""
import ctypes as C
import os
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbayesembler_b200.so")

EXPORTED_SYMBOLS = [
    "bayes_gibbs_create", "bayes_gibbs_destroy", "bayes_gibbs_set_stream", "bayes_gibbs_get_stream",
    "bayes_gibbs_set_host_threads", "bayes_last_error", "bayes_gibbs_submit", "bayes_gibbs_upload", "bayes_gibbs_run",
    "bayes_gibbs_download", "bayes_gibbs_sync", "bayes_gibbs_run_batch", "bayes_gibbs_clear", "bayes_gibbs_fetch",
    "bayes_gibbs_fetch_all",
    "bayes_gibbs_locus_info", "bayes_gibbs_chain_info", "bayes_gibbs_num_loci", "bayes_gibbs_num_jobs", "bayes_gibbs_unit_stats", "bayes_gibbs_phase_seconds", "bayes_gibbs_last_launches",
    "bayes_gibbs_set_trace", "bayes_gibbs_fetch_trace", "bayes_gibbs_set_calibration", "bayes_gibbs_set_mode", "bayes_gibbs_get_mode",
    "bayes_top_marginals", "bayes_step_generate_assignment", "bayes_step_assignment",
    "bayes_step_init_assignment", "bayes_step_generate_expression", "bayes_step_variates", "bayes_min_read_cover",
    "bayes_simplex_table", "bayes_chain_key", "bayes_locus_cost", "bayes_partition_lpt", "bayes_philox4x32",
    "bayes_seq_model_create_gaussian", "bayes_seq_model_destroy", "bayes_seq_model_three_prime_prob", "bayes_seq_model_length_prob",
    "bayes_seq_model_effective_length", "bayes_fragment_model_estimate", "bayes_collapsed_map_build", "bayes_collapsed_map_destroy",
    "bayes_collapsed_map_num_candidates", "bayes_collapsed_map_num_rows", "bayes_collapsed_map_nnz", "bayes_collapsed_map_export",
    "bayes_collapsed_map_stats", "bayes_fetch_fragment_lengths",
    "bayes_all_paths_build", "bayes_path_set_destroy", "bayes_path_set_num_candidates", "bayes_path_set_num_exons",
    "bayes_path_set_export", "bayes_path_set_stats", "bayes_path_set_graph_name", "bayes_path_set_reference",
    "bayes_path_set_strand", "bayes_path_set_candidate_name",
    "bayes_bam_front_open", "bayes_bam_front_close", "bayes_bam_front_counts", "bayes_bam_front_num_alignments", "bayes_bam_front_write_sam",
    "bayes_bam_front_graphs", "bayes_sequencing_probability",
]


class BayesError(RuntimeError):
    pass


# sampling modes (include/bayesembler_b200.h)
MODE_FAST, MODE_REPLAY = 0, 1
_MODES = {"fast": MODE_FAST, "replay": MODE_REPLAY, MODE_FAST: MODE_FAST, MODE_REPLAY: MODE_REPLAY}


class LocusDesc(C.Structure):
    """bayes_locus_desc"""
    _fields_ = [("T", C.c_int32), ("R", C.c_int32), ("row_ptr", C.c_void_p), ("col_idx", C.c_void_p), ("val", C.c_void_p),
                ("row_count", C.c_void_p), ("efflen", C.c_void_p), ("total_num_fragments", C.c_double), ("gamma", C.c_double),
                ("pi", C.c_double), ("burn_in", C.c_int32), ("samples", C.c_int32), ("seed", C.c_uint64),
                ("n_chains", C.c_int32), ("reserved", C.c_int32)]


class RawLocusDesc(C.Structure):
    """bayes_raw_locus"""
    _fields_ = [("n_cand", C.c_int32), ("exon_ptr", C.c_void_p), ("exon_start", C.c_void_p), ("exon_end", C.c_void_p),
                ("cand_length", C.c_void_p), ("is_premrna", C.c_void_p), ("subgraph_id", C.c_void_p), ("n_subgraphs", C.c_int32),
                ("strand_plus", C.c_int32), ("n_frag", C.c_int32), ("read_pos", C.c_void_p), ("cigar_ptr", C.c_void_p),
                ("cigar_op", C.c_void_p), ("cigar_len", C.c_void_p), ("frag_prob", C.c_void_p)]


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
    p = path or os.environ.get("BAYES_B200_LIB") or LIB_PATH  # BAYES_B200_LIB: an experiment build of the same sources (debug bounds, phase clocks)
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
        "bayes_gibbs_unit_stats": (i64, [vp, vp, vp, vp]),
        "bayes_gibbs_phase_seconds": (C.c_int, [vp, vp]),
        "bayes_gibbs_last_launches": (i32, [vp]),
        "bayes_gibbs_set_trace": (C.c_int, [vp, i64, i32, i32]),
        "bayes_gibbs_set_calibration": (C.c_int, [vp, i32]),
        "bayes_gibbs_set_mode": (C.c_int, [vp, i32]),
        "bayes_gibbs_get_mode": (i32, [vp]),
        "bayes_step_assignment": (C.c_int, [C.c_int, C.POINTER(LocusDesc), vp, u64, C.c_uint32, i32, vp]),
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
        "bayes_seq_model_create_gaussian": (C.c_int, [dbl, dbl, i32, C.POINTER(vp)]),
        "bayes_seq_model_destroy": (None, [vp]),
        "bayes_seq_model_three_prime_prob": (dbl, [vp, i32, i32]),
        "bayes_seq_model_length_prob": (dbl, [vp, i32, i32]),
        "bayes_seq_model_effective_length": (dbl, [vp, i32]),
        "bayes_fragment_model_estimate": (C.c_int, [i32, vp, vp, C.POINTER(dbl), C.POINTER(dbl)]),
        "bayes_collapsed_map_build": (C.c_int, [C.POINTER(RawLocusDesc), vp, i32, vp, C.POINTER(vp)]),
        "bayes_collapsed_map_destroy": (None, [vp]),
        "bayes_collapsed_map_num_candidates": (i32, [vp]),
        "bayes_collapsed_map_num_rows": (i32, [vp]),
        "bayes_collapsed_map_nnz": (i64, [vp]),
        "bayes_collapsed_map_export": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
        "bayes_collapsed_map_stats": (C.c_int, [vp, vp]),
        "bayes_fetch_fragment_lengths": (i64, [C.POINTER(RawLocusDesc), vp, i64]),
        "bayes_all_paths_build": (C.c_int, [C.c_char_p, i32, i32, C.POINTER(vp)]),
        "bayes_path_set_destroy": (None, [vp]),
        "bayes_path_set_num_candidates": (i32, [vp]),
        "bayes_path_set_num_exons": (i32, [vp]),
        "bayes_path_set_export": (C.c_int, [vp, vp, vp, vp, vp, vp]),
        "bayes_path_set_stats": (C.c_int, [vp, vp]),
        "bayes_path_set_graph_name": (C.c_char_p, [vp]),
        "bayes_path_set_reference": (C.c_char_p, [vp]),
        "bayes_path_set_strand": (C.c_char_p, [vp]),
        "bayes_path_set_candidate_name": (C.c_int, [vp, i32, C.c_char_p, i32]),
        "bayes_bam_front_open": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(vp)]),
        "bayes_bam_front_close": (None, [vp]),
        "bayes_bam_front_counts": (C.c_int, [vp, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(dbl)]),
        "bayes_bam_front_num_alignments": (i64, [vp, i32]),
        "bayes_bam_front_write_sam": (C.c_int, [vp, i32, C.c_char_p, i32]),
        "bayes_bam_front_graphs": (C.c_int, [vp, i32, C.c_char_p, C.c_char_p, dbl, i32, C.c_char_p, i32, C.POINTER(i32), C.POINTER(i32),
                                             C.POINTER(i32), C.POINTER(i32)]),
        "bayes_sequencing_probability": (C.c_int, [C.c_char_p, C.c_char_p, i32, C.c_char_p, vp, C.POINTER(dbl)]),
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
        raise BayesError("%s failed (%d): %s" % (what, rc, load_library().bayes_last_error().decode("utf-8", "replace")))  # (messages may quote bytes of a corrupt input file)
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

    def __init__(self, device=0, stream=None, host_threads=0, mode=None):
        self.lib = load_library()
        h = C.c_void_p()
        _check(self.lib.bayes_gibbs_create(device, C.byref(h)), "bayes_gibbs_create")
        self.h = h
        if mode is not None:
            self.set_mode(mode)
        self._T = []
        self._chains = []
        self._lock = threading.Lock()
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

    def _note(self, desc_array, first=None):
        """Pointer to a bayes_locus_desc[] and, once the ids are known, the sizes the fetch calls need."""
        if isinstance(desc_array, np.ndarray):  # structured array laid out as bayes_locus_desc[] (synth.DESC_DTYPE)
            ptr = C.c_void_p(desc_array.ctypes.data)
            shapes = [(int(t), int(c)) for t, c in zip(desc_array["T"], desc_array["n_chains"])]
        else:
            ptr = C.cast(desc_array, C.c_void_p)
            shapes = [(d.T, d.n_chains) for d in desc_array]
        if first is not None:
            with self._lock:  # submit may be called from several threads; ids come from the library
                need = first + len(shapes)
                if len(self._T) < need:
                    self._T.extend([0] * (need - len(self._T)))
                    self._chains.extend([0] * (need - len(self._chains)))
                for i, (t, c) in enumerate(shapes):
                    self._T[first + i] = t
                    self._chains[first + i] = c
        return ptr

    def submit(self, desc_array):
        first = C.c_int64(0)
        _check(self.lib.bayes_gibbs_submit(self.h, len(desc_array), self._note(desc_array), C.byref(first)), "bayes_gibbs_submit")
        self._note(desc_array, first.value)
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
        _check(self.lib.bayes_gibbs_run_batch(self.h, len(desc_array), self._note(desc_array), C.byref(first)), "bayes_gibbs_run_batch")
        self._note(desc_array, first.value)
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
        n = self.lib.bayes_gibbs_unit_stats(self.h, None, None, None)
        shape = np.zeros((n, 8), np.int32)
        clock = np.zeros((n, 3), np.uint64)
        pred = np.zeros((n, 5))
        _check(self.lib.bayes_gibbs_unit_stats(self.h, _ptr(shape), _ptr(clock), _ptr(pred)), "bayes_gibbs_unit_stats")
        self.unit_predicted_s, self.unit_features = pred[:, 0].copy(), pred[:, 1:].copy()
        return shape, clock

    def phase_seconds(self):
        out = np.zeros(6)
        _check(self.lib.bayes_gibbs_phase_seconds(self.h, _ptr(out)), "bayes_gibbs_phase_seconds")
        return dict(zip(["prepare", "pack", "h2d_enqueue", "launch", "d2h_enqueue", "sync"], out.tolist()))

    def last_launches(self):
        return self.lib.bayes_gibbs_last_launches(self.h)

    def set_calibration(self, iterations):
        """-1 = automatic, 0 = off, n = calibration pass of n iterations (bayes_gibbs_set_calibration)."""
        _check(self.lib.bayes_gibbs_set_calibration(self.h, int(iterations)), "bayes_gibbs_set_calibration")

    def set_mode(self, mode):
        """"fast" (production, the default) or "replay" (the reference's order of draws); before upload."""
        _check(self.lib.bayes_gibbs_set_mode(self.h, _MODES[mode]), "bayes_gibbs_set_mode")

    def get_mode(self):
        return "replay" if self.lib.bayes_gibbs_get_mode(self.h) == MODE_REPLAY else "fast"

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


def step_assignment(loc, theta, key, iteration=0, mode="fast", device=0):
    """generateAssignment (theta given) or initAssignment (theta None) in the given sampling mode."""
    d = make_desc(loc, 0, 1, 0, pi=0.5)
    theta = None if theta is None else np.ascontiguousarray(theta, np.float64)
    counts = np.zeros(loc.T, np.int32)
    _check(load_library().bayes_step_assignment(device, C.byref(d), _ptr(theta), int(key), iteration, _MODES[mode], _ptr(counts)),
           "bayes_step_assignment")
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


# ---- likelihood builder --------------------------------------------------------------------------
def make_raw_desc(raw):
    """bayes_raw_locus for an object with the flat arrays of the same names (e.g. oracle.f1_py.RawLocus)."""
    d = RawLocusDesc(raw.n_cand, raw.exon_ptr.ctypes.data, raw.exon_start.ctypes.data, raw.exon_end.ctypes.data,
                     raw.cand_length.ctypes.data, raw.is_premrna.ctypes.data, raw.subgraph_id.ctypes.data, raw.n_subgraphs,
                     raw.strand_plus, raw.n_frag, raw.read_pos.ctypes.data, raw.cigar_ptr.ctypes.data, raw.cigar_op.ctypes.data,
                     raw.cigar_len.ctypes.data, raw.frag_prob.ctypes.data)
    d._keep = raw
    return d


class SequencingModel(object):
    """bayes_seq_model: GaussianFragmentLengthModel + SequencingModel tables (sequencingModel.cpp:34-76)."""

    def __init__(self, mean, sd, max_transcript_length):
        self.lib = load_library()
        h = C.c_void_p()
        _check(self.lib.bayes_seq_model_create_gaussian(mean, sd, max_transcript_length, C.byref(h)), "bayes_seq_model_create_gaussian")
        self.h = h
        self.max_len = max_transcript_length

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.bayes_seq_model_destroy(self.h)
            self.h = None

    def three_prime_prob(self, pos, tran_length):
        return self.lib.bayes_seq_model_three_prime_prob(self.h, pos, tran_length)

    def length_prob(self, fragment_length, pos):
        return self.lib.bayes_seq_model_length_prob(self.h, fragment_length, pos)

    def effective_length(self, length):
        return self.lib.bayes_seq_model_effective_length(self.h, length)


def estimate_fragment_model(hist):
    lengths = np.ascontiguousarray(sorted(hist), np.uint32)
    counts = np.ascontiguousarray([hist[int(k)] for k in lengths], np.uint32)
    med, mad = C.c_double(0), C.c_double(0)
    _check(load_library().bayes_fragment_model_estimate(len(lengths), _ptr(lengths), _ptr(counts), C.byref(med), C.byref(mad)),
           "bayes_fragment_model_estimate")
    return med.value, mad.value


def build_collapsed_map(raw, model, no_coverage_exclude=False):
    """AlignmentParser::calculateFragTranProbabilities through the C ABI -> dict with the CSR arrays and a Locus factory."""
    L = load_library()
    d = make_raw_desc(raw)
    cx = np.bincount(raw.subgraph_id, minlength=raw.n_subgraphs).astype(np.int32)
    h = C.c_void_p()
    rc = _check(L.bayes_collapsed_map_build(C.byref(d), model.h, int(no_coverage_exclude), _ptr(cx), C.byref(h)), "bayes_collapsed_map_build")
    try:
        T, R, nnz = L.bayes_collapsed_map_num_candidates(h), L.bayes_collapsed_map_num_rows(h), L.bayes_collapsed_map_nnz(h)
        out = dict(return_code=rc, T=T, R=R, complexity=cx, row_ptr=np.zeros(R + 1, np.int32), col_idx=np.zeros(nnz, np.int32),
                   val=np.zeros(nnz), counts=np.zeros(R, np.int32), new_to_old=np.zeros(T, np.int32), efflen=np.zeros(T),
                   stats=np.zeros(10, np.int64))
        if rc == 0:
            _check(L.bayes_collapsed_map_export(h, _ptr(out["row_ptr"]), _ptr(out["col_idx"]), _ptr(out["val"]), _ptr(out["counts"]),
                                                _ptr(out["new_to_old"]), _ptr(out["efflen"])), "bayes_collapsed_map_export")
        L.bayes_collapsed_map_stats(h, _ptr(out["stats"]))
    finally:
        L.bayes_collapsed_map_destroy(h)
    return out


def fetch_fragment_lengths(raw):
    L = load_library()
    d = make_raw_desc(raw)
    n = _check(L.bayes_fetch_fragment_lengths(C.byref(d), None, 0), "bayes_fetch_fragment_lengths")
    out = np.zeros(n, np.uint32)
    _check(L.bayes_fetch_fragment_lengths(C.byref(d), _ptr(out), n), "bayes_fetch_fragment_lengths")
    return out


def all_paths(dot, no_pre_mrna=False, max_candidate_number=100):
    """AllPaths(dot).findFullPaths(no_pre_mrna, max_candidate_number, log) through the C ABI -> dict with the candidate
    arrays bayes_raw_locus takes, names, graph name / reference / strand and the numbers of the log block."""
    L = load_library()
    h = C.c_void_p()
    _check(L.bayes_all_paths_build(dot.encode() if isinstance(dot, str) else dot, int(bool(no_pre_mrna)), int(max_candidate_number),
                                   C.byref(h)), "bayes_all_paths_build")
    try:
        n, ne = L.bayes_path_set_num_candidates(h), L.bayes_path_set_num_exons(h)
        out = dict(n_cand=n, exon_ptr=np.zeros(n + 1, np.int32), exon_start=np.zeros(ne, np.int32), exon_end=np.zeros(ne, np.int32),
                   cand_length=np.zeros(n, np.int32), is_premrna=np.zeros(n, np.int32), stats=np.zeros(6, np.int32),
                   graph_name=L.bayes_path_set_graph_name(h).decode(), reference=L.bayes_path_set_reference(h).decode(),
                   strand=L.bayes_path_set_strand(h).decode())
        _check(L.bayes_path_set_export(h, _ptr(out["exon_ptr"]), _ptr(out["exon_start"]), _ptr(out["exon_end"]), _ptr(out["cand_length"]),
                                       _ptr(out["is_premrna"])), "bayes_path_set_export")
        _check(L.bayes_path_set_stats(h, _ptr(out["stats"])), "bayes_path_set_stats")
        buf = C.create_string_buffer(len(out["graph_name"]) + 32)
        names = []
        for i in range(n):
            _check(L.bayes_path_set_candidate_name(h, i, buf, len(buf)), "bayes_path_set_candidate_name")
            names.append(buf.value.decode())
        out["names"] = names
    finally:
        L.bayes_path_set_destroy(h)
    return out


class BamFront(object):
    """BAM front end (row f4, csrc/bamfront.cpp): duplicate removal and strand split of a coordinate-sorted paired-end BAM, SAM text
    for CEM processsam, instance file -> graph file."""

    def __init__(self, bam_path, strand_specific="unstranded"):
        self.lib = load_library()
        self.h = C.c_void_p()
        _check(self.lib.bayes_bam_front_open(os.fsencode(bam_path), strand_specific.encode(), C.byref(self.h)), "bayes_bam_front_open")

    def close(self):
        if self.h:
            self.lib.bayes_bam_front_close(self.h)
            self.h = C.c_void_p()

    def counts(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        _check(self.lib.bayes_bam_front_counts(self.h, C.byref(a), C.byref(b), C.byref(c)), "bayes_bam_front_counts")
        return dict(total_mapped_pairs=a.value, unique_pairs=b.value, pairs_written=c.value)

    def num_alignments(self, strand_file=0):
        return int(self.lib.bayes_bam_front_num_alignments(self.h, strand_file))

    def write_sam(self, path, strand_file=0, with_header=False):
        _check(self.lib.bayes_bam_front_write_sam(self.h, strand_file, os.fsencode(path), int(with_header)), "bayes_bam_front_write_sam")

    def graphs(self, instance_path, graph_file_path, strand=".", strand_file=0, exon_base_coverage=0.05, no_pre_mrna=False, append=False):
        n = [C.c_int32() for _ in range(4)]
        _check(self.lib.bayes_bam_front_graphs(self.h, strand_file, os.fsencode(instance_path), strand.encode(), exon_base_coverage, int(no_pre_mrna),
                                               os.fsencode(graph_file_path), int(append), *[C.byref(x) for x in n]), "bayes_bam_front_graphs")
        return dict(zip(["graphs", "with_fragments", "excluded", "cem_graphs"], [x.value for x in n]))


def sequencing_probability(md_tag, qualities, cigar):
    """calculateSequencingProbability (assembler.cpp:1322-1417); cigar = [(op, length), ...]."""
    L = load_library()
    ops = "".join(o for o, _ in cigar).encode()
    lens = np.asarray([l for _, l in cigar], np.int32)
    out = C.c_double()
    _check(L.bayes_sequencing_probability(md_tag.encode(), qualities.encode(), len(cigar), ops, _ptr(lens), C.byref(out)), "bayes_sequencing_probability")
    return out.value
