"""ctypes bindings for the CPU oracle (oracle/_build/libbayesoracle.so) and, when it
has been built, the shim-compiled UNMODIFIED reference (oracle/_ref/libbayesref.so).

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libbayesoracle.so")
REF_SO = os.path.join(HERE, "_ref", "libbayesref.so")
REF_SRC = os.environ.get("BAYES_REF_SRC", "/root/reference/src")

SEQ, SITE, FAST = 0, 1, 2  # FAST: the production sampling mode of the product (bayes_oracle.c, assignment_fast)

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")


def build(ref=True):
    """Compile the C restatement and, when /root/reference is present, oracle/_ref."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref and os.path.isdir(REF_SRC):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref", "REF_SRC=" + REF_SRC])


class Locus(object):
    """One CollapsedMap (utils.h:129-136) in CSR form."""

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
        assert len(self.row_ptr) == self.R + 1 and len(self.efflen) == self.T

    @property
    def nnz(self):
        return int(self.row_ptr[-1])

    @property
    def n_frag(self):
        return int(self.row_count.sum())

    def dense(self):
        P = np.zeros((self.R, self.T))
        for i in range(self.R):
            s, e = self.row_ptr[i], self.row_ptr[i + 1]
            P[i, self.col_idx[s:e]] = self.val[s:e]
        return P


class _OrcLocus(C.Structure):
    _fields_ = [("T", C.c_int), ("R", C.c_int), ("row_ptr", C.c_void_p), ("col_idx", C.c_void_p), ("val", C.c_void_p),
                ("row_count", C.c_void_p), ("efflen", C.c_void_p), ("adjusted_fragment_count", C.c_double),
                ("gamma", C.c_double)]


def _orc(loc):
    return _OrcLocus(loc.T, loc.R, loc.row_ptr.ctypes.data, loc.col_idx.ctypes.data, loc.val.ctypes.data,
                     loc.row_count.ctypes.data, loc.efflen.ctypes.data, loc.n_total, loc.gamma)


def _opt(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle(object):
    def __init__(self, path=ORACLE_SO):
        if not os.path.exists(path):
            build(ref=False)
        self.lib = L = C.CDLL(path)
        L.orc_min_read_cover.restype = C.c_int
        L.orc_min_read_cover.argtypes = [C.POINTER(_OrcLocus), C.c_int, C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64)]
        L.orc_simplex_table.restype = None
        L.orc_simplex_table.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, _f64p, _i32p]
        L.orc_sample_simplex.restype = C.c_int
        L.orc_sample_simplex.argtypes = [_f64p, _i32p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint32]
        L.orc_init_assignment.restype = C.c_int
        L.orc_init_assignment.argtypes = [C.POINTER(_OrcLocus), C.c_int, C.c_uint64, _i32p, C.POINTER(C.c_uint64)]
        L.orc_generate_assignment.restype = C.c_int
        L.orc_generate_assignment.argtypes = [C.POINTER(_OrcLocus), _f64p, C.c_int, C.c_uint64, C.c_uint32, _i32p,
                                              C.POINTER(C.c_uint64)]
        L.orc_generate_expression.restype = C.c_int
        L.orc_generate_expression.argtypes = [C.c_int, _i32p, C.c_int, C.c_double, C.c_int, C.c_uint64, C.c_uint32, _f64p,
                                              C.c_void_p, C.POINTER(C.c_uint64)]
        L.orc_run_sampler.restype = C.c_int
        L.orc_run_sampler.argtypes = [C.POINTER(_OrcLocus), C.c_double, C.c_int, C.c_int, C.c_int, C.c_uint64,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        L.orc_top_marginals.restype = C.c_int
        L.orc_top_marginals.argtypes = [C.c_int, C.c_int, _i32p, _f64p, _f64p, C.c_double, _i32p, _f64p, _f64p]
        L.orc_test_gamma.argtypes = [C.c_double, C.c_int, C.c_uint64, C.c_int, _f64p]
        L.orc_test_binomial.argtypes = [C.c_int, C.c_double, C.c_int, C.c_uint64, C.c_int, _i32p]
        L.orc_test_uniform_int.argtypes = [C.c_uint32, C.c_int, C.c_uint64, C.c_int, _i32p]
        L.orc_test_philox.argtypes = [C.c_uint64, _u32p, _u32p]
        L.orc_test_mt.argtypes = [C.c_uint32, C.c_int, _u32p]

    def min_read_cover(self, loc, mode, seed):
        flags = np.zeros(loc.T, np.int32)
        w = C.c_uint64(0)
        o = _orc(loc)
        n = self.lib.orc_min_read_cover(C.byref(o), mode, seed, flags.ctypes.data, C.byref(w))
        return n, flags, w.value

    def simplex_table(self, pi, gamma, T, n_frag):
        table = np.zeros((T, T))
        row_len = np.zeros(T, np.int32)
        self.lib.orc_simplex_table(pi, gamma, T, n_frag, table, row_len)
        return table, row_len

    def sample_simplex(self, table, row_len, count_plus, mode, seed, iteration=0):
        return self.lib.orc_sample_simplex(table, row_len, table.shape[0], count_plus, mode, seed, iteration)

    def init_assignment(self, loc, mode, seed):
        counts = np.zeros(loc.T, np.int32)
        w = C.c_uint64(0)
        o = _orc(loc)
        self.lib.orc_init_assignment(C.byref(o), mode, seed, counts, C.byref(w))
        return counts, w.value

    def generate_assignment(self, loc, theta, mode, seed, iteration=0):
        counts = np.zeros(loc.T, np.int32)
        w = C.c_uint64(0)
        o = _orc(loc)
        self.lib.orc_generate_assignment(C.byref(o), np.ascontiguousarray(theta, np.float64), mode, seed, iteration, counts,
                                         C.byref(w))
        return counts, w.value

    def generate_expression(self, counts, b, gamma, mode, seed, iteration=0):
        counts = np.ascontiguousarray(counts, np.int32)
        T = len(counts)
        theta = np.zeros(T)
        order = np.full(T, -1, np.int32)
        w = C.c_uint64(0)
        n = self.lib.orc_generate_expression(T, counts, b, gamma, mode, seed, iteration, theta, order.ctypes.data, C.byref(w))
        return theta, order[:n], w.value

    def run_sampler(self, loc, burn, samples, mode, seed, pi=-1.0, trace_iters=0):
        T = loc.T
        occ = np.zeros(T, np.int32)
        mean = np.zeros(T)
        m2 = np.zeros(T)
        tr = max(trace_iters, 0)
        theta_tr = np.zeros((tr, T)) if tr else None
        counts_tr = np.zeros((tr, T), np.int32) if tr else None
        b_tr = np.zeros(tr, np.int32) if tr else None
        init_counts = np.zeros(T, np.int32)
        pi_out = C.c_double(0)
        w = C.c_uint64(0)
        o = _orc(loc)
        self.lib.orc_run_sampler(C.byref(o), pi, burn, samples, mode, seed, _opt(occ), _opt(mean), _opt(m2), tr,
                                 _opt(theta_tr), _opt(counts_tr), _opt(b_tr), _opt(init_counts), C.byref(pi_out), C.byref(w))
        return dict(occ=occ, mean=mean, m2=m2, theta_trace=theta_tr, counts_trace=counts_tr, b_trace=b_tr,
                    init_counts=init_counts, pi=pi_out.value, words=w.value)

    def top_marginals(self, samples, occ, mean, m2, threshold):
        T = len(occ)
        idx = np.zeros(T, np.int32)
        mo = np.zeros(T)
        sd = np.zeros(T)
        with np.errstate(all="ignore"):
            n = self.lib.orc_top_marginals(T, samples, np.ascontiguousarray(occ, np.int32), np.ascontiguousarray(mean),
                                           np.ascontiguousarray(m2), threshold, idx, mo, sd)
        return idx[:n], mo[:n], sd[:n]

    def gamma(self, alpha, mode, seed, n):
        out = np.zeros(n)
        self.lib.orc_test_gamma(alpha, mode, seed, n, out)
        return out

    def binomial(self, nn, p, mode, seed, n):
        out = np.zeros(n, np.int32)
        self.lib.orc_test_binomial(nn, p, mode, seed, n, out)
        return out

    def uniform_int(self, rng, mode, seed, n):
        out = np.zeros(n, np.int32)
        self.lib.orc_test_uniform_int(rng, mode, seed, n, out)
        return out

    def philox(self, key, ctr):
        out = np.zeros(4, np.uint32)
        self.lib.orc_test_philox(key, np.ascontiguousarray(ctr, np.uint32), out)
        return out

    def mt(self, seed, n):
        out = np.zeros(n, np.uint32)
        self.lib.orc_test_mt(seed, n, out)
        return out


GTF_ARGTYPES = [C.c_int, C.c_int, C.c_int, _i32p, _f64p, C.c_int, _i32p, _i32p, _i32p, _i32p, _i32p, _f64p, C.c_int, C.c_double,
                C.c_double, C.c_int, C.c_char_p, C.c_int]


def call_gtf(fn, case, mode):
    """Shared argument marshalling of ref_gtf (oracle/_ref) and facade_gtf (tests/cpp/facade_driver.cpp)."""
    buf = C.create_string_buffer(1 << 20)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)  # noqa: E731
    n = fn(case["T"], case["samples"], len(case["mask"]), i32(case["mask"]), np.ascontiguousarray(case["values"], np.float64),
           len(case["n_vert"]), i32(case["n_vert"]), i32(case["starts"]), i32(case["ends"]), i32(case["premrna"]), i32(case["old_idx"]),
           np.ascontiguousarray(case["efflen"], np.float64), int(case["total_num_fragments"]), float(case["count_threshold"]),
           float(case["marginal_threshold"]), mode, buf, len(buf))
    assert n >= 0
    return buf.value.decode()


def gtf_case(seed, T=7, samples=40, n_cand=9):
    """Deterministic GibbsOutput + candidate set for the GTF writers."""
    rng = np.random.default_rng(seed)
    mask = (rng.random((samples, T)) < rng.random(T)[None, :]).astype(np.int32)
    mask[:, 0] = 1
    values = rng.gamma(2.0, 20.0, (samples, T)) * rng.random(T)[None, :]
    n_vert = rng.integers(1, 6, n_cand)
    starts, ends = [], []
    for nv in n_vert:
        pos = 1000 + np.cumsum(rng.integers(50, 4000, 2 * nv))
        starts += pos[0::2].tolist()
        ends += pos[1::2].tolist()
    premrna = np.zeros(n_cand, np.int32)
    premrna[rng.integers(n_cand)] = 1
    old_idx = np.sort(rng.choice(n_cand, T, replace=False))
    return dict(T=T, samples=samples, mask=mask, values=values, n_vert=n_vert, starts=starts, ends=ends, premrna=premrna,
                old_idx=old_idx, efflen=rng.uniform(300, 4000, T), total_num_fragments=2000000, count_threshold=12.0,
                marginal_threshold=0.5)


class Reference(object):
    """The unmodified reference sampler TUs behind oracle/ref_driver.cpp (sequential MT19937 only)."""

    def __init__(self, path=REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path + " not built (needs /root/reference; run make -C oracle ref)")
        self.lib = L = C.CDLL(path)
        csr = [C.c_int, C.c_int, _i32p, _i32p, _f64p, _i32p]
        L.ref_run_locus.restype = C.c_int
        L.ref_run_locus.argtypes = csr + [_f64p, C.c_double, C.c_double, C.c_uint32, C.c_int, C.c_int, _i32p, _f64p, _f64p,
                                          C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        L.ref_min_read_cover.restype = C.c_int
        L.ref_min_read_cover.argtypes = csr + [C.c_uint32, C.POINTER(C.c_uint64)]
        L.ref_init_assignment.restype = C.c_int
        L.ref_init_assignment.argtypes = csr + [C.c_uint32, _i32p, C.POINTER(C.c_uint64)]
        L.ref_generate_assignment.restype = C.c_int
        L.ref_generate_assignment.argtypes = csr + [_f64p, C.c_uint32, _i32p, C.POINTER(C.c_uint64)]
        L.ref_generate_expression.restype = C.c_int
        L.ref_generate_expression.argtypes = [C.c_int, _i32p, C.c_int, C.c_int, C.c_double, C.c_uint32, _f64p, _i32p,
                                              C.POINTER(C.c_uint64)]
        L.ref_simplex_table.restype = C.c_int
        L.ref_simplex_table.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, _f64p, _i32p]
        L.ref_sample_simplex.restype = C.c_int
        L.ref_sample_simplex.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_uint32]
        L.ref_top_marginals.restype = C.c_int
        L.ref_top_marginals.argtypes = [C.c_int, C.c_int, _i32p, _f64p, _f64p, C.c_double, _i32p, _f64p, _f64p]
        L.ref_gtf.restype = C.c_int
        L.ref_gtf.argtypes = GTF_ARGTYPES
        L.ref_run_batch.restype = C.c_double
        L.ref_run_batch.argtypes = [C.c_int, _i32p, C.c_void_p, C.c_void_p, C.c_void_p, _i32p, _i32p, _f64p, _i32p, _f64p,
                                    C.c_double, C.c_double, _u32p, _i32p, _i32p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]

    @staticmethod
    def _csr(loc):
        return (loc.R, loc.T, loc.row_ptr, loc.col_idx, loc.val, loc.row_count)

    def run_locus(self, loc, burn, samples, seed):
        T = loc.T
        occ = np.zeros(T, np.int32)
        mean = np.zeros(T)
        m2 = np.zeros(T)
        cover = C.c_int(0)
        pi = C.c_double(0)
        w = C.c_uint64(0)
        self.lib.ref_run_locus(*self._csr(loc), loc.efflen, loc.n_total, loc.gamma, seed, burn, samples, occ, mean, m2,
                               C.byref(cover), C.byref(pi), C.byref(w))
        return dict(occ=occ, mean=mean, m2=m2, cover=cover.value, pi=pi.value, words=w.value)

    def min_read_cover(self, loc, seed):
        w = C.c_uint64(0)
        n = self.lib.ref_min_read_cover(*self._csr(loc), seed, C.byref(w))
        return n, w.value

    def init_assignment(self, loc, seed):
        counts = np.zeros(loc.T, np.int32)
        w = C.c_uint64(0)
        self.lib.ref_init_assignment(*self._csr(loc), seed, counts, C.byref(w))
        return counts, w.value

    def generate_assignment(self, loc, theta, seed):
        counts = np.zeros(loc.T, np.int32)
        w = C.c_uint64(0)
        self.lib.ref_generate_assignment(*self._csr(loc), np.ascontiguousarray(theta, np.float64), seed, counts, C.byref(w))
        return counts, w.value

    def generate_expression(self, counts, n_frag, b, gamma, seed):
        counts = np.ascontiguousarray(counts, np.int32)
        T = len(counts)
        theta = np.zeros(T)
        order = np.full(T, -1, np.int32)
        w = C.c_uint64(0)
        n = self.lib.ref_generate_expression(T, counts, n_frag, b, gamma, seed, theta, order, C.byref(w))
        return theta, order[:n], w.value

    def simplex_table(self, pi, gamma, T, n_frag):
        table = np.zeros((T, T))
        row_len = np.zeros(T, np.int32)
        self.lib.ref_simplex_table(pi, gamma, T, n_frag, table, row_len)
        return table, row_len

    def sample_simplex(self, pi, gamma, T, n_frag, count_plus, seed):
        return self.lib.ref_sample_simplex(pi, gamma, T, n_frag, count_plus, seed)

    def top_marginals(self, samples, occ, mean, m2, threshold):
        T = len(occ)
        idx = np.zeros(T, np.int32)
        mo = np.zeros(T)
        sd = np.zeros(T)
        n = self.lib.ref_top_marginals(T, samples, np.ascontiguousarray(occ, np.int32), np.ascontiguousarray(mean),
                                       np.ascontiguousarray(m2), threshold, idx, mo, sd)
        return idx[:n], mo[:n], sd[:n]

    def gtf(self, case, mode):
        return call_gtf(self.lib.ref_gtf, case, mode)

    def run_batch(self, loci, seeds, burn, samples, n_threads, want_output=False):
        """CPU baseline: thread pool over loci (assembler.cpp:2017-2026). Returns (seconds, outputs|None)."""
        n = len(loci)
        T_arr = np.array([l.T for l in loci], np.int32)
        row_off = np.zeros(n + 1, np.int64)
        cell_off = np.zeros(n + 1, np.int64)
        t_off = np.zeros(n + 1, np.int64)
        for i, l in enumerate(loci):
            row_off[i + 1] = row_off[i] + l.R
            cell_off[i + 1] = cell_off[i] + l.nnz
            t_off[i + 1] = t_off[i] + l.T
        row_ptr = np.concatenate([l.row_ptr for l in loci]).astype(np.int32)
        col = np.concatenate([l.col_idx for l in loci]).astype(np.int32)
        val = np.concatenate([l.val for l in loci]).astype(np.float64)
        cnt = np.concatenate([l.row_count for l in loci]).astype(np.int32)
        eff = np.concatenate([l.efflen for l in loci]).astype(np.float64)
        occ = np.zeros(int(t_off[-1]), np.int32) if want_output else None
        mean = np.zeros(int(t_off[-1])) if want_output else None
        m2 = np.zeros(int(t_off[-1])) if want_output else None
        secs = self.lib.ref_run_batch(n, T_arr, row_off.ctypes.data, cell_off.ctypes.data, t_off.ctypes.data, row_ptr, col,
                                      val, cnt, eff, loci[0].n_total, loci[0].gamma, np.ascontiguousarray(seeds, np.uint32),
                                      np.ascontiguousarray(burn, np.int32), np.ascontiguousarray(samples, np.int32),
                                      n_threads, _opt(occ), _opt(mean), _opt(m2))
        out = None
        if want_output:
            out = [dict(occ=occ[t_off[i]:t_off[i + 1]], mean=mean[t_off[i]:t_off[i + 1]], m2=m2[t_off[i]:t_off[i + 1]])
                   for i in range(n)]
        return secs, out


def random_locus(rng, T, R, max_count=60, density=0.4, gamma=1.0, n_total=1000000, single_frac=0.2):
    """Small random CollapsedMap for unit tests: every row has >= 1 cell, rows are normalised."""
    row_ptr = [0]
    cols, vals = [], []
    for _ in range(R):
        if rng.random() < single_frac or T == 1:
            c = np.array([rng.integers(T)])
        else:
            k = max(2, min(T, int(rng.binomial(T, density)))) if T > 1 else 1
            c = np.sort(rng.choice(T, size=k, replace=False))
        v = rng.random(len(c)) ** 3 + 1e-3
        v = v / v.sum()
        cols.extend(c.tolist())
        vals.extend(v.tolist())
        row_ptr.append(len(cols))
    counts = np.maximum(1, (rng.pareto(1.2, R) * 4).astype(np.int64)).clip(1, max_count)
    efflen = rng.uniform(200, 3000, T)
    return Locus(T, row_ptr, cols, vals, counts, efflen, n_total, gamma)
