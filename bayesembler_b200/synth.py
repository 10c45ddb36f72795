"""Synthetic CollapsedMap workloads for bench.py and the full-size tests (see csrc/synth.cpp and SURVEY.md 8d).

Not part of the sampler: it only manufactures inputs, on the host, for the five BASELINE.json configurations.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import Locus, LocusDesc

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "synth.cpp")
LIB = os.path.join(HERE, "libbayessynth.so")

# BASELINE.json configs: (loci, fragments)
CONFIGS = {1: (200, 1_000_000), 2: (20_000, 20_000_000), 3: (500, 5_000_000), 4: (40_000, 100_000_000), 5: (200, 1_000_000)}

DESC_DTYPE = np.dtype([("T", "<i4"), ("R", "<i4"), ("row_ptr", "<u8"), ("col_idx", "<u8"), ("val", "<u8"), ("row_count", "<u8"),
                       ("efflen", "<u8"), ("total_num_fragments", "<f8"), ("gamma", "<f8"), ("pi", "<f8"), ("burn_in", "<i4"),
                       ("samples", "<i4"), ("seed", "<u8"), ("n_chains", "<i4"), ("reserved", "<i4")])
assert DESC_DTYPE.itemsize == C.sizeof(LocusDesc)


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    subprocess.check_call(["/usr/bin/g++", "-O3", "-std=gnu++17", "-fPIC", "-shared", "-pthread", "-o", LIB, SRC])
    return LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.synth_generate.restype = C.c_void_p
        L.synth_generate.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_double, C.c_int, C.c_void_p]
        L.synth_size.restype = C.c_int64
        L.synth_size.argtypes = [C.c_void_p, C.c_int]
        L.synth_export.restype = None
        L.synth_export.argtypes = [C.c_void_p] * 12
        L.synth_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


class SynthBatch(object):
    """n loci as concatenated CSR arrays; locus i owns rows row_off[i]:row_off[i+1] etc."""

    def __init__(self, cfg, n_loci=None, first=0, mean_frags=None, n_threads=0, gamma=1.0, base_iterations=1000,
                 scale_iterations=60, ids=None):
        L = _load()
        full_n, full_f = CONFIGS[cfg]
        if ids is not None:
            ids = np.ascontiguousarray(ids, np.int64)
            n_loci = len(ids)
        n = full_n if n_loci is None else int(n_loci)
        mean = (full_f / full_n) if mean_frags is None else float(mean_frags)
        n_threads = n_threads or (os.cpu_count() or 1)
        h = L.synth_generate(cfg, first, n, mean, n_threads, None if ids is None else ids.ctypes.data)
        try:
            nr, nc, nt = L.synth_size(h, 1), L.synth_size(h, 2), L.synth_size(h, 3)
            self.n = n
            self.cfg = cfg
            self.T = np.zeros(n, np.int32)
            self.R = np.zeros(n, np.int32)
            self.n_frag = np.zeros(n, np.int64)
            self.row_off = np.zeros(n + 1, np.int64)
            self.cell_off = np.zeros(n + 1, np.int64)
            self.t_off = np.zeros(n + 1, np.int64)
            self.row_ptr = np.zeros(nr + n, np.int32)
            self.col = np.zeros(nc, np.int32)
            self.val = np.zeros(nc, np.float64)
            self.count = np.zeros(nr, np.int32)
            self.efflen = np.zeros(nt, np.float64)
            L.synth_export(h, *[a.ctypes.data for a in (self.T, self.R, self.n_frag, self.row_off, self.cell_off, self.t_off,
                                                        self.row_ptr, self.col, self.val, self.count, self.efflen)])
        finally:
            L.synth_free(h)
        self.gamma = gamma
        # adjusted_fragment_count of the whole library (assembler.cpp:612-619): all fragments of the FULL configuration
        self.n_total = float(full_f)
        # assembler.cpp:1598-1599 with max_subgraph_complexity = T (one sub-graph per locus)
        self.burn = (base_iterations + scale_iterations * self.T).astype(np.int32)
        self.samples = (10 * self.burn).astype(np.int32)
        self.ids = np.arange(first, first + n, dtype=np.int64) if ids is None else ids
        self.seeds = (np.uint64(0x5EED0000) + np.uint64(100003) * self.ids.astype(np.uint64))

    def locus(self, i):
        r0, r1 = self.row_off[i], self.row_off[i + 1]
        c0, c1 = self.cell_off[i], self.cell_off[i + 1]
        t0, t1 = self.t_off[i], self.t_off[i + 1]
        return Locus(self.T[i], self.row_ptr[r0 + i:r1 + i + 1], self.col[c0:c1], self.val[c0:c1], self.count[r0:r1],
                     self.efflen[t0:t1], self.n_total, self.gamma)

    def desc_array(self, n_chains=1, subset=None, pi=-1.0):
        """numpy array laid out as bayes_locus_desc[], pointing into this batch's arrays (keep the batch alive)."""
        idx = np.arange(self.n) if subset is None else np.asarray(subset)
        d = np.zeros(len(idx), DESC_DTYPE)
        d["T"] = self.T[idx]
        d["R"] = self.R[idx]
        d["row_ptr"] = self.row_ptr.ctypes.data + 4 * (self.row_off[idx] + idx)
        d["col_idx"] = self.col.ctypes.data + 4 * self.cell_off[idx]
        d["val"] = self.val.ctypes.data + 8 * self.cell_off[idx]
        d["row_count"] = self.count.ctypes.data + 4 * self.row_off[idx]
        d["efflen"] = self.efflen.ctypes.data + 8 * self.t_off[idx]
        d["total_num_fragments"] = self.n_total
        d["gamma"] = self.gamma
        d["pi"] = pi
        d["burn_in"] = self.burn[idx]
        d["samples"] = self.samples[idx]
        d["seed"] = self.seeds[idx]
        d["n_chains"] = n_chains
        return d

    # ---- workload accounting ----
    def stats(self, n_chains=1, subset=None):
        idx = np.arange(self.n) if subset is None else np.asarray(subset)
        iters = (self.burn[idx].astype(np.int64) + self.samples[idx]) * n_chains
        samp = self.samples[idx].astype(np.int64) * n_chains
        nnz = (self.cell_off[idx + 1] - self.cell_off[idx])
        R = self.R[idx].astype(np.int64)
        T = self.T[idx].astype(np.int64)
        multi = T > 1  # single-candidate loci are closed form in the reference too (gibbsSampler.cpp:57-67)
        return dict(
            loci=int(len(idx)), jobs=int(len(idx) * n_chains), fragments=int(self.n_frag[idx].sum()),
            draws=float((iters * self.n_frag[idx])[multi].sum()),
            row_visits=float((iters * R)[multi].sum()), cell_visits=float((iters * nnz)[multi].sum()),
            iterations=float(iters[multi].sum()),
            # algorithmic bytes, f64 values / u32 columns: 12 nnz + 8 R + 24 T per iteration, + 40 T when sampling
            algo_bytes=float((iters * (12 * nnz + 8 * R + 24 * T) + samp * 40 * T)[multi].sum()),
            rows=int(R.sum()), cells=int(nnz.sum()), candidates=int(T.sum()))


def write_graph_file(path, graphs, unique_pairs, mapped_pairs=None):
    """The input of the host driver (host/assembler.hpp, bayesembler_b200_cli --graph-file): what the reference's grapherCallback
    queues per graph.  graphs = [(dot_strings, fragments, single_path)], fragments = an object with the fragment arrays of
    bayes_raw_locus (read_pos, cigar_ptr, cigar_op, cigar_len, frag_prob)."""
    with open(path, "w") as f:
        f.write("# bayesembler_b200 graph file\n")
        f.write("fragments %d %d\n" % (unique_pairs, unique_pairs if mapped_pairs is None else mapped_pairs))
        for dots, fr, single in graphs:
            n_frag = 0 if fr is None else len(fr.frag_prob)
            f.write("graph %d %d %d\n" % (len(dots), n_frag, int(bool(single))))
            for d in dots:
                assert d.endswith("}\n")
                f.write(d)
            for i in range(n_frag):
                mates = []
                for m in (2 * i, 2 * i + 1):
                    a, b = fr.cigar_ptr[m], fr.cigar_ptr[m + 1]
                    mates.append("%d %s" % (fr.read_pos[m], "".join("%d%s" % (fr.cigar_len[k], chr(fr.cigar_op[k])) for k in range(a, b))))
                f.write("%s %s %.17g\n" % (mates[0], mates[1], fr.frag_prob[i]))
