"""Likelihood builder -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

 * RawLocus / random_raw_locus: candidates (exon chains in genome coordinates) and paired-end fragment alignments
   (positions + CIGARs) at the level AlignmentParser::calculateFragTranProbabilities consumes (alignmentParser.cpp:251);
 * ReferenceF1: ctypes binding of oracle/_ref/libbayesref_f1.so, the UNMODIFIED reference translation units
   alignmentParser.cpp / sequencingModel.cpp / fragmentLengthModel.cpp behind oracle/ref_f1_driver.cpp;
 * OracleF1: ctypes binding of the plain-C restatement oracle/f1_oracle.c.
Only tests/ and bench-style tools may import this module.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_F1_SO = os.path.join(HERE, "_ref", "libbayesref_f1.so")
ORACLE_F1_SO = os.path.join(HERE, "_build", "libbayesoracle_f1.so")

READ = 76


class RawLocus(object):
    """Flat arrays in the layout of bayes_raw_locus (include/bayesembler_b200.h)."""

    def __init__(self, exon_ptr, exon_start, exon_end, cand_length, is_premrna, subgraph_id, strand_plus, read_pos, cigar_ptr,
                 cigar_op, cigar_len, frag_prob):
        i32 = lambda a: np.ascontiguousarray(a, np.int32)  # noqa: E731
        self.exon_ptr, self.exon_start, self.exon_end = i32(exon_ptr), i32(exon_start), i32(exon_end)
        self.cand_length, self.is_premrna, self.subgraph_id = i32(cand_length), i32(is_premrna), i32(subgraph_id)
        self.strand_plus = int(strand_plus)
        self.read_pos = np.ascontiguousarray(read_pos, np.uint32)
        self.cigar_ptr = i32(cigar_ptr)
        self.cigar_op = np.ascontiguousarray(cigar_op, np.int8)
        self.cigar_len = np.ascontiguousarray(cigar_len, np.uint32)
        self.frag_prob = np.ascontiguousarray(frag_prob, np.float64)
        self.n_cand = len(self.cand_length)
        self.n_frag = len(self.frag_prob)
        self.n_subgraphs = int(self.subgraph_id.max()) + 1 if self.n_cand else 0

    def complexity(self):
        return np.bincount(self.subgraph_id, minlength=self.n_subgraphs).astype(np.int32)


def _read_cigar(exons, tpos, length):
    """Genome position (0-based) and CIGAR of `length` transcript bases starting at transcript offset tpos."""
    ops, lens = [], []
    off = 0
    pos = None
    left = length
    for k, (s, e) in enumerate(exons):
        el = e - s + 1
        if tpos < off + el and left > 0:
            a = max(tpos, off)
            take = min(left, off + el - a)
            if pos is None:
                pos = s + (a - off) - 1
            else:
                ops.append(ord("N"))
                lens.append(s - exons[k - 1][1] - 1)
            ops.append(ord("M"))
            lens.append(take)
            left -= take
            tpos = a + take
        off += el
    assert left == 0 and pos is not None
    return pos, ops, lens


def random_raw_locus(rng, n_exons=6, n_cand=5, n_frag=400, strand_plus=True, premrna=True, indel_rate=0.03, junk_rate=0.05,
                     frag_mean=200.0, frag_sd=20.0, two_subgraphs=False):
    """A synthetic locus: exons, candidates = distinct exon subsets (+ optional pre-mRNA), fragments sampled from the candidates."""
    starts, pos = [], 1000
    exons = []
    for _ in range(n_exons):
        length = int(np.clip(np.exp(np.log(150) + 0.6 * rng.normal()), 80, 1200))
        exons.append((pos, pos + length - 1))
        pos += length + int(rng.integers(60, 900))
    cands = []
    masks = set()
    tries = 0
    while len(cands) < n_cand and tries < 200:
        tries += 1
        f = 0 if rng.random() < 0.7 else int(rng.integers(0, n_exons // 2 + 1))
        g = n_exons - 1 if rng.random() < 0.7 else int(rng.integers(f, n_exons))
        m = tuple(e for e in range(f, g + 1) if e in (f, g) or rng.random() < 0.7)
        if m in masks or sum(exons[e][1] - exons[e][0] + 1 for e in m) < 320:
            continue
        masks.add(m)
        cands.append([exons[e] for e in m])
    if not cands:
        cands.append(list(exons))
    is_pre = [0] * len(cands)
    if premrna:
        cands.append([(exons[0][0], exons[-1][1])])
        is_pre.append(1)
    sub = [0] * len(cands)
    if two_subgraphs:
        sub = [i % 2 for i in range(len(cands))]
    return _sample_fragments(rng, cands, is_pre, sub, n_frag, strand_plus, premrna, indel_rate, junk_rate, frag_mean, frag_sd)


def _sample_fragments(rng, cands, is_pre, sub, n_frag, strand_plus, premrna_last, indel_rate, junk_rate, frag_mean, frag_sd):
    """Paired-end fragments sampled from candidates (lists of (start, end) exons) with random abundances."""
    lengths = [sum(e - s + 1 for s, e in c) for c in cands]
    rho = rng.gamma(0.5, size=len(cands)) * (rng.random(len(cands)) > 0.3)
    if premrna_last:
        rho[-1] *= 0.05
    if rho.sum() == 0:
        rho[0] = 1.0
    w = rho * np.array(lengths)
    w = w / w.sum()
    read_pos, cigar_ptr, cigar_op, cigar_len, prob = [], [0], [], [], []
    for _ in range(n_frag):
        t = int(rng.choice(len(cands), p=w))
        L = lengths[t]
        fl = int(np.clip(round(frag_mean + frag_sd * rng.normal()), READ + 5, L))
        s = int(rng.integers(0, L - fl + 1))
        for m, (tp, rl) in enumerate([(s, READ), (s + fl - READ, READ)]):
            p, ops, lens = _read_cigar(cands[t], tp, rl)
            if rng.random() < junk_rate:
                p += int(rng.integers(1, 40))  # shifted read: usually incompatible with every candidate
            if rng.random() < indel_rate and lens[0] > 20:  # TopHat-style I / D inside the first match
                cut = int(rng.integers(5, lens[0] - 5))
                op = ord("I") if rng.random() < 0.5 else ord("D")
                ops = [ord("M"), op, ord("M")] + ops[1:]
                lens = [cut, int(rng.integers(1, 3)), lens[0] - cut] + lens[1:]
            read_pos.append(p)
            cigar_op += ops
            cigar_len += lens
            cigar_ptr.append(len(cigar_op))
        prob.append(float(rng.uniform(0.2, 1.0)))
    exon_ptr = np.cumsum([0] + [len(c) for c in cands])
    es = [s for c in cands for s, _ in c]
    ee = [e for c in cands for _, e in c]
    return RawLocus(exon_ptr, es, ee, lengths, is_pre, sub, strand_plus, read_pos, cigar_ptr, cigar_op, cigar_len, prob)


def fragments_for_candidates(rng, exon_ptr, exon_start, exon_end, cand_length, is_premrna, n_frag=400, strand_plus=True, indel_rate=0.0,
                             junk_rate=0.02, frag_mean=200.0, frag_sd=20.0):
    """Fragments for GIVEN candidates in the flat layout of bayes_path_set_export (row f2 -> row f1)."""
    cands = [[(int(exon_start[k]), int(exon_end[k])) for k in range(exon_ptr[i], exon_ptr[i + 1])] for i in range(len(cand_length))]
    assert [sum(e - s + 1 for s, e in c) for c in cands] == [int(x) for x in cand_length]
    return _sample_fragments(rng, cands, [int(x) for x in is_premrna], [0] * len(cands), n_frag, strand_plus, False, indel_rate, junk_rate,
                             frag_mean, frag_sd)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _call_collapsed_map(fn, loc, frag_mean, frag_sd, max_len, no_coverage_exclude, r_cap=None):
    r_cap = r_cap or max(16, loc.n_frag)
    T, R = C.c_int(0), C.c_int(0)
    P = np.zeros((r_cap, max(1, loc.n_cand)))
    counts = np.zeros(r_cap, np.int32)
    new_to_old = np.zeros(max(1, loc.n_cand), np.int32)
    efflen = np.zeros(max(1, loc.n_cand))
    cx = loc.complexity()
    rc = fn(loc.n_cand, _p(loc.exon_ptr), _p(loc.exon_start), _p(loc.exon_end), _p(loc.cand_length), _p(loc.is_premrna),
            _p(loc.subgraph_id), loc.strand_plus, loc.n_frag, _p(loc.read_pos), _p(loc.cigar_ptr), _p(loc.cigar_op), _p(loc.cigar_len),
            _p(loc.frag_prob), C.c_double(frag_mean), C.c_double(frag_sd), int(max_len), int(no_coverage_exclude), loc.n_subgraphs, _p(cx),
            r_cap, C.byref(T), C.byref(R), _p(P), _p(counts), _p(new_to_old), _p(efflen))
    t, r = T.value, R.value
    return dict(return_code=rc, T=t, R=r, P=P.ravel()[:r * t].reshape(r, t).copy(), counts=counts[:r].copy(), new_to_old=new_to_old[:t].copy(),
                efflen=efflen[:t].copy(), complexity=cx)


_CM_ARGS = [C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                                                                   C.c_void_p, C.c_int, C.POINTER(C.c_int),
                                                                                   C.POINTER(C.c_int)] + [C.c_void_p] * 4


class _F1(object):
    prefix = ""

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = L = C.CDLL(path)
        fn = getattr(L, self.prefix + "collapsed_map")
        fn.restype = C.c_int
        fn.argtypes = _CM_ARGS
        self._cm = fn
        st = getattr(L, self.prefix + "sequencing_tables")
        st.argtypes = [C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        self._st = st
        em = getattr(L, self.prefix + "estimate_fragment_model")
        em.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        self._em = em

    def collapsed_map(self, loc, frag_mean=200.0, frag_sd=20.0, max_len=None, no_coverage_exclude=False):
        max_len = max_len or int(loc.cand_length.max())
        return _call_collapsed_map(self._cm, loc, frag_mean, frag_sd, max_len, no_coverage_exclude)

    def sequencing_tables(self, mean, sd, max_len):
        a, b, c = np.zeros(max_len), np.zeros(max_len), np.zeros(max_len)
        self._st(mean, sd, max_len, _p(a), _p(b), _p(c))
        return dict(three_prime=a, length_prob=b, efflen=c)

    def estimate_fragment_model(self, hist):
        lengths = np.ascontiguousarray(sorted(hist), np.uint32)
        counts = np.ascontiguousarray([hist[int(k)] for k in lengths], np.uint32)
        med, mad = C.c_double(0), C.c_double(0)
        self._em(len(lengths), _p(lengths), _p(counts), C.byref(med), C.byref(mad))
        return med.value, mad.value


class ReferenceF1(_F1):
    prefix = "ref_"

    def __init__(self, path=REF_F1_SO):
        _F1.__init__(self, path)
        self.lib.ref_fetch_fragment_lengths.restype = C.c_int

    def fragment_length_histogram(self, loc):
        cap = 100000
        ln, ct = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
        n = self.lib.ref_fetch_fragment_lengths(loc.n_cand, _p(loc.exon_ptr), _p(loc.exon_start), _p(loc.exon_end), _p(loc.cand_length),
                                                loc.strand_plus, loc.n_frag, _p(loc.read_pos), _p(loc.cigar_ptr), _p(loc.cigar_op),
                                                _p(loc.cigar_len), _p(ln), _p(ct), cap)
        return dict(zip(ln[:n].tolist(), ct[:n].tolist()))


class OracleF1(_F1):
    prefix = "orc_"

    def __init__(self, path=ORACLE_F1_SO):
        _F1.__init__(self, path)
        self.lib.orc_fetch_fragment_lengths.restype = C.c_long

    def fragment_lengths(self, loc):
        cap = loc.n_frag * max(1, loc.n_cand)
        out = np.zeros(max(1, cap), np.uint32)
        n = self.lib.orc_fetch_fragment_lengths(loc.n_cand, _p(loc.exon_ptr), _p(loc.exon_start), _p(loc.exon_end), loc.strand_plus,
                                                loc.n_frag, _p(loc.read_pos), _p(loc.cigar_ptr), _p(loc.cigar_op), _p(loc.cigar_len),
                                                _p(out), C.c_long(cap))
        return out[:n]
