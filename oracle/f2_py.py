"""Candidate enumeration -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

 * all_paths_oracle: plain-Python restatement of AllPaths (allPaths.cpp:42-343) that follows the reference step by step
   (recursive enumeration, junction threshold raised by one per aborted attempt); small graphs only;
 * ReferenceF2: ctypes binding of oracle/_ref/libbayesref_f2.so, the UNMODIFIED reference translation unit allPaths.cpp
   behind oracle/ref_f2_driver.cpp, compiled against the Boost.Graph stand-in oracle/shim/orc_graph_shim.hpp;
 * random_splice_graph / dot_text: seeded splice graphs in the DOT mini-format the reference writes
   (assembler.cpp:764-861).
Boost.Graph is absent here and unpinned by the reference; the vertex order read_graphviz produces (node ids sorted as
strings) is taken from its published implementation -- parity is pinned against the reference's own code compiled
against that stand-in, unpinned against a Boost-linked binary.
Only tests/ and bench-style tools may import this module.
"""
import ctypes as C
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_F2_SO = os.path.join(HERE, "_ref", "libbayesref_f2.so")


# ------------------------------------------------------------------------------------------------------------ generator
def dot_text(name, reference, strand, exons, junctions, pre_mrna=True, left=None, right=None):
    """The text graphConstructorCallback assembles (assembler.cpp:764, 824, 849-858): exons = [(start, end)] numbered
    from 0, junctions = [(i, j, coverage)] in the order they are to be written."""
    out = ["digraph %s {\n" % name]
    for i, (s, e) in enumerate(exons):
        out.append('%d [label="%s;%s;%d;%d"]\n' % (i, reference, strand, s, e))
    if pre_mrna:
        lo = min(s for s, _ in exons) if left is None else left
        hi = max(e for _, e in exons) if right is None else right
        out.append('pre_mrna [label="%s;%s;%d;%d"]\n' % (reference, strand, lo, hi))
    for i, j, c in junctions:
        out.append('%d->%d [coverage="%d"]\n' % (i, j, c))
    out.append("}\n")
    return "".join(out)


def random_splice_graph(seed, n_exons, p_edge=0.35, p_adjacent=0.2, max_cov=40, pre_mrna=True, name=None, strand="+"):
    """A random DAG over n_exons exons (edges i -> j, i < j), junction coverages with many ties, some exons directly
    adjacent (end + 1 == next start: merged in the candidate, allPaths.cpp:237-244), junction statements shuffled the way
    an unordered_map would leave them (assembler.cpp:855-858)."""
    rng = np.random.RandomState(seed)
    pos = 1000 + int(rng.randint(0, 5000))
    exons = []
    for i in range(n_exons):
        length = int(rng.randint(30, 400))
        exons.append((pos, pos + length - 1))
        pos += length + (0 if rng.rand() < p_adjacent else int(rng.randint(50, 3000)))
    junctions = []
    for i in range(n_exons):
        for j in range(i + 1, n_exons):
            near = j - i <= 3
            if rng.rand() < (p_edge if near else p_edge * 0.15):
                junctions.append((i, j, int(rng.randint(1, max_cov + 1)) if rng.rand() < 0.7 else int(rng.choice([1, 2, 5]))))
    rng.shuffle(junctions)
    junctions = [tuple(int(x) for x in j) for j in junctions]
    return dot_text(name or ("graph_%d_%s" % (seed, "plus" if strand == "+" else "minus")), "chr%d" % (1 + seed % 22), strand, exons,
                    junctions, pre_mrna=pre_mrna)


# ------------------------------------------------------------------------------------------------------------ restatement
class _TooMany(Exception):
    pass


def _parse(dot):
    """parseGraphviz + parseVertexProperties (allPaths.cpp:52-113).  Vertices in the order of their ids as strings."""
    m = re.search(r"^digraph\s(\S+)\s\{$", dot, re.M)  # :69-74
    assert m, "no 'digraph <name> {' line"
    name = m.group(1)
    body = dot[dot.index("{") + 1:dot.rindex("}")]
    labels, edges, ids = {}, [], set()
    for line in body.split("\n"):
        line = line.strip()
        if not line:
            continue
        me = re.match(r'^(\S+?)\s*->\s*(\S+?)\s*\[\s*coverage\s*=\s*"?(-?\d+)"?\s*\]$', line)
        if me:
            edges.append((me.group(1), me.group(2), int(me.group(3))))
            ids.update(me.group(1, 2))
            continue
        mn = re.match(r'^(\S+?)\s*\[\s*label\s*=\s*"([^"]*)"\s*\]$', line)
        assert mn, "unparsable DOT statement %r" % line
        labels[mn.group(1)] = mn.group(2)
        ids.add(mn.group(1))
    order = sorted(ids)  # std::map<node_name, ...> order (bytewise)
    index = dict((v, i) for i, v in enumerate(order))
    verts = []
    for v in order:
        f = labels[v].split(";")  # :88
        verts.append(dict(name=v, reference=f[0], strand=f[1], start=int(f[2]), end=int(f[3])))
    adj = [[] for _ in order]
    for a, b, c in edges:
        adj[index[a]].append([index[b], c])
    return name, verts, adj


def all_paths_oracle(dot, no_pre_mrna=False, max_candidate_number=100):
    """AllPaths(dot).findFullPaths(no_pre_mrna, max_candidate_number, log), step for step."""
    graph_name, verts, adj = _parse(dot)
    V = len(verts)
    n_edges = sum(len(a) for a in adj)
    non_source = set(t for a in adj for t, _ in a)  # :120-130
    non_sinks = set(i for i, a in enumerate(adj) if a)
    stats = [V, V - len(non_source), V - len(non_sinks), 0, 0]
    if not non_source and not non_sinks and not no_pre_mrna and V == 2:  # :136-167
        by = dict((v["name"], v) for v in verts)
        if "0" in by and "pre_mrna" in by and by["0"]["start"] == by["pre_mrna"]["start"] and by["0"]["end"] == by["pre_mrna"]["end"]:
            k = [v["name"] for v in verts].index("pre_mrna")
            del verts[k]
            del adj[k]
    junction_threshold = 1
    too_many = True
    cands = []
    while too_many:  # :173
        cands = []
        for a in adj:  # :181-194
            a[:] = [e for e in a if not junction_threshold > e[1]]
        non_source = set(t for a in adj for t, _ in a)  # :196-199
        path_count = [0]

        def find(s, path):  # findAllPathsSource :307-343
            path = path + [s]
            if not adj[s]:
                path_count[0] += 1
                if path_count[0] > max_candidate_number:
                    raise _TooMany()
                return [path]
            out = []
            for t, _ in adj[s]:
                out.extend(find(t, path))
            return out

        too_many = False
        for v in range(len(verts)):  # :202
            if v in non_source:
                continue
            try:
                paths = find(v, [])
            except _TooMany:
                too_many = True
                junction_threshold += 1
                stats[3] += 1
                break
            for p in paths:  # :221-286
                exons = []
                length = 0
                for j, u in enumerate(p):
                    length += verts[u]["end"] - verts[u]["start"] + 1
                    if j > 0 and verts[p[j - 1]]["end"] + 1 == verts[u]["start"]:
                        exons[-1][1] = verts[u]["end"]
                    else:
                        exons.append([verts[u]["start"], verts[u]["end"]])
                pre = verts[p[0]]["name"] == "pre_mrna"
                if pre:
                    assert len(p) == 1
                    if no_pre_mrna:
                        continue
                idx = len(cands)
                cands.append(dict(idx=idx, name=graph_name + ("_pre" if pre else "_seq%d" % idx), length=length, exons=exons, pre=pre))
        if sum(len(a) for a in adj) == 0:  # :291-294
            too_many = False
    stats[4] = len(cands)
    return dict(graph_name=graph_name, reference=verts[0]["reference"], strand=verts[0]["strand"], candidates=cands, stats=stats,
                n_edges_in=n_edges)


def flatten(res):
    """Candidate list of all_paths_oracle -> the flat arrays of bayes_path_set_export."""
    c = res["candidates"]
    ptr = np.zeros(len(c) + 1, np.int32)
    st, en = [], []
    for i, x in enumerate(c):
        for s, e in x["exons"]:
            st.append(s)
            en.append(e)
        ptr[i + 1] = len(st)
    return dict(n_cand=len(c), exon_ptr=ptr, exon_start=np.array(st, np.int32), exon_end=np.array(en, np.int32),
                cand_length=np.array([x["length"] for x in c], np.int32), is_premrna=np.array([int(x["pre"]) for x in c], np.int32),
                names=[x["name"] for x in c], graph_name=res["graph_name"], stats=np.array(res["stats"], np.int32))


# ------------------------------------------------------------------------------------------------------------ reference
class ReferenceF2(object):
    """The UNMODIFIED allPaths.cpp (oracle/_ref/libbayesref_f2.so)."""

    def __init__(self, path=REF_F2_SO):
        self.lib = C.CDLL(path)
        self.lib.ref_all_paths.restype = C.c_int

    def all_paths(self, dot, no_pre_mrna=False, max_candidate_number=100, cap_cand=20000, cap_exons=2000000):
        z = lambda n: np.zeros(n, np.int32)  # noqa: E731
        ptr, st, en, ln, pre, idx, stats = z(cap_cand + 1), z(cap_exons), z(cap_exons), z(cap_cand), z(cap_cand), z(cap_cand), z(5)
        names = C.create_string_buffer(64 * cap_cand + 64)
        gname = C.create_string_buffer(512)
        P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        n = self.lib.ref_all_paths(dot.encode(), int(bool(no_pre_mrna)), int(max_candidate_number), cap_cand, cap_exons, P(ptr), P(st),
                                   P(en), P(ln), P(pre), P(idx), names, len(names), gname, len(gname), P(stats))
        assert n >= 0, "capacity too small"
        ne = int(ptr[n])
        assert np.array_equal(idx[:n], np.arange(n))
        return dict(n_cand=n, exon_ptr=ptr[:n + 1].copy(), exon_start=st[:ne].copy(), exon_end=en[:ne].copy(), cand_length=ln[:n].copy(),
                    is_premrna=pre[:n].copy(), names=names.value.decode().split("\n")[:-1], graph_name=gname.value.decode(),
                    stats=stats.copy())
