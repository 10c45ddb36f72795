"""Host driver and command line: bayesembler_b200_cli = the reference's options (main.cpp:52-154) in
front of Assembler::runBayesemblerBatched (host/assembler.hpp) = bayesemblerCallback + fragmentLengthCallback + outputCallback
(assembler.cpp:1420-1839) from the splice-graph stage on.

CPU: option parsing (names, defaults, checks, help), the graph-file reader, the graphs that never reach the sampler (return
codes 2 / 3 / 4) and the refusal to sample without a GPU.  GPU: a full run whose assembly.gtf / candidates.gtf /
fragment_lengths.txt are compared, byte for byte, with the oracles chained the same way (oracle/f2_py -> oracle/f1_oracle.c ->
oracle/bayes_oracle.c in the sampling mode the driver ran in (fast by default, BAYES_MODE=replay for the other) -> the GTF writer that tests/test_host_facade.py pins against the reference's)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import f1_py, f2_py

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cli(lib):
    from bayesembler_b200 import _build
    return _build.build_cli()


def run(cli, *args, cwd=None, env=None):
    p = subprocess.run([cli] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=cwd, timeout=600,
                       env=None if env is None else dict(os.environ, **env))
    return p.returncode, p.stdout, p.stderr


# ------------------------------------------------------------------------------------------------------------ options
def test_help_lists_the_reference_options(cli):
    rc, out, _ = run(cli)
    assert rc == 1 and "== Required arguments ==" in out and "--bam-file" in out and "--seed" not in out  # main.cpp:119-122
    rc, out, _ = run(cli, "-h")
    assert rc == 1 and "--confidence-threshold" in out and "--seed" not in out
    rc, adv, _ = run(cli, "--advanced")
    assert rc == 1
    for name in ("bam-file", "output-prefix", "num-threads", "strand-specific", "confidence-threshold", "count-threshold", "no-pre-mRNA", "seed",
                 "output-mode", "library-size", "keep-temp-files", "max-candidate-number", "dirichlet-parameter", "frag-mean", "frag-sd",
                 "gibbs-iteration-scaling"):  # every option of main.cpp:64-92
        assert "--" + name in adv, name
    for letter in "bopscfm":
        assert "-%s [" % letter in adv


@pytest.mark.parametrize("args,msg", [
    ((), None),
    (("--output-prefix", "x"), "'--bam-file' is required"),
    (("-b", "reads.bam"), "--cem-processsam-path"),
    (("-b", "reads.bam", "-s", "second", "--instance-file", "x.instance"), "--instance-file-plus and --instance-file-minus"),
    (("-g", "g.txt", "--strand-specific", "third"), "--strand-specific argument need to be either"),
    (("-g", "g.txt", "--output-mode", "verbose"), "--output_mode argument need to be either"),
    (("-g", "g.txt", "--frag-mean", "200"), "--frag-mean and --frag-sd must be given together"),
    (("-g", "g.txt", "--num-threads", "many"), "is invalid"),
    (("-g", "g.txt", "--seed"), "is missing"),
    (("-g", "g.txt", "--no-such-option"), "unrecognised option"),
    (("-g", "/nonexistent/graphs.txt", "--frag-mean", "200", "--frag-sd", "20"), "cannot open graph file"),
])
def test_option_errors(cli, args, msg):
    rc, out, err = run(cli, *args)
    assert rc == 1
    if msg:
        assert msg in err, err


# ------------------------------------------------------------------------------------------------------------ graph files
def _fragments(rng, dot, n_frag, no_pre_mrna=False, max_cand=100):
    """Fragments sampled from the candidates of a graph (oracle enumeration), in the arrays of bayes_raw_locus."""
    ps = f2_py.flatten(f2_py.all_paths_oracle(dot, no_pre_mrna, max_cand))
    keep = [i for i in range(ps["n_cand"]) if ps["cand_length"][i] >= 400]
    assert keep, "graph without a candidate long enough to carry 2 x 76 bp fragments"
    ptr = [0]
    st, en = [], []
    for i in keep:
        st += ps["exon_start"][ps["exon_ptr"][i]:ps["exon_ptr"][i + 1]].tolist()
        en += ps["exon_end"][ps["exon_ptr"][i]:ps["exon_ptr"][i + 1]].tolist()
        ptr.append(len(st))
    return f1_py.fragments_for_candidates(rng, np.array(ptr), st, en, ps["cand_length"][keep], ps["is_premrna"][keep], n_frag=n_frag,
                                          strand_plus="_plus" in dot.split("\n")[0])


def _long_exon_graph(seed, n_exons, p_edge, name, strand="+", pos=5000):
    """random_splice_graph with exons long enough for paired-end fragments"""
    rng = np.random.RandomState(seed)
    exons = []
    for _ in range(n_exons):
        length = int(rng.randint(250, 900))
        exons.append((pos, pos + length - 1))
        pos += length + int(rng.randint(80, 2000))
    junctions = [(i, j, int(rng.randint(1, 30))) for i in range(n_exons) for j in range(i + 1, n_exons) if rng.rand() < (p_edge if j - i <= 2 else 0.1)]
    rng.shuffle(junctions)
    return f2_py.dot_text(name, "chr7", strand, exons, [tuple(int(x) for x in j) for j in junctions])


def _make_run(tmp_path, seed=5, n_multi=6, n_single=3):
    from bayesembler_b200 import synth
    rng = np.random.default_rng(seed)
    graphs = []
    for g in range(n_multi):
        dot = _long_exon_graph(100 + g, 3 + g % 5, 0.7, "graph_%d_%s" % (g, "plus" if g % 2 == 0 else "minus"), "+" if g % 2 == 0 else "-")
        graphs.append(([dot], _fragments(rng, dot, 150 + 120 * g), False))
    for g in range(n_single):  # one long exon: the fragment-length pass reads these (assembler.cpp:1651-1727)
        dot = f2_py.dot_text("graph_%d_plus" % (50 + g), "chr7", "+", [(90000 + 9000 * g, 90000 + 9000 * g + 2999 + 400 * g)], [])
        graphs.append(([dot], _fragments(rng, dot, 500), True))
    # a graph whose fragments match nothing (candidate coverage, return code 3 or 4) and a two-sub-graph locus
    lone = f2_py.dot_text("graph_70_plus", "chr7", "+", [(200000, 200800)], [])
    far = _fragments(rng, f2_py.dot_text("graph_71_plus", "chr7", "+", [(300000, 301500)], []), 40)
    graphs.append(([lone], far, False))
    d1 = _long_exon_graph(300, 4, 0.8, "graph_80_plus")
    d2 = _long_exon_graph(301, 3, 0.8, "graph_81_plus", pos=150000)  # second sub-graph, far downstream
    f12 = _fragments(rng, d1, 400)
    graphs.append(([d1, d2], f12, False))
    path = os.path.join(str(tmp_path), "graphs.txt")
    n_pairs = sum(len(g[1].frag_prob) for g in graphs)
    synth.write_graph_file(path, graphs, n_pairs, n_pairs + 1000)
    return path, graphs, n_pairs


def test_graph_file_errors(cli, tmp_path):
    bad = os.path.join(str(tmp_path), "bad.txt")
    for text, msg in (("graph 1 0 0\ndigraph g {\n}\n", "no 'fragments' line"),
                      ("fragments 10\ngraph 1 1 0\ndigraph g {\n0 [label=\"c;+;1;500\"]\n}\n", "fragment line expected"),
                      ("fragments 10\ngraph 1 1 0\ndigraph g {\n0 [label=\"c;+;1;500\"]\n}\n10 76M 200 7x6M 1.0\n", "malformed CIGAR"),
                      ("fragments 10\ngraph 1 0 0\ndigraph g {\n0 [label=\"c;+;1;500\"]\n", "without a closing"),
                      ("fragments 10\nlocus 1 0 0\n", "expected, found 'locus'")):
        open(bad, "w").write(text)
        rc, out, err = run(cli, "-g", bad, "--frag-mean", "200", "--frag-sd", "20")
        assert rc == 1 and msg in err, (msg, err)


def test_graphs_that_never_reach_the_sampler(cli, tmp_path):
    """Return codes 2-4 are decided on the host (assembler.cpp:1498-1531): such a run needs no GPU and writes an empty assembly."""
    from bayesembler_b200 import synth
    rng = np.random.default_rng(1)
    lone = f2_py.dot_text("graph_70_plus", "chr7", "+", [(200000, 200800)], [])
    far = _fragments(rng, f2_py.dot_text("graph_71_plus", "chr7", "+", [(300000, 301500)], []), 40)
    path = os.path.join(str(tmp_path), "g.txt")
    synth.write_graph_file(path, [([lone], far, False), ([lone.replace("graph_70", "graph_72")], None, False)], 40)
    prefix = os.path.join(str(tmp_path), "out", "run1_")
    rc, out, err = run(cli, "-g", path, "-o", prefix, "--frag-mean", "200", "--frag-sd", "20", "--output-mode", "full", "--seed", "3", "-p", "3")
    assert rc == 0, err
    assert "2 graph(s) were parsed in total" in out and "0 graph(s) were assembled successfully" in out
    assert open(prefix + "assembly.gtf").read() == "" and open(prefix + "candidates.gtf").read() == ""
    log = open(prefix + "graph_logs.txt").read()
    assert "###### graph_70_plus ######" in log and "Seeding graph random number generator with 4" in log  # --seed + graphs left (:1437)
    assert "Number of vertices in graph: 2" in log


def test_no_cpu_sampler_behind_the_cli(cli, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    path, graphs, n_pairs = _make_run(tmp_path, n_multi=1, n_single=0)
    rc, out, err = run(cli, "-g", path, "--frag-mean", "200", "--frag-sd", "20", "-o", os.path.join(str(tmp_path), "x_"))
    assert rc == 1 and "no CUDA device" in err and "no CPU path" in err


# ------------------------------------------------------------------------------------------------------------ full run (GPU)
def _expected(lib, oracle, facade, graphs, n_pairs, seed, frag_mean, frag_sd, no_pre=False, max_cand=100, base=1000, scale=60,
              count_threshold=12.0, marginal_threshold=0.5, replay=False):
    """The reference's control flow (assembler.cpp:1420-1631) on the oracles."""
    from oracle.oracle_py import FAST, SITE, Locus
    orc_f1 = f1_py.OracleF1()
    order = [g for g in graphs if not g[2]] + [g for g in graphs if g[2]]
    ens, cand, codes = [], [], []
    remaining = len(order)
    for dots, fr, single in order:
        remaining -= 1
        graph_seed = seed + remaining
        ptr, st, en, ln, pre, sub, names, gnames = [0], [], [], [], [], [], [], []
        ref = strand = None
        for k, dot in enumerate(dots):
            res = f2_py.all_paths_oracle(dot, no_pre, max_cand)
            ps = f2_py.flatten(res)
            for i in range(ps["n_cand"]):
                st += ps["exon_start"][ps["exon_ptr"][i]:ps["exon_ptr"][i + 1]].tolist()
                en += ps["exon_end"][ps["exon_ptr"][i]:ps["exon_ptr"][i + 1]].tolist()
                ptr.append(len(st))
                gnames.append(res["graph_name"])
            ln += ps["cand_length"].tolist()
            pre += ps["is_premrna"].tolist()
            sub += [k] * ps["n_cand"]
            names += ps["names"]
            ref, strand = ref or res["reference"], strand or res["strand"]
        if not ln:
            codes.append(2)
            continue
        raw = f1_py.RawLocus(ptr, st, en, ln, pre, sub, strand == "+", fr.read_pos, fr.cigar_ptr, fr.cigar_op, fr.cigar_len, fr.frag_prob)
        cm = orc_f1.collapsed_map(raw, frag_mean, frag_sd, max_len=max(ln))
        if cm["return_code"] != 0:
            codes.append(cm["return_code"])
            continue
        rp, ci, va = [0], [], []
        for i in range(cm["R"]):
            nz = np.nonzero(cm["P"][i])[0]
            ci += nz.tolist()
            va += cm["P"][i, nz].tolist()
            rp.append(len(ci))
        loc = Locus(cm["T"], rp, ci, va, cm["counts"], cm["efflen"], int(n_pairs))
        burn = base + scale * int(cm["complexity"].max())
        o = oracle.run_sampler(loc, burn, 10 * burn, SITE if replay else FAST, lib.chain_key(graph_seed & 0xFFFFFFFF, 0))
        texts = []
        for mode in (0, 1):
            buf = C.create_string_buffer(1 << 22)
            i32 = lambda a: np.ascontiguousarray(a, np.int32)  # noqa: E731
            # a candidate keeps the graph name of ITS sub-graph: the writer is called per sub-graph name through the names
            n = facade.facade_gtf_from_accumulators(
                cm["T"], 10 * burn, i32(o["occ"]).ctypes.data_as(C.c_void_p), np.ascontiguousarray(o["mean"]).ctypes.data_as(C.c_void_p),
                np.ascontiguousarray(o["m2"]).ctypes.data_as(C.c_void_p), len(ln), i32(ptr).ctypes.data_as(C.c_void_p),
                i32(st).ctypes.data_as(C.c_void_p), i32(en).ctypes.data_as(C.c_void_p), i32(pre).ctypes.data_as(C.c_void_p),
                "\n".join(names).encode(), gnames[0].encode(), ref.encode(), strand.encode(), i32(cm["new_to_old"]).ctypes.data_as(C.c_void_p),
                np.ascontiguousarray(cm["efflen"]).ctypes.data_as(C.c_void_p), int(n_pairs), C.c_double(count_threshold),
                C.c_double(marginal_threshold), mode, buf, len(buf))
            assert n >= 0 or (mode == 0 and n in (-5, -7)), n
            texts.append((n, buf.value.decode() if n >= 0 else ""))
        code = 0 if texts[0][0] >= 0 else -texts[0][0]
        codes.append(code)
        if code == 0:
            ens.append(texts[0][1])
        cand.append(texts[1][1])
    return "".join(ens), "".join(cand), codes


@pytest.mark.gpu
def test_cli_run_matches_the_chained_oracles(cli, lib, oracle, facade, tmp_path):
    path, graphs, n_pairs = _make_run(tmp_path)
    single_graph = [g for g in graphs if len(g[0]) == 1]  # one graph name per locus keeps the expected writer call simple
    from bayesembler_b200 import synth
    synth.write_graph_file(path, single_graph, n_pairs, n_pairs + 1000)
    prefix = os.path.join(str(tmp_path), "res", "")
    rc, out, err = run(cli, "-g", path, "-o", prefix, "--seed", "11", "--frag-mean", "200", "--frag-sd", "20", "--output-mode", "full")
    assert rc == 0, err
    want_ens, want_cand, codes = _expected(lib, oracle, facade, single_graph, n_pairs, 11, 200.0, 20.0)
    assert codes.count(0) >= 5
    assert open(prefix + "assembly.gtf").read() == want_ens
    assert open(prefix + "candidates.gtf").read() == want_cand
    assert "%d graph(s) were assembled successfully" % codes.count(0) in out
    assert "%d graph(s) were not assembled due to insufficient candidate coverage" % codes.count(3) in out
    # the same run in the replay sampling mode (the reference's order of draws)
    rc, _, err = run(cli, "-g", path, "-o", prefix + "replay_", "--seed", "11", "--frag-mean", "200", "--frag-sd", "20", env={"BAYES_MODE": "replay"})
    assert rc == 0, err
    assert open(prefix + "replay_assembly.gtf").read() == _expected(lib, oracle, facade, single_graph, n_pairs, 11, 200.0, 20.0, replay=True)[0]
    # the same run with the fragment length distribution estimated from the single-path graphs (assembler.cpp:1919-1987)
    rc, out2, err = run(cli, "-g", path, "-o", prefix + "est_", "--seed", "11", "--output-mode", "full", "--num-threads", "4")
    assert rc == 0, err
    hist = {}
    for dots, fr, single in single_graph:
        if single:
            ps = f2_py.flatten(f2_py.all_paths_oracle(dots[0], True, 100))
            raw = f1_py.RawLocus(ps["exon_ptr"], ps["exon_start"], ps["exon_end"], ps["cand_length"], ps["is_premrna"], [0] * ps["n_cand"], True,
                                 fr.read_pos, fr.cigar_ptr, fr.cigar_op, fr.cigar_len, fr.frag_prob)
            for v in lib.fetch_fragment_lengths(raw).tolist():
                hist[v] = hist.get(v, 0) + 1
    lines = open(prefix + "est_fragment_lengths.txt").read().split("\n")
    assert lines[0] == "length\tcount" and lines[1:-1] == ["%d\t%d" % (k, hist[k]) for k in sorted(hist)]
    med, mad = f1_py.OracleF1().estimate_fragment_model(hist)
    want_ens2, _, _ = _expected(lib, oracle, facade, single_graph, n_pairs, 11, med, 1.4826 * mad)
    assert open(prefix + "est_assembly.gtf").read() == want_ens2


@pytest.mark.gpu
def test_cli_devices_shards_give_identical_files(cli, tmp_path):
    """--devices 0,1: graphs sharded over the GPUs by cost (bayes_partition_lpt), one batch per device; the files do not change."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    path, graphs, n_pairs = _make_run(tmp_path)
    out = {}
    for tag, extra in (("one", ("--device", "0")), ("two", ("--devices", "0,1"))):
        prefix = os.path.join(str(tmp_path), tag + "_")
        rc, text, err = run(cli, "-g", path, "-o", prefix, "--seed", "4", "--frag-mean", "200", "--frag-sd", "20", "--output-mode", "full", *extra)
        assert rc == 0, err
        out[tag] = [open(prefix + f).read() for f in ("assembly.gtf", "candidates.gtf", "graph_logs.txt")]
        if tag == "two":
            assert "sharded over 2 device(s)" in text
    assert out["one"] == out["two"] and len(out["one"][0]) > 0
