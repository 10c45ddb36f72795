"""BAM front end (csrc/bamfront.cpp) against the plain-Python restatement of the reference's
assembler.cpp:84-624, 627-1164, 1167-1417 (oracle/f4_py.py) on small synthetic BAMs.

PARITY UNPINNED for this row: BamTools, Boost, samtools and CEM processsam are absent, the reference's front end cannot be
built here and it ships no fixture; both sides are readings of the same source, written independently (C++ / Python).
Everything is host code: these tests need no GPU.
"""
import os
import subprocess

import numpy as np
import pytest

import bayesembler_b200 as bb
from oracle import f4_py


def _read_graph_file(path):
    graphs, head = [], None
    lines = open(path).read().split("\n")
    i = 0
    while i < len(lines):
        ln = lines[i]
        i += 1
        if not ln or ln.startswith("#"):
            continue
        w = ln.split()
        if w[0] == "fragments":
            head = (float(w[1]), float(w[2]))
            continue
        assert w[0] == "graph"
        n_dot, n_frag, single = int(w[1]), int(w[2]), int(w[3])
        dots = []
        for _ in range(n_dot):
            d = ""
            while True:
                d += lines[i] + "\n"
                i += 1
                if lines[i - 1] == "}":
                    break
            dots.append(d)
        frags = []
        for _ in range(n_frag):
            p1, c1, p2, c2, pr = lines[i].split()
            frags.append((int(p1), c1, int(p2), c2, float(pr)))
            i += 1
        graphs.append(dict(dot=dots, single=bool(single), frags=frags))
    return head, graphs


def _cigar_text(c):
    return "".join("%d%s" % (l, o) for o, l in c)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_sequencing_probability(seed):
    rng = np.random.default_rng(seed)
    for _ in range(200):
        n = int(rng.integers(20, 101))
        qual = "".join(chr(33 + int(q)) for q in rng.integers(2, 42, n))
        kind = rng.integers(4)
        if kind == 0:
            cigar, md = [("M", n)], str(n)
        elif kind == 1:
            a = int(rng.integers(1, n - 1))
            cigar, md = [("M", n)], "%dT%d" % (a, n - a - 1)
        elif kind == 2:
            a = int(rng.integers(2, n - 6))
            cigar, md = [("M", a), ("I", 3), ("M", n - a - 3)], str(n - 3)
        else:
            a = int(rng.integers(2, n - 2))
            cigar, md = [("M", a), ("D", 2), ("M", n - a)], "%d^AC%d" % (a, n - a)
        want = f4_py.sequencing_probability(md, qual, cigar)
        got = bb.sequencing_probability(md, qual, cigar)
        assert got == pytest.approx(want, rel=1e-13)
    # by hand: two matches at Q10 and Q20, one mismatch at Q30
    assert bb.sequencing_probability("2A0", "+5?", [("M", 3)]) == pytest.approx((1 - 0.1) * (1 - 0.01) * (0.001 / 3), rel=1e-13)
    with pytest.raises(bb.BayesError, match="disagree|past the end"):
        bb.sequencing_probability("10", "IIIII", [("M", 5)])


@pytest.mark.parametrize("strand_specific", ["unstranded", "first", "second"])
@pytest.mark.parametrize("seed", [11, 12])
def test_duplicate_removal_and_strand_split(tmp_path, seed, strand_specific):
    rng = np.random.default_rng(seed)
    refs, recs, _, _ = f4_py.synth_dataset(rng, n_loci=7, pairs_per_locus=80, stranded=strand_specific != "unstranded")
    bam = str(tmp_path / "in.bam")
    f4_py.write_bam(bam, refs, recs)
    _, parsed = f4_py.read_bam(bam)
    want_sets, want_counts = f4_py.dedup(parsed, strand_specific)
    front = bb.BamFront(bam, strand_specific)
    assert front.counts() == want_counts
    assert want_counts["unique_pairs"] > 0 and want_counts["pairs_written"] < want_counts["total_mapped_pairs"]  # duplicates were there
    for s, want in enumerate(want_sets):
        assert front.num_alignments(s) == len(want)
        sam = str(tmp_path / ("set%d.sam" % s))
        front.write_sam(sam, s)
        got = []
        for ln in open(sam):
            f = ln.rstrip("\n").split("\t")
            got.append(dict(name=f[0], flag=int(f[1]), ref=refs.index(next(r for r in refs if r[0] == f[2])), pos=int(f[3]) - 1, mate_pos=int(f[7]) - 1,
                            tlen=int(f[8]), cigar=f4_py.parse_cigar(f[5])))
            assert f[6] == "=" and len(f[9]) == len(f[10]) == 76 and any(t.startswith("NH:i:") for t in f[11:]) and any(t.startswith("MD:Z:") for t in f[11:])
        assert f4_py.canonical_sam(got) == f4_py.canonical_sam(want)
    front.close()


@pytest.mark.parametrize("no_pre_mrna", [False, True])
@pytest.mark.parametrize("seed", [21, 22, 23])
def test_instance_file_to_graph_file(tmp_path, seed, no_pre_mrna):
    rng = np.random.default_rng(seed)
    refs, recs, instance, loci = f4_py.synth_dataset(rng, n_loci=8, pairs_per_locus=70)
    bam, inst, out = str(tmp_path / "in.bam"), str(tmp_path / "in.instance"), str(tmp_path / "graphs.txt")
    f4_py.write_bam(bam, refs, recs)
    open(inst, "w").write(instance)
    front = bb.BamFront(bam, "unstranded")
    n = front.graphs(inst, out, strand=".", exon_base_coverage=0.05, no_pre_mrna=no_pre_mrna)
    front.close()
    _, parsed = f4_py.read_bam(bam)
    (written,), counts = f4_py.dedup(parsed, "unstranded")
    unsorted, excluded, cem = f4_py.parse_instance(instance, ".", 0.05, no_pre_mrna)
    graphs = f4_py.collapse(unsorted, no_pre_mrna)
    assert n["cem_graphs"] == cem == len(loci) and n["excluded"] == excluded == 1 and n["graphs"] == len(graphs)
    want = []
    for g in graphs:
        fr = f4_py.fetch(written, [r[0] for r in refs].index(g["reference"]), g)
        if fr:
            want.append((g, fr))
    head, got = _read_graph_file(out)
    assert head == (counts["unique_pairs"], counts["total_mapped_pairs"])
    assert n["with_fragments"] == len(got) == len(want) and len(want) >= 4
    for (g, fr), h in zip(want, got):
        assert [f4_py.canonical_dot(d) for d in h["dot"]] == [f4_py.canonical_dot(d) for d in g["dot"]]
        assert h["single"] == g["single"]
        a = sorted((p1, _cigar_text(c1), p2, _cigar_text(c2)) for p1, c1, p2, c2, _ in fr)
        b = sorted(x[:4] for x in h["frags"])
        assert a == b
        pa = dict(((p1, _cigar_text(c1), p2, _cigar_text(c2)), pr) for p1, c1, p2, c2, pr in fr)
        for x in h["frags"]:
            if len(fr) == len(pa):  # (two distinct pairs with equal positions and CIGARs differ in their NH / names only)
                assert x[4] == pytest.approx(pa[x[:4]], rel=1e-12)


def test_overlapping_graphs_are_collapsed(tmp_path):
    inst = "\n".join([
        "Instance\t1", "Boundary\tchr1\t100\t900\t+", "Reads\t5", "Segs\t2", "100\t300\t201\t0\t0\t0\t0\t0.1\t0", "600\t900\t301\t0\t0\t0\t0\t0.1\t0", "Refs\t0",
        "SGTypes\t1", "1 1 5", "PETypes\t0\t0",
        "Instance\t2", "Boundary\tchr1\t800\t1500\t+", "Reads\t7", "Segs\t2", "800\t1000\t201\t0\t0\t0\t0\t0.2\t0", "1300\t1500\t201\t0\t0\t0\t0\t0.97\t0", "Refs\t0",
        "SGTypes\t2", "1 1 3", "1 0 4", "PETypes\t0\t0",
        "Instance\t3", "Boundary\tchr1\t5000\t5400\t-", "Reads\t2", "Segs\t1", "5000\t5400\t401\t0\t0\t0\t0\t0.0\t0", "Refs\t0", "SGTypes\t1", "1 2", "PETypes\t0\t0"]) + "\n"
    unsorted, excluded, cem = f4_py.parse_instance(inst, ".", 0.05, False)
    graphs = f4_py.collapse(unsorted, False)
    assert cem == 3 and excluded == 0 and len(graphs) == 2
    merged = graphs[0]
    assert merged["left"] == 100 and merged["right"] == 1500 and len(merged["dot"]) == 2 and not merged["single"] and merged["read_count"] == 12
    # the second exon of instance 2 has 97 % of its bases uncovered: excluded, its junction with it
    assert '1 [label' not in merged["dot"][1] and "->" not in merged["dot"][1]
    assert merged["dot"][0] == ('digraph graph_1_plus {\n0 [label="chr1;+;100;300"]\n1 [label="chr1;+;600;900"]\npre_mrna [label="chr1;+;100;900"]\n'
                                '0->1 [coverage="5"]\n}\n')
    # the product, fed alignments that cover both graphs
    recs = []
    for i, (p1, p2) in enumerate([(120, 650), (150, 700), (820, 900), (5010, 5200)]):
        q = "I" * 50
        recs.append(dict(name="p%d" % i, flag=0x1 | 0x2 | 0x20 | 0x40, ref="chr1", pos=p1, mate_pos=p2, tlen=p2 + 50 - p1, cigar=[("M", 50)], qual=q,
                         tags=[("NH", "C", 1), ("MD", "Z", "50")]))
        recs.append(dict(name="p%d" % i, flag=0x1 | 0x2 | 0x10 | 0x80, ref="chr1", pos=p2, mate_pos=p1, tlen=-(p2 + 50 - p1), cigar=[("M", 50)], qual=q,
                         tags=[("NH", "C", 1), ("MD", "Z", "50")]))
    recs.sort(key=lambda r: r["pos"])
    bam, ip, out = str(tmp_path / "x.bam"), str(tmp_path / "x.instance"), str(tmp_path / "g.txt")
    f4_py.write_bam(bam, [("chr1", 10000)], recs)
    open(ip, "w").write(inst)
    front = bb.BamFront(bam)
    n = front.graphs(ip, out)
    front.close()
    head, got = _read_graph_file(out)
    assert n == dict(graphs=2, with_fragments=2, excluded=0, cem_graphs=3) and head == (4.0, 4.0)
    assert len(got[0]["dot"]) == 2 and got[0]["dot"][0] == merged["dot"][0] and len(got[0]["frags"]) == 3 and len(got[1]["frags"]) == 1
    assert got[0]["frags"][0][:4] == (120, "50M", 650, "50M") and got[0]["frags"][0][4] == pytest.approx((1 - 10 ** -4.0) ** 100, rel=1e-12)


def test_errors(tmp_path):
    with pytest.raises(bb.BayesError, match="cannot open"):
        bb.BamFront(str(tmp_path / "missing.bam"))
    junk = tmp_path / "junk.bam"
    junk.write_bytes(b"this is not a compressed file at all")
    with pytest.raises(bb.BayesError, match="BGZF|BAM"):
        bb.BamFront(str(junk))
    with pytest.raises(bb.BayesError, match="strand_specific"):
        bb.BamFront(str(junk), "sideways")
    rng = np.random.default_rng(5)
    refs, recs, instance, _ = f4_py.synth_dataset(rng, n_loci=2, pairs_per_locus=20)
    unsorted = str(tmp_path / "unsorted.bam")
    f4_py.write_bam(unsorted, refs, recs[::-1])
    with pytest.raises(bb.BayesError, match="sorted|mate"):
        bb.BamFront(unsorted)
    no_nh = [dict(r, tags=[t for t in r["tags"] if t[0] != "NH"]) for r in recs]
    p = str(tmp_path / "no_nh.bam")
    f4_py.write_bam(p, refs, no_nh)
    with pytest.raises(bb.BayesError, match="NH"):
        bb.BamFront(p)
    cut = tmp_path / "cut.bam"
    good = str(tmp_path / "good.bam")
    f4_py.write_bam(good, refs, recs)
    cut.write_bytes(open(good, "rb").read()[:400])
    with pytest.raises(bb.BayesError, match="truncated|inflate"):
        bb.BamFront(str(cut))
    front = bb.BamFront(good)
    with pytest.raises(bb.BayesError, match="cannot open instance"):
        front.graphs(str(tmp_path / "none.instance"), str(tmp_path / "g.txt"))
    bad = tmp_path / "bad.instance"
    bad.write_text(instance.replace("chr1", "chrZ"))
    with pytest.raises(bb.BayesError, match="reference sequence"):
        front.graphs(str(bad), str(tmp_path / "g.txt"))
    front.close()


def test_cli_refuses_a_bam_without_graphs(tmp_path):
    cli = os.path.join(os.path.dirname(bb.__file__), "bayesembler_b200_cli")
    if not os.path.exists(cli):
        from bayesembler_b200 import _build
        _build.build_cli()
    r = subprocess.run([cli, "--bam-file", "x.bam", "-o", str(tmp_path) + "/"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "processsam" in r.stderr and "--instance-file" in r.stderr
    r = subprocess.run([cli, "--bam-file", "x.bam", "-s", "first", "--instance-file-plus", "a", "-o", str(tmp_path) + "/"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "--instance-file-minus" in r.stderr


@pytest.mark.gpu
def test_cli_from_a_bam_file(tmp_path):
    """bayesembler_b200_cli --bam-file + --instance-file: the whole chain BAM -> duplicate removal -> graphs -> fragments -> candidates ->
    likelihoods -> sampler (GPU) -> assembly.gtf, equal to the run from the graph file the front end writes."""
    cli = os.path.join(os.path.dirname(bb.__file__), "bayesembler_b200_cli")
    rng = np.random.default_rng(99)
    refs, recs, instance, _ = f4_py.synth_dataset(rng, n_loci=8, pairs_per_locus=150)
    bam, inst = str(tmp_path / "lib.bam"), str(tmp_path / "lib.instance")
    f4_py.write_bam(bam, refs, recs)
    open(inst, "w").write(instance)
    common = ["--seed", "7", "--frag-mean", "220", "--frag-sd", "25", "--output-mode", "full"]
    a = subprocess.run([cli, "--bam-file", bam, "--instance-file", inst, "-o", str(tmp_path) + "/a_"] + common, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert a.returncode == 0, a.stderr
    assert "Removed duplicates from" in a.stdout and "collapsed them to" in a.stdout
    b = subprocess.run([cli, "--graph-file", str(tmp_path) + "/a_graphs.txt", "-o", str(tmp_path) + "/b_"] + common, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert b.returncode == 0, b.stderr
    gtf = open(str(tmp_path) + "/a_assembly.gtf").read()
    assert gtf == open(str(tmp_path) + "/b_assembly.gtf").read()
    assert gtf.count("\ttranscript\t") >= 3 or gtf.count("\texon\t") >= 6
    assert open(str(tmp_path) + "/a_candidates.gtf").read() == open(str(tmp_path) + "/b_candidates.gtf").read()


@pytest.mark.parametrize("strand_specific", ["unstranded", "first"])
def test_cli_front_end_writes_the_graph_file(tmp_path, strand_specific):
    """The command line's BAM stage runs on the host: whatever happens afterwards (here, on a box without a GPU, the sampler refuses
    to start), <prefix>graphs.txt holds the graphs of every strand set, equal to what the C ABI writes when called directly."""
    import torch
    cli = os.path.join(os.path.dirname(bb.__file__), "bayesembler_b200_cli")
    if not os.path.exists(cli):
        from bayesembler_b200 import _build
        _build.build_cli()
    rng = np.random.default_rng(31)
    stranded = strand_specific != "unstranded"
    refs, recs, instance, loci = f4_py.synth_dataset(rng, n_loci=6, pairs_per_locus=60, stranded=stranded)
    bam = str(tmp_path / "lib.bam")
    f4_py.write_bam(bam, refs, recs)
    # one instance file per strand set, as processsam -d + / -d - would write them
    blocks = instance.strip("\n").split("Instance\t")[1:]
    args = ["--bam-file", bam, "-s", strand_specific, "-o", str(tmp_path) + "/o_", "--frag-mean", "220", "--frag-sd", "25", "--seed", "3"]
    front = bb.BamFront(bam, strand_specific)
    want = str(tmp_path / "want.txt")
    if stranded:
        files = {}
        for sign, name in (("+", "plus"), ("-", "minus")):
            files[name] = str(tmp_path / (name + ".instance"))
            open(files[name], "w").write("".join("Instance\t" + b for b in blocks if "\t%s\n" % sign in b.split("\n")[1] + "\n"))
            args += ["--instance-file-" + name, files[name]]
        front.graphs(files["plus"], want, strand="+", strand_file=0)
        front.graphs(files["minus"], want, strand="-", strand_file=1, append=True)
    else:
        inst = str(tmp_path / "lib.instance")
        open(inst, "w").write(instance)
        args += ["--instance-file", inst]
        front.graphs(inst, want, strand=".")
    front.close()
    r = subprocess.run([cli] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert "Removed duplicates from" in r.stdout and "collapsed them to" in r.stdout, (r.stdout, r.stderr)
    if not torch.cuda.is_available():
        assert r.returncode == 1 and "no CUDA device" in r.stderr
    else:
        assert r.returncode == 0, r.stderr
    got = open(str(tmp_path) + "/o_graphs.txt").read()
    assert got == open(want).read() and got.count("\ngraph ") >= 3
