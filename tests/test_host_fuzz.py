"""Malformed input never takes the process down: the reference aborts through `assert` (allPaths.cpp:73-79,
alignmentParser.cpp:61-118); the library must come back with BAYES_E_INVALID and a message, or with a valid answer.
Seeded mutation fuzzing of the two host stages that parse caller-supplied structure (rows f1 and f2)."""
import numpy as np

from oracle import f1_py, f2_py


def test_fuzz_dot_text(lib):
    rng = np.random.default_rng(0)
    alphabet = list('0123456789->[]{}";=,abcdefghijklmnopqrstuvwxyz_ \n')
    outcomes = {"ok": 0, "invalid": 0}
    for _ in range(1500):
        s = list(f2_py.random_splice_graph(int(rng.integers(0, 50)), int(rng.integers(2, 9))))
        for _ in range(int(rng.integers(1, 6))):
            op, i = int(rng.integers(0, 3)), int(rng.integers(0, len(s)))
            if op == 0 and len(s) > 1:
                del s[i]
            elif op == 1:
                s.insert(i, alphabet[int(rng.integers(0, len(alphabet)))])
            else:
                s[i] = alphabet[int(rng.integers(0, len(alphabet)))]
        try:
            got = lib.all_paths("".join(s), bool(rng.integers(0, 2)), int(rng.integers(0, 50)))
            assert got["n_cand"] == len(got["names"]) == len(got["cand_length"])
            outcomes["ok"] += 1
        except lib.BayesError as e:
            assert "bayes_all_paths_build" in str(e)
            outcomes["invalid"] += 1
    assert outcomes["ok"] > 50 and outcomes["invalid"] > 500, outcomes


def test_fuzz_raw_locus(lib):
    rng = np.random.default_rng(1)
    seen = set()
    for _ in range(400):
        raw = f1_py.random_raw_locus(rng, n_exons=int(rng.integers(2, 7)), n_cand=int(rng.integers(1, 6)), n_frag=int(rng.integers(1, 60)))
        k = int(rng.integers(0, 6))
        if k == 0:
            raw.read_pos[rng.integers(0, len(raw.read_pos))] = rng.integers(0, 2 ** 31)
        elif k == 1:
            raw.cigar_len[rng.integers(0, len(raw.cigar_len))] = rng.integers(0, 100000)
        elif k == 2:
            raw.cigar_op[rng.integers(0, len(raw.cigar_op))] = int(rng.choice([ord(x) for x in "MNIDSHX=P*"]))
        elif k == 3:
            raw.exon_end[rng.integers(0, len(raw.exon_end))] -= int(rng.integers(0, 5000))
        elif k == 4:
            raw.frag_prob[rng.integers(0, len(raw.frag_prob))] = float(rng.choice([0.0, -1.0, np.nan, np.inf, 1e-320]))
        else:
            raw.cand_length[rng.integers(0, len(raw.cand_length))] = int(rng.integers(-5, 10))
        try:
            model = lib.SequencingModel(200.0, 20.0, max(1, int(raw.cand_length.max())))
            got = lib.build_collapsed_map(raw, model)
            assert got["return_code"] in (0, 3, 4)
            if got["return_code"] == 0:
                assert got["counts"].min() >= 1 and np.all(np.isfinite(got["val"])) and got["row_ptr"][-1] == len(got["val"])
            seen.add(got["return_code"])
        except lib.BayesError as e:
            seen.add("invalid")
            assert "bayes_" in str(e)
    assert 0 in seen and "invalid" in seen


def test_fuzz_bam_and_instance_files(tmp_path):
    """Row f4: byte mutations of a BAM file's decompressed stream (re-compressed into valid BGZF, so the record decoder sees them) and
    line mutations of a CEM instance file.  The reference asserts all over this stage (assembler.cpp:329-352, :700-1010); the library
    answers BAYES_E_INVALID / BAYES_E_NUMERIC or a valid result, and never takes the process down."""
    import gzip
    import bayesembler_b200 as bb
    from oracle import f4_py
    rng = np.random.default_rng(2)
    refs, recs, instance, _ = f4_py.synth_dataset(rng, n_loci=3, pairs_per_locus=25)
    good = str(tmp_path / "good.bam")
    f4_py.write_bam(good, refs, recs)
    raw = bytearray(gzip.open(good, "rb").read())
    inst_path = str(tmp_path / "good.instance")
    open(inst_path, "w").write(instance)
    outcomes = {"ok": 0, "invalid": 0}
    for trial in range(300):
        b = bytearray(raw)
        for _ in range(int(rng.integers(1, 4))):
            k, i = int(rng.integers(0, 4)), int(rng.integers(0, len(b)))
            if k == 0:
                b[i] = int(rng.integers(0, 256))
            elif k == 1:
                b[i:i + 4] = int(rng.choice([0, 1, 0x7FFFFFFF, 0xFFFFFFFF, 0x80000000, 65536])).to_bytes(4, "little")
            elif k == 2:
                del b[i:i + int(rng.integers(1, 40))]
            else:
                b = b[:max(8, i)]
        p = str(tmp_path / "m.bam")
        with open(p, "wb") as f:
            for j in range(0, len(b), 60000):
                f.write(f4_py._bgzf_block(bytes(b[j:j + 60000])))
            f.write(f4_py._bgzf_block(b""))
        try:
            front = bb.BamFront(p, str(rng.choice(["unstranded", "first", "second"])))
            c = front.counts()
            assert c["unique_pairs"] <= c["total_mapped_pairs"] + 1
            try:
                front.graphs(inst_path, str(tmp_path / "g.txt"))
            except bb.BayesError as e:
                assert "bayes_bam_front_graphs" in str(e)
            front.close()
            outcomes["ok"] += 1
        except bb.BayesError as e:
            assert "bayes_bam_front_open" in str(e)
            outcomes["invalid"] += 1
    assert outcomes["ok"] > 5 and outcomes["invalid"] > 100, outcomes
    front = bb.BamFront(good)
    lines = instance.split("\n")
    alphabet = list("0123456789\t +-.ISRBPT")
    seen = {"ok": 0, "invalid": 0}
    for trial in range(300):
        ls = list(lines)
        for _ in range(int(rng.integers(1, 4))):
            k, i = int(rng.integers(0, 4)), int(rng.integers(0, len(ls)))
            if k == 0:
                del ls[i]
            elif k == 1:
                ls.insert(i, ls[int(rng.integers(0, len(ls)))])
            elif k == 2 and ls[i]:
                c = list(ls[i])
                c[int(rng.integers(0, len(c)))] = alphabet[int(rng.integers(0, len(alphabet)))]
                ls[i] = "".join(c)
            else:
                ls[i] = ls[i][:int(rng.integers(0, len(ls[i]) + 1))]
        ip = str(tmp_path / "m.instance")
        open(ip, "w").write("\n".join(ls))
        try:
            n = front.graphs(ip, str(tmp_path / "g2.txt"))
            assert n["graphs"] >= n["with_fragments"] >= 0
            seen["ok"] += 1
        except bb.BayesError as e:
            assert "bayes_bam_front_graphs" in str(e)
            seen["invalid"] += 1
    front.close()
    assert seen["ok"] > 20 and seen["invalid"] > 20, seen
