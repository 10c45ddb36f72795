"""TEST INFRASTRUCTURE -- plain-Python restatement of the reference's BAM front end and a writer of
small synthetic BAM / CEM-instance fixtures.  Only tests/ may import this module; nothing in the product does.

PARITY UNPINNED: the reference's front end (assembler.cpp:84-624, 627-1164, 1167-1417) needs BamTools, Boost, samtools and
CEM processsam, none of which exist here, and the reference ships no BAM fixture or golden vector for it.  What this module
pins is the product (csrc/bamfront.cpp) against an independent, line-by-line reading of the reference's source:

    dedup()                  generateSpliceGraphs :329-380 / :400-520, markDuplicates :157-249, add/writeUniqueReads :84-135
    parse_instance()         graphConstructorCallback :692-1010
    collapse()               :1012-1125
    fetch()                  grapherCallback :1167-1319 (region overlap as in BamTools 2.x: starts before `right`, and starts or ends
                             at / after `left`)
    sequencing_probability() calculateSequencingProbability :1322-1417

Two orders in the reference are those of hash containers (pairs completed at the same position, :128-137; junction lines of a
DOT string, :872-875).  The product reproduces them with the reference's own container types; this restatement does not model a
hash table, so the tests compare those two groups as sets (canonical()).
"""
import gzip
import re
import struct
import zlib

OPS = "MIDNSHP=X"


# ---------------------------------------------------------------------------------------------------------- BAM in / out
def _bgzf_block(data):
    comp = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = comp.compress(data) + comp.flush()
    bsize = len(body) + 25
    return (struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize) + body +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def parse_cigar(text):
    return [(m.group(2), int(m.group(1))) for m in re.finditer(r"(\d+)([MIDNSHP=X])", text)]


def encode_record(r, ref_index):
    name = r["name"].encode() + b"\0"
    cigar = r["cigar"] if not isinstance(r["cigar"], str) else parse_cigar(r["cigar"])
    seq = r.get("seq", "A" * len(r.get("qual", "")))
    qual = r.get("qual", "")
    codes = "=ACMGRSVTWYHKDBN"
    packed = bytearray((len(seq) + 1) // 2)
    for i, c in enumerate(seq):
        packed[i >> 1] |= codes.index(c) << (4 if i % 2 == 0 else 0)
    tags = b""
    for key, typ, val in r.get("tags", []):
        tags += key.encode() + typ.encode()
        if typ == "Z":
            tags += val.encode() + b"\0"
        elif typ == "A":
            tags += val.encode()
        else:
            tags += struct.pack("<" + {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}[typ], val)
    body = struct.pack("<iiBBHHHiiii", ref_index[r["ref"]], r["pos"], len(name), r.get("mapq", 50), 4680, len(cigar), r["flag"], len(seq),
                       ref_index[r.get("mate_ref", r["ref"])], r["mate_pos"], r["tlen"])
    body += name + b"".join(struct.pack("<I", (l << 4) | OPS.index(o)) for o, l in cigar) + bytes(packed)
    body += bytes((ord(c) - 33) for c in qual) + tags
    return struct.pack("<i", len(body)) + body


def write_bam(path, refs, records, header_text=None):
    """refs = [(name, length)]; records = dicts (name, flag, ref, pos, mate_pos, tlen, cigar, qual, tags) in file order."""
    ref_index = {n: i for i, (n, _) in enumerate(refs)}
    text = (header_text if header_text is not None else
            "@HD\tVN:1.0\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % (n, l) for n, l in refs)).encode()
    parts = [b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(refs))]
    for n, l in refs:
        parts.append(struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", l))
    for r in records:
        parts.append(encode_record(r, ref_index))
    raw = b"".join(parts)
    with open(path, "wb") as f:
        for i in range(0, len(raw), 60000):
            f.write(_bgzf_block(raw[i:i + 60000]))
        f.write(_bgzf_block(b""))  # the BGZF end-of-file marker


def read_bam(path):
    raw = gzip.open(path, "rb").read()
    assert raw[:4] == b"BAM\1"
    l_text, = struct.unpack_from("<i", raw, 4)
    at = 8 + l_text
    n_ref, = struct.unpack_from("<i", raw, at)
    at += 4
    refs = []
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<i", raw, at)
        refs.append(raw[at + 4:at + 4 + l_name - 1].decode())
        at += 8 + l_name
    recs = []
    while at < len(raw):
        block, = struct.unpack_from("<i", raw, at)
        b = raw[at + 4:at + 4 + block]
        at += 4 + block
        ref, pos, l_name, mapq, _bin, n_cigar, flag, l_seq, mref, mpos, tlen = struct.unpack_from("<iiBBHHHiiii", b, 0)
        p = 32
        name = b[p:p + l_name - 1].decode()
        p += l_name
        cigar = []
        for _ in range(n_cigar):
            v, = struct.unpack_from("<I", b, p)
            cigar.append((OPS[v & 15], v >> 4))
            p += 4
        p += (l_seq + 1) // 2
        qual = "".join(chr(c + 33) for c in b[p:p + l_seq])
        p += l_seq
        tags = {}
        while p < len(b):
            key, typ = b[p:p + 2].decode(), chr(b[p + 2])
            p += 3
            if typ in "ZH":
                e = b.index(b"\0", p)
                tags[key] = b[p:e].decode()
                p = e + 1
            elif typ == "A":
                tags[key] = chr(b[p])
                p += 1
            else:
                fmt = {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}[typ]
                tags[key], = struct.unpack_from("<" + fmt, b, p)
                p += struct.calcsize(fmt)
        recs.append(dict(name=name, flag=flag, ref=ref, pos=pos, mate_ref=mref, mate_pos=mpos, tlen=tlen, cigar=cigar, qual=qual, tags=tags,
                         mapq=mapq))
    return refs, recs


# ---------------------------------------------------------------------------------------------- duplicate removal, strands
def _dedup_cigar(cigar):  # generateCigarString :138-154
    out = ""
    for o, l in cigar:
        if o in "MD":
            out += "%dM" % l
        elif o == "N":
            out += "%dN" % l
        elif o != "I":
            raise ValueError("Unhandled cigar string symbol")
    return out


class _Strand(object):
    def __init__(self):
        self.first_reads = {}   # position -> {(name, hi): alignment}
        self.read_pairs = {}    # (pos, insert, cigar, mate cigar) -> (first, second)
        self.unique = []        # (position, insertion order, alignment)
        self.seq = 0
        self.cur_pos, self.cur_ref = -1, -1
        self.written, self.nd_nm, self.paired = [], 0, 0

    def add_unique(self):  # :124-135
        for first, second in self.read_pairs.values():
            for a in (first, second):
                self.unique.append((a["pos"], self.seq, a))
                self.seq += 1
        self.read_pairs = {}

    def write_unique(self):  # :84-121
        left_most = min(self.first_reads) if self.first_reads else 1 << 32
        self.unique.sort(key=lambda t: (t[0], t[1]))
        keep = []
        for t in self.unique:
            if t[0] < left_most:
                if t[2]["tags"]["NH"] == 1:
                    self.nd_nm += 1
                self.written.append(t[2])
            else:
                keep.append(t)
        self.unique = keep

    def mark(self, cur):  # :157-249
        rid = (cur["name"], cur["tags"].get("HI", 0))
        at_mate = self.first_reads.get(cur["mate_pos"])
        if at_mate is None:
            self.first_reads.setdefault(cur["pos"], {})[rid] = cur
            return
        first = at_mate.get(rid)
        if first is None:
            assert cur["pos"] == cur["mate_pos"]
            at_mate[rid] = cur
            return
        pi = (first["pos"], abs(first["tlen"]), _dedup_cigar(first["cigar"]), _dedup_cigar(cur["cigar"]))
        prev = self.read_pairs.get(pi)
        if prev is None:
            self.read_pairs[pi] = (first, cur)
        elif first["tags"]["NH"] == 1 and prev[0]["tags"]["NH"] > 1:
            self.read_pairs[pi] = (first, cur)
        del at_mate[rid]
        if not at_mate:
            del self.first_reads[cur["mate_pos"]]

    def take(self, cur):  # :343-368
        if self.cur_pos != cur["pos"] or self.cur_ref != cur["ref"]:
            self.paired += len(self.read_pairs)
            self.add_unique()
            self.write_unique()
            if self.cur_ref != cur["ref"]:
                self.cur_ref = cur["ref"]
                assert not self.unique and not self.first_reads
            else:
                assert self.cur_pos < cur["pos"]
        self.mark(cur)
        self.cur_pos = cur["pos"]

    def finish(self):
        self.paired += len(self.read_pairs)
        self.add_unique()
        self.write_unique()
        assert not self.unique and not self.first_reads


def dedup(recs, strand_specific="unstranded"):
    """-> (strand sets [unstranded] or [plus, minus] of written alignments, counters)"""
    plus, minus = _Strand(), _Strand()
    w_first, w_second = (plus, minus) if strand_specific == "first" else (minus, plus)
    total = 0
    for a in recs:
        f = a["flag"]
        assert f & 1 and not f & 0x200
        if f & 0x4 or f & 0x100 or f & 0x8:
            continue
        total += 1
        rev, mrev, first = bool(f & 0x10), bool(f & 0x20), bool(f & 0x40)
        if not (rev != mrev and a["ref"] == a["mate_ref"]):
            continue
        if abs(a["tlen"]) > 700000:
            continue
        if strand_specific == "unstranded":
            plus.take(a)
        elif (first and rev and a["pos"] >= a["mate_pos"]) or (not first and not rev and a["pos"] <= a["mate_pos"]):
            w_first.take(a)
        elif (first and not rev and a["pos"] <= a["mate_pos"]) or (not first and rev and a["pos"] >= a["mate_pos"]):
            w_second.take(a)
    sets = [plus] if strand_specific == "unstranded" else [plus, minus]
    for s in sets:
        s.finish()
    counts = dict(total_mapped_pairs=float(total // 2), unique_pairs=float(sum(s.nd_nm for s in sets) // 2),
                  pairs_written=float(sum(s.paired for s in sets)))
    return [s.written for s in sets], counts


# --------------------------------------------------------------------------------------------------------- read probability
def sequencing_probability(md, qual, cigar):  # :1322-1417
    probability = 1.0
    deletions, insertions, running = [], [], 0
    for o, l in cigar:
        if o == "M":
            running += l
        elif o == "D":
            deletions.append(running)
        elif o == "I":
            for _ in range(l):
                insertions.append(running)
                running += 1
    assert running == len(qual)
    deletions.append(len(qual) + 1)
    insertions.append(len(qual) + 1)
    d_it = i_it = q = 0
    numbers = [int(x) for x in re.findall(r"\d+", md)]
    for k, num in enumerate(numbers):
        for _ in range(num):
            while insertions[i_it] == q:
                i_it += 1
                q += 1
            probability *= (1 - pow(10, -(ord(qual[q]) - 33) / 10.0))
            q += 1
        while insertions[i_it] == q:
            i_it += 1
            q += 1
        if q < len(qual):
            if deletions[d_it] == q:
                d_it += 1
            else:
                probability *= pow(10, -(ord(qual[q]) - 33) / 10.0) / 3
                q += 1
        else:
            assert k == len(numbers) - 1
    assert insertions[i_it] == len(qual) + 1 and deletions[d_it] == len(qual) + 1 and q == len(qual)
    return probability


# ------------------------------------------------------------------------------------------------------------ instance file
def parse_instance(text, strand=".", exon_base_coverage=0.05, no_pre_mrna=False):
    """-> (per-reference lists of graphs, instances dropped for lack of a strand, instances).  A graph is a dict with dot (list of
    strings), reference, strand, left, right, single, read_count; junction lines of a DOT string are sorted (see module docstring)."""
    pats = dict(instance=r"^Instance\s+(\d+)$", boundary=r"^Boundary\s+([\w|-]+)\s+(\d+)\s+(\d+)\s+([\+|\.|-])$", segs=r"^Segs\s+(\d+)$",
                refs=r"^Refs\s+(\d+)$", sg=r"^SGTypes\s+(\d+)$", pe=r"^PETypes\s+(\d+)\s+\d+$", reads=r"^Reads\s+(\d+)$")
    exon_on = exon_off = path_on = path_off = False
    unsorted, temp = [], []
    graph_count = excluded_count = 0
    left = right = ref_exons = ref_paths = real_exon = real_path = 0
    reads = 0
    ref_id, last_ref, ref_strand = "", "", ""
    dot, junctions, excluded, vin, vout, single = "", {}, set(), {}, {}, True
    for line in text.split("\n"):
        m = re.match(pats["instance"], line)
        if m:
            graph_count += 1
            assert dot == "" and not junctions
            dot = "digraph graph_" + m.group(1)
            single = True
        m = re.match(pats["boundary"], line)
        if m:
            ref_id, ref_strand = m.group(1), m.group(4)
            dot += {"+": "_plus {\n", "-": "_minus {\n"}.get(ref_strand, "_dot {\n")
            left, right = int(m.group(2)), int(m.group(3))
        m = re.match(pats["reads"], line)
        if m:
            reads = int(m.group(1))
        if re.match(pats["refs"], line):
            exon_on, exon_off = False, True
            assert ref_exons == real_exon
        if exon_on and not exon_off:
            f = line.split("\t")
            if exon_base_coverage <= (1 - float(f[7])):
                left, right = min(left, int(f[0])), max(right, int(f[1]))
                dot += '%d [label="%s;%s;%s;%s"]\n' % (real_exon, ref_id, ref_strand, f[0], f[1])
            else:
                excluded.add(real_exon)
            real_exon += 1
        m = re.match(pats["segs"], line)
        if m:
            exon_on, exon_off, real_exon, ref_exons = True, False, 0, int(m.group(1))
        if re.match(pats["pe"], line):
            path_on, path_off = False, True
            assert ref_paths == real_path
            if not no_pre_mrna:
                dot += 'pre_mrna [label="%s;%s;%d;%d"]\n' % (ref_id, ref_strand, left, right)
            for j in sorted(junctions):
                dot += '%s [coverage="%d"]\n' % (j, junctions[j])
            dot += "}\n"
            if ref_id != last_ref:
                if temp:
                    unsorted.append(temp)
                    temp = []
                last_ref = ref_id
            if ref_strand in "+-":
                temp.append(dict(dot=[dot], reference=ref_id, strand=ref_strand, left=left, right=right, single=single, read_count=reads))
            else:
                assert strand == "."
                excluded_count += 1
            dot, junctions, excluded, vin, vout = "", {}, set(), {}, {}
        if path_on and not path_off:
            f = line.split(" ")
            assert len(f) == real_exon + 1
            first = -1
            for i in range(real_exon):
                if i in excluded:
                    continue
                if f[i][:1] == "1":
                    first = i
                    break
            last = first
            for i in range(first + 1, real_exon):
                if i in excluded:
                    continue
                if f[i][:1] == "1":
                    key = "%d->%d" % (last, i)
                    if single:
                        s = vin.setdefault(i, set())
                        if last not in s:
                            s.add(last)
                            if len(s) > 1:
                                single = False
                        s = vout.setdefault(last, set())
                        if i not in s:
                            s.add(i)
                            if len(s) > 1:
                                single = False
                    last = i
                    junctions[key] = junctions.get(key, 0) + int(f[real_exon])
            real_path += 1
        m = re.match(pats["sg"], line)
        if m:
            path_on, path_off, real_path, ref_paths = True, False, 0, int(m.group(1))
    if temp:
        unsorted.append(temp)
    return unsorted, excluded_count, graph_count


def collapse(unsorted, no_pre_mrna=False):  # :1012-1125
    out = []
    if no_pre_mrna:
        for ref in unsorted:
            out.extend(ref)
        return out
    last, first = {}, {"+": True, "-": True}
    for ref in unsorted:
        for g in sorted(ref, key=lambda g: g["left"]):  # list::sort is stable, so is sorted()
            s = g["strand"]
            if not first[s] and g["reference"] == last[s]["reference"] and g["left"] <= last[s]["right"]:
                last[s]["right"] = max(g["right"], last[s]["right"])
                last[s]["single"] = False
                last[s]["dot"] = last[s]["dot"] + [g["dot"][0]]
                last[s]["read_count"] += g["read_count"]
            else:
                if not first[s]:
                    out.append(last[s])
                first[s] = False
                last[s] = dict(g)
    for s in "+-":
        if not first[s]:
            out.append(last[s])
    return out


# ------------------------------------------------------------------------------------------------------------ region fetch
def _end(a):
    return a["pos"] + sum(l for o, l in a["cigar"] if o in "MDN=X")


def fetch(written, ref_index, g):  # :1167-1319; written = one strand set in file order, ref_index = BAM index of g's reference
    single, out = {}, []
    for a in written:
        if a["ref"] != ref_index or not (a["pos"] < g["right"] and (a["pos"] >= g["left"] or _end(a) >= g["left"])):
            continue
        if a["flag"] & 0x8 or a["tags"]["NH"] > 1:
            continue
        if a["mate_pos"] + 1 > g["right"] or a["mate_pos"] + 1 < g["left"]:
            continue
        p = sequencing_probability(a["tags"]["MD"], a["qual"], a["cigar"])
        mate = single.pop(a["name"], None)
        if mate is not None:
            assert mate["pos"] <= a["pos"] and mate["pos"] == a["mate_pos"]
            out.append((mate["pos"], mate["cigar"], a["pos"], a["cigar"], mate["p"] * p))
        else:
            single[a["name"]] = dict(pos=a["pos"], cigar=a["cigar"], p=p)
    return out


# --------------------------------------------------------------------------------------------------------- synthetic data
def synth_dataset(rng, n_loci=6, pairs_per_locus=60, read_len=76, refs=(("chr1", 400000), ("chrX_random-2", 250000)), stranded=False):
    """Small paired-end library over a few multi-exon loci: (refs, coordinate-sorted records, CEM-instance text, loci).
    Reads are 2 x read_len with M / N CIGARs (a few with I or D), MD tags with mismatches, varying base qualities, exact
    duplicates of some pairs, some multi-mapping pairs (NH = 2, HI), a few alignments the reference filters out."""
    recs, loci, instance = [], [], []
    serial = [0]

    def place(ref, exons, start_off, length):
        """CIGAR of `length` transcript bases starting start_off into the exon chain -> (pos, cigar)"""
        cigar, pos, left = [], None, length
        off = start_off
        for k, (s, e) in enumerate(exons):
            el = e - s + 1
            if off >= el:
                off -= el
                continue
            if pos is None:
                pos = s + off - 1  # 0-based
            take = min(left, el - off)
            cigar.append(("M", take))
            left -= take
            off = 0
            if left == 0:
                break
            cigar.append(("N", exons[k + 1][0] - e - 1))
        return pos, cigar

    inst = 0
    for li in range(n_loci):
        ref, ref_len = refs[li % len(refs)]
        base = 5000 + 30000 * (li // len(refs)) + (0 if li % 3 else 2000)
        n_ex = int(rng.integers(2, 6))
        exons, at = [], base
        for _ in range(n_ex):
            l = int(rng.integers(150, 500))
            exons.append((at, at + l - 1))
            at += l + int(rng.integers(80, 600))
        strand = "+" if (li % 2 == 0 or not stranded) and li % 4 != 3 else "-"
        if not stranded and li == n_loci - 1:
            strand = "."
        # isoforms: all exons, and (if possible) one skipping an inner exon
        paths = [list(range(n_ex))]
        if n_ex >= 3:
            skip = int(rng.integers(1, n_ex - 1))
            paths.append([i for i in range(n_ex) if i != skip])
        counts = [0] * len(paths)
        for _ in range(pairs_per_locus):
            pi = int(rng.integers(len(paths)))
            chain = [exons[i] for i in paths[pi]]
            tlen_tx = sum(e - s + 1 for s, e in chain)
            frag = int(min(tlen_tx, max(2 * read_len - 20, rng.normal(220, 25))))
            if frag < read_len:
                continue
            off = int(rng.integers(0, tlen_tx - frag + 1))
            p1, c1 = place(ref, chain, off, read_len)
            p2, c2 = place(ref, chain, off + frag - read_len, read_len)
            counts[pi] += 1
            serial[0] += 1
            name = "r%05d" % serial[0]
            nh = 2 if rng.random() < 0.08 else 1

            def md_and_cigar(c):
                c = list(c)
                md, qual = "", "".join(chr(33 + int(q)) for q in rng.integers(12, 41, read_len))
                if len(c) == 1 and rng.random() < 0.15:  # an insertion or a deletion inside an unspliced read
                    a = int(rng.integers(10, read_len - 20))
                    if rng.random() < 0.5:
                        c = [("M", a), ("I", 2), ("M", read_len - a - 2)]
                        md = "%d" % (read_len - 2)
                    else:
                        c = [("M", a), ("D", 1), ("M", read_len - a)]
                        md = "%d^A%d" % (a, read_len - a)
                elif rng.random() < 0.3:
                    a = int(rng.integers(1, read_len - 1))
                    md = "%dC%d" % (a, read_len - a - 1)
                else:
                    md = "%d" % read_len
                return c, md, qual

            c1, md1, q1 = md_and_cigar(c1)
            c2, md2, q2 = md_and_cigar(c2)
            end2 = p2 + sum(l for o, l in c2 if o in "MDN")
            tl = end2 - p1
            minus = strand == "-"
            # dUTP-like ("first"): first mate reverse on the plus strand
            f1 = 0x1 | 0x2 | 0x20 | (0x80 if not minus else 0x40)
            f2 = 0x1 | 0x2 | 0x10 | (0x40 if not minus else 0x80)
            copies = 1 + (1 if rng.random() < 0.2 else 0) + (1 if rng.random() < 0.05 else 0)
            for cp in range(copies):
                nm = name if cp == 0 else "%s_dup%d" % (name, cp)
                nh_c = nh if cp == 0 else (1 if rng.random() < 0.5 else 2)
                tags1 = [("NH", "C", nh_c), ("MD", "Z", md1)] + ([("HI", "C", 0)] if nh_c > 1 else [])
                tags2 = [("NH", "C", nh_c), ("MD", "Z", md2)] + ([("HI", "C", 0)] if nh_c > 1 else [])
                recs.append(dict(name=nm, flag=f1, ref=ref, pos=p1, mate_pos=p2, tlen=tl, cigar=c1, qual=q1, tags=tags1))
                recs.append(dict(name=nm, flag=f2, ref=ref, pos=p2, mate_pos=p1, tlen=-tl, cigar=c2, qual=q2, tags=tags2))
        # alignments the reference skips: mate unmapped, secondary, both mates on one strand
        recs.append(dict(name="u%d" % li, flag=0x1 | 0x8 | 0x40, ref=ref, pos=base + 5, mate_pos=base + 5, tlen=0, cigar=[("M", read_len)],
                         qual="I" * read_len, tags=[("NH", "C", 1), ("MD", "Z", str(read_len))]))
        recs.append(dict(name="s%d" % li, flag=0x1 | 0x2 | 0x100 | 0x20 | 0x40, ref=ref, pos=base + 9, mate_pos=base + 200, tlen=276, cigar=[("M", read_len)],
                         qual="I" * read_len, tags=[("NH", "C", 3), ("HI", "C", 1), ("MD", "Z", str(read_len))]))
        recs.append(dict(name="o%d" % li, flag=0x1 | 0x2 | 0x40, ref=ref, pos=base + 11, mate_pos=base + 210, tlen=280, cigar=[("M", read_len)],
                         qual="I" * read_len, tags=[("NH", "C", 1), ("MD", "Z", str(read_len))]))
        recs.append(dict(name="o%d" % li, flag=0x1 | 0x2 | 0x80, ref=ref, pos=base + 210, mate_pos=base + 11, tlen=-280, cigar=[("M", read_len)],
                         qual="I" * read_len, tags=[("NH", "C", 1), ("MD", "Z", str(read_len))]))
        inst += 1
        lines = ["Instance\t%d" % inst, "Boundary\t%s\t%d\t%d\t%s" % (ref, exons[0][0], exons[-1][1], strand), "ReadLen\t%d" % read_len,
                 "Reads\t%d" % sum(counts), "Segs\t%d" % n_ex]
        for k, (s, e) in enumerate(exons):
            cov = 0.99 if (k == n_ex - 1 and li % 5 == 4) else float(rng.uniform(0.0, 0.6))  # field 8: share of the segment's bases WITHOUT coverage
            lines.append("%d\t%d\t%d\t0\t0\t0\t0\t%.4f\t0" % (s, e, e - s + 1, cov))
        lines += ["Refs\t0", "Reads\t%d" % sum(counts), "SGTypes\t%d" % len(paths)]
        for p, c in zip(paths, counts):
            lines.append(" ".join("1" if i in p else "0" for i in range(n_ex)) + " %d" % c)
        lines += ["PETypes\t0\t0"]
        instance += lines
        loci.append(dict(ref=ref, exons=exons, strand=strand, paths=paths))
    ref_index = {n: i for i, (n, _) in enumerate(refs)}
    order = sorted(range(len(recs)), key=lambda i: (ref_index[recs[i]["ref"]], recs[i]["pos"], i))
    return list(refs), [recs[i] for i in order], "\n".join(instance) + "\n", loci


# ------------------------------------------------------------------------------------------------------- canonical forms
def canonical_sam(records):
    """alignments as comparable tuples; groups with equal (reference, position) are sorted (hash order in the reference)"""
    key = [(r["ref"], r["pos"], r["name"], r["flag"], r["mate_pos"], r["tlen"], "".join("%d%s" % (l, o) for o, l in r["cigar"])) for r in records]
    assert [k[:2] for k in key] == sorted(k[:2] for k in key)
    return sorted(key)


def canonical_dot(dot):
    head = [l for l in dot.split("\n") if "->" not in l]
    return "\n".join(head) + "\n" + "\n".join(sorted(l for l in dot.split("\n") if "->" in l))
