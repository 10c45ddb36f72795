// BAM front end: everything the reference does between "a TopHat2 BAM" and "a queue of (DOT strings,
// fragment alignments) per graph" -- host C++, same shared library, C ABI bayes_bam_front_* (include/bayesembler_b200.h).
//
//   reference (assembler.cpp)                                             here
//   BamTools::BamReader                                                   BgzfStream + read_bam_record: BGZF members inflated with zlib, BAM records decoded
//   generateSpliceGraphs :272-624: filter, strand split, markDuplicates   Front::scan -> Strand::take / mark_duplicates / add_unique_reads / write_unique_reads
//     :84-249, counters :369-382, :610-619                                (the de-duplicated alignments stay in memory, in the order the reference WRITES them)
//   samtools view > .sam (:662-676)                                       write_sam: the SAM text of those alignments (the input of CEM processsam)
//   processsam (CEM, not part of the reference repository)                NOT restated: its instance file is an input (or the tool is run when a path is given)
//   graphConstructorCallback :692-1125: instance file -> DOT strings,     parse_instance + collapse_graphs
//     single-path flags, overlapping graphs collapsed
//   grapherCallback :1167-1319: region fetch, mate pairing                fetch_fragments (overlap rule of BamTools' region filter, see below)
//   calculateSequencingProbability :1322-1417                             sequencing_probability
//
// Where the ORDER of the reference's output depends on the iteration order of a hash container (pairs completed at the same
// position, :128-137; junction lines of a DOT string, :872-875) the same container types are used with the same hash
// functions -- std::tr1::unordered_map, and for the two Boost hashers of assembler.h:69-96 the published hash_combine /
// hash_range of Boost <= 1.80 -- so that the order is the reference's when it is built with this standard library.
// Unpinned: BamTools, Boost and CEM are absent, no reference binary can be built or run here; the oracle of this
// row is a plain-Python restatement (oracle/f4_py.py) on small synthetic BAMs.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <fstream>
#include <functional>
#include <limits>
#include <list>
#include <map>
#include <regex>
#include <sstream>
#include <string>
#include <thread>
#include <tr1/unordered_map>
#include <tr1/unordered_set>
#include <vector>

#include "internal.h"

namespace bayes {
namespace {

struct FrontError {
    int code;
    std::string msg;
};
[[noreturn]] void fail(int code, const std::string &msg) { throw FrontError{code, msg}; }

// ---- BGZF: a series of gzip members (RFC 1952), each at most 64 KB of data -------------------------------------------------
// Every BGZF member says how long it is (extra subfield "BC"), so a batch of members is cut out of the file without
// inflating anything and inflated by a pool of host threads, one member each; the records are then parsed in file order.
// A gzip stream without the subfield (not BGZF proper) is inflated as one sequential stream instead.
class BgzfStream {
   public:
    explicit BgzfStream(const char *path) : f_(fopen(path, "rb")) {
        if (!f_) fail(BAYES_E_INVALID, std::string("cannot open ") + path);
        memset(&zs_, 0, sizeof(zs_));
        if (inflateInit2(&zs_, 15 + 16) != Z_OK) fail(BAYES_E_NOMEM, "inflateInit2 failed");
        in_.resize(1 << 16);
        out_.resize(1 << 17);
        n_threads_ = (int)std::min<unsigned>(16u, std::max<unsigned>(1u, std::thread::hardware_concurrency()));
        // BGZF proper?  (peek at the first member's header)
        unsigned char h[18];
        const size_t k = fread(h, 1, sizeof h, f_);
        rewind(f_);
        blocked_ = k == sizeof h && h[0] == 31 && h[1] == 139 && (h[3] & 4) && h[12] == 'B' && h[13] == 'C' && h[14] == 2 && h[15] == 0;
    }
    ~BgzfStream() {
        inflateEnd(&zs_);
        if (f_) fclose(f_);
    }
    // exactly n bytes, or false at a clean end of file (n bytes cut short is an error)
    bool read(void *dst, size_t n) {
        unsigned char *d = (unsigned char *)dst;
        size_t got = 0;
        while (got < n) {
            if (out_pos_ == out_len_ && !(blocked_ ? refill_blocks() : refill())) {
                if (got == 0) return false;
                fail(BAYES_E_INVALID, "truncated BAM file");
            }
            const size_t k = std::min(n - got, out_len_ - out_pos_);
            memcpy(d + got, out_.data() + out_pos_, k);
            out_pos_ += k;
            got += k;
        }
        return true;
    }

   private:
    // a batch of members: cut, inflated in parallel, concatenated
    bool refill_blocks() {
        const int kBatch = 64 * n_threads_;
        struct Member {
            size_t in_off, in_len, out_off, out_len;
        };
        std::vector<Member> members;
        comp_.clear();
        size_t out_total = 0;
        for (int m = 0; m < kBatch; m++) {
            unsigned char h[12];
            const size_t k = fread(h, 1, sizeof h, f_);
            if (k == 0) break;
            if (k != sizeof h || h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) fail(BAYES_E_INVALID, "not a BGZF block (header)");
            const size_t xlen = (size_t)h[10] | ((size_t)h[11] << 8);
            std::vector<unsigned char> extra(xlen);
            if (xlen && fread(extra.data(), 1, xlen, f_) != xlen) fail(BAYES_E_INVALID, "truncated BGZF block");
            long bsize = -1;
            for (size_t i = 0; i + 4 <= xlen;) {
                const size_t slen = (size_t)extra[i + 2] | ((size_t)extra[i + 3] << 8);
                if (extra[i] == 'B' && extra[i + 1] == 'C' && slen == 2 && i + 6 <= xlen) bsize = (long)extra[i + 4] | ((long)extra[i + 5] << 8);
                i += 4 + slen;
            }
            const long cdata = bsize - (long)xlen - 19;
            if (bsize < 0 || cdata < 0) fail(BAYES_E_INVALID, "not a BGZF block (no length)");
            const size_t at = comp_.size();
            comp_.resize(at + (size_t)cdata + 8);
            if (fread(comp_.data() + at, 1, (size_t)cdata + 8, f_) != (size_t)cdata + 8) fail(BAYES_E_INVALID, "truncated BGZF block");
            const size_t isize = le32(comp_.data() + at + cdata + 4);
            if (isize > (1u << 16)) fail(BAYES_E_INVALID, "BGZF block larger than 64 KB");
            members.push_back(Member{at, (size_t)cdata, out_total, isize});
            out_total += isize;
        }
        if (members.empty()) return false;
        if (out_.size() < out_total) out_.resize(out_total);
        std::atomic<size_t> next(0);
        std::atomic<int> bad(0);
        auto work = [&]() {
            z_stream z;
            memset(&z, 0, sizeof z);
            if (inflateInit2(&z, -15) != Z_OK) { bad = 1; return; }
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= members.size()) break;
                const Member &mb = members[i];
                inflateReset(&z);
                z.next_in = comp_.data() + mb.in_off;
                z.avail_in = (uInt)mb.in_len;
                z.next_out = out_.data() + mb.out_off;
                z.avail_out = (uInt)mb.out_len;
                const int rc = mb.out_len == 0 && mb.in_len <= 2 ? Z_STREAM_END : inflate(&z, Z_FINISH);
                if (rc != Z_STREAM_END || z.avail_out != 0 ||
                    (uint32_t)crc32(crc32(0L, Z_NULL, 0), out_.data() + mb.out_off, (uInt)mb.out_len) != le32(comp_.data() + mb.in_off + mb.in_len))
                    bad = 1;
            }
            inflateEnd(&z);
        };
        const int n_thr = (int)std::min<size_t>((size_t)n_threads_, members.size());
        std::vector<std::thread> pool;
        for (int t = 1; t < n_thr; t++) pool.emplace_back(work);
        work();
        for (auto &t : pool) t.join();
        if (bad) fail(BAYES_E_INVALID, "corrupt BGZF block (inflate or CRC failed)");
        out_pos_ = 0;
        out_len_ = out_total;
        return out_total > 0 || refill_blocks();  // (a batch of empty members, e.g. the end-of-file marker alone: look further)
    }
    static uint32_t le32(const unsigned char *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
    bool refill() {
        out_pos_ = out_len_ = 0;
        while (out_len_ == 0) {
            if (zs_.avail_in == 0) {
                const size_t k = fread(in_.data(), 1, in_.size(), f_);
                if (k == 0) {
                    if (mid_member_) fail(BAYES_E_INVALID, "truncated BGZF block");
                    return false;
                }
                zs_.next_in = in_.data();
                zs_.avail_in = (uInt)k;
            }
            zs_.next_out = out_.data();
            zs_.avail_out = (uInt)out_.size();
            const int rc = inflate(&zs_, Z_NO_FLUSH);
            if (rc != Z_OK && rc != Z_STREAM_END && rc != Z_BUF_ERROR) fail(BAYES_E_INVALID, "not a BGZF / gzip stream (inflate failed)");
            out_len_ = out_.size() - zs_.avail_out;
            mid_member_ = rc != Z_STREAM_END;
            if (rc == Z_STREAM_END) inflateReset(&zs_);  // next member
        }
        return true;
    }
    FILE *f_;
    z_stream zs_;
    std::vector<unsigned char> in_, out_, comp_;
    size_t out_pos_ = 0, out_len_ = 0;
    bool mid_member_ = false, blocked_ = false;
    int n_threads_ = 1;
};

// ---- BAM records ---------------------------------------------------------------------------------------------------------------
struct CigarOp {
    char type;
    uint32_t length;
};

struct BamRec {
    int32_t ref_id = -1, pos = -1, mate_ref = -1, mate_pos = -1, tlen = 0;
    uint16_t flag = 0;
    uint8_t mapq = 0;
    std::string name, qual, md;  // qual: phred + 33 text like BamTools' Qualities
    std::vector<CigarOp> cigar;
    std::string seq_packed, tags_raw;  // 4-bit bases and the optional fields as they are in the file: turned into SAM text only when written
    int32_t l_seq = 0;
    uint32_t nh = 0, hi = 0;
    bool has_nh = false, has_hi = false, has_md = false;
    bool paired() const { return flag & 0x1; }
    bool mapped() const { return !(flag & 0x4); }
    bool mate_mapped() const { return !(flag & 0x8); }
    bool reverse() const { return flag & 0x10; }
    bool mate_reverse() const { return flag & 0x20; }
    bool first_mate() const { return flag & 0x40; }
    bool primary() const { return !(flag & 0x100); }
    bool failed_qc() const { return flag & 0x200; }
    int32_t end_pos() const {  // BamAlignment::GetEndPosition(): half-open, reference-consuming operations
        int32_t e = pos;
        for (const CigarOp &c : cigar)
            if (c.type == 'M' || c.type == 'D' || c.type == 'N' || c.type == '=' || c.type == 'X') e += (int32_t)c.length;
        return e;
    }
};

struct BamHeader {
    std::string text;
    std::vector<std::string> ref_name;
    std::vector<int32_t> ref_len;
};

template <typename T>
T le(const unsigned char *p) {
    T v;
    memcpy(&v, p, sizeof(T));
    return v;
}

void read_header(BgzfStream &in, BamHeader *h) {
    unsigned char b[8];
    if (!in.read(b, 4) || memcmp(b, "BAM\1", 4) != 0) fail(BAYES_E_INVALID, "not a BAM file (magic)");
    if (!in.read(b, 4)) fail(BAYES_E_INVALID, "truncated BAM header");
    const int32_t l_text = le<int32_t>(b);
    if (l_text < 0) fail(BAYES_E_INVALID, "bad BAM header");
    h->text.resize(l_text);
    if (l_text && !in.read(&h->text[0], l_text)) fail(BAYES_E_INVALID, "truncated BAM header");
    while (!h->text.empty() && h->text.back() == '\0') h->text.pop_back();
    if (!in.read(b, 4)) fail(BAYES_E_INVALID, "truncated BAM header");
    const int32_t n_ref = le<int32_t>(b);
    for (int32_t i = 0; i < n_ref; i++) {
        if (!in.read(b, 4)) fail(BAYES_E_INVALID, "truncated BAM header");
        const int32_t l_name = le<int32_t>(b);
        if (l_name < 1 || l_name > 1 << 20) fail(BAYES_E_INVALID, "bad reference name length");
        std::string name(l_name, '\0');
        if (!in.read(&name[0], l_name) || !in.read(b, 4)) fail(BAYES_E_INVALID, "truncated BAM header");
        name.resize(strlen(name.c_str()));
        h->ref_name.push_back(name);
        h->ref_len.push_back(le<int32_t>(b));
    }
}

void append_tag_value(std::string *out, char type, const unsigned char *&p, const unsigned char *end, long long *ival, bool *is_int) {
    char buf[64];
    *is_int = false;
    auto need = [&](size_t n) { if ((size_t)(end - p) < n) fail(BAYES_E_INVALID, "truncated optional field"); };
    switch (type) {
        case 'A': need(1); out->append("A:").push_back((char)*p); p += 1; return;
        case 'c': need(1); *ival = (int8_t)*p; p += 1; break;
        case 'C': need(1); *ival = *p; p += 1; break;
        case 's': need(2); *ival = le<int16_t>(p); p += 2; break;
        case 'S': need(2); *ival = le<uint16_t>(p); p += 2; break;
        case 'i': need(4); *ival = le<int32_t>(p); p += 4; break;
        case 'I': need(4); *ival = le<uint32_t>(p); p += 4; break;
        case 'f': need(4); snprintf(buf, sizeof buf, "f:%g", le<float>(p)); out->append(buf); p += 4; return;
        case 'Z':
        case 'H': {
            const unsigned char *q = (const unsigned char *)memchr(p, 0, end - p);
            if (!q) fail(BAYES_E_INVALID, "unterminated string field");
            out->push_back(type);
            out->push_back(':');
            out->append((const char *)p, q - p);
            p = q + 1;
            return;
        }
        case 'B': {
            need(5);
            const char sub = (char)p[0];
            const int32_t cnt = le<int32_t>(p + 1);
            p += 5;
            out->append("B:").push_back(sub);
            for (int32_t i = 0; i < cnt; i++) {
                std::string one;
                long long v = 0;
                bool isi = false;
                append_tag_value(&one, sub, p, end, &v, &isi);
                out->push_back(',');
                if (isi) out->append(std::to_string(v));
                else out->append(one.substr(2));
            }
            return;
        }
        default: fail(BAYES_E_INVALID, std::string("unknown optional field type '") + type + "'");
    }
    *is_int = true;
    out->append("i:").append(std::to_string(*ival));
}

// the optional fields of a record as SAM text (tab separated), from their raw bytes
std::string tags_sam_text(const std::string &raw) {
    std::string out;
    const unsigned char *p = (const unsigned char *)raw.data(), *end = p + raw.size();
    while (p < end) {
        if (end - p < 3) fail(BAYES_E_INVALID, "truncated optional field");
        const char t0 = (char)p[0], t1 = (char)p[1], type = (char)p[2];
        p += 3;
        std::string text;
        long long ival = 0;
        bool is_int = false;
        append_tag_value(&text, type, p, end, &ival, &is_int);
        if (!out.empty()) out.push_back('\t');
        out.push_back(t0);
        out.push_back(t1);
        out.push_back(':');
        out.append(text);
    }
    return out;
}

std::string seq_text(const BamRec &r);

bool read_bam_record(BgzfStream &in, BamRec *r, std::vector<unsigned char> &buf) {
    unsigned char b4[4];
    if (!in.read(b4, 4)) return false;
    const int32_t block = le<int32_t>(b4);
    if (block < 32 || block > (1 << 28)) fail(BAYES_E_INVALID, "bad BAM record size");
    buf.resize(block);
    if (!in.read(buf.data(), block)) fail(BAYES_E_INVALID, "truncated BAM record");
    const unsigned char *p = buf.data(), *end = p + block;
    *r = BamRec();
    r->ref_id = le<int32_t>(p);
    r->pos = le<int32_t>(p + 4);
    const uint8_t l_name = p[8];
    r->mapq = p[9];
    const uint16_t n_cigar = le<uint16_t>(p + 12);
    r->flag = le<uint16_t>(p + 14);
    const int32_t l_seq = le<int32_t>(p + 16);
    r->mate_ref = le<int32_t>(p + 20);
    r->mate_pos = le<int32_t>(p + 24);
    r->tlen = le<int32_t>(p + 28);
    p += 32;
    if (l_seq < 0 || (size_t)(end - p) < (size_t)l_name + 4u * n_cigar + (size_t)(l_seq + 1) / 2 + (size_t)l_seq) fail(BAYES_E_INVALID, "BAM record shorter than its fields");
    r->name.assign((const char *)p, l_name ? l_name - 1 : 0);
    p += l_name;
    static const char ops[] = "MIDNSHP=X";
    for (int i = 0; i < n_cigar; i++, p += 4) {
        const uint32_t v = le<uint32_t>(p);
        if ((v & 15u) > 8u) fail(BAYES_E_INVALID, "unknown CIGAR operation");
        r->cigar.push_back(CigarOp{ops[v & 15u], v >> 4});
    }
    r->l_seq = l_seq;
    r->seq_packed.assign((const char *)p, (size_t)(l_seq + 1) / 2);
    p += (l_seq + 1) / 2;
    r->qual.resize(l_seq);
    for (int32_t i = 0; i < l_seq; i++) r->qual[i] = (char)(p[i] + 33);
    p += l_seq;
    r->tags_raw.assign((const char *)p, (size_t)(end - p));
    // the three tags the front end itself reads: NH, HI (integers of any width), MD (string); the others are only walked over
    while (p < end) {
        if (end - p < 3) fail(BAYES_E_INVALID, "truncated optional field");
        const char t0 = (char)p[0], t1 = (char)p[1], type = (char)p[2];
        p += 3;
        const bool nh = t0 == 'N' && t1 == 'H', hi = t0 == 'H' && t1 == 'I', md = t0 == 'M' && t1 == 'D' && type == 'Z';
        if (nh || hi || md || type == 'B') {
            std::string text;
            long long ival = 0;
            bool is_int = false;
            append_tag_value(&text, type, p, end, &ival, &is_int);
            if (nh && is_int) { r->nh = (uint32_t)ival; r->has_nh = true; }
            if (hi && is_int) { r->hi = (uint32_t)ival; r->has_hi = true; }
            if (md) { r->md = text.substr(2); r->has_md = true; }
        } else {
            size_t skip = 0;
            switch (type) {
                case 'A': case 'c': case 'C': skip = 1; break;
                case 's': case 'S': skip = 2; break;
                case 'i': case 'I': case 'f': skip = 4; break;
                case 'Z': case 'H': {
                    const unsigned char *q = (const unsigned char *)memchr(p, 0, end - p);
                    if (!q) fail(BAYES_E_INVALID, "unterminated string field");
                    skip = (size_t)(q - p) + 1;
                    break;
                }
                default: fail(BAYES_E_INVALID, std::string("unknown optional field type '") + type + "'");
            }
            if ((size_t)(end - p) < skip) fail(BAYES_E_INVALID, "truncated optional field");
            p += skip;
        }
    }
    return true;
}

std::string seq_text(const BamRec &r) {
    static const char bases[] = "=ACMGRSVTWYHKDBN";
    std::string s((size_t)r.l_seq, 'N');
    const unsigned char *p = (const unsigned char *)r.seq_packed.data();
    for (int32_t i = 0; i < r.l_seq; i++) s[i] = bases[(p[i >> 1] >> ((~i & 1) << 2)) & 15];
    return s;
}

// ---- duplicate removal (assembler.cpp:84-249) -------------------------------------------------------------------------------
// Boost <= 1.80: hash_value(unsigned) is the value; hash_value(std::string) = hash_range over the characters with
// hash_combine(seed, v): seed ^= v + 0x9e3779b9 + (seed << 6) + (seed >> 2)
size_t boost_hash_string(const std::string &s) {
    size_t seed = 0;
    for (char c : s) seed ^= (size_t)c + 0x9e3779b9 + (seed << 6) + (seed >> 2);
    return seed;
}

struct ReadId {  // assembler.h:58-67
    std::string name;
    uint32_t hi_tag;
    bool operator==(const ReadId &o) const { return name == o.name && hi_tag == o.hi_tag; }
};
struct ReadIdHasher {  // assembler.h:69-75
    size_t operator()(const ReadId &ri) const { return (boost_hash_string(ri.name) ^ ((size_t)ri.hi_tag << 1)) >> 1; }
};
struct PairInfo {  // assembler.h:77-88
    uint32_t pos, insert;
    std::string cigar, m_cigar;
    bool operator==(const PairInfo &o) const { return pos == o.pos && insert == o.insert && cigar == o.cigar && m_cigar == o.m_cigar; }
};
struct PairInfoHasher {  // assembler.h:90-96
    size_t operator()(const PairInfo &pi) const {
        return ((((((size_t)pi.pos ^ ((size_t)pi.insert << 1)) >> 1) ^ (boost_hash_string(pi.cigar) >> 1)) << 1) ^ (boost_hash_string(pi.m_cigar) >> 1));
    }
};
typedef std::pair<BamRec *, BamRec *> ReadPair;
typedef std::tr1::unordered_map<ReadId, BamRec *, ReadIdHasher> ReadIDs;
typedef std::map<uint32_t, ReadIDs> FirstReads;
typedef std::tr1::unordered_map<PairInfo, ReadPair, PairInfoHasher> ReadPairs;
typedef std::multimap<uint32_t, BamRec *> UniqueReads;

std::string dedup_cigar_string(const std::vector<CigarOp> &cigar) {  // generateCigarString, assembler.cpp:130-154
    std::string s;
    char num[16];
    for (const CigarOp &c : cigar) {
        if (c.type == 'M' || c.type == 'D' || c.type == 'N') {
            s.append(num, (size_t)snprintf(num, sizeof num, "%u", c.length));
            s.push_back(c.type == 'N' ? 'N' : 'M');
        } else if (c.type != 'I') {
            fail(BAYES_E_INVALID, std::string("Unhandled cigar string symbol '") + c.type + "'!");
        }
    }
    return s;
}

struct Strand {
    FirstReads first_reads;
    ReadPairs read_pairs;
    UniqueReads unique_reads;
    int current_position = -1, cur_reference = -1;
    std::vector<BamRec> written;  // what the reference saves to its *_nd_*.bam, in that order
    double nd_nm_reads = 0;
    unsigned long num_reads_paired = 0;

    void add_unique_reads() {  // assembler.cpp:124-135
        for (ReadPairs::iterator it = read_pairs.begin(); it != read_pairs.end(); ++it) {
            unique_reads.insert(unique_reads.end(), std::make_pair((uint32_t)it->second.first->pos, it->second.first));
            unique_reads.insert(unique_reads.end(), std::make_pair((uint32_t)it->second.second->pos, it->second.second));
        }
        read_pairs.clear();
    }
    void write_unique_reads() {  // assembler.cpp:84-121
        const uint32_t left_most = first_reads.empty() ? std::numeric_limits<uint32_t>::max() : first_reads.begin()->first;
        while (!unique_reads.empty() && unique_reads.begin()->first < left_most) {
            BamRec *a = unique_reads.begin()->second;
            if (!a->has_nh) fail(BAYES_E_INVALID, "alignment " + a->name + " has no NH tag");
            if (!a->has_hi && a->nh != 1) fail(BAYES_E_INVALID, "multi-mapping alignment " + a->name + " has no HI tag");
            if (a->nh == 1) nd_nm_reads += 1;
            written.push_back(std::move(*a));
            delete a;
            unique_reads.erase(unique_reads.begin());
        }
    }
    void mark_duplicates(BamRec &cur) {  // assembler.cpp:157-249 (the record is moved into the containers, not copied: the caller reads the next one into it)
        ReadId ri;
        ri.name = cur.name;
        ri.hi_tag = cur.has_hi ? cur.hi : 0;
        FirstReads::iterator fit = first_reads.find((uint32_t)cur.mate_pos);
        if (fit == first_reads.end()) {
            if (first_reads[(uint32_t)cur.pos].count(ri)) fail(BAYES_E_INVALID, "alignment " + cur.name + " seen twice at one position");
            first_reads[(uint32_t)cur.pos].insert(std::make_pair(ri, new BamRec(std::move(cur))));
            return;
        }
        ReadIDs::iterator pit = fit->second.find(ri);
        if (pit == fit->second.end()) {
            if (cur.pos != cur.mate_pos) fail(BAYES_E_INVALID, "mate of " + cur.name + " not found where it should be (is the BAM sorted by coordinate?)");
            fit->second.insert(std::make_pair(ri, new BamRec(std::move(cur))));
            return;
        }
        BamRec *first = pit->second;
        PairInfo pi;
        pi.pos = (uint32_t)first->pos;
        pi.insert = (uint32_t)abs(first->tlen);
        pi.cigar = dedup_cigar_string(first->cigar);
        pi.m_cigar = dedup_cigar_string(cur.cigar);
        if ((uint32_t)cur.mate_pos != pi.pos) fail(BAYES_E_INVALID, "mate position of " + cur.name + " does not match its mate");
        ReadPairs::iterator rit = read_pairs.find(pi);
        if (rit == read_pairs.end()) {
            read_pairs.insert(std::make_pair(pi, ReadPair(first, new BamRec(std::move(cur)))));
        } else {
            if (!first->has_nh || !cur.has_nh || !rit->second.first->has_nh) fail(BAYES_E_INVALID, "alignment " + cur.name + " has no NH tag");
            if (first->nh != cur.nh) fail(BAYES_E_INVALID, "mates of " + cur.name + " disagree on NH");
            if (first->nh == 1 && rit->second.first->nh > 1) {  // a unique pair replaces a multi-mapping duplicate
                delete rit->second.first;
                delete rit->second.second;
                rit->second = ReadPair(first, new BamRec(std::move(cur)));
            } else {
                delete first;
            }
        }
        fit->second.erase(pit);
        if (fit->second.empty()) first_reads.erase(fit);
    }
    void take(BamRec &cur) {  // assembler.cpp:343-368
        if (current_position != cur.pos || cur_reference != cur.ref_id) {
            num_reads_paired += read_pairs.size();
            add_unique_reads();
            write_unique_reads();
            if (cur_reference != cur.ref_id) {
                cur_reference = cur.ref_id;
                if (!unique_reads.empty() || !first_reads.empty()) fail(BAYES_E_INVALID, "alignments without their mates at the end of a reference sequence");
            } else if (!(current_position < cur.pos)) {
                fail(BAYES_E_INVALID, "BAM file is not sorted by coordinate");
            }
        }
        const int32_t pos = cur.pos;
        mark_duplicates(cur);
        current_position = pos;
    }
    void finish() {  // assembler.cpp:371-379
        num_reads_paired += read_pairs.size();
        add_unique_reads();
        write_unique_reads();
        if (!unique_reads.empty() || !first_reads.empty()) fail(BAYES_E_INVALID, "alignments without their mates at the end of the file");
    }
    ~Strand() {
        for (auto &kv : first_reads)
            for (auto &e : kv.second) delete e.second;
        for (auto &kv : read_pairs) { delete kv.second.first; delete kv.second.second; }
        for (auto &kv : unique_reads) delete kv.second;
    }
};

// ---- read probability (assembler.cpp:1322-1417) -------------------------------------------------------------------------------
// 10^(-q / 10) for every phred score a quality character can hold: the reference calls pow() per base (:1381, :1397), the same
// expression evaluated once per score gives the same doubles
struct PhredTable {
    double err[256];
    PhredTable() {
        for (int c = 0; c < 256; c++) err[c] = pow(10, -((int)(signed char)c - 33) / (double)10);  // ((int) qualities[i] - 33 on a signed char, like the reference)
    }
};
const PhredTable g_phred;

double sequencing_probability(const std::string &md_tag, const std::string &qualities, const std::vector<CigarOp> &cigar) {
    double probability = 1;
    std::vector<unsigned short> deletions, insertions;
    unsigned short running_length = 0;
    for (const CigarOp &c : cigar) {
        if (c.type == 'M') {
            running_length += c.length;
        } else if (c.type == 'D') {
            deletions.push_back(running_length);
        } else if (c.type == 'I') {
            for (uint32_t i = 0; i < c.length; i++) {
                insertions.push_back(running_length);
                running_length++;
            }
        }
    }
    if (running_length != qualities.size()) fail(BAYES_E_INVALID, "CIGAR and base qualities disagree on the read length");
    deletions.push_back(qualities.size() + 1);
    insertions.push_back(qualities.size() + 1);
    size_t del_it = 0, ins_it = 0;
    std::vector<int> numbers;  // the digit runs of the MD tag (boost::regex "\d+" tokens)
    for (size_t i = 0; i < md_tag.size();) {
        if (md_tag[i] >= '0' && md_tag[i] <= '9') {
            size_t j = i;
            while (j < md_tag.size() && md_tag[j] >= '0' && md_tag[j] <= '9') j++;
            numbers.push_back(atoi(md_tag.substr(i, j - i).c_str()));
            i = j;
        } else {
            i++;
        }
    }
    unsigned short q_idx = 0;
    auto err_at = [&](unsigned short i) -> double {
        if (i >= qualities.size()) fail(BAYES_E_INVALID, "MD tag runs past the end of the read");
        return g_phred.err[(unsigned char)qualities[i]];
    };
    for (size_t k = 0; k < numbers.size(); k++) {
        for (unsigned short i = 0; i < numbers[k]; i++) {
            while (insertions[ins_it] == q_idx) { ins_it++; q_idx++; }
            probability *= (1 - err_at(q_idx));
            q_idx++;
        }
        while (insertions[ins_it] == q_idx) { ins_it++; q_idx++; }
        if (q_idx < qualities.size()) {
            if (deletions[del_it] == q_idx) {
                del_it++;
            } else {
                probability *= err_at(q_idx) / 3;
                q_idx++;
            }
        } else if (k + 1 != numbers.size()) {
            fail(BAYES_E_INVALID, "MD tag runs past the end of the read");
        }
    }
    if (insertions[ins_it] != qualities.size() + 1 || deletions[del_it] != qualities.size() + 1 || q_idx != qualities.size())
        fail(BAYES_E_INVALID, "MD tag, CIGAR and base qualities disagree");
    if (!(probability < 1) || !(probability >= 2.2250738585072014e-308)) fail(BAYES_E_NUMERIC, "read probability outside [DBL_MIN, 1)");
    return probability;
}

// ---- CEM instance file -> graphs (assembler.cpp:692-1125) -----------------------------------------------------------------------
struct GraphInfo {  // utils.h:80-89
    std::list<std::string> dot_strings;
    std::string reference, strand;
    int left = 0, right = 0;
    bool single_path_graph = true;
    unsigned read_count = 0;
};

std::vector<std::string> split_any(const std::string &s, char sep) {  // boost::split(.., is_any_of(sep)): empty tokens are kept
    std::vector<std::string> out(1);
    for (char c : s) {
        if (c == sep) out.push_back(std::string());
        else out.back().push_back(c);
    }
    return out;
}

void parse_instance(std::istream &in, const std::string &strand, double exon_base_coverage, bool no_pre_mrna, std::list<std::list<GraphInfo> > *unsorted_graphs,
                    int *stats_dot_excluded, int *cem_graphs) {
    static const std::regex instance_pat("^Instance\\s+(\\d+)$"), boundary_pat("^Boundary\\s+([\\w|-]+)\\s+(\\d+)\\s+(\\d+)\\s+([\\+|\\.|-])$"),
        exon_start_pat("^Segs\\s+(\\d+)$"), exon_end_pat("^Refs\\s+(\\d+)$"), path_start_pat("^SGTypes\\s+(\\d+)$"),
        path_end_pat("^PETypes\\s+(\\d+)\\s+\\d+$"), read_count_pat("^Reads\\s+(\\d+)$");
    std::smatch m;
    bool exon_line_start_pass = false, exon_line_end_pass = false, path_line_start_pass = false, path_line_end_pass = false;
    int graph_count = 0, current_left_boundary = 0, current_right_boundary = 0, reference_exon_no = 0, reference_path_no = 0, real_exon_no = 0,
        real_path_no = 0;
    unsigned current_read_count = 0;
    std::string current_ref_id, last_ref_id, current_ref_strand;
    std::list<GraphInfo> temp_graphs;
    std::tr1::unordered_map<std::string, int> junctions;
    std::tr1::unordered_set<unsigned> excluded_exons;
    std::tr1::unordered_map<int, std::tr1::unordered_set<int> > vertex_in_edges, vertex_out_edges;
    bool single_path_graph = true;
    std::string line;
    std::stringstream dot;
    long line_no = 0;
    auto bad = [&](const std::string &what) { fail(BAYES_E_INVALID, "instance file line " + std::to_string(line_no) + ": " + what); };
    while (std::getline(in, line)) {
        line_no++;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (std::regex_match(line, m, instance_pat)) {
            graph_count++;
            if (!dot.str().empty() || !junctions.empty()) bad("a new instance starts before the previous one ended");
            dot << "digraph graph_" << m[1];
            single_path_graph = true;
        }
        if (std::regex_match(line, m, boundary_pat)) {
            current_ref_id = m[1];
            current_ref_strand = m[4];
            dot << (current_ref_strand == "+" ? "_plus {\n" : current_ref_strand == "-" ? "_minus {\n" : "_dot {\n");
            current_left_boundary = atoi(m[2].str().c_str());
            current_right_boundary = atoi(m[3].str().c_str());
        }
        if (std::regex_match(line, m, read_count_pat)) current_read_count = (unsigned)strtoul(m[1].str().c_str(), 0, 10);
        if (std::regex_match(line, m, exon_end_pat)) {
            exon_line_start_pass = false;
            exon_line_end_pass = true;
            if (reference_exon_no != real_exon_no) bad("number of segment lines differs from the Segs count");
        }
        if (exon_line_start_pass && !exon_line_end_pass) {
            const std::vector<std::string> f = split_any(line, '\t');
            if (f.size() < 8) bad("segment line with fewer than 8 tab-separated fields");
            if (exon_base_coverage <= (1 - atof(f[7].c_str()))) {
                const int l = atoi(f[0].c_str()), r = atoi(f[1].c_str());
                if (l < current_left_boundary) current_left_boundary = l;
                if (r > current_right_boundary) current_right_boundary = r;
                dot << real_exon_no << " [label=\"" << current_ref_id << ";" << current_ref_strand << ";" << f[0] << ";" << f[1] << "\"]\n";
            } else {
                excluded_exons.insert((unsigned)real_exon_no);
            }
            real_exon_no += 1;
        }
        if (std::regex_match(line, m, exon_start_pat)) {
            exon_line_start_pass = true;
            exon_line_end_pass = false;
            real_exon_no = 0;
            reference_exon_no = atoi(m[1].str().c_str());
        }
        if (std::regex_match(line, m, path_end_pat)) {
            path_line_start_pass = false;
            path_line_end_pass = true;
            if (reference_path_no != real_path_no) bad("number of path lines differs from the SGTypes count");
            if (!no_pre_mrna) dot << "pre_mrna [label=\"" << current_ref_id << ";" << current_ref_strand << ";" << current_left_boundary << ";" << current_right_boundary << "\"]\n";
            for (std::tr1::unordered_map<std::string, int>::iterator it = junctions.begin(); it != junctions.end(); ++it)
                dot << it->first << " [coverage=\"" << it->second << "\"]\n";
            dot << "}\n";
            if (current_ref_id != last_ref_id) {
                if (!temp_graphs.empty()) {
                    unsorted_graphs->push_back(temp_graphs);
                    temp_graphs.clear();
                }
                last_ref_id = current_ref_id;
            }
            if (current_ref_strand == "+" || current_ref_strand == "-") {
                GraphInfo g;
                g.reference = current_ref_id;
                g.strand = current_ref_strand;
                g.left = current_left_boundary;
                g.right = current_right_boundary;
                g.single_path_graph = single_path_graph;
                g.dot_strings.push_back(dot.str());
                g.read_count = current_read_count;
                temp_graphs.push_back(g);
            } else {
                if (strand != ".") bad("an instance without a strand in strand-specific data");
                (*stats_dot_excluded)++;
            }
            dot.clear();
            dot.str(std::string());
            junctions.clear();
            excluded_exons.clear();
            vertex_in_edges.clear();
            vertex_out_edges.clear();
        }
        if (path_line_start_pass && !path_line_end_pass) {
            const std::vector<std::string> f = split_any(line, ' ');
            if ((int)f.size() != real_exon_no + 1) bad("path line with a wrong number of fields");
            int first_idx = -1;
            for (int i = 0; i < real_exon_no; i++) {
                if (excluded_exons.count((unsigned)i) > 0) continue;
                if (!f[i].empty() && f[i][0] == '1') { first_idx = i; break; }
            }
            int last_idx = first_idx;
            for (int i = first_idx + 1; i < real_exon_no; i++) {
                if (excluded_exons.count((unsigned)i) > 0) continue;
                if (!f[i].empty() && f[i][0] == '1') {
                    std::stringstream js;
                    js << last_idx << "->" << i;
                    const std::string cur_junction = js.str();
                    // in- and out-degrees: a graph in which no vertex has two of either is a single path (fragment length estimation)
                    if (single_path_graph) {
                        if (vertex_in_edges[i].insert(last_idx).second && vertex_in_edges[i].size() > 1) single_path_graph = false;
                        if (vertex_out_edges[last_idx].insert(i).second && vertex_out_edges[last_idx].size() > 1) single_path_graph = false;
                    }
                    last_idx = i;
                    if (junctions.count(cur_junction) > 0) junctions[cur_junction] += atoi(f[real_exon_no].c_str());
                    else junctions[cur_junction] = atoi(f[real_exon_no].c_str());
                }
            }
            real_path_no += 1;
        }
        if (std::regex_match(line, m, path_start_pat)) {
            path_line_start_pass = true;
            path_line_end_pass = false;
            real_path_no = 0;
            reference_path_no = atoi(m[1].str().c_str());
        }
    }
    if (!temp_graphs.empty()) unsorted_graphs->push_back(temp_graphs);
    *cem_graphs = graph_count;
}

bool graph_compare_pos(const GraphInfo &a, const GraphInfo &b) { return a.left < b.left; }  // assembler.cpp:72-75

// overlapping graphs of one strand become one graph with several DOT strings (assembler.cpp:1012-1125)
void collapse_graphs(std::list<std::list<GraphInfo> > *unsorted_graphs, bool no_pre_mrna, std::list<GraphInfo> *output_graphs) {
    if (no_pre_mrna) {
        for (auto &ref : *unsorted_graphs)
            for (auto &g : ref) output_graphs->push_back(g);
        return;
    }
    GraphInfo last[2];
    bool first[2] = {true, true};
    for (auto &ref : *unsorted_graphs) {
        ref.sort(graph_compare_pos);
        for (auto &g : ref) {
            const int s = g.strand == "+" ? 0 : 1;
            if (!first[s] && g.reference == last[s].reference && g.left <= last[s].right) {
                last[s].right = std::max(g.right, last[s].right);
                last[s].single_path_graph = false;
                last[s].dot_strings.push_back(g.dot_strings.front());
                last[s].read_count += g.read_count;
            } else {
                if (!first[s]) output_graphs->push_back(last[s]);
                first[s] = false;
                last[s] = g;
            }
        }
    }
    if (!first[0]) output_graphs->push_back(last[0]);
    if (!first[1]) output_graphs->push_back(last[1]);
}

// ---- per-graph fragments (grapherCallback, assembler.cpp:1167-1319) -------------------------------------------------------------
struct Fragment {
    uint32_t pos1, pos2;
    std::vector<CigarOp> cigar1, cigar2;
    double probability;
};

struct ReadInfo {
    int32_t position, fragment_length;
    std::vector<CigarOp> cigar;
    double probability;
};

// `recs`: the de-duplicated alignments of the strand file, in file order (sorted by reference, then position).
// BamTools region [left, right) on one reference: an alignment is returned when it starts before `right` and either starts
// at / after `left` or ends at / after it (BamReaderPrivate::IsOverlap of BamTools 2.x; BamTools is absent here: unpinned).
void fetch_fragments(const std::vector<BamRec> &recs, const std::vector<size_t> &ref_first, int ref_id, const GraphInfo &g, std::vector<Fragment> *out) {
    out->clear();
    std::tr1::unordered_map<std::string, ReadInfo> single_mates;
    const size_t lo = ref_first[ref_id], hi = ref_first[ref_id + 1];
    // alignments are sorted by start; none is longer than the 700 000 the dedup pass allows a pair to span, spliced reads included
    size_t i = lo;
    {
        size_t a = lo, b = hi;  // first record with pos >= left - 1 000 000
        const int32_t from = g.left > 1000000 ? g.left - 1000000 : 0;
        while (a < b) {
            const size_t mid = (a + b) / 2;
            if (recs[mid].pos < from) a = mid + 1; else b = mid;
        }
        i = a;
    }
    for (; i < hi && recs[i].pos < g.right; i++) {
        const BamRec &cur = recs[i];
        if (!(cur.pos >= g.left || cur.end_pos() >= g.left)) continue;
        if (!cur.mate_mapped()) continue;
        if (!cur.has_nh) fail(BAYES_E_INVALID, "alignment " + cur.name + " has no NH tag");
        if (cur.nh > 1) continue;
        if (cur.mate_pos + 1 > g.right || cur.mate_pos + 1 < g.left) continue;
        if (!cur.has_md) fail(BAYES_E_INVALID, "alignment " + cur.name + " has no MD tag");
        std::tr1::unordered_map<std::string, ReadInfo>::iterator it = single_mates.find(cur.name);
        if (it != single_mates.end()) {
            const ReadInfo &mate = it->second;
            if (!(mate.position <= cur.pos) || mate.position != cur.mate_pos) fail(BAYES_E_INVALID, "mates of " + cur.name + " disagree on their positions");
            if (mate.fragment_length != -cur.tlen && mate.position != cur.pos) fail(BAYES_E_INVALID, "mates of " + cur.name + " disagree on the insert size");
            Fragment f;
            f.pos1 = (uint32_t)mate.position;
            f.cigar1 = mate.cigar;
            f.pos2 = (uint32_t)cur.pos;
            f.cigar2 = cur.cigar;
            f.probability = mate.probability * sequencing_probability(cur.md, cur.qual, cur.cigar);
            out->push_back(f);
            single_mates.erase(it);
        } else {
            ReadInfo ri;
            ri.position = cur.pos;
            ri.fragment_length = cur.tlen;
            ri.cigar = cur.cigar;
            ri.probability = sequencing_probability(cur.md, cur.qual, cur.cigar);
            single_mates.insert(std::make_pair(cur.name, ri));
        }
    }
}

std::string cigar_text(const std::vector<CigarOp> &c) {
    std::string s;
    for (const CigarOp &o : c) s += std::to_string(o.length) + o.type;
    return s.empty() ? "*" : s;
}

}  // namespace
}  // namespace bayes

using namespace bayes;

struct bayes_bam_front {
    BamHeader header;
    std::string strand_specific;
    Strand strand[2];  // unstranded: [0]; strand-specific: [0] = plus file, [1] = minus file
    std::vector<size_t> ref_first[2];
    unsigned long total_mapped_pair_reads = 0;
};

namespace {

int guarded(const char *what, const std::function<void()> &fn) {
    try {
        fn();
        return BAYES_OK;
    } catch (const FrontError &e) {
        set_error(std::string(what) + ": " + e.msg);
        return e.code;
    } catch (const std::exception &e) {
        set_error(std::string(what) + ": " + e.what());
        return BAYES_E_INVALID;
    }
}

// generateSpliceGraphs, assembler.cpp:329-380 (unstranded) / :400-520 (strand-specific)
void scan(bayes_bam_front *F, const char *path) {
    BgzfStream in(path);
    read_header(in, &F->header);
    const bool unstranded = F->strand_specific == "unstranded";
    // strand-specific: "first" = dUTP-like orientation, the plus file takes the pairs the reference's writer_pair.first takes
    Strand *plus = &F->strand[0], *minus = &F->strand[1];
    Strand *w_first = F->strand_specific == "first" ? plus : minus, *w_second = F->strand_specific == "first" ? minus : plus;
    BamRec cur;
    std::vector<unsigned char> buf;
    while (read_bam_record(in, &cur, buf)) {
        if (!cur.paired()) fail(BAYES_E_INVALID, "alignment " + cur.name + " is not from a paired-end read");
        if (cur.failed_qc()) fail(BAYES_E_INVALID, "alignment " + cur.name + " failed quality control");
        if (!(cur.mapped() && cur.primary() && cur.mate_mapped())) continue;
        F->total_mapped_pair_reads += 1;
        if (!((cur.reverse() != cur.mate_reverse()) && cur.ref_id == cur.mate_ref)) continue;
        if (abs(cur.tlen) > 700000) continue;  // same threshold as CEM
        if (unstranded) {
            F->strand[0].take(cur);
        } else if ((cur.first_mate() && cur.reverse() && cur.pos >= cur.mate_pos) || (!cur.first_mate() && !cur.reverse() && cur.pos <= cur.mate_pos)) {
            w_first->take(cur);
        } else if ((cur.first_mate() && !cur.reverse() && cur.pos <= cur.mate_pos) || (!cur.first_mate() && cur.reverse() && cur.pos >= cur.mate_pos)) {
            w_second->take(cur);
        }
    }
    for (int s = 0; s < (unstranded ? 1 : 2); s++) {
        Strand &S = F->strand[s];
        S.finish();
        // first record of every reference in the written order
        const size_t n_ref = F->header.ref_name.size();
        F->ref_first[s].assign(n_ref + 1, S.written.size());
        size_t at = 0;
        for (size_t r = 0; r < n_ref; r++) {
            while (at < S.written.size() && S.written[at].ref_id < (int32_t)r) at++;
            F->ref_first[s][r] = at;
        }
    }
}

}  // namespace

extern "C" {

int bayes_bam_front_open(const char *bam_path, const char *strand_specific, bayes_bam_front **out) {
    if (!bam_path || !strand_specific || !out) {
        set_error("bayes_bam_front_open: null argument");
        return BAYES_E_INVALID;
    }
    *out = nullptr;
    const std::string s = strand_specific;
    if (s != "unstranded" && s != "first" && s != "second") {
        set_error("bayes_bam_front_open: strand_specific must be \"unstranded\", \"first\" or \"second\"");
        return BAYES_E_INVALID;
    }
    bayes_bam_front *F = new (std::nothrow) bayes_bam_front();
    if (!F) return BAYES_E_NOMEM;
    F->strand_specific = s;
    const int rc = guarded("bayes_bam_front_open", [&]() { scan(F, bam_path); });
    if (rc != BAYES_OK) {
        delete F;
        return rc;
    }
    *out = F;
    return BAYES_OK;
}

void bayes_bam_front_close(bayes_bam_front *f) { delete f; }

int bayes_bam_front_counts(bayes_bam_front *f, double *total_mapped_pairs, double *unique_pairs, double *pairs_written) {
    if (!f) return BAYES_E_INVALID;
    double nd = 0, written = 0;
    for (int s = 0; s < 2; s++) {
        nd += f->strand[s].nd_nm_reads;
        written += (double)f->strand[s].num_reads_paired;
    }
    // assembler.cpp:384-388: both counts are of alignments, two per pair
    if (total_mapped_pairs) *total_mapped_pairs = (double)(f->total_mapped_pair_reads / 2);
    if (unique_pairs) *unique_pairs = floor(nd / 2);
    if (pairs_written) *pairs_written = written;
    return BAYES_OK;
}

int64_t bayes_bam_front_num_alignments(bayes_bam_front *f, int32_t strand_file) {
    if (!f || strand_file < 0 || strand_file > 1) return BAYES_E_INVALID;
    return (int64_t)f->strand[strand_file].written.size();
}

int bayes_bam_front_write_sam(bayes_bam_front *f, int32_t strand_file, const char *sam_path, int32_t with_header) {
    if (!f || strand_file < 0 || strand_file > 1 || !sam_path) {
        set_error("bayes_bam_front_write_sam: bad arguments");
        return BAYES_E_INVALID;
    }
    return guarded("bayes_bam_front_write_sam", [&]() {
        std::ofstream out(sam_path);
        if (!out) fail(BAYES_E_INVALID, std::string("cannot write ") + sam_path);
        if (with_header && !f->header.text.empty()) out << f->header.text << (f->header.text.back() == '\n' ? "" : "\n");
        auto ref = [&](int32_t id) { return id < 0 || id >= (int32_t)f->header.ref_name.size() ? std::string("*") : f->header.ref_name[id]; };
        for (const BamRec &r : f->strand[strand_file].written) {
            out << r.name << '\t' << r.flag << '\t' << ref(r.ref_id) << '\t' << r.pos + 1 << '\t' << (int)r.mapq << '\t' << cigar_text(r.cigar) << '\t'
                << (r.mate_ref < 0 ? std::string("*") : r.mate_ref == r.ref_id ? std::string("=") : ref(r.mate_ref)) << '\t' << r.mate_pos + 1 << '\t' << r.tlen << '\t'
                << (r.l_seq == 0 ? std::string("*") : seq_text(r)) << '\t' << (r.qual.empty() || (unsigned char)(r.qual[0] - 33) == 0xff ? std::string("*") : r.qual);
            if (!r.tags_raw.empty()) out << '\t' << tags_sam_text(r.tags_raw);
            out << '\n';
        }
    });
}

int bayes_bam_front_graphs(bayes_bam_front *f, int32_t strand_file, const char *instance_path, const char *strand, double exon_base_coverage,
                           int32_t no_pre_mrna, const char *graph_file_path, int32_t append, int32_t *n_graphs, int32_t *n_with_fragments,
                           int32_t *n_excluded, int32_t *cem_graphs) {
    if (!f || strand_file < 0 || strand_file > 1 || !instance_path || !strand || !graph_file_path) {
        set_error("bayes_bam_front_graphs: bad arguments");
        return BAYES_E_INVALID;
    }
    return guarded("bayes_bam_front_graphs", [&]() {
        std::ifstream in(instance_path);
        if (!in) fail(BAYES_E_INVALID, std::string("cannot open instance file ") + instance_path);
        std::list<std::list<GraphInfo> > unsorted;
        std::list<GraphInfo> graphs;
        int excluded = 0, cem = 0;
        parse_instance(in, strand, exon_base_coverage, no_pre_mrna != 0, &unsorted, &excluded, &cem);
        collapse_graphs(&unsorted, no_pre_mrna != 0, &graphs);
        std::ofstream out(graph_file_path, append ? std::ios::app : std::ios::trunc);
        if (!out) fail(BAYES_E_INVALID, std::string("cannot write ") + graph_file_path);
        if (!append) {
            double total = 0, unique = 0;
            bayes_bam_front_counts(f, &total, &unique, nullptr);
            char head[128];
            snprintf(head, sizeof head, "# bayesembler_b200 graph file\nfragments %.0f %.0f\n", unique, total);
            out << head;
        }
        std::map<std::string, int> ref_id;
        for (size_t r = 0; r < f->header.ref_name.size(); r++) ref_id[f->header.ref_name[r]] = (int)r;
        // the graphs are independent: host threads fetch and format them, the file is written in graph order
        std::vector<const GraphInfo *> gv;
        std::vector<int> gref;
        for (const GraphInfo &g : graphs) {
            std::map<std::string, int>::iterator it = ref_id.find(g.reference);
            if (it == ref_id.end()) fail(BAYES_E_INVALID, "instance file names a reference sequence the BAM header does not have: " + g.reference);
            gv.push_back(&g);
            gref.push_back(it->second);
        }
        std::vector<std::string> text(gv.size());
        std::vector<std::string> errors(gv.size());
        std::vector<int> codes(gv.size(), 0);
        std::atomic<size_t> next(0);
        auto work = [&]() {
            std::vector<Fragment> frags;
            char num[64];
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= gv.size()) return;
                try {
                    const GraphInfo &g = *gv[i];
                    fetch_fragments(f->strand[strand_file].written, f->ref_first[strand_file], gref[i], g, &frags);
                    if (frags.empty()) continue;  // assembler.cpp:1262-1280: a graph without fragments is dropped
                    std::string &t = text[i];
                    t = "graph " + std::to_string(g.dot_strings.size()) + " " + std::to_string(frags.size()) + " " + (g.single_path_graph ? "1" : "0") + "\n";
                    for (const std::string &d : g.dot_strings) t += d;
                    for (const Fragment &fr : frags) {
                        snprintf(num, sizeof num, "%.17g", fr.probability);
                        t += std::to_string(fr.pos1) + " " + cigar_text(fr.cigar1) + " " + std::to_string(fr.pos2) + " " + cigar_text(fr.cigar2) + " " + num + "\n";
                    }
                } catch (const FrontError &e) {
                    codes[i] = e.code;
                    errors[i] = e.msg;
                }
            }
        };
        const int n_thr = (int)std::min<size_t>(std::min<unsigned>(16u, std::max<unsigned>(1u, std::thread::hardware_concurrency())), std::max<size_t>(1, gv.size()));
        std::vector<std::thread> pool;
        for (int t = 1; t < n_thr; t++) pool.emplace_back(work);
        work();
        for (auto &t : pool) t.join();
        int with = 0;
        for (size_t i = 0; i < gv.size(); i++) {
            if (codes[i]) fail(codes[i], errors[i]);
            if (text[i].empty()) continue;
            with++;
            out << text[i];
        }
        if (n_graphs) *n_graphs = (int32_t)graphs.size();
        if (n_with_fragments) *n_with_fragments = with;
        if (n_excluded) *n_excluded = excluded;
        if (cem_graphs) *cem_graphs = cem;
    });
}

int bayes_sequencing_probability(const char *md_tag, const char *qualities, int32_t n_cigar, const char *cigar_op, const int32_t *cigar_len, double *out) {
    if (!md_tag || !qualities || n_cigar < 0 || (n_cigar > 0 && (!cigar_op || !cigar_len)) || !out) {
        set_error("bayes_sequencing_probability: bad arguments");
        return BAYES_E_INVALID;
    }
    return guarded("bayes_sequencing_probability", [&]() {
        std::vector<CigarOp> c;
        for (int i = 0; i < n_cigar; i++) c.push_back(CigarOp{cigar_op[i], (uint32_t)cigar_len[i]});
        *out = sequencing_probability(md_tag, qualities, c);
    });
}

}  // extern "C"
