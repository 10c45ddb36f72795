// Synthetic workload generator for bench.py and the full-size tests (host only; NOT part of the sampler).
//
// The reference ships no data and the BAM-level tools it shells out to are absent, so inputs are generated at the
// candidate + fragment level and turned into CollapsedMaps:
//   * a locus has E exons (log-normal lengths), candidates are distinct ordered exon subsets, true abundances are
//     Dirichlet(0.5) with ~40 % of the candidates switched off;
//   * paired-end fragments (2 x 76 bp mates, fragment length ~ N(200, 20)) are drawn from the expressed candidates in
//     proportion to abundance x effective length, uniformly along the transcript;
//   * a fragment is compatible with a candidate when both mates' exon chains are contiguous in it; the likelihood is
//     pmf(implied fragment length) / effective length, which is what calculateThreePrimeProb x calculateLengthProb
//     reduce to (alignmentParser.cpp:314-318, sequencingModel.cpp:59-76; the read probability cancels in :369);
//   * candidates without fragments or without coverage of their first/last exon are excluded (a simplification of
//     alignmentParser.cpp:425-467), rows are normalised (:369) and identical rows collapsed with a count (:373-411).
// Every locus is regenerable from (cfg, locus index) alone: seed = 0xB2000000 + 1000 cfg + locus.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

struct Rng {
    uint64_t s[4];
    static uint64_t splitmix(uint64_t &x) {
        uint64_t z = (x += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    explicit Rng(uint64_t seed) { for (int i = 0; i < 4; i++) s[i] = splitmix(seed); }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {  // xoshiro256**
        const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uni() { return ((next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
    int below(int n) { return (int)(uni() * n); }
    double normal() { return sqrt(-2.0 * log(uni())) * cos(6.283185307179586 * uni()); }
    double gamma(double a) {
        if (a < 1.0) return gamma(a + 1.0) * pow(uni(), 1.0 / a);
        const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
        for (;;) {
            const double x = normal();
            double v = 1.0 + c * x;
            if (v <= 0) continue;
            v = v * v * v;
            const double u = uni();
            if (log(u) < 0.5 * x * x + d * (1.0 - v + log(v))) return d * v;
        }
    }
};

struct SeqModel {  // sequencingModel.cpp:34-57 for a Gaussian(200, 20) fragment length model
    std::vector<double> pmf, cum, cum3;
    explicit SeqModel(int max_len) : pmf(max_len + 1, 0.0), cum(max_len + 1, 0.0), cum3(max_len + 1, 0.0) {
        const double mean = 200.0, sd = 20.0, nc = 1.0 / (sd * sqrt(2.0 * 3.14159265358979323846));
        for (int i = 1; i <= max_len; i++) {
            pmf[i] = nc * exp(-((i - mean) * (i - mean)) / (2.0 * sd * sd));
            cum[i] = cum[i - 1] + pmf[i];
            cum3[i] = cum3[i - 1] + cum[i];
        }
    }
    double efflen(int L) const { return cum3[L]; }
};

const int kRead = 76;
const int kMaxLen = 80000;

struct Transcript {
    std::vector<int> exons;  // gene exon indices, ascending
    std::vector<int> off;    // transcript offset of each exon, plus total length at the end
    int pos_of[32];          // position of gene exon in this transcript or -1
    int length() const { return off.back(); }
};

struct LocusOut {
    int T = 0;
    std::vector<int32_t> row_ptr, col, count;
    std::vector<double> val, efflen;
    int64_t n_frag = 0;
};

struct Cfg {
    int e_lo, e_hi, t_lo, t_hi;
    bool wide;       // cfg3: hundreds of candidates
    double incl, sigma;
};

Cfg cfg_of(int cfg) {
    switch (cfg) {
        case 3: return Cfg{13, 18, 200, 1000, true, 0.5, 0.5};
        case 4: return Cfg{3, 20, 1, 30, false, 0.75, 1.2};
        default: return Cfg{3, 20, 1, 30, false, 0.75, 0.5};
    }
}

void generate_locus(int cfg, int64_t locus, double mean_frags, const SeqModel &sm, LocusOut *out) {
    const Cfg C = cfg_of(cfg);
    // cfg5 = the loci of cfg1 (run with 64 chains)
    Rng rng(0xB2000000ull + 1000ull * (uint64_t)(cfg == 5 ? 1 : cfg) + (uint64_t)locus);
    const int E = C.e_lo + rng.below(C.e_hi - C.e_lo + 1);
    std::vector<int> exon_len(E);
    for (int e = 0; e < E; e++) {
        double l = exp(log(150.0) + 0.7 * rng.normal());
        exon_len[e] = (int)std::min(3000.0, std::max(30.0, l));
    }
    int t_want;
    if (C.wide) t_want = C.t_lo + rng.below(C.t_hi - C.t_lo + 1);
    else t_want = std::min(C.t_hi, 1 + (int)floor(-log(rng.uni()) * 6.0));
    // candidates: distinct exon subsets
    std::vector<uint32_t> masks;
    for (int tries = 0; (int)masks.size() < t_want && tries < 40 * t_want + 200; tries++) {
        const int f = rng.uni() < 0.7 ? 0 : rng.below(E / 2 + 1);
        const int g = rng.uni() < 0.7 ? E - 1 : f + rng.below(E - f);
        uint32_t m = (1u << f) | (1u << g);
        for (int e = f + 1; e < g; e++)
            if (rng.uni() < C.incl) m |= 1u << e;
        int len = 0;
        for (int e = 0; e < E; e++)
            if (m >> e & 1u) len += exon_len[e];
        if (len < 300 || len > kMaxLen) continue;
        if (std::find(masks.begin(), masks.end(), m) == masks.end()) masks.push_back(m);
    }
    if (masks.empty()) {  // fall back to the full-length transcript, padded so that it can hold a fragment
        exon_len[0] = std::max(exon_len[0], 400);
        masks.push_back((E >= 32 ? 0xffffffffu : (1u << E) - 1u));
    }
    const int T0 = (int)masks.size();
    std::vector<Transcript> tr(T0);
    for (int t = 0; t < T0; t++) {
        Transcript &x = tr[t];
        for (int e = 0; e < 32; e++) x.pos_of[e] = -1;
        int o = 0;
        for (int e = 0; e < E; e++)
            if (masks[t] >> e & 1u) {
                x.pos_of[e] = (int)x.exons.size();
                x.exons.push_back(e);
                x.off.push_back(o);
                o += exon_len[e];
            }
        x.off.push_back(o);
    }
    // abundances: Dirichlet(0.5), ~40 % switched off
    std::vector<double> rho(T0), w(T0);
    double wsum = 0;
    bool any = false;
    for (int t = 0; t < T0; t++) {
        rho[t] = rng.uni() < 0.4 ? 0.0 : rng.gamma(0.5);
        any |= rho[t] > 0;
    }
    if (!any) rho[rng.below(T0)] = 1.0;
    for (int t = 0; t < T0; t++) { w[t] = rho[t] * sm.efflen(tr[t].length()); wsum += w[t]; }
    for (int t = 0; t < T0; t++) w[t] = (t ? w[t - 1] : 0.0) + w[t] / wsum;

    double nf = mean_frags * exp(-0.5 * C.sigma * C.sigma + C.sigma * rng.normal());
    const int64_t n_frag = (int64_t)std::max(20.0, nf);

    // pass 1: fragments -> (compatible candidate, implied length) lists, collapsed by identity
    std::unordered_map<std::string, int> row_of;
    std::vector<std::vector<int32_t> > rows;  // flattened (t, len) pairs
    std::vector<int32_t> row_count;
    std::vector<int> min_first(T0, 1 << 30), max_last(T0, -1);
    std::vector<int32_t> key;
    for (int64_t i = 0; i < n_frag; i++) {
        const double u = rng.uni();
        int t = (int)(std::lower_bound(w.begin(), w.end(), u) - w.begin());
        if (t >= T0) t = T0 - 1;
        const Transcript &src = tr[t];
        const int L = src.length();
        int len = (int)lround(200.0 + 20.0 * rng.normal());
        len = std::max(2 * kRead, std::min(len, L));
        const int s = rng.below(L - len + 1);
        // exon positions (in src) of the four mate ends
        auto exon_at = [&](int p) { return (int)(std::upper_bound(src.off.begin(), src.off.end() - 1, p) - src.off.begin()) - 1; };
        const int a1 = exon_at(s), b1 = exon_at(s + kRead - 1), a2 = exon_at(s + len - kRead), b2 = exon_at(s + len - 1);
        key.clear();
        for (int c = 0; c < T0; c++) {
            const Transcript &x = tr[c];
            const int p1 = x.pos_of[src.exons[a1]], p2 = x.pos_of[src.exons[a2]];
            if (p1 < 0 || p2 < 0) continue;
            bool ok = p1 + (b1 - a1) < (int)x.exons.size() && p2 + (b2 - a2) < (int)x.exons.size();
            for (int k = 1; ok && k <= b1 - a1; k++) ok = x.exons[p1 + k] == src.exons[a1 + k];
            for (int k = 1; ok && k <= b2 - a2; k++) ok = x.exons[p2 + k] == src.exons[a2 + k];
            if (!ok) continue;
            const int start = x.off[p1] + (s - src.off[a1]);
            const int end = x.off[p2] + (s + len - src.off[a2]);
            const int l2 = end - start;
            if (l2 < 2 * kRead || l2 > x.length() || sm.pmf[l2] < 1e-300) continue;
            key.push_back(c);
            key.push_back(l2);
            min_first[c] = std::min(min_first[c], p1);
            max_last[c] = std::max(max_last[c], p2 + (b2 - a2));
        }
        if (key.empty()) continue;
        std::string ks((const char *)key.data(), key.size() * sizeof(int32_t));
        auto it = row_of.find(ks);
        if (it == row_of.end()) {
            row_of.emplace(std::move(ks), (int)rows.size());
            rows.push_back(key);
            row_count.push_back(1);
        } else {
            row_count[it->second]++;
        }
    }
    // exclusion + column compaction
    std::vector<int> new_idx(T0, -1);
    int T = 0;
    for (int c = 0; c < T0; c++) {
        const bool keep = max_last[c] >= 0 && min_first[c] == 0 && max_last[c] == (int)tr[c].exons.size() - 1;
        if (keep) new_idx[c] = T++;
    }
    if (T == 0) {  // keep the best supported candidate so that the locus is not empty
        int best = 0;
        for (int c = 1; c < T0; c++) if (max_last[c] > max_last[best]) best = c;
        new_idx[best] = T++;
    }
    out->T = T;
    out->efflen.resize(T);
    for (int c = 0; c < T0; c++)
        if (new_idx[c] >= 0) out->efflen[new_idx[c]] = sm.efflen(tr[c].length());
    // pass 2: renormalise over the kept candidates (alignmentParser.cpp:572-589) and collapse rows whose VALUES agree
    // (:590-619 compares values with double_compare; rows that differ only in a fragment length shared by all their
    // candidates are equal after normalisation).  Values are matched on their leading 40 mantissa bits.
    std::unordered_map<std::string, int> row2;
    out->row_ptr.assign(1, 0);
    std::vector<int32_t> k2;
    std::vector<double> pv;
    std::vector<int64_t> qk;
    for (size_t r = 0; r < rows.size(); r++) {
        k2.clear();
        for (size_t j = 0; j < rows[r].size(); j += 2)
            if (new_idx[rows[r][j]] >= 0) { k2.push_back(new_idx[rows[r][j]]); k2.push_back(rows[r][j + 1]); }
        if (k2.empty()) continue;
        pv.clear();
        double norm = 0;
        for (size_t j = 0; j < k2.size(); j += 2) {
            pv.push_back(sm.pmf[k2[j + 1]] / out->efflen[k2[j]]);
            norm += pv.back();
        }
        qk.clear();
        for (size_t j = 0; j < pv.size(); j++) {
            pv[j] /= norm;
            int ex;
            const double m = frexp(pv[j], &ex);
            qk.push_back(k2[2 * j]);
            qk.push_back(((int64_t)ex << 44) ^ (int64_t)llround(ldexp(m, 40)));
        }
        std::string ks((const char *)qk.data(), qk.size() * sizeof(int64_t));
        auto it = row2.find(ks);
        if (it != row2.end()) { out->count[it->second] += row_count[r]; out->n_frag += row_count[r]; continue; }
        row2.emplace(std::move(ks), (int)out->count.size());
        for (size_t j = 0; j < pv.size(); j++) {
            out->col.push_back(k2[2 * j]);
            out->val.push_back(pv[j]);
        }
        out->row_ptr.push_back((int32_t)out->val.size());
        out->count.push_back(row_count[r]);
        out->n_frag += row_count[r];
    }
}

struct Batch {
    int64_t n = 0;
    std::vector<int32_t> T, R;
    std::vector<int64_t> row_off, cell_off, t_off, n_frag;
    std::vector<int32_t> row_ptr, col, count;
    std::vector<double> val, efflen;
};

}  // namespace

extern "C" {

// Generate n loci of configuration cfg with mean_frags fragments per locus on n_threads threads: the loci ids[0..n)
// when ids is given, else [first, first + n).
void *synth_generate(int cfg, int64_t first, int64_t n, double mean_frags, int n_threads, const int64_t *ids) {
    static const SeqModel sm(kMaxLen);
    std::vector<LocusOut> loci(n);
    std::atomic<int64_t> next(0);
    auto work = [&]() {
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n) return;
            generate_locus(cfg, ids ? ids[i] : first + i, mean_frags, sm, &loci[i]);
        }
    };
    n_threads = std::max(1, n_threads);
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; t++) pool.emplace_back(work);
    for (auto &t : pool) t.join();
    Batch *b = new Batch();
    b->n = n;
    b->row_off.assign(1, 0); b->cell_off.assign(1, 0); b->t_off.assign(1, 0);
    for (int64_t i = 0; i < n; i++) {
        const LocusOut &l = loci[i];
        b->T.push_back(l.T);
        b->R.push_back((int32_t)l.count.size());
        b->n_frag.push_back(l.n_frag);
        b->row_ptr.insert(b->row_ptr.end(), l.row_ptr.begin(), l.row_ptr.end());
        b->col.insert(b->col.end(), l.col.begin(), l.col.end());
        b->val.insert(b->val.end(), l.val.begin(), l.val.end());
        b->count.insert(b->count.end(), l.count.begin(), l.count.end());
        b->efflen.insert(b->efflen.end(), l.efflen.begin(), l.efflen.end());
        b->row_off.push_back(b->row_off.back() + (int64_t)l.count.size());
        b->cell_off.push_back(b->cell_off.back() + (int64_t)l.val.size());
        b->t_off.push_back(b->t_off.back() + l.T);
    }
    return b;
}

int64_t synth_size(void *h, int what) {
    Batch *b = (Batch *)h;
    switch (what) {
        case 0: return b->n;
        case 1: return b->row_off.back();
        case 2: return b->cell_off.back();
        case 3: return b->t_off.back();
        default: return -1;
    }
}

// Copy out; row_ptr has R_l + 1 entries per locus starting at row_off[l] + l.
void synth_export(void *h, int32_t *T, int32_t *R, int64_t *n_frag, int64_t *row_off, int64_t *cell_off, int64_t *t_off,
                  int32_t *row_ptr, int32_t *col, double *val, int32_t *count, double *efflen) {
    Batch *b = (Batch *)h;
    memcpy(T, b->T.data(), b->T.size() * 4);
    memcpy(R, b->R.data(), b->R.size() * 4);
    memcpy(n_frag, b->n_frag.data(), b->n_frag.size() * 8);
    memcpy(row_off, b->row_off.data(), b->row_off.size() * 8);
    memcpy(cell_off, b->cell_off.data(), b->cell_off.size() * 8);
    memcpy(t_off, b->t_off.data(), b->t_off.size() * 8);
    memcpy(row_ptr, b->row_ptr.data(), b->row_ptr.size() * 4);
    memcpy(col, b->col.data(), b->col.size() * 4);
    memcpy(val, b->val.data(), b->val.size() * 8);
    memcpy(count, b->count.data(), b->count.size() * 4);
    memcpy(efflen, b->efflen.data(), b->efflen.size() * 8);
}

void synth_free(void *h) { delete (Batch *)h; }

}  // extern "C"
