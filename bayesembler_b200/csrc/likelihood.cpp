// Likelihood / compatibility builder (host, C++): candidates + paired-end fragment alignments -> CollapsedMap in CSR form,
// the input of the Gibbs sampler.  SURVEY.md 8f row f1; replaces, behind the C ABI of include/bayesembler_b200.h,
//   GaussianFragmentLengthModel        fragmentLengthModel.cpp:40-161
//   SequencingModel                    sequencingModel.cpp:34-76
//   AlignmentParser::matchFragmentToCandidate / calculateFragTranProbabilities / fetchFragmentLengths
//                                      alignmentParser.cpp:42-249 / 251-689 / 692-725
// Same arithmetic in the same order as the reference, so compatibility sets, row order, counts and exclusions are
// identical and the probabilities agree to the last bit wherever the reference's dense sums run in column order.  What
// changes is the data structure: rows are stored sparse, and the reference's linear search for an equal row
// (O(R^2 T), alignmentParser.cpp:373-389 and :590-606) becomes a hash lookup that returns the same row -- see find_row().
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "internal.h"

namespace {

const double kDoubleUnderflow = DBL_MIN;  // utils.h:52

inline bool double_compare(double a, double b) {  // utils.h:58-61
    return (a == b) || (fabs(a - b) < fabs(std::min(a, b)) * DBL_EPSILON * 100);
}

}  // namespace

struct bayes_seq_model {
    // sequencingModel.cpp:34-57
    std::vector<double> frag_length_prob, frag_length_cum_prob, three_prime_cum_prob;
    double three_prime_prob(int three_prime_pos, int tran_length) const { return frag_length_cum_prob[three_prime_pos - 1] / three_prime_cum_prob[tran_length - 1]; }
    double length_prob(int fragment_length, int three_prime_pos) const { return frag_length_prob[fragment_length - 1] / frag_length_cum_prob[three_prime_pos - 1]; }
    double effective_length(int transcript_length) const { return three_prime_cum_prob[transcript_length - 1]; }
    int max_len() const { return (int)frag_length_prob.size(); }
};

struct bayes_collapsed_map {
    int32_t return_code = 0, T = 0;
    std::vector<int32_t> row_ptr, col_idx, row_count, new_to_old;
    std::vector<double> val, efflen;
    int64_t stats[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
};

namespace {

struct FragmentMatch {
    int first_exon_idx, last_exon_idx, fragment_length, three_prime_position;
};

struct Cand {
    const int32_t *start, *end;  // exons, 1-based inclusive (Vertex.start / .end)
    int n_exons, length;
    bool premrna;
};

// AlignmentParser::matchFragmentToCandidate (alignmentParser.cpp:42-249); returns false for a malformed CIGAR symbol
// through *bad (the reference exits).
bool match_fragment(const bayes_raw_locus &L, int frag, const Cand &c, bool plus, FragmentMatch *out, bool *bad) {
    const uint32_t pos1 = L.read_pos[2 * frag];
    int current_pos = (int)pos1 + 1;
    unsigned first_exon_idx = 0, current_exon_idx = 0;
    bool skip_fragment = true;
    for (unsigned pair_idx = 0; pair_idx < 2; pair_idx++) {
        bool first_match = true;
        if (pair_idx == 1) current_pos = (int)L.read_pos[2 * frag + 1] + 1;
        const int c0 = L.cigar_ptr[2 * frag + pair_idx], c1 = L.cigar_ptr[2 * frag + pair_idx + 1];
        for (int j = c0; j < c1; j++) {
            skip_fragment = true;
            const char type = L.cigar_op[j];
            const int len = (int)L.cigar_len[j];
            if (type == 'M') {
                if (first_match) {  // search for the exon where the alignment starts (:77-94)
                    for (unsigned k = first_exon_idx; k < (unsigned)c.n_exons; k++) {
                        if (current_pos >= c.start[k] && current_pos <= c.end[k]) {
                            if (pair_idx == 0) first_exon_idx = k;
                            current_pos += len - 1;
                            current_exon_idx = k;
                            first_match = false;
                            skip_fragment = false;
                            break;
                        }
                    }
                } else if (L.cigar_op[j - 1] != 'N') {  // after an insertion / deletion (:97-106)
                    if (current_pos <= c.end[current_exon_idx]) {
                        current_pos += len - 1;
                        skip_fragment = false;
                    }
                } else {  // splice acceptor check (:109-118)
                    if (c.start[current_exon_idx] == current_pos) {
                        current_pos += len - 1;
                        skip_fragment = false;
                    }
                }
            } else if (type == 'N') {  // splice donor check (:122-137)
                if (c.end[current_exon_idx] == current_pos) {
                    current_exon_idx += 1;
                    current_pos += len + 1;
                    skip_fragment = current_exon_idx >= (unsigned)c.n_exons;
                }
            } else if (type == 'I') {  // TopHat's definition (:139-147)
                if (current_pos <= c.end[current_exon_idx]) skip_fragment = false;
            } else if (type == 'D') {  // (:149-156)
                if (current_pos <= c.end[current_exon_idx]) {
                    current_pos += len + 1;
                    skip_fragment = false;
                }
            } else {
                *bad = true;
                return false;
            }
            if (skip_fragment) break;
        }
        if (!skip_fragment && current_pos > c.end[current_exon_idx]) skip_fragment = true;  // final right boundary (:172-182)
        if (skip_fragment) break;
    }
    if (skip_fragment) return false;
    FragmentMatch m;
    m.first_exon_idx = (int)first_exon_idx;
    m.last_exon_idx = (int)current_exon_idx;
    const int read1 = (int)pos1 + 1;
    if (first_exon_idx == current_exon_idx) {  // fragment length (:196-211)
        m.fragment_length = current_pos - read1 + 1;
    } else {
        m.fragment_length = c.end[first_exon_idx] - read1 + 1;
        for (unsigned j = first_exon_idx + 1; j < current_exon_idx; j++) m.fragment_length += c.end[j] - c.start[j] + 1;
        m.fragment_length += current_pos - c.start[current_exon_idx] + 1;
    }
    if (plus) {  // last transcript position of the second read (:216-225)
        m.three_prime_position = current_pos - c.start[current_exon_idx] + 1;
        for (unsigned j = 0; j < current_exon_idx; j++) m.three_prime_position += c.end[j] - c.start[j] + 1;
    } else {  // first transcript position of the first read, counted from the other end (:228-238)
        m.three_prime_position = c.end[first_exon_idx] - read1 + 1;
        for (unsigned j = first_exon_idx + 1; j < (unsigned)c.n_exons; j++) m.three_prime_position += c.end[j] - c.start[j] + 1;
    }
    *out = m;
    return true;
}

// Sparse row store with the reference's "first equal row" lookup.
// The reference compares a new row with every stored row, cell by cell, under double_compare (relative 1e-14) and takes
// the FIRST match.  Two rows can only match when their supports (non-zero columns) are identical -- double_compare(0, x)
// is false for x != 0 -- and their first non-zero values agree to 1e-14, so rows are bucketed by (support, first value
// quantised to 1e-9 relative); a lookup checks the value's bin and its two neighbours with the full cell-by-cell test
// and returns the smallest row index among the matches: the row the linear search would have stopped at.
struct RowStore {
    std::vector<int32_t> ptr{0}, col;
    std::vector<double> val;
    std::unordered_map<uint64_t, std::vector<int32_t> > buckets;

    int rows() const { return (int)ptr.size() - 1; }
    static int64_t bin_of(double v) { return (int64_t)floor(log(v) * 1e9); }
    static uint64_t key_of(const int32_t *c, int n, int64_t bin) {
        uint64_t h = 0xcbf29ce484222325ull ^ (uint64_t)n;
        for (int i = 0; i < n; i++) h = (h ^ (uint64_t)(uint32_t)c[i]) * 0x100000001b3ull;
        h = (h ^ (uint64_t)bin) * 0x100000001b3ull;
        return h ^ (h >> 29);
    }
    bool equal(int r, const int32_t *c, const double *v, int n) const {
        if (ptr[r + 1] - ptr[r] != n) return false;
        const int32_t *rc = col.data() + ptr[r];
        const double *rv = val.data() + ptr[r];
        for (int i = 0; i < n; i++)
            if (rc[i] != c[i] || !double_compare(rv[i], v[i])) return false;
        return true;
    }
    // smallest stored row equal to (c, v); -1 when none.  n == 0 (an all-zero row) matches the first all-zero row.
    int find_row(const int32_t *c, const double *v, int n) const {
        int best = -1;
        const int64_t bin = n ? bin_of(v[0]) : 0;
        for (int64_t b = bin - 1; b <= bin + 1; b++) {
            auto it = buckets.find(key_of(c, n, b));
            if (it == buckets.end()) continue;
            for (int32_t r : it->second) {
                if (best >= 0 && r >= best) break;  // bucket lists are ascending
                if (equal(r, c, v, n)) { best = r; break; }
            }
            if (n == 0) break;
        }
        return best;
    }
    int append(const int32_t *c, const double *v, int n) {
        const int r = rows();
        col.insert(col.end(), c, c + n);
        val.insert(val.end(), v, v + n);
        ptr.push_back((int32_t)val.size());
        buckets[key_of(c, n, n ? bin_of(v[0]) : 0)].push_back(r);
        return r;
    }
};

}  // namespace

using bayes::set_error;

extern "C" {

// GaussianFragmentLengthModel(mean, sd) (fragmentLengthModel.cpp:40-69) tabulated by SequencingModel's constructor
// (sequencingModel.cpp:34-57) up to max_transcript_length.
int bayes_seq_model_create_gaussian(double mean, double sd, int32_t max_transcript_length, bayes_seq_model **out) {
    if (!out || max_transcript_length < 1 || !(sd > 0)) {
        set_error("bayes_seq_model_create_gaussian: need sd > 0 and max_transcript_length >= 1");
        return BAYES_E_INVALID;
    }
    bayes_seq_model *m = new (std::nothrow) bayes_seq_model();
    if (!m) return BAYES_E_NOMEM;
    const double var = pow(sd, 2);
    const double norm_const = 1 / (sd * sqrt(2 * 3.14159265358979323846264338327950288));
    auto pmf = [&](unsigned fragment_length) { return norm_const * exp(-(pow((double)fragment_length - mean, 2) / (2 * var))); };
    const int n = max_transcript_length;
    m->frag_length_prob.assign(n, 0);
    m->frag_length_cum_prob.assign(n, 0);
    m->three_prime_cum_prob.assign(n, 0);
    m->frag_length_prob[0] = pmf(1);
    m->frag_length_cum_prob[0] = m->frag_length_prob[0];
    m->three_prime_cum_prob[0] = m->frag_length_prob[0];
    for (int i = 1; i < n; i++) {
        m->frag_length_prob[i] = pmf(i + 1);
        m->frag_length_cum_prob[i] = m->frag_length_prob[i] + m->frag_length_cum_prob[i - 1];
        m->three_prime_cum_prob[i] = m->frag_length_cum_prob[i] + m->three_prime_cum_prob[i - 1];
    }
    *out = m;
    return BAYES_OK;
}

void bayes_seq_model_destroy(bayes_seq_model *m) { delete m; }

// SequencingModel::calculateThreePrimeProb / calculateLengthProb / getEffectiveLength (sequencingModel.cpp:59-76)
double bayes_seq_model_three_prime_prob(const bayes_seq_model *m, int32_t three_prime_pos, int32_t tran_length) { return m->three_prime_prob(three_prime_pos, tran_length); }
double bayes_seq_model_length_prob(const bayes_seq_model *m, int32_t fragment_length, int32_t three_prime_pos) { return m->length_prob(fragment_length, three_prime_pos); }
double bayes_seq_model_effective_length(const bayes_seq_model *m, int32_t transcript_length) { return m->effective_length(transcript_length); }

// GaussianFragmentLengthModel::estimateParameters (fragmentLengthModel.cpp:71-161): median and median absolute deviation
// of a fragment length histogram; the model then uses mean = median, sd = 1.4826 * MAD (:56-59).  Fewer than 1000
// observations: BAYES_E_INVALID (the reference prints an error and exits).
int bayes_fragment_model_estimate(int32_t n, const uint32_t *length, const uint32_t *count, double *median_out, double *mad_out) {
    if (n < 1 || !length || !count || !median_out || !mad_out) {
        set_error("bayes_fragment_model_estimate: bad arguments");
        return BAYES_E_INVALID;
    }
    std::map<unsigned, unsigned> fragment_lengths;
    for (int i = 0; i < n; i++) fragment_lengths[length[i]] = count[i];
    bool ok = true;
    auto median_of = [&ok](const auto &observations) -> double {  // estimateMedian (:71-118)
        int observation_count = 0;
        for (const auto &kv : observations) observation_count += kv.second;
        if (observation_count == 0) { ok = false; return 0.0; }
        int median_observation_idx = observation_count % 2 == 0 ? (int)round((double)observation_count / 2.0) : (int)ceil((double)observation_count / 2.0);
        auto it = observations.begin();
        int running = it->second;
        while (running < median_observation_idx) {
            ++it;
            running += it->second;
        }
        if (observation_count % 2 == 0 && running == median_observation_idx) {
            auto median = it->first;
            ++it;
            median += it->first;
            return (double)median / 2;
        }
        return (double)it->first;
    };
    const double median = median_of(fragment_lengths);
    std::map<double, unsigned> abs_distances;
    int observation_count = 0;
    for (const auto &kv : fragment_lengths) {
        observation_count += kv.second;
        abs_distances[fabs((double)kv.first - median)] += kv.second;
    }
    const double mad = median_of(abs_distances);
    if (!ok || observation_count < 1000) {
        set_error("bayes_fragment_model_estimate: fewer than 1000 fragment length observations (use --frag-mean / --frag-sd)");
        return BAYES_E_INVALID;
    }
    *median_out = median;
    *mad_out = mad;
    return BAYES_OK;
}

static int check_raw(const bayes_raw_locus *L, const char *who) {
    if (!L || L->n_cand < 1 || L->n_frag < 0 || !L->exon_ptr || !L->exon_start || !L->exon_end || !L->cand_length || !L->is_premrna ||
        (L->n_frag > 0 && (!L->read_pos || !L->cigar_ptr || !L->cigar_op || !L->cigar_len || !L->frag_prob))) {
        set_error(std::string(who) + ": null pointer or empty candidate list");
        return BAYES_E_INVALID;
    }
    for (int i = 0; i < L->n_cand; i++)
        if (L->exon_ptr[i + 1] <= L->exon_ptr[i] || L->cand_length[i] < 1) {
            set_error(std::string(who) + ": candidate without exons or with length < 1");
            return BAYES_E_INVALID;
        }
    for (int i = 0; i < 2 * L->n_frag; i++) {
        const int c0 = L->cigar_ptr[i], c1 = L->cigar_ptr[i + 1];
        if (c1 <= c0 || L->cigar_op[c0] != 'M' || L->cigar_op[c1 - 1] != 'M') {  // reference: assert (:67-68)
            set_error(std::string(who) + ": every CIGAR must start and end with M");
            return BAYES_E_INVALID;
        }
    }
    // FragmentAlignment::probability is a product of base-call probabilities (assembler.cpp calculateSequencingProbability):
    // a number in [0, 1]; NaN or infinity would poison every row it is merged into
    for (int i = 0; i < L->n_frag; i++)
        if (!(L->frag_prob[i] >= 0.0) || !std::isfinite(L->frag_prob[i])) {
            set_error(std::string(who) + ": fragment probability is negative or not finite");
            return BAYES_E_INVALID;
        }
    return BAYES_OK;
}

// AlignmentParser::calculateFragTranProbabilities (alignmentParser.cpp:251-689).  Returns the CollapsedMap's return_code
// (0, 3 = every candidate excluded, 4 = no rows) or a negative BAYES_E_*.  subgraph_complexity (n_subgraphs entries, may
// be NULL) is decremented once per excluded candidate (:501-502).
int bayes_collapsed_map_build(const bayes_raw_locus *L, const bayes_seq_model *model, int32_t no_coverage_exclude,
                              int32_t *subgraph_complexity, bayes_collapsed_map **out) {
    int rc = check_raw(L, "bayes_collapsed_map_build");
    if (rc) return rc;
    if (!model || !out) {
        set_error("bayes_collapsed_map_build: null model / out");
        return BAYES_E_INVALID;
    }
    const int n_cand = L->n_cand;
    std::vector<Cand> cands(n_cand);
    for (int i = 0; i < n_cand; i++) {
        cands[i] = Cand{L->exon_start + L->exon_ptr[i], L->exon_end + L->exon_ptr[i], L->exon_ptr[i + 1] - L->exon_ptr[i], L->cand_length[i],
                        L->is_premrna[i] != 0};
        if (cands[i].length > model->max_len()) {
            set_error("bayes_collapsed_map_build: candidate longer than the sequencing model's tables");
            return BAYES_E_INVALID;
        }
    }
    bayes_collapsed_map *cm = new (std::nothrow) bayes_collapsed_map();
    if (!cm) return BAYES_E_NOMEM;
    const bool plus = L->strand_plus != 0;

    std::vector<char> excluded(n_cand, 0);
    int n_excluded = 0;
    std::vector<int> exclude_downstream(n_cand, -1);
    std::vector<double> exclude_upstream(n_cand, 2e9);  // the reference stores 2e9 in an int-valued map (:269); compared as a number
    std::vector<unsigned> candidate_counts(n_cand, 0);
    int64_t matched = 0, unmatched = 0, matches = 0, underflow_skipped = 0, internal_exclude = 0, no_map_exclude = 0, upstream_exclude = 0,
            downstream_exclude = 0;

    RowStore rows;
    std::vector<uint32_t> counts;
    std::vector<int32_t> tc;
    std::vector<double> tv;
    for (int f = 0; f < L->n_frag; f++) {
        tc.clear();
        tv.clear();
        double norm_const = 0;
        bool found = false;
        for (int i = 0; i < n_cand; i++) {
            if (excluded[i]) continue;
            FragmentMatch fm;
            bool bad = false;
            if (!match_fragment(*L, f, cands[i], plus, &fm, &bad)) {
                if (bad) {
                    delete cm;
                    set_error("bayes_collapsed_map_build: unhandled CIGAR symbol");
                    return BAYES_E_INVALID;
                }
                continue;
            }
            if (fm.fragment_length > cands[i].length || fm.three_prime_position > cands[i].length || fm.three_prime_position < 1 ||
                fm.fragment_length < 1) {  // reference: asserts (:213, :241-242); positions index the tables
                delete cm;
                set_error("bayes_collapsed_map_build: fragment longer than its candidate");
                return BAYES_E_INVALID;
            }
            found = true;
            matches++;
            candidate_counts[i]++;
            double p = 1;
            p *= L->frag_prob[f];
            p *= model->three_prime_prob(fm.three_prime_position, cands[i].length);
            p *= model->length_prob(fm.fragment_length, fm.three_prime_position);
            tc.push_back(i);
            tv.push_back(p);
            norm_const += p;
            if (!no_coverage_exclude) {  // coverage exclusion bookkeeping (:325-353)
                if (fm.first_exon_idx <= exclude_downstream[i] || candidate_counts[i] == 1) {
                    if (fm.last_exon_idx > exclude_downstream[i]) exclude_downstream[i] = fm.last_exon_idx;
                    if (fm.first_exon_idx < exclude_upstream[i]) exclude_upstream[i] = fm.first_exon_idx;
                } else if (!cands[i].premrna) {
                    excluded[i] = 1;
                    n_excluded++;
                    internal_exclude++;
                }
            }
        }
        if (!found) {
            unmatched++;
            continue;
        }
        matched++;
        if (norm_const < kDoubleUnderflow) {
            underflow_skipped++;
            continue;
        }
        // normalise (:369); cells that are exactly zero are not part of a sparse row
        int n = 0;
        for (size_t k = 0; k < tv.size(); k++) {
            const double v = tv[k] / norm_const;
            if (v != 0.0) {
                tc[n] = tc[k];
                tv[n] = v;
                n++;
            }
        }
        const int r = rows.find_row(tc.data(), tv.data(), n);
        if (r >= 0) {
            counts[r]++;
        } else {
            rows.append(tc.data(), tv.data(), n);
            counts.push_back(1);
        }
    }

    // exclusion on flank coverage and total fragment count (:425-467)
    if (!no_coverage_exclude) {
        for (int i = 0; i < n_cand; i++) {
            if (excluded[i]) continue;
            if (exclude_downstream[i] == -1) {
                no_map_exclude++;
                excluded[i] = 1;
                n_excluded++;
            } else if (!cands[i].premrna) {
                if (exclude_upstream[i] > 0) {
                    upstream_exclude++;
                    excluded[i] = 1;
                    n_excluded++;
                } else if (exclude_downstream[i] < cands[i].n_exons - 1) {
                    downstream_exclude++;
                    excluded[i] = 1;
                    n_excluded++;
                }
            }
        }
    }
    int64_t st[10] = {L->n_frag, matched, unmatched, 0, matches, 0, n_excluded, no_map_exclude, internal_exclude, upstream_exclude};
    (void)downstream_exclude;
    if (n_excluded == n_cand) {
        cm->return_code = 3;
        memcpy(cm->stats, st, sizeof(st));
        cm->stats[3] = underflow_skipped;
        *out = cm;
        return 3;
    }
    // column compaction and the index map (:484-560)
    std::vector<int32_t> new_idx(n_cand, -1);
    int T = 0;
    for (int i = 0; i < n_cand; i++) {
        if (excluded[i]) {
            if (subgraph_complexity && L->subgraph_id) subgraph_complexity[L->subgraph_id[i]]--;
            continue;
        }
        new_idx[i] = T++;
        cm->new_to_old.push_back(i);
        cm->efflen.push_back(model->effective_length(cands[i].length));
    }
    cm->T = T;

    // final normalisation over the kept candidates and second collapse (:572-619); skipped when nothing was excluded
    const int R0 = rows.rows();
    std::vector<char> row_excluded(R0, 0);
    RowStore fin;            // row i of `fin` is what the reference's matrix holds in row i (kept columns) after the loop
    if (n_excluded > 0) {
        for (int i = 0; i < R0; i++) {
            tc.clear();
            tv.clear();
            double row_sum = 0;
            for (int k = rows.ptr[i]; k < rows.ptr[i + 1]; k++) {
                const int c = new_idx[rows.col[k]];
                if (c < 0) continue;
                tc.push_back(c);
                tv.push_back(rows.val[k]);
                row_sum += rows.val[k];
            }
            if (row_sum < kDoubleUnderflow) {
                underflow_skipped += counts[i];
                row_excluded[i] = 1;
                fin.append(tc.data(), tv.data(), (int)tc.size());  // stays in the matrix with its old values
                continue;
            }
            std::vector<double> norm(tv.size());
            for (size_t k = 0; k < tv.size(); k++) norm[k] = tv[k] / row_sum;
            const int j = fin.find_row(tc.data(), norm.data(), (int)tc.size());
            if (j >= 0) {
                counts[j] += counts[i];  // also when row j itself was dropped: the reference loses these fragments (:610)
                row_excluded[i] = 1;
                fin.append(tc.data(), tv.data(), (int)tc.size());
            } else {
                fin.append(tc.data(), norm.data(), (int)tc.size());
            }
        }
    }
    // final container (:621-649), CSR over the non-zero cells
    cm->row_ptr.assign(1, 0);
    for (int i = 0; i < R0; i++) {
        if (row_excluded[i]) continue;
        if (n_excluded > 0) {
            for (int k = fin.ptr[i]; k < fin.ptr[i + 1]; k++) {
                cm->col_idx.push_back(fin.col[k]);
                cm->val.push_back(fin.val[k]);
            }
        } else {
            for (int k = rows.ptr[i]; k < rows.ptr[i + 1]; k++) {
                cm->col_idx.push_back(rows.col[k]);  // nothing excluded: new index == old index
                cm->val.push_back(rows.val[k]);
            }
        }
        cm->row_ptr.push_back((int32_t)cm->val.size());
        cm->row_count.push_back((int32_t)counts[i]);
    }
    memcpy(cm->stats, st, sizeof(st));
    cm->stats[3] = underflow_skipped;
    cm->stats[5] = (int64_t)cm->row_count.size();
    cm->return_code = cm->row_count.empty() ? 4 : 0;
    *out = cm;
    return cm->return_code;
}

void bayes_collapsed_map_destroy(bayes_collapsed_map *cm) { delete cm; }
int32_t bayes_collapsed_map_num_candidates(const bayes_collapsed_map *cm) { return cm ? cm->T : 0; }
int32_t bayes_collapsed_map_num_rows(const bayes_collapsed_map *cm) { return cm ? (int32_t)cm->row_count.size() : 0; }
int64_t bayes_collapsed_map_nnz(const bayes_collapsed_map *cm) { return cm ? (int64_t)cm->val.size() : 0; }

// CSR arrays (R + 1, nnz, nnz, R), new -> old candidate index and effective length (T each); any pointer may be NULL
int bayes_collapsed_map_export(const bayes_collapsed_map *cm, int32_t *row_ptr, int32_t *col_idx, double *val, int32_t *row_count,
                               int32_t *new_to_old, double *efflen) {
    if (!cm) return BAYES_E_INVALID;
    if (row_ptr) std::copy(cm->row_ptr.begin(), cm->row_ptr.end(), row_ptr);
    if (col_idx) std::copy(cm->col_idx.begin(), cm->col_idx.end(), col_idx);
    if (val) std::copy(cm->val.begin(), cm->val.end(), val);
    if (row_count) std::copy(cm->row_count.begin(), cm->row_count.end(), row_count);
    if (new_to_old) std::copy(cm->new_to_old.begin(), cm->new_to_old.end(), new_to_old);
    if (efflen) std::copy(cm->efflen.begin(), cm->efflen.end(), efflen);
    return BAYES_OK;
}

// the numbers of the reference's "BAM-FILE PARSING STATS" log block (:662-681): fragments parsed, matched, unmatched,
// removed by underflow, fragment-candidate matches, unique rows, candidates excluded, of which: no mapped fragments,
// internal coverage, upstream coverage (downstream = excluded - the other three)
int bayes_collapsed_map_stats(const bayes_collapsed_map *cm, int64_t *stats10) {
    if (!cm || !stats10) return BAYES_E_INVALID;
    memcpy(stats10, cm->stats, sizeof(cm->stats));
    return BAYES_OK;
}

// AlignmentParser::fetchFragmentLengths (alignmentParser.cpp:692-725): the implied fragment length of every
// (fragment, candidate) match, in fragment-major order.  Returns the number of matches (lengths_out may be NULL to count).
int64_t bayes_fetch_fragment_lengths(const bayes_raw_locus *L, uint32_t *lengths_out, int64_t cap) {
    if (check_raw(L, "bayes_fetch_fragment_lengths")) return BAYES_E_INVALID;
    const bool plus = L->strand_plus != 0;
    int64_t n = 0;
    for (int f = 0; f < L->n_frag; f++)
        for (int i = 0; i < L->n_cand; i++) {
            const Cand c{L->exon_start + L->exon_ptr[i], L->exon_end + L->exon_ptr[i], L->exon_ptr[i + 1] - L->exon_ptr[i], L->cand_length[i],
                         L->is_premrna[i] != 0};
            FragmentMatch fm;
            bool bad = false;
            if (match_fragment(*L, f, c, plus, &fm, &bad)) {
                if (lengths_out && n < cap) lengths_out[n] = (uint32_t)fm.fragment_length;
                n++;
            } else if (bad) {
                set_error("bayes_fetch_fragment_lengths: unhandled CIGAR symbol");
                return BAYES_E_INVALID;
            }
        }
    return n;
}

}  // extern "C"
