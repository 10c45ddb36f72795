// Host-side pieces of the hot path (K0 in SURVEY.md 2b): minimum read cover, simplex-size CDF table, job cost
// model and the cost-balanced partition of loci over GPUs.  Plain C++, no CUDA.
#include <float.h>
#include <math.h>

#include <algorithm>
#include <numeric>
#include <queue>

#include "internal.h"
#include "rng.cuh"

namespace bayes {

// AssignmentGenerator::calcMinimumReadCover (assignmentGenerator.cpp:52-91): greedy set cover of the rows by the
// columns of the support matrix (P >= double_underflow, :61).  Each round takes the column that covers most of the
// still-uncovered rows (:65-66); ties are broken by one uniform_int over the ascending list of maximisers
// (:69-82).  The reference recomputes a (1 x R)*(R x T) product per round; here the per-column cover counts are
// kept incrementally through a column-major copy of the support, which yields the same counts.
int min_read_cover(const bayes_locus_desc &in, uint64_t key, int32_t *in_cover) {
    const int T = in.T, R = in.R;
    std::vector<int32_t> col_ptr(T + 1, 0);
    for (int i = 0; i < R; i++)
        for (int k = in.row_ptr[i]; k < in.row_ptr[i + 1]; k++)
            if (in.val[k] >= DBL_MIN) col_ptr[in.col_idx[k] + 1]++;
    for (int t = 0; t < T; t++) col_ptr[t + 1] += col_ptr[t];
    std::vector<int32_t> col_rows(col_ptr[T]);
    {
        std::vector<int32_t> fill(col_ptr.begin(), col_ptr.end() - 1);
        for (int i = 0; i < R; i++)
            for (int k = in.row_ptr[i]; k < in.row_ptr[i + 1]; k++)
                if (in.val[k] >= DBL_MIN) col_rows[fill[in.col_idx[k]]++] = i;
    }
    std::vector<int32_t> covered(T);
    for (int t = 0; t < T; t++) covered[t] = col_ptr[t + 1] - col_ptr[t];
    std::vector<uint8_t> uncovered(R, 1);
    std::vector<int32_t> max_idx;
    max_idx.reserve(T);
    if (in_cover) std::fill(in_cover, in_cover + T, 0);
    int remaining = R, cover = 0;
    uint32_t round = 0;
    while (remaining > 0) {
        int max_val = 0;
        for (int t = 0; t < T; t++) max_val = std::max(max_val, covered[t]);
        if (max_val <= 0) return BAYES_E_INVALID;  // a row without support (reference: assert :67)
        max_idx.clear();
        for (int t = 0; t < T; t++)
            if (covered[t] == max_val) max_idx.push_back(t);
        const Site s = make_site(key, PH_COVER, 0, round++, 0);
        const int pick = max_idx[uniform_int(s, (uint32_t)max_idx.size())];
        if (in_cover) in_cover[pick] = 1;
        cover++;
        for (int p = col_ptr[pick]; p < col_ptr[pick + 1]; p++) {
            const int i = col_rows[p];
            if (!uncovered[i]) continue;
            uncovered[i] = 0;
            remaining--;
            for (int k = in.row_ptr[i]; k < in.row_ptr[i + 1]; k++)
                if (in.val[k] >= DBL_MIN) covered[in.col_idx[k]]--;
        }
    }
    return cover;
}

static inline bool double_compare(double a, double b) {  // utils.h:58-61
    return (a == b) || (fabs(a - b) < fabs(std::min(a, b)) * DBL_EPSILON * 100);
}

static inline double lgam(double x) {
    int sign;
    return lgamma_r(x, &sign);
}

// FixedBinomialFixedGammaSimplexSizeGenerator ctor (simplexSizeGenerator.cpp:117-187).  Row k-1 (k = |S+|) is the CDF
// over simplex sizes j = k..T of  log C(T-k, j-k) + j log pi + (T-j) log(1-pi) + lgamma(j g) - lgamma(N_f + j g),
// accumulated with log(1 + exp(.)) (:166) until the running sum stops moving under double_compare (:169-172) and
// exp-normalised by the final sum (:176-180).  The O(T^2) lgamma calls of the reference collapse to O(T) table
// look-ups of the same lgamma values, so every entry is bit-identical to evaluating the formula in place.
void simplex_table(double pi, double gamma, int T, int num_fragments, std::vector<int32_t> *row_off, std::vector<double> *table) {
    std::vector<double> lfact(T + 2), lt(T + 1);
    for (int m = 0; m <= T + 1; m++) lfact[m] = lgam((double)m);  // lfact[m] = lgamma(m)
    for (int j = 1; j <= T; j++) lt[j] = lgam(j * gamma) - lgam(num_fragments + j * gamma);
    const double lpi = log(pi), l1pi = log(1 - pi);
    row_off->assign(1, 0);
    table->clear();
    std::vector<double> row;
    for (int k = 1; k <= T; k++) {
        row.clear();
        double row_sum = 0.0 + (k * lpi + (T - k) * l1pi) + lt[k];
        row.push_back(row_sum);
        for (int j = k + 1; j <= T; j++) {
            const double cardinal = lfact[T - k + 1] - (lfact[j - k + 1] + lfact[T - j + 1]);
            const double prob_z = j * lpi + (T - j) * l1pi;
            const double prob_eq = cardinal + prob_z + lt[j];
            row_sum += log(1 + exp(prob_eq - row_sum));
            row.push_back(row_sum);
            if (double_compare(row[row.size() - 1], row[row.size() - 2])) break;
        }
        for (size_t j = 0; j < row.size(); j++) table->push_back(exp(row[j] - row_sum));
        row_off->push_back((int32_t)table->size());
    }
}

}  // namespace bayes

using namespace bayes;

extern "C" {

int bayes_min_read_cover(const bayes_locus_desc *locus, uint64_t key, int32_t *in_cover) {
    std::string err;
    if (!locus || validate_locus(*locus, &err) != 0) {
        set_error("bayes_min_read_cover: " + err);
        return BAYES_E_INVALID;
    }
    return min_read_cover(*locus, key, in_cover);
}

int64_t bayes_simplex_table(double pi, double gamma, int32_t T, int32_t num_fragments, int64_t *row_off, double *table) {
    if (T < 1 || !(pi > 0) || !(pi < 1) || !(gamma > 0)) {
        set_error("bayes_simplex_table: need T >= 1, 0 < pi < 1, gamma > 0");
        return BAYES_E_INVALID;
    }
    std::vector<int32_t> ro;
    std::vector<double> tv;
    simplex_table(pi, gamma, T, num_fragments, &ro, &tv);
    if (row_off)
        for (int i = 0; i <= T; i++) row_off[i] = ro[i];
    if (table) std::copy(tv.begin(), tv.end(), table);
    return (int64_t)tv.size();
}

// GibbsOutput::getTopMarginals (gibbsOutput.cpp:77-106)
int bayes_top_marginals(int32_t T, int32_t samples, const int32_t *occ, const double *mean, const double *m2,
                        double marginal_threshold, int32_t *idx_out, double *mean_out, double *sd_out) {
    const double count_threshold = marginal_threshold * samples;
    int n = 0;
    for (int i = 0; i < T; i++) {
        if (occ[i] >= count_threshold) {
            if (idx_out) idx_out[n] = i;
            if (mean_out) mean_out[n] = mean[i];
            if (sd_out) sd_out[n] = sqrt(m2[i] / (occ[i] - 1));
            n++;
        }
    }
    return n;
}

uint64_t bayes_chain_key(uint64_t seed, int32_t chain) {
    if (chain == 0) return seed;
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)chain;  // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

double bayes_locus_cost(const bayes_locus_desc *locus) {
    if (!locus || locus->R < 0) return 0;
    const double cells = locus->row_ptr ? (double)locus->row_ptr[locus->R] : 0;
    const double iters = (double)locus->burn_in + locus->samples;
    return iters * (cells + 4.0 * locus->R + 8.0 * locus->T) * (locus->n_chains > 0 ? locus->n_chains : 1);
}

// Longest-processing-time-first: items in descending cost, each to the currently lightest part.
int bayes_partition_lpt(int64_t n, const double *cost, int32_t n_parts, int32_t *part_out) {
    if (n < 0 || n_parts < 1 || !part_out || (n > 0 && !cost)) {
        set_error("bayes_partition_lpt: bad arguments");
        return BAYES_E_INVALID;
    }
    std::vector<int64_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return cost[a] > cost[b]; });
    typedef std::pair<double, int32_t> Load;  // (load, part); min-heap, ties to the lowest part index
    std::priority_queue<Load, std::vector<Load>, std::greater<Load> > heap;
    for (int32_t p = 0; p < n_parts; p++) heap.push(Load(0.0, p));
    for (int64_t i : order) {
        Load l = heap.top();
        heap.pop();
        part_out[i] = l.second;
        l.first += cost[i];
        heap.push(l);
    }
    return BAYES_OK;
}

void bayes_philox4x32(uint64_t key, const uint32_t ctr[4], uint32_t out[4]) {
    const u32x4 r = philox4x32_10((uint32_t)key, (uint32_t)(key >> 32), ctr[0], ctr[1], ctr[2], ctr[3]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

}  // extern "C"
