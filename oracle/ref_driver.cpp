/*
 * oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C-callable driver around the UNMODIFIED reference sampler translation units
 * (valueContainer, expressionGenerator, simplexSizeGenerator, gammaGenerator,
 * gibbsOutput, assignmentGenerator, gibbsSampler from /root/reference/src),
 * compiled against the stand-in headers in oracle/shim/ into
 * oracle/_ref/libbayesref.so by oracle/Makefile.  No reference source is copied.
 *
 * ref_run_locus() reproduces the hot-path call site of
 * Assembler::bayesemblerCallback, assembler.cpp:1548-1609 (generators on the
 * stack, min read cover -> pi, burn-in/sample counts supplied by the caller,
 * runSampler, accumulator read-back).  The per-function entry points exist so
 * that the C restatement (oracle/bayes_oracle.c) can be pinned function by
 * function, and ref_run_batch() is the CPU baseline: a pool of worker threads
 * popping loci from a queue like assembler.cpp:1422-1439 / 2017-2026.
 */
#include <chrono>
#include <cstring>
#include <mutex>
#include <sstream>
#include <stdint.h>
#include <thread>
#include <vector>

/* read-back of GibbsOutput's accumulators and of the simplex CDF table, which the
 * reference only exposes as formatted text; access control does not change layout */
#define private public
#include <gibbsOutput.h>
#include <simplexSizeGenerator.h>
#undef private
#include <gibbsSampler.h>

namespace {

struct DenseLocus {
    Eigen::MatrixXd P;
    std::vector<int> counts;
    tr1::unordered_map<int, pair<int, double> > idx_map;
    int num_fragments;
};

void build_dense(int R, int T, const int *row_ptr, const int *col_idx, const double *val, const int *row_count,
                 const double *efflen, DenseLocus *d) {
    d->P = Eigen::MatrixXd::Zero(R, T);
    for (int i = 0; i < R; i++)
        for (int k = row_ptr[i]; k < row_ptr[i + 1]; k++) d->P(i, col_idx[k]) = val[k];
    d->counts.assign(row_count, row_count + R);
    d->num_fragments = 0;
    for (int i = 0; i < R; i++) d->num_fragments += row_count[i];
    for (int t = 0; t < T; t++) d->idx_map[t] = pair<int, double>(t, efflen[t]);
}

/* assembler.cpp:1548-1603 */
void run_dense(DenseLocus &d, int T, double adjusted_fragment_count, double gamma, uint32_t seed, int burn, int samples,
               int *occ, double *mean, double *m2, int *cover_size, double *pi_out, uint64_t *words) {
    boost::random::mt19937 mt_rng;
    mt_rng.seed(seed);
    stringstream log_stream;

    AssignmentGenerator assignment_generator(&mt_rng, &d.P, &d.counts);
    assignment_generator.calcMinimumReadCover(log_stream);
    double msc_to_all_ratio = double(assignment_generator.getMinimumReadCoverSize()) / T;
    if (msc_to_all_ratio > double_almost_one) msc_to_all_ratio = double_almost_one;

    FixedBinomialFixedGammaSimplexSizeGenerator fixed_simplex_size_generator(msc_to_all_ratio, gamma, T, &mt_rng, d.num_fragments);
    FixedGammaGenerator fixed_gamma_generator(gamma);
    ExpressionGenerator expression_generator(d.num_fragments, T, &mt_rng);
    GibbsSampler gibbs_sampler(&fixed_simplex_size_generator, &fixed_gamma_generator, &expression_generator,
                               &assignment_generator, d.idx_map, adjusted_fragment_count, d.num_fragments, T, &mt_rng, false);
    GibbsOutput gibbs_output(samples, T);
    gibbs_sampler.runSampler(&gibbs_output, burn, samples, log_stream);

    for (int t = 0; t < T; t++) {
        if (occ) occ[t] = gibbs_output.marginal_occurence(t);
        if (mean) mean[t] = gibbs_output.marginal_mean_cumulant(t);
        if (m2) m2[t] = gibbs_output.marginal_var_cumulant(t);
    }
    if (cover_size) *cover_size = assignment_generator.getMinimumReadCoverSize();
    if (pi_out) *pi_out = msc_to_all_ratio;
    if (words) *words = mt_rng.words_consumed();
}

} // namespace

extern "C" {

int ref_run_locus(int R, int T, const int *row_ptr, const int *col_idx, const double *val, const int *row_count,
                  const double *efflen, double adjusted_fragment_count, double gamma, uint32_t seed, int burn, int samples,
                  int *occ, double *mean, double *m2, int *cover_size, double *pi_out, uint64_t *words) {
    DenseLocus d;
    build_dense(R, T, row_ptr, col_idx, val, row_count, efflen, &d);
    run_dense(d, T, adjusted_fragment_count, gamma, seed, burn, samples, occ, mean, m2, cover_size, pi_out, words);
    return 0;
}

/* AssignmentGenerator::calcMinimumReadCover, assignmentGenerator.cpp:52-96 */
int ref_min_read_cover(int R, int T, const int *row_ptr, const int *col_idx, const double *val, const int *row_count,
                       uint32_t seed, uint64_t *words) {
    std::vector<double> efflen(T, 1.0);
    DenseLocus d;
    build_dense(R, T, row_ptr, col_idx, val, row_count, efflen.data(), &d);
    boost::random::mt19937 mt_rng;
    mt_rng.seed(seed);
    stringstream log_stream;
    AssignmentGenerator ag(&mt_rng, &d.P, &d.counts);
    ag.calcMinimumReadCover(log_stream);
    if (words) *words = mt_rng.words_consumed();
    return ag.getMinimumReadCoverSize();
}

/* AssignmentGenerator::initAssignment, assignmentGenerator.cpp:213-260 */
int ref_init_assignment(int R, int T, const int *row_ptr, const int *col_idx, const double *val, const int *row_count,
                        uint32_t seed, int *counts_out, uint64_t *words) {
    std::vector<double> efflen(T, 1.0);
    DenseLocus d;
    build_dense(R, T, row_ptr, col_idx, val, row_count, efflen.data(), &d);
    boost::random::mt19937 mt_rng;
    mt_rng.seed(seed);
    stringstream log_stream;
    AssignmentGenerator ag(&mt_rng, &d.P, &d.counts);
    CountValueContainer c = ag.initAssignment(log_stream);
    for (int t = 0; t < T; t++) counts_out[t] = c.getCount(t);
    if (words) *words = mt_rng.words_consumed();
    return c.getPlusSize();
}

/* AssignmentGenerator::generateAssignment, assignmentGenerator.cpp:263-373 */
int ref_generate_assignment(int R, int T, const int *row_ptr, const int *col_idx, const double *val, const int *row_count,
                            const double *theta, uint32_t seed, int *counts_out, uint64_t *words) {
    std::vector<double> efflen(T, 1.0);
    DenseLocus d;
    build_dense(R, T, row_ptr, col_idx, val, row_count, efflen.data(), &d);
    boost::random::mt19937 mt_rng;
    mt_rng.seed(seed);
    AssignmentGenerator ag(&mt_rng, &d.P, &d.counts);
    ExpressionValueContainer e(T);
    for (int t = 0; t < T; t++)
        if (theta[t] > 0) { e.setValue(theta[t], t); e.setBinaryOn(t); e.addToPlus(t); }
    CountValueContainer c = ag.generateAssignment(e);
    for (int t = 0; t < T; t++) counts_out[t] = c.getCount(t);
    if (words) *words = mt_rng.words_consumed();
    return c.getPlusSize();
}

/* ExpressionGenerator::generateExpression, expressionGenerator.cpp:77-132; counts_in must have >= 1 positive entry */
int ref_generate_expression(int T, const int *counts_in, int num_fragments, int b, double gamma, uint32_t seed,
                            double *theta_out, int *plus_order_out, uint64_t *words) {
    boost::random::mt19937 mt_rng;
    mt_rng.seed(seed);
    CountValueContainer c(T);
    for (int t = 0; t < T; t++)
        if (counts_in[t] > 0) { c.addToCount(counts_in[t], t); c.addToPlus(t); }
    c.fetchNullSet();
    ExpressionGenerator eg(num_fragments, T, &mt_rng);
    ExpressionValueContainer e = eg.generateExpression(b, gamma, c);
    for (int t = 0; t < T; t++) theta_out[t] = e.getValue(t);
    if (plus_order_out) for (int i = 0; i < e.getPlusSize(); i++) plus_order_out[i] = e.getPlus(i);
    if (words) *words = mt_rng.words_consumed();
    return e.getPlusSize();
}

/* FixedBinomialFixedGammaSimplexSizeGenerator ctor, simplexSizeGenerator.cpp:117-187.
 * table_out is T x T row-major (row k-1 = |S+| = k), row_len_out[T]. */
int ref_simplex_table(double pi, double gamma, int T, int num_fragments, double *table_out, int *row_len_out) {
    boost::random::mt19937 mt_rng;
    FixedBinomialFixedGammaSimplexSizeGenerator g(pi, gamma, T, &mt_rng, num_fragments);
    for (int i = 0; i < T; i++) {
        row_len_out[i] = (int)g.simplex_prob_matrix[i].size();
        for (size_t j = 0; j < g.simplex_prob_matrix[i].size(); j++) table_out[(size_t)i * T + j] = g.simplex_prob_matrix[i][j];
    }
    return 0;
}

/* ...::sampleSimplexSize via initSimplexSize, simplexSizeGenerator.cpp:201-224 */
int ref_sample_simplex(double pi, double gamma, int T, int num_fragments, int count_plus_size, uint32_t seed) {
    boost::random::mt19937 mt_rng;
    mt_rng.seed(seed);
    FixedBinomialFixedGammaSimplexSizeGenerator g(pi, gamma, T, &mt_rng, num_fragments);
    return g.initSimplexSize(count_plus_size, gamma).first;
}

/* GibbsOutput::getTopMarginals, gibbsOutput.cpp:77-106, on caller-supplied accumulators */
int ref_top_marginals(int T, int samples, const int *occ, const double *mean, const double *m2, double threshold,
                      int *idx_out, double *mean_out, double *sd_out) {
    GibbsOutput go(samples, T);
    for (int t = 0; t < T; t++) {
        go.marginal_occurence(t) = occ[t];
        go.marginal_mean_cumulant(t) = mean[t];
        go.marginal_var_cumulant(t) = m2[t];
    }
    Ensemble e = go.getTopMarginals(threshold);
    for (size_t i = 0; i < e.indices.size(); i++) {
        idx_out[i] = e.indices[i];
        mean_out[i] = e.means(i);
        sd_out[i] = e.sd(i);
    }
    return (int)e.indices.size();
}

/* GibbsOutput::addSample (gibbsOutput.cpp:51-75) n_add times, then getTopMarginals + writeEnsembleGtf (mode 0,
 * gibbsOutput.cpp:109-141) or writeCandidateGtf (mode 1, :143-166) on candidates built from flat arrays.
 * Returns the text length, or -(needed) when cap is too small. */
int ref_gtf(int T, int samples, int n_add, const int *mask, const double *values, int n_cand, const int *n_vert, const int *starts,
            const int *ends, const int *premrna, const int *old_idx, const double *efflen, int total_num_fragments,
            double count_threshold, double marginal_threshold, int mode, char *out, int cap) {
    GibbsOutput go(samples, T);
    for (int s = 0; s < n_add; s++) {
        vector<int> plus;
        for (int t = T - 1; t >= 0; t--)
            if (mask[(size_t)s * T + t]) plus.push_back(t);
        go.addSample(plus, vector<double>(values + (size_t)s * T, values + (size_t)(s + 1) * T));
    }
    tr1::unordered_map<int, pair<int, double> > idx_map;
    for (int t = 0; t < T; t++) idx_map[t] = pair<int, double>(old_idx[t], efflen[t]);
    vector<Candidate> cands(n_cand);
    int at = 0;
    for (int i = 0; i < n_cand; i++) {
        stringstream nm;
        nm << "graph1_cand" << i;
        cands[i].idx = i;
        cands[i].name = nm.str();
        cands[i].graph_name = "graph1";
        cands[i].length = 0;
        cands[i].isPreMrna = premrna[i] != 0;
        for (int j = 0; j < n_vert[i]; j++, at++) {
            Vertex v;
            v.reference = "chr1";
            v.strand = (i & 1) ? "-" : "+";
            v.start = starts[at];
            v.end = ends[at];
            cands[i].vertices.push_back(v);
            cands[i].length += ends[at] - starts[at] + 1;
        }
    }
    string text = mode == 0 ? go.writeEnsembleGtf(go.getTopMarginals(marginal_threshold), cands, idx_map, total_num_fragments, count_threshold)
                            : go.writeCandidateGtf(cands, idx_map, total_num_fragments);
    if ((int)text.size() + 1 > cap) return -(int)text.size() - 1;
    memcpy(out, text.c_str(), text.size() + 1);
    return (int)text.size();
}

/*
 * CPU baseline: n_loci loci, concatenated CSR (loc_ptr[n+1] into rows, nnz_ptr[n+1] into cells, t_ptr[n+1] into
 * efflen / outputs), worker pool of n_threads.  Dense matrices are built before the clock starts (the reference's
 * CollapsedMap arrives dense); returns wall seconds of the sampler section alone.
 */
double ref_run_batch(int n_loci, const int *T_arr, const int64_t *row_off, const int64_t *cell_off, const int64_t *t_off,
                     const int *row_ptr_cat, const int *col_idx_cat, const double *val_cat, const int *row_count_cat,
                     const double *efflen_cat, double adjusted_fragment_count, double gamma, const uint32_t *seeds,
                     const int *burn, const int *samples, int n_threads, int *occ_cat, double *mean_cat, double *m2_cat) {
    std::vector<DenseLocus> dense(n_loci);
    for (int l = 0; l < n_loci; l++) {
        int R = (int)(row_off[l + 1] - row_off[l]);
        /* per-locus row_ptr is stored with R+1 entries at offset row_off[l] + l */
        build_dense(R, T_arr[l], row_ptr_cat + row_off[l] + l, col_idx_cat + cell_off[l], val_cat + cell_off[l],
                    row_count_cat + row_off[l], efflen_cat + t_off[l], &dense[l]);
    }
    std::mutex lock;
    int next = 0;
    auto worker = [&]() {
        for (;;) {
            int l;
            {
                std::lock_guard<std::mutex> g(lock);
                if (next >= n_loci) return;
                l = next++;
            }
            run_dense(dense[l], T_arr[l], adjusted_fragment_count, gamma, seeds[l], burn[l], samples[l],
                      occ_cat ? occ_cat + t_off[l] : 0, mean_cat ? mean_cat + t_off[l] : 0, m2_cat ? m2_cat + t_off[l] : 0, 0, 0, 0);
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int i = 0; i < n_threads; i++) pool.emplace_back(worker);
    for (auto &t : pool) t.join();
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

} // extern "C"
