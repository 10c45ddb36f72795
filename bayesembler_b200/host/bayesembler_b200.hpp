// Host-side mirror of the reference's sampler interface on top of the C ABI (include/bayesembler_b200.h).
//
// Same class names, constructor argument order and meaning, and error behaviour as the reference objects built at the
// hot-path call site, assembler.cpp:1548-1609, so that Assembler::bayesemblerCallback keeps its shape:
//
//   reference (src/)                                                      here
//   AssignmentGenerator(mt_rng_pt_t, Eigen::MatrixXd*, vector<int>*)      AssignmentGenerator(mt_rng_pt_t, const Matrix*, const vector<int>*)
//     (assignmentGenerator.h:48)                                            Matrix: anything with rows(), cols(), operator()(i, j) -- Eigen::MatrixXd works
//   FixedBinomialFixedGammaSimplexSizeGenerator(pi, gamma, T, rng, N_f)   same (simplexSizeGenerator.h:67)
//   FixedGammaGenerator(gamma)                                            same (gammaGenerator.h:71)
//   ExpressionGenerator(N_f, T, rng)                                      same (expressionGenerator.h:45)
//   GibbsSampler(ssg*, gg*, eg*, ag*, idx->(old idx, efflen), N_total,    same (gibbsSampler.h:51); `int` truncation of N_total kept
//                N_f, T, rng, assignment_minimum)
//   GibbsOutput(samples, T), getTopMarginals, writeEnsembleGtf,           same (gibbsOutput.h:46-50); Ensemble holds std::vector<double>
//   writeCandidateGtf                                                      instead of Eigen::VectorXd
//   AllPaths(dot).findFullPaths(no_pre_mrna, max_candidates, log)         same (allPaths.h:95-98); malformed DOT throws instead of asserting
//   void GibbsSampler::runSampler(GibbsOutput*, burn, samples, log)       same, synchronous, ONE locus on the GPU (correct, latency-bound);
//                                                                         GibbsBatch::add(...) + run() is the batched form the GPU wants
//
// The one type that changes is the random engine: the reference hands every generator a boost::random::mt19937*
// seeded with graph_seed (assembler.cpp:1437-1445); a sequential engine cannot feed a GPU, so mt_rng_t here only
// carries the seed and the device derives counter-based Philox streams from it.
//
// Header-only; link against bayesembler_b200/libbayesembler_b200.so.  There is no CPU sampler behind these classes:
// without a CUDA device runSampler / GibbsBatch::run throw std::runtime_error carrying bayes_last_error().
#ifndef BAYESEMBLER_B200_HPP
#define BAYESEMBLER_B200_HPP

#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/bayesembler_b200.h"

namespace bayesembler_b200 {

// ---- utils.h:85-136 ------------------------------------------------------------------------------------------------
struct Ensemble {  // utils.h:91-97
    std::vector<int> indices;
    std::vector<double> means;
    std::vector<double> sd;
};
struct Vertex {  // utils.h:100-106
    std::string reference;
    std::string strand;
    int start;
    int end;
};
struct Candidate {  // utils.h:109-117
    int idx;
    std::string name;
    std::string graph_name;
    int length;
    std::vector<Vertex> vertices;
    bool isPreMrna;
};
typedef std::map<int, std::pair<int, double> > IdxMap;  // new idx -> (old idx, effective length); reference: tr1::unordered_map

// Stand-in for boost::random::mt19937: carries graph_seed only (see header comment).
struct mt_rng_t {
    uint64_t graph_seed;
    mt_rng_t() : graph_seed(5489u) {}
    explicit mt_rng_t(uint64_t s) : graph_seed(s) {}
    void seed(uint64_t s) { graph_seed = s; }
};
typedef mt_rng_t *mt_rng_pt_t;

inline void check(int rc, const char *what) {
    if (rc < 0) throw std::runtime_error(std::string(what) + ": " + bayes_last_error());
}

// ---- assignmentGenerator.h -------------------------------------------------------------------------------------------
class AssignmentGenerator {
   public:
    template <class Matrix>
    AssignmentGenerator(mt_rng_pt_t rng, const Matrix *collapsed_map, const std::vector<int> *collapsed_counts)
        : mt_rng_pt(rng), num_rows((int)collapsed_map->rows()), num_transcripts((int)collapsed_map->cols()), cover_size(-1) {
        // dense (zeros = incompatible, utils.h:131) -> CSR over the non-zero cells, columns ascending
        row_ptr.assign(1, 0);
        for (int i = 0; i < num_rows; i++) {
            for (int j = 0; j < num_transcripts; j++) {
                const double v = (*collapsed_map)(i, j);
                if (v != 0.0) {
                    col_idx.push_back(j);
                    val.push_back(v);
                }
            }
            row_ptr.push_back((int32_t)val.size());
        }
        row_count.assign(collapsed_counts->begin(), collapsed_counts->end());
    }
    // CSR form, for callers that never materialise the dense matrix
    AssignmentGenerator(mt_rng_pt_t rng, int T, const std::vector<int32_t> &row_ptr_in, const std::vector<int32_t> &col_idx_in,
                        const std::vector<double> &val_in, const std::vector<int32_t> &counts)
        : mt_rng_pt(rng), num_rows((int)counts.size()), num_transcripts(T), cover_size(-1), row_ptr(row_ptr_in), col_idx(col_idx_in),
          val(val_in), row_count(counts) {}

    // assignmentGenerator.cpp:52-91
    void calcMinimumReadCover(std::stringstream &log_stream) {
        bayes_locus_desc d = desc();
        const int n = bayes_min_read_cover(&d, bayes_chain_key(mt_rng_pt->graph_seed, 0), 0);
        check(n, "calcMinimumReadCover");
        cover_size = n;
        log_stream << "\nCalculating minimum read cover" << std::endl;                                               // :54
        log_stream << "Number of candidates in minimum read cover: " << cover_size << "\n" << std::endl;              // :90
    }
    int getMinimumReadCoverSize() const { return cover_size; }  // :93-96

    bayes_locus_desc desc() const {
        bayes_locus_desc d;
        d.T = num_transcripts;
        d.R = num_rows;
        d.row_ptr = row_ptr.data();
        d.col_idx = col_idx.data();
        d.val = val.data();
        d.row_count = row_count.data();
        d.efflen = 0;
        d.total_num_fragments = 1;
        d.gamma = 1;
        d.pi = -1;
        d.burn_in = 0;
        d.samples = 1;
        d.seed = mt_rng_pt->graph_seed;
        d.n_chains = 1;
        d.reserved = 0;
        return d;
    }
    int getNumFragments() const {
        long n = 0;
        for (size_t i = 0; i < row_count.size(); i++) n += row_count[i];
        return (int)n;
    }

   private:
    mt_rng_pt_t mt_rng_pt;
    int num_rows, num_transcripts, cover_size;
    std::vector<int32_t> row_ptr, col_idx;
    std::vector<double> val;
    std::vector<int32_t> row_count;
};

// ---- simplexSizeGenerator.h ------------------------------------------------------------------------------------------
class SimplexSizeGenerator {
   public:
    virtual std::string getParameterString() = 0;
    virtual double getPi() const = 0;
    bool isFixed() const { return true; }
    virtual ~SimplexSizeGenerator() {}
};

class FixedBinomialFixedGammaSimplexSizeGenerator : public SimplexSizeGenerator {
   public:
    FixedBinomialFixedGammaSimplexSizeGenerator(double pi_in, double gamma_in, int num_transcripts_in, mt_rng_pt_t, int num_fragments_in)
        : pi(pi_in), gamma(gamma_in), num_transcripts(num_transcripts_in), num_fragments(num_fragments_in) {}
    std::string getParameterString() {  // simplexSizeGenerator.cpp:189-199
        std::stringstream s;
        s << "FixedBinomialFixedGamma simplex size generator (pi=" << pi << ", gamma=" << gamma << ")";
        return s.str();
    }
    double getPi() const { return pi; }
    // the CDF table of the ctor (simplexSizeGenerator.cpp:117-187), row k-1 = |S+| = k
    std::vector<std::vector<double> > simplexProbMatrix() const {
        std::vector<int64_t> off(num_transcripts + 1);
        const int64_t n = bayes_simplex_table(pi, gamma, num_transcripts, num_fragments, 0, 0);
        check((int)n, "bayes_simplex_table");
        std::vector<double> tab(n);
        bayes_simplex_table(pi, gamma, num_transcripts, num_fragments, off.data(), tab.data());
        std::vector<std::vector<double> > rows(num_transcripts);
        for (int k = 0; k < num_transcripts; k++) rows[k].assign(tab.begin() + off[k], tab.begin() + off[k + 1]);
        return rows;
    }

   private:
    double pi, gamma;
    int num_transcripts, num_fragments;
};

// ---- gammaGenerator.h --------------------------------------------------------------------------------------------------
class GammaGenerator {
   public:
    virtual double initGamma() = 0;
    bool isFixed() const { return true; }
    virtual ~GammaGenerator() {}
};
class FixedGammaGenerator : public GammaGenerator {  // gammaGenerator.cpp:159-171
   public:
    explicit FixedGammaGenerator(double gamma_in) : gamma(gamma_in) {}
    double initGamma() { return gamma; }

   private:
    double gamma;
};

// ---- expressionGenerator.h ---------------------------------------------------------------------------------------------
class ExpressionGenerator {
   public:
    ExpressionGenerator(int num_fragments_in, int num_transcripts_in, mt_rng_pt_t) : num_transcripts(num_transcripts_in), num_fragments(num_fragments_in) {}
    int numTranscripts() const { return num_transcripts; }
    int numFragments() const { return num_fragments; }

   private:
    int num_transcripts, num_fragments;
};

// ---- gibbsOutput.h -----------------------------------------------------------------------------------------------------
class GibbsOutput {
    int num_iterations, num_transcripts;

   public:
    GibbsOutput(int num_iterations_in, int num_transcripts_in)  // gibbsOutput.cpp:40-49
        : num_iterations(num_iterations_in), num_transcripts(num_transcripts_in), marginal_occurence(num_transcripts_in, 0),
          marginal_mean_cumulant(num_transcripts_in, 0.0), marginal_var_cumulant(num_transcripts_in, 0.0) {}

    // gibbsOutput.cpp:51-75 (host form; the device runs the same recurrence fused into its E-step)
    void addSample(std::vector<int> plus, const std::vector<double> &values) {
        for (size_t i = 0; i < plus.size(); i++) {
            const int idx = plus[i];
            marginal_occurence[idx]++;
            if (marginal_occurence[idx] == 1) {
                marginal_mean_cumulant[idx] = values[idx];
            } else {
                const double delta = values[idx] - marginal_mean_cumulant[idx];
                marginal_mean_cumulant[idx] += delta / marginal_occurence[idx];
                marginal_var_cumulant[idx] += delta * (values[idx] - marginal_mean_cumulant[idx]);
            }
        }
    }

    Ensemble getTopMarginals(double marginal_threshold) const {  // gibbsOutput.cpp:77-106
        Ensemble e;
        e.indices.resize(num_transcripts);
        e.means.resize(num_transcripts);
        e.sd.resize(num_transcripts);
        const int n = bayes_top_marginals(num_transcripts, num_iterations, marginal_occurence.data(), marginal_mean_cumulant.data(),
                                          marginal_var_cumulant.data(), marginal_threshold, e.indices.data(), e.means.data(), e.sd.data());
        e.indices.resize(n);
        e.means.resize(n);
        e.sd.resize(n);
        return e;
    }

    // gibbsOutput.cpp:109-141.  total_num_fragments is an int here exactly as in the reference signature.
    std::string writeEnsembleGtf(const Ensemble &top_ensemble, const std::vector<Candidate> &candidates, const IdxMap &idx_map,
                                 int total_num_fragments, double count_threshold) const {
        std::stringstream out;
        for (size_t i = 0; i < top_ensemble.indices.size(); i++) {
            const int new_idx = top_ensemble.indices[i];
            const std::pair<int, double> &m = idx_map.at(new_idx);
            const double expected_count = top_ensemble.means[i] * m.second * total_num_fragments / 1e9;
            if (expected_count < count_threshold) continue;
            const Candidate &c = candidates[m.first];
            if (c.idx != m.first) throw std::logic_error("writeEnsembleGtf: candidate index mismatch");  // reference: assert :127
            if (c.isPreMrna) continue;
            for (size_t j = 0; j < c.vertices.size(); j++)
                exon_line(out, c, c.vertices[j], (double)marginal_occurence[new_idx] / num_iterations, top_ensemble.means[i],
                          top_ensemble.sd[i], expected_count);
        }
        return out.str();
    }

    // gibbsOutput.cpp:143-166
    std::string writeCandidateGtf(const std::vector<Candidate> &candidates, const IdxMap &idx_map, int total_num_fragments) const {
        std::stringstream out;
        for (int i = 0; i < num_transcripts; i++) {
            const std::pair<int, double> &m = idx_map.at(i);
            const double expected_count = marginal_mean_cumulant[i] * m.second * total_num_fragments / 1e9;
            const Candidate &c = candidates[m.first];
            if (c.idx != m.first) throw std::logic_error("writeCandidateGtf: candidate index mismatch");
            for (size_t j = 0; j < c.vertices.size(); j++)
                exon_line(out, c, c.vertices[j], (double)marginal_occurence[i] / num_iterations, marginal_mean_cumulant[i],
                          sqrt(marginal_var_cumulant[i] / (marginal_occurence[i] - 1)), expected_count);
        }
        return out.str();
    }

    int numIterations() const { return num_iterations; }
    int numTranscripts() const { return num_transcripts; }
    std::vector<int32_t> marginal_occurence;
    std::vector<double> marginal_mean_cumulant;
    std::vector<double> marginal_var_cumulant;

   private:
    static void exon_line(std::stringstream &out, const Candidate &c, const Vertex &v, double confidence, double fpkm, double fpkm_sd,
                          double expected_count) {
        out << v.reference << "\tBayesembler\texon\t" << v.start << "\t" << v.end << "\t.\t" << v.strand << "\t.\tgene_id \"" << c.graph_name
            << "\"; transcript_id \"" << c.name << "\"; transcript_confidence \"" << confidence << "\"; FPKM \"" << fpkm << "\"; FPKM_sd \""
            << fpkm_sd << "\"; expected_count \"" << expected_count << "\";" << std::endl;
    }
};

// ---- gibbsSampler.h ----------------------------------------------------------------------------------------------------
class GibbsSampler {
   public:
    GibbsSampler(SimplexSizeGenerator *ssg, GammaGenerator *gg, ExpressionGenerator *eg, AssignmentGenerator *ag, const IdxMap &idx_map,
                 int total_num_fragments_in, int num_fragments_in, int num_transcripts_in, mt_rng_pt_t rng, bool assignment_minimum_in)
        : simplex_size_generator(ssg), gamma_generator(gg), expression_generator(eg), assignment_generator(ag),
          new_idx_to_old_idx_effective_length(idx_map), total_num_fragments(total_num_fragments_in), num_fragments(num_fragments_in),
          num_transcripts(num_transcripts_in), mt_rng_pt(rng), assignment_minimum(assignment_minimum_in) {
        if (assignment_minimum) throw std::invalid_argument("GibbsSampler: assignment_minimum is unreachable in v1.2.0 (assembler.cpp:1577) and not built");
        efflen.resize(num_transcripts);
        for (int t = 0; t < num_transcripts; t++) efflen[t] = idx_map.at(t).second;  // gibbsSampler.cpp:73-78
    }

    // bayes_locus_desc of this sampler: what the reference's objects at assembler.cpp:1548-1577 hold between them
    bayes_locus_desc desc(int burn_in, int samples, int n_chains = 1) const {
        bayes_locus_desc d = assignment_generator->desc();
        d.efflen = efflen.data();
        d.total_num_fragments = (double)total_num_fragments;
        d.gamma = gamma_generator->initGamma();
        d.pi = simplex_size_generator->getPi();
        d.burn_in = burn_in;
        d.samples = samples;
        d.seed = mt_rng_pt->graph_seed;
        d.n_chains = n_chains;
        return d;
    }

    // gibbsSampler.cpp:55-148 for ONE locus on device `device`: correct, but a single dependent chain cannot fill a GPU --
    // use GibbsBatch wherever the caller can defer reading the GibbsOutput.
    void runSampler(GibbsOutput *gibbs_output, int burn_in, int samples, std::stringstream &log_stream, int device = 0) const {
        bayes_gibbs_ctx *ctx = 0;
        check(bayes_gibbs_create(device, &ctx), "bayes_gibbs_create");
        bayes_locus_desc d = desc(burn_in, samples);
        int64_t id = 0;
        int rc = bayes_gibbs_run_batch(ctx, 1, &d, &id);
        if (rc == 0)
            rc = bayes_gibbs_fetch(ctx, id, 0, gibbs_output->marginal_occurence.data(), gibbs_output->marginal_mean_cumulant.data(),
                                   gibbs_output->marginal_var_cumulant.data(), 0);
        std::string err = rc ? bayes_last_error() : "";
        bayes_gibbs_destroy(ctx);
        if (rc) throw std::runtime_error("runSampler: " + err);
        writeSamplerLog(log_stream, num_transcripts, burn_in, samples);
    }

    // The lines runSampler writes to the graph's log (gibbsSampler.cpp:62, :99-143; assignmentGenerator.cpp:215), without the
    // time stamps: on the GPU the iterations of all graphs run as one batch, so the text is written once the batch is back.
    static void writeSamplerLog(std::ostream &log_stream, int num_transcripts, int burn_in, int samples) {
        if (num_transcripts == 1) {
            log_stream << "\nSingle candidate - skipping gibbs sampling" << std::endl;
            return;
        }
        const int print_frequency = 2000;
        log_stream << "Initialising read assignments uniformly" << std::endl;
        log_stream << "\nPerforming " << burn_in << " burn-in iterations" << std::endl;
        for (int i = print_frequency; i < burn_in; i += print_frequency) log_stream << "Completed " << i << " burn-in iterations" << std::endl;
        log_stream << "Completed " << burn_in << " burn-in iterations" << std::endl;
        log_stream << "\nPerforming " << samples << " gibbs iterations" << std::endl;
        for (int i = print_frequency; i < samples; i += print_frequency) log_stream << "Completed " << i << " gibbs iterations" << std::endl;
        log_stream << "Completed " << samples << " gibbs iterations" << std::endl;
    }

    int numTranscripts() const { return num_transcripts; }

   private:
    SimplexSizeGenerator *simplex_size_generator;
    GammaGenerator *gamma_generator;
    ExpressionGenerator *expression_generator;
    AssignmentGenerator *assignment_generator;
    IdxMap new_idx_to_old_idx_effective_length;
    int total_num_fragments;  // int on purpose: gibbsSampler.h:51 truncates the caller's double
    int num_fragments;
    int num_transcripts;
    mt_rng_pt_t mt_rng_pt;
    bool assignment_minimum;
    std::vector<double> efflen;
};

// The batched form: the reference's N worker threads each run one locus start to finish (assembler.cpp:2017-2026); a GPU
// wants all loci of a pass at once.  add() copies what the sampler needs (the generator objects may die afterwards),
// run() executes every queued sampler in one submit/upload/run/download and fills the GibbsOutputs.
class GibbsBatch {
   public:
    explicit GibbsBatch(int device = 0) : ctx(0) { check(bayes_gibbs_create(device, &ctx), "bayes_gibbs_create"); }
    ~GibbsBatch() { bayes_gibbs_destroy(ctx); }
    void add(const GibbsSampler &sampler, GibbsOutput *out, int burn_in, int samples) {
        bayes_locus_desc d = sampler.desc(burn_in, samples);
        int64_t id = 0;
        check(bayes_gibbs_submit(ctx, 1, &d, &id), "bayes_gibbs_submit");  // packs now; the caller's buffers are free again
        outputs.push_back(out);
    }
    void run() {
        check(bayes_gibbs_upload(ctx), "bayes_gibbs_upload");
        check(bayes_gibbs_run(ctx), "bayes_gibbs_run");
        check(bayes_gibbs_download(ctx), "bayes_gibbs_download");
        check(bayes_gibbs_sync(ctx), "bayes_gibbs_sync");
        for (size_t i = 0; i < outputs.size(); i++)
            check(bayes_gibbs_fetch(ctx, (int64_t)i, 0, outputs[i]->marginal_occurence.data(), outputs[i]->marginal_mean_cumulant.data(),
                                    outputs[i]->marginal_var_cumulant.data(), 0),
                  "bayes_gibbs_fetch");
        check(bayes_gibbs_clear(ctx), "bayes_gibbs_clear");
        outputs.clear();
    }
    size_t size() const { return outputs.size(); }

   private:
    GibbsBatch(const GibbsBatch &);
    GibbsBatch &operator=(const GibbsBatch &);
    bayes_gibbs_ctx *ctx;
    std::vector<GibbsOutput *> outputs;
};


// =====================================================================================================================
// Likelihood builder: the reference classes that produce the sampler's input, on the C ABI's host
// functions (bayes_seq_model_*, bayes_collapsed_map_*).  Same constructor / method signatures as the reference.
// =====================================================================================================================
namespace BamTools {  // utils.h:63 only needs CigarOp from BamTools
struct CigarOp {
    char Type;
    uint32_t Length;
    CigarOp(char t = 0, uint32_t l = 0) : Type(t), Length(l) {}
};
}  // namespace BamTools

typedef std::pair<unsigned, std::vector<BamTools::CigarOp> > ReadData;  // utils.h:63
struct FragmentAlignment {                                               // utils.h:73-77
    std::pair<ReadData, ReadData> read_data;
    double probability;
};

// fragmentLengthModel.h:41-69
class FragmentLengthModel {
   public:
    virtual double pmf(unsigned) = 0;
    virtual ~FragmentLengthModel() {}
};
class GaussianFragmentLengthModel : public FragmentLengthModel {
   public:
    GaussianFragmentLengthModel() : mean(0), sd(1) {}
    GaussianFragmentLengthModel(double mean_in, double sd_in) : mean(mean_in), sd(sd_in) {}   // fragmentLengthModel.cpp:40-49
    explicit GaussianFragmentLengthModel(const std::map<unsigned, unsigned> &fragment_lengths) {  // :52-64
        std::pair<double, double> p = estimateParameters(fragment_lengths);
        mean = p.first;
        sd = 1.4826 * p.second;
    }
    std::pair<double, double> estimateParameters(const std::map<unsigned, unsigned> &fragment_lengths) const {  // :120-161
        std::vector<uint32_t> len, cnt;
        for (std::map<unsigned, unsigned>::const_iterator it = fragment_lengths.begin(); it != fragment_lengths.end(); ++it) {
            len.push_back(it->first);
            cnt.push_back(it->second);
        }
        double median = 0, mad = 0;
        check(bayes_fragment_model_estimate((int32_t)len.size(), len.data(), cnt.data(), &median, &mad), "estimateParameters");
        return std::make_pair(median, mad);
    }
    double pmf(unsigned fragment_length) {  // :66-69
        const double var = pow(sd, 2), norm_const = 1 / (sd * sqrt(2 * 3.14159265358979323846264338327950288));
        return norm_const * exp(-(pow((double)fragment_length - mean, 2) / (2 * var)));
    }
    double getMean() const { return mean; }
    double getSd() const { return sd; }

   private:
    double mean, sd;
};

// sequencingModel.h:43-59; only the Gaussian model is wired into the CLI (assembler.cpp:1984-1987)
class SequencingModel {
   public:
    SequencingModel(FragmentLengthModel *fragment_length_model_pt, int max_transcript_length, std::stringstream &log_stream) : model(0) {
        GaussianFragmentLengthModel *g = dynamic_cast<GaussianFragmentLengthModel *>(fragment_length_model_pt);
        if (!g) throw std::invalid_argument("SequencingModel: only GaussianFragmentLengthModel is supported");
        check(bayes_seq_model_create_gaussian(g->getMean(), g->getSd(), max_transcript_length, &model), "bayes_seq_model_create_gaussian");
        log_stream << "\nTabulating fragment length and 3'-position probabilities up to transcript lengths of " << max_transcript_length
                   << std::endl;  // sequencingModel.cpp:36
    }
    ~SequencingModel() { bayes_seq_model_destroy(model); }
    double calculateThreePrimeProb(int three_prime_pos, int tran_length) { return bayes_seq_model_three_prime_prob(model, three_prime_pos, tran_length); }
    double calculateLengthProb(int fragment_length, int three_prime_pos) { return bayes_seq_model_length_prob(model, fragment_length, three_prime_pos); }
    double getEffectiveLength(int transcript_length) { return bayes_seq_model_effective_length(model, transcript_length); }
    const bayes_seq_model *handle() const { return model; }

   private:
    SequencingModel(const SequencingModel &);
    SequencingModel &operator=(const SequencingModel &);
    bayes_seq_model *model;
};

// utils.h:129-136 with the probability matrix in CSR form (the reference's is a dense Eigen::MatrixXd); it offers the
// rows() / cols() / (i, j) accessors AssignmentGenerator's templated constructor reads, and AssignmentGenerator also
// takes it directly.
struct CollapsedMap {
    struct Csr {
        int T;
        std::vector<int32_t> row_ptr, col_idx;
        std::vector<double> val;
        int rows() const { return (int)row_ptr.size() - 1; }
        int cols() const { return T; }
        double operator()(int i, int j) const {
            const int32_t *b = col_idx.data() + row_ptr[i], *e = col_idx.data() + row_ptr[i + 1];
            const int32_t *p = std::lower_bound(b, e, (int32_t)j);
            return p != e && *p == j ? val[p - col_idx.data()] : 0.0;
        }
    } probability_matrix;
    std::vector<int> counts;
    IdxMap new_idx_to_old_idx_effective_length;
    int return_code;
};

// alignmentParser.h:44-63
class AlignmentParser {
   public:
    // alignmentParser.cpp:251-689; like the reference it deletes the FragmentAlignment objects it is given
    CollapsedMap calculateFragTranProbabilities(std::vector<Candidate> &candidates, std::vector<FragmentAlignment *> *fragment_alignments,
                                                SequencingModel *sequencing_model, bool no_coverage_exclude, std::stringstream &log_stream,
                                                std::map<std::string, int> *subgraph_complexity) {
        Flat f(candidates, fragment_alignments);
        std::vector<int32_t> complexity(f.graph_names.size(), 0);
        for (size_t g = 0; g < f.graph_names.size(); g++) complexity[g] = (*subgraph_complexity)[f.graph_names[g]];
        bayes_collapsed_map *cm = 0;
        const int rc = bayes_collapsed_map_build(&f.raw, sequencing_model->handle(), no_coverage_exclude ? 1 : 0, complexity.data(), &cm);
        check(rc, "calculateFragTranProbabilities");
        CollapsedMap out;
        out.return_code = rc;
        out.probability_matrix.T = bayes_collapsed_map_num_candidates(cm);
        int64_t st[10];
        bayes_collapsed_map_stats(cm, st);
        if (rc == 0) {
            const int R = bayes_collapsed_map_num_rows(cm), T = out.probability_matrix.T;
            const int64_t nnz = bayes_collapsed_map_nnz(cm);
            out.probability_matrix.row_ptr.resize(R + 1);
            out.probability_matrix.col_idx.resize(nnz);
            out.probability_matrix.val.resize(nnz);
            std::vector<int32_t> counts(R), new_to_old(T);
            std::vector<double> efflen(T);
            bayes_collapsed_map_export(cm, out.probability_matrix.row_ptr.data(), out.probability_matrix.col_idx.data(), out.probability_matrix.val.data(),
                                       counts.data(), new_to_old.data(), efflen.data());
            out.counts.assign(counts.begin(), counts.end());
            for (int t = 0; t < T; t++) out.new_idx_to_old_idx_effective_length[t] = std::make_pair((int)new_to_old[t], efflen[t]);
            for (size_t g = 0; g < f.graph_names.size(); g++) (*subgraph_complexity)[f.graph_names[g]] = complexity[g];
        } else {
            out.probability_matrix.row_ptr.assign(1, 0);
        }
        bayes_collapsed_map_destroy(cm);
        // the graph's log, as the reference words it (alignmentParser.cpp:273, :478, :484, :659, :668-684), without the time stamps
        log_stream << "Matching fragments with candidates" << std::endl;
        if (rc == 3) {
            log_stream << "\nNo candidates had sufficient fragment coverage to continue - exiting!" << std::endl;
        } else {
            log_stream << "Excluding " << st[6] << " of " << candidates.size() << " candidates and renormalising conditional probability matrix " << std::endl;
            if (rc == 4) {
                log_stream << "\nNo rows in conditional probability matrix - exiting!" << std::endl;
            } else {
                log_stream << "\n=== BAM-FILE PARSING STATS ===\n" << std::endl;
                log_stream << "  " << st[0] << " fragment(s) were parsed in total." << std::endl;
                log_stream << "\n  - Fragment-candidate match stats:" << std::endl;
                log_stream << "  " << st[1] << " fragment(s) matched a transcript." << std::endl;
                log_stream << "  " << st[2] << " fragment(s) dit not match a transcript." << std::endl;
                log_stream << "  " << st[3] << " fragments were removed due to probability underflow." << std::endl;
                log_stream << "  " << st[4] << " fragment-candidate match(es) were found." << std::endl;
                log_stream << "  " << st[5] << " unique rows in probability matrix." << std::endl;
                log_stream << "\n  - Candidate exclusion stats:" << std::endl;
                log_stream << "  " << st[6] << " candidates were excluded due to low fragment coverage." << std::endl;
                log_stream << "  " << st[7] << " candidates were excluded due to lack of mapped fragments." << std::endl;
                log_stream << "  " << st[8] << " candidates were excluded due to lack of internal fragment coverage." << std::endl;
                log_stream << "  " << st[9] << " candidates were excluded due to lack of upstream fragment coverage." << std::endl;
                log_stream << "  " << st[6] - st[7] - st[8] - st[9] << " candidates were excluded due to lack of downstream fragment coverage." << std::endl;
                log_stream << std::endl;
            }
        }
        for (size_t i = 0; i < fragment_alignments->size(); i++) delete (*fragment_alignments)[i];
        return out;
    }

    // alignmentParser.cpp:692-725
    void fetchFragmentLengths(std::map<unsigned, unsigned> *fragment_length_map, std::vector<Candidate> &candidates,
                              std::vector<FragmentAlignment *> *fragment_alignments, std::stringstream &) {
        Flat f(candidates, fragment_alignments);
        const int64_t n = bayes_fetch_fragment_lengths(&f.raw, 0, 0);
        check((int)n, "fetchFragmentLengths");
        std::vector<uint32_t> lengths(n);
        bayes_fetch_fragment_lengths(&f.raw, lengths.data(), n);
        for (int64_t i = 0; i < n; i++) (*fragment_length_map)[lengths[i]]++;
        for (size_t i = 0; i < fragment_alignments->size(); i++) delete (*fragment_alignments)[i];
    }

   private:
    // vector<Candidate> + vector<FragmentAlignment*> -> the flat arrays of bayes_raw_locus
    struct Flat {
        bayes_raw_locus raw;
        std::vector<int32_t> exon_ptr, exon_start, exon_end, cand_length, is_premrna, subgraph_id, cigar_ptr;
        std::vector<uint32_t> read_pos, cigar_len;
        std::vector<char> cigar_op;
        std::vector<double> frag_prob;
        std::vector<std::string> graph_names;
        Flat(const std::vector<Candidate> &candidates, const std::vector<FragmentAlignment *> *frags) {
            exon_ptr.push_back(0);
            for (size_t i = 0; i < candidates.size(); i++) {
                const Candidate &c = candidates[i];
                for (size_t k = 0; k < c.vertices.size(); k++) {
                    exon_start.push_back(c.vertices[k].start);
                    exon_end.push_back(c.vertices[k].end);
                }
                exon_ptr.push_back((int32_t)exon_start.size());
                cand_length.push_back(c.length);
                is_premrna.push_back(c.isPreMrna ? 1 : 0);
                size_t g = 0;
                while (g < graph_names.size() && graph_names[g] != c.graph_name) g++;
                if (g == graph_names.size()) graph_names.push_back(c.graph_name);
                subgraph_id.push_back((int32_t)g);
            }
            cigar_ptr.push_back(0);
            for (size_t i = 0; i < frags->size(); i++) {
                const FragmentAlignment &f = *(*frags)[i];
                for (int m = 0; m < 2; m++) {
                    const ReadData &rd = m == 0 ? f.read_data.first : f.read_data.second;
                    read_pos.push_back(rd.first);
                    for (size_t k = 0; k < rd.second.size(); k++) {
                        cigar_op.push_back(rd.second[k].Type);
                        cigar_len.push_back(rd.second[k].Length);
                    }
                    cigar_ptr.push_back((int32_t)cigar_op.size());
                }
                frag_prob.push_back(f.probability);
            }
            raw.n_cand = (int32_t)candidates.size();
            raw.exon_ptr = exon_ptr.data();
            raw.exon_start = exon_start.data();
            raw.exon_end = exon_end.data();
            raw.cand_length = cand_length.data();
            raw.is_premrna = is_premrna.data();
            raw.subgraph_id = subgraph_id.data();
            raw.n_subgraphs = (int32_t)graph_names.size();
            raw.strand_plus = !candidates.empty() && !candidates.front().vertices.empty() && candidates.front().vertices.front().strand == "+" ? 1 : 0;
            raw.n_frag = (int32_t)frags->size();
            raw.read_pos = read_pos.data();
            raw.cigar_ptr = cigar_ptr.data();
            raw.cigar_op = cigar_op.data();
            raw.cigar_len = cigar_len.data();
            raw.frag_prob = frag_prob.data();
        }
    };
};

// allPaths.h:43-98
class AllPaths {
   public:
    // allPaths.cpp:42-49: the reference parses the DOT text here and aborts (assert) on a malformed one; the text is kept
    // and parsed by findFullPaths, which throws std::runtime_error instead
    explicit AllPaths(std::string dotfile_in) : dot(dotfile_in), junction_threshold(1) {}

    // allPaths.cpp:116-304.  Same candidates, names, order and log lines (without the reference's time stamps).
    std::vector<Candidate> findFullPaths(bool no_pre_mrna, int max_candidate_number, std::stringstream &log_stream) {
        bayes_path_set *ps = 0;
        check(bayes_all_paths_build(dot.c_str(), no_pre_mrna ? 1 : 0, max_candidate_number, &ps), "findFullPaths");
        int32_t st[6];
        bayes_path_set_stats(ps, st);
        const int n = bayes_path_set_num_candidates(ps), ne = bayes_path_set_num_exons(ps);
        std::vector<int32_t> exon_ptr(n + 1), exon_start(ne), exon_end(ne), length(n), pre(n);
        bayes_path_set_export(ps, exon_ptr.data(), exon_start.data(), exon_end.data(), length.data(), pre.data());
        const std::string graph_name = bayes_path_set_graph_name(ps), reference = bayes_path_set_reference(ps), strand = bayes_path_set_strand(ps);
        log_stream << "\nNumber of vertices in graph: " << st[0] << std::endl;
        log_stream << "Number of sources in graph: " << st[1] << std::endl;
        log_stream << "Number of sinks in graph: " << st[2] << std::endl;
        for (int t = 1; t <= st[3]; t++)
            log_stream << "Too many candidates to continue - adjusting junction threshold from " << t << " to " << t + 1 << std::endl;
        junction_threshold = st[5];
        std::vector<Candidate> out(n);
        std::vector<char> name(graph_name.size() + 32);
        for (int i = 0; i < n; i++) {
            Candidate &c = out[i];
            c.idx = i;
            bayes_path_set_candidate_name(ps, i, name.data(), (int32_t)name.size());
            c.name = name.data();
            c.graph_name = graph_name;
            c.length = length[i];
            c.isPreMrna = pre[i] != 0;
            for (int k = exon_ptr[i]; k < exon_ptr[i + 1]; k++) {
                Vertex v;
                v.reference = reference;
                v.strand = strand;
                v.start = exon_start[k];
                v.end = exon_end[k];
                c.vertices.push_back(v);
            }
        }
        bayes_path_set_destroy(ps);
        log_stream << "Number of paths in splice-graph (transcripts): " << n << std::endl;
        return out;
    }

    int junctionThreshold() const { return junction_threshold; }  // private in the reference; exposed for logs / tests

   private:
    std::string dot;
    int junction_threshold;
};

}  // namespace bayesembler_b200

#endif  // BAYESEMBLER_B200_HPP
