// C-callable test driver around the C++ host facade (bayesembler_b200/host/bayesembler_b200.hpp); built by
// tests/test_host_facade.py with g++ and linked against libbayesembler_b200.so.  facade_run_loci() is written the way
// Assembler::bayesemblerCallback builds and runs its sampler objects (assembler.cpp:1548-1609).
#include <stdlib.h>
#include <string.h>

#include "../../bayesembler_b200/host/bayesembler_b200.hpp"

using namespace bayesembler_b200;

namespace {

// dense matrix with the accessors AssignmentGenerator's templated ctor needs (what Eigen::MatrixXd offers)
struct Dense {
    int R, T;
    std::vector<double> a;
    Dense(int R_, int T_) : R(R_), T(T_), a((size_t)R_ * T_, 0.0) {}
    int rows() const { return R; }
    int cols() const { return T; }
    double operator()(int i, int j) const { return a[(size_t)j * R + i]; }
    double &operator()(int i, int j) { return a[(size_t)j * R + i]; }
};

std::vector<Candidate> make_candidates(int n_cand, const int *n_vert, const int *starts, const int *ends, const int *premrna) {
    std::vector<Candidate> c(n_cand);
    int at = 0;
    for (int i = 0; i < n_cand; i++) {
        std::stringstream nm;
        nm << "graph1_cand" << i;
        c[i].idx = i;
        c[i].name = nm.str();
        c[i].graph_name = "graph1";
        c[i].length = 0;
        c[i].isPreMrna = premrna[i] != 0;
        for (int j = 0; j < n_vert[i]; j++, at++) {
            Vertex v;
            v.reference = "chr1";
            v.strand = (i & 1) ? "-" : "+";
            v.start = starts[at];
            v.end = ends[at];
            c[i].vertices.push_back(v);
            c[i].length += ends[at] - starts[at] + 1;
        }
    }
    return c;
}

}  // namespace

extern "C" {

// GibbsOutput: n_add addSample calls (mask/values rows of length T), then getTopMarginals + writeEnsembleGtf (mode 0) or
// writeCandidateGtf (mode 1).  Returns the text length (or -needed when cap is too small).
int facade_gtf(int T, int samples, int n_add, const int *mask, const double *values, int n_cand, const int *n_vert, const int *starts,
               const int *ends, const int *premrna, const int *old_idx, const double *efflen, int total_num_fragments,
               double count_threshold, double marginal_threshold, int mode, char *out, int cap) {
    GibbsOutput go(samples, T);
    for (int s = 0; s < n_add; s++) {
        std::vector<int> plus;
        for (int t = T - 1; t >= 0; t--)
            if (mask[(size_t)s * T + t]) plus.push_back(t);
        go.addSample(plus, std::vector<double>(values + (size_t)s * T, values + (size_t)(s + 1) * T));
    }
    IdxMap idx_map;
    for (int t = 0; t < T; t++) idx_map[t] = std::make_pair(old_idx[t], efflen[t]);
    const std::vector<Candidate> cands = make_candidates(n_cand, n_vert, starts, ends, premrna);
    const std::string text = mode == 0 ? go.writeEnsembleGtf(go.getTopMarginals(marginal_threshold), cands, idx_map, total_num_fragments, count_threshold)
                                       : go.writeCandidateGtf(cands, idx_map, total_num_fragments);
    if ((int)text.size() + 1 > cap) return -(int)text.size() - 1;
    memcpy(out, text.c_str(), text.size() + 1);
    return (int)text.size();
}

// n loci (concatenated CSR, per-locus row_ptr with R+1 entries at row_off[l] + l), each driven like
// assembler.cpp:1548-1609.  batched = 0: GibbsSampler::runSampler per locus; 1: GibbsBatch.  Returns 0 or -1 (message in err).
int facade_run_loci(int n, const int *T_arr, const int64_t *row_off, const int64_t *cell_off, const int64_t *t_off, const int *row_ptr,
                    const int *col_idx, const double *val, const int *row_count, const double *efflen, double adjusted_fragment_count,
                    double gamma, const uint64_t *seeds, const int *burn, const int *samples, int batched, int *occ, double *mean,
                    double *m2, double *pi_out, char *err, int err_cap) {
    try {
        GibbsBatch *batch = batched ? new GibbsBatch(0) : 0;
        std::vector<GibbsOutput *> outs;
        for (int l = 0; l < n; l++) {
            const int T = T_arr[l], R = (int)(row_off[l + 1] - row_off[l]);
            const int *rp = row_ptr + row_off[l] + l;
            Dense P(R, T);
            for (int i = 0; i < R; i++)
                for (int k = rp[i]; k < rp[i + 1]; k++) P(i, col_idx[cell_off[l] + k]) = val[cell_off[l] + k];
            std::vector<int> counts(row_count + row_off[l], row_count + row_off[l + 1]);
            IdxMap idx_map;
            for (int t = 0; t < T; t++) idx_map[t] = std::make_pair(t, efflen[t_off[l] + t]);
            int num_fragments = 0;
            for (int i = 0; i < R; i++) num_fragments += counts[i];
            std::stringstream log;

            mt_rng_t mt_rng;
            mt_rng.seed(seeds[l]);
            AssignmentGenerator assignment_generator(&mt_rng, &P, &counts);
            assignment_generator.calcMinimumReadCover(log);
            double msc_to_all_ratio = double(assignment_generator.getMinimumReadCoverSize()) / T;
            if (msc_to_all_ratio > 1.0 - 2.220446049250313e-16) msc_to_all_ratio = 1.0 - 2.220446049250313e-16;
            FixedBinomialFixedGammaSimplexSizeGenerator fixed_simplex_size_generator(msc_to_all_ratio, gamma, T, &mt_rng, num_fragments);
            FixedGammaGenerator fixed_gamma_generator(gamma);
            ExpressionGenerator expression_generator(num_fragments, T, &mt_rng);
            GibbsSampler gibbs_sampler(&fixed_simplex_size_generator, &fixed_gamma_generator, &expression_generator, &assignment_generator,
                                       idx_map, adjusted_fragment_count, num_fragments, T, &mt_rng, false);
            GibbsOutput *gibbs_output = new GibbsOutput(samples[l], T);
            outs.push_back(gibbs_output);
            if (batch) batch->add(gibbs_sampler, gibbs_output, burn[l], samples[l]);
            else gibbs_sampler.runSampler(gibbs_output, burn[l], samples[l], log);
            if (pi_out) pi_out[l] = msc_to_all_ratio;
        }
        if (batch) {
            batch->run();
            delete batch;
        }
        for (int l = 0; l < n; l++) {
            for (int t = 0; t < T_arr[l]; t++) {
                occ[t_off[l] + t] = outs[l]->marginal_occurence[t];
                mean[t_off[l] + t] = outs[l]->marginal_mean_cumulant[t];
                m2[t_off[l] + t] = outs[l]->marginal_var_cumulant[t];
            }
            delete outs[l];
        }
        return 0;
    } catch (const std::exception &e) {
        if (err && err_cap > 0) {
            strncpy(err, e.what(), err_cap - 1);
            err[err_cap - 1] = 0;
        }
        return -1;
    }
}

// Likelihood builder through the facade classes, the way assembler.cpp:1511-1514 calls them; same argument list as
// ref_collapsed_map (oracle/ref_f1_driver.cpp).
int facade_collapsed_map(int n_cand, const int *exon_ptr, const int *exon_start, const int *exon_end, const int *cand_length,
                         const int *is_premrna, const int *subgraph_id, int strand_plus, int n_frag, const uint32_t *read_pos,
                         const int *cigar_ptr, const char *cigar_op, const uint32_t *cigar_len, const double *frag_prob, double frag_mean,
                         double frag_sd, int max_transcript_length, int no_coverage_exclude, int n_subgraphs, int *subgraph_complexity,
                         int r_cap, int *T_out, int *R_out, double *P_out, int *counts_out, int *new_to_old, double *efflen_out) {
    std::vector<Candidate> candidates(n_cand);
    for (int i = 0; i < n_cand; i++) {
        Candidate &c = candidates[i];
        std::stringstream nm, gn;
        nm << "cand" << i;
        gn << "g" << subgraph_id[i];
        c.idx = i;
        c.name = nm.str();
        c.graph_name = gn.str();
        c.length = cand_length[i];
        c.isPreMrna = is_premrna[i] != 0;
        for (int k = exon_ptr[i]; k < exon_ptr[i + 1]; k++) {
            Vertex v;
            v.reference = "chr1";
            v.strand = strand_plus ? "+" : "-";
            v.start = exon_start[k];
            v.end = exon_end[k];
            c.vertices.push_back(v);
        }
    }
    std::vector<FragmentAlignment *> frags(n_frag);
    for (int i = 0; i < n_frag; i++) {
        FragmentAlignment *f = new FragmentAlignment;
        for (int m = 0; m < 2; m++) {
            ReadData &rd = m == 0 ? f->read_data.first : f->read_data.second;
            rd.first = read_pos[2 * i + m];
            for (int k = cigar_ptr[2 * i + m]; k < cigar_ptr[2 * i + m + 1]; k++) rd.second.push_back(BamTools::CigarOp(cigar_op[k], cigar_len[k]));
        }
        f->probability = frag_prob[i];
        frags[i] = f;
    }
    std::map<std::string, int> complexity;
    for (int g = 0; g < n_subgraphs; g++) {
        std::stringstream gn;
        gn << "g" << g;
        complexity[gn.str()] = subgraph_complexity[g];
    }
    std::stringstream log;
    GaussianFragmentLengthModel flm(frag_mean, frag_sd);
    SequencingModel sm(&flm, max_transcript_length, log);
    AlignmentParser parser;
    CollapsedMap cm = parser.calculateFragTranProbabilities(candidates, &frags, &sm, no_coverage_exclude != 0, log, &complexity);
    for (int g = 0; g < n_subgraphs; g++) {
        std::stringstream gn;
        gn << "g" << g;
        subgraph_complexity[g] = complexity[gn.str()];
    }
    *T_out = 0;
    *R_out = 0;
    if (cm.return_code != 0) return cm.return_code;
    const int R = cm.probability_matrix.rows(), T = cm.probability_matrix.cols();
    *T_out = T;
    *R_out = R;
    if (R > r_cap) return -1;
    for (int i = 0; i < R; i++) {
        counts_out[i] = cm.counts[i];
        for (int j = 0; j < T; j++) P_out[(size_t)i * T + j] = cm.probability_matrix(i, j);
    }
    for (int t = 0; t < T; t++) {
        new_to_old[t] = cm.new_idx_to_old_idx_effective_length[t].first;
        efflen_out[t] = cm.new_idx_to_old_idx_effective_length[t].second;
    }
    // the CollapsedMap plugs into the sampler classes as the reference's does (assembler.cpp:1548)
    mt_rng_t rng(1);
    AssignmentGenerator ag(&rng, &cm.probability_matrix, &cm.counts);
    return ag.getNumFragments() > 0 ? 0 : -2;
}

// the CDF table as the facade's FixedBinomialFixedGammaSimplexSizeGenerator exposes it (row lengths only + flat values)
int facade_simplex_rows(double pi, double gamma, int T, int num_fragments, int *row_len, double *flat) {
    mt_rng_t rng;
    FixedBinomialFixedGammaSimplexSizeGenerator g(pi, gamma, T, &rng, num_fragments);
    std::vector<std::vector<double> > rows = g.simplexProbMatrix();
    int at = 0;
    for (int k = 0; k < T; k++) {
        row_len[k] = (int)rows[k].size();
        for (size_t j = 0; j < rows[k].size(); j++) flat[at++] = rows[k][j];
    }
    return at;
}

// AllPaths of the facade driven like assembler.cpp:1462-1474, flattened with the signature of oracle/ref_f2_driver.cpp's
// ref_all_paths so that the same Python binding reads both
int facade_all_paths(const char *dot, int no_pre_mrna, int max_candidate_number, int cap_cand, int cap_exons, int *exon_ptr, int *exon_start,
                     int *exon_end, int *cand_length, int *is_premrna, int *cand_idx, char *names, int names_cap, char *graph_name,
                     int graph_name_cap, int *stats5) {
    std::stringstream log;
    std::vector<Candidate> cands;
    try {
        AllPaths paths((std::string(dot)));
        cands = paths.findFullPaths(no_pre_mrna != 0, max_candidate_number, log);
    } catch (const std::runtime_error &) {
        return -2;
    }
    if ((int)cands.size() > cap_cand) return -1;
    int ne = 0;
    std::string all_names;
    exon_ptr[0] = 0;
    graph_name[0] = 0;
    for (size_t i = 0; i < cands.size(); i++) {
        if (ne + (int)cands[i].vertices.size() > cap_exons) return -1;
        for (size_t j = 0; j < cands[i].vertices.size(); j++, ne++) {
            exon_start[ne] = cands[i].vertices[j].start;
            exon_end[ne] = cands[i].vertices[j].end;
        }
        exon_ptr[i + 1] = ne;
        cand_length[i] = cands[i].length;
        is_premrna[i] = cands[i].isPreMrna ? 1 : 0;
        cand_idx[i] = cands[i].idx;
        all_names += cands[i].name + "\n";
        if (i == 0) strncpy(graph_name, cands[i].graph_name.c_str(), graph_name_cap - 1), graph_name[graph_name_cap - 1] = 0;
    }
    if ((int)all_names.size() + 1 > names_cap) return -1;
    memcpy(names, all_names.c_str(), all_names.size() + 1);
    // the numbers of the log block, read back from the text the facade wrote
    const char *keys[5] = {"Number of vertices in graph: ", "Number of sources in graph: ", "Number of sinks in graph: ", 0,
                           "Number of paths in splice-graph (transcripts): "};
    for (int k = 0; k < 5; k++) stats5[k] = 0;
    std::string line;
    while (std::getline(log, line)) {
        for (int k = 0; k < 5; k++)
            if (keys[k] && line.find(keys[k]) != std::string::npos) stats5[k] = atoi(line.c_str() + line.find(keys[k]) + strlen(keys[k]));
        if (line.find("Too many candidates to continue") != std::string::npos) stats5[3]++;
    }
    return (int)cands.size();
}

// GibbsOutput with GIVEN accumulators -> getTopMarginals + writeEnsembleGtf (mode 0) / writeCandidateGtf (mode 1), candidates flat
// as in bayes_path_set_export with their names '\n'-joined.  Returns the text length, -1 when cap is too small, -5 / -7 for the
// soft codes of assembler.cpp:1609-1625 (mode 0 only).
int facade_gtf_from_accumulators(int T, int samples, const int *occ, const double *mean, const double *m2, int n_cand, const int *exon_ptr,
                                 const int *starts, const int *ends, const int *premrna, const char *names, const char *graph_name,
                                 const char *reference, const char *strand, const int *old_idx, const double *efflen, int total_num_fragments,
                                 double count_threshold, double marginal_threshold, int mode, char *out, int cap) {
    GibbsOutput go(samples, T);
    for (int t = 0; t < T; t++) {
        go.marginal_occurence[t] = occ[t];
        go.marginal_mean_cumulant[t] = mean[t];
        go.marginal_var_cumulant[t] = m2[t];
    }
    std::vector<Candidate> cands(n_cand);
    std::stringstream nm(names);
    for (int i = 0; i < n_cand; i++) {
        cands[i].idx = i;
        std::getline(nm, cands[i].name);
        cands[i].graph_name = graph_name;
        cands[i].isPreMrna = premrna[i] != 0;
        cands[i].length = 0;
        for (int k = exon_ptr[i]; k < exon_ptr[i + 1]; k++) {
            Vertex v;
            v.reference = reference;
            v.strand = strand;
            v.start = starts[k];
            v.end = ends[k];
            cands[i].vertices.push_back(v);
            cands[i].length += ends[k] - starts[k] + 1;
        }
    }
    IdxMap m;
    for (int t = 0; t < T; t++) m[t] = std::make_pair(old_idx[t], efflen[t]);
    std::string text;
    if (mode == 0) {
        Ensemble top = go.getTopMarginals(marginal_threshold);
        if (top.indices.empty()) return -5;
        text = go.writeEnsembleGtf(top, cands, m, total_num_fragments, count_threshold);
        if (text.empty()) return -7;
    } else {
        text = go.writeCandidateGtf(cands, m, total_num_fragments);
    }
    if ((int)text.size() + 1 > cap) return -1;
    memcpy(out, text.c_str(), text.size() + 1);
    return (int)text.size();
}

}  // extern "C"
