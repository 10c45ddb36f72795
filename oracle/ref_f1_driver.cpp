/*
 * oracle/ref_f1_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C-callable driver around the UNMODIFIED reference translation units of the likelihood builder (SURVEY.md 8f, row f1:
 * alignmentParser.cpp, sequencingModel.cpp, fragmentLengthModel.cpp from /root/reference/src), compiled against the
 * stand-in headers in oracle/shim/ into oracle/_ref/libbayesref_f1.so by oracle/Makefile.  No reference source is copied.
 * ref_collapsed_map() is the call Assembler::bayesemblerCallback makes at assembler.cpp:1511-1514.
 */
#include <cstring>
#include <sstream>
#include <stdint.h>
#include <vector>

#include <alignmentParser.h>
#include <sequencingModel.h>

extern "C" {

/* sequencingModel.cpp:34-76 on a Gaussian fragment length model (fragmentLengthModel.cpp:40-69): the three tables */
int ref_sequencing_tables(double mean, double sd, int max_transcript_length, double *three_prime_prob_at_len, double *length_prob_200,
                          double *efflen) {
    stringstream log;
    streambuf *old = cout.rdbuf(log.rdbuf());
    GaussianFragmentLengthModel flm(mean, sd);
    cout.rdbuf(old);
    SequencingModel sm(&flm, max_transcript_length, log);
    for (int l = 1; l <= max_transcript_length; l++) {
        efflen[l - 1] = sm.getEffectiveLength(l);
        three_prime_prob_at_len[l - 1] = sm.calculateThreePrimeProb(l, max_transcript_length);
        length_prob_200[l - 1] = sm.calculateLengthProb(l < 200 ? l : 200, l);
    }
    return 0;
}

/* GaussianFragmentLengthModel(map) (fragmentLengthModel.cpp:54-161): median / MAD estimate -> (mean, sd) */
int ref_estimate_fragment_model(int n, const uint32_t *length, const uint32_t *count, double *median, double *mad) {
    map<uint, uint> m;
    for (int i = 0; i < n; i++) m[length[i]] = count[i];
    stringstream log;
    streambuf *old = cout.rdbuf(log.rdbuf());
    GaussianFragmentLengthModel flm;
    pair<double, double> p = flm.estimateParameters(m);
    cout.rdbuf(old);
    *median = p.first;
    *mad = p.second;
    return 0;
}

/*
 * AlignmentParser::calculateFragTranProbabilities (alignmentParser.cpp:251-689).
 * Candidates: exon_ptr[n_cand + 1] into exon_start / exon_end (1-based inclusive, as Vertex.start / .end), length, isPreMrna,
 * subgraph id (graph_name "g<id>").  Fragments: read_pos[2 i], read_pos[2 i + 1] = 0-based leftmost positions of the two
 * mates (FragmentAlignment.read_data.{first,second}.first), their CIGARs at cigar_ptr[2 i], cigar_ptr[2 i + 1].
 * Outputs: dense row-major P (R x T), counts, new -> old candidate index, effective lengths, the decremented
 * subgraph complexities.  Returns CollapsedMap.return_code, or -1 when R exceeds r_cap.
 */
int ref_collapsed_map(int n_cand, const int *exon_ptr, const int *exon_start, const int *exon_end, const int *cand_length,
                      const int *is_premrna, const int *subgraph_id, int strand_plus, int n_frag, const uint32_t *read_pos,
                      const int *cigar_ptr, const char *cigar_op, const uint32_t *cigar_len, const double *frag_prob, double frag_mean,
                      double frag_sd, int max_transcript_length, int no_coverage_exclude, int n_subgraphs, int *subgraph_complexity,
                      int r_cap, int *T_out, int *R_out, double *P_out, int *counts_out, int *new_to_old, double *efflen_out) {
    vector<Candidate> candidates(n_cand);
    for (int i = 0; i < n_cand; i++) {
        Candidate &c = candidates[i];
        c.idx = i;
        stringstream nm, gn;
        nm << "cand" << i;
        gn << "g" << subgraph_id[i];
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
    vector<FragmentAlignment *> frags(n_frag);
    for (int i = 0; i < n_frag; i++) {
        FragmentAlignment *f = new FragmentAlignment;  /* the parser deletes them (alignmentParser.cpp:378, 418) */
        for (int m = 0; m < 2; m++) {
            ReadData &rd = m == 0 ? f->read_data.first : f->read_data.second;
            rd.first = read_pos[2 * i + m];
            for (int k = cigar_ptr[2 * i + m]; k < cigar_ptr[2 * i + m + 1]; k++) rd.second.push_back(BamTools::CigarOp(cigar_op[k], cigar_len[k]));
        }
        f->probability = frag_prob[i];
        frags[i] = f;
    }
    tr1::unordered_map<string, int> complexity;
    for (int g = 0; g < n_subgraphs; g++) {
        stringstream gn;
        gn << "g" << g;
        complexity[gn.str()] = subgraph_complexity[g];
    }
    stringstream log;
    streambuf *old = cout.rdbuf(log.rdbuf());
    GaussianFragmentLengthModel flm(frag_mean, frag_sd);
    cout.rdbuf(old);
    SequencingModel sm(&flm, max_transcript_length, log);
    AlignmentParser parser;
    CollapsedMap cm = parser.calculateFragTranProbabilities(candidates, &frags, &sm, no_coverage_exclude != 0, log, &complexity);
    for (int g = 0; g < n_subgraphs; g++) {
        stringstream gn;
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
    return 0;
}

/* AlignmentParser::fetchFragmentLengths (alignmentParser.cpp:692-725): histogram of the implied fragment lengths of all
 * (fragment, candidate) matches as ascending (length, count) pairs; returns the number of distinct lengths. */
int ref_fetch_fragment_lengths(int n_cand, const int *exon_ptr, const int *exon_start, const int *exon_end, const int *cand_length,
                               int strand_plus, int n_frag, const uint32_t *read_pos, const int *cigar_ptr, const char *cigar_op,
                               const uint32_t *cigar_len, uint32_t *length_out, uint32_t *count_out, int cap) {
    vector<Candidate> candidates(n_cand);
    for (int i = 0; i < n_cand; i++) {
        Candidate &c = candidates[i];
        c.idx = i;
        c.length = cand_length[i];
        c.isPreMrna = false;
        for (int k = exon_ptr[i]; k < exon_ptr[i + 1]; k++) {
            Vertex v;
            v.reference = "chr1";
            v.strand = strand_plus ? "+" : "-";
            v.start = exon_start[k];
            v.end = exon_end[k];
            c.vertices.push_back(v);
        }
    }
    vector<FragmentAlignment *> frags(n_frag);
    for (int i = 0; i < n_frag; i++) {
        FragmentAlignment *f = new FragmentAlignment;
        for (int m = 0; m < 2; m++) {
            ReadData &rd = m == 0 ? f->read_data.first : f->read_data.second;
            rd.first = read_pos[2 * i + m];
            for (int k = cigar_ptr[2 * i + m]; k < cigar_ptr[2 * i + m + 1]; k++) rd.second.push_back(BamTools::CigarOp(cigar_op[k], cigar_len[k]));
        }
        f->probability = 1.0;
        frags[i] = f;
    }
    map<uint, uint> hist;
    stringstream log;
    AlignmentParser parser;
    parser.fetchFragmentLengths(&hist, candidates, &frags, log);
    int n = 0;
    for (map<uint, uint>::iterator it = hist.begin(); it != hist.end(); it++, n++)
        if (n < cap) { length_out[n] = it->first; count_out[n] = it->second; }
    return n;
}

} /* extern "C" */
