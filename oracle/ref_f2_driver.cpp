/*
 * oracle/ref_f2_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C-callable driver around the UNMODIFIED reference translation unit allPaths.cpp (candidate
 * enumeration), compiled where it lies against the Boost.Graph / Boost.Regex stand-ins of oracle/shim/orc_graph_shim.hpp
 * into oracle/_ref/libbayesref_f2.so by oracle/Makefile.  No reference source is copied.
 * ref_all_paths() is the pair of calls Assembler::bayesemblerCallback makes per DOT string at assembler.cpp:1462-1463.
 */
#include <cstring>
#include <sstream>
#include <stdint.h>
#include <string>
#include <vector>

#include <allPaths.h>

extern "C" {

/*
 * AllPaths(dot).findFullPaths(no_pre_mrna, max_candidate_number, log) (allPaths.cpp:42-49, 116-304).
 * Candidates come back flat: exons (Candidate::vertices, adjacent exons already merged, :237-244) of candidate i are
 * [exon_ptr[i], exon_ptr[i+1]); names '\n'-joined.  stats5 = {vertices, sources, sinks, junction-threshold adjustments,
 * paths} as logged at :132-134, :317, :301.  Returns the number of candidates, or -1 when a capacity is too small.
 */
int ref_all_paths(const char *dot, int no_pre_mrna, int max_candidate_number, int cap_cand, int cap_exons, int *exon_ptr, int *exon_start,
                  int *exon_end, int *cand_length, int *is_premrna, int *cand_idx, char *names, int names_cap, char *graph_name,
                  int graph_name_cap, int *stats5) {
    stringstream log;
    AllPaths paths((string(dot)));
    vector<Candidate> cands = paths.findFullPaths(no_pre_mrna != 0, max_candidate_number, log);
    if ((int)cands.size() > cap_cand) return -1;
    int ne = 0;
    string all_names;
    exon_ptr[0] = 0;
    for (size_t i = 0; i < cands.size(); i++) {
        const Candidate &c = cands[i];
        if (ne + (int)c.vertices.size() > cap_exons) return -1;
        for (size_t j = 0; j < c.vertices.size(); j++) {
            exon_start[ne] = c.vertices[j].start;
            exon_end[ne] = c.vertices[j].end;
            ne++;
        }
        exon_ptr[i + 1] = ne;
        cand_length[i] = c.length;
        is_premrna[i] = c.isPreMrna ? 1 : 0;
        cand_idx[i] = c.idx;
        all_names += c.name;
        all_names += '\n';
        if (i == 0) {
            strncpy(graph_name, c.graph_name.c_str(), graph_name_cap - 1);
            graph_name[graph_name_cap - 1] = 0;
        }
    }
    if ((int)all_names.size() + 1 > names_cap) return -1;
    memcpy(names, all_names.c_str(), all_names.size() + 1);
    if (cands.empty() && graph_name_cap > 0) graph_name[0] = 0;
    /* the numbers of the log block */
    for (int k = 0; k < 5; k++) stats5[k] = 0;
    string line;
    while (getline(log, line)) {
        const char *keys[] = {"Number of vertices in graph: ", "Number of sources in graph: ", "Number of sinks in graph: ", NULL,
                              "Number of paths in splice-graph (transcripts): "};
        for (int k = 0; k < 5; k++) {
            if (!keys[k]) continue;
            size_t p = line.find(keys[k]);
            if (p != string::npos) stats5[k] = atoi(line.c_str() + p + strlen(keys[k]));
        }
        if (line.find("Too many candidates to continue") != string::npos) stats5[3]++;
    }
    return (int)cands.size();
}

}  // extern "C"
