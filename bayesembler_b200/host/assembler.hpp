// Host driver from the splice-graph stage on: the reference's command line, its per-graph worker
// (Assembler::bayesemblerCallback, assembler.cpp:1420-1648), its fragment-length pass (fragmentLengthCallback, :1651-1744)
// and its output writer (outputCallback, :1746-1839) on top of the facade classes of bayesembler_b200.hpp.
//
//   reference                                                        here
//   main.cpp:52-154        boost::program_options                    parseCommandLine(): same long / short names, defaults, checks
//   OptionsContainer       utils.h:140-170                           same fields (+ graph_file, device)
//   Assembler(options).runBayesemblerThreaded()  assembler.cpp:1841  Assembler(options).runBayesemblerBatched()
//   N worker threads, one graph each, start to finish                every graph is prepared on the host (AllPaths ->
//                                                                    AlignmentParser -> read cover), ALL samplers run as one
//                                                                    GPU batch (GibbsBatch), then the writer runs per graph
//   assembly.gtf / candidates.gtf / graph_logs.txt /                 same files, same text (log lines carry no time stamps)
//   fragment_lengths.txt  (:1748-1760, :1965-1978)
//
// The stages before the splice graphs (row f4, assembler.cpp:84-1417) are in the library too (csrc/bamfront.cpp, C ABI
// bayes_bam_front_*): BAM reading, duplicate removal, strand split, instance file -> DOT strings, per-graph fragment fetch, read
// probabilities -- everything but CEM processsam itself, which is not part of the reference's repository: with --bam-file the
// driver either runs it (--cem-processsam-path, like the reference, :678-690) or takes the instance file(s) it wrote
// (--instance-file / --instance-file-plus / --instance-file-minus).  Either way the front end writes a "graph file" holding what
// grapherCallback queues per graph (pair<list<string>, vector<FragmentAlignment*>>, :1167-1417), which is also an input format
// of its own (--graph-file):
//
//     fragments <unique non-redundant mapped read pairs> <all mapped read pairs>        (assembler.cpp:610-619)
//     graph <dot strings> <fragments> <single-path 0|1>                                  (one block per graph)
//     digraph <name> { ... }                                                             (DOT text, a line "}" ends it)
//     <pos1> <cigar1> <pos2> <cigar2> <probability>                                      (0-based positions, CIGARs like 30M120N46M)
//
// --bam-file without processsam and without instance files is refused with a message saying so.
#ifndef BAYESEMBLER_B200_ASSEMBLER_HPP
#define BAYESEMBLER_B200_ASSEMBLER_HPP

#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <fstream>
#include <iostream>
#include <list>
#include <atomic>
#include <chrono>
#include <memory>
#include <thread>

#include "bayesembler_b200.hpp"

namespace bayesembler_b200 {

struct OptionsContainer {  // utils.h:140-170
    std::string bam_file;
    std::string graph_file;  // this driver's input (see header comment)
    std::string instance_file, instance_file_plus, instance_file_minus;  // what CEM processsam wrote for the de-duplicated alignments
    std::string samtools_path;
    std::string cem_processsam_path;
    int num_threads = 1;
    std::string strand_specific = "unstranded";
    double marginal_threshold = 0.5;
    double count_threshold = 12;
    bool no_pre_mrna = false;
    int seed = 0;
    std::string output_prefix;
    std::string output_mode = "assembly";
    int total_num_fragments = 0;
    bool keep_temp_files = false;
    int max_candidate_number = 100;
    double exon_base_coverage = 0.05;
    double gamma = 1;
    double frag_mean = 0;
    double frag_sd = 0;
    bool estimate_fragment_length = true;
    unsigned fragment_est_min_transcript_length = 2500;
    unsigned gibbs_base_iterations = 1000;
    unsigned gibbs_scale_iterations = 60;
    int device = 0;  // CUDA device the sampler batch runs on
    std::string devices;  // "0,1,2,3": shard the graphs over several GPUs by cost (overrides device)
};

// ---- main.cpp:52-154 ---------------------------------------------------------------------------------------------------
inline void printHelp(std::ostream &out, bool advanced) {
    out << "## Bayesembler help ##:\n\n"
           "  -h [ --help ]                                   produce help message for base options\n"
           "  -a [ --advanced ]                               produce help message for advanced options\n\n"
           "== Required arguments ==:\n"
           "  -b [ --bam-file ] arg                           TopHat2 bamfile containing the mapped paired-end reads to the reference genome\n"
           "                                                  (with --cem-processsam-path, or with the instance file(s) CEM processsam wrote)\n"
           "  -g [ --graph-file ] arg                         splice graphs + fragment alignments, the stage after the BAM front end\n"
           "  --instance-file arg                             CEM instance file of the de-duplicated alignments (unstranded data)\n"
           "  --instance-file-plus / --instance-file-minus    the same for the two strand sets of strand-specific data\n"
           "  --cem-processsam-path arg                       CEM processsam executable: run on the de-duplicated alignments when no\n"
           "                                                  instance file is given\n\n"
           "== Optional arguments ==:\n"
           "  -o [ --output-prefix ] arg                      prefix to be used on all output filenames (OBS: to use output directory <dir>\n"
           "                                                  without any prefix on the output files, use <dir/>)\n"
           "  -p [ --num-threads ] arg (=1)                   number of host threads preparing graphs (paths, likelihoods, read cover)\n"
           "  -s [ --strand-specific ] arg (=unstranded)      data is strand-specific. Use \"first\" to indicate mate orientation as in the dUTP\n"
           "                                                  protocol or \"second\" if opposite\n"
           "  -c [ --confidence-threshold ] arg (=0.5)        exclude candidates with a confidence below <confidence-threshold>\n"
           "  -f [ --count-threshold ] arg (=12)              exclude candidates with an expected fragment count below <count-threshold>\n"
           "  -m [ --no-pre-mRNA ] [=arg(=1)] (=0)            do not include pre-mRNA in the set of candidates\n";
    if (!advanced) return;
    out << "\n== Advanced arguments ==:\n"
           "  --seed arg (=unix time)                         seed for pseudo-random number generator\n"
           "  --output-mode arg (=assembly)                   output a gtf file of the assembly. Use \"full\" to output additional information\n"
           "                                                  on the assembly, candidates and fragment lengths\n"
           "  --library-size arg (=0)                         the total number of sequenced paired-end reads used for FPKM normalisation. Use\n"
           "                                                  value of 0 (default) to normalise using the number of mapped paired-end reads\n"
           "  --keep-temp-files [=arg(=1)] (=0)               keep filtered bam, instance and processsam log-file\n"
           "  --max-candidate-number arg (=100)               maximum number of candidates per splice-graph (used for graph pruning)\n"
           "  --dirichlet-parameter arg (=1)                  abundance parameter (symmetric Dirichlet concentration parameter gamma)\n"
           "  --frag-mean arg                                 disable internal fragment length mean estimation and use <frag-mean>\n"
           "  --frag-sd arg                                   disable internal fragment length sd estimation and use <frag-sd>\n"
           "  --gibbs-iteration-scaling arg (=60)             scaling factor used to calculate number of Gibbs iterations (burn-in &\n"
           "                                                  sample-size) using: burn-in = 1000 + <gibbs-iteration-scaling> *\n"
           "                                                  number_of_candidates, sample-size = 10 * burn-in\n"
           "  --device arg (=0)                               CUDA device of the sampler batch\n"
           "  --devices arg                                   comma-separated CUDA devices: the graphs are sharded over them by cost\n"
           "                                                  (greedy, largest first), one host thread and one batch per device\n";
}

// Returns 0 = run, 1 = help was printed (the reference returns 1 too, main.cpp:119-127), -1 = error (message on err).
inline int parseCommandLine(int argc, char *const argv[], OptionsContainer *opt, std::ostream &out, std::ostream &err) {
    struct Spec {
        const char *name;
        char letter;
        int kind;  // 0 flag (help), 1 string, 2 int, 3 double, 4 bool with implicit true, 5 unsigned
        void *dst;
    };
    OptionsContainer &o = *opt;
    o.seed = (int)time(NULL);  // main.cpp:83
    bool help = false, advanced = false, have_mean = false, have_sd = false;
    const Spec specs[] = {{"help", 'h', 0, &help}, {"advanced", 'a', 0, &advanced}, {"bam-file", 'b', 1, &o.bam_file},
                          {"graph-file", 'g', 1, &o.graph_file}, {"output-prefix", 'o', 1, &o.output_prefix}, {"num-threads", 'p', 2, &o.num_threads},
                          {"strand-specific", 's', 1, &o.strand_specific}, {"confidence-threshold", 'c', 3, &o.marginal_threshold},
                          {"count-threshold", 'f', 3, &o.count_threshold}, {"no-pre-mRNA", 'm', 4, &o.no_pre_mrna}, {"seed", 0, 2, &o.seed},
                          {"output-mode", 0, 1, &o.output_mode}, {"library-size", 0, 2, &o.total_num_fragments},
                          {"keep-temp-files", 0, 4, &o.keep_temp_files}, {"max-candidate-number", 0, 2, &o.max_candidate_number},
                          {"dirichlet-parameter", 0, 3, &o.gamma}, {"frag-mean", 0, 3, &o.frag_mean}, {"frag-sd", 0, 3, &o.frag_sd},
                          {"gibbs-iteration-scaling", 0, 5, &o.gibbs_scale_iterations}, {"device", 0, 2, &o.device},
                          {"devices", 0, 1, &o.devices}, {"instance-file", 0, 1, &o.instance_file},
                          {"instance-file-plus", 0, 1, &o.instance_file_plus}, {"instance-file-minus", 0, 1, &o.instance_file_minus},
                          {"cem-processsam-path", 0, 1, &o.cem_processsam_path}, {"samtools-path", 0, 1, &o.samtools_path}};
    const int n_specs = (int)(sizeof(specs) / sizeof(specs[0]));
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        const Spec *s = 0;
        std::string value;
        bool has_value = false;
        if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
            const size_t eq = a.find('=');
            const std::string name = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
            for (int k = 0; k < n_specs; k++)
                if (name == specs[k].name) s = &specs[k];
            if (eq != std::string::npos) { value = a.substr(eq + 1); has_value = true; }
        } else if (a.size() >= 2 && a[0] == '-') {
            for (int k = 0; k < n_specs; k++)
                if (specs[k].letter && a[1] == specs[k].letter) s = &specs[k];
            if (a.size() > 2) { value = a.substr(2); has_value = true; }
        }
        if (!s) { err << "ERROR: unrecognised option '" << a << "'" << std::endl; return -1; }
        if (s->kind == 0) { *(bool *)s->dst = true; continue; }
        if (s->kind == 4) {  // implicit_value(true): only an attached value is read
            *(bool *)s->dst = !has_value || (value != "0" && value != "false" && value != "no" && value != "off");
            continue;
        }
        if (!has_value) {
            if (i + 1 >= argc) { err << "ERROR: the required argument for option '--" << s->name << "' is missing" << std::endl; return -1; }
            value = argv[++i];
        }
        char *end = 0;
        if (s->kind == 1) *(std::string *)s->dst = value;
        else if (s->kind == 2) *(int *)s->dst = (int)strtol(value.c_str(), &end, 10);
        else if (s->kind == 5) *(unsigned *)s->dst = (unsigned)strtoul(value.c_str(), &end, 10);
        else *(double *)s->dst = strtod(value.c_str(), &end);
        if (s->kind != 1 && (end == value.c_str() || *end)) { err << "ERROR: the argument ('" << value << "') for option '--" << s->name << "' is invalid" << std::endl; return -1; }
        if (s->dst == &o.frag_mean) have_mean = true;
        if (s->dst == &o.frag_sd) have_sd = true;
    }
    if (help || argc == 1) { printHelp(out, false); return 1; }
    if (advanced) { printHelp(out, true); return 1; }
    if (o.bam_file.empty() && o.graph_file.empty()) { err << "ERROR: the option '--bam-file' is required but missing" << std::endl; return -1; }
    if (o.strand_specific != "unstranded" && o.strand_specific != "first" && o.strand_specific != "second") {  // main.cpp:143-147
        err << "\nERROR: --strand-specific argument need to be either \"unstranded\", \"first\" or \"second\"" << std::endl;
        return -1;
    }
    if (o.output_mode != "assembly" && o.output_mode != "full") {  // main.cpp:149-153
        err << "\nERROR: --output_mode argument need to be either \"assembly\" or \"full\"" << std::endl;
        return -1;
    }
    if (have_mean != have_sd) {  // main.cpp:157 (assert)
        err << "ERROR: --frag-mean and --frag-sd must be given together" << std::endl;
        return -1;
    }
    o.estimate_fragment_length = !have_mean;  // main.cpp:155-166
    if (o.graph_file.empty()) {
        const bool stranded = o.strand_specific != "unstranded";
        const bool have_instances = stranded ? (!o.instance_file_plus.empty() && !o.instance_file_minus.empty()) : !o.instance_file.empty();
        if (!have_instances && o.cem_processsam_path.empty()) {
            err << "ERROR: --bam-file needs the splice graphs of CEM processsam, which is not part of Bayesembler: give its path "
                   "(--cem-processsam-path) or the instance file(s) it wrote for the de-duplicated alignments ("
                << (stranded ? "--instance-file-plus and --instance-file-minus" : "--instance-file") << "), or start from --graph-file" << std::endl;
            return -1;
        }
    }
    return 0;
}

// ---- what grapherCallback queues per graph (assembler.cpp:1167-1417) ---------------------------------------------------
struct GraphInput {
    std::list<std::string> dot_strings;
    std::vector<FragmentAlignment> fragments;
    bool single_path_graph = false;
};

struct AssemblyInfo {  // utils.h:119-127
    std::string name, ensemble, candidates, assembly_log;
    int return_code = 0;
};

inline bool parseCigar(const std::string &text, std::vector<BamTools::CigarOp> *out) {
    out->clear();
    size_t i = 0;
    while (i < text.size()) {
        size_t j = i;
        while (j < text.size() && text[j] >= '0' && text[j] <= '9') j++;
        if (j == i || j >= text.size() || !strchr("MIDNSHP=X", text[j])) return false;
        out->push_back(BamTools::CigarOp(text[j], (uint32_t)strtoul(text.substr(i, j - i).c_str(), 0, 10)));
        i = j + 1;
    }
    return !out->empty();
}

// Reads a graph file (header comment).  Throws std::runtime_error with the line number on malformed input.
inline std::vector<GraphInput> readGraphFile(std::istream &in, double *unique_pairs, double *mapped_pairs) {
    std::vector<GraphInput> graphs;
    std::string line;
    long line_no = 0;
    auto fail = [&](const std::string &what) { throw std::runtime_error("graph file line " + std::to_string(line_no) + ": " + what); };
    *unique_pairs = *mapped_pairs = 0;
    bool have_header = false;
    while (std::getline(in, line)) {
        line_no++;
        if (line.empty() || line[0] == '#') continue;
        std::stringstream ls(line);
        std::string word;
        ls >> word;
        if (word == "fragments") {
            if (!(ls >> *unique_pairs)) fail("'fragments <unique pairs> [<mapped pairs>]' expected");
            if (!(ls >> *mapped_pairs)) *mapped_pairs = *unique_pairs;
            have_header = true;
            continue;
        }
        if (word != "graph") fail("'graph <dot strings> <fragments> <single-path>' expected, found '" + word + "'");
        long n_dot = 0, n_frag = 0;
        int single = 0;
        if (!(ls >> n_dot >> n_frag >> single) || n_dot < 1 || n_frag < 0) fail("'graph <dot strings> <fragments> <single-path>' expected");
        GraphInput g;
        g.single_path_graph = single != 0;
        for (long d = 0; d < n_dot; d++) {
            std::string dot;
            bool closed = false;
            while (std::getline(in, line)) {
                line_no++;
                dot += line;
                dot += '\n';
                if (line == "}") { closed = true; break; }
            }
            if (!closed) fail("DOT text without a closing '}' line");
            g.dot_strings.push_back(dot);
        }
        g.fragments.resize(n_frag);
        for (long f = 0; f < n_frag; f++) {
            if (!std::getline(in, line)) fail("fragment line expected");
            line_no++;
            std::stringstream fs(line);
            unsigned p1 = 0, p2 = 0;
            std::string c1, c2;
            double prob = 0;
            if (!(fs >> p1 >> c1 >> p2 >> c2 >> prob)) fail("'<pos1> <cigar1> <pos2> <cigar2> <probability>' expected");
            FragmentAlignment &fa = g.fragments[f];
            fa.read_data.first.first = p1;
            fa.read_data.second.first = p2;
            fa.probability = prob;
            if (!parseCigar(c1, &fa.read_data.first.second) || !parseCigar(c2, &fa.read_data.second.second)) fail("malformed CIGAR");
        }
        graphs.push_back(g);
    }
    if (!have_header) { line_no = 0; fail("no 'fragments' line"); }
    return graphs;
}

class Assembler {
   public:
    explicit Assembler(const OptionsContainer &o) : options_variables(o) {}

    // generateSpliceGraphs + graphConstructorCallback + grapherCallback (assembler.cpp:272-1319) through the library's BAM front
    // end: leaves the graphs with their fragments in <output-prefix>graphs.txt and points options_variables.graph_file at it
    void generateSpliceGraphs(std::ostream &cout_) {
        OptionsContainer &o = options_variables;
        cout_ << "Removing duplicate reads" << std::endl;
        bayes_bam_front *front = 0;
        check(bayes_bam_front_open(o.bam_file.c_str(), o.strand_specific.c_str(), &front), "bayes_bam_front_open");
        struct Closer {
            bayes_bam_front *f;
            ~Closer() { bayes_bam_front_close(f); }
        } closer = {front};
        double total = 0, unique = 0, written = 0;
        check(bayes_bam_front_counts(front, &total, &unique, &written), "bayes_bam_front_counts");
        cout_ << "Removed duplicates from " << total << " mapped read pairs" << std::endl;
        cout_ << "Wrote " << written << " read pairs used for splice-graph construction" << std::endl;
        // file names like the reference's (:274-300): <prefix><bam name without its last extension>_nd_<strand>
        std::string stem = o.bam_file.substr(o.bam_file.rfind('/') == std::string::npos ? 0 : o.bam_file.rfind('/') + 1);
        if (stem.rfind('.') != std::string::npos) stem = stem.substr(0, stem.rfind('.'));
        const bool stranded = o.strand_specific != "unstranded";
        const char *set_name[2] = {stranded ? "_nd_plus" : "_nd_unstranded", "_nd_minus"};
        const char *set_strand[2] = {stranded ? "+" : ".", "-"};
        const std::string given[2] = {stranded ? o.instance_file_plus : o.instance_file, o.instance_file_minus};
        const std::string graph_file = o.output_prefix + "graphs.txt";
        int graph_count = 0, cem_total = 0, excluded_total = 0;
        for (int s = 0; s < (stranded ? 2 : 1); s++) {
            std::string instance = given[s];
            const std::string base = o.output_prefix + stem + set_name[s];
            if (instance.empty()) {
                // samtools view > .sam, then processsam (:662-690)
                const std::string sam = base + ".sam", log = base + "_instance_log.txt";
                cout_ << "Generating splice-graphs from " << base << " using cem" << std::endl;
                check(bayes_bam_front_write_sam(front, s, sam.c_str(), 0), "bayes_bam_front_write_sam");
                const std::string cmd = "nice " + o.cem_processsam_path + " --no-coverage -c 1 -g 1 -u 0.05 -d " + set_strand[s] + " -o " + base + " " + sam + " > " + log + " 2>&1";
                if (system(cmd.c_str()) != 0) throw std::runtime_error("CEM processsam failed (see " + log + "): " + cmd);
                instance = base + ".instance";
                if (!o.keep_temp_files) remove(sam.c_str());
            }
            int32_t n_graphs = 0, n_with = 0, n_excluded = 0, cem = 0;
            check(bayes_bam_front_graphs(front, s, instance.c_str(), set_strand[s], o.exon_base_coverage, o.no_pre_mrna ? 1 : 0, graph_file.c_str(), s > 0 ? 1 : 0,
                                         &n_graphs, &n_with, &n_excluded, &cem), "bayes_bam_front_graphs");
            cout_ << "Parsed " << cem << " graph(s) from cem instance file" << std::endl;
            graph_count += n_graphs;
            cem_total += cem;
            excluded_total += n_excluded;
        }
        cout_ << "\nParsed " << cem_total << " splice graph(s) from cem instance file" << (stranded ? "s" : "") << " and collapsed them to " << graph_count
              << " assembly graph(s)";
        if (!stranded) cout_ << " (" << excluded_total << " graph(s) excluded due to inference issues resulting from unstranded data).";
        cout_ << std::endl;
        if (graph_count == 0) throw std::runtime_error("No graphs remaining after exclusion due to inference issues resulting from unstranded data");
        o.graph_file = graph_file;
    }

    // assembler.cpp:1841-2030 from the point where the splice graphs exist; returns the number of graphs assembled (code 0)
    int runBayesemblerBatched(std::ostream &cout_) {
        typedef std::chrono::steady_clock Clock;
        auto seconds = [](Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
        const Clock::time_point t_start = Clock::now();
        if (options_variables.graph_file.empty()) generateSpliceGraphs(cout_);
        std::ifstream in(options_variables.graph_file.c_str());
        if (!in) throw std::runtime_error("cannot open graph file " + options_variables.graph_file);
        double unique_pairs = 0, mapped_pairs = 0;
        std::vector<GraphInput> graphs = readGraphFile(in, &unique_pairs, &mapped_pairs);
        double adjusted_fragment_count = unique_pairs;  // assembler.cpp:612-619
        if (options_variables.total_num_fragments != 0) adjusted_fragment_count = options_variables.total_num_fragments * unique_pairs / mapped_pairs;
        cout_ << unique_pairs << " unique, non-redundant read pairs being used for quantification" << std::endl;
        cout_ << adjusted_fragment_count << " read pairs being used for FPKM normalisation\n" << std::endl;

        // multi-path graphs first, then the single-path ones (:1851-1916, :2001-2009)
        std::vector<const GraphInput *> sorted;
        int single_graph_number = 0;
        for (size_t i = 0; i < graphs.size(); i++)
            if (!graphs[i].single_path_graph) sorted.push_back(&graphs[i]);
        for (size_t i = 0; i < graphs.size(); i++)
            if (graphs[i].single_path_graph) { sorted.push_back(&graphs[i]); single_graph_number++; }

        // fragment length model (:1919-1993)
        std::unique_ptr<GaussianFragmentLengthModel> fragment_length_model;
        if (options_variables.estimate_fragment_length) {
            if (single_graph_number == 0) throw std::runtime_error("no single-path graph to estimate the fragment length distribution from; give --frag-mean and --frag-sd");
            std::map<unsigned, unsigned> output_map;
            unsigned single_path_long_candidates = 0;
            for (size_t i = 0; i < graphs.size(); i++)
                if (graphs[i].single_path_graph) fragmentLengthCallback(graphs[i], &output_map, &single_path_long_candidates);
            cout_ << "Estimating fragment length distribution from " << single_path_long_candidates << " transcripts longer than "
                  << options_variables.fragment_est_min_transcript_length << " nucleotides" << std::endl;
            if (options_variables.output_mode == "full") {
                std::ofstream f((options_variables.output_prefix + "fragment_lengths.txt").c_str());
                f << "length\tcount" << std::endl;
                for (std::map<unsigned, unsigned>::iterator it = output_map.begin(); it != output_map.end(); ++it) f << it->first << "\t" << it->second << std::endl;
            }
            if (output_map.empty()) throw std::runtime_error("no fragment of a single-path transcript longer than the minimum length; give --frag-mean and --frag-sd");
            fragment_length_model.reset(new GaussianFragmentLengthModel(output_map));
            cout_ << "Estimated fragment length distribution: mean " << fragment_length_model->getMean() << ", sd " << fragment_length_model->getSd() << std::endl;
        } else {
            cout_ << "Using user-defined fragment length distribution parameters." << std::endl;
            fragment_length_model.reset(new GaussianFragmentLengthModel(options_variables.frag_mean, options_variables.frag_sd));
        }
        const int graph_number = (int)sorted.size();
        cout_ << "\nStarting Bayesembler on " << graph_number - single_graph_number << " multi-path graph(s) and " << single_graph_number
              << " single-path graph(s)" << std::endl;

        const Clock::time_point t_model = Clock::now();
        // ---- bayesemblerCallback, host part, per graph (:1420-1599); the samplers are queued, not run
        // --num-threads host threads take graphs off one counter, like the reference's workers take them off the queue (:1424-1436)
        std::vector<Prepared> prepared(graph_number);
        int n_samplers = 0;
        {
            std::atomic<int> next_graph(0);
            std::vector<std::string> errors(std::max(1, options_variables.num_threads));
            auto work = [&](int tid) {
                try {
                    for (;;) {
                        const int g = next_graph.fetch_add(1);
                        if (g >= graph_number) return;
                        // the seed of a graph is --seed + graphs left in the queue after it (:1435-1437)
                        prepareGraph(*sorted[g], options_variables.seed + (graph_number - 1 - g), adjusted_fragment_count, fragment_length_model.get(),
                                     &prepared[g]);
                    }
                } catch (std::exception &e) {
                    errors[tid] = e.what();
                }
            };
            std::vector<std::thread> pool;
            for (int t = 1; t < (int)errors.size(); t++) pool.emplace_back(work, t);
            work(0);
            for (size_t t = 0; t < pool.size(); t++) pool[t].join();
            for (size_t t = 0; t < errors.size(); t++)
                if (!errors[t].empty()) throw std::runtime_error(errors[t]);
            for (int g = 0; g < graph_number; g++) n_samplers += prepared[g].sampler ? 1 : 0;
        }
        const Clock::time_point t_prepared = Clock::now();
        // ---- every sampler of the run as one GPU batch (gibbsSampler.cpp:55-148 for all graphs at once)
        if (n_samplers > 0) {
            // graphs are independent (assembler.cpp:1441-1445): with several devices they are sharded by cost, largest first to
            // the lightest device (bayes_partition_lpt -- the reference's largest-first queue, :1886-1916, turned into shards);
            // one host thread and one batch per device, no exchange between them
            std::vector<int> devices = deviceList();
            std::vector<int> idx;
            std::vector<double> cost;
            for (int g = 0; g < graph_number; g++)
                if (prepared[g].sampler) {
                    bayes_locus_desc d = prepared[g].sampler->desc(prepared[g].burn_in, prepared[g].samples);
                    idx.push_back(g);
                    cost.push_back(bayes_locus_cost(&d));
                }
            std::vector<int32_t> part(idx.size(), 0);
            check(bayes_partition_lpt((int64_t)idx.size(), cost.data(), (int32_t)devices.size(), part.data()), "bayes_partition_lpt");
            std::vector<std::string> errors(devices.size());
            std::vector<std::thread> workers;
            for (size_t p = 0; p < devices.size(); p++)
                workers.emplace_back([&, p]() {
                    try {
                        size_t mine = 0;
                        for (size_t k = 0; k < idx.size(); k++) mine += part[k] == (int32_t)p;
                        if (mine == 0) return;
                        GibbsBatch batch(devices[p]);
                        for (size_t k = 0; k < idx.size(); k++)
                            if (part[k] == (int32_t)p) {
                                Prepared &P = prepared[idx[k]];
                                batch.add(*P.sampler, P.gibbs_output.get(), P.burn_in, P.samples);
                            }
                        batch.run();
                    } catch (std::exception &e) {
                        errors[p] = e.what();
                    }
                });
            for (size_t p = 0; p < workers.size(); p++) workers[p].join();
            for (size_t p = 0; p < errors.size(); p++)
                if (!errors[p].empty()) throw std::runtime_error("device " + std::to_string(devices[p]) + ": " + errors[p]);
            if (devices.size() > 1) cout_ << n_samplers << " sampler(s) sharded over " << devices.size() << " device(s)" << std::endl;
        }
        const Clock::time_point t_sampled = Clock::now();
        // ---- bayesemblerCallback, after the sampler (:1604-1631), and outputCallback (:1746-1839)
        std::ofstream ensemble_file((options_variables.output_prefix + "assembly.gtf").c_str());
        std::ofstream candidates_file, log_file;
        const bool full = options_variables.output_mode == "full";
        if (full) {
            candidates_file.open((options_variables.output_prefix + "candidates.gtf").c_str());
            log_file.open((options_variables.output_prefix + "graph_logs.txt").c_str());
        }
        std::vector<int> stats_list(8, 0);
        for (int g = 0; g < graph_number; g++) {
            finishGraph(&prepared[g], adjusted_fragment_count);
            const AssemblyInfo &a = prepared[g].assembly;
            if (a.return_code == 0) ensemble_file << a.ensemble;
            stats_list[a.return_code]++;
            if (full) {
                candidates_file << a.candidates;
                log_file << "\n###### " << a.name << " ######\n" << a.assembly_log << std::endl;
            }
        }
        cout_ << "\n======= FINAL STATS =======\n\n";
        cout_ << "\t" << graph_number << " graph(s) were parsed in total" << std::endl;
        cout_ << "\t" << stats_list[0] << " graph(s) were assembled successfully" << std::endl;
        cout_ << "\t" << 0 << " graph(s) were not assembled due to non successfully assigned fragments" << std::endl;
        cout_ << "\t" << stats_list[2] << " graph(s) were not assembled due to insufficient exon coverage" << std::endl;
        cout_ << "\t" << stats_list[3] << " graph(s) were not assembled due to insufficient candidate coverage" << std::endl;
        cout_ << "\t" << stats_list[4] << " graph(s) were not assembled as no fragments had sufficient mapping probability" << std::endl;
        cout_ << "\t" << stats_list[5] << " graph(s) were not assembled as no transcripts had probability higher than "
              << options_variables.marginal_threshold << std::endl;
        cout_ << "\t" << stats_list[7] << " graph(s) were not assembled due to transcripts being pre-mRNA or having an expected count lower than "
              << options_variables.count_threshold << std::endl;
        cout_ << std::endl;
        const Clock::time_point t_end = Clock::now();
        cout_ << "Timing: input + fragment length model " << seconds(t_start, t_model) << " s, graph preparation on " << options_variables.num_threads
              << " host thread(s) " << seconds(t_model, t_prepared) << " s, " << n_samplers << " sampler(s) on the GPU " << seconds(t_prepared, t_sampled)
              << " s, output " << seconds(t_sampled, t_end) << " s" << std::endl;
        return stats_list[0];
    }

    // One graph up to the point where its sampler would run; exposed for tests.
    struct Prepared {
        AssemblyInfo assembly;
        std::stringstream log;
        std::vector<Candidate> all_candidates;
        CollapsedMap frag_tran_map;
        mt_rng_t mt_rng;
        std::unique_ptr<AssignmentGenerator> assignment_generator;
        std::unique_ptr<FixedBinomialFixedGammaSimplexSizeGenerator> simplex_size_generator;
        std::unique_ptr<FixedGammaGenerator> gamma_generator;
        std::unique_ptr<ExpressionGenerator> expression_generator;
        std::unique_ptr<GibbsSampler> sampler;
        std::unique_ptr<GibbsOutput> gibbs_output;
        int burn_in = 0, samples = 0;
    };

    void prepareGraph(const GraphInput &graph, int graph_seed, double adjusted_fragment_count, FragmentLengthModel *fragment_length_model_pt,
                      Prepared *out) {
        Prepared &P = *out;
        std::stringstream &log = P.log;
        P.mt_rng.seed((uint64_t)(uint32_t)graph_seed);  // :1444-1445
        log << "\nSeeding graph random number generator with " << graph_seed << std::endl;
        int max_transcript_length = 0, sub_graph_count = 0, candidate_counter = 0;
        std::map<std::string, int> subgraph_complexity;
        std::stringstream assembly_id_ss;
        for (std::list<std::string>::const_iterator dot_iter = graph.dot_strings.begin(); dot_iter != graph.dot_strings.end(); ++dot_iter) {  // :1459-1493
            log << "\nCalculating paths for subgraph " << sub_graph_count << std::endl;
            AllPaths paths(*dot_iter);
            std::vector<Candidate> current = paths.findFullPaths(options_variables.no_pre_mrna, options_variables.max_candidate_number, log);
            for (size_t i = 0; i < current.size(); i++) {
                Candidate c = current[i];
                c.idx = candidate_counter++;
                max_transcript_length = std::max(max_transcript_length, c.length);
                P.all_candidates.push_back(c);
            }
            if (!current.empty()) {
                assembly_id_ss << current.front().graph_name;
                if (!subgraph_complexity.insert(std::make_pair(current.front().graph_name, (int)current.size())).second)
                    throw std::runtime_error("two sub-graphs of one graph share the name " + current.front().graph_name);  // reference: assert :1484
            }
            if (sub_graph_count < (int)graph.dot_strings.size() - 1) assembly_id_ss << " ";
            sub_graph_count++;
        }
        P.assembly.name = assembly_id_ss.str();
        if (P.all_candidates.empty()) { P.assembly.return_code = 2; return; }  // :1498-1507
        // :1511-1514; AlignmentParser deletes the fragments it is given, like the reference's
        SequencingModel sequencing_model(fragment_length_model_pt, max_transcript_length, log);
        std::vector<FragmentAlignment *> frags;
        for (size_t i = 0; i < graph.fragments.size(); i++) frags.push_back(new FragmentAlignment(graph.fragments[i]));
        AlignmentParser alignment_parser;
        P.frag_tran_map = alignment_parser.calculateFragTranProbabilities(P.all_candidates, &frags, &sequencing_model, false, log, &subgraph_complexity);
        if (P.frag_tran_map.return_code != 0) { P.assembly.return_code = P.frag_tran_map.return_code; return; }  // :1519-1531
        const int num_transcripts = P.frag_tran_map.probability_matrix.cols();
        int num_fragments = 0;
        for (size_t i = 0; i < P.frag_tran_map.counts.size(); i++) num_fragments += P.frag_tran_map.counts[i];
        log << "\nNumber of candidates used for further analysis: " << num_transcripts << std::endl;
        log << "Number of fragments used for further analysis: " << num_fragments << std::endl;
        // :1548-1577
        std::vector<int32_t> counts32(P.frag_tran_map.counts.begin(), P.frag_tran_map.counts.end());
        P.assignment_generator.reset(new AssignmentGenerator(&P.mt_rng, num_transcripts, P.frag_tran_map.probability_matrix.row_ptr,
                                                             P.frag_tran_map.probability_matrix.col_idx, P.frag_tran_map.probability_matrix.val, counts32));
        P.assignment_generator->calcMinimumReadCover(log);
        double msc_to_all_ratio = double(P.assignment_generator->getMinimumReadCoverSize()) / num_transcripts;
        const double double_almost_one = 1 - 2.220446049250313e-16;  // utils.h:55
        if (msc_to_all_ratio > double_almost_one) msc_to_all_ratio = double_almost_one;
        log << "Initialising Gibbs sampler simplex size generator with minimum read based fixed pi and fixed gamma (pi=" << msc_to_all_ratio
            << ", gamma=" << options_variables.gamma << ")" << std::endl;
        P.simplex_size_generator.reset(new FixedBinomialFixedGammaSimplexSizeGenerator(msc_to_all_ratio, options_variables.gamma, num_transcripts, &P.mt_rng, num_fragments));
        log << "Initialising Gibbs sampler gamma generator with fixed gamma (gamma=" << options_variables.gamma << ")" << std::endl;
        P.gamma_generator.reset(new FixedGammaGenerator(options_variables.gamma));
        log << "Initialising Gibbs sampler expression generator with symmetric Dirichlet prior" << std::endl;
        P.expression_generator.reset(new ExpressionGenerator(num_fragments, num_transcripts, &P.mt_rng));
        P.sampler.reset(new GibbsSampler(P.simplex_size_generator.get(), P.gamma_generator.get(), P.expression_generator.get(), P.assignment_generator.get(),
                                         P.frag_tran_map.new_idx_to_old_idx_effective_length, (int)adjusted_fragment_count, num_fragments, num_transcripts,
                                         &P.mt_rng, false));
        int max_subgraph_complexity = 0;  // :1579-1599
        for (std::map<std::string, int>::iterator it = subgraph_complexity.begin(); it != subgraph_complexity.end(); ++it)
            max_subgraph_complexity = std::max(max_subgraph_complexity, it->second);
        P.burn_in = (int)(options_variables.gibbs_base_iterations + options_variables.gibbs_scale_iterations * max_subgraph_complexity);
        P.samples = (int)(10 * options_variables.gibbs_base_iterations + 10 * options_variables.gibbs_scale_iterations * max_subgraph_complexity);
        P.gibbs_output.reset(new GibbsOutput(P.samples, num_transcripts));
        GibbsSampler::writeSamplerLog(log, num_transcripts, P.burn_in, P.samples);  // the lines of runSampler (gibbsSampler.cpp:62, :99-143)
    }

    void finishGraph(Prepared *p, double adjusted_fragment_count) {  // :1604-1631
        Prepared &P = *p;
        if (P.sampler) {
            P.assembly.return_code = 0;
            Ensemble top_marginals = P.gibbs_output->getTopMarginals(options_variables.marginal_threshold);
            if (top_marginals.indices.size() < 1) {
                P.log << "\nNo transcripts with a marginal probability over " << options_variables.marginal_threshold << " found - exiting!" << std::endl;
                P.assembly.return_code = 5;
            } else {
                P.assembly.ensemble = P.gibbs_output->writeEnsembleGtf(top_marginals, P.all_candidates, P.frag_tran_map.new_idx_to_old_idx_effective_length,
                                                                       (int)adjusted_fragment_count, options_variables.count_threshold);
                if (P.assembly.ensemble.size() < 1) {
                    P.log << "\nNo transcripts with an expected count threshold over " << options_variables.count_threshold << " found - exiting!" << std::endl;
                    P.assembly.return_code = 7;
                }
            }
            if (options_variables.output_mode == "full")
                P.assembly.candidates = P.gibbs_output->writeCandidateGtf(P.all_candidates, P.frag_tran_map.new_idx_to_old_idx_effective_length, (int)adjusted_fragment_count);
        }
        P.assembly.assembly_log = P.log.str();
    }

    // fragmentLengthCallback for one single-path graph (:1651-1727)
    void fragmentLengthCallback(const GraphInput &graph, std::map<unsigned, unsigned> *output_map, unsigned *single_path_long_candidates) {
        std::stringstream log;
        std::vector<Candidate> all_candidates;
        int candidate_counter = 0;
        if (graph.dot_strings.size() != 1) throw std::runtime_error("a single-path graph must have exactly one DOT string");  // reference: assert :1683
        AllPaths paths(graph.dot_strings.front());
        std::vector<Candidate> current = paths.findFullPaths(true, options_variables.max_candidate_number, log);
        for (size_t i = 0; i < current.size(); i++) {
            if ((unsigned)current[i].length < options_variables.fragment_est_min_transcript_length) continue;
            (*single_path_long_candidates)++;
            Candidate c = current[i];
            c.idx = candidate_counter++;
            all_candidates.push_back(c);
        }
        if (all_candidates.empty()) return;
        std::vector<FragmentAlignment *> frags;
        for (size_t i = 0; i < graph.fragments.size(); i++) frags.push_back(new FragmentAlignment(graph.fragments[i]));
        AlignmentParser alignment_parser;
        alignment_parser.fetchFragmentLengths(output_map, all_candidates, &frags, log);
    }

   private:
    std::vector<int> deviceList() const {
        std::vector<int> out;
        std::stringstream ss(options_variables.devices);
        std::string tok;
        while (std::getline(ss, tok, ','))
            if (!tok.empty()) out.push_back(atoi(tok.c_str()));
        if (out.empty()) out.push_back(options_variables.device);
        return out;
    }

    OptionsContainer options_variables;
};

}  // namespace bayesembler_b200

#endif  // BAYESEMBLER_B200_ASSEMBLER_HPP
