/*
 * bayesembler_b200.h -- C ABI of the B200-native per-locus Gibbs sampler.
 *
 * Drop-in boundary for ONE path of bioinformatics-centre/bayesembler v1.2.0: the block
 * assembler.cpp:1548-1609 inside Assembler::bayesemblerCallback, i.e. everything between "a CollapsedMap
 * exists" and "GibbsOutput holds the posterior".  The reference has no FFI of its own; each entry point
 * below names the reference interface it stands in for (paths relative to the reference's src/).
 *
 * Plain C, no C++/torch types.  Every function returns 0 on success or a negative BAYES_E_* code and never
 * throws; bayes_last_error() gives the message of the calling thread's last failure.  There is no CPU
 * fallback: without a usable CUDA device bayes_gibbs_create() fails.
 *
 * Usage (one context per GPU, the reference's N worker threads become one batched call):
 *     bayes_gibbs_create(device, &ctx);
 *     bayes_gibbs_submit(ctx, n, loci, &first_id);      // pack + min read cover + simplex tables (host, threaded)
 *     bayes_gibbs_upload(ctx);                          // H2D, async on the context stream
 *     bayes_gibbs_run(ctx);                             // all Gibbs iterations of all (locus, chain) jobs, async
 *     bayes_gibbs_download(ctx);                        // D2H of the accumulators, async
 *     bayes_gibbs_sync(ctx);
 *     bayes_gibbs_fetch(ctx, id, chain, occ, mean, m2); // GibbsOutput state of one job
 *     bayes_gibbs_destroy(ctx);
 */
#ifndef BAYESEMBLER_B200_H
#define BAYESEMBLER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BAYES_OK 0
#define BAYES_E_INVALID (-1)  /* bad argument / malformed locus      */
#define BAYES_E_CUDA (-2)     /* CUDA runtime error                  */
#define BAYES_E_STATE (-3)    /* call out of order                   */
#define BAYES_E_NOMEM (-4)
#define BAYES_E_NODEVICE (-5) /* no CUDA device: there is no CPU path */
#define BAYES_E_NUMERIC (-6)  /* a state the reference asserts against (a row without any candidate of non-zero
                                 probability, assignmentGenerator.cpp:284; a Gamma draw below DBL_MIN,
                                 expressionGenerator.cpp:115): reported by bayes_gibbs_sync after a download */

/* Sampling modes (bayes_gibbs_set_mode).  Both draw every iteration from exactly the conditional distributions of the
 * reference (gibbsSampler.cpp:103-141); they differ in WHICH random numbers feed which draw.
 *   BAYES_MODE_FAST    production mode, the default: rows with >= 20 fragments split their fragments over the cells in
 *                      play by a balanced binary tree of binomials (depth log2 k instead of the reference's chain of k
 *                      dependent conditional binomials, assignmentGenerator.cpp:336-365; all nodes of a tree level of
 *                      all rows are drawn at once), rows with fewer fragments locate words of a per-iteration pool of
 *                      Philox blocks in their CDF (:289-334).
 *   BAYES_MODE_REPLAY  draw for draw the reference's order of draws (one site per (row, column) of the chain), the
 *                      mode the replay parity tests pin against the reference's own code.
 * The CPU oracle restates both (oracle/bayes_oracle.c: ORC_WS_FAST, ORC_WS_SITE). */
#define BAYES_MODE_FAST 0
#define BAYES_MODE_REPLAY 1

typedef struct bayes_gibbs_ctx bayes_gibbs_ctx;

/*
 * One locus = one CollapsedMap (utils.h:129-136) plus the sampler arguments the call site derives for it.
 * The probability matrix is passed as CSR over its non-zero cells (the reference stores it dense,
 * column-major, zeros = incompatible); columns must be strictly ascending inside a row, every row needs at
 * least one cell >= DBL_MIN and a count >= 1.
 */
typedef struct {
    int32_t T;                  /* candidates = probability_matrix.cols()                                   */
    int32_t R;                  /* collapsed rows = probability_matrix.rows() = counts.size()               */
    const int32_t *row_ptr;     /* R+1                                                                      */
    const int32_t *col_idx;     /* nnz                                                                      */
    const double *val;          /* nnz, CollapsedMap::probability_matrix(i,j)                               */
    const int32_t *row_count;   /* R,   CollapsedMap::counts                                                */
    const double *efflen;       /* T,   new_idx_to_old_idx_effective_length[t].second                       */
    double total_num_fragments; /* adjusted_fragment_count; truncated to int as GibbsSampler's ctor does
                                   (gibbsSampler.h:51, assembler.cpp:1577)                                  */
    double gamma;               /* --dirichlet-parameter (assembler.cpp:1569)                               */
    double pi;                  /* < 0: derive from the minimum read cover (assembler.cpp:1551-1559), the
                                   reference's behaviour; otherwise use this value                          */
    int32_t burn_in;            /* gibbs_burn_in  (assembler.cpp:1598)                                      */
    int32_t samples;            /* gibbs_samples  (assembler.cpp:1599)                                      */
    uint64_t seed;              /* graph_seed (assembler.cpp:1437); Philox key of chain c is derived from
                                   (seed, c) by bayes_chain_key()                                           */
    int32_t n_chains;           /* independent chains of this locus (reference: 1); each chain is a complete
                                   independent run (own read cover when pi < 0, own Philox key)             */
    int32_t reserved;
} bayes_locus_desc;

/* ---- context ------------------------------------------------------------------------------------------ */
int bayes_gibbs_create(int device, bayes_gibbs_ctx **out);
void bayes_gibbs_destroy(bayes_gibbs_ctx *ctx);
/* launch on a caller-owned cudaStream_t (e.g. torch.cuda.Stream().cuda_stream) instead of the context's own */
int bayes_gibbs_set_stream(bayes_gibbs_ctx *ctx, void *cuda_stream);
void *bayes_gibbs_get_stream(bayes_gibbs_ctx *ctx);
/* host worker threads used by submit (0 = all hardware threads) */
int bayes_gibbs_set_host_threads(bayes_gibbs_ctx *ctx, int n_threads);
const char *bayes_last_error(void);

/* ---- the sampler (replaces the objects built at assembler.cpp:1548-1577 and GibbsSampler::runSampler,
 *      gibbsSampler.cpp:55-148, for a whole batch of loci) ------------------------------------------------ */
/* Host side of K0: folds/sorts/packs the rows, runs AssignmentGenerator::calcMinimumReadCover
 * (assignmentGenerator.cpp:52-96) when pi < 0 and builds the FixedBinomialFixedGammaSimplexSizeGenerator
 * CDF table (simplexSizeGenerator.cpp:117-187).  Input buffers may be freed when it returns.  Locus ids are
 * consecutive from *first_id. */
int bayes_gibbs_submit(bayes_gibbs_ctx *ctx, int32_t n_loci, const bayes_locus_desc *loci, int64_t *first_id);
/* Forms the units, packs them and enqueues their H2D copies on the context stream.  For a batch of several units per SM it also
 * runs the calibration pass (bayes_gibbs_set_calibration): the first iterations of every unit are timed on the device and the
 * batch is re-formed, which SYNCHRONISES the context stream before it returns -- unless the context has calibrated the same batch
 * (same shapes, iteration counts and mode) before: then the corrections are reused and the call only enqueues.  An upload while
 * the previous batch of the context is still uploaded synchronises first (it rewrites the pinned staging buffers). */
int bayes_gibbs_upload(bayes_gibbs_ctx *ctx);
/* initAssignment (assignmentGenerator.cpp:213-260), initSimplexSize (simplexSizeGenerator.cpp:201-206) and
 * burn_in + samples iterations of generateExpression / generateAssignment / generateSimplexSize / addSample
 * per job, one thread block per (locus, chain).  May be called repeatedly; every call restarts all chains. */
int bayes_gibbs_run(bayes_gibbs_ctx *ctx);
int bayes_gibbs_download(bayes_gibbs_ctx *ctx);
int bayes_gibbs_sync(bayes_gibbs_ctx *ctx);
/* submit + upload + run + download + sync in one call (the end-to-end path with host buffers) */
int bayes_gibbs_run_batch(bayes_gibbs_ctx *ctx, int32_t n_loci, const bayes_locus_desc *loci, int64_t *first_id);
/* forget all submitted loci (device buffers are kept for reuse) */
int bayes_gibbs_clear(bayes_gibbs_ctx *ctx);

/* GibbsOutput's state (gibbsOutput.cpp:40-75): marginal_occurence, marginal_mean_cumulant,
 * marginal_var_cumulant, T entries each; any pointer may be NULL.  final_counts = map_counts after the
 * last iteration (gibbsSampler.cpp:145). */
int bayes_gibbs_fetch(bayes_gibbs_ctx *ctx, int64_t locus_id, int32_t chain, int32_t *occ, double *mean, double *m2,
                      int32_t *final_counts);
/* The same for every job at once: locus after locus in submission order, chains of a locus back to back, T entries
 * per job (what the reference's output thread would drain, assembler.cpp:1746-1839). */
int bayes_gibbs_fetch_all(bayes_gibbs_ctx *ctx, int32_t *occ, double *mean, double *m2, int32_t *final_counts);
/* pi actually used by chain 0, its |minimum read cover| (0 when pi was given), folded single-cell rows */
int bayes_gibbs_locus_info(bayes_gibbs_ctx *ctx, int64_t locus_id, double *pi, int32_t *cover_size, int32_t *rows_folded);
/* Every chain is one independent run of the reference's call site: unless pi was given, chain c derives pi from its
 * own minimum read cover (tie breaks keyed by bayes_chain_key(seed, c), assembler.cpp:1551-1559). */
int bayes_gibbs_chain_info(bayes_gibbs_ctx *ctx, int64_t locus_id, int32_t chain, double *pi, int32_t *cover_size);
int64_t bayes_gibbs_num_loci(bayes_gibbs_ctx *ctx);
int64_t bayes_gibbs_num_jobs(bayes_gibbs_ctx *ctx);
/* Scheduling diagnostics of the last run, one entry per thread-block unit: shape8 = {wide, threads, slots, row slabs,
 * padded cells, iterations, T of slot 0, matrix in shared memory | rows per slab << 8}, clock3 = {start ns, end ns, SM id} (globaltimer).
 * predicted_s = 5 doubles per unit: the scheduler's time-model estimate (s) and its inputs (chain slab columns,
 * uniform slabs, uniform Philox blocks, uniform block-columns).  Returns the number of units; call with all pointers NULL
 * to size the buffers. */
int64_t bayes_gibbs_unit_stats(bayes_gibbs_ctx *ctx, int32_t *shape8, uint64_t *clock3, double *predicted_s);
/* Wall-clock seconds of the host phases of the last batch: {prepare (fold rows, read cover, CDF tables), unit formation
 * and packing, H2D enqueue, launch enqueue, D2H enqueue, final synchronise}. */
int bayes_gibbs_phase_seconds(bayes_gibbs_ctx *ctx, double *out6);
/* Calibration pass of bayes_gibbs_upload: every unit runs its first `iterations` Gibbs iterations once, timed on the device,
 * and the batch is then sized and ordered with the measured per-iteration times (results do not depend on it: draws are
 * keyed by site).  -1 = automatic (256 iterations when the batch is several times what the GPU holds at once), 0 = off. */
int bayes_gibbs_set_calibration(bayes_gibbs_ctx *ctx, int32_t iterations);
/* Sampling mode of the batches uploaded from now on (before bayes_gibbs_upload); the environment variable BAYES_MODE
 * (fast | replay) sets the default of new contexts. */
int bayes_gibbs_set_mode(bayes_gibbs_ctx *ctx, int32_t mode);
int32_t bayes_gibbs_get_mode(bayes_gibbs_ctx *ctx);
/* kernels launched by the last bayes_gibbs_run() */
int32_t bayes_gibbs_last_launches(bayes_gibbs_ctx *ctx);

/* Replay/debug: record theta and map_counts of the first trace_iters iterations of one job during the next
 * run (call after submit, before upload).  theta_trace / counts_trace: trace_iters*T, b_trace: trace_iters,
 * init_counts: T (initAssignment's result). */
int bayes_gibbs_set_trace(bayes_gibbs_ctx *ctx, int64_t locus_id, int32_t chain, int32_t trace_iters);
int bayes_gibbs_fetch_trace(bayes_gibbs_ctx *ctx, double *theta_trace, int32_t *counts_trace, int32_t *b_trace,
                            int32_t *init_counts);

/* GibbsOutput::getTopMarginals (gibbsOutput.cpp:77-106) on fetched accumulators; returns the number kept */
int bayes_top_marginals(int32_t T, int32_t samples, const int32_t *occ, const double *mean, const double *m2,
                        double marginal_threshold, int32_t *idx_out, double *mean_out, double *sd_out);

/* ---- single sampler steps on the GPU, for step-level parity (same device code as bayes_gibbs_run) ------- */
/* AssignmentGenerator::generateAssignment(expression) (assignmentGenerator.cpp:263-373): theta[T] in,
 * counts_out[T]; draw sites keyed (ASSIGN, iteration, row, col) under `key`. */
int bayes_step_generate_assignment(int device, const bayes_locus_desc *locus, const double *theta, uint64_t key,
                                   uint32_t iteration, int32_t *counts_out);
/* AssignmentGenerator::initAssignment (assignmentGenerator.cpp:213-260) */
int bayes_step_init_assignment(int device, const bayes_locus_desc *locus, uint64_t key, int32_t *counts_out);
/* generateAssignment (theta != NULL) or initAssignment (theta == NULL) in the given sampling mode; BAYES_MODE_REPLAY is
 * what the two functions above run */
int bayes_step_assignment(int device, const bayes_locus_desc *locus, const double *theta, uint64_t key, uint32_t iteration,
                          int32_t mode, int32_t *counts_out);
/* ExpressionGenerator::generateExpression(b, gamma, map_counts) (expressionGenerator.cpp:77-132) */
int bayes_step_generate_expression(int device, int32_t T, const int32_t *counts, int32_t b, double gamma, uint64_t key,
                                   uint32_t iteration, double *theta_out);
/* raw variates, n draws; draw i uses site (phase, 0, i, 0): gamma -> PH_GAMMA, binomial -> PH_ASSIGN,
 * uniform_int -> PH_NULLPICK */
int bayes_step_variates(int device, int kind, double a, double p, uint64_t key, int32_t n, double *out);

/* ---- host-side pieces of the path (C++, no GPU needed) -------------------------------------------------- */
/* AssignmentGenerator::calcMinimumReadCover + getMinimumReadCoverSize (assignmentGenerator.cpp:52-96);
 * tie breaks from sites (COVER, 0, round) under `key`.  in_cover: T flags or NULL.  Returns |cover| or <0. */
int bayes_min_read_cover(const bayes_locus_desc *locus, uint64_t key, int32_t *in_cover);
/* FixedBinomialFixedGammaSimplexSizeGenerator ctor (simplexSizeGenerator.cpp:117-187): row k-1 (|S+| = k)
 * is table[row_off[k-1] .. row_off[k]).  Call with table == NULL to size it: returns total entries. */
int64_t bayes_simplex_table(double pi, double gamma, int32_t T, int32_t num_fragments, int64_t *row_off, double *table);
/* Philox key of chain c of a locus seeded with `seed` */
uint64_t bayes_chain_key(uint64_t seed, int32_t chain);
/* cost model used to order jobs and to shard loci across GPUs: iterations * (cells + rows + candidates) */
double bayes_locus_cost(const bayes_locus_desc *locus);
/* cost-balanced greedy (LPT) partition of n items into n_parts (mirrors the reference's largest-first work
 * queue, assembler.cpp:1886-1916); part_out[i] in [0, n_parts) */
int bayes_partition_lpt(int64_t n, const double *cost, int32_t n_parts, int32_t *part_out);
/* Philox4x32-10 block (known-answer tests) */
void bayes_philox4x32(uint64_t key, const uint32_t ctr[4], uint32_t out[4]);

/* ---- likelihood / compatibility builder (host, C++): produces the sampler's input ------------ */
/* One locus at the level AlignmentParser::calculateFragTranProbabilities consumes (alignmentParser.cpp:251):
 * candidates = vector<Candidate> (utils.h:109-117), fragments = vector<FragmentAlignment*> (utils.h:73-77). */
typedef struct {
    int32_t n_cand;
    const int32_t *exon_ptr;     /* n_cand + 1: exons (Candidate::vertices) of candidate i are [exon_ptr[i], exon_ptr[i+1]) */
    const int32_t *exon_start;   /* Vertex::start, 1-based inclusive                                                     */
    const int32_t *exon_end;     /* Vertex::end                                                                          */
    const int32_t *cand_length;  /* Candidate::length                                                                    */
    const int32_t *is_premrna;   /* Candidate::isPreMrna                                                                 */
    const int32_t *subgraph_id;  /* index of Candidate::graph_name among the locus' sub-graphs (may be NULL)             */
    int32_t n_subgraphs;
    int32_t strand_plus;         /* candidates.front().vertices.front().strand == "+" (alignmentParser.cpp:273)           */
    int32_t n_frag;
    const uint32_t *read_pos;    /* 2 n_frag: read_data.first.first, read_data.second.first (0-based leftmost position)   */
    const int32_t *cigar_ptr;    /* 2 n_frag + 1 offsets into cigar_op / cigar_len, mates of a fragment back to back      */
    const char *cigar_op;        /* BamTools::CigarOp::Type: M N I D                                                     */
    const uint32_t *cigar_len;   /* BamTools::CigarOp::Length                                                            */
    const double *frag_prob;     /* FragmentAlignment::probability                                                       */
} bayes_raw_locus;

typedef struct bayes_seq_model bayes_seq_model;
typedef struct bayes_collapsed_map bayes_collapsed_map;

/* GaussianFragmentLengthModel(mean, sd) (fragmentLengthModel.cpp:40-69) tabulated by SequencingModel's constructor
 * (sequencingModel.cpp:34-57) for transcript lengths up to max_transcript_length */
int bayes_seq_model_create_gaussian(double mean, double sd, int32_t max_transcript_length, bayes_seq_model **out);
void bayes_seq_model_destroy(bayes_seq_model *model);
/* SequencingModel::calculateThreePrimeProb / calculateLengthProb / getEffectiveLength (sequencingModel.cpp:59-76) */
double bayes_seq_model_three_prime_prob(const bayes_seq_model *model, int32_t three_prime_pos, int32_t tran_length);
double bayes_seq_model_length_prob(const bayes_seq_model *model, int32_t fragment_length, int32_t three_prime_pos);
double bayes_seq_model_effective_length(const bayes_seq_model *model, int32_t transcript_length);
/* GaussianFragmentLengthModel::estimateParameters (fragmentLengthModel.cpp:71-161): median and median absolute deviation
 * of a fragment length histogram (mean = median, sd = 1.4826 * MAD, :56-59) */
int bayes_fragment_model_estimate(int32_t n, const uint32_t *length, const uint32_t *count, double *median_out, double *mad_out);
/* AlignmentParser::calculateFragTranProbabilities (alignmentParser.cpp:251-689).  Returns CollapsedMap::return_code
 * (0, 3 = every candidate excluded, 4 = no rows) or a negative BAYES_E_*; subgraph_complexity (n_subgraphs entries or
 * NULL) is decremented per excluded candidate (:501-502) -- its maximum sets the iteration count (assembler.cpp:1598). */
int bayes_collapsed_map_build(const bayes_raw_locus *locus, const bayes_seq_model *model, int32_t no_coverage_exclude,
                              int32_t *subgraph_complexity, bayes_collapsed_map **out);
void bayes_collapsed_map_destroy(bayes_collapsed_map *cm);
int32_t bayes_collapsed_map_num_candidates(const bayes_collapsed_map *cm);
int32_t bayes_collapsed_map_num_rows(const bayes_collapsed_map *cm);
int64_t bayes_collapsed_map_nnz(const bayes_collapsed_map *cm);
/* CollapsedMap (utils.h:129-136) as the CSR arrays bayes_locus_desc takes: row_ptr R+1, col_idx / val nnz, row_count R,
 * new -> old candidate index and effective length T each; any pointer may be NULL */
int bayes_collapsed_map_export(const bayes_collapsed_map *cm, int32_t *row_ptr, int32_t *col_idx, double *val, int32_t *row_count,
                               int32_t *new_to_old, double *efflen);
/* the numbers of the "BAM-FILE PARSING STATS" log block (alignmentParser.cpp:662-681): fragments parsed, matched,
 * unmatched, removed by underflow, fragment-candidate matches, unique rows, candidates excluded, of which without
 * mapped fragments / internal coverage / upstream coverage */
int bayes_collapsed_map_stats(const bayes_collapsed_map *cm, int64_t *stats10);
/* AlignmentParser::fetchFragmentLengths (alignmentParser.cpp:692-725): implied fragment length of every (fragment,
 * candidate) match in fragment-major order; returns the number of matches (lengths_out may be NULL) */
int64_t bayes_fetch_fragment_lengths(const bayes_raw_locus *locus, uint32_t *lengths_out, int64_t cap);

/* ---- candidate enumeration (host, C++): splice graph -> candidate transcripts ---------------- */
/* AllPaths(dot).findFullPaths(no_pre_mrna, max_candidate_number, log) (allPaths.h:95-98, allPaths.cpp:42-343; call site
 * assembler.cpp:1462-1463).  `dot` is one splice graph in the DOT mini-format the reference writes (assembler.cpp:764-861):
 *     digraph <name> {   <id> [label="ref;strand;start;end"] ...   pre_mrna [label=...]   <id>-><id> [coverage="n"] ...   }
 * Every source-to-sink path becomes a candidate; while there are more than max_candidate_number of them the junction
 * threshold rises and junctions with less coverage are dropped (allPaths.cpp:173-295).  Candidates come out in the
 * reference's order (sources in vertex order = node ids sorted as strings; depth first in edge-statement order). */
typedef struct bayes_path_set bayes_path_set;
int bayes_all_paths_build(const char *dot, int32_t no_pre_mrna, int32_t max_candidate_number, bayes_path_set **out);
void bayes_path_set_destroy(bayes_path_set *ps);
int32_t bayes_path_set_num_candidates(const bayes_path_set *ps);
int32_t bayes_path_set_num_exons(const bayes_path_set *ps);
/* vector<Candidate> (utils.h:109-117) as the arrays bayes_raw_locus takes: exon_ptr n_cand + 1, exon_start / exon_end (adjacent
 * exons merged, allPaths.cpp:237-244), Candidate::length, Candidate::isPreMrna; any pointer may be NULL */
int bayes_path_set_export(const bayes_path_set *ps, int32_t *exon_ptr, int32_t *exon_start, int32_t *exon_end, int32_t *cand_length,
                          int32_t *is_premrna);
/* the numbers of the log block (allPaths.cpp:132-134, :317, :301): vertices, sources, sinks of the graph as read,
 * junction-threshold adjustments, paths, and the final junction threshold */
int bayes_path_set_stats(const bayes_path_set *ps, int32_t *stats6);
/* Candidate::graph_name and the Vertex::reference / Vertex::strand shared by all vertices (valid until destroy) */
const char *bayes_path_set_graph_name(const bayes_path_set *ps);
const char *bayes_path_set_reference(const bayes_path_set *ps);
const char *bayes_path_set_strand(const bayes_path_set *ps);
/* Candidate::name of candidate i ("<graph>_seq<i>" or "<graph>_pre", allPaths.cpp:254, :270); returns its length */
int bayes_path_set_candidate_name(const bayes_path_set *ps, int32_t i, char *buf, int32_t cap);

/* ---- BAM front end (host, C++): TopHat2 BAM -> graphs with their fragment alignments -------------- */
/* What the reference does before its graph queue exists (assembler.cpp:84-624, 627-1164, 1167-1417), without BamTools /
 * samtools: a coordinate-sorted paired-end BAM is read (BGZF through zlib), the alignments the reference keeps are kept
 * (mapped, primary, mate mapped, mates on opposite strands of one reference, |insert| <= 700 000; generateSpliceGraphs
 * :329-352 / :400-470), duplicate pairs are removed (markDuplicates :157-249: same position, insert and CIGARs; a uniquely
 * mapping pair replaces a multi-mapping one) and, for strand-specific data ("first" | "second"), split into a plus and a minus
 * set (:424-470).  The de-duplicated alignments stay in memory in the order the reference writes them to its *_nd_*.bam. */
typedef struct bayes_bam_front bayes_bam_front;
int bayes_bam_front_open(const char *bam_path, const char *strand_specific, bayes_bam_front **out);
void bayes_bam_front_close(bayes_bam_front *f);
/* :384-388, :610-619: mapped read pairs, unique non-redundant (NH == 1) pairs, pairs written for graph construction */
int bayes_bam_front_counts(bayes_bam_front *f, double *total_mapped_pairs, double *unique_pairs, double *pairs_written);
/* alignments of a strand set: 0 = unstranded or plus, 1 = minus */
int64_t bayes_bam_front_num_alignments(bayes_bam_front *f, int32_t strand_file);
/* the strand set as SAM text, i.e. `samtools view` of the reference's *_nd_*.bam (:662-676): the input of CEM processsam,
 * which is not part of the reference repository and not restated here */
int bayes_bam_front_write_sam(bayes_bam_front *f, int32_t strand_file, const char *sam_path, int32_t with_header);
/* graphConstructorCallback :692-1125 + grapherCallback :1167-1319: parse the instance file processsam wrote for this strand set
 * (`strand` = "+", "-" or "." as passed to processsam -d), build the DOT strings, collapse overlapping graphs, fetch every
 * graph's fragments (both mates inside the graph, NH == 1) with their read probabilities (calculateSequencingProbability
 * :1322-1417) and write / append them to `graph_file_path` in the format `bayesembler_b200_cli --graph-file` reads.
 * Counts out (any may be NULL): graphs after collapsing, those with fragments, instances dropped for lack of a strand,
 * instances in the file. */
int bayes_bam_front_graphs(bayes_bam_front *f, int32_t strand_file, const char *instance_path, const char *strand, double exon_base_coverage,
                           int32_t no_pre_mrna, const char *graph_file_path, int32_t append, int32_t *n_graphs, int32_t *n_with_fragments,
                           int32_t *n_excluded, int32_t *cem_graphs);
/* calculateSequencingProbability (assembler.cpp:1322-1417): probability of the read's bases from its MD tag, its base
 * qualities (phred + 33 text) and its CIGAR (operations / lengths) */
int bayes_sequencing_probability(const char *md_tag, const char *qualities, int32_t n_cigar, const char *cigar_op, const int32_t *cigar_len, double *out);

#ifdef __cplusplus
}
#endif
#endif /* BAYESEMBLER_B200_H */
