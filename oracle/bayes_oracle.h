/*
 * oracle/bayes_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the reference's per-locus Gibbs sampler, the
 * checker the CUDA path is compared with.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * Parity status: PINNED against the reference itself -- tests/test_oracle_pin.py
 * compares every function below, in ORC_WS_SEQ mode, with the UNMODIFIED
 * reference translation units compiled here (oracle/_ref, recipe oracle/Makefile)
 * on the same MT19937 stream: identical integer outputs, identical number of
 * 32-bit words consumed, floating point equal to 1e-12.  The uniform->variate
 * transforms (Boost.Random in the reference; Boost is absent and unpinned) are
 * those of oracle/orc_rng.h on BOTH sides; see that header.
 */
#ifndef BAYES_ORACLE_H
#define BAYES_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* one CollapsedMap (utils.h:129-136) in CSR form; columns ascending inside a row */
typedef struct {
    int T;                 /* candidates (columns)              */
    int R;                 /* collapsed rows                    */
    const int *row_ptr;    /* R+1                               */
    const int *col_idx;    /* nnz                               */
    const double *val;     /* nnz, row-normalised probabilities */
    const int *row_count;  /* R, fragments per collapsed row    */
    const double *efflen;  /* T effective lengths               */
    double adjusted_fragment_count; /* truncated to int exactly as GibbsSampler's ctor does (gibbsSampler.h:51) */
    double gamma;          /* --dirichlet-parameter             */
} orc_locus;

/* rng_mode: 0 = ORC_WS_SEQ (MT19937 seeded with (uint32)seed), 1 = ORC_WS_SITE (Philox key = seed) */

/* assignmentGenerator.cpp:52-96.  in_cover: T flags or NULL.  Returns |cover|. */
int orc_min_read_cover(const orc_locus *loc, int rng_mode, uint64_t seed, int *in_cover, uint64_t *words);

/* simplexSizeGenerator.cpp:117-187.  table: T*T doubles, row k-1 at table + (k-1)*T; row_len: T ints. */
void orc_simplex_table(double pi, double gamma, int T, int num_fragments, double *table, int *row_len);

/* simplexSizeGenerator.cpp:201-224 (one draw; site = (SIMPLEX, iteration)). */
int orc_sample_simplex(const double *table, const int *row_len, int T, int count_plus_size, int rng_mode, uint64_t seed,
                       uint32_t iteration);

/* assignmentGenerator.cpp:213-260.  Returns |S+|. */
int orc_init_assignment(const orc_locus *loc, int rng_mode, uint64_t seed, int *counts_out, uint64_t *words);

/* assignmentGenerator.cpp:263-373.  Returns |S+|. */
int orc_generate_assignment(const orc_locus *loc, const double *theta, int rng_mode, uint64_t seed, uint32_t iteration,
                            int *counts_out, uint64_t *words);

/* expressionGenerator.cpp:77-132.  plus_order_out: b ints or NULL.  Returns |S+_expr|. */
int orc_generate_expression(int T, const int *counts, int b, double gamma, int rng_mode, uint64_t seed, uint32_t iteration,
                            double *theta_out, int *plus_order_out, uint64_t *words);

/*
 * gibbsSampler.cpp:55-148 driven like assembler.cpp:1548-1603.
 *   pi < 0  -> min read cover -> pi = |cover|/T clamped to 1-eps (assembler.cpp:1551-1559); the cover consumes RNG first
 *   pi >= 0 -> given (no cover, nothing consumed)
 * Outputs (any may be NULL): occ/mean/m2 = GibbsOutput accumulators (gibbsOutput.cpp:51-75), T each;
 * traces for the first trace_iters iterations: theta_trace[it*T+t], counts_trace[it*T+t] (counts AFTER the
 * assignment of iteration it), b_trace[it] (simplex size USED by iteration it); init_counts[T]; final pi; words.
 */
int orc_run_sampler(const orc_locus *loc, double pi, int burn_in, int samples, int rng_mode, uint64_t seed,
                    int *occ, double *mean, double *m2,
                    int trace_iters, double *theta_trace, int *counts_trace, int *b_trace, int *init_counts,
                    double *pi_out, uint64_t *words);

/* gibbsOutput.cpp:77-106.  Returns number kept. */
int orc_top_marginals(int T, int samples, const int *occ, const double *mean, const double *m2, double threshold,
                      int *idx_out, double *mean_out, double *sd_out);

/* raw generators, for the variate unit tests (fills out[n]) */
void orc_test_gamma(double alpha, int rng_mode, uint64_t seed, int n, double *out);
void orc_test_binomial(int nn, double p, int rng_mode, uint64_t seed, int n, int *out);
void orc_test_uniform_int(uint32_t range, int rng_mode, uint64_t seed, int n, int *out);
void orc_test_philox(uint64_t key, const uint32_t ctr[4], uint32_t out[4]);
void orc_test_mt(uint32_t seed, int n, uint32_t *out);

#ifdef __cplusplus
}
#endif
#endif
