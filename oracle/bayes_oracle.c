/*
 * oracle/bayes_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's per-locus Gibbs sampler (see
 * bayes_oracle.h for scope and pinning).  Each function cites the reference
 * lines it follows (paths relative to /root/reference/src).  The matrix is
 * CSR instead of the reference's dense column-major Eigen::MatrixXd; absent
 * cells are exact zeros in the reference, which add 0 to every sum, fail every
 * ">= double_underflow" test and never receive a draw, so iterating stored
 * cells in ascending column order is the same arithmetic in the same order.
 */
#include "bayes_oracle.h"

#include <assert.h>
#include <float.h>
#include <math.h>
#include <stdlib.h>

#include "orc_rng.h"

static const double double_underflow = DBL_MIN;              /* utils.h:52 */
static const double double_precision = DBL_EPSILON;          /* utils.h:54 */
static const double double_almost_one = 1.0 - DBL_EPSILON;   /* utils.h:55 */
static const double accumulated_precision_error = 100;       /* utils.h:56 */

/* utils.h:58-61 */
static int double_compare(double a, double b) {
    double mn = a < b ? a : b;
    return (a == b) || (fabs(a - b) < fabs(mn) * double_precision * accumulated_precision_error);
}

typedef struct {
    orc_ws ws;
    orc_mt mt;
} rng_t;

/*
 * Modes.  ORC_WS_SEQ (the reference's one sequential MT19937) and ORC_WS_SITE (the same draws from site-keyed Philox
 * streams) MUST share every line of the sampler below: the only place that tells them apart is the word source in
 * orc_rng.h.  That is what makes the chain  reference (SEQ) == oracle (SEQ),  oracle (SITE) == GPU replay mode  sound.
 * Never add a branch on SEQ vs SITE to this file.
 * ORC_WS_FAST is the product's production sampling mode: the same conditional distributions drawn in an order that
 * suits the GPU.  It replaces exactly two functions (init_assignment / generate_assignment -> assignment_fast) and
 * draws from sites like ORC_WS_SITE; everything else is shared.
 *
 * RULE: no function below may branch on the word-source mode except through orc_ws_site() / orc_ws_next() (and the
 * two dispatch points that select assignment_fast).  The parity argument reference(SEQ) == oracle(SEQ) and
 * oracle(SITE) == GPU(SITE) needs the two modes to run the SAME statements, differing only in where the words come from.
 */
static void rng_init(rng_t *r, int mode, uint64_t seed) {
    if (mode == ORC_WS_SEQ) {
        orc_mt_seed(&r->mt, (uint32_t)seed);
        orc_ws_init_seq(&r->ws, &r->mt);
    } else {
        orc_ws_init_site(&r->ws, seed);
    }
}

/* ------------------------------------------------------------------------------------------------
 * AssignmentGenerator::calcMinimumReadCover, assignmentGenerator.cpp:52-91
 * greedy set cover of rows by columns on the support matrix (P >= double_underflow, :61); every round
 * takes the column covering most still-uncovered rows (:65-66), ties broken by one uniform_int over the
 * ascending list of maximisers (:69-82), and un-marks the rows it covers (:85).
 * ---------------------------------------------------------------------------------------------- */
static int min_read_cover(const orc_locus *L, orc_ws *ws, int *in_cover) {
    const int R = L->R, T = L->T;
    int *logical = (int *)malloc(sizeof(int) * (R > 0 ? R : 1));
    int *covered = (int *)malloc(sizeof(int) * T);
    int *max_idx = (int *)malloc(sizeof(int) * T);
    int remaining = R, cover = 0;
    uint32_t round = 0;
    for (int i = 0; i < R; i++) logical[i] = 1;
    if (in_cover) for (int t = 0; t < T; t++) in_cover[t] = 0;
    while (remaining > 0) {
        for (int t = 0; t < T; t++) covered[t] = 0;
        for (int i = 0; i < R; i++)
            if (logical[i])
                for (int k = L->row_ptr[i]; k < L->row_ptr[i + 1]; k++)
                    if (L->val[k] >= double_underflow) covered[L->col_idx[k]] += 1;
        int max_val = covered[0];
        for (int t = 1; t < T; t++) if (covered[t] > max_val) max_val = covered[t];
        assert(max_val > 0);
        int n_max = 0;
        for (int t = 0; t < T; t++) if (covered[t] == max_val) max_idx[n_max++] = t;
        orc_ws_site(ws, ORC_PH_COVER, 0, round++, 0);
        int max_pos = max_idx[orc_uniform_int(ws, (uint32_t)n_max)];
        if (in_cover) in_cover[max_pos] = 1;
        cover++;
        for (int i = 0; i < R; i++)
            if (logical[i])
                for (int k = L->row_ptr[i]; k < L->row_ptr[i + 1]; k++)
                    if (L->col_idx[k] == max_pos && L->val[k] >= double_underflow) { logical[i] = 0; remaining--; break; }
    }
    free(logical); free(covered); free(max_idx);
    return cover;
}

int orc_min_read_cover(const orc_locus *loc, int rng_mode, uint64_t seed, int *in_cover, uint64_t *words) {
    rng_t r;
    rng_init(&r, rng_mode, seed);
    int c = min_read_cover(loc, &r.ws, in_cover);
    if (words) *words = r.ws.words;
    return c;
}

/* ------------------------------------------------------------------------------------------------
 * FixedBinomialFixedGammaSimplexSizeGenerator ctor, simplexSizeGenerator.cpp:117-187
 * row (k-1), k = |S+| = 1..T: CDF over simplex sizes j = k..T of
 *   log C(T-k, j-k) + j log(pi) + (T-j) log(1-pi) + lgamma(j gamma) - lgamma(N_f + j gamma)
 * accumulated by log(1+exp(.)) (:166), stopped early when the running sum stops changing under
 * double_compare (:169-172), then exp-normalised by the last running sum (:176-180).
 * ---------------------------------------------------------------------------------------------- */
static double lgam(double x) { int s; return lgamma_r(x, &s); }

void orc_simplex_table(double pi, double gamma, int T, int num_fragments, double *table, int *row_len) {
    for (int i = 0; i < T; i++) {
        const int k = i + 1;
        double *row = table + (size_t)i * T;
        int len = 0;
        double cardinal_eq_z_log = 0;
        double prob_z_log = k * log(pi) + (T - k) * log(1 - pi);
        double prob_t_log = lgam(k * gamma) - lgam(num_fragments + k * gamma);
        double prob_eq_z_log = cardinal_eq_z_log + prob_z_log + prob_t_log;
        double row_sum = prob_eq_z_log;
        row[len++] = row_sum;
        for (int j = k + 1; j < T + 1; j++) {
            cardinal_eq_z_log = lgam(T - k + 1) - (lgam(j - k + 1) + lgam(T - j + 1));
            prob_z_log = j * log(pi) + (T - j) * log(1 - pi);
            prob_t_log = lgam(j * gamma) - lgam(num_fragments + j * gamma);
            prob_eq_z_log = cardinal_eq_z_log + prob_z_log + prob_t_log;
            row_sum += log(1 + exp(prob_eq_z_log - row_sum));
            row[len++] = row_sum;
            if (double_compare(row[len - 1], row[len - 2])) break;
        }
        for (int j = 0; j < len; j++) row[j] = exp(row[j] - row_sum);
        assert(row[len - 1] > double_almost_one);
        row_len[i] = len;
    }
}

/* simplexSizeGenerator.cpp:214-224: one uniform_01, std::upper_bound on row |S+|-1, + |S+| */
static int sample_simplex(const double *table, const int *row_len, int T, int count_plus_size, orc_ws *ws, uint32_t iteration) {
    const double *row = table + (size_t)(count_plus_size - 1) * T;
    const int len = row_len[count_plus_size - 1];
    orc_ws_site(ws, ORC_PH_SIMPLEX, iteration, 0, 0);
    double u = orc_u01(orc_ws_next(ws));
    int lo = 0, hi = len; /* first element > u */
    while (lo < hi) {
        int mid = (lo + hi) / 2;
        if (u < row[mid]) hi = mid; else lo = mid + 1;
    }
    return lo + count_plus_size;
}

int orc_sample_simplex(const double *table, const int *row_len, int T, int count_plus_size, int rng_mode, uint64_t seed,
                       uint32_t iteration) {
    rng_t r;
    rng_init(&r, rng_mode, seed);
    return sample_simplex(table, row_len, T, count_plus_size, &r.ws, iteration);
}

static int assignment_fast(const orc_locus *L, const double *theta, orc_ws *ws, uint32_t iteration, int init, int *counts);

static int plus_size(const int *counts, int T) {
    int n = 0;
    for (int t = 0; t < T; t++) n += counts[t] > 0;
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * AssignmentGenerator::initAssignment, assignmentGenerator.cpp:213-260
 * per row a chain of conditional binomials over the columns with P_ij >= double_underflow, using P_ij
 * itself (rows are already normalised; no division by the row sum here), divider starting at 1 (:222)
 * and reduced by every column's probability (:249).  A binomial is drawn even when nothing remains
 * (:230-233); with n == 0 our binomial consumes nothing on either side.
 * If rounding leaves fragments unassigned the reference aborts (assert :253); to conserve the total the
 * restatement (and the product) give them to the last eligible column instead.
 * ---------------------------------------------------------------------------------------------- */
static void init_assignment(const orc_locus *L, orc_ws *ws, int *counts) {
    for (int t = 0; t < L->T; t++) counts[t] = 0;
    for (int i = 0; i < L->R; i++) {
        int num_samples = L->row_count[i];
        double divider = 1;
        int last = -1;
        for (int k = L->row_ptr[i]; k < L->row_ptr[i + 1]; k++) {
            const int j = L->col_idx[k];
            const double norm_probability = L->val[k];
            if (norm_probability >= double_underflow) {
                orc_ws_site(ws, ORC_PH_INIT_ASSIGN, 0, (uint32_t)i, (uint32_t)j);
                int num_maps = orc_binomial(ws, num_samples, norm_probability / divider);
                counts[j] += num_maps;
                num_samples -= num_maps;
                last = j;
            }
            divider -= norm_probability;
        }
        if (num_samples > 0 && last >= 0) counts[last] += num_samples;
    }
}

int orc_init_assignment(const orc_locus *loc, int rng_mode, uint64_t seed, int *counts_out, uint64_t *words) {
    rng_t r;
    rng_init(&r, rng_mode, seed);
    if (rng_mode == ORC_WS_FAST) assignment_fast(loc, 0, &r.ws, 0, 1, counts_out);
    else init_assignment(loc, &r.ws, counts_out);
    if (words) *words = r.ws.words;
    return plus_size(counts_out, loc->T);
}

/* ------------------------------------------------------------------------------------------------
 * Production ("fast") mode of the product, ORC_WS_FAST: an exact draw from the same per-row multinomials
 * as initAssignment / generateAssignment (assignmentGenerator.cpp:213-260, 263-373) that is not tied to the
 * reference's sequential order of draws (outside replay the sampler may pick any exact
 * multinomial per row).  Per row i with n fragments, over its stored non-zero cells in ascending column order:
 *   p_j = P_ij * theta_j (init: theta == 1); the cells IN PLAY are those with p_j > 0, their running sums
 *   cum_1 < ... < cum_k, Z = the sum of the row (in the order of the product's warp scans, see below);
 *   rows with a single stored non-zero cell give it everything (whatever theta is: they are folded by the packer);
 *   k == 1: everything goes to that cell;
 *   n < 20 (the reference's threshold, :289): the row owns n consecutive 32-bit words of the iteration's word
 *     POOL (rows with n < 20 in row order, words of block b = Philox block (b & 4095) of site
 *     (POOL, iteration, b >> 12)); word w lands in the first cell j < k with
 *     w * 2^-32 < cum_j * (1/Z), i.e. w <= ceil(cum_j (1/Z) 2^32 - 1), else in cell k (:302-334 with unsorted
 *     uniforms and one reciprocal per row);
 *   20 <= n <= direct_max(T): fragment by fragment, one word each from the row's own site (DIRECT, iteration, row), located
 *     like the words of the small rows;
 *   n > direct_max(T): a balanced binary TREE over the cells in play instead of the chain of conditional binomials
 *     (:336-365): node (a, len) holds m fragments for cells [a, a + len); its left half [a, a + len/2) receives
 *     Binomial(m, (cum(a + len/2) - cum(a)) / (cum(a + len) - cum(a))), the right half the rest; one site per node
 *     (TREE | node << 8, iteration, row, chunk), node = a | (len - 1) << 12.  Depth log2 k instead of k dependent
 *     draws.  The row's fragments enter the tree in ceil(n / 64) independent chunks of equal size +- 1 (n <= 8192; a
 *     larger row enters whole).
 * ---------------------------------------------------------------------------------------------- */
static uint32_t sat_u32_ceil(double x) { /* __double2uint_ru */
    if (!(x > 0.0)) return 0u;
    if (x >= 4294967295.0) return 0xffffffffu;
    return (uint32_t)ceil(x);
}

static void tree_split(orc_ws *ws, uint32_t phase, uint32_t iteration, uint32_t row, uint32_t chunk, const double *cum,
                       const int *cell_col, int a, int len, int m, int *counts) {
    if (m <= 0) return;
    if (len == 1) { counts[cell_col[a]] += m; return; }
    const int half = len >> 1;
    const double Pa = a > 0 ? cum[a - 1] : 0.0, Pm = cum[a + half - 1], Pb = cum[a + len - 1];
    const double p_left = (Pm - Pa) / (Pb - Pa);
    orc_ws_site(ws, phase | ((uint32_t)a | ((uint32_t)(len - 1) << 12)) << 8, iteration, row, chunk);
    const int left = orc_binomial(ws, m, p_left);
    tree_split(ws, phase, iteration, row, chunk, cum, cell_col, a, half, left, counts);
    tree_split(ws, phase, iteration, row, chunk, cum, cell_col, a + half, len - half, m - left, counts);
}

/* chunks a row with n >= 20 fragments is drawn in: ceil(n / 64) while that is at most 128 (every binomial then has
 * n <= 64, the inversion regime); a row above 8192 fragments is drawn whole (its binomials are BTRD's either way, and
 * one is cheaper than 128) */
static int tree_chunks(int n) {
    const int m = (n + 63) / 64;
    return m <= 128 ? m : 1;
}

/* Rows with >= 20 fragments and at most this many are drawn fragment by fragment (one uniform word each, located in the
 * row's running sums like the words of the small rows), the larger ones by the splitting tree.  A tree costs about a binomial
 * per cell in play and chunk of 64 fragments and is log2(cells) dependent levels deep; a fragment costs one binary search and
 * depends on nothing, which is cheaper for all but the largest rows of a locus with hundreds of candidates. */
static int direct_max(int T) { return T > 32 ? 4096 : 0; }

static int assignment_fast(const orc_locus *L, const double *theta, orc_ws *ws, uint32_t iteration, int init, int *counts) {
    const uint32_t ph_pool = init ? ORC_PH_INIT_POOL : ORC_PH_POOL, ph_tree = init ? ORC_PH_INIT_TREE : ORC_PH_TREE;
    const uint32_t ph_direct = init ? ORC_PH_INIT_DIRECT : ORC_PH_DIRECT;
    double *cum = (double *)malloc(sizeof(double) * (size_t)(L->T > 0 ? L->T : 1));
    int *cell_col = (int *)malloc(sizeof(int) * (size_t)(L->T > 0 ? L->T : 1));
    uint64_t pool_at = 0; /* next word of the pool */
    int bad = 0;
    for (int t = 0; t < L->T; t++) counts[t] = 0;
    for (int i = 0; i < L->R; i++) {
        const int beg = L->row_ptr[i], end = L->row_ptr[i + 1];
        const int n = L->row_count[i];
        int stored = 0, only = -1;
        for (int k = beg; k < end; k++) if (L->val[k] != 0.0) { stored++; only = L->col_idx[k]; }
        if (stored == 1) { counts[only] += n; continue; }
        int kk = 0;
        double run = 0;
        if (L->T <= 32) {
            /* Loci with up to 32 candidates: the product sums the row with a warp-shuffle inclusive scan over the row's
             * stored cells (distances 1, 2, 4, 8, 16; a cell with theta == 0 contributes an exact zero), i.e. in a fixed
             * tree order instead of left to right.  Same real-number sums, a different but fully specified rounding. */
            double x[32];
            int cc[32], m = 0;
            for (int k = beg; k < end; k++) {
                if (L->val[k] == 0.0) continue;
                cc[m] = L->col_idx[k];
                x[m++] = L->val[k] * (init ? 1.0 : theta[L->col_idx[k]]);
            }
            int live[32];
            for (int j = 0; j < m; j++) live[j] = x[j] > 0.0;
            for (int d = 1; d < 32; d <<= 1)
                for (int j = m - 1; j >= d; j--) x[j] += x[j - d];
            for (int j = 0; j < m; j++)
                if (live[j]) { cum[kk] = x[j]; cell_col[kk] = cc[j]; kk++; }
            run = m > 0 ? x[m - 1] : 0.0; /* the row sum is the scan's value at the row's last stored cell */
        } else {
            /* Loci with more than 32 candidates (gibbs_rows.cu, one warp per row): the lanes hold the row's stored cells whose
             * candidate is in the simplex (theta != 0), 32 at a time, in ascending candidate order within every block of 1024
             * candidates; a chunk of 32 is summed by a warp-shuffle inclusive scan (distances 1, 2, 4, 8, 16), the chunks are
             * chained left to right: cum = (sum of the chunks before) + (scan value inside the chunk). */
            int k = beg;
            while (k < end) {
                double x[32];
                int cc[32], live[32], m = 0;
                const int block = L->col_idx[k] >> 10;
                for (; k < end && m < 32 && (L->col_idx[k] >> 10) == block; k++) {
                    if (L->val[k] == 0.0) continue;
                    if (!init && theta[L->col_idx[k]] == 0.0) continue;
                    cc[m] = L->col_idx[k];
                    x[m++] = L->val[k] * (init ? 1.0 : theta[L->col_idx[k]]);
                }
                if (m == 0) continue;
                for (int j = 0; j < m; j++) live[j] = x[j] > 0.0;
                for (int d = 1; d < 32; d <<= 1)
                    for (int j = m - 1; j >= d; j--) x[j] += x[j - d];
                for (int j = 0; j < m; j++)
                    if (live[j]) { cum[kk] = run + x[j]; cell_col[kk] = cc[j]; kk++; }
                run = run + x[m - 1];
            }
        }
        const uint64_t my_words = pool_at;
        if (n < 20) pool_at += (uint64_t)n; /* the row's words are its own whether it needs them or not */
        if (kk == 0) { bad = 1; continue; } /* reference: assert(normalisation_constant >= double_underflow) (:284) */
        if (kk == 1) { counts[cell_col[0]] += n; continue; }
        if (n < 20) {
            const double rZ = 1.0 / run;
            for (int f = 0; f < n; f++) {
                const uint64_t word = my_words + (uint64_t)f, blk = word >> 2;
                orc_ws_site(ws, ph_pool, iteration, (uint32_t)(blk >> 12), 0);
                ws->k = (uint32_t)((blk & 4095u) << 2); /* block (blk & 4095) of the site */
                uint32_t b[4];
                orc_ws_block(ws, b);
                const uint32_t w = b[word & 3u];
                int j = 0;
                for (; j < kk - 1; j++)
                    if (w <= sat_u32_ceil(fma(cum[j] * rZ, 4294967296.0, -1.0))) break;
                counts[cell_col[j]] += 1;
            }
        } else if (n <= direct_max(L->T)) {
            /* fragment by fragment: fragment f draws word (f & 3) of block (f >> 2) of the row's own site and lands like the
             * words of a row with < 20 fragments do */
            const double rZ = 1.0 / run;
            for (int f = 0; f < n; f++) {
                const uint32_t blk = (uint32_t)f >> 2;
                orc_ws_site(ws, ph_direct, iteration, (uint32_t)i, blk >> 12);
                ws->k = (blk & 4095u) << 2;
                uint32_t b[4];
                orc_ws_block(ws, b);
                const uint32_t w = b[f & 3];
                int j = 0;
                for (; j < kk - 1; j++)
                    if (w <= sat_u32_ceil(fma(cum[j] * rZ, 4294967296.0, -1.0))) break;
                counts[cell_col[j]] += 1;
            }
        } else {
            /* the n fragments are drawn as independent chunks of at most 64 (the sum of independent multinomials with the
             * same probabilities is the multinomial of the sum): every binomial then stays in the small-n inversion regime */
            const int m = tree_chunks(n);
            for (int c = 0; c < m; c++)
                tree_split(ws, ph_tree, iteration, (uint32_t)i, (uint32_t)c, cum, cell_col, 0, kk, n / m + (c < n % m), counts);
        }
    }
    free(cum); free(cell_col);
    return bad;
}

static int cmp_double(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

/* ------------------------------------------------------------------------------------------------
 * AssignmentGenerator::generateAssignment, assignmentGenerator.cpp:263-373
 * per row: p_j = P_ij * theta_j (:280), Z = sum p (:282); with fewer than 20 fragments (:289) draw that
 * many uniform_01, sort them and walk the CDF cum += p_j / Z (:291-334), each uniform going to the first
 * column whose cumulative sum exceeds it; otherwise a chain of conditional binomials with
 * norm_p = p_j / Z, p = norm_p / divider over columns with norm_p >= double_underflow (:336-365).
 * ---------------------------------------------------------------------------------------------- */
static void generate_assignment(const orc_locus *L, const double *theta, orc_ws *ws, uint32_t iteration, int *counts) {
    double uniform_samples[20];
    for (int t = 0; t < L->T; t++) counts[t] = 0;
    for (int i = 0; i < L->R; i++) {
        const int beg = L->row_ptr[i], end = L->row_ptr[i + 1];
        double normalisation_constant = 0;
        for (int k = beg; k < end; k++) normalisation_constant += L->val[k] * theta[L->col_idx[k]];
        assert(normalisation_constant >= double_underflow);
        int num_samples = L->row_count[i];
        if (num_samples < 20) {
            orc_ws_site(ws, ORC_PH_ASSIGN, iteration, (uint32_t)i, 0);
            for (int j = 0; j < num_samples; j++) uniform_samples[j] = orc_u01(orc_ws_next(ws));
            qsort(uniform_samples, (size_t)num_samples, sizeof(double), cmp_double);
            double cum_sum = 0;
            int index_num = 0;
            for (int k = beg; k < end && index_num < num_samples; k++) {
                cum_sum += (L->val[k] * theta[L->col_idx[k]]) / normalisation_constant;
                int samp_counter = 0;
                while (index_num < num_samples && uniform_samples[index_num] < cum_sum) { samp_counter++; index_num++; }
                counts[L->col_idx[k]] += samp_counter;
            }
            /* a uniform >= the final cum_sum is silently dropped by the reference (:302-334); cannot happen
             * with 32-bit uniforms (max 1-2^-32) unless the row sum is off by more than 2^-32 */
        } else {
            double divider = 1;
            int last = -1;
            for (int k = beg; k < end; k++) {
                const int j = L->col_idx[k];
                const double norm_probability = (L->val[k] * theta[j]) / normalisation_constant;
                if (norm_probability >= double_underflow) {
                    orc_ws_site(ws, ORC_PH_ASSIGN, iteration, (uint32_t)i, (uint32_t)j);
                    int num_maps = orc_binomial(ws, num_samples, norm_probability / divider);
                    counts[j] += num_maps;
                    num_samples -= num_maps;
                    last = j;
                }
                divider -= norm_probability;
            }
            if (num_samples > 0 && last >= 0) counts[last] += num_samples; /* reference: assert (:363) */
        }
    }
}

int orc_generate_assignment(const orc_locus *loc, const double *theta, int rng_mode, uint64_t seed, uint32_t iteration,
                            int *counts_out, uint64_t *words) {
    rng_t r;
    rng_init(&r, rng_mode, seed);
    if (rng_mode == ORC_WS_FAST) assignment_fast(loc, theta, &r.ws, iteration, 0, counts_out);
    else generate_assignment(loc, theta, &r.ws, iteration, counts_out);
    if (words) *words = r.ws.words;
    return plus_size(counts_out, loc->T);
}

/* ------------------------------------------------------------------------------------------------
 * ExpressionGenerator::generateExpression, expressionGenerator.cpp:77-132
 * Gamma(count_t + gamma, 1) for t in S+ ascending (:85-98; CountValueContainer::fetchNullSet sorted it,
 * valueContainer.cpp:141); then for every extra simplex slot one uniform_int over the current ascending
 * null list (:106-109), a Gamma(gamma, 1) for the picked transcript (:113) and removal of the pick from
 * the list (:126); finally theta = g / sum g (:129, valueContainer.cpp:216-224).
 * ---------------------------------------------------------------------------------------------- */
static int generate_expression(int T, const int *counts, int b, double gamma, orc_ws *ws, uint32_t iteration,
                               double *theta, int *plus_order, int *null_scratch) {
    double norm_const_expression = 0;
    int n_plus = 0, n_null = 0;
    for (int t = 0; t < T; t++) theta[t] = 0;
    for (int t = 0; t < T; t++) {
        if (counts[t] > 0) {
            orc_ws_site(ws, ORC_PH_GAMMA, iteration, (uint32_t)t, 0);
            double gamma_sample = orc_gamma(ws, counts[t] + gamma);
            norm_const_expression += gamma_sample;
            theta[t] = gamma_sample;
            if (plus_order) plus_order[n_plus] = t;
            n_plus++;
        } else {
            null_scratch[n_null++] = t;
        }
    }
    const int count_plus = n_plus;
    for (int i = count_plus; i < b; i++) {
        orc_ws_site(ws, ORC_PH_NULLPICK, iteration, (uint32_t)(i - count_plus), 0);
        int s_null_idx = orc_uniform_int(ws, (uint32_t)n_null);
        int trans_id = null_scratch[s_null_idx];
        orc_ws_site(ws, ORC_PH_GAMMA, iteration, (uint32_t)trans_id, 0);
        double gamma_base_sample = orc_gamma(ws, gamma);
        assert(gamma_base_sample >= double_underflow);
        norm_const_expression += gamma_base_sample;
        theta[trans_id] = gamma_base_sample;
        if (plus_order) plus_order[n_plus] = trans_id;
        n_plus++;
        for (int k = s_null_idx; k + 1 < n_null; k++) null_scratch[k] = null_scratch[k + 1];
        n_null--;
    }
    assert(norm_const_expression >= double_underflow);
    for (int t = 0; t < T; t++) if (theta[t] != 0) theta[t] /= norm_const_expression;
    return n_plus;
}

int orc_generate_expression(int T, const int *counts, int b, double gamma, int rng_mode, uint64_t seed, uint32_t iteration,
                            double *theta_out, int *plus_order_out, uint64_t *words) {
    rng_t r;
    rng_init(&r, rng_mode, seed);
    int *scratch = (int *)malloc(sizeof(int) * T);
    int n = generate_expression(T, counts, b, gamma, &r.ws, iteration, theta_out, plus_order_out, scratch);
    free(scratch);
    if (words) *words = r.ws.words;
    return n;
}

/* GibbsOutput::addSample, gibbsOutput.cpp:51-75: Welford mean / M2 of FPKM conditional on inclusion */
static void add_sample(int T, const double *theta, const double *fpkm, int *occ, double *mean, double *m2) {
    for (int idx = 0; idx < T; idx++) {
        if (theta[idx] == 0) continue; /* not in expression.plus */
        occ[idx]++;
        if (occ[idx] == 1) {
            mean[idx] = fpkm[idx];
        } else {
            double delta = fpkm[idx] - mean[idx];
            mean[idx] += delta / occ[idx];
            m2[idx] += delta * (fpkm[idx] - mean[idx]);
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * GibbsSampler::runSampler, gibbsSampler.cpp:55-148, set up like assembler.cpp:1548-1603
 * ---------------------------------------------------------------------------------------------- */
int orc_run_sampler(const orc_locus *L, double pi, int burn_in, int samples, int rng_mode, uint64_t seed,
                    int *occ, double *mean, double *m2,
                    int trace_iters, double *theta_trace, int *counts_trace, int *b_trace, int *init_counts,
                    double *pi_out, uint64_t *words) {
    const int T = L->T;
    rng_t r;
    rng_init(&r, rng_mode, seed);
    int num_fragments = 0;
    for (int i = 0; i < L->R; i++) num_fragments += L->row_count[i];
    const int total_num_fragments = (int)L->adjusted_fragment_count; /* double -> int at the ctor call, assembler.cpp:1577 */

    if (pi < 0) {
        int cover = min_read_cover(L, &r.ws, 0);           /* assembler.cpp:1551 */
        pi = (double)cover / T;                             /* :1552 */
        if (pi > double_almost_one) pi = double_almost_one; /* :1554-1557 */
    }
    if (pi_out) *pi_out = pi;

    int *occ_l = (int *)calloc((size_t)T, sizeof(int));
    double *mean_l = (double *)calloc((size_t)T, sizeof(double));
    double *m2_l = (double *)calloc((size_t)T, sizeof(double));

    if (T == 1) { /* gibbsSampler.cpp:57-67 */
        double single_fpkm = (num_fragments * 1e9) / (total_num_fragments * L->efflen[0]);
        double one = 1.0;
        for (int i = 0; i < samples; i++) add_sample(1, &one, &single_fpkm, occ_l, mean_l, m2_l);
    } else {
        double *table = (double *)malloc(sizeof(double) * (size_t)T * T);
        int *row_len = (int *)malloc(sizeof(int) * T);
        orc_simplex_table(pi, L->gamma, T, num_fragments, table, row_len); /* assembler.cpp:1563 */
        int *counts = (int *)malloc(sizeof(int) * T);
        int *scratch = (int *)malloc(sizeof(int) * T);
        double *theta = (double *)malloc(sizeof(double) * T);
        double *fpkm = (double *)malloc(sizeof(double) * T);

        if (rng_mode == ORC_WS_FAST) assignment_fast(L, 0, &r.ws, 0, 1, counts);
        else init_assignment(L, &r.ws, counts);                                             /* gibbsSampler.cpp:90 */
        if (init_counts) for (int t = 0; t < T; t++) init_counts[t] = counts[t];
        int b = sample_simplex(table, row_len, T, plus_size(counts, T), &r.ws, 0);          /* :96 */

        const int total = burn_in + samples;
        for (int it = 0; it < total; it++) {
            if (it < trace_iters && b_trace) b_trace[it] = b;
            generate_expression(T, counts, b, L->gamma, &r.ws, (uint32_t)it, theta, 0, scratch); /* :105 / :126 */
            if (rng_mode == ORC_WS_FAST) assignment_fast(L, theta, &r.ws, (uint32_t)it, 0, counts);
            else generate_assignment(L, theta, &r.ws, (uint32_t)it, counts);                      /* :106 / :127 */
            b = sample_simplex(table, row_len, T, plus_size(counts, T), &r.ws, (uint32_t)it + 1); /* :107 / :128 */
            if (it < trace_iters) {
                if (theta_trace) for (int t = 0; t < T; t++) theta_trace[(size_t)it * T + t] = theta[t];
                if (counts_trace) for (int t = 0; t < T; t++) counts_trace[(size_t)it * T + t] = counts[t];
            }
            if (it >= burn_in) {
                /* ExpressionValueContainer::getFPKM, valueContainer.cpp:226-239 */
                for (int t = 0; t < T; t++)
                    fpkm[t] = theta[t] != 0 ? (theta[t] * num_fragments * 1e9) / (L->efflen[t] * total_num_fragments) : 0;
                add_sample(T, theta, fpkm, occ_l, mean_l, m2_l);                                  /* :131-133 */
            }
        }
        free(table); free(row_len); free(counts); free(scratch); free(theta); free(fpkm);
    }
    for (int t = 0; t < T; t++) {
        if (occ) occ[t] = occ_l[t];
        if (mean) mean[t] = mean_l[t];
        if (m2) m2[t] = m2_l[t];
    }
    free(occ_l); free(mean_l); free(m2_l);
    if (words) *words = r.ws.words;
    return 0;
}

/* GibbsOutput::getTopMarginals, gibbsOutput.cpp:77-106 */
int orc_top_marginals(int T, int samples, const int *occ, const double *mean, const double *m2, double threshold,
                      int *idx_out, double *mean_out, double *sd_out) {
    double count_threshold = threshold * samples;
    int n = 0;
    for (int i = 0; i < T; i++) {
        if (occ[i] >= count_threshold) {
            idx_out[n] = i;
            mean_out[n] = mean[i];
            sd_out[n] = sqrt(m2[i] / (occ[i] - 1));
            n++;
        }
    }
    return n;
}

/* ------------------------------------------------------------------------------- variate test hooks */
void orc_test_gamma(double alpha, int rng_mode, uint64_t seed, int n, double *out) {
    rng_t r;
    rng_init(&r, rng_mode, seed);
    for (int i = 0; i < n; i++) {
        orc_ws_site(&r.ws, ORC_PH_GAMMA, 0, (uint32_t)i, 0);
        out[i] = orc_gamma(&r.ws, alpha);
    }
}

void orc_test_binomial(int nn, double p, int rng_mode, uint64_t seed, int n, int *out) {
    rng_t r;
    rng_init(&r, rng_mode, seed);
    for (int i = 0; i < n; i++) {
        orc_ws_site(&r.ws, ORC_PH_ASSIGN, 0, (uint32_t)i, 0);
        out[i] = orc_binomial(&r.ws, nn, p);
    }
}

void orc_test_uniform_int(uint32_t range, int rng_mode, uint64_t seed, int n, int *out) {
    rng_t r;
    rng_init(&r, rng_mode, seed);
    for (int i = 0; i < n; i++) {
        orc_ws_site(&r.ws, ORC_PH_NULLPICK, 0, (uint32_t)i, 0);
        out[i] = orc_uniform_int(&r.ws, range);
    }
}

void orc_test_philox(uint64_t key, const uint32_t ctr[4], uint32_t out[4]) {
    uint32_t k[2] = { (uint32_t)key, (uint32_t)(key >> 32) };
    orc_philox4x32_10(k, ctr, out);
}

void orc_test_mt(uint32_t seed, int n, uint32_t *out) {
    orc_mt m;
    orc_mt_seed(&m, seed);
    for (int i = 0; i < n; i++) out[i] = orc_mt_next(&m);
}
