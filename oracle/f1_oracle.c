/*
 * oracle/f1_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's likelihood builder, the checker the product's
 * bayes_collapsed_map_build() is compared with.  It keeps the reference's own data structure -- a dense row-major
 * matrix searched linearly for an equal row -- so that it is an independent check of the product's sparse / hashed
 * implementation.  Each function cites the reference lines it follows (paths relative to /root/reference/src).
 *
 * Parity status: PINNED against the reference itself -- tests/test_f1_likelihood.py compares every function below with
 * the UNMODIFIED reference translation units alignmentParser.cpp, sequencingModel.cpp, fragmentLengthModel.cpp compiled
 * here (oracle/_ref/libbayesref_f1.so, recipe oracle/Makefile) live and through tests/golden/f1_vectors.npz: identical
 * indices and counts, identical doubles.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const double double_underflow = DBL_MIN;        /* utils.h:52 */
static const double double_precision = DBL_EPSILON;    /* utils.h:54 */
static const double accumulated_precision_error = 100; /* utils.h:56 */

/* utils.h:58-61 */
static int double_compare(double a, double b) {
    double mn = a < b ? a : b;
    return (a == b) || (fabs(a - b) < fabs(mn) * double_precision * accumulated_precision_error);
}

/* GaussianFragmentLengthModel::pmf, fragmentLengthModel.cpp:40-49, 66-69 */
typedef struct {
    double mean, var, norm_const;
} gauss_model;

static void gauss_init(gauss_model *g, double mean, double sd) {
    g->mean = mean;
    g->var = pow(sd, 2);
    g->norm_const = 1 / (sd * sqrt(2 * 3.14159265358979323846264338327950288));
}
static double gauss_pmf(const gauss_model *g, unsigned fragment_length) {
    return g->norm_const * exp(-(pow((double)fragment_length - g->mean, 2) / (2 * g->var)));
}

/* SequencingModel, sequencingModel.cpp:34-76 */
typedef struct {
    int n;
    double *frag_length_prob, *frag_length_cum_prob, *three_prime_cum_prob;
} seq_model;

static void seq_init(seq_model *s, const gauss_model *g, int max_transcript_length) {
    s->n = max_transcript_length;
    s->frag_length_prob = (double *)calloc(s->n, sizeof(double));
    s->frag_length_cum_prob = (double *)calloc(s->n, sizeof(double));
    s->three_prime_cum_prob = (double *)calloc(s->n, sizeof(double));
    s->frag_length_prob[0] = gauss_pmf(g, 1);
    s->frag_length_cum_prob[0] = s->frag_length_prob[0];
    s->three_prime_cum_prob[0] = s->frag_length_prob[0];
    for (int i = 1; i < s->n; i++) {
        s->frag_length_prob[i] = gauss_pmf(g, i + 1);
        s->frag_length_cum_prob[i] = s->frag_length_prob[i] + s->frag_length_cum_prob[i - 1];
        s->three_prime_cum_prob[i] = s->frag_length_cum_prob[i] + s->three_prime_cum_prob[i - 1];
    }
}
static void seq_free(seq_model *s) { free(s->frag_length_prob); free(s->frag_length_cum_prob); free(s->three_prime_cum_prob); }
static double seq_three_prime(const seq_model *s, int pos, int len) { return s->frag_length_cum_prob[pos - 1] / s->three_prime_cum_prob[len - 1]; }
static double seq_length_prob(const seq_model *s, int fl, int pos) { return s->frag_length_prob[fl - 1] / s->frag_length_cum_prob[pos - 1]; }
static double seq_efflen(const seq_model *s, int len) { return s->three_prime_cum_prob[len - 1]; }

int orc_sequencing_tables(double mean, double sd, int max_transcript_length, double *three_prime_prob_at_len, double *length_prob_200,
                          double *efflen) {
    gauss_model g;
    seq_model s;
    gauss_init(&g, mean, sd);
    seq_init(&s, &g, max_transcript_length);
    for (int l = 1; l <= max_transcript_length; l++) {
        efflen[l - 1] = seq_efflen(&s, l);
        three_prime_prob_at_len[l - 1] = seq_three_prime(&s, l, max_transcript_length);
        length_prob_200[l - 1] = seq_length_prob(&s, l < 200 ? l : 200, l);
    }
    seq_free(&s);
    return 0;
}

/* GaussianFragmentLengthModel::estimateMedian, fragmentLengthModel.cpp:71-118, on (value, count) pairs sorted by value */
static double estimate_median(const double *value, const unsigned *count, int n) {
    int observation_count = 0;
    for (int i = 0; i < n; i++) observation_count += count[i];
    int median_observation_idx = observation_count % 2 == 0 ? (int)round((double)observation_count / 2.0) : (int)ceil((double)observation_count / 2.0);
    int it = 0, running = count[0];
    while (running < median_observation_idx) {
        it++;
        running += count[it];
    }
    if (observation_count % 2 == 0 && running == median_observation_idx) return (value[it] + value[it + 1]) / 2;
    return value[it];
}

static int cmp_pair(const void *a, const void *b) {
    double x = ((const double *)a)[0], y = ((const double *)b)[0];
    return x < y ? -1 : x > y;
}

/* GaussianFragmentLengthModel::estimateParameters, fragmentLengthModel.cpp:120-161; lengths ascending and distinct */
int orc_estimate_fragment_model(int n, const uint32_t *length, const uint32_t *count, double *median_out, double *mad_out) {
    double *v = (double *)malloc(sizeof(double) * n);
    for (int i = 0; i < n; i++) v[i] = (double)length[i];
    const double median = estimate_median(v, count, n);
    /* map<double, uint> abs_distances: merge equal distances, ascending */
    double *pairs = (double *)malloc(sizeof(double) * 2 * n);
    for (int i = 0; i < n; i++) { pairs[2 * i] = fabs((double)length[i] - median); pairs[2 * i + 1] = (double)count[i]; }
    qsort(pairs, n, 2 * sizeof(double), cmp_pair);
    double *dv = (double *)malloc(sizeof(double) * n);
    unsigned *dc = (unsigned *)malloc(sizeof(unsigned) * n);
    int m = 0;
    for (int i = 0; i < n; i++) {
        if (m > 0 && dv[m - 1] == pairs[2 * i]) dc[m - 1] += (unsigned)pairs[2 * i + 1];
        else { dv[m] = pairs[2 * i]; dc[m] = (unsigned)pairs[2 * i + 1]; m++; }
    }
    *median_out = median;
    *mad_out = estimate_median(dv, dc, m);
    free(v); free(pairs); free(dv); free(dc);
    return 0;
}

typedef struct {
    int first_exon_idx, last_exon_idx, fragment_length, three_prime_position;
} frag_match;

/* AlignmentParser::matchFragmentToCandidate, alignmentParser.cpp:42-249 */
static int match_fragment(const int *es, const int *ee, int n_exons, uint32_t pos1, uint32_t pos2, const char *op1, const uint32_t *len1,
                          int n1, const char *op2, const uint32_t *len2, int n2, int plus, frag_match *out) {
    int current_pos = (int)pos1 + 1;
    unsigned first_exon_idx = 0, current_exon_idx = 0;
    const char *op = op1;
    const uint32_t *ln = len1;
    int n = n1;
    int skip_fragment = 1;
    for (unsigned pair_idx = 0; pair_idx < 2; pair_idx++) {
        int first_match = 1;
        if (pair_idx == 1) { current_pos = (int)pos2 + 1; op = op2; ln = len2; n = n2; }
        for (int j = 0; j < n; j++) {
            skip_fragment = 1;
            if (op[j] == 'M') {
                if (first_match) {
                    for (unsigned k = first_exon_idx; k < (unsigned)n_exons; k++) {
                        if (current_pos >= es[k] && current_pos <= ee[k]) {
                            if (pair_idx == 0) first_exon_idx = k;
                            current_pos += (int)ln[j] - 1;
                            current_exon_idx = k;
                            first_match = 0;
                            skip_fragment = 0;
                            break;
                        }
                    }
                } else if (op[j - 1] != 'N') {
                    if (current_pos <= ee[current_exon_idx]) { current_pos += (int)ln[j] - 1; skip_fragment = 0; }
                } else {
                    if (es[current_exon_idx] == current_pos) { current_pos += (int)ln[j] - 1; skip_fragment = 0; }
                }
            } else if (op[j] == 'N') {
                if (ee[current_exon_idx] == current_pos) {
                    current_exon_idx += 1;
                    current_pos += (int)ln[j] + 1;
                    skip_fragment = current_exon_idx >= (unsigned)n_exons ? 1 : 0;
                }
            } else if (op[j] == 'I') {
                if (current_pos <= ee[current_exon_idx]) skip_fragment = 0;
            } else if (op[j] == 'D') {
                if (current_pos <= ee[current_exon_idx]) { current_pos += (int)ln[j] + 1; skip_fragment = 0; }
            } else {
                return -1;
            }
            if (skip_fragment) break;
        }
        if (!skip_fragment && current_pos > ee[current_exon_idx]) skip_fragment = 1;
        if (skip_fragment) break;
    }
    if (skip_fragment) return 0;
    out->first_exon_idx = (int)first_exon_idx;
    out->last_exon_idx = (int)current_exon_idx;
    const int read1 = (int)pos1 + 1;
    if (first_exon_idx == current_exon_idx) {
        out->fragment_length = current_pos - read1 + 1;
    } else {
        out->fragment_length = ee[first_exon_idx] - read1 + 1;
        for (unsigned j = first_exon_idx + 1; j < current_exon_idx; j++) out->fragment_length += ee[j] - es[j] + 1;
        out->fragment_length += current_pos - es[current_exon_idx] + 1;
    }
    if (plus) {
        out->three_prime_position = current_pos - es[current_exon_idx] + 1;
        for (unsigned j = 0; j < current_exon_idx; j++) out->three_prime_position += ee[j] - es[j] + 1;
    } else {
        out->three_prime_position = ee[first_exon_idx] - read1 + 1;
        for (unsigned j = first_exon_idx + 1; j < (unsigned)n_exons; j++) out->three_prime_position += ee[j] - es[j] + 1;
    }
    return 1;
}

/* AlignmentParser::calculateFragTranProbabilities, alignmentParser.cpp:251-689; same argument list as ref_collapsed_map
 * (oracle/ref_f1_driver.cpp) */
int orc_collapsed_map(int n_cand, const int *exon_ptr, const int *exon_start, const int *exon_end, const int *cand_length,
                      const int *is_premrna, const int *subgraph_id, int strand_plus, int n_frag, const uint32_t *read_pos,
                      const int *cigar_ptr, const char *cigar_op, const uint32_t *cigar_len, const double *frag_prob, double frag_mean,
                      double frag_sd, int max_transcript_length, int no_coverage_exclude, int n_subgraphs, int *subgraph_complexity,
                      int r_cap, int *T_out, int *R_out, double *P_out, int *counts_out, int *new_to_old, double *efflen_out) {
    (void)n_subgraphs;
    gauss_model g;
    seq_model sm;
    gauss_init(&g, frag_mean, frag_sd);
    seq_init(&sm, &g, max_transcript_length);
    const int C = n_cand;
    int cap = 500, row_idx = 0;
    double *P = (double *)calloc((size_t)cap * C, sizeof(double));
    unsigned *collapsed = (unsigned *)calloc(cap, sizeof(unsigned));
    char *excl = (char *)calloc(C, 1);
    int *down = (int *)malloc(sizeof(int) * C), *up = (int *)malloc(sizeof(int) * C);
    unsigned *cand_counts = (unsigned *)calloc(C, sizeof(unsigned));
    double *temp = (double *)malloc(sizeof(double) * C);
    for (int i = 0; i < C; i++) { up[i] = 2000000000; down[i] = -1; }
    int n_excl = 0, rc = 0;
    *T_out = 0;
    *R_out = 0;
    for (int f = 0; f < n_frag; f++) {                                                   /* :285-420 */
        int found = 0;
        double norm_const = 0;
        for (int i = 0; i < C; i++) temp[i] = 0;
        for (int i = 0; i < C; i++) {
            if (excl[i]) continue;
            frag_match m;
            const int a = cigar_ptr[2 * f], b = cigar_ptr[2 * f + 1], c = cigar_ptr[2 * f + 2];
            int ok = match_fragment(exon_start + exon_ptr[i], exon_end + exon_ptr[i], exon_ptr[i + 1] - exon_ptr[i], read_pos[2 * f],
                                    read_pos[2 * f + 1], cigar_op + a, cigar_len + a, b - a, cigar_op + b, cigar_len + b, c - b, strand_plus, &m);
            if (ok < 0) { rc = -2; goto done; }
            if (!ok) continue;
            found = 1;
            cand_counts[i]++;
            double p = 1;
            p *= frag_prob[f];
            p *= seq_three_prime(&sm, m.three_prime_position, cand_length[i]);
            p *= seq_length_prob(&sm, m.fragment_length, m.three_prime_position);
            temp[i] = p;
            norm_const += p;
            if (!no_coverage_exclude) {                                                    /* :325-353 */
                if (m.first_exon_idx <= down[i] || cand_counts[i] == 1) {
                    if (m.last_exon_idx > down[i]) down[i] = m.last_exon_idx;
                    if (m.first_exon_idx < up[i]) up[i] = m.first_exon_idx;
                } else if (!is_premrna[i]) {
                    excl[i] = 1;
                    n_excl++;
                }
            }
        }
        if (!found) continue;
        if (norm_const < double_underflow) continue;
        for (int j = 0; j < C; j++) temp[j] = temp[j] / norm_const;                        /* :369 */
        int i, row_match = 0;
        for (i = 0; i < row_idx; i++) {                                                    /* :373-389 */
            row_match = 1;
            for (int j = 0; j < C; j++)
                if (!double_compare(P[(size_t)i * C + j], temp[j])) { row_match = 0; break; }
            if (row_match) break;
        }
        if (row_match) {
            collapsed[i]++;
        } else {
            if (row_idx == cap) {
                cap += 500;
                P = (double *)realloc(P, sizeof(double) * (size_t)cap * C);
                collapsed = (unsigned *)realloc(collapsed, sizeof(unsigned) * cap);
            }
            memcpy(P + (size_t)row_idx * C, temp, sizeof(double) * C);
            collapsed[row_idx] = 1;
            row_idx++;
        }
    }
    if (!no_coverage_exclude) {                                                            /* :425-467 */
        for (int i = 0; i < C; i++) {
            if (excl[i]) continue;
            if (down[i] == -1) { excl[i] = 1; n_excl++; }
            else if (!is_premrna[i]) {
                const int n_exons = exon_ptr[i + 1] - exon_ptr[i];
                if (up[i] > 0) { excl[i] = 1; n_excl++; }
                else if (down[i] < n_exons - 1) { excl[i] = 1; n_excl++; }
            }
        }
    }
    if (n_excl == C) { rc = 3; goto done; }                                                /* :475-481 */
    {
        /* :484-560: subgraph complexity, index map, column compaction */
        int num_cols = 0;
        for (int i = 0; i < C; i++) {
            if (excl[i]) { if (subgraph_complexity && subgraph_id) subgraph_complexity[subgraph_id[i]]--; continue; }
            new_to_old[num_cols] = i;
            efflen_out[num_cols] = seq_efflen(&sm, cand_length[i]);
            for (int r = 0; r < row_idx; r++) P[(size_t)r * C + num_cols] = P[(size_t)r * C + i];
            num_cols++;
        }
        /* :572-619 */
        char *row_excl = (char *)calloc(row_idx > 0 ? row_idx : 1, 1);
        if (n_excl > 0) {
            double *norm_row = (double *)malloc(sizeof(double) * num_cols);
            for (int i = 0; i < row_idx; i++) {
                double row_sum = 0;
                for (int k = 0; k < num_cols; k++) row_sum += P[(size_t)i * C + k];
                if (row_sum < double_underflow) { row_excl[i] = 1; continue; }
                for (int k = 0; k < num_cols; k++) norm_row[k] = P[(size_t)i * C + k] / row_sum;
                int j, row_match = 0;
                for (j = 0; j < i; j++) {
                    row_match = 1;
                    for (int k = 0; k < num_cols; k++)
                        if (!double_compare(P[(size_t)j * C + k], norm_row[k])) { row_match = 0; break; }
                    if (row_match) break;
                }
                if (row_match) { collapsed[j] += collapsed[i]; row_excl[i] = 1; }
                else memcpy(P + (size_t)i * C, norm_row, sizeof(double) * num_cols);
            }
            free(norm_row);
        }
        /* :621-649 */
        int R = 0;
        for (int i = 0; i < row_idx; i++) {
            if (row_excl[i]) continue;
            if (R < r_cap) {
                for (int k = 0; k < num_cols; k++) P_out[(size_t)R * num_cols + k] = P[(size_t)i * C + k];
                counts_out[R] = (int)collapsed[i];
            }
            R++;
        }
        free(row_excl);
        *T_out = num_cols;
        *R_out = R;
        rc = R == 0 ? 4 : (R > r_cap ? -1 : 0);
        if (rc == 4) { *T_out = 0; }
    }
done:
    free(P); free(collapsed); free(excl); free(down); free(up); free(cand_counts); free(temp);
    seq_free(&sm);
    return rc;
}

/* AlignmentParser::fetchFragmentLengths, alignmentParser.cpp:692-725 */
long orc_fetch_fragment_lengths(int n_cand, const int *exon_ptr, const int *exon_start, const int *exon_end, int strand_plus, int n_frag,
                                const uint32_t *read_pos, const int *cigar_ptr, const char *cigar_op, const uint32_t *cigar_len,
                                uint32_t *lengths_out, long cap) {
    long n = 0;
    for (int f = 0; f < n_frag; f++)
        for (int i = 0; i < n_cand; i++) {
            frag_match m;
            const int a = cigar_ptr[2 * f], b = cigar_ptr[2 * f + 1], c = cigar_ptr[2 * f + 2];
            if (match_fragment(exon_start + exon_ptr[i], exon_end + exon_ptr[i], exon_ptr[i + 1] - exon_ptr[i], read_pos[2 * f],
                               read_pos[2 * f + 1], cigar_op + a, cigar_len + a, b - a, cigar_op + b, cigar_len + b, c - b, strand_plus, &m) > 0) {
                if (n < cap) lengths_out[n] = (uint32_t)m.fragment_length;
                n++;
            }
        }
    return n;
}
