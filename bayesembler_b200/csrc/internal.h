// Internal structures shared by the host packer, the C ABI and the CUDA kernels.
#pragma once

#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/bayesembler_b200.h"

namespace bayes {

constexpr int kMaxT = 4096;          // per-transcript state lives in shared memory (44 B per candidate)
constexpr int kMaxBlock = 256;       // threads per (locus, chain) block: one thread per packed row up to this many
constexpr int kSmallCount = 20;      // assignmentGenerator.cpp:289: rows with fewer fragments draw uniforms
constexpr int kTableSmemEntries = 1024;   // simplex CDF tables up to this many entries are staged in shared memory
constexpr int kMatrixSmemBytes = 96 * 1024;  // loci whose packed matrix exceeds this stream it from L2/HBM

// Device-side descriptor of one packed locus.  All *_off fields index the context-wide concatenated arrays.
struct LocusDev {
    int32_t T;          // candidates
    int32_t R;          // packed rows (rows with >= 2 cells; single-cell rows are folded into base counts)
    int32_t n_slabs;    // ceil(R / 32) slabs of 32 rows, stored column-major inside a slab (SELL-32)
    int32_t N_f;        // fragments in the locus (sum of all row counts, folded rows included)
    int32_t N_total;    // (int) total_num_fragments
    int32_t burn;
    int32_t samples;
    int32_t tab_len;    // entries in the simplex CDF table
    int32_t n_cells;    // padded cells = 32 * sum of slab widths
    int32_t mat_smem;   // 1: matrix staged in shared memory
    int32_t tab_smem;   // 1: CDF table staged in shared memory
    int32_t smem_bytes; // dynamic shared memory this locus needs
    double gamma;
    int64_t slab_off;   // -> slab_cell_off[n_slabs + 1] (cell offsets relative to cell_off)
    int64_t cell_off;   // -> val / col
    int64_t row_off;    // -> row_n / row_nnz / row_id   (n_slabs * 32 entries, padded rows have n = nnz = 0)
    int64_t t_off;      // -> efflen / base_counts
    int64_t tab_off;    // -> tab_val
    int64_t tabrow_off; // -> tab_row[T + 1]
};

struct JobDev {
    int32_t locus;
    int32_t chain;
    uint64_t key;       // Philox key of the chain
    int64_t out_off;    // -> out_occ / out_mean / out_m2 / out_counts (T entries each)
};

struct KernelArgs {
    const LocusDev *loci;
    const JobDev *jobs;
    const double *val;
    const uint16_t *col;
    const int32_t *slab_cell_off;
    const int32_t *row_n;
    const int32_t *row_nnz;
    const int32_t *row_id;
    const double *efflen;
    const int32_t *base_counts;
    const double *tab_val;
    const int32_t *tab_row;
    int32_t *out_occ;
    double *out_mean;
    double *out_m2;
    int32_t *out_counts;
    // replay trace of one job (trace_job < 0: off)
    int64_t trace_job;
    int32_t trace_iters;
    double *tr_theta;
    int32_t *tr_counts;
    int32_t *tr_b;
    int32_t *tr_init;
};

// Host-side result of packing one locus.
struct PackedLocus {
    LocusDev d;
    std::vector<double> val;
    std::vector<uint16_t> col;
    std::vector<int32_t> slab_cell_off, row_n, row_nnz, row_id, base_counts, tab_row;
    std::vector<double> efflen, tab_val;
    double pi = 0;
    int32_t cover = 0;
    int32_t folded = 0;
    int32_t n_chains = 1;
    uint64_t seed = 0;
    double cost = 0;
    int32_t block = 32;  // threads per block chosen for this locus
};

// pack.cpp
int pack_locus(const bayes_locus_desc &in, PackedLocus *out, std::string *err);
int validate_locus(const bayes_locus_desc &in, std::string *err);
size_t locus_smem_bytes(const LocusDev &d);

// host_steps.cpp
int min_read_cover(const bayes_locus_desc &in, uint64_t key, int32_t *in_cover);
void simplex_table(double pi, double gamma, int T, int num_fragments, std::vector<int32_t> *row_off, std::vector<double> *table);

// gibbs_kernel.cu
int launch_gibbs(const KernelArgs &args, int64_t job_base, int64_t n_jobs, int block, bool mat_smem, size_t smem_bytes, void *stream);
int launch_step_assignment(const KernelArgs &args, bool init, const double *d_theta, uint32_t iteration, int32_t *d_counts_out,
                           size_t smem_bytes, void *stream);
int launch_step_expression(int T, const int32_t *d_counts, int b, double gamma, uint64_t key, uint32_t iteration, double *d_theta,
                           void *stream);
int launch_step_variates(int kind, double a, double p, uint64_t key, int n, double *d_out, void *stream);
int configure_kernels();

void set_error(const std::string &msg);

}  // namespace bayes
