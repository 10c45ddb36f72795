// Internal structures shared by the host packer, the C ABI and the CUDA kernels.
//
// Terminology
//   slot  one chain of one locus: the unit of the reference's runSampler call (gibbsSampler.cpp:55-148)
//   unit  what ONE thread block runs: either one slot with any number of candidates ("wide" unit), or up to 16 slots
//         whose candidates fit side by side in the 32 lanes of a warp ("bundle"): slot g owns lanes
//         [g * Tpad, (g + 1) * Tpad), Tpad = power of two >= T.  Rows of all slots of a unit are pooled and sorted
//         together so that the 32 rows a warp walks are alike, whichever locus they came from.
#pragma once

#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/bayesembler_b200.h"
#include "layout.h"

namespace bayes {

constexpr int kMaxT = 4096;             // candidates per locus (per-candidate state lives in shared memory)
constexpr int kMaxBlock = 512;          // threads per unit
constexpr int kWarpsPerSM = 28;         // warps an SM keeps resident in a full batch (replay kernels: 72 registers per thread; fast mode: shared memory of ~17 tiled units): the scheduler's capacity
constexpr int kMaxSlots = 16;           // slots per bundle (T = 2 -> Tpad = 2 -> 16 slots)
#ifndef BAYES_TREE_CHUNKS
#define BAYES_TREE_CHUNKS 128
#endif
constexpr int kTreeChunks = BAYES_TREE_CHUNKS;        // fast mode: a row enters its splitting tree in ceil(n / 64) independent chunks of fragments while that is at most kTreeChunks, else whole
constexpr int kRowsDirectMax = 4096;    // fast mode, T > 32: rows with 20 .. this many fragments are drawn fragment by fragment, larger ones by the splitting tree
constexpr int kMaxCluster = 16;         // CTAs of the thread-block cluster a wide locus may be spread over (fast mode)
constexpr int kSmallCount = 20;         // assignmentGenerator.cpp:289: rows with fewer fragments draw uniforms
constexpr int kTableSmemEntries = 2048; // simplex CDF tables up to this many entries are staged in shared memory
constexpr int kMatrixSmemBytes = 96 * 1024;  // units whose packed matrix exceeds this stream it from L2/HBM
constexpr uint32_t kRowSlotShift = 27;  // row_info = slot << 27 | original row index
// A-step work items (one lane each) = collapsed rows.  A row with n >= kSmallCount fragments is a chain item (one
// conditional binomial per cell in play); a row with fewer is a uniform item that locates its n uniforms in the row's
// CDF, four (one Philox block) at a time.  row_nnz word of an item: cells | Philox blocks << 16 | chain << 24 |
// may-have-denormal-range-cells << 25
constexpr uint32_t kItemNnzMask = 0xFFFFu;
constexpr uint32_t kItemBlockShift = 16, kItemChainShift = 24, kItemTinyShift = 25;
// fast mode, bundle tiles: cell meta word
constexpr uint32_t kMetaHeadShift = 5, kMetaNShift = 6, kMetaWoffShift = 11, kMetaOrdShift = 6;
constexpr int64_t kMaxTilePoolWords = (int64_t)1 << 21;  // pool words a tiled unit can address

// Device-side descriptor of one unit.  All *_off fields index the context-wide concatenated arrays.
struct UnitDev {
    int32_t n_slots;    // 1 for a wide unit
    int32_t Tpad;       // lanes per slot (bundle) or T (wide unit)
    int32_t wide;       // 1: wide unit (gibbs_wide_kernel), 0: bundle (gibbs_bundle_kernel)
    int32_t n_slabs;    // ceil(pooled rows / slab_h) slabs of slab_h rows, column-major inside a slab (SELL-h)
    int32_t slab_h;     // rows per slab = lanes a warp uses in the A-step: 32, or 16 / 8 / 4 for latency-critical units
    int32_t n_cells;    // padded cells = slab_h * sum of slab widths
    int32_t burn;
    int32_t samples;
    int32_t tab_len;    // entries of all simplex CDF tables of the unit
    int32_t mat_smem;   // 1: matrix staged in shared memory
    int32_t tab_smem;   // 1: CDF tables staged in shared memory
    int32_t smem_bytes; // dynamic shared memory this unit needs
    int32_t block;      // threads
    int64_t slab_off;   // -> slab_cell_off[n_slabs + 1] (cell offsets relative to cell_off)
    int64_t cell_off;   // -> val / col
    int64_t row_off;    // -> row_n / row_nnz / row_info (n_slabs * slab_h entries, padded rows have n = nnz = 0)
    int64_t lane_off;   // -> efflen / base_counts: 32 entries (bundle, by lane) or T entries (wide)
    int64_t tab_off;    // -> tab_val (tables of the slots back to back)
    int64_t tabrow_off; // -> tab_row: slot g at g * (Tpad + 1), T + 1 offsets relative to the unit's tab_off
    int64_t mask_off;   // -> row_mask (wide units), in 64-bit words: ceil(T / 64) words per padded row, [slab][word][lane]
    // fast (production) sampling mode, gibbs_fast.cu
    int32_t n_chain_slabs;  // the first slabs hold the rows with >= 20 fragments (tree of binomials), the others the pooled-uniform rows
    int32_t n_chain_cells;  // padded cells of the chain slabs = entries of the per-iteration scratch (cum, col) lists
    int32_t qcap;           // capacity of one task queue buffer (nodes of one tree level over all chain rows)
    int32_t pool_blocks;    // Philox blocks of the unit's uniform-word pool (4 words each)
    int32_t pool_smem;      // 1: the pool is generated into shared memory once per iteration; 0: rows compute their blocks themselves
    int32_t scr_smem;       // tiled units: 1: scratch lists, chain-row table and task queues in shared memory
    int64_t woff_off;       // -> row_woff (n_slabs * slab_h entries): first pool word of a uniform row
    int64_t pooldesc_off;   // -> pool_desc (pool_blocks entries): slot << 24 | block of the slot's pool
    int64_t scr_off;        // -> scr_cum / scr_col (units whose matrix is not in shared memory keep them in global memory)
    int64_t q_off;          // -> queue words, 6 * qcap per unit (idem)
    // fast mode, bundles: lane-per-cell TILES instead of SELL slabs (gibbs_fast.cu).  A tile is 32 consecutive cells; the cells
    // of a row are contiguous and never straddle a tile; the first n_chain_tiles tiles hold the rows with >= 20 fragments.
    // n_cells = 32 * n_tiles, n_chain_cells = 32 * n_chain_tiles for such a unit; val is indexed like meta.
    int32_t n_tiles, n_chain_tiles, n_chain_rows, tile_pad;
    int64_t meta_off;       // -> cell_meta (n_cells): candidate lane | head << 5 | (uniform rows: fragments << 6 | pool word << 11;
                            //    chain rows: chain-row ordinal << 6)
    int64_t ctab_off;       // -> chain_tab (3 words per chain row: fragments, slot << 27 | row, first cell of the row)
    // fast mode, loci with more than 32 candidates: cooperative rows over a thread-block cluster (gibbs_rows.cu).  n_cta > 0
    // marks such a unit: n_slabs = rows, slab_h = 1 (slab_cell_off = first cell of every row), masks row-major; CTA j of the
    // cluster owns rows [cta_row0[j], cta_row0[j + 1]), the first cta_chain[j] of them with >= 20 fragments, and the Philox
    // blocks pool_desc[cta_pool0[j] .. cta_pool0[j + 1]) (row_woff = word within that pool).
    int32_t n_cta, rows_max, cells_max, ccells_max, pool_max, chain_max;
    int32_t cta_row0[kMaxCluster + 1], cta_chain[kMaxCluster], cta_pool0[kMaxCluster + 1];
    int32_t cta_dtask0[kMaxCluster + 1];  // -> chain_tab (from ctab_off): the CTA's direct-draw tasks, (row - first row of the CTA) << 16 | block of 4 fragments
    int64_t gpool_off;      // -> gpool (16-byte blocks), n_cta * (pool_max + 1) per unit whose pool is not in shared memory
    // per slot
    int32_t T[kMaxSlots];
    int32_t N_f[kMaxSlots];      // fragments of the locus (folded rows included)
    int32_t N_total[kMaxSlots];  // (int) total_num_fragments
    double gamma[kMaxSlots];
    uint64_t key[kMaxSlots];     // Philox key of the chain
    int64_t out_off[kMaxSlots];  // -> out_occ / out_mean / out_m2 / out_counts (T entries)
    int32_t pool_first[kMaxSlots];  // fast mode: first block of the slot in the unit's uniform-word pool
};

struct KernelArgs {
    const UnitDev *units;
    const double *val;
    const uint16_t *col;
    const int32_t *slab_cell_off;
    const int32_t *row_n;
    const int32_t *row_nnz;
    const uint32_t *row_info;
    const uint32_t *row_mask;
    const double *efflen;
    const int32_t *base_counts;
    const double *tab_val;
    const int32_t *tab_row;
    int32_t *out_occ;
    double *out_mean;
    double *out_m2;
    int32_t *out_counts;
    uint64_t *unit_clock;  // per unit: globaltimer ns at block start, at block end, SM id (3 words)
    // fast mode
    const int32_t *row_woff;
    const uint32_t *cell_meta;
    const uint32_t *chain_tab;
    const uint32_t *pool_desc;
    double *scr_cum;
    uint16_t *scr_col;
    uint32_t *queue;
    uint32_t *gpool;    // uniform-word pools of the cluster units that do not keep them in shared memory
    int32_t *unit_err;  // per unit: != 0 when a row had no candidate with a non-zero probability (reference: assert, assignmentGenerator.cpp:284)
    // calibration pass (capi.cu, calibrate_units): run only the first iter_cap iterations of every chain and put the
    // time stamp of iteration iter_cap / 2 in the "start" word, so that end - start times the second half; 0 = off
    int32_t iter_cap;
    // replay trace of one slot (trace_unit < 0: off)
    int64_t trace_unit;
    int32_t trace_slot;
    int32_t trace_iters;
    double *tr_theta;
    int32_t *tr_counts;
    int32_t *tr_b;
    int32_t *tr_init;
};

// Host-side result of preparing one locus (everything that does not depend on which unit it lands in).
// Simplex-size prior of a chain: pi and the CDF table built from it.  Every chain is one independent run of the
// reference's call site, so (unless the caller fixes pi) chain c derives pi from ITS OWN minimum read cover, whose
// random tie breaks are keyed by the chain (assembler.cpp:1551-1559); chains with equal cover size share a table.
struct SimplexPrior {
    double pi = 0;
    int32_t cover = 0;
    std::vector<double> tab_val;
    std::vector<int32_t> tab_row;
};

struct PreLocus {
    int32_t T = 0, n_frag = 0, N_total = 0, burn = 0, samples = 0, n_chains = 1, folded = 0;
    double gamma = 1;
    uint64_t seed = 0;
    std::vector<double> efflen;
    std::vector<int32_t> base_counts;
    std::vector<SimplexPrior> priors;   // distinct priors of the locus
    std::vector<int32_t> chain_prior;   // chain -> index into priors
    const SimplexPrior &prior(int chain) const { return priors[chain_prior[chain]]; }
    // multi-cell rows in processing order (binomial-chain rows first, wider first, larger counts first)
    std::vector<int32_t> row_id, row_n, row_ptr;  // row_ptr: rows + 1
    // fast mode: first word of the row in the locus' uniform-word pool (rows with < 20 fragments, in ORIGINAL row order;
    // -1 for the others) and the size of that pool
    std::vector<int32_t> row_woff;
    int64_t pool_words = 0;
    std::vector<int32_t> cell_col;
    std::vector<double> cell_val;
    double cost = 0;  // per chain
};

struct SlotRef {
    const PreLocus *locus;
    int32_t chain;
    int64_t out_off;
};

// Host-side packed unit.
struct PackedUnit {
    UnitDev d;
    std::vector<double> val, efflen, tab_val;
    std::vector<uint16_t> col;
    std::vector<int32_t> slab_cell_off, row_n, row_nnz, base_counts, tab_row;
    std::vector<uint32_t> row_info, row_mask, pool_desc, cell_meta, chain_tab;
    std::vector<int32_t> row_woff;
    double cost = 0;
    double features[4] = {0, 0, 0, 0};  // the time model's inputs (capi.cu)
};

// shape of one A-step item before packing (pack.cpp builds them, capi.cu sizes units with them)
struct ItemShape {
    int32_t slot, r, n, nnz, blocks;
    bool chain;
};

// pack.cpp
void build_items(const std::vector<SlotRef> &slots, std::vector<ItemShape> *items, int *n_chain);
int validate_locus(const bayes_locus_desc &in, std::string *err);
int prepare_locus(const bayes_locus_desc &in, PreLocus *out, std::string *err);
int slot_lanes(int T);  // Tpad of a bundle slot (power of two >= T), 0 when T > 32
void pack_unit(const std::vector<SlotRef> &slots, bool wide, int slab_h, bool fast, PackedUnit *out);
void pack_unit_rows(const std::vector<SlotRef> &slots, int n_cta, PackedUnit *out);  // fast mode, T > 32 (gibbs_rows.cu)
extern bool g_tile_stream;  // fast mode: bundles stream their tiles from global memory (L2) instead of staging them in shared memory
int tree_chunks(int n);   // fast mode: independent chunks a row with n >= 20 fragments enters its splitting tree in (<= 64 fragments each while that takes at most 32 chunks, else the row enters whole)
size_t max_unit_smem();  // dynamic shared memory a unit may use (opt-in limit of the device, 227 KB on sm_100)

// host_steps.cpp
int min_read_cover(const bayes_locus_desc &in, uint64_t key, int32_t *in_cover);
void simplex_table(double pi, double gamma, int T, int num_fragments, std::vector<int32_t> *row_off, std::vector<double> *table);

// gibbs_kernel.cu
int launch_units(const KernelArgs &args, int64_t unit_base, int64_t n_units, int block, bool wide, bool mat_smem, size_t smem_bytes,
                 void *stream);
// gibbs_fast.cu
int launch_units_fast(const KernelArgs &args, int64_t unit_base, int64_t n_units, int block, bool wide, bool mat_smem, bool scr_smem,
                      size_t smem_bytes, void *stream, int n_cta = 1);
int launch_step_assignment_fast(const KernelArgs &args, bool init, bool tiles, const double *d_theta, uint32_t iteration, int32_t *d_counts_out,
                                size_t smem_bytes, void *stream);
int configure_fast_kernels();
// gibbs_rows.cu
int launch_units_rows(const KernelArgs &args, int64_t unit_base, int64_t n_units, int block, int n_cta, size_t smem_bytes, void *stream);
int launch_step_assignment_rows(const KernelArgs &args, bool init, const double *d_theta, uint32_t iteration, int32_t *d_counts_out,
                                size_t smem_bytes, void *stream);
int configure_rows_kernels();
int rows_max_cluster(int block, size_t smem_bytes);  // largest power-of-two cluster the device co-schedules for the rows kernel at this footprint (1 when clusters are unavailable)
int launch_stagger(unsigned ns, void *stream);
int launch_step_assignment(const KernelArgs &args, bool init, const double *d_theta, uint32_t iteration, int32_t *d_counts_out,
                           size_t smem_bytes, void *stream);
int launch_step_expression(int T, const int32_t *d_counts, int b, double gamma, uint64_t key, uint32_t iteration, double *d_theta,
                           void *stream);
int launch_step_variates(int kind, double a, double p, uint64_t key, int n, double *d_out, void *stream);
int configure_kernels();

void set_error(const std::string &msg);

}  // namespace bayes
