// Shared-memory carve-up of one unit's thread block; used by the host (to size the launch) and by the kernels.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define BAYES_LAYOUT_HD __host__ __device__ __forceinline__
#else
#define BAYES_LAYOUT_HD inline
#endif

namespace bayes {

struct SmemLayout {
    // byte offsets from the start of dynamic shared memory; 8-byte members first
    uint32_t theta, den, mean, m2, tab_val, val, key, amask;               // double / uint64
    uint32_t counts, base, occ, tab_row, mask, slab, row_n, row_nnz, row_info, next_slab;  // int32
    uint32_t col, act;                                                     // uint16
    uint32_t total;
};

// nT: per-candidate entries (32 for a bundle, T for a wide unit); wide units also keep den / mean / m2 / base / occ and
// the S+ / picked masks in shared memory, a bundle keeps them in the registers of warp 0.
BAYES_LAYOUT_HD SmemLayout unit_layout(bool wide, int nT, int n_slots, int Tpad, int n_slabs, int slab_h, int n_cells, int tab_len,
                                       bool tab_smem, bool mat_smem) {
    SmemLayout L;
    const uint32_t uT = (uint32_t)nT, Rp = (uint32_t)n_slabs * (uint32_t)slab_h, nch = (uT + 31u) / 32u;
    const uint32_t n_tabrow = (uint32_t)n_slots * ((uint32_t)Tpad + 1u);
    uint32_t o = 0;
    L.theta = o; o += 8u * uT;
    L.den = o; o += wide ? 8u * uT : 0u;
    L.mean = o; o += wide ? 8u * uT : 0u;
    L.m2 = o; o += wide ? 8u * uT : 0u;
    L.tab_val = o; o += tab_smem ? 8u * (uint32_t)tab_len : 0u;
    L.val = o; o += mat_smem ? 8u * (uint32_t)n_cells : 0u;
    L.key = o; o += 8u * (uint32_t)n_slots;
    L.amask = o; o += wide ? 8u * ((uT + 63u) / 64u) : 0u;
    L.next_slab = o; o += 8u;
    L.counts = o; o += 4u * uT;
    L.base = o; o += wide ? 4u * uT : 0u;
    L.occ = o; o += wide ? 4u * uT : 0u;
    L.tab_row = o; o += tab_smem ? 4u * n_tabrow : 0u;
    L.mask = o; o += wide ? 4u * (2u * nch + 2u) : 0u;
    L.slab = o; o += mat_smem ? 4u * ((uint32_t)n_slabs + 1u) : 0u;
    L.row_n = o; o += mat_smem ? 4u * Rp : 0u;
    L.row_nnz = o; o += mat_smem ? 4u * Rp : 0u;
    L.row_info = o; o += mat_smem ? 4u * Rp : 0u;
    L.col = o; o += mat_smem ? 2u * (uint32_t)n_cells : 0u;
    o = (o + 1u) & ~1u;
    L.act = o; o += wide ? 2u * uT : 0u;
    L.total = (o + 15u) & ~15u;
    return L;
}

// ---- fast (production) sampling mode, gibbs_fast.cu ---------------------------------------------------------------
// Every unit keeps the per-candidate state (nT entries: 32 lanes of a bundle, T of a wide unit) in shared memory, the
// E-step loads what it needs per iteration.  Units whose matrix is staged in shared memory (mat_smem) also keep there:
// the per-iteration scratch lists of the chain rows (running sums + candidates of the cells in play), the two task
// queue buffers of the splitting trees, the rows' pool offsets.  The uniform-word pool is a per-iteration cache of Philox
// blocks in shared memory; a unit whose pool does not fit computes the blocks where it needs them (same words).
struct FastLayout {
    uint32_t theta, den, mean, m2, tab_val, val, scr_cum, key, sgamma, amask;  // 8-byte
    uint32_t counts, base, occ, tab_row, mask, slab, row_n, row_nnz, row_info, row_woff, pool, pool_desc, queue, sT, sNf, ctl;  // 4-byte
    uint32_t col, scr_col, act;  // 2-byte
    uint32_t total;
};

constexpr int kFastCtlWords = 8;  // ctl: next slab, three rotating task counters, error flag
constexpr int kTileClusterBytes = 256;  // a tiled unit spread over a cluster: two buffers of 32 partial counts, after TileLayout::total

BAYES_LAYOUT_HD FastLayout fast_layout(bool wide, int nT, int n_slots, int Tpad, int n_slabs, int slab_h, int n_cells, int tab_len,
                                       bool tab_smem, bool mat_smem, int n_chain_cells, int qcap, int pool_blocks, bool pool_smem) {
    FastLayout L;
    const uint32_t uT = (uint32_t)nT, Rp = (uint32_t)n_slabs * (uint32_t)slab_h, nch = (uT + 31u) / 32u;
    const uint32_t n_tabrow = (uint32_t)n_slots * ((uint32_t)Tpad + 1u);
    const uint32_t cells = mat_smem ? (uint32_t)n_cells : 0u, ccells = mat_smem ? (uint32_t)n_chain_cells : 0u;
    const uint32_t rows = mat_smem ? Rp : 0u, qwords = mat_smem ? 6u * (uint32_t)qcap : 0u;
    uint32_t o = 0;
    L.theta = o; o += 8u * uT;
    L.den = o; o += 8u * uT;
    L.mean = o; o += 8u * uT;
    L.m2 = o; o += 8u * uT;
    L.tab_val = o; o += tab_smem ? 8u * (uint32_t)tab_len : 0u;
    L.val = o; o += 8u * cells;
    L.scr_cum = o; o += 8u * ccells;
    L.key = o; o += 8u * (uint32_t)n_slots;
    L.sgamma = o; o += 8u * (uint32_t)n_slots;
    L.amask = o; o += wide ? 8u * ((uT + 63u) / 64u) : 0u;
    o = (o + 15u) & ~15u;
    L.pool = o; o += pool_smem ? 16u * ((uint32_t)pool_blocks + 1u) : 0u;  // one spare block: a row reads its words four at a time
    L.counts = o; o += 4u * uT;
    L.base = o; o += 4u * uT;
    L.occ = o; o += 4u * uT;
    L.tab_row = o; o += tab_smem ? 4u * n_tabrow : 0u;
    L.mask = o; o += wide ? 4u * (2u * nch + 2u) : 0u;
    L.slab = o; o += mat_smem ? 4u * ((uint32_t)n_slabs + 1u) : 0u;
    L.row_n = o; o += 4u * rows;
    L.row_nnz = o; o += 4u * rows;
    L.row_info = o; o += 4u * rows;
    L.row_woff = o; o += 4u * rows;
    L.pool_desc = o; o += pool_smem ? 4u * (uint32_t)pool_blocks : 0u;
    L.queue = o; o += 4u * qwords;
    L.sT = o; o += 4u * (uint32_t)n_slots;
    L.sNf = o; o += 4u * (uint32_t)n_slots;
    L.ctl = o; o += 4u * (uint32_t)kFastCtlWords;
    L.col = o; o += 2u * cells;
    L.scr_col = o; o += 2u * ccells;
    o = (o + 1u) & ~1u;
    L.act = o; o += wide ? 2u * uT : 0u;
    L.total = (o + 15u) & ~15u;
    return L;
}

// ---- fast mode, bundles: lane-per-cell tiles --------------------------------------------------------------------------
struct TileLayout {
    uint32_t theta, den, mean, m2, tab_val, val, scr_cum, key, sgamma, pool;  // 8-byte (pool: 16)
    uint32_t counts, base, occ, tab_row, meta, ctab, pool_desc, queue, sT, sNf, ctl;  // 4-byte
    uint32_t scr_col;  // 2-byte
    uint32_t total;
};

// mat_smem: the tiles (val, meta) are staged in shared memory, else streamed from global memory (L2) with a one-tile
// prefetch; scr_smem: the chain rows' scratch lists, their table and the task queues are in shared memory (else global).
BAYES_LAYOUT_HD TileLayout tile_layout(int n_slots, int Tpad, int n_tiles, int n_chain_tiles, int n_chain_rows, int tab_len, bool tab_smem,
                                       bool mat_smem, bool scr_smem, int qcap, int pool_blocks, bool pool_smem) {
    TileLayout L;
    const uint32_t n_tabrow = (uint32_t)n_slots * ((uint32_t)Tpad + 1u);
    const uint32_t cells = mat_smem ? 32u * (uint32_t)n_tiles : 0u, ccells = scr_smem ? 32u * (uint32_t)n_chain_tiles : 0u;
    uint32_t o = 0;
    L.theta = o; o += 8u * 32u;
    L.den = o; o += 8u * 32u;
    L.mean = o; o += 8u * 32u;
    L.m2 = o; o += 8u * 32u;
    L.tab_val = o; o += tab_smem ? 8u * (uint32_t)tab_len : 0u;
    L.val = o; o += 8u * cells;
    L.scr_cum = o; o += 8u * ccells;
    L.key = o; o += 8u * (uint32_t)n_slots;
    L.sgamma = o; o += 8u * (uint32_t)n_slots;
    o = (o + 15u) & ~15u;
    L.pool = o; o += pool_smem ? 16u * ((uint32_t)pool_blocks + 1u) : 0u;
    L.counts = o; o += 4u * 32u;
    L.base = o; o += 4u * 32u;
    L.occ = o; o += 4u * 32u;
    L.tab_row = o; o += tab_smem ? 4u * n_tabrow : 0u;
    L.meta = o; o += 4u * cells;
    L.ctab = o; o += scr_smem ? 12u * (uint32_t)n_chain_rows : 0u;
    L.pool_desc = o; o += pool_smem ? 4u * (uint32_t)pool_blocks : 0u;
    L.queue = o; o += scr_smem ? 16u * (uint32_t)qcap : 0u;  // two buffers of two words per task
    L.sT = o; o += 4u * (uint32_t)n_slots;
    L.sNf = o; o += 4u * (uint32_t)n_slots;
    L.ctl = o; o += 4u * (uint32_t)kFastCtlWords;
    L.scr_col = o; o += 2u * ccells;
    L.total = (o + 15u) & ~15u;
    return L;
}

// ---- fast mode, loci with more than 32 candidates: cooperative rows over a thread-block cluster (gibbs_rows.cu) ----------
// Every CTA of the cluster keeps the per-candidate state (the E-step is replicated), its own share of the rows (matrix
// values, candidate masks, row table: `mat`), the scratch lists / task queues of its rows with >= 20 fragments (`scr`),
// its pool of uniform words (`pool`) and, in a cluster of more than one CTA, two buffers of partial counts that the
// other CTAs read through distributed shared memory.
struct RowsLayout {
    uint32_t theta, den, mean, m2, val, rmask, scr_cum, amask, norm, bar, ch_rz;   // 8-byte
    uint32_t pool;                                                           // 16-byte
    uint32_t counts, base, occ, part, mask, actoff, ch_kk, rcell, row_n, row_info, row_woff, pool_desc, queue, ctl;  // 4-byte
    uint32_t scr_col, act;                                                   // 2-byte
    uint32_t total;
};

constexpr int kRowsCtlWords = 16;

BAYES_LAYOUT_HD RowsLayout rows_layout(int T, int n_cta, int rows_max, int cells_max, int ccells_max, int chain_max, int qcap, int pool_blocks,
                                       bool mat, bool scr, bool pool) {
    RowsLayout L;
    const uint32_t uT = (uint32_t)T, nch = (uT + 31u) / 32u, nw = (uT + 63u) / 64u;
    const uint32_t rows = mat ? (uint32_t)rows_max : 0u, cells = mat ? (uint32_t)cells_max : 0u, cc = scr ? (uint32_t)ccells_max : 0u;
    uint32_t o = 0;
    L.theta = o; o += 8u * uT;
    L.den = o; o += 8u * uT;
    L.mean = o; o += 8u * uT;
    L.m2 = o; o += 8u * uT;
    L.val = o; o += mat ? 8u * (cells + 1u) : 0u;  // (one spare entry: the copy lands 16-byte congruent with its source)
    L.rmask = o; o += 8u * rows * nw;
    L.scr_cum = o; o += 8u * cc;
    L.amask = o; o += 8u * nw;
    L.norm = o; o += 8u;
    L.bar = o; o += 8u;  // mbarrier of the bulk copy that stages the matrix
    L.ch_rz = o; o += 8u * (uint32_t)chain_max;  // rows with >= 20 fragments: 1 / row sum and cells in play of this iteration
    o = (o + 15u) & ~15u;
    L.pool = o; o += pool ? 16u * ((uint32_t)pool_blocks + 1u) : 0u;  // one spare block: a row reads its words four at a time
    L.counts = o; o += 4u * uT;
    L.base = o; o += 4u * uT;
    L.occ = o; o += 4u * uT;
    L.part = o; o += n_cta > 1 ? 8u * uT : 0u;
    L.mask = o; o += 4u * (2u * nch + 2u);
    L.actoff = o; o += 4u * (nch + 2u);
    L.ch_kk = o; o += 4u * (uint32_t)chain_max;
    L.rcell = o; o += mat ? 4u * (rows + 1u) : 0u;
    L.row_n = o; o += 4u * rows;
    L.row_info = o; o += 4u * rows;
    L.row_woff = o; o += 4u * rows;
    L.pool_desc = o;  // (read from global memory: once per iteration, behind the null-candidate selection)
    L.queue = o; o += scr ? 24u * (uint32_t)qcap : 0u;  // two buffers of three words per task
    L.ctl = o; o += 4u * (uint32_t)kRowsCtlWords;
    L.scr_col = o; o += 2u * cc;
    o = (o + 1u) & ~1u;
    L.act = o; o += 2u * uT;
    L.total = (o + 15u) & ~15u;
    return L;
}

}  // namespace bayes
