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

}  // namespace bayes
