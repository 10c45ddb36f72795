// Shared-memory carve-up of one (locus, chain) thread block; used by the host (to size the launch) and by the kernel.
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
    uint32_t theta, den, mean, m2, tab_val, val;        // double
    uint32_t counts, base, occ, tab_row, mask, slab, row_n, row_nnz, row_id;  // int32
    uint32_t col;                                       // uint16
    uint32_t total;
};

BAYES_LAYOUT_HD SmemLayout smem_layout(int T, int n_slabs, int n_cells, int tab_len, bool tab_smem, bool mat_smem) {
    SmemLayout L;
    const uint32_t uT = (uint32_t)T, Rp = (uint32_t)n_slabs * 32u, nch = (uT + 31u) / 32u;
    uint32_t o = 0;
    L.theta = o; o += 8u * uT;
    L.den = o; o += 8u * uT;
    L.mean = o; o += 8u * uT;
    L.m2 = o; o += 8u * uT;
    L.tab_val = o; o += tab_smem ? 8u * (uint32_t)tab_len : 0u;
    L.val = o; o += mat_smem ? 8u * (uint32_t)n_cells : 0u;
    L.counts = o; o += 4u * uT;
    L.base = o; o += 4u * uT;
    L.occ = o; o += 4u * uT;
    L.tab_row = o; o += tab_smem ? 4u * (uT + 1u) : 0u;
    L.mask = o; o += 4u * (2u * nch + 2u);
    L.slab = o; o += mat_smem ? 4u * ((uint32_t)n_slabs + 1u) : 0u;
    L.row_n = o; o += mat_smem ? 4u * Rp : 0u;
    L.row_nnz = o; o += mat_smem ? 4u * Rp : 0u;
    L.row_id = o; o += mat_smem ? 4u * Rp : 0u;
    L.col = o; o += mat_smem ? 2u * (uint32_t)n_cells : 0u;
    L.total = (o + 15u) & ~15u;
    return L;
}

}  // namespace bayes
