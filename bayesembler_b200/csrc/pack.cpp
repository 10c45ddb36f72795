// Host packer: one CollapsedMap (CSR) -> the device layout of the Gibbs kernel.
//
//  * rows with a single cell always send all their fragments to that cell's candidate (p/Z == 1 in
//    generateAssignment, assignmentGenerator.cpp:280-334/346-349; in initAssignment whatever the binomial leaves is
//    the row's own remainder), so they are folded into per-candidate base counts and never visited again;
//  * the remaining rows are ordered by (binomial-chain rows first, wider rows first, larger counts first) so that the
//    32 rows a warp walks together take the same branch and have similar lengths;
//  * 32 consecutive rows form a slab stored column-major (cell k of lane l at slab + 32 k + l, SELL-32), so a warp
//    reading "cell k of my row" touches 32 consecutive doubles / 32 consecutive uint16 columns: conflict-free in
//    shared memory, fully coalesced from L2/HBM.
#include <float.h>
#include <math.h>

#include <algorithm>
#include <numeric>

#include "internal.h"
#include "layout.h"

namespace bayes {

int validate_locus(const bayes_locus_desc &in, std::string *err) {
    if (in.T < 1 || in.R < 1) { *err = "locus needs T >= 1 and R >= 1"; return BAYES_E_INVALID; }
    if (!in.row_ptr || !in.col_idx || !in.val || !in.row_count) { *err = "null CSR pointer"; return BAYES_E_INVALID; }
    if (in.row_ptr[0] != 0) { *err = "row_ptr[0] != 0"; return BAYES_E_INVALID; }
    for (int i = 0; i < in.R; i++) {
        if (in.row_ptr[i + 1] < in.row_ptr[i]) { *err = "row_ptr not monotone"; return BAYES_E_INVALID; }
        if (in.row_count[i] < 1) { *err = "row count < 1"; return BAYES_E_INVALID; }
        bool support = false;
        int prev = -1;
        for (int k = in.row_ptr[i]; k < in.row_ptr[i + 1]; k++) {
            const int c = in.col_idx[k];
            if (c <= prev || c >= in.T) { *err = "columns must be strictly ascending and < T"; return BAYES_E_INVALID; }
            prev = c;
            if (!(in.val[k] >= 0.0) || !std::isfinite(in.val[k])) { *err = "probability not finite / negative"; return BAYES_E_INVALID; }
            support |= in.val[k] >= DBL_MIN;
        }
        if (!support) { *err = "row without a cell >= DBL_MIN"; return BAYES_E_INVALID; }
    }
    return BAYES_OK;
}

// Same carve-up as the kernel uses (layout.h).
size_t locus_smem_bytes(const LocusDev &d) {
    return smem_layout(d.T, d.n_slabs, d.n_cells, d.tab_len, d.tab_smem != 0, d.mat_smem != 0).total;
}

int pack_locus(const bayes_locus_desc &in, PackedLocus *out, std::string *err) {
    int rc = validate_locus(in, err);
    if (rc != 0) return rc;
    if (in.T > kMaxT) { *err = "T exceeds the supported maximum (4096 candidates)"; return BAYES_E_INVALID; }
    if (!in.efflen) { *err = "null efflen"; return BAYES_E_INVALID; }
    if (!(in.gamma >= DBL_MIN)) { *err = "gamma must be > 0"; return BAYES_E_INVALID; }  // assembler.cpp:1567
    if (in.burn_in < 0 || in.samples < 1) { *err = "need burn_in >= 0 and samples >= 1"; return BAYES_E_INVALID; }
    if (!(in.total_num_fragments >= 1.0) || in.total_num_fragments > 2147483647.0) { *err = "total_num_fragments out of int range"; return BAYES_E_INVALID; }
    if (in.n_chains < 1) { *err = "n_chains < 1"; return BAYES_E_INVALID; }

    const int T = in.T, R = in.R;
    PackedLocus &P = *out;
    P = PackedLocus();
    P.n_chains = in.n_chains;
    P.seed = in.seed;
    P.efflen.assign(in.efflen, in.efflen + T);
    P.base_counts.assign(T, 0);

    struct Row { int32_t id, n, nnz; };
    std::vector<Row> rows;
    rows.reserve(R);
    int64_t n_frag = 0;
    for (int i = 0; i < R; i++) {
        int nz = 0, last = -1;
        for (int k = in.row_ptr[i]; k < in.row_ptr[i + 1]; k++)
            if (in.val[k] != 0.0) { nz++; last = in.col_idx[k]; }
        n_frag += in.row_count[i];
        if (nz == 1) { P.base_counts[last] += in.row_count[i]; P.folded++; }
        else rows.push_back(Row{i, in.row_count[i], nz});
    }
    if (n_frag > 2147483647) { *err = "more than 2^31-1 fragments in one locus"; return BAYES_E_INVALID; }
    std::stable_sort(rows.begin(), rows.end(), [](const Row &a, const Row &b) {
        const bool ca = a.n >= kSmallCount, cb = b.n >= kSmallCount;
        if (ca != cb) return ca;
        if (a.nnz != b.nnz) return a.nnz > b.nnz;
        return a.n > b.n;
    });

    const int Rm = (int)rows.size();
    const int n_slabs = (Rm + 31) / 32;
    P.slab_cell_off.assign(n_slabs + 1, 0);
    P.row_n.assign((size_t)n_slabs * 32, 0);
    P.row_nnz.assign((size_t)n_slabs * 32, 0);
    P.row_id.assign((size_t)n_slabs * 32, 0);
    for (int s = 0; s < n_slabs; s++) {
        int w = 0;
        for (int l = 0; l < 32 && s * 32 + l < Rm; l++) w = std::max(w, rows[s * 32 + l].nnz);
        P.slab_cell_off[s + 1] = P.slab_cell_off[s] + 32 * w;
    }
    const int n_cells = P.slab_cell_off[n_slabs];
    P.val.assign(n_cells, 0.0);
    P.col.assign(n_cells, 0);
    for (int r = 0; r < Rm; r++) {
        const Row &row = rows[r];
        const int s = r / 32, l = r % 32;
        P.row_n[r] = row.n;
        P.row_nnz[r] = row.nnz;
        P.row_id[r] = row.id;
        int k2 = 0;
        for (int k = in.row_ptr[row.id]; k < in.row_ptr[row.id + 1]; k++) {
            if (in.val[k] == 0.0) continue;
            const size_t at = (size_t)P.slab_cell_off[s] + 32 * (size_t)k2 + l;
            P.val[at] = in.val[k];
            P.col[at] = (uint16_t)in.col_idx[k];
            k2++;
        }
    }

    // pi from the minimum read cover unless supplied (assembler.cpp:1551-1559)
    double pi = in.pi;
    if (pi < 0) {
        const int cover = min_read_cover(in, bayes_chain_key(in.seed, 0), nullptr);
        if (cover < 0) { *err = "minimum read cover failed"; return BAYES_E_INVALID; }
        P.cover = cover;
        pi = (double)cover / T;
        if (pi > 1.0 - DBL_EPSILON) pi = 1.0 - DBL_EPSILON;
    }
    if (!(pi >= DBL_MIN) || !(pi <= 1.0 - DBL_EPSILON)) { *err = "pi out of (0, 1-eps]"; return BAYES_E_INVALID; }
    P.pi = pi;
    if (T > 1) simplex_table(pi, in.gamma, T, (int)n_frag, &P.tab_row, &P.tab_val);
    else { P.tab_row.assign(2, 0); }

    LocusDev &d = P.d;
    d = LocusDev();
    d.T = T;
    d.R = Rm;
    d.n_slabs = n_slabs;
    d.N_f = (int32_t)n_frag;
    d.N_total = (int32_t)in.total_num_fragments;
    d.burn = in.burn_in;
    d.samples = in.samples;
    d.tab_len = (int32_t)P.tab_val.size();
    d.n_cells = n_cells;
    d.gamma = in.gamma;
    d.tab_smem = d.tab_len <= kTableSmemEntries;
    const size_t mat_bytes = (size_t)n_cells * 10 + (size_t)n_slabs * 32 * 12 + 4 * ((size_t)n_slabs + 1);
    d.mat_smem = mat_bytes <= (size_t)kMatrixSmemBytes;
    d.smem_bytes = (int32_t)locus_smem_bytes(d);

    int block = 32;
    while (block < Rm && block < kMaxBlock) block *= 2;
    P.block = block;
    const double iters = (double)in.burn_in + in.samples;
    P.cost = iters * ((double)n_cells + 8.0 * Rm + 16.0 * T + 64.0);
    return BAYES_OK;
}

}  // namespace bayes
