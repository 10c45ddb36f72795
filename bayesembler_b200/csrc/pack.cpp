// Host packer: CollapsedMaps (CSR) -> the device layout of the Gibbs kernels.
//
// prepare_locus (per locus, thread pool):
//  * rows with a single cell always send all their fragments to that cell's candidate (p/Z == 1 in
//    generateAssignment, assignmentGenerator.cpp:280-334/346-349; in initAssignment whatever the binomial leaves is
//    the row's own remainder), so they are folded into per-candidate base counts and never visited again;
//  * minimum read cover -> pi (assembler.cpp:1551-1559) and the simplex-size CDF table (simplexSizeGenerator.cpp:117-187).
// pack_unit (per thread block):
//  * the rows of all slots of the unit are pooled and ordered by (binomial-chain rows first, wider rows first, larger
//    counts first) so that the 32 rows a warp walks together take the same branch and have similar lengths;
//  * h consecutive rows (h = 32, or 16 / 8 / 4 for latency-critical units that spread their rows over more warps) form
//    a slab stored column-major (cell k of lane l at slab + h k + l, SELL-h), so a warp reading "cell k of my row"
//    touches consecutive doubles / consecutive uint16 columns: conflict-free in shared memory, coalesced from L2/HBM;
//  * in a bundle the column of a cell is the LANE of its candidate (slot * Tpad + t), so the assignment step indexes
//    one pooled theta / counts vector whichever locus a row belongs to.
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

size_t max_unit_smem() { return 227 * 1024; }
int tree_chunks(int n) { return (n + 63) / 64 <= kTreeChunks ? (n + 63) / 64 : 1; }
bool g_tile_stream = true;

int slot_lanes(int T) {
    if (T > 32) return 0;
    int p = 2;
    while (p < T) p *= 2;
    return p;
}

int prepare_locus(const bayes_locus_desc &in, PreLocus *out, std::string *err) {
    int rc = validate_locus(in, err);
    if (rc != 0) return rc;
    if (in.T > kMaxT) { *err = "T exceeds the supported maximum (4096 candidates)"; return BAYES_E_INVALID; }
    if (!in.efflen) { *err = "null efflen"; return BAYES_E_INVALID; }
    for (int t = 0; t < in.T; t++)  // SequencingModel::getEffectiveLength is a sum of probabilities (sequencingModel.cpp:47-57): > 0
        if (!(in.efflen[t] > 0.0) || !std::isfinite(in.efflen[t])) { *err = "effective length must be finite and > 0"; return BAYES_E_INVALID; }
    if (!(in.gamma >= DBL_MIN)) { *err = "gamma must be > 0"; return BAYES_E_INVALID; }  // assembler.cpp:1567
    if (in.burn_in < 0 || in.samples < 1) { *err = "need burn_in >= 0 and samples >= 1"; return BAYES_E_INVALID; }
    if (!(in.total_num_fragments >= 1.0) || in.total_num_fragments > 2147483647.0) { *err = "total_num_fragments out of int range"; return BAYES_E_INVALID; }
    if (in.n_chains < 1) { *err = "n_chains < 1"; return BAYES_E_INVALID; }
    if (in.R >= (1 << kRowSlotShift)) { *err = "too many rows"; return BAYES_E_INVALID; }

    const int T = in.T, R = in.R;
    PreLocus &P = *out;
    P = PreLocus();
    P.T = T;
    P.N_total = (int32_t)in.total_num_fragments;
    P.burn = in.burn_in;
    P.samples = in.samples;
    P.n_chains = in.n_chains;
    P.gamma = in.gamma;
    P.seed = in.seed;
    P.efflen.assign(in.efflen, in.efflen + T);
    P.base_counts.assign(T, 0);

    struct Row { int32_t id, n, nnz, woff; };
    std::vector<Row> rows;
    rows.reserve(R);
    int64_t n_frag = 0, pool_words = 0;
    for (int i = 0; i < R; i++) {
        int nz = 0, last = -1;
        for (int k = in.row_ptr[i]; k < in.row_ptr[i + 1]; k++)
            if (in.val[k] != 0.0) { nz++; last = in.col_idx[k]; }
        n_frag += in.row_count[i];
        if (nz == 1) { P.base_counts[last] += in.row_count[i]; P.folded++; continue; }
        // fast mode: a row with fewer than kSmallCount fragments owns that many consecutive words of the locus' pool
        int32_t woff = -1;
        if (in.row_count[i] < kSmallCount) {
            if (pool_words > 2000000000) { *err = "too many fragments in rows with < 20 fragments"; return BAYES_E_INVALID; }
            woff = (int32_t)pool_words;
            pool_words += in.row_count[i];
        }
        rows.push_back(Row{i, in.row_count[i], nz, woff});
    }
    P.pool_words = pool_words;
    if (n_frag > 2147483647) { *err = "more than 2^31-1 fragments in one locus"; return BAYES_E_INVALID; }
    P.n_frag = (int32_t)n_frag;
    std::stable_sort(rows.begin(), rows.end(), [](const Row &a, const Row &b) {
        const bool ca = a.n >= kSmallCount, cb = b.n >= kSmallCount;
        if (ca != cb) return ca;
        if (a.nnz != b.nnz) return a.nnz > b.nnz;
        return a.n > b.n;
    });
    P.row_ptr.assign(1, 0);
    for (const Row &row : rows) {
        P.row_id.push_back(row.id);
        P.row_n.push_back(row.n);
        P.row_woff.push_back(row.woff);
        for (int k = in.row_ptr[row.id]; k < in.row_ptr[row.id + 1]; k++) {
            if (in.val[k] == 0.0) continue;
            P.cell_col.push_back(in.col_idx[k]);
            P.cell_val.push_back(in.val[k]);
        }
        P.row_ptr.push_back((int32_t)P.cell_val.size());
    }

    // pi from the minimum read cover unless supplied (assembler.cpp:1551-1559), one cover per chain
    P.chain_prior.assign(in.n_chains, 0);
    auto add_prior = [&](double pi, int cover, std::string *e) -> int {
        if (!(pi >= DBL_MIN) || !(pi <= 1.0 - DBL_EPSILON)) { *e = "pi out of (0, 1-eps]"; return -1; }
        for (size_t i = 0; i < P.priors.size(); i++)
            if (P.priors[i].pi == pi && P.priors[i].cover == cover) return (int)i;
        P.priors.push_back(SimplexPrior());
        SimplexPrior &sp = P.priors.back();
        sp.pi = pi;
        sp.cover = cover;
        if (T > 1) simplex_table(pi, in.gamma, T, (int)n_frag, &sp.tab_row, &sp.tab_val);
        else sp.tab_row.assign(2, 0);
        return (int)P.priors.size() - 1;
    };
    if (in.pi < 0) {
        for (int c = 0; c < in.n_chains; c++) {
            const int cover = min_read_cover(in, bayes_chain_key(in.seed, c), nullptr);
            if (cover < 0) { *err = "minimum read cover failed"; return BAYES_E_INVALID; }
            double pi = (double)cover / T;
            if (pi > 1.0 - DBL_EPSILON) pi = 1.0 - DBL_EPSILON;
            const int idx = add_prior(pi, cover, err);
            if (idx < 0) return BAYES_E_INVALID;
            P.chain_prior[c] = idx;
        }
    } else {
        if (add_prior(in.pi, 0, err) < 0) return BAYES_E_INVALID;
    }

    const double iters = (double)in.burn_in + in.samples;
    P.cost = iters * ((double)P.cell_val.size() + 8.0 * rows.size() + 16.0 * T + 64.0);
    return BAYES_OK;
}

// A-step items of a unit in processing order: chain items first (wider first, larger counts first: the number of
// conditional binomials a lane runs is bounded by its cells), then uniform items (more Philox blocks first, then wider),
// so that the lanes of a slab loop the same number of times.
void build_items(const std::vector<SlotRef> &slots, std::vector<ItemShape> *items, int *n_chain) {
    items->clear();
    for (int g = 0; g < (int)slots.size(); g++) {
        const PreLocus &L = *slots[g].locus;
        for (size_t r = 0; r < L.row_id.size(); r++) {
            const int32_t n = L.row_n[r], nnz = L.row_ptr[r + 1] - L.row_ptr[r];
            const bool chain = n >= kSmallCount;
            items->push_back(ItemShape{g, (int32_t)r, n, nnz, chain ? 0 : (n + 3) / 4, chain});
        }
    }
    std::stable_sort(items->begin(), items->end(), [](const ItemShape &a, const ItemShape &b) {
        if (a.chain != b.chain) return a.chain;
        if (a.chain) return a.nnz != b.nnz ? a.nnz > b.nnz : a.n > b.n;
        return a.blocks != b.blocks ? a.blocks > b.blocks : a.nnz > b.nnz;
    });
    int nc = 0;
    for (const ItemShape &it : *items) nc += it.chain;
    *n_chain = nc;
}

// Per-slot tail of a packed unit: per-candidate constants, CDF tables, keys.
static void pack_slots(const std::vector<SlotRef> &slots, bool wide, int Tpad, int nT, PackedUnit &U) {
    const int n_slots = (int)slots.size();
    U.efflen.assign(nT, 1.0);
    U.base_counts.assign(nT, 0);
    U.tab_row.assign((size_t)n_slots * (Tpad + 1), 0);
    UnitDev &d = U.d;
    for (int g = 0; g < n_slots; g++) {
        const PreLocus &L = *slots[g].locus;
        const int lane0 = wide ? 0 : g * Tpad;
        for (int t = 0; t < L.T; t++) {
            U.efflen[lane0 + t] = L.efflen[t];
            U.base_counts[lane0 + t] = L.base_counts[t];
        }
        const SimplexPrior &sp = L.prior(slots[g].chain);
        const int32_t tab_base = (int32_t)U.tab_val.size();
        for (int t = 0; t <= L.T; t++) U.tab_row[(size_t)g * (Tpad + 1) + t] = tab_base + sp.tab_row[t];
        U.tab_val.insert(U.tab_val.end(), sp.tab_val.begin(), sp.tab_val.end());
        d.T[g] = L.T;
        d.N_f[g] = L.n_frag;
        d.N_total[g] = L.N_total;
        d.gamma[g] = L.gamma;
        d.key[g] = bayes_chain_key(L.seed, slots[g].chain);
        d.out_off[g] = slots[g].out_off;
        U.cost = std::max(U.cost, L.cost);
    }
    U.cost *= 1.0 + 0.25 * (n_slots - 1);
    d.n_slots = n_slots;
    d.Tpad = Tpad;
    d.wide = wide ? 1 : 0;
    d.burn = slots[0].locus->burn;
    d.samples = slots[0].locus->samples;
    d.tab_len = (int32_t)U.tab_val.size();
}

// Fast mode, bundle: lane-per-cell tiles (gibbs_fast.cu).  The rows of all slots are pooled, chain rows (>= 20 fragments)
// first, then the others with the rows that need most Philox blocks first; every row goes to the first tile of its kind
// that still has room for all its cells (first fit over rows in decreasing order): the rows with several Philox blocks
// fill the first tiles among themselves, the one-block rows that follow fill the gaps, so a tile wastes few lanes and
// few tiles loop over more than one block.
static void pack_unit_tiles(const std::vector<SlotRef> &slots, PackedUnit *out) {
    PackedUnit &U = *out;
    U = PackedUnit();
    const int n_slots = (int)slots.size();
    const int Tpad = slot_lanes(slots[0].locus->T);
    std::vector<ItemShape> items;
    int n_chain = 0;
    build_items(slots, &items, &n_chain);
    const int n_items = (int)items.size();

    std::vector<int64_t> pool_base(n_slots, 0);
    int64_t pool_blocks = 0, qcap = 0;
    for (int g = 0; g < n_slots; g++) {
        pool_base[g] = pool_blocks;
        const int64_t nb = (slots[g].locus->pool_words + 3) / 4;
        for (int64_t b = 0; b < nb; b++) U.pool_desc.push_back(((uint32_t)g << 24) | (uint32_t)b);
        pool_blocks += nb;
    }
    // tiles
    std::vector<int> fill;         // lanes used per tile
    std::vector<int> item_tile(n_items), item_lane(n_items);
    int n_chain_tiles = 0;
    auto place = [&](int first, int last, int tile0) {
        int open0 = tile0;  // tiles before this one are full
        for (int i = first; i < last; i++) {
            const int need = items[i].nnz;
            int t = -1;
            while (open0 < (int)fill.size() && fill[open0] >= 31) open0++;  // a row has at least two cells
            for (int k = open0; k < (int)fill.size(); k++)
                if (fill[k] + need <= 32) { t = k; break; }
            if (t < 0) { fill.push_back(0); t = (int)fill.size() - 1; }
            item_tile[i] = t;
            item_lane[i] = fill[t];
            fill[t] += need;
        }
    };
    place(0, n_chain, 0);
    n_chain_tiles = (int)fill.size();
    place(n_chain, n_items, n_chain_tiles);
    const int n_tiles = (int)fill.size();
    const size_t n_cells = (size_t)n_tiles * 32;
    U.val.assign(n_cells, 0.0);
    // padding lane: its own one-cell segment (head), no fragments, probability 0
    U.cell_meta.assign(n_cells, 1u << kMetaHeadShift);
    for (size_t i = 0; i < (size_t)n_chain_tiles * 32; i++) U.cell_meta[i] |= 0x3FFFFFFu << kMetaOrdShift;  // chain tiles: padding ordinal
    U.chain_tab.assign((size_t)3 * n_chain, 0u);
    for (int i = 0; i < n_items; i++) {
        const ItemShape &it = items[i];
        const PreLocus &L = *slots[it.slot].locus;
        const size_t cell0 = (size_t)item_tile[i] * 32 + item_lane[i];
        const int lane0 = it.slot * Tpad;
        uint32_t hi;
        if (it.chain) {
            hi = (uint32_t)i << kMetaOrdShift;  // chain items come first: i is the chain-row ordinal
            U.chain_tab[3 * i] = (uint32_t)it.n;
            U.chain_tab[3 * i + 1] = ((uint32_t)it.slot << kRowSlotShift) | (uint32_t)L.row_id[it.r];
            U.chain_tab[3 * i + 2] = (uint32_t)cell0;
            qcap += (int64_t)tree_chunks(it.n) * std::max(1, it.nnz / 2);
        } else {
            const int64_t woff = 4 * pool_base[it.slot] + L.row_woff[it.r];
            hi = ((uint32_t)it.n << kMetaNShift) | ((uint32_t)woff << kMetaWoffShift);
        }
        for (int k = 0; k < it.nnz; k++) {
            U.val[cell0 + k] = L.cell_val[L.row_ptr[it.r] + k];
            U.cell_meta[cell0 + k] = (uint32_t)(lane0 + L.cell_col[L.row_ptr[it.r] + k]) | ((k == 0 ? 1u : 0u) << kMetaHeadShift) | hi;
        }
    }
    pack_slots(slots, false, Tpad, 32, U);
    UnitDev &d = U.d;
    for (int g = 0; g < n_slots; g++) d.pool_first[g] = (int32_t)pool_base[g];
    d.n_slabs = 0;
    d.slab_h = 32;
    d.n_cells = (int32_t)n_cells;
    d.n_tiles = n_tiles;
    d.n_chain_tiles = n_chain_tiles;
    d.n_chain_cells = 32 * n_chain_tiles;
    d.n_chain_rows = n_chain;
    d.n_chain_slabs = 0;
    d.qcap = (int32_t)std::max<int64_t>(qcap, 1);
    d.pool_blocks = (int32_t)pool_blocks;
    const size_t limit = max_unit_smem();
    // The tiles are read once per iteration, front to back, coalesced: streamed from global memory (they stay in L2) behind a
    // one-tile prefetch they cost no shared memory, and a unit can pool as many slots as its 32 candidate lanes hold.  What is
    // accessed at random -- scratch lists, chain-row table, task queues, the word pool, the CDF tables -- stays in shared memory.
    auto total = [&](bool tab, bool mat, bool scr, bool pool) -> size_t {
        return tile_layout(n_slots, Tpad, n_tiles, n_chain_tiles, n_chain, d.tab_len, tab, mat, scr, d.qcap, d.pool_blocks, pool).total;
    };
    bool tab = d.tab_len <= kTableSmemEntries, mat = !g_tile_stream && n_cells * 12 <= (size_t)kMatrixSmemBytes, scr = true, pool = pool_blocks <= 8192;
    if (total(tab, mat, scr, pool) > limit && tab && total(false, mat, scr, pool) <= limit) tab = false;
    if (total(tab, mat, scr, pool) > limit) mat = false;
    if (total(tab, mat, scr, pool) > limit) tab = false;
    if (total(tab, mat, scr, pool) > limit) pool = false;
    if (total(tab, mat, scr, pool) > limit) scr = false;
    d.tab_smem = tab;
    d.mat_smem = mat;
    d.scr_smem = scr;
    d.pool_smem = pool;
    d.smem_bytes = (int32_t)total(tab, mat, scr, pool);
    d.block = 32;
}

// pool words a bundle of these slots addresses (fast mode): must stay below kMaxTilePoolWords
int64_t bundle_pool_words(const std::vector<SlotRef> &slots) {
    int64_t w = 0;
    for (const SlotRef &s : slots) w += 4 * ((s.locus->pool_words + 3) / 4);
    return w;
}

void pack_unit(const std::vector<SlotRef> &slots, bool wide, int slab_h, bool fast, PackedUnit *out) {
    if (fast && !wide) {
        pack_unit_tiles(slots, out);
        return;
    }
    const int H = slab_h;
    PackedUnit &U = *out;
    U = PackedUnit();
    const int n_slots = (int)slots.size();
    const PreLocus &L0 = *slots[0].locus;
    const int Tpad = wide ? L0.T : slot_lanes(L0.T);
    const int nT = wide ? L0.T : 32;

    std::vector<ItemShape> items;
    int n_chain = 0;
    build_items(slots, &items, &n_chain);
    // chain items fill their own slabs; the uniform items start on a slab boundary (a slab is one kind of work)
    const int n_items = (int)items.size();
    const int chain_slabs = (n_chain + H - 1) / H, unif_slabs = (n_items - n_chain + H - 1) / H;
    const int n_slabs = chain_slabs + unif_slabs;
    auto position = [&](int i) { return i < n_chain ? i : chain_slabs * H + (i - n_chain); };  // item -> padded row
    U.slab_cell_off.assign(n_slabs + 1, 0);
    U.row_n.assign((size_t)n_slabs * H, 0);
    U.row_nnz.assign((size_t)n_slabs * H, 0);
    U.row_info.assign((size_t)n_slabs * H, 0);
    std::vector<int> width(n_slabs, 0);
    for (int i = 0; i < n_items; i++) width[position(i) / H] = std::max(width[position(i) / H], items[i].nnz);
    for (int s = 0; s < n_slabs; s++) U.slab_cell_off[s + 1] = U.slab_cell_off[s] + H * width[s];
    const int n_cells = U.slab_cell_off[n_slabs];
    U.val.assign(n_cells, 0.0);
    U.col.assign(n_cells, 0);
    // wide units: T bits per padded row saying which candidates the row has a cell for (gibbs_kernel.cu, assign_rows_wide)
    const int n_words = wide ? (L0.T + 63) / 64 : 0;  // 64-bit words, stored as pairs of uint32 (low half first)
    U.row_mask.assign((size_t)n_slabs * H * n_words * 2, 0u);
    // fast mode: the uniform-word pools of the slots back to back, whole Philox blocks per slot
    std::vector<int64_t> pool_base(n_slots, 0);  // first block of the slot
    int64_t pool_blocks = 0, qcap = 0;
    if (fast) {
        U.row_woff.assign((size_t)n_slabs * H, 0);
        for (int g = 0; g < n_slots; g++) {
            pool_base[g] = pool_blocks;
            const int64_t nb = (slots[g].locus->pool_words + 3) / 4;
            for (int64_t b = 0; b < nb; b++) U.pool_desc.push_back(((uint32_t)g << 24) | (uint32_t)b);
            pool_blocks += nb;
        }
    }
    for (int i = 0; i < n_items; i++) {
        const ItemShape &it = items[i];
        const PreLocus &L = *slots[it.slot].locus;
        const int r = position(i), s = r / H, l = r % H;
        U.row_n[r] = it.n;
        if (fast) {
            if (it.chain) qcap += (int64_t)tree_chunks(it.n) * std::max(1, it.nnz / 2);  // nodes of the deepest tree level with two or more cells, per chunk
            else U.row_woff[r] = (int32_t)(4 * pool_base[it.slot] + L.row_woff[it.r]);
        }
        // a product P_ij theta_j can only reach the denormal range from a tiny P_ij, or from a tiny theta_j, which needs a
        // Gamma(< 1) draw: flag those rows for the exact (slow) in-play rule of the A-step
        bool may_be_tiny = L.gamma < 1.0;
        for (int k = 0; k < it.nnz; k++) may_be_tiny |= L.cell_val[L.row_ptr[it.r] + k] < 1e-150;
        U.row_nnz[r] = (int32_t)((uint32_t)it.nnz | ((uint32_t)it.blocks << kItemBlockShift) | ((uint32_t)(it.chain ? 1 : 0) << kItemChainShift) |
                                 ((uint32_t)(may_be_tiny ? 1 : 0) << kItemTinyShift));
        U.row_info[r] = ((uint32_t)it.slot << kRowSlotShift) | (uint32_t)L.row_id[it.r];
        const int lane0 = wide ? 0 : it.slot * Tpad;
        for (int k = 0; k < it.nnz; k++) {
            const size_t at = (size_t)U.slab_cell_off[s] + H * (size_t)k + l;
            U.val[at] = L.cell_val[L.row_ptr[it.r] + k];
            U.col[at] = (uint16_t)(lane0 + L.cell_col[L.row_ptr[it.r] + k]);
            if (wide) {
                const int c = L.cell_col[L.row_ptr[it.r] + k];
                U.row_mask[2 * (((size_t)s * n_words + (c >> 6)) * H + l) + ((c >> 5) & 1)] |= 1u << (c & 31);
            }
        }
    }

    pack_slots(slots, wide, Tpad, nT, U);
    UnitDev &d = U.d;
    for (int g = 0; g < n_slots; g++) d.pool_first[g] = (int32_t)pool_base[g];
    d.n_slabs = n_slabs;
    d.slab_h = H;
    d.n_cells = n_cells;
    d.n_chain_slabs = chain_slabs;
    d.n_chain_cells = U.slab_cell_off[chain_slabs];
    d.qcap = (int32_t)std::max<int64_t>(qcap, 1);
    d.pool_blocks = (int32_t)pool_blocks;
    // What goes to shared memory is decided on the unit's WHOLE footprint against the opt-in limit of the device: first
    // the matrix (with, in fast mode, the scratch lists and task queues that go where the matrix goes), then the CDF
    // tables; whatever does not fit is read from global memory (L2) instead.
    const size_t limit = max_unit_smem();
    auto total = [&](bool tab, bool mat, bool pool) -> size_t {
        return fast ? fast_layout(wide, nT, n_slots, Tpad, n_slabs, H, n_cells, d.tab_len, tab, mat, d.n_chain_cells, d.qcap, d.pool_blocks, pool).total
                    : unit_layout(wide, nT, n_slots, Tpad, n_slabs, H, n_cells, d.tab_len, tab, mat).total;
    };
    const size_t mat_bytes = (size_t)n_cells * 10 + (size_t)n_slabs * H * 12 + 4 * ((size_t)n_slabs + 1);
    bool tab = d.tab_len <= kTableSmemEntries, mat = mat_bytes <= (size_t)kMatrixSmemBytes, pool = fast && pool_blocks <= 8192;
    if (total(tab, mat, pool) > limit && tab && total(false, mat, pool) <= limit) tab = false;
    if (total(tab, mat, pool) > limit) mat = false;
    if (total(tab, mat, pool) > limit) tab = false;
    if (total(tab, mat, pool) > limit) pool = false;
    d.tab_smem = tab;
    d.mat_smem = mat;
    d.scr_smem = mat;
    d.pool_smem = pool;
    d.smem_bytes = (int32_t)total(tab, mat, pool);  // per-candidate state alone is < 200 KB at T = 4096: always fits
    d.block = 32;  // build_units decides
}

// Fast mode, locus with more than 32 candidates: cooperative rows over a thread-block cluster (gibbs_rows.cu).  Row-major
// matrix (values only: the candidate of a cell is the position of its bit in the row's candidate mask), row-major masks.
// The rows are dealt over the n_cta CTAs of the cluster (rows with >= 20 fragments first, wider rows first, round robin:
// every CTA gets the same mix), and stored CTA by CTA so that a CTA's share is one contiguous block it can keep resident in
// its shared memory.  Every CTA has its own pool of uniform words: the Philox blocks its rows with < 20 fragments touch, row
// after row (a block shared by two rows is generated twice, which is cheaper than sharing it across CTAs).
void pack_unit_rows(const std::vector<SlotRef> &slots, int n_cta, PackedUnit *out) {
    PackedUnit &U = *out;
    U = PackedUnit();
    const PreLocus &L = *slots[0].locus;
    const int T = L.T;
    std::vector<ItemShape> items;
    int n_chain = 0;
    build_items(slots, &items, &n_chain);
    const int n_items = (int)items.size();
    n_cta = std::max(1, std::min(n_cta, kMaxCluster));
    std::vector<std::vector<int> > mine(n_cta);
    std::vector<int> chain_of(n_cta, 0);
    for (int i = 0; i < n_chain; i++) { mine[i % n_cta].push_back(i); chain_of[i % n_cta]++; }
    for (int i = n_chain; i < n_items; i++) mine[(i - n_chain) % n_cta].push_back(i);
    const int n_words = (T + 63) / 64;
    UnitDev &d = U.d;
    U.slab_cell_off.assign(1, 0);
    U.row_mask.assign((size_t)n_items * n_words * 2, 0u);
    int rows_max = 0, cells_max = 0, ccells_max = 0, pool_max = 0, chain_max = 0;
    int64_t qcap_max = 1;
    int row = 0;
    for (int j = 0; j < n_cta; j++) {
        d.cta_row0[j] = row;
        d.cta_dtask0[j] = (int32_t)U.chain_tab.size();
        chain_max = std::max(chain_max, chain_of[j]);
        d.cta_chain[j] = chain_of[j];
        d.cta_pool0[j] = (int32_t)U.pool_desc.size();
        const int cell_first = U.slab_cell_off.back();
        int ccells = 0;
        int64_t qcap = 0;
        for (size_t q = 0; q < mine[j].size(); q++, row++) {
            const ItemShape &it = items[mine[j][q]];
            U.row_n.push_back(it.n);
            U.row_nnz.push_back((int32_t)((uint32_t)it.nnz | ((uint32_t)it.blocks << kItemBlockShift) | ((uint32_t)(it.chain ? 1 : 0) << kItemChainShift)));
            U.row_info.push_back((uint32_t)L.row_id[it.r]);
            int32_t woff = 0;
            if (it.chain) {
                if (it.n <= kRowsDirectMax) {
                    for (int b = 0; 4 * b < it.n; b++) U.chain_tab.push_back(((uint32_t)q << 16) | (uint32_t)b);  // q < 2^16: a CTA's rows with >= 20 fragments
                } else {
                    qcap += (int64_t)tree_chunks(it.n) * std::max(1, it.nnz / 2);  // nodes of the deepest tree level with two or more cells, per chunk
                }
                ccells += it.nnz;
            } else {
                const int64_t w0 = L.row_woff[it.r], b0 = w0 / 4, b1 = (w0 + it.n - 1) / 4;
                woff = (int32_t)(4 * ((int64_t)U.pool_desc.size() - d.cta_pool0[j]) + (w0 & 3));
                for (int64_t b = b0; b <= b1; b++) U.pool_desc.push_back((uint32_t)b);
            }
            U.row_woff.push_back(woff);
            for (int k = 0; k < it.nnz; k++) {
                const int c = L.cell_col[L.row_ptr[it.r] + k];
                U.val.push_back(L.cell_val[L.row_ptr[it.r] + k]);
                U.row_mask[2 * ((size_t)row * n_words + (c >> 6)) + ((c >> 5) & 1)] |= 1u << (c & 31);
            }
            U.slab_cell_off.push_back((int32_t)U.val.size());
        }
        rows_max = std::max(rows_max, (int)mine[j].size());
        cells_max = std::max(cells_max, U.slab_cell_off.back() - cell_first);
        ccells_max = std::max(ccells_max, ccells);
        pool_max = std::max(pool_max, (int)U.pool_desc.size() - d.cta_pool0[j]);
        qcap_max = std::max(qcap_max, qcap);
    }
    for (int j = n_cta; j <= kMaxCluster; j++) d.cta_row0[j] = row;
    for (int j = n_cta; j < kMaxCluster; j++) d.cta_chain[j] = 0;
    for (int j = n_cta; j <= kMaxCluster; j++) d.cta_pool0[j] = (int32_t)U.pool_desc.size();
    for (int j = n_cta; j <= kMaxCluster; j++) d.cta_dtask0[j] = (int32_t)U.chain_tab.size();
    pack_slots(slots, true, T, T, U);
    d.pool_first[0] = 0;
    d.n_cta = n_cta;
    d.n_slabs = n_items;
    d.slab_h = 1;
    d.n_cells = (int32_t)U.val.size();
    d.n_chain_slabs = 0;
    d.n_chain_cells = d.n_cells;  // entries of the global scratch lists of a unit that does not keep them in shared memory
    d.qcap = (int32_t)qcap_max;
    d.pool_blocks = (int32_t)U.pool_desc.size();
    d.rows_max = rows_max;
    d.cells_max = cells_max;
    d.ccells_max = ccells_max;
    d.pool_max = pool_max;
    d.chain_max = chain_max;
    // Shared memory, in the order it pays: the scratch lists and task queues of the splitting trees (random access on the
    // critical path), the pool of uniform words, then the matrix itself (streamed reads; L2 serves them otherwise).
    const size_t limit = max_unit_smem();
    auto total = [&](bool mat, bool scr, bool pool) -> size_t {
        return rows_layout(T, n_cta, rows_max, cells_max, ccells_max, chain_max, d.qcap, pool_max, mat, scr, pool).total;
    };
    bool scr = total(false, true, false) <= limit;
    bool pool = total(false, scr, true) <= limit;
    bool mat = total(true, scr, pool) <= limit;
    d.tab_smem = 0;
    d.mat_smem = mat;
    d.scr_smem = scr;
    d.pool_smem = pool;
    d.smem_bytes = (int32_t)total(mat, scr, pool);
    d.block = 32;  // build_units decides
}

}  // namespace bayes
