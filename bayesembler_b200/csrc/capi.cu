// C ABI of the B200 Gibbs sampler (include/bayesembler_b200.h): context, host preparation on a thread pool, unit
// formation, device buffers, launches.  Everything that computes on the device is in gibbs_kernel.cu; nothing here
// falls back to the CPU.
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <map>
#include <numeric>
#include <queue>
#include <mutex>
#include <thread>
#include <tuple>

#include "internal.h"

namespace bayes {

static thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }

namespace {

int cuda_rc(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return BAYES_OK;
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return BAYES_E_CUDA;
}
#define CU(call)                                   \
    do {                                           \
        int _rc = cuda_rc((call), #call);          \
        if (_rc != BAYES_OK) return _rc;           \
    } while (0)

// growable device buffer
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return BAYES_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const size_t want = std::max<size_t>(n + n / 8, 256);
        CU(cudaMalloc((void **)&p, want * sizeof(T)));
        cap = want;
        return BAYES_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// growable pinned host buffer
template <typename T>
struct PinBuf {
    T *p = nullptr;
    size_t cap = 0, n = 0;
    int resize(size_t want) {
        if (want > cap) {
            T *q = nullptr;
            const size_t c = std::max<size_t>(want + want / 4, 256);
            CU(cudaMallocHost((void **)&q, c * sizeof(T)));
            if (p) {
                memcpy(q, p, n * sizeof(T));
                cudaFreeHost(p);
            }
            p = q;
            cap = c;
        }
        n = want;
        return BAYES_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = n = 0;
    }
};

struct Launch {
    int64_t unit_base, n_units;
    int block;
    bool wide, mat_smem, scr_smem;
    size_t smem;
    int n_cta;  // > 0: cooperative-rows units (gibbs_rows.cu), CTAs per cluster
};

}  // namespace
}  // namespace bayes

using namespace bayes;

// shape of a unit before it is packed: its A-step items in processing order (build_units)
struct UnitShape {
    std::vector<int32_t> nnz;     // per item
    std::vector<int32_t> blocks;  // per item: Philox blocks of a uniform item
    int n_chain = 0;
    double iters = 0;
    bool wide = false;
};

struct bayes_gibbs_ctx {
    int device = 0, sm_count = 148;
    std::vector<UnitShape> shape_cache;    // build_units: kept from the first build of an upload for the calibrated rebuild
    std::vector<PackedUnit> packed_cache;  // idem, in unit formation order
    std::vector<int> packed_h;             // slab height each cached unit was packed with
    cudaStream_t own_stream = nullptr, stream = nullptr;
    int host_threads = 0;
    bool uploaded = false, ran = false, downloaded = false;

    std::mutex submit_lock;         // bayes_gibbs_submit may be called from several host threads (one worker per graph in the reference)
    std::vector<PreLocus> loci;     // per submitted locus
    std::vector<int64_t> out_base;  // locus -> offset of chain 0 in the outputs (chains of a locus are contiguous)
    int64_t out_len = 0;            // total T over all (locus, chain)

    // staging (pinned), concatenated over units
    PinBuf<UnitDev> h_units;
    PinBuf<double> h_val, h_efflen, h_tab_val;
    PinBuf<uint16_t> h_col;
    PinBuf<int32_t> h_slab, h_row_n, h_row_nnz, h_base, h_tab_row;
    PinBuf<uint32_t> h_row_info, h_row_mask, h_pool_desc, h_cell_meta, h_chain_tab;
    PinBuf<int32_t> h_row_woff;
    // results (pinned)
    PinBuf<int32_t> h_occ, h_counts, h_unit_err;
    PinBuf<double> h_mean, h_m2;

    DevBuf<UnitDev> d_units;
    DevBuf<double> d_val, d_efflen, d_tab_val, d_mean, d_m2, d_tr_theta;
    DevBuf<uint16_t> d_col;
    DevBuf<int32_t> d_slab, d_row_n, d_row_nnz, d_base, d_tab_row, d_occ, d_counts, d_tr_counts, d_tr_b, d_tr_init;
    DevBuf<uint32_t> d_row_info, d_row_mask, d_pool_desc, d_queue, d_cell_meta, d_chain_tab;
    DevBuf<int32_t> d_row_woff, d_unit_err;
    DevBuf<double> d_scr_cum;
    DevBuf<uint16_t> d_scr_col;
    DevBuf<uint32_t> d_gpool;        // uniform-word pools of the cooperative-rows units that do not keep them in shared memory
    int64_t gpool_len = 0;           // in 16-byte blocks
    DevBuf<uint64_t> d_unit_clock;
    int mode = BAYES_MODE_FAST;      // sampling mode of the next upload (bayes_gibbs_set_mode)
    int built_mode = BAYES_MODE_FAST;  // mode the uploaded units were packed for
    int64_t scr_len = 0, q_len = 0;  // global scratch of the fast-mode units whose matrix is not in shared memory

    std::vector<Launch> launches;
    std::vector<int64_t> unit_group;    // launch order -> index of the unit in formation order (stable across rebuilds)
    std::vector<double> unit_model_us;  // time model's per-iteration estimate at the unit's shape, before calibration
    int32_t calibration = -1;           // iterations of the calibration pass: -1 = automatic, 0 = off
    bool calibrated = false;
    // the per-unit corrections of the last calibration pass and a fingerprint of the batch they belong to: a caller that uploads the
    // same batch again (other seeds, the next pass of a pipeline over the same library) does not pay for the pass, and its upload stays
    // asynchronous
    std::vector<double> calib_scale;
    uint64_t calib_fp = 0;
    std::vector<double> unit_pred;  // time model's estimate per unit (seconds), launch order
    std::vector<double> unit_feat;  // 4 per unit: chain columns, uniform slabs, uniform blocks, uniform block-columns
    int32_t last_launches = 0;
    // wall-clock seconds of the host phases of the last batch: prepare (fold, cover, tables), unit formation + packing,
    // H2D enqueue, launch enqueue, D2H enqueue, final synchronise (= device time not hidden behind the host)
    double phase_s[6] = {0, 0, 0, 0, 0, 0};

    int64_t trace_locus = -1;
    int32_t trace_chain = 0, trace_iters = 0, trace_T = 0, trace_slot = 0;
    int64_t trace_unit = -1;

    // launch classes run concurrently: each on its own auxiliary stream, forked from / joined to `stream` by events
    std::vector<cudaStream_t> aux;
    std::vector<cudaEvent_t> join_ev, gate_ev;
    cudaEvent_t fork_ev = nullptr;
    cudaStream_t gate = nullptr;  // releases the class launches one by one (bayes_gibbs_run)
};

namespace {

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int use_device(bayes_gibbs_ctx *ctx) { return cuda_rc(cudaSetDevice(ctx->device), "cudaSetDevice"); }

int host_thread_count(const bayes_gibbs_ctx *ctx, int64_t work) {
    int n = ctx->host_threads > 0 ? ctx->host_threads : (int)std::thread::hardware_concurrency();
    return (int)std::max<int64_t>(1, std::min<int64_t>(n, work));
}

template <typename F>
void parallel_for(int n_threads, int64_t n, F fn) {
    if (n_threads <= 1) {
        for (int64_t i = 0; i < n; i++) fn(i);
        return;
    }
    std::atomic<int64_t> next(0);
    auto work = [&]() {
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n) return;
            fn(i);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; t++) pool.emplace_back(work);
    for (auto &t : pool) t.join();
}

template <typename T>
int append(PinBuf<T> &dst, const std::vector<T> &src) {
    const size_t at = dst.n;
    int rc = dst.resize(at + src.size());
    if (rc) return rc;
    if (!src.empty()) memcpy(dst.p + at, src.data(), src.size() * sizeof(T));
    return BAYES_OK;
}

template <typename T>
int h2d(DevBuf<T> &d, const PinBuf<T> &h, cudaStream_t st) {
    int rc = d.reserve(h.n);
    if (rc) return rc;
    if (h.n) CU(cudaMemcpyAsync(d.p, h.p, h.n * sizeof(T), cudaMemcpyHostToDevice, st));
    return BAYES_OK;
}

int append_unit(bayes_gibbs_ctx *ctx, PackedUnit &U) {
    int rc;
    U.d.slab_off = (int64_t)ctx->h_slab.n;
    U.d.cell_off = (int64_t)ctx->h_val.n;
    U.d.row_off = (int64_t)ctx->h_row_n.n;
    U.d.lane_off = (int64_t)ctx->h_efflen.n;
    U.d.tab_off = (int64_t)ctx->h_tab_val.n;
    U.d.tabrow_off = (int64_t)ctx->h_tab_row.n;
    U.d.mask_off = (int64_t)(ctx->h_row_mask.n / 2);  // in 64-bit words
    U.d.woff_off = (int64_t)ctx->h_row_woff.n;
    U.d.pooldesc_off = (int64_t)ctx->h_pool_desc.n;
    U.d.meta_off = (int64_t)ctx->h_cell_meta.n;
    U.d.ctab_off = (int64_t)ctx->h_chain_tab.n;
    U.d.scr_off = ctx->scr_len;
    U.d.q_off = ctx->q_len;
    U.d.gpool_off = ctx->gpool_len;
    if ((!U.row_woff.empty() || !U.cell_meta.empty()) && !U.d.scr_smem) {  // fast mode, scratch in global memory
        ctx->scr_len += U.d.n_chain_cells;
        ctx->q_len += 6 * (int64_t)U.d.qcap * std::max(1, U.d.n_cta);
    }
    if (U.d.n_cta > 0 && !U.d.pool_smem) ctx->gpool_len += (int64_t)U.d.n_cta * (U.d.pool_max + 1);
    if ((rc = append(ctx->h_slab, U.slab_cell_off)) || (rc = append(ctx->h_val, U.val)) || (rc = append(ctx->h_col, U.col)) ||
        (rc = append(ctx->h_row_n, U.row_n)) || (rc = append(ctx->h_row_nnz, U.row_nnz)) || (rc = append(ctx->h_row_info, U.row_info)) || (rc = append(ctx->h_row_mask, U.row_mask)) ||
        (rc = append(ctx->h_row_woff, U.row_woff)) || (rc = append(ctx->h_pool_desc, U.pool_desc)) ||
        (rc = append(ctx->h_cell_meta, U.cell_meta)) || (rc = append(ctx->h_chain_tab, U.chain_tab)) ||
        (rc = append(ctx->h_efflen, U.efflen)) || (rc = append(ctx->h_base, U.base_counts)) ||
        (rc = append(ctx->h_tab_val, U.tab_val)) || (rc = append(ctx->h_tab_row, U.tab_row)))
        return rc;
    const size_t ui = ctx->h_units.n;
    if ((rc = ctx->h_units.resize(ui + 1))) return rc;
    ctx->h_units.p[ui] = U.d;
    return BAYES_OK;
}

// Form the units: every (locus, chain) with more than one candidate is a slot; slots with T <= 32 that share Tpad and
// the iteration counts are bundled side by side (most expensive first, so a bundle's slots finish together), the
// others get a wide unit each.  Units are then ordered so that each launch is one (kernel, block size, matrix
// placement) class with its most expensive units first: the hardware block scheduler hands blocks out in index order,
// which makes this the largest-first queue of the reference (assembler.cpp:1886-1916).
int build_units(bayes_gibbs_ctx *ctx, const std::vector<double> *group_scale) {
    ctx->h_units.n = ctx->h_val.n = ctx->h_efflen.n = ctx->h_tab_val.n = ctx->h_col.n = 0;
    ctx->h_slab.n = ctx->h_row_n.n = ctx->h_row_nnz.n = ctx->h_row_info.n = ctx->h_row_mask.n = ctx->h_base.n = ctx->h_tab_row.n = 0;
    ctx->h_row_woff.n = ctx->h_pool_desc.n = ctx->h_cell_meta.n = ctx->h_chain_tab.n = 0;
    ctx->scr_len = ctx->q_len = ctx->gpool_len = 0;
    const bool fast = ctx->mode == BAYES_MODE_FAST;
    ctx->built_mode = ctx->mode;
    ctx->launches.clear();
    ctx->trace_unit = -1;

    std::vector<std::vector<SlotRef> > groups;  // one per unit
    std::vector<char> group_wide;
    typedef std::tuple<int, int, int> BundleKey;  // (Tpad, burn, samples)
    std::map<BundleKey, std::vector<SlotRef> > pools;
    for (size_t l = 0; l < ctx->loci.size(); l++) {
        const PreLocus &L = ctx->loci[l];
        if (L.T == 1) continue;  // closed form on the host (gibbsSampler.cpp:57-67)
        for (int c = 0; c < L.n_chains; c++) {
            SlotRef s{&L, c, ctx->out_base[l] + (int64_t)c * L.T};
            const int tp = slot_lanes(L.T);
            if (tp == 0) {
                groups.push_back(std::vector<SlotRef>(1, s));
                group_wide.push_back(1);
            } else {
                pools[BundleKey(tp, L.burn, L.samples)].push_back(s);
            }
        }
    }
    for (auto &kv : pools) {
        std::vector<SlotRef> &v = kv.second;
        const int per = 32 / std::get<0>(kv.first);
        std::stable_sort(v.begin(), v.end(), [](const SlotRef &a, const SlotRef &b) { return a.locus->cost > b.locus->cost; });
        // A bundle takes slots until the lanes of warp 0 are used up or (fast mode) its shared-memory footprint reaches the
        // budget of one warp: the SM holds 227 KB and 32 warps, and a unit that needs more than its share per warp either idles
        // warps (when it gets them) or keeps warp slots empty (when it does not).
        const char *ub = getenv("BAYES_UNIT_SMEM");  // experiment knob, bytes
        const double budget = ub ? atof(ub) : (fast ? 12288.0 : 1e12);
        for (size_t i = 0; i < v.size();) {
            size_t j = i;
            double bytes = 2048.0;  // per-candidate state of the 32 lanes, keys, control words
            int64_t pool_words = 0, chain_rows = 0;
            while (j < v.size() && (int)(j - i) < per) {
                const PreLocus &L = *v[j].locus;
                int64_t chain = 0, chain_cells = 0;
                for (size_t r = 0; r < L.row_n.size(); r++)
                    if (L.row_n[r] >= kSmallCount) { chain++; chain_cells += L.row_ptr[r + 1] - L.row_ptr[r]; }
                // tiles: 12 bytes per cell (+ a few padding lanes), 10 per scratch entry, 16 per queued node, 5 per pool word, the CDF table
                const double b = (g_tile_stream ? 0.0 : 13.0 * L.cell_val.size()) + 18.0 * chain_cells + 12.0 * chain + 5.0 * L.pool_words + 8.0 * L.prior(v[j].chain).tab_val.size() + 64.0;
                const int64_t pw = 4 * ((L.pool_words + 3) / 4);
                if (j > i && (bytes + b > budget || (fast && (pool_words + pw >= kMaxTilePoolWords || chain_rows + chain >= (1 << 15))))) break;
                bytes += b;
                pool_words += pw;
                chain_rows += chain;
                j++;
            }
            if (fast && (pool_words >= kMaxTilePoolWords || chain_rows >= (1 << 15))) {
                set_error("a locus with at most 32 candidates has more rows than the fast sampling mode addresses (use BAYES_MODE_REPLAY)");
                return BAYES_E_INVALID;
            }
            groups.push_back(std::vector<SlotRef>(v.begin() + i, v.begin() + j));
            group_wide.push_back(0);
            i = j;
        }
    }
    const int64_t n_units = (int64_t)groups.size();
    if (group_scale && (int64_t)group_scale->size() != n_units) group_scale = nullptr;  // (corrections of another batch: none)
    std::vector<int> unit_cta(n_units, 0);  // > 0: cooperative-rows unit (fast mode, T > 32), CTAs of its cluster

    // ---- shape of every unit before it is packed: its A-step items in processing order (chain items, then uniform
    // items, pack.cpp) -> slab widths for any slab height
    // (a rebuild after the calibration pass forms the same units: their shapes and packed matrices are kept)
    const bool reuse = group_scale && (int64_t)ctx->shape_cache.size() == n_units && (int64_t)ctx->packed_cache.size() == n_units;
    std::vector<UnitShape> &shapes = ctx->shape_cache;
    if (!reuse) {
        shapes.assign(n_units, UnitShape());
        ctx->packed_cache.clear();
        ctx->packed_h.clear();
    }
    if (!reuse) parallel_for(host_thread_count(ctx, n_units), n_units, [&](int64_t i) {
        UnitShape &S = shapes[i];
        std::vector<ItemShape> items;
        build_items(groups[i], &items, &S.n_chain);
        S.nnz.reserve(items.size());
        S.blocks.reserve(items.size());
        for (const ItemShape &it : items) {
            S.nnz.push_back(it.nnz);
            S.blocks.push_back(it.blocks);
        }
        S.iters = (double)groups[i][0].locus->burn + groups[i][0].locus->samples;
        S.wide = group_wide[i] != 0;
    });
    // Time model in microseconds per iteration, fitted by least squares to the measured durations of the 4 183 units of
    // config 2 on a FULL B200 (tools/schedule_report.py -> gpurun_out/unit_stats.npz, profiles/README.md): a chain slab
    // costs chain_us(h) per column (one conditional binomial of its slowest lane per step: fewer lanes, fewer divergent
    // paths), a uniform slab a fixed part (row sum) plus, per Philox block of its busiest lane, a fixed part and a
    // little per column; the one-warp E-step, the two block barriers and the loop overhead cost kEstepUs.  The warps of
    // a unit take slabs from a shared counter, heaviest first (greedy list scheduling).
    auto chain_us = [](int h) { return h >= 32 ? 1.3 : h >= 16 ? 1.15 : h >= 8 ? 0.95 : 0.9; };
    const double kEstepUs = 12.5, kUniformSlabUs = 0.8, kUniformBlockUs = 1.55, kUniformColumnUs = 0.16;
    auto slab_layout = [](const UnitShape &S, int h, int *chain_slabs, int *n_slabs) {
        const int n_items = (int)S.nnz.size();
        *chain_slabs = (S.n_chain + h - 1) / h;
        *n_slabs = *chain_slabs + (n_items - S.n_chain + h - 1) / h;
    };
    // Fast mode (gibbs_fast.cu): a chain slab only compacts its rows' cells in play (cheap, per column); the binomials
    // are the nodes of the rows' splitting trees, processed level by level by ALL threads of the unit, 32 nodes per warp
    // pass whichever row they belong to; a uniform slab reads its words from the pool (generated during the E-step by
    // the other warps, after it by a one-warp unit).  A row's cells in play are estimated as half its stored cells.
    const double kFastEstepUs = 9.0, kFastChainColUs = 0.10, kFastLevelUs = 1.9, kFastBarrierUs = 0.25, kFastUniformSlabUs = 0.7,
                 kFastUniformBlockUs = 0.35, kFastUniformColumnUs = 0.10, kFastPoolUs = 0.6;
    auto tree_batches = [](const UnitShape &S, int lanes, int *levels) {
        // nodes with two or more cells per tree level, summed over the chain rows; returns the warp passes they need
        double passes = 0;
        int lv = 0;
        for (int d = 0; d < 13; d++) {
            int64_t nodes = 0;
            for (int i = 0; i < S.n_chain; i++) {
                const int k = std::max(2, (S.nnz[i] + 1) / 2);
                const int64_t w = (int64_t)1 << d;
                if (w < k) nodes += std::min<int64_t>(w, k - w);
            }
            if (nodes == 0) break;
            lv++;
            passes += (double)((nodes + lanes - 1) / lanes);
        }
        *levels = lv;
        return passes;
    };
    // Bundles in fast mode are lane-per-cell tiles: a tile costs the same whatever its rows look like (one scan), a tile of
    // pooled-uniform rows a little more per extra Philox block of its busiest row; tiles are handed out to the warps.
    // (least squares over the 4 977 units of config 2 on a full B200: ~18 us fixed, ~1.1 us per tile and warp, ~3 us per tree pass)
    const double kTileEstepUs = 18.0, kTileChainUs = 1.0, kTileUniformUs = 1.1, kTileBlockUs = 0.3, kTileLevelUs = 3.0;
    auto iteration_us_tiles = [&](const UnitShape &S, int warps) {
        const int n_items = (int)S.nnz.size();
        double chain_cells = 0, unif_cells = 0, extra = 0, blocks_total = 0;
        for (int i = 0; i < n_items; i++) {
            if (i < S.n_chain) chain_cells += S.nnz[i];
            else {
                unif_cells += S.nnz[i];
                blocks_total += S.blocks[i];
                extra += (S.blocks[i] - 1) * (double)S.nnz[i];
            }
        }
        const double chain_tiles = ceil(chain_cells / 30.0), unif_tiles = ceil(unif_cells / 30.0);
        const double rows = chain_tiles * kTileChainUs + unif_tiles * kTileUniformUs + extra / 30.0 * kTileBlockUs;
        int levels = 0;
        const double passes = tree_batches(S, 32 * warps, &levels);
        const double pool = warps > 1 ? 0.0 : kFastPoolUs * ceil(blocks_total / 32.0);
        return kTileEstepUs + pool + std::max(rows / warps, std::min(rows, kTileUniformUs)) + passes * kTileLevelUs +
               (levels + 2) * kFastBarrierUs * (warps > 1 ? 1.0 : 0.2) + (warps > 32 ? 3.0 : 0.0);  // (> 32 warps: a cluster, its barrier and count exchange)
    };
    // Cooperative rows (gibbs_rows.cu): the replicated E-step and the cluster barrier, one warp per row (the rows are dealt over
    // all warps of all CTAs of the cluster), the tree levels of the CTA with most of them.
    const double kRowsEstepUs = 14.0, kRowsRowUs = 0.6;
    auto iteration_us_rows = [&](const UnitShape &S, int warps, int n_cta) {
        const int n_items = (int)S.nnz.size();
        double wide_rows = 0;  // rows whose cells in play may not fit one round are walked twice
        for (int i = S.n_chain; i < n_items; i++) wide_rows += S.nnz[i] > 64 ? 1.0 : 0.0;
        int levels = 0;
        const double passes = tree_batches(S, 32 * warps * n_cta, &levels);
        return kRowsEstepUs + ceil((n_items + wide_rows) / ((double)warps * n_cta)) * kRowsRowUs + passes * kFastLevelUs + (levels + 2) * kFastBarrierUs;
    };
    auto iteration_us_fast = [&](const UnitShape &S, int warps, int h) {
        if (!S.wide) return iteration_us_tiles(S, warps);
        if (unit_cta[&S - shapes.data()] > 0) return iteration_us_rows(S, warps, unit_cta[&S - shapes.data()]);
        int chain_slabs, n_slabs;
        slab_layout(S, h, &chain_slabs, &n_slabs);
        const int n_items = (int)S.nnz.size();
        std::vector<double> load(warps, 0.0);
        double worst = 0, blocks_total = 0;
        for (int s = 0; s < n_slabs; s++) {
            const bool chain = s < chain_slabs;
            const int first = chain ? s * h : S.n_chain + (s - chain_slabs) * h, last = chain ? S.n_chain : n_items;
            int width = 0, blocks = 0;
            for (int l = 0; l < h && first + l < last; l++) {
                width = std::max(width, S.nnz[first + l]);
                blocks = std::max(blocks, S.blocks[first + l]);
                blocks_total += S.blocks[first + l];
            }
            const double us = chain ? 0.3 + width * kFastChainColUs : kFastUniformSlabUs + blocks * (kFastUniformBlockUs + width * kFastUniformColumnUs);
            double &l0 = *std::min_element(load.begin(), load.end());
            l0 += us;
            worst = std::max(worst, l0);
        }
        int levels = 0;
        const double passes = tree_batches(S, 32 * warps, &levels);
        const double pool = warps > 1 ? 0.0 : kFastPoolUs * ceil(blocks_total / 32.0);
        return kFastEstepUs + pool + worst + passes * kFastLevelUs + (levels + 2) * kFastBarrierUs * (warps > 1 ? 1.0 : 0.2);
    };
    auto iteration_us_replay = [&](const UnitShape &S, int warps, int h) {
        int chain_slabs, n_slabs;
        slab_layout(S, h, &chain_slabs, &n_slabs);
        const int n_items = (int)S.nnz.size();
        std::vector<double> load(warps, 0.0);  // list scheduling in slab order
        double worst = 0;
        for (int s = 0; s < n_slabs; s++) {
            const bool chain = s < chain_slabs;
            const int first = chain ? s * h : S.n_chain + (s - chain_slabs) * h, last = chain ? S.n_chain : n_items;
            int width = 0, blocks = 0;
            for (int l = 0; l < h && first + l < last; l++) {
                width = std::max(width, S.nnz[first + l]);
                blocks = std::max(blocks, S.blocks[first + l]);
            }
            const double us = chain ? width * chain_us(h) : kUniformSlabUs + blocks * (kUniformBlockUs + width * kUniformColumnUs);
            double &l0 = *std::min_element(load.begin(), load.end());
            l0 += us;
            worst = std::max(worst, l0);
        }
        return worst + kEstepUs;
    };
    auto iteration_us = [&](const UnitShape &S, int warps, int h) { return fast ? iteration_us_fast(S, warps, h) : iteration_us_replay(S, warps, h); };
    // group_scale (calibrate_units): measured / modelled per-iteration time of the unit at its first shape
    auto unit_time = [&](const UnitShape &S, int warps, int h) {
        const double scale = group_scale ? (*group_scale)[&S - shapes.data()] : 1.0;
        return 1e-6 * S.iters * iteration_us(S, warps, h) * scale;
    };
    // totals the model is made of (exposed through bayes_gibbs_unit_stats for calibration)
    auto unit_features = [&](const UnitShape &S, int h, double *f4) {
        int chain_slabs, n_slabs;
        slab_layout(S, h, &chain_slabs, &n_slabs);
        const int n_items = (int)S.nnz.size();
        f4[0] = f4[1] = f4[2] = f4[3] = 0;
        for (int s = 0; s < n_slabs; s++) {
            const bool chain = s < chain_slabs;
            const int first = chain ? s * h : S.n_chain + (s - chain_slabs) * h, last = chain ? S.n_chain : n_items;
            int width = 0, blocks = 0;
            for (int l = 0; l < h && first + l < last; l++) {
                width = std::max(width, S.nnz[first + l]);
                blocks = std::max(blocks, S.blocks[first + l]);
            }
            if (chain) f4[0] += width;
            else { f4[1] += 1; f4[2] += blocks; f4[3] += (double)blocks * width; }
        }
        if (fast) {  // fast mode: uniform slabs and blocks share one entry, the freed one holds the tree's warp passes at one warp
            int levels = 0;
            f4[1] = f4[1] + 0.001 * f4[2];
            f4[2] = tree_batches(S, 32, &levels);
        }
    };
    auto smem_estimate = [&](const UnitShape &S, int h) {
        int chain_slabs, n_slabs;
        slab_layout(S, h, &chain_slabs, &n_slabs);
        const int n_items = (int)S.nnz.size();
        double cells = 0;
        for (int s = 0; s < n_slabs; s++) {
            const bool chain = s < chain_slabs;
            const int first = chain ? s * h : S.n_chain + (s - chain_slabs) * h, last = chain ? S.n_chain : n_items;
            int width = 0;
            for (int l = 0; l < h && first + l < last; l++) width = std::max(width, S.nnz[first + l]);
            cells += (double)h * width;
        }
        return (fast ? 12.0 : 10.0) * cells + (fast ? 16.0 : 12.0) * h * n_slabs + (fast ? 2048.0 : 1024.0);
    };

    // Warps and slab height per unit.  Throughput is best with ONE warp per unit walking full 32-row slabs: every extra
    // warp is parked at the block barrier during the one-warp E-step, and short slabs leave lanes idle.  Extra warps /
    // shorter slabs only buy latency, which matters for the units that would otherwise outlast the rest of the batch.
    // Greedy critical-path reduction: while the longest unit exceeds what the GPU needs for the whole batch anyway
    // (occupied warp-seconds / warps it keeps resident), move that unit one step down the ladder below.
    std::vector<int> unit_warps(n_units, 1), unit_h(n_units, 32);
    // Fast mode, loci with more than 32 candidates: cooperative rows over a thread-block cluster (gibbs_rows.cu).  Such a locus
    // is one long chain of dependent iterations, and what an iteration costs is latency: the more CTAs share its rows, the
    // shorter the iteration -- and the more SMs it keeps from the other loci.  So the cluster is as large as the SMs allow when
    // all such loci of the batch are to run at once (8 loci: 16 CTAs each; 500 loci: one CTA each, matrix streamed from L2).
    int64_t n_rows_units = 0;
    for (int64_t i = 0; i < n_units; i++) n_rows_units += fast && group_wide[i];
    int rows_cta = 1;
    if (n_rows_units > 0) {
        const char *fc = getenv("BAYES_FORCE_CTAS");  // experiment knob
        const int cap = rows_max_cluster(kMaxBlock, max_unit_smem());
        int want = fc ? atoi(fc) : (int)std::min<int64_t>(kMaxCluster, ctx->sm_count / n_rows_units);
        while (rows_cta * 2 <= want && rows_cta * 2 <= cap) rows_cta *= 2;
    }
    {
        // (fast mode: the latency of a unit is in its tree levels, which all threads share whatever the slab height: more warps only)
        static const int ladder_replay[][2] = {{1, 32}, {2, 32}, {4, 32}, {8, 32}, {8, 16}, {16, 16}, {16, 8}};
        // (the tiled kernel runs at 64 registers per thread: a monster locus may take a whole SM's 32 warps)
        // (beyond 32 warps a tiled unit is spread over a thread-block cluster of 2, 4 or 8 CTAs of 32 warps: gibbs_fast.cu)
        static const int ladder_fast[][2] = {{1, 32}, {2, 32}, {4, 32}, {8, 32}, {16, 32}, {32, 32}, {64, 32}, {128, 32}, {256, 32}};
        const int (*ladder)[2] = fast ? ladder_fast : ladder_replay;
        const char *mc = getenv("BAYES_MAX_BUNDLE_CTAS");  // experiment knob: 1 = no clusters for bundles
        const int max_bundle_cta = mc ? std::max(1, std::min(8, atoi(mc))) : (n_units < 12 * (int64_t)ctx->sm_count ? 8 : 1);
        const int n_ladder_all = fast ? (max_bundle_cta >= 8 ? 9 : max_bundle_cta >= 4 ? 8 : max_bundle_cta >= 2 ? 7 : 6) : 7;
        const double capacity = (double)ctx->sm_count * kWarpsPerSM;  // resident warps at the kernels' register count
        const char *spw = getenv("BAYES_SMEM_PER_WARP");  // experiment knob
        const double smem_per_warp = spw ? atof(spw) : (fast ? 16384.0 : 8192.0);
        const char *ce = getenv("BAYES_CRIT");  // experiment knob
        // Fast mode, two regimes (measured on config 4 and on its eighth, the share of one GPU of eight; profiles/README.md):
        //  * a batch of many more units than the SMs hold at once is bound by throughput, and the SMs hold about 17 tiled units each
        //    (shared memory): units of two or four warps are what fills the warp slots, so units are widened generously (0.5: 4 980 ms per
        //    step of config 4 against 5 520 ms at 1.0);
        //  * a batch the GPU holds at once is bound by its longest units, and widening costs throughput (a unit on 32 warps runs an
        //    iteration 1.6 times faster than on one, for 32 times the issue slots): only the units that would outlast the batch are
        //    widened, those as far as a cluster of eight CTAs (1.0 with clusters: 1 040 ms for 5 000 loci against 1 213 ms at 0.5 without).
        const bool small_batch = fast && n_units < 12 * (int64_t)ctx->sm_count;
        const double crit = ce ? atof(ce) : (fast ? (small_batch ? 1.0 : 0.5) : 0.75);
        std::vector<int> step(n_units, 0);
        std::vector<double> t(n_units);
        double occupied = 0;  // sum of warps * time
        typedef std::pair<double, int64_t> Item;
        std::priority_queue<Item> heap;
        for (int64_t i = 0; i < n_units; i++) {
            // A unit's matrix sits in shared memory whatever its block size; the SM runs out of shared memory before it
            // runs out of warp slots when units of much more than 227 KB / 32 keep one warp each, so such units start with as
            // many warps as their footprint pays for (they are parked warps otherwise nobody could use).
            const UnitShape &S = shapes[i];
            int chain_slabs, n_slabs;
            slab_layout(S, 32, &chain_slabs, &n_slabs);
            const double smem_est = smem_estimate(S, 32);
            int k = 0;
            // (tiled units stream their matrix: their footprint does not grow with it, one warp unless they are critical)
            const bool tiled = fast && !S.wide;
            while (!tiled && k + 1 < 4 && ladder[k][0] * smem_per_warp < smem_est && ladder[k + 1][0] < 2 * n_slabs) k++;
            step[i] = k;
            t[i] = unit_time(S, ladder[k][0], ladder[k][1]);
            occupied += ladder[k][0] * t[i];
            heap.push(Item(t[i], i));
        }
        while (!heap.empty()) {
            const Item top = heap.top();
            const int64_t i = top.second;
            // (fast mode: a unit that starts first runs its whole life on a full GPU, up to twice its calibrated time: a tighter bound)
            if (top.first <= crit * occupied / capacity) break;  // the batch is throughput-bound from here on
            heap.pop();
            // next rung that actually shortens the unit
            int k = step[i] + 1;
            double tk = 0;
            const int n_ladder = n_ladder_all;
            for (; k < n_ladder; k++) {
                int chain_slabs, n_slabs;
                slab_layout(shapes[i], ladder[k][1], &chain_slabs, &n_slabs);
                if (!fast && ladder[k][0] >= 2 * n_slabs && ladder[k][0] > ladder[step[i]][0]) continue;  // idle warps only
                if (fast && shapes[i].wide && ladder[k][0] > kMaxBlock / 32) break;                       // the wide kernel: 128 registers
                tk = unit_time(shapes[i], ladder[k][0], ladder[k][1]);
                if (tk < 0.93 * t[i]) break;
            }
            if (k >= n_ladder_all || (fast && shapes[i].wide && ladder[k][0] > kMaxBlock / 32)) continue;  // cannot be shortened further: it stays out of the heap
            occupied += ladder[k][0] * tk - ladder[step[i]][0] * t[i];
            step[i] = k;
            t[i] = tk;
            heap.push(Item(tk, i));
        }
        const char *fw = getenv("BAYES_FORCE_WARPS"), *fh = getenv("BAYES_FORCE_H");  // experiment knobs
        for (int64_t i = 0; i < n_units; i++) {
            unit_warps[i] = fw ? std::max(1, std::min(atoi(fw), (fast && !shapes[i].wide ? 8192 : kMaxBlock) / 32)) : ladder[step[i]][0];
            unit_h[i] = fh ? atoi(fh) : ladder[step[i]][1];
            // A locus with hundreds of candidates (config-3 class) is one long dependent chain of iterations that each walk
            // a thousand rows: all 16 warps, and the tallest slabs that still give every warp one -- its A-step is a scan of
            // the rows' candidate masks (gibbs_kernel.cu, assign_rows_wide), which all lanes of a slab do in lock step, so
            // short slabs only idle lanes (measured on the largest config-3 locus: 4.4 of 32 lanes active at h = 8).
            if (fast && group_wide[i]) {
                // cooperative rows: one warp per row, 16 warps unless the locus has few rows
                unit_cta[i] = rows_cta;
                unit_h[i] = 1;
                const int rows = (int)shapes[i].nnz.size();
                int w = fw ? atoi(fw) : 16;
                while (w > 1 && (int64_t)(w / 2) * rows_cta * 2 > rows) w /= 2;
                unit_warps[i] = std::max(1, std::min(w, kMaxBlock / 32));
            } else if (group_wide[i] && groups[i][0].locus->T > 128 && !fw && !fh) {
                int h = 32, chain_slabs = 0, n_slabs = 0;
                for (; h > 8; h /= 2) {
                    slab_layout(shapes[i], h, &chain_slabs, &n_slabs);
                    if (n_slabs >= kMaxBlock / 32) break;
                }
                slab_layout(shapes[i], h, &chain_slabs, &n_slabs);
                unit_h[i] = h;
                unit_warps[i] = std::max(1, std::min(kMaxBlock / 32, n_slabs));
            }
        }
    }

    std::vector<PackedUnit> &packed = ctx->packed_cache;
    packed.resize(n_units);
    ctx->packed_h.resize(n_units, 0);
    parallel_for(host_thread_count(ctx, n_units), n_units, [&](int64_t i) {
        if (unit_cta[i] > 0) {
            if (!reuse || ctx->packed_h[i] != -unit_cta[i]) {
                pack_unit_rows(groups[i], unit_cta[i], &packed[i]);
                // (a cluster whose replicated per-candidate state leaves no room: one CTA, state only)
                if (packed[i].d.smem_bytes > (int32_t)max_unit_smem() && unit_cta[i] > 1) pack_unit_rows(groups[i], 1, &packed[i]);
            }
            ctx->packed_h[i] = -unit_cta[i];
        } else {
            if (!reuse || ctx->packed_h[i] != unit_h[i]) pack_unit(groups[i], group_wide[i] != 0, unit_h[i], fast, &packed[i]);
            ctx->packed_h[i] = unit_h[i];
            // a tiled unit above 32 warps: a cluster of CTAs of 32 warps
            if (packed[i].d.n_cta > 1) packed[i].d.smem_bytes -= kTileClusterBytes;  // (a cached unit that was a cluster in the first build)
            packed[i].d.n_cta = 0;
            if (unit_warps[i] > 32) {
                if (packed[i].d.smem_bytes + kTileClusterBytes <= (int32_t)max_unit_smem()) {
                    packed[i].d.n_cta = unit_warps[i] / 32;
                    packed[i].d.smem_bytes += kTileClusterBytes;
                }
                unit_warps[i] = 32;
            }
        }
        packed[i].d.block = 32 * unit_warps[i];
        packed[i].cost = unit_time(shapes[i], unit_warps[i] * std::max(1, unit_cta[i] > 0 ? 1 : packed[i].d.n_cta), unit_h[i]);  // estimated duration: the launch-order key
        unit_features(shapes[i], unit_h[i], packed[i].features);
    });

    // One launch per class = (kernel, block size, matrix placement, shared-memory bucket) and duration band.  The bucket
    // matters: a launch has ONE dynamic shared-memory size, and sizing it for the largest unit of a broad class would
    // cut the number of resident units per SM several-fold.  The band matters: the hardware block scheduler hands
    // blocks out launch by launch, so the launches are issued longest units first across ALL classes -- the
    // largest-first queue of the reference (assembler.cpp:1886-1916) -- with the blocks of 8 or more warps, which only
    // fit an SM that is still nearly empty, ahead of everything else.
    const int32_t smem_limit = (int32_t)max_unit_smem();
    auto smem_bucket = [smem_limit](int32_t bytes) -> int32_t {
        int32_t step = bytes <= 8 * 1024 ? 2048 : bytes <= 32 * 1024 ? 4096 : 16 * 1024;
        return std::min(smem_limit, (bytes + step - 1) / step * step);
    };
    for (int64_t i = 0; i < n_units; i++)
        if (unit_cta[i] > 0 && packed[i].d.chain_max >= (1 << 16)) {  // direct-draw tasks name a row of their CTA in 16 bits
            set_error("a locus with more than 32 candidates has more than 65 535 rows of 20 or more fragments per CTA: more than the fast sampling mode addresses (use BAYES_MODE_REPLAY)");
            return BAYES_E_INVALID;
        }
    for (int64_t i = 0; i < n_units; i++)
        if (packed[i].d.smem_bytes > smem_limit) {
            set_error("a unit needs " + std::to_string(packed[i].d.smem_bytes) + " bytes of shared memory, more than the device offers");
            return BAYES_E_INVALID;
        }
    auto band_of = [](const PackedUnit &u) -> int {
        if (u.d.block >= 256) return 1000;                            // critical, hard to place later
        return (int)floor(log(std::max(u.cost, 1e-6)) / log(1.4));   // +-18 % of duration per band
    };
    typedef std::tuple<int, int, int, int, int> ClassKey;  // (-band, wide | CTAs per cluster << 1, mat_smem | scr_smem << 1, block, smem bucket)
    auto class_of = [&](const PackedUnit &u) {
        // (critical band: the largest blocks first -- a 1024-thread block needs a whole SM's registers and would wait, with
        // everything queued behind it, for an SM to drain completely if smaller blocks were spread over all SMs before it)
        // (clusters first, the largest ahead: they need that many SMs of one GPC free at the same time)
        return ClassKey(-band_of(u), u.d.wide + 2 * (kMaxCluster - u.d.n_cta), u.d.mat_smem | (u.d.scr_smem << 1), u.d.block >= 256 ? -u.d.block : u.d.block, smem_bucket(u.d.smem_bytes));
    };
    std::vector<ClassKey> key(n_units);
    for (int64_t i = 0; i < n_units; i++) key[i] = class_of(packed[i]);
    std::vector<int64_t> order(n_units);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
        if (key[a] != key[b]) return key[a] < key[b];
        return packed[a].cost > packed[b].cost;
    });
    int rc;
    ctx->unit_pred.clear();
    ctx->unit_feat.clear();
    ctx->unit_group.assign(order.begin(), order.end());
    ctx->unit_model_us.clear();
    for (int64_t i = 0; i < n_units; i++) {
        if ((rc = append_unit(ctx, packed[order[i]]))) return rc;
        ctx->unit_pred.push_back(packed[order[i]].cost);
        ctx->unit_model_us.push_back(iteration_us(shapes[order[i]], unit_warps[order[i]] * (unit_cta[order[i]] > 0 ? 1 : std::max(1, packed[order[i]].d.n_cta)), unit_h[order[i]]));
        for (int k = 0; k < 4; k++) ctx->unit_feat.push_back(packed[order[i]].features[k]);
    }
    const UnitDev *units = ctx->h_units.p;
    for (int64_t i = 0; i < n_units;) {
        int64_t j = i;
        const ClassKey &k = key[order[i]];
        while (j < n_units && key[order[j]] == k) j++;
        ctx->launches.push_back(Launch{i, j - i, units[i].block, units[i].wide != 0, units[i].mat_smem != 0, units[i].scr_smem != 0, (size_t)std::get<4>(k), units[i].n_cta});
        i = j;
    }
    if (ctx->trace_locus >= 0) {
        const int64_t want = ctx->out_base[ctx->trace_locus] + (int64_t)ctx->trace_chain * ctx->trace_T;
        for (int64_t i = 0; i < n_units && ctx->trace_unit < 0; i++)
            for (int g = 0; g < units[i].n_slots; g++)
                if (units[i].out_off[g] == want) {
                    ctx->trace_unit = i;
                    ctx->trace_slot = g;
                    break;
                }
    }
    return BAYES_OK;
}

KernelArgs kernel_args(bayes_gibbs_ctx *ctx) {
    KernelArgs a;
    memset(&a, 0, sizeof(a));
    a.units = ctx->d_units.p;
    a.val = ctx->d_val.p;
    a.col = ctx->d_col.p;
    a.slab_cell_off = ctx->d_slab.p;
    a.row_n = ctx->d_row_n.p;
    a.row_nnz = ctx->d_row_nnz.p;
    a.row_info = ctx->d_row_info.p;
    a.row_mask = ctx->d_row_mask.p;
    a.efflen = ctx->d_efflen.p;
    a.base_counts = ctx->d_base.p;
    a.tab_val = ctx->d_tab_val.p;
    a.tab_row = ctx->d_tab_row.p;
    a.out_occ = ctx->d_occ.p;
    a.out_mean = ctx->d_mean.p;
    a.out_m2 = ctx->d_m2.p;
    a.out_counts = ctx->d_counts.p;
    a.unit_clock = ctx->d_unit_clock.p;
    a.row_woff = ctx->d_row_woff.p;
    a.pool_desc = ctx->d_pool_desc.p;
    a.cell_meta = ctx->d_cell_meta.p;
    a.chain_tab = ctx->d_chain_tab.p;
    a.scr_cum = ctx->d_scr_cum.p;
    a.scr_col = ctx->d_scr_col.p;
    a.queue = ctx->d_queue.p;
    a.gpool = ctx->d_gpool.p;
    a.unit_err = ctx->d_unit_err.p;
    a.trace_unit = ctx->trace_unit;
    a.trace_slot = ctx->trace_slot;
    a.trace_iters = ctx->trace_iters;
    a.tr_theta = ctx->d_tr_theta.p;
    a.tr_counts = ctx->d_tr_counts.p;
    a.tr_b = ctx->d_tr_b.p;
    a.tr_init = ctx->d_tr_init.p;
    return a;
}

int upload_units(bayes_gibbs_ctx *ctx) {
    int rc;
    cudaStream_t st = ctx->stream;
    if ((rc = h2d(ctx->d_units, ctx->h_units, st)) || (rc = h2d(ctx->d_val, ctx->h_val, st)) || (rc = h2d(ctx->d_col, ctx->h_col, st)) ||
        (rc = h2d(ctx->d_slab, ctx->h_slab, st)) || (rc = h2d(ctx->d_row_n, ctx->h_row_n, st)) ||
        (rc = h2d(ctx->d_row_nnz, ctx->h_row_nnz, st)) || (rc = h2d(ctx->d_row_info, ctx->h_row_info, st)) ||
        (rc = h2d(ctx->d_row_mask, ctx->h_row_mask, st)) || (rc = h2d(ctx->d_row_woff, ctx->h_row_woff, st)) ||
        (rc = h2d(ctx->d_pool_desc, ctx->h_pool_desc, st)) || (rc = h2d(ctx->d_cell_meta, ctx->h_cell_meta, st)) ||
        (rc = h2d(ctx->d_chain_tab, ctx->h_chain_tab, st)) || (rc = ctx->d_scr_cum.reserve((size_t)ctx->scr_len)) ||
        (rc = ctx->d_scr_col.reserve((size_t)ctx->scr_len)) || (rc = ctx->d_queue.reserve((size_t)ctx->q_len)) || (rc = ctx->d_gpool.reserve(4 * (size_t)ctx->gpool_len)) ||
        (rc = ctx->d_unit_err.reserve(ctx->h_units.n + 1)) || (rc = ctx->h_unit_err.resize(ctx->h_units.n + 1)) ||
        (rc = h2d(ctx->d_efflen, ctx->h_efflen, st)) || (rc = h2d(ctx->d_base, ctx->h_base, st)) ||
        (rc = h2d(ctx->d_tab_val, ctx->h_tab_val, st)) || (rc = h2d(ctx->d_tab_row, ctx->h_tab_row, st)))
        return rc;
    return BAYES_OK;
}

}  // namespace

extern "C" {

const char *bayes_last_error(void) { return g_last_error.c_str(); }

int bayes_gibbs_create(int device, bayes_gibbs_ctx **out) {
    if (!out) {
        set_error("bayes_gibbs_create: null out pointer");
        return BAYES_E_INVALID;
    }
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev < 1) {
        set_error(std::string("bayes_gibbs_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU path");
        return BAYES_E_NODEVICE;
    }
    if (device < 0 || device >= n_dev) {
        set_error("bayes_gibbs_create: device index out of range");
        return BAYES_E_INVALID;
    }
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("bayes_gibbs_create: kernels are built for sm_100a only (found sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + ")");
        return BAYES_E_NODEVICE;
    }
    int rc = configure_kernels();
    if (rc) return rc;
    if ((rc = configure_fast_kernels())) return rc;
    if ((rc = configure_rows_kernels())) return rc;
    bayes_gibbs_ctx *ctx = new (std::nothrow) bayes_gibbs_ctx();
    if (!ctx) return BAYES_E_NOMEM;
    if (const char *m = getenv("BAYES_MODE")) ctx->mode = strcmp(m, "replay") == 0 ? BAYES_MODE_REPLAY : BAYES_MODE_FAST;
    if (const char *m = getenv("BAYES_TILE_STREAM")) g_tile_stream = atoi(m) != 0;  // experiment knob
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
    rc = cuda_rc(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking), "cudaStreamCreate");
    if (rc) {
        delete ctx;
        return rc;
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return BAYES_OK;
}

void bayes_gibbs_destroy(bayes_gibbs_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->h_units.release(); ctx->h_val.release(); ctx->h_efflen.release(); ctx->h_tab_val.release(); ctx->h_col.release();
    ctx->h_slab.release(); ctx->h_row_n.release(); ctx->h_row_nnz.release(); ctx->h_row_info.release(); ctx->h_base.release();
    ctx->h_tab_row.release(); ctx->h_occ.release(); ctx->h_counts.release(); ctx->h_mean.release(); ctx->h_m2.release();
    ctx->h_row_mask.release(); ctx->d_row_mask.release();
    ctx->h_pool_desc.release(); ctx->h_row_woff.release(); ctx->h_unit_err.release(); ctx->h_cell_meta.release(); ctx->h_chain_tab.release();
    ctx->d_cell_meta.release(); ctx->d_chain_tab.release();
    ctx->d_pool_desc.release(); ctx->d_queue.release(); ctx->d_row_woff.release(); ctx->d_unit_err.release();
    ctx->d_scr_cum.release(); ctx->d_scr_col.release(); ctx->d_gpool.release();
    ctx->d_units.release(); ctx->d_val.release(); ctx->d_efflen.release(); ctx->d_tab_val.release(); ctx->d_mean.release();
    ctx->d_m2.release(); ctx->d_tr_theta.release(); ctx->d_col.release(); ctx->d_slab.release(); ctx->d_row_n.release();
    ctx->d_row_nnz.release(); ctx->d_row_info.release(); ctx->d_base.release(); ctx->d_tab_row.release(); ctx->d_occ.release();
    ctx->d_unit_clock.release();
    ctx->d_counts.release(); ctx->d_tr_counts.release(); ctx->d_tr_b.release(); ctx->d_tr_init.release();
    for (cudaStream_t st : ctx->aux) cudaStreamDestroy(st);
    for (cudaEvent_t ev : ctx->join_ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : ctx->gate_ev) cudaEventDestroy(ev);
    if (ctx->gate) cudaStreamDestroy(ctx->gate);
    if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int bayes_gibbs_set_stream(bayes_gibbs_ctx *ctx, void *cuda_stream) {
    if (!ctx) return BAYES_E_INVALID;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return BAYES_OK;
}

void *bayes_gibbs_get_stream(bayes_gibbs_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int bayes_gibbs_set_host_threads(bayes_gibbs_ctx *ctx, int n_threads) {
    if (!ctx || n_threads < 0) return BAYES_E_INVALID;
    ctx->host_threads = n_threads;
    return BAYES_OK;
}

int bayes_gibbs_clear(bayes_gibbs_ctx *ctx) {
    if (!ctx) return BAYES_E_INVALID;
    int rc = use_device(ctx);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->loci.clear();
    ctx->phase_s[0] = 0;
    ctx->out_base.clear();
    ctx->out_len = 0;
    ctx->h_units.n = 0;
    ctx->launches.clear();
    ctx->uploaded = ctx->ran = ctx->downloaded = false;
    ctx->trace_locus = -1;
    ctx->trace_unit = -1;
    ctx->trace_iters = 0;
    return BAYES_OK;
}

int bayes_gibbs_submit(bayes_gibbs_ctx *ctx, int32_t n_loci, const bayes_locus_desc *loci, int64_t *first_id) {
    if (!ctx || n_loci < 0 || (n_loci > 0 && !loci)) {
        set_error("bayes_gibbs_submit: bad arguments");
        return BAYES_E_INVALID;
    }
    if (ctx->uploaded) {
        set_error("bayes_gibbs_submit: batch already uploaded; call bayes_gibbs_clear first");
        return BAYES_E_STATE;
    }
    const double t0 = now_s();
    // fold rows, read cover, CDF table: independent per locus
    std::vector<PreLocus> pre(n_loci);
    std::vector<int> prc(n_loci, 0);
    std::vector<std::string> perr(n_loci);
    parallel_for(host_thread_count(ctx, n_loci), n_loci, [&](int64_t i) { prc[i] = prepare_locus(loci[i], &pre[i], &perr[i]); });
    for (int32_t i = 0; i < n_loci; i++)
        if (prc[i] != BAYES_OK) {
            set_error("bayes_gibbs_submit: locus " + std::to_string(i) + ": " + perr[i]);
            return prc[i];
        }
    std::lock_guard<std::mutex> guard(ctx->submit_lock);  // the preparation above ran unlocked; ids are handed out here
    if (ctx->uploaded) {
        set_error("bayes_gibbs_submit: batch already uploaded; call bayes_gibbs_clear first");
        return BAYES_E_STATE;
    }
    if (first_id) *first_id = (int64_t)ctx->loci.size();
    ctx->loci.reserve(ctx->loci.size() + n_loci);
    for (int32_t i = 0; i < n_loci; i++) {
        ctx->out_base.push_back(ctx->out_len);
        ctx->out_len += (int64_t)pre[i].T * pre[i].n_chains;
        ctx->loci.push_back(std::move(pre[i]));
    }
    ctx->phase_s[0] += now_s() - t0;
    return BAYES_OK;
}

static int calibrate_units(bayes_gibbs_ctx *ctx, int32_t iters, std::vector<double> *group_scale);

// what the unit time model sees of a batch (shapes, iteration counts, mode, calibration length), hashed
static uint64_t batch_fingerprint(const bayes_gibbs_ctx *ctx, int32_t calib) {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&h](uint64_t v) { h = (h ^ v) * 1099511628211ull; };
    mix((uint64_t)ctx->mode);
    mix((uint64_t)calib);
    mix((uint64_t)ctx->loci.size());
    for (const PreLocus &L : ctx->loci) {
        mix((uint64_t)L.T | ((uint64_t)L.n_chains << 32));
        mix((uint64_t)(uint32_t)L.burn | ((uint64_t)(uint32_t)L.samples << 32));
        mix((uint64_t)L.row_id.size() | ((uint64_t)L.cell_val.size() << 24));
        mix((uint64_t)L.n_frag | ((uint64_t)L.pool_words << 32));
    }
    return h ? h : 1;
}

int bayes_gibbs_upload(bayes_gibbs_ctx *ctx) {
    if (!ctx) return BAYES_E_INVALID;
    int rc = use_device(ctx);
    if (rc) return rc;
    if (ctx->uploaded) CU(cudaStreamSynchronize(ctx->stream));  // a second upload rewrites the pinned staging buffers
    const double t0 = now_s();
    // the same batch as last time (as far as the time model can tell): its calibration still holds, no timing pass, no synchronisation
    int32_t calib_want = ctx->calibration;
    if (const char *e = getenv("BAYES_CALIB_ITERS")) calib_want = atoi(e);
    const uint64_t fp = batch_fingerprint(ctx, calib_want);
    const bool cached = calib_want != 0 && ctx->trace_locus < 0 && fp == ctx->calib_fp && !ctx->calib_scale.empty();
    if ((rc = build_units(ctx, cached ? &ctx->calib_scale : nullptr))) return rc;
    double t1 = now_s();
    ctx->phase_s[1] = t1 - t0;
    if (ctx->trace_locus >= 0) {
        const size_t n = (size_t)ctx->trace_iters * ctx->trace_T;
        if ((rc = ctx->d_tr_theta.reserve(n)) || (rc = ctx->d_tr_counts.reserve(n)) || (rc = ctx->d_tr_b.reserve(ctx->trace_iters)) ||
            (rc = ctx->d_tr_init.reserve(ctx->trace_T)))
            return rc;
    }
    if ((rc = upload_units(ctx))) return rc;
    if ((rc = ctx->d_unit_clock.reserve(3 * ctx->h_units.n + 3))) return rc;
    const size_t n_out = (size_t)ctx->out_len;
    if ((rc = ctx->d_occ.reserve(n_out)) || (rc = ctx->d_counts.reserve(n_out)) || (rc = ctx->d_mean.reserve(n_out)) ||
        (rc = ctx->d_m2.reserve(n_out)))
        return rc;
    if ((rc = ctx->h_occ.resize(n_out)) || (rc = ctx->h_counts.resize(n_out)) || (rc = ctx->h_mean.resize(n_out)) ||
        (rc = ctx->h_m2.resize(n_out)))
        return rc;
    ctx->phase_s[2] = now_s() - t1;
    // calibration pass (see calibrate_units): worth it once units have to share SMs
    ctx->calibrated = false;
    int32_t calib = ctx->calibration;
    if (const char *e = getenv("BAYES_CALIB_ITERS")) calib = atoi(e);  // experiment knob
    // (from two units per SM on: below that nothing waits for anything; at 8 units per SM -- one eighth of config 4 -- it is worth 9 % of the step)
    if (calib < 0) calib = (int64_t)ctx->h_units.n >= 2 * (int64_t)ctx->sm_count ? 256 : 0;
    if (cached) {
        ctx->calibrated = true;
    } else if (calib > 0 && ctx->h_units.n > 0 && ctx->trace_locus < 0) {
        t1 = now_s();
        std::vector<double> scale;
        if ((rc = calibrate_units(ctx, calib, &scale))) return rc;
        ctx->calib_scale = scale;
        ctx->calib_fp = fp;
        const double t2 = now_s();
        if ((rc = build_units(ctx, &scale))) return rc;
        const double t3 = now_s();
        if ((rc = upload_units(ctx))) return rc;
        ctx->phase_s[1] += t3 - t2;
        ctx->phase_s[2] += (t2 - t1) + (now_s() - t3);
        ctx->calibrated = true;
    }
    std::vector<UnitShape>().swap(ctx->shape_cache);
    std::vector<PackedUnit>().swap(ctx->packed_cache);
    ctx->uploaded = true;
    ctx->ran = ctx->downloaded = false;
    return BAYES_OK;
}

// Issue every class launch of the batch on its own stream, longest units first (iter_cap > 0: calibration pass).
static int issue_launches(bayes_gibbs_ctx *ctx, int32_t iter_cap) {
    int rc;
    KernelArgs args = kernel_args(ctx);
    args.iter_cap = iter_cap;
    ctx->last_launches = 0;
    CU(cudaMemsetAsync(ctx->d_unit_err.p, 0, (ctx->h_units.n + 1) * sizeof(int32_t), ctx->stream));
    const size_t n_l = ctx->launches.size();
    if (!ctx->fork_ev) CU(cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
    while (ctx->aux.size() < n_l) {
        // launch i carries the class with the i-th longest unit: the earlier the launch, the higher the stream priority,
        // so that the block scheduler places the units on the critical path before it fills the SMs with short ones
        cudaStream_t st;
        cudaEvent_t ev;
        int least = 0, greatest = 0;
        CU(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        const int prio = std::min(least, greatest + (int)ctx->aux.size());
        CU(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prio));
        CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ctx->aux.push_back(st);
        ctx->join_ev.push_back(ev);
    }
    // The block scheduler hands out the blocks of concurrent launches in the order in which the launches become
    // runnable, and the launches are sorted longest units first.  When the batch is queued behind other work all of
    // them would become runnable at the same instant (and be served in no particular order), so a gate stream releases
    // them one by one, a few microseconds apart: gate kernel i -> event i -> launch i on its own stream.
    if (n_l > 1) {
        if (!ctx->gate) CU(cudaStreamCreateWithFlags(&ctx->gate, cudaStreamNonBlocking));
        while (ctx->gate_ev.size() < n_l) {
            cudaEvent_t ev;
            CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            ctx->gate_ev.push_back(ev);
        }
        CU(cudaEventRecord(ctx->fork_ev, ctx->stream));
        CU(cudaStreamWaitEvent(ctx->gate, ctx->fork_ev, 0));
    }
    for (size_t i = 0; i < n_l; i++) {
        const Launch &l = ctx->launches[i];
        cudaStream_t st = n_l > 1 ? ctx->aux[i] : ctx->stream;
        if (n_l > 1) {
            if (i > 0) {
                if ((rc = launch_stagger(3000u, ctx->gate))) return rc;
                ctx->last_launches++;
            }
            CU(cudaEventRecord(ctx->gate_ev[i], ctx->gate));
            CU(cudaStreamWaitEvent(st, ctx->gate_ev[i], 0));
        }
        if (l.n_cta > 0 && l.wide) rc = launch_units_rows(args, l.unit_base, l.n_units, l.block, l.n_cta, l.smem, st);
        else if (ctx->built_mode == BAYES_MODE_FAST) rc = launch_units_fast(args, l.unit_base, l.n_units, l.block, l.wide, l.mat_smem, l.scr_smem, l.smem, st, std::max(1, l.n_cta));
        else rc = launch_units(args, l.unit_base, l.n_units, l.block, l.wide, l.mat_smem, l.smem, st);
        if (rc) return rc;
        if (n_l > 1) {
            CU(cudaEventRecord(ctx->join_ev[i], st));
            CU(cudaStreamWaitEvent(ctx->stream, ctx->join_ev[i], 0));
        }
        ctx->last_launches++;
    }
    return BAYES_OK;
}

// Calibration pass.  The time model only sees the SHAPE of a unit; what a conditional binomial or a CDF walk costs also
// depends on the fragment counts and on how sparse the expression vector gets, which leaves +-30 % per unit -- enough
// to start a long unit late and to leave the GPU half empty for the last quarter of the batch.  So every unit first runs
// its first `iters` iterations (about 1 % of a chain) with the GPU as full as it will be, the second half is timed on the
// device, and the units are then sized, widened and ordered again with measured / modelled as a per-unit correction.
// Measured on config 2 (tools/calib_ab.py, profiles/README.md): measured / predicted unit duration 0.52-1.07 (5-95 %)
// without, 0.73-1.27 with; step 1.89 s -> 1.78 s; 256 or 1024 iterations make no difference.  Tried on top and rejected:
// widening the units that start last (they hold more warps but do not finish sooner), 32 hardware queues
// (CUDA_DEVICE_MAX_CONNECTIONS: no change), 5.8-12 KB of shared memory per warp instead of 8 (within +-2 %).
static int calibrate_units(bayes_gibbs_ctx *ctx, int32_t iters, std::vector<double> *group_scale) {
    int rc;
    const int64_t n = (int64_t)ctx->h_units.n;
    if ((rc = issue_launches(ctx, iters))) return rc;
    std::vector<uint64_t> clock(3 * (size_t)n);
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(clock.data(), ctx->d_unit_clock.p, clock.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    group_scale->assign((size_t)n, 1.0);
    for (int64_t i = 0; i < n; i++) {
        const UnitDev &u = ctx->h_units.p[i];
        const int total = std::min(iters, u.burn + u.samples);
        const int timed = total - (total >> 1);
        const double us = 1e-3 * (double)(clock[3 * i + 1] - clock[3 * i]) / std::max(1, timed);
        const double scale = us / std::max(1e-3, ctx->unit_model_us[i]);
        (*group_scale)[ctx->unit_group[i]] = std::min(4.0, std::max(0.25, scale));
    }
    return BAYES_OK;
}

int bayes_gibbs_set_calibration(bayes_gibbs_ctx *ctx, int32_t iterations) {
    if (!ctx || iterations < -1) {
        set_error("bayes_gibbs_set_calibration: bad arguments");
        return BAYES_E_INVALID;
    }
    ctx->calibration = iterations;
    return BAYES_OK;
}

int bayes_gibbs_set_mode(bayes_gibbs_ctx *ctx, int32_t mode) {
    if (!ctx || (mode != BAYES_MODE_FAST && mode != BAYES_MODE_REPLAY)) {
        set_error("bayes_gibbs_set_mode: bad arguments");
        return BAYES_E_INVALID;
    }
    if (ctx->uploaded) {
        set_error("bayes_gibbs_set_mode: must be called before bayes_gibbs_upload");
        return BAYES_E_STATE;
    }
    ctx->mode = mode;
    return BAYES_OK;
}

int32_t bayes_gibbs_get_mode(bayes_gibbs_ctx *ctx) { return ctx ? ctx->mode : BAYES_E_INVALID; }

int bayes_gibbs_run(bayes_gibbs_ctx *ctx) {
    if (!ctx) return BAYES_E_INVALID;
    if (!ctx->uploaded) {
        set_error("bayes_gibbs_run: call bayes_gibbs_upload first");
        return BAYES_E_STATE;
    }
    int rc = use_device(ctx);
    if (rc) return rc;
    const double t0 = now_s();
    if ((rc = issue_launches(ctx, 0))) return rc;
    ctx->ran = true;
    ctx->downloaded = false;
    ctx->phase_s[3] = now_s() - t0;
    return BAYES_OK;
}

int bayes_gibbs_download(bayes_gibbs_ctx *ctx) {
    if (!ctx) return BAYES_E_INVALID;
    if (!ctx->ran) {
        set_error("bayes_gibbs_download: call bayes_gibbs_run first");
        return BAYES_E_STATE;
    }
    int rc = use_device(ctx);
    if (rc) return rc;
    const double t0 = now_s();
    const size_t n = (size_t)ctx->out_len;
    if (n) {
        CU(cudaMemcpyAsync(ctx->h_occ.p, ctx->d_occ.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(ctx->h_mean.p, ctx->d_mean.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(ctx->h_m2.p, ctx->d_m2.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(ctx->h_counts.p, ctx->d_counts.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (ctx->h_units.n)
        CU(cudaMemcpyAsync(ctx->h_unit_err.p, ctx->d_unit_err.p, ctx->h_units.n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->downloaded = true;
    ctx->phase_s[4] = now_s() - t0;
    return BAYES_OK;
}

int bayes_gibbs_sync(bayes_gibbs_ctx *ctx) {
    if (!ctx) return BAYES_E_INVALID;
    int rc = use_device(ctx);
    if (rc) return rc;
    const double t0 = now_s();
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->phase_s[5] = now_s() - t0;
    if (ctx->downloaded)
        for (size_t i = 0; i < ctx->h_units.n; i++)
            if (ctx->h_unit_err.p[i] != 0) {
                // the reference aborts here: assert(normalisation_constant >= double_underflow) (assignmentGenerator.cpp:284),
                // assert(gamma_base_sample >= double_underflow) (expressionGenerator.cpp:115)
                const int code = ctx->h_unit_err.p[i];
                set_error(code == 1   ? "a row had no candidate with a non-zero probability under the sampled expression (unit " + std::to_string(i) + ")"
                          : code == 2 ? "a Gamma draw underflowed (unit " + std::to_string(i) + ")"
                                      : "internal bounds check " + std::to_string(code) + " failed (unit " + std::to_string(i) + ")");
                return BAYES_E_NUMERIC;
            }
    return BAYES_OK;
}

// wall-clock seconds of the host phases of the last batch (see bayes_gibbs_ctx::phase_s)
int bayes_gibbs_phase_seconds(bayes_gibbs_ctx *ctx, double *out6) {
    if (!ctx || !out6) return BAYES_E_INVALID;
    for (int i = 0; i < 6; i++) out6[i] = ctx->phase_s[i];
    return BAYES_OK;
}

int bayes_gibbs_run_batch(bayes_gibbs_ctx *ctx, int32_t n_loci, const bayes_locus_desc *loci, int64_t *first_id) {
    int rc;
    if ((rc = bayes_gibbs_submit(ctx, n_loci, loci, first_id))) return rc;
    if ((rc = bayes_gibbs_upload(ctx))) return rc;
    if ((rc = bayes_gibbs_run(ctx))) return rc;
    if ((rc = bayes_gibbs_download(ctx))) return rc;
    return bayes_gibbs_sync(ctx);
}

int bayes_gibbs_fetch(bayes_gibbs_ctx *ctx, int64_t locus_id, int32_t chain, int32_t *occ, double *mean, double *m2,
                      int32_t *final_counts) {
    if (!ctx || locus_id < 0 || locus_id >= (int64_t)ctx->loci.size()) {
        set_error("bayes_gibbs_fetch: bad locus id");
        return BAYES_E_INVALID;
    }
    const PreLocus &L = ctx->loci[locus_id];
    if (chain < 0 || chain >= L.n_chains) {
        set_error("bayes_gibbs_fetch: bad chain");
        return BAYES_E_INVALID;
    }
    if (L.T == 1) {
        // gibbsSampler.cpp:57-67: `samples` identical samples of (num_fragments * 1e9) / (total_num_fragments * efflen)
        // accumulated by addSample (gibbsOutput.cpp:51-75): mean = that value, M2 = 0
        if (occ) occ[0] = L.samples;
        if (mean) mean[0] = (L.n_frag * 1e9) / (L.N_total * L.efflen[0]);
        if (m2) m2[0] = 0.0;
        if (final_counts) final_counts[0] = L.n_frag;
        return BAYES_OK;
    }
    if (!ctx->downloaded) {
        set_error("bayes_gibbs_fetch: results not downloaded (run, download, sync first)");
        return BAYES_E_STATE;
    }
    const int64_t off = ctx->out_base[locus_id] + (int64_t)chain * L.T;
    if (occ) memcpy(occ, ctx->h_occ.p + off, L.T * sizeof(int32_t));
    if (mean) memcpy(mean, ctx->h_mean.p + off, L.T * sizeof(double));
    if (m2) memcpy(m2, ctx->h_m2.p + off, L.T * sizeof(double));
    if (final_counts) memcpy(final_counts, ctx->h_counts.p + off, L.T * sizeof(int32_t));
    return BAYES_OK;
}

int bayes_gibbs_fetch_all(bayes_gibbs_ctx *ctx, int32_t *occ, double *mean, double *m2, int32_t *final_counts) {
    if (!ctx) return BAYES_E_INVALID;
    int64_t at = 0;
    for (int64_t l = 0; l < (int64_t)ctx->loci.size(); l++) {
        const PreLocus &L = ctx->loci[l];
        for (int32_t c = 0; c < L.n_chains; c++) {
            int rc = bayes_gibbs_fetch(ctx, l, c, occ ? occ + at : nullptr, mean ? mean + at : nullptr, m2 ? m2 + at : nullptr,
                                       final_counts ? final_counts + at : nullptr);
            if (rc) return rc;
            at += L.T;
        }
    }
    return BAYES_OK;
}

int bayes_gibbs_locus_info(bayes_gibbs_ctx *ctx, int64_t locus_id, double *pi, int32_t *cover_size, int32_t *rows_folded) {
    if (!ctx || locus_id < 0 || locus_id >= (int64_t)ctx->loci.size()) {
        set_error("bayes_gibbs_locus_info: bad locus id");
        return BAYES_E_INVALID;
    }
    const PreLocus &L = ctx->loci[locus_id];
    if (pi) *pi = L.prior(0).pi;
    if (cover_size) *cover_size = L.prior(0).cover;
    if (rows_folded) *rows_folded = L.folded;
    return BAYES_OK;
}

int bayes_gibbs_chain_info(bayes_gibbs_ctx *ctx, int64_t locus_id, int32_t chain, double *pi, int32_t *cover_size) {
    if (!ctx || locus_id < 0 || locus_id >= (int64_t)ctx->loci.size() || chain < 0 || chain >= ctx->loci[locus_id].n_chains) {
        set_error("bayes_gibbs_chain_info: bad locus id / chain");
        return BAYES_E_INVALID;
    }
    const SimplexPrior &sp = ctx->loci[locus_id].prior(chain);
    if (pi) *pi = sp.pi;
    if (cover_size) *cover_size = sp.cover;
    return BAYES_OK;
}

// Diagnostics of the last run: per unit {wide, block, n_slots, n_slabs, n_cells, iterations, T of slot 0, mat_smem} and
// {start ns, end ns, SM id}.  Returns the number of units (call with NULL pointers to size the buffers).
int64_t bayes_gibbs_unit_stats(bayes_gibbs_ctx *ctx, int32_t *shape8, uint64_t *clock3, double *predicted_s) {
    if (!ctx) return BAYES_E_INVALID;
    const int64_t n = (int64_t)ctx->h_units.n;
    if (!shape8 && !clock3 && !predicted_s) return n;
    if (!ctx->ran) {
        set_error("bayes_gibbs_unit_stats: no run yet");
        return BAYES_E_STATE;
    }
    if (use_device(ctx)) return BAYES_E_CUDA;
    if (cuda_rc(cudaStreamSynchronize(ctx->stream), "sync")) return BAYES_E_CUDA;
    if (shape8)
        for (int64_t i = 0; i < n; i++) {
            const UnitDev &u = ctx->h_units.p[i];
            int32_t *o = shape8 + 8 * i;
            o[0] = u.wide; o[1] = u.block; o[2] = u.n_slots; o[3] = u.n_slabs; o[4] = u.n_cells; o[5] = u.burn + u.samples; o[6] = u.T[0];
            o[7] = u.mat_smem | (u.slab_h << 8) | (((u.smem_bytes + 1023) >> 10) << 16);  // + shared memory of the unit, KB
        }
    if (predicted_s)
        for (int64_t i = 0; i < n; i++) {
            predicted_s[5 * i] = i < (int64_t)ctx->unit_pred.size() ? ctx->unit_pred[i] : 0.0;
            for (int k = 0; k < 4; k++) predicted_s[5 * i + 1 + k] = 4 * i + k < (int64_t)ctx->unit_feat.size() ? ctx->unit_feat[4 * i + k] : 0.0;
        }
    if (clock3 && n && cuda_rc(cudaMemcpy(clock3, ctx->d_unit_clock.p, 3 * n * sizeof(uint64_t), cudaMemcpyDeviceToHost), "unit clock D2H"))
        return BAYES_E_CUDA;
    return n;
}

int64_t bayes_gibbs_num_loci(bayes_gibbs_ctx *ctx) { return ctx ? (int64_t)ctx->loci.size() : 0; }
int64_t bayes_gibbs_num_jobs(bayes_gibbs_ctx *ctx) {
    if (!ctx) return 0;
    int64_t n = 0;
    for (const PreLocus &L : ctx->loci) n += L.n_chains;
    return n;
}
int32_t bayes_gibbs_last_launches(bayes_gibbs_ctx *ctx) { return ctx ? ctx->last_launches : 0; }

int bayes_gibbs_set_trace(bayes_gibbs_ctx *ctx, int64_t locus_id, int32_t chain, int32_t trace_iters) {
    if (!ctx || locus_id < 0 || locus_id >= (int64_t)ctx->loci.size() || trace_iters < 1) {
        set_error("bayes_gibbs_set_trace: bad arguments");
        return BAYES_E_INVALID;
    }
    if (ctx->uploaded) {
        set_error("bayes_gibbs_set_trace: must be called before bayes_gibbs_upload");
        return BAYES_E_STATE;
    }
    const PreLocus &L = ctx->loci[locus_id];
    if (L.T == 1 || chain < 0 || chain >= L.n_chains) {
        set_error("bayes_gibbs_set_trace: locus has no device job for that chain");
        return BAYES_E_INVALID;
    }
    ctx->trace_locus = locus_id;
    ctx->trace_chain = chain;
    ctx->trace_iters = trace_iters;
    ctx->trace_T = L.T;
    return BAYES_OK;
}

int bayes_gibbs_fetch_trace(bayes_gibbs_ctx *ctx, double *theta_trace, int32_t *counts_trace, int32_t *b_trace, int32_t *init_counts) {
    if (!ctx || ctx->trace_locus < 0 || !ctx->ran) {
        set_error("bayes_gibbs_fetch_trace: no traced run");
        return BAYES_E_STATE;
    }
    int rc = use_device(ctx);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    const size_t n = (size_t)ctx->trace_iters * ctx->trace_T;
    if (theta_trace) CU(cudaMemcpy(theta_trace, ctx->d_tr_theta.p, n * sizeof(double), cudaMemcpyDeviceToHost));
    if (counts_trace) CU(cudaMemcpy(counts_trace, ctx->d_tr_counts.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (b_trace) CU(cudaMemcpy(b_trace, ctx->d_tr_b.p, ctx->trace_iters * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (init_counts) CU(cudaMemcpy(init_counts, ctx->d_tr_init.p, ctx->trace_T * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return BAYES_OK;
}

// ---- single steps --------------------------------------------------------------------------------------------------
static int step_assignment(int device, const bayes_locus_desc *locus, const double *theta, uint64_t key, uint32_t iteration,
                           int32_t *counts_out, bool init, int mode = BAYES_MODE_REPLAY) {
    if (mode != BAYES_MODE_FAST && mode != BAYES_MODE_REPLAY) {
        set_error("bayes_step_assignment: bad mode");
        return BAYES_E_INVALID;
    }
    const bool fast = mode == BAYES_MODE_FAST;
    if (!locus || !counts_out || (!init && !theta)) {
        set_error("bayes_step_*_assignment: null argument");
        return BAYES_E_INVALID;
    }
    bayes_gibbs_ctx *ctx = nullptr;
    int rc = bayes_gibbs_create(device, &ctx);
    if (rc) return rc;
    auto fail = [&](int code) { bayes_gibbs_destroy(ctx); return code; };
    bayes_locus_desc d = *locus;
    if (d.pi < 0) d.pi = 0.5;  // the cover is not part of this step
    d.n_chains = 1;
    if (d.burn_in < 0) d.burn_in = 0;
    if (d.samples < 1) d.samples = 1;
    d.seed = key;  // chain 0 uses the seed itself as its Philox key
    PreLocus P;
    std::string err;
    if ((rc = prepare_locus(d, &P, &err))) {
        set_error("bayes_step_*_assignment: " + err);
        return fail(rc);
    }
    const int T = d.T;
    PackedUnit U;
    // replay mode and wide loci: wide layout (columns stay candidate indices); fast mode with T <= 32: the tiles of a bundle
    const bool tiles = fast && T <= 32, rows = fast && T > 32;
    if (rows) pack_unit_rows(std::vector<SlotRef>(1, SlotRef{&P, 0, 0}), 1, &U);
    else pack_unit(std::vector<SlotRef>(1, SlotRef{&P, 0, 0}), !tiles, 32, fast, &U);
    U.d.mat_smem = 0;  // the step kernels read the matrix (and the fast-mode scratch) from global memory
    U.d.scr_smem = 0;
    U.d.pool_smem = 0;
    ctx->scr_len = ctx->q_len = ctx->gpool_len = 0;
    if ((rc = append_unit(ctx, U))) return fail(rc);
    if ((rc = upload_units(ctx))) return fail(rc);
    cudaStream_t st = ctx->stream;
    if ((rc = ctx->d_mean.reserve(T)) || (rc = ctx->d_occ.reserve(T))) return fail(rc);
    if (!init && (rc = cuda_rc(cudaMemcpyAsync(ctx->d_mean.p, theta, T * sizeof(double), cudaMemcpyHostToDevice, st), "theta H2D")))
        return fail(rc);
    const KernelArgs args = kernel_args(ctx);
    const size_t smem = 12 * (size_t)T + 8 * (((size_t)T + 63) / 64) + 48 + 4 * kFastCtlWords;
    if ((rc = cuda_rc(cudaMemsetAsync(ctx->d_unit_err.p, 0, sizeof(int32_t), st), "memset"))) return fail(rc);
    if (rows) rc = launch_step_assignment_rows(args, init, ctx->d_mean.p, iteration, ctx->d_occ.p, 12 * (size_t)T + 8 * (((size_t)T + 63) / 64) + 4 * kRowsCtlWords + 12 * (size_t)U.d.chain_max + 16, st);
    else if (fast) rc = launch_step_assignment_fast(args, init, tiles, ctx->d_mean.p, iteration, ctx->d_occ.p, smem, st);
    else rc = launch_step_assignment(args, init, ctx->d_mean.p, iteration, ctx->d_occ.p, smem, st);
    if (rc) return fail(rc);
    if ((rc = cuda_rc(cudaMemcpyAsync(counts_out, ctx->d_occ.p, T * sizeof(int32_t), cudaMemcpyDeviceToHost, st), "counts D2H")))
        return fail(rc);
    int32_t unit_err = 0;
    if ((rc = cuda_rc(cudaMemcpyAsync(&unit_err, ctx->d_unit_err.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st), "err D2H"))) return fail(rc);
    if ((rc = cuda_rc(cudaStreamSynchronize(st), "step sync"))) return fail(rc);
    bayes_gibbs_destroy(ctx);
    if (unit_err != 0) {
        set_error("bayes_step_assignment: a row had no candidate with a non-zero probability (reference: assert, assignmentGenerator.cpp:284)");
        return BAYES_E_NUMERIC;
    }
    return BAYES_OK;
}

int bayes_step_generate_assignment(int device, const bayes_locus_desc *locus, const double *theta, uint64_t key,
                                   uint32_t iteration, int32_t *counts_out) {
    return step_assignment(device, locus, theta, key, iteration, counts_out, false);
}

int bayes_step_init_assignment(int device, const bayes_locus_desc *locus, uint64_t key, int32_t *counts_out) {
    return step_assignment(device, locus, nullptr, key, 0, counts_out, true);
}

int bayes_step_assignment(int device, const bayes_locus_desc *locus, const double *theta, uint64_t key, uint32_t iteration, int32_t mode,
                          int32_t *counts_out) {
    return step_assignment(device, locus, theta, key, iteration, counts_out, theta == nullptr, mode);
}

int bayes_step_generate_expression(int device, int32_t T, const int32_t *counts, int32_t b, double gamma, uint64_t key,
                                   uint32_t iteration, double *theta_out) {
    if (T < 1 || T > kMaxT || !counts || !theta_out || !(gamma > 0)) {
        set_error("bayes_step_generate_expression: bad arguments");
        return BAYES_E_INVALID;
    }
    int n_plus = 0;
    for (int t = 0; t < T; t++) n_plus += counts[t] > 0;
    if (n_plus < 1 || b < n_plus || b > T) {
        set_error("bayes_step_generate_expression: need 1 <= |S+| <= b <= T");
        return BAYES_E_INVALID;
    }
    bayes_gibbs_ctx *ctx = nullptr;
    int rc = bayes_gibbs_create(device, &ctx);
    if (rc) return rc;
    auto fail = [&](int code) { bayes_gibbs_destroy(ctx); return code; };
    cudaStream_t st = ctx->stream;
    if ((rc = ctx->d_occ.reserve(T)) || (rc = ctx->d_mean.reserve(T))) return fail(rc);
    if ((rc = cuda_rc(cudaMemcpyAsync(ctx->d_occ.p, counts, T * sizeof(int32_t), cudaMemcpyHostToDevice, st), "counts H2D"))) return fail(rc);
    if ((rc = launch_step_expression(T, ctx->d_occ.p, b, gamma, key, iteration, ctx->d_mean.p, st))) return fail(rc);
    if ((rc = cuda_rc(cudaMemcpyAsync(theta_out, ctx->d_mean.p, T * sizeof(double), cudaMemcpyDeviceToHost, st), "theta D2H"))) return fail(rc);
    if ((rc = cuda_rc(cudaStreamSynchronize(st), "step sync"))) return fail(rc);
    bayes_gibbs_destroy(ctx);
    return BAYES_OK;
}

int bayes_step_variates(int device, int kind, double a, double p, uint64_t key, int32_t n, double *out) {
    if (n < 1 || !out || kind < 0 || kind > 2) {
        set_error("bayes_step_variates: bad arguments");
        return BAYES_E_INVALID;
    }
    bayes_gibbs_ctx *ctx = nullptr;
    int rc = bayes_gibbs_create(device, &ctx);
    if (rc) return rc;
    auto fail = [&](int code) { bayes_gibbs_destroy(ctx); return code; };
    if ((rc = ctx->d_mean.reserve(n))) return fail(rc);
    if ((rc = launch_step_variates(kind, a, p, key, n, ctx->d_mean.p, ctx->stream))) return fail(rc);
    if ((rc = cuda_rc(cudaMemcpyAsync(out, ctx->d_mean.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream), "variates D2H"))) return fail(rc);
    if ((rc = cuda_rc(cudaStreamSynchronize(ctx->stream), "step sync"))) return fail(rc);
    bayes_gibbs_destroy(ctx);
    return BAYES_OK;
}

// Development aid (host only): pack the given loci (all with the same number of candidate lanes) as ONE bundle and report
// the shared-memory layout of the unit: out[0..] = total, matrix cells, padded rows, chain cells, qcap, pool blocks, table entries
int bayes_debug_unit_layout(int32_t n_loci, const bayes_locus_desc *loci, int32_t fast, int32_t *out8) {
    std::vector<PreLocus> pre(n_loci);
    std::string err;
    for (int i = 0; i < n_loci; i++)
        if (int rc = prepare_locus(loci[i], &pre[i], &err)) { set_error(err); return rc; }
    std::vector<SlotRef> slots;
    for (int i = 0; i < n_loci; i++) slots.push_back(SlotRef{&pre[i], 0, 0});
    PackedUnit U;
    if (fast && pre[0].T > 32) pack_unit_rows(slots, 1, &U);
    else pack_unit(slots, pre[0].T > 32, 32, fast != 0, &U);
    out8[0] = U.d.smem_bytes; out8[1] = U.d.n_cells; out8[2] = U.d.n_slabs * U.d.slab_h; out8[3] = U.d.n_chain_cells;
    out8[4] = U.d.qcap; out8[5] = U.d.pool_blocks; out8[6] = U.d.tab_len; out8[7] = U.d.mat_smem | (U.d.tab_smem << 1) | (U.d.pool_smem << 2);
    return BAYES_OK;
}

}  // extern "C"
