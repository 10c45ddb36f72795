// Fast (production) sampling mode of the persistent Gibbs kernels for sm_100a.
//
// Same model, same conditional distributions and the same E-step as the replay-mode kernels (gibbs_kernel.cu), but the
// A-step no longer mirrors the ORDER in which the reference draws (assignmentGenerator.cpp:263-373); SURVEY.md section 7
// allows any exact multinomial per row outside replay.  What changes, and why:
//
//  * rows with >= 20 fragments: the reference's chain of conditional binomials (:336-365) is k strictly dependent draws
//    for a row with k cells in play -- it set both the critical path of config 2 and the iteration time of config 3.
//    Here a row's cells in play are compacted into a per-iteration scratch list of running sums, and the multinomial
//    is drawn by a balanced binary TREE of binomials over that list (node (a, len): left half gets
//    Binomial(m, mass(left) / mass(node))).  All nodes of one tree level, over all rows of the unit, are independent:
//    they are queued in shared memory and processed level by level with one lane per node, densely packed,
//    whichever row they came from.  Depth log2 k instead of k, and full warps instead of lanes waiting for the
//    deepest row of their slab.
//  * rows with < 20 fragments (:289-334): the 32-bit words they locate in their CDF come from a per-iteration POOL of
//    Philox blocks of the slot -- fragments of all such rows back to back, four to a block, instead of at least one
//    block per row (the median row has ONE fragment) -- generated with full lanes, by the warps that would otherwise
//    idle during the one-warp E-step; one reciprocal per row and un-normalised running sums instead of a division
//    per cell.
//  * bundles (loci with up to 32 candidates) drop the lane-per-row SELL slabs for lane-per-cell TILES: 32 consecutive cells,
//    rows contiguous and never straddling a tile.  One pass per tile with all lanes busy whatever the row lengths:
//    p = P_ij theta_j, a segmented warp-shuffle inclusive scan gives every cell its running sum and every row its
//    total, ballots name the cells in play, and each cell takes the pool words that fall between its neighbour's
//    threshold and its own.  No padding to the widest row of a slab (SELL-32 doubled the footprint of small units),
//    no lanes waiting for the longest row.  The scan fixes the order of the floating-point additions of a row sum;
//    the oracle sums in the same order.
//  * per-candidate state of a bundle (Welford accumulators, constants) lives in shared memory instead of in registers
//    of warp 0 that every other thread had to allocate as well.
// The CPU oracle restates this mode draw for draw (oracle/bayes_oracle.c, assignment_fast); tests compare the two.
// No tensor cores: nothing here is a dense contraction.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <float.h>
#include <stdio.h>
#include <stdlib.h>

#include "internal.h"
#include "layout.h"
#include "rng.cuh"
#include "gibbs_common.cuh"

#ifndef BAYES_FAST_MAXNREG
#define BAYES_FAST_MAXNREG 64
#endif

namespace cg = cooperative_groups;

namespace bayes {

namespace {

constexpr uint32_t kFull = 0xffffffffu;
// ctl words
constexpr int kCtlSlab = 0, kCtlQ0 = 1, kCtlErr = 4;

#ifdef BAYES_DEBUG_BOUNDS
// debug build: every index derived from packed data is checked; a violation lands in the unit's error word
#define BAYES_CHECK(cond, code) do { if (!(cond)) V.ctl[kCtlErr] = (code); } while (0)
#else
#define BAYES_CHECK(cond, code) do { } while (0)
#endif

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Pool of the iteration: block b of the unit = block (desc & 0xffffff) of the pool of slot (desc >> 24).
__device__ __forceinline__ void fast_pool_gen(uint32_t *pool, const uint32_t *pool_desc, int pool_blocks, const uint64_t *skey, uint32_t iteration,
                                              bool init, int first_thread, int n_threads) {
    const uint32_t ph = init ? PH_INIT_POOL : PH_POOL;
    for (int b = (int)threadIdx.x - first_thread; b < pool_blocks; b += n_threads) {
        const uint32_t desc = pool_desc[b], blk = desc & 0xffffffu;
        const u32x4 w = make_site(skey[desc >> 24], ph, iteration, blk >> 12, 0).block(blk & 4095u);
        reinterpret_cast<uint4 *>(pool)[b] = make_uint4(w.x, w.y, w.z, w.w);
    }
}

// Where the words of one Philox block (up to `left` of them) land in a row's CDF: number of words below the threshold
__device__ __forceinline__ int words_below(const u32x4 &w, int left, uint32_t t) {
    return (int)(w.x <= t) + ((left > 1) & (int)(w.y <= t)) + ((left > 2) & (int)(w.z <= t)) + ((left > 3) & (int)(w.w <= t));
}

// ------------------------------------------------------------------------------------------------------------------
// Bundles: lane-per-cell tiles.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lanemask_le() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_le;" : "=r"(m));
    return m;
}

constexpr uint32_t kPadOrd = 0x3FFFFFFu;  // chain tiles: ordinal of a padding lane

struct TileView {
    const double *val;       // n_tiles * 32
    const uint32_t *meta;    // idem: candidate lane | head << 5 | (fragments << 6 | pool word << 11  or  chain ordinal << 6)
    const uint32_t *ctab;    // 3 words per chain row: fragments, slot << 27 | row, first cell
    int n_tiles, n_chain_tiles;
    double *scr_cum;         // n_chain_tiles * 32: running sums of the cells in play of every chain row, compacted to the row's first cells
    uint16_t *scr_col;       // their candidate lanes
    uint32_t *queue;         // two buffers of 2 * qcap words: fragments | ordinal | chunk << 15 | first cell << 22 | (cells - 1) << 27
    int qcap;
    const uint32_t *pool;    // this iteration's uniform words (shared memory) or nullptr
    const int32_t *pool_first;
    int sh;                  // log2 of the lanes per slot
    int max_row;             // lanes per slot = most cells a row can have
    int *ctl;
};

__device__ __forceinline__ void tile_push(const TileView &V, int cnt, int buf, bool want, uint32_t n, uint32_t w1) {
    const uint32_t m = __ballot_sync(kFull, want);
    if (m == 0u) return;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(&V.ctl[cnt], __popc(m));
    base = __shfl_sync(kFull, base, leader);
    if (want) {
        const int idx = base + __popc(m & lanemask_lt());
        BAYES_CHECK(idx < V.qcap, 111);
        uint32_t *q = V.queue + (size_t)buf * 2 * V.qcap;
        q[idx] = n;
        q[V.qcap + idx] = w1;
    }
}

__device__ __forceinline__ u32x4 tile_pool_words4(const TileView &V, const uint64_t *skey, uint32_t slot, uint32_t ph_pool, uint32_t iteration, int w) {
    u32x4 o;
    if (V.pool) {
        o.x = V.pool[w]; o.y = V.pool[w + 1]; o.z = V.pool[w + 2]; o.w = V.pool[w + 3];
        return o;
    }
    const int wl = w - 4 * V.pool_first[slot];
    const uint32_t b0 = (uint32_t)wl >> 2, off = (uint32_t)wl & 3u;
    const uint64_t key = skey[slot];
    const u32x4 A = make_site(key, ph_pool, iteration, b0 >> 12, 0).block(b0 & 4095u);
    if (off == 0u) return A;
    const u32x4 B = make_site(key, ph_pool, iteration, (b0 + 1u) >> 12, 0).block((b0 + 1u) & 4095u);
    const uint32_t a[8] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w};
    o.x = a[off]; o.y = a[off + 1]; o.z = a[off + 2]; o.w = a[off + 3];
    return o;
}

// A-step, first part: every tile in one pass.  Chain tiles compact the cells in play of their rows into the scratch
// lists and queue the roots of the splitting trees; the other tiles locate their rows' pool words right away.
// (a unit spread over a cluster of n_cta CTAs: tile t goes to CTA t mod n_cta, so the chain tiles -- the first ones -- and with them the
// splitting trees are dealt over all CTAs)
__device__ __forceinline__ void tile_rows(const TileView &V, const double *theta, int32_t *counts, const uint64_t *skey, uint32_t iteration,
                                          const bool init, const int rank = 0, const int n_cta = 1) {
    const int lane = threadIdx.x & 31, warp = (int)(threadIdx.x >> 5) * n_cta + rank, n_warps = (int)(blockDim.x >> 5) * n_cta;
    const uint32_t ph_pool = init ? PH_INIT_POOL : PH_POOL;
    const uint32_t le = lanemask_le(), lt = lanemask_lt();
    // Tiles cost about the same, so the warps take them round robin; the next tile of the warp is loaded while this one is
    // processed (the tiles of a streaming unit come from L2).
    int ti = warp;
    double v_next = 0.0;
    uint32_t m_next = 0u;
    if (ti < V.n_tiles) {
        v_next = V.val[ti * 32 + lane];
        m_next = V.meta[ti * 32 + lane];
    }
    for (; ti < V.n_tiles; ti += n_warps) {
        const uint32_t meta = m_next;
        const double v = v_next;
        if (ti + n_warps < V.n_tiles) {
            v_next = V.val[(ti + n_warps) * 32 + lane];
            m_next = V.meta[(ti + n_warps) * 32 + lane];
        }
        const int c = (int)(meta & 31u);
        const double p = v * theta[c];
        // segments = rows: first lane of mine, last lane of mine
        const uint32_t hm = __ballot_sync(kFull, (meta >> kMetaHeadShift) & 1u);
        const int seg_start = 31 - __clz(hm & le);
        const uint32_t above = hm & ~le;
        const int seg_end = above ? __ffs(above) - 2 : 31;
        const uint32_t segmask = (seg_end == 31 ? kFull : ((2u << seg_end) - 1u)) & ~((1u << seg_start) - 1u);
        // inclusive scan within the row (the order of these additions is part of the mode's definition: the oracle sums alike)
        // (a row has at most as many cells as its slot has candidate lanes: the steps beyond that add nothing)
        double cum = p;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            if (d >= V.max_row) break;
            const double up = __shfl_up_sync(kFull, cum, d);
            if (lane - d >= seg_start) cum += up;
        }
        const bool inplay = p > 0.0;
        const uint32_t ip = __ballot_sync(kFull, inplay) & segmask;  // the cells in play of my row
        if (ti < V.n_chain_tiles) {
            const uint32_t ord = meta >> kMetaOrdShift;
            const int k = __popc(ip), k2 = __popc(ip & lt);
            const int base = ti * 32 + seg_start;
            if (inplay) {
                V.scr_cum[base + k2] = cum;
                V.scr_col[base + k2] = (uint16_t)c;
            }
            const bool head = lane == seg_start && ord != kPadOrd;
            uint32_t n = 0;
            if ((head || (inplay && k == 1)) && ord != kPadOrd) n = V.ctab[3 * ord];
            if (inplay && k == 1) atomicAdd(&counts[c], (int)n);
            if (head && k == 0) V.ctl[kCtlErr] = 1;  // reference: assert(normalisation_constant >= double_underflow) (:284)
            // the row's fragments enter the tree as independent chunks of at most 64: every binomial stays in the inversion regime
            // (a row above 64 kTreeChunks = 8 192 fragments enters whole: its binomials are BTRD's either way, and 128 larger chunks of
            // it were measured slower.  Below that every binomial stays in the inversion regime: a single row of 2 500 fragments
            // entering whole put one BTRD binomial per level on the critical path of its locus, 20 us of a 28 us iteration)
            const int m = head && k >= 2 ? (((int)n + 63) >> 6 <= kTreeChunks ? ((int)n + 63) >> 6 : 1) : 0;
            const int m_max = __reduce_max_sync(kFull, m);
            if (m_max <= 4) {
                // the usual tile: a few rows of a few chunks each, the heads push side by side
                for (int ch = 0; ch < m_max; ch++)
                    tile_push(V, kCtlQ0, 0, ch < m, n / (uint32_t)max(m, 1) + (uint32_t)(ch < (int)(n % (uint32_t)max(m, 1))),
                              ord | ((uint32_t)ch << 15) | ((uint32_t)(k - 1) << 27));
            } else {
                // a row of many chunks: all lanes push its roots, 32 chunks at a time
                uint32_t heads = __ballot_sync(kFull, m > 0);
                while (heads) {
                    const int src = __ffs(heads) - 1;
                    heads &= heads - 1u;
                    const int mb = __shfl_sync(kFull, m, src);
                    const uint32_t nb = __shfl_sync(kFull, n, src);
                    const uint32_t wb = __shfl_sync(kFull, ord | ((uint32_t)(k - 1) << 27), src);
                    for (int c0 = 0; c0 < mb; c0 += 32) {
                        const int ch = c0 + lane;
                        tile_push(V, kCtlQ0, 0, ch < mb, nb / (uint32_t)mb + (uint32_t)(ch < (int)(nb % (uint32_t)mb)), wb | ((uint32_t)ch << 15));
                    }
                }
            }
        } else {
            const int n = (int)((meta >> kMetaNShift) & 31u);  // 0: padding lane
            const int w0 = (int)(meta >> kMetaWoffShift);
            const double Z = __shfl_sync(kFull, cum, seg_end);
            if (n > 0 && ip == 0u && lane == seg_start) V.ctl[kCtlErr] = 1;
            const bool is_last = inplay && (ip & ~le) == 0u;  // the last cell in play takes what the others left
            const uint32_t lower = ip & lt;                   // cells in play before mine
            const int src = lower ? 31 - __clz(lower) : lane;
            const double rZ = __drcp_rn(Z);
            // word * 2^-32 < cum / Z  <=>  word <= ceil(cum (1/Z) 2^32 - 1); the conversion saturates
            const uint32_t t = __double2uint_ru(__fma_rn(cum * rZ, 4294967296.0, -1.0));
            const int n_max = __reduce_max_sync(kFull, n);
            for (int j0 = 0; j0 < n_max; j0 += 4) {
                const int left = n - j0 < 4 ? n - j0 : 4;
                int below = 0;
                if (inplay && left > 0) {
                    if (is_last) {
                        below = left;
                    } else {
                        const u32x4 w = tile_pool_words4(V, skey, (uint32_t)c >> V.sh, ph_pool, iteration, w0 + j0);
                        below = words_below(w, left, t);
                    }
                }
                const int prev = __shfl_sync(kFull, below, src);
                const int delta = lower ? below - prev : below;
                if (delta > 0) atomicAdd(&counts[c], delta);
            }
        }
    }
}

// A-step, second part: the splitting trees, level by level.  One lane per node of the level, whichever row it belongs to; children
// with two or more cells are queued for the next level, single cells receive their fragments.  Three rotating counters (this level,
// next level, the one being reset) and two queue buffers need one block barrier per level.
__device__ __forceinline__ void tile_tree_levels(const TileView &V, int32_t *counts, const uint64_t *skey, uint32_t iteration, const bool init,
                                                 long long *n_levels = nullptr) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ph_tree = init ? PH_INIT_TREE : PH_TREE;
    const int qcap = V.qcap;
    int cur = kCtlQ0, nxt = kCtlQ0 + 1, old = kCtlQ0 + 2, buf = 0;
    for (;;) {
        const int Q = V.ctl[cur];
        if (Q == 0) break;
        if (threadIdx.x == 0) V.ctl[old] = 0;
        const uint32_t *q = V.queue + (size_t)buf * 2 * qcap;
        if (n_levels) { n_levels[0] += 1; n_levels[1] += Q; }
        for (int i0 = warp * 32; i0 < Q; i0 += (int)blockDim.x) {
            const int i = i0 + lane;
            bool push_l = false, push_r = false;
            uint32_t n_l = 0, n_r = 0, w_l = 0, w_r = 0;
            if (i < Q) {
                const uint32_t n = q[i], w1 = q[qcap + i];
                const uint32_t ord = w1 & 0x7FFFu, och = w1 & 0x3FFFFFu;  // och: ordinal and chunk, inherited by the children
                const int a = (int)((w1 >> 22) & 31u), len = (int)(w1 >> 27) + 1, half = len >> 1, len_r = len - half;
                const int base = (int)V.ctab[3 * ord + 2];
                BAYES_CHECK(base + a + len - 1 < 32 * V.n_chain_tiles, 116);
                const double *sc = V.scr_cum + base;
                const double Pa = a ? sc[a - 1] : 0.0, Pm = sc[a + half - 1], Pb = sc[a + len - 1];
                const double p_left = (Pm - Pa) / (Pb - Pa);
                const uint32_t info = V.ctab[3 * ord + 1];
                const uint32_t node = (uint32_t)a | ((uint32_t)(len - 1) << 12);
                const Site site = make_site(skey[info >> kRowSlotShift], ph_tree | (node << 8), iteration, info & ((1u << kRowSlotShift) - 1u),
                                            (w1 >> 15) & 127u);
                n_l = (uint32_t)binomial_variate(site, (int)n, p_left);
                n_r = n - n_l;
                if (n_l) {
                    if (half == 1) atomicAdd(&counts[V.scr_col[base + a]], (int)n_l);
                    else { push_l = true; w_l = och | ((uint32_t)a << 22) | ((uint32_t)(half - 1) << 27); }
                }
                if (n_r) {
                    if (len_r == 1) atomicAdd(&counts[V.scr_col[base + a + half]], (int)n_r);
                    else { push_r = true; w_r = och | ((uint32_t)(a + half) << 22) | ((uint32_t)(len_r - 1) << 27); }
                }
            }
            tile_push(V, nxt, buf ^ 1, push_l, n_l, w_l);
            tile_push(V, nxt, buf ^ 1, push_r, n_r, w_r);
        }
        __syncthreads();
        const int t = old;
        old = cur;
        cur = nxt;
        nxt = t;
        buf ^= 1;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// E-step of a bundle (warp 0): lane = slot * Tpad + t.  ExpressionGenerator::generateExpression
// (expressionGenerator.cpp:77-132) for every slot of the bundle at once, + getFPKM / addSample when sampling
// (valueContainer.cpp:226-239, gibbsOutput.cpp:51-75).  Same draws as bundle_e_step of the replay kernels; the
// per-lane state is read from / written to shared memory.
// ------------------------------------------------------------------------------------------------------------------
struct FastBundleE {
    double *theta, *den, *mean, *m2;
    int32_t *counts, *base, *occ;
    const uint64_t *skey;
    const double *sgamma;
    const int32_t *sT, *sNf;
    const double *tab_val;
    const int32_t *tab_row;
    int *ctl;
    int Tpad, sh, n_slots;
};

__device__ __forceinline__ int fast_bundle_e_step(const FastBundleE &E, uint32_t iteration, bool sampling) {
    const int lane = threadIdx.x & 31;
    const int Tpad = E.Tpad, g = lane >> E.sh, t = lane & (Tpad - 1);
    const bool slot_ok = g < E.n_slots;
    const int Tg = slot_ok ? E.sT[g] : 0;
    const bool valid = t < Tg;
    const uint64_t key = slot_ok ? E.skey[g] : 0ull;
    const double gamma = slot_ok ? E.sgamma[g] : 1.0;
    const uint32_t segmask = Tpad == 32 ? kFull : (((1u << Tpad) - 1u) << (g * Tpad));
    const uint32_t validmask = Tg == 32 ? kFull : (((1u << Tg) - 1u) << (g * Tpad));
    const int cnt = E.counts[lane];
    const uint32_t pm = __ballot_sync(kFull, valid && cnt > 0) & segmask;
    const int n_plus = __popc(pm);
    int b = 0, n_pick = 0;
    if (Tg > 0) {
        b = sample_simplex(E.tab_val, E.tab_row + g * (Tpad + 1), n_plus, key, iteration);
        n_pick = b - n_plus;
    }
    // expressionGenerator.cpp:104-127: the uniform_int draws are independent (site = pick ordinal): lane t of the slot makes
    // the draw of pick t, only the selection is sequential
    uint32_t my_pick = 0u;
    if (valid && t < n_pick) my_pick = uniform_int(make_site(key, PH_NULLPICK, iteration, (uint32_t)t, 0), (uint32_t)(Tg - n_plus - t));
    uint32_t sel = 0u;
    const int max_pick = __reduce_max_sync(kFull, n_pick);
    const int lane0 = lane - t;
    for (int k = 0; k < max_pick; k++) {
        const uint32_t s = __shfl_sync(kFull, my_pick, (lane0 + k) & 31);
        if (k < n_pick) {
            const uint32_t nullm = validmask & ~(pm | sel);
            sel |= 1u << __fns(nullm, 0, (int)s + 1);
        }
    }
    const bool active = ((pm | sel) >> lane) & 1u;
    double gv = 0.0;
    if (active) {
        gv = gamma_variate(make_site(key, PH_GAMMA, iteration, (uint32_t)t, 0), (double)cnt + gamma);
        if (!(gv >= DBL_MIN)) E.ctl[kCtlErr] = 2;  // reference: assert(gamma_base_sample >= double_underflow) (expressionGenerator.cpp:115)
    }
    double norm = gv;
    for (int o = Tpad >> 1; o > 0; o >>= 1) norm += __shfl_xor_sync(kFull, norm, o);
    const double th = active ? gv / norm : 0.0;  // valueContainer.cpp:216-224
    E.theta[lane] = th;
    E.counts[lane] = E.base[lane];  // fresh CountValueContainer for the next assignment (:266) + folded rows
    if (sampling && active) {
        const double fpkm = (th * (double)E.sNf[g] * 1e9) / E.den[lane];  // valueContainer.cpp:235
        const int o = E.occ[lane] + 1;                                     // gibbsOutput.cpp:60
        E.occ[lane] = o;
        if (o == 1) {
            E.mean[lane] = fpkm;
        } else {
            const double mean = E.mean[lane], delta = fpkm - mean, mu = mean + delta / o;
            E.mean[lane] = mu;
            E.m2[lane] += delta * (fpkm - mu);
        }
    }
    return b;
}

template <bool MAT_SMEM, bool SCR_SMEM>
__device__ __forceinline__ TileView tile_stage(const KernelArgs &A, const UnitDev &U, const TileLayout &lay, unsigned char *smem, const int rank) {
    TileView V;
    V.n_tiles = U.n_tiles;
    V.n_chain_tiles = U.n_chain_tiles;
    V.qcap = U.qcap;
    V.ctl = (int *)(smem + lay.ctl);
    V.pool_first = U.pool_first;
    {
        int sh = 1;
        while ((1 << sh) < U.Tpad) sh++;
        V.sh = sh;
        V.max_row = U.Tpad;
    }
    if (MAT_SMEM) {
        double *sv = (double *)(smem + lay.val);
        uint32_t *sm = (uint32_t *)(smem + lay.meta);
        block_copy(sv, A.val + U.cell_off, U.n_cells);
        block_copy(sm, A.cell_meta + U.meta_off, U.n_cells);
        V.val = sv; V.meta = sm;
    } else {
        V.val = A.val + U.cell_off;
        V.meta = A.cell_meta + U.meta_off;
    }
    if (SCR_SMEM) {
        uint32_t *st = (uint32_t *)(smem + lay.ctab);
        block_copy(st, A.chain_tab + U.ctab_off, 3 * U.n_chain_rows);
        V.ctab = st;
        V.scr_cum = (double *)(smem + lay.scr_cum);
        V.scr_col = (uint16_t *)(smem + lay.scr_col);
        V.queue = (uint32_t *)(smem + lay.queue);
    } else {
        V.ctab = A.chain_tab + U.ctab_off;
        V.scr_cum = A.scr_cum + U.scr_off;  // (a cluster shares the lists: every CTA writes and reads the rows of its own tiles only)
        V.scr_col = A.scr_col + U.scr_off;
        V.queue = A.queue + U.q_off + (size_t)rank * 6 * U.qcap;  // (its own task queues)
    }
    uint32_t *pool = nullptr;
    if (U.pool_smem) {
        pool = (uint32_t *)(smem + lay.pool);
        block_copy((uint32_t *)(smem + lay.pool_desc), A.pool_desc + U.pooldesc_off, U.pool_blocks);
        if (threadIdx.x < 4) pool[4 * U.pool_blocks + threadIdx.x] = 0u;  // the spare block
    }
    V.pool = pool;
    return V;
}

// ------------------------------------------------------------------------------------------------------------------
// A latency-critical unit may be spread over a thread-block cluster (build_units, the ladder's last rungs): every CTA runs the
// same E-step on the same counts (same theta, nothing to exchange), takes every n_cta-th tile -- and with its chain tiles their
// splitting trees --, accumulates the fragments it places in a partial count vector of its own, and after one cluster barrier
// per iteration adds up the partial vectors of all CTAs through distributed shared memory (double-buffered, like gibbs_rows.cu).
// (CLUSTER is a template parameter so that the thousands of one-CTA units of a batch carry none of this: it cost them 3 % otherwise)
template <bool MAT_SMEM, bool SCR_SMEM, bool CLUSTER>
__global__ void __maxnreg__(BAYES_FAST_MAXNREG) gibbs_fast_bundle_kernel(const __grid_constant__ KernelArgs A, const int64_t unit_base) {
    extern __shared__ __align__(16) unsigned char smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int n_cta = CLUSTER ? (int)cluster.num_blocks() : 1, rank = CLUSTER ? (int)cluster.block_rank() : 0;
    const bool lead = rank == 0;
    const int64_t ui = unit_base + blockIdx.x / n_cta;
    if (lead && threadIdx.x == 0) A.unit_clock[3 * ui] = global_ns();
    const UnitDev &U = A.units[ui];
    const int n_slots = U.n_slots, Tpad = U.Tpad;
    const TileLayout lay = tile_layout(n_slots, Tpad, U.n_tiles, U.n_chain_tiles, U.n_chain_rows, U.tab_len, U.tab_smem != 0, MAT_SMEM, SCR_SMEM,
                                       U.qcap, U.pool_blocks, U.pool_smem != 0);
    uint64_t *skey = (uint64_t *)(smem + lay.key);
    FastBundleE E;
    E.theta = (double *)(smem + lay.theta);
    E.den = (double *)(smem + lay.den);
    E.mean = (double *)(smem + lay.mean);
    E.m2 = (double *)(smem + lay.m2);
    E.counts = (int32_t *)(smem + lay.counts);
    E.base = (int32_t *)(smem + lay.base);
    E.occ = (int32_t *)(smem + lay.occ);
    E.skey = skey;
    double *sgamma = (double *)(smem + lay.sgamma);
    int32_t *sT = (int32_t *)(smem + lay.sT), *sNf = (int32_t *)(smem + lay.sNf);
    E.sgamma = sgamma;
    E.sT = sT;
    E.sNf = sNf;
    E.Tpad = Tpad;
    E.n_slots = n_slots;
    {
        int sh = 1;
        while ((1 << sh) < Tpad) sh++;
        E.sh = sh;
    }
    if (U.tab_smem) {
        double *tv = (double *)(smem + lay.tab_val);
        int32_t *tr = (int32_t *)(smem + lay.tab_row);
        block_copy(tv, A.tab_val + U.tab_off, U.tab_len);
        block_copy(tr, A.tab_row + U.tabrow_off, n_slots * (Tpad + 1));
        E.tab_val = tv;
        E.tab_row = tr;
    } else {
        E.tab_val = A.tab_val + U.tab_off;
        E.tab_row = A.tab_row + U.tabrow_off;
    }
    TileView V = tile_stage<MAT_SMEM, SCR_SMEM>(A, U, lay, smem, rank);
    E.ctl = V.ctl;
    uint32_t *pool = (uint32_t *)(smem + lay.pool);
    const uint32_t *pool_desc = (const uint32_t *)(smem + lay.pool_desc);
    const int pool_blocks = U.pool_smem ? U.pool_blocks : 0;
    if (threadIdx.x < (unsigned)n_slots) {
        skey[threadIdx.x] = U.key[threadIdx.x];
        sgamma[threadIdx.x] = U.gamma[threadIdx.x];
        sT[threadIdx.x] = U.T[threadIdx.x];
        sNf[threadIdx.x] = U.N_f[threadIdx.x];
    }
    const int lane = threadIdx.x & 31;
    const int g = lane >> E.sh, t = lane & (Tpad - 1);
    const bool valid = g < n_slots && t < U.T[g < n_slots ? g : 0];
    if (threadIdx.x < 32) {
        const int32_t bc = A.base_counts[U.lane_off + lane];
        E.theta[lane] = 1.0;  // initAssignment runs the A-step on P_ij itself
        E.counts[lane] = bc;
        E.base[lane] = bc;
        E.occ[lane] = 0;
        E.mean[lane] = 0.0;
        E.m2[lane] = 0.0;
        E.den[lane] = A.efflen[U.lane_off + lane] * (double)(g < n_slots ? U.N_total[g] : 1);  // valueContainer.cpp:235 denominator
    }
    if (threadIdx.x < kFastCtlWords) V.ctl[threadIdx.x] = 0;

    int32_t *part = (int32_t *)(smem + lay.total);  // cluster: two buffers of 32 partial counts (kTileClusterBytes, only then)
    if (n_cta > 1 && threadIdx.x < 64) part[threadIdx.x] = 0;
    int cur = 0;
    int32_t *acc = n_cta > 1 ? part : E.counts;
    const bool tracing = lead && A.trace_unit == ui;
    const bool my_trace = tracing && threadIdx.x < 32 && g == A.trace_slot && valid;
    const int Tg = valid ? U.T[g] : 0;
    const int burn = U.burn;
    const int total = A.iter_cap > 0 && A.iter_cap < U.burn + U.samples ? A.iter_cap : U.burn + U.samples;
    const int stamp_at = A.iter_cap > 0 ? total >> 1 : -2;  // calibration pass: time the second half only
    const int n_threads = blockDim.x;
    __syncthreads();
#ifdef BAYES_TILE_PHASES  // development aid: thread 0 times the phases of every iteration (SM clock); large blocks print the averages
    long long ph_t = clock64(), ph_a[3] = {0, 0, 0}, ph_levels[2] = {0, 0};
#define TILE_PHASE(i) do { if (threadIdx.x == 0) { const long long now_ = clock64(); ph_a[i] += now_ - ph_t; ph_t = now_; } } while (0)
#else
#define TILE_PHASE(i) do { } while (0)
#endif

    // it == -1 is initAssignment (gibbsSampler.cpp:90): the same A-step code on theta == 1, no E-step before it
    for (int it = -1; it < total; it++) {
        const uint32_t iter = (uint32_t)(it < 0 ? 0 : it);
        if (threadIdx.x == 0) {
            V.ctl[kCtlSlab] = 0;
            V.ctl[kCtlQ0] = 0;
            V.ctl[kCtlQ0 + 1] = 0;
            V.ctl[kCtlQ0 + 2] = 0;
            if (lead && it == stamp_at) A.unit_clock[3 * ui] = global_ns();
        }
        if (it >= 0 && threadIdx.x < 32) {
            if (my_trace && it == 0) A.tr_init[t] = E.counts[lane];
            if (my_trace && it > 0 && it - 1 < A.trace_iters) A.tr_counts[(size_t)(it - 1) * Tg + t] = E.counts[lane];
            const int b = fast_bundle_e_step(E, iter, it >= burn || (stamp_at >= 0 && it >= stamp_at));
            if (my_trace && it < A.trace_iters) {
                A.tr_theta[(size_t)it * Tg + t] = E.theta[lane];
                if (t == 0) A.tr_b[it] = b;
            }
        }
        // the iteration's uniform words: by the warps that have no E-step to run (all of them for initAssignment)
        if (it < 0 || n_threads == 32) fast_pool_gen(pool, pool_desc, pool_blocks, skey, iter, it < 0, 0, n_threads);
        else if (threadIdx.x >= 32) fast_pool_gen(pool, pool_desc, pool_blocks, skey, iter, false, 32, n_threads - 32);
        __syncthreads();
        TILE_PHASE(0);
        tile_rows(V, E.theta, acc, skey, iter, it < 0, rank, n_cta);
        __syncthreads();
        TILE_PHASE(1);
#ifdef BAYES_TILE_PHASES
        tile_tree_levels(V, acc, skey, iter, it < 0, threadIdx.x == 0 ? ph_levels : nullptr);
#else
        tile_tree_levels(V, acc, skey, iter, it < 0);
#endif
        if (CLUSTER && n_cta > 1) {
            cluster.sync();
            if (threadIdx.x < 32) {
                int s = 0;
                for (int j = 0; j < n_cta; j++) s += cluster.map_shared_rank(part, j)[cur * 32 + lane];
                E.counts[lane] += s;
                part[(cur ^ 1) * 32 + lane] = 0;  // read by the other CTAs one iteration ago, before the barrier above
            }
            cur ^= 1;
            acc = part + cur * 32;
            __syncthreads();
        }
        TILE_PHASE(2);
    }
#ifdef BAYES_TILE_PHASES
    if (threadIdx.x == 0 && n_threads >= 256) {
        const double k = 1.0 / (1.965 * (total + 1));
        printf("tile phases, ns/iteration (%d slots, T0=%d, %d tiles (%d chain), %d threads): E-step / pool %.0f | tiles %.0f | trees %.0f (%.1f levels, %.0f nodes per "
               "iteration)\n", n_slots, U.T[0], U.n_tiles, U.n_chain_tiles, n_threads, ph_a[0] * k, ph_a[1] * k, ph_a[2] * k, (double)ph_levels[0] / (total + 1),
               (double)ph_levels[1] / (total + 1));
    }
#endif
    if (my_trace && total - 1 < A.trace_iters) A.tr_counts[(size_t)(total - 1) * Tg + t] = E.counts[lane];

    if (lead && threadIdx.x < 32 && valid) {
        const int64_t o = U.out_off[g] + t;
        A.out_occ[o] = E.occ[lane];
        A.out_mean[o] = E.mean[lane];
        A.out_m2[o] = E.m2[lane];
        A.out_counts[o] = E.counts[lane];
    }
    int err = V.ctl[kCtlErr];
    if (CLUSTER && n_cta > 1) {
        cluster.sync();
        if (lead && threadIdx.x == 0)
            for (int j = 1; j < n_cta; j++) {
                const int e = cluster.map_shared_rank(V.ctl, j)[kCtlErr];
                if (e) err = e;
            }
        cluster.sync();  // nobody leaves while its shared memory may still be read
    }
    if (lead && threadIdx.x == 0) {
        A.unit_err[ui] = err;
        A.unit_clock[3 * ui + 1] = global_ns();
        A.unit_clock[3 * ui + 2] = sm_id();
    }
}

// ---- single A-step in fast mode of a locus with up to 32 candidates: the tiles of a one-slot bundle, two warps ----------------
template <bool INIT>
__global__ void step_assignment_tiles_kernel(const __grid_constant__ KernelArgs A, const double *theta_in, uint32_t iteration, int32_t *counts_out) {
    __shared__ double theta[32];
    __shared__ int32_t counts[32];
    __shared__ uint64_t skey[1];
    __shared__ int ctl[kFastCtlWords];
    const UnitDev &U = A.units[0];
    const int T = U.T[0];
    if (threadIdx.x < 32) {
        theta[threadIdx.x] = (int)threadIdx.x < T ? (INIT ? 1.0 : theta_in[threadIdx.x]) : 0.0;
        counts[threadIdx.x] = A.base_counts[U.lane_off + threadIdx.x];
    }
    if (threadIdx.x == 0) skey[0] = U.key[0];
    if (threadIdx.x < kFastCtlWords) ctl[threadIdx.x] = 0;
    TileView V;
    V.n_tiles = U.n_tiles;
    V.n_chain_tiles = U.n_chain_tiles;
    V.qcap = U.qcap;
    V.ctl = ctl;
    V.pool_first = U.pool_first;
    V.sh = 5;  // one slot: every candidate lane belongs to slot 0
    V.max_row = 32;
    V.val = A.val + U.cell_off;
    V.meta = A.cell_meta + U.meta_off;
    V.ctab = A.chain_tab + U.ctab_off;
    V.scr_cum = A.scr_cum + U.scr_off;
    V.scr_col = A.scr_col + U.scr_off;
    V.queue = A.queue + U.q_off;
    V.pool = nullptr;
    __syncthreads();
    tile_rows(V, theta, counts, skey, iteration, INIT);
    __syncthreads();
    tile_tree_levels(V, counts, skey, iteration, INIT);
    __syncthreads();
    if ((int)threadIdx.x < T) counts_out[threadIdx.x] = counts[threadIdx.x];
    if (threadIdx.x == 0) A.unit_err[0] = ctl[kCtlErr];
}

}  // namespace

static int check_fast(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return BAYES_OK;
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return BAYES_E_CUDA;
}

int configure_fast_kernels() {
    int rc;
    const char *co = getenv("BAYES_CARVEOUT");  // experiment knob, percent (see configure_kernels)
    const int carveout = co ? atoi(co) : 65;
    const void *fns[6] = {(const void *)gibbs_fast_bundle_kernel<true, true, false>, (const void *)gibbs_fast_bundle_kernel<false, true, false>,
                          (const void *)gibbs_fast_bundle_kernel<false, false, false>, (const void *)gibbs_fast_bundle_kernel<true, true, true>,
                          (const void *)gibbs_fast_bundle_kernel<false, true, true>, (const void *)gibbs_fast_bundle_kernel<false, false, true>};
    for (const void *f : fns) {
        if ((rc = check_fast(cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, carveout), "carveout"))) return rc;
        if ((rc = check_fast(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_unit_smem()), "cudaFuncSetAttribute"))) return rc;
    }
    return BAYES_OK;
}

int launch_units_fast(const KernelArgs &args, int64_t unit_base, int64_t n_units, int block, bool wide, bool mat_smem, bool scr_smem,
                      size_t smem_bytes, void *stream, int n_cta) {
    if (n_units <= 0) return BAYES_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)n_units;
    if (!wide && n_cta > 1) {  // critical bundles spread over a thread-block cluster
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(n_units * n_cta), 1, 1);
        cfg.blockDim = dim3((unsigned)block, 1, 1);
        cfg.dynamicSmemBytes = smem_bytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)n_cta;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e;
        if (mat_smem) e = cudaLaunchKernelEx(&cfg, gibbs_fast_bundle_kernel<true, true, true>, args, unit_base);
        else if (scr_smem) e = cudaLaunchKernelEx(&cfg, gibbs_fast_bundle_kernel<false, true, true>, args, unit_base);
        else e = cudaLaunchKernelEx(&cfg, gibbs_fast_bundle_kernel<false, false, true>, args, unit_base);
        return check_fast(e, "gibbs fast kernel launch (cluster)");
    }
    if (wide) {  // loci with more than 32 candidates run in gibbs_rows.cu (launch_units_rows)
        set_error("launch_units_fast: wide units belong to the cooperative-rows kernel");
        return BAYES_E_STATE;
    }
    if (mat_smem) gibbs_fast_bundle_kernel<true, true, false><<<grid, block, smem_bytes, st>>>(args, unit_base);
    else if (scr_smem) gibbs_fast_bundle_kernel<false, true, false><<<grid, block, smem_bytes, st>>>(args, unit_base);
    else gibbs_fast_bundle_kernel<false, false, false><<<grid, block, smem_bytes, st>>>(args, unit_base);
    return check_fast(cudaGetLastError(), "gibbs fast kernel launch");
}

int launch_step_assignment_fast(const KernelArgs &args, bool init, bool tiles, const double *d_theta, uint32_t iteration, int32_t *d_counts_out,
                                size_t smem_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (tiles) {
        if (init) step_assignment_tiles_kernel<true><<<1, 64, 0, st>>>(args, d_theta, iteration, d_counts_out);
        else step_assignment_tiles_kernel<false><<<1, 64, 0, st>>>(args, d_theta, iteration, d_counts_out);
        return check_fast(cudaGetLastError(), "step_assignment_tiles_kernel launch");
    }
    (void)smem_bytes;
    set_error("launch_step_assignment_fast: loci with more than 32 candidates take launch_step_assignment_rows");
    return BAYES_E_STATE;
}

}  // namespace bayes
