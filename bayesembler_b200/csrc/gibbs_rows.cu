// Fast (production) sampling mode for loci with MORE than 32 candidates (config-3 class: 200-1000 candidate paths, a
// thousand collapsed rows, 0.4-2.5 MB of likelihood matrix, 10^5-10^6 strictly sequential Gibbs iterations).
//
// What bounds such a locus is the latency of ONE iteration, so the kernel is organised around that:
//
//  * One locus = one thread-block CLUSTER (1, 2, 4, 8 or 16 CTAs, chosen by the host from how many such loci compete for
//    the SMs).  The rows are dealt over the CTAs; each CTA keeps its share of the matrix (values, candidate masks, row
//    table) resident in its own shared memory, so an iteration reads nothing from L2 / HBM.
//  * The E-step (ExpressionGenerator::generateExpression, expressionGenerator.cpp:77-132; simplex size,
//    simplexSizeGenerator.cpp:209-224; getFPKM + GibbsOutput::addSample) is REPLICATED: every CTA computes the same theta
//    from the same counts with the same Philox sites, so theta and the simplex mask are never exchanged.  Inside a CTA it
//    runs on all warps (one Gamma draw per thread); only the null-candidate selection is sequential, and that is a
//    warp-wide rank select (prefix counts in the lanes, one ballot per pick) instead of a walk over the chunks.
//  * The A-step (AssignmentGenerator::generateAssignment, assignmentGenerator.cpp:263-373) is WARP PER ROW: the lanes first
//    hold the row's candidate-mask words (ANDed with the simplex mask: the cells in play, a few tens of the row's
//    hundreds of stored cells), a shuffle prefix sum ranks them, then the lanes hold the cells in play themselves, 32 at
//    a time: p = P_ij theta_j, warp-shuffle inclusive scan, thresholds against the row's pool words (rows with < 20
//    fragments) or compaction into the scratch list the splitting trees work on (rows with >= 20 fragments; same trees as
//    gibbs_fast.cu).  Stored cells that are not in play are never touched.
//  * Counts: every CTA accumulates the fragments of ITS rows into a partial vector in its own shared memory; after one
//    cluster barrier every CTA adds up the partial vectors of all CTAs through distributed shared memory (integer sums:
//    the order does not matter).  The partial vectors are double-buffered, so one cluster barrier per iteration suffices.
//
// The order of the floating-point additions of a row sum is fixed by the scan (chunks of 32 cells in play, Hillis-Steele
// inside a chunk, chunks chained left to right); the CPU oracle sums in the same order (oracle/bayes_oracle.c,
// assignment_fast) and the parity tests compare the two draw for draw.  No tensor cores: nothing here is a dense contraction.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <float.h>
#include <stdio.h>
#include <stdlib.h>

#include "internal.h"
#include "layout.h"
#include "rng.cuh"
#include "gibbs_common.cuh"

namespace cg = cooperative_groups;

#ifndef BAYES_ROWS_MAXNREG
#define BAYES_ROWS_MAXNREG 128
#endif

namespace bayes {

namespace {

constexpr uint32_t kFull = 0xffffffffu;
// ctl words
constexpr int kCtlQ0 = 1, kCtlErr = 4, kCtlNAct = 5, kCtlB = 6;

#ifdef BAYES_DEBUG_BOUNDS
#define ROWS_CHECK(cond, code) do { if (!(cond)) V.ctl[kCtlErr] = (code); } while (0)
#else
#define ROWS_CHECK(cond, code) do { } while (0)
#endif

// Development aid: -DBAYES_ROWS_PHASES makes thread 0 of the cluster's first CTA time the phases of every iteration (SM clock)
// and print the averages when the unit ends.
#ifdef BAYES_ROWS_PHASES
struct PhaseClock {
    long long t, a[10];
};
#define ROWS_PHASE(pc, i) do { if ((pc) && threadIdx.x == 0) { const long long now_ = clock64(); (pc)->a[i] += now_ - (pc)->t; (pc)->t = now_; } } while (0)
#else
struct PhaseClock;
#define ROWS_PHASE(pc, i) do { } while (0)
#endif

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ int words_below(const u32x4 &w, int left, uint32_t t) {
    return (int)(w.x <= t) + ((left > 1) & (int)(w.y <= t)) + ((left > 2) & (int)(w.z <= t)) + ((left > 3) & (int)(w.w <= t));
}

// What one CTA works on.  Row / cell indices are those of the unit; the staged copies start at row_base / cell_base.
struct RowsView {
    const double *val;         // cell - cell_base
    const uint64_t *rmask;     // (row - row_base) * n_words + word
    const int32_t *rcell;      // row - row_base -> first cell of the row (one more entry than rows)
    const int32_t *row_n;      // fragments of the row
    const uint32_t *row_info;  // slot << 27 | original row index
    const int32_t *row_woff;   // rows with < 20 fragments: first word of the row in the CTA's pool
    int row_base, cell_base;
    int row_begin, row_end, chain_end;  // this CTA's rows; [row_begin, chain_end) have >= 20 fragments
    int n_words;
    double *scr_cum;           // rows with >= 20 fragments: running sums of the cells in play, entry (first cell - scr_base + i)
    uint16_t *scr_col;         // their candidates
    int scr_base, scr_len;
    uint32_t *queue;           // two buffers of 3 * qcap words: fragments | node | row
    int qcap;
    const uint32_t *pool;      // this iteration's uniform words of the CTA's rows
    double *ch_rz;             // rows with >= 20 fragments (row - row_begin): 1 / row sum of this iteration
    int *ch_kk;                // their cells in play
    const uint32_t *dtask;     // direct-draw tasks of the CTA: (row - row_begin) << 16 | block of 4 fragments
    int n_dtask;
    int *ctl;
};

// Warp-aggregated append of at most one task per lane to queue buffer `buf`, counted by ctl word `cnt`.  All 32 lanes call.
__device__ __forceinline__ void push_tasks(const RowsView &V, int cnt, int buf, bool want, uint32_t n, uint32_t node, uint32_t row) {
    const uint32_t m = __ballot_sync(kFull, want);
    if (m == 0u) return;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(&V.ctl[cnt], __popc(m));
    base = __shfl_sync(kFull, base, leader);
    if (want) {
        const int idx = base + __popc(m & lanemask_lt());
        ROWS_CHECK(idx < V.qcap, 201);
        uint32_t *q = V.queue + (size_t)buf * 3 * V.qcap;
        q[idx] = n;
        q[V.qcap + idx] = node;
        q[2 * V.qcap + idx] = row;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// A-step, first part: one warp per row.
// ------------------------------------------------------------------------------------------------------------------
// One group of 32 mask words (32 bits each: 1024 candidates) of a row, one word per lane.
struct MaskGroup {
    uint32_t m_all, m;  // my word: the candidates the row has a cell for; those of them in the simplex (= the cells in play)
    uint32_t excl;      // stored cells before my word << 16 | cells in play before my word, within the group
    int kbase;          // stored cells of the row before this group
    int w0;             // first word of the group
    int Kg, stored;     // cells in play / stored cells of the group
};

__device__ __forceinline__ MaskGroup mask_group(uint32_t m_all, uint32_t m, int w0, int kbase, int lane) {
    MaskGroup G;
    G.m_all = m_all;
    G.m = m;
    const uint32_t pk = ((uint32_t)__popc(m_all) << 16) | (uint32_t)__popc(m);
    uint32_t incl = pk;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += up;
    }
    const uint32_t tot = __shfl_sync(kFull, incl, 31);
    G.excl = incl - pk;
    G.kbase = kbase;
    G.w0 = w0;
    G.Kg = (int)(tot & 0xffffu);
    G.stored = (int)(tot >> 16);
    return G;
}

// Round j0 of a group: lane l takes the (j0 + l)-th cell in play of the group; returns p = P_ij theta_j and the candidate.
__device__ __forceinline__ double round_p(const MaskGroup &G, int j0, int lane, const double *val, const double *theta, int &c) {
    const int j = j0 + lane;
    // the word my cell sits in: the last lane whose first cell in play is not after mine
    const int my_jb = (int)(G.excl & 0xffffu);
    int l = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
        const int jb = __shfl_sync(kFull, my_jb, l + step);
        if (jb <= j) l += step;
    }
    const uint32_t ms = __shfl_sync(kFull, G.m, l), mas = __shfl_sync(kFull, G.m_all, l), ex = __shfl_sync(kFull, G.excl, l);
    double p = 0.0;
    c = 0;
    if (j < G.Kg) {
        const int bit = (int)__fns(ms, 0, j - (int)(ex & 0xffffu) + 1);
        const int k = G.kbase + (int)(ex >> 16) + __popc(mas & ((1u << bit) - 1u));
        c = ((G.w0 + l) << 5) + bit;
        p = val[k] * theta[c];
    }
    return p;
}

// inclusive scan over the lanes (the order of these additions is part of the mode's definition: the oracle sums alike)
__device__ __forceinline__ double warp_scan(double x, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double up = __shfl_up_sync(kFull, x, d);
        if (lane >= d) x += up;
    }
    return x;
}

// Rows with < 20 fragments: where the row's n pool words land among the cells in play of one round (bal != 0).  Cell i of
// the row takes the words w with cum_(i-1) / Z <= w 2^-32 < cum_i / Z, the last cell in play what the others left.
__device__ __forceinline__ void round_thresholds(double cum, int c, uint32_t bal, int kk, int kk_total, double rZ, int n, const uint32_t *pw, int32_t *acc,
                                                 int &below_carry, int lane, uint32_t lt) {
    const bool inplay = (bal >> lane) & 1u;
    const int pos = kk + __popc(bal & lt);
    int below = 0;
    if (inplay) {
        if (pos == kk_total - 1) {
            below = n;
        } else {
            // word * 2^-32 < cum / Z  <=>  word <= ceil(cum (1/Z) 2^32 - 1); the conversion saturates
            const uint32_t t = __double2uint_ru(__fma_rn(cum * rZ, 4294967296.0, -1.0));
            for (int b = 0; b < n; b += 4) {
                const u32x4 wv = {pw[b], pw[b + 1], pw[b + 2], pw[b + 3]};
                below += words_below(wv, n - b < 4 ? n - b : 4, t);
            }
        }
    }
    const uint32_t lower = bal & lt;
    int prev = __shfl_sync(kFull, below, lower ? 31 - __clz(lower) : 0);
    if (!lower) prev = below_carry;
    if (inplay && below > prev) atomicAdd(&acc[c], below - prev);
    below_carry = __shfl_sync(kFull, below, 31 - __clz(bal));
}

constexpr int kRowsCachedRounds = 3;  // rows with < 20 fragments and up to 96 cells in play keep their running sums in registers

__device__ __forceinline__ void rows_assign(const RowsView &V, const uint64_t *amask64, const double *theta, int32_t *acc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const uint32_t lt = lanemask_lt();
    const int n_w32 = 2 * V.n_words;
    const uint32_t *amask = (const uint32_t *)amask64;
    // the row after this one: its table entries and first mask words are on their way while this one is worked on
    int r = V.row_begin + warp;
    int nx_n = 0, nx_cell0 = 0, nx_woff = 0;
    uint32_t nx_mall = 0u;
    if (r < V.row_end) {
        const int ri = r - V.row_base;
        nx_n = V.row_n[ri];
        nx_cell0 = V.rcell[ri];
        nx_woff = V.row_woff[ri];
        if (lane < n_w32) nx_mall = ((const uint32_t *)(V.rmask + (size_t)ri * V.n_words))[lane];
    }
    for (; r < V.row_end; r += n_warps) {
        const int ri = r - V.row_base;
        const int n = nx_n, cell0 = nx_cell0;
        const uint32_t *pw = V.pool + nx_woff;
        const uint32_t mall0 = nx_mall;
        if (r + n_warps < V.row_end) {
            const int rj = ri + n_warps;
            nx_n = V.row_n[rj];
            nx_cell0 = V.rcell[rj];
            nx_woff = V.row_woff[rj];
            if (lane < n_w32) nx_mall = ((const uint32_t *)(V.rmask + (size_t)rj * V.n_words))[lane];
        }
        const bool chain = r < V.chain_end;
        const double *val = V.val + (cell0 - V.cell_base);
        const uint32_t *rm = (const uint32_t *)(V.rmask + (size_t)ri * V.n_words);
        MaskGroup G = mask_group(mall0, lane < n_w32 ? mall0 & amask[lane] : 0u, 0, 0, lane);
        if (chain) {
            // rows with >= 20 fragments: the running sums of the cells in play go to the row's scratch list, the splitting tree
            // (rows_tree_levels) works on that list
            const int sb = cell0 - V.scr_base;
            double carry = 0.0;
            int kk = 0, c_only = 0;
            for (int w0 = 0;;) {
                for (int j0 = 0; j0 < G.Kg; j0 += 32) {
                    int c;
                    const double p = round_p(G, j0, lane, val, theta, c);
                    const double cum = carry + warp_scan(p, lane);
                    carry = __shfl_sync(kFull, cum, G.Kg - 1 - j0 < 31 ? G.Kg - 1 - j0 : 31);
                    const bool inplay = p > 0.0;
                    const uint32_t bal = __ballot_sync(kFull, inplay);
                    if (inplay) {
                        const int at = sb + kk + __popc(bal & lt);
                        ROWS_CHECK(at >= 0 && at < V.scr_len, 203);
                        V.scr_cum[at] = cum;
                        V.scr_col[at] = (uint16_t)c;
                    }
                    if (bal) c_only = __shfl_sync(kFull, c, 31 - __clz(bal));
                    kk += __popc(bal);
                }
                w0 += 32;
                if (w0 >= n_w32) break;
                const int w = w0 + lane;
                const uint32_t ma = w < n_w32 ? rm[w] : 0u;
                G = mask_group(ma, w < n_w32 ? ma & amask[w] : 0u, w0, G.kbase + G.stored, lane);
            }
            if (n <= kRowsDirectMax && lane == 0) {  // drawn fragment by fragment once all rows have their lists (rows_direct)
                V.ch_kk[r - V.row_begin] = kk;
                V.ch_rz[r - V.row_begin] = __drcp_rn(carry);
            }
            if (kk == 0) {  // reference: assert(normalisation_constant >= double_underflow) (assignmentGenerator.cpp:284)
                if (lane == 0) V.ctl[kCtlErr] = 1;
            } else if (kk == 1) {
                if (lane == 0) atomicAdd(&acc[c_only], n);
            } else if (n > kRowsDirectMax) {
                // the row's fragments enter the tree as ceil(n / 64) independent chunks while that is at most kTreeChunks, else whole
                const int mch = (n + 63) >> 6 <= kTreeChunks ? (n + 63) >> 6 : 1;
                for (int c0 = 0; c0 < mch; c0 += 32) {
                    const int ch = c0 + lane;
                    push_tasks(V, kCtlQ0, 0, ch < mch, (uint32_t)(n / mch + (ch < n % mch)), ((uint32_t)(kk - 1) << 12) | ((uint32_t)ch << 24), (uint32_t)r);
                }
            }
            continue;
        }
        if (n_w32 <= 32 && G.Kg <= 32 * kRowsCachedRounds) {
            // the usual row: a few tens of cells in play; their running sums stay in registers until the row sum is known
            double cum[kRowsCachedRounds];
            int cc[kRowsCachedRounds];
            uint32_t bal[kRowsCachedRounds];
            double carry = 0.0;
            int kk_total = 0;
#pragma unroll
            for (int rd = 0; rd < kRowsCachedRounds; rd++) {
                cum[rd] = 0.0;
                cc[rd] = 0;
                bal[rd] = 0u;
                if (rd * 32 < G.Kg) {
                    const double p = round_p(G, rd * 32, lane, val, theta, cc[rd]);
                    cum[rd] = carry + warp_scan(p, lane);
                    carry = __shfl_sync(kFull, cum[rd], G.Kg - 1 - rd * 32 < 31 ? G.Kg - 1 - rd * 32 : 31);
                    bal[rd] = __ballot_sync(kFull, p > 0.0);
                    kk_total += __popc(bal[rd]);
                }
            }
            if (kk_total == 0) {
                if (lane == 0) V.ctl[kCtlErr] = 1;
                continue;
            }
            const double rZ = __drcp_rn(carry);
            int kk = 0, below_carry = 0;
#pragma unroll
            for (int rd = 0; rd < kRowsCachedRounds; rd++)
                if (bal[rd]) {
                    round_thresholds(cum[rd], cc[rd], bal[rd], kk, kk_total, rZ, n, pw, acc, below_carry, lane, lt);
                    kk += __popc(bal[rd]);
                }
            continue;
        }
        // a row with more cells in play than that (the first iterations, when the simplex still holds most candidates) is walked
        // twice: once for its sum, once -- the same arithmetic again -- for the thresholds
        double rZ = 0.0;
        int kk_total = 0;
        for (int pass = 0; pass < 2; pass++) {
            double carry = 0.0;
            int kk = 0, below_carry = 0;
            if (pass == 1) G = mask_group(mall0, lane < n_w32 ? mall0 & amask[lane] : 0u, 0, 0, lane);
            for (int w0 = 0;;) {
                for (int j0 = 0; j0 < G.Kg; j0 += 32) {
                    int c;
                    const double p = round_p(G, j0, lane, val, theta, c);
                    const double cum = carry + warp_scan(p, lane);
                    carry = __shfl_sync(kFull, cum, G.Kg - 1 - j0 < 31 ? G.Kg - 1 - j0 : 31);
                    const uint32_t bal = __ballot_sync(kFull, p > 0.0);
                    if (pass == 1 && bal) round_thresholds(cum, c, bal, kk, kk_total, rZ, n, pw, acc, below_carry, lane, lt);
                    kk += __popc(bal);
                }
                w0 += 32;
                if (w0 >= n_w32) break;
                const int w = w0 + lane;
                const uint32_t ma = w < n_w32 ? rm[w] : 0u;
                G = mask_group(ma, w < n_w32 ? ma & amask[w] : 0u, w0, G.kbase + G.stored, lane);
            }
            if (kk == 0) {
                if (lane == 0) V.ctl[kCtlErr] = 1;
                break;
            }
            rZ = __drcp_rn(carry);
            kk_total = kk;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// A-step, second part: the rows with 20 .. kRowsDirectMax fragments, fragment by fragment.  One lane per block of four
// fragments, whichever row they belong to: one Philox block of the row's own site, four binary searches (in step, for the
// memory-level parallelism) of the words in the row's running sums -- the thresholds of the small rows, word <=
// ceil(cum (1/Z) 2^32 - 1) -- and one shared-memory atomic per fragment.  Nothing here depends on anything but the lists the
// first part left: no levels, no barriers, every lane of the CTA busy.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void rows_direct(const RowsView &V, int32_t *acc, const uint64_t key, uint32_t iteration, const bool init) {
    const uint32_t ph = init ? PH_INIT_DIRECT : PH_DIRECT;
    for (int task = threadIdx.x; task < V.n_dtask; task += (int)blockDim.x) {
        const uint32_t e = V.dtask[task];
        const int rl = (int)(e >> 16);
        const uint32_t blk = e & 0xffffu;
        const int kk = V.ch_kk[rl];
        if (kk < 2) continue;  // no cell in play (flagged) or a single one (took everything already)
        const int ri = V.row_begin + rl - V.row_base;
        const int n = V.row_n[ri];
        const int sb = V.rcell[ri] - V.scr_base;
        const double rZ = V.ch_rz[rl];
        const double *sc = V.scr_cum + sb;
        const u32x4 wv = make_site(key, ph, iteration, V.row_info[ri] & ((1u << kRowSlotShift) - 1u), blk >> 12).block(blk & 4095u);
        const uint32_t w[4] = {wv.x, wv.y, wv.z, wv.w};
        const int left = n - 4 * (int)blk < 4 ? n - 4 * (int)blk : 4;
        // number of thresholds below the word = first cell whose threshold the word does not exceed (the last cell has none)
        int lo[4] = {0, 0, 0, 0};
        const int n_search = kk - 1;
        for (int step = 1 << (31 - __clz(n_search)); step > 0; step >>= 1) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int cand = lo[q] + step;
                if (cand <= n_search) {
                    ROWS_CHECK(sb + cand - 1 >= 0 && sb + cand - 1 < V.scr_len, 205);
                    const uint32_t t = __double2uint_ru(__fma_rn(sc[cand - 1] * rZ, 4294967296.0, -1.0));
                    if (w[q] > t) lo[q] = cand;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (q < left) atomicAdd(&acc[V.scr_col[sb + lo[q]]], 1);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// A-step, third part: the splitting trees of the rows with >= 20 fragments, level by level (the scheme of
// fast_tree_levels in gibbs_fast.cu: one lane per node of the level whichever row it belongs to, three rotating
// counters, two queue buffers, one block barrier per level).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void rows_tree_levels(const RowsView &V, int32_t *acc, const uint64_t key, uint32_t iteration, const bool init) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ph_tree = init ? PH_INIT_TREE : PH_TREE;
    const int qcap = V.qcap;
    int cur = kCtlQ0, nxt = kCtlQ0 + 1, old = kCtlQ0 + 2, buf = 0;
    for (;;) {
        const int Q = V.ctl[cur];
        if (Q == 0) break;
        if (threadIdx.x == 0) V.ctl[old] = 0;
        const uint32_t *q = V.queue + (size_t)buf * 3 * qcap;
        for (int i0 = warp * 32; i0 < Q; i0 += (int)blockDim.x) {
            const int i = i0 + lane;
            bool push_l = false, push_r = false;
            uint32_t n_l = 0, n_r = 0, node_l = 0, node_r = 0, row = 0;
            if (i < Q) {
                const uint32_t n = q[i], node = q[qcap + i];
                row = q[2 * qcap + i];
                const uint32_t chunk = node >> 24;
                const int a = (int)(node & 4095u), len = (int)((node >> 12) & 4095u) + 1, half = len >> 1, len_r = len - half;
                const int ri = (int)row - V.row_base;
                const int sb = V.rcell[ri] - V.scr_base;
                ROWS_CHECK(sb >= 0 && sb + a + len - 1 < V.scr_len, 204);
                const double *sc = V.scr_cum + sb;
                const double Pa = a ? sc[a - 1] : 0.0, Pm = sc[a + half - 1], Pb = sc[a + len - 1];
                const double p_left = (Pm - Pa) / (Pb - Pa);
                const uint32_t info = V.row_info[ri];
                const Site site = make_site(key, ph_tree | ((node & 0xFFFFFFu) << 8), iteration, info & ((1u << kRowSlotShift) - 1u), chunk);
                n_l = (uint32_t)binomial_variate(site, (int)n, p_left);
                n_r = n - n_l;
                if (n_l) {
                    if (half == 1) atomicAdd(&acc[V.scr_col[sb + a]], (int)n_l);
                    else { push_l = true; node_l = (uint32_t)a | ((uint32_t)(half - 1) << 12) | (chunk << 24); }
                }
                if (n_r) {
                    if (len_r == 1) atomicAdd(&acc[V.scr_col[sb + a + half]], (int)n_r);
                    else { push_r = true; node_r = (uint32_t)(a + half) | ((uint32_t)(len_r - 1) << 12) | (chunk << 24); }
                }
            }
            push_tasks(V, nxt, buf ^ 1, push_l, n_l, node_l, row);
            push_tasks(V, nxt, buf ^ 1, push_r, n_r, node_r, row);
        }
        __syncthreads();
        const int t = old;
        old = cur;
        cur = nxt;
        nxt = t;
        buf ^= 1;
    }
}

// FixedBinomialFixedGammaSimplexSizeGenerator::sampleSimplexSize (simplexSizeGenerator.cpp:214-224) by a whole warp: the
// table lives in global memory (L2), so instead of ten dependent probes the 32 lanes probe 32 evenly spaced entries at once.
// Same result as sample_simplex (gibbs_common.cuh): first entry of the CDF row of |S+| that exceeds the uniform.
__device__ __forceinline__ int sample_simplex_warp(const double *tab_val, const int32_t *tab_row, int n_plus, uint64_t key, uint32_t iteration, int lane) {
    const double u = u01(make_site(key, PH_SIMPLEX, iteration, 0, 0).block(0).x);
    if (n_plus < 1) n_plus = 1;  // only after fragments were lost to a flagged error (BAYES_E_NUMERIC): stay inside the table
    int lo = tab_row[n_plus - 1], hi = tab_row[n_plus];
    const int first = lo;
    while (lo < hi) {
        const int step = (hi - lo + 31) >> 5;
        const int idx = lo + (lane + 1) * step - 1;
        const bool above = idx < hi ? u < tab_val[idx] : true;
        const int f = __ffs(__ballot_sync(kFull, above)) - 1;  // lane 31 probes at or beyond the end: some lane always answers
        const int idx_f = lo + (f + 1) * step - 1;
        lo += f * step;
        hi = idx_f < hi ? idx_f : hi;
        if (step == 1) {
            lo = hi;
            break;
        }
    }
    return (lo - first) + n_plus;
}

// The CTA's share of the matrix goes to shared memory as ONE bulk asynchronous copy (cp.async.bulk, the TMA unit's 1-D form),
// completion counted in bytes by an mbarrier.  Source and destination have to be 16-byte aligned and the size a multiple of 16:
// the destination is shifted by one entry when needed so that both agree modulo 16, an odd first / last entry is copied by hand.
__device__ __forceinline__ const double *stage_values_bulk(double *dst, const double *src, int n, uint64_t *bar) {
    if ((((uintptr_t)src) ^ ((uintptr_t)dst)) & 8u) dst += 1;
    const int head = (((uintptr_t)src) & 8u) ? 1 : 0;
    const int bulk = n > head ? (n - head) & ~1 : 0;
    const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_addr), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (head && n > 0) dst[0] = src[0];
        if (n > 0 && head + bulk < n) dst[n - 1] = src[n - 1];
        const uint32_t bytes = (uint32_t)bulk * 8u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(bytes) : "memory");
        if (bytes)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(dst + head)),
                         "l"(src + head), "r"(bytes), "r"(bar_addr)
                         : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar_addr), "r"(0) : "memory");
    return dst;
}

// the iteration's uniform words of this CTA's rows: block b of the CTA's pool = Philox block (desc & 0xffffff) of the locus' pool
__device__ __forceinline__ void rows_pool_gen(uint32_t *pool, const uint32_t *pool_desc, int pool_blocks, const uint64_t key, uint32_t iteration,
                                              bool init, int first_thread, int n_threads) {
    const uint32_t ph = init ? PH_INIT_POOL : PH_POOL;
    for (int b = (int)threadIdx.x - first_thread; b < pool_blocks; b += n_threads) {
        if (b < 0) continue;
        const uint32_t blk = pool_desc[b];
        const u32x4 w = make_site(key, ph, iteration, blk >> 12, 0).block(blk & 4095u);
        reinterpret_cast<uint4 *>(pool)[b] = make_uint4(w.x, w.y, w.z, w.w);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// E-step on all warps of the CTA (every CTA of the cluster runs it on the same counts and gets the same theta).
// Same draws, the same sites and the same order of the floating-point sums as wide_e_step (gibbs_common.cuh).
// Welford accumulators of candidate t are kept by CTA (t mod n_cta).  Ends with a block barrier.
// ------------------------------------------------------------------------------------------------------------------
struct RowsState {
    double *theta, *den, *mean, *m2, *norm;
    int32_t *counts, *occ;
    const int32_t *base;
    uint32_t *mask;    // [0, nch): S+ of the counts; [nch, 2 nch): picked null candidates
    int32_t *actoff;   // first entry of every 32-candidate chunk in act
    uint16_t *act;     // members of the simplex, ascending
    uint64_t *amask;   // the simplex as a bit mask, 64 candidates per word
    const double *tab_val;
    const int32_t *tab_row;
    int *ctl;
    int T, nch, N_f;
    double gamma;
};

__device__ __forceinline__ int rows_e_step(const RowsState &E, const uint64_t key, uint32_t iteration, bool sampling, int rank, int n_cta,
                                           uint32_t *pool, const uint32_t *pool_desc, int pool_blocks, PhaseClock *pc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5, n_threads = blockDim.x;
    const int T = E.T, nch = E.nch;
    const uint32_t lt = lanemask_lt();
    // 1. S+ of the counts, theta cleared
    for (int c = warp; c < nch; c += n_warps) {
        const int t = c * 32 + lane;
        const uint32_t pm = __ballot_sync(kFull, t < T && E.counts[t] > 0);
        if (lane == 0) {
            E.mask[c] = pm;
            E.mask[nch + c] = 0u;
        }
        if (t < T) E.theta[t] = 0.0;
    }
    __syncthreads();
    ROWS_PHASE(pc, 0);
    // 2. warp 0: simplex size and the null candidates that join S+; the other warps: this iteration's uniform words
    if (warp == 0) {
        const int G = (nch + 31) >> 5;  // chunks per lane (<= 4: T <= 4096), lane l holds chunks [l G, (l + 1) G)
        uint32_t nullw[4], pmw[4], picked[4];
        int cnt = 0, plus = 0;
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const int c = lane * G + g;
            uint32_t pm = 0u, vmask = 0u;
            if (g < G && c < nch) {
                pm = E.mask[c];
                vmask = (c == nch - 1 && (T & 31)) ? (1u << (T & 31)) - 1u : kFull;
            }
            pmw[g] = pm;
            nullw[g] = ~pm & vmask;
            picked[g] = 0u;
            cnt += __popc(nullw[g]);
            plus += __popc(pm);
        }
        for (int o = 16; o > 0; o >>= 1) plus += __shfl_xor_sync(kFull, plus, o);
        const int n_plus = plus;
        const int b = sample_simplex_warp(E.tab_val, E.tab_row, n_plus, key, iteration, lane);
        const int n_null = T - n_plus, n_pick = b - n_plus;
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += up;
        }
        // expressionGenerator.cpp:104-127: pick k takes the (draw k)-th of the null candidates still left, in ascending order.
        // The draws are independent (site = pick ordinal): 32 at a time, one per lane; the selection is sequential.
        for (int k0 = 0; k0 < n_pick; k0 += 32) {
            uint32_t my_pick = 0u;
            if (k0 + lane < n_pick)
                my_pick = uniform_int(make_site(key, PH_NULLPICK, iteration, (uint32_t)(k0 + lane), 0), (uint32_t)(n_null - k0 - lane));
            const int kn = n_pick - k0 < 32 ? n_pick - k0 : 32;
            for (int k = 0; k < kn; k++) {
                const int rem = (int)__shfl_sync(kFull, my_pick, k);
                const uint32_t ge = __ballot_sync(kFull, incl > rem);
                if (ge == 0u) continue;  // cannot happen: rem < null candidates left
                const int owner = __ffs(ge) - 1;
                if (lane == owner) {
                    int r2 = rem - (incl - cnt);
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        const int pc = __popc(nullw[g]);
                        if (r2 >= 0 && r2 < pc) {
                            const uint32_t bit = __fns(nullw[g], 0, r2 + 1);
                            nullw[g] &= ~(1u << bit);
                            picked[g] |= 1u << bit;
                            r2 = -1;
                        } else if (r2 >= 0) {
                            r2 -= pc;
                        }
                    }
                    cnt--;
                }
                if (lane >= owner) incl--;
            }
        }
        int members = 0;
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const int c = lane * G + g;
            if (g < G && c < nch) E.mask[nch + c] = picked[g];
            members += __popc(pmw[g] | picked[g]);
        }
        int mi = members;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(kFull, mi, d);
            if (lane >= d) mi += up;
        }
        int at = mi - members;
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const int c = lane * G + g;
            if (g < G && c < nch) {
                E.actoff[c] = at;
                at += __popc(pmw[g] | picked[g]);
            }
        }
        if (lane == 31) E.ctl[kCtlNAct] = mi;
        if (lane == 0) E.ctl[kCtlB] = b;
        if (n_threads == 32) rows_pool_gen(pool, pool_desc, pool_blocks, key, iteration, false, 0, 32);
    } else {
        rows_pool_gen(pool, pool_desc, pool_blocks, key, iteration, false, 32, n_threads - 32);
    }
    __syncthreads();
    ROWS_PHASE(pc, 1);
    // 3. the members of the simplex, ascending, and the simplex mask the A-step ANDs the rows' masks with
    for (int c = warp; c < nch; c += n_warps) {
        const uint32_t m = E.mask[c] | E.mask[nch + c];
        if ((m >> lane) & 1u) E.act[E.actoff[c] + __popc(m & lt)] = (uint16_t)(c * 32 + lane);
    }
    for (int w = threadIdx.x; w < (nch + 1) / 2; w += n_threads) {
        const uint32_t lo = E.mask[2 * w] | E.mask[nch + 2 * w];
        const uint32_t hi = 2 * w + 1 < nch ? E.mask[2 * w + 1] | E.mask[nch + 2 * w + 1] : 0u;
        E.amask[w] = ((uint64_t)hi << 32) | lo;
    }
    __syncthreads();
    ROWS_PHASE(pc, 2);
    // 4. one Gamma(count + gamma) draw per member
    const int n_act = E.ctl[kCtlNAct];
    for (int i = threadIdx.x; i < n_act; i += n_threads) {
        const int t = E.act[i];
        const double gv = gamma_variate(make_site(key, PH_GAMMA, iteration, (uint32_t)t, 0), (double)E.counts[t] + E.gamma);
        if (!(gv >= DBL_MIN)) E.ctl[kCtlErr] = 2;  // reference: assert(gamma_base_sample >= double_underflow) (expressionGenerator.cpp:115)
        E.theta[t] = gv;
    }
    __syncthreads();
    ROWS_PHASE(pc, 3);
    // 5. the normalising sum, in the order it always had (per lane over the chunks, then across the lanes); fresh counts
    if (warp == 0) {
        double part = 0.0;
        for (int c = 0; c < nch; c++) {
            const int t = c * 32 + lane;
            if (t < T) part += E.theta[t];
        }
        const double norm = warp_sum(part);
        if (lane == 0) *E.norm = norm;
    }
    __syncthreads();
    ROWS_PHASE(pc, 4);
    for (int t = threadIdx.x; t < T; t += n_threads) E.counts[t] = E.base[t];  // fresh CountValueContainer (+ the folded rows)
    const double norm = *E.norm;
    for (int i = threadIdx.x; i < n_act; i += n_threads) {
        const int t = E.act[i];
        const double g = E.theta[t];
        const double th = g != 0.0 ? g / norm : 0.0;  // valueContainer.cpp:216-224
        E.theta[t] = th;
        if (sampling && g != 0.0 && t % n_cta == rank) {
            const double fpkm = (th * (double)E.N_f * 1e9) / E.den[t];  // valueContainer.cpp:235
            const int o = ++E.occ[t];                                    // gibbsOutput.cpp:60
            if (o == 1) {
                E.mean[t] = fpkm;
            } else {
                const double delta = fpkm - E.mean[t];
                const double mu = E.mean[t] + delta / o;
                E.mean[t] = mu;
                E.m2[t] += delta * (fpkm - mu);
            }
        }
    }
    __syncthreads();
    ROWS_PHASE(pc, 5);
    return E.ctl[kCtlB];
}

// ------------------------------------------------------------------------------------------------------------------
__global__ void __maxnreg__(BAYES_ROWS_MAXNREG) gibbs_rows_kernel(const __grid_constant__ KernelArgs A, const int64_t unit_base) {
    extern __shared__ __align__(16) unsigned char smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int n_cta = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int64_t ui = unit_base + blockIdx.x / n_cta;
    const bool lead = rank == 0;
    if (lead && threadIdx.x == 0) A.unit_clock[3 * ui] = global_ns();
    const UnitDev &U = A.units[ui];
    const int T = U.T[0];
    const int n_threads = blockDim.x;
    const RowsLayout lay = rows_layout(T, n_cta, U.rows_max, U.cells_max, U.ccells_max, U.chain_max, U.qcap, U.pool_max, U.mat_smem != 0, U.scr_smem != 0,
                                       U.pool_smem != 0);
    RowsState E;
    E.theta = (double *)(smem + lay.theta);
    E.den = (double *)(smem + lay.den);
    E.mean = (double *)(smem + lay.mean);
    E.m2 = (double *)(smem + lay.m2);
    E.norm = (double *)(smem + lay.norm);
    E.counts = (int32_t *)(smem + lay.counts);
    E.occ = (int32_t *)(smem + lay.occ);
    int32_t *base = (int32_t *)(smem + lay.base);
    E.base = base;
    E.mask = (uint32_t *)(smem + lay.mask);
    E.actoff = (int32_t *)(smem + lay.actoff);
    E.act = (uint16_t *)(smem + lay.act);
    E.amask = (uint64_t *)(smem + lay.amask);
    E.tab_val = A.tab_val + U.tab_off;
    E.tab_row = A.tab_row + U.tabrow_off;
    E.T = T;
    E.nch = (T + 31) >> 5;
    E.N_f = U.N_f[0];
    E.gamma = U.gamma[0];
    int *ctl = (int *)(smem + lay.ctl);
    E.ctl = ctl;
    int32_t *part = (int32_t *)(smem + lay.part);  // two buffers of T partial counts (clusters of more than one CTA)
    const uint64_t key = U.key[0];

    const double n_total = (double)U.N_total[0];
    for (int t = threadIdx.x; t < T; t += n_threads) {
        const int32_t bc = A.base_counts[U.lane_off + t];
        base[t] = bc;
        E.counts[t] = bc;
        E.occ[t] = 0;
        E.mean[t] = 0.0;
        E.m2[t] = 0.0;
        E.theta[t] = 1.0;  // initAssignment runs the A-step on P_ij itself
        E.den[t] = A.efflen[U.lane_off + t] * n_total;  // valueContainer.cpp:235 denominator
        E.act[t] = (uint16_t)t;  // initAssignment: every candidate is "in the simplex"
        if (n_cta > 1) {
            part[t] = 0;
            part[T + t] = 0;
        }
    }
    for (int w = threadIdx.x; w < (T + 63) >> 6; w += n_threads)
        E.amask[w] = w == ((T + 63) >> 6) - 1 && (T & 63) ? (1ull << (T & 63)) - 1ull : ~0ull;
    if (threadIdx.x < kRowsCtlWords) ctl[threadIdx.x] = 0;

    // ---- this CTA's rows ----
    RowsView V;
    V.row_begin = U.cta_row0[rank];
    V.row_end = U.cta_row0[rank + 1];
    V.chain_end = V.row_begin + U.cta_chain[rank];
    V.n_words = (T + 63) >> 6;
    V.qcap = U.qcap;
    V.ctl = ctl;
    const int32_t *g_rcell = A.slab_cell_off + U.slab_off;
    const int cell_begin = g_rcell[V.row_begin], cell_end = g_rcell[V.row_end];
    if (U.mat_smem) {
        const int rows = V.row_end - V.row_begin;
        double *sv = (double *)(smem + lay.val);
        uint64_t *sm = (uint64_t *)(smem + lay.rmask);
        int32_t *rc = (int32_t *)(smem + lay.rcell), *rn = (int32_t *)(smem + lay.row_n), *rw = (int32_t *)(smem + lay.row_woff);
        uint32_t *rinfo = (uint32_t *)(smem + lay.row_info);
        const double *staged = stage_values_bulk(sv, A.val + U.cell_off + cell_begin, cell_end - cell_begin, (uint64_t *)(smem + lay.bar));
        block_copy(sm, (const uint64_t *)A.row_mask + U.mask_off + (size_t)V.row_begin * V.n_words, rows * V.n_words);
        block_copy(rc, g_rcell + V.row_begin, rows + 1);
        block_copy(rn, A.row_n + U.row_off + V.row_begin, rows);
        block_copy(rinfo, A.row_info + U.row_off + V.row_begin, rows);
        block_copy(rw, A.row_woff + U.woff_off + V.row_begin, rows);
        V.val = staged; V.rmask = sm; V.rcell = rc; V.row_n = rn; V.row_info = rinfo; V.row_woff = rw;
        V.row_base = V.row_begin;
        V.cell_base = cell_begin;
    } else {
        V.val = A.val + U.cell_off;
        V.rmask = (const uint64_t *)A.row_mask + U.mask_off;
        V.rcell = g_rcell;
        V.row_n = A.row_n + U.row_off;
        V.row_info = A.row_info + U.row_off;
        V.row_woff = A.row_woff + U.woff_off;
        V.row_base = 0;
        V.cell_base = 0;
    }
    if (U.scr_smem) {
        V.scr_cum = (double *)(smem + lay.scr_cum);
        V.scr_col = (uint16_t *)(smem + lay.scr_col);
        V.scr_base = cell_begin;
        V.scr_len = U.ccells_max;
        V.queue = (uint32_t *)(smem + lay.queue);
    } else {
        V.scr_cum = A.scr_cum + U.scr_off;
        V.scr_col = A.scr_col + U.scr_off;
        V.scr_base = 0;
        V.scr_len = U.n_cells;
        V.queue = A.queue + U.q_off + (size_t)rank * 6 * U.qcap;
    }
    uint32_t *pool = U.pool_smem ? (uint32_t *)(smem + lay.pool) : A.gpool + 4 * ((size_t)U.gpool_off + (size_t)rank * (U.pool_max + 1));
    const uint32_t *pool_desc = A.pool_desc + U.pooldesc_off + U.cta_pool0[rank];
    const int pool_blocks = U.cta_pool0[rank + 1] - U.cta_pool0[rank];
    if (threadIdx.x < 4) pool[4 * pool_blocks + threadIdx.x] = 0u;  // the spare block
    V.pool = pool;
    V.ch_rz = (double *)(smem + lay.ch_rz);
    V.ch_kk = (int *)(smem + lay.ch_kk);
    V.dtask = A.chain_tab + U.ctab_off + U.cta_dtask0[rank];
    V.n_dtask = U.cta_dtask0[rank + 1] - U.cta_dtask0[rank];

    const bool tracing = lead && A.trace_unit == ui;
    const int burn = U.burn;
    const int total = A.iter_cap > 0 && A.iter_cap < U.burn + U.samples ? A.iter_cap : U.burn + U.samples;
    const int stamp_at = A.iter_cap > 0 ? total >> 1 : -2;  // calibration pass: time the second half only
    int cur = 0;
    int32_t *acc = n_cta > 1 ? part : E.counts;
    __syncthreads();
#ifdef BAYES_ROWS_PHASES
    PhaseClock pclock;
    PhaseClock *pc = lead ? &pclock : nullptr;
    for (int i = 0; i < 10; i++) pclock.a[i] = 0;
    pclock.t = clock64();
#else
    PhaseClock *pc = nullptr;
#endif

    // it == -1 is initAssignment (gibbsSampler.cpp:90): the same A-step code on theta == 1, no E-step before it
    for (int it = -1; it < total; it++) {
        const uint32_t iter = (uint32_t)(it < 0 ? 0 : it);
        if (threadIdx.x == 0) {
            ctl[kCtlQ0] = 0;
            ctl[kCtlQ0 + 1] = 0;
            ctl[kCtlQ0 + 2] = 0;
            if (lead && it == stamp_at) A.unit_clock[3 * ui] = global_ns();
        }
        if (it >= 0) {
            if (tracing && it == 0)
                for (int t = threadIdx.x; t < T; t += n_threads) A.tr_init[t] = E.counts[t];
            if (tracing && it > 0 && it - 1 < A.trace_iters)
                for (int t = threadIdx.x; t < T; t += n_threads) A.tr_counts[(size_t)(it - 1) * T + t] = E.counts[t];
            const int b = rows_e_step(E, key, iter, it >= burn || (stamp_at >= 0 && it >= stamp_at), rank, n_cta, pool, pool_desc, pool_blocks, pc);
            if (tracing && it < A.trace_iters) {
                for (int t = threadIdx.x; t < T; t += n_threads) A.tr_theta[(size_t)it * T + t] = E.theta[t];
                if (threadIdx.x == 0) A.tr_b[it] = b;
            }
        } else {
            ctl[kCtlNAct] = T;
            rows_pool_gen(pool, pool_desc, pool_blocks, key, iter, true, 0, n_threads);
            __syncthreads();
        }
        ROWS_PHASE(pc, 9);
        rows_assign(V, E.amask, E.theta, acc);
        __syncthreads();
        ROWS_PHASE(pc, 6);
        rows_direct(V, acc, key, iter, it < 0);
        rows_tree_levels(V, acc, key, iter, it < 0);
        if (n_cta == 1) __syncthreads();  // the direct draws are in before the counts are read (a cluster meets at its barrier below)
        ROWS_PHASE(pc, 7);
        if (n_cta > 1) {
            // every CTA adds up the partial counts of all CTAs (distributed shared memory); only members of the simplex got any
            cluster.sync();
            // (one (member, CTA) pair per thread: the remote loads are all independent)
            const int n_act = ctl[kCtlNAct];
            const int sh = 31 - __clz(n_cta);
            for (int i = threadIdx.x; i < (n_act << sh); i += n_threads) {
                const int t = E.act[i >> sh];
                const int v = cluster.map_shared_rank(part, i & (n_cta - 1))[cur * T + t];
                if (v) atomicAdd(&E.counts[t], v);
            }
            // the other buffer was read by the other CTAs one iteration ago (before the barrier above): clear it for the next one
            cur ^= 1;
            for (int t = threadIdx.x; t < T; t += n_threads) part[cur * T + t] = 0;
            acc = part + cur * T;
            __syncthreads();
        }
        ROWS_PHASE(pc, 8);
    }
#ifdef BAYES_ROWS_PHASES
    if (lead && threadIdx.x == 0) {
        const double k = 1.0 / (1.965 * (total + 1));  // ns per iteration at 1965 MHz
        printf("rows phases, ns/iteration (T=%d, %d CTAs x %d threads): S+ %.0f | simplex+picks / pool %.0f | members %.0f | gamma %.0f | norm %.0f | theta+welford %.0f | "
               "rows %.0f | trees %.0f | cluster reduce %.0f | other %.0f\n", T, n_cta, n_threads, pclock.a[0] * k, pclock.a[1] * k, pclock.a[2] * k, pclock.a[3] * k,
               pclock.a[4] * k, pclock.a[5] * k, pclock.a[6] * k, pclock.a[7] * k, pclock.a[8] * k, pclock.a[9] * k);
    }
#endif
    if (tracing && total - 1 < A.trace_iters)
        for (int t = threadIdx.x; t < T; t += n_threads) A.tr_counts[(size_t)(total - 1) * T + t] = E.counts[t];

    const int64_t o = U.out_off[0];
    for (int t = threadIdx.x; t < T; t += n_threads) {
        if (t % n_cta == rank) {
            A.out_occ[o + t] = E.occ[t];
            A.out_mean[o + t] = E.mean[t];
            A.out_m2[o + t] = E.m2[t];
        }
        if (lead) A.out_counts[o + t] = E.counts[t];
    }
    int err = ctl[kCtlErr];
    if (n_cta > 1) {
        cluster.sync();
        if (lead && threadIdx.x == 0)
            for (int j = 1; j < n_cta; j++) {
                const int e = cluster.map_shared_rank(ctl, j)[kCtlErr];
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

// ---- single A-step in fast mode of a locus with more than 32 candidates: the same device code, one CTA, matrix, scratch and
// pool in global memory ----
template <bool INIT>
__global__ void step_assignment_rows_kernel(const __grid_constant__ KernelArgs A, const double *theta_in, uint32_t iteration, int32_t *counts_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const UnitDev &U = A.units[0];
    const int T = U.T[0];
    const int n_words = (T + 63) >> 6;
    double *theta = (double *)smem;
    uint64_t *amask = (uint64_t *)(theta + T);
    int32_t *counts = (int32_t *)(amask + n_words);
    int *ctl = (int *)(counts + T);
    int *ch_kk = ctl + kRowsCtlWords;
    double *ch_rz = (double *)(smem + ((12 * (size_t)T + 8 * (size_t)n_words + 4 * (size_t)kRowsCtlWords + 4 * (size_t)U.chain_max + 7) & ~(size_t)7));
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        theta[t] = INIT ? 1.0 : theta_in[t];
        counts[t] = A.base_counts[U.lane_off + t];
    }
    if (threadIdx.x < kRowsCtlWords) ctl[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x < 32)
        for (int w = 0; w < n_words; w++) {
            const int t0 = w * 64 + (int)threadIdx.x;
            const uint32_t lo = __ballot_sync(kFull, t0 < T && theta[t0] != 0.0);
            const uint32_t hi = __ballot_sync(kFull, t0 + 32 < T && theta[t0 + 32] != 0.0);
            if (threadIdx.x == 0) amask[w] = ((uint64_t)hi << 32) | lo;
        }
    const uint64_t key = U.key[0];
    RowsView V;
    V.row_begin = U.cta_row0[0];
    V.row_end = U.cta_row0[1];
    V.chain_end = V.row_begin + U.cta_chain[0];
    V.n_words = n_words;
    V.qcap = U.qcap;
    V.ctl = ctl;
    V.val = A.val + U.cell_off;
    V.rmask = (const uint64_t *)A.row_mask + U.mask_off;
    V.rcell = A.slab_cell_off + U.slab_off;
    V.row_n = A.row_n + U.row_off;
    V.row_info = A.row_info + U.row_off;
    V.row_woff = A.row_woff + U.woff_off;
    V.row_base = 0;
    V.cell_base = 0;
    V.scr_cum = A.scr_cum + U.scr_off;
    V.scr_col = A.scr_col + U.scr_off;
    V.scr_base = 0;
    V.scr_len = U.n_cells;
    V.queue = A.queue + U.q_off;
    uint32_t *pool = A.gpool + 4 * (size_t)U.gpool_off;
    const int pool_blocks = U.cta_pool0[1] - U.cta_pool0[0];
    rows_pool_gen(pool, A.pool_desc + U.pooldesc_off + U.cta_pool0[0], pool_blocks, key, iteration, INIT, 0, blockDim.x);
    if (threadIdx.x < 4) pool[4 * pool_blocks + threadIdx.x] = 0u;
    V.pool = pool;
    V.ch_rz = ch_rz;
    V.ch_kk = ch_kk;
    V.dtask = A.chain_tab + U.ctab_off + U.cta_dtask0[0];
    V.n_dtask = U.cta_dtask0[1] - U.cta_dtask0[0];
    __syncthreads();
    rows_assign(V, amask, theta, counts);
    __syncthreads();
    rows_direct(V, counts, key, iteration, INIT);
    rows_tree_levels(V, counts, key, iteration, INIT);
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) counts_out[t] = counts[t];
    if (threadIdx.x == 0) A.unit_err[0] = ctl[kCtlErr];
}

int check_rows(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return BAYES_OK;
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return BAYES_E_CUDA;
}

}  // namespace

int configure_rows_kernels() {
    int rc;
    if ((rc = check_rows(cudaFuncSetAttribute(gibbs_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_unit_smem()), "cudaFuncSetAttribute")))
        return rc;
    if ((rc = check_rows(cudaFuncSetAttribute(gibbs_rows_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1), "non-portable cluster size"))) return rc;
    if ((rc = check_rows(cudaFuncSetAttribute(step_assignment_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024), "cudaFuncSetAttribute")))
        return rc;
    return check_rows(cudaFuncSetAttribute(step_assignment_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024), "cudaFuncSetAttribute");
}

// Largest cluster the device can co-schedule for the rows kernel at `block` threads and `smem_bytes` of shared memory per CTA.
int rows_max_cluster(int block, size_t smem_bytes) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kMaxCluster, 1, 1);
    cfg.blockDim = dim3((unsigned)block, 1, 1);
    cfg.dynamicSmemBytes = smem_bytes;
    int size = 0;
    if (cudaOccupancyMaxPotentialClusterSize(&size, gibbs_rows_kernel, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    int p = 1;
    while (2 * p <= size && 2 * p <= kMaxCluster) p *= 2;
    return p;
}

int launch_units_rows(const KernelArgs &args, int64_t unit_base, int64_t n_units, int block, int n_cta, size_t smem_bytes, void *stream) {
    if (n_units <= 0) return BAYES_OK;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_units * n_cta), 1, 1);
    cfg.blockDim = dim3((unsigned)block, 1, 1);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)n_cta;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return check_rows(cudaLaunchKernelEx(&cfg, gibbs_rows_kernel, args, unit_base), "gibbs rows kernel launch");
}

int launch_step_assignment_rows(const KernelArgs &args, bool init, const double *d_theta, uint32_t iteration, int32_t *d_counts_out,
                                size_t smem_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (init) step_assignment_rows_kernel<true><<<1, 128, smem_bytes, st>>>(args, d_theta, iteration, d_counts_out);
    else step_assignment_rows_kernel<false><<<1, 128, smem_bytes, st>>>(args, d_theta, iteration, d_counts_out);
    return check_rows(cudaGetLastError(), "step_assignment_rows_kernel launch");
}

}  // namespace bayes
