// Persistent Gibbs sampler kernels for sm_100a: one thread block per UNIT for ALL of its iterations.
//
// A chain is 17 600 .. 671 000 strictly dependent iterations (gibbsSampler.cpp:103-141), so a launch per iteration
// would be launch-bound; instead a block keeps everything it needs in shared memory / registers and loops:
//   E  warp 0:  |S+| from the counts, simplex size b (1 uniform + CDF search, simplexSizeGenerator.cpp:214-224), the
//               b-|S+| null picks (expressionGenerator.cpp:104-127), one Gamma draw per member of the simplex, warp
//               shuffle reduction of the sum, theta = g / sum (expressionGenerator.cpp:77-132), FPKM + Welford update
//               of the posterior accumulators when sampling (valueContainer.cpp:226-239, gibbsOutput.cpp:51-75);
//   A  all threads, one thread per collapsed row, 32 rows of a SELL-32 slab per warp: Z = sum_j P_ij theta_j, then
//               either n < 20 uniforms located in the row's CDF or the chain of conditional binomials
//               (assignmentGenerator.cpp:263-373); counts accumulate with shared-memory atomics.
// Two unit shapes (internal.h):
//   bundle  up to 16 chains (of different loci) whose candidates sit side by side in the 32 lanes of warp 0; the
//           E-step is segmented by slot and keeps all per-candidate state in registers; rows of all slots are pooled,
//           so the data-dependent branches of the A-step find many more lanes taking the same path;
//   wide    one chain of a locus with more than 32 candidates; the E-step loops over 32-candidate chunks with the
//           per-candidate state in shared memory.
// The matrix, theta, counts and the CDF tables live in shared memory for the whole chain; units too large for that
// stream their cells from L2/HBM with fully coalesced slab-column reads.  No tensor cores: nothing here is a dense
// contraction.
#include <cuda_runtime.h>
#include <float.h>
#include <stdlib.h>

#include "internal.h"
#include "layout.h"
#include "rng.cuh"
#include "gibbs_common.cuh"

// Registers per thread: 72 keeps 28 warps per SM resident (internal.h: kWarpsPerSM) and removes most of the spills that 64
// (32 warps) forces into the inner loops.  Measured on config 2: 64 -> 2.21 s per step, 72 -> 2.01 s, 80 -> 2.09 s,
// 96 -> 2.16 s (gpurun_out/exp_regs.txt, profiles/README.md).
// Unroll factor of the row-sum loop: rows are short (a handful of cells), so the compiler's default 8-way unrolling with
// its remainder loop only costs instruction-cache space: 1 measured best on config 2 (8: 2.04 s, 4: 1.98 s, 2: 1.94 s, 1: 1.91 s).
#ifndef BAYES_Z_UNROLL
#define BAYES_Z_UNROLL 1
#endif
#ifndef BAYES_MAXNREG
#define BAYES_MAXNREG 72
#endif
// A wide unit is one locus with hundreds of candidates on 8-16 warps; a batch has few of them and they are limited by their
// own dependent chain, not by how many fit an SM: no register cap worth spilling for.
#ifndef BAYES_WIDE_MAXNREG
#define BAYES_WIDE_MAXNREG 128
#endif

namespace bayes {

namespace {

constexpr int kZUnroll = BAYES_Z_UNROLL;

// Read-only view of the packed matrix; pointers are shared-memory or global depending on the instantiation.
struct MatView {
    const double *val;
    const uint16_t *col;
    const int32_t *slab;     // n_slabs + 1 cell offsets
    const int32_t *row_n;
    const int32_t *row_nnz;
    const uint32_t *row_info;  // slot << 27 | original row index
    int n_slabs;
    int h;  // rows per slab (lanes >= h idle in the A-step)
    // wide units only: which candidates each row has a cell for, T bits per row as n_words 64-bit words, word w of the row
    // in lane l of slab s at row_mask[(s * n_words + w) * h + l] (global memory)
    const uint64_t *row_mask;
    int n_words;
};

// ------------------------------------------------------------------------------------------------------------------
// A-step.  One lane per collapsed row ("item", internal.h): a row with n >= 20 fragments runs the conditional binomials
// of assignmentGenerator.cpp:336-365, a row with n < 20 locates its n uniforms in the row's CDF (:289-334), one Philox
// block = four uniforms at a time.  Rows of one kind fill whole slabs and are sorted by length / block count, so the
// lanes of a warp run the same code about the same number of times.
// init = AssignmentGenerator::initAssignment (:213-260): binomial chain on P_ij itself for EVERY row, no row
// normalisation; the caller passes theta == 1 (P_ij * 1.0 is P_ij exactly) and Z is pinned to 1, so that the one copy of
// this code (instruction-cache footprint) serves both.
// ------------------------------------------------------------------------------------------------------------------

__device__ __forceinline__ void assign_rows(const MatView &M, const double *theta, int32_t *counts, const uint64_t *skey,
                                            int lanes_per_slot, uint32_t iteration, const bool init, int *next_slab, int32_t *err) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const uint32_t phase = init ? PH_INIT_ASSIGN : PH_ASSIGN;
    const int H = M.h;
    // Slabs are ordered heaviest first.  A one-warp unit walks them all; the warps of a larger unit take the next slab
    // from a shared counter as they become free (the draws are site-keyed and the counts integer atomics, so who
    // processes which slab does not change the result).
    const bool dynamic = next_slab != nullptr && n_warps > 1;
    for (int s = dynamic ? 0 : warp;; s += n_warps) {
        if (dynamic) {
            if (lane == 0) s = atomicAdd(next_slab, 1);
            s = __shfl_sync(0xffffffffu, s, 0);
        }
        if (s >= M.n_slabs) break;
        if (lane >= H) continue;
        const int r = s * H + lane;
        const uint32_t item = (uint32_t)M.row_nnz[r];
        const int nnz = (int)(item & kItemNnzMask);
        if (nnz == 0) continue;  // padding lane
        const bool chain = init || ((item >> kItemChainShift) & 1u);
        const int n = M.row_n[r];  // fragments of the row
        const uint32_t info = M.row_info[r];
        const uint32_t id = info & ((1u << kRowSlotShift) - 1u);
        const double *val = M.val + M.slab[s] + lane;
        const uint16_t *col = M.col + M.slab[s] + lane;

        // Z = sum_j P_ij theta_j (:280-282) and the first / last cell that can receive fragments.  A cell is in play
        // when its normalised probability is > 0 (CDF walk, :304-306) resp. >= double_underflow (chain, :344).
        double Z = 0.0;
        int kl = -1, in_play = 0;
#pragma unroll kZUnroll
        for (int k = 0; k < nnz; k++) {
            const double p = val[H * k] * theta[col[H * k]];
            Z += p;
            if (p != 0.0) {
                kl = k;
                in_play++;
            }
        }
        if (init) Z = 1.0;  // initAssignment uses P_ij as it is (:226)
        // Denormal-range cells (p < 1e-300 needs an input probability below ~1e-280, or a Dirichlet parameter < 1): the
        // packer flags rows that can have them, and only those pay for the exact rule "in play iff p / Z >=
        // double_underflow" with true divisions.
        bool tiny = false;
        if ((item >> kItemTinyShift) & 1u) {
            tiny = !(Z >= 1e-300);
            for (int k = 0; k < nnz; k++) {
                const double p = val[H * k] * theta[col[H * k]];
                tiny |= p != 0.0 && p < 1e-300;
            }
            if (chain && tiny) {
                kl = -1;
                in_play = 0;
                for (int k = 0; k < nnz; k++) {
                    const double p = val[H * k] * theta[col[H * k]];
                    if (p != 0.0 && p / Z >= DBL_MIN) {
                        kl = k;
                        in_play++;
                    }
                }
            }
        }
        if (kl < 0) {  // reference: assert(normalisation_constant >= double_underflow) (:284); reported as BAYES_E_NUMERIC
            *err = 1;
            continue;
        }
        const int c_last = col[H * kl];
        const double rZ = __drcp_rn(Z);

        // Whatever the cells before the last one in play do not take ends up in that last cell: in the chain the
        // reference asserts that nothing is left after it (:253, :363), and in the CDF walk its cumulative sum is 1 up
        // to rounding while every uniform is <= 1 - 2^-32.  So the last cell never needs a draw of its own.
        const uint32_t slot = info >> kRowSlotShift;
        const uint64_t key = skey[slot];
        if (!chain) {
            // :289-334.  The reference sorts the n uniforms and walks the CDF once; which column a uniform lands in does
            // not depend on the visiting order, so they are taken four at a time (one Philox block).
            int rem = n;
            if (in_play > 1) {  // with a single cell in play everything goes to it
                const Site site = make_site(key, phase, iteration, id, 0);
                const int blocks = (int)((item >> kItemBlockShift) & 7u);
                for (int j = 0; j < blocks; j++) {
                    const u32x4 w = site.block((uint32_t)j);
                    const int left = n - 4 * j < 4 ? n - 4 * j : 4;
                    double cum = 0.0;
                    int prev = 0;
                    // every lane first skips ahead to ITS next cell with a non-zero probability (theta is zero outside
                    // the current simplex) and only then do the lanes meet for the threshold arithmetic
                    for (int k = 0;; k++) {
                        int c = 0;
                        double p = 0.0;
                        for (; k < kl; k++) {
                            c = col[H * k];
                            p = val[H * k] * theta[c];
                            if (p != 0.0) break;
                        }
                        if (k >= kl) break;
                        cum += tiny ? p / Z : div_by(p, Z, rZ);                                   // :304
                        // u = w * 2^-32 < cum  <=>  w < ceil(cum * 2^32)  <=>  w <= ceil(cum * 2^32 - 1)   // :306
                        // (the conversion saturates, so cum >= 1 gives 0xffffffff: every word is below)
                        const uint32_t t = __double2uint_ru(__fma_rn(cum, 4294967296.0, -1.0));
                        int below = 0;
                        if (cum > 0.0)
                            below = (int)(w.x <= t) + ((left > 1) & (int)(w.y <= t)) + ((left > 2) & (int)(w.z <= t)) + ((left > 3) & (int)(w.w <= t));
                        if (below > prev) atomicAdd(&counts[c], below - prev);              // :323-324
                        rem -= below - prev;
                        prev = below;
                        if (below >= left) break;                                             // :327-331
                    }
                }
            }
            if (rem > 0) atomicAdd(&counts[c_last], rem);
        } else {
            // :336-365 (INIT: :222-251)
            // Every lane first skips ahead to ITS next cell in play and only then do the lanes meet again for the
            // binomial, so a warp runs the (expensive, divergent) draw once per in-play cell of its deepest row
            // instead of once per stored cell.
            int rem = n;
            if (in_play > 1) {
                const int col0 = (int)slot * lanes_per_slot;  // lane of the slot's candidate 0 (0 in a wide unit)
                double divider = 1.0;
                int k = 0;
                while (rem > 0) {
                    int c = 0;
                    double np = 0.0;
                    for (; k < kl; k++) {
                        c = col[H * k];
                        const double p = val[H * k] * theta[c];
                        if (p == 0.0) continue;
                        np = tiny ? p / Z : div_by(p, Z, rZ);
                        if (np >= DBL_MIN) break;
                        divider -= np;  // below double_underflow: no draw, but the divider still shrinks (:360)
                    }
                    if (k >= kl) break;
                    const Site site = make_site(key, phase, iteration, id, (uint32_t)(c - col0));
                    const int got = binomial_variate(site, rem, np / divider);
                    if (got > 0) atomicAdd(&counts[c], got);
                    rem -= got;
                    divider -= np;
                    k++;
                }
            }
            if (rem > 0) atomicAdd(&counts[c_last], rem);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// A-step of a wide unit.  With hundreds of candidates the expression vector is almost empty (the simplex holds a few tens
// of them), so nearly every product P_ij theta_j of a row is zero, and walking all cells of a row to find the few that
// are not -- twice per iteration, each step a dependent load from L2 -- is what a wide unit used to spend its time on
// (745 us per iteration on config 3).  Here every row carries a bit mask of its candidates; ANDed with the mask of the
// current simplex (wide_e_step leaves it in shared memory) it names the cells in play directly: cell k of a row is the
// k-th set bit of its mask.  Same arithmetic on the same cells in the same (ascending candidate) order as assign_rows,
// so the results are identical: the cells skipped are exactly those whose product is zero.
// ------------------------------------------------------------------------------------------------------------------
// The cells of one row whose product P_ij theta_j is not zero, in ascending candidate order.  The first walk (the row sum)
// scans the masks and remembers up to kRowCache cells with their products; every later walk of the same row (one per
// Philox block of a uniform item, one for a binomial chain) replays that list instead of scanning and gathering again.
// Rows with more cells in play (the first iterations of a chain, before the simplex has thinned out) are scanned each time.
constexpr int kRowCache = 24;
struct RowWalk {
    ActiveCells scan;
    const double *val, *theta;
    uint32_t *kc;  // k << 12 | c
    double *pv;
    int n, i;
    bool cached;
    __device__ __forceinline__ void start() {
        i = 0;
        if (!cached) scan.start();
    }
    __device__ __forceinline__ bool next(int &k, int &c, double &p) {
        if (cached) {
            if (i >= n) return false;
            k = (int)(kc[i] >> 12);
            c = (int)(kc[i] & 4095u);
            p = pv[i];
            i++;
            return true;
        }
        while (scan.next(k, c)) {
            p = val[scan.h * k] * theta[c];
            if (p != 0.0) return true;
        }
        return false;
    }
};

__device__ __forceinline__ void assign_rows_wide(const MatView &M, const uint64_t *amask, const double *theta, int32_t *counts, uint64_t key,
                                                 uint32_t iteration, const bool init, int *next_slab, int32_t *err) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const uint32_t phase = init ? PH_INIT_ASSIGN : PH_ASSIGN;
    const int H = M.h;
    const bool dynamic = next_slab != nullptr && n_warps > 1;
    uint32_t kc[kRowCache];
    double pv[kRowCache];
    for (int s = dynamic ? 0 : warp;; s += n_warps) {
        if (dynamic) {
            if (lane == 0) s = atomicAdd(next_slab, 1);
            s = __shfl_sync(0xffffffffu, s, 0);
        }
        if (s >= M.n_slabs) break;
        if (lane >= H) continue;
        const int r = s * H + lane;
        const uint32_t item = (uint32_t)M.row_nnz[r];
        if ((item & kItemNnzMask) == 0) continue;  // padding lane
        const bool chain = init || ((item >> kItemChainShift) & 1u);
        const int n = M.row_n[r];
        const uint32_t id = M.row_info[r] & ((1u << kRowSlotShift) - 1u);
        const double *val = M.val + M.slab[s] + lane;
        RowWalk cells;
        cells.scan.rm = M.row_mask + (size_t)s * M.n_words * H + lane;
        cells.scan.amask = amask;
        cells.scan.h = H;
        cells.scan.n_words = M.n_words;
        cells.val = val;
        cells.theta = theta;
        cells.kc = kc;
        cells.pv = pv;

        // Z, the last cell in play and how many there are (assign_rows: :280-282, :304-306, :344)
        double Z = 0.0, p;
        int kl = -1, c_last = 0, in_play = 0, k, c;
        cells.scan.start();
        while (cells.scan.next(k, c)) {
            p = val[H * k] * theta[c];
            Z += p;
            if (p != 0.0) {
                kl = k;
                c_last = c;
                if (in_play < kRowCache) {
                    kc[in_play] = ((uint32_t)k << 12) | (uint32_t)c;
                    pv[in_play] = p;
                }
                in_play++;
            }
        }
        cells.n = in_play;
        cells.cached = in_play <= kRowCache;
        if (init) Z = 1.0;
        bool tiny = false;
        if ((item >> kItemTinyShift) & 1u) {
            tiny = !(Z >= 1e-300);
            cells.start();
            while (cells.next(k, c, p)) tiny |= p < 1e-300;
            if (chain && tiny) {
                kl = -1;
                in_play = 0;
                cells.start();
                while (cells.next(k, c, p))
                    if (p / Z >= DBL_MIN) {
                        kl = k;
                        c_last = c;
                        in_play++;
                    }
            }
        }
        if (kl < 0) {
            *err = 1;
            continue;
        }
        const double rZ = __drcp_rn(Z);
        int rem = n;
        if (!chain) {
            if (in_play > 1) {
                const Site site = make_site(key, phase, iteration, id, 0);
                const int blocks = (int)((item >> kItemBlockShift) & 7u);
                for (int j = 0; j < blocks; j++) {
                    const u32x4 w = site.block((uint32_t)j);
                    const int left = n - 4 * j < 4 ? n - 4 * j : 4;
                    double cum = 0.0;
                    int prev = 0;
                    cells.start();
                    while (cells.next(k, c, p) && k < kl) {
                        cum += tiny ? p / Z : div_by(p, Z, rZ);
                        const uint32_t t = __double2uint_ru(__fma_rn(cum, 4294967296.0, -1.0));
                        int below = 0;
                        if (cum > 0.0)
                            below = (int)(w.x <= t) + ((left > 1) & (int)(w.y <= t)) + ((left > 2) & (int)(w.z <= t)) + ((left > 3) & (int)(w.w <= t));
                        if (below > prev) atomicAdd(&counts[c], below - prev);
                        rem -= below - prev;
                        prev = below;
                        if (below >= left) break;
                    }
                }
            }
        } else if (in_play > 1) {
            double divider = 1.0;
            cells.start();
            while (rem > 0 && cells.next(k, c, p) && k < kl) {
                const double np = tiny ? p / Z : div_by(p, Z, rZ);
                if (np >= DBL_MIN) {
                    const Site site = make_site(key, phase, iteration, id, (uint32_t)c);
                    const int got = binomial_variate(site, rem, np / divider);
                    if (got > 0) atomicAdd(&counts[c], got);
                    rem -= got;
                }
                divider -= np;  // below double_underflow: no draw, but the divider still shrinks (:360)
            }
        }
        if (rem > 0) atomicAdd(&counts[c_last], rem);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// E-step of a bundle (warp 0): lane = slot * Tpad + t.  Per-lane constants and the posterior accumulators live in
// registers for the whole chain.
// ------------------------------------------------------------------------------------------------------------------
struct BundleLane {
    // constants
    uint64_t key;        // Philox key of the lane's slot
    double gamma, den;   // den = efflen * N_total (valueContainer.cpp:235)
    const int32_t *trow; // CDF row offsets of the slot
    uint32_t segmask;    // lanes of the slot
    uint32_t validmask;  // lanes of the slot that hold a candidate
    int32_t Tg, t, base, N_f;
    int32_t half;        // Tpad / 2: first xor distance of the segmented sum
    bool valid;          // lane holds a candidate
    // state: GibbsOutput accumulators (gibbsOutput.cpp:40-75)
    int32_t occ;
    double mean, m2;
};

// ExpressionGenerator::generateExpression (expressionGenerator.cpp:77-132) for every slot of the bundle at once,
// + getFPKM / addSample when sampling.  Returns the simplex size used by the lane's slot.
__device__ __forceinline__ int bundle_e_step(BundleLane &L, double *theta, int32_t *counts, const double *tab_val, uint32_t iteration,
                                             bool sampling) {
    const int lane = threadIdx.x & 31;
    const int cnt = counts[lane];
    const uint32_t pm = __ballot_sync(0xffffffffu, L.valid && cnt > 0) & L.segmask;
    const int n_plus = __popc(pm);
    int b = 0, n_pick = 0;
    if (L.Tg > 0) {
        b = sample_simplex(tab_val, L.trow, n_plus, L.key, iteration);
        n_pick = b - n_plus;
    }
    // expressionGenerator.cpp:104-127: one uniform_int over the current ascending null list per extra slot.  The draws
    // themselves are independent (site = pick ordinal), so lane t of the slot makes the draw of pick t and only the
    // cheap selection below is sequential.
    uint32_t my_pick = 0u;
    if (L.valid && L.t < n_pick)
        my_pick = uniform_int(make_site(L.key, PH_NULLPICK, iteration, (uint32_t)L.t, 0), (uint32_t)(L.Tg - n_plus - L.t));
    uint32_t sel = 0u;
    const int max_pick = __reduce_max_sync(0xffffffffu, n_pick);
    const int lane0 = lane - L.t;  // first lane of the slot
    for (int k = 0; k < max_pick; k++) {
        const uint32_t s = __shfl_sync(0xffffffffu, my_pick, (lane0 + k) & 31);
        if (k < n_pick) {
            const uint32_t nullm = L.validmask & ~(pm | sel);
            sel |= 1u << __fns(nullm, 0, (int)s + 1);
        }
    }
    // :85-98 and :113: Gamma(count + gamma) on S+, Gamma(gamma) on the picked nulls
    const bool active = ((pm | sel) >> lane) & 1u;
    double g = 0.0;
    if (active) g = gamma_variate(make_site(L.key, PH_GAMMA, iteration, (uint32_t)L.t, 0), (double)cnt + L.gamma);
    double norm = g;
    for (int o = L.half; o > 0; o >>= 1) norm += __shfl_xor_sync(0xffffffffu, norm, o);
    const double th = active ? g / norm : 0.0;  // valueContainer.cpp:216-224
    theta[lane] = th;
    counts[lane] = L.base;  // fresh CountValueContainer for the next assignment (:266) + folded rows
    if (sampling && active) {
        const double fpkm = (th * (double)L.N_f * 1e9) / L.den;  // valueContainer.cpp:235
        const int o = ++L.occ;                                   // gibbsOutput.cpp:60
        if (o == 1) {
            L.mean = fpkm;
        } else {
            const double delta = fpkm - L.mean;
            L.mean += delta / o;
            L.m2 += delta * (fpkm - L.mean);
        }
    }
    return b;
}

template <bool MAT_SMEM>
__device__ __forceinline__ MatView stage_matrix(const KernelArgs &A, const UnitDev &U, const SmemLayout &lay, unsigned char *smem) {
    MatView M;
    M.n_slabs = U.n_slabs;
    M.h = U.slab_h;
    M.row_mask = (const uint64_t *)A.row_mask + U.mask_off;
    M.n_words = (U.Tpad + 63) >> 6;
    if (MAT_SMEM) {
        double *sv = (double *)(smem + lay.val);
        uint16_t *sc = (uint16_t *)(smem + lay.col);
        int32_t *ss = (int32_t *)(smem + lay.slab);
        int32_t *rn = (int32_t *)(smem + lay.row_n);
        int32_t *rz = (int32_t *)(smem + lay.row_nnz);
        uint32_t *ri = (uint32_t *)(smem + lay.row_info);
        block_copy(sv, A.val + U.cell_off, U.n_cells);
        block_copy(sc, A.col + U.cell_off, U.n_cells);
        block_copy(ss, A.slab_cell_off + U.slab_off, U.n_slabs + 1);
        block_copy(rn, A.row_n + U.row_off, U.n_slabs * U.slab_h);
        block_copy(rz, A.row_nnz + U.row_off, U.n_slabs * U.slab_h);
        block_copy(ri, A.row_info + U.row_off, U.n_slabs * U.slab_h);
        M.val = sv; M.col = sc; M.slab = ss; M.row_n = rn; M.row_nnz = rz; M.row_info = ri;
    } else {
        M.val = A.val + U.cell_off;
        M.col = A.col + U.cell_off;
        M.slab = A.slab_cell_off + U.slab_off;
        M.row_n = A.row_n + U.row_off;
        M.row_nnz = A.row_nnz + U.row_off;
        M.row_info = A.row_info + U.row_off;
    }
    return M;
}

// ------------------------------------------------------------------------------------------------------------------
template <bool MAT_SMEM>
__global__ void __maxnreg__(BAYES_MAXNREG) gibbs_bundle_kernel(const __grid_constant__ KernelArgs A, const int64_t unit_base) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int64_t ui = unit_base + blockIdx.x;
    if (threadIdx.x == 0) A.unit_clock[3 * ui] = global_ns();
    const UnitDev &U = A.units[ui];
    const int n_slots = U.n_slots, Tpad = U.Tpad;
    const SmemLayout lay = unit_layout(false, 32, n_slots, Tpad, U.n_slabs, U.slab_h, U.n_cells, U.tab_len, U.tab_smem != 0, MAT_SMEM);
    double *theta = (double *)(smem + lay.theta);
    int32_t *counts = (int32_t *)(smem + lay.counts);
    uint64_t *skey = (uint64_t *)(smem + lay.key);
    int *next_slab = (int *)(smem + lay.next_slab);
    const double *tab_val;
    const int32_t *tab_row;
    if (U.tab_smem) {
        double *tv = (double *)(smem + lay.tab_val);
        int32_t *tr = (int32_t *)(smem + lay.tab_row);
        block_copy(tv, A.tab_val + U.tab_off, U.tab_len);
        block_copy(tr, A.tab_row + U.tabrow_off, n_slots * (Tpad + 1));
        tab_val = tv;
        tab_row = tr;
    } else {
        tab_val = A.tab_val + U.tab_off;
        tab_row = A.tab_row + U.tabrow_off;
    }
    const MatView M = stage_matrix<MAT_SMEM>(A, U, lay, smem);
    if (threadIdx.x < (unsigned)n_slots) skey[threadIdx.x] = U.key[threadIdx.x];

    // lane constants of warp 0
    const int lane = threadIdx.x & 31;
    BundleLane L;
    int g = 0;
    {
        int sh = 1;
        while ((1 << sh) < Tpad) sh++;
        g = lane >> sh;
        L.t = lane & (Tpad - 1);
        const bool slot_ok = g < n_slots;
        L.Tg = slot_ok ? U.T[g] : 0;
        L.valid = L.t < L.Tg;
        L.key = slot_ok ? U.key[g] : 0ull;
        L.gamma = slot_ok ? U.gamma[g] : 1.0;
        L.N_f = slot_ok ? U.N_f[g] : 0;
        L.den = A.efflen[U.lane_off + lane] * (double)(slot_ok ? U.N_total[g] : 1);
        L.base = A.base_counts[U.lane_off + lane];
        L.trow = tab_row + (slot_ok ? g : 0) * (Tpad + 1);
        L.segmask = Tpad == 32 ? 0xffffffffu : (((1u << Tpad) - 1u) << (g * Tpad));
        L.validmask = L.Tg == 32 ? 0xffffffffu : (((1u << L.Tg) - 1u) << (g * Tpad));
        L.half = Tpad >> 1;
        L.occ = 0;
        L.mean = 0.0;
        L.m2 = 0.0;
    }
    if (threadIdx.x < 32) {
        theta[lane] = 1.0;  // initAssignment runs the A-step on P_ij itself
        counts[lane] = L.base;
    }

    const bool tracing = A.trace_unit == ui;
    const bool my_trace = tracing && threadIdx.x < 32 && g == A.trace_slot && L.valid;
    const int burn = U.burn;
    const int total = A.iter_cap > 0 && A.iter_cap < U.burn + U.samples ? A.iter_cap : U.burn + U.samples;
    const int stamp_at = A.iter_cap > 0 ? total >> 1 : -2;  // calibration pass: time the second half only

    // it == -1 is initAssignment (gibbsSampler.cpp:90): the same A-step code on theta == 1, no E-step before it
    for (int it = -1; it < total; it++) {
        if (threadIdx.x == 0) {
            *next_slab = 0;
            if (it == stamp_at) A.unit_clock[3 * ui] = global_ns();
        }
        if (it >= 0 && threadIdx.x < 32) {
            if (my_trace && it == 0) A.tr_init[L.t] = counts[lane];
            if (my_trace && it > 0 && it - 1 < A.trace_iters) A.tr_counts[(size_t)(it - 1) * L.Tg + L.t] = counts[lane];
            // (the timed half of a calibration pass runs as sampling iterations: nine tenths of a chain are)
            const int b = bundle_e_step(L, theta, counts, tab_val, (uint32_t)it, it >= burn || (stamp_at >= 0 && it >= stamp_at));
            if (my_trace && it < A.trace_iters) {
                A.tr_theta[(size_t)it * L.Tg + L.t] = theta[lane];
                if (L.t == 0) A.tr_b[it] = b;
            }
        }
        __syncthreads();
        assign_rows(M, theta, counts, skey, Tpad, (uint32_t)(it < 0 ? 0 : it), it < 0, next_slab, A.unit_err + ui);
        __syncthreads();
    }
    if (my_trace && total - 1 < A.trace_iters) A.tr_counts[(size_t)(total - 1) * L.Tg + L.t] = counts[lane];

    if (threadIdx.x < 32 && L.valid) {
        const int64_t o = U.out_off[g] + L.t;
        A.out_occ[o] = L.occ;
        A.out_mean[o] = L.mean;
        A.out_m2[o] = L.m2;
        A.out_counts[o] = counts[lane];
    }
    if (threadIdx.x == 0) {
        A.unit_clock[3 * ui + 1] = global_ns();
        A.unit_clock[3 * ui + 2] = sm_id();
    }
}

// ------------------------------------------------------------------------------------------------------------------
template <bool MAT_SMEM>
__global__ void __maxnreg__(BAYES_WIDE_MAXNREG) gibbs_wide_kernel(const __grid_constant__ KernelArgs A, const int64_t unit_base) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int64_t ui = unit_base + blockIdx.x;
    if (threadIdx.x == 0) A.unit_clock[3 * ui] = global_ns();
    const UnitDev &U = A.units[ui];
    const int T = U.T[0];
    const SmemLayout lay = unit_layout(true, T, 1, U.Tpad, U.n_slabs, U.slab_h, U.n_cells, U.tab_len, U.tab_smem != 0, MAT_SMEM);

    WideState E;
    E.theta = (double *)(smem + lay.theta);
    E.den = (double *)(smem + lay.den);
    E.mean = (double *)(smem + lay.mean);
    E.m2 = (double *)(smem + lay.m2);
    E.counts = (int32_t *)(smem + lay.counts);
    E.occ = (int32_t *)(smem + lay.occ);
    int32_t *base = (int32_t *)(smem + lay.base);
    E.base = base;
    E.mask = (uint32_t *)(smem + lay.mask);
    E.act = (uint16_t *)(smem + lay.act);
    E.amask = (uint64_t *)(smem + lay.amask);
    E.T = T;
    E.nch = (T + 31) >> 5;
    E.N_f = U.N_f[0];
    E.gamma = U.gamma[0];
    uint64_t *skey = (uint64_t *)(smem + lay.key);
    int *next_slab = (int *)(smem + lay.next_slab);
    const uint64_t key = U.key[0];
    if (threadIdx.x == 0) skey[0] = key;

    const double n_total = (double)U.N_total[0];
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        const int32_t bc = A.base_counts[U.lane_off + t];
        base[t] = bc;
        E.counts[t] = bc;
        E.occ[t] = 0;
        E.mean[t] = 0.0;
        E.m2[t] = 0.0;
        E.theta[t] = 1.0;  // initAssignment runs the A-step on P_ij itself
        E.den[t] = A.efflen[U.lane_off + t] * n_total;  // valueContainer.cpp:235 denominator
    }
    // initAssignment (theta == 1): every candidate is "in the simplex"
    for (int w = threadIdx.x; w < (T + 63) >> 6; w += blockDim.x)
        E.amask[w] = w == ((T + 63) >> 6) - 1 && (T & 63) ? (1ull << (T & 63)) - 1ull : ~0ull;
    if (U.tab_smem) {
        double *tv = (double *)(smem + lay.tab_val);
        int32_t *tr = (int32_t *)(smem + lay.tab_row);
        block_copy(tv, A.tab_val + U.tab_off, U.tab_len);
        block_copy(tr, A.tab_row + U.tabrow_off, T + 1);
        E.tab_val = tv;
        E.tab_row = tr;
    } else {
        E.tab_val = A.tab_val + U.tab_off;
        E.tab_row = A.tab_row + U.tabrow_off;
    }
    const MatView M = stage_matrix<MAT_SMEM>(A, U, lay, smem);

    const bool tracing = A.trace_unit == ui;
    const int burn = U.burn;
    const int total = A.iter_cap > 0 && A.iter_cap < U.burn + U.samples ? A.iter_cap : U.burn + U.samples;
    const int stamp_at = A.iter_cap > 0 ? total >> 1 : -2;  // calibration pass: time the second half only

    // it == -1 is initAssignment (gibbsSampler.cpp:90): the same A-step code on theta == 1, no E-step before it
    for (int it = -1; it < total; it++) {
        if (threadIdx.x == 0) {
            *next_slab = 0;
            if (it == stamp_at) A.unit_clock[3 * ui] = global_ns();
        }
        if (it >= 0 && threadIdx.x < 32) {
            const bool tr = tracing && it < A.trace_iters;
            if (tracing && it == 0)
                for (int t = threadIdx.x; t < T; t += 32) A.tr_init[t] = E.counts[t];
            if (tracing && it > 0 && it - 1 < A.trace_iters)
                for (int t = threadIdx.x; t < T; t += 32) A.tr_counts[(size_t)(it - 1) * T + t] = E.counts[t];
            const int b = wide_e_step(E, key, (uint32_t)it, it >= burn || (stamp_at >= 0 && it >= stamp_at), tr ? A.tr_theta + (size_t)it * T : nullptr);
            if (tr && threadIdx.x == 0) A.tr_b[it] = b;
        }
        __syncthreads();
        assign_rows_wide(M, E.amask, E.theta, E.counts, key, (uint32_t)(it < 0 ? 0 : it), it < 0, next_slab, A.unit_err + ui);
        __syncthreads();
    }
    if (tracing && total - 1 < A.trace_iters)
        for (int t = threadIdx.x; t < T; t += blockDim.x) A.tr_counts[(size_t)(total - 1) * T + t] = E.counts[t];

    const int64_t o = U.out_off[0];
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        A.out_occ[o + t] = E.occ[t];
        A.out_mean[o + t] = E.mean[t];
        A.out_m2[o + t] = E.m2[t];
        A.out_counts[o + t] = E.counts[t];
    }
    if (threadIdx.x == 0) {
        A.unit_clock[3 * ui + 1] = global_ns();
        A.unit_clock[3 * ui + 2] = sm_id();
    }
}

// ---- single-step kernels (same device code, one block) ------------------------------------------------------------
template <bool INIT>
__global__ void step_assignment_kernel(const __grid_constant__ KernelArgs A, const double *theta_in, uint32_t iteration, int32_t *counts_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const UnitDev &U = A.units[0];
    const int T = U.T[0];
    double *theta = (double *)smem;
    uint64_t *skey = (uint64_t *)(theta + T);
    // simplex mask of the given expression vector: candidates with theta != 0 (assign_rows_wide)
    const int n_words = (T + 63) >> 6;
    uint64_t *amask = (uint64_t *)(skey + 1);
    int32_t *counts = (int32_t *)(amask + n_words);
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        theta[t] = INIT ? 1.0 : theta_in[t];
        counts[t] = A.base_counts[U.lane_off + t];
    }
    if (threadIdx.x == 0) skey[0] = U.key[0];
    __syncthreads();
    if (threadIdx.x < 32)
        for (int w = 0; w < n_words; w++) {
            const int t0 = w * 64 + (int)threadIdx.x;
            const uint32_t lo = __ballot_sync(0xffffffffu, t0 < T && theta[t0] != 0.0);
            const uint32_t hi = __ballot_sync(0xffffffffu, t0 + 32 < T && theta[t0 + 32] != 0.0);
            if (threadIdx.x == 0) amask[w] = ((uint64_t)hi << 32) | lo;
        }
    MatView M;
    M.n_slabs = U.n_slabs;
    M.h = U.slab_h;
    M.row_mask = (const uint64_t *)A.row_mask + U.mask_off;
    M.n_words = n_words;
    M.val = A.val + U.cell_off;
    M.col = A.col + U.cell_off;
    M.slab = A.slab_cell_off + U.slab_off;
    M.row_n = A.row_n + U.row_off;
    M.row_nnz = A.row_nnz + U.row_off;
    M.row_info = A.row_info + U.row_off;
    __syncthreads();
    assign_rows_wide(M, amask, theta, counts, skey[0], iteration, INIT, nullptr, A.unit_err);
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) counts_out[t] = counts[t];
}

// generateExpression with a GIVEN simplex size b: a one-row CDF table ((b - |S+|) zeros, then 2.0) makes the upper bound
// of any uniform land on b.  T <= 32 runs the bundle E-step (one slot), larger T the wide E-step.
__global__ void step_expression_kernel(int T, const int32_t *counts_in, int b, double gamma, uint64_t key, uint32_t iteration, double *theta_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int nT = T <= 32 ? 32 : T;
    const int nch = (nT + 31) >> 5;
    double *theta = (double *)smem;
    double *tab = theta + nT;  // T + 2 entries
    int32_t *counts = (int32_t *)(tab + T + 2);
    int32_t *base = counts + nT;
    int32_t *tab_row = base + nT;  // T + 1
    uint32_t *mask = (uint32_t *)(tab_row + T + 1);
    const int lane = threadIdx.x;
    for (int t = lane; t < nT; t += 32) { counts[t] = t < T ? counts_in[t] : 0; base[t] = 0; theta[t] = 0.0; }
    __syncwarp();
    int n_plus = 0;
    for (int t = 0; t < T; t++) n_plus += counts[t] > 0;
    const int extra = b - n_plus;
    for (int t = lane; t <= T; t += 32) tab_row[t] = t >= n_plus ? extra + 1 : 0;
    for (int t = lane; t <= extra; t += 32) tab[t] = t == extra ? 2.0 : 0.0;
    __syncwarp();
    if (T <= 32) {
        BundleLane L;
        L.key = key; L.gamma = gamma; L.den = 1.0; L.trow = tab_row; L.segmask = 0xffffffffu;
        L.validmask = T == 32 ? 0xffffffffu : (1u << T) - 1u;
        L.Tg = T; L.t = lane; L.base = 0; L.N_f = 0; L.half = 16; L.valid = lane < T; L.occ = 0; L.mean = 0.0; L.m2 = 0.0;
        bundle_e_step(L, theta, counts, tab, iteration, false);
        __syncwarp();
        if (lane < T) theta_out[lane] = theta[lane];
    } else {
        WideState E;
        E.theta = theta; E.den = theta; E.mean = theta; E.m2 = theta;
        E.counts = counts; E.occ = counts; E.base = base; E.mask = mask; E.act = (uint16_t *)(mask + 2 * nch + 2);
        E.amask = (uint64_t *)(smem + ((8 * (nT + (size_t)T + 2) + 4 * (2 * nT + (size_t)T + 1) + 4 * (2 * (size_t)nch + 2) + 2 * nT + 15) & ~(size_t)7));
        E.tab_val = tab; E.tab_row = tab_row;
        E.T = T; E.nch = nch; E.N_f = 0; E.gamma = gamma;
        wide_e_step(E, key, iteration, false, theta_out);
    }
}

__global__ void step_variates_kernel(int kind, double a, double p, uint64_t key, int n, double *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (kind == 0) out[i] = gamma_variate(make_site(key, PH_GAMMA, 0, (uint32_t)i, 0), a);
    else if (kind == 1) out[i] = (double)binomial_variate(make_site(key, PH_ASSIGN, 0, (uint32_t)i, 0), (int)a, p);
    else out[i] = (double)uniform_int(make_site(key, PH_NULLPICK, 0, (uint32_t)i, 0), (uint32_t)a);
}

}  // namespace

static int check(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return BAYES_OK;
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return BAYES_E_CUDA;
}

int configure_kernels() {
    int rc;
    // Every sampler kernel asks for the same carve-out of the unified L1 / shared memory, so launches of different classes
    // can share an SM instead of waiting for it to drain and be reconfigured.  65 % (164 KB shared, 64 KB L1) measured
    // best on config 2: the registers spilled around the out-of-line RNG calls then hit L1 instead of L2 (100 %: 1.97 s
    // per step, 80 %: 1.90 s, 65 %: 1.86 s, 55 %: 2.01 s; gpurun_out / profiles/README.md).
    const char *co = getenv("BAYES_CARVEOUT");  // experiment knob, percent
    const int carveout = co ? atoi(co) : 65;
    if ((rc = check(cudaFuncSetAttribute(gibbs_bundle_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout), "carveout"))) return rc;
    if ((rc = check(cudaFuncSetAttribute(gibbs_bundle_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout), "carveout"))) return rc;
    if ((rc = check(cudaFuncSetAttribute(gibbs_wide_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout), "carveout"))) return rc;
    if ((rc = check(cudaFuncSetAttribute(gibbs_wide_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout), "carveout"))) return rc;
    if ((rc = check(cudaFuncSetAttribute(gibbs_bundle_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024), "cudaFuncSetAttribute"))) return rc;
    if ((rc = check(cudaFuncSetAttribute(gibbs_bundle_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024), "cudaFuncSetAttribute"))) return rc;
    if ((rc = check(cudaFuncSetAttribute(gibbs_wide_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024), "cudaFuncSetAttribute"))) return rc;
    if ((rc = check(cudaFuncSetAttribute(gibbs_wide_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024), "cudaFuncSetAttribute"))) return rc;
    if ((rc = check(cudaFuncSetAttribute(step_assignment_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024), "cudaFuncSetAttribute"))) return rc;
    return check(cudaFuncSetAttribute(step_assignment_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024), "cudaFuncSetAttribute");
}

int launch_units(const KernelArgs &args, int64_t unit_base, int64_t n_units, int block, bool wide, bool mat_smem, size_t smem_bytes,
                 void *stream) {
    if (n_units <= 0) return BAYES_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)n_units;
    if (wide) {
        if (mat_smem) gibbs_wide_kernel<true><<<grid, block, smem_bytes, st>>>(args, unit_base);
        else gibbs_wide_kernel<false><<<grid, block, smem_bytes, st>>>(args, unit_base);
    } else {
        if (mat_smem) gibbs_bundle_kernel<true><<<grid, block, smem_bytes, st>>>(args, unit_base);
        else gibbs_bundle_kernel<false><<<grid, block, smem_bytes, st>>>(args, unit_base);
    }
    return check(cudaGetLastError(), "gibbs kernel launch");
}

// One thread that waits `ns` nanoseconds: spaces the start of consecutive class launches (capi.cu, bayes_gibbs_run)
__global__ void stagger_kernel(unsigned ns) {
    const uint64_t t0 = global_ns();
    while (global_ns() - t0 < ns) {
    }
}

int launch_stagger(unsigned ns, void *stream) {
    stagger_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(ns);
    return check(cudaGetLastError(), "stagger_kernel launch");
}

int launch_step_assignment(const KernelArgs &args, bool init, const double *d_theta, uint32_t iteration, int32_t *d_counts_out,
                           size_t smem_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (init) step_assignment_kernel<true><<<1, 128, smem_bytes, st>>>(args, d_theta, iteration, d_counts_out);
    else step_assignment_kernel<false><<<1, 128, smem_bytes, st>>>(args, d_theta, iteration, d_counts_out);
    return check(cudaGetLastError(), "step_assignment_kernel launch");
}

int launch_step_expression(int T, const int32_t *d_counts, int b, double gamma, uint64_t key, uint32_t iteration, double *d_theta,
                           void *stream) {
    const size_t nT = T <= 32 ? 32 : (size_t)T;
    const size_t smem = 8 * (nT + (size_t)T + 2) + 4 * (2 * nT + (size_t)T + 1) + 4 * (2 * ((nT + 31) / 32) + 2) + 2 * nT + 32 + 8 * ((nT + 63) / 64 + 1);
    step_expression_kernel<<<1, 32, smem, (cudaStream_t)stream>>>(T, d_counts, b, gamma, key, iteration, d_theta);
    return check(cudaGetLastError(), "step_expression_kernel launch");
}

int launch_step_variates(int kind, double a, double p, uint64_t key, int n, double *d_out, void *stream) {
    step_variates_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(kind, a, p, key, n, d_out);
    return check(cudaGetLastError(), "step_variates_kernel launch");
}

}  // namespace bayes
