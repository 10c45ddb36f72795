// Persistent per-(locus, chain) Gibbs sampler for sm_100a.
//
// One thread block owns one chain of one locus for ALL of its iterations (17 600 .. 671 000 dependent iterations per
// locus; a launch per iteration would be launch-bound).  Per iteration (gibbsSampler.cpp:103-141):
//   E  warp 0:  |S+| from the counts, simplex size b (1 uniform + CDF search, simplexSizeGenerator.cpp:214-224), the
//               b-|S+| null picks (expressionGenerator.cpp:104-127), one Gamma draw per member of the simplex, warp
//               shuffle reduction of the sum, theta = g / sum (expressionGenerator.cpp:77-132), FPKM + Welford update
//               of the posterior accumulators when sampling (valueContainer.cpp:226-239, gibbsOutput.cpp:51-75);
//   A  all threads, one thread per collapsed row, 32 rows of a SELL-32 slab per warp: Z = sum_j P_ij theta_j, then
//               either n < 20 uniforms located in the row's CDF or the chain of conditional binomials
//               (assignmentGenerator.cpp:263-373); counts accumulate with shared-memory atomics.
// The locus' matrix, theta, counts, the CDF table and the accumulators live in shared memory for the whole chain;
// loci too large for that stream their cells from L2/HBM with fully coalesced slab-column reads.
// No tensor cores: there is no dense contraction anywhere on this path.
#include <cuda_runtime.h>
#include <float.h>

#include "internal.h"
#include "layout.h"
#include "rng.cuh"

namespace bayes {

namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Read-only view of the packed matrix; pointers are shared-memory or global depending on the instantiation.
struct MatView {
    const double *val;
    const uint16_t *col;
    const int32_t *slab;     // n_slabs + 1 cell offsets
    const int32_t *row_n;
    const int32_t *row_nnz;
    const int32_t *row_id;
    int n_slabs;
};

// ------------------------------------------------------------------------------------------------------------------
// A-step.  INIT = AssignmentGenerator::initAssignment (assignmentGenerator.cpp:213-260): binomial chain on P_ij itself,
// no row normalisation.  Otherwise generateAssignment (:263-373).
// ------------------------------------------------------------------------------------------------------------------
template <bool INIT>
__device__ __forceinline__ void assign_rows(const MatView &M, const double *theta, int32_t *counts, uint64_t key, uint32_t iteration) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const uint32_t phase = INIT ? PH_INIT_ASSIGN : PH_ASSIGN;
    for (int s = warp; s < M.n_slabs; s += n_warps) {
        const int r = s * 32 + lane;
        const int n = M.row_n[r];
        const int nnz = M.row_nnz[r];
        if (nnz == 0) continue;  // padding lane of the last slab
        const uint32_t id = (uint32_t)M.row_id[r];
        const double *val = M.val + M.slab[s] + lane;
        const uint16_t *col = M.col + M.slab[s] + lane;

        double Z = 1.0;
        if (!INIT) {
            Z = 0.0;
            for (int k = 0; k < nnz; k++) Z += val[32 * k] * theta[col[32 * k]];  // :280-282
        }

        if (!INIT && n < kSmallCount) {
            // :289-334.  The reference sorts the n uniforms and walks the CDF once; which column a uniform lands in
            // does not depend on the order they are visited in, so they are taken four at a time (one Philox block).
            const Site site = make_site(key, phase, iteration, id, 0);
            for (int j = 0; 4 * j < n; j++) {
                const u32x4 w = site.block((uint32_t)j);
                const int left = n - 4 * j;
                double cum = 0.0;
                int below_prev = 0;
                for (int k = 0; k < nnz; k++) {
                    const int c = col[32 * k];
                    cum += (val[32 * k] * theta[c]) / Z;                        // :304
                    // u = w * 2^-32 < cum  <=>  w < ceil(cum * 2^32)            // :306
                    const unsigned long long thr = __double2ull_ru(cum * 4294967296.0);
                    int below = (int)((unsigned long long)w.x < thr);
                    below += (left > 1) & (int)((unsigned long long)w.y < thr);
                    below += (left > 2) & (int)((unsigned long long)w.z < thr);
                    below += (left > 3) & (int)((unsigned long long)w.w < thr);
                    const int got = below - below_prev;
                    if (got > 0) atomicAdd(&counts[c], got);                     // :323-324
                    below_prev = below;
                    if (below >= (left < 4 ? left : 4)) break;                   // :327-331
                }
            }
        } else {
            // :336-365 (INIT: :222-251)
            double divider = 1.0;
            int rem = n, last = -1;
            for (int k = 0; k < nnz && rem > 0; k++) {
                const int c = col[32 * k];
                const double np = INIT ? val[32 * k] : (val[32 * k] * theta[c]) / Z;
                if (np >= DBL_MIN) {
                    const Site site = make_site(key, phase, iteration, id, (uint32_t)c);
                    const int got = binomial_variate(site, rem, np / divider);
                    if (got > 0) atomicAdd(&counts[c], got);
                    rem -= got;
                    last = c;
                }
                divider -= np;
            }
            // the reference asserts nothing is left (:253, :363); keep the total conserved instead of aborting
            if (rem > 0 && last >= 0) atomicAdd(&counts[last], rem);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// E-step pieces (warp 0 only; all 32 lanes execute the scalar parts redundantly, so nothing is broadcast)
// ------------------------------------------------------------------------------------------------------------------
// FixedBinomialFixedGammaSimplexSizeGenerator::sampleSimplexSize (simplexSizeGenerator.cpp:214-224)
__device__ __forceinline__ int sample_simplex(const double *tab_val, const int32_t *tab_row, int n_plus, uint64_t key, uint32_t iteration) {
    const double u = u01(make_site(key, PH_SIMPLEX, iteration, 0, 0).block(0).x);
    int lo = tab_row[n_plus - 1];
    const int first = lo;
    int hi = tab_row[n_plus];
    while (lo < hi) {  // std::upper_bound: first entry > u
        const int mid = (lo + hi) >> 1;
        if (u < tab_val[mid]) hi = mid; else lo = mid + 1;
    }
    return (lo - first) + n_plus;
}

struct EState {
    double *theta, *den, *mean, *m2;
    int32_t *counts, *occ;
    const int32_t *base;
    uint32_t *mask;  // [0, nch): S+ of the counts; [nch, 2 nch): picked null transcripts
    const double *tab_val;
    const int32_t *tab_row;
    int T, nch, N_f;
    double gamma;
};

// ExpressionGenerator::generateExpression (expressionGenerator.cpp:77-132) + getFPKM / addSample when sampling.
// Returns the simplex size used (for the trace).
__device__ __forceinline__ int e_step(const EState &E, uint64_t key, uint32_t iteration, bool sampling, double *tr_theta) {
    const int lane = threadIdx.x & 31;
    const int T = E.T, nch = E.nch;
    int n_plus = 0;
    for (int c = 0; c < nch; c++) {
        const int t = c * 32 + lane;
        const uint32_t pm = __ballot_sync(0xffffffffu, t < T && E.counts[t] > 0);
        if (lane == 0) { E.mask[c] = pm; E.mask[nch + c] = 0u; }
        n_plus += __popc(pm);
    }
    __syncwarp();
    const int b = sample_simplex(E.tab_val, E.tab_row, n_plus, key, iteration);

    // expressionGenerator.cpp:104-127: one uniform_int over the current ascending null list per extra slot
    const int n_null = T - n_plus;
    for (int k = 0; k < b - n_plus; k++) {
        uint32_t rem = uniform_int(make_site(key, PH_NULLPICK, iteration, (uint32_t)k, 0), (uint32_t)(n_null - k));
        int pick = -1;
        for (int c = 0; c < nch; c++) {
            uint32_t w = ~(E.mask[c] | E.mask[nch + c]);
            if (c == nch - 1 && (T & 31)) w &= (1u << (T & 31)) - 1u;
            const uint32_t pc = (uint32_t)__popc(w);
            if (rem < pc) { pick = c * 32 + (int)__fns(w, 0, (int)rem + 1); break; }
            rem -= pc;
        }
        __syncwarp();
        if (lane == 0 && pick >= 0) E.mask[nch + (pick >> 5)] |= 1u << (pick & 31);
        __syncwarp();
    }

    // :85-98 and :113: Gamma(count + gamma) on S+, Gamma(gamma) on the picked nulls
    double part = 0.0;
    for (int c = 0; c < nch; c++) {
        const int t = c * 32 + lane;
        const bool active = ((E.mask[c] | E.mask[nch + c]) >> lane) & 1u;
        double g = 0.0;
        if (active) g = gamma_variate(make_site(key, PH_GAMMA, iteration, (uint32_t)t, 0), (double)E.counts[t] + E.gamma);
        if (t < T) E.theta[t] = g;
        part += g;
    }
    const double norm = warp_sum(part);
    for (int c = 0; c < nch; c++) {
        const int t = c * 32 + lane;
        if (t < T) {
            const double g = E.theta[t];
            const double th = g != 0.0 ? g / norm : 0.0;  // valueContainer.cpp:216-224
            E.theta[t] = th;
            E.counts[t] = E.base[t];  // fresh CountValueContainer for the next assignment (:266) + folded rows
            if (tr_theta) tr_theta[t] = th;
            if (sampling && g != 0.0) {
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
    }
    return b;
}

template <typename T>
__device__ __forceinline__ void block_copy(T *dst, const T *src, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------------------------------
template <bool MAT_SMEM>
__global__ void __launch_bounds__(kMaxBlock) gibbs_chain_kernel(const __grid_constant__ KernelArgs A, const int64_t job_base) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int64_t job_idx = job_base + blockIdx.x;
    const JobDev job = A.jobs[job_idx];
    const LocusDev L = A.loci[job.locus];
    const int T = L.T;
    const SmemLayout lay = smem_layout(T, L.n_slabs, L.n_cells, L.tab_len, L.tab_smem != 0, MAT_SMEM);

    EState E;
    E.theta = (double *)(smem + lay.theta);
    E.den = (double *)(smem + lay.den);
    E.mean = (double *)(smem + lay.mean);
    E.m2 = (double *)(smem + lay.m2);
    E.counts = (int32_t *)(smem + lay.counts);
    E.occ = (int32_t *)(smem + lay.occ);
    int32_t *base = (int32_t *)(smem + lay.base);
    E.base = base;
    E.mask = (uint32_t *)(smem + lay.mask);
    E.T = T;
    E.nch = (T + 31) >> 5;
    E.N_f = L.N_f;
    E.gamma = L.gamma;

    // stage the locus
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        const int32_t bc = A.base_counts[L.t_off + t];
        base[t] = bc;
        E.counts[t] = bc;
        E.occ[t] = 0;
        E.mean[t] = 0.0;
        E.m2[t] = 0.0;
        E.theta[t] = 0.0;
        E.den[t] = A.efflen[L.t_off + t] * (double)L.N_total;  // valueContainer.cpp:235 denominator
    }
    if (L.tab_smem) {
        double *tv = (double *)(smem + lay.tab_val);
        int32_t *tr = (int32_t *)(smem + lay.tab_row);
        block_copy(tv, A.tab_val + L.tab_off, L.tab_len);
        block_copy(tr, A.tab_row + L.tabrow_off, T + 1);
        E.tab_val = tv;
        E.tab_row = tr;
    } else {
        E.tab_val = A.tab_val + L.tab_off;
        E.tab_row = A.tab_row + L.tabrow_off;
    }
    MatView M;
    M.n_slabs = L.n_slabs;
    if (MAT_SMEM) {
        double *sv = (double *)(smem + lay.val);
        uint16_t *sc = (uint16_t *)(smem + lay.col);
        int32_t *ss = (int32_t *)(smem + lay.slab);
        int32_t *rn = (int32_t *)(smem + lay.row_n);
        int32_t *rz = (int32_t *)(smem + lay.row_nnz);
        int32_t *ri = (int32_t *)(smem + lay.row_id);
        block_copy(sv, A.val + L.cell_off, L.n_cells);
        block_copy(sc, A.col + L.cell_off, L.n_cells);
        block_copy(ss, A.slab_cell_off + L.slab_off, L.n_slabs + 1);
        block_copy(rn, A.row_n + L.row_off, L.n_slabs * 32);
        block_copy(rz, A.row_nnz + L.row_off, L.n_slabs * 32);
        block_copy(ri, A.row_id + L.row_off, L.n_slabs * 32);
        M.val = sv; M.col = sc; M.slab = ss; M.row_n = rn; M.row_nnz = rz; M.row_id = ri;
    } else {
        M.val = A.val + L.cell_off;
        M.col = A.col + L.cell_off;
        M.slab = A.slab_cell_off + L.slab_off;
        M.row_n = A.row_n + L.row_off;
        M.row_nnz = A.row_nnz + L.row_off;
        M.row_id = A.row_id + L.row_off;
    }
    __syncthreads();

    const bool tracing = A.trace_job == job_idx;
    const int total = L.burn + L.samples;

    assign_rows<true>(M, E.theta, E.counts, job.key, 0u);  // gibbsSampler.cpp:90
    __syncthreads();
    if (tracing)
        for (int t = threadIdx.x; t < T; t += blockDim.x) A.tr_init[t] = E.counts[t];

    for (int it = 0; it < total; it++) {
        if (threadIdx.x < 32) {
            const bool tr = tracing && it < A.trace_iters;
            if (tracing && it > 0 && it - 1 < A.trace_iters)
                for (int t = threadIdx.x; t < T; t += 32) A.tr_counts[(size_t)(it - 1) * T + t] = E.counts[t];
            const int b = e_step(E, job.key, (uint32_t)it, it >= L.burn, tr ? A.tr_theta + (size_t)it * T : nullptr);
            if (tr && threadIdx.x == 0) A.tr_b[it] = b;
        }
        __syncthreads();
        assign_rows<false>(M, E.theta, E.counts, job.key, (uint32_t)it);
        __syncthreads();
    }
    if (tracing && total - 1 < A.trace_iters)
        for (int t = threadIdx.x; t < T; t += blockDim.x) A.tr_counts[(size_t)(total - 1) * T + t] = E.counts[t];

    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        A.out_occ[job.out_off + t] = E.occ[t];
        A.out_mean[job.out_off + t] = E.mean[t];
        A.out_m2[job.out_off + t] = E.m2[t];
        A.out_counts[job.out_off + t] = E.counts[t];
    }
}

// ---- single-step kernels (same device code, one block) ------------------------------------------------------------
template <bool INIT>
__global__ void step_assignment_kernel(const __grid_constant__ KernelArgs A, const double *theta_in, uint32_t iteration, int32_t *counts_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const LocusDev L = A.loci[0];
    const JobDev job = A.jobs[0];
    double *theta = (double *)smem;
    int32_t *counts = (int32_t *)(theta + L.T);
    for (int t = threadIdx.x; t < L.T; t += blockDim.x) {
        theta[t] = INIT ? 0.0 : theta_in[t];
        counts[t] = A.base_counts[L.t_off + t];
    }
    MatView M;
    M.n_slabs = L.n_slabs;
    M.val = A.val + L.cell_off;
    M.col = A.col + L.cell_off;
    M.slab = A.slab_cell_off + L.slab_off;
    M.row_n = A.row_n + L.row_off;
    M.row_nnz = A.row_nnz + L.row_off;
    M.row_id = A.row_id + L.row_off;
    __syncthreads();
    assign_rows<INIT>(M, theta, counts, job.key, iteration);
    __syncthreads();
    for (int t = threadIdx.x; t < L.T; t += blockDim.x) counts_out[t] = counts[t];
}

__global__ void step_expression_kernel(int T, const int32_t *counts_in, int b, double gamma, uint64_t key, uint32_t iteration, double *theta_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int nch = (T + 31) >> 5;
    double *theta = (double *)smem;
    double *tab = theta + T;      // T + 2 entries: a one-row CDF table that forces the simplex size to b
    int32_t *counts = (int32_t *)(tab + T + 2);
    int32_t *base = counts + T;
    int32_t *tab_row = base + T;  // T + 1
    uint32_t *mask = (uint32_t *)(tab_row + T + 1);
    const int lane = threadIdx.x;
    for (int t = lane; t < T; t += 32) { counts[t] = counts_in[t]; base[t] = 0; theta[t] = 0.0; }
    __syncwarp();
    int n_plus = 0;
    for (int t = 0; t < T; t++) n_plus += counts[t] > 0;
    // row |S+|-1 = (b - |S+|) zeros, then 2.0: the upper bound of any u in [0,1) is at offset b - |S+|
    const int extra = b - n_plus;
    for (int t = lane; t <= T; t += 32) tab_row[t] = t >= n_plus ? extra + 1 : 0;
    for (int t = lane; t <= extra; t += 32) tab[t] = t == extra ? 2.0 : 0.0;
    __syncwarp();
    EState E;
    E.theta = theta; E.den = theta; E.mean = theta; E.m2 = theta;
    E.counts = counts; E.occ = counts; E.base = base; E.mask = mask;
    E.tab_val = tab; E.tab_row = tab_row;
    E.T = T; E.nch = nch; E.N_f = 0; E.gamma = gamma;
    e_step(E, key, iteration, false, theta_out);
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
    int rc = check(cudaFuncSetAttribute(gibbs_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024), "cudaFuncSetAttribute");
    if (rc) return rc;
    rc = check(cudaFuncSetAttribute(gibbs_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024), "cudaFuncSetAttribute");
    if (rc) return rc;
    rc = check(cudaFuncSetAttribute(step_assignment_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024), "cudaFuncSetAttribute");
    if (rc) return rc;
    return check(cudaFuncSetAttribute(step_assignment_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024), "cudaFuncSetAttribute");
}

int launch_gibbs(const KernelArgs &args, int64_t job_base, int64_t n_jobs, int block, bool mat_smem, size_t smem_bytes, void *stream) {
    if (n_jobs <= 0) return BAYES_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (mat_smem) gibbs_chain_kernel<true><<<(unsigned)n_jobs, block, smem_bytes, st>>>(args, job_base);
    else gibbs_chain_kernel<false><<<(unsigned)n_jobs, block, smem_bytes, st>>>(args, job_base);
    return check(cudaGetLastError(), "gibbs_chain_kernel launch");
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
    const size_t smem = 8 * (2 * (size_t)T + 2) + 4 * (3 * (size_t)T + 1) + 4 * (2 * (((size_t)T + 31) / 32) + 2) + 16;
    step_expression_kernel<<<1, 32, smem, (cudaStream_t)stream>>>(T, d_counts, b, gamma, key, iteration, d_theta);
    return check(cudaGetLastError(), "step_expression_kernel launch");
}

int launch_step_variates(int kind, double a, double p, uint64_t key, int n, double *d_out, void *stream) {
    step_variates_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(kind, a, p, key, n, d_out);
    return check(cudaGetLastError(), "step_variates_kernel launch");
}

}  // namespace bayes
