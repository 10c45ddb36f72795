// Device code shared by the replay-mode kernels (gibbs_kernel.cu) and the fast-mode kernels (gibbs_fast.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "rng.cuh"

namespace bayes {
namespace {

__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t sm_id() {
    uint32_t s;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    return s;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// p / Z, correctly rounded, from rZ = RN(1 / Z): quotient estimate, exact residual, one fused correction (Markstein).
// One reciprocal per row replaces a division per cell.
__device__ __forceinline__ double div_by(double p, double Z, double rZ) {
    const double q0 = p * rZ;
    return __fma_rn(__fma_rn(-q0, Z, p), rZ, q0);
}

struct ActiveCells {
    const uint64_t *rm;     // the row's mask words, stride h
    const uint64_t *amask;  // simplex mask, 64 candidates per word (shared memory, written by the E-step)
    int h, n_words, w, kbase;
    uint64_t m_all, m;
    __device__ __forceinline__ void start() {
        w = -1;
        kbase = 0;
        m_all = 0ull;
        m = 0ull;
    }
    // next cell of the row whose candidate is in the simplex: k = its index in the row, c = its candidate
    __device__ __forceinline__ bool next(int &k, int &c) {
        while (m == 0ull) {
            kbase += __popcll(m_all);
            if (++w >= n_words) return false;
            m_all = __ldg((const unsigned long long *)rm + (size_t)w * h);
            m = m_all & amask[w];
        }
        const int bit = __ffsll((long long)m) - 1;
        m &= m - 1ull;
        k = kbase + __popcll(m_all & ((1ull << bit) - 1ull));
        c = (w << 6) + bit;
        return true;
    }
};

// FixedBinomialFixedGammaSimplexSizeGenerator::sampleSimplexSize (simplexSizeGenerator.cpp:214-224): one uniform_01,
// std::upper_bound on the CDF row of |S+|, + |S+|
__device__ __forceinline__ int sample_simplex(const double *tab_val, const int32_t *tab_row, int n_plus, uint64_t key, uint32_t iteration) {
    const double u = u01(make_site(key, PH_SIMPLEX, iteration, 0, 0).block(0).x);
    if (n_plus < 1) n_plus = 1;  // only after fragments were lost to a flagged error (BAYES_E_NUMERIC): stay inside the table
    int lo = tab_row[n_plus - 1];
    const int first = lo;
    int hi = tab_row[n_plus];
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (u < tab_val[mid]) hi = mid; else lo = mid + 1;
    }
    return (lo - first) + n_plus;
}

// ------------------------------------------------------------------------------------------------------------------
// E-step of a wide unit (warp 0, 32 candidates per chunk, state in shared memory)
// ------------------------------------------------------------------------------------------------------------------
struct WideState {
    double *theta, *den, *mean, *m2;
    int32_t *counts, *occ;
    const int32_t *base;
    uint32_t *mask;  // [0, nch): S+ of the counts; [nch, 2 nch): picked null transcripts
    uint16_t *act;   // members of the simplex, ascending (scratch of the E-step)
    uint64_t *amask; // the simplex as a bit mask, 64 candidates per word: what the A-step ANDs the rows' masks with
    const double *tab_val;
    const int32_t *tab_row;
    int T, nch, N_f;
    double gamma;
};

__device__ __forceinline__ int wide_e_step(const WideState &E, uint64_t key, uint32_t iteration, bool sampling, double *tr_theta) {
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
    const int n_null = T - n_plus;
    const int n_pick = b - n_plus;
    // the draws are independent (site = pick ordinal): 32 at a time, one per lane; the selection is sequential
    for (int k0 = 0; k0 < n_pick; k0 += 32) {
        uint32_t my_pick = 0u;
        if (k0 + lane < n_pick)
            my_pick = uniform_int(make_site(key, PH_NULLPICK, iteration, (uint32_t)(k0 + lane), 0), (uint32_t)(n_null - k0 - lane));
        const int kn = n_pick - k0 < 32 ? n_pick - k0 : 32;
        for (int k = 0; k < kn; k++) {
            uint32_t rem = __shfl_sync(0xffffffffu, my_pick, k);
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
    }
    // The simplex holds b of the T candidates, a few per 32-candidate chunk: the draws and the divisions below run over
    // a compacted list of its members (b / 32 rounds with full lanes) instead of over every chunk with a few lanes each.
    // Sums keep the order they always had (per lane over the chunks, then across the lanes).
    int n_act = 0;
    for (int c = 0; c < nch; c++) {
        const int t = c * 32 + lane;
        const uint32_t m = E.mask[c] | E.mask[nch + c];
        if ((m >> lane) & 1u) E.act[n_act + __popc(m & ((1u << lane) - 1u))] = (uint16_t)t;
        n_act += __popc(m);
        if (t < T) E.theta[t] = 0.0;
    }
    __syncwarp();
    for (int i0 = 0; i0 < n_act; i0 += 32) {
        const int i = i0 + lane;
        if (i < n_act) {
            const int t = E.act[i];
            E.theta[t] = gamma_variate(make_site(key, PH_GAMMA, iteration, (uint32_t)t, 0), (double)E.counts[t] + E.gamma);
        }
    }
    __syncwarp();
    double part = 0.0;
    for (int c = 0; c < nch; c++) {
        const int t = c * 32 + lane;
        if (t < T) {
            part += E.theta[t];
            E.counts[t] = E.base[t];
        }
    }
    const double norm = warp_sum(part);
    for (int i0 = 0; i0 < n_act; i0 += 32) {
        const int i = i0 + lane;
        if (i < n_act) {
            const int t = E.act[i];
            const double g = E.theta[t];
            const double th = g != 0.0 ? g / norm : 0.0;
            E.theta[t] = th;
            if (sampling && g != 0.0) {
                const double fpkm = (th * (double)E.N_f * 1e9) / E.den[t];
                const int o = ++E.occ[t];
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
    for (int w = lane; w < (nch + 1) / 2; w += 32) {
        const uint32_t lo = E.mask[2 * w] | E.mask[nch + 2 * w];
        const uint32_t hi = 2 * w + 1 < nch ? E.mask[2 * w + 1] | E.mask[nch + 2 * w + 1] : 0u;
        E.amask[w] = ((uint64_t)hi << 32) | lo;
    }
    if (tr_theta) {
        __syncwarp();
        for (int t = lane; t < T; t += 32) tr_theta[t] = E.theta[t];
    }
    return b;
}

template <typename T>
__device__ __forceinline__ void block_copy(T *dst, const T *src, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

}  // namespace
}  // namespace bayes
