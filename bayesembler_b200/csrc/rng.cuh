// Counter-based random numbers and uniform->variate transforms of the B200 sampler.
//
// The reference draws through Boost.Random on one sequential mt19937 per locus (assignmentGenerator.cpp:37-40,
// expressionGenerator.cpp:37-40, simplexSizeGenerator.cpp:39-42).  A sequential stream cannot feed thousands of
// threads, so every draw SITE owns a Philox4x32-10 stream instead (Salmon et al., SC'11):
//     block j of site (phase, iteration, index, sub) = philox(key, {(sub << 12) | j, index, iteration, phase})
// and the transforms are the published algorithms spelled out in DESIGN.md ("Variates"): Lemire multiply-shift
// uniform_int, Marsaglia-Tsang gamma over a Box-Muller normal, BINV / BTRD binomial.  The CPU oracle implements the
// same specification independently (oracle/orc_rng.h); parity tests compare the two draw for draw.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define BAYES_HD __host__ __device__ __forceinline__
#define BAYES_HD_NOINLINE inline __host__ __device__ __noinline__
#else
#define BAYES_HD inline
#define BAYES_HD_NOINLINE inline
#endif

namespace bayes {

enum Phase : uint32_t {
    PH_COVER = 1,        // calcMinimumReadCover tie break, index = greedy round
    PH_INIT_ASSIGN = 2,  // initAssignment chain, index = row, sub = column
    PH_SIMPLEX = 3,      // simplex-size uniform of the iteration that uses b
    PH_GAMMA = 4,        // expression gamma draw, index = transcript
    PH_NULLPICK = 5,     // null-transcript pick, index = pick ordinal
    PH_ASSIGN = 6,       // generateAssignment, index = row, sub = column (chain) or 0 (uniforms)
    // fast (production) sampling mode, gibbs_fast.cu
    PH_POOL = 7,         // generateAssignment, words of the rows with < 20 fragments: block b of the iteration = block (b & 4095) of index (b >> 12)
    PH_INIT_POOL = 8,    // the same for initAssignment
    PH_TREE = 9,         // generateAssignment, one binomial per node of a row's splitting tree: index = row, phase word = PH_TREE | node << 8
    PH_INIT_TREE = 10,   // the same for initAssignment
    PH_DIRECT = 11,      // generateAssignment, rows with >= 20 fragments drawn fragment by fragment: index = row, word f of the row = word (f & 3) of block (f >> 2)
    PH_INIT_DIRECT = 12  // the same for initAssignment
};

struct u32x4 {
    uint32_t x, y, z, w;
};

BAYES_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// Out of line on purpose (as are the log, the gamma and the binomial below): the sampler's hot loop has to fit the
// 32 KB instruction cache, and six inlined copies of the ten rounds alone were 7 KB.
BAYES_HD_NOINLINE u32x4 philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    u32x4 o = {c0, c1, c2, c3};
    return o;
}

// Draw site: key + the three fixed counter words; block(j) yields 4 fresh words.
struct Site {
    uint32_t k0, k1, sub12, index, iteration, phase;
    BAYES_HD u32x4 block(uint32_t j) const { return philox4x32_10(k0, k1, sub12 | j, index, iteration, phase); }
};

BAYES_HD Site make_site(uint64_t key, uint32_t phase, uint32_t iteration, uint32_t index, uint32_t sub) {
    Site s = {(uint32_t)key, (uint32_t)(key >> 32), sub << 12, index, iteration, phase};
    return s;
}

// uniform_01<mt19937*>: one 32-bit word -> [0,1)
BAYES_HD double u01(uint32_t w) { return (double)w * 2.3283064365386963e-10; }
BAYES_HD double u32_open(uint32_t w) { return ((double)w + 0.5) * 2.3283064365386963e-10; }
BAYES_HD double u53(uint32_t hi, uint32_t lo) {
    const uint64_t x = (((uint64_t)hi << 32) | lo) >> 11;
    return ((double)x + 0.5) * 1.1102230246251565e-16;
}

// uniform_int_distribution(0, n-1): words 0,1,2,... of the site (unbiased multiply-shift)
BAYES_HD uint32_t uniform_int(const Site &s, uint32_t n) {
    if (n <= 1) return 0;
    uint32_t k = 0;
    u32x4 b = s.block(0);
    uint64_t m = (uint64_t)b.x * n;
    uint32_t l = (uint32_t)m;
    if (l < n) {
        const uint32_t t = (0u - n) % n;
        while (l < t) {
            k++;
            if ((k & 3u) == 0) b = s.block(k >> 2);
            const uint32_t w = (k & 3u) == 1 ? b.y : (k & 3u) == 2 ? b.z : (k & 3u) == 3 ? b.w : b.x;
            m = (uint64_t)w * n;
            l = (uint32_t)m;
        }
    }
    return (uint32_t)(m >> 32);
}

// one shared copy of the double-precision log (about 100 instructions when inlined)
BAYES_HD_NOINLINE double dlog(double x) { return log(x); }

BAYES_HD double cos_2pi(double f) {
#if defined(__CUDA_ARCH__)
    return cospi(2.0 * f);
#else
    return cos(6.283185307179586 * f);
#endif
}

// gamma_distribution(alpha, 1): one block per attempt (expressionGenerator.cpp:89-92, 100-113)
BAYES_HD_NOINLINE double gamma_variate(const Site &s, double alpha) {
    uint32_t j = 0;
    if (alpha == 1.0) {
        const u32x4 b = s.block(0);
        return -dlog(u53(b.x, b.y));
    }
    double boost = 1.0, a = alpha;
    if (alpha < 1.0) {
        const u32x4 b = s.block(j++);
        boost = exp(dlog(u53(b.x, b.y)) / alpha);
        a = alpha + 1.0;
    }
    const double d = a - 1.0 / 3.0;
    const double c = 1.0 / sqrt(9.0 * d);
    for (;;) {
        const u32x4 b = s.block(j++);
        const double u1 = u53(b.x, b.y);
        const double x = sqrt(-2.0 * dlog(u1)) * cos_2pi(u01(b.z));
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        const double u = u32_open(b.w);
        const double x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2) return d * v * boost;
        if (dlog(u) < 0.5 * x2 + d * (1.0 - v + dlog(v))) return d * v * boost;
    }
}

BAYES_HD double stirling_fc(int k) {
    if (k < 10) {
        // 1/(12(k+1)) - 1/(360(k+1)^3) + ... is not accurate enough below 10: tabulated (Hoermann 1993, table 1)
        const double t0 = 0.08106146679532726, t1 = 0.04134069595540929, t2 = 0.02767792568499834, t3 = 0.02079067210376509,
                     t4 = 0.01664469118982119, t5 = 0.01387612882307075, t6 = 0.01189670994589177, t7 = 0.01041126526197209,
                     t8 = 0.009255462182712733, t9 = 0.008330563433362871;
        const double lo = k < 1 ? t0 : k < 2 ? t1 : k < 3 ? t2 : k < 4 ? t3 : t4;
        const double hi = k < 6 ? t5 : k < 7 ? t6 : k < 8 ? t7 : k < 9 ? t8 : t9;
        return k < 5 ? lo : hi;
    }
    const double kp1 = (double)k + 1.0, kp1sq = kp1 * kp1;
    return (1.0 / 12 - (1.0 / 360 - 1.0 / 1260 / kp1sq) / kp1sq) / kp1;
}

BAYES_HD double recip(double x) {
#if defined(__CUDA_ARCH__)
    return __drcp_rn(x);  // correctly rounded, i.e. the same double as 1.0 / x on the host
#else
    return 1.0 / x;
#endif
}

// (1-q)^n by binary exponentiation (IEEE multiplications only: bit-identical to the oracle's orc_ipow)
BAYES_HD double ipow(double base, int n) {
    double result = 1.0;
    while (n > 0) {
        if (n & 1) result *= base;
        base *= base;
        n >>= 1;
    }
    return result;
}

// BTRD (Hoermann 1993), q <= 0.5, n*q >= 10; one block per attempt.  Same operation order as oracle/orc_rng.h.
// Out of line: about 900 instructions that the fast sampling mode almost never runs (its binomials have n <= 64) and that
// must not sit in the instruction stream of the hot loop (the executed footprint has to fit the 32 KB instruction cache).
BAYES_HD_NOINLINE int binomial_btrd(const Site &s, int n, double q) {
    const double npq = n * q * (1.0 - q);
    const double spq = sqrt(npq);
    const double bb = 1.15 + 2.53 * spq;
    const double inv_bb = recip(bb);
    const double a = -0.0873 + 0.0248 * bb + 0.01 * q;
    const double c = n * q + 0.5;
    const double v_r = 0.92 - 4.2 * inv_bb;
    const double inv_vr = recip(v_r);
    const double r = q / (1.0 - q);
    const double alpha = (2.83 + 5.1 * inv_bb) * spq;
    const int m = (int)floor((n + 1) * q);
    const double nr = (n + 1) * r;
    for (uint32_t j = 0;; j++) {
        const u32x4 b = s.block(j);
        double v = u53(b.x, b.y);
        const double u2 = u53(b.z, b.w);
        double u;
        if (v <= 0.86 * v_r) {
            u = v * inv_vr - 0.43;
            return (int)floor((2.0 * a / (0.5 - fabs(u)) + bb) * u + c);
        }
        if (v >= v_r) {
            u = u2 - 0.5;
        } else {
            u = v * inv_vr - 0.93;
            u = (u < 0.0 ? -0.5 : 0.5) - u;
            v = u2 * v_r;
        }
        const double us = 0.5 - fabs(u);
        const double kd = floor((2.0 * a / us + bb) * u + c);
        if (kd < 0.0 || kd > (double)n) continue;
        const int k = (int)kd;
        v = v * alpha / (a / (us * us) + bb);
        const int km = k > m ? k - m : m - k;
        if (km <= 15) {
            double num = 1.0, den = 1.0;
            if (m < k) {
                for (int i = m + 1; i <= k; i++) { num *= (nr - r * i); den *= (double)i; }
                if (v * den <= num) return k;
            } else {
                for (int i = k + 1; i <= m; i++) { num *= (nr - r * i); den *= (double)i; }
                if (v * num <= den) return k;
            }
            continue;
        }
        v = dlog(v);
        const double kmd = (double)km;
        const double rho = (kmd / npq) * (((kmd / 3.0 + 0.625) * kmd + 1.0 / 6.0) / npq + 0.5);
        const double t = -kmd * kmd / (2.0 * npq);
        if (v < t - rho) return k;
        if (v > t + rho) continue;
        const double nm = (double)(n - m + 1);
        const double h = (m + 0.5) * dlog((m + 1) / (r * nm)) + stirling_fc(m) + stirling_fc(n - m);
        const double nk = (double)(n - k + 1);
        if (v <= h + (n + 1) * dlog(nm / nk) + (k + 0.5) * dlog(nk * r / (k + 1)) - stirling_fc(k) - stirling_fc(n - k)) return k;
    }
}

// n * min(p, 1-p) below this: inversion (BINV), else BTRD.  Hoermann's BTRD needs >= 10 (measured best switch point on B200: 32); inversion costs one short
// loop iteration per unit of n*q but has a single, branch-free body, which is what 32 lanes in lockstep want.
#ifndef BAYES_BINV_NQ
#define BAYES_BINV_NQ 32.0
#endif

// 1/k for the inversion recurrence: the correctly rounded reciprocal, so a table entry is the same double as 1.0 / k
#if defined(__CUDACC__)
static __constant__ double kInvTable[64] = {
    0.0,      1.0 / 1,  1.0 / 2,  1.0 / 3,  1.0 / 4,  1.0 / 5,  1.0 / 6,  1.0 / 7,  1.0 / 8,  1.0 / 9,  1.0 / 10, 1.0 / 11, 1.0 / 12,
    1.0 / 13, 1.0 / 14, 1.0 / 15, 1.0 / 16, 1.0 / 17, 1.0 / 18, 1.0 / 19, 1.0 / 20, 1.0 / 21, 1.0 / 22, 1.0 / 23, 1.0 / 24, 1.0 / 25,
    1.0 / 26, 1.0 / 27, 1.0 / 28, 1.0 / 29, 1.0 / 30, 1.0 / 31, 1.0 / 32, 1.0 / 33, 1.0 / 34, 1.0 / 35, 1.0 / 36, 1.0 / 37, 1.0 / 38,
    1.0 / 39, 1.0 / 40, 1.0 / 41, 1.0 / 42, 1.0 / 43, 1.0 / 44, 1.0 / 45, 1.0 / 46, 1.0 / 47, 1.0 / 48, 1.0 / 49, 1.0 / 50, 1.0 / 51,
    1.0 / 52, 1.0 / 53, 1.0 / 54, 1.0 / 55, 1.0 / 56, 1.0 / 57, 1.0 / 58, 1.0 / 59, 1.0 / 60, 1.0 / 61, 1.0 / 62, 1.0 / 63};
#endif

BAYES_HD double inv_int(int k) {
#if defined(__CUDA_ARCH__)
    return __drcp_rn((double)k);
#else
    return 1.0 / (double)k;
#endif
}

// binomial_distribution(n, p) (assignmentGenerator.cpp:230-233, 346-349)
BAYES_HD_NOINLINE int binomial_variate(const Site &s, int n, double p) {
    if (n <= 0 || !(p > 0.0)) return 0;
    if (p >= 1.0) return n;
    const bool flip = p > 0.5;
    const double q = flip ? 1.0 - p : p;
    int k;
    if ((double)n * q < BAYES_BINV_NQ) {
        // BINV: P(0) = (1-q)^n, P(k) = P(k-1) * ((n-k+1) s / k).  The factor does not depend on P(k-1), so the
        // dependent chain per step is one multiplication and one subtraction.
        const u32x4 b = s.block(0);
        double u = u53(b.x, b.y);
        const double sr = q / (1.0 - q);
        double r = ipow(1.0 - q, n);
        double left = (double)n;  // n - k + 1 of the next step, kept as a double (exact) to save a conversion per step
        k = 0;
#if defined(__CUDA_ARCH__)
        // the first 63 steps take 1/k from the constant table: lanes run the recurrence in lockstep from k = 1, so the
        // read is a broadcast, and the loop body has no branch
        const int k_tab = n < 63 ? n : 63;
        while (u > r && k < k_tab) {
            u -= r;
            k++;
            r = r * ((left * sr) * kInvTable[k]);
            left -= 1.0;
        }
#endif
        while (u > r && k < n) {
            u -= r;
            k++;
            r = r * ((left * sr) * inv_int(k));
            left -= 1.0;
        }
    } else {
        k = binomial_btrd(s, n, q);
    }
    return flip ? n - k : k;
}

}  // namespace bayes
