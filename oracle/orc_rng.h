/*
 * oracle/orc_rng.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Random-word sources and uniform->variate transforms of the CPU oracle.
 *
 * Why this file exists: the reference draws every variate through Boost.Random
 * (uniform_01, uniform_int_distribution, binomial_distribution,
 * gamma_distribution on one boost::random::mt19937 per locus; call sites
 * assignmentGenerator.cpp:37-40,81-82,230-233,276,295,346-349,
 * expressionGenerator.cpp:37-40,89-92,100-113, simplexSizeGenerator.cpp:218-220).
 * Boost is a third-party dependency that is NOT under /root/reference, is not
 * vendored and not version-pinned (there is no build file at all), so the
 * uniform->variate transforms are not defined by the reference.  They are
 * defined HERE, from the published algorithms:
 *   - MT19937: Matsumoto & Nishimura 1998 (init_genrand form, what
 *     mt19937::seed(uint32) is specified to do);
 *   - Philox4x32-10: Salmon, Moraes, Dror, Shaw, SC'11;
 *   - uniform_int: Lemire 2019 multiply-shift with rejection (unbiased);
 *   - gamma: Marsaglia & Tsang 2000 (alpha>=1), alpha<1 by the U^(1/alpha)
 *     boost, alpha==1 by inversion; normal by Box-Muller (cosine branch);
 *   - binomial: inversion (BINV, Kachitvichyanukul & Schmeiser 1988) when
 *     n*min(p,1-p) < ORC_BINV_NQ (32), else BTRD (Hoermann 1993).
 * The same file backs (a) the Boost stand-in headers under oracle/shim/ that
 * the UNMODIFIED reference translation units are compiled against
 * (oracle/_ref) and (b) the C restatement in oracle/bayes_oracle.c, so that
 * the two can be compared draw for draw.  The CUDA product has its OWN
 * independent implementation of the same spec (bayesembler_b200/csrc/); only
 * tests/, smoke() and bench.py's cpu_baseline leg may use anything in oracle/.
 *
 * Word sources ("where do the 32-bit words come from"):
 *   ORC_WS_SEQ   one sequential MT19937 stream, words consumed in call order
 *                (exactly the reference's single-RNG-per-locus behaviour);
 *   ORC_WS_SITE  counter-based: every draw site (phase, iteration, index, sub)
 *                owns the Philox4x32-10 stream
 *                    word k = philox(key, ctr = {(sub<<12) | (k>>2), index,
 *                                                iteration, phase})[k & 3]
 *                so that sites can be evaluated in any order / in parallel.
 * All variates consume whole 4-word blocks, except uniform_01 and
 * uniform_int, which consume single words.
 */
#ifndef ORC_RNG_H
#define ORC_RNG_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ MT19937 */
typedef struct {
    uint32_t mt[624];
    int idx;
} orc_mt;

static inline void orc_mt_seed(orc_mt *s, uint32_t seed) {
    s->mt[0] = seed;
    for (int i = 1; i < 624; i++)
        s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
    s->idx = 624;
}

static inline uint32_t orc_mt_next(orc_mt *s) {
    if (s->idx >= 624) {
        for (int i = 0; i < 624; i++) {
            uint32_t y = (s->mt[i] & 0x80000000u) | (s->mt[(i + 1) % 624] & 0x7fffffffu);
            uint32_t v = s->mt[(i + 397) % 624] ^ (y >> 1);
            if (y & 1u) v ^= 0x9908b0dfu;
            s->mt[i] = v;
        }
        s->idx = 0;
    }
    uint32_t y = s->mt[s->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* ------------------------------------------------------------ Philox4x32-10 */
static inline void orc_philox4x32_10(const uint32_t key[2], const uint32_t ctr[4], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* draw-site phases (shared numbering with the CUDA product) */
enum {
    ORC_PH_COVER = 1,       /* calcMinimumReadCover tie break, index = greedy round    */
    ORC_PH_INIT_ASSIGN = 2, /* initAssignment binomial chain, index = row, sub = col   */
    ORC_PH_SIMPLEX = 3,     /* simplex size uniform, iteration = iteration that USES b */
    ORC_PH_GAMMA = 4,       /* expression gamma draw, index = transcript               */
    ORC_PH_NULLPICK = 5,    /* null-transcript pick, index = pick ordinal              */
    ORC_PH_ASSIGN = 6,      /* generateAssignment, index = row, sub = col (chain) or 0 */
    /* production ("fast") sampling mode, ORC_WS_FAST: same distributions, different draw sites (bayes_oracle.c) */
    ORC_PH_POOL = 7,        /* generateAssignment, words of all rows with < 20 fragments: block b of the iteration is
                               block (b & 4095) of site index (b >> 12)                                          */
    ORC_PH_INIT_POOL = 8,   /* the same for initAssignment                                                    */
    ORC_PH_TREE = 9,        /* generateAssignment, one binomial per node of a row's splitting tree: index = row,
                               phase word = ORC_PH_TREE | node << 8, node = first cell | (cells - 1) << 12       */
    ORC_PH_INIT_TREE = 10,  /* the same for initAssignment                                                    */
    ORC_PH_DIRECT = 11,     /* generateAssignment, rows with >= 20 fragments drawn fragment by fragment: index = row,
                               fragment f takes word (f & 3) of block (f >> 2) of the row's site                  */
    ORC_PH_INIT_DIRECT = 12 /* the same for initAssignment                                                    */
};

/* --------------------------------------------------------------- word source */
enum { ORC_WS_SEQ = 0, ORC_WS_SITE = 1, ORC_WS_FAST = 2 /* sites like ORC_WS_SITE; only bayes_oracle.c tells them apart */ };

typedef struct {
    int mode;
    orc_mt *mt;         /* ORC_WS_SEQ */
    uint64_t words;     /* words handed out so far (both modes; consumption check) */
    uint32_t key[2];    /* ORC_WS_SITE */
    uint32_t ctr[4];    /* block counter of the current site (ctr[0] low 12 bits = block) */
    uint32_t k;         /* next word index inside the current site */
    uint32_t buf[4];
} orc_ws;

static inline void orc_ws_init_seq(orc_ws *w, orc_mt *mt) {
    memset(w, 0, sizeof(*w));
    w->mode = ORC_WS_SEQ;
    w->mt = mt;
}

static inline void orc_ws_init_site(orc_ws *w, uint64_t key) {
    memset(w, 0, sizeof(*w));
    w->mode = ORC_WS_SITE; /* ORC_WS_FAST draws from sites in exactly the same way */
    w->key[0] = (uint32_t)key;
    w->key[1] = (uint32_t)(key >> 32);
}

/* position the source at the start of a draw site (no-op for the sequential stream) */
static inline void orc_ws_site(orc_ws *w, uint32_t phase, uint32_t iteration, uint32_t index, uint32_t sub) {
    if (w->mode != ORC_WS_SITE) return;
    w->ctr[0] = sub << 12;
    w->ctr[1] = index;
    w->ctr[2] = iteration;
    w->ctr[3] = phase;
    w->k = 0;
}

static inline uint32_t orc_ws_next(orc_ws *w) {
    w->words++;
    if (w->mode == ORC_WS_SEQ) return orc_mt_next(w->mt);
    if ((w->k & 3u) == 0) {
        uint32_t c[4] = { w->ctr[0] + (w->k >> 2), w->ctr[1], w->ctr[2], w->ctr[3] };
        orc_philox4x32_10(w->key, c, w->buf);
    }
    return w->buf[w->k++ & 3u];
}

static inline void orc_ws_block(orc_ws *w, uint32_t out[4]) {
    out[0] = orc_ws_next(w); out[1] = orc_ws_next(w);
    out[2] = orc_ws_next(w); out[3] = orc_ws_next(w);
}

/* ----------------------------------------------------------------- uniforms */
/* uniform_01<mt19937*>: one 32-bit word -> [0,1)   (assignmentGenerator.cpp:276,295) */
static inline double orc_u01(uint32_t w) { return (double)w * 2.3283064365386963e-10; /* 2^-32 */ }
/* 32-bit open-interval uniform in (0,1) */
static inline double orc_u32_open(uint32_t w) { return ((double)w + 0.5) * 2.3283064365386963e-10; }
/* 53-bit open-interval uniform in (0,1) from two words */
static inline double orc_u53(uint32_t hi, uint32_t lo) {
    uint64_t x = (((uint64_t)hi << 32) | lo) >> 11;
    return ((double)x + 0.5) * 1.1102230246251565e-16; /* 2^-53 */
}

/* uniform_int_distribution(0, n-1)   (assignmentGenerator.cpp:81-82, expressionGenerator.cpp:106-108) */
static inline int orc_uniform_int(orc_ws *w, uint32_t n) {
    if (n <= 1) return 0; /* no randomness needed, no word consumed */
    uint32_t x = orc_ws_next(w);
    uint64_t m = (uint64_t)x * n;
    uint32_t l = (uint32_t)m;
    if (l < n) {
        uint32_t t = (0u - n) % n;
        while (l < t) {
            x = orc_ws_next(w);
            m = (uint64_t)x * n;
            l = (uint32_t)m;
        }
    }
    return (int)(m >> 32);
}

/* -------------------------------------------------------------------- gamma */
/* gamma_distribution(alpha, 1)   (expressionGenerator.cpp:89-92, 100-113) */
static inline double orc_gamma(orc_ws *w, double alpha) {
    uint32_t b[4];
    if (alpha == 1.0) {
        orc_ws_block(w, b);
        return -log(orc_u53(b[0], b[1]));
    }
    double boost = 1.0, a = alpha;
    if (alpha < 1.0) {
        orc_ws_block(w, b);
        boost = exp(log(orc_u53(b[0], b[1])) / alpha);
        a = alpha + 1.0;
    }
    const double d = a - 1.0 / 3.0;
    const double c = 1.0 / sqrt(9.0 * d);
    for (;;) {
        orc_ws_block(w, b);
        double u1 = orc_u53(b[0], b[1]);
        double x = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * orc_u01(b[2]));
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        double u = orc_u32_open(b[3]);
        double x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2) return d * v * boost;
        if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) return d * v * boost;
    }
}

/* ----------------------------------------------------------------- binomial */
static inline double orc_stirling_fc(int k) {
    static const double tab[10] = {
        0.08106146679532726, 0.04134069595540929, 0.02767792568499834, 0.02079067210376509,
        0.01664469118982119, 0.01387612882307075, 0.01189670994589177, 0.01041126526197209,
        0.009255462182712733, 0.008330563433362871 };
    if (k < 10) return tab[k];
    double kp1 = (double)k + 1.0, kp1sq = kp1 * kp1;
    return (1.0 / 12 - (1.0 / 360 - 1.0 / 1260 / kp1sq) / kp1sq) / kp1;
}

/* (1-q)^n by binary exponentiation, least significant bit first: IEEE multiplications only, so the CUDA product
 * reproduces it bit for bit (no libm call whose last bit could differ between host and device) */
static inline double orc_ipow(double base, int n) {
    double result = 1.0;
    while (n > 0) {
        if (n & 1) result *= base;
        base *= base;
        n >>= 1;
    }
    return result;
}

/* BTRD, Hoermann 1993; requires q <= 0.5 and n*q >= 10; one 4-word block per attempt.  Divisions that share a
 * denominator are written as one reciprocal and multiplications, and the recursive evaluation of f(k)/f(m) (step 3.1)
 * keeps numerator and denominator separately: the algorithm is unchanged, it is only cheaper on the GPU. */
static inline int orc_btrd(orc_ws *w, int n, double q) {
    const double npq = n * q * (1.0 - q);
    const double spq = sqrt(npq);
    const double bb = 1.15 + 2.53 * spq;
    const double inv_bb = 1.0 / bb;
    const double a = -0.0873 + 0.0248 * bb + 0.01 * q;
    const double c = n * q + 0.5;
    const double v_r = 0.92 - 4.2 * inv_bb;
    const double inv_vr = 1.0 / v_r;
    const double r = q / (1.0 - q);
    const double alpha = (2.83 + 5.1 * inv_bb) * spq;
    const int m = (int)floor((n + 1) * q);
    const double nr = (n + 1) * r;
    uint32_t b[4];
    for (;;) {
        orc_ws_block(w, b);
        double v = orc_u53(b[0], b[1]);
        double u2 = orc_u53(b[2], b[3]);
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
        double us = 0.5 - fabs(u);
        double kd = floor((2.0 * a / us + bb) * u + c);
        if (kd < 0.0 || kd > (double)n) continue;
        int k = (int)kd;
        v = v * alpha / (a / (us * us) + bb);
        int km = k > m ? k - m : m - k;
        if (km <= 15) {
            /* f(k)/f(m) = prod (nr/i - r) = prod (nr - r i) / prod i */
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
        v = log(v);
        double kmd = (double)km;
        double rho = (kmd / npq) * (((kmd / 3.0 + 0.625) * kmd + 1.0 / 6.0) / npq + 0.5);
        double t = -kmd * kmd / (2.0 * npq);
        if (v < t - rho) return k;
        if (v > t + rho) continue;
        double nm = (double)(n - m + 1);
        double h = (m + 0.5) * log((m + 1) / (r * nm)) + orc_stirling_fc(m) + orc_stirling_fc(n - m);
        double nk = (double)(n - k + 1);
        if (v <= h + (n + 1) * log(nm / nk) + (k + 0.5) * log(nk * r / (k + 1)) - orc_stirling_fc(k) - orc_stirling_fc(n - k))
            return k;
    }
}

/* n * min(p, 1-p) below this: inversion (BINV), else BTRD (which needs >= 10) */
#ifndef ORC_BINV_NQ
#define ORC_BINV_NQ 32.0
#endif

/* binomial_distribution(n, p)   (assignmentGenerator.cpp:230-233, 346-349) */
static inline int orc_binomial(orc_ws *w, int n, double p) {
    if (n <= 0 || !(p > 0.0)) return 0; /* also NaN; nothing consumed */
    if (p >= 1.0) return n;             /* nothing consumed */
    const int flip = p > 0.5;
    const double q = flip ? 1.0 - p : p;
    int k;
    if ((double)n * q < ORC_BINV_NQ) {
        /* BINV: sequential search from 0 with P(0) = (1-q)^n, P(k) = P(k-1) ((n-k+1) s (1/k)), s = q/(1-q) */
        uint32_t b[4];
        orc_ws_block(w, b);
        double u = orc_u53(b[0], b[1]);
        const double s = q / (1.0 - q);
        double r = orc_ipow(1.0 - q, n);
        k = 0;
        while (u > r && k < n) {
            u -= r;
            k++;
            r = r * (((double)(n - k + 1) * s) * (1.0 / (double)k));
        }
    } else {
        k = orc_btrd(w, n, q);
    }
    return flip ? n - k : k;
}

#ifdef __cplusplus
}
#endif
#endif /* ORC_RNG_H */
