/*
 * oracle/shim/orc_boost_shim.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Stand-in for the subset of Boost.Random / Boost.Math / lexical_cast that the
 * seven sampler translation units of the reference use.  Boost itself is absent
 * from this image and is not pinned by the reference (no build file), so the
 * UNMODIFIED reference sources under /root/reference/src are compiled against
 * these few classes instead (recipe: oracle/Makefile -> oracle/_ref/).
 *
 * Every distribution forwards to the published-algorithm variates in
 * oracle/orc_rng.h over a sequential MT19937 word stream, i.e. one 32-bit word
 * per uniform_01 call and whole 4-word blocks per gamma/binomial attempt.  The
 * engine counts the words it hands out so that the C restatement
 * (oracle/bayes_oracle.c, ORC_WS_SEQ mode) can be checked to consume exactly
 * the same stream.
 */
#ifndef ORC_BOOST_SHIM_HPP
#define ORC_BOOST_SHIM_HPP

#include <algorithm>
#include <cassert>
#include <cmath>
#include <sstream>
#include <stdint.h>
#include <string>

#include "../orc_rng.h"

namespace boost {
namespace random {

class mt19937 {
  public:
    typedef uint32_t result_type;
    mt19937() { orc_mt_seed(&state_, 5489u); orc_ws_init_seq(&ws_, &state_); }
    explicit mt19937(uint32_t s) { orc_mt_seed(&state_, s); orc_ws_init_seq(&ws_, &state_); }
    mt19937(const mt19937 &o) : state_(o.state_), ws_(o.ws_) { ws_.mt = &state_; }
    mt19937 &operator=(const mt19937 &o) { state_ = o.state_; ws_ = o.ws_; ws_.mt = &state_; return *this; }
    void seed(uint32_t s) { orc_mt_seed(&state_, s); orc_ws_init_seq(&ws_, &state_); }
    uint32_t operator()() { return orc_ws_next(&ws_); }
    orc_ws *word_source() { return &ws_; }
    uint64_t words_consumed() const { return ws_.words; }

  private:
    orc_mt state_;
    orc_ws ws_;
};

template <class Engine> class uniform_01;
template <> class uniform_01<mt19937 *> {
  public:
    explicit uniform_01(mt19937 *e) : eng_(e) {}
    double operator()() { return orc_u01((*eng_)()); }

  private:
    mt19937 *eng_;
};

template <class IntType = int> class uniform_int_distribution {
  public:
    uniform_int_distribution(IntType lo, IntType hi) : lo_(lo), hi_(hi) {}
    IntType operator()(mt19937 &eng) {
        return lo_ + (IntType)orc_uniform_int(eng.word_source(), (uint32_t)(hi_ - lo_) + 1u);
    }

  private:
    IntType lo_, hi_;
};

template <class IntType = int, class RealType = double> class binomial_distribution {
  public:
    typedef IntType result_type;
    binomial_distribution(IntType n, RealType p) : n_(n), p_(p) {}
    IntType operator()(mt19937 &eng) { return (IntType)orc_binomial(eng.word_source(), (int)n_, (double)p_); }

  private:
    IntType n_;
    RealType p_;
};

template <class RealType = double> class gamma_distribution {
  public:
    typedef RealType result_type;
    explicit gamma_distribution(RealType alpha, RealType beta = 1.0) : alpha_(alpha), beta_(beta) {}
    RealType operator()(mt19937 &eng) { return beta_ * orc_gamma(eng.word_source(), (double)alpha_); }

  private:
    RealType alpha_, beta_;
};

/* only reached from LogNormalHyperPrior::init (gammaGenerator.cpp:206-216), dead code in v1.2.0 */
template <class RealType = double> class lognormal_distribution {
  public:
    typedef RealType result_type;
    lognormal_distribution(RealType m, RealType s) : m_(m), s_(s) {}
    RealType operator()(mt19937 &eng) {
        uint32_t b[4];
        orc_ws_block(eng.word_source(), b);
        double z = std::sqrt(-2.0 * std::log(orc_u53(b[0], b[1]))) * std::cos(6.283185307179586 * orc_u01(b[2]));
        return std::exp(m_ + s_ * z);
    }

  private:
    RealType m_, s_;
};

template <class EnginePtr, class Dist> class variate_generator {
  public:
    variate_generator(EnginePtr e, Dist d) : eng_(e), dist_(d) {}
    typename Dist::result_type operator()() { return dist_(*eng_); }

  private:
    EnginePtr eng_;
    Dist dist_;
};

} // namespace random

namespace math {

inline double lgamma(double x) { int sign; return ::lgamma_r(x, &sign); } /* re-entrant: baseline runs loci on a thread pool */

template <class RealType = double> class beta_distribution {
  public:
    beta_distribution(RealType a = 1, RealType b = 1) : a_(a), b_(b) {}

  private:
    RealType a_, b_;
};

namespace constants {
template <class T> inline T pi() { return (T)3.14159265358979323846264338327950288; }
} // namespace constants

} // namespace math

template <class Target, class Source> inline Target lexical_cast(const Source &s) {
    std::stringstream ss;
    ss << s;
    return ss.str();
}

} // namespace boost

#endif
