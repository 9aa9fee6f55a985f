// ORACLE — TEST INFRASTRUCTURE ONLY (not part of the product path).
//
// CPU restatement of the third-party random-number stack the reference's
// `simulate` path depends on.  None of these crates is vendored under
// /root/reference (Cargo.toml:25-47 gives only semver ranges, Cargo.lock is
// git-ignored), so their published algorithms are restated here and pinned by
// the reference's own seed-42 golden vectors (see
// tests/test_oracle_golden.py):
//
//   rand 0.8 / rand_core 0.6 / rand_chacha 0.3
//       StdRng = ChaCha12, SeedableRng::seed_from_u64 (PCG32 key expansion),
//       BlockRng word stream, Rng::gen_range (UniformInt::sample_single_inclusive),
//       Rng::gen_bool (Bernoulli), Standard f64, Open01, WeightedIndex<f64>
//   rand_distr 0.4
//       Geometric, Gamma (Marsaglia-Tsang), StandardNormal (256-layer ziggurat),
//       Beta (Cheng BB / BC)
//
// Reference call sites: src/lib.rs:22,31; src/simulate/mod.rs:29,183,201,258,261,314;
// src/simulate/error.rs:293; src/simulate/fragments.rs:57,59,68,88,162;
// src/references.rs:41,87-89; src/model/error.rs:99-100,120,126;
// src/model/quality.rs:125-126; src/model/adapter.rs:35,45,68-70,84-86;
// src/model/glitch.rs:23-40,56-58; src/model/length.rs:29,43; src/model/identity.rs:36,50.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <vector>

namespace orc {

static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static inline uint32_t rotr32(uint32_t x, unsigned r) {
  r &= 31;
  return r ? ((x >> r) | (x << (32 - r))) : x;
}

// rand::rngs::StdRng (rand 0.8) == rand_chacha::ChaCha12Rng.
// The generator is exposed as a flat stream of 32-bit words.
class StdRng {
 public:
  StdRng() : pos_(16) { std::memset(key_, 0, sizeof key_); block_ = 0; }

  // rand_core::SeedableRng::seed_from_u64: PCG32 output fills the 32-byte key.
  static StdRng seed_from_u64(uint64_t state) {
    StdRng r;
    const uint64_t MUL = 6364136223846793005ULL, INC = 11634580027462260723ULL;
    for (int i = 0; i < 8; ++i) {
      state = state * MUL + INC;
      uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
      uint32_t rot = (uint32_t)(state >> 59);
      r.key_[i] = rotr32(xorshifted, rot);
    }
    r.block_ = 0;
    r.pos_ = 16;
    return r;
  }

  uint32_t next_u32() {
    if (pos_ >= 16) refill();
    ++words_;
    return buf_[pos_++];
  }
  // BlockRng::next_u64: two consecutive words, low word first, no alignment.
  uint64_t next_u64() {
    uint64_t lo = next_u32();
    uint64_t hi = next_u32();
    return (hi << 32) | lo;
  }
  // number of 32-bit words consumed so far (the device path addresses the
  // stream by this index)
  uint64_t words_consumed() const { return words_; }

  // ---- rand::Rng helpers -------------------------------------------------
  // Standard f64: 53 random bits * 2^-53
  double gen_f64() { return (double)(next_u64() >> 11) * (1.0 / 9007199254740992.0); }
  // Open01 f64
  double open01() {
    uint64_t bits = (next_u64() >> 12) | 0x3FF0000000000000ULL;
    double f;
    std::memcpy(&f, &bits, 8);
    return f - (1.0 - 2.220446049250313e-16 / 2.0);
  }
  // gen_range on a 64-bit integer type (usize/u64), inclusive bounds.
  uint64_t gen_range_u64(uint64_t low, uint64_t high_incl) {
    uint64_t range = high_incl - low + 1;
    if (range == 0) return next_u64();
    uint64_t zone = (range << __builtin_clzll(range)) - 1;
    for (;;) {
      uint64_t v = next_u64();
      unsigned __int128 m = (unsigned __int128)v * range;
      uint64_t hi = (uint64_t)(m >> 64), lo = (uint64_t)m;
      if (lo <= zone) return low + hi;
    }
  }
  // gen_range on a 32-bit integer type (i32/u32), inclusive bounds: u32 draws.
  uint32_t gen_range_u32(uint32_t low, uint32_t high_incl) {
    uint32_t range = high_incl - low + 1;
    if (range == 0) return next_u32();
    uint32_t zone = (range << __builtin_clz(range)) - 1;
    for (;;) {
      uint32_t v = next_u32();
      uint64_t m = (uint64_t)v * range;
      uint32_t hi = (uint32_t)(m >> 32), lo = (uint32_t)m;
      if (lo <= zone) return low + hi;
    }
  }
  // Rng::gen_bool(p) = Bernoulli::new(p).unwrap().sample(): p == 1.0 consumes
  // nothing; p outside [0,1] panics in the reference.
  bool gen_bool(double p) {
    if (!(p >= 0.0 && p < 1.0)) {
      if (p == 1.0) return true;
      throw std::runtime_error("gen_bool: p outside [0,1] (reference panics)");
    }
    uint64_t p_int = (uint64_t)(p * 18446744073709551616.0);
    return next_u64() < p_int;
  }

 private:
  void refill() {
    static const uint32_t C[4] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574};
    uint32_t s[16], x[16];
    for (int i = 0; i < 4; ++i) s[i] = C[i];
    for (int i = 0; i < 8; ++i) s[4 + i] = key_[i];
    s[12] = (uint32_t)block_;
    s[13] = (uint32_t)(block_ >> 32);
    s[14] = 0;
    s[15] = 0;
    std::memcpy(x, s, sizeof x);
#define ORC_QR(a, b, c, d)                    \
  x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16); \
  x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12); \
  x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);  \
  x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);
    for (int r = 0; r < 6; ++r) {  // 12 rounds = 6 double rounds
      ORC_QR(0, 4, 8, 12) ORC_QR(1, 5, 9, 13) ORC_QR(2, 6, 10, 14) ORC_QR(3, 7, 11, 15)
      ORC_QR(0, 5, 10, 15) ORC_QR(1, 6, 11, 12) ORC_QR(2, 7, 8, 13) ORC_QR(3, 4, 9, 14)
    }
#undef ORC_QR
    for (int i = 0; i < 16; ++i) buf_[i] = x[i] + s[i];
    ++block_;
    pos_ = 0;
  }
  uint32_t key_[8];
  uint64_t block_;
  uint32_t buf_[16];
  int pos_;
  uint64_t words_ = 0;
};

// rand::distributions::WeightedIndex<f64>
struct WeightedIndex {
  std::vector<double> cum;  // n-1 partial sums
  double total = 0.0;
  double scale = 0.0;  // UniformFloat::new(0, total).scale
  explicit WeightedIndex(const std::vector<double>& w) { build(w.data(), w.size()); }
  WeightedIndex(const double* w, size_t n) { build(w, n); }
  void build(const double* w, size_t n) {
    if (n == 0) throw std::runtime_error("WeightedIndex: no item");
    cum.reserve(n - 1);
    total = w[0];
    if (!(w[0] >= 0.0)) throw std::runtime_error("WeightedIndex: invalid weight");
    for (size_t i = 1; i < n; ++i) {
      if (!(w[i] >= 0.0)) throw std::runtime_error("WeightedIndex: invalid weight");
      cum.push_back(total);
      total += w[i];
    }
    if (total == 0.0) throw std::runtime_error("WeightedIndex: all weights zero");
    // Uniform::new(0.0, total)
    const double max_rand = 1.0 - 2.220446049250313e-16;
    scale = total;
    while (scale * max_rand + 0.0 >= total) scale = std::nextafter(scale, 0.0);
  }
  size_t sample(StdRng& rng) const {
    uint64_t bits = (rng.next_u64() >> 12) | 0x3FF0000000000000ULL;
    double v12;
    std::memcpy(&v12, &bits, 8);
    double chosen = (v12 - 1.0) * scale + 0.0;
    // partition_point(|w| w <= chosen)
    size_t lo = 0, hi = cum.size();
    while (lo < hi) {
      size_t mid = lo + (hi - lo) / 2;
      if (cum[mid] <= chosen) lo = mid + 1; else hi = mid;
    }
    return lo;
  }
};

// compiler-rt __powidf2: what f64::powi lowers to for a run-time exponent.
static inline double powi(double a, int b) {
  const bool recip = b < 0;
  double r = 1.0;
  for (;;) {
    if (b & 1) r *= a;
    b /= 2;
    if (b == 0) break;
    a *= a;
  }
  return recip ? 1.0 / r : r;
}

// rand_distr::Geometric
struct Geometric {
  double p = 1.0, pi = 1.0;
  uint64_t k = 0;
  Geometric() {}
  explicit Geometric(double p_) {
    if (!std::isfinite(p_) || p_ < 0.0 || p_ > 1.0) throw std::runtime_error("Geometric: invalid p");
    p = p_;
    if (p == 0.0 || p >= 2.0 / 3.0) {
      pi = p;
      k = 0;
    } else {
      k = 1;
      pi = powi(1.0 - p, 2);
      while (pi > 0.5) {
        ++k;
        pi = pi * pi;
      }
    }
  }
  uint64_t sample(StdRng& rng) const {
    if (p >= 2.0 / 3.0) {  // branch not pinned by any reference golden vector
      uint64_t failures = 0;
      for (;;) {
        double u = rng.gen_f64();
        if (u <= p) break;
        ++failures;
      }
      return failures;
    }
    if (p == 0.0) return UINT64_MAX;
    uint64_t d = 0;
    while (rng.gen_f64() < pi) ++d;
    uint64_t m;
    for (;;) {
      m = rng.next_u64() & ((1ULL << k) - 1);
      double p_reject = (m <= (uint64_t)INT32_MAX) ? powi(1.0 - p, (int)m) : std::pow(1.0 - p, (double)m);
      double u = rng.gen_f64();
      if (u < p_reject) break;
    }
    return (d << k) + m;
  }
};

// rand_distr ziggurat tables for the normal distribution.  The crate ships
// them as literals produced by its ziggurat_tables.py ('%.18f' of the
// recurrence below); they are not on disk here, so they are regenerated
//  and pinned by src/model/length.rs:64-113.
struct ZigNorm {
  double X[257], F[257];
  static constexpr double R = 3.654152885361008796;
  ZigNorm() {
    const double V = 0.00492867323399;
    auto f = [](double x) { return std::exp(-x * x / 2.0); };
    auto finv = [](double y) { return std::sqrt(-2.0 * std::log(y)); };
    X[0] = V / f(3.6541528853610088);
    X[1] = 3.6541528853610088;
    for (int i = 2; i < 256; ++i) {
      double last = X[i - 1];
      X[i] = finv(V / last + f(last));
    }
    X[256] = 0.0;
    for (int i = 0; i <= 256; ++i) F[i] = f(X[i]);
    // the crate's literals are the '%.18f' renderings of those values
    char buf[64];
    for (int i = 0; i <= 256; ++i) {
      std::snprintf(buf, sizeof buf, "%.18f", X[i]);
      X[i] = std::strtod(buf, nullptr);
      std::snprintf(buf, sizeof buf, "%.18f", F[i]);
      F[i] = std::strtod(buf, nullptr);
    }
  }
  static const ZigNorm& get() {
    static ZigNorm z;
    return z;
  }
};

static inline double standard_normal(StdRng& rng) {
  const ZigNorm& Z = ZigNorm::get();
  for (;;) {
    uint64_t bits = rng.next_u64();
    size_t i = bits & 0xff;
    uint64_t fb = (bits >> 12) | 0x4000000000000000ULL;  // exponent 1 -> [2,4)
    double u24;
    std::memcpy(&u24, &fb, 8);
    double u = u24 - 3.0;
    double x = u * Z.X[i];
    if (std::fabs(x) < Z.X[i + 1]) return x;
    if (i == 0) {
      double xx = 1.0, yy = 0.0;
      while (-2.0 * yy < xx * xx) {
        double x_ = rng.open01();
        double y_ = rng.open01();
        xx = std::log(x_) / ZigNorm::R;
        yy = std::log(y_);
      }
      return (u < 0.0) ? xx - ZigNorm::R : ZigNorm::R - xx;
    }
    if (Z.F[i + 1] + (Z.F[i] - Z.F[i + 1]) * rng.gen_f64() < std::exp(-x * x / 2.0)) return x;
  }
}

// rand_distr::Gamma.  shape == 1 (Exp ziggurat) is not pinned
// by any reference test and is refused loudly.
struct Gamma {
  double shape, scale, c, d, inv_shape;
  int mode;  // 0 large, 1 small
  Gamma() : shape(2), scale(1), c(0), d(0), inv_shape(0), mode(0) {}
  Gamma(double shape_, double scale_) : shape(shape_), scale(scale_) {
    if (!(shape > 0.0) || !(scale > 0.0)) throw std::runtime_error("Gamma: invalid parameter");
    if (shape == 1.0) throw std::runtime_error("Gamma: shape == 1 (Exp branch) is not restated");
    if (shape < 1.0) {
      mode = 1;
      inv_shape = 1.0 / shape;
      double s = shape + 1.0;
      d = s - 1.0 / 3.0;
      c = 1.0 / std::sqrt(9.0 * d);
    } else {
      mode = 0;
      inv_shape = 0.0;
      d = shape - 1.0 / 3.0;
      c = 1.0 / std::sqrt(9.0 * d);
    }
  }
  double sample_large(StdRng& rng) const {
    for (;;) {
      double x = standard_normal(rng);
      double v_cbrt = 1.0 + c * x;
      if (v_cbrt <= 0.0) continue;
      double v = v_cbrt * v_cbrt * v_cbrt;
      double u = rng.open01();
      double x_sqr = x * x;
      if (u < 1.0 - 0.0331 * x_sqr * x_sqr || std::log(u) < 0.5 * x_sqr + d * (1.0 - v + std::log(v)))
        return d * v * scale;
    }
  }
  double sample(StdRng& rng) const {
    if (mode == 1) {  // branch not pinned by any reference golden vector
      double u = rng.open01();
      return sample_large(rng) * std::pow(u, inv_shape);
    }
    return sample_large(rng);
  }
};

// rand_distr::Beta 0.4.x, Cheng 1978 BB / BC
struct Beta {
  double a, b;
  bool switched;
  bool bb;
  double alpha, beta, gamma;          // BB
  double delta, kappa1, kappa2;       // BC (alpha, beta shared)
  Beta() : a(1), b(1), switched(false), bb(false), alpha(0), beta(0), gamma(0), delta(0), kappa1(0), kappa2(0) {}
  Beta(double alpha0, double beta0) {
    if (!(alpha0 > 0.0) || !(beta0 > 0.0)) throw std::runtime_error("Beta: invalid parameter");
    double a0 = alpha0, b0 = beta0;
    if (a0 < b0) { a = a0; b = b0; switched = false; } else { a = b0; b = a0; switched = true; }
    gamma = delta = kappa1 = kappa2 = 0.0;
    if (a > 1.0) {
      bb = true;
      alpha = a + b;
      beta = std::sqrt((alpha - 2.0) / (2.0 * a * b - alpha));
      gamma = a + 1.0 / beta;
    } else {
      bb = false;
      double na = b, nb = a;
      a = na; b = nb; switched = !switched;
      alpha = a + b;
      beta = 1.0 / b;
      delta = 1.0 + a - b;
      kappa1 = delta * (1.0 / 18.0 / 4.0 + 3.0 / 18.0 / 4.0 * b) / (a * beta - 14.0 / 18.0);
      kappa2 = 0.25 + (0.5 + 0.25 / delta) * b;
    }
  }
  double sample(StdRng& rng) const {
    double w = 0.0;
    if (bb) {
      for (;;) {
        double u1 = rng.open01();
        double u2 = rng.open01();
        double v = beta * std::log(u1 / (1.0 - u1));
        w = a * std::exp(v);
        double z = u1 * u1 * u2;
        double r = gamma * v - std::log(4.0);
        double s = a + r - w;
        if (s + 1.0 + std::log(5.0) >= 5.0 * z) break;
        double t = std::log(z);
        if (s >= t) break;
        if (!(r + alpha * std::log(alpha / (b + w)) < t)) break;
      }
    } else {
      for (;;) {
        double z;
        double u1 = rng.open01();
        double u2 = rng.open01();
        if (u1 < 0.5) {
          double y = u1 * u2;
          z = u1 * y;
          if (0.25 * u2 + z - y >= kappa1) continue;
        } else {
          z = u1 * u1 * u2;
          if (z <= 0.25) {
            double v = beta * std::log(u1 / (1.0 - u1));
            w = a * std::exp(v);
            break;
          }
          if (z >= kappa2) continue;
        }
        double v = beta * std::log(u1 / (1.0 - u1));
        w = a * std::exp(v);
        if (!(alpha * (std::log(alpha / (b + w)) + v) - std::log(4.0) < std::log(z))) break;
      }
    }
    if (!switched) {
      if (w == std::numeric_limits<double>::infinity()) return 1.0;
      return w / (b + w);
    }
    return b / (b + w);
  }
};

}  // namespace orc
