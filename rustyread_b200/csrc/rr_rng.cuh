// rustyread_b200 — random-number stack shared by the host feeder and the device kernels.
//
// What the reference gets from the crates rand 0.8 / rand_chacha 0.3 / rand_distr 0.4
// (Cargo.toml:28-29): StdRng = ChaCha12 seeded through SeedableRng::seed_from_u64, uniform
// integer ranges, Bernoulli, Open01, WeightedIndex, Geometric and Beta.  The generator is
// treated as a counter-addressable stream of 32-bit words (block b, word w = stream word
// 16 b + w), which is what lets a warp speculate ahead and a qscore thread jump to its draw.
//
// Reference call sites this serves on the device: src/lib.rs:18-45, src/simulate/mod.rs:244-271,
// src/simulate/error.rs:250-302, src/model/error.rs:93-135, src/model/glitch.rs:51-64,
// src/model/adapter.rs:64-93, src/model/quality.rs:113-133.
#pragma once
#include <cstdint>
#include <cmath>

#if defined(__CUDACC__)
#define RR_HD __host__ __device__ __forceinline__
#define RR_UNROLL _Pragma("unroll")
#else
#define RR_HD inline
#define RR_UNROLL
#endif

namespace rr {

RR_HD uint32_t rotl(uint32_t x, int r) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(x, x, r);
#else
  return (x << r) | (x >> (32 - r));
#endif
}

// SeedableRng::seed_from_u64: eight PCG32 outputs form the ChaCha key.
RR_HD void chacha_key_from_u64(uint64_t state, uint32_t key[8]) {
  const uint64_t MUL = 6364136223846793005ULL, INC = 11634580027462260723ULL;
  RR_UNROLL
  for (int i = 0; i < 8; ++i) {
    state = state * MUL + INC;
    uint32_t xs = (uint32_t)(((state >> 18) ^ state) >> 27);
    uint32_t rot = (uint32_t)(state >> 59);
    key[i] = (xs >> rot) | (xs << ((32 - rot) & 31));
  }
}

#define RR_QR(a, b, c, d)                 \
  a += b; d = rotl(d ^ a, 16);            \
  c += d; b = rotl(b ^ c, 12);            \
  a += b; d = rotl(d ^ a, 8);             \
  c += d; b = rotl(b ^ c, 7);

// One ChaCha12 block (64-bit block counter in words 12-13, stream id 0).
RR_HD void chacha12_block(const uint32_t key[8], uint64_t counter, uint32_t out[16]) {
  const uint32_t c0 = 0x61707865u, c1 = 0x3320646eu, c2 = 0x79622d32u, c3 = 0x6b206574u;
  uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3;
  uint32_t x4 = key[0], x5 = key[1], x6 = key[2], x7 = key[3];
  uint32_t x8 = key[4], x9 = key[5], x10 = key[6], x11 = key[7];
  uint32_t x12 = (uint32_t)counter, x13 = (uint32_t)(counter >> 32), x14 = 0, x15 = 0;
  RR_UNROLL
  for (int r = 0; r < 6; ++r) {
    RR_QR(x0, x4, x8, x12) RR_QR(x1, x5, x9, x13) RR_QR(x2, x6, x10, x14) RR_QR(x3, x7, x11, x15)
    RR_QR(x0, x5, x10, x15) RR_QR(x1, x6, x11, x12) RR_QR(x2, x7, x8, x13) RR_QR(x3, x4, x9, x14)
  }
  out[0] = x0 + c0; out[1] = x1 + c1; out[2] = x2 + c2; out[3] = x3 + c3;
  out[4] = x4 + key[0]; out[5] = x5 + key[1]; out[6] = x6 + key[2]; out[7] = x7 + key[3];
  out[8] = x8 + key[4]; out[9] = x9 + key[5]; out[10] = x10 + key[6]; out[11] = x11 + key[7];
  out[12] = x12 + (uint32_t)counter; out[13] = x13 + (uint32_t)(counter >> 32); out[14] = x14; out[15] = x15;
}

// ---- scalar conversions (rand 0.8) ---------------------------------------------------------
RR_HD double bits_to_f64(uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)b);
#else
  double d;
  __builtin_memcpy(&d, &b, 8);
  return d;
#endif
}
RR_HD double u64_to_f64_standard(uint64_t v) { return (double)(v >> 11) * 1.1102230246251565e-16; }          // Standard
RR_HD double u64_to_open01(uint64_t v) {                                                                         // Open01
  return bits_to_f64((v >> 12) | 0x3FF0000000000000ULL) - (1.0 - 1.1102230246251565e-16);
}
RR_HD double u64_to_unit12(uint64_t v) { return bits_to_f64((v >> 12) | 0x3FF0000000000000ULL) - 1.0; }      // UniformFloat value0_1
RR_HD uint64_t bernoulli_threshold(double p) { return (uint64_t)(p * 18446744073709551616.0); }                 // Bernoulli::new

RR_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}
RR_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)x);
#else
  return __builtin_clzll(x);
#endif
}
RR_HD int clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __clz((int)x);
#else
  return __builtin_clz(x);
#endif
}

// compiler-rt __powidf2 (what f64::powi lowers to)
RR_HD double powi(double a, int b) {
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

// ---- distribution parameters (constructed on the host, sampled on either side) ------------------------
struct GeometricParams {  // rand_distr::Geometric::new
  double p, pi;
  uint64_t k;
};
inline bool geometric_new(double p, GeometricParams& g) {
  if (!std::isfinite(p) || p < 0.0 || p > 1.0) return false;
  g.p = p;
  if (p == 0.0 || p >= 2.0 / 3.0) {
    g.pi = p;
    g.k = 0;
  } else {
    g.k = 1;
    g.pi = powi(1.0 - p, 2);
    while (g.pi > 0.5) {
      ++g.k;
      g.pi = g.pi * g.pi;
    }
  }
  return true;
}

struct BetaParams {  // rand_distr::Beta::new (Cheng BB / BC)
  double a, b, alpha, beta, gamma, delta, kappa1, kappa2;
  int switched, bb;
};
inline bool beta_new(double alpha0, double beta0, BetaParams& B) {
  if (!(alpha0 > 0.0) || !(beta0 > 0.0)) return false;
  if (alpha0 < beta0) { B.a = alpha0; B.b = beta0; B.switched = 0; } else { B.a = beta0; B.b = alpha0; B.switched = 1; }
  B.gamma = B.delta = B.kappa1 = B.kappa2 = 0.0;
  if (B.a > 1.0) {
    B.bb = 1;
    B.alpha = B.a + B.b;
    B.beta = std::sqrt((B.alpha - 2.0) / (2.0 * B.a * B.b - B.alpha));
    B.gamma = B.a + 1.0 / B.beta;
  } else {
    B.bb = 0;
    double na = B.b, nb = B.a;
    B.a = na; B.b = nb; B.switched = !B.switched;
    B.alpha = B.a + B.b;
    B.beta = 1.0 / B.b;
    B.delta = 1.0 + B.a - B.b;
    B.kappa1 = B.delta * (1.0 / 18.0 / 4.0 + 3.0 / 18.0 / 4.0 * B.b) / (B.a * B.beta - 14.0 / 18.0);
    B.kappa2 = 0.25 + (0.5 + 0.25 / B.delta) * B.b;
  }
  return true;
}

// Generic samplers over any generator exposing next_u64().
template <class R>
RR_HD uint64_t geometric_sample(const GeometricParams& g, R& rng) {
  if (g.p >= 2.0 / 3.0) {
    uint64_t failures = 0;
    for (;;) {
      double u = u64_to_f64_standard(rng.next_u64());
      if (u <= g.p) break;
      ++failures;
    }
    return failures;
  }
  if (g.p == 0.0) return ~0ULL;
  uint64_t d = 0;
  while (u64_to_f64_standard(rng.next_u64()) < g.pi) ++d;
  uint64_t m;
  const double q = 1.0 - g.p;
  for (;;) {
    m = rng.next_u64() & ((1ULL << g.k) - 1);
    double p_reject = (m <= 0x7fffffffULL) ? powi(q, (int)m) : pow(q, (double)m);
    double u = u64_to_f64_standard(rng.next_u64());
    if (u < p_reject) break;
  }
  return (d << g.k) + m;
}

template <class R>
RR_HD double beta_sample(const BetaParams& B, R& rng) {
  double w = 0.0;
  if (B.bb) {
    for (;;) {
      double u1 = u64_to_open01(rng.next_u64());
      double u2 = u64_to_open01(rng.next_u64());
      double v = B.beta * log(u1 / (1.0 - u1));
      w = B.a * exp(v);
      double z = u1 * u1 * u2;
      double r = B.gamma * v - log(4.0);
      double s = B.a + r - w;
      if (s + 1.0 + log(5.0) >= 5.0 * z) break;
      double t = log(z);
      if (s >= t) break;
      if (!(r + B.alpha * log(B.alpha / (B.b + w)) < t)) break;
    }
  } else {
    for (;;) {
      double z;
      double u1 = u64_to_open01(rng.next_u64());
      double u2 = u64_to_open01(rng.next_u64());
      if (u1 < 0.5) {
        double y = u1 * u2;
        z = u1 * y;
        if (0.25 * u2 + z - y >= B.kappa1) continue;
      } else {
        z = u1 * u1 * u2;
        if (z <= 0.25) {
          double v = B.beta * log(u1 / (1.0 - u1));
          w = B.a * exp(v);
          break;
        }
        if (z >= B.kappa2) continue;
      }
      double v = B.beta * log(u1 / (1.0 - u1));
      w = B.a * exp(v);
      if (!(B.alpha * (log(B.alpha / (B.b + w)) + v) - log(4.0) < log(z))) break;
    }
  }
  if (!B.switched) {
    if (w > 1.7976931348623157e308) return 1.0;  // w == +inf
    return w / (B.b + w);
  }
  return B.b / (B.b + w);
}

// WeightedIndex<f64>: cumulative sums (sequential, left to right) + UniformFloat scale.
// cum_out gets n entries: cum[0..n-2] are the partial sums searched, cum[n-1] = total.
inline bool weighted_index_build(const double* w, size_t n, double* cum_out, double* scale_out) {
  if (n == 0) return false;
  double total = w[0];
  if (!(w[0] >= 0.0)) return false;
  for (size_t i = 1; i < n; ++i) {
    if (!(w[i] >= 0.0)) return false;
    cum_out[i - 1] = total;
    total += w[i];
  }
  cum_out[n - 1] = total;
  if (total == 0.0) return false;
  const double max_rand = 1.0 - 2.220446049250313e-16;
  double scale = total;
  while (scale * max_rand + 0.0 >= total) scale = std::nextafter(scale, 0.0);
  *scale_out = scale;
  return true;
}

}  // namespace rr
