// rustyread_b200 — host side of `rustyread simulate` (include/rustyread_b200_host.h).
//
// C++ stand-in for the Rust host code around the hot path: main_rng schedule, model construction and
// loading, reference ingest, the serial Fragments feeder and the serial FASTQ writer.  Per read this is
// O(1) work (about ten draws of main_rng, one MD5, one line of text); all per-base work is on the GPU.
//
// Reference: src/simulate/mod.rs:27-217, src/simulate/fragments.rs:13-165, src/simulate/description.rs:59-109,
// src/references.rs:26-171, src/model/{length,identity,error,quality}.rs, src/cli/simulate.rs:12-80.
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include <cerrno>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/uio.h>
#include <thread>
#include <unistd.h>

#include "../../include/rustyread_b200_host.h"
#include "rr_rng.cuh"

#define RR_API extern "C" __attribute__((visibility("default")))

namespace {

using rr::BetaParams;

// ---- main_rng: StdRng consumed word by word on the host ---------------------------------------------
struct HostRng {
  uint32_t key[8];
  uint32_t buf[16];
  uint64_t block = 0;
  int pos = 16;
  uint64_t words = 0;
  explicit HostRng(uint64_t seed) { rr::chacha_key_from_u64(seed, key); }
  uint32_t next_u32() {
    if (pos == 16) { rr::chacha12_block(key, block++, buf); pos = 0; }
    ++words;
    return buf[pos++];
  }
  uint64_t next_u64() { uint64_t lo = next_u32(); return lo | ((uint64_t)next_u32() << 32); }
  bool gen_bool(double p, bool* ok) {  // Rng::gen_bool; *ok = false where the reference panics
    if (p == 1.0) return true;
    if (!(p >= 0.0 && p < 1.0)) { *ok = false; return false; }
    return next_u64() < rr::bernoulli_threshold(p);
  }
  uint64_t range_u64(uint64_t low, uint64_t high_incl) {
    uint64_t range = high_incl - low + 1;
    if (range == 0) return next_u64();
    uint64_t zone = (range << rr::clz64(range)) - 1;
    for (;;) {
      uint64_t v = next_u64();
      if (v * range <= zone) return low + rr::mulhi64(v, range);
    }
  }
  uint32_t range_u32(uint32_t low, uint32_t high_incl) {
    uint32_t range = high_incl - low + 1;
    if (range == 0) return next_u32();
    uint32_t zone = (range << rr::clz32(range)) - 1;
    for (;;) {
      uint64_t m = (uint64_t)next_u32() * range;
      if ((uint32_t)m <= zone) return low + (uint32_t)(m >> 32);
    }
  }
  uint8_t base() {  // random_base, src/lib.rs:18-23 (NUCS = A C T G)
    static const char N[4] = {'A', 'C', 'T', 'G'};
    return (uint8_t)N[range_u64(0, 3)];
  }
};

// ---- StandardNormal: rand_distr's 256-layer ziggurat ------------------------------------------------------
// The crate's table literals are '%.18f' renderings of this recurrence (ziggurat_tables.py); they are not on
// disk in the reference checkout, so they are regenerated at start-up.
struct Ziggurat {
  double x[257], f[257];
  static constexpr double R = 3.654152885361008796;
  Ziggurat() {
    const double v = 0.00492867323399;
    auto pdf = [](double t) { return std::exp(-t * t / 2.0); };
    x[0] = v / pdf(R);
    x[1] = R;
    for (int i = 2; i < 256; ++i) x[i] = std::sqrt(-2.0 * std::log(v / x[i - 1] + pdf(x[i - 1])));
    x[256] = 0.0;
    char tmp[64];
    for (int i = 0; i <= 256; ++i) {
      double fi = pdf(x[i]);
      std::snprintf(tmp, sizeof tmp, "%.18f", x[i]);
      x[i] = std::strtod(tmp, nullptr);
      std::snprintf(tmp, sizeof tmp, "%.18f", fi);
      f[i] = std::strtod(tmp, nullptr);
    }
  }
  double sample(HostRng& g) const {
    for (;;) {
      uint64_t bits = g.next_u64();
      unsigned i = bits & 0xff;
      double u = rr::bits_to_f64((bits >> 12) | 0x4000000000000000ULL) - 3.0;
      double xv = u * x[i];
      if (std::fabs(xv) < x[i + 1]) return xv;
      if (i == 0) {
        double a = 1.0, b = 0.0;
        while (-2.0 * b < a * a) {
          double a_ = rr::u64_to_open01(g.next_u64());
          double b_ = rr::u64_to_open01(g.next_u64());
          a = std::log(a_) / R;
          b = std::log(b_);
        }
        return u < 0.0 ? a - R : R - a;
      }
      if (f[i + 1] + (f[i] - f[i + 1]) * rr::u64_to_f64_standard(g.next_u64()) < std::exp(-xv * xv / 2.0)) return xv;
    }
  }
};
const Ziggurat& zig() { static Ziggurat z; return z; }

// ---- model::Length (src/model/length.rs) : Gamma(k, theta) by Marsaglia-Tsang ------------------------------
struct LengthModel {
  double mean = 0, scale = 0, c = 0, d = 0, inv_shape = 0;
  int mode = -1;  // -1 constant, 0 shape > 1, 1 shape < 1
  const char* init(double mean_, double stdev) {
    if (mean_ <= 0.0 || stdev < 0.0) return "length: mean must be > 0 and stdev >= 0";
    mean = mean_;
    if (stdev == 0.0) { mode = -1; return nullptr; }
    double shape = std::pow(mean, 2.0) / std::pow(stdev, 2.0);
    scale = std::pow(stdev, 2.0) / mean;
    if (!(shape > 0.0) || !(scale > 0.0)) return "length: invalid Gamma parameters";
    if (shape == 1.0) return "length: stdev == mean selects rand_distr's Exp sampler, which is not implemented";
    mode = shape < 1.0 ? 1 : 0;
    double s = mode ? shape + 1.0 : shape;
    inv_shape = mode ? 1.0 / shape : 0.0;
    d = s - 1.0 / 3.0;
    c = 1.0 / std::sqrt(9.0 * d);
    return nullptr;
  }
  double gamma_large(HostRng& g) const {
    for (;;) {
      double x = zig().sample(g);
      double vc = 1.0 + c * x;
      if (vc <= 0.0) continue;
      double v = vc * vc * vc;
      double u = rr::u64_to_open01(g.next_u64());
      double x2 = x * x;
      if (u < 1.0 - 0.0331 * x2 * x2 || std::log(u) < 0.5 * x2 + d * (1.0 - v + std::log(v))) return d * v * scale;
    }
  }
  uint64_t get(HostRng& g) const {  // length.rs:38-47
    double v;
    if (mode < 0) v = mean;
    else if (mode == 1) { double u = rr::u64_to_open01(g.next_u64()); v = gamma_large(g) * std::pow(u, inv_shape); }
    else v = gamma_large(g);
    return (uint64_t)std::round(v);
  }
};

// ---- model::Identity (src/model/identity.rs) -----------------------------------------------------------------------
struct IdentityModel {
  double mean = 0, max = 0;
  bool has = false;
  BetaParams beta{};
  const char* init(double mean_, double max_, double stdev) {
    if (mean_ <= 0.0 || stdev <= 0.0) return "identity: mean and stdev must be > 0";
    mean = mean_ / 100.0;
    max = max_ / 100.0;
    stdev /= 100.0;
    has = std::fabs(mean - max) > 2.220446049250313e-16;
    if (has) {
      double a = (((1.0 - (mean / max)) / std::pow(stdev / max, 2.0)) - (max / mean)) * std::pow(mean / max, 2.0);
      double b = a * ((max / mean) - 1.0);
      if (!rr::beta_new(a, b, beta)) return "identity: invalid Beta parameters";
    }
    return nullptr;
  }
  double get(HostRng& g) const { return has ? max * rr::beta_sample(beta, g) : mean; }
};

// ---- WeightedIndex over contigs --------------------------------------------------------------------------------------
struct Weighted {
  std::vector<double> cum;
  double scale = 0;
  bool build(const std::vector<double>& w) {
    cum.resize(w.size());
    return !w.empty() && rr::weighted_index_build(w.data(), w.size(), cum.data(), &scale);
  }
  size_t sample(HostRng& g) const {
    double chosen = rr::u64_to_unit12(g.next_u64()) * scale + 0.0;
    size_t lo = 0, hi = cum.size() - 1;
    while (lo < hi) {
      size_t mid = (lo + hi) / 2;
      if (cum[mid] <= chosen) lo = mid + 1; else hi = mid;
    }
    return lo;
  }
};

struct Contig {
  std::string id;
  std::string seq;       // empty in a session made from an index (rr_sim_create_indexed): the bases live on the device only
  uint64_t len = 0;      // number of bases
  bool circular = false;
  double weight = 0;     // depth x length, then adjusted (References::adjust_depth)
  double weight0 = 0;    // depth x length as read from the header
};

// ---- MD5 (RFC 1321) for Uuid::new_v3 ------------------------------------------------------------------------------------
struct Md5 {
  uint32_t h[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
  static uint32_t rl(uint32_t x, int s) { return (x << s) | (x >> (32 - s)); }
  void block(const uint8_t* p) {
    static const uint32_t K[64] = {
        0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af, 0xffff5bb1,
        0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453,
        0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a, 0xfffa3942,
        0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05,
        0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d,
        0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
    static const uint8_t S[16] = {7, 12, 17, 22, 5, 9, 14, 20, 4, 11, 16, 23, 6, 10, 15, 21};
    uint32_t m[16];
    for (int i = 0; i < 16; ++i) m[i] = (uint32_t)p[4 * i] | (uint32_t)p[4 * i + 1] << 8 | (uint32_t)p[4 * i + 2] << 16 | (uint32_t)p[4 * i + 3] << 24;
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3];
    for (int i = 0; i < 64; ++i) {
      uint32_t f;
      int g;
      switch (i >> 4) {
        case 0: f = (b & c) | (~b & d); g = i; break;
        case 1: f = (d & b) | (~d & c); g = (5 * i + 1) & 15; break;
        case 2: f = b ^ c ^ d; g = (3 * i + 5) & 15; break;
        default: f = c ^ (b | ~d); g = (7 * i) & 15; break;
      }
      uint32_t t = d;
      d = c;
      c = b;
      b = b + rl(a + f + K[i] + m[g], S[(i >> 4) * 4 + (i & 3)]);
      a = t;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d;
  }
};
// Uuid::new_v3(&NAMESPACE_X500, name16).to_hyphenated() appended to `out` (36 chars).  32 bytes hash in one block.
void uuid_v3_x500(const uint8_t name[16], std::string& out) {
  uint8_t blk[64] = {0x6b, 0xa7, 0xb8, 0x14, 0x9d, 0xad, 0x11, 0xd1, 0x80, 0xb4, 0x00, 0xc0, 0x4f, 0xd4, 0x30, 0xc8};
  std::memcpy(blk + 16, name, 16);
  blk[32] = 0x80;
  blk[56] = 0x00;  // 256 bits = 0x0100 little-endian
  blk[57] = 0x01;
  Md5 m;
  m.block(blk);
  uint8_t d[16];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) d[4 * i + j] = (uint8_t)(m.h[i] >> (8 * j));
  d[6] = (d[6] & 0x0f) | 0x30;
  d[8] = (d[8] & 0x3f) | 0x80;
  static const char* hex = "0123456789abcdef";
  for (int i = 0; i < 16; ++i) {
    if (i == 4 || i == 6 || i == 8 || i == 10) out.push_back('-');
    out.push_back(hex[d[i] >> 4]);
    out.push_back(hex[d[i] & 15]);
  }
}

void append_u64(std::string& s, uint64_t v) {
  char b[24];
  auto r = std::to_chars(b, b + sizeof b, v);
  s.append(b, r.ptr);
}
// Rust `{}` of an f64: shortest digits that round-trip, positional notation
void append_f64(std::string& s, double v) {
  if (std::isnan(v)) { s += "NaN"; return; }
  if (std::isinf(v)) { s += v < 0 ? "-inf" : "inf"; return; }
  char b[400];
  auto r = std::to_chars(b, b + sizeof b, v, std::chars_format::fixed);
  s.append(b, r.ptr);
}

std::vector<std::string_view> split_view(std::string_view s, char sep, bool keep_empty) {
  std::vector<std::string_view> out;
  size_t p = 0;
  while (p <= s.size()) {
    size_t q = s.find(sep, p);
    if (q == std::string_view::npos) q = s.size();
    if (keep_empty || q > p) out.push_back(s.substr(p, q - p));
    p = q + 1;
  }
  return out;
}
std::string_view chomp(std::string_view l) {
  if (!l.empty() && l.back() == '\r') l.remove_suffix(1);
  return l;
}

}  // namespace

struct rr_sim {
  std::string err;
  HostRng rng{0};
  LengthModel length;
  IdentityModel identity;
  std::vector<Contig> contigs;
  Weighted contig_dist;
  rr_sim_params p{};
  std::string start_adapter, end_adapter;
  // error model as loaded (error.rs:26-82)
  uint32_t k = 7;
  bool error_random = true;
  std::vector<uint8_t> em_keys, em_alt_bytes;
  std::vector<uint32_t> em_alt_begin{0}, em_alt_off{0};
  std::vector<double> em_weights;
  // qscore model as loaded (quality.rs:26-110)
  uint32_t q_max_k = 1;
  std::vector<uint8_t> q_cigars, q_scores;
  std::vector<uint32_t> q_cigar_off{0}, q_score_off{0};
  std::vector<double> q_weights;
  // last batch
  std::vector<rr_desc> items;
  std::string fastq;
  // threads rr_sim_write_fastq may use on a regular file (rr_sim_set_writer_threads; 1 = one writev stream)
  int writer_threads = 1;
};

static std::string g_sim_error;
static int sfail(rr_sim* s, int code, const std::string& m) {
  if (s) s->err = m; else g_sim_error = m;
  return code;
}

// read_reference (src/references.rs:135-171) on FASTA text: id = first word, the rest is the description
static const char* parse_fasta(std::string_view text, std::vector<Contig>& out) {
  Contig cur;
  bool have = false;
  std::string desc;
  auto flush = [&]() -> const char* {
    if (!have) return nullptr;
    double w = 1.0;
    for (size_t p = desc.find("depth="); p != std::string::npos; p = desc.find("depth=", p + 1)) {  // regex depth=([\d.]+)
      size_t q = p + 6, e = q;
      while (e < desc.size() && ((desc[e] >= '0' && desc[e] <= '9') || desc[e] == '.')) ++e;
      if (e > q) {
        char* endp = nullptr;
        std::string num = desc.substr(q, e - q);
        w = std::strtod(num.c_str(), &endp);
        if (endp != num.c_str() + num.size()) return "reference: cannot parse depth= value";  // f64::from_str error
        break;
      }
    }
    cur.circular = desc.find("circular=true") != std::string::npos;
    cur.len = cur.seq.size();
    cur.weight = cur.weight0 = w * (double)cur.seq.size();
    out.push_back(std::move(cur));
    return nullptr;
  };
  size_t i = 0;
  while (i < text.size()) {
    size_t e = text.find('\n', i);
    if (e == std::string_view::npos) e = text.size();
    std::string_view line = text.substr(i, e - i);
    i = e + 1;
    while (!line.empty() && (line.back() == '\r' || line.back() == ' ' || line.back() == '\t')) line.remove_suffix(1);
    if (!line.empty() && line[0] == '>') {
      if (const char* m = flush()) return m;
      cur = Contig();
      have = true;
      line.remove_prefix(1);
      size_t sp = 0;
      while (sp < line.size() && line[sp] != ' ' && line[sp] != '\t') ++sp;
      cur.id.assign(line.substr(0, sp));
      desc.assign(sp < line.size() ? line.substr(sp + 1) : std::string_view());
      // the record's sequence lines end at the next "\n>": reserve once instead of growing by doubling (3.1 Gbp references)
      size_t nxt = text.find("\n>", i > 0 ? i - 1 : 0);
      cur.seq.reserve((nxt == std::string_view::npos ? text.size() : nxt) - std::min(i, text.size()) + 1);
    } else if (have) {
      cur.seq.append(line);
    }
  }
  if (const char* m = flush()) return m;
  return out.empty() ? "reference: no sequence found" : nullptr;
}

// random_error (src/model/error.rs:116-135) against main_rng, used by the error-model loader
static std::string host_random_error(std::string_view kmer, HostRng& g) {
  size_t pos = (size_t)g.range_u64(0, kmer.size() - 1);
  std::string out(kmer.substr(0, pos));
  switch (g.range_u32(1, 3)) {
    case 1: {
      uint8_t b;
      do b = g.base(); while (b == (uint8_t)kmer[pos]);
      out.push_back((char)b);
      break;
    }
    case 2: out.push_back((char)g.base()); out.push_back(kmer[pos]); break;
    default: break;
  }
  out.append(kmer.substr(pos + 1));
  return out;
}

// Error::from_stream (src/model/error.rs:26-82)
static const char* load_error_model(rr_sim* s, std::string_view text) {
  std::unordered_map<std::string, size_t> seen;  // key -> slot (HashMap::insert replaces)
  struct Slot { std::string key; std::vector<std::string> alts; std::vector<double> w; };
  std::vector<Slot> slots;
  size_t klen = 0;
  for (std::string_view line : split_view(text, '\n', false)) {
    line = chomp(line);
    if (line.empty()) continue;
    Slot sl;
    std::vector<std::string> names;
    std::vector<double> probs;
    for (std::string_view f : split_view(line, ';', true)) {
      if (f.empty()) continue;
      size_t c = f.find(',');
      if (c == std::string_view::npos) return "error model: field without ',' (Model::ErrorParsing)";
      std::string_view num = f.substr(c + 1);
      size_t c2 = num.find(',');
      if (c2 != std::string_view::npos) num = num.substr(0, c2);
      std::string ns(num);
      char* endp = nullptr;
      double p = std::strtod(ns.c_str(), &endp);
      if (ns.empty() || endp != ns.c_str() + ns.size()) return "error model: cannot parse probability";
      names.emplace_back(f.substr(0, c));
      probs.push_back(p);
    }
    if (names.empty()) continue;
    double sum = 0.0;
    for (double p : probs) sum += p;
    if (sum < 1.0) {
      if (names[0].empty()) return "error model: empty k-mer";
      names.push_back(host_random_error(names[0], s->rng));
      probs.push_back(1.0 - sum);
    }
    sl.key = names[0];
    sl.alts.assign(names.begin() + 1, names.end());
    sl.w.assign(probs.begin() + 1, probs.end());
    klen = sl.key.size();
    auto it = seen.find(sl.key);
    if (it == seen.end()) { seen[sl.key] = slots.size(); slots.push_back(std::move(sl)); }
    else slots[it->second] = std::move(sl);
  }
  s->k = (uint32_t)klen;
  s->error_random = false;
  for (auto& sl : slots) {
    if (sl.key.size() != klen) continue;  // can never equal a k-byte window of the read
    s->em_keys.insert(s->em_keys.end(), sl.key.begin(), sl.key.end());
    for (size_t j = 0; j < sl.alts.size(); ++j) {
      s->em_alt_bytes.insert(s->em_alt_bytes.end(), sl.alts[j].begin(), sl.alts[j].end());
      s->em_alt_off.push_back((uint32_t)s->em_alt_bytes.size());
      s->em_weights.push_back(sl.w[j]);
    }
    s->em_alt_begin.push_back((uint32_t)s->em_weights.size());
  }
  return klen ? nullptr : "error model: no record";
}

static void add_uniform_entry(rr_sim* s, const char* cigar, int lo, int hi) {
  s->q_cigars.insert(s->q_cigars.end(), cigar, cigar + std::strlen(cigar));
  s->q_cigar_off.push_back((uint32_t)s->q_cigars.size());
  for (int v = lo; v <= hi; ++v) { s->q_scores.push_back((uint8_t)v); s->q_weights.push_back(1.0); }
  s->q_score_off.push_back((uint32_t)s->q_scores.size());
}

// Quality::from_stream (src/model/quality.rs:26-77)
static const char* load_qscore_model(rr_sim* s, std::string_view text) {
  bool first = true;
  size_t max_k = 0;
  for (std::string_view line : split_view(text, '\n', false)) {
    line = chomp(line);
    if (line.empty()) continue;
    if (first) { first = false; continue; }  // the `overall` record is skipped (quality.rs:37)
    auto f = split_view(line, ';', true);
    if (f.size() < 3) return "qscore model: record with fewer than 3 fields (Model::QualityParsing)";
    s->q_cigars.insert(s->q_cigars.end(), f[0].begin(), f[0].end());
    s->q_cigar_off.push_back((uint32_t)s->q_cigars.size());
    max_k = std::max(max_k, f[0].size());
    for (std::string_view sw : split_view(f[2], ',', true)) {
      if (sw.empty()) continue;
      size_t c = sw.find(':');
      if (c == std::string_view::npos) return "qscore model: score without ':'";
      std::string a(sw.substr(0, c)), b(sw.substr(c + 1));
      size_t c2 = b.find(':');
      if (c2 != std::string::npos) b.resize(c2);
      char* e1 = nullptr;
      char* e2 = nullptr;
      unsigned long sc = std::strtoul(a.c_str(), &e1, 10);
      double w = std::strtod(b.c_str(), &e2);
      if (a.empty() || b.empty() || e1 != a.c_str() + a.size() || e2 != b.c_str() + b.size() || sc > 255) return "qscore model: cannot parse score:weight";
      s->q_scores.push_back((uint8_t)sc);
      s->q_weights.push_back(w);
    }
    s->q_score_off.push_back((uint32_t)s->q_scores.size());
  }
  s->q_max_k = (uint32_t)max_k;
  return nullptr;
}

RR_API void rr_sim_default_params(rr_sim_params* p) {
  if (!p) return;
  std::memset(p, 0, sizeof *p);
  p->seed = 42;
  p->length_mean = 15000; p->length_stdev = 13000;
  p->identity_mean = 85; p->identity_max = 95; p->identity_stdev = 5;
  p->start_adapter_seq = "AATGTACTTCGTTCAGTTACGTATTGCT";
  p->end_adapter_seq = "GCAATACGTAACTGAACGAAGT";
  p->start_adapter_rate = 90; p->start_adapter_amount = 60;
  p->end_adapter_rate = 50; p->end_adapter_amount = 20;
  p->junk_reads = 1; p->random_reads = 1; p->chimera = 1;
  p->glitch_rate = 10000; p->glitch_size = 25; p->glitch_skip = 25;
  p->small_plasmid_bias = 0;
  p->error_model = nullptr;
  p->qscore_model = "random";
  p->qscore_model_len = 6;
}

RR_API const char* rr_sim_last_error(const rr_sim* s) { return s ? s->err.c_str() : g_sim_error.c_str(); }

RR_API void rr_sim_destroy(rr_sim* s) { delete s; }

RR_API int64_t rr_quantity_to_bases(const char* text, uint64_t genome) {
  if (!text) return -1;
  std::string t(text);
  if (t.empty()) return -1;
  uint64_t mul = 1;
  bool cov = false;
  switch (t.back()) {
    case 'x': cov = true; t.pop_back(); break;
    case 'G': mul = 1000000000ULL; t.pop_back(); break;
    case 'M': mul = 1000000ULL; t.pop_back(); break;
    case 'K': mul = 1000ULL; t.pop_back(); break;
    default: break;
  }
  if (t.empty()) return -1;
  size_t b = t[0] == '+' ? 1 : 0;  // u64::from_str accepts a leading '+'
  if (b == t.size()) return -1;
  uint64_t v = 0;
  auto r = std::from_chars(t.data() + b, t.data() + t.size(), v);
  if (r.ec != std::errc() || r.ptr != t.data() + t.size()) return -1;
  return (int64_t)(cov ? v * genome : v * mul);
}

// Everything of simulate() up to the batch loop (mod.rs:28-127), in the reference's main_rng order; `fill` supplies the contigs
// (from FASTA text, or from an index when the bases are already on the device).
template <class Fill>
static int create_session(const rr_sim_params* p, rr_sim** out, Fill fill) {
  *out = nullptr;
  rr_sim* s = new rr_sim();
  auto bail = [&](const std::string& m) { int rc = sfail(nullptr, RR_ERR_ARG, m); delete s; return rc; };
  s->p = *p;
  s->start_adapter = p->start_adapter_seq ? p->start_adapter_seq : "";
  s->end_adapter = p->end_adapter_seq ? p->end_adapter_seq : "";
  s->rng = HostRng(p->seed);                                                             // mod.rs:28-37
  if (const char* m = s->length.init(p->length_mean, p->length_stdev)) return bail(m);  // mod.rs:40
  if (const char* m = fill(s->contigs)) return bail(m);
  {  // References::adjust_depth (references.rs:95-132): 100 000 lengths from main_rng, always
    std::vector<uint64_t> lens(100000);
    uint64_t total = 0;
    for (auto& l : lens) { l = s->length.get(s->rng); total += l; }
    for (auto& c : s->contigs) {
      uint64_t rl = c.len;
      if (!p->small_plasmid_bias && c.circular) {
        uint64_t pass = 0;
        for (uint64_t l : lens) if (l <= rl) pass += l;
        if (pass == 0) return bail("reference: a circular contig is shorter than every sampled fragment (Cli::SmallPlasmidBias)");
        c.weight *= (double)total / (double)pass;
      }
      if (!c.circular) {
        uint64_t pass = 0;
        for (uint64_t l : lens) pass += std::min(l, rl);
        c.weight *= (double)total / (double)pass;
      }
    }
    std::vector<double> w;
    for (auto& c : s->contigs) w.push_back(c.weight);
    if (!s->contig_dist.build(w)) return bail("reference: invalid contig weights (WeightedIndex::new)");
  }
  if (const char* m = s->identity.init(p->identity_mean, p->identity_max, p->identity_stdev)) return bail(m);  // mod.rs:58-63
  // Adapter::new / Glitch::new are validated by rr_set_adapter / rr_set_glitch at attach time (no main_rng draws)
  if (p->error_model && !(p->error_model_len == 6 && std::memcmp(p->error_model, "random", 6) == 0)) {         // mod.rs:88-104
    if (const char* m = load_error_model(s, std::string_view(p->error_model, p->error_model_len))) return bail(m);
  } else {
    s->k = 7;
    s->error_random = true;
  }
  std::string_view qm(p->qscore_model ? p->qscore_model : "random", p->qscore_model ? p->qscore_model_len : 6);
  if (qm == "ideal") {                                                                                           // mod.rs:109-111
    s->q_max_k = 9;
    add_uniform_entry(s, "X", 1, 3); add_uniform_entry(s, "I", 1, 3); add_uniform_entry(s, "=", 4, 7);
    add_uniform_entry(s, "===", 8, 20); add_uniform_entry(s, "=====", 21, 30); add_uniform_entry(s, "=======", 31, 40);
    add_uniform_entry(s, "=========", 41, 50);
  } else if (qm == "random") {                                                                                    // mod.rs:112-114
    s->q_max_k = 1;
    add_uniform_entry(s, "=", 1, 20); add_uniform_entry(s, "X", 1, 20); add_uniform_entry(s, "I", 1, 20);
  } else if (const char* m = load_qscore_model(s, qm)) return bail(m);
  *out = s;
  return RR_OK;
}

RR_API int rr_sim_create(const rr_sim_params* p, const char* fasta, uint64_t fasta_len, rr_sim** out) {
  if (!p || !fasta || !out) return sfail(nullptr, RR_ERR_ARG, "rr_sim_create: NULL argument");
  return create_session(p, out, [&](std::vector<Contig>& c) { return parse_fasta(std::string_view(fasta, fasta_len), c); });
}

RR_API int rr_sim_create_indexed(const rr_sim_params* p, uint64_t n, const char* const* ids, const uint64_t* lengths, const uint8_t* circular,
                                 const double* weights, rr_sim** out) {
  if (!p || !out || !ids || !lengths || !circular || !weights) return sfail(nullptr, RR_ERR_ARG, "rr_sim_create_indexed: NULL argument");
  return create_session(p, out, [&](std::vector<Contig>& c) -> const char* {
    if (n == 0) return "reference: no sequence found";
    for (uint64_t i = 0; i < n; ++i) {
      Contig x;
      x.id = ids[i] ? ids[i] : "";
      x.len = lengths[i];
      x.circular = circular[i] != 0;
      x.weight = x.weight0 = weights[i];
      c.push_back(std::move(x));
    }
    return nullptr;
  });
}

RR_API int rr_sim_contig_index(const rr_sim* s, uint64_t i, const char** id, uint64_t* length, int* circular, double* weight) {
  if (!s || i >= s->contigs.size()) return RR_ERR_ARG;
  const Contig& c = s->contigs[i];
  if (id) *id = c.id.c_str();
  if (length) *length = c.len;
  if (circular) *circular = c.circular;
  if (weight) *weight = c.weight0;
  return RR_OK;
}

RR_API uint64_t rr_sim_genome_length(const rr_sim* s) {
  uint64_t t = 0;
  if (s) for (auto& c : s->contigs) t += c.len;
  return t;
}
RR_API uint32_t rr_sim_k(const rr_sim* s) { return s ? s->k : 0; }
RR_API uint64_t rr_sim_n_refs(const rr_sim* s) { return s ? s->contigs.size() : 0; }
RR_API uint64_t rr_sim_rng_words(const rr_sim* s) { return s ? s->rng.words : 0; }
RR_API const uint8_t* rr_sim_ref(const rr_sim* s, uint64_t i, uint64_t* len, int* circular, double* weight, const char** id) {
  if (!s || i >= s->contigs.size()) return nullptr;
  const Contig& c = s->contigs[i];
  if (len) *len = c.len;
  if (circular) *circular = c.circular;
  if (weight) *weight = c.weight;
  if (id) *id = c.id.c_str();
  return (const uint8_t*)c.seq.data();
}

static int attach(rr_sim* s, rr_ctx* ctx, const void* device_bases) {
  if (!s || !ctx) return RR_ERR_ARG;
  std::vector<const uint8_t*> ptr;
  std::vector<uint64_t> len;
  std::vector<uint8_t> circ;
  bool have_bases = true;
  for (auto& c : s->contigs) {
    ptr.push_back((const uint8_t*)c.seq.data());
    len.push_back(c.len);
    circ.push_back(c.circular);
    have_bases = have_bases && c.seq.size() == c.len;
  }
  int rc;
  if (device_bases) rc = rr_set_references_device(ctx, len.size(), len.data(), circ.data(), device_bases);
  else if (!have_bases) return sfail(s, RR_ERR_ARG, "rr_sim_attach: this session was made from an index and holds no bases (use rr_sim_attach_device)");
  else rc = rr_set_references(ctx, ptr.size(), ptr.data(), len.data(), circ.data());
  if (!rc) {
    if (s->error_random) rc = rr_set_error_model(ctx, s->k, 0, nullptr, nullptr, nullptr, nullptr, nullptr);
    else rc = rr_set_error_model(ctx, s->k, s->em_alt_begin.size() - 1, s->em_keys.data(), s->em_alt_begin.data(), s->em_alt_off.data(),
                                 s->em_alt_bytes.data(), s->em_weights.data());
  }
  if (!rc) rc = rr_set_qscore_model(ctx, s->q_max_k, s->q_cigar_off.size() - 1, s->q_cigars.data(), s->q_cigar_off.data(), s->q_score_off.data(),
                                    s->q_scores.data(), s->q_weights.data());
  if (!rc) rc = rr_set_adapter(ctx, (const uint8_t*)s->start_adapter.data(), s->start_adapter.size(), (const uint8_t*)s->end_adapter.data(),
                               s->end_adapter.size(), s->p.start_adapter_rate, s->p.start_adapter_amount, s->p.end_adapter_rate, s->p.end_adapter_amount);
  if (!rc) rc = rr_set_glitch(ctx, s->p.glitch_rate, s->p.glitch_size, s->p.glitch_skip);
  if (rc) s->err = std::string("rr_sim_attach: ") + rr_last_error(ctx);
  return rc;
}

RR_API int rr_sim_attach(const rr_sim* s, rr_ctx* ctx) { return attach(const_cast<rr_sim*>(s), ctx, nullptr); }

RR_API int rr_sim_attach_device(const rr_sim* s, rr_ctx* ctx, const void* device_bases) {
  if (!device_bases) return RR_ERR_ARG;
  return attach(const_cast<rr_sim*>(s), ctx, device_bases);
}

// Fragments::generate_fragment (src/simulate/fragments.rs:72-114); returns the counted length
static uint64_t feed_fragment(rr_sim* s, bool* ok, uint32_t& type, uint32_t& ref_idx, uint32_t& strand, uint64_t& start, uint64_t& end) {
  HostRng& g = s->rng;
  if (g.gen_bool(s->p.junk_reads / 100.0, ok)) type = RR_JUNK;
  else if (g.gen_bool(s->p.random_reads / 100.0, ok)) type = RR_RANDOM;
  else type = RR_REAL;
  uint64_t length = s->length.get(g);
  ref_idx = 0;
  if (type != RR_REAL) { strand = RR_NOSTRAND; start = 0; end = length; return length; }
  // The reference re-draws the contig until the fragment fits (fragments.rs:80-85) and never leaves that loop when the
  // length reaches every selectable contig and all of them are circular (many small plasmids with the default lengths):
  // fail loudly instead of hanging.
  {
    bool possible = false;
    for (const Contig& c : s->contigs)
      if (c.weight > 0.0 && !(length >= c.len && c.circular)) { possible = true; break; }
    if (!possible) { *ok = false; s->err = "fragment of " + std::to_string(length) + " bases fits no contig: every selectable contig is circular and shorter (the reference loops forever here; use a shorter --length or add a linear contig)"; return 0; }
  }
  size_t idx;
  for (;;) {  // choose_reference (references.rs:83-92): strand from an i32 range (one u32 draw), then the contig
    strand = g.range_u32(0, 1) == 0 ? RR_PLUS : RR_MINUS;
    idx = s->contig_dist.sample(g);
    const Contig& c = s->contigs[idx];
    if (!(length >= c.len && c.circular)) break;  // fragment_is_possible (fragments.rs:117-123)
  }
  const Contig& c = s->contigs[idx];
  const uint64_t rl = c.len;
  ref_idx = (uint32_t)idx;
  uint64_t b = g.range_u64(0, rl - 1);
  start = b;
  if (b + length < rl) { end = b + length; return length; }
  if (c.circular) { end = length - (rl - b); return length; }
  end = rl - 1;  // fragments.rs:97-102: end = len - 1 but the counted length is len - begin
  return rl - b;
}

RR_API int rr_sim_feed(rr_sim* s, uint64_t base_loop, const rr_desc** descs, uint64_t* n) {
  if (!s || !descs || !n) return RR_ERR_ARG;
  s->items.clear();
  s->err.clear();
  uint64_t target = base_loop;
  bool ok = true;
  while (target != 0) {  // Fragments::next (fragments.rs:132-164)
    rr_desc d;
    std::memset(&d, 0, sizeof d);
    uint64_t l1 = feed_fragment(s, &ok, d.type, d.ref_idx, d.strand, d.start, d.end), l2 = 0;
    if (s->rng.gen_bool(s->p.chimera / 100.0, &ok)) {
      d.has_chimera = 1;
      l2 = feed_fragment(s, &ok, d.ctype, d.cref_idx, d.cstrand, d.cstart, d.cend);
    } else d.cstrand = RR_NOSTRAND;
    uint64_t tt = l1 + l2;
    target = tt > target ? 0 : target - tt;
    d.identity = s->identity.get(s->rng);
    d.seed = s->rng.next_u64();
    if (!ok) return sfail(s, RR_ERR_ARG, !s->err.empty() && s->err.rfind("fragment of ", 0) == 0 ? "rr_sim_feed: " + s->err
                                         : std::string("rr_sim_feed: junk/random/chimera rate outside 0..100 (gen_bool panics in the reference)"));
    s->items.push_back(d);
  }
  *descs = s->items.data();
  *n = s->items.size();
  return RR_OK;
}

static void append_origin(const rr_sim* s, std::string& o, uint32_t type, uint32_t ref, uint32_t strand, uint64_t start, uint64_t end) {
  if (type == RR_JUNK) { o += "junk_seq"; return; }
  if (type == RR_RANDOM) { o += "random_seq"; return; }
  o += s->contigs[ref].id;
  o += strand == RR_PLUS ? ",+strand," : ",-strand,";
  append_u64(o, start);
  o += '-';
  append_u64(o, end);
}
// Description Display (description.rs:93-109); `length` / `identity` are the post-generation values
static void append_description(const rr_sim* s, std::string& o, const rr_desc& d, uint64_t length, double identity) {
  append_origin(s, o, d.type, d.ref_idx, d.strand, d.start, d.end);
  if (d.has_chimera) {
    o += " chimera ";
    append_origin(s, o, d.ctype, d.cref_idx, d.cstrand, d.cstart, d.cend);
  }
  o += " length=";
  append_u64(o, d.end - d.start);  // usize arithmetic wraps in release builds (Cargo.toml:58)
  o += " error-free_length=";
  append_u64(o, length);
  o += " read_identity=";
  append_f64(o, identity);
  o += '%';
}

RR_API int rr_sim_item_description(rr_sim* s, uint64_t i, char* out, uint64_t cap) {
  if (!s || !out || i >= s->items.size()) return RR_ERR_ARG;
  const rr_desc& d = s->items[i];
  // Description::new(.., tt_length, identity): not kept per item; callers use this for origin text only
  std::string t;
  append_description(s, t, d, 0, d.identity);
  if (t.size() + 1 > cap) return RR_ERR_ARG;
  std::memcpy(out, t.c_str(), t.size() + 1);
  return RR_OK;
}

// Writes iov[0, n) to fd: at the file position (at < 0, writev) or at offset `at` (pwritev).  Short writes are continued.
static const char* write_all(int fd, iovec* iov, size_t n, off_t at, uint64_t& bytes) {
  size_t i = 0;
  while (i < n) {
    const int cnt = (int)std::min<size_t>(n - i, 1000);  // IOV_MAX is 1024
    ssize_t w = at < 0 ? ::writev(fd, iov + i, cnt) : ::pwritev(fd, iov + i, cnt, at);
    if (w < 0) {
      if (errno == EINTR) continue;
      return std::strerror(errno);
    }
    bytes += (uint64_t)w;
    if (at >= 0) at += w;
    size_t left = (size_t)w;
    while (left > 0 && i < n) {
      if (left >= iov[i].iov_len) { left -= iov[i].iov_len; ++i; }
      else { iov[i].iov_base = (char*)iov[i].iov_base + left; iov[i].iov_len -= left; left = 0; }
    }
    while (i < n && iov[i].iov_len == 0) ++i;
  }
  return nullptr;
}

RR_API int rr_sim_set_writer_threads(rr_sim* s, int threads) {
  if (!s || threads < 1) return RR_ERR_ARG;
  s->writer_threads = threads;
  return RR_OK;
}

RR_API int rr_sim_write_fastq(rr_sim* s, const rr_batch* b, uint64_t first, int fd, uint64_t* n_records, uint64_t* n_bases, uint64_t* n_bytes) {
  if (!s || !b) return RR_ERR_ARG;
  if (first + b->n_reads > s->items.size()) return sfail(s, RR_ERR_ARG, "rr_sim_write_fastq: batch does not match the last feed");
  const uint32_t k = b->k;
  // Pass 1 (serial: read names come from main_rng in write order, mod.rs:199-203): the two header lines of every emitted record
  // in one string, and the record's five pieces as iovecs that point into it and into the batch arenas (no copy of the payload).
  std::string& hdr = s->fastq;
  hdr.clear();
  std::vector<size_t> hoff;     // per record: offset of "@uuid comment\n", of "\n+ comment\n", and the end
  std::vector<uint64_t> idx;    // emitted records
  uint64_t bases = 0;
  static const char NL = '\n';
  std::string comment;
  for (uint64_t i = 0; i < b->n_reads; ++i) {
    const uint64_t l = b->lengths[i];
    if (l <= 14) continue;  // mod.rs:192-194
    if (l < 2ULL * k) return sfail(s, RR_ERR_READ, "rr_sim_write_fastq: read shorter than 2k (the reference panics slicing it)");
    uint64_t lo = s->rng.next_u64(), hi = s->rng.next_u64();
    uint8_t name[16];
    for (int x = 0; x < 8; ++x) { name[x] = (uint8_t)(hi >> (56 - 8 * x)); name[8 + x] = (uint8_t)(lo >> (56 - 8 * x)); }
    comment.clear();
    append_description(s, comment, s->items[first + i], l, b->real_identity[i]);
    hoff.push_back(hdr.size());
    hdr += '@';
    uuid_v3_x500(name, hdr);
    hdr += ' ';
    hdr += comment;
    hdr += '\n';
    hoff.push_back(hdr.size());
    hdr += "\n+ ";
    hdr += comment;
    hdr += '\n';
    hoff.push_back(hdr.size());
    idx.push_back(i);
    bases += l - 2ULL * k;
  }
  const size_t nrec = idx.size();
  std::vector<iovec> iov(5 * nrec);
  std::vector<uint64_t> start(nrec + 1, 0);  // byte offset of every record inside this part
  for (size_t r = 0; r < nrec; ++r) {
    const uint64_t j = idx[r], n = b->lengths[j] - 2ULL * k;
    const char* h = hdr.data();
    iov[5 * r + 0] = {(void*)(h + hoff[3 * r]), hoff[3 * r + 1] - hoff[3 * r]};
    iov[5 * r + 1] = {(void*)(b->seq + b->offsets[j] + k), (size_t)n};
    iov[5 * r + 2] = {(void*)(h + hoff[3 * r + 1]), hoff[3 * r + 2] - hoff[3 * r + 1]};
    iov[5 * r + 3] = {(void*)(b->qual + b->offsets[j] + k), (size_t)n};
    iov[5 * r + 4] = {(void*)&NL, 1};
    start[r + 1] = start[r] + (hoff[3 * r + 2] - hoff[3 * r]) + 2 * n + 1;
  }
  // Pass 2.  One writev stream fills the page cache at 2-4 GB/s, and several pwritev streams into ONE file do no better (buffered
  // writes serialise on the inode lock).  What does scale is the copy itself: with rr_sim_set_writer_threads(> 1) a regular
  // file is grown to the end of this part, the part's byte range is mapped, and the threads copy contiguous ranges of records
  // into the mapping (page faults and copies of different ranges proceed in parallel).  Pipes, append-mode descriptors and
  // small parts go through writev at the file position.  Requires that no other process resizes the file meanwhile.
  uint64_t bytes = 0;
  const uint64_t total = start[nrec];
  int threads = s->writer_threads;
  off_t base = -1;
  if (threads > 1 && total >= (4u << 20)) {
    struct stat st;
    const int fl = ::fcntl(fd, F_GETFL);
    if (::fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && fl >= 0 && !(fl & O_APPEND) && (fl & O_ACCMODE) == O_RDWR) {
      base = ::lseek(fd, 0, SEEK_CUR);
      if (base >= 0 && st.st_size < base + (off_t)total && ::ftruncate(fd, base + (off_t)total) != 0) base = -1;
    }
  }
  char* map = nullptr;
  size_t map_len = 0, map_skew = 0;
  if (base >= 0) {
    const long page = ::sysconf(_SC_PAGESIZE);
    map_skew = (size_t)(base % page);
    map_len = map_skew + total;
    void* m = ::mmap(nullptr, map_len, PROT_READ | PROT_WRITE, MAP_SHARED, fd, base - (off_t)map_skew);
    if (m == MAP_FAILED) base = -1; else map = (char*)m;
  }
  if (base < 0) {
    if (const char* m = write_all(fd, iov.data(), iov.size(), -1, bytes)) return sfail(s, RR_ERR_ARG, std::string("rr_sim_write_fastq: writev failed: ") + m);
  } else {
    threads = (int)std::min<uint64_t>((uint64_t)threads, std::max<uint64_t>(1, total >> 20));
    std::vector<size_t> cut(threads + 1, nrec);
    cut[0] = 0;
    for (int t = 1; t < threads; ++t)
      cut[t] = (size_t)(std::lower_bound(start.begin(), start.end(), total * (uint64_t)t / (uint64_t)threads) - start.begin());
    for (int t = 1; t <= threads; ++t) cut[t] = std::max(cut[t], cut[t - 1]);
    std::vector<std::thread> pool;
    auto work = [&](int t) {
      char* dst = map + map_skew + start[cut[t]];
      for (size_t x = 5 * cut[t]; x < 5 * cut[t + 1]; ++x) {
        std::memcpy(dst, iov[x].iov_base, iov[x].iov_len);
        dst += iov[x].iov_len;
      }
    };
    for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    ::munmap(map, map_len);
    bytes = total;
    if (::lseek(fd, base + (off_t)total, SEEK_SET) < 0) return sfail(s, RR_ERR_ARG, std::string("rr_sim_write_fastq: lseek failed: ") + std::strerror(errno));
  }
  if (bytes != total) return sfail(s, RR_ERR_ARG, "rr_sim_write_fastq: short write");
  if (n_records) *n_records = nrec;
  if (n_bases) *n_bases = bases;
  if (n_bytes) *n_bytes = bytes;
  return RR_OK;
}

RR_API int rr_sim_fastq_bytes(rr_sim* s, uint64_t n, const uint64_t* lengths, const double* real_identity, uint32_t k, uint64_t first,
                              uint64_t* n_bytes, uint64_t* n_records, uint64_t* n_bases) {
  if (!s || (n && (!lengths || !real_identity))) return RR_ERR_ARG;
  if (first + n > s->items.size()) return sfail(s, RR_ERR_ARG, "rr_sim_fastq_bytes: range does not match the last feed");
  uint64_t bytes = 0, recs = 0, bases = 0;
  std::string comment;
  for (uint64_t i = 0; i < n; ++i) {
    const uint64_t l = lengths[i];
    if (l <= 14) continue;  // mod.rs:192-194
    if (l < 2ULL * k) return sfail(s, RR_ERR_READ, "rr_sim_fastq_bytes: read shorter than 2k (the reference panics slicing it)");
    comment.clear();
    append_description(s, comment, s->items[first + i], l, real_identity[i]);
    const uint64_t nb = l - 2ULL * k;
    bytes += 1 + 36 + 1 + comment.size() + 1 + nb + 3 + comment.size() + 1 + nb + 1;  // "@uuid comment\n" seq "\n+ comment\n" qual "\n"
    ++recs;
    bases += nb;
  }
  if (n_bytes) *n_bytes = bytes;
  if (n_records) *n_records = recs;
  if (n_bases) *n_bases = bases;
  return RR_OK;
}

RR_API int rr_sim_skip_names(rr_sim* s, uint64_t n_records) {
  if (!s) return RR_ERR_ARG;
  for (uint64_t i = 0; i < 2 * n_records; ++i) s->rng.next_u64();
  return RR_OK;
}

RR_API int rr_sim_fastq(rr_sim* s, const rr_batch* b, uint64_t first, const char** text, uint64_t* len, uint64_t* n_records, uint64_t* n_bases) {
  if (!s || !b || !text || !len) return RR_ERR_ARG;
  if (first + b->n_reads > s->items.size()) return sfail(s, RR_ERR_ARG, "rr_sim_fastq: batch does not match the last feed");
  const uint32_t k = b->k;
  std::string& o = s->fastq;
  o.clear();
  o.reserve((size_t)(2 * b->total_bases + 400 * b->n_reads + 64));
  uint64_t recs = 0, bases = 0;
  std::string comment;
  for (uint64_t i = 0; i < b->n_reads; ++i) {
    const uint64_t l = b->lengths[i];
    if (l <= 14) continue;  // mod.rs:192-194 (a literal 14)
    if (l < 2ULL * k) return sfail(s, RR_ERR_READ, "rr_sim_fastq: read shorter than 2k (the reference panics slicing it)");
    uint64_t lo = s->rng.next_u64(), hi = s->rng.next_u64();  // main_rng.gen::<u128>() (mod.rs:201): low half first
    uint8_t name[16];
    for (int x = 0; x < 8; ++x) { name[x] = (uint8_t)(hi >> (56 - 8 * x)); name[8 + x] = (uint8_t)(lo >> (56 - 8 * x)); }  // to_be_bytes
    comment.clear();
    append_description(s, comment, s->items[first + i], l, b->real_identity[i]);
    const char* sq = (const char*)b->seq + b->offsets[i] + k;
    const char* ql = (const char*)b->qual + b->offsets[i] + k;
    const uint64_t n = l - 2ULL * k;
    o += '@';
    uuid_v3_x500(name, o);
    o += ' ';
    o += comment;
    o += '\n';
    o.append(sq, n);
    o += "\n+ ";
    o += comment;
    o += '\n';
    o.append(ql, n);
    o += '\n';
    ++recs;
    bases += n;
  }
  *text = o.data();
  *len = o.size();
  if (n_records) *n_records = recs;
  if (n_bases) *n_bases = bases;
  return RR_OK;
}
