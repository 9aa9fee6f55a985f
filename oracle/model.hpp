// ORACLE — TEST INFRASTRUCTURE ONLY (not part of the product path).
//
// Restatement of the reference's model objects:
//   Length    src/model/length.rs:13-47
//   Identity  src/model/identity.rs:14-54
//   Adapter   src/model/adapter.rs:12-97      (incl. the double "/100" quirk at :34-38)
//   Glitch    src/model/glitch.rs:12-64
//   Error     src/model/error.rs:19-135
//   Quality   src/model/quality.rs:19-138
//   random_base / random_base_diff / random_seq   src/lib.rs:15-45
#pragma once
#include <memory>
#include <sstream>
#include <string>
#include <unordered_map>

#include "align.hpp"
#include "rng.hpp"

namespace orc {

// src/lib.rs:15 — note the order
static const uint8_t NUCS[4] = {'A', 'C', 'T', 'G'};

inline uint8_t random_base(StdRng& rng) { return NUCS[rng.gen_range_u64(0, 3)]; }  // lib.rs:18-23
inline uint8_t random_base_diff(uint8_t nuc, StdRng& rng) {                          // lib.rs:26-37
  for (;;) {
    uint64_t idx = rng.gen_range_u64(0, 3);
    if (NUCS[idx] != nuc) return NUCS[idx];
  }
}
inline Bytes random_seq(size_t length, StdRng& rng) {  // lib.rs:40-45
  Bytes s(length);
  for (size_t i = 0; i < length; ++i) s[i] = random_base(rng);
  return s;
}

struct Length {  // src/model/length.rs
  double mean;
  bool has_dist;
  Gamma dist;
  Length(double mean_, double stdev) : mean(mean_), has_dist(false) {
    if (mean <= 0.0 || stdev < 0.0) throw std::runtime_error("Length: parameters must be > 0");
    if (stdev != 0.0) {
      double k = std::pow(mean, 2.0) / std::pow(stdev, 2.0);
      double t = std::pow(stdev, 2.0) / mean;
      dist = Gamma(k, t);
      has_dist = true;
    }
  }
  uint64_t get_length(StdRng& rng) const {
    double v = has_dist ? dist.sample(rng) : mean;
    return (uint64_t)std::round(v);
  }
};

struct Identity {  // src/model/identity.rs
  double mean, max;
  bool has_dist;
  Beta dist;
  Identity(double mean_, double max_, double stdev) : has_dist(false) {
    if (mean_ <= 0.0 || stdev <= 0.0) throw std::runtime_error("Identity: parameters must be > 0");
    mean = mean_ / 100.0;
    max = max_ / 100.0;
    stdev /= 100.0;
    if (std::fabs(mean - max) > 2.220446049250313e-16) {
      double beta_a = (((1.0 - (mean / max)) / std::pow(stdev / max, 2.0)) - (max / mean)) * std::pow(mean / max, 2.0);
      double beta_b = beta_a * ((max / mean) - 1.0);
      dist = Beta(beta_a, beta_b);
      has_dist = true;
    }
  }
  double get_identity(StdRng& rng) const { return has_dist ? max * dist.sample(rng) : mean; }
};

struct Adapter {  // src/model/adapter.rs
  Bytes start, end;
  double start_rate, end_rate;
  bool has_start, has_end;
  Beta start_dist, end_dist;
  Adapter(const Bytes& s, const Bytes& e, double srate, double samount, double erate, double eamount)
      : start(s), end(e), has_start(false), has_end(false) {
    if (!(start.empty() || srate == 0.0 || samount == 0.0)) {
      samount /= 100.0;
      start_dist = Beta(2.0 * samount, 2.0 - 2.0 * (samount / 100.0));  // adapter.rs:34-38 quirk
      has_start = true;
    }
    if (!(end.empty() || erate == 0.0 || eamount == 0.0)) {
      eamount /= 100.0;
      end_dist = Beta(2.0 * eamount, 2.0 - 2.0 * eamount);
      has_end = true;
    }
    start_rate = srate / 100.0;
    end_rate = erate / 100.0;
  }
  Bytes get_start(StdRng& rng) const {
    if (rng.gen_bool(start_rate) && has_start) {
      size_t n = (size_t)std::round((double)start.size() * start_dist.sample(rng));
      return Bytes(start.begin(), start.begin() + n);
    }
    return Bytes();
  }
  Bytes get_end(StdRng& rng) const {
    if (rng.gen_bool(end_rate) && has_end) {
      size_t n = (size_t)std::round((double)end.size() * end_dist.sample(rng));
      return Bytes(end.begin(), end.begin() + n);
    }
    return Bytes();
  }
  size_t max_len() const { return start.size() + end.size(); }
};

struct Glitch {  // src/model/glitch.rs
  bool has_distance;
  Geometric distance, length_size, length_skip;
  Glitch(double rate, double size, double skip) : has_distance(false) {
    if (rate != 0.0) {
      distance = Geometric(rate > 1.0 ? 1.0 / rate : 1.0);
      has_distance = true;
    }
    length_size = Geometric(size > 1.0 ? 1.0 / size : 1.0);
    length_skip = Geometric(skip > 1.0 ? 1.0 / skip : 1.0);
  }
  // glitch.rs:51-64 — draw order: distance, skip, size, bases
  bool get_glitch(StdRng& rng, size_t& begin, size_t& end, Bytes& seq) const {
    if (!has_distance) return false;
    begin = (size_t)distance.sample(rng);
    end = begin + (size_t)length_skip.sample(rng);
    seq = random_seq((size_t)length_size.sample(rng), rng);
    return true;
  }
};

// FxHash-flavoured string hash so that the CPU baseline is not handicapped by
// std::hash (the reference uses rustc_hash::FxHashMap).
struct BytesHash {
  size_t operator()(const std::string& s) const {
    uint64_t h = 0;
    for (unsigned char c : s) h = (((h << 5) | (h >> 59)) ^ c) * 0x517cc1b727220a95ULL;
    return (size_t)h;
  }
};

// minimal csv reader matching csv::ReaderBuilder{delimiter b';', no headers,
// flexible}: records split on \n / \r\n, empty lines skipped.
inline std::vector<std::vector<std::string>> read_csv(const std::string& text) {
  std::vector<std::vector<std::string>> recs;
  size_t i = 0, n = text.size();
  while (i < n) {
    size_t e = text.find('\n', i);
    if (e == std::string::npos) e = n;
    std::string line = text.substr(i, e - i);
    if (!line.empty() && line.back() == '\r') line.pop_back();
    i = e + 1;
    if (line.empty()) continue;
    std::vector<std::string> f;
    size_t p = 0;
    for (;;) {
      size_t q = line.find(';', p);
      if (q == std::string::npos) { f.push_back(line.substr(p)); break; }
      f.push_back(line.substr(p, q - p));
      p = q + 1;
    }
    recs.push_back(f);
  }
  return recs;
}

inline Bytes random_error(const uint8_t* kmer, size_t klen, StdRng& rng) {  // model/error.rs:116-135
  size_t error_pos = (size_t)rng.gen_range_u64(0, klen - 1);
  Bytes nk;
  nk.reserve(klen + 1);
  nk.insert(nk.end(), kmer, kmer + error_pos);
  switch (rng.gen_range_u32(1, 3)) {  // i32 range -> u32 draws
    case 1: nk.push_back(random_base_diff(kmer[error_pos], rng)); break;
    case 2: nk.push_back(random_base(rng)); nk.push_back(kmer[error_pos]); break;
    default: break;
  }
  nk.insert(nk.end(), kmer + error_pos + 1, kmer + klen);
  return nk;
}

struct ErrorModel {  // src/model/error.rs
  size_t length;
  bool is_random;
  struct Entry { std::vector<Bytes> alts; std::vector<uint64_t> edit; std::vector<double> prob; };
  std::unordered_map<std::string, Entry, BytesHash> data;
  std::vector<std::string> key_order;  // insertion order, for exporting to device tables

  static ErrorModel random(size_t k) { ErrorModel m; m.length = k; m.is_random = true; return m; }

  static ErrorModel from_text(const std::string& text, StdRng& rng) {  // error.rs:26-82
    ErrorModel m;
    m.is_random = false;
    m.length = 0;
    for (auto& rec : read_csv(text)) {
      Entry e;
      for (auto& row : rec) {
        if (row.empty()) continue;
        size_t c = row.find(',');
        if (c == std::string::npos) throw std::runtime_error("error model: parse error");
        std::string kmer = row.substr(0, c);
        size_t c2 = row.find(',', c + 1);
        double p = std::strtod(row.substr(c + 1, c2 == std::string::npos ? std::string::npos : c2 - c - 1).c_str(), nullptr);
        uint64_t dist = e.alts.empty() ? 0
            : edit_distance(e.alts[0].data(), e.alts[0].size(), (const uint8_t*)kmer.data(), kmer.size());
        e.alts.push_back(Bytes(kmer.begin(), kmer.end()));
        e.edit.push_back(dist);
        e.prob.push_back(p);
      }
      if (e.alts.empty()) continue;
      double sum = 0.0;
      for (double p : e.prob) sum += p;
      if (sum < 1.0) {
        e.alts.push_back(random_error(e.alts[0].data(), e.alts[0].size(), rng));
        e.edit.push_back(1);
        e.prob.push_back(1.0 - sum);
      }
      Bytes key = e.alts[0];
      e.alts.erase(e.alts.begin());
      e.edit.erase(e.edit.begin());
      e.prob.erase(e.prob.begin());
      m.length = key.size();
      std::string ks(key.begin(), key.end());
      if (!m.data.count(ks)) m.key_order.push_back(ks);
      m.data[ks] = e;
    }
    return m;
  }

  size_t k() const { return length; }

  Bytes add_errors_to_kmer(const uint8_t* kmer, size_t klen, StdRng& rng) const {  // error.rs:93-107
    if (!is_random) {
      auto it = data.find(std::string((const char*)kmer, klen));
      if (it != data.end()) {
        WeightedIndex dist(it->second.prob);  // rebuilt per call, as in the reference
        return it->second.alts[dist.sample(rng)];
      }
    }
    return random_error(kmer, klen, rng);
  }
};

struct QualityModel {  // src/model/quality.rs
  size_t max_k_;
  struct Entry { std::vector<uint8_t> scores; std::vector<double> weights; };
  std::unordered_map<std::string, Entry, BytesHash> data;
  std::vector<std::string> key_order;

  void insert(const std::string& k, int lo, int hi) {
    Entry e;
    for (int s = lo; s <= hi; ++s) { e.scores.push_back((uint8_t)s); e.weights.push_back(1.0); }
    if (!data.count(k)) key_order.push_back(k);
    data[k] = e;
  }
  static QualityModel random() {  // quality.rs:80-91
    QualityModel m; m.max_k_ = 1;
    m.insert("=", 1, 20); m.insert("X", 1, 20); m.insert("I", 1, 20);
    return m;
  }
  static QualityModel ideal() {  // quality.rs:94-110
    QualityModel m; m.max_k_ = 9;
    m.insert("X", 1, 3); m.insert("I", 1, 3);
    m.insert("=", 4, 7); m.insert("===", 8, 20); m.insert("=====", 21, 30);
    m.insert("=======", 31, 40); m.insert("=========", 41, 50);
    return m;
  }
  static QualityModel from_text(const std::string& text) {  // quality.rs:26-77
    QualityModel m; m.max_k_ = 0;
    auto recs = read_csv(text);
    for (size_t r = 1; r < recs.size(); ++r) {  // first record ("overall") skipped
      auto& rec = recs[r];
      if (rec.size() < 3) throw std::runtime_error("quality model: parse error");
      const std::string& cigar = rec[0];
      Entry e;
      const std::string& sw = rec[2];
      size_t p = 0;
      while (p <= sw.size()) {
        size_t q = sw.find(',', p);
        if (q == std::string::npos) q = sw.size();
        std::string row = sw.substr(p, q - p);
        p = q + 1;
        if (row.empty()) continue;
        size_t c = row.find(':');
        if (c == std::string::npos) throw std::runtime_error("quality model: parse error");
        e.scores.push_back((uint8_t)std::strtoul(row.substr(0, c).c_str(), nullptr, 10));
        size_t c2 = row.find(':', c + 1);
        e.weights.push_back(std::strtod(row.substr(c + 1, c2 == std::string::npos ? std::string::npos : c2 - c - 1).c_str(), nullptr));
      }
      if (cigar.size() > m.max_k_) m.max_k_ = cigar.size();
      if (!m.data.count(cigar)) m.key_order.push_back(cigar);
      m.data[cigar] = e;
    }
    if (!m.data.count("=") || !m.data.count("X") || !m.data.count("I"))
      throw std::runtime_error("quality model: must contain '=', 'X' and 'I'");
    return m;
  }
  size_t max_k() const { return max_k_; }

  // quality.rs:113-133 ; returns -1 / -2 for the two reference error variants
  int get_qscore(const uint8_t* cigar, size_t len, StdRng& rng) const {
    const uint8_t* c = cigar;
    size_t n = len;
    while (n > 0) {
      if (n % 2 == 0) return -1;  // QualityCigarLenNotOdd
      auto it = data.find(std::string((const char*)c, n));
      if (it != data.end()) {
        WeightedIndex dist(it->second.weights);  // rebuilt per call, as in the reference
        return it->second.scores[dist.sample(rng)] + 33;
      }
      c += 1;
      n -= 2;
    }
    return -2;  // QualityNotMinimalCigarString
  }
};

}  // namespace orc
