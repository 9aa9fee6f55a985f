// ORACLE — TEST INFRASTRUCTURE ONLY (not part of the product path).
//
// Restatement of the reference's `simulate` per-read pipeline and its feeder:
//   Change / Changes / add_change / sequence / add_glitches / add_error
//                                         src/simulate/error.rs:13-302
//   generate_quality                      src/simulate/quality.rs:14-38
//   ReadType / Origin / Description       src/simulate/description.rs:11-109
//   Reference / References                src/references.rs:17-171
//   Fragments                             src/simulate/fragments.rs:13-165
//   generate_read, add_fragment, ...      src/simulate/mod.rs:223-339
//   simulate (orchestrator + FASTQ writer) src/simulate/mod.rs:27-217
// All usize arithmetic wraps (Cargo.toml:58 overflow-checks = false).
#pragma once
#include <atomic>
#include <charconv>
#include <chrono>
#include <fstream>
#include <thread>

#include "md5.hpp"
#include "model.hpp"

namespace orc {

// ---------------------------------------------------------------------------
// src/simulate/error.rs
// ---------------------------------------------------------------------------
struct Change {
  size_t begin = 0, end = 0;
  Bytes seq, cigar;
  uint64_t edit_distance = 0;

  Change() {}
  Change(size_t b, size_t e, const Bytes& s) : begin(b), end(e), seq(s) {}
  static Change from_seq(size_t b, size_t e, const Bytes& s, const uint8_t* orig, size_t orig_len) {
    Change c(b, e, s);
    c.update_align(orig, orig_len);
    return c;
  }
  size_t end_raw() const { return end; }
  size_t end_err() const { return begin + seq.size(); }
  double edit() const { return (double)edit_distance; }

  bool contain(const Change& o) const {  // error.rs:83-88 (one clause duplicated in the reference)
    return o.end_raw() <= end_err() || o.end_raw() <= end_raw() || o.end_err() <= end_err() ||
           o.end_err() <= end_err();
  }
  bool overlap(const Change& o) const {  // error.rs:92-98
    if (contain(o)) return false;
    return o.begin < end_err() || o.begin < end_raw();
  }
  // error.rs:105-152 ; slices are resolved as (pointer,len); an out-of-range
  // slice panics (aborts) in the reference and throws here.
  double merge(const Change& o, const uint8_t* raw, size_t raw_len) {
    if (!overlap(o)) return 0.0;
    size_t rs, re;       // new_raw = raw[rs..re]
    size_t es;           // new_err = o.seq[es..]
    if (o.begin < end_err()) {
      size_t ovl = end_err() - o.begin;
      if (o.seq.size() > ovl) seq.insert(seq.end(), o.seq.begin() + ovl, o.seq.end());
      size_t old_end = end_raw();
      end = end_raw() + o.end_raw() - o.begin - ovl;
      if (end < raw_len - 1) {
        if (old_end < end) { rs = old_end; re = end; } else { rs = end; re = old_end; }
      } else { rs = old_end; re = raw_len; }
      es = ovl;
    } else {
      seq.insert(seq.end(), o.seq.begin(), o.seq.end());
      size_t old_end = end_raw();
      end = o.begin + o.seq.size();
      if (end < raw_len - 1) {
        if (old_end < end) { rs = old_end; re = end; } else { rs = end; re = old_end; }
      } else { rs = old_end; re = raw_len; }
      es = 0;
    }
    if (rs > re || re > raw_len) throw std::runtime_error("merge: raw slice out of range (reference panics)");
    if (es > o.seq.size()) throw std::runtime_error("merge: err slice out of range (reference panics)");
    Bytes cg;
    uint64_t ed = align(o.seq.data() + es, o.seq.size() - es, raw + rs, re - rs, cg);
    cigar.insert(cigar.end(), cg.begin(), cg.end());
    edit_distance += ed;
    return (double)ed;
  }
  void update_align(const uint8_t* orig, size_t orig_len) {  // error.rs:155-164
    size_t b = begin, e = (end_raw() > orig_len) ? orig_len : end_raw();
    if (b > e) throw std::runtime_error("update_align: slice out of range (reference panics)");
    edit_distance = align(seq.data(), seq.size(), orig + b, e - b, cigar);
  }
  bool operator==(const Change& o) const {
    return begin == o.begin && end == o.end && seq == o.seq && cigar == o.cigar && edit_distance == o.edit_distance;
  }
};
typedef std::vector<Change> Changes;

inline double add_change(Changes& cs, Change change, const uint8_t* raw, size_t raw_len) {  // error.rs:173-193
  if (!cs.empty()) {
    Change& last = cs.back();
    if (last.contain(change)) return 0.0;
    if (last.overlap(change)) return last.merge(change, raw, raw_len);
  }
  change.update_align(raw, raw_len);
  double e = change.edit();
  cs.push_back(std::move(change));
  return e;
}

inline double number_of_edit(double target, size_t length) { return (1.0 - target) * (double)length; }  // error.rs:196-198

inline void add_glitches(const uint8_t* raw, size_t raw_len, Changes& changes, const Glitch& model, StdRng& rng) {  // error.rs:250-264
  (void)raw;
  size_t position = 0, begin, end;
  Bytes seq;
  while (model.get_glitch(rng, begin, end, seq)) {
    position += begin;
    if (position > raw_len || position + end > raw_len) break;
    changes.push_back(Change(position, position + (end - begin), seq));
  }
}

inline void add_error(size_t k, double target, const uint8_t* raw, size_t raw_len, Changes& changes,
                      const ErrorModel& em, StdRng& rng) {  // error.rs:267-302
  Changes glitches(changes.rbegin(), changes.rend());
  changes.clear();
  if (raw_len < k) throw std::runtime_error("add_error: raw shorter than k (reference wraps)");
  for (size_t pos = 0; pos < raw_len - k; ++pos) {
    if (!glitches.empty() && pos == glitches.back().begin) {
      target -= add_change(changes, glitches.back(), raw, raw_len);
      glitches.pop_back();
    }
    if (target < 0.0 || target > (double)(raw_len - pos)) break;
    if (rng.gen_bool(target / (double)(raw_len - pos))) {
      Bytes kmer = em.add_errors_to_kmer(raw + pos, k, rng);
      target -= add_change(changes, Change(pos, pos + k, kmer), raw, raw_len);
    }
    if (target < 0.0 || target > (double)(raw_len - pos)) break;
  }
}

// error.rs:201-247
inline void sequence(double identity, const uint8_t* seq, size_t len, const ErrorModel& em, const Glitch& gm,
                     StdRng& rng, Bytes& err, Bytes& cig, double& real_id) {
  size_t k = em.k();
  double target = number_of_edit(identity, len);
  Changes changes;
  add_glitches(seq, len, changes, gm, rng);
  add_error(k, target, seq, len, changes, em, rng);
  double real_edit = 0.0;
  size_t pos_in_raw = 0;
  err.clear();
  cig.clear();
  err.reserve(len + len / 4);
  cig.reserve(len + len / 4);
  for (const Change& c : changes) {
    if (c.begin < pos_in_raw) continue;
    real_edit += c.edit();
    err.insert(err.end(), seq + pos_in_raw, seq + c.begin);
    cig.insert(cig.end(), c.begin - pos_in_raw, '=');
    err.insert(err.end(), c.seq.begin(), c.seq.end());
    cig.insert(cig.end(), c.cigar.begin(), c.cigar.end());
    pos_in_raw = c.end_raw();
  }
  if (pos_in_raw < len) {
    err.insert(err.end(), seq + pos_in_raw, seq + len);
    cig.insert(cig.end(), len - pos_in_raw, '=');
  }
  real_id = 1.0 - (real_edit / (double)len);
}

// ---------------------------------------------------------------------------
// src/simulate/quality.rs:14-38
// ---------------------------------------------------------------------------
inline int generate_quality(const Bytes& cigar, const QualityModel& model, StdRng& rng, Bytes& qual) {
  qual.clear();
  qual.reserve(cigar.size());
  size_t margin = (model.max_k() - 1) / 2;
  size_t n = cigar.size();
  for (size_t i = 0; i < n; ++i) {
    if (cigar[i] == 'D') continue;
    size_t start, end;
    if (i < margin) { start = 0; end = i + i + 1; }
    else if (i >= n - margin) { start = i - (n - i - 1); end = n; }
    else { start = i - margin; end = i + margin + 1; }
    if (end > n || start > end) throw std::runtime_error("generate_quality: window out of range (reference panics)");
    int q = model.get_qscore(cigar.data() + start, end - start, rng);
    if (q < 0) return q;
    qual.push_back((uint8_t)q);
  }
  return 0;
}

// ---------------------------------------------------------------------------
// src/simulate/description.rs
// ---------------------------------------------------------------------------
enum ReadType { REAL = 0, JUNK = 1, RANDOM = 2 };

struct Origin {
  std::string ref_id;
  char strand = '*';
  size_t start = 0, end = 0;
  ReadType read_type = REAL;
  static Origin reference(const std::string& id, char strand, size_t s, size_t e) {
    Origin o; o.ref_id = id; o.strand = strand; o.start = s; o.end = e; o.read_type = REAL; return o;
  }
  static Origin junk(size_t len) { Origin o; o.end = len; o.read_type = JUNK; return o; }
  static Origin random(size_t len) { Origin o; o.end = len; o.read_type = RANDOM; return o; }
  std::string to_string() const {  // description.rs:59-71
    if (read_type == JUNK) return "junk_seq";
    if (read_type == RANDOM) return "random_seq";
    return ref_id + "," + strand + "strand," + std::to_string(start) + "-" + std::to_string(end);
  }
};

// Rust `{}` for f64: shortest round-trip digits, never an exponent
inline std::string fmt_f64(double v) {
  if (std::isnan(v)) return "NaN";
  if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
  char buf[512];
  auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
  return std::string(buf, r.ptr);
}

struct Description {
  Origin origin;
  bool has_chimera = false;
  Origin chimera;
  size_t length = 0;
  double identity = 0.0;
  std::string to_string() const {  // description.rs:93-109
    std::string s;
    if (has_chimera) s = origin.to_string() + " chimera " + chimera.to_string() + " ";
    else s = origin.to_string() + " ";
    s += "length=" + std::to_string((uint64_t)(origin.end - origin.start));  // wraps like usize
    s += " error-free_length=" + std::to_string(length);
    s += " read_identity=" + fmt_f64(identity) + "%";
    return s;
  }
};

// ---------------------------------------------------------------------------
// src/references.rs
// ---------------------------------------------------------------------------
inline Bytes revcomp(const Bytes& s) {  // bio::alphabets::dna::revcomp (IUPAC aware, case preserving)
  static uint8_t comp[256];
  static bool init = false;
  if (!init) {
    for (int i = 0; i < 256; ++i) comp[i] = (uint8_t)i;
    const char* a = "AGCTYRWSKMDVHBN";
    const char* b = "TCGARYWSMKHBDVN";
    for (int i = 0; a[i]; ++i) {
      comp[(uint8_t)a[i]] = (uint8_t)b[i];
      comp[(uint8_t)(a[i] + 32)] = (uint8_t)(b[i] + 32);
    }
    init = true;
  }
  Bytes r(s.size());
  for (size_t i = 0; i < s.size(); ++i) r[i] = comp[s[s.size() - 1 - i]];
  return r;
}

struct Reference {
  std::string id;
  Bytes seq, rc;
  bool circular;
};

struct References {
  std::vector<Reference> sequences;
  std::vector<double> prob;
  std::unique_ptr<WeightedIndex> dist;

  // references.rs:135-171 ; header id = first word, desc = rest (bio fasta reader)
  static void read_reference(const std::string& text, std::vector<Reference>& seqs, std::vector<double>& prob) {
    size_t i = 0, n = text.size();
    Reference cur;
    std::string desc;
    bool have = false;
    auto flush = [&]() {
      if (!have) return;
      double weight = 1.0;
      size_t p = desc.find("depth=");
      while (p != std::string::npos) {  // regex depth=([\d.]+): first match with >= 1 char
        size_t q = p + 6, e = q;
        while (e < desc.size() && (std::isdigit((unsigned char)desc[e]) || desc[e] == '.')) ++e;
        if (e > q) { weight = std::strtod(desc.substr(q, e - q).c_str(), nullptr); break; }
        p = desc.find("depth=", p + 1);
      }
      cur.circular = desc.find("circular=true") != std::string::npos;
      cur.rc = revcomp(cur.seq);
      prob.push_back(weight * (double)cur.seq.size());
      seqs.push_back(cur);
    };
    while (i < n) {
      size_t e = text.find('\n', i);
      if (e == std::string::npos) e = n;
      std::string line = text.substr(i, e - i);
      i = e + 1;
      while (!line.empty() && std::isspace((unsigned char)line.back())) line.pop_back();
      if (!line.empty() && line[0] == '>') {
        flush();
        have = true;
        cur = Reference();
        std::string h = line.substr(1);
        size_t sp = 0;
        while (sp < h.size() && !std::isspace((unsigned char)h[sp])) ++sp;
        cur.id = h.substr(0, sp);
        desc = sp < h.size() ? h.substr(sp + 1) : std::string();
      } else if (have) {
        cur.seq.insert(cur.seq.end(), line.begin(), line.end());
      }
    }
    flush();
  }
  static References from_text(const std::string& text) {  // references.rs:49-59
    References r;
    read_reference(text, r.sequences, r.prob);
    r.dist.reset(new WeightedIndex(r.prob));
    return r;
  }
  // references.rs:62-80, 95-132
  static References from_text_adjusted_weight(const std::string& text, bool small_plasmid_bias, const Length& lm, StdRng& rng) {
    References r;
    read_reference(text, r.sequences, r.prob);
    std::vector<uint64_t> lengths(100000);
    for (auto& l : lengths) l = lm.get_length(rng);
    uint64_t total = 0;
    for (auto l : lengths) total += l;
    for (size_t i = 0; i < r.sequences.size(); ++i) {
      const Reference& ref = r.sequences[i];
      uint64_t rl = ref.seq.size();
      if (!small_plasmid_bias && ref.circular) {
        uint64_t passing = 0;
        for (auto l : lengths) if (l <= rl) passing += l;
        if (passing == 0) throw std::runtime_error("small plasmid bias: no fragment fits");
        r.prob[i] *= (double)total / (double)passing;
      }
      if (!ref.circular) {
        uint64_t passing = 0;
        for (auto l : lengths) passing += std::min(l, rl);
        r.prob[i] *= (double)total / (double)passing;
      }
    }
    r.dist.reset(new WeightedIndex(r.prob));
    return r;
  }
  // references.rs:83-92 — strand is an i32 range (u32 draw), then the weighted index
  std::pair<size_t, char> choose_reference(StdRng& rng) const {
    char strand = rng.gen_range_u32(0, 1) == 0 ? '+' : '-';
    size_t idx = dist->sample(rng);
    return {idx, strand};
  }
};

// ---------------------------------------------------------------------------
// src/simulate/fragments.rs
// ---------------------------------------------------------------------------
struct WorkItem {
  size_t ref_idx = 0, ref_idx2 = 0;
  Description desc;
  uint64_t seed = 0;
};

struct Fragments {
  uint64_t target;
  double junk_rate, random_rate, chimera_rate;
  const References& refs;
  const Length& lm;
  const Identity& im;
  StdRng& rng;
  Fragments(uint64_t target_, double junk, double random, double chimera, const References& r, const Length& l,
            const Identity& i, StdRng& g)
      : target(target_), junk_rate(junk / 100.0), random_rate(random / 100.0), chimera_rate(chimera / 100.0),
        refs(r), lm(l), im(i), rng(g) {}

  ReadType get_read_type() {  // fragments.rs:56-64
    if (rng.gen_bool(junk_rate)) return JUNK;
    if (rng.gen_bool(random_rate)) return RANDOM;
    return REAL;
  }
  bool is_chimera() { return rng.gen_bool(chimera_rate); }

  void generate_fragment(size_t& ref_index, size_t& real_length, Origin& origin) {  // fragments.rs:72-114
    ReadType rt = get_read_type();
    size_t length = (size_t)lm.get_length(rng);
    if (rt == REAL) {
      {  // the reference never leaves the loop below when no selectable contig can host the fragment: throw instead
        bool possible = false;
        for (size_t ri = 0; ri < refs.sequences.size(); ++ri) {
          const Reference& r = refs.sequences[ri];
          if (refs.prob[ri] > 0.0 && !(length >= r.seq.size() && r.circular)) { possible = true; break; }
        }
        if (!possible) throw std::runtime_error("generate_fragment: fragment fits no contig (all circular and shorter): the reference loops forever");
      }
      auto ch = refs.choose_reference(rng);
      const Reference* ref = &refs.sequences[ch.first];
      while (length >= ref->seq.size() && ref->circular) {  // !fragment_is_possible
        ch = refs.choose_reference(rng);
        ref = &refs.sequences[ch.first];
      }
      size_t rl = ref->seq.size();
      size_t try_begin = (size_t)rng.gen_range_u64(0, rl - 1);
      size_t begin, end;
      if (try_begin + length < rl) { begin = try_begin; end = try_begin + length; real_length = length; }
      else if (ref->circular) { begin = try_begin; end = length - (rl - try_begin); real_length = length; }
      else { begin = try_begin; end = rl - 1; real_length = rl - try_begin; }
      ref_index = ch.first;
      origin = Origin::reference(ref->id, ch.second, begin, end);
    } else if (rt == JUNK) {
      ref_index = 0; real_length = length; origin = Origin::junk(length);
    } else {
      ref_index = 0; real_length = length; origin = Origin::random(length);
    }
  }

  bool next(WorkItem& w) {  // fragments.rs:132-164
    if (target == 0) return false;
    size_t l1 = 0, l2 = 0;
    w = WorkItem();
    generate_fragment(w.ref_idx, l1, w.desc.origin);
    if (is_chimera()) {
      generate_fragment(w.ref_idx2, l2, w.desc.chimera);
      w.desc.has_chimera = true;
    }
    size_t tt = l1 + l2;
    if ((uint64_t)tt > target) target = 0; else target -= tt;
    w.desc.length = tt;
    w.desc.identity = im.get_identity(rng);
    w.seed = rng.next_u64();
    return true;
  }
};

// ---------------------------------------------------------------------------
// src/simulate/mod.rs:223-339
// ---------------------------------------------------------------------------
static const double CHIMERA_START_ADAPTER_CHANCE = 0.25;  // lib.rs:12
static const double CHIMERA_END_ADAPTER_CHANCE = 0.25;    // lib.rs:13

inline void add_junk(Bytes& raw, size_t length, StdRng& rng) {  // mod.rs:310-317
  size_t unit = (size_t)rng.gen_range_u64(1, 5);
  Bytes small = random_seq(unit, rng);
  size_t reps = length / small.size();
  for (size_t r = 0; r < reps; ++r) raw.insert(raw.end(), small.begin(), small.end());
}
inline void add_random(Bytes& raw, size_t length, StdRng& rng) {  // mod.rs:319-324
  Bytes s = random_seq(length, rng);
  raw.insert(raw.end(), s.begin(), s.end());
}
inline void add_real_fragment(Bytes& raw, const Origin& o, const Reference& ref) {  // mod.rs:326-339
  const Bytes& local = (o.strand == '+') ? ref.seq : ref.rc;
  if (o.start < o.end) {
    if (o.end > local.size()) throw std::runtime_error("add_real_fragment: slice out of range");
    raw.insert(raw.end(), local.begin() + o.start, local.begin() + o.end);
  } else if (ref.circular) {
    raw.insert(raw.end(), local.begin() + o.start, local.end());
    raw.insert(raw.end(), local.begin(), local.begin() + o.end);
  }
}
inline void add_fragment(Bytes& raw, const Origin& o, const Reference& ref, StdRng& rng) {  // mod.rs:295-308
  switch (o.read_type) {
    case JUNK: add_junk(raw, o.end, rng); break;
    case RANDOM: add_random(raw, o.end, rng); break;
    default: add_real_fragment(raw, o, ref); break;
  }
}

struct Models {
  const Adapter& adapter;
  const ErrorModel& error;
  const Glitch& glitch;
  const QualityModel& qscore;
};

// Builds the raw (error-free) fragment: mod.rs:243-271
inline void build_raw_fragment(const Reference& r0, const Reference& r1, const Description& d, const Models& m,
                               StdRng& rng, Bytes& raw) {
  size_t k = m.error.k();
  raw.clear();
  Bytes t = random_seq(k, rng);
  raw.insert(raw.end(), t.begin(), t.end());
  t = m.adapter.get_start(rng);
  raw.insert(raw.end(), t.begin(), t.end());
  add_fragment(raw, d.origin, r0, rng);
  if (d.has_chimera) {
    if (rng.gen_bool(CHIMERA_END_ADAPTER_CHANCE)) { t = m.adapter.get_end(rng); raw.insert(raw.end(), t.begin(), t.end()); }
    if (rng.gen_bool(CHIMERA_START_ADAPTER_CHANCE)) { t = m.adapter.get_start(rng); raw.insert(raw.end(), t.begin(), t.end()); }
    add_fragment(raw, d.chimera, r1, rng);
  }
  t = m.adapter.get_end(rng);
  raw.insert(raw.end(), t.begin(), t.end());
  t = random_seq(k, rng);
  raw.insert(raw.end(), t.begin(), t.end());
}

// mod.rs:223-293 ; returns 0 or the (negative) get_qscore error
inline int generate_read(const Reference& r0, const Reference& r1, Description& d, const Models& m, StdRng rng,
                         Bytes& err, Bytes& qual, Bytes* cigar_out = nullptr, Bytes* raw_out = nullptr) {
  Bytes raw, cigar;
  build_raw_fragment(r0, r1, d, m, rng, raw);
  double real_id;
  sequence(d.identity, raw.data(), raw.size(), m.error, m.glitch, rng, err, cigar, real_id);
  int rc = generate_quality(cigar, m.qscore, rng, qual);
  if (rc < 0) return rc;
  if (qual.size() != err.size()) qual.resize(err.size(), '!');
  d.identity = real_id * 100.0;
  d.length = err.size();
  if (cigar_out) *cigar_out = cigar;
  if (raw_out) *raw_out = raw;
  return 0;
}

// ---------------------------------------------------------------------------
// src/cli/simulate.rs:12-125 (Quantity) and src/simulate/mod.rs:27-217
// ---------------------------------------------------------------------------
struct Quantity {
  uint64_t value = 0;
  bool coverage = false;
  static Quantity parse(const std::string& s) {  // cli/simulate.rs:12-80
    Quantity q;
    if (s.empty()) throw std::runtime_error("quantity: empty");
    char last = s.back();
    std::string num = s;
    uint64_t mul = 1;
    if (last == 'x') { q.coverage = true; num.pop_back(); }
    else if (last == 'K') { mul = 1000ULL; num.pop_back(); }
    else if (last == 'M') { mul = 1000000ULL; num.pop_back(); }
    else if (last == 'G') { mul = 1000000000ULL; num.pop_back(); }
    if (num.empty() || num.find_first_not_of("0123456789") != std::string::npos) throw std::runtime_error("quantity: parse error");
    q.value = std::strtoull(num.c_str(), nullptr, 10) * mul;
    return q;
  }
  uint64_t number_of_base(uint64_t genome) const { return coverage ? value * genome : value; }
};

struct SimParams {
  std::string reference_text;
  uint64_t seed = 42;
  double length_mean = 15000, length_stdev = 13000;
  double id_mean = 85, id_max = 95, id_stdev = 5;
  std::string error_model = "random";   // "random" or model text prefixed by "text:"
  std::string qscore_model = "random";  // "random", "ideal" or "text:..."
  double start_rate = 90, start_amount = 60, end_rate = 50, end_amount = 20;
  std::string start_adapter_seq = "AATGTACTTCGTTCAGTTACGTATTGCT";
  std::string end_adapter_seq = "GCAATACGTAACTGAACGAAGT";
  double junk = 1, random = 1, chimera = 1;
  double glitch_rate = 10000, glitch_size = 25, glitch_skip = 25;
  bool small_plasmid_bias = false;
  std::string quantity = "1x";
  std::string nb_base_store;  // empty = none
  int threads = 1;
  std::string output_path;  // empty = discard (count only)
};

struct SimStats {
  uint64_t reads = 0, bases = 0, fastq_bytes = 0;
  double generate_seconds = 0.0, total_seconds = 0.0;
};

inline std::string format_record(const std::string& uuid, const Description& d, const Bytes& seq, const Bytes& qual, size_t k) {
  std::string c = d.to_string();
  std::string r;
  r.reserve(2 * seq.size() + 2 * c.size() + 64);
  r += "@"; r += uuid; r += " "; r += c; r += "\n";
  r.append((const char*)seq.data() + k, seq.size() - 2 * k);
  r += "\n+ "; r += c; r += "\n";
  r.append((const char*)qual.data() + k, seq.size() - 2 * k);
  r += "\n";
  return r;
}

inline SimStats simulate(const SimParams& p, std::string* capture = nullptr) {
  auto t0 = std::chrono::steady_clock::now();
  SimStats st;
  StdRng main_rng = StdRng::seed_from_u64(p.seed);
  Length length(p.length_mean, p.length_stdev);
  References refs = References::from_text_adjusted_weight(p.reference_text, p.small_plasmid_bias, length, main_rng);
  Identity identity(p.id_mean, p.id_max, p.id_stdev);
  Adapter adapter(Bytes(p.start_adapter_seq.begin(), p.start_adapter_seq.end()),
                  Bytes(p.end_adapter_seq.begin(), p.end_adapter_seq.end()), p.start_rate, p.start_amount, p.end_rate,
                  p.end_amount);
  Glitch glitches(p.glitch_rate, p.glitch_size, p.glitch_skip);
  ErrorModel error = (p.error_model == "random") ? ErrorModel::random(7) : ErrorModel::from_text(p.error_model.substr(5), main_rng);
  size_t k = error.k();
  QualityModel qscore = (p.qscore_model == "ideal")    ? QualityModel::ideal()
                        : (p.qscore_model == "random") ? QualityModel::random()
                                                       : QualityModel::from_text(p.qscore_model.substr(5));
  Models models{adapter, error, glitches, qscore};

  uint64_t len_ref = 0;
  for (auto& r : refs.sequences) len_ref += r.seq.size();
  uint64_t total_base = Quantity::parse(p.quantity).number_of_base(len_ref);
  uint64_t base_limit = p.nb_base_store.empty() ? total_base : Quantity::parse(p.nb_base_store).number_of_base(len_ref);
  uint64_t base_produce = 0;

  std::ofstream out;
  if (!p.output_path.empty()) out.open(p.output_path, std::ios::binary);

  struct Result { Description d; Bytes seq, qual; };
  while (base_produce < total_base) {
    uint64_t base_loop = (base_limit > total_base - base_produce) ? total_base - base_produce : base_limit;
    base_produce += base_loop;
    auto g0 = std::chrono::steady_clock::now();
    // serial feeder (the reference bridges this iterator to rayon under a mutex)
    std::vector<WorkItem> items;
    {
      Fragments fr(base_loop, p.junk, p.random, p.chimera, refs, length, identity, main_rng);
      WorkItem w;
      while (fr.next(w)) items.push_back(w);
    }
    std::vector<Result> results(items.size());
    std::atomic<size_t> next{0};
    auto worker = [&]() {
      for (;;) {
        size_t i = next.fetch_add(1);
        if (i >= items.size()) break;
        Result& r = results[i];
        r.d = items[i].desc;
        int rc = generate_read(refs.sequences[items[i].ref_idx], refs.sequences[items[i].ref_idx2], r.d, models,
                               StdRng::seed_from_u64(items[i].seed), r.seq, r.qual);
        if (rc < 0) { std::fprintf(stderr, "generate_read failed (%d); the reference aborts here\n", rc); std::abort(); }
      }
    };
    int nt = std::max(1, p.threads);
    if (nt == 1) worker();
    else {
      std::vector<std::thread> th;
      for (int t = 0; t < nt; ++t) th.emplace_back(worker);
      for (auto& t : th) t.join();
    }
    st.generate_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - g0).count();
    // serial writer, canonical (iterator) order
    for (auto& r : results) {
      if (r.seq.size() <= 14) continue;
      uint64_t lo = main_rng.next_u64(), hi = main_rng.next_u64();  // gen::<u128>(): low half first
      uint8_t name[16];
      for (int b = 0; b < 8; ++b) { name[b] = (uint8_t)(hi >> (56 - 8 * b)); name[8 + b] = (uint8_t)(lo >> (56 - 8 * b)); }
      std::string rec = format_record(uuid_v3_x500(name, 16), r.d, r.seq, r.qual, k);
      st.reads += 1;
      st.bases += r.seq.size() - 2 * k;
      st.fastq_bytes += rec.size();
      if (out.is_open()) out.write(rec.data(), rec.size());
      if (capture) *capture += rec;
    }
  }
  st.total_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return st;
}

}  // namespace orc
