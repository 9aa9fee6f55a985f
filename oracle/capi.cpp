// ORACLE — TEST INFRASTRUCTURE ONLY (not part of the product path).
//
// C-ABI test hooks over the CPU restatement (rng.hpp, align.hpp, model.hpp,
// simulate.hpp) so that tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline leg can drive it through ctypes.  Nothing under rustyread_b200/
// links or loads this library.
#include <mutex>

#include "simulate.hpp"

using namespace orc;

#define ORC_API extern "C" __attribute__((visibility("default")))

static thread_local std::string g_err;
static char* dup_str(const std::string& s) {
  char* p = (char*)std::malloc(s.size() + 1);
  std::memcpy(p, s.data(), s.size());
  p[s.size()] = 0;
  return p;
}
static std::vector<std::string> split(const std::string& s, char sep) {
  std::vector<std::string> out;
  size_t p = 0;
  while (p <= s.size()) {
    size_t q = s.find(sep, p);
    if (q == std::string::npos) q = s.size();
    if (q > p) out.push_back(s.substr(p, q - p));
    p = q + 1;
  }
  return out;
}
#define ORC_BEGIN try {
#define ORC_END(fail) } catch (const std::exception& e) { g_err = e.what(); return fail; }

ORC_API const char* orc_last_error() { return g_err.c_str(); }
ORC_API void orc_free(void* p) { std::free(p); }

// ---- rng -------------------------------------------------------------------
ORC_API void* orc_rng_new(uint64_t seed) { return new StdRng(StdRng::seed_from_u64(seed)); }
ORC_API void orc_rng_free(void* r) { delete (StdRng*)r; }
ORC_API uint32_t orc_rng_next_u32(void* r) { return ((StdRng*)r)->next_u32(); }
ORC_API uint64_t orc_rng_next_u64(void* r) { return ((StdRng*)r)->next_u64(); }
ORC_API uint64_t orc_rng_words(void* r) { return ((StdRng*)r)->words_consumed(); }
ORC_API double orc_rng_gen_f64(void* r) { return ((StdRng*)r)->gen_f64(); }
ORC_API int orc_rng_gen_bool(void* r, double p) { ORC_BEGIN return ((StdRng*)r)->gen_bool(p) ? 1 : 0; ORC_END(-1) }
ORC_API uint64_t orc_rng_gen_range_u64(void* r, uint64_t lo, uint64_t hi) { return ((StdRng*)r)->gen_range_u64(lo, hi); }
ORC_API uint32_t orc_rng_gen_range_u32(void* r, uint32_t lo, uint32_t hi) { return ((StdRng*)r)->gen_range_u32(lo, hi); }
ORC_API void orc_rng_words_block(uint64_t seed, uint64_t first_word, uint64_t n, uint32_t* out) {
  StdRng r = StdRng::seed_from_u64(seed);
  for (uint64_t i = 0; i < first_word; ++i) r.next_u32();
  for (uint64_t i = 0; i < n; ++i) out[i] = r.next_u32();
}
ORC_API int orc_random_base(void* r) { return random_base(*(StdRng*)r); }
ORC_API int orc_random_base_diff(void* r, int nuc) { return random_base_diff((uint8_t)nuc, *(StdRng*)r); }
ORC_API void orc_random_seq(void* r, uint64_t n, uint8_t* out) {
  Bytes s = random_seq(n, *(StdRng*)r);
  if (n) std::memcpy(out, s.data(), n);
}
ORC_API double orc_number_of_edit(double target, uint64_t len) { return number_of_edit(target, len); }

// ---- distributions ----------------------------------------------------------
ORC_API void* orc_weighted_new(const double* w, uint64_t n) { ORC_BEGIN return new WeightedIndex(w, n); ORC_END(nullptr) }
ORC_API void orc_weighted_free(void* d) { delete (WeightedIndex*)d; }
ORC_API uint64_t orc_weighted_sample(void* d, void* r) { return ((WeightedIndex*)d)->sample(*(StdRng*)r); }
ORC_API double orc_weighted_scale(void* d) { return ((WeightedIndex*)d)->scale; }
ORC_API uint64_t orc_geometric_sample(double p, void* r) { ORC_BEGIN return Geometric(p).sample(*(StdRng*)r); ORC_END(0) }
ORC_API void orc_geometric_params(double p, double* pi, uint64_t* k) { Geometric g(p); *pi = g.pi; *k = g.k; }
ORC_API double orc_beta_sample(double a, double b, void* r) { ORC_BEGIN return Beta(a, b).sample(*(StdRng*)r); ORC_END(-1.0) }
ORC_API double orc_gamma_sample(double shape, double scale, void* r) { ORC_BEGIN return Gamma(shape, scale).sample(*(StdRng*)r); ORC_END(-1.0) }
ORC_API double orc_standard_normal(void* r) { return standard_normal(*(StdRng*)r); }

// ---- models -----------------------------------------------------------------
ORC_API void* orc_length_new(double mean, double stdev) { ORC_BEGIN return new Length(mean, stdev); ORC_END(nullptr) }
ORC_API void orc_length_free(void* m) { delete (Length*)m; }
ORC_API uint64_t orc_length_get(void* m, void* r) { return ((Length*)m)->get_length(*(StdRng*)r); }
ORC_API void* orc_identity_new(double mean, double max, double stdev) { ORC_BEGIN return new Identity(mean, max, stdev); ORC_END(nullptr) }
ORC_API void orc_identity_free(void* m) { delete (Identity*)m; }
ORC_API double orc_identity_get(void* m, void* r) { return ((Identity*)m)->get_identity(*(StdRng*)r); }

ORC_API void* orc_adapter_new(const char* start, const char* end, double sr, double sa, double er, double ea) {
  ORC_BEGIN return new Adapter(Bytes(start, start + std::strlen(start)), Bytes(end, end + std::strlen(end)), sr, sa, er, ea); ORC_END(nullptr)
}
ORC_API void orc_adapter_free(void* m) { delete (Adapter*)m; }
ORC_API int orc_adapter_get(void* m, int which_end, void* r, uint8_t* out) {
  Bytes s = which_end ? ((Adapter*)m)->get_end(*(StdRng*)r) : ((Adapter*)m)->get_start(*(StdRng*)r);
  if (!s.empty()) std::memcpy(out, s.data(), s.size());
  return (int)s.size();
}
ORC_API uint64_t orc_adapter_max_len(void* m) { return ((Adapter*)m)->max_len(); }
// constructed Beta parameters, for handing to the device library
ORC_API void orc_adapter_params(void* m_, double* out) {
  Adapter* m = (Adapter*)m_;
  const Beta* bs[2] = {&m->start_dist, &m->end_dist};
  out[0] = m->start_rate; out[1] = m->end_rate; out[2] = m->has_start; out[3] = m->has_end;
  for (int i = 0; i < 2; ++i) {
    const Beta& b = *bs[i];
    double* o = out + 4 + 10 * i;
    o[0] = b.a; o[1] = b.b; o[2] = b.switched; o[3] = b.bb; o[4] = b.alpha; o[5] = b.beta; o[6] = b.gamma;
    o[7] = b.delta; o[8] = b.kappa1; o[9] = b.kappa2;
  }
}

ORC_API void* orc_glitch_new(double rate, double size, double skip) { ORC_BEGIN return new Glitch(rate, size, skip); ORC_END(nullptr) }
ORC_API void orc_glitch_free(void* m) { delete (Glitch*)m; }
// returns seq length or -1 if the model has no distance distribution
ORC_API int64_t orc_glitch_get(void* m, void* r, uint64_t* begin, uint64_t* end, uint8_t* out, uint64_t cap) {
  size_t b, e;
  Bytes s;
  if (!((Glitch*)m)->get_glitch(*(StdRng*)r, b, e, s)) return -1;
  *begin = b; *end = e;
  if (s.size() > cap) return -2;
  if (!s.empty()) std::memcpy(out, s.data(), s.size());
  return (int64_t)s.size();
}

ORC_API void* orc_error_model_random(uint64_t k) { return new ErrorModel(ErrorModel::random(k)); }
ORC_API void* orc_error_model_from_text(const char* text, void* r) {
  ORC_BEGIN return new ErrorModel(ErrorModel::from_text(text, *(StdRng*)r)); ORC_END(nullptr)
}
ORC_API void orc_error_model_free(void* m) { delete (ErrorModel*)m; }
ORC_API uint64_t orc_error_model_k(void* m) { return ((ErrorModel*)m)->k(); }
ORC_API int orc_error_model_add(void* m, void* r, const uint8_t* kmer, uint64_t klen, uint8_t* out) {
  Bytes s = ((ErrorModel*)m)->add_errors_to_kmer(kmer, klen, *(StdRng*)r);
  if (!s.empty()) std::memcpy(out, s.data(), s.size());
  return (int)s.size();
}
ORC_API int orc_random_error(void* r, const uint8_t* kmer, uint64_t klen, uint8_t* out) {
  Bytes s = random_error(kmer, klen, *(StdRng*)r);
  if (!s.empty()) std::memcpy(out, s.data(), s.size());
  return (int)s.size();
}
// dump of the loaded model: "KEY\tALT:edit:prob,ALT:edit:prob...\n" in insertion order
ORC_API char* orc_error_model_dump(void* m_) {
  ErrorModel* m = (ErrorModel*)m_;
  std::string s;
  char buf[64];
  for (auto& k : m->key_order) {
    auto& e = m->data[k];
    s += k + "\t";
    for (size_t i = 0; i < e.alts.size(); ++i) {
      if (i) s += ",";
      s.append((const char*)e.alts[i].data(), e.alts[i].size());
      std::snprintf(buf, sizeof buf, ":%llu:%.17g", (unsigned long long)e.edit[i], e.prob[i]);
      s += buf;
    }
    s += "\n";
  }
  return dup_str(s);
}

ORC_API void* orc_quality_model_new(const char* kind, const char* text) {
  ORC_BEGIN 
      std::string k = kind;
      if (k == "random") return new QualityModel(QualityModel::random());
      if (k == "ideal") return new QualityModel(QualityModel::ideal());
      return new QualityModel(QualityModel::from_text(text)); ORC_END(nullptr)
}
ORC_API void orc_quality_model_free(void* m) { delete (QualityModel*)m; }
ORC_API uint64_t orc_quality_model_max_k(void* m) { return ((QualityModel*)m)->max_k(); }
ORC_API int orc_get_qscore(void* m, void* r, const uint8_t* cigar, uint64_t len) {
  return ((QualityModel*)m)->get_qscore(cigar, len, *(StdRng*)r);
}
ORC_API int64_t orc_generate_quality(void* m, void* r, const uint8_t* cigar, uint64_t len, uint8_t* out) {
  ORC_BEGIN 
      Bytes q; Bytes c(cigar, cigar + len);
      int rc = generate_quality(c, *(QualityModel*)m, *(StdRng*)r, q);
      if (rc < 0) return rc;
      if (!q.empty()) std::memcpy(out, q.data(), q.size());
      return (int64_t)q.size(); ORC_END(-3)
}
// "CIGAR\tscore:weight,...\n" in insertion order
ORC_API char* orc_quality_model_dump(void* m_) {
  QualityModel* m = (QualityModel*)m_;
  std::string s;
  char buf[64];
  for (auto& k : m->key_order) {
    auto& e = m->data[k];
    s += k + "\t";
    for (size_t i = 0; i < e.scores.size(); ++i) {
      std::snprintf(buf, sizeof buf, "%s%d:%.17g", i ? "," : "", (int)e.scores[i], e.weights[i]);
      s += buf;
    }
    s += "\n";
  }
  return dup_str(s);
}

// ---- alignment ----------------------------------------------------------------
ORC_API uint64_t orc_edit_distance(const uint8_t* a, uint64_t n, const uint8_t* b, uint64_t m) { return edit_distance(a, n, b, m); }
ORC_API uint64_t orc_align(const uint8_t* q, uint64_t m, const uint8_t* t, uint64_t n, uint8_t* cigar, uint64_t* clen) {
  Bytes c;
  uint64_t e = align(q, m, t, n, c);
  if (!c.empty()) std::memcpy(cigar, c.data(), c.size());
  *clen = c.size();
  return e;
}

// ---- Change / Changes ---------------------------------------------------------
static std::string dump_changes(const Changes& cs) {
  std::string s;
  for (auto& c : cs) {
    s += std::to_string(c.begin) + "\t" + std::to_string(c.end) + "\t";
    s.append((const char*)c.seq.data(), c.seq.size());
    s += "\t";
    s.append((const char*)c.cigar.data(), c.cigar.size());
    s += "\t" + std::to_string(c.edit_distance) + "\n";
  }
  return s;
}
// A tiny script interface for the Change goldens (src/simulate/error.rs:317-621).
// ops: one per line,  "from_seq begin end SEQ" | "raw begin end SEQ CIGAR edit" (push onto the stack, SEQ/CIGAR '-' = empty)
//      "contain" / "overlap" / "merge" act on (stack[-2], stack[-1]) and pop the second for merge
//      "add_change" moves stack[-1] into the Changes list ; output: one line per op result, then "--" and dumps
ORC_API char* orc_change_script(const uint8_t* raw, uint64_t raw_len, const char* script) {
  ORC_BEGIN 
      std::vector<Change> stack; Changes list; std::string out;
      // NB: no iostream number extraction here — numpy's statically linked libstdc++ exports
      // GNU_UNIQUE locale ids that break std::num_get inside a ctypes-loaded library.
      for (auto& line : split(script, '\n')) {
        std::vector<std::string> tk = split(line, ' ');
        if (tk.empty()) continue;
        const std::string& op = tk[0];
        if (op == "from_seq" || op == "raw") {
          size_t b = std::strtoull(tk.at(1).c_str(), nullptr, 10), e = std::strtoull(tk.at(2).c_str(), nullptr, 10);
          std::string s = tk.at(3);
          if (s == "-") s.clear();
          Change c(b, e, Bytes(s.begin(), s.end()));
          if (op == "from_seq") c.update_align(raw, raw_len);
          else { std::string cg = tk.at(4); if (cg == "-") cg.clear(); c.cigar = Bytes(cg.begin(), cg.end()); c.edit_distance = std::strtoull(tk.at(5).c_str(), nullptr, 10); }
          stack.push_back(c);
        } else if (op == "contain") { out += stack[stack.size() - 2].contain(stack.back()) ? "1\n" : "0\n"; }
        else if (op == "overlap") { out += stack[stack.size() - 2].overlap(stack.back()) ? "1\n" : "0\n"; }
        else if (op == "merge") {
          Change o = stack.back(); stack.pop_back();
          double e = stack.back().merge(o, raw, raw_len);
          out += fmt_f64(e) + "\n";
        } else if (op == "add_change") {
          Change c = stack.back(); stack.pop_back();
          out += fmt_f64(add_change(list, c, raw, raw_len)) + "\n";
        }
      }
      out += "--\n" + dump_changes(stack) + "--\n" + dump_changes(list);
      return dup_str(out); ORC_END(nullptr)
}
ORC_API char* orc_add_glitches(void* r, void* glitch, const uint8_t* raw, uint64_t len, int update_align) {
  ORC_BEGIN 
      Changes cs; add_glitches(raw, len, cs, *(Glitch*)glitch, *(StdRng*)r);
      if (update_align) for (auto& c : cs) c.update_align(raw, len);
      return dup_str(dump_changes(cs)); ORC_END(nullptr)
}
ORC_API char* orc_add_error(void* r, void* em, uint64_t k, double target, const uint8_t* raw, uint64_t len) {
  ORC_BEGIN 
      Changes cs; add_error(k, target, raw, len, cs, *(ErrorModel*)em, *(StdRng*)r);
      return dup_str(dump_changes(cs)); ORC_END(nullptr)
}
// sequence(): outputs err, cigar (caller buffers of capacity cap), returns 0 / -1
ORC_API int orc_sequence(void* r, double identity, const uint8_t* raw, uint64_t len, void* em, void* gm, uint8_t* err,
                         uint64_t* err_len, uint8_t* cig, uint64_t* cig_len, double* real_id, uint64_t cap) {
  ORC_BEGIN 
      Bytes e; Bytes c;
      sequence(identity, raw, len, *(ErrorModel*)em, *(Glitch*)gm, *(StdRng*)r, e, c, *real_id);
      if (e.size() > cap || c.size() > cap) { g_err = "capacity"; return -2; }
      if (!e.empty()) std::memcpy(err, e.data(), e.size());
      if (!c.empty()) std::memcpy(cig, c.data(), c.size());
      *err_len = e.size(); *cig_len = c.size();
      return 0; ORC_END(-1)
}

// ---- description / misc ----------------------------------------------------------
ORC_API char* orc_fmt_f64(double v) { return dup_str(fmt_f64(v)); }
ORC_API char* orc_description(int type, const char* ref_id, int strand, uint64_t start, uint64_t end, int has_chim,
                              int ctype, const char* cref_id, int cstrand, uint64_t cstart, uint64_t cend,
                              uint64_t length, double identity) {
  Description d;
  d.origin.read_type = (ReadType)type; d.origin.ref_id = ref_id; d.origin.strand = (char)strand; d.origin.start = start; d.origin.end = end;
  d.has_chimera = has_chim;
  d.chimera.read_type = (ReadType)ctype; d.chimera.ref_id = cref_id; d.chimera.strand = (char)cstrand; d.chimera.start = cstart; d.chimera.end = cend;
  d.length = length; d.identity = identity;
  return dup_str(d.to_string());
}
ORC_API char* orc_uuid_v3_x500(const uint8_t* name, uint64_t len) { return dup_str(uuid_v3_x500(name, len)); }
ORC_API void orc_md5(const uint8_t* msg, uint64_t len, uint8_t* out) { md5(msg, len, out); }
ORC_API int64_t orc_quantity(const char* s, uint64_t genome) { ORC_BEGIN return (int64_t)Quantity::parse(s).number_of_base(genome); ORC_END(-1) }
ORC_API void orc_revcomp(const uint8_t* s, uint64_t n, uint8_t* out) {
  Bytes r = revcomp(Bytes(s, s + n));
  if (n) std::memcpy(out, r.data(), n);
}

// ---- workload: references + models + feeder + per-read generation ---------------------
struct Workload {
  StdRng main_rng;
  std::unique_ptr<Length> length;
  std::unique_ptr<Identity> identity;
  std::unique_ptr<Adapter> adapter;
  std::unique_ptr<Glitch> glitch;
  std::unique_ptr<ErrorModel> error;
  std::unique_ptr<QualityModel> qscore;
  References refs;
  double junk = 0, random = 0, chimera = 0;
  std::vector<WorkItem> items;
  struct Out { Description d; Bytes seq, qual, cigar, raw; int rc = 0; };
  std::vector<Out> outs;
  std::string text;
};

ORC_API void* orc_workload_new(uint64_t seed) {
  Workload* w = new Workload();
  w->main_rng = StdRng::seed_from_u64(seed);
  return w;
}
ORC_API void orc_workload_free(void* w) { delete (Workload*)w; }
ORC_API void* orc_workload_rng(void* w) { return &((Workload*)w)->main_rng; }
ORC_API int orc_workload_set_length(void* w, double mean, double stdev) { ORC_BEGIN ((Workload*)w)->length.reset(new Length(mean, stdev)); return 0; ORC_END(-1) }
ORC_API int orc_workload_set_identity(void* w, double mean, double max, double stdev) { ORC_BEGIN ((Workload*)w)->identity.reset(new Identity(mean, max, stdev)); return 0; ORC_END(-1) }
ORC_API int orc_workload_set_adapter(void* w, const char* s, const char* e, double sr, double sa, double er, double ea) {
  ORC_BEGIN ((Workload*)w)->adapter.reset(new Adapter(Bytes(s, s + std::strlen(s)), Bytes(e, e + std::strlen(e)), sr, sa, er, ea)); return 0; ORC_END(-1)
}
ORC_API int orc_workload_set_glitch(void* w, double rate, double size, double skip) { ORC_BEGIN ((Workload*)w)->glitch.reset(new Glitch(rate, size, skip)); return 0; ORC_END(-1) }
// text == NULL -> Error::random(7); otherwise Error::from_stream (consumes main_rng)
ORC_API int orc_workload_set_error_model(void* w_, const char* text) {
  Workload* w = (Workload*)w_;
  ORC_BEGIN w->error.reset(new ErrorModel(text ? ErrorModel::from_text(text, w->main_rng) : ErrorModel::random(7))); return 0; ORC_END(-1)
}
ORC_API int orc_workload_set_qscore_model(void* w_, const char* kind, const char* text) {
  Workload* w = (Workload*)w_;
  ORC_BEGIN 
      std::string k = kind;
      if (k == "random") w->qscore.reset(new QualityModel(QualityModel::random()));
      else if (k == "ideal") w->qscore.reset(new QualityModel(QualityModel::ideal()));
      else w->qscore.reset(new QualityModel(QualityModel::from_text(text)));
      return 0; ORC_END(-1)
}
// adjusted != 0 -> from_stream_adjusted_weight (needs the length model; consumes main_rng)
ORC_API int orc_workload_set_references(void* w_, const char* fasta, uint64_t len, int adjusted, int small_plasmid_bias) {
  Workload* w = (Workload*)w_;
  ORC_BEGIN 
      std::string text(fasta, len);
      if (adjusted) w->refs = References::from_text_adjusted_weight(text, small_plasmid_bias, *w->length, w->main_rng);
      else w->refs = References::from_text(text);
      return 0; ORC_END(-1)
}
ORC_API void orc_workload_set_rates(void* w, double junk, double random, double chimera) {
  ((Workload*)w)->junk = junk; ((Workload*)w)->random = random; ((Workload*)w)->chimera = chimera;
}
ORC_API uint64_t orc_workload_n_refs(void* w) { return ((Workload*)w)->refs.sequences.size(); }
ORC_API uint64_t orc_workload_ref_len(void* w, uint64_t i) { return ((Workload*)w)->refs.sequences[i].seq.size(); }
ORC_API const uint8_t* orc_workload_ref_seq(void* w, uint64_t i) { return ((Workload*)w)->refs.sequences[i].seq.data(); }
ORC_API int orc_workload_ref_circular(void* w, uint64_t i) { return ((Workload*)w)->refs.sequences[i].circular; }
ORC_API const char* orc_workload_ref_id(void* w, uint64_t i) { return ((Workload*)w)->refs.sequences[i].id.c_str(); }
ORC_API double orc_workload_ref_weight(void* w, uint64_t i) { return ((Workload*)w)->refs.prob[i]; }
ORC_API void orc_workload_choose_reference(void* w_, void* r, uint64_t* idx, int* strand) {
  auto p = ((Workload*)w_)->refs.choose_reference(*(StdRng*)r);
  *idx = p.first; *strand = p.second;
}

// Run the feeder (Fragments iterator) for one batch of `target` bases, at most max_items items
// (0 = until exhausted).  Appends to / replaces the item list.
ORC_API uint64_t orc_workload_feed(void* w_, uint64_t target, uint64_t max_items) {
  Workload* w = (Workload*)w_;
  w->items.clear();
  w->outs.clear();
  try {
    Fragments fr(target, w->junk, w->random, w->chimera, w->refs, *w->length, *w->identity, w->main_rng);
    WorkItem it;
    while ((max_items == 0 || w->items.size() < max_items) && fr.next(it)) w->items.push_back(it);
  } catch (const std::exception& e) {  // a fragment no contig can host: the reference loops forever there
    g_err = e.what();
    w->items.clear();
    return ~0ULL;
  }
  return w->items.size();
}
// one descriptor = 12 u64: type, ref_idx, strand('+'=0,'-'=1,'*'=2), start, end, has_chimera, ctype, cref_idx, cstrand, cstart, cend, seed ; plus identity (f64) and length
ORC_API void orc_workload_items(void* w_, uint64_t* desc12, double* identity, uint64_t* length) {
  Workload* w = (Workload*)w_;
  auto sc = [](char c) -> uint64_t { return c == '+' ? 0 : (c == '-' ? 1 : 2); };
  for (size_t i = 0; i < w->items.size(); ++i) {
    const WorkItem& it = w->items[i];
    uint64_t* d = desc12 + 12 * i;
    d[0] = it.desc.origin.read_type; d[1] = it.ref_idx; d[2] = sc(it.desc.origin.strand); d[3] = it.desc.origin.start; d[4] = it.desc.origin.end;
    d[5] = it.desc.has_chimera; d[6] = it.desc.chimera.read_type; d[7] = it.ref_idx2; d[8] = sc(it.desc.chimera.strand);
    d[9] = it.desc.chimera.start; d[10] = it.desc.chimera.end; d[11] = it.seed;
    identity[i] = it.desc.identity;
    length[i] = it.desc.length;
  }
}
// inverse of orc_workload_items: replace the fed items by caller-supplied descriptors (same 12-u64 layout), so that a
// test can run generate_read of the oracle on work items produced elsewhere (e.g. by the product's own feeder)
ORC_API void orc_workload_load_items(void* w_, const uint64_t* desc12, const double* identity, uint64_t n) {
  Workload* w = (Workload*)w_;
  w->items.assign(n, WorkItem());
  w->outs.clear();
  auto cs = [](uint64_t s) -> char { return s == 0 ? '+' : (s == 1 ? '-' : '*'); };
  for (uint64_t i = 0; i < n; ++i) {
    const uint64_t* d = desc12 + 12 * i;
    WorkItem& it = w->items[i];
    it.desc.origin.read_type = (ReadType)d[0]; it.ref_idx = d[1]; it.desc.origin.strand = cs(d[2]); it.desc.origin.start = d[3]; it.desc.origin.end = d[4];
    if (d[0] == REAL) it.desc.origin.ref_id = w->refs.sequences[d[1]].id;
    it.desc.has_chimera = d[5] != 0;
    it.desc.chimera.read_type = (ReadType)d[6]; it.ref_idx2 = d[7]; it.desc.chimera.strand = cs(d[8]); it.desc.chimera.start = d[9]; it.desc.chimera.end = d[10];
    if (it.desc.has_chimera && d[6] == REAL) it.desc.chimera.ref_id = w->refs.sequences[d[7]].id;
    it.seed = d[11];
    it.desc.identity = identity[i];
  }
}
// test hook: overwrite the target identity of fed item i (edge cases such as 0.0 / 1.0 that Identity never draws)
ORC_API void orc_workload_set_item_identity(void* w_, uint64_t i, double identity) { ((Workload*)w_)->items[i].desc.identity = identity; }
// generate_read for every fed item on `threads` CPU threads; returns seconds spent, <0 on failure
ORC_API double orc_workload_generate(void* w_, int threads, int keep_cigar_raw) {
  Workload* w = (Workload*)w_;
  w->outs.assign(w->items.size(), Workload::Out());
  Models models{*w->adapter, *w->error, *w->glitch, *w->qscore};
  std::atomic<size_t> next{0};
  std::atomic<int> failed{0};
  std::mutex emu;
  auto t0 = std::chrono::steady_clock::now();
  auto worker = [&]() {
    for (;;) {
      size_t i = next.fetch_add(1);
      if (i >= w->items.size()) break;
      Workload::Out& o = w->outs[i];
      o.d = w->items[i].desc;
      try {
        o.rc = generate_read(w->refs.sequences[w->items[i].ref_idx], w->refs.sequences[w->items[i].ref_idx2], o.d, models,
                             StdRng::seed_from_u64(w->items[i].seed), o.seq, o.qual, keep_cigar_raw ? &o.cigar : nullptr,
                             keep_cigar_raw ? &o.raw : nullptr);
      } catch (const std::exception& e) {
        std::lock_guard<std::mutex> g(emu);
        w->text = e.what();
        o.rc = -9;
      }
      if (o.rc < 0) failed = 1;
    }
  };
  int nt = std::max(1, threads);
  if (nt == 1) worker();
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back(worker);
    for (auto& t : th) t.join();
  }
  double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (failed) { g_err = w->text.empty() ? "get_qscore failed" : w->text; return -1.0; }
  return s;
}
// which: 0 seq, 1 qual, 2 cigar, 3 raw
static const Bytes& out_field(const Workload::Out& o, int which) {
  return which == 0 ? o.seq : which == 1 ? o.qual : which == 2 ? o.cigar : o.raw;
}
ORC_API uint64_t orc_workload_out_len(void* w, uint64_t i, int which) { return out_field(((Workload*)w)->outs[i], which).size(); }
ORC_API void orc_workload_out_lens(void* w_, int which, uint64_t* lens) {
  Workload* w = (Workload*)w_;
  for (size_t i = 0; i < w->outs.size(); ++i) lens[i] = out_field(w->outs[i], which).size();
}
ORC_API void orc_workload_out_copy(void* w_, int which, uint8_t* arena, const uint64_t* offsets) {
  Workload* w = (Workload*)w_;
  for (size_t i = 0; i < w->outs.size(); ++i) {
    const Bytes& b = out_field(w->outs[i], which);
    if (!b.empty()) std::memcpy(arena + offsets[i], b.data(), b.size());
  }
}
ORC_API void orc_workload_out_identity(void* w_, double* id) {
  Workload* w = (Workload*)w_;
  for (size_t i = 0; i < w->outs.size(); ++i) id[i] = w->outs[i].d.identity;
}
ORC_API char* orc_workload_out_description(void* w, uint64_t i) { return dup_str(((Workload*)w)->outs[i].d.to_string()); }
ORC_API char* orc_workload_item_description(void* w, uint64_t i) { return dup_str(((Workload*)w)->items[i].desc.to_string()); }

// ---- whole simulate() ------------------------------------------------------------------------
// params: "key=value\n" lines (keys = SimParams field names); reference text passed separately.
// Returns stats via out[5] = reads, bases, fastq_bytes, generate_seconds, total_seconds; the FASTQ text is
// returned (malloc'd) when capture != 0.
ORC_API char* orc_simulate(const char* kv, const char* fasta, uint64_t fasta_len, const char* error_text,
                           const char* qscore_text, int capture, double* out) {
  ORC_BEGIN 
      SimParams p;
      p.reference_text.assign(fasta, fasta_len);
      auto nums = [](const std::string& v) { std::vector<double> r; for (auto& t : split(v, ',')) r.push_back(std::strtod(t.c_str(), nullptr)); return r; };
      for (auto& line : split(kv, '\n')) {
        size_t eq = line.find('=');
        if (eq == std::string::npos) continue;
        std::string k = line.substr(0, eq), v = line.substr(eq + 1);
        if (k == "seed") p.seed = std::strtoull(v.c_str(), nullptr, 10);
        else if (k == "length") { auto n = nums(v); p.length_mean = n.at(0); p.length_stdev = n.at(1); }
        else if (k == "identity") { auto n = nums(v); p.id_mean = n.at(0); p.id_max = n.at(1); p.id_stdev = n.at(2); }
        else if (k == "start_adapter") { auto n = nums(v); p.start_rate = n.at(0); p.start_amount = n.at(1); }
        else if (k == "end_adapter") { auto n = nums(v); p.end_rate = n.at(0); p.end_amount = n.at(1); }
        else if (k == "start_adapter_seq") p.start_adapter_seq = v;
        else if (k == "end_adapter_seq") p.end_adapter_seq = v;
        else if (k == "junk_reads") p.junk = std::strtod(v.c_str(), nullptr);
        else if (k == "random_reads") p.random = std::strtod(v.c_str(), nullptr);
        else if (k == "chimera") p.chimera = std::strtod(v.c_str(), nullptr);
        else if (k == "glitches") { auto n = nums(v); p.glitch_rate = n.at(0); p.glitch_size = n.at(1); p.glitch_skip = n.at(2); }
        else if (k == "small_plasmid_bias") p.small_plasmid_bias = v == "1" || v == "true";
        else if (k == "quantity") p.quantity = v;
        else if (k == "number_base_store") p.nb_base_store = v;
        else if (k == "threads") p.threads = std::atoi(v.c_str());
        else if (k == "output") p.output_path = v;
        else if (k == "error_model") p.error_model = v;     // random | file
        else if (k == "qscore_model") p.qscore_model = v;   // random | ideal | file
        else throw std::runtime_error("orc_simulate: unknown key " + k);
      }
      if (p.error_model != "random") p.error_model = std::string("text:") + (error_text ? error_text : "");
      if (p.qscore_model != "random" && p.qscore_model != "ideal") p.qscore_model = std::string("text:") + (qscore_text ? qscore_text : "");
      std::string cap;
      SimStats st = simulate(p, capture ? &cap : nullptr);
      out[0] = (double)st.reads; out[1] = (double)st.bases; out[2] = (double)st.fastq_bytes;
      out[3] = st.generate_seconds; out[4] = st.total_seconds;
      return dup_str(cap); ORC_END(nullptr)
}
