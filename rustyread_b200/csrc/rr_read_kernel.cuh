// rustyread_b200 — the per-read kernel: assemble (K2) -> glitches + error walk (K3) with inline
// alignment (K4) -> streamed stitch (K5).  See rr_device.cuh for the layout of the work.
#pragma once
#include "rr_device.cuh"

namespace rr {

// What an out-of-line stage hands back: the generator state (kept in registers by the caller) + one value.
struct StageOut {
  uint64_t base, cur, value;
  uint32_t status;
};

__device__ __forceinline__ uint8_t complement(uint8_t c) {
  // bio::alphabets::dna::complement (IUPAC aware, case preserving); used by Reference::new (src/references.rs:27)
  switch (c) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
    case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
    case 'Y': return 'R'; case 'R': return 'Y'; case 'K': return 'M'; case 'M': return 'K';
    case 'D': return 'H'; case 'H': return 'D'; case 'V': return 'B'; case 'B': return 'V';
    case 'y': return 'r'; case 'r': return 'y'; case 'k': return 'm'; case 'm': return 'k';
    case 'd': return 'h'; case 'h': return 'd'; case 'v': return 'b'; case 'b': return 'v';
    default: return c;  // W, S, N and anything unknown map to themselves
  }
}

// Warp copy of a forward-strand reference slice: after a byte-wise head that brings dst to a 16-byte boundary, every
// lane stores one 128-bit word per step, built from three aligned 64-bit loads of the (differently aligned) source with
// funnel shifts; the warp moves 512 contiguous bytes per step.  The source may be read up to 15 bytes past src + n
// (the reference arena is allocated with slack).
__device__ __forceinline__ void warp_copy_vec(uint8_t* dst, const uint8_t* src, uint64_t n, int lane) {
  const uint64_t head = min(n, (uint64_t)((16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u));
  if ((uint64_t)lane < head) dst[lane] = src[lane];
  const uint64_t nvec = (n - head) >> 4;
  const uint8_t* sb = src + head;
  const uint32_t sh = (uint32_t)((uintptr_t)sb & 7u) * 8u;
  const uint64_t* sa = reinterpret_cast<const uint64_t*>(sb - (sh >> 3));
  uint4* da = reinterpret_cast<uint4*>(dst + head);
  for (uint64_t v = lane; v < nvec; v += 32) {
    const uint64_t a = sa[2 * v], b = sa[2 * v + 1];
    uint64_t o0 = a, o1 = b;
    if (sh) {
      const uint64_t c = sa[2 * v + 2];
      o0 = (a >> sh) | (b << (64u - sh));
      o1 = (b >> sh) | (c << (64u - sh));
    }
    da[v] = make_uint4((uint32_t)o0, (uint32_t)(o0 >> 32), (uint32_t)o1, (uint32_t)(o1 >> 32));
  }
  const uint64_t done = head + (nvec << 4);
  if (done + (uint64_t)lane < n) dst[done + lane] = src[done + lane];  // tail: fewer than 16 bytes
}

// Adapter::get_start / get_end (src/model/adapter.rs:64-93): Bernoulli(rate), then a Beta-distributed
// prefix of the adapter.  Returns the number of bytes appended.
__device__ __forceinline__ uint32_t add_adapter(WarpRng& rng, const DeviceModels& M, bool end_side, uint8_t* dst, int lane) {
  double rate = end_side ? M.end_rate : M.start_rate;
  int has = end_side ? M.has_end : M.has_start;
  if (!rng.gen_bool(rate) || !has) return 0;
  uint32_t alen = end_side ? M.end_len : M.start_len;
  const uint8_t* a = M.adapter_bytes + (end_side ? M.start_len : 0);
  double frac = beta_sample(end_side ? M.end_beta : M.start_beta, rng);
  uint32_t n = (uint32_t)round((double)alen * frac);
  if (n > alen) n = alen;
  for (uint32_t i = lane; i < n; i += 32) dst[i] = a[i];
  return n;
}

// add_fragment (src/simulate/mod.rs:295-339).  `tmp` = 8 bytes of shared memory.
__device__ __forceinline__ uint64_t add_fragment(WarpRng& rng, const DeviceModels& M, uint32_t type, uint32_t ref_idx, uint32_t strand,
                                                 uint64_t start, uint64_t end, uint8_t* dst, uint8_t* tmp, int lane) {
  if (type == RR_JUNK) {  // mod.rs:310-317
    uint64_t unit = rng.gen_range_u64(1, 5);
    __syncwarp();
    rng.random_seq(unit, tmp);
    __syncwarp();
    uint64_t total = (end / unit) * unit;
    for (uint64_t i = lane; i < total; i += 32) dst[i] = tmp[i % unit];
    return total;
  }
  if (type == RR_RANDOM) {  // mod.rs:319-324
    rng.random_seq(end, dst);
    return end;
  }
  // mod.rs:326-339 — `revcomp` is indexed with the same coordinates as `seq`
  const uint64_t r0 = M.ref_off[ref_idx];
  const uint64_t rl = M.ref_off[ref_idx + 1] - r0;
  const uint8_t* ref = M.ref_bytes + r0;
  const bool rc = strand == RR_MINUS;
  uint64_t n1, n2 = 0;
  if (start < end) n1 = end - start;
  else if (M.ref_circ[ref_idx]) { n1 = rl - start; n2 = end; }
  else return 0;
  if (!rc) {
    warp_copy_vec(dst, ref + start, n1, lane);
    if (n2) warp_copy_vec(dst + n1, ref, n2, lane);
  } else {
    for (uint64_t i = lane; i < n1; i += 32) dst[i] = complement(ref[rl - 1 - (start + i)]);
    for (uint64_t i = lane; i < n2; i += 32) dst[n1 + i] = complement(ref[rl - 1 - i]);
  }
  return n1 + n2;
}

// ---- K2 (out of line): the raw fragment of generate_read, src/simulate/mod.rs:243-271.  value = raw length ----
__device__ __noinline__ StageOut stage_assemble(WarpRng rng, const ReadKernelParams& P, WarpSmem* ws, uint32_t ridx, uint8_t* raw) {
  const DeviceModels& M = P.M;
  const int lane = rng.lane;
  const rr_desc d = P.descs[ridx];
  const uint32_t k = M.k;
  uint64_t L = 0;
  rng.random_seq(k, raw);
  L += k;
  L += add_adapter(rng, M, false, raw + L, lane);
  L += add_fragment(rng, M, d.type, d.ref_idx, d.strand, d.start, d.end, raw + L, ws->alt, lane);
  if (d.has_chimera) {
    if (rng.gen_bool(0.25)) L += add_adapter(rng, M, true, raw + L, lane);   // CHIMERA_END_ADAPTER_CHANCE, lib.rs:13
    if (rng.gen_bool(0.25)) L += add_adapter(rng, M, false, raw + L, lane);  // CHIMERA_START_ADAPTER_CHANCE, lib.rs:12
    L += add_fragment(rng, M, d.ctype, d.cref_idx, d.cstrand, d.cstart, d.cend, raw + L, ws->alt, lane);
  }
  L += add_adapter(rng, M, true, raw + L, lane);
  rng.random_seq(k, raw + L);
  L += k;
  __syncwarp();
  StageOut o;
  o.base = rng.base;
  o.cur = rng.cursor();
  o.value = L;
  o.status = ST_OK;
  return o;
}

// ---- add_glitches (out of line), src/simulate/error.rs:250-264: positions, skips and inserted bases are drawn up front.
// value = number of glitches recorded in the warp's scratch ----
__device__ __noinline__ StageOut stage_glitches(WarpRng rng, const ReadKernelParams& P, uint32_t gwarp, uint64_t L) {
  const DeviceModels& M = P.M;
  const int lane = rng.lane;
  GlitchRec* grec = P.sc_grec + (uint64_t)gwarp * P.sc_grec_cap;
  uint8_t* gseq = P.sc_gseq + (uint64_t)gwarp * P.sc_gseq_cap;
  uint32_t ng = 0, status = ST_OK;
  uint64_t position = 0, gbytes = 0;
  for (;;) {
    uint64_t begin = geometric_sample(M.g_dist, rng);
    uint64_t end = begin + geometric_sample(M.g_skip, rng);
    uint64_t size = geometric_sample(M.g_size, rng);
    if (gbytes + size > P.sc_gseq_cap || ng >= P.sc_grec_cap) { status = ST_GLITCH_OVERFLOW; break; }
    __syncwarp();
    rng.random_seq(size, gseq + gbytes);
    position += begin;
    if (position > L || position + end > L) break;
    if (lane == 0) {
      GlitchRec g;
      g.begin = position;
      g.end = position + (end - begin);
      g.seq_off = (uint32_t)gbytes;
      g.seq_len = (uint32_t)size;
      grec[ng] = g;
    }
    ++ng;
    gbytes += size;
  }
  __syncwarp();
  StageOut o;
  o.base = rng.base;
  o.cur = rng.cursor();
  o.value = ng;
  o.status = status;
  return o;
}

// Error::add_errors_to_kmer on a k-mer that is in the table (src/model/error.rs:96-100): one WeightedIndex draw.
// Out of line: only file models reach it.  value = alternative length (bytes in ws->alt).
__device__ __noinline__ StageOut stage_alt_from_table(WarpRng rng, const DeviceModels& M, WarpSmem* ws, int key) {
  const int lane = rng.lane;
  uint32_t a0 = M.em_alt_begin[key], a1 = M.em_alt_begin[key + 1];
  uint32_t na = a1 - a0;
  double chosen = u64_to_unit12(rng.next_u64()) * M.em_scale[key] + 0.0;
  uint32_t idx = 0;  // number of partial sums <= chosen among the first na - 1
  for (uint32_t b = 0; b + 1 < na; b += 32) {
    uint32_t i = b + lane;
    bool le = (i + 1 < na) && (M.em_cum[a0 + i] <= chosen);
    idx += __popc(__ballot_sync(0xffffffffu, le));
  }
  uint32_t o0 = M.em_alt_off[a0 + idx], o1 = M.em_alt_off[a0 + idx + 1];
  uint32_t n = o1 - o0;
  __syncwarp();
  for (uint32_t i = lane; i < n; i += 32) ws->alt[i] = M.em_alt_bytes[o0 + i];
  __syncwarp();
  StageOut o;
  o.base = rng.base;
  o.cur = rng.cursor();
  o.value = n;
  o.status = a0 + idx;  // global id of the chosen alternative (indexes the precomputed alignments)
  return o;
}

// State of `changes.last()` plus the streamed stitch of sequence() (src/simulate/error.rs:219-246).
// A Change only ever grows by appending to seq / cigar (merge, error.rs:105-152), and only the last
// one can change, so err / CIGAR are written once, in order, as the walk proceeds.
// All positions are 32-bit: the host refuses reads whose raw capacity reaches 2^31 bases.
struct Stitch {
  const uint8_t* raw;
  uint8_t* err;
  uint8_t* cig;
  uint32_t L;
  uint32_t cap;
  uint32_t err_len, cig_len;
  uint32_t pos_in_raw;   // sequence(): end of the last emitted change
  uint32_t real_edit;
  bool has_last, l_emit;
  uint32_t lb, le, lseq_len, ledit;
  uint32_t status;
};

// Capacity is checked once per event by the caller (add_change) or by st_tail.
__device__ __forceinline__ void st_emit_gap(Stitch& S, uint32_t n, int lane) {  // raw[pos_in_raw .. +n) copied, '=' in the CIGAR
  const uint8_t* src = S.raw + S.pos_in_raw;
  uint8_t* de = S.err + S.err_len;
  uint8_t* dc = S.cig + S.cig_len;
  if ((uint32_t)lane < n) {
    de[lane] = src[lane];
    dc[lane] = '=';
  }
  for (uint32_t i = lane + 32; i < n; i += 32) {
    de[i] = src[i];
    dc[i] = '=';
  }
  S.err_len += n;
  S.cig_len += n;
}

__device__ __forceinline__ void st_emit_seq(Stitch& S, const uint8_t* q, uint32_t n, int lane) {
  uint8_t* de = S.err + S.err_len;
  if ((uint32_t)lane < n) de[lane] = q[lane];
  for (uint32_t i = lane + 32; i < n; i += 32) de[i] = q[i];
  S.err_len += n;
}

// AbsChanges::add_change (src/simulate/error.rs:173-193) with contain / overlap / merge /
// update_align (error.rs:83-164).  Returns the edit count to subtract from the target.
// One alignment call site: the merge and the push cases only differ in which slices are aligned.
// `fk` != 0 says that oseq is raw[ob..oe) with ONE random edit of kind fk (1 substitution, 2 insertion, 3 deletion) at
// offset fep (random_error, src/model/error.rs:116-135).  When such a change is pushed, or merged into a last change
// whose raw and err ends coincide, the piece that gets aligned is a k-mer tail against the same tail with at most one
// edit, and the CIGAR of the reference's LCS alignment has a closed form, so no DP is run: the suffix after the edit
// matches on the diagonal, the edit then slides left through equal neighbours (a match always takes the diagonal),
// and where it stops the prefixes are identical, which forces X / I / D there.  Checked against the DP + tie-break
// exhaustively for all k-mers up to k = 7 over 1-, 2- and 4-letter alphabets and for every piece offset.
// fk == 4: oseq is alternative `fep` of the error-model table for the k-mer raw[ob..oe); its alignments against the key
// (whole, and every tail piece) were computed at load time with the same rule, so they are copied instead of recomputed.
__device__ __forceinline__ uint32_t add_change(Stitch& S, WarpSmem* ws, const ReadKernelParams& P, uint32_t gwarp, uint32_t ob, uint32_t oe,
                                               const uint8_t* oseq, uint32_t olen, uint32_t fk, uint32_t fep, int lane) {
  bool merge = false, piece_fast = fk != 0;  // closed form applies to a pushed change, and to some merges (below)
  uint32_t es = 0, rs, re;
  if (S.has_last) {
    const uint32_t l_end_err = S.lb + S.lseq_len;
    if (oe <= l_end_err || oe <= S.le || ob + olen <= l_end_err) return 0;  // contain (error.rs:83-88)
    if (ob < l_end_err || ob < S.le) {                                       // overlap -> merge (error.rs:105-152)
      merge = true;
      const uint32_t old_end = S.le;
      if (ob < l_end_err) {
        es = l_end_err - ob;  // olen > es holds because the change is not contained
        S.le = S.le + oe - ob - es;
        // If the last change consumed as many raw bases as it produced (old_end == l_end_err), the appended piece
        // oseq[es..] faces exactly raw[ob + es .. oe): the same k-mer tail, so the single-edit closed form still holds.
        piece_fast = piece_fast && old_end == l_end_err && es <= oe - ob;  // (a longer overlap makes the raw slice run backwards)
      } else {
        S.le = ob + olen;
        piece_fast = false;
      }
      if (S.le < S.L - 1) {
        if (old_end < S.le) { rs = old_end; re = S.le; } else { rs = S.le; re = old_end; }
      } else { rs = old_end; re = S.L; piece_fast = false; }
      if (rs > re || re > S.L || es > olen) { S.status = ST_REF_PANIC; return 0; }
      S.lseq_len += olen - es;
    }
  }
  uint32_t gap = 0;
  if (!merge) {  // push a new change (update_align: error.rs:155-164)
    if (S.has_last && S.l_emit) {  // the previous change is final: sequence() accounts for it
      S.real_edit += S.ledit;
      S.pos_in_raw = S.le;
    }
    rs = ob;
    re = oe > S.L ? S.L : oe;
    if (rs > re) { S.status = ST_REF_PANIC; return 0; }
    S.has_last = true;
    S.l_emit = ob >= S.pos_in_raw;  // sequence(): `if change.begin() < pos_in_raw { continue }`
    S.lb = ob;
    S.le = oe;
    S.lseq_len = olen;
    S.ledit = 0;
    gap = ob - S.pos_in_raw;  // only meaningful when l_emit
  }
  const uint32_t add = olen - es, tn = re - rs;
  uint8_t* dst = nullptr;
  if (S.l_emit) {
    // one capacity check per event: the CIGAR of this piece has at most add + tn ops
    if (S.err_len + gap + add > S.cap || S.cig_len + gap + add + tn > S.cap) { S.status = ST_OUT_OVERFLOW; return 0; }
    if (gap) st_emit_gap(S, gap, lane);
    st_emit_seq(S, oseq + es, add, lane);
    dst = S.cig + S.cig_len;
  }
  if (piece_fast && fk == 4) {
    const DeviceModels& M = P.M;
    const uint32_t pi = M.em_piece_begin[fep] + es;
    const uint32_t o0 = M.em_piece_off[pi], total = M.em_piece_off[pi + 1] - o0;
    const uint32_t ed = M.em_piece_edits[pi];
    if (dst) for (uint32_t x = lane; x < total; x += 32) dst[x] = M.em_piece_ops[o0 + x];
    if (S.l_emit) S.cig_len += total;
    S.ledit += ed;
    return ed;
  }
  if (piece_fast) {
    const uint8_t* q = oseq + es;
    const uint8_t* t = S.raw + rs;
    const uint32_t e = fep >= es ? fep - es : 0;    // where the edit sits inside the piece (an edit left of it acts at 0)
    const bool has_edit = !(fk == 1 && fep < es);   // a substitution in the overlapped part leaves the piece identical
    uint32_t at = e;                                // where the non-'=' op lands
    if (fk != 1) {
      // slide left while the base left of the edit equals the one it would be swapped with
      const bool c = (uint32_t)lane >= 1 && (uint32_t)lane <= e && (fk == 2 ? q[lane] == t[lane - 1] : t[lane - 1] == t[lane]);
      const uint32_t B = __ballot_sync(0xffffffffu, c);
      at = e - __clz(~(B << (31 - e)));
    }
    const uint32_t total = fk == 2 ? add : tn;  // substitution / deletion: one op per raw base, insertion: one per piece base
    if (dst && (uint32_t)lane < total) dst[lane] = (has_edit && (uint32_t)lane == at) ? (fk == 1 ? 'X' : fk == 2 ? 'I' : 'D') : '=';
    if (S.l_emit) S.cig_len += total;
    const uint32_t ed = has_edit ? 1u : 0u;
    S.ledit += ed;
    return ed;
  }
  const uint64_t a = warp_align(ws, P, gwarp, oseq + es, add, S.raw + rs, tn, dst, lane);
  if (align_status(a)) { S.status = align_status(a); return 0; }
  if (S.l_emit) S.cig_len += align_nops(a);
  S.ledit += align_edits(a);
  return align_edits(a);
}

__device__ __forceinline__ void process_read(const ReadKernelParams& P, WarpSmem* ws, uint32_t gwarp, uint32_t ridx, int lane) {
  const DeviceModels& M = P.M;
  const uint32_t k = M.k;
  WarpRng rng;
  rng.seed(P.descs[ridx].seed, ws->rng, lane);

  uint8_t* raw = P.raw + P.raw_off[ridx];
  StageOut so = stage_assemble(rng, P, ws, ridx, raw);
  rng.set_cursor(so.base, so.cur);
  const uint32_t L = (uint32_t)so.value;

  Stitch S;
  S.raw = raw;
  S.L = L;
  S.err = P.err + P.out_off[ridx];
  S.cig = P.cig + P.out_off[ridx];
  S.cap = (uint32_t)(P.out_off[ridx + 1] - P.out_off[ridx]);
  S.err_len = S.cig_len = 0;
  S.pos_in_raw = 0;
  S.real_edit = 0;
  S.has_last = false;
  S.l_emit = false;
  S.lb = S.le = S.lseq_len = S.ledit = 0;
  S.status = ST_OK;

  // ---- K3: error::sequence (src/simulate/error.rs:201-217) ----
  double target = (1.0 - P.descs[ridx].identity) * (double)L;  // number_of_edit, error.rs:196-198
  const GlitchRec* grec = P.sc_grec + (uint64_t)gwarp * P.sc_grec_cap;
  const uint8_t* gseq = P.sc_gseq + (uint64_t)gwarp * P.sc_gseq_cap;
  uint32_t ng = 0;
  if (M.has_distance) {
    so = stage_glitches(rng, P, gwarp, L);
    rng.set_cursor(so.base, so.cur);
    ng = (uint32_t)so.value;
    S.status = so.status;
  }

  // add_error (error.rs:267-302): one Bernoulli(target / (L - pos)) per position.  32 positions are
  // tried at once against the current target; the first lane with an event wins.
  if (S.status == ST_OK) {
    const uint32_t nlim = L - k;
    uint32_t pos = 0;
    uint32_t g = 0;
    uint32_t gbegin = ng ? (uint32_t)grec[0].begin : ~0u;
    bool glitch_checked = false;
    while (pos < nlim) {
      rng.need(64);
      const uint32_t pi = pos + (uint32_t)lane;
      const bool inb = pi < nlim;
      const double rem = (double)(L - pi);
      const bool brk = target < 0.0 || target > rem;
      const bool gl = (pi == gbegin) && !(lane == 0 && glitch_checked);
      const double p = target / rem;
      const bool one = (p == 1.0);
      const uint64_t v = rng.peek64_at(rng.off + 2 * (uint32_t)lane);
      const bool hit = one || v < bernoulli_threshold(p);
      const uint32_t m_any = __ballot_sync(0xffffffffu, !inb || gl || brk || hit);
      if (m_any == 0) {
        pos += 32;
        rng.off += 64;
        glitch_checked = false;
        continue;
      }
      const int f = __ffs(m_any) - 1;
      const uint32_t fl = (inb ? 1u : 0u) | (gl ? 2u : 0u) | (brk ? 4u : 0u) | (one ? 8u : 0u);
      const uint32_t ff = __shfl_sync(0xffffffffu, fl, f);
      if (f > 0) {
        pos += (uint32_t)f;
        rng.off += 2 * (uint32_t)f;
        glitch_checked = false;
      }
      if (!(ff & 1u)) break;  // pos == L - k
      uint32_t ob, oe, olen, fk = 0, fep = 0;
      const uint8_t* oseq;
      const bool is_glitch = (ff & 2u) != 0;
      if (is_glitch) {  // glitch due at pos (error.rs:282-287)
        const GlitchRec gr = grec[g];
        ob = (uint32_t)gr.begin;
        oe = (uint32_t)gr.end;
        oseq = gseq + gr.seq_off;
        olen = gr.seq_len;
        ++g;
        gbegin = (g < ng) ? (uint32_t)grec[g].begin : ~0u;
      } else {
        if (ff & 4u) break;  // error.rs:289-291
        if (!(ff & 8u)) rng.off += 2;
        // Error::add_errors_to_kmer (src/model/error.rs:93-107) -> ws->alt
        const uint8_t* kmer = raw + pos;
        int key = -1;
        __syncwarp();
        if (M.em_index) {  // 2-bit code of the k-mer; any byte outside ACGT misses the table
          uint32_t part = 0, bad = 0;
          if ((uint32_t)lane < k) {
            uint8_t c = kmer[lane];
            uint32_t s = c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 4u;
            bad = s > 3u;
            part = (s & 3u) << (2 * (k - 1 - lane));
          }
          uint32_t code = __reduce_or_sync(0xffffffffu, part);
          if (!__any_sync(0xffffffffu, bad)) key = M.em_index[code];
        }
        if (key >= 0) {
          so = stage_alt_from_table(rng, M, ws, key);
          rng.set_cursor(so.base, so.cur);
          olen = (uint32_t)so.value;
          fk = 4;
          fep = so.status;
        } else {  // random_error (src/model/error.rs:116-135)
          const uint32_t ep = (uint32_t)rng.gen_range_u64(0, k - 1);
          const uint32_t kind = rng.gen_range_u32(1, 3);  // 1 substitution, 2 insertion, 3 deletion
          uint8_t nb = 0;
          if (kind == 1) nb = rng.random_base_diff(kmer[ep]);
          else if (kind == 2) nb = rng.random_base();
          const uint32_t n = kind == 1 ? k : (kind == 2 ? k + 1 : k - 1);
          if ((uint32_t)lane < n) {
            const uint32_t i = lane;
            uint8_t c;
            if (kind == 1) c = (i == ep) ? nb : kmer[i];
            else if (kind == 2) c = (i < ep) ? kmer[i] : (i == ep ? nb : kmer[i - 1]);
            else c = (i < ep) ? kmer[i] : kmer[i + 1];
            ws->alt[i] = c;
          }
          __syncwarp();
          olen = n;
          fk = kind;
          fep = ep;
        }
        ob = pos;
        oe = pos + k;
        oseq = ws->alt;
      }
      target -= (double)add_change(S, ws, P, gwarp, ob, oe, oseq, olen, fk, fep, lane);
      if (S.status) break;
      if (is_glitch) {
        glitch_checked = true;  // same pos again, with the new target, glitch check done
      } else {
        if (target < 0.0 || target > (double)(L - pos)) break;  // error.rs:298-300
        pos += 1;
        glitch_checked = false;
      }
    }
  }

  // ---- K5: tail of the stitch (error.rs:241-246) ----
  if (S.status == ST_OK) {
    if (S.has_last && S.l_emit) {
      S.real_edit += S.ledit;
      S.pos_in_raw = S.le;
    }
    if (S.pos_in_raw < L) {
      const uint32_t n = L - S.pos_in_raw;
      if (S.err_len + n > S.cap || S.cig_len + n > S.cap) S.status = ST_OUT_OVERFLOW;
      else st_emit_gap(S, n, lane);
    }
  }
  __syncwarp();
  if (lane == 0) {
    P.raw_len[ridx] = L;
    P.err_len[ridx] = S.status ? 0 : S.err_len;
    P.cig_len[ridx] = S.status ? 0 : S.cig_len;
    P.q_cursor[ridx] = rng.cursor();
    P.real_id[ridx] = (1.0 - ((double)S.real_edit / (double)L)) * 100.0;  // error.rs:246, mod.rs:289
    P.status[ridx] = S.status;
  }
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, RR_BLOCKS_PER_SM) rr_read_kernel(const __grid_constant__ ReadKernelParams P) {
  __shared__ WarpSmem smem[WARPS_PER_BLOCK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t gwarp = blockIdx.x * WARPS_PER_BLOCK + warp;
  WarpSmem* ws = &smem[warp];
  for (;;) {
    uint32_t slot = 0;
    if (lane == 0) slot = atomicAdd(P.queue, 1u);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot >= P.n_reads) break;
    process_read(P, ws, gwarp, P.order[slot], lane);
    __syncwarp();
  }
}

}  // namespace rr
