// rustyread_b200 — the per-read kernel: assemble (K2) -> glitches + error walk (K3) with inline
// alignment (K4) -> streamed stitch (K5).  See rr_device.cuh for the layout of the work.
#pragma once
#include "rr_device.cuh"

namespace rr {

enum : int { EM_RANDOM = 0, EM_PACKED = 1, EM_TABLES = 2 };  // Error::random(k) / file model in packed records / file model in plain tables

// What an out-of-line stage hands back: the generator's offset (its base lives in shared memory) + one value.
struct StageOut {
  uint64_t value;
  uint32_t off, status;
};
__device__ __forceinline__ WarpRng stage_rng(uint32_t ws32, uint32_t off, int lane) {
  WarpRng rng;
  rng.ws = smem_ptr(ws32);
  rng.off = off;
  rng.lane = lane;
  return rng;
}

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

// Warp copy of a reverse-strand slice: dst[i] = complement(ref[top - i]) for i < n (`revcomp` indexed with forward
// coordinates, src/simulate/mod.rs:327-338).  Four output bytes per lane and step: two aligned 32-bit loads + a funnel shift
// fetch the four source bytes, one byte permute reverses them, and the complement of A / C / G / T is bit arithmetic
// ('A' 0x41 ^ 'T' 0x54 = 0x15, 'C' 0x43 ^ 'G' 0x47 = 0x04; bit 1 tells the pairs apart); a word that holds any other byte
// (N, IUPAC codes, lower case) goes through complement() byte by byte.  Reads up to 3 bytes around the slice (the
// reference arena is allocated with slack; its base is aligned).
__device__ __forceinline__ void warp_copy_revcomp(uint8_t* dst, const uint8_t* ref, uint64_t top, uint64_t n, int lane) {
  const uint64_t head = min(n, (uint64_t)((4u - (uint32_t)((uintptr_t)dst & 3u)) & 3u));
  if ((uint64_t)lane < head) dst[lane] = complement(ref[top - lane]);
  const uint64_t nw = (n - head) >> 2;
  uint32_t* dw = reinterpret_cast<uint32_t*>(dst + head);
  for (uint64_t j = lane; j < nw; j += 32) {
    const uint8_t* lo = ref + (top - head - 4 * j - 3);  // lowest address of the four source bytes
    const uint32_t sh = (uint32_t)((uintptr_t)lo & 3u) * 8u;
    const uint32_t* la = reinterpret_cast<const uint32_t*>(lo - (sh >> 3));
    const uint32_t v = __funnelshift_r(la[0], sh ? la[1] : 0u, sh);
    uint32_t w = __byte_perm(v, 0, 0x0123);  // byte i of w = source byte top - head - 4j - i
    const uint32_t y = w ^ 0x41414141u, b0 = y & 0x01010101u;  // the test of raw_is_acgt
    if ((y & 0xE8E8E8E8u) | (b0 ^ ((y >> 4) & 0x01010101u)) | (b0 ^ ((y >> 2) & ~(y >> 1) & 0x01010101u))) {
      uint32_t r = 0;
#pragma unroll 1
      for (int x = 0; x < 4; ++x) r |= (uint32_t)complement((uint8_t)(w >> (8 * x))) << (8 * x);
      w = r;
    } else {
      w ^= 0x15151515u ^ (((w >> 1) & 0x01010101u) * 0x11u);
    }
    dw[j] = w;
  }
  const uint64_t done = head + (nw << 2);
  if (done + (uint64_t)lane < n) dst[done + lane] = complement(ref[top - done - lane]);  // fewer than 4 bytes
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
    warp_copy_revcomp(dst, ref, rl - 1 - start, n1, lane);
    if (n2) warp_copy_revcomp(dst + n1, ref, rl - 1, n2, lane);
  }
  return n1 + n2;
}

// True when every byte of raw[0, L) is one of A, C, G, T (raw is 16-byte aligned).  Four bytes at a time: with
// y = x ^ 0x41414141 the valid bytes are 0x00, 0x02, 0x06, 0x15, i.e. bits 3, 5, 6, 7 clear, bit 4 == bit 0, and
// (bit 2 and not bit 1) == bit 0.
__device__ __forceinline__ bool raw_is_acgt(const uint8_t* raw, uint32_t L, int lane) {
  uint32_t bad = 0;
  const uint4* v = reinterpret_cast<const uint4*>(raw);
  for (uint32_t i = lane; i < L / 16; i += 32) {
    const uint4 q = v[i];
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t y = w[j] ^ 0x41414141u;
      const uint32_t b0 = y & 0x01010101u;
      bad |= (y & 0xE8E8E8E8u) | (b0 ^ ((y >> 4) & 0x01010101u)) | (b0 ^ ((y >> 2) & ~(y >> 1) & 0x01010101u));
    }
  }
  const uint32_t t = (L & ~15u) + (uint32_t)lane;  // fewer than 16 bytes left
  if (t < L) {
    const uint8_t c = raw[t];
    bad |= !(c == 'A' || c == 'C' || c == 'G' || c == 'T');
  }
  return !__any_sync(0xffffffffu, bad != 0);
}

// ---- K2 (out of line): the raw fragment of generate_read, src/simulate/mod.rs:243-271.  value = raw length ----
__device__ __noinline__ StageOut stage_assemble(uint32_t ws32, uint32_t off, int lane, const ReadKernelParams& P, uint32_t ridx, uint8_t* raw) {
  const DeviceModels& M = P.M;
  WarpRng rng = stage_rng(ws32, off, lane);
  WarpSmem* ws = rng.ws;
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
  o.off = rng.off;
  o.value = L;
  o.status = raw_is_acgt(raw, (uint32_t)L, lane) ? 0u : 1u;  // 1: some byte outside ACGT (alignments must check their windows)
  return o;
}

// ---- add_glitches (out of line), src/simulate/error.rs:250-264: positions, skips and inserted bases are drawn up front.
// value = number of glitches recorded in the warp's scratch ----
__device__ __noinline__ StageOut stage_glitches(uint32_t ws32, uint32_t off, int lane, const ReadKernelParams& P, uint32_t gwarp, uint64_t L) {
  const DeviceModels& M = P.M;
  WarpRng rng = stage_rng(ws32, off, lane);
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
  o.off = rng.off;
  o.value = ng;
  o.status = status;
  return o;
}

// State of `changes.last()` plus the streamed stitch of sequence() (src/simulate/error.rs:219-246).
// A Change only ever grows by appending to seq / cigar (merge, error.rs:105-152), and only the last
// one can change, so err / CIGAR are written once, in order, as the walk proceeds: into the warp's staging rings
// (WarpSmem::errb / cigb), from which complete 256-byte chunks leave for HBM.
// All positions are 32-bit: the host refuses reads whose raw capacity reaches 2^30 bases.
// The last change is (le = its end in raw, lee = its end in err coordinates = begin + seq.len(), ledit).
struct Stitch {
  uint32_t err_len, cig_len;
  uint32_t pos_in_raw;   // sequence(): end of the last emitted change
  uint32_t real_edit;
  uint32_t le, lee, ledit;
  uint32_t flags;        // F_HAS_LAST | F_EMIT | F_GLITCH_CHECKED | F_REPLAY | F_REPLAYED | F_DIRTY (bytes outside ACGT around) | status << 8
};
enum : uint32_t { F_HAS_LAST = 1u, F_EMIT = 2u, F_GLITCH_CHECKED = 4u, F_REPLAY = 8u, F_REPLAYED = 16u, F_DIRTY = 32u, F_STATUS_SHIFT = 8u };

// Everything the error walk carries from one position to the next.  The walk keeps it in registers; around EVERY
// out-of-line call inside the loop it is parked in WarpSmem::st and read back, so that no value is live across a call:
// the calling convention would otherwise push the whole state to the local-memory stack, and ptxas then reloads it at
// every use in the hot path (first ncu captures of this kernel: 35 LDL/STL per event).
struct Walk {
  uint32_t pos, off;       // position in raw; next word of the read's stream (relative to WarpSmem::rng_base)
  uint32_t L, gbegin;      // raw length; raw position of the next glitch (~0: none)
  uint32_t ridx;
  double target;
  const uint8_t* raw;
  Stitch S;
};
enum : int { ST_L = 0, ST_CAP = 1, ST_ERR_LEN = 2, ST_CIG_LEN = 3, ST_POS_IN_RAW = 4, ST_REAL_EDIT = 5, ST_LE = 6, ST_LEE = 7, ST_LEDIT = 8,
             ST_FLAGS = 9, ST_POS = 10, ST_OFF = 11, ST_G = 12, ST_NG = 13, ST_TARGET = 14 /* 2 words */, ST_GBEGIN = 16, ST_RIDX = 17,
             ST_RAW = 18 /* 2 words */ };

__device__ __forceinline__ void walk_save(WarpSmem* ws, const Walk& W) {  // every lane writes the same values
  __syncwarp();
  uint32_t* st = ws->st;
  st[ST_L] = W.L; st[ST_ERR_LEN] = W.S.err_len; st[ST_CIG_LEN] = W.S.cig_len; st[ST_POS_IN_RAW] = W.S.pos_in_raw; st[ST_REAL_EDIT] = W.S.real_edit;
  st[ST_LE] = W.S.le; st[ST_LEE] = W.S.lee; st[ST_LEDIT] = W.S.ledit; st[ST_FLAGS] = W.S.flags; st[ST_POS] = W.pos; st[ST_OFF] = W.off;
  st[ST_GBEGIN] = W.gbegin; st[ST_RIDX] = W.ridx;
  *reinterpret_cast<double*>(&st[ST_TARGET]) = W.target;
  *reinterpret_cast<const uint8_t**>(&st[ST_RAW]) = W.raw;
  __syncwarp();
}
__device__ __forceinline__ void walk_load(const WarpSmem* ws, Walk& W) {
  __syncwarp();
  const uint32_t* st = ws->st;
  W.L = st[ST_L]; W.S.err_len = st[ST_ERR_LEN]; W.S.cig_len = st[ST_CIG_LEN]; W.S.pos_in_raw = st[ST_POS_IN_RAW]; W.S.real_edit = st[ST_REAL_EDIT];
  W.S.le = st[ST_LE]; W.S.lee = st[ST_LEE]; W.S.ledit = st[ST_LEDIT]; W.S.flags = st[ST_FLAGS]; W.pos = st[ST_POS]; W.off = st[ST_OFF];
  W.gbegin = st[ST_GBEGIN]; W.ridx = st[ST_RIDX];
  W.target = *reinterpret_cast<const double*>(&st[ST_TARGET]);
  W.raw = *reinterpret_cast<const uint8_t* const*>(&st[ST_RAW]);
  __syncwarp();
}

__device__ __forceinline__ uint32_t st_status(const Stitch& S) { return S.flags >> F_STATUS_SHIFT; }
__device__ __forceinline__ void st_fail(Stitch& S, uint32_t status) {
  if (status && !st_status(S)) S.flags |= status << F_STATUS_SHIFT;
}

__device__ __forceinline__ uint4 eq16() { return make_uint4(0x3d3d3d3du, 0x3d3d3d3du, 0x3d3d3d3du, 0x3d3d3d3du); }

// Sends the chunks completed since (e_old, c_old) to HBM and re-arms the CIGAR ring with '='.  Both rings are indexed
// by absolute position modulo RING, so a chunk is CHUNK-aligned on both sides.  The capacity test lives here: a
// read that comes within two chunks of its planned capacity is reported (the host re-plans the batch with more slack).
__device__ __forceinline__ uint32_t flush_chunks(WarpSmem* ws, const ReadKernelParams& P, uint32_t ridx, uint32_t e_old, uint32_t e_new,
                                                 uint32_t c_old, uint32_t c_new, int lane) {
  __syncwarp();
  const uint64_t off = P.out_off[ridx];
  if (lane < (int)(CHUNK / 16u)) {
    uint4* eg = reinterpret_cast<uint4*>(P.err + off);
    uint4* cg = reinterpret_cast<uint4*>(P.cig + off);
    for (uint32_t c = e_old & ~(CHUNK - 1u); c < (e_new & ~(CHUNK - 1u)); c += CHUNK)
      eg[(c >> 4) + lane] = *reinterpret_cast<const uint4*>(&ws->errb[(c & (RING - 1u)) + 16u * lane]);
    for (uint32_t c = c_old & ~(CHUNK - 1u); c < (c_new & ~(CHUNK - 1u)); c += CHUNK) {
      uint4* sp = reinterpret_cast<uint4*>(&ws->cigb[(c & (RING - 1u)) + 16u * lane]);
      cg[(c >> 4) + lane] = *sp;
      *sp = eq16();
    }
  }
  __syncwarp();
  return (max(e_new, c_new) + 2u * CHUNK > ws->st[ST_CAP]) ? (uint32_t)ST_OUT_OVERFLOW : (uint32_t)ST_OK;
}

// Appends n positions to the stitch, at most 64 per step with a flush in between (cold path: glitch inserts, large
// alignments, the tail of the read).  err_src != nullptr: those bytes go to the sequence; cig_src != nullptr:
// those ops go to the CIGAR; cig_eq: the CIGAR advances over its '=' pre-fill instead.
__device__ __forceinline__ void stitch_append(Stitch& S, WarpSmem* ws, const ReadKernelParams& P, uint32_t ridx, const uint8_t* err_src,
                                              const uint8_t* cig_src, bool cig_eq, uint32_t n, int lane) {
  for (uint32_t done = 0; done < n && st_status(S) == ST_OK; done += 64) {
    const uint32_t m = min(64u, n - done);
    const uint32_t e_old = S.err_len, c_old = S.cig_len;
    for (uint32_t i = lane; i < m; i += 32) {
      if (err_src) ws->errb[(e_old + i) & (RING - 1u)] = err_src[done + i];
      if (cig_src) ws->cigb[(c_old + i) & (RING - 1u)] = cig_src[done + i];
    }
    if (err_src) S.err_len += m;
    if (cig_src || cig_eq) S.cig_len += m;
    if (((e_old ^ S.err_len) | (c_old ^ S.cig_len)) >= CHUNK) st_fail(S, flush_chunks(ws, P, ridx, e_old, S.err_len, c_old, S.cig_len, lane));
    __syncwarp();
  }
}

// The bookkeeping half of AbsChanges::add_change (src/simulate/error.rs:173-193) with contain / overlap / merge /
// update_align (error.rs:83-164): decides what happens to a change [ob, oe) carrying olen bytes and updates the state
// of `changes.last()`.  What is left to the caller is to append the piece bytes [es, olen) after `gap` unchanged
// bases raw[gap_src ..), and to align them against raw[rs, re).
struct ChangePlan {
  uint32_t es, rs, re, gap, gap_src;
  bool contained, fast;
};
// `fk` != 0 says that the piece is raw[ob..oe) with ONE random edit of kind fk (1 substitution, 2 insertion, 3 deletion)
// at offset fep (random_error, src/model/error.rs:116-135), or (fk == 4) alternative `fep` of the error-model table for
// the k-mer raw[ob..oe).  `fast` comes back true when the piece that gets aligned is a k-mer tail against the same tail:
// a pushed change, or a merge into a last change whose raw and err ends coincide.  Then the CIGAR of the reference's LCS
// alignment has a closed form (single edit) or was computed at load time (table), and no DP is run.
__device__ __forceinline__ ChangePlan plan_change(Stitch& S, uint32_t L, uint32_t ob, uint32_t oe, uint32_t olen, uint32_t fk) {
  ChangePlan c;
  c.es = 0;
  c.gap = 0;
  c.gap_src = 0;
  c.contained = false;
  c.fast = fk != 0;
  c.rs = c.re = 0;
  bool merge = false;
  if (S.flags & F_HAS_LAST) {
    const uint32_t l_end_err = S.lee;
    if (oe <= l_end_err || oe <= S.le || ob + olen <= l_end_err) {  // contain (error.rs:83-88)
      c.contained = true;
      return c;
    }
    if (ob < l_end_err || ob < S.le) {  // overlap -> merge (error.rs:105-152)
      merge = true;
      const uint32_t old_end = S.le;
      if (ob < l_end_err) {
        c.es = l_end_err - ob;  // olen > es holds because the change is not contained
        S.le = S.le + oe - ob - c.es;
        // If the last change consumed as many raw bases as it produced (old_end == l_end_err), the appended piece
        // oseq[es..] faces exactly raw[ob + es .. oe): the same k-mer tail.  (A longer overlap makes the raw slice run backwards.)
        c.fast = c.fast && old_end == l_end_err && c.es <= oe - ob;
      } else {
        S.le = ob + olen;
        c.fast = false;
      }
      if (S.le < L - 1) {
        if (old_end < S.le) { c.rs = old_end; c.re = S.le; } else { c.rs = S.le; c.re = old_end; }
      } else { c.rs = old_end; c.re = L; c.fast = false; }
      if (c.rs > c.re || c.re > L || c.es > olen) { st_fail(S, ST_REF_PANIC); c.contained = true; return c; }
      S.lee += olen - c.es;
    }
  }
  if (!merge) {  // push a new change (update_align: error.rs:155-164)
    if ((S.flags & (F_HAS_LAST | F_EMIT)) == (F_HAS_LAST | F_EMIT)) {  // the previous change is final: sequence() accounts for it
      S.real_edit += S.ledit;
      S.pos_in_raw = S.le;
    }
    c.rs = ob;
    c.re = oe > L ? L : oe;
    if (c.rs > c.re) { st_fail(S, ST_REF_PANIC); c.contained = true; return c; }
    // sequence(): `if change.begin() < pos_in_raw { continue }`
    S.flags = (S.flags & ~(F_HAS_LAST | F_EMIT)) | F_HAS_LAST | (ob >= S.pos_in_raw ? F_EMIT : 0u);
    S.le = oe;
    S.lee = ob + olen;
    S.ledit = 0;
    c.gap_src = S.pos_in_raw;
    c.gap = ob - S.pos_in_raw;  // only meaningful when F_EMIT
  }
  return c;
}

// ---- out-of-line stages of the walk: walk state in / out through WarpSmem::st ----

// refill of the resident stream words (every ~25 events)
__device__ __noinline__ void walk_refill(uint32_t ws32, int lane) {
  WarpSmem* ws = smem_ptr(ws32);
  const uint32_t off = rng_refill(ws32, ws->st[ST_OFF], lane);
  __syncwarp();
  ws->st[ST_OFF] = off;
  __syncwarp();
}

// Cold form of add_change: the piece lives in memory (glitch inserts, table alternatives longer than a warp).
// Returns the edit count.
__device__ __noinline__ uint32_t add_change_cold(uint32_t ws32, int lane, const ReadKernelParams& P, uint32_t ob, uint32_t oe, const uint8_t* oseq,
                                                 uint32_t olen) {
  WarpSmem* ws = smem_ptr(ws32);
  const uint32_t gwarp = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  Walk W;
  walk_load(ws, W);
  Stitch& S = W.S;
  uint32_t edits = 0;
  const ChangePlan c = plan_change(S, W.L, ob, oe, olen, 0);
  if (!c.contained) {
    const uint32_t add = olen - c.es, tn = c.re - c.rs;
    const bool emit = (S.flags & F_EMIT) != 0;
    uint8_t* fwd = P.sc_ops + (uint64_t)gwarp * P.sc_ops_cap + P.sc_ops_cap / 2;  // ops in forward order (second half of the scratch)
    if (emit) {
      if (c.gap) stitch_append(S, ws, P, W.ridx, W.raw + c.gap_src, nullptr, true, c.gap, lane);
      stitch_append(S, ws, P, W.ridx, oseq + c.es, nullptr, false, add, lane);
    }
    __syncwarp();
    const uint64_t a = warp_align(ws, P, gwarp, oseq + c.es, add, W.raw + c.rs, tn, emit ? fwd : nullptr, lane);
    __syncwarp();
    if (align_status(a)) st_fail(S, align_status(a));
    else {
      if (emit && st_status(S) == ST_OK) stitch_append(S, ws, P, W.ridx, nullptr, fwd, false, align_nops(a), lane);
      S.ledit += align_edits(a);
      edits = align_edits(a);
    }
  }
  walk_save(ws, W);
  return edits;
}

// The unchanged tail raw[pos_in_raw, L) of the read, then the partial last chunks of both rings.
__device__ __noinline__ void walk_finish(uint32_t ws32, int lane, const ReadKernelParams& P) {
  WarpSmem* ws = smem_ptr(ws32);
  Walk W;
  walk_load(ws, W);
  Stitch& S = W.S;
  if (S.pos_in_raw < W.L) stitch_append(S, ws, P, W.ridx, W.raw + S.pos_in_raw, nullptr, true, W.L - S.pos_in_raw, lane);
  // the chunk past max(len) stays inside the capacity: see flush_chunks
  if (st_status(S) == ST_OK) (void)flush_chunks(ws, P, W.ridx, S.err_len, S.err_len + CHUNK, S.cig_len, S.cig_len + CHUNK, lane);
  walk_save(ws, W);
}

// Alignment of a register-resident piece that the register aligner cannot take (more than 16 ops, or a byte outside
// ACGT): the piece goes through shared memory to the general aligners and its ops are appended to the CIGAR.
// The event's err bytes are already staged: the chunks they completed are flushed here as well (e_old, c_old = the ring
// positions at the start of the event).  Returns the edit count; lengths / status through the walk state.
__device__ __noinline__ uint32_t align_piece_cold(uint32_t ws32, int lane, const ReadKernelParams& P, uint32_t c, uint32_t es, uint32_t olen,
                                                  uint32_t rs, uint32_t tn, uint32_t e_old, uint32_t c_old) {
  WarpSmem* ws = smem_ptr(ws32);
  const uint32_t gwarp = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  Walk W;
  walk_load(ws, W);
  Stitch& S = W.S;
  const bool emit = (S.flags & F_EMIT) != 0;
  if ((uint32_t)lane < olen) ws->alt[lane] = (uint8_t)c;  // olen <= 32
  __syncwarp();
  uint8_t* fwd = P.sc_ops + (uint64_t)gwarp * P.sc_ops_cap + P.sc_ops_cap / 2;
  const uint64_t a = warp_align(ws, P, gwarp, ws->alt + es, olen - es, W.raw + rs, tn, emit ? fwd : nullptr, lane);
  __syncwarp();
  uint32_t edits = 0;
  if (align_status(a)) st_fail(S, align_status(a));
  else {
    if (emit) {
      if (((e_old ^ S.err_len) | (c_old ^ S.cig_len)) >= CHUNK) st_fail(S, flush_chunks(ws, P, W.ridx, e_old, S.err_len, c_old, S.cig_len, lane));
      stitch_append(S, ws, P, W.ridx, nullptr, fwd, false, align_nops(a), lane);
    }
    edits = align_edits(a);
    S.ledit += edits;
  }
  walk_save(ws, W);
  return edits;
}

// Hot form of add_change: the piece (olen <= 32 bytes) lives in registers, lane i holding byte i in `c`; `kb` holds the
// k-mer raw[ob + lane] it replaces.  Returns the edit count to subtract from the target.
//
// Closed form for a single random edit (fast, fk 1..3), in k-mer coordinates (piece offset j = lane - es): walking the
// reference's traceback from the end, the common suffix after the edit is taken on the diagonal (a match always wins the
// tie-break), the edit then slides left while the neighbouring bases are equal (still matches), and where it stops the
// two prefixes are identical, which forces X / I / D there and '=' before it.  Checked against the DP + tie-break
// exhaustively for all k-mers up to k = 7 over 1-, 2- and 4-letter alphabets and for every piece offset.
// `rec` (fk == 4 with packed tables, else 0): lane i < 16 holds word i of the alternative's record (DeviceModels::em_arec).
// EM: the error model the kernel is instantiated for (EM_RANDOM / EM_PACKED / EM_TABLES); only EM_RANDOM and misses of the
// file models take fk 1..3, only the file models take fk 4.
template <int EM>
__device__ __forceinline__ uint32_t add_change_reg(Walk& W, WarpSmem* ws, const ReadKernelParams& P, uint32_t ob, uint32_t oe, uint32_t c, uint32_t olen,
                                                   uint32_t fk, uint32_t fep, uint32_t kb, int lane, uint32_t rec = 0) {
  Stitch& S = W.S;
  ChangePlan cp = plan_change(S, W.L, ob, oe, olen, fk);
  if (cp.contained) return 0;
  uint32_t pk_ops = 0, pk_meta = 0;  // packed tables: the precomputed alignment of this alternative at piece offset es
  if (EM == EM_PACKED && fk == 4 && cp.fast) {
    pk_ops = __shfl_sync(0xffffffffu, rec, 4 + (int)(cp.es & 7u));
    pk_meta = (__shfl_sync(0xffffffffu, rec, 12 + (int)((cp.es & 7u) >> 1)) >> (16u * (cp.es & 1u))) & 0xffffu;
    if (cp.es > 7u || !(pk_meta & 0x8000u)) cp.fast = false;  // more than 16 ops: aligned like any other merge
  }
  const uint32_t es = cp.es, add = olen - es, tn = cp.re - cp.rs;
  const bool emit = (S.flags & F_EMIT) != 0;
  uint32_t ed = 0, nops = 0;
  bool cold = false;
  // A merge without a closed form whose window is small enough for the table of alignments (rr_device.cuh, LCS_MAX_SUM):
  // the entry is fetched now, so that the L2 / HBM round trip overlaps the emission below.  Key: the 2-bit codes of the
  // piece (lane es + i holds base i) above those of raw[rs, re) (lane j loads base j).
  uint32_t lcs_e = 0;  // 0: not in the table (an entry always carries its op count in the top bits)
  if (!cp.fast && add - 1u < LCS_MAX_M && tn - 1u < LCS_MAX_N) {
    const uint32_t lbase = c_lcs_base[(add - 1u) * LCS_MAX_N + (tn - 1u)];
    if (lbase != LCS_NO_TABLE) {
      const uint32_t tb = (uint32_t)lane < tn ? (uint32_t)W.raw[cp.rs + lane] : (uint32_t)'A';
      const bool inq = (uint32_t)lane >= es && (uint32_t)lane < olen;
      bool clean = true;
      if (S.flags & F_DIRTY) {  // bytes outside ACGT around: the table only knows the four bases ('A' 0x41, 'C' 0x43, 'G' 0x47, 'T' 0x54)
        const uint32_t dq = (inq ? c : (uint32_t)'A') - 65u, dt = tb - 65u;
        clean = !__any_sync(0xffffffffu, dq > 19u || dt > 19u || !((0x80045u >> dq) & (0x80045u >> dt) & 1u));
      }
      if (clean) {
        uint32_t part = (uint32_t)lane < tn ? ((tb >> 1) & 3u) << (2u * (uint32_t)lane) : 0u;
        if (inq) part |= ((c >> 1) & 3u) << (2u * (tn + (uint32_t)lane - es));
        lcs_e = __ldg(P.M.lcs_tab + lbase + __reduce_or_sync(0xffffffffu, part));
      }
    }
  }
  // One emission loop with a single flush site: the unchanged stretch before the change goes out 32 bases per step
  // (it is usually shorter than that), the piece and its ops go with the last step.
  uint32_t gap_left = emit ? cp.gap : 0u, gap_src = cp.gap_src;
  uint32_t e_old, c_old;
  for (;;) {
    e_old = S.err_len;
    c_old = S.cig_len;
    const uint32_t gn = min(gap_left, 32u);
    if ((uint32_t)lane < gn) ws->errb[(S.err_len + lane) & (RING - 1u)] = W.raw[gap_src + lane];
    S.err_len += gn;
    S.cig_len += gn;  // the CIGAR ring is pre-filled with '='
    gap_src += gn;
    gap_left -= gn;
    if (gap_left == 0) {
      if (emit) {
        if ((uint32_t)lane >= es && (uint32_t)lane < olen) ws->errb[(S.err_len + lane - es) & (RING - 1u)] = (uint8_t)c;
        S.err_len += add;
      }
      if (EM == EM_PACKED && cp.fast && fk == 4) {  // table alternative against its own key, packed tables: the ops came with the record
        nops = pk_meta & 31u;
        ed = (pk_meta >> 5) & 31u;
        if (emit && (uint32_t)lane < nops) ws->cigb[(S.cig_len + lane) & (RING - 1u)] = (uint8_t)(0x4449583Du >> (8u * ((pk_ops >> (2u * lane)) & 3u)));
      } else if (EM == EM_TABLES && cp.fast && fk == 4) {  // ... unpacked tables: alignments precomputed at load time
        const DeviceModels& M = P.M;
        const uint32_t pi = M.em_piece_begin[fep] + es;
        const uint32_t o0 = M.em_piece_off[pi];
        nops = M.em_piece_off[pi + 1] - o0;
        ed = M.em_piece_edits[pi];
        if (emit) for (uint32_t x = lane; x < nops; x += 32) ws->cigb[(S.cig_len + x) & (RING - 1u)] = M.em_piece_ops[o0 + x];
      } else if (cp.fast) {
        const uint32_t e = fep >= es ? fep - es : 0;    // where the edit sits inside the piece (an edit left of it acts at 0)
        const bool has_edit = !(fk == 1 && fep < es);   // a substitution in the overlapped part leaves the piece identical
        uint32_t at = e;                                // where the non-'=' op lands
        if (fk != 1) {
          // slide left while the base left of the edit equals the one it would be swapped with
          const uint32_t kprev = __shfl_up_sync(0xffffffffu, kb, 1);
          const bool sl = (uint32_t)lane > es && (uint32_t)lane <= es + e && (fk == 2 ? c == kprev : kb == kprev);
          const uint32_t B = __ballot_sync(0xffffffffu, sl) >> es;
          at = e - __clz(~(B << (31 - e)));
        }
        nops = fk == 2 ? add : tn;  // substitution / deletion: one op per raw base, insertion: one per piece base
        ed = has_edit ? 1u : 0u;
        if (emit && has_edit) ws->cigb[(S.cig_len + at) & (RING - 1u)] = (uint8_t)(0x44495800u >> (8u * fk));  // fk 1 'X', 2 'I', 3 'D'
      } else if ((add == 0 || tn == 0) && add + tn <= 32u) {  // all deletions / all insertions
        nops = add + tn;
        ed = nops;
        if (emit && (uint32_t)lane < nops) ws->cigb[(S.cig_len + lane) & (RING - 1u)] = add == 0 ? 'D' : 'I';
      } else if (lcs_e) {  // a merge without a closed form, small enough for the table of alignments (fetched above)
        nops = lcs_e >> 28;
        ed = __popc((lcs_e | (lcs_e >> 1)) & 0x01555555u);  // ops other than '='
        if (emit && (uint32_t)lane < nops) ws->cigb[(S.cig_len + lane) & (RING - 1u)] = (uint8_t)(0x4449583Du >> (8u * ((lcs_e >> (2u * lane)) & 3u)));
      } else {
        // outside the table (2 % of these merges: more than LCS_MAX_SUM bases, or a byte outside ACGT): the register aligner.
        // (Sending these to the out-of-line aligners instead was measured: 46.8 -> 49.2 ms, the parked walk state costs more.)
        uint64_t a = ~0ULL;
        if (add + tn <= 16u && add != 0 && tn != 0) {
          const uint32_t qc = __shfl_sync(0xffffffffu, c, (lane - 1 + (int)es) & 31);
          const uint32_t tc = (uint32_t)lane < tn ? W.raw[cp.rs + lane] : (uint32_t)'A';
          a = warp_align_small(qc, add, tc, tn, emit ? ws->cigb : nullptr, S.cig_len, (S.flags & F_DIRTY) != 0, lane);
        }
        if (a == ~0ULL) cold = true;  // more than 16 ops or a byte outside ACGT
        else {
          nops = align_nops(a);
          ed = align_edits(a);
        }
      }
      if (emit && !cold) S.cig_len += nops;
    }
    if (cold) break;
    if (((e_old ^ S.err_len) | (c_old ^ S.cig_len)) >= CHUNK) st_fail(S, flush_chunks(ws, P, W.ridx, e_old, S.err_len, c_old, S.cig_len, lane));
    if (gap_left == 0) break;
  }
  if (cold) {
    walk_save(ws, W);
    ed = align_piece_cold(smem_addr(ws), lane, P, c, es, olen, cp.rs, tn, e_old, c_old);
    walk_load(ws, W);
    return ed;
  }
  S.ledit += ed;
  return ed;
}

// KC: the model's k when the kernel is instantiated for it (7: Error::random and every badread model), 0 = read it from the model
template <int EM, int KC>
__device__ __forceinline__ void process_read(const ReadKernelParams& P, WarpSmem* ws, uint32_t ridx, int lane) {
  const DeviceModels& M = P.M;
  const uint32_t k = KC ? (uint32_t)KC : M.k;
  const uint32_t ws32 = smem_addr(ws);
  Walk W;
  {
    WarpRng rng;
    rng.seed(P.descs[ridx].seed, ws, lane);
    uint8_t* raw = P.raw + P.raw_off[ridx];
    StageOut so = stage_assemble(ws32, rng.off, lane, P, ridx, raw);
    W.off = so.off;
    W.L = (uint32_t)so.value;
    W.raw = raw;
    W.S.flags = (so.status || !M.em_alts_acgt) ? F_DIRTY : 0u;
  }
  W.ridx = ridx;
  W.pos = 0;
  W.gbegin = ~0u;
  W.S.err_len = W.S.cig_len = 0;
  W.S.pos_in_raw = 0;
  W.S.real_edit = 0;
  W.S.le = W.S.lee = W.S.ledit = 0;
  __syncwarp();
  ws->st[ST_CAP] = (uint32_t)(P.out_off[ridx + 1] - P.out_off[ridx]);
  for (uint32_t i = lane; i < RING / 16u; i += 32) *reinterpret_cast<uint4*>(&ws->cigb[16u * i]) = eq16();
  __syncwarp();

  // ---- K3: error::sequence (src/simulate/error.rs:201-217) ----
  W.target = (1.0 - P.descs[ridx].identity) * (double)W.L;  // number_of_edit, error.rs:196-198
  if (M.has_distance) {
    const uint32_t gwarp = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const StageOut so = stage_glitches(ws32, W.off, lane, P, gwarp, W.L);
    W.off = so.off;
    st_fail(W.S, so.status);
    __syncwarp();
    ws->st[ST_G] = 0;
    ws->st[ST_NG] = (uint32_t)so.value;
    if (so.value) W.gbegin = (uint32_t)P.sc_grec[(uint64_t)gwarp * P.sc_grec_cap].begin;
  }

  // add_error (error.rs:267-302): one Bernoulli(target / (L - pos)) per position.  32 positions are
  // tried at once against the current target; the first lane with an event wins.
  if (st_status(W.S) == ST_OK) {
    while (W.pos < W.L - k) {
      if (W.off + 64 + HOT_FIRST > RNG_WORDS || (W.S.flags & F_REPLAY)) {  // 64 words for the 32 positions + the first attempts of an event's draws
        const bool replay = (W.S.flags & F_REPLAY) != 0;
        W.S.flags &= ~F_REPLAY;
        walk_save(ws, W);
        walk_refill(ws32, lane);
        walk_load(ws, W);
        if (replay) W.S.flags |= F_REPLAYED;  // remembered until the replayed event completes
      }
      const uint32_t pi = W.pos + (uint32_t)lane;
      const double rem = (double)(W.L - pi);
      const double p = walk_div(W.target, rem);
      const uint32_t wo = W.off + 2 * (uint32_t)lane;
      const uint64_t v = (uint64_t)ws->rng[wo] | ((uint64_t)ws->rng[wo + 1] << 32);
      // A position is `rare` when something other than a plain Bernoulli(p) draw with 0 <= p < 1 happens there: the end
      // of the walk (pi == L - k), a glitch due, the break of error.rs:289-291 (target < 0 or target > rem, i.e. p > 1), or
      // p == 1 (gen_bool draws nothing).  Only then are the flags of the winning lane needed.
      const bool inb = pi < W.L - k;
      const bool gl = (pi == W.gbegin) && !(lane == 0 && (W.S.flags & F_GLITCH_CHECKED));
      const bool rare = !inb || gl || !(p < 1.0) || W.target < 0.0;
      const uint32_t m_r = __ballot_sync(0xffffffffu, rare);
      const uint32_t m_any = m_r | __ballot_sync(0xffffffffu, v < bernoulli_threshold(p));
      if (m_any == 0) {
        W.pos += 32;
        W.off += 64;
        W.S.flags &= ~F_GLITCH_CHECKED;
        continue;
      }
      const int f = __ffs(m_any) - 1;
      W.pos += (uint32_t)f;
      W.off += 2 * (uint32_t)f;
      if (f > 0) W.S.flags &= ~F_GLITCH_CHECKED;
      bool is_glitch = false;
      uint32_t consume = 2;  // words of the Bernoulli draw
      if ((m_r >> f) & 1u) {
        const bool brk = W.target < 0.0 || W.target > rem;
        const uint32_t fl = (inb ? 1u : 0u) | (gl ? 2u : 0u) | (brk ? 4u : 0u) | (p == 1.0 ? 8u : 0u);
        const uint32_t ff = __shfl_sync(0xffffffffu, fl, f);
        if (!(ff & 1u)) break;  // pos == L - k
        is_glitch = (ff & 2u) != 0;
        if (!is_glitch) {
          if (ff & 4u) break;  // error.rs:289-291
          if (ff & 8u) consume = 0;
        }
      }
      uint32_t edits;
      if (is_glitch) {  // glitch due at pos (error.rs:282-287)
        const uint32_t gwarp = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
        const GlitchRec* grec = P.sc_grec + (uint64_t)gwarp * P.sc_grec_cap;
        const uint32_t g = ws->st[ST_G];
        const GlitchRec gr = grec[g];
        W.gbegin = (g + 1 < ws->st[ST_NG]) ? (uint32_t)grec[g + 1].begin : ~0u;
        walk_save(ws, W);
        ws->st[ST_G] = g + 1;
        edits = add_change_cold(ws32, lane, P, (uint32_t)gr.begin, (uint32_t)gr.end, P.sc_gseq + (uint64_t)gwarp * P.sc_gseq_cap + gr.seq_off, gr.seq_len);
        walk_load(ws, W);
        W.S.flags |= F_GLITCH_CHECKED;  // same pos again, with the new target, glitch check done
      } else {
        const uint32_t off_ev = W.off;  // where a replay of this event starts
        W.off += consume;
        // Error::add_errors_to_kmer (src/model/error.rs:93-107): the altered k-mer is built in registers, lane i = byte i
        const uint32_t kb = (uint32_t)lane < k ? (uint32_t)W.raw[W.pos + lane] : 0u;
        uint32_t c = 0, olen, fk, fep;
        bool ok = true;
        int key = -1;
        uint2 hd = make_uint2(0u, 0u);  // packed tables: alternatives of the k-mer (0: not in the model), id of the first one
        double scale = 0.0, cm = 0.0;
        const double* cump = nullptr;
        if (EM != EM_RANDOM) {  // 2-bit code of the k-mer; any byte outside ACGT misses the table
          uint32_t part = 0, bad = 0;
          if ((uint32_t)lane < k) {
            uint32_t s = kb == 'A' ? 0u : kb == 'C' ? 1u : kb == 'G' ? 2u : kb == 'T' ? 3u : 4u;
            bad = s > 3u;
            part = (s & 3u) << (2 * (k - 1 - lane));
          }
          const uint32_t code = __reduce_or_sync(0xffffffffu, part);
          if (!__any_sync(0xffffffffu, bad)) {
            if (EM == EM_PACKED) {  // key record, directly indexed; its first 32 partial sums travel with the header
              const uint8_t* kp = M.em_kblob + (size_t)code * M.em_kstride;
              hd = __ldg(reinterpret_cast<const uint2*>(kp));
              scale = __ldg(reinterpret_cast<const double*>(kp + 8));
              cump = reinterpret_cast<const double*>(kp + 16);
              cm = __ldg(cump + lane);
            } else key = M.em_index[code];
          }
        }
        if (EM == EM_PACKED && hd.x != 0) {  // one WeightedIndex draw over the alternatives of the k-mer (src/model/error.rs:96-100)
          const uint32_t na = hd.x;
          const uint64_t v2 = (uint64_t)ws->rng[W.off] | ((uint64_t)ws->rng[W.off + 1] << 32);  // first attempt: resident (HOT_FIRST)
          W.off += 2;
          const double chosen = u64_to_unit12(v2) * scale + 0.0;
          uint32_t idx = 0;  // number of partial sums <= chosen among the first na - 1
          for (uint32_t b = 0; b + 1 < na; b += 32) {
            const uint32_t i = b + lane;
            if (b) cm = (i + 1 < na) ? __ldg(cump + i) : 0.0;
            idx += __popc(__ballot_sync(0xffffffffu, (i + 1 < na) && (cm <= chosen)));
          }
          fk = 4;
          fep = hd.y + idx;  // global id of the chosen alternative
          const uint32_t rec = (uint32_t)lane < EM_AREC_WORDS ? __ldg(M.em_arec + (size_t)fep * EM_AREC_WORDS + lane) : 0u;
          olen = __shfl_sync(0xffffffffu, rec, 3) >> 24;
          c = (__shfl_sync(0xffffffffu, rec, lane >> 2) >> (8u * ((uint32_t)lane & 3u))) & 0xffu;  // lane i = byte i (i < olen <= 15)
          edits = add_change_reg<EM>(W, ws, P, W.pos, W.pos + k, c, olen, fk, fep, kb, lane, rec);
        } else if (EM == EM_TABLES && key >= 0) {  // one WeightedIndex draw over the alternatives of the k-mer (src/model/error.rs:96-100)
          const uint32_t a0 = M.em_alt_begin[key], na = M.em_alt_begin[key + 1] - a0;
          const uint64_t v2 = (uint64_t)ws->rng[W.off] | ((uint64_t)ws->rng[W.off + 1] << 32);  // first attempt: resident (HOT_FIRST)
          W.off += 2;
          const double chosen = u64_to_unit12(v2) * M.em_scale[key] + 0.0;
          uint32_t idx = 0;  // number of partial sums <= chosen among the first na - 1
          for (uint32_t b = 0; b + 1 < na; b += 32) {
            const uint32_t i = b + lane;
            const bool le = (i + 1 < na) && (M.em_cum[a0 + i] <= chosen);
            idx += __popc(__ballot_sync(0xffffffffu, le));
          }
          fk = 4;
          fep = a0 + idx;  // global id of the chosen alternative (it indexes the precomputed alignments)
          const uint32_t o0 = M.em_alt_off[fep];
          olen = M.em_alt_off[fep + 1] - o0;
          if (olen <= 32u) {
            if ((uint32_t)lane < olen) c = M.em_alt_bytes[o0 + lane];
            edits = add_change_reg<EM>(W, ws, P, W.pos, W.pos + k, c, olen, fk, fep, kb, lane);
          } else {  // an alternative longer than a warp: through memory
            walk_save(ws, W);
            edits = add_change_cold(ws32, lane, P, W.pos, W.pos + k, M.em_alt_bytes + o0, olen);
            walk_load(ws, W);
          }
        } else {  // random_error (src/model/error.rs:116-135)
          const uint32_t ep = hot_index(ws, W.off, ok, k);
          const uint32_t kind = ok ? hot_range_u32(ws, W.off, ok, 1, 3) : 1u;  // 1 substitution, 2 insertion, 3 deletion
          if (kind == 1) {
            const uint32_t nb = ok ? hot_base_diff(ws, W.off, ok, __shfl_sync(0xffffffffu, kb, ep)) : 0u;
            c = (uint32_t)lane == ep ? nb : kb;
            olen = k;
          } else if (kind == 2) {
            const uint32_t nb = ok ? hot_base_diff(ws, W.off, ok, 256u) : 0u;
            const uint32_t kprev = __shfl_up_sync(0xffffffffu, kb, 1);
            c = (uint32_t)lane < ep ? kb : ((uint32_t)lane == ep ? nb : kprev);
            olen = k + 1;
          } else {
            const uint32_t knext = __shfl_down_sync(0xffffffffu, kb, 1);
            c = (uint32_t)lane < ep ? kb : knext;
            olen = k - 1;
          }
          if (!ok) {  // the draws ran past the resident words: refill and replay this event from its Bernoulli draw
            if (W.S.flags & F_REPLAYED) { st_fail(W.S, ST_RNG_EXHAUSTED); break; }  // > 400 words of rejections in one event
            W.off = off_ev;
            W.S.flags |= F_REPLAY;
            continue;
          }
          fk = kind;
          fep = ep;
          edits = add_change_reg<EM>(W, ws, P, W.pos, W.pos + k, c, olen, fk, fep, kb, lane);
        }
        W.S.flags &= ~F_REPLAYED;
      }
      W.target -= (double)edits;
      if (st_status(W.S)) break;
      if (!is_glitch) {
        if (W.target < 0.0 || W.target > (double)(W.L - W.pos)) break;  // error.rs:298-300
        W.pos += 1;
        W.S.flags &= ~F_GLITCH_CHECKED;
      }
    }
  }

  // ---- K5: tail of the stitch (error.rs:241-246) ----
  if (st_status(W.S) == ST_OK) {
    if ((W.S.flags & (F_HAS_LAST | F_EMIT)) == (F_HAS_LAST | F_EMIT)) {
      W.S.real_edit += W.S.ledit;
      W.S.pos_in_raw = W.S.le;
    }
    walk_save(ws, W);
    walk_finish(ws32, lane, P);
    walk_load(ws, W);
  }
  __syncwarp();
  if (lane == 0) {
    const uint32_t status = st_status(W.S);
    P.raw_len[W.ridx] = W.L;
    P.err_len[W.ridx] = status ? 0 : W.S.err_len;
    P.cig_len[W.ridx] = status ? 0 : W.S.cig_len;
    P.q_cursor[W.ridx] = ws->rng_base + W.off;
    P.real_id[W.ridx] = (1.0 - ((double)W.S.real_edit / (double)W.L)) * 100.0;  // error.rs:246, mod.rs:289
    P.status[W.ridx] = status;
  }
}

// One instantiation per kind of error model, so that the hot loop of each holds only its own path (the bench's Error::random
// kernel carries no table code at all).
template <int EM, int KC>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, RR_BLOCKS_PER_SM) rr_read_kernel(const __grid_constant__ ReadKernelParams P) {
  // dynamic shared memory (WARPS_PER_BLOCK x WarpSmem is past the 48 KB a kernel may declare statically)
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  WarpSmem* smem = reinterpret_cast<WarpSmem*>(smem_dyn);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpSmem* ws = &smem[warp];
  // fast-lane SMs (see ReadKernelParams): a wrong guess about the numbering of the SMs only changes who takes which read
  bool fast = false;
  if (P.n_fast_sms) {
    uint32_t smid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    fast = smid % P.fast_stride == 0 && smid / P.fast_stride < P.n_fast_sms;
    if (fast && (uint32_t)warp >= P.fast_warps) return;
  }
  bool first_open = true;  // the queue this warp prefers still has entries
  for (;;) {
    uint32_t slot = ~0u;
    if (lane == 0) {
      // fast lanes: heavy reads first, then the rest; the others: the rest first, then whatever heavy reads are left
      if (first_open) {
        slot = fast ? atomicAdd(P.queue + 2, 1u) : P.n_heavy + atomicAdd(P.queue, 1u);
        if (slot >= (fast ? P.n_heavy : P.n_reads)) slot = ~0u;
      }
      if (slot == ~0u) {
        slot = fast ? P.n_heavy + atomicAdd(P.queue, 1u) : atomicAdd(P.queue + 2, 1u);
        if (slot >= (fast ? P.n_reads : P.n_heavy)) slot = ~0u;
      }
    }
    const uint32_t got = __shfl_sync(0xffffffffu, slot, 0);
    if (got == ~0u) break;
    if (fast ? got >= P.n_heavy : got < P.n_heavy) first_open = false;
    process_read<EM, KC>(P, ws, P.order[got], lane);
    __syncwarp();
  }
}

}  // namespace rr
