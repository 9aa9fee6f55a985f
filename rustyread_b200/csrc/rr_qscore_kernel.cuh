// rustyread_b200 — K6: quality::generate_quality (src/simulate/quality.rs:14-38) +
// Quality::get_qscore (src/model/quality.rs:113-133), and K7a: packing of the sequences (same kernel).
//
// get_qscore consumes exactly one u64 of the read's stream per non-'D' CIGAR position, so
// position i with rank r (non-'D' positions before it) uses stream words q_cursor + 2r, +1.
// A block walks one read in tiles: prefix-rank the tile, generate exactly the ChaCha blocks the
// tile needs (one block per thread), then every thread samples its positions independently.
//
// The "strip one char from both ends until the window is known" fallback of get_qscore is
// resolved once on the host into dense tables: for every odd length l <= max_k, a 4^l-entry
// table (alphabet = X I D) -> distribution id.
#pragma once
#include "rr_device.cuh"

namespace rr {

#ifndef RR_Q_THREADS
#define RR_Q_THREADS 128  // measured at 200x: 64 / 128 / 256 / 512 threads (all with 2048 resident threads per SM) -> 5.41 / 5.37 / 5.60 / 6.23 ms
#endif
constexpr int Q_THREADS = RR_Q_THREADS;
constexpr int Q_PER_THREAD = 8;
constexpr int Q_TILE = Q_THREADS * Q_PER_THREAD - 8;  // multiple of 8; 2 * Q_TILE + 15 words fit in Q_THREADS ChaCha blocks (one per thread)
static_assert(2 * Q_TILE + 15 <= 16 * Q_THREADS, "a tile's draws must fit the blocks its threads generate");
constexpr int Q_HALO = 8;     // (max_k - 1) / 2 <= 7
constexpr uint32_t QD_CACHE = 128;   // distributions kept in shared memory when the whole model fits (random / ideal / small files)
constexpr uint32_t QS_CACHE = 2048;  // ... and their scores

// One distribution of the model: where its scores / partial sums start, how many there are (top bit: every weight is
// exactly 1.0, so the partial sums are 1, 2, 3, ... and the search is a truncation), and the UniformFloat scale.
struct __align__(16) QDist {
  uint32_t d0, nd_unit;
  double scale;
};

struct QscoreModelDev {
  uint32_t max_k;
  uint32_t n_dist, n_scores;
  const uint32_t* tab_off;   // indexed by window length l (odd) -> offset into tab
  const uint16_t* tab;       // dist id, 0xFFFF = unknown
  const QDist* dist;         // n_dist
  const uint8_t* scores;
  const double* cum;         // per dist: partial sums, last = total
  // Per distribution Q_GUIDE + 1 bounds of the WeightedIndex search: guide[g] = number of partial sums <= (g / Q_GUIDE) * scale,
  // guide[Q_GUIDE] = count - 1.  A draw u in [g / Q_GUIDE, (g + 1) / Q_GUIDE) selects an index in [guide[g], guide[g + 1]]
  // (u * scale and the count are monotone), so the search over ~45 scores shrinks to the one or two that share a bucket.
  const uint16_t* guide;
  // The same model as one record per distribution, for models that do not fit shared memory (a draw then costs three dependent
  // L2 round trips instead of five): tab32 maps a window straight to its record (offset in 16-byte units, Q_NO_REC = unknown);
  // record = {f64 scale, u32 count | unit flag, u32 0}, 17 u16 search bounds (48 bytes), then count pairs {f64 partial sum, u64 score}.
  const uint32_t* tab32;
  const uint8_t* qrec;
};
constexpr uint32_t Q_NO_REC = 0xffffffffu, Q_REC_HEAD = 4;  // 16-byte units before the pairs
constexpr uint32_t Q_GUIDE = 16;

struct QscoreKernelParams {
  QscoreModelDev Q;
  const rr_desc* descs;
  const uint32_t* order;
  uint32_t n_reads;
  uint32_t* queue;
  const uint64_t* out_off;     // cigar arena offsets
  const uint8_t* cig;
  const uint64_t* cig_len;
  const uint64_t* err_len;
  const uint64_t* q_cursor;
  const uint64_t* packed_off;  // where read i's quality string (and packed sequence) goes
  uint8_t* qual;
  uint32_t* status;
  // K7a, fused: the block that draws a read's qualities also moves its sequence from the err arena to the packed one
  const uint8_t* err;
  uint8_t* seq;
};

// K1: max_k == 1 (Quality::random / ideal): the window is the position itself.  CACHED: the whole model sits in shared memory.
template <bool K1, bool CACHED>
#ifndef RR_Q_MINB
#define RR_Q_MINB (2048 / RR_Q_THREADS)  // 32 registers, 64 warps per SM: the kernel wants occupancy (256 threads: 8.7 / 6.4 / 5.8 / 5.6 ms at 200x for 1 / 4 / 6 / 8 blocks)
#endif
__global__ void __launch_bounds__(Q_THREADS, RR_Q_MINB) rr_qscore_kernel(const __grid_constant__ QscoreKernelParams P) {
  __shared__ __align__(16) uint32_t s_words[Q_THREADS * 16];
  __shared__ __align__(16) uint8_t s_cig[Q_HALO + Q_THREADS * Q_PER_THREAD + Q_HALO];  // every thread loads its 8 ops (+ 8 on either side), also the ones past the tile
  __shared__ __align__(16) QDist s_dist[CACHED ? QD_CACHE : 1];
  __shared__ uint8_t s_scores[CACHED ? QS_CACHE : 1];
  __shared__ uint32_t s_warp_sum[Q_THREADS / 32];
  __shared__ uint32_t s_slot;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const QscoreModelDev& Q = P.Q;
  const uint32_t margin = K1 ? 0u : (Q.max_k - 1) / 2;
  if (CACHED) {
    for (uint32_t i = tid; i < Q.n_dist; i += Q_THREADS) s_dist[i] = Q.dist[i];
    for (uint32_t i = tid; i < Q.n_scores; i += Q_THREADS) s_scores[i] = Q.scores[i];
  }
  uint32_t dsym0 = 0xFFFFu, dsym1 = 0xFFFFu, dsym2 = 0xFFFFu;  // K1: distribution (CACHED) or record (!CACHED) of '=', 'X', 'I'
  if (K1) {
    const uint32_t o = Q.tab_off[1];
    dsym0 = CACHED ? (uint32_t)Q.tab[o] : Q.tab32[o];
    dsym1 = CACHED ? (uint32_t)Q.tab[o + 1] : Q.tab32[o + 1];
    dsym2 = CACHED ? (uint32_t)Q.tab[o + 2] : Q.tab32[o + 2];
  }

  for (;;) {
    __syncthreads();
    if (tid == 0) s_slot = atomicAdd(P.queue, 1u);
    __syncthreads();
    const uint32_t slot = s_slot;
    if (slot >= P.n_reads) break;
    const uint32_t ridx = P.order[slot];
    if (P.status[ridx] != ST_OK) continue;
    const uint32_t n = (uint32_t)P.cig_len[ridx];  // < 2^31 (the host refuses longer reads)
    const uint8_t* cig = P.cig + P.out_off[ridx];
    uint8_t* qual = P.qual + P.packed_off[ridx];
    const uint32_t elen = (uint32_t)P.err_len[ridx];
    const uint64_t qcur = P.q_cursor[ridx];
    uint32_t key[8];
    chacha_key_from_u64(P.descs[ridx].seed, key);
    {  // err arena -> packed sequence arena (both 16-byte aligned per read); the loads are in flight while the tiles below compute
      const uint4* src = reinterpret_cast<const uint4*>(P.err + P.out_off[ridx]);
      uint4* dst = reinterpret_cast<uint4*>(P.seq + P.packed_off[ridx]);
      const uint32_t nv = (elen + 15u) / 16u;  // capacity is padded to 16 on both sides
      for (uint32_t i = tid; i < nv; i += Q_THREADS) {
        uint4 v = src[i];
        if (16u * i + 16u > elen) {  // last word: the padding after the read is zeroed (what lies beyond err_len in the arena is stale staging data)
          uint32_t w[4] = {v.x, v.y, v.z, v.w};
          const uint32_t keep = elen - 16u * i;
#pragma unroll
          for (uint32_t j = 0; j < 4; ++j) w[j] = keep >= 4 * j + 4 ? w[j] : (keep > 4 * j ? (w[j] & ((1u << (8 * (keep - 4 * j))) - 1u)) : 0u);
          v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        dst[i] = v;
      }
    }

    uint32_t rank0 = 0;  // non-'D' positions before this tile
    for (uint32_t t0 = 0; t0 < n; t0 += Q_TILE) {
      const uint32_t tn = min((uint32_t)Q_TILE, n - t0);
      __syncthreads();
      // stage the tile (+ halo) of the CIGAR
      // body: one aligned 8-byte load per thread (read offsets are 16-byte aligned, tiles are multiples of 8)
      {
        const uint32_t p8 = tid * 8;
        if (p8 + 8 <= tn) {
          *reinterpret_cast<uint2*>(s_cig + Q_HALO + p8) = *reinterpret_cast<const uint2*>(cig + t0 + p8);
        } else {
          for (uint32_t x = p8; x < tn; ++x) s_cig[Q_HALO + x] = cig[t0 + x];
        }
        if (!K1 && tid < 2 * Q_HALO) {  // halo: 8 positions before the tile, 8 after
          const int64_t gi = tid < Q_HALO ? (int64_t)t0 - Q_HALO + tid : (int64_t)t0 + tn + (tid - Q_HALO);
          s_cig[tid < Q_HALO ? tid : Q_HALO + tn + (tid - Q_HALO)] = (gi >= 0 && gi < (int64_t)n) ? cig[gi] : 0;
        }
      }
      __syncthreads();
      // rank the non-'D' positions: 8 consecutive positions per thread
      const uint32_t p0 = tid * Q_PER_THREAD;
      uint32_t flags = 0;
      const uint2 cw = *reinterpret_cast<const uint2*>(s_cig + Q_HALO + p0);  // this thread's 8 ops
#pragma unroll
      for (int x = 0; x < Q_PER_THREAD; ++x) {
        const uint32_t c = ((x < 4 ? cw.x : cw.y) >> (8 * (x & 3))) & 0xffu;
        if (p0 + x < tn && c != 'D') flags |= 1u << x;
      }
      uint32_t mine = __popc(flags);
      uint32_t incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      if (lane == 31) s_warp_sum[wid] = incl;
      __syncthreads();
      uint32_t wbase = 0, total = 0;
#pragma unroll
      for (int w = 0; w < Q_THREADS / 32; ++w) {
        uint32_t sw = s_warp_sum[w];
        if (w < wid) wbase += sw;
        total += sw;
      }
      const uint32_t excl = wbase + incl - mine;
      // the ChaCha blocks holding words [qcur + 2 rank0, qcur + 2 (rank0 + total))
      const uint64_t w_first = qcur + 2 * (uint64_t)rank0;
      const uint64_t b_first = w_first >> 4;
      if (total) {
        const uint64_t b_last = (w_first + 2 * (uint64_t)total - 1) >> 4;
        if (b_first + (uint64_t)tid <= b_last) {
          uint32_t out[16];
          chacha12_block(key, b_first + (uint64_t)tid, out);
          uint4* dst = reinterpret_cast<uint4*>(s_words + (uint32_t)tid * 16u);
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = make_uint4(out[4 * j], out[4 * j + 1], out[4 * j + 2], out[4 * j + 3]);
        }
      }
      __syncthreads();
      const bool edge = !K1 && (t0 < margin || t0 + tn + margin > n);  // tiles that touch an end of the read shrink their windows
      // The 2-bit symbols ('=' 0, 'X' 1, 'I' 2, 'D' / outside the read 3) of ops p0 - 8 .. p0 + 15, 4 bytes at a time: with
      // '=' 0x3D, 'X' 0x58, 'I' 0x49, 'D' 0x44 the symbol is (~bit4, ~bit0) of the byte.  A window is then a shift and a mask.
      uint64_t syms = 0;
      if (!K1) {
        const uint2* wp = reinterpret_cast<const uint2*>(s_cig + Q_HALO + p0 - 8);
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          const uint2 w2 = wp[g];
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const uint32_t y = ~(h2 ? w2.y : w2.x);
            const uint32_t b = (y & 0x01010101u) | ((y >> 3) & 0x02020202u);  // one symbol in the low 2 bits of every byte
            const uint32_t pk = (b | (b >> 6) | (b >> 12) | (b >> 18)) & 0xffu;
            syms |= (uint64_t)pk << (8 * (2 * g + h2));
          }
        }
      }
      uint32_t woff = (uint32_t)(w_first & 15ULL) + 2 * excl;
      uint32_t qi = rank0 + excl;
#pragma unroll
      for (int x = 0; x < Q_PER_THREAD; ++x) {
        if (!(flags & (1u << x))) continue;
        uint32_t dist;
        if constexpr (K1) {
          const uint32_t c = ((x < 4 ? cw.x : cw.y) >> (8 * (x & 3))) & 0xffu;
          dist = c == '=' ? dsym0 : c == 'X' ? dsym1 : dsym2;
        } else {
          // centred window, shrunk symmetrically at both ends (quality.rs:22-34)
          uint32_t half = margin;
          if (edge) {
            const uint32_t i = t0 + p0 + x;  // position in the CIGAR
            if (i < margin) half = i;
            else if (i >= n - margin) half = n - i - 1;  // wraps like the reference's usize when n < margin
            if (i < half || i + half + 1 > n) {          // the slice cigar[start..end] is out of range: the reference panics
              P.status[ridx] = ST_QUAL_WINDOW;
              half = 0;
            }
          }
          const uint32_t len = 2 * half + 1;
          // ops p0 + x - half .. p0 + x + half = symbols 8 + x - half .. of the thread's 24-symbol word
          const uint32_t code = (uint32_t)(syms >> (2 * (8 + x - half))) & ((1u << (2 * len)) - 1u);
          dist = CACHED ? (uint32_t)Q.tab[Q.tab_off[len] + code] : __ldg(Q.tab32 + Q.tab_off[len] + code);
        }
        uint32_t qv = '!';
        if (dist == (CACHED ? 0xFFFFu : Q_NO_REC)) {
          P.status[ridx] = ST_QUAL;
        } else {
          const uint64_t v = (uint64_t)s_words[woff] | ((uint64_t)s_words[woff + 1] << 32);
          if constexpr (CACHED) {
            const QDist d = s_dist[dist];
            const uint32_t hi = (d.nd_unit & 0x7fffffffu) - 1;
            const double chosen = u64_to_unit12(v) * d.scale + 0.0;
            uint32_t lo = 0;  // partition_point over the first nd - 1 partial sums
            if (d.nd_unit >> 31) {
              lo = min((uint32_t)chosen, hi);  // partial sums are exactly 1, 2, ..: #{i : i + 1 <= chosen} = floor(chosen)
            } else {
              const uint16_t* gd = Q.guide + (size_t)dist * (Q_GUIDE + 1u) + (uint32_t)(u64_to_unit12(v) * (double)Q_GUIDE);
              lo = gd[0];
              uint32_t up = gd[1];
              while (lo < up) {
                uint32_t mid = (lo + up) >> 1;
                if (Q.cum[d.d0 + mid] <= chosen) lo = mid + 1; else up = mid;
              }
            }
            qv = s_scores[d.d0 + lo] + 33u;
          } else {  // one record per distribution: header + search bounds, then {partial sum, score} pairs
            const uint4* rec = reinterpret_cast<const uint4*>(Q.qrec) + dist;
            const double u = u64_to_unit12(v);
            const uint16_t* gd = reinterpret_cast<const uint16_t*>(rec + 1) + (uint32_t)(u * (double)Q_GUIDE);
            const uint4 h = __ldg(rec);
            uint32_t lo = __ldg(gd), up = __ldg(gd + 1);  // in flight together with the header
            const uint32_t hi = (h.z & 0x7fffffffu) - 1;
            const double chosen = u * __hiloint2double((int)h.y, (int)h.x) + 0.0;
            if (h.z >> 31) {
              lo = min((uint32_t)chosen, hi);
            } else {
              while (lo < up) {
                const uint32_t mid = (lo + up) >> 1;
                if (__ldg(reinterpret_cast<const double*>(rec + Q_REC_HEAD + mid)) <= chosen) lo = mid + 1; else up = mid;
              }
            }
            qv = (__ldg(reinterpret_cast<const uint32_t*>(rec + Q_REC_HEAD + lo) + 2) & 0xffu) + 33u;
          }
        }
        if (qi < elen) qual[qi] = (uint8_t)qv;
        woff += 2;
        ++qi;
      }
      rank0 += total;
    }
    // mod.rs:284-287: the quality string is resized to the sequence length with '!'; the padding up to the record's
    // 16-byte boundary is zeroed like the sequence's (above), so that the drained arena is deterministic
    for (uint32_t x = min(rank0, elen) + tid; x < ((elen + 15u) & ~15u); x += Q_THREADS) qual[x] = x < elen ? (uint8_t)'!' : (uint8_t)0;
  }
}

// packed offsets: exclusive scan of round_up(err_len, 16); single block, n up to a few million
__global__ void __launch_bounds__(1024) rr_scan_kernel(const uint64_t* __restrict__ err_len, const uint32_t* __restrict__ status,
                                                        uint32_t n, uint32_t k, uint64_t* __restrict__ packed_off,
                                                        uint64_t* __restrict__ totals /* [0] bytes, [1] bases, [2] failed */) {
  __shared__ uint64_t s_part[1024];
  __shared__ uint64_t s_bases[1024];
  __shared__ uint64_t s_fail[1024];
  const uint32_t tid = threadIdx.x;
  const uint32_t per = (n + 1023) / 1024;
  const uint32_t b = tid * per, e = min(n, b + per);
  uint64_t sum = 0, bases = 0, fail = 0;
  for (uint32_t i = b; i < e; ++i) {
    uint64_t l = err_len[i];
    sum += (l + 15) & ~15ULL;
    if (l > 14 && l > 2ULL * k) bases += l - 2ULL * k;  // writer: mod.rs:192-208
    fail += status[i] != ST_OK;
  }
  s_part[tid] = sum;
  s_bases[tid] = bases;
  s_fail[tid] = fail;
  __syncthreads();
  if (tid == 0) {
    uint64_t acc = 0, ab = 0, af = 0;
    for (int i = 0; i < 1024; ++i) {
      uint64_t x = s_part[i];
      s_part[i] = acc;
      acc += x;
      ab += s_bases[i];
      af += s_fail[i];
    }
    packed_off[n] = acc;
    totals[0] = acc;
    totals[1] = ab;
    totals[2] = af;
  }
  __syncthreads();
  uint64_t off = s_part[tid];
  for (uint32_t i = b; i < e; ++i) {
    packed_off[i] = off;
    off += (err_len[i] + 15) & ~15ULL;
  }
}

}  // namespace rr
