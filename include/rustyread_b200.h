/*
 * rustyread_b200 — C ABI of the B200-native `simulate` hot path.
 *
 * The reference (natir/rustyread, Rust) exposes no FFI / plugin interface.  The seam this
 * library plugs into is the body of the rayon closure at src/simulate/mod.rs:172-186, i.e.
 * `generate_read` (src/simulate/mod.rs:223-293) applied to every work item the serial
 * `Fragments` iterator yields (src/simulate/fragments.rs:132-164).  The entry points below are
 * the batched form of that closure: what a Rust `extern "C"` block in src/simulate/mod.rs
 * would bind (see INTEGRATION.md for the stub).  Plain pointers and sizes only.
 *
 * There is NO CPU fallback: every rr_* compute entry point runs hand-written sm_100a CUDA
 * kernels and fails (non-zero return + rr_last_error) when no CUDA device is usable.
 */
#ifndef RUSTYREAD_B200_H
#define RUSTYREAD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rr_ctx rr_ctx;

/* return codes */
enum {
  RR_OK = 0,
  RR_ERR_CUDA = 1,        /* CUDA runtime error or no device */
  RR_ERR_ARG = 2,         /* invalid argument (the reference would reject or panic on it) */
  RR_ERR_UNSUPPORTED = 3, /* valid for the reference but outside what the device tables hold */
  RR_ERR_READ = 4         /* a read hit a condition on which the reference aborts (see rr_last_error) */
};

/* read types: src/simulate/description.rs:11-16 */
enum { RR_REAL = 0, RR_JUNK = 1, RR_RANDOM = 2 };
/* strands: '+' / '-' of Origin.strand (description.rs:22); junk/random carry '*' */
enum { RR_PLUS = 0, RR_MINUS = 1, RR_NOSTRAND = 2 };

/*
 * One work item = what Fragments::next yields: (ref_idx, ref_idx2, Description, seed)
 * src/simulate/fragments.rs:130,153-163 ; Origin fields: src/simulate/description.rs:19-26.
 * For RR_JUNK / RR_RANDOM `start` is 0 and `end` is the requested length (description.rs:39-57).
 */
typedef struct rr_desc {
  uint32_t type;        /* origin.read_type */
  uint32_t ref_idx;     /* index into the contigs given to rr_set_references */
  uint32_t strand;      /* origin.strand */
  uint32_t has_chimera; /* description.chimera.is_some() */
  uint32_t ctype;       /* chimera.read_type */
  uint32_t cref_idx;
  uint32_t cstrand;
  uint32_t reserved;
  uint64_t start, end;   /* origin.start / origin.end */
  uint64_t cstart, cend; /* chimera.start / chimera.end */
  double identity;       /* description.identity (target identity in [0,1]) */
  uint64_t seed;         /* per-read seed: StdRng::seed_from_u64(seed), mod.rs:183 */
} rr_desc;

/*
 * Result of one batch = Vec<(Description, Seq, Quality)> of mod.rs:163-187, flattened.
 * All pointers are library-owned pinned host memory, valid until the next rr_generate /
 * rr_download / rr_destroy on the same context.  Read i occupies
 * seq[offsets[i] .. offsets[i] + lengths[i]) and the same range of qual (records are padded to
 * 16 bytes so that the device packs them with 128-bit stores); these are the FULL erroneous
 * fragments including the k random flank bases that the writer trims (mod.rs:204-208).
 * lengths[i] = description.length, real_identity[i] = description.identity after
 * mod.rs:289-290 (percent).
 */
typedef struct rr_batch {
  uint64_t n_reads;
  const uint64_t* offsets; /* n_reads + 1 ; offsets[n_reads] = total_bytes */
  const uint64_t* lengths; /* n_reads */
  const uint8_t* seq;
  const uint8_t* qual;
  const double* real_identity;
  uint64_t total_bytes;
  uint64_t total_bases; /* what the writer would emit: sum of lengths[i] - 2k over lengths[i] > 14 */
  uint32_t k;           /* error_model.k() */
} rr_batch;

/* ---- context --------------------------------------------------------------------------- */
/* One context per GPU; owns the stream, device tables and pinned staging. Not re-entrant. */
int rr_create(int device, rr_ctx** out);
void rr_destroy(rr_ctx* ctx);
const char* rr_last_error(const rr_ctx* ctx); /* ctx may be NULL: error of a failed rr_create */
int rr_device_count(void);

/* ---- immutable inputs (the `&` borrows of generate_read, mod.rs:224-230) --------------------- */
/* References::sequences (src/references.rs:17-35).  Bytes are taken as read: no case folding
 * (references.rs:162-166).  The reverse complement (references.rs:27) is computed on the fly. */
int rr_set_references(rr_ctx* ctx, uint64_t n, const uint8_t* const* seq, const uint64_t* len,
                      const uint8_t* circular);
/* The same from DEVICE memory of the context's GPU: `device_bases` holds the n contigs back to back (sum of len bytes),
 * e.g. the receive buffer of an NCCL broadcast from the rank that read the FASTA.  Copied device to device; the caller
 * may free its buffer when the call returns. */
int rr_set_references_device(rr_ctx* ctx, uint64_t n, const uint64_t* len, const uint8_t* circular,
                             const void* device_bases);
/* Device address and size of the context's reference arena (contigs back to back), valid until the references are set
 * again or the context is destroyed: the send buffer of that broadcast. */
int rr_references_device(const rr_ctx* ctx, const void** device_bases, uint64_t* bytes);

/* model::Error (src/model/error.rs:19-22), AFTER loading: for key i (k bytes at keys + i*k) the
 * alternatives alt_begin[i] .. alt_begin[i+1], alternative j = alt_bytes[alt_off[j] .. alt_off[j+1])
 * with weight weights[j] (error.rs:65-73: random alt appended, key removed).  n_keys == 0 is
 * Error::random(k) (error.rs:85-90). */
int rr_set_error_model(rr_ctx* ctx, uint32_t k, uint64_t n_keys, const uint8_t* keys,
                       const uint32_t* alt_begin, const uint32_t* alt_off, const uint8_t* alt_bytes,
                       const double* weights);

/* model::Quality (src/model/quality.rs:19-22): entry i has key cigars[cigar_off[i] .. cigar_off[i+1])
 * and (scores, weights)[score_off[i] .. score_off[i+1]).  max_k as quality.rs:136-138. */
int rr_set_qscore_model(rr_ctx* ctx, uint32_t max_k, uint64_t n, const uint8_t* cigars,
                        const uint32_t* cigar_off, const uint32_t* score_off, const uint8_t* scores,
                        const double* weights);

/* model::Adapter::new (src/model/adapter.rs:23-61): same arguments, same units (percent). */
int rr_set_adapter(rr_ctx* ctx, const uint8_t* start, uint64_t start_len, const uint8_t* end,
                   uint64_t end_len, double start_rate, double start_amount, double end_rate,
                   double end_amount);

/* model::Glitch::new (src/model/glitch.rs:20-48): same arguments. */
int rr_set_glitch(rr_ctx* ctx, double rate, double size, double skip);

/* ---- the hot path ------------------------------------------------------------------------------ */
/* generate_read for n work items (host buffers in, host buffers out; copies inside). */
int rr_generate(rr_ctx* ctx, uint64_t n, const rr_desc* descs, rr_batch* out);

/* The same in three steps, so that a caller can keep a batch resident in HBM:
 *   rr_upload   H2D of the work items (+ capacity planning)
 *   rr_run      the kernels only (assemble -> error walk -> qscores -> pack); repeatable
 *   rr_download D2H of the packed result into pinned memory
 */
int rr_upload(rr_ctx* ctx, uint64_t n, const rr_desc* descs);
int rr_run(rr_ctx* ctx);
int rr_download(rr_ctx* ctx, rr_batch* out);

/* Asynchronous, double-buffered form, so that the host can format and write batch i (mod.rs:191-212) while
 * batch i+1 is generated: rr_submit queues H2D + kernels and returns at once (at most two batches in flight);
 * rr_collect waits for the OLDEST submitted batch and drains it through pinned memory on a copy stream, which
 * overlaps the kernels of the batch submitted after it.  The result stays valid until that slot is reused,
 * i.e. until the second rr_submit after the one that produced it.  rr_generate == rr_submit + rr_collect. */
int rr_submit(rr_ctx* ctx, uint64_t n, const rr_desc* descs);
int rr_collect(rr_ctx* ctx, rr_batch* out);

/* Streaming form of rr_collect for a caller that consumes the batch read by read, as the FASTQ writer does
 * (mod.rs:191-212): the oldest submitted batch is handed over in consecutive parts of about 32 MiB of sequence through a
 * ring of three small pinned buffers, so no whole-batch pinned staging is allocated (page-locking gigabytes costs
 * seconds) and the D2H of part p + 1 runs while the caller formats part p.
 *   rr_collect_begin  waits for the batch's kernels; *n_reads = reads of the batch
 *   rr_collect_part   next part: an rr_batch whose read 0 is read *first of the batch (offsets relative to the part's
 *                     arenas); valid until the next rr_collect_part.  part->n_reads == 0: the batch is finished and
 *                     counts as collected. */
int rr_collect_begin(rr_ctx* ctx, uint64_t* n_reads, uint64_t* total_bases);
int rr_collect_part(rr_ctx* ctx, rr_batch* part, uint64_t* first);
/* Between rr_collect_begin and the last rr_collect_part: the lengths and identities (mod.rs:289-290) of ALL reads of the
 * draining batch, library-owned.  A writer that shares its output file with other processes sizes its records from these
 * before the first part arrives (rr_sim_fastq_bytes). */
int rr_collect_meta(rr_ctx* ctx, const uint64_t** lengths, const double** real_identity);

/* number of kernel launches issued by this context so far, and device milliseconds spent in the
 * last rr_run (CUDA events on the context's stream): [0] whole run, [1] read kernel (assemble +
 * glitches + error walk + alignment + stitch), [2] scan + qscore kernel (which also packs the sequences), [3] 0 (the pack step no longer has a launch of its own) */
uint64_t rr_kernel_launches(const rr_ctx* ctx);
int rr_last_run_ms(const rr_ctx* ctx, float* ms4);
/* bases emitted by the last run: sum over reads of max(0, len - 2k) for len > 14 (mod.rs:192-208) */
uint64_t rr_last_run_bases(const rr_ctx* ctx);
/* CUDA stream the kernels are launched on (cudaStream_t), for external event timing */
void* rr_stream(const rr_ctx* ctx);

/* ---- debug / parity hooks (used by tests only; they run the device code paths) ------------------------ */
/* alignment::align (src/alignment.rs:13-41) on the device: cigar must hold m + n bytes */
int rr_debug_align(rr_ctx* ctx, const uint8_t* query, uint32_t m, const uint8_t* target, uint32_t n,
                   uint8_t* cigar, uint32_t* cigar_len, uint32_t* edits);
/* first n 32-bit words of StdRng::seed_from_u64(seed), starting at word `first` */
int rr_debug_rng_words(rr_ctx* ctx, uint64_t seed, uint64_t first, uint32_t n, uint32_t* out);
/* parity hook: the walk's division (reciprocal + Newton fast path) against the compiler's IEEE `/` on n pseudo-random
 * operand pairs of the walk's ranges; *mismatches = number of results that differ in any bit (must be 0) */
int rr_debug_div_check(rr_ctx* ctx, uint64_t seed, uint64_t n, uint64_t* mismatches);
/* capacity-planning slack for the next batches (<= 0 keeps it) and the number of re-plans so far: lets a test force
 * the "estimate too small -> re-plan with twice the slack" path */
/* size of the parts of the streaming drain (default 32 MiB), so that tests can force many parts on a small batch */
int rr_debug_set_part_bytes(rr_ctx* ctx, uint64_t bytes);
int rr_debug_set_slack(rr_ctx* ctx, double slack, uint64_t* replans_so_far);
/* intermediate products of the last rr_run for read i: which = 0 raw fragment, 1 CIGAR */
int rr_debug_fetch(rr_ctx* ctx, uint64_t i, int which, uint8_t* out, uint64_t cap, uint64_t* len);

#ifdef __cplusplus
}
#endif
#endif /* RUSTYREAD_B200_H */
