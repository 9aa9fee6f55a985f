/*
 * rustyread_b200 — host side of `rustyread simulate` behind a C ABI.
 *
 * This is the C++ stand-in for the reference's Rust host code around the hot path (the Rust
 * toolchain is not available in the build image): the orchestrator of src/simulate/mod.rs:27-217
 * split at the seams a Rust caller would keep:
 *
 *   rr_sim_create   mod.rs:28-127   main_rng, Length, References (+ adjust_depth), Identity, Adapter,
 *                                   Glitch, Error and Quality models — in the reference's main_rng order
 *   rr_sim_attach   (new)           hands those immutable objects to a device context (rr_set_*)
 *   rr_sim_feed     mod.rs:163-170  the serial Fragments iterator for one batch (src/simulate/fragments.rs)
 *   rr_sim_fastq    mod.rs:191-212  the serial FASTQ writer: len <= 14 filter, UUIDv3 names from main_rng,
 *                                   Description text (src/simulate/description.rs:59-109), k-base trimming
 *
 * Everything here is O(1) work per read; the per-base work is in rustyread_b200.h.
 */
#ifndef RUSTYREAD_B200_HOST_H
#define RUSTYREAD_B200_HOST_H

#include "rustyread_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rr_sim rr_sim;

/* cli::simulate::Command (src/cli/simulate.rs:162-239), already parsed; percentages as given on the CLI */
typedef struct rr_sim_params {
  uint64_t seed;                                  /* --seed */
  double length_mean, length_stdev;               /* --length 15000,13000 */
  double identity_mean, identity_max, identity_stdev; /* --identity 85,95,5 */
  const char* start_adapter_seq;                  /* --start_adapter_seq */
  const char* end_adapter_seq;                    /* --end_adapter_seq */
  double start_adapter_rate, start_adapter_amount; /* --start_adapter 90,60 */
  double end_adapter_rate, end_adapter_amount;     /* --end_adapter 50,20 */
  double junk_reads, random_reads, chimera;       /* --junk_reads 1 --random_reads 1 --chimera 1 */
  double glitch_rate, glitch_size, glitch_skip;   /* --glitches 10000,25,25 */
  int small_plasmid_bias;                         /* --small_plasmid_bias */
  const char* error_model;                        /* NULL or "random" = Error::random(7); else the model TEXT */
  uint64_t error_model_len;
  const char* qscore_model;                       /* "random", "ideal", or the model TEXT */
  uint64_t qscore_model_len;
} rr_sim_params;

void rr_sim_default_params(rr_sim_params* p); /* the CLI defaults, error/qscore model = random */

int rr_sim_create(const rr_sim_params* p, const char* fasta, uint64_t fasta_len, rr_sim** out);
void rr_sim_destroy(rr_sim* s);
const char* rr_sim_last_error(const rr_sim* s); /* s may be NULL: error of a failed rr_sim_create */

uint64_t rr_sim_genome_length(const rr_sim* s); /* sum of contig lengths (mod.rs:129-133) */
uint32_t rr_sim_k(const rr_sim* s);
/* Quantity::from_str + number_of_base (src/cli/simulate.rs:12-80): "50", "50K", "50M", "50G", "50x"; -1 on error */
int64_t rr_quantity_to_bases(const char* text, uint64_t genome_length);

/* The same session without the bases: contig ids, lengths, circular flags and header weights (depth= x length) as
 * rr_sim_contig_index reports them for a session made from FASTA text.  main_rng is drawn from in the same order
 * (adjust_depth needs lengths only), so the feeder and the writer of such a session are indistinguishable from the
 * original's.  For the ranks of a multi-GPU run that receive the reference device to device (rr_sim_attach_device)
 * instead of parsing the FASTA again. */
int rr_sim_create_indexed(const rr_sim_params* p, uint64_t n, const char* const* ids, const uint64_t* lengths, const uint8_t* circular,
                          const double* weights, rr_sim** out);
int rr_sim_contig_index(const rr_sim* s, uint64_t i, const char** id, uint64_t* length, int* circular, double* weight);

/* References + the four models -> device context */
int rr_sim_attach(const rr_sim* s, rr_ctx* ctx);
/* ... with the concatenated bases of all contigs already in device memory of the context's GPU (rr_set_references_device) */
int rr_sim_attach_device(const rr_sim* s, rr_ctx* ctx, const void* device_bases);

/* One batch of the Fragments iterator for `base_loop` bases (mod.rs:153-170).  The returned array is owned by
 * the session and valid until the next rr_sim_feed. */
int rr_sim_feed(rr_sim* s, uint64_t base_loop, const rr_desc** descs, uint64_t* n);

/* FASTQ text of `batch`, which must hold the results of items [first, first + batch->n_reads) of the last
 * rr_sim_feed, shards being passed in increasing `first` order (read names are drawn from main_rng in write
 * order).  The text is owned by the session and valid until the next call. */
int rr_sim_fastq(rr_sim* s, const rr_batch* batch, uint64_t first, const char** text, uint64_t* len,
                 uint64_t* n_records, uint64_t* n_bases);

/* The same records written straight to file descriptor `fd` with writev(2): headers come from a small scratch
 * buffer, sequence and quality bytes are sent from the batch's (pinned) arenas without an intermediate copy. */
int rr_sim_write_fastq(rr_sim* s, const rr_batch* batch, uint64_t first, int fd, uint64_t* n_records, uint64_t* n_bases,
                       uint64_t* n_bytes);
/* threads > 1: on a regular file opened read-write (not for append) rr_sim_write_fastq grows the file to the end of the part,
 * maps the part's byte range and copies contiguous ranges of records into the mapping from `threads` threads, then leaves the
 * file position after the part (one writev stream fills the page cache at 2-4 GB/s and several streams into one file do no
 * better; the parallel copy does).  Only for a file no other process resizes meanwhile.  Default 1: one writev stream, which is
 * also what pipes, append-mode descriptors and small parts always get.  The bytes written are the same either way. */
int rr_sim_set_writer_threads(rr_sim* s, int threads);

/* Bytes, records and bases of the text rr_sim_fastq / rr_sim_write_fastq produce for items [first, first + n) of the last
 * feed, given their final lengths and identities; formats the comments only and draws nothing from main_rng (a read
 * name is always 36 characters).  Lets several writers place their records in one file without a gather. */
int rr_sim_fastq_bytes(rr_sim* s, uint64_t n, const uint64_t* lengths, const double* real_identity, uint32_t k, uint64_t first,
                       uint64_t* n_bytes, uint64_t* n_records, uint64_t* n_bases);

/* Advance main_rng past the read-name draws of `n_records` emitted reads (2 u64 each, mod.rs:201) without
 * formatting anything: what the ranks that do not write the file call to stay in step with the writer rank. */
int rr_sim_skip_names(rr_sim* s, uint64_t n_records);

/* Description::to_string of item i of the last feed, before generation (origin/chimera/length/identity) */
int rr_sim_item_description(rr_sim* s, uint64_t i, char* out, uint64_t cap);

/* introspection for parity tests: the loaded tables exactly as handed to the device */
uint64_t rr_sim_n_refs(const rr_sim* s);
const uint8_t* rr_sim_ref(const rr_sim* s, uint64_t i, uint64_t* len, int* circular, double* weight, const char** id);
/* main_rng: number of 32-bit words consumed so far (to check the reference's draw order) */
uint64_t rr_sim_rng_words(const rr_sim* s);

#ifdef __cplusplus
}
#endif
#endif
