// rustyread_b200 — context, device tables, launch plumbing and the C ABI (include/rustyread_b200.h).
// There is no CPU fallback anywhere in this file: without a usable CUDA device rr_create fails.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <numeric>
#include <string>
#include <vector>

#include "rr_qscore_kernel.cuh"
#include "rr_read_kernel.cuh"

using namespace rr;

#define RR_API extern "C" __attribute__((visibility("default")))

static constexpr size_t READ_KERNEL_SMEM = (size_t)WARPS_PER_BLOCK * sizeof(WarpSmem);
static_assert(READ_KERNEL_SMEM * RR_BLOCKS_PER_SM + 1024 * RR_BLOCKS_PER_SM <= 228 * 1024, "RR_BLOCKS_PER_SM blocks must fit the SM's shared memory");

static std::string g_create_error;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  bool ensure(size_t bytes) {
    if (bytes <= cap) return true;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); return false; }
    cap = want;
    return true;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return (T*)p; }
};
struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  bool ensure(size_t bytes) {
    if (bytes <= cap) return true;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return false; }
    cap = want;
    return true;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return (T*)p; }
};

struct rr_ctx {
  int device = 0;
  int n_sm = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  uint64_t launches = 0;

  // models
  DeviceModels M{};
  QscoreModelDev Q{};
  bool have_refs = false, have_error = false, have_qscore = false, have_adapter = false, have_glitch = false;
  std::vector<uint64_t> ref_len;
  std::vector<uint8_t> ref_circ;
  double glitch_mean_dist = 0, glitch_mean_size = 0;
  DevBuf d_ref, d_ref_off, d_ref_circ;
  DevBuf d_em_index, d_em_alt_begin, d_em_alt_off, d_em_alt_bytes, d_em_cum, d_em_scale;
  DevBuf d_em_piece_begin, d_em_piece_off, d_em_piece_edits, d_em_piece_ops;
  DevBuf d_em_kblob, d_em_arec;
  DevBuf d_adapter;
  DevBuf d_lcs_tab;
  DevBuf d_q_tab_off, d_q_tab, d_q_dist_off, d_q_scores, d_q_cum, d_q_scale, d_q_unit, d_q_guide, d_q_tab32, d_q_rec;

  // Up to two batches in flight (rr_submit / rr_collect): all per-batch state, arenas and scratch included, lives in a
  // slot; the D2H of batch i (copy stream) overlaps the kernels of batch i+1 (the other slot's compute stream).
  struct Slot {
    uint64_t n = 0;
    std::vector<rr_desc> descs;
    std::vector<uint64_t> h_raw_off, h_out_off;
    std::vector<uint32_t> h_order;
    double slack = 1.5;
    uint64_t sc_grec_cap = 0, sc_gseq_cap = 0;
    uint32_t read_grid = 0;
    uint32_t n_heavy = 0, n_fast_sms = 0, fast_stride = 1;  // fast lanes of the read kernel (ReadKernelParams)
    DevBuf d_descs, d_order, d_queue, d_raw_off, d_out_off;
    DevBuf d_raw_len, d_err_len, d_cig_len, d_q_cursor, d_real_id, d_status, d_packed_off, d_totals;
    DevBuf d_seq, d_qual;
    // every slot has its own compute stream, working arenas and per-warp scratch: the kernels of two batches in flight may
    // share the GPU (the persistent blocks of the next batch's read kernel move in as the tail of this one's retire)
    cudaStream_t stream = nullptr;
    DevBuf d_raw, d_err, d_cig;
    DevBuf d_sc_rows, d_sc_ops, d_sc_grec, d_sc_gseq;
    uint64_t sc_grec_alloc = 0, sc_gseq_alloc = 0;  // per-warp capacities the scratch is allocated for
    // Heaviest read's work over a warp's share of the batch's work (plan()).  Above g_overlap_ratio (2: the chain of the heaviest read is then about 1.5 x the balanced kernel time) the batch ends in a long
    // tail (a few warps walking the heaviest reads, the rest of the GPU idle), and the next batch's read kernel is allowed to
    // start behind it; otherwise the next batch waits, so that this batch's qscore kernel is not locked out by the next read
    // kernel's persistent blocks (measured on the bench workload: 18.2 -> 14.3 Gbases/s host to host when they overlap).
    double tail_ratio = 0;
    bool done_recorded = false;
    PinBuf h_seq, h_qual, h_off, h_len, h_id, h_totals, h_status;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t done = nullptr;
    bool prepared = false, enqueued = false, ran = false;
    float ms[4] = {0, 0, 0, 0};
    uint64_t bases = 0, bytes = 0;
  };
  Slot slot[2];
  uint64_t n_submitted = 0, n_collected = 0;
  // streaming drain (rr_collect_begin / rr_collect_part): a ring of small pinned buffers instead of whole-batch staging
  struct Part {
    PinBuf seq, qual;
    std::vector<uint64_t> off;  // record offsets relative to the part
    cudaEvent_t ev = nullptr;
    uint64_t first = 0, last = 0;
  };
  static constexpr int N_PARTS = 3;
  uint64_t part_bytes = 32ull << 20;  // rr_debug_set_part_bytes
  Part part[N_PARTS];
  bool draining = false;
  uint64_t drain_queued = 0, drain_returned = 0, drain_next_read = 0;  // parts queued / handed out; first read not yet queued
  int last_slot = 0;  // slot of the most recently completed run (debug fetch, timings)
  cudaStream_t copy_stream = nullptr;
  double initial_slack = 1.5;  // capacity planning slack (rr_debug_set_slack)
  uint64_t n_replans = 0;      // how many times a batch had to be re-planned with more slack
  DevBuf d_sc_rows, d_sc_ops;  // scratch of rr_debug_align
  uint64_t sc_rows_cap = 32768, sc_ops_cap = 16384;  // per-warp alignment scratch (words / bytes)
};

// RR_TRACE=1: host-side timeline of the batch pipeline on stderr (developer aid, off by default)
static const bool g_trace = std::getenv("RR_TRACE") != nullptr;
static const uint32_t g_fast_warps = std::getenv("RR_FAST_WARPS") ? (uint32_t)std::atoi(std::getenv("RR_FAST_WARPS")) : (uint32_t)std::max(1, WARPS_PER_BLOCK / 2);
static const double g_fast_threshold = std::getenv("RR_FAST_THRESHOLD") ? std::atof(std::getenv("RR_FAST_THRESHOLD")) : 0.7;
static const bool g_no_fast_lanes = std::getenv("RR_NO_FAST_LANES") != nullptr;  // developer aid: A/B of the read kernel's fast lanes
static const double g_overlap_ratio = std::getenv("RR_OVERLAP_RATIO") ? std::atof(std::getenv("RR_OVERLAP_RATIO")) : 2.0;  // see Slot::tail_ratio
static void trace(const char* what, uint64_t batch) {
  if (!g_trace) return;
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  std::fprintf(stderr, "[rr %.3f ms] batch %llu %s\n", ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6, (unsigned long long)batch, what);
}

static int fail(rr_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg; else g_create_error = msg;
  return code;
}
#define CK(call)                                                                               \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      cudaGetLastError();                                                                      \
      return fail(ctx, RR_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));       \
    }                                                                                          \
  } while (0)
#define ENSURE(buf, bytes) \
  do { if (!(buf).ensure(bytes)) return fail(ctx, RR_ERR_CUDA, "out of memory allocating " #buf); } while (0)

template <class T>
static bool upload(rr_ctx* ctx, DevBuf& b, const std::vector<T>& v) {
  size_t bytes = std::max<size_t>(1, v.size()) * sizeof(T);
  if (!b.ensure(bytes)) return false;
  if (!v.empty() && cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return false;
  return cudaStreamSynchronize(ctx->stream) == cudaSuccess;
}

RR_API int rr_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// ---- table of small alignments (rr_device.cuh, LCS_MAX_SUM) ----------------------------------------------------------
// One thread per (query, target) pair of a shape: alignment::align (src/alignment.rs:13-41) = LCS recurrence + the traceback
// preference diag >= left && diag >= up -> diag, else up >= left -> I, else D (the same rule as host_align / the DP kernels).
__global__ void __launch_bounds__(256) rr_lcs_table_kernel(uint32_t* __restrict__ tab, uint32_t m, uint32_t n) {
  const uint32_t count = 1u << (2u * (m + n));
  for (uint32_t key = blockIdx.x * blockDim.x + threadIdx.x; key < count; key += gridDim.x * blockDim.x) {
    const uint32_t tcode = key & ((1u << (2u * n)) - 1u), qcode = key >> (2u * n);
    uint8_t S[LCS_MAX_M + 1][LCS_MAX_N + 1];
    for (uint32_t j = 0; j <= n; ++j) S[0][j] = 0;
    for (uint32_t i = 1; i <= m; ++i) {
      S[i][0] = 0;
      const uint32_t qi = (qcode >> (2u * (i - 1u))) & 3u;
      for (uint32_t j = 1; j <= n; ++j) {
        const uint32_t eq = qi == ((tcode >> (2u * (j - 1u))) & 3u) ? 1u : 0u;
        S[i][j] = (uint8_t)max((uint32_t)S[i - 1][j - 1] + eq, max((uint32_t)S[i - 1][j], (uint32_t)S[i][j - 1]));
      }
    }
    uint32_t i = m, j = n, cnt = 0, rev = 0;  // ops in traceback order, 2 bits each
    while (i > 0 || j > 0) {
      uint32_t op;
      if (i == 0) { op = 3; --j; }
      else if (j == 0) { op = 2; --i; }
      else {
        const bool eq = ((qcode >> (2u * (i - 1u))) & 3u) == ((tcode >> (2u * (j - 1u))) & 3u);
        const uint32_t dg = (uint32_t)S[i - 1][j - 1] + (eq ? 1u : 0u), up = S[i - 1][j], left = S[i][j - 1];
        if (dg >= left && dg >= up) { op = eq ? 0u : 1u; --i; --j; }
        else if (up >= left) { op = 2; --i; }
        else { op = 3; --j; }
      }
      rev |= op << (2u * cnt);
      ++cnt;
    }
    uint32_t fwd = 0;
    for (uint32_t x = 0; x < cnt; ++x) fwd |= ((rev >> (2u * (cnt - 1u - x))) & 3u) << (2u * x);
    tab[key] = fwd | (cnt << 28);
  }
}

static int build_lcs_table(rr_ctx* ctx) {
  std::vector<uint32_t> base(LCS_MAX_M * LCS_MAX_N, LCS_NO_TABLE);
  uint64_t total = 0;
  // RR_LCS_MIN (developer aid): smallest m + n that goes through the table; smaller windows keep the bit-vector DP
  static const uint32_t min_sum = std::getenv("RR_LCS_MIN") ? (uint32_t)std::atoi(std::getenv("RR_LCS_MIN")) : LCS_MIN_SUM;
  // shapes in order of m + n: the small windows, which take most of the lookups (76 % within 10 bases), sit together in
  // the first 33 MiB of the table
  for (uint32_t sum = std::max(2u, min_sum); sum <= LCS_MAX_SUM; ++sum) {
    for (uint32_t m = 1; m <= LCS_MAX_M; ++m) {
      const uint32_t n = sum - m;
      if (n < 1 || n > LCS_MAX_N || m >= sum) continue;
      base[(m - 1) * LCS_MAX_N + (n - 1)] = (uint32_t)total;
      total += 1ull << (2 * sum);
    }
  }
  if (total >= 0xffffffffull) return fail(ctx, RR_ERR_UNSUPPORTED, "alignment table: too many entries for 32-bit offsets");
  ENSURE(ctx->d_lcs_tab, std::max<uint64_t>(total, 1) * 4);
  CK(cudaMemcpyToSymbol(c_lcs_base, base.data(), base.size() * sizeof(uint32_t)));
  for (uint32_t m = 1; m <= LCS_MAX_M; ++m)
    for (uint32_t n = 1; n <= LCS_MAX_N; ++n) {
      const uint32_t b = base[(m - 1) * LCS_MAX_N + (n - 1)];
      if (b == LCS_NO_TABLE) continue;
      const uint64_t count = 1ull << (2 * (m + n));
      const uint32_t grid = (uint32_t)std::min<uint64_t>((count + 255) / 256, (uint64_t)ctx->n_sm * 32);
      rr_lcs_table_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->d_lcs_tab.as<uint32_t>() + b, m, n);
      CK(cudaGetLastError());
    }
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->M.lcs_tab = ctx->d_lcs_tab.as<uint32_t>();
  return RR_OK;
}

RR_API int rr_create(int device, rr_ctx** out) {
  rr_ctx* ctx = nullptr;
  if (!out) return fail(nullptr, RR_ERR_ARG, "rr_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(nullptr, RR_ERR_CUDA, "rr_create: no CUDA device available (this library has no CPU fallback)");
  }
  if (device < 0 || device >= n) return fail(nullptr, RR_ERR_ARG, "rr_create: device index out of range");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail(nullptr, RR_ERR_CUDA, "rr_create: kernels are built for sm_100a only");
  // the read kernel's per-block shared memory (8 x WarpSmem) is dynamic: opt in past 48 KB, and keep the carve-out large
  // enough for RR_BLOCKS_PER_SM resident blocks
  for (auto* kern : {rr_read_kernel<EM_RANDOM, 7>, rr_read_kernel<EM_PACKED, 7>, rr_read_kernel<EM_TABLES, 7>,
                     rr_read_kernel<EM_RANDOM, 0>, rr_read_kernel<EM_PACKED, 0>, rr_read_kernel<EM_TABLES, 0>}) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)READ_KERNEL_SMEM));
    // RR_SMEM_CARVEOUT (developer A/B): percent of the SM's L1 / shared memory array preferred as shared memory
    static const int carveout = std::getenv("RR_SMEM_CARVEOUT") ? std::atoi(std::getenv("RR_SMEM_CARVEOUT")) : (int)cudaSharedmemCarveoutMaxShared;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carveout));
  }
  ctx = new rr_ctx();
  ctx->device = device;
  ctx->n_sm = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return fail(nullptr, RR_ERR_CUDA, "cudaStreamCreate failed"); }
  for (auto& sl : ctx->slot) {
    for (auto& ev : sl.ev) cudaEventCreate(&ev);
    cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming);
  }
  ctx->slot[0].stream = ctx->stream;  // rr_upload / rr_run (and rr_stream) work on slot 0
  if (cudaStreamCreateWithFlags(&ctx->slot[1].stream, cudaStreamNonBlocking) != cudaSuccess) { rr_destroy(ctx); return fail(nullptr, RR_ERR_CUDA, "cudaStreamCreate failed"); }
  if (build_lcs_table(ctx) != RR_OK) {
    g_create_error = "rr_create: " + ctx->err;
    rr_destroy(ctx);
    return RR_ERR_CUDA;
  }
  *out = ctx;
  return RR_OK;
}

RR_API void rr_destroy(rr_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->slot[1].stream) cudaStreamSynchronize(ctx->slot[1].stream);
  cudaStreamSynchronize(ctx->copy_stream);
  DevBuf* bufs[] = {&ctx->d_ref, &ctx->d_ref_off, &ctx->d_ref_circ, &ctx->d_em_index, &ctx->d_em_alt_begin, &ctx->d_em_alt_off,
                    &ctx->d_em_alt_bytes, &ctx->d_em_cum, &ctx->d_em_scale, &ctx->d_em_piece_begin, &ctx->d_em_piece_off, &ctx->d_em_piece_edits,
                    &ctx->d_em_piece_ops, &ctx->d_em_kblob, &ctx->d_em_arec, &ctx->d_adapter, &ctx->d_lcs_tab, &ctx->d_q_tab_off, &ctx->d_q_tab,
                    &ctx->d_q_dist_off, &ctx->d_q_scores, &ctx->d_q_cum, &ctx->d_q_scale, &ctx->d_q_unit, &ctx->d_q_guide, &ctx->d_q_tab32, &ctx->d_q_rec,
                    &ctx->d_sc_rows, &ctx->d_sc_ops};
  for (auto* b : bufs) b->release();
  for (auto& sl : ctx->slot) {
    DevBuf* sb[] = {&sl.d_descs, &sl.d_order, &sl.d_queue, &sl.d_raw_off, &sl.d_out_off, &sl.d_raw_len, &sl.d_err_len, &sl.d_cig_len,
                    &sl.d_q_cursor, &sl.d_real_id, &sl.d_status, &sl.d_packed_off, &sl.d_totals, &sl.d_seq, &sl.d_qual,
                    &sl.d_raw, &sl.d_err, &sl.d_cig, &sl.d_sc_rows, &sl.d_sc_ops, &sl.d_sc_grec, &sl.d_sc_gseq};
    for (auto* b : sb) b->release();
    PinBuf* pins[] = {&sl.h_seq, &sl.h_qual, &sl.h_off, &sl.h_len, &sl.h_id, &sl.h_totals, &sl.h_status};
    for (auto* b : pins) b->release();
    for (auto& ev : sl.ev) if (ev) cudaEventDestroy(ev);
    if (sl.done) cudaEventDestroy(sl.done);
  }
  for (auto& pt : ctx->part) {
    pt.seq.release();
    pt.qual.release();
    if (pt.ev) cudaEventDestroy(pt.ev);
  }
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->slot[1].stream) cudaStreamDestroy(ctx->slot[1].stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  delete ctx;
}

RR_API const char* rr_last_error(const rr_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
RR_API uint64_t rr_kernel_launches(const rr_ctx* ctx) { return ctx ? ctx->launches : 0; }
RR_API uint64_t rr_last_run_bases(const rr_ctx* ctx) { return ctx ? ctx->slot[ctx->last_slot].bases : 0; }
RR_API void* rr_stream(const rr_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
RR_API int rr_last_run_ms(const rr_ctx* ctx, float* ms4) {
  if (!ctx || !ms4) return RR_ERR_ARG;
  for (int i = 0; i < 4; ++i) ms4[i] = ctx->slot[ctx->last_slot].ms[i];
  return RR_OK;
}

// ---- references -------------------------------------------------------------------------------------
RR_API int rr_set_references(rr_ctx* ctx, uint64_t n, const uint8_t* const* seq, const uint64_t* len, const uint8_t* circular) {
  if (!ctx) return RR_ERR_ARG;
  if (n == 0 || !seq || !len) return fail(ctx, RR_ERR_ARG, "rr_set_references: no contigs");
  if (n > 0xffffffffULL) return fail(ctx, RR_ERR_UNSUPPORTED, "rr_set_references: too many contigs");
  CK(cudaSetDevice(ctx->device));
  std::vector<uint64_t> off(n + 1, 0);
  for (uint64_t i = 0; i < n; ++i) off[i + 1] = off[i] + len[i];
  ENSURE(ctx->d_ref, std::max<uint64_t>(16, off[n]));
  for (uint64_t i = 0; i < n; ++i)
    if (len[i]) CK(cudaMemcpyAsync(ctx->d_ref.as<uint8_t>() + off[i], seq[i], len[i], cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->ref_len.assign(len, len + n);
  ctx->ref_circ.assign(n, 0);
  if (circular) for (uint64_t i = 0; i < n; ++i) ctx->ref_circ[i] = circular[i] ? 1 : 0;
  if (!upload(ctx, ctx->d_ref_off, off) || !upload(ctx, ctx->d_ref_circ, ctx->ref_circ)) return fail(ctx, RR_ERR_CUDA, "rr_set_references: upload failed");
  ctx->M.ref_bytes = ctx->d_ref.as<uint8_t>();
  ctx->M.ref_off = ctx->d_ref_off.as<uint64_t>();
  ctx->M.ref_circ = ctx->d_ref_circ.as<uint8_t>();
  ctx->M.n_refs = (uint32_t)n;
  ctx->have_refs = true;
  return RR_OK;
}

RR_API int rr_set_references_device(rr_ctx* ctx, uint64_t n, const uint64_t* len, const uint8_t* circular, const void* device_bases) {
  if (!ctx) return RR_ERR_ARG;
  if (n == 0 || !len || !device_bases) return fail(ctx, RR_ERR_ARG, "rr_set_references_device: no contigs");
  if (n > 0xffffffffULL) return fail(ctx, RR_ERR_UNSUPPORTED, "rr_set_references_device: too many contigs");
  CK(cudaSetDevice(ctx->device));
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, device_bases) != cudaSuccess || (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged) ||
      at.device != ctx->device) {
    cudaGetLastError();
    return fail(ctx, RR_ERR_ARG, "rr_set_references_device: not device memory of this context's GPU");
  }
  std::vector<uint64_t> off(n + 1, 0);
  for (uint64_t i = 0; i < n; ++i) off[i + 1] = off[i] + len[i];
  ENSURE(ctx->d_ref, std::max<uint64_t>(16, off[n]));
  if (off[n]) CK(cudaMemcpyAsync(ctx->d_ref.p, device_bases, off[n], cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->ref_len.assign(len, len + n);
  ctx->ref_circ.assign(n, 0);
  if (circular) for (uint64_t i = 0; i < n; ++i) ctx->ref_circ[i] = circular[i] ? 1 : 0;
  if (!upload(ctx, ctx->d_ref_off, off) || !upload(ctx, ctx->d_ref_circ, ctx->ref_circ)) return fail(ctx, RR_ERR_CUDA, "rr_set_references_device: upload failed");
  ctx->M.ref_bytes = ctx->d_ref.as<uint8_t>();
  ctx->M.ref_off = ctx->d_ref_off.as<uint64_t>();
  ctx->M.ref_circ = ctx->d_ref_circ.as<uint8_t>();
  ctx->M.n_refs = (uint32_t)n;
  ctx->have_refs = true;
  return RR_OK;
}

RR_API int rr_references_device(const rr_ctx* ctx, const void** device_bases, uint64_t* bytes) {
  if (!ctx || !device_bases || !bytes) return RR_ERR_ARG;
  if (!ctx->have_refs) return RR_ERR_ARG;
  uint64_t total = 0;
  for (uint64_t l : ctx->ref_len) total += l;
  *device_bases = ctx->d_ref.p;
  *bytes = total;
  return RR_OK;
}

// alignment::align on the host (src/alignment.rs:13-41: bio's global aligner with match 1, everything else 0 = LCS with
// the traceback preference diag >= left && diag >= up -> diag, else up >= left -> I, else D).  Used once per model
// alternative at load time; the kernels use bit-vector forms of the same rule.
static uint32_t host_align(const uint8_t* q, uint32_t m, const uint8_t* t, uint32_t n, std::vector<uint8_t>& ops) {
  std::vector<uint16_t> S((size_t)(m + 1) * (n + 1), 0);
  const size_t W = n + 1;
  for (uint32_t i = 1; i <= m; ++i)
    for (uint32_t j = 1; j <= n; ++j)
      S[i * W + j] = std::max<uint16_t>((uint16_t)(S[(i - 1) * W + j - 1] + (q[i - 1] == t[j - 1] ? 1 : 0)), std::max(S[(i - 1) * W + j], S[i * W + j - 1]));
  std::vector<uint8_t> rev;
  uint32_t i = m, j = n, edits = 0;
  while (i > 0 || j > 0) {
    uint8_t op;
    if (i == 0) { op = 'D'; --j; }
    else if (j == 0) { op = 'I'; --i; }
    else {
      const bool eq = q[i - 1] == t[j - 1];
      const int dg = S[(i - 1) * W + j - 1] + (eq ? 1 : 0), up = S[(i - 1) * W + j], left = S[i * W + j - 1];
      if (dg >= left && dg >= up) { op = eq ? '=' : 'X'; --i; --j; }
      else if (up >= left) { op = 'I'; --i; }
      else { op = 'D'; --j; }
    }
    edits += op != '=';
    rev.push_back(op);
  }
  ops.insert(ops.end(), rev.rbegin(), rev.rend());
  return edits;
}

// ---- error model ---------------------------------------------------------------------------------------
RR_API int rr_set_error_model(rr_ctx* ctx, uint32_t k, uint64_t n_keys, const uint8_t* keys, const uint32_t* alt_begin,
                              const uint32_t* alt_off, const uint8_t* alt_bytes, const double* weights) {
  if (!ctx) return RR_ERR_ARG;
  if (k == 0 || k > 31) return fail(ctx, RR_ERR_UNSUPPORTED, "rr_set_error_model: k must be in 1..31");
  CK(cudaSetDevice(ctx->device));
  ctx->M.k = k;
  ctx->M.em_alts_acgt = 1;
  if (n_keys == 0) {  // Error::random(k)
    ctx->M.em_index = nullptr;
    ctx->M.em_kblob = nullptr;
    ctx->have_error = true;
    return RR_OK;
  }
  if (!keys || !alt_begin || !alt_off || !alt_bytes || !weights) return fail(ctx, RR_ERR_ARG, "rr_set_error_model: NULL table");
  if (k > 12) return fail(ctx, RR_ERR_UNSUPPORTED, "rr_set_error_model: file models need k <= 12 (dense 4^k index)");
  const uint64_t n_codes = 1ULL << (2 * k);
  std::vector<int32_t> index(n_codes, -1);
  const uint32_t n_alts = alt_begin[n_keys];
  std::vector<double> cum(n_alts), scale(n_keys);
  for (uint64_t i = 0; i < n_keys; ++i) {
    uint32_t code = 0;
    for (uint32_t x = 0; x < k; ++x) {
      uint8_t c = keys[i * k + x];
      uint32_t s = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4;
      if (s > 3) return fail(ctx, RR_ERR_UNSUPPORTED, "rr_set_error_model: key with a byte outside ACGT");
      code |= s << (2 * (k - 1 - x));
    }
    index[code] = (int32_t)i;  // later duplicates replace earlier ones, like HashMap::insert
    uint32_t a0 = alt_begin[i], a1 = alt_begin[i + 1];
    if (a1 <= a0) return fail(ctx, RR_ERR_ARG, "rr_set_error_model: key without alternatives (the reference panics when it is sampled)");
    if (!weighted_index_build(weights + a0, a1 - a0, cum.data() + a0, &scale[i]))
      return fail(ctx, RR_ERR_ARG, "rr_set_error_model: invalid weights (WeightedIndex::new fails in the reference)");
    for (uint32_t j = a0; j < a1; ++j)
      if (alt_off[j + 1] - alt_off[j] > (uint32_t)ALT_MAX) return fail(ctx, RR_ERR_UNSUPPORTED, "rr_set_error_model: alternative longer than 64 bytes");
  }
  std::vector<uint32_t> ab(alt_begin, alt_begin + n_keys + 1), ao(alt_off, alt_off + n_alts + 1);
  std::vector<uint8_t> bytes(alt_bytes, alt_bytes + alt_off[n_alts]);
  for (uint8_t c : bytes)
    if (!(c == 'A' || c == 'C' || c == 'G' || c == 'T')) ctx->M.em_alts_acgt = 0;
  // precomputed alignments: alternative j (of key i) against its key, for every piece offset es < len(alt), es <= k
  std::vector<uint32_t> piece_begin(n_alts + 1, 0), piece_off(1, 0);
  std::vector<uint8_t> piece_edits, piece_ops;
  for (uint64_t i = 0; i < n_keys; ++i) {
    const uint8_t* key = keys + i * k;
    for (uint32_t j = alt_begin[i]; j < alt_begin[i + 1]; ++j) {
      const uint8_t* alt = alt_bytes + alt_off[j];
      const uint32_t al = alt_off[j + 1] - alt_off[j];
      piece_begin[j] = (uint32_t)piece_edits.size();
      for (uint32_t es = 0; es < std::max(1u, al) && es <= k; ++es) {
        const uint32_t e = host_align(alt + std::min(es, al), al - std::min(es, al), key + es, k - es, piece_ops);
        piece_edits.push_back((uint8_t)std::min(e, 255u));
        piece_off.push_back((uint32_t)piece_ops.size());
      }
    }
  }
  piece_begin[n_alts] = (uint32_t)piece_edits.size();
  if (!upload(ctx, ctx->d_em_index, index) || !upload(ctx, ctx->d_em_alt_begin, ab) || !upload(ctx, ctx->d_em_alt_off, ao) ||
      !upload(ctx, ctx->d_em_alt_bytes, bytes) || !upload(ctx, ctx->d_em_cum, cum) || !upload(ctx, ctx->d_em_scale, scale) ||
      !upload(ctx, ctx->d_em_piece_begin, piece_begin) || !upload(ctx, ctx->d_em_piece_off, piece_off) ||
      !upload(ctx, ctx->d_em_piece_edits, piece_edits) || !upload(ctx, ctx->d_em_piece_ops, piece_ops))
    return fail(ctx, RR_ERR_CUDA, "rr_set_error_model: upload failed");
  // packed records (DeviceModels::em_kblob): same content, two dependent loads per event
  ctx->M.em_kblob = nullptr;
  ctx->M.em_arec = nullptr;
  ctx->M.em_kstride = 0;
  bool packable = k <= EM_PACK_K;
  for (uint32_t j = 0; j < n_alts && packable; ++j) packable = alt_off[j + 1] - alt_off[j] <= EM_PACK_ALT;
  if (packable) {
    uint32_t max_na = 0;
    for (uint64_t i = 0; i < n_keys; ++i) max_na = std::max(max_na, alt_begin[i + 1] - alt_begin[i]);
    const uint32_t stride = 16 + 8 * std::max(32u, (max_na + 1u) & ~1u);  // every lane may read one sum past the header
    if ((uint64_t)stride * n_codes > (1ull << 31)) packable = false;      // a model with thousands of alternatives per k-mer: plain tables
    if (packable) {
      std::vector<uint32_t> arec((size_t)n_alts * EM_AREC_WORDS, 0);
      std::vector<uint8_t> blob((size_t)stride * n_codes, 0);
      for (uint64_t code = 0; code < n_codes; ++code) {
        const int32_t i = index[code];
        if (i < 0) continue;  // alternatives = 0: not in the model
        const uint32_t a0 = alt_begin[i], na = alt_begin[i + 1] - a0;
        uint8_t* at = &blob[(size_t)stride * code];
        std::memcpy(at, &na, 4);
        std::memcpy(at + 4, &a0, 4);
        std::memcpy(at + 8, &scale[i], 8);
        std::memcpy(at + 16, &cum[a0], (size_t)na * 8);
      }
      for (uint32_t j = 0; j < n_alts; ++j) {
        uint32_t* r = &arec[(size_t)j * EM_AREC_WORDS];
        uint8_t* rb = reinterpret_cast<uint8_t*>(r);
        const uint32_t al = alt_off[j + 1] - alt_off[j];
        std::memcpy(rb, alt_bytes + alt_off[j], al);
        rb[15] = (uint8_t)al;
        uint16_t* meta = reinterpret_cast<uint16_t*>(r + 12);
        for (uint32_t es = 0; es < 8; ++es) {
          const uint32_t pi = piece_begin[j] + es;
          if (pi >= piece_begin[j + 1]) break;  // piece offsets that cannot occur for this alternative
          const uint32_t o0 = piece_off[pi], no = piece_off[pi + 1] - o0;
          if (no > 16) continue;  // entry stays invalid: the kernel aligns such a piece like any other merge
          uint32_t ops = 0;
          for (uint32_t x = 0; x < no; ++x) {
            const uint8_t op = piece_ops[o0 + x];
            ops |= (op == '=' ? 0u : op == 'X' ? 1u : op == 'I' ? 2u : 3u) << (2 * x);
          }
          r[4 + es] = ops;
          meta[es] = (uint16_t)(no | ((uint32_t)piece_edits[pi] << 5) | 0x8000u);
        }
      }
      if (!upload(ctx, ctx->d_em_kblob, blob) || !upload(ctx, ctx->d_em_arec, arec)) return fail(ctx, RR_ERR_CUDA, "rr_set_error_model: upload failed");
      if (!std::getenv("RR_NO_PACKED_MODEL")) {  // developer A/B: the unpacked tables
        ctx->M.em_kblob = ctx->d_em_kblob.as<uint8_t>();
        ctx->M.em_kstride = stride;
        ctx->M.em_arec = ctx->d_em_arec.as<uint32_t>();
      }
    }
  }
  ctx->M.em_index = ctx->d_em_index.as<int32_t>();
  ctx->M.em_alt_begin = ctx->d_em_alt_begin.as<uint32_t>();
  ctx->M.em_alt_off = ctx->d_em_alt_off.as<uint32_t>();
  ctx->M.em_alt_bytes = ctx->d_em_alt_bytes.as<uint8_t>();
  ctx->M.em_cum = ctx->d_em_cum.as<double>();
  ctx->M.em_scale = ctx->d_em_scale.as<double>();
  ctx->M.em_piece_begin = ctx->d_em_piece_begin.as<uint32_t>();
  ctx->M.em_piece_off = ctx->d_em_piece_off.as<uint32_t>();
  ctx->M.em_piece_edits = ctx->d_em_piece_edits.as<uint8_t>();
  ctx->M.em_piece_ops = ctx->d_em_piece_ops.as<uint8_t>();
  ctx->have_error = true;
  return RR_OK;
}

// ---- qscore model ------------------------------------------------------------------------------------------
RR_API int rr_set_qscore_model(rr_ctx* ctx, uint32_t max_k, uint64_t n, const uint8_t* cigars, const uint32_t* cigar_off,
                               const uint32_t* score_off, const uint8_t* scores, const double* weights) {
  if (!ctx) return RR_ERR_ARG;
  if (n == 0 || !cigars || !cigar_off || !score_off || !scores || !weights) return fail(ctx, RR_ERR_ARG, "rr_set_qscore_model: NULL table");
  if (max_k == 0) return fail(ctx, RR_ERR_ARG, "rr_set_qscore_model: max_k == 0");
  if (max_k > 12) return fail(ctx, RR_ERR_UNSUPPORTED, "rr_set_qscore_model: max_k > 12 (dense 4^l window tables)");
  if (n > 65000) return fail(ctx, RR_ERR_UNSUPPORTED, "rr_set_qscore_model: more than 65000 entries");
  CK(cudaSetDevice(ctx->device));
  std::map<std::string, uint32_t> key2dist;
  for (uint64_t i = 0; i < n; ++i) key2dist[std::string((const char*)cigars + cigar_off[i], cigar_off[i + 1] - cigar_off[i])] = (uint32_t)i;
  if (!key2dist.count("=") || !key2dist.count("X") || !key2dist.count("I"))
    return fail(ctx, RR_ERR_ARG, "rr_set_qscore_model: model must contain '=', 'X' and 'I' (QualityNotMinimalCigarString)");
  std::vector<uint32_t> dist_off(score_off, score_off + n + 1);
  const uint32_t n_sc = score_off[n];
  std::vector<double> cum(std::max<uint32_t>(1, n_sc)), scale(n);
  std::vector<char> dist_ok(n, 1);
  std::vector<uint8_t> unit(n, 0);
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t a0 = score_off[i], a1 = score_off[i + 1];
    if (a1 <= a0 || !weighted_index_build(weights + a0, a1 - a0, cum.data() + a0, &scale[i])) { dist_ok[i] = 0; continue; }
    bool u = (a1 - a0) < (1u << 20);
    for (uint32_t j = a0; j < a1 && u; ++j) u = weights[j] == 1.0;
    unit[i] = u ? 1 : 0;  // partial sums are then exactly 1, 2, 3, ... (Quality::random / ideal)
  }
  const uint32_t margin = (max_k - 1) / 2;
  std::vector<uint32_t> tab_off(2 * margin + 2, 0);
  std::vector<uint16_t> tab;
  static const char SYM[4] = {'=', 'X', 'I', 'D'};
  for (uint32_t l = 1; l <= 2 * margin + 1; l += 2) {
    tab_off[l] = (uint32_t)tab.size();
    const uint64_t cnt = 1ULL << (2 * l);
    std::string w(l, '=');
    for (uint64_t code = 0; code < cnt; ++code) {
      for (uint32_t y = 0; y < l; ++y) w[y] = SYM[(code >> (2 * y)) & 3];
      // get_qscore: strip one char from both ends until the window is known (quality.rs:118-130)
      uint16_t id = 0xFFFF;
      size_t b = 0, len = l;
      for (;;) {
        auto it = key2dist.find(w.substr(b, len));
        if (it != key2dist.end()) { id = (uint16_t)it->second; break; }
        if (len <= 2) break;
        b += 1;
        len -= 2;
      }
      if (id != 0xFFFF && !dist_ok[id]) {
        // WeightedIndex::new(..).unwrap() would panic when this entry is sampled; only fatal if reachable
        if (w[l / 2] != 'D') return fail(ctx, RR_ERR_ARG, "rr_set_qscore_model: entry '" + w.substr(b, len) + "' has invalid weights");
        id = 0xFFFF;
      }
      tab.push_back(id);
    }
  }
  std::vector<uint8_t> sc(scores, scores + n_sc);
  std::vector<QDist> qd(n);
  for (uint64_t i = 0; i < n; ++i) {
    qd[i].d0 = score_off[i];
    qd[i].nd_unit = (score_off[i + 1] - score_off[i]) | (unit[i] ? 0x80000000u : 0u);
    qd[i].scale = dist_ok[i] ? scale[i] : 0.0;
  }
  // bounds of the per-draw search (QscoreModelDev::guide): the products are formed exactly as the kernel forms `chosen`
  std::vector<uint16_t> guide(n * (Q_GUIDE + 1), 0);
  for (uint64_t i = 0; i < n; ++i) {
    const uint32_t a0 = score_off[i], nd = score_off[i + 1] - score_off[i];
    if (!dist_ok[i] || nd == 0) continue;
    if (nd > 65535) return fail(ctx, RR_ERR_UNSUPPORTED, "rr_set_qscore_model: a distribution with more than 65535 scores");
    for (uint32_t g = 0; g < Q_GUIDE; ++g) {
      const double x = ((double)g / (double)Q_GUIDE) * scale[i] + 0.0;
      uint32_t c = 0;
      while (c + 1 < nd && cum[a0 + c] <= x) ++c;  // partial sums are non-decreasing
      guide[i * (Q_GUIDE + 1) + g] = (uint16_t)c;
    }
    guide[i * (Q_GUIDE + 1) + Q_GUIDE] = (uint16_t)(nd - 1);
  }
  // one record per distribution + the window table pointing straight at the records (QscoreModelDev::qrec)
  std::vector<uint32_t> rec_off(n, Q_NO_REC), tab32(tab.size(), Q_NO_REC);
  std::vector<uint8_t> qrec;
  for (uint64_t i = 0; i < n; ++i) {
    const uint32_t a0 = score_off[i], nd = score_off[i + 1] - score_off[i];
    if (!dist_ok[i] || nd == 0) continue;
    rec_off[i] = (uint32_t)(qrec.size() / 16);
    const size_t at = qrec.size();
    qrec.resize(at + 16 * (size_t)(Q_REC_HEAD + nd), 0);
    std::memcpy(&qrec[at], &scale[i], 8);
    std::memcpy(&qrec[at + 8], &qd[i].nd_unit, 4);
    std::memcpy(&qrec[at + 16], &guide[i * (Q_GUIDE + 1)], (Q_GUIDE + 1) * 2);
    for (uint32_t j = 0; j < nd; ++j) {
      const uint64_t sc8 = scores[a0 + j];
      std::memcpy(&qrec[at + 16 * (Q_REC_HEAD + j)], &cum[a0 + j], 8);
      std::memcpy(&qrec[at + 16 * (Q_REC_HEAD + j) + 8], &sc8, 8);
    }
  }
  if (qrec.size() / 16 >= Q_NO_REC) return fail(ctx, RR_ERR_UNSUPPORTED, "rr_set_qscore_model: model too large");
  qrec.resize(qrec.size() + 64, 0);
  for (size_t x = 0; x < tab.size(); ++x) if (tab[x] != 0xFFFF) tab32[x] = rec_off[tab[x]];
  if (!upload(ctx, ctx->d_q_tab32, tab32) || !upload(ctx, ctx->d_q_rec, qrec)) return fail(ctx, RR_ERR_CUDA, "rr_set_qscore_model: upload failed");
  ctx->Q.tab32 = ctx->d_q_tab32.as<uint32_t>();
  ctx->Q.qrec = ctx->d_q_rec.as<uint8_t>();
  if (!upload(ctx, ctx->d_q_tab_off, tab_off) || !upload(ctx, ctx->d_q_tab, tab) || !upload(ctx, ctx->d_q_dist_off, qd) ||
      !upload(ctx, ctx->d_q_scores, sc) || !upload(ctx, ctx->d_q_cum, cum) || !upload(ctx, ctx->d_q_guide, guide))
    return fail(ctx, RR_ERR_CUDA, "rr_set_qscore_model: upload failed");
  ctx->Q.max_k = max_k;
  ctx->Q.n_dist = (uint32_t)n;
  ctx->Q.n_scores = n_sc;
  ctx->Q.tab_off = ctx->d_q_tab_off.as<uint32_t>();
  ctx->Q.tab = ctx->d_q_tab.as<uint16_t>();
  ctx->Q.dist = ctx->d_q_dist_off.as<QDist>();
  ctx->Q.scores = ctx->d_q_scores.as<uint8_t>();
  ctx->Q.cum = ctx->d_q_cum.as<double>();
  ctx->Q.guide = ctx->d_q_guide.as<uint16_t>();
  ctx->have_qscore = true;
  return RR_OK;
}

// ---- adapter / glitch -----------------------------------------------------------------------------------------
RR_API int rr_set_adapter(rr_ctx* ctx, const uint8_t* start, uint64_t start_len, const uint8_t* end, uint64_t end_len,
                          double start_rate, double start_amount, double end_rate, double end_amount) {
  if (!ctx) return RR_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  // Adapter::new, src/model/adapter.rs:23-61 (including the double division by 100 at :34-38)
  DeviceModels& M = ctx->M;
  M.has_start = M.has_end = 0;
  if (!(start_len == 0 || start_rate == 0.0 || start_amount == 0.0)) {
    double a = start_amount / 100.0;
    if (!beta_new(2.0 * a, 2.0 - 2.0 * (a / 100.0), M.start_beta)) return fail(ctx, RR_ERR_ARG, "rr_set_adapter: invalid start Beta parameters");
    M.has_start = 1;
  }
  if (!(end_len == 0 || end_rate == 0.0 || end_amount == 0.0)) {
    double a = end_amount / 100.0;
    if (!beta_new(2.0 * a, 2.0 - 2.0 * a, M.end_beta)) return fail(ctx, RR_ERR_ARG, "rr_set_adapter: invalid end Beta parameters");
    M.has_end = 1;
  }
  M.start_rate = start_rate / 100.0;
  M.end_rate = end_rate / 100.0;
  // gen_bool panics outside [0,1] (rand Bernoulli::new)
  if (!(M.start_rate >= 0.0 && M.start_rate <= 1.0) || !(M.end_rate >= 0.0 && M.end_rate <= 1.0))
    return fail(ctx, RR_ERR_ARG, "rr_set_adapter: rate outside 0..100 (gen_bool panics in the reference)");
  std::vector<uint8_t> bytes;
  if (start_len) bytes.insert(bytes.end(), start, start + start_len);
  if (end_len) bytes.insert(bytes.end(), end, end + end_len);
  if (!upload(ctx, ctx->d_adapter, bytes)) return fail(ctx, RR_ERR_CUDA, "rr_set_adapter: upload failed");
  M.adapter_bytes = ctx->d_adapter.as<uint8_t>();
  M.start_len = (uint32_t)start_len;
  M.end_len = (uint32_t)end_len;
  ctx->have_adapter = true;
  return RR_OK;
}

RR_API int rr_set_glitch(rr_ctx* ctx, double rate, double size, double skip) {
  if (!ctx) return RR_ERR_ARG;
  DeviceModels& M = ctx->M;
  // Glitch::new, src/model/glitch.rs:20-48
  M.has_distance = 0;
  if (rate != 0.0) {
    if (!geometric_new(rate > 1.0 ? 1.0 / rate : 1.0, M.g_dist)) return fail(ctx, RR_ERR_ARG, "rr_set_glitch: invalid rate");
    M.has_distance = 1;
  }
  if (!geometric_new(size > 1.0 ? 1.0 / size : 1.0, M.g_size)) return fail(ctx, RR_ERR_ARG, "rr_set_glitch: invalid size");
  if (!geometric_new(skip > 1.0 ? 1.0 / skip : 1.0, M.g_skip)) return fail(ctx, RR_ERR_ARG, "rr_set_glitch: invalid skip");
  ctx->glitch_mean_dist = M.has_distance ? (1.0 - M.g_dist.p) / M.g_dist.p : 0.0;
  ctx->glitch_mean_size = (1.0 - M.g_size.p) / M.g_size.p;
  ctx->have_glitch = true;
  return RR_OK;
}

// ---- batch planning ------------------------------------------------------------------------------------------
static inline uint64_t align16(uint64_t x) { return (x + 15) & ~15ULL; }
static inline uint64_t align128(uint64_t x) { return (x + 127) & ~127ULL; }
typedef rr_ctx::Slot Slot;

static int frag_len(rr_ctx* ctx, uint32_t type, uint32_t ref_idx, uint64_t start, uint64_t end, uint64_t& out) {
  if (type == RR_JUNK || type == RR_RANDOM) { out = end; return RR_OK; }
  if (type != RR_REAL) return fail(ctx, RR_ERR_ARG, "rr_upload: unknown read type");
  if (ref_idx >= ctx->ref_len.size()) return fail(ctx, RR_ERR_ARG, "rr_upload: ref_idx out of range");
  uint64_t rl = ctx->ref_len[ref_idx];
  if (start < end) {
    if (end > rl) return fail(ctx, RR_ERR_ARG, "rr_upload: fragment end beyond the contig (the reference panics)");
    out = end - start;
  } else if (ctx->ref_circ[ref_idx]) {
    if (start > rl || end > rl) return fail(ctx, RR_ERR_ARG, "rr_upload: fragment beyond the contig (the reference panics)");
    out = rl - start + end;
  } else out = 0;
  return RR_OK;
}

static int plan(rr_ctx* ctx, Slot& sl) {
  const uint64_t n = sl.n;
  const DeviceModels& M = ctx->M;
  const uint64_t adapters = (uint64_t)M.start_len + M.end_len;
  sl.h_raw_off.assign(n + 1, 0);
  sl.h_out_off.assign(n + 1, 0);
  sl.h_order.resize(n);
  // Processing order = longest estimated WORK first.  The walk costs about 400 warp instructions per error event and
  // about 5 per base (profiles/), and the events of a read are a serial chain, so the read with the most events (not
  // the most bases) is the critical path of a batch and has to start first.
  std::vector<double> work(n);
  double max_ng = 0;
  for (uint64_t i = 0; i < n; ++i) {
    const rr_desc& d = sl.descs[i];
    uint64_t f1 = 0, f2 = 0;
    int rc = frag_len(ctx, d.type, d.ref_idx, d.start, d.end, f1);
    if (rc) return rc;
    if (d.has_chimera && (rc = frag_len(ctx, d.ctype, d.cref_idx, d.cstart, d.cend, f2))) return rc;
    if (!(d.identity <= 1.0 && d.identity >= -1.0)) return fail(ctx, RR_ERR_ARG, "rr_upload: identity outside [0,1]");
    uint64_t rcap = 2ULL * M.k + f1 + f2 + 2 * adapters;
    if (rcap >= (1ULL << 30)) return fail(ctx, RR_ERR_UNSUPPORTED, "rr_upload: a single read of 2^30 bases or more");
    double target = std::max(0.0, (1.0 - d.identity)) * (double)rcap;
    work[i] = 80.0 * target + (double)rcap;
    double ng = 0, gl = 0;
    if (M.has_distance) {
      ng = (double)rcap / (ctx->glitch_mean_dist + 1.0);
      ng = (ng + 6.0 * std::sqrt(ng) + 30.0) * sl.slack;
      gl = ng * (ctx->glitch_mean_size + 1.0);
      max_ng = std::max(max_ng, ng);
    }
    // err / CIGAR leave the kernel in CHUNK-byte (256) chunks and the capacity test keeps two chunks of margin (flush_chunks)
    uint64_t ocap = align128(rcap + (uint64_t)std::ceil(target) + (uint64_t)gl + (uint64_t)(64 * sl.slack) + ALT_MAX + 16 + 3 * CHUNK);
    sl.h_raw_off[i + 1] = sl.h_raw_off[i] + align16(rcap + 16);
    sl.h_out_off[i + 1] = sl.h_out_off[i] + ocap;
  }
  std::iota(sl.h_order.begin(), sl.h_order.end(), 0u);
  std::stable_sort(sl.h_order.begin(), sl.h_order.end(), [&](uint32_t a, uint32_t b) { return work[a] > work[b]; });
  // Fast lanes: reads that hold more than 0.7 of a warp's share of the batch's work would set the kernel's length at the
  // issue rate of a fully occupied scheduler; enough half-occupancy SMs take them to run each at about twice that rate.
  {
    const double n_warps = (double)ctx->n_sm * RR_BLOCKS_PER_SM * WARPS_PER_BLOCK;
    double total = 0;
    for (uint64_t i = 0; i < n; ++i) total += work[i];
    const double share = total / n_warps;
    sl.tail_ratio = (n && share > 0) ? work[sl.h_order[0]] / share : 0.0;  // heaviest read against a warp's share of the batch
    if (g_trace) std::fprintf(stderr, "[rr] plan: %llu reads, heaviest read / warp share = %.2f (next batch may overlap above %.2f)\n",
                              (unsigned long long)n, sl.tail_ratio, g_overlap_ratio);
    uint64_t nh = 0;
    double heavy = 0;
    while (nh < n && work[sl.h_order[nh]] > g_fast_threshold * share) heavy += work[sl.h_order[nh++]];
    const uint32_t max_fast = (uint32_t)std::max(1, ctx->n_sm / 6);
    uint32_t nf = nh ? (uint32_t)std::min<double>(max_fast, std::ceil(heavy / (0.4 * (WARPS_PER_BLOCK / (double)g_fast_warps) * share) / (RR_BLOCKS_PER_SM * g_fast_warps))) : 0;
    if (g_no_fast_lanes || n < (uint64_t)n_warps) nf = 0;  // fewer reads than warps: the batch is latency-bound whatever is done
    sl.n_fast_sms = nf;
    sl.n_heavy = nf ? (uint32_t)nh : 0;
    sl.fast_stride = nf ? (uint32_t)std::max(1, ctx->n_sm / (int)nf) : 1;
  }
  sl.sc_grec_cap = (uint64_t)max_ng + 8;
  sl.sc_gseq_cap = align16((uint64_t)(max_ng * (ctx->glitch_mean_size + 1.0)) + 64);
  return RR_OK;
}

// H2D of the plan + (re)allocation of everything the kernels of this batch touch (all of it owned by the slot, which is
// idle here: its previous batch has been collected).
static int upload_plan(rr_ctx* ctx, Slot& sl) {
  const uint64_t n = sl.n;
  uint32_t want_blocks = (uint32_t)std::min<uint64_t>((uint64_t)ctx->n_sm * RR_BLOCKS_PER_SM, (n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
  sl.read_grid = std::max<uint32_t>(1, want_blocks);  // persistent grid: 4 blocks of 8 warps per SM, fewer if the batch is small
  const uint64_t n_warps = (uint64_t)ctx->n_sm * RR_BLOCKS_PER_SM * WARPS_PER_BLOCK;
  sl.sc_grec_alloc = std::max(sl.sc_grec_alloc, sl.sc_grec_cap);
  sl.sc_gseq_alloc = std::max(sl.sc_gseq_alloc, sl.sc_gseq_cap);
  ENSURE(sl.d_raw, sl.h_raw_off[n] + 16);
  ENSURE(sl.d_err, sl.h_out_off[n] + 16);
  ENSURE(sl.d_cig, sl.h_out_off[n] + 16);
  ENSURE(sl.d_sc_rows, n_warps * ctx->sc_rows_cap * 8);
  ENSURE(sl.d_sc_ops, n_warps * ctx->sc_ops_cap);
  ENSURE(sl.d_sc_grec, n_warps * sl.sc_grec_alloc * sizeof(GlitchRec));
  ENSURE(sl.d_sc_gseq, n_warps * sl.sc_gseq_alloc);
  ENSURE(sl.d_raw_off, (n + 1) * 8);
  ENSURE(sl.d_out_off, (n + 1) * 8);
  ENSURE(sl.d_order, n * 4);
  // worst-case packed size (capacities), so that no host round trip is needed between the kernels of a batch
  ENSURE(sl.d_seq, sl.h_out_off[n] + 16);
  ENSURE(sl.d_qual, sl.h_out_off[n] + 16);
  CK(cudaMemcpyAsync(sl.d_raw_off.p, sl.h_raw_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, sl.stream));
  CK(cudaMemcpyAsync(sl.d_out_off.p, sl.h_out_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, sl.stream));
  CK(cudaMemcpyAsync(sl.d_order.p, sl.h_order.data(), n * 4, cudaMemcpyHostToDevice, sl.stream));
  return RR_OK;
}

// plan + H2D of one batch into a slot (no kernels yet)
static int prepare(rr_ctx* ctx, Slot& sl, uint64_t n, const rr_desc* descs) {
  if (!(ctx->have_refs && ctx->have_error && ctx->have_qscore && ctx->have_adapter && ctx->have_glitch))
    return fail(ctx, RR_ERR_ARG, "rr_upload: references and all four models must be set first");
  if (n == 0 || !descs) return fail(ctx, RR_ERR_ARG, "rr_upload: empty batch");
  if (n > 0x7fffffffULL) return fail(ctx, RR_ERR_UNSUPPORTED, "rr_upload: more than 2^31 reads in one batch");
  CK(cudaSetDevice(ctx->device));
  sl.prepared = sl.enqueued = sl.ran = false;
  sl.n = n;
  sl.descs.assign(descs, descs + n);
  sl.slack = ctx->initial_slack;
  int rc = plan(ctx, sl);
  if (rc) return rc;
  ENSURE(sl.d_descs, n * sizeof(rr_desc));
  ENSURE(sl.d_queue, 16);
  ENSURE(sl.d_raw_len, n * 8);
  ENSURE(sl.d_err_len, n * 8);
  ENSURE(sl.d_cig_len, n * 8);
  ENSURE(sl.d_q_cursor, n * 8);
  ENSURE(sl.d_real_id, n * 8);
  ENSURE(sl.d_status, n * 4);
  ENSURE(sl.d_packed_off, (n + 1) * 8);
  ENSURE(sl.d_totals, 32);
  ENSURE(sl.h_totals, 32);
  ENSURE(sl.h_status, n * 4);
  CK(cudaMemcpyAsync(sl.d_descs.p, sl.descs.data(), n * sizeof(rr_desc), cudaMemcpyHostToDevice, sl.stream));
  rc = upload_plan(ctx, sl);
  if (rc) return rc;
  sl.prepared = true;
  return RR_OK;
}

// queue the three kernels of a batch (read, scan, qscore + pack) on the compute stream; returns without waiting
static int enqueue(rr_ctx* ctx, Slot& sl) {
  const uint32_t n = (uint32_t)sl.n;
  ReadKernelParams P{};
  P.M = ctx->M;
  P.descs = sl.d_descs.as<rr_desc>();
  P.order = sl.d_order.as<uint32_t>();
  P.n_reads = n;
  P.queue = sl.d_queue.as<uint32_t>();
  P.n_heavy = sl.n_heavy;
  P.n_fast_sms = sl.n_fast_sms;
  P.fast_stride = sl.fast_stride;
  P.fast_warps = g_fast_warps;
  P.raw_off = sl.d_raw_off.as<uint64_t>();
  P.out_off = sl.d_out_off.as<uint64_t>();
  P.raw = sl.d_raw.as<uint8_t>();
  P.err = sl.d_err.as<uint8_t>();
  P.cig = sl.d_cig.as<uint8_t>();
  P.raw_len = sl.d_raw_len.as<uint64_t>();
  P.err_len = sl.d_err_len.as<uint64_t>();
  P.cig_len = sl.d_cig_len.as<uint64_t>();
  P.q_cursor = sl.d_q_cursor.as<uint64_t>();
  P.real_id = sl.d_real_id.as<double>();
  P.status = sl.d_status.as<uint32_t>();
  P.sc_rows = sl.d_sc_rows.as<uint64_t>();  P.sc_rows_cap = ctx->sc_rows_cap;
  P.sc_ops = sl.d_sc_ops.as<uint8_t>();     P.sc_ops_cap = ctx->sc_ops_cap;
  P.sc_grec = sl.d_sc_grec.as<GlitchRec>(); P.sc_grec_cap = sl.sc_grec_alloc;
  P.sc_gseq = sl.d_sc_gseq.as<uint8_t>();   P.sc_gseq_cap = sl.sc_gseq_alloc;

  {  // order against the other slot's batch (see Slot::tail_ratio)
    Slot& other = ctx->slot[1 - (int)(&sl - ctx->slot)];
    if (other.done_recorded && !(other.tail_ratio > g_overlap_ratio)) CK(cudaStreamWaitEvent(sl.stream, other.done, 0));
  }
  CK(cudaMemsetAsync(sl.d_queue.p, 0, 16, sl.stream));
  CK(cudaEventRecord(sl.ev[0], sl.stream));
  const bool k7 = ctx->M.k == 7;
  auto* read_kernel = !ctx->M.em_index ? (k7 ? rr_read_kernel<EM_RANDOM, 7> : rr_read_kernel<EM_RANDOM, 0>)
                      : ctx->M.em_kblob ? (k7 ? rr_read_kernel<EM_PACKED, 7> : rr_read_kernel<EM_PACKED, 0>)
                                       : (k7 ? rr_read_kernel<EM_TABLES, 7> : rr_read_kernel<EM_TABLES, 0>);
  read_kernel<<<sl.read_grid, WARPS_PER_BLOCK * 32, READ_KERNEL_SMEM, sl.stream>>>(P);
  CK(cudaGetLastError());
  CK(cudaEventRecord(sl.ev[1], sl.stream));
  rr_scan_kernel<<<1, 1024, 0, sl.stream>>>(sl.d_err_len.as<uint64_t>(), sl.d_status.as<uint32_t>(), n, ctx->M.k,
                                              sl.d_packed_off.as<uint64_t>(), sl.d_totals.as<uint64_t>());
  CK(cudaGetLastError());
  QscoreKernelParams QP{};
  QP.Q = ctx->Q;
  QP.descs = P.descs;
  QP.order = P.order;
  QP.n_reads = n;
  QP.queue = sl.d_queue.as<uint32_t>() + 1;
  QP.out_off = P.out_off;
  QP.cig = P.cig;
  QP.cig_len = P.cig_len;
  QP.err_len = P.err_len;
  QP.q_cursor = P.q_cursor;
  QP.packed_off = sl.d_packed_off.as<uint64_t>();
  QP.qual = sl.d_qual.as<uint8_t>();
  QP.status = P.status;
  QP.err = P.err;
  QP.seq = sl.d_seq.as<uint8_t>();
  uint32_t qgrid = (uint32_t)std::min<uint64_t>((uint64_t)ctx->n_sm * (2048 / Q_THREADS), n);  // every resident slot of the SMs, dynamic queue
  const bool q_k1 = ctx->Q.max_k == 1, q_cached = ctx->Q.n_dist <= QD_CACHE && ctx->Q.n_scores <= QS_CACHE;
  if (q_k1 && q_cached) rr_qscore_kernel<true, true><<<qgrid, Q_THREADS, 0, sl.stream>>>(QP);
  else if (q_k1) rr_qscore_kernel<true, false><<<qgrid, Q_THREADS, 0, sl.stream>>>(QP);
  else if (q_cached) rr_qscore_kernel<false, true><<<qgrid, Q_THREADS, 0, sl.stream>>>(QP);
  else rr_qscore_kernel<false, false><<<qgrid, Q_THREADS, 0, sl.stream>>>(QP);
  CK(cudaGetLastError());
  CK(cudaEventRecord(sl.ev[2], sl.stream));
  CK(cudaEventRecord(sl.ev[3], sl.stream));
  CK(cudaMemcpyAsync(sl.h_totals.p, sl.d_totals.p, 24, cudaMemcpyDeviceToHost, sl.stream));
  CK(cudaMemcpyAsync(sl.h_status.p, sl.d_status.p, (size_t)n * 4, cudaMemcpyDeviceToHost, sl.stream));
  CK(cudaEventRecord(sl.done, sl.stream));
  sl.done_recorded = true;
  ctx->launches += 3;
  sl.enqueued = true;
  return RR_OK;
}

// wait for the batch; on a capacity overflow re-plan with more slack and run again (synchronously)
static int finish(rr_ctx* ctx, Slot& sl) {
  for (int attempt = 0; attempt < 12; ++attempt) {
    CK(cudaEventSynchronize(sl.done));
    float a = 0, b = 0, c = 0, w = 0;
    cudaEventElapsedTime(&w, sl.ev[0], sl.ev[3]);
    cudaEventElapsedTime(&a, sl.ev[0], sl.ev[1]);
    cudaEventElapsedTime(&b, sl.ev[1], sl.ev[2]);
    cudaEventElapsedTime(&c, sl.ev[2], sl.ev[3]);
    sl.ms[0] = w; sl.ms[1] = a; sl.ms[2] = b; sl.ms[3] = c;
    sl.bytes = sl.h_totals.as<uint64_t>()[0];
    sl.bases = sl.h_totals.as<uint64_t>()[1];
    const uint32_t* st = sl.h_status.as<uint32_t>();
    uint64_t grow = 0;
    for (uint64_t i = 0; i < sl.n; ++i) {
      if (st[i] == ST_OK) continue;
      if (st[i] == ST_OUT_OVERFLOW || st[i] == ST_GLITCH_OVERFLOW) { ++grow; continue; }
      char buf[256];
      const char* what = st[i] == ST_ALIGN_TOO_LARGE ? "alignment window beyond the bit-vector scratch (2048 columns)"
                         : st[i] == ST_REF_PANIC     ? "slice out of range in Change::merge/update_align (the reference panics here)"
                         : st[i] == ST_RNG_EXHAUSTED ? "an error event rejected more than 400 stream words in a row (internal limit; probability < 2^-130)"
                         : st[i] == ST_QUAL_WINDOW   ? "CIGAR window slice out of range in generate_quality (src/simulate/quality.rs:22-34: the reference panics here)"
                                                     : "qscore window not in the model (reference: QualityNotMinimalCigarString)";
      std::snprintf(buf, sizeof buf, "read %llu (seed %llu): %s", (unsigned long long)i, (unsigned long long)sl.descs[i].seed, what);
      sl.enqueued = false;
      return fail(ctx, RR_ERR_READ, buf);
    }
    if (grow == 0) {
      sl.ran = true;
      sl.enqueued = false;
      ctx->last_slot = (int)(&sl - ctx->slot);
      return RR_OK;
    }
    sl.slack *= 2.0;  // capacity estimate too small for `grow` reads
    ctx->n_replans += 1;
    int rc = plan(ctx, sl);
    if (rc) return rc;
    rc = upload_plan(ctx, sl);
    if (rc) return rc;
    rc = enqueue(ctx, sl);
    if (rc) return rc;
  }
  sl.enqueued = false;
  return fail(ctx, RR_ERR_READ, "rr_run: output capacity still too small after 12 attempts");
}

// D2H of the packed result on the copy stream (the batch's kernels are known to be complete)
static int fetch(rr_ctx* ctx, Slot& sl, rr_batch* out) {
  const uint64_t n = sl.n, total = sl.bytes;
  ENSURE(sl.h_off, (n + 1) * 8);
  ENSURE(sl.h_len, n * 8);
  ENSURE(sl.h_id, n * 8);
  ENSURE(sl.h_seq, total + 16);
  ENSURE(sl.h_qual, total + 16);
  cudaStream_t cs = ctx->copy_stream;
  CK(cudaMemcpyAsync(sl.h_off.p, sl.d_packed_off.p, (n + 1) * 8, cudaMemcpyDeviceToHost, cs));
  CK(cudaMemcpyAsync(sl.h_len.p, sl.d_err_len.p, n * 8, cudaMemcpyDeviceToHost, cs));
  CK(cudaMemcpyAsync(sl.h_id.p, sl.d_real_id.p, n * 8, cudaMemcpyDeviceToHost, cs));
  if (total) {
    CK(cudaMemcpyAsync(sl.h_seq.p, sl.d_seq.p, total, cudaMemcpyDeviceToHost, cs));
    CK(cudaMemcpyAsync(sl.h_qual.p, sl.d_qual.p, total, cudaMemcpyDeviceToHost, cs));
  }
  CK(cudaStreamSynchronize(cs));
  out->n_reads = n;
  out->offsets = sl.h_off.as<uint64_t>();
  out->lengths = sl.h_len.as<uint64_t>();
  out->seq = sl.h_seq.as<uint8_t>();
  out->qual = sl.h_qual.as<uint8_t>();
  out->real_identity = sl.h_id.as<double>();
  out->total_bytes = total;
  out->total_bases = sl.bases;
  out->k = ctx->M.k;
  return RR_OK;
}

static int require_idle(rr_ctx* ctx, const char* who) {
  if (ctx->n_submitted != ctx->n_collected) return fail(ctx, RR_ERR_ARG, std::string(who) + ": batches submitted with rr_submit are still in flight (rr_collect them first)");
  return RR_OK;
}

RR_API int rr_upload(rr_ctx* ctx, uint64_t n, const rr_desc* descs) {
  if (!ctx) return RR_ERR_ARG;
  int rc = require_idle(ctx, "rr_upload");
  if (rc) return rc;
  rc = prepare(ctx, ctx->slot[0], n, descs);
  if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->slot[0].stream));
  return RR_OK;
}

RR_API int rr_run(rr_ctx* ctx) {
  if (!ctx) return RR_ERR_ARG;
  int rc = require_idle(ctx, "rr_run");
  if (rc) return rc;
  Slot& sl = ctx->slot[0];
  if (!sl.prepared) return fail(ctx, RR_ERR_ARG, "rr_run: nothing uploaded");
  CK(cudaSetDevice(ctx->device));
  rc = enqueue(ctx, sl);
  if (rc) return rc;
  return finish(ctx, sl);
}

RR_API int rr_download(rr_ctx* ctx, rr_batch* out) {
  if (!ctx || !out) return RR_ERR_ARG;
  Slot& sl = ctx->slot[0];
  if (!sl.ran) return fail(ctx, RR_ERR_ARG, "rr_download: no completed rr_run");
  CK(cudaSetDevice(ctx->device));
  return fetch(ctx, sl, out);
}

RR_API int rr_submit(rr_ctx* ctx, uint64_t n, const rr_desc* descs) {
  if (!ctx) return RR_ERR_ARG;
  if (ctx->n_submitted - ctx->n_collected >= 2) return fail(ctx, RR_ERR_ARG, "rr_submit: two batches already in flight (rr_collect one first)");
  Slot& sl = ctx->slot[ctx->n_submitted & 1];
  trace("submit: begin", ctx->n_submitted);
  int rc = prepare(ctx, sl, n, descs);
  if (rc) return rc;
  trace("submit: plan + H2D queued", ctx->n_submitted);
  rc = enqueue(ctx, sl);
  if (rc) return rc;
  trace("submit: kernels queued", ctx->n_submitted);
  ctx->n_submitted += 1;
  return RR_OK;
}

RR_API int rr_collect(rr_ctx* ctx, rr_batch* out) {
  if (!ctx || !out) return RR_ERR_ARG;
  if (ctx->n_submitted == ctx->n_collected) return fail(ctx, RR_ERR_ARG, "rr_collect: nothing submitted");
  if (ctx->draining) return fail(ctx, RR_ERR_ARG, "rr_collect: a batch is being drained with rr_collect_part");
  CK(cudaSetDevice(ctx->device));
  Slot& sl = ctx->slot[ctx->n_collected & 1];
  ctx->n_collected += 1;
  trace("collect: begin", ctx->n_collected - 1);
  int rc = finish(ctx, sl);
  if (rc) return rc;
  trace("collect: kernels done", ctx->n_collected - 1);
  rc = fetch(ctx, sl, out);
  trace("collect: D2H done", ctx->n_collected - 1);
  return rc;
}

// ---- streaming drain -----------------------------------------------------------------------------------------------------
// queue the D2H of the next part of the draining batch into ring slot (drain_queued % N_PARTS)
static int queue_part(rr_ctx* ctx, Slot& sl) {
  const uint64_t* off = sl.h_off.as<uint64_t>();
  const uint64_t first = ctx->drain_next_read;
  uint64_t last = first + 1;
  while (last < sl.n && off[last + 1] - off[first] <= ctx->part_bytes) ++last;
  rr_ctx::Part& pt = ctx->part[ctx->drain_queued % rr_ctx::N_PARTS];
  const uint64_t bytes = off[last] - off[first];
  ENSURE(pt.seq, std::max<uint64_t>(bytes, ctx->part_bytes) + 16);
  ENSURE(pt.qual, std::max<uint64_t>(bytes, ctx->part_bytes) + 16);
  if (!pt.ev) CK(cudaEventCreateWithFlags(&pt.ev, cudaEventDisableTiming));
  if (bytes) {
    CK(cudaMemcpyAsync(pt.seq.p, sl.d_seq.as<uint8_t>() + off[first], bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
    CK(cudaMemcpyAsync(pt.qual.p, sl.d_qual.as<uint8_t>() + off[first], bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
  }
  CK(cudaEventRecord(pt.ev, ctx->copy_stream));
  pt.first = first;
  pt.last = last;
  pt.off.resize(last - first + 1);
  for (uint64_t i = first; i <= last; ++i) pt.off[i - first] = off[i] - off[first];
  ctx->drain_next_read = last;
  ctx->drain_queued += 1;
  return RR_OK;
}

RR_API int rr_collect_begin(rr_ctx* ctx, uint64_t* n_reads, uint64_t* total_bases) {
  if (!ctx) return RR_ERR_ARG;
  if (ctx->n_submitted == ctx->n_collected) return fail(ctx, RR_ERR_ARG, "rr_collect_begin: nothing submitted");
  if (ctx->draining) return fail(ctx, RR_ERR_ARG, "rr_collect_begin: the previous batch is still being drained (call rr_collect_part until it returns no reads)");
  CK(cudaSetDevice(ctx->device));
  Slot& sl = ctx->slot[ctx->n_collected & 1];
  int rc = finish(ctx, sl);
  if (rc) { ctx->n_collected += 1; return rc; }
  const uint64_t n = sl.n;
  ENSURE(sl.h_off, (n + 1) * 8);
  ENSURE(sl.h_len, n * 8);
  ENSURE(sl.h_id, n * 8);
  cudaStream_t cs = ctx->copy_stream;
  CK(cudaMemcpyAsync(sl.h_off.p, sl.d_packed_off.p, (n + 1) * 8, cudaMemcpyDeviceToHost, cs));
  CK(cudaMemcpyAsync(sl.h_len.p, sl.d_err_len.p, n * 8, cudaMemcpyDeviceToHost, cs));
  CK(cudaMemcpyAsync(sl.h_id.p, sl.d_real_id.p, n * 8, cudaMemcpyDeviceToHost, cs));
  CK(cudaStreamSynchronize(cs));
  ctx->draining = true;
  ctx->drain_queued = ctx->drain_returned = 0;
  ctx->drain_next_read = 0;
  for (int i = 0; i < rr_ctx::N_PARTS && ctx->drain_next_read < n; ++i)
    if ((rc = queue_part(ctx, sl))) return rc;
  if (n_reads) *n_reads = n;
  if (total_bases) *total_bases = sl.bases;
  return RR_OK;
}

RR_API int rr_collect_meta(rr_ctx* ctx, const uint64_t** lengths, const double** real_identity) {
  if (!ctx) return RR_ERR_ARG;
  if (!ctx->draining) return fail(ctx, RR_ERR_ARG, "rr_collect_meta: no batch is being drained (rr_collect_begin first)");
  const Slot& sl = ctx->slot[ctx->n_collected & 1];
  if (lengths) *lengths = sl.h_len.as<uint64_t>();
  if (real_identity) *real_identity = sl.h_id.as<double>();
  return RR_OK;
}

RR_API int rr_collect_part(rr_ctx* ctx, rr_batch* out, uint64_t* first) {
  if (!ctx || !out) return RR_ERR_ARG;
  if (!ctx->draining) return fail(ctx, RR_ERR_ARG, "rr_collect_part: no batch is being drained (rr_collect_begin first)");
  CK(cudaSetDevice(ctx->device));
  Slot& sl = ctx->slot[ctx->n_collected & 1];
  // the part handed out by the previous call is finished with: its ring slot takes the next pending part
  if (ctx->drain_returned > 0 && ctx->drain_next_read < sl.n) {
    int rc = queue_part(ctx, sl);
    if (rc) return rc;
  }
  std::memset(out, 0, sizeof *out);
  out->k = ctx->M.k;
  if (ctx->drain_returned == ctx->drain_queued) {  // everything handed out: the batch is collected
    ctx->draining = false;
    ctx->n_collected += 1;
    if (first) *first = sl.n;
    return RR_OK;
  }
  rr_ctx::Part& pt = ctx->part[ctx->drain_returned % rr_ctx::N_PARTS];
  CK(cudaEventSynchronize(pt.ev));
  ctx->drain_returned += 1;
  const uint64_t* len = sl.h_len.as<uint64_t>();
  uint64_t bases = 0;
  for (uint64_t i = pt.first; i < pt.last; ++i)
    if (len[i] > 14 && len[i] > 2ull * ctx->M.k) bases += len[i] - 2ull * ctx->M.k;  // writer: mod.rs:192-208
  out->n_reads = pt.last - pt.first;
  out->offsets = pt.off.data();
  out->lengths = len + pt.first;
  out->seq = pt.seq.as<uint8_t>();
  out->qual = pt.qual.as<uint8_t>();
  out->real_identity = sl.h_id.as<double>() + pt.first;
  out->total_bytes = pt.off.back();
  out->total_bases = bases;
  if (first) *first = pt.first;
  return RR_OK;
}

RR_API int rr_generate(rr_ctx* ctx, uint64_t n, const rr_desc* descs, rr_batch* out) {
  if (!ctx) return RR_ERR_ARG;
  int rc = require_idle(ctx, "rr_generate");
  if (rc) return rc;
  rc = rr_submit(ctx, n, descs);
  if (rc) return rc;
  return rr_collect(ctx, out);
}

// ---- debug / parity hooks ---------------------------------------------------------------------------------------
__global__ void rr_debug_align_kernel(ReadKernelParams P, const uint8_t* q, uint32_t m, const uint8_t* t, uint32_t n, uint8_t* cigar,
                                      uint32_t* out) {
  __shared__ WarpSmem ws;
  uint64_t a = warp_align(&ws, P, 0, q, m, t, n, cigar, threadIdx.x & 31);
  if (threadIdx.x == 0) { out[0] = align_nops(a); out[1] = align_edits(a); out[2] = align_status(a); }
}

RR_API int rr_debug_align(rr_ctx* ctx, const uint8_t* query, uint32_t m, const uint8_t* target, uint32_t n, uint8_t* cigar,
                          uint32_t* cigar_len, uint32_t* edits) {
  if (!ctx) return RR_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  DevBuf dq, dt, dc, dout;
  ENSURE(dq, m + 16); ENSURE(dt, n + 16); ENSURE(dc, (size_t)m + n + 16); ENSURE(dout, 16);
  ENSURE(ctx->d_sc_rows, ctx->sc_rows_cap * 8);
  ENSURE(ctx->d_sc_ops, ctx->sc_ops_cap);
  if (m) CK(cudaMemcpyAsync(dq.p, query, m, cudaMemcpyHostToDevice, ctx->stream));
  if (n) CK(cudaMemcpyAsync(dt.p, target, n, cudaMemcpyHostToDevice, ctx->stream));
  ReadKernelParams P{};
  P.sc_rows = ctx->d_sc_rows.as<uint64_t>(); P.sc_rows_cap = ctx->sc_rows_cap;
  P.sc_ops = ctx->d_sc_ops.as<uint8_t>();    P.sc_ops_cap = ctx->sc_ops_cap;
  rr_debug_align_kernel<<<1, 32, 0, ctx->stream>>>(P, dq.as<uint8_t>(), m, dt.as<uint8_t>(), n, dc.as<uint8_t>(), dout.as<uint32_t>());
  CK(cudaGetLastError());
  ctx->launches += 1;
  uint32_t res[3];
  CK(cudaMemcpyAsync(res, dout.p, 12, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  int rc = RR_OK;
  if (res[2] != ST_OK) rc = fail(ctx, RR_ERR_UNSUPPORTED, "rr_debug_align: window beyond the bit-vector scratch");
  else {
    if (res[0]) CK(cudaMemcpy(cigar, dc.p, res[0], cudaMemcpyDeviceToHost));
    *cigar_len = res[0];
    *edits = res[1];
  }
  dq.release(); dt.release(); dc.release(); dout.release();
  return rc;
}

__global__ void rr_debug_rng_kernel(uint64_t seed, uint64_t first, uint32_t n, uint32_t* out) {
  __shared__ WarpSmem ws;
  WarpRng rng;
  rng.seed(seed, &ws, threadIdx.x & 31);
  rng.seek(first);
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t v = rng.next_u32();
    if (threadIdx.x == 0) out[i] = v;
  }
}

RR_API int rr_debug_rng_words(rr_ctx* ctx, uint64_t seed, uint64_t first, uint32_t n, uint32_t* out) {
  if (!ctx || !out) return RR_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  DevBuf d;
  ENSURE(d, (size_t)n * 4 + 16);
  rr_debug_rng_kernel<<<1, 32, 0, ctx->stream>>>(seed, first, n, d.as<uint32_t>());
  CK(cudaGetLastError());
  ctx->launches += 1;
  CK(cudaMemcpyAsync(out, d.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  d.release();
  return RR_OK;
}

// walk_div against the compiler's `/` on pseudo-random operands of the walk's ranges (plus the edges); counts bit mismatches
__global__ void rr_debug_div_kernel(uint64_t seed, uint64_t n, unsigned long long* mismatches) {
  unsigned long long bad = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t x = (i + 1) * 0x9E3779B97F4A7C15ULL + seed;
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL; x ^= x >> 27; x *= 0x94D049BB133111EBULL; x ^= x >> 31;
    const uint32_t r_i = 1u + (uint32_t)(x % 0x7fffffffULL);
    uint64_t y = x * 0xD6E8FEB86659FD93ULL + 1;
    y ^= y >> 32;
    double t;
    const uint32_t sel = (uint32_t)(y & 7);
    const double u = (double)(y >> 11) * 1.1102230246251565e-16;  // [0, 1)
    if (sel == 0) t = 0.0;
    else if (sel == 1) t = (double)r_i;                         // p == 1
    else if (sel == 2) t = u * 9.5367431640625e-07;             // tiny targets
    else if (sel == 3) t = ldexp(u + 0.5, -105 + (int)((y >> 3) % 136));
    else if (sel == 4) t = (double)r_i * u;                     // p uniform in [0, 1)
    else if (sel == 5) t = (double)(r_i / 2) + u;               // near one half
    else t = u * 2147483648.0;
    const double a = walk_div(t, (double)r_i), b = t / (double)r_i;
    bad += __double_as_longlong(a) != __double_as_longlong(b);
  }
  if (bad) atomicAdd(mismatches, bad);
}

RR_API int rr_debug_div_check(rr_ctx* ctx, uint64_t seed, uint64_t n, uint64_t* mismatches) {
  if (!ctx || !mismatches) return RR_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  DevBuf d;
  ENSURE(d, 16);
  CK(cudaMemsetAsync(d.p, 0, 16, ctx->stream));
  rr_debug_div_kernel<<<ctx->n_sm * 8, 256, 0, ctx->stream>>>(seed, n, d.as<unsigned long long>());
  CK(cudaGetLastError());
  ctx->launches += 1;
  CK(cudaMemcpyAsync(mismatches, d.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  d.release();
  return RR_OK;
}

RR_API int rr_debug_set_part_bytes(rr_ctx* ctx, uint64_t bytes) {
  if (!ctx || bytes == 0) return RR_ERR_ARG;
  if (ctx->draining) return fail(ctx, RR_ERR_ARG, "rr_debug_set_part_bytes: a batch is being drained");
  ctx->part_bytes = bytes;
  return RR_OK;
}

RR_API int rr_debug_set_slack(rr_ctx* ctx, double slack, uint64_t* replans_so_far) {
  if (!ctx) return RR_ERR_ARG;
  if (slack > 0.0) ctx->initial_slack = slack;
  if (replans_so_far) *replans_so_far = ctx->n_replans;
  return RR_OK;
}

RR_API int rr_debug_fetch(rr_ctx* ctx, uint64_t i, int which, uint8_t* out, uint64_t cap, uint64_t* len) {
  if (!ctx || !out || !len) return RR_ERR_ARG;
  const Slot& sl = ctx->slot[ctx->last_slot];
  if (!sl.ran || i >= sl.n) return fail(ctx, RR_ERR_ARG, "rr_debug_fetch: no such read");
  CK(cudaSetDevice(ctx->device));
  uint64_t l = 0;
  const DevBuf& lens = which == 0 ? sl.d_raw_len : sl.d_cig_len;
  CK(cudaMemcpy(&l, lens.as<uint64_t>() + i, 8, cudaMemcpyDeviceToHost));
  if (l > cap) return fail(ctx, RR_ERR_ARG, "rr_debug_fetch: buffer too small");
  const uint8_t* src = which == 0 ? sl.d_raw.as<uint8_t>() + sl.h_raw_off[i] : sl.d_cig.as<uint8_t>() + sl.h_out_off[i];
  if (l) CK(cudaMemcpy(out, src, l, cudaMemcpyDeviceToHost));
  *len = l;
  return RR_OK;
}
