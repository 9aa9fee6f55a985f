"""ctypes binding of include/rustyread_b200.h (the drop-in boundary for src/simulate/mod.rs:172-186)."""
import ctypes as C
import os

import numpy as np

from . import _build

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)

RR_REAL, RR_JUNK, RR_RANDOM = 0, 1, 2
RR_PLUS, RR_MINUS, RR_NOSTRAND = 0, 1, 2

# numpy mirror of struct rr_desc (80 bytes)
DESC_DTYPE = np.dtype([
    ("type", "<u4"), ("ref_idx", "<u4"), ("strand", "<u4"), ("has_chimera", "<u4"),
    ("ctype", "<u4"), ("cref_idx", "<u4"), ("cstrand", "<u4"), ("reserved", "<u4"),
    ("start", "<u8"), ("end", "<u8"), ("cstart", "<u8"), ("cend", "<u8"),
    ("identity", "<f8"), ("seed", "<u8"),
])


class rr_batch(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("offsets", u64p), ("lengths", u64p), ("seq", u8p), ("qual", u8p),
                ("real_identity", f64p), ("total_bytes", C.c_uint64), ("total_bases", C.c_uint64), ("k", C.c_uint32)]


class RRError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("rustyread_b200 error %d: %s" % (code, msg))
        self.code = code


EXPORTS = {
    "rr_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "rr_destroy": (None, [C.c_void_p]),
    "rr_last_error": (C.c_char_p, [C.c_void_p]),
    "rr_device_count": (C.c_int, []),
    "rr_set_references": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p), u64p, C.c_void_p]),
    "rr_set_references_device": (C.c_int, [C.c_void_p, C.c_uint64, u64p, C.c_void_p, C.c_void_p]),
    "rr_references_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), u64p]),
    "rr_set_error_model": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rr_set_qscore_model": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rr_set_adapter": (C.c_int, [C.c_void_p, C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_double]),
    "rr_set_glitch": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double]),
    "rr_generate": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(rr_batch)]),
    "rr_upload": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p]),
    "rr_run": (C.c_int, [C.c_void_p]),
    "rr_download": (C.c_int, [C.c_void_p, C.POINTER(rr_batch)]),
    "rr_submit": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p]),
    "rr_collect": (C.c_int, [C.c_void_p, C.POINTER(rr_batch)]),
    "rr_collect_begin": (C.c_int, [C.c_void_p, u64p, u64p]),
    "rr_collect_part": (C.c_int, [C.c_void_p, C.POINTER(rr_batch), u64p]),
    "rr_collect_meta": (C.c_int, [C.c_void_p, C.POINTER(u64p), C.POINTER(f64p)]),
    "rr_kernel_launches": (C.c_uint64, [C.c_void_p]),
    "rr_last_run_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "rr_last_run_bases": (C.c_uint64, [C.c_void_p]),
    "rr_stream": (C.c_void_p, [C.c_void_p]),
    "rr_debug_align": (C.c_int, [C.c_void_p, C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_void_p, u32p, u32p]),
    "rr_debug_rng_words": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]),
    "rr_debug_div_check": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, u64p]),
    "rr_debug_set_part_bytes": (C.c_int, [C.c_void_p, C.c_uint64]),
    "rr_debug_set_slack": (C.c_int, [C.c_void_p, C.c_double, u64p]),
    "rr_debug_fetch": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_uint64, u64p]),
}


class Library:
    """The loaded C-ABI library.  Loading never falls back to anything else."""

    def __init__(self, path=None):
        self.path = path or _build.LIB
        if not os.path.exists(self.path):
            raise RRError(1, "CUDA library %s is missing: run __graft_entry__.build() (no CPU fallback exists)" % self.path)
        self.dll = C.CDLL(self.path)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(self.dll, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args

    def device_count(self):
        return self.dll.rr_device_count()


_lib = None


def load_library(path=None):
    global _lib
    if _lib is None or (path and _lib.path != path):
        _lib = Library(path)
    return _lib


class Batch:
    """Host view of one generated batch (numpy views over the library's pinned buffers)."""

    def __init__(self, b):
        n = int(b.n_reads)
        self.struct = b
        self.n = n
        self.k = int(b.k)
        self.total_bytes = int(b.total_bytes)
        self.total_bases = int(b.total_bases)
        self.offsets = np.ctypeslib.as_array(b.offsets, shape=(n + 1,))
        self.lengths = np.ctypeslib.as_array(b.lengths, shape=(n,))
        self.real_identity = np.ctypeslib.as_array(b.real_identity, shape=(n,))
        tb = max(1, self.total_bytes)
        self.seq = np.ctypeslib.as_array(b.seq, shape=(tb,))
        self.qual = np.ctypeslib.as_array(b.qual, shape=(tb,))

    def read(self, i):
        o, l = int(self.offsets[i]), int(self.lengths[i])
        return bytes(self.seq[o:o + l]), bytes(self.qual[o:o + l])


class Context:
    """Host-side mirror of the reference objects `generate_read` borrows (src/simulate/mod.rs:223-231)."""

    def __init__(self, device=0, lib=None):
        self.lib = lib or load_library()
        h = C.c_void_p()
        rc = self.lib.dll.rr_create(device, C.byref(h))
        if rc:
            raise RRError(rc, self.lib.dll.rr_last_error(None).decode())
        self.h = h
        self.k = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.dll.rr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise RRError(rc, self.lib.dll.rr_last_error(self.h).decode())

    # -- immutable inputs ---------------------------------------------------------------
    def set_references(self, seqs, circular=None):
        """seqs: list of bytes (References::sequences[i].seq); circular: list of bool."""
        n = len(seqs)
        keep = [np.frombuffer(s, dtype=np.uint8) if len(s) else np.zeros(1, np.uint8) for s in seqs]
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in keep])
        lens = (C.c_uint64 * n)(*[len(s) for s in seqs])
        circ = (C.c_uint8 * n)(*[1 if (circular and circular[i]) else 0 for i in range(n)])
        self._ck(self.lib.dll.rr_set_references(self.h, n, ptrs, lens, circ))

    def references_device(self):
        """(device address, bytes) of the reference arena: contigs back to back (the send buffer of a broadcast)."""
        ptr, n = C.c_void_p(), C.c_uint64()
        self._ck(self.lib.dll.rr_references_device(self.h, C.byref(ptr), C.byref(n)))
        return int(ptr.value or 0), int(n.value)

    def set_error_model_random(self, k=7):
        """model::Error::random(k) — src/model/error.rs:85-90."""
        self._ck(self.lib.dll.rr_set_error_model(self.h, k, 0, None, None, None, None, None))
        self.k = k

    def set_error_model(self, k, entries):
        """entries: list of (key bytes, [(alt bytes, weight), ...]) as loaded by Error::from_stream."""
        keys = b"".join(e[0] for e in entries)
        alt_begin = np.zeros(len(entries) + 1, np.uint32)
        alts, weights = [], []
        for i, (_key, al) in enumerate(entries):
            alt_begin[i + 1] = alt_begin[i] + len(al)
            for a, w in al:
                alts.append(a)
                weights.append(w)
        alt_off = np.zeros(len(alts) + 1, np.uint32)
        np.cumsum([len(a) for a in alts], out=alt_off[1:]) if alts else None
        bytes_ = np.frombuffer(b"".join(alts) + b"\0", dtype=np.uint8)
        w = np.asarray(weights + [0.0], dtype=np.float64)
        kb = np.frombuffer(keys + b"\0", dtype=np.uint8)
        self._ck(self.lib.dll.rr_set_error_model(self.h, k, len(entries), kb.ctypes.data, alt_begin.ctypes.data,
                                                 alt_off.ctypes.data, bytes_.ctypes.data, w.ctypes.data))
        self.k = k

    def set_qscore_model(self, max_k, entries):
        """entries: list of (cigar bytes, [score...], [weight...]) as loaded by Quality::from_stream."""
        cig = b"".join(e[0] for e in entries)
        cigar_off = np.zeros(len(entries) + 1, np.uint32)
        score_off = np.zeros(len(entries) + 1, np.uint32)
        scores, weights = [], []
        for i, (c, s, w) in enumerate(entries):
            cigar_off[i + 1] = cigar_off[i] + len(c)
            score_off[i + 1] = score_off[i] + len(s)
            scores += list(s)
            weights += list(w)
        sc = np.asarray(scores + [0], dtype=np.uint8)
        w = np.asarray(weights + [0.0], dtype=np.float64)
        cb = np.frombuffer(cig + b"\0", dtype=np.uint8)
        self._ck(self.lib.dll.rr_set_qscore_model(self.h, max_k, len(entries), cb.ctypes.data, cigar_off.ctypes.data,
                                                  score_off.ctypes.data, sc.ctypes.data, w.ctypes.data))

    def set_qscore_model_random(self):
        """model::Quality::random() — src/model/quality.rs:80-91."""
        e = [(c, list(range(1, 21)), [1.0] * 20) for c in (b"=", b"X", b"I")]
        self.set_qscore_model(1, e)

    def set_qscore_model_ideal(self):
        """model::Quality::ideal() — src/model/quality.rs:94-110."""
        r = lambda a, b: (list(range(a, b + 1)), [1.0] * (b - a + 1))
        e = [(b"X",) + r(1, 3), (b"I",) + r(1, 3), (b"=",) + r(4, 7), (b"===",) + r(8, 20), (b"=====",) + r(21, 30),
             (b"=======",) + r(31, 40), (b"=========",) + r(41, 50)]
        self.set_qscore_model(9, e)

    def set_adapter(self, start, end, start_rate, start_amount, end_rate, end_amount):
        self._ck(self.lib.dll.rr_set_adapter(self.h, start, len(start), end, len(end), start_rate, start_amount, end_rate, end_amount))

    def set_glitch(self, rate, size, skip):
        self._ck(self.lib.dll.rr_set_glitch(self.h, rate, size, skip))

    # -- the hot path ------------------------------------------------------------------------
    def generate(self, descs):
        """generate_read over a batch of work items (numpy array of DESC_DTYPE); host in, host out."""
        descs = np.ascontiguousarray(descs, dtype=DESC_DTYPE)
        b = rr_batch()
        self._ck(self.lib.dll.rr_generate(self.h, len(descs), descs.ctypes.data, C.byref(b)))
        return Batch(b)

    def upload(self, descs):
        descs = np.ascontiguousarray(descs, dtype=DESC_DTYPE)
        self._ck(self.lib.dll.rr_upload(self.h, len(descs), descs.ctypes.data))

    def run(self):
        self._ck(self.lib.dll.rr_run(self.h))

    def download(self):
        b = rr_batch()
        self._ck(self.lib.dll.rr_download(self.h, C.byref(b)))
        return Batch(b)

    def submit(self, descs):
        """Asynchronous rr_generate, first half: queues H2D + kernels (at most two batches in flight)."""
        descs = np.ascontiguousarray(descs, dtype=DESC_DTYPE)
        self._ck(self.lib.dll.rr_submit(self.h, len(descs), descs.ctypes.data))

    def collect(self):
        """Second half: waits for the oldest submitted batch and drains it to pinned host memory."""
        b = rr_batch()
        self._ck(self.lib.dll.rr_collect(self.h, C.byref(b)))
        return Batch(b)

    def collect_begin(self):
        """Wait for the oldest submitted batch; returns (n_reads, total_bases, lengths, real_identity) — the two arrays are
        copies of the whole batch's metadata (rr_collect_meta).  Follow with parts()."""
        n, bases = C.c_uint64(), C.c_uint64()
        self._ck(self.lib.dll.rr_collect_begin(self.h, C.byref(n), C.byref(bases)))
        lp, ip = u64p(), f64p()
        self._ck(self.lib.dll.rr_collect_meta(self.h, C.byref(lp), C.byref(ip)))
        nn = int(n.value)
        lengths = np.ctypeslib.as_array(lp, shape=(nn,)).copy() if nn else np.zeros(0, np.uint64)
        ident = np.ctypeslib.as_array(ip, shape=(nn,)).copy() if nn else np.zeros(0, np.float64)
        return nn, int(bases.value), lengths, ident

    def parts(self):
        """After collect_begin(): yields (first_read, Batch) for consecutive parts of the draining batch."""
        while True:
            b = rr_batch()
            first = C.c_uint64()
            self._ck(self.lib.dll.rr_collect_part(self.h, C.byref(b), C.byref(first)))
            if b.n_reads == 0:
                return
            yield int(first.value), Batch(b)

    def collect_parts(self):
        """Streaming form of collect(): yields (first_read, Batch) for consecutive parts of the oldest submitted batch."""
        n, bases = C.c_uint64(), C.c_uint64()
        self._ck(self.lib.dll.rr_collect_begin(self.h, C.byref(n), C.byref(bases)))
        while True:
            b = rr_batch()
            first = C.c_uint64()
            self._ck(self.lib.dll.rr_collect_part(self.h, C.byref(b), C.byref(first)))
            if b.n_reads == 0:
                return
            yield int(first.value), Batch(b)

    def kernel_launches(self):
        return int(self.lib.dll.rr_kernel_launches(self.h))

    def last_run_ms(self):
        ms = (C.c_float * 4)()
        self._ck(self.lib.dll.rr_last_run_ms(self.h, ms))
        return dict(total=ms[0], read_kernel=ms[1], qscore=ms[2], pack=ms[3])

    def last_run_bases(self):
        return int(self.lib.dll.rr_last_run_bases(self.h))

    # -- debug / parity hooks ------------------------------------------------------------------
    def debug_align(self, query, target):
        out = (C.c_uint8 * (len(query) + len(target) + 1))()
        n, e = C.c_uint32(), C.c_uint32()
        self._ck(self.lib.dll.rr_debug_align(self.h, query, len(query), target, len(target), out, C.byref(n), C.byref(e)))
        return e.value, bytes(out[:n.value]).decode()

    def debug_rng_words(self, seed, first, n):
        out = np.zeros(n, dtype=np.uint32)
        self._ck(self.lib.dll.rr_debug_rng_words(self.h, seed, first, n, out.ctypes.data))
        return out

    def debug_div_check(self, seed, n):
        """Number of (dividend, divisor) pairs on which the walk's division differs from IEEE `/` in any bit."""
        bad = C.c_uint64()
        self._ck(self.lib.dll.rr_debug_div_check(self.h, seed, n, C.byref(bad)))
        return bad.value

    def debug_set_part_bytes(self, nbytes):
        self._ck(self.lib.dll.rr_debug_set_part_bytes(self.h, nbytes))

    def debug_set_slack(self, slack=0.0):
        n = C.c_uint64()
        self._ck(self.lib.dll.rr_debug_set_slack(self.h, slack, C.byref(n)))
        return n.value

    def debug_fetch(self, i, which, cap=1 << 22):
        out = (C.c_uint8 * cap)()
        n = C.c_uint64()
        self._ck(self.lib.dll.rr_debug_fetch(self.h, i, {"raw": 0, "cigar": 1}[which], out, cap, C.byref(n)))
        return bytes(out[:n.value])


def descs_to_item_table(descs):
    """rr_desc records -> the (n,12) u64 table + identities used by the oracle's test hooks (test helper)."""
    d12 = np.zeros((len(descs), 12), dtype=np.uint64)
    for j, f in enumerate(("type", "ref_idx", "strand", "start", "end", "has_chimera", "ctype", "cref_idx", "cstrand", "cstart", "cend", "seed")):
        d12[:, j] = descs[f]
    return d12, np.ascontiguousarray(descs["identity"], dtype=np.float64)


def descs_from_oracle_items(d12, identity):
    """Convert the oracle feeder's (n,12) u64 item table into rr_desc records (test helper)."""
    n = len(d12)
    out = np.zeros(n, dtype=DESC_DTYPE)
    out["type"] = d12[:, 0]
    out["ref_idx"] = d12[:, 1]
    out["strand"] = d12[:, 2]
    out["start"] = d12[:, 3]
    out["end"] = d12[:, 4]
    out["has_chimera"] = d12[:, 5]
    out["ctype"] = d12[:, 6]
    out["cref_idx"] = d12[:, 7]
    out["cstrand"] = d12[:, 8]
    out["cstart"] = d12[:, 9]
    out["cend"] = d12[:, 10]
    out["seed"] = d12[:, 11]
    out["identity"] = identity
    return out


# ---------------------------------------------------------------------------------------------------
# host side of `rustyread simulate` (include/rustyread_b200_host.h)
# ---------------------------------------------------------------------------------------------------
class rr_sim_params(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("length_mean", C.c_double), ("length_stdev", C.c_double),
                ("identity_mean", C.c_double), ("identity_max", C.c_double), ("identity_stdev", C.c_double),
                ("start_adapter_seq", C.c_char_p), ("end_adapter_seq", C.c_char_p),
                ("start_adapter_rate", C.c_double), ("start_adapter_amount", C.c_double),
                ("end_adapter_rate", C.c_double), ("end_adapter_amount", C.c_double),
                ("junk_reads", C.c_double), ("random_reads", C.c_double), ("chimera", C.c_double),
                ("glitch_rate", C.c_double), ("glitch_size", C.c_double), ("glitch_skip", C.c_double),
                ("small_plasmid_bias", C.c_int),
                ("error_model", C.c_char_p), ("error_model_len", C.c_uint64),
                ("qscore_model", C.c_char_p), ("qscore_model_len", C.c_uint64)]


HOST_EXPORTS = {
    "rr_sim_default_params": (None, [C.POINTER(rr_sim_params)]),
    "rr_sim_create": (C.c_int, [C.POINTER(rr_sim_params), C.c_char_p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "rr_sim_destroy": (None, [C.c_void_p]),
    "rr_sim_last_error": (C.c_char_p, [C.c_void_p]),
    "rr_sim_genome_length": (C.c_uint64, [C.c_void_p]),
    "rr_sim_k": (C.c_uint32, [C.c_void_p]),
    "rr_quantity_to_bases": (C.c_int64, [C.c_char_p, C.c_uint64]),
    "rr_sim_attach": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rr_sim_attach_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "rr_sim_create_indexed": (C.c_int, [C.POINTER(rr_sim_params), C.c_uint64, C.POINTER(C.c_char_p), u64p, C.c_void_p, f64p, C.POINTER(C.c_void_p)]),
    "rr_sim_contig_index": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_char_p), u64p, C.POINTER(C.c_int), f64p]),
    "rr_sim_feed": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p), u64p]),
    "rr_sim_fastq": (C.c_int, [C.c_void_p, C.POINTER(rr_batch), C.c_uint64, C.POINTER(C.c_void_p), u64p, u64p, u64p]),
    "rr_sim_write_fastq": (C.c_int, [C.c_void_p, C.POINTER(rr_batch), C.c_uint64, C.c_int, u64p, u64p, u64p]),
    "rr_sim_skip_names": (C.c_int, [C.c_void_p, C.c_uint64]),
    "rr_sim_set_writer_threads": (C.c_int, [C.c_void_p, C.c_int]),
    "rr_sim_fastq_bytes": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, u64p, u64p, u64p]),
    "rr_sim_item_description": (C.c_int, [C.c_void_p, C.c_uint64, C.c_char_p, C.c_uint64]),
    "rr_sim_n_refs": (C.c_uint64, [C.c_void_p]),
    "rr_sim_ref": (C.c_void_p, [C.c_void_p, C.c_uint64, u64p, C.POINTER(C.c_int), f64p, C.POINTER(C.c_char_p)]),
    "rr_sim_rng_words": (C.c_uint64, [C.c_void_p]),
}
EXPORTS.update(HOST_EXPORTS)


class Simulation:
    """Mirror of the reference's `simulate::simulate` orchestrator (src/simulate/mod.rs:27-217):
    builds main_rng, the models and the references in the reference's order, then feeds batches of
    work items and formats FASTQ records.  Flags are the CLI's (src/cli/simulate.rs:162-239)."""

    def __init__(self, fasta, seed=42, length=(15000, 13000), identity=(85, 95, 5),
                 start_adapter=(90, 60), end_adapter=(50, 20),
                 start_adapter_seq="AATGTACTTCGTTCAGTTACGTATTGCT", end_adapter_seq="GCAATACGTAACTGAACGAAGT",
                 junk_reads=1, random_reads=1, chimera=1, glitches=(10000, 25, 25), small_plasmid_bias=False,
                 error_model="random", qscore_model="random", lib=None, index=None):
        self.lib = lib or load_library()
        if isinstance(fasta, str):
            fasta = fasta.encode()
        p = rr_sim_params()
        self.lib.dll.rr_sim_default_params(C.byref(p))
        p.seed = seed
        p.length_mean, p.length_stdev = length
        p.identity_mean, p.identity_max, p.identity_stdev = identity
        self._keep = [start_adapter_seq.encode(), end_adapter_seq.encode()]
        p.start_adapter_seq, p.end_adapter_seq = self._keep
        p.start_adapter_rate, p.start_adapter_amount = start_adapter
        p.end_adapter_rate, p.end_adapter_amount = end_adapter
        p.junk_reads, p.random_reads, p.chimera = junk_reads, random_reads, chimera
        p.glitch_rate, p.glitch_size, p.glitch_skip = glitches
        p.small_plasmid_bias = int(small_plasmid_bias)
        em = None if error_model in (None, "random") else (error_model.encode() if isinstance(error_model, str) else error_model)
        qm = (qscore_model.encode() if isinstance(qscore_model, str) else qscore_model)
        self._keep += [em, qm]
        p.error_model, p.error_model_len = em, (len(em) if em else 0)
        p.qscore_model, p.qscore_model_len = qm, len(qm)
        h = C.c_void_p()
        if index is not None:  # contigs without bases: [(id, length, circular, header weight)], see contig_index()
            n = len(index)
            ids = (C.c_char_p * n)(*[i[0].encode("utf-8", "surrogateescape") for i in index])
            lens = (C.c_uint64 * n)(*[int(i[1]) for i in index])
            circ = (C.c_uint8 * n)(*[1 if i[2] else 0 for i in index])
            wts = (C.c_double * n)(*[float(i[3]) for i in index])
            rc = self.lib.dll.rr_sim_create_indexed(C.byref(p), n, ids, lens, circ, wts, C.byref(h))
        else:
            rc = self.lib.dll.rr_sim_create(C.byref(p), fasta, len(fasta), C.byref(h))
        if rc:
            raise RRError(rc, self.lib.dll.rr_sim_last_error(None).decode())
        self.h = h
        self.k = int(self.lib.dll.rr_sim_k(h))
        self.genome_length = int(self.lib.dll.rr_sim_genome_length(h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.dll.rr_sim_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise RRError(rc, self.lib.dll.rr_sim_last_error(self.h).decode())

    def quantity(self, text):
        v = self.lib.dll.rr_quantity_to_bases(text.encode(), self.genome_length)
        if v < 0:
            raise RRError(2, "cannot parse quantity %r" % text)
        return int(v)

    def attach(self, ctx, device_bases=None):
        """References + models -> device context; device_bases: device address of the concatenated contigs when they are
        already on the context's GPU (a session made from an index has no host copy)."""
        if device_bases is None:
            self._ck(self.lib.dll.rr_sim_attach(self.h, ctx.h))
        else:
            self._ck(self.lib.dll.rr_sim_attach_device(self.h, ctx.h, C.c_void_p(device_bases)))
        ctx.k = self.k

    def contig_index(self):
        """[(id, length, circular, header weight = depth x length)] — all a session needs of the reference besides the bases."""
        out = []
        for i in range(self.lib.dll.rr_sim_n_refs(self.h)):
            rid, ln, circ, w = C.c_char_p(), C.c_uint64(), C.c_int(), C.c_double()
            self._ck(self.lib.dll.rr_sim_contig_index(self.h, i, C.byref(rid), C.byref(ln), C.byref(circ), C.byref(w)))
            out.append((rid.value.decode("utf-8", "surrogateescape"), int(ln.value), bool(circ.value), float(w.value)))
        return out

    def feed(self, base_loop):
        """One batch of the Fragments iterator; returns a numpy array of DESC_DTYPE (a copy)."""
        p, n = C.c_void_p(), C.c_uint64()
        self._ck(self.lib.dll.rr_sim_feed(self.h, base_loop, C.byref(p), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, dtype=DESC_DTYPE)
        buf = (C.c_uint8 * (n.value * DESC_DTYPE.itemsize)).from_address(p.value)
        return np.frombuffer(buf, dtype=DESC_DTYPE).copy()

    def fastq(self, batch_struct, first=0):
        """FASTQ text for a device batch (rr_batch struct) holding items [first, first+n) of the last feed."""
        t, n, recs, bases = C.c_void_p(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._ck(self.lib.dll.rr_sim_fastq(self.h, C.byref(batch_struct), first, C.byref(t), C.byref(n), C.byref(recs), C.byref(bases)))
        return C.string_at(t.value, n.value), recs.value, bases.value

    def write_fastq(self, batch_struct, fd, first=0):
        """Write the records of a device batch to file descriptor fd (writev straight from the pinned arenas)."""
        recs, bases, nbytes = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._ck(self.lib.dll.rr_sim_write_fastq(self.h, C.byref(batch_struct), first, fd, C.byref(recs), C.byref(bases), C.byref(nbytes)))
        return recs.value, bases.value, nbytes.value

    def set_writer_threads(self, threads):
        """Threads write_fastq may use on a regular file (contiguous record ranges sent to their offsets with pwritev)."""
        self._ck(self.lib.dll.rr_sim_set_writer_threads(self.h, int(threads)))

    def fastq_bytes(self, lengths, real_identity, first=0):
        """(bytes, records, bases) of the FASTQ text for items [first, first + n) with these final lengths / identities."""
        lengths = np.ascontiguousarray(lengths, np.uint64)
        real_identity = np.ascontiguousarray(real_identity, np.float64)
        nb, nr, nbs = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._ck(self.lib.dll.rr_sim_fastq_bytes(self.h, len(lengths), lengths.ctypes.data, real_identity.ctypes.data, self.k, first,
                                                 C.byref(nb), C.byref(nr), C.byref(nbs)))
        return int(nb.value), int(nr.value), int(nbs.value)

    def skip_names(self, n_records):
        self._ck(self.lib.dll.rr_sim_skip_names(self.h, n_records))

    def refs(self):
        out = []
        for i in range(self.lib.dll.rr_sim_n_refs(self.h)):
            ln, circ, w, rid = C.c_uint64(), C.c_int(), C.c_double(), C.c_char_p()
            p = self.lib.dll.rr_sim_ref(self.h, i, C.byref(ln), C.byref(circ), C.byref(w), C.byref(rid))
            out.append(dict(id=rid.value.decode("utf-8", "surrogateescape"), seq=C.string_at(p, ln.value), circular=bool(circ.value), weight=w.value))
        return out

    def rng_words(self):
        return int(self.lib.dll.rr_sim_rng_words(self.h))


def batch_struct_from_arrays(offsets, lengths, seq, qual, real_identity, k):
    """Build an rr_batch over caller-owned numpy arrays (keeps references alive on the struct)."""
    b = rr_batch()
    keep = [np.ascontiguousarray(offsets, np.uint64), np.ascontiguousarray(lengths, np.uint64), np.ascontiguousarray(seq, np.uint8),
            np.ascontiguousarray(qual, np.uint8), np.ascontiguousarray(real_identity, np.float64)]
    b.n_reads = len(keep[1])
    b.offsets = keep[0].ctypes.data_as(u64p)
    b.lengths = keep[1].ctypes.data_as(u64p)
    b.seq = keep[2].ctypes.data_as(u8p)
    b.qual = keep[3].ctypes.data_as(u8p)
    b.real_identity = keep[4].ctypes.data_as(f64p)
    b.total_bytes = int(keep[0][-1]) if len(keep[0]) else 0
    b.total_bases = int(sum(int(x) - 2 * k for x in keep[1] if x > 14))
    b.k = k
    b._keep = keep
    return b
