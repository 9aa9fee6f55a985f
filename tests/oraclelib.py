"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "liboracle.so")

u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)


def build(force=False):
    srcs = [os.path.join(ROOT, "oracle", f) for f in os.listdir(os.path.join(ROOT, "oracle")) if f.endswith((".cpp", ".hpp"))]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return LIB


def _b(x):
    if isinstance(x, str):
        x = x.encode()
    return bytes(x)


def _buf(b):
    b = _b(b)
    return (C.c_uint8 * max(1, len(b))).from_buffer_copy(b + (b"\0" if not b else b""))


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        L = lib
        sig = {
            "orc_last_error": (C.c_char_p, []),
            "orc_free": (None, [C.c_void_p]),
            "orc_rng_new": (C.c_void_p, [C.c_uint64]),
            "orc_rng_free": (None, [C.c_void_p]),
            "orc_rng_next_u32": (C.c_uint32, [C.c_void_p]),
            "orc_rng_next_u64": (C.c_uint64, [C.c_void_p]),
            "orc_rng_words": (C.c_uint64, [C.c_void_p]),
            "orc_rng_gen_f64": (C.c_double, [C.c_void_p]),
            "orc_rng_gen_bool": (C.c_int, [C.c_void_p, C.c_double]),
            "orc_rng_gen_range_u64": (C.c_uint64, [C.c_void_p, C.c_uint64, C.c_uint64]),
            "orc_rng_gen_range_u32": (C.c_uint32, [C.c_void_p, C.c_uint32, C.c_uint32]),
            "orc_rng_words_block": (None, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]),
            "orc_random_base": (C.c_int, [C.c_void_p]),
            "orc_random_base_diff": (C.c_int, [C.c_void_p, C.c_int]),
            "orc_random_seq": (None, [C.c_void_p, C.c_uint64, C.c_void_p]),
            "orc_number_of_edit": (C.c_double, [C.c_double, C.c_uint64]),
            "orc_weighted_new": (C.c_void_p, [C.c_void_p, C.c_uint64]),
            "orc_weighted_free": (None, [C.c_void_p]),
            "orc_weighted_sample": (C.c_uint64, [C.c_void_p, C.c_void_p]),
            "orc_weighted_scale": (C.c_double, [C.c_void_p]),
            "orc_geometric_sample": (C.c_uint64, [C.c_double, C.c_void_p]),
            "orc_geometric_params": (None, [C.c_double, f64p, u64p]),
            "orc_beta_sample": (C.c_double, [C.c_double, C.c_double, C.c_void_p]),
            "orc_gamma_sample": (C.c_double, [C.c_double, C.c_double, C.c_void_p]),
            "orc_standard_normal": (C.c_double, [C.c_void_p]),
            "orc_length_new": (C.c_void_p, [C.c_double, C.c_double]),
            "orc_length_free": (None, [C.c_void_p]),
            "orc_length_get": (C.c_uint64, [C.c_void_p, C.c_void_p]),
            "orc_identity_new": (C.c_void_p, [C.c_double, C.c_double, C.c_double]),
            "orc_identity_free": (None, [C.c_void_p]),
            "orc_identity_get": (C.c_double, [C.c_void_p, C.c_void_p]),
            "orc_adapter_new": (C.c_void_p, [C.c_char_p, C.c_char_p, C.c_double, C.c_double, C.c_double, C.c_double]),
            "orc_adapter_free": (None, [C.c_void_p]),
            "orc_adapter_get": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
            "orc_adapter_max_len": (C.c_uint64, [C.c_void_p]),
            "orc_adapter_params": (None, [C.c_void_p, C.c_void_p]),
            "orc_glitch_new": (C.c_void_p, [C.c_double, C.c_double, C.c_double]),
            "orc_glitch_free": (None, [C.c_void_p]),
            "orc_glitch_get": (C.c_int64, [C.c_void_p, C.c_void_p, u64p, u64p, C.c_void_p, C.c_uint64]),
            "orc_error_model_random": (C.c_void_p, [C.c_uint64]),
            "orc_error_model_from_text": (C.c_void_p, [C.c_char_p, C.c_void_p]),
            "orc_error_model_free": (None, [C.c_void_p]),
            "orc_error_model_k": (C.c_uint64, [C.c_void_p]),
            "orc_error_model_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
            "orc_random_error": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
            "orc_error_model_dump": (C.c_void_p, [C.c_void_p]),
            "orc_quality_model_new": (C.c_void_p, [C.c_char_p, C.c_char_p]),
            "orc_quality_model_free": (None, [C.c_void_p]),
            "orc_quality_model_max_k": (C.c_uint64, [C.c_void_p]),
            "orc_get_qscore": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
            "orc_generate_quality": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
            "orc_quality_model_dump": (C.c_void_p, [C.c_void_p]),
            "orc_edit_distance": (C.c_uint64, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]),
            "orc_align": (C.c_uint64, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, u64p]),
            "orc_change_script": (C.c_void_p, [C.c_void_p, C.c_uint64, C.c_char_p]),
            "orc_add_glitches": (C.c_void_p, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]),
            "orc_add_error": (C.c_void_p, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_double, C.c_void_p, C.c_uint64]),
            "orc_sequence": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, u64p, C.c_void_p, u64p, f64p, C.c_uint64]),
            "orc_fmt_f64": (C.c_void_p, [C.c_double]),
            "orc_description": (C.c_void_p, [C.c_int, C.c_char_p, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_double]),
            "orc_uuid_v3_x500": (C.c_void_p, [C.c_void_p, C.c_uint64]),
            "orc_md5": (None, [C.c_void_p, C.c_uint64, C.c_void_p]),
            "orc_quantity": (C.c_int64, [C.c_char_p, C.c_uint64]),
            "orc_revcomp": (None, [C.c_void_p, C.c_uint64, C.c_void_p]),
            "orc_workload_new": (C.c_void_p, [C.c_uint64]),
            "orc_workload_free": (None, [C.c_void_p]),
            "orc_workload_rng": (C.c_void_p, [C.c_void_p]),
            "orc_workload_set_length": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
            "orc_workload_set_identity": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double]),
            "orc_workload_set_adapter": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_double, C.c_double, C.c_double, C.c_double]),
            "orc_workload_set_glitch": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double]),
            "orc_workload_set_error_model": (C.c_int, [C.c_void_p, C.c_char_p]),
            "orc_workload_set_qscore_model": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
            "orc_workload_set_references": (C.c_int, [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int, C.c_int]),
            "orc_workload_set_rates": (None, [C.c_void_p, C.c_double, C.c_double, C.c_double]),
            "orc_workload_n_refs": (C.c_uint64, [C.c_void_p]),
            "orc_workload_ref_len": (C.c_uint64, [C.c_void_p, C.c_uint64]),
            "orc_workload_ref_seq": (C.c_void_p, [C.c_void_p, C.c_uint64]),
            "orc_workload_ref_circular": (C.c_int, [C.c_void_p, C.c_uint64]),
            "orc_workload_ref_id": (C.c_char_p, [C.c_void_p, C.c_uint64]),
            "orc_workload_ref_weight": (C.c_double, [C.c_void_p, C.c_uint64]),
            "orc_workload_choose_reference": (None, [C.c_void_p, C.c_void_p, u64p, C.POINTER(C.c_int)]),
            "orc_workload_feed": (C.c_uint64, [C.c_void_p, C.c_uint64, C.c_uint64]),
            "orc_workload_items": (None, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
            "orc_workload_generate": (C.c_double, [C.c_void_p, C.c_int, C.c_int]),
            "orc_workload_load_items": (None, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
            "orc_workload_set_item_identity": (None, [C.c_void_p, C.c_uint64, C.c_double]),
            "orc_workload_out_len": (C.c_uint64, [C.c_void_p, C.c_uint64, C.c_int]),
            "orc_workload_out_lens": (None, [C.c_void_p, C.c_int, C.c_void_p]),
            "orc_workload_out_copy": (None, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
            "orc_workload_out_identity": (None, [C.c_void_p, C.c_void_p]),
            "orc_workload_out_description": (C.c_void_p, [C.c_void_p, C.c_uint64]),
            "orc_workload_item_description": (C.c_void_p, [C.c_void_p, C.c_uint64]),
            "orc_simulate": (C.c_void_p, [C.c_char_p, C.c_char_p, C.c_uint64, C.c_char_p, C.c_char_p, C.c_int, f64p]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args

    # ---- helpers -------------------------------------------------------------
    def err(self):
        return self.lib.orc_last_error().decode()

    def take_str(self, p):
        if not p:
            raise RuntimeError(self.err())
        s = C.string_at(p)
        self.lib.orc_free(p)
        return s.decode("utf-8", "surrogateescape")

    def rng(self, seed):
        return Rng(self, seed)

    def align(self, q, t):
        q, t = _b(q), _b(t)
        out = (C.c_uint8 * (len(q) + len(t) + 1))()
        n = C.c_uint64()
        e = self.lib.orc_align(_buf(q), len(q), _buf(t), len(t), out, C.byref(n))
        return e, bytes(out[: n.value]).decode()

    def edit_distance(self, a, b):
        a, b = _b(a), _b(b)
        return self.lib.orc_edit_distance(_buf(a), len(a), _buf(b), len(b))

    def workload(self, seed):
        return Workload(self, seed)

    def simulate(self, params, fasta, error_text=None, qscore_text=None, capture=False):
        kv = "".join("%s=%s\n" % (k, v) for k, v in params.items()).encode()
        out = (C.c_double * 5)()
        fasta = _b(fasta)
        p = self.lib.orc_simulate(kv, fasta, len(fasta), _b(error_text) if error_text is not None else None,
                                  _b(qscore_text) if qscore_text is not None else None, 1 if capture else 0, out)
        text = self.take_str(p)
        stats = dict(reads=int(out[0]), bases=int(out[1]), fastq_bytes=int(out[2]), generate_seconds=out[3], total_seconds=out[4])
        return text, stats


class Rng:
    def __init__(self, o, seed):
        self.o = o
        self.h = o.lib.orc_rng_new(seed)

    def __del__(self):
        try:
            self.o.lib.orc_rng_free(self.h)
        except Exception:
            pass

    def next_u32(self):
        return self.o.lib.orc_rng_next_u32(self.h)

    def next_u64(self):
        return self.o.lib.orc_rng_next_u64(self.h)

    def words(self):
        return self.o.lib.orc_rng_words(self.h)

    def gen_f64(self):
        return self.o.lib.orc_rng_gen_f64(self.h)

    def random_seq(self, n):
        out = (C.c_uint8 * max(1, n))()
        self.o.lib.orc_random_seq(self.h, n, out)
        return bytes(out[:n])


class Workload:
    """References + models + feeder + per-read CPU generation (the parity checker's side)."""

    def __init__(self, o, seed):
        self.o = o
        self.L = o.lib
        self.h = self.L.orc_workload_new(seed)
        self.n = 0

    def __del__(self):
        try:
            self.L.orc_workload_free(self.h)
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(self.o.err())

    def set_length(self, mean, stdev):
        self._ck(self.L.orc_workload_set_length(self.h, mean, stdev))

    def set_identity(self, mean, mx, stdev):
        self._ck(self.L.orc_workload_set_identity(self.h, mean, mx, stdev))

    def set_adapter(self, s, e, sr, sa, er, ea):
        self._ck(self.L.orc_workload_set_adapter(self.h, _b(s), _b(e), sr, sa, er, ea))

    def set_glitch(self, rate, size, skip):
        self._ck(self.L.orc_workload_set_glitch(self.h, rate, size, skip))

    def set_error_model(self, text=None):
        self._ck(self.L.orc_workload_set_error_model(self.h, _b(text) if text is not None else None))

    def set_qscore_model(self, kind, text=None):
        self._ck(self.L.orc_workload_set_qscore_model(self.h, _b(kind), _b(text) if text is not None else None))

    def set_references(self, fasta, adjusted=False, small_plasmid_bias=False):
        fasta = _b(fasta)
        self._ck(self.L.orc_workload_set_references(self.h, fasta, len(fasta), int(adjusted), int(small_plasmid_bias)))

    def set_rates(self, junk, random, chimera):
        self.L.orc_workload_set_rates(self.h, junk, random, chimera)

    def rng_handle(self):
        return self.L.orc_workload_rng(self.h)

    def refs(self):
        out = []
        for i in range(self.L.orc_workload_n_refs(self.h)):
            n = self.L.orc_workload_ref_len(self.h, i)
            seq = C.string_at(self.L.orc_workload_ref_seq(self.h, i), n)
            out.append(dict(id=self.L.orc_workload_ref_id(self.h, i).decode("utf-8", "surrogateescape"), seq=seq,
                            circular=bool(self.L.orc_workload_ref_circular(self.h, i)),
                            weight=self.L.orc_workload_ref_weight(self.h, i)))
        return out

    def feed(self, target, max_items=0):
        n = self.L.orc_workload_feed(self.h, target, max_items)
        if n == 2 ** 64 - 1:
            self.n = 0
            raise RuntimeError(self.o.err())
        self.n = n
        return self.n

    def items(self):
        d = np.zeros((self.n, 12), dtype=np.uint64)
        ident = np.zeros(self.n, dtype=np.float64)
        length = np.zeros(self.n, dtype=np.uint64)
        if self.n:
            self.L.orc_workload_items(self.h, d.ctypes.data, ident.ctypes.data, length.ctypes.data)
        return d, ident, length

    def load_items(self, d12, identity):
        """Replace the fed items by caller-supplied descriptors ((n,12) u64 table + identities)."""
        d12 = np.ascontiguousarray(d12, dtype=np.uint64)
        identity = np.ascontiguousarray(identity, dtype=np.float64)
        self.n = len(d12)
        self.L.orc_workload_load_items(self.h, d12.ctypes.data, identity.ctypes.data, self.n)

    def set_item_identity(self, i, identity):
        self.L.orc_workload_set_item_identity(self.h, i, identity)

    def generate(self, threads=1, keep_cigar_raw=True):
        s = self.L.orc_workload_generate(self.h, threads, int(keep_cigar_raw))
        if s < 0:
            raise RuntimeError(self.o.err())
        return s

    def out(self, which):
        """(arena uint8, offsets uint64[n+1]) for which in seq/qual/cigar/raw."""
        w = {"seq": 0, "qual": 1, "cigar": 2, "raw": 3}[which]
        lens = np.zeros(self.n, dtype=np.uint64)
        if self.n:
            self.L.orc_workload_out_lens(self.h, w, lens.ctypes.data)
        off = np.zeros(self.n + 1, dtype=np.uint64)
        np.cumsum(lens, out=off[1:])
        arena = np.zeros(max(1, int(off[-1])), dtype=np.uint8)
        if self.n:
            self.L.orc_workload_out_copy(self.h, w, arena.ctypes.data, off.ctypes.data)
        return arena, off

    def out_identity(self):
        a = np.zeros(self.n, dtype=np.float64)
        if self.n:
            self.L.orc_workload_out_identity(self.h, a.ctypes.data)
        return a

    def out_description(self, i):
        return self.o.take_str(self.L.orc_workload_out_description(self.h, i))

    def item_description(self, i):
        return self.o.take_str(self.L.orc_workload_item_description(self.h, i))


_cached = None


def load():
    global _cached
    if _cached is None:
        build()
        _cached = Oracle(C.CDLL(LIB))
    return _cached
