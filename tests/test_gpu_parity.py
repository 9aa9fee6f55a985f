"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the reference's goldens.
Bit-exact: every byte of every sequence and quality string, and every f64 identity, must be equal."""
import random

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rr():
    import rustyread_b200

    return rustyread_b200


@pytest.fixture(scope="module")
def ctx(rr):
    c = rr.Context(0)
    yield c
    c.close()


# ---- K1: ChaCha12 word stream --------------------------------------------------------------------
@pytest.mark.parametrize("seed,first,n", [(42, 0, 2000), (0, 0, 64), (2 ** 64 - 1, 0, 100), (17195042692806716983, 12345, 1500), (7, 511, 3)])
def test_rng_words(orc, ctx, seed, first, n):
    want = np.zeros(n, dtype=np.uint32)
    orc.lib.orc_rng_words_block(seed, first, n, want.ctypes.data)
    got = ctx.debug_rng_words(seed, first, n)
    assert np.array_equal(got, want)


# ---- the scan's division: reciprocal + Newton fast path vs the compiler's IEEE `/`, bit for bit -------------------
@pytest.mark.parametrize("seed", [1, 42])
def test_walk_division_is_ieee_exact(ctx, seed):
    assert ctx.debug_div_check(seed, 1 << 28) == 0


# ---- K4: alignment -----------------------------------------------------------------------------------
def test_align_goldens(ctx, goldens):  # src/alignment.rs:47-86
    for q, t, e, cigar in goldens["alignment.align"]:
        assert ctx.debug_align(q.encode(), t.encode()) == (e, cigar)


def test_align_glitch_goldens(orc, ctx, goldens):  # windows of src/simulate/error.rs:646-713
    g = goldens["simulate.error.glitches"]
    raw = orc.rng(42).random_seq(g["raw_len"])
    for c in g["truth"]:
        got = ctx.debug_align(c["seq"].encode(), raw[c["begin"]:c["end"]])
        assert got == (c["edit_distance"], c["cigar"])


@pytest.mark.parametrize("mmax,nmax,alphabet,count", [
    (12, 12, "ACGT", 400),     # k-mer sized windows (fast path)
    (40, 64, "AC", 200),       # low-entropy: many ties
    (127, 64, "ACGT", 60),     # fast-path limits
    (128, 65, "ACGT", 40),     # just past the fast path
    (300, 300, "ACGTN", 40),   # glitch-sized windows, multi-word rows, non-ACGT bytes
    (900, 1100, "ACGT", 6),
    (5, 2048, "ACGT", 4),
    (1500, 3, "ACGT", 4),
])
def test_align_fuzz(orc, ctx, mmax, nmax, alphabet, count):
    rnd = random.Random(mmax * 7919 + nmax)
    for it in range(count):
        m = rnd.randint(0, mmax) if it % 5 else mmax
        n = rnd.randint(0, nmax) if it % 7 else nmax
        q = "".join(rnd.choice(alphabet) for _ in range(m))
        if it % 3 == 0 and m:  # related sequences, like a real error window
            t = list(q)
            for _ in range(rnd.randint(0, 3)):
                op = rnd.choice("sid")
                p = rnd.randrange(len(t) + 1)
                if op == "s" and p < len(t):
                    t[p] = rnd.choice(alphabet)
                elif op == "i":
                    t.insert(p, rnd.choice(alphabet))
                elif op == "d" and p < len(t):
                    del t[p]
            t = "".join(t)[:nmax]
        else:
            t = "".join(rnd.choice(alphabet) for _ in range(n))
        assert ctx.debug_align(q.encode(), t.encode()) == orc.align(q, t), (q, t)


def test_align_too_large_fails_loudly(rr, ctx):
    with pytest.raises(rr.RRError):
        ctx.debug_align(b"A" * 10, b"C" * 3000)


# ---- full reads: the reference's own 20-read golden --------------------------------------------------------
def test_produce_read_golden(orc, goldens):  # src/simulate/mod.rs:433-919
    spec = dict(length=(8.0, 2.0), identity=(85.0, 95.0, 5.0), adapter=("TACGTATTGCT", "TACGTATTGCT", 90.0, 60.0, 50.0, 20.0),
                glitch=(50.0, 5.0, 5.0), rates=(10.0, 10.0, 25.0), adjusted=False)
    w, ctx, _ = helpers.make_pair(orc, goldens["simulate.FASTA"], spec)
    w.feed(10_000, 20)
    d, ident, _ = w.items()
    import rustyread_b200.capi as capi

    batch = ctx.generate(capi.descs_from_oracle_items(d, ident))
    truth = goldens["simulate.produce_read"]
    for i, (desc, wseq, wqual) in enumerate(truth):
        s, q = batch.read(i)
        assert list(s) == wseq, i
        assert list(q) == wqual, i
        assert batch.real_identity[i] == float(desc["identity"]["f64"]), i
    w.generate(threads=2)
    helpers.compare_batch(w, batch, ctx)
    ctx.close()


def _run_pair(orc, fasta, spec, target, max_items=0, threads=8):
    import rustyread_b200.capi as capi

    w, ctx, s = helpers.make_pair(orc, fasta, spec)
    n = w.feed(target, max_items)
    d, ident, _ = w.items()
    batch = ctx.generate(capi.descs_from_oracle_items(d, ident))
    w.generate(threads=threads)
    helpers.compare_batch(w, batch, ctx)
    return w, ctx, batch, n


def test_shape_5mbp_golden(orc, goldens):  # src/simulate/mod.rs:922-1023 (58 reads, 1 048 243 bases)
    import ctypes as C
    import math

    w0 = orc.workload(42)
    w0.set_length(18200.0, 15500.0)
    rh = w0.rng_handle()
    parts = []
    for n in (5_000_000, 80_000, 75_000):
        out = (C.c_uint8 * n)()
        orc.lib.orc_random_seq(rh, n, out)
        parts.append(">%s\n%s\n" % ("{:_}".format(n), bytes(out).decode()))
    fasta = "".join(parts)
    # the reference test draws the contigs from the same rng that then feeds the run: replay that state
    import rustyread_b200 as rrm
    import rustyread_b200.capi as capi

    w0.set_references(fasta, adjusted=True)
    w0.set_identity(87.0, 100.0, 9.0)
    w0.set_adapter("TACGTATTGCT", "TACGTATTGCT", 90.0, 60.0, 50.0, 20.0)
    w0.set_error_model(None)
    w0.set_qscore_model("random")
    w0.set_glitch(50.0, 5.0, 5.0)
    w0.set_rates(1.0, 1.0, 1.0)
    n = w0.feed(1_000_000)
    d, ident, _ = w0.items()
    ctx = rrm.Context(0)
    refs = w0.refs()
    ctx.set_references([r["seq"] for r in refs], [r["circular"] for r in refs])
    ctx.set_error_model_random(7)
    ctx.set_qscore_model_random()
    ctx.set_adapter(b"TACGTATTGCT", b"TACGTATTGCT", 90.0, 60.0, 50.0, 20.0)
    ctx.set_glitch(50.0, 5.0, 5.0)
    batch = ctx.generate(capi.descs_from_oracle_items(d, ident))
    lengths = [float(x) for x in batch.lengths]
    sum_len = 0.0
    for x in lengths:
        sum_len += x
    sum_id = 0.0
    for x in batch.real_identity:
        sum_id += float(x)
    avg_id = sum_id / n
    acc = 0.0
    for x in batch.real_identity:
        acc += (float(x) - avg_id) ** 2.0
    want = goldens["simulate.shape"]
    assert [sum_len, sum_len / n, math.sqrt(sum_len / n), sum_id, avg_id, math.sqrt(acc / n)] == want
    w0.generate(threads=8)
    helpers.compare_batch(w0, batch, ctx)
    ctx.close()


# ---- BASELINE.json configs at sizes the oracle finishes in seconds ----------------------------------------------
def test_config1_default(orc):
    fasta = helpers.random_fasta(orc, [400_000])
    _, ctx, batch, n = _run_pair(orc, fasta, {}, 4_000_000)
    assert n > 150
    ctx.close()


def test_config2_full_size_5mbp_200x(orc):
    """BASELINE.json configs[1] at its full size (the bench workload): 5 Mbp reference, 200x, seed 42, defaults —
    every byte of all ~66 000 reads (1.0 Gbases) against the oracle, which needs ~10 s on the box's host threads."""
    import os

    import rustyread_b200.capi as capi

    fasta = helpers.random_fasta(orc, [5_000_000])
    w, ctx, s = helpers.make_pair(orc, fasta, {})
    n = w.feed(1_000_000_000)
    d, ident, _ = w.items()
    batch = ctx.generate(capi.descs_from_oracle_items(d, ident))
    w.generate(threads=os.cpu_count() or 8, keep_cigar_raw=False)
    helpers.compare_batch(w, batch, None, check_intermediate=False)
    assert n > 60_000 and batch.total_bases > 990_000_000
    ctx.close()


def test_config4_ultralong_low_identity(orc):
    fasta = helpers.random_fasta(orc, [600_000])
    spec = dict(length=(50000.0, 40000.0), identity=(75.0, 90.0, 8.0), glitch=(2000.0, 50.0, 50.0), rates=(1.0, 1.0, 5.0))
    _, ctx, batch, n = _run_pair(orc, fasta, spec, 6_000_000)
    assert n > 60
    ctx.close()


def test_config5_plasmids_divergent(orc):
    rnd = random.Random(5)
    contigs = [int(2000 * (25 ** rnd.random())) for _ in range(60)]
    fasta = helpers.random_fasta(orc, contigs, circular=set(range(60)), depth={3: "2.5", 7: "0.5"})
    spec = dict(rates=(10.0, 10.0, 1.0), small_plasmid_bias=True, length=(6000.0, 5000.0))
    _, ctx, batch, n = _run_pair(orc, fasta, spec, 3_000_000)
    assert n > 300
    ctx.close()


def test_mixed_linear_circular_chimeras(orc):
    fasta = helpers.random_fasta(orc, [120_000, 3_000, 50_000, 800, 20_000], circular={1, 3, 4}, depth={2: "3"})
    spec = dict(rates=(5.0, 5.0, 30.0), length=(4000.0, 3500.0), identity=(80.0, 99.0, 10.0), glitch=(500.0, 30.0, 20.0))
    _, ctx, batch, n = _run_pair(orc, fasta, spec, 2_000_000)
    ctx.close()


def test_file_models_fixtures(orc, goldens):
    # the two inline model fixtures of the reference (src/model/error.rs:156, src/model/quality.rs:146-150)
    fasta = "".join(">c%d\n%s\n" % (i, "ACAGTTG" * (3000 + 100 * i)) for i in range(3)) + helpers.random_fasta(orc, [50_000])
    spec = dict(error_model=goldens["model.error.MODEL"], qscore_model=goldens["model.quality.MODEL"], length=(3000.0, 2500.0))
    _, ctx, batch, n = _run_pair(orc, fasta, spec, 1_500_000)
    ctx.close()


def _synthetic_models(k=7, seed=9):
    rnd = random.Random(seed)
    lines = []
    for code in range(4 ** k):
        if rnd.random() < 0.25:
            continue  # a quarter of the k-mers miss the table -> random_error
        kmer = "".join("ACGT"[(code >> (2 * (k - 1 - i))) & 3] for i in range(k))
        alts, total = [], 0.0
        na = rnd.randint(1, 25)
        for _ in range(na):
            a = list(kmer)
            for _ in range(rnd.randint(1, 3)):
                op = rnd.choice("sid")
                p = rnd.randrange(len(a)) if a else 0
                if op == "s" and a:
                    a[p] = rnd.choice("ACGT")
                elif op == "i":
                    a.insert(p, rnd.choice("ACGT"))
                elif a:
                    del a[p]
            p_ = round(rnd.random() * 0.9 / na, 6)
            total += p_
            alts.append("%s,%s" % ("".join(a), repr(p_)))
        lines.append("%s,%s;%s;" % (kmer, repr(round(max(0.0, 0.05), 6)), ";".join(alts)))
    error_text = "\n".join(lines) + "\n"
    q = ["overall;1000;" + ",".join("%d:%s" % (s, 0.01) for s in range(1, 40)) + ","]
    seen = set()

    def add(c):
        if c in seen:
            return
        seen.add(c)
        n = rnd.randint(3, 45)
        scores = rnd.sample(range(1, 60), n)
        q.append("%s;%d;%s," % (c, rnd.randint(1, 10 ** 6), ",".join("%d:%s" % (s, repr(round(rnd.random(), 6) + 1e-6)) for s in scores)))

    for c in "=XI":
        add(c)
    for l in (3, 5, 7, 9):
        for _ in range({3: 40, 5: 200, 7: 600, 9: 1500}[l]):
            w = [rnd.choice("====XID") for _ in range(l)]
            if w[l // 2] == "D":
                w[l // 2] = "="
            add("".join(w))
    add("=" * 9)
    return error_text, "\n".join(q) + "\n"


def test_synthetic_badread_scale_models(orc):
    error_text, qscore_text = _synthetic_models()
    fasta = helpers.random_fasta(orc, [300_000])
    spec = dict(error_model=error_text, qscore_model=qscore_text, length=(8000.0, 7000.0))
    _, ctx, batch, n = _run_pair(orc, fasta, spec, 2_500_000)
    ctx.close()


def _free_port():
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _long_alt_model(k, seed=5):
    """Every k-mer in the table; alternatives of 0..64 bytes (longer than a warp: the memory path of add_change), several edits."""
    rnd = random.Random(seed)
    lines = []
    for code in range(4 ** k):
        kmer = "".join("ACGT"[(code >> (2 * (k - 1 - i))) & 3] for i in range(k))
        alts = []
        na = rnd.randint(1, 6)
        for _ in range(na):
            r = rnd.random()
            if r < 0.15:
                a = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(33, 64)))  # through memory
            elif r < 0.3:
                a = kmer[: rnd.randint(0, k - 1)]  # a prefix (possibly empty): all deletions
            elif r < 0.45:
                a = kmer + "".join(rnd.choice("ACGT") for _ in range(rnd.randint(8, 24)))  # more than 16 ops: the mid-size aligner
            else:
                a = list(kmer)
                a[rnd.randrange(k)] = rnd.choice("ACGT")
                a.insert(rnd.randrange(k), rnd.choice("ACGT"))
                a = "".join(a)
            alts.append("%s,%s" % (a, repr(round(0.9 / na, 6))))
        lines.append("%s,0.05;%s;" % (kmer, ";".join(alts)))
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("k", [3, 5])
def test_table_alternatives_longer_than_a_warp_and_other_k(orc, k):
    fasta = helpers.random_fasta(orc, [120_000])
    spec = dict(error_model=_long_alt_model(k), length=(4000.0, 3000.0), identity=(80.0, 95.0, 6.0), glitch=(800.0, 30.0, 30.0))
    _, ctx, batch, n = _run_pair(orc, fasta, spec, 600_000)
    ctx.close()


def _packed_edge_model(k, seed=9):
    """A model that fits the packed records (k <= 7, alternatives <= 15 bytes) and exercises their edges: alternatives of
    10..15 unrelated bases (more than 16 ops against the key: the record's entry is invalid and the piece is aligned like any
    other merge), empty and one-base alternatives, keys with more than 32 alternatives (a second chunk of partial sums), and
    k-mers that are missing from the table."""
    rnd = random.Random(seed)
    lines = []
    for code in range(4 ** k):
        if rnd.random() < 0.2:
            continue  # not in the table: random_error
        kmer = "".join("ACGT"[(code >> (2 * (k - 1 - i))) & 3] for i in range(k))
        alts = []
        na = rnd.choice([1, 2, 3, 5, 40])
        for _ in range(na):
            r = rnd.random()
            if r < 0.25:
                a = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(10, 15)))
            elif r < 0.35:
                a = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(0, 1)))
            elif r < 0.5:
                a = kmer + "".join(rnd.choice("ACGT") for _ in range(rnd.randint(1, 15 - k)))
            else:
                a = list(kmer)
                a[rnd.randrange(k)] = rnd.choice("ACGT")
                if rnd.random() < 0.5:
                    a.insert(rnd.randrange(k), rnd.choice("ACGT"))
                else:
                    del a[rnd.randrange(len(a))]
                a = "".join(a)
            alts.append("%s,%s" % (a, repr(round(0.9 / na, 6))))
        lines.append("%s,0.05;%s;" % (kmer, ";".join(alts)))
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("k", [4, 7])
def test_packed_error_model_edges(orc, k):
    fasta = helpers.random_fasta(orc, [150_000])
    spec = dict(error_model=_packed_edge_model(k), length=(4000.0, 3000.0), identity=(78.0, 95.0, 6.0), glitch=(900.0, 30.0, 30.0))
    _, ctx, batch, n = _run_pair(orc, fasta, spec, 700_000)
    ctx.close()


def test_qscore_window_panic_is_reported_for_the_same_read(orc):
    """Reads of a handful of bases with a k = 3 model and 9-wide qscore windows: `generate_quality` slices its CIGAR out of range
    (src/simulate/quality.rs:22-34, usize arithmetic) and the reference panics.  The device reports exactly the first read the
    oracle fails on, names the place, and everything before that read is generated normally."""
    import re

    import rustyread_b200 as rr
    import rustyread_b200.capi as capi

    fasta = helpers.random_fasta(orc, [30_000])
    adapter = ("AATGTACTTCGTTCAGTTACGTATTGCT", "GCAATACGTAACTGAACGAAGT", 10.0, 60.0, 10.0, 20.0)
    spec = dict(seed=1, length=(12.0, 10.0), error_model=_long_alt_model(3, seed=1), qscore_model="ideal", identity=(70.0, 88.0, 9.0), adapter=adapter)
    w, ctx, s = helpers.make_pair(orc, fasta, spec)
    n = w.feed(3000)
    d, ident, _ = w.items()
    descs = capi.descs_from_oracle_items(d, ident)
    with pytest.raises(rr.RRError) as e:
        ctx.generate(descs)
    assert "generate_quality" in str(e.value)
    r = int(re.search(r"read (\d+) ", str(e.value)).group(1))
    w.load_items(d[r:r + 1], ident[r:r + 1])
    with pytest.raises(RuntimeError) as oe:
        w.generate(threads=1)
    assert "generate_quality" in str(oe.value)
    assert r > 0
    w.load_items(d[:r], ident[:r])
    w.generate(threads=4)  # no error before read r
    helpers.compare_batch(w, ctx.generate(descs[:r]), ctx)
    ctx.close()


def test_reference_with_bytes_outside_acgt(orc):
    # N runs, IUPAC codes and lower case pass through untouched (src/references.rs:162-166); k-mers that contain them miss the
    # error-model table and alignments that meet them leave the register aligner for the general one
    fasta = helpers.random_fasta(orc, [150_000])
    rnd = random.Random(3)
    lines = fasta.split("\n")
    for i in range(1, len(lines)):
        if lines[i] and rnd.random() < 0.2:
            row = list(lines[i])
            for _ in range(rnd.randint(1, 6)):
                row[rnd.randrange(len(row))] = rnd.choice("NnRYacgtKMW")
            lines[i] = "".join(row)
    error_text, qscore_text = _synthetic_models()
    for spec in (dict(length=(6000.0, 5000.0)), dict(error_model=error_text, qscore_model=qscore_text, length=(6000.0, 5000.0))):
        _, ctx, batch, n = _run_pair(orc, "\n".join(lines), spec, 900_000)
        ctx.close()


def test_ideal_qscore_and_no_glitch_no_adapter(orc):
    fasta = helpers.random_fasta(orc, [200_000])
    spec = dict(qscore_model="ideal", glitch=(0.0, 0.0, 0.0), adapter=("", "", 0.0, 0.0, 0.0, 0.0), rates=(0.0, 0.0, 0.0),
                length=(5000.0, 0.0), identity=(90.0, 90.0, 5.0))
    _, ctx, batch, n = _run_pair(orc, fasta, spec, 1_000_000)
    ctx.close()


def test_tiny_and_ragged_reads(orc, goldens):
    # reads of a handful of bases (the writer's `len <= 14` filter lives above this boundary)
    spec = dict(length=(3.0, 2.0), adapter=("TACGTATTGCT", "TACGTATTGCT", 100.0, 60.0, 100.0, 20.0), glitch=(10.0, 3.0, 3.0),
                rates=(20.0, 20.0, 50.0), adjusted=False)
    _, ctx, batch, n = _run_pair(orc, goldens["simulate.FASTA"], spec, 3_000, threads=2)
    assert n > 200
    ctx.close()


def test_run_is_repeatable_and_staged_api(orc):
    import rustyread_b200.capi as capi

    fasta = helpers.random_fasta(orc, [100_000])
    w, ctx, s = helpers.make_pair(orc, fasta, {})
    w.feed(500_000)
    d, ident, _ = w.items()
    descs = capi.descs_from_oracle_items(d, ident)
    ctx.upload(descs)
    ctx.run()
    a = ctx.download()
    sa, la = a.seq.copy(), a.lengths.copy()
    l0 = ctx.kernel_launches()
    ctx.run()
    b = ctx.download()
    assert ctx.kernel_launches() - l0 == 3
    assert np.array_equal(la, b.lengths) and np.array_equal(sa[:a.total_bytes], b.seq[:b.total_bytes])
    ms = ctx.last_run_ms()
    assert ms["total"] > 0 and ms["read_kernel"] > 0
    assert ctx.last_run_bases() == int(sum(int(x) - 14 for x in la if x > 14))
    ctx.close()


def test_invalid_inputs_fail_loudly(rr, orc):
    import rustyread_b200.capi as capi

    ctx = rr.Context(0)
    d = np.zeros(1, dtype=capi.DESC_DTYPE)
    with pytest.raises(rr.RRError):
        ctx.generate(d)  # models not set
    ctx.set_references([b"ACGT" * 10], [False])
    ctx.set_error_model_random(7)
    ctx.set_qscore_model_random()
    ctx.set_adapter(b"", b"", 0.0, 0.0, 0.0, 0.0)
    ctx.set_glitch(0.0, 0.0, 0.0)
    d["ref_idx"] = 3
    d["end"] = 5
    d["identity"] = 0.9
    with pytest.raises(rr.RRError):
        ctx.generate(d)  # ref_idx out of range
    d["ref_idx"] = 0
    d["end"] = 500
    with pytest.raises(rr.RRError):
        ctx.generate(d)  # slice past the contig: the reference panics
    with pytest.raises(rr.RRError):
        ctx.set_qscore_model(1, [(b"=", [1], [1.0])])  # missing X / I
    ctx.close()


def test_whole_simulate_fastq_is_byte_identical(orc):
    """Product path end to end (host feeder -> kernels -> FASTQ writer) vs the oracle's simulate():
    every byte of the FASTQ text, including read names and descriptions, over two --number_base_store batches."""
    import rustyread_b200.capi as capi

    fasta = helpers.random_fasta(orc, [80_000, 6_000, 900], circular={1, 2})
    params = dict(seed=11, length="2500,2200", identity="80,97,8", quantity="12x", junk_reads=4, random_reads=4, chimera=10,
                  glitches="800,25,25", threads=8, number_base_store="600K")
    text, stats = orc.simulate(params, fasta, capture=True)
    sim = capi.Simulation(fasta, seed=11, length=(2500, 2200), identity=(80, 97, 8), junk_reads=4, random_reads=4, chimera=10, glitches=(800, 25, 25))
    ctx = capi.Context(0)
    sim.attach(ctx)
    total, limit = sim.quantity("12x"), sim.quantity("600K")
    produced, out, recs = 0, b"", 0
    while produced < total:
        base_loop = min(limit, total - produced)
        produced += base_loop
        batch = ctx.generate(sim.feed(base_loop))
        t, r, _ = sim.fastq(batch.struct, 0)
        out += t
        recs += r
    assert recs == stats["reads"]
    assert out.decode() == text
    ctx.close()


def test_submit_collect_pipeline_matches_sync(orc):
    """rr_submit / rr_collect (double buffered) returns exactly what rr_generate returns, batch after batch."""
    import rustyread_b200.capi as capi

    fasta = helpers.random_fasta(orc, [150_000])
    sim = capi.Simulation(fasta, seed=5, length=(4000, 3500), glitches=(1000, 25, 25))
    ctx = capi.Context(0)
    sim.attach(ctx)
    batches = [sim.feed(300_000 + 50_000 * i) for i in range(4)]
    want = []
    for d in batches:
        b = ctx.generate(d)
        want.append((b.lengths.copy(), b.offsets.copy(), b.seq[:b.total_bytes].copy(), b.qual[:b.total_bytes].copy(), b.real_identity.copy()))
    ctx.submit(batches[0])
    for i in range(4):
        if i + 1 < 4:
            ctx.submit(batches[i + 1])
        b = ctx.collect()
        w = want[i]
        assert np.array_equal(b.lengths, w[0]) and np.array_equal(b.real_identity, w[4])
        for r in range(b.n):
            o, l = int(b.offsets[r]), int(b.lengths[r])
            wo = int(w[1][r])
            assert np.array_equal(b.seq[o:o + l], w[2][wo:wo + l]) and np.array_equal(b.qual[o:o + l], w[3][wo:wo + l])
    import rustyread_b200 as rr
    with pytest.raises(rr.RRError):
        ctx.collect()  # nothing in flight
    ctx.submit(batches[0])
    ctx.submit(batches[1])
    with pytest.raises(rr.RRError):
        ctx.submit(batches[2])  # two already in flight
    ctx.collect()
    ctx.collect()
    ctx.close()


def test_streaming_drain_matches_whole_batch_collect(orc):
    """rr_collect_begin / rr_collect_part hand over exactly what rr_collect does, part by part (many parts forced)."""
    import rustyread_b200.capi as capi

    fasta = helpers.random_fasta(orc, [150_000])
    w, ctx, s = helpers.make_pair(orc, fasta, dict(length=(3000.0, 2500.0)))
    w.feed(1_500_000)
    d, ident, _ = w.items()
    descs = capi.descs_from_oracle_items(d, ident)
    whole = ctx.generate(descs)
    want = [(bytes(whole.seq[int(o):int(o) + int(l)]), bytes(whole.qual[int(o):int(o) + int(l)]), float(i))
            for o, l, i in zip(whole.offsets[:-1], whole.lengths, whole.real_identity)]
    for part_bytes in (40_000, 7, 1 << 30):  # several reads per part; one read per part; one part
        ctx.debug_set_part_bytes(part_bytes)
        ctx.submit(descs)
        ctx.submit(descs[: len(descs) // 2])  # a second batch in flight while the first is drained
        got, nparts = [], 0
        for first, b in ctx.collect_parts():
            assert first == len(got)
            nparts += 1
            got += [(bytes(b.seq[int(o):int(o) + int(l)]), bytes(b.qual[int(o):int(o) + int(l)]), float(i))
                    for o, l, i in zip(b.offsets[:-1], b.lengths, b.real_identity)]
        assert got == want
        assert nparts == (1 if part_bytes == 1 << 30 else nparts) and (part_bytes != 7 or nparts == len(descs))
        second = ctx.collect()
        assert second.n == len(descs) // 2 and np.array_equal(second.lengths, whole.lengths[: len(descs) // 2])
    with pytest.raises(capi.RRError):
        next(ctx.collect_parts())  # nothing submitted
    ctx.close()


def test_collect_meta_sizes_records_before_the_first_part(orc, tmp_path):
    """What the parallel multi-rank writer relies on, on one GPU: after rr_collect_begin, rr_collect_meta gives the whole
    batch's lengths / identities, rr_sim_fastq_bytes sizes the records from them without touching main_rng, and the parts
    written with writev at an offset of a shared file add up to exactly that many bytes = the text rr_sim_fastq produces."""
    import os

    import rustyread_b200.capi as capi

    fasta = helpers.random_fasta(orc, [90_000, 4_000], circular={1})
    sim = capi.Simulation(fasta, seed=11, length=(2500, 2000), junk_reads=4, random_reads=4, chimera=8)
    ctx = capi.Context(0)
    sim.attach(ctx)
    ctx.debug_set_part_bytes(50_000)
    descs = sim.feed(900_000)
    lo = len(descs) // 3  # a shard that does not start at item 0
    ctx.submit(descs[lo:])
    n, bases, lens, ident = ctx.collect_begin()
    assert n == len(descs) - lo
    words = sim.rng_words()
    nbytes, recs, nb = sim.fastq_bytes(lens, ident, first=lo)
    assert sim.rng_words() == words and nb == bases and recs == int(np.count_nonzero(lens > 14))
    path = tmp_path / "shared.fq"
    prefix = b"x" * 123  # stands for the records of the ranks before this one
    path.write_bytes(prefix)
    fd = os.open(path, os.O_WRONLY)
    os.lseek(fd, len(prefix), os.SEEK_SET)
    wrote = 0
    for first, part in ctx.parts():
        wrote += sim.write_fastq(part.struct, fd, first=lo + first)[2]
    os.close(fd)
    assert wrote == nbytes and os.path.getsize(path) == len(prefix) + nbytes
    # the same records through the in-memory formatter (a twin simulation: same seed, same main_rng position)
    sim2 = capi.Simulation(fasta, seed=11, length=(2500, 2000), junk_reads=4, random_reads=4, chimera=8)
    assert np.array_equal(sim2.feed(900_000), descs)
    whole = ctx.generate(descs[lo:])
    assert np.array_equal(whole.lengths, lens) and np.array_equal(whole.real_identity, ident)
    text, r2, b2 = sim2.fastq(whole.struct, first=lo)
    assert (len(text), r2, b2) == (nbytes, recs, nb)
    assert path.read_bytes() == prefix + text
    with pytest.raises(capi.RRError):
        ctx.collect_begin()  # nothing submitted
    ctx.close()


@pytest.mark.parametrize("quantity,store", [("15x", "500K"), ("90x", None)])
def test_cli_simulate_matches_oracle(orc, tmp_path, quantity, store):
    """`python -m rustyread_b200 simulate` (the drop-in CLI) writes the same file as the oracle's simulate(); the second case is
    large enough (17 MB of FASTQ in one part) for the threaded writer, which maps the output file and copies into it."""
    import subprocess
    import sys

    fasta = helpers.random_fasta(orc, [90_000, 7_000], circular={1})
    ref = tmp_path / "ref.fasta"
    ref.write_text(fasta)
    out = tmp_path / "reads.fastq"
    cmd = [sys.executable, "-m", "rustyread_b200", "simulate", "--reference", str(ref), "--quantity", quantity, "--seed", "42", "--error_model", "random",
           "--qscore_model", "random", "--output", str(out), "--chimera", "5"] + (["--number_base_store", store] if store else [])
    env = dict(__import__("os").environ, RUSTYREAD_B200_WRITER_THREADS="4")
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=helpers.ROOT if hasattr(helpers, "ROOT") else None, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    params = dict(seed=42, quantity=quantity, chimera=5, threads=8)
    if store:
        params["number_base_store"] = store
    text, stats = orc.simulate(params, fasta, capture=True)
    assert out.read_text() == text and stats["reads"] > 50


def test_identity_edge_cases_and_stuck_glitches(orc):
    """identity 0.0 drives gen_bool with p == 1.0 (no draw consumed, rand Bernoulli ALWAYS_TRUE) until the budget
    moves off the diagonal; identity 1.0 is a walk with p == 0 at every position; a glitch rate of 2 makes
    zero-distance glitches (which block all later ones) the common case."""
    import rustyread_b200.capi as capi

    fasta = helpers.random_fasta(orc, [60_000])
    w, ctx, s = helpers.make_pair(orc, fasta, dict(length=(1500.0, 1200.0), glitch=(2.0, 3.0, 3.0), rates=(5.0, 5.0, 20.0)))
    n = w.feed(300_000)
    d, ident, _ = w.items()
    for i in range(n):
        if i % 3 == 0:
            ident[i] = 0.0
        elif i % 3 == 1:
            ident[i] = 1.0
        else:
            ident[i] = 0.5
        w.set_item_identity(i, float(ident[i]))
    batch = ctx.generate(capi.descs_from_oracle_items(d, ident))
    w.generate(threads=8)
    helpers.compare_batch(w, batch, ctx)
    assert n > 100
    ctx.close()


def test_multi_contig_genome_scale_layout(orc):
    """BASELINE config 3 in miniature: 24 linear contigs with human-like relative lengths (3.1 Mbp total), 4x."""
    rel = [248, 242, 198, 190, 181, 171, 159, 145, 138, 133, 135, 133, 114, 107, 102, 90, 83, 80, 58, 64, 46, 50, 156, 57]
    fasta = helpers.random_fasta(orc, [r * 1000 for r in rel])
    _, ctx, batch, n = _run_pair(orc, fasta, {}, 12_000_000)
    assert n > 500
    ctx.close()


def test_pipelined_chunks_and_writev_writer(orc, tmp_path):
    """simulate() on one GPU: a reference batch split into several device chunks (double buffered), records written
    with writev straight from pinned memory; the file must equal the oracle's simulate() text."""
    from rustyread_b200 import simulate as S

    fasta = helpers.random_fasta(orc, [120_000, 9_000], circular={1})
    out = tmp_path / "reads.fastq"
    with open(out, "wb") as f:
        reads, bases = S.simulate(fasta, "20x", f, seed=3, error_model="random", qscore_model="ideal", chimera=8.0, glitches="700,20,20",
                                  number_base_store="1200K", chunk_bases=150_000)
    text, stats = orc.simulate(dict(seed=3, quantity="20x", chimera=8, glitches="700,20,20", number_base_store="1200K", threads=8, qscore_model="ideal"), fasta, capture=True)
    assert (reads, bases) == (stats["reads"], stats["bases"])
    assert out.read_text() == text


def test_capacity_overflow_is_replanned_not_truncated(orc):
    """With an absurdly small planning slack the first attempts overflow (output and glitch scratch); the library must
    re-plan with more slack until everything fits and still return the exact bytes."""
    import rustyread_b200.capi as capi

    fasta = helpers.random_fasta(orc, [80_000])
    w, ctx, s = helpers.make_pair(orc, fasta, dict(length=(3000.0, 2500.0), glitch=(40.0, 30.0, 5.0), identity=(70.0, 90.0, 10.0)))
    ctx.debug_set_slack(0.01)
    n = w.feed(600_000)
    d, ident, _ = w.items()
    batch = ctx.generate(capi.descs_from_oracle_items(d, ident))
    assert ctx.debug_set_slack() >= 1  # at least one re-plan happened
    w.generate(threads=8)
    helpers.compare_batch(w, batch, ctx)
    ctx.close()


@pytest.mark.parametrize("mode", ["parallel", "parallel-rounds", "single", "every-rank-parses"])
def test_two_gpu_cli_matches_oracle(orc, rr, tmp_path, mode):
    """torchrun with 2 ranks (one process per GPU): sharded generation; FASTQ == the oracle's simulate() with the parallel
    writer (every rank streams its pieces to their final offsets; one or several rounds per batch) and with the single writer.
    Rank 1 gets the reference from rank 0's GPU (contig index + one NCCL broadcast) except in the last mode, where it parses the
    FASTA itself."""
    import os
    import subprocess
    import sys

    if rr.load_library().device_count() < 2:
        pytest.skip("needs 2 GPUs")
    fasta = helpers.random_fasta(orc, [150_000, 8_000], circular={1})
    ref = tmp_path / "ref.fasta"
    ref.write_text(fasta)
    out = tmp_path / "reads.fastq"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           "-m", "rustyread_b200", "simulate", "--reference", str(ref), "--quantity", "12x", "--seed", "21", "--error_model", "random",
           "--qscore_model", "random", "--output", str(out), "--number_base_store", "700K"]
    env = dict(os.environ)
    if mode == "single":
        env["RUSTYREAD_B200_SINGLE_WRITER"] = "1"
    if mode == "parallel-rounds":
        env["RUSTYREAD_B200_CHUNK_BASES"] = "90000"
    if mode == "every-rank-parses":
        env["RUSTYREAD_B200_NO_SHARED_REFERENCE"] = "1"
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=helpers.ROOT, timeout=900, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    text, stats = orc.simulate(dict(seed=21, quantity="12x", number_base_store="700K", threads=8), fasta, capture=True)
    assert out.read_text() == text and stats["reads"] > 50
