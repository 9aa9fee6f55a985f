"""CPU tests of the product's host side (feeder, loaders, reference ingest, FASTQ writer) against the oracle,
and of the C-ABI surface.  No compute entry point is called here (there is no GPU on this box)."""
import ctypes as C
import os
import random
import re

import numpy as np
import pytest

import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def rr():
    import rustyread_b200

    rustyread_b200.build_library()
    return rustyread_b200


def test_library_exports_every_declared_symbol(rr):
    lib = rr.load_library()
    declared = set()
    for h in ("rustyread_b200.h", "rustyread_b200_host.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        declared |= set(re.findall(r"\b(rr_[a-z0-9_]+)\s*\(", text))
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(lib.dll, name), name
    import rustyread_b200.capi as capi

    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)


def test_no_cpu_fallback(rr):
    lib = rr.load_library()
    if lib.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(rr.RRError) as e:
        rr.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_desc_struct_layout(rr):
    import rustyread_b200.capi as capi

    assert capi.DESC_DTYPE.itemsize == 80
    assert capi.DESC_DTYPE.fields["start"][1] == 32 and capi.DESC_DTYPE.fields["seed"][1] == 72


def test_quantity(rr):  # src/cli/simulate.rs:247-290
    q = lambda s, g=1000: rr.load_library().dll.rr_quantity_to_bases(s.encode(), g)
    assert (q("50"), q("50K"), q("50M"), q("50G"), q("50x")) == (50, 50_000, 50_000_000, 50_000_000_000, 50_000)
    assert q("bépo") == -1 and q("50k") == -1 and q("x") == -1 and q("") == -1 and q("1.5x") == -1


SPECS = {
    "default": ({}, [300_000], set(), 3_000_000),
    "cfg4": (dict(length=(50000.0, 40000.0), identity=(75.0, 90.0, 8.0), glitch=(2000.0, 50.0, 50.0), rates=(1.0, 1.0, 5.0)), [500_000], set(), 5_000_000),
    "plasmids_bias": (dict(rates=(10.0, 10.0, 1.0), small_plasmid_bias=True, length=(6000.0, 5000.0)), None, None, 2_000_000),
    "plasmids_nobias": (dict(rates=(10.0, 10.0, 20.0), small_plasmid_bias=False, length=(6000.0, 5000.0)), None, None, 2_000_000),
    "fixed": (dict(length=(5000.0, 0.0), identity=(90.0, 90.0, 5.0), rates=(0.0, 0.0, 0.0)), [100_000, 20_000], {1}, 1_000_000),
    "small_shape": (dict(length=(3000.0, 6000.0)), [100_000], set(), 1_000_000),  # Gamma shape < 1 (unpinned branch)
}


def _fasta(orc, name):
    spec, contigs, circ, target = SPECS[name]
    if contigs is None:
        rnd = random.Random(5)
        contigs = [int(2000 * (25 ** rnd.random())) for _ in range(40)]
        return helpers.random_fasta(orc, contigs, circular=set(range(0, 40, 2)), depth={3: "2.5", 8: "0.5"})
    return helpers.random_fasta(orc, contigs, circular=circ)


def _sim(rr, fasta, s):
    import rustyread_b200.capi as capi

    a = s["adapter"]
    return capi.Simulation(fasta, seed=s["seed"], length=s["length"], identity=s["identity"], start_adapter=(a[2], a[3]), end_adapter=(a[4], a[5]),
                           start_adapter_seq=a[0], end_adapter_seq=a[1], junk_reads=s["rates"][0], random_reads=s["rates"][1], chimera=s["rates"][2],
                           glitches=s["glitch"], small_plasmid_bias=s["small_plasmid_bias"],
                           error_model=s["error_model"] or "random", qscore_model=s["qscore_model"])


def _oracle_workload(orc, fasta, spec):
    s = dict(helpers.DEFAULT)
    s.update(spec)
    w = orc.workload(s["seed"])
    w.set_length(*s["length"])
    w.set_references(fasta, adjusted=True, small_plasmid_bias=s["small_plasmid_bias"])
    w.set_identity(*s["identity"])
    w.set_adapter(*s["adapter"])
    w.set_glitch(*s["glitch"])
    w.set_error_model(s["error_model"])
    w.set_qscore_model(s["qscore_model"] if s["qscore_model"] in ("random", "ideal") else "file", None if s["qscore_model"] in ("random", "ideal") else s["qscore_model"])
    w.set_rates(*s["rates"])
    return w, s


@pytest.mark.parametrize("name", sorted(SPECS))
def test_feeder_matches_oracle(rr, orc, name):
    import rustyread_b200.capi as capi

    spec, _, _, target = SPECS[name]
    fasta = _fasta(orc, name)
    w, s = _oracle_workload(orc, fasta, spec)
    sim = _sim(rr, fasta, s)
    # reference ingest: ids, bytes, circular flags and adjusted weights
    for a, b in zip(sim.refs(), w.refs()):
        assert a == b
    for batch in range(2):  # two consecutive batches share main_rng (--number_base_store semantics)
        n = w.feed(target)
        d, ident, _ = w.items()
        want = capi.descs_from_oracle_items(d, ident)
        got = sim.feed(target)
        assert len(got) == n
        for f in capi.DESC_DTYPE.names:
            if f in ("reserved",):
                continue
            assert np.array_equal(got[f], want[f]), (name, batch, f)
    assert sim.rng_words() == orc.lib.orc_rng_words(w.rng_handle())


def test_feeder_with_file_error_model_draws(rr, orc, goldens):
    # Error::from_stream consumes main_rng (src/model/error.rs:65-69) before the first fragment is drawn
    import rustyread_b200.capi as capi

    fasta = helpers.random_fasta(orc, [50_000])
    spec = dict(error_model=goldens["model.error.MODEL"], qscore_model=goldens["model.quality.MODEL"], length=(3000.0, 2500.0))
    w, s = _oracle_workload(orc, fasta, spec)
    sim = _sim(rr, fasta, s)
    assert sim.k == 7
    n = w.feed(500_000)
    d, ident, _ = w.items()
    got = sim.feed(500_000)
    want = capi.descs_from_oracle_items(d, ident)
    assert np.array_equal(got["seed"], want["seed"]) and np.array_equal(got["identity"], want["identity"])


def _batch_from_oracle(w, k):
    import rustyread_b200.capi as capi

    seq, so = w.out("seq")
    qual, _ = w.out("qual")
    lens = (so[1:] - so[:-1]).astype(np.uint64)
    return capi.batch_struct_from_arrays(so, lens, seq, qual, w.out_identity(), k)


@pytest.mark.parametrize("store", ["", "150K"])
def test_fastq_writer_matches_oracle_simulate(rr, orc, store):
    # whole-run text: names (UUIDv3 of main_rng draws), descriptions, trimming, the len <= 14 filter, batching
    fasta = helpers.random_fasta(orc, [60_000, 4_000], circular={1})
    params = dict(seed=7, length="2000,1800", identity="85,95,5", quantity="8x", junk_reads=5, random_reads=5, chimera=10,
                  glitches="1000,25,25", threads=4)
    if store:
        params["number_base_store"] = store
    text, stats = orc.simulate(params, fasta, capture=True)
    spec = dict(seed=7, length=(2000.0, 1800.0), rates=(5.0, 5.0, 10.0), glitch=(1000.0, 25.0, 25.0))
    w, s = _oracle_workload(orc, fasta, spec)
    sim = _sim(rr, fasta, s)
    total = sim.quantity("8x")
    limit = sim.quantity(store) if store else total
    produced, out, recs, bases = 0, b"", 0, 0
    while produced < total:  # the batch loop of src/simulate/mod.rs:153-214
        base_loop = min(limit, total - produced)
        produced += base_loop
        descs = sim.feed(base_loop)
        n = w.feed(base_loop)
        assert n == len(descs)
        w.generate(threads=4, keep_cigar_raw=False)
        # the oracle workload's main rng must follow the writer's u128 draws too: replay them
        half = len(descs) // 2
        b = _batch_from_oracle(w, 7)
        # exercise sharded formatting: two shards in order
        import rustyread_b200.capi as capi

        seq, so = w.out("seq")
        qual, _ = w.out("qual")
        ident = w.out_identity()
        lens = (so[1:] - so[:-1]).astype(np.uint64)
        for lo, hi in ((0, half), (half, len(descs))):
            sb = capi.batch_struct_from_arrays(so[lo:hi + 1] - so[lo], lens[lo:hi], seq[int(so[lo]):int(so[hi])] if hi > lo else seq[:1], qual[int(so[lo]):int(so[hi])] if hi > lo else qual[:1], ident[lo:hi], 7)
            words = sim.rng_words()
            assert sim.fastq_bytes(lens[lo:hi], ident[lo:hi], first=lo)[1:] == (int(np.count_nonzero(lens[lo:hi] > 14)), int(sum(int(x) - 14 for x in lens[lo:hi] if x > 14)))
            assert sim.rng_words() == words  # sizing draws nothing from main_rng
            t, r, bs = sim.fastq(sb, first=lo)
            assert sim.fastq_bytes(lens[lo:hi], ident[lo:hi], first=lo) == (len(t), r, bs)
            out += t
            recs += r
            bases += bs
            for _ in range(r):  # keep the oracle-side main rng in step (2 u64 per emitted read)
                orc.lib.orc_rng_next_u64(w.rng_handle())
                orc.lib.orc_rng_next_u64(w.rng_handle())
    assert recs == stats["reads"] and bases == stats["bases"]
    assert out.decode() == text


def test_fragment_that_fits_no_contig_fails_loudly_instead_of_hanging(rr, orc):
    """fragments.rs:80-85 re-draws the contig until the fragment fits; with only circular contigs that are all shorter than
    the drawn length the reference never leaves that loop (BASELINE config 5 with the default --length does this within
    ~100 reads).  Product feeder and oracle both refuse loudly; with a length model that fits they agree item by item."""
    import rustyread_b200.capi as capi

    fasta = helpers.random_fasta(orc, [3000, 5000, 8000], circular={0, 1, 2})
    sim = capi.Simulation(fasta, seed=1, length=(15000, 13000), small_plasmid_bias=True)
    with pytest.raises(capi.RRError, match="fits no contig"):
        sim.feed(2_000_000)
    w = orc.workload(1)
    w.set_length(15000.0, 13000.0)
    w.set_references(fasta, adjusted=True, small_plasmid_bias=True)
    w.set_identity(85.0, 95.0, 5.0)
    w.set_rates(1.0, 1.0, 1.0)
    with pytest.raises(RuntimeError, match="fits no contig"):
        w.feed(2_000_000)
    # one linear contig makes every length possible again
    fasta2 = helpers.random_fasta(orc, [3000, 5000, 8000], circular={0, 1})
    sim2 = capi.Simulation(fasta2, seed=1, length=(15000, 13000), small_plasmid_bias=True)
    w2 = orc.workload(1)
    w2.set_length(15000.0, 13000.0)
    w2.set_references(fasta2, adjusted=True, small_plasmid_bias=True)
    w2.set_identity(85.0, 95.0, 5.0)
    w2.set_rates(1.0, 1.0, 1.0)
    d = sim2.feed(300_000)
    n = w2.feed(300_000)
    od, oident, _ = w2.items()
    pd, pident = capi.descs_to_item_table(d)
    assert n == len(d) > 20 and np.array_equal(od, pd) and np.array_equal(oident, pident)


def test_loader_rejects_what_the_reference_rejects(rr, orc):
    import rustyread_b200.capi as capi

    fasta = helpers.random_fasta(orc, [10_000])
    with pytest.raises(rr.RRError):
        capi.Simulation(fasta, length=(0, 0))
    with pytest.raises(rr.RRError):
        capi.Simulation(fasta, identity=(0, 0, 0))
    with pytest.raises(rr.RRError):
        capi.Simulation(b"")
    with pytest.raises(rr.RRError):
        capi.Simulation(fasta, qscore_model="overall;3;1:0.5,\n=;1;x:0.5,\n")


def test_session_from_contig_index_is_the_same_session(orc, rr):
    """rr_sim_create_indexed (a rank that receives the bases device to device): same main_rng schedule, same work items and
    the same FASTQ text as the session made from the FASTA itself."""
    import rustyread_b200.capi as capi

    fasta = helpers.random_fasta(orc, [40_000, 3_000, 9_000], circular={1}, depth={2: "2.5"})
    kw = dict(seed=11, length=(1500, 1300), junk_reads=5, random_reads=5, chimera=10, glitches=(800, 25, 25))
    a = capi.Simulation(fasta, **kw)
    idx = a.contig_index()
    assert [c[1] for c in idx] == [40_000, 3_000, 9_000] and [c[2] for c in idx] == [False, True, False]
    assert idx[2][3] == 2.5 * 9_000 and idx[0][0] == "contig_0"
    b = capi.Simulation(None, index=idx, **kw)
    assert b.genome_length == a.genome_length and a.rng_words() == b.rng_words()
    assert [(r["id"], r["circular"], r["weight"]) for r in a.refs()] == [(r["id"], r["circular"], r["weight"]) for r in b.refs()]
    rs = np.random.RandomState(4)
    for _ in range(2):
        da, db = a.feed(150_000), b.feed(150_000)
        assert len(da) > 20 and np.array_equal(da, db)
        # any results will do for the writer: the text depends on the session only through descriptors, ids and main_rng
        lens = (rs.randint(10, 400, size=len(da))).astype(np.uint64)
        off = np.concatenate([[0], np.cumsum((lens + 15) // 16 * 16)]).astype(np.uint64)
        seq = rs.choice(np.frombuffer(b"ACGT", np.uint8), size=int(off[-1]))
        qual = rs.randint(34, 70, size=int(off[-1])).astype(np.uint8)
        ident = rs.uniform(80, 99, size=len(da))
        ta = a.fastq(capi.batch_struct_from_arrays(off, lens, seq, qual, ident, a.k))
        tb = b.fastq(capi.batch_struct_from_arrays(off, lens, seq, qual, ident, b.k))
        assert ta == tb and ta[1] > 10
        assert a.rng_words() == b.rng_words()
    with pytest.raises(capi.RRError):
        capi.Simulation(None, index=[], **kw)


def test_threaded_file_writer_writes_the_same_bytes(orc, rr, tmp_path):
    """rr_sim_write_fastq: several threads sending record ranges to their offsets (regular file), one writev stream (pipe,
    append mode, one thread) and rr_sim_fastq's text are the same bytes; the file position ends after the part."""
    import rustyread_b200.capi as capi

    fasta = helpers.random_fasta(orc, [60_000])
    kw = dict(seed=5, length=(9000, 7000))
    rs = np.random.RandomState(8)

    def session():
        s = capi.Simulation(fasta, **kw)
        d = s.feed(12_000_000)
        return s, d

    s0, d = session()
    lens = rs.randint(0, 20_000, size=len(d)).astype(np.uint64)
    lens[::17] = rs.randint(0, 15, size=len(lens[::17]))  # reads of 14 bases or fewer are dropped (mod.rs:192-194)
    lens[(lens > 14) & (lens < 2 * s0.k)] = 2 * s0.k
    off = np.concatenate([[0], np.cumsum((lens + 15) // 16 * 16)]).astype(np.uint64)
    seq = rs.choice(np.frombuffer(b"ACGT", np.uint8), size=int(off[-1]))
    qual = rs.randint(34, 70, size=int(off[-1])).astype(np.uint8)
    ident = rs.uniform(80, 99, size=len(d))
    batch = lambda s: capi.batch_struct_from_arrays(off, lens, seq, qual, ident, s.k)
    text, recs, bases = s0.fastq(batch(s0))
    assert len(text) > (8 << 20) and recs > 500  # large enough for the threaded path
    for mode in ("threads", "one-thread", "append", "pipe", "offset"):
        s, _ = session()
        s.set_writer_threads(1 if mode == "one-thread" else 5)
        path = tmp_path / ("out_" + mode)
        if mode == "pipe":
            import threading

            r, w = os.pipe()
            got = []
            t = threading.Thread(target=lambda: got.append(os.fdopen(r, "rb").read()))
            t.start()
            res = s.write_fastq(batch(s), w)
            os.close(w)
            t.join()
            data = got[0]
        else:
            pre = b"HEADER\n" if mode == "offset" else b""
            fd = os.open(str(path), os.O_RDWR | os.O_CREAT | (os.O_APPEND if mode == "append" else 0), 0o644)
            os.write(fd, pre)
            res = s.write_fastq(batch(s), fd)
            assert os.lseek(fd, 0, os.SEEK_CUR) == len(pre) + len(text)
            os.write(fd, b"TAIL")
            os.close(fd)
            data = path.read_bytes()
            assert data.startswith(pre) and data.endswith(b"TAIL")
            data = data[len(pre):-4]
        assert res == (recs, bases, len(text)), mode
        assert data == text, mode
        assert s.rng_words() == s0.rng_words()
