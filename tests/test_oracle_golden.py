"""Pins the CPU oracle against every golden vector the reference's own tests hold for the
`simulate` path.  Expected values come from
tests/golden/reference_goldens.json (extracted by tests/golden/extract_goldens.py from
/root/reference/src/**/*.rs `#[cfg(test)]` modules); a few small Change expectations that are
spread over many assert lines are transcribed inline with their file:line.
"""
import ctypes as C
import math

import numpy as np
import pytest

from oraclelib import _b, _buf


def f(x):
    return float(x["f64"]) if isinstance(x, dict) else float(x)


def parse_changes(text):
    out = []
    for line in text.splitlines():
        b, e, seq, cig, ed = line.split("\t")
        out.append(dict(begin=int(b), end=int(e), seq=seq, cigar=cig, edit_distance=int(ed)))
    return out


# ---- Appendix C #1, #2: ChaCha12 + seeding + gen_range(u64) -------------------------------------
def test_random_base(orc, goldens):  # src/lib.rs:85-97
    r = orc.rng(42)
    got = [orc.lib.orc_random_base(r.h) for _ in range(20)]
    assert got == goldens["lib.random_base"]


def test_random_base_diff(orc, goldens):  # src/lib.rs:99-135
    r = orc.rng(42)
    g = goldens["lib.random_base_diff"]
    for nuc, truth in zip(g["nucs"], g["truth"]):
        assert [orc.lib.orc_random_base_diff(r.h, ord(nuc)) for _ in range(20)] == truth


def test_random_seq(orc, goldens):  # src/lib.rs:137-141, src/simulate/mod.rs:363-371
    assert orc.rng(42).random_seq(20).decode() == goldens["lib.random_seq"]
    assert orc.rng(42).random_seq(100).decode() == goldens["simulate.random_seq"]


# ---- #3: gen::<f64> ---------------------------------------------------------------------------
def test_identity2edit(orc, goldens):  # src/simulate/error.rs:623-643
    r = orc.rng(42)
    got = [orc.lib.orc_number_of_edit(r.gen_f64(), 100) for _ in range(10)]
    assert got == [f(x) for x in goldens["simulate.error.identity2edit"]]


# ---- #4: mixed u32/u64 draws ----------------------------------------------------------------------
def test_random_error(orc, goldens):  # src/model/error.rs:142-154
    r = orc.rng(42)
    out = (C.c_uint8 * 16)()
    for want, kmer in goldens["model.error.random_error"]:
        n = orc.lib.orc_random_error(r.h, _buf(kmer), len(kmer), out)
        assert bytes(out[:n]).decode() == want


def test_error_model_random(orc, goldens):  # src/model/error.rs:197-221
    r = orc.rng(42)
    m = orc.lib.orc_error_model_random(7)
    g = goldens["model.error.random"]
    out = (C.c_uint8 * 16)()
    for want, _edit in g["truth"]:
        n = orc.lib.orc_error_model_add(m, r.h, _buf(g["kmer"]), 7, out)
        assert bytes(out[:n]).decode() == want
    orc.lib.orc_error_model_free(m)


# ---- #5: WeightedIndex -------------------------------------------------------------------------------
def test_read_error_model(orc, goldens):  # src/model/error.rs:158-195
    r = orc.rng(42)
    m = orc.lib.orc_error_model_from_text(goldens["model.error.MODEL"].encode(), r.h)
    assert m, orc.err()
    assert orc.lib.orc_error_model_k(m) == 7
    out = (C.c_uint8 * 16)()
    for kmer, want, _edit in goldens["model.error.read_error_model"]:
        n = orc.lib.orc_error_model_add(m, r.h, _buf(kmer), len(kmer), out)
        assert bytes(out[:n]).decode() == want
    # edit distances computed at load (error.rs:55-59): ACGGTTG 1, ACAGG 2, random alt 1
    dump = orc.take_str(orc.lib.orc_error_model_dump(m))
    edits = {a.split(":")[0]: int(a.split(":")[1]) for a in dump.strip().split("\t")[1].split(",")}
    for _kmer, want, edit in goldens["model.error.read_error_model"]:
        assert edits[want] == edit
    orc.lib.orc_error_model_free(m)


@pytest.mark.parametrize("kind", ["file", "random", "ideal"])
def test_quality_models(orc, goldens, kind):  # src/model/quality.rs:152-206, 208-262, 264-318
    r = orc.rng(42)
    text = goldens["model.quality.MODEL"].encode() if kind == "file" else None
    m = orc.lib.orc_quality_model_new(kind.encode(), text)
    assert m, orc.err()
    key = {"file": "model.quality.read_quality_model", "random": "model.quality.random", "ideal": "model.quality.ideal"}[kind]
    for c, truth in zip("=IX", goldens[key]):
        got = [orc.lib.orc_get_qscore(m, r.h, _buf(c), 1) for _ in range(100)]
        assert got == truth
    q = lambda s: orc.lib.orc_get_qscore(m, r.h, _buf(s), len(s))
    assert q("==") < 0 and q("===") >= 33 and q("=I=") >= 33 and q("=X=") >= 33
    assert q("") < 0 and q("bepo") < 0
    orc.lib.orc_quality_model_free(m)


# ---- #6: gen_bool, strand u32 draw, contig weights --------------------------------------------------------
def _wl(orc, fasta, adjusted=False, length=(8.0, 2.0), identity=(85.0, 95.0, 5.0), seed=42, bias=False):
    w = orc.workload(seed)
    w.set_length(*length)
    w.set_identity(*identity)
    w.set_references(fasta, adjusted=adjusted, small_plasmid_bias=bias)
    return w


def test_references_read(orc, goldens):  # src/references.rs:202-269
    w = _wl(orc, goldens["references.FASTA"])
    refs = w.refs()
    truth = goldens["references.read_reference"]
    assert len(refs) == len(truth)
    for got, want in zip(refs, truth):
        assert got["id"] == want["id"] and got["circular"] == want["circular"]
        assert list(got["seq"]) == want["seq"]
        rc = (C.c_uint8 * len(got["seq"]))()
        orc.lib.orc_revcomp(_buf(got["seq"]), len(got["seq"]), rc)
        assert list(rc) == want["revcomp"]


@pytest.mark.parametrize("adjusted,key", [(False, "references.get_reference"), (True, "references.read_reference_adjust")])
def test_choose_reference(orc, goldens, adjusted, key):  # src/references.rs:272-293, 296-324
    w = _wl(orc, goldens["references.FASTA"], adjusted=adjusted, length=(10.0, 8.0) if adjusted else (8.0, 2.0))
    got = []
    idx, strand = C.c_uint64(), C.c_int()
    for _ in range(10):
        orc.lib.orc_workload_choose_reference(w.h, w.rng_handle(), C.byref(idx), C.byref(strand))
        got.append([idx.value, chr(strand.value)])
    assert got == goldens[key]


# ---- #7: Gamma / ziggurat ------------------------------------------------------------------------------------
def test_length_values(orc, goldens):  # src/model/length.rs:64-80
    r = orc.rng(42)
    g = goldens["model.length.get_value"]
    m = orc.lib.orc_length_new(*g["params"])
    assert [orc.lib.orc_length_get(m, r.h) for _ in range(100)] == g["truth"]
    orc.lib.orc_length_free(m)


def _stats(v):
    s = 0.0
    for x in v:
        s += x
    avg = s / len(v)
    acc = 0.0
    for x in v:
        acc += (x - avg) ** 2.0
    return s, avg, math.sqrt(acc / len(v))


def test_length_shape(orc, goldens):  # src/model/length.rs:83-113
    r = orc.rng(42)
    want = goldens["model.length.shape"]
    for i, params in enumerate([(15000.0, 13000.0), (18200.0, 15500.0)]):
        m = orc.lib.orc_length_new(*params)
        data = [float(orc.lib.orc_length_get(m, r.h)) for _ in range(100_000)]
        assert list(_stats(data)) == want[3 * i:3 * i + 3]
        orc.lib.orc_length_free(m)


# ---- #8: Beta BB --------------------------------------------------------------------------------------------------
def test_identity_values(orc, goldens):  # src/model/identity.rs:82-202
    r = orc.rng(42)
    m = orc.lib.orc_identity_new(85.0, 85.0, 5.0)
    assert abs(0.85 - orc.lib.orc_identity_get(m, r.h)) < 2.3e-16
    m2 = orc.lib.orc_identity_new(85.0, 95.0, 5.0)
    got = [orc.lib.orc_identity_get(m2, r.h) for _ in range(100)]
    assert got == [f(x) for x in goldens["model.identity.get_value"]]


def test_identity_shape(orc, goldens):  # src/model/identity.rs:205-243
    r = orc.rng(42)
    want = goldens["model.identity.shape"]
    for i, params in enumerate([(90.0, 100.0, 5.0), (87.0, 100.0, 9.0), (75.0, 100.0, 15.0)]):
        m = orc.lib.orc_identity_new(*params)
        data = [orc.lib.orc_identity_get(m, r.h) for _ in range(100_000)]
        assert list(_stats(data)) == want[3 * i:3 * i + 3]


# ---- #9: Beta BB + BC via adapters -----------------------------------------------------------------------------------
def test_adapter(orc, goldens):  # src/model/adapter.rs:111-177
    g = goldens["model.adapter.generate"]
    r = orc.rng(42)
    m = orc.lib.orc_adapter_new(g["start"].encode(), g["end"].encode(), *g["params"])
    out = (C.c_uint8 * 64)()
    starts = [bytes(out[:orc.lib.orc_adapter_get(m, 0, r.h, out)]).decode() for _ in range(20)]
    ends = [bytes(out[:orc.lib.orc_adapter_get(m, 1, r.h, out)]).decode() for _ in range(20)]
    assert starts == g["starts"] and ends == g["ends"]
    assert orc.lib.orc_adapter_max_len(m) == 20


# ---- #10: Geometric + random_seq -----------------------------------------------------------------------------------------
def test_glitch_values(orc, goldens):  # src/model/glitch.rs:81-220
    g = goldens["model.glitch.get_value"]
    r = orc.rng(42)
    m = orc.lib.orc_glitch_new(*g["params"])
    out = (C.c_uint8 * 4096)()
    b, e = C.c_uint64(), C.c_uint64()
    for wb, we, wseq in g["truth"]:
        n = orc.lib.orc_glitch_get(m, r.h, C.byref(b), C.byref(e), out, 4096)
        assert (b.value, e.value, list(out[:n])) == (wb, we, wseq)
    for params in [(10000.0, 25.0, 0.0), (10000.0, 0.0, 25.0), (0.0, 25.0, 25.0), (0.0, 0.0, 0.0)]:  # glitch.rs:73-79
        assert orc.lib.orc_glitch_new(*params)


# ---- #11: Levenshtein / aligner tie-break ------------------------------------------------------------------------------------
def test_alignment(orc, goldens):  # src/alignment.rs:47-86
    for a, b, d in goldens["alignment.edit_distance"]:
        assert orc.edit_distance(a, b) == d
    for q, t, e, cigar in goldens["alignment.align"]:
        assert orc.align(q, t) == (e, cigar)


# ---- #12, #13: Change build / contain / overlap / merge / add_change --------------------------------------------------------------
def _script(orc, raw, script):
    raw = _b(raw)
    text = orc.take_str(orc.lib.orc_change_script(_buf(raw), len(raw), script.encode()))
    res, stack, lst = text.split("--\n")
    return res.split(), parse_changes(stack), parse_changes(lst)


def test_build_change(orc):  # src/simulate/error.rs:317-345
    raw = "GGCGGTGTCCTTGCTAACAT"
    _, st, _ = _script(orc, raw, "from_seq 3 10 GGTAACC\nfrom_seq 6 13 GTTTG\nfrom_seq 6 13 GTCAACTTG\n")
    assert st[0] == dict(begin=3, end=10, seq="GGTAACC", cigar="===XX==", edit_distance=2)
    assert st[1] == dict(begin=6, end=13, seq="GTTTG", cigar="==DD===", edit_distance=2)
    assert st[2] == dict(begin=6, end=13, seq="GTCAACTTG", cigar="===II====", edit_distance=2)


def test_contain_overlap(orc):  # src/simulate/error.rs:347-381
    res, _, _ = _script(orc, "TCAGGAAGATCCACA", "from_seq 3 15 GGCCGAT\nfrom_seq 6 13 AGACCA\ncontain\n")
    assert res == ["1"]
    res, _, _ = _script(orc, "TTTGCACTTGTTGCCACAGG", "from_seq 2 9 TGCCCTT\nfrom_seq 11 18 TGCCTCA\ncontain\noverlap\n")
    assert res == ["0", "0"]
    res, _, _ = _script(orc, "TCAGGAAGATCCACA", "from_seq 3 10 GGCCGAT\nfrom_seq 6 13 AGACCA\noverlap\n")
    assert res == ["1"]


def test_merge_change(orc):  # src/simulate/error.rs:383-442
    _, st, _ = _script(orc, "TCAGGAAGATCCACA", "from_seq 3 10 GGCCGAT\nfrom_seq 6 13 AGACCA\nmerge\n")
    assert st == [dict(begin=3, end=13, seq="GGCCGATCA", cigar="==XX===D==", edit_distance=3)]
    _, st, _ = _script(orc, "GTGACCGTTTCTC", "from_seq 2 9 GATT\nfrom_seq 6 13 GTAACTC\nmerge\n")
    assert st == [dict(begin=2, end=13, seq="GATTGTAACTC", cigar="==DDD==I=II===", edit_distance=6)]
    _, st, _ = _script(orc, "TTTGCACTTGTTGCCACAGG", "from_seq 2 9 TGCCCTT\nfrom_seq 11 18 TGCCTCA\nmerge\nfrom_seq 11 18 TGCCTCA\n")
    assert st[0] == dict(begin=2, end=9, seq="TGCCCTT", cigar="===X===", edit_distance=1)
    assert st[1] == dict(begin=11, end=18, seq="TGCCTCA", cigar="====X==", edit_distance=1)


def test_merge_containment(orc):  # src/simulate/error.rs:444-487
    s = "raw 2 9 TATGCGTC ===I==== 1\nraw 2 2 T I 1\nmerge\nraw 4 11 TGCG ====DDD 3\ncontain\nmerge\n"
    res, st, _ = _script(orc, "GCTATGCGTCAG", s)
    assert res[1] == "1"
    assert st == [dict(begin=2, end=9, seq="TATGCGTC", cigar="===I====", edit_distance=1)]


def test_changes(orc):  # src/simulate/error.rs:489-621
    items = [(37, 50, "A", 13), (90, 100, "B", 10), (10, 30, "C", 20), (60, 70, "D", 10), (25, 35, "E", 10), (30, 40, "F", 10),
             (90, 95, "G", 5), (105, 120, "H", 15), (90, 110, "I", 20), (140, 150, "J", 10), (135, 145, "K", 10)]
    items.sort(key=lambda x: x[0])  # stable, like sort_by_key
    script = "".join("raw %d %d %s - 12\nadd_change\n" % (b, e, c * n) for b, e, c, n in items)
    res, _, lst = _script(orc, "X" * 150, script)
    assert [float(x) for x in res] == [20.0, 5.0, 5.0, 10.0, 10.0, 10.0, 0.0, 10.0, 10.0, 10.0, 5.0]
    assert lst == [
        dict(begin=10, end=50, seq="C" * 20 + "E" * 5 + "F" * 5 + "A" * 10, cigar="X" * 40, edit_distance=40),
        dict(begin=60, end=70, seq="D" * 10, cigar="X" * 10, edit_distance=10),
        dict(begin=90, end=120, seq="B" * 10 + "I" * 10 + "H" * 10, cigar="X" * 30, edit_distance=30),
        dict(begin=135, end=150, seq="K" * 10 + "J" * 5, cigar="X" * 15, edit_distance=15),
    ]


# ---- #14, #15, #16: add_glitches, add_error, sequence -------------------------------------------------------------------------------
def test_add_glitches(orc, goldens):  # src/simulate/error.rs:646-713
    g = goldens["simulate.error.glitches"]
    r = orc.rng(42)
    m = orc.lib.orc_glitch_new(*g["params"])
    raw = r.random_seq(g["raw_len"])
    got = parse_changes(orc.take_str(orc.lib.orc_add_glitches(r.h, m, _buf(raw), len(raw), 1)))
    assert got == g["truth"]


def test_add_error(orc, goldens):  # src/simulate/error.rs:715-784
    g = goldens["simulate.error.errors"]
    r = orc.rng(42)
    m = orc.lib.orc_error_model_random(7)
    raw = r.random_seq(g["raw_len"])
    got = parse_changes(orc.take_str(orc.lib.orc_add_error(r.h, m, g["k"], g["target"], _buf(raw), len(raw))))
    assert got == g["truth"]


def test_sequence_all(orc, goldens):  # src/simulate/error.rs:926-947
    g = goldens["simulate.error.all"]
    r = orc.rng(42)
    em = orc.lib.orc_error_model_random(7)
    gm = orc.lib.orc_glitch_new(*g["glitch"])
    for ident, want in zip(g["identities"], g["reads"]):
        raw = r.random_seq(g["raw_len"])
        cap = 4 * len(raw)
        err, cig = (C.c_uint8 * cap)(), (C.c_uint8 * cap)()
        el, cl, rid = C.c_uint64(), C.c_uint64(), C.c_double()
        rc = orc.lib.orc_sequence(r.h, ident, _buf(raw), len(raw), em, gm, err, C.byref(el), cig, C.byref(cl), C.byref(rid), cap)
        assert rc == 0, orc.err()
        assert bytes(err[:el.value]).decode() == want["err"]
        assert bytes(cig[:cl.value]).decode() == want["cigar"]
        assert rid.value == want["real_id"]


# ---- #17: generate_quality with a file model ---------------------------------------------------------------------------------------------
def test_generate_quality(orc, goldens):  # src/simulate/quality.rs:58-74
    g = goldens["simulate.quality"]
    r = orc.rng(42)
    m = orc.lib.orc_quality_model_new(b"file", g["MODEL"].encode())
    _, cigar = orc.align(g["err"], g["raw"])
    out = (C.c_uint8 * (len(cigar) + 1))()
    n = orc.lib.orc_generate_quality(m, r.h, _buf(cigar), len(cigar), out)
    assert n == len(g["err"])
    assert bytes(out[:n]).decode() == g["qual"]


# ---- #18: Fragments ------------------------------------------------------------------------------------------------------------------------
TYPES = {"Real": 0, "Junk": 1, "Random": 2}
STRANDS = {"+": 0, "-": 1, "*": 2}


def test_fragments_iter(orc, goldens):  # src/simulate/fragments.rs:497-707
    w = _wl(orc, goldens["fragments.FASTA"])
    w.set_rates(20.0, 20.0, 20.0)
    w.feed(10_000, 10)
    d, ident, length = w.items()
    refs = w.refs()
    truth = goldens["fragments.iter"]
    assert len(truth) == 10
    for i, (r1, r2, desc, seed) in enumerate(truth):
        o = desc["origin"]
        assert d[i, 1] == r1 and d[i, 7] == r2 and int(d[i, 11]) == seed
        assert (d[i, 0], d[i, 2], d[i, 3], d[i, 4]) == (TYPES[o["read_type"]], STRANDS[o["strand"]], o["start"], o["end"])
        if o["read_type"] == "Real":
            assert refs[r1]["id"] == o["ref_id"]
        c = desc["chimera"]
        assert bool(d[i, 5]) == (c is not None)
        if c is not None:
            assert (d[i, 6], d[i, 8], d[i, 9], d[i, 10]) == (TYPES[c["read_type"]], STRANDS[c["strand"]], c["start"], c["end"])
        assert length[i] == desc["length"] and ident[i] == f(desc["identity"])


def test_fragments_read_type_and_chimera(orc, goldens):  # src/simulate/fragments.rs:198-251
    # get_read_type / is_chimera are single gen_bool draws on the shared stream
    r = orc.rng(42)
    got = []
    for _ in range(10):
        if orc.lib.orc_rng_gen_bool(r.h, 0.2):
            got.append("Junk")
        elif orc.lib.orc_rng_gen_bool(r.h, 0.2):
            got.append("Random")
        else:
            got.append("Real")
    assert got == goldens["fragments.read_type"]
    r = orc.rng(42)
    assert [bool(orc.lib.orc_rng_gen_bool(r.h, 0.2)) for _ in range(10)] == goldens["fragments.chimera"]


def _synthetic_fasta(rng, contigs):
    return "".join(">%s\n%s\n" % ("{:_}".format(n), rng.random_seq(n).decode()) for n in contigs)


def test_fragments_shape(orc, goldens):  # src/simulate/fragments.rs:709-910
    g = goldens["fragments.shape"]
    w = orc.workload(42)
    w.set_length(18200.0, 15500.0)
    w.set_identity(87.0, 100.0, 9.0)
    rh = w.rng_handle()

    class R:  # view of the workload's main rng
        h = rh
    lm = orc.lib.orc_length_new(18200.0, 15500.0)
    data = [float(orc.lib.orc_length_get(lm, rh)) for _ in range(100_000)]
    assert list(_stats(data)) == g["scalars"][0:3]
    im = orc.lib.orc_identity_new(87.0, 100.0, 9.0)
    data = [orc.lib.orc_identity_get(im, rh) for _ in range(100_000)]
    assert list(_stats(data)) == g["scalars"][3:6]
    parts = []
    for n in g["contigs"]:
        out = (C.c_uint8 * n)()
        orc.lib.orc_random_seq(rh, n, out)
        parts.append(">%s\n%s\n" % ("{:_}".format(n), bytes(out).decode()))
    w.set_references("".join(parts), adjusted=True)
    w.set_rates(1.0, 1.0, 1.0)
    w.feed(sum(g["contigs"]) * 50)
    d, ident, _ = w.items()
    types = [int((d[:, 0] == t).sum()) for t in (0, 1, 2)] + [int(d[:, 5].sum())]
    assert types == g["type_count"]
    lengths = [[] for _ in g["contigs"]]
    for row in d:
        lengths[int(row[1])].append(float(int(row[4]) - int(row[3])))
        if row[5]:
            lengths[int(row[7])].append(float(int(row[10]) - int(row[9])))
    assert [len(x) for x in lengths] == g["read_per_ref"]
    fsum = lambda v: _stats(v)[0]
    assert [fsum(x) / len(x) for x in lengths] == g["mean_len_per_ref"]
    sum_len = 0.0
    for x in lengths:
        sum_len += fsum(x)
    avg_len = sum_len / len(lengths)
    acc = 0.0
    for x in lengths:
        s = 0.0
        for v in x:
            s += (v - avg_len) ** 2.0
        acc += s
    std_len = math.sqrt(acc / len(lengths))
    assert [sum_len, avg_len, std_len] == g["scalars"][6:9]
    assert list(_stats(list(ident))) == g["scalars"][9:12]


# ---- #19: junk / random / real fragment extraction ------------------------------------------------------------------------------------------------
def _full_workload(orc, fasta, rates, adjusted=False, length=(8.0, 2.0), identity=(85.0, 95.0, 5.0), w=None):
    w = w or _wl(orc, fasta, adjusted=adjusted, length=length, identity=identity)
    w.set_adapter("TACGTATTGCT", "TACGTATTGCT", 90.0, 60.0, 50.0, 20.0)
    w.set_error_model(None)
    w.set_qscore_model("random")
    w.set_glitch(50.0, 5.0, 5.0)
    w.set_rates(*rates)
    return w


def test_real_fragments(orc, goldens):  # src/simulate/mod.rs:397-431
    w = _wl(orc, goldens["simulate.FASTA"])
    w.set_rates(0.0, 0.0, 0.0)
    w.feed(10_000, 10)
    d, _, _ = w.items()
    refs = w.refs()
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    got = []
    for row in d:
        seq = refs[int(row[1])]["seq"]
        local = seq if row[2] == 0 else seq.translate(comp)[::-1]
        s, e = int(row[3]), int(row[4])
        if s < e:
            got.append(list(local[s:e]))
        elif refs[int(row[1])]["circular"]:
            got.append(list(local[s:] + local[:e]))
        else:
            got.append([])
    assert got == goldens["simulate.read_reference"]


def test_junk_seq(orc, goldens):  # src/simulate/mod.rs:354-361 ; add_junk = gen_range(1..=5) then random_seq(unit)
    r = orc.rng(42)
    unit = orc.lib.orc_rng_gen_range_u64(r.h, 1, 5)
    small = r.random_seq(unit)
    assert (small * (100 // len(small))).decode() == goldens["simulate.junk_seq"]


# ---- #20: 20 full reads -----------------------------------------------------------------------------------------------------------------------------
def test_produce_read(orc, goldens):  # src/simulate/mod.rs:433-919
    w = _full_workload(orc, goldens["simulate.FASTA"], (10.0, 10.0, 25.0))
    w.feed(10_000, 20)
    w.generate(threads=2)
    seq, so = w.out("seq")
    qual, qo = w.out("qual")
    d, _, _ = w.items()
    ident = w.out_identity()
    refs = w.refs()
    truth = goldens["simulate.produce_read"]
    assert len(truth) == 20
    for i, (desc, wseq, wqual) in enumerate(truth):
        assert list(seq[int(so[i]):int(so[i + 1])]) == wseq, i
        assert list(qual[int(qo[i]):int(qo[i + 1])]) == wqual, i
        assert ident[i] == f(desc["identity"]), i
        assert int(so[i + 1] - so[i]) == desc["length"]
        o = desc["origin"]
        assert (d[i, 0], d[i, 2], d[i, 3], d[i, 4]) == (TYPES[o["read_type"]], STRANDS[o["strand"]], o["start"], o["end"])
        if o["read_type"] == "Real":
            assert refs[int(d[i, 1])]["id"] == o["ref_id"]
        assert bool(d[i, 5]) == (desc["chimera"] is not None)


# ---- #21: 5 Mbp long-read regime ------------------------------------------------------------------------------------------------------------------------
def test_shape_5mbp(orc, goldens):  # src/simulate/mod.rs:922-1023
    w = orc.workload(42)
    w.set_length(18200.0, 15500.0)
    w.set_identity(87.0, 100.0, 9.0)
    rh = w.rng_handle()
    parts = []
    for n in (5_000_000, 80_000, 75_000):
        out = (C.c_uint8 * n)()
        orc.lib.orc_random_seq(rh, n, out)
        parts.append(">%s\n%s\n" % ("{:_}".format(n), bytes(out).decode()))
    w.set_references("".join(parts), adjusted=True)
    _full_workload(orc, None, (1.0, 1.0, 1.0), w=w)
    n = w.feed(1_000_000)
    w.generate(threads=4, keep_cigar_raw=False)
    _, so = w.out("seq")
    lengths = [float(int(so[i + 1] - so[i])) for i in range(n)]
    ids = list(w.out_identity())
    sum_len = _stats(lengths)[0]
    avg_len = sum_len / len(lengths)
    std_len = math.sqrt(sum_len / len(lengths))
    sum_id, avg_id, std_id = _stats(ids)
    assert [sum_len, avg_len, std_len, sum_id, avg_id, std_id] == goldens["simulate.shape"]


# ---- #22: Description text --------------------------------------------------------------------------------------------------------------------------------
def test_description(orc):  # src/simulate/description.rs:115-166
    D = lambda *a: orc.take_str(orc.lib.orc_description(*a))
    ref = "bépo".encode()
    assert D(0, ref, ord("+"), 100, 400, 0, 0, b"", ord("*"), 0, 0, 301, 99.99) == \
        "bépo,+strand,100-400 length=300 error-free_length=301 read_identity=99.99%"
    assert D(0, ref, ord("+"), 100, 400, 1, 0, ref, ord("+"), 100, 400, 301, 99.99) == \
        "bépo,+strand,100-400 chimera bépo,+strand,100-400 length=300 error-free_length=301 read_identity=99.99%"
    assert D(1, ref, ord("+"), 100, 400, 1, 0, ref, ord("+"), 100, 400, 301, 99.99) == \
        "junk_seq chimera bépo,+strand,100-400 length=300 error-free_length=301 read_identity=99.99%"
    assert D(1, ref, ord("+"), 100, 400, 0, 0, b"", 0, 0, 0, 301, 99.99) == "junk_seq length=300 error-free_length=301 read_identity=99.99%"
    assert D(2, ref, ord("+"), 100, 400, 0, 0, b"", 0, 0, 0, 301, 99.99) == "random_seq length=300 error-free_length=301 read_identity=99.99%"
    # usize wrap-around of `length=` for circular wrap reads
    assert "length=%d " % (2 ** 64 - 2) in D(0, b"c", ord("+"), 4, 2, 0, 0, b"", 0, 0, 0, 10, 90.0)
    assert orc.take_str(orc.lib.orc_fmt_f64(90.0)) == "90"  # unpinned by the reference
    assert orc.take_str(orc.lib.orc_fmt_f64(76.66666666666666)) == "76.66666666666666"


# ---- #23: CLI parsers ------------------------------------------------------------------------------------------------------------------------------------------
def test_quantity(orc):  # src/cli/simulate.rs:247-338
    q = lambda s, g=1000: orc.lib.orc_quantity(s.encode(), g)
    assert q("50") == 50 and q("50K") == 50_000 and q("50M") == 50_000_000 and q("50G") == 50_000_000_000
    assert q("50x") == 50_000 and q("bépo") == -1 and q("50k") == -1 and q("x") == -1


# ---- md5 / uuid (UNPINNED by the reference; RFC 1321 known answers) ---------------------------------------------------------------------------------------------
def test_md5_uuid(orc):
    out = (C.c_uint8 * 16)()
    for msg, want in [(b"", "d41d8cd98f00b204e9800998ecf8427e"), (b"abc", "900150983cd24fb0d6963f7d28e17f72"),
                      (b"12345678901234567890123456789012345678901234567890123456789012345678901234567890", "57edf4a22be3c955ac49da2e2107b67a")]:
        orc.lib.orc_md5(_buf(msg), len(msg), out)
        assert bytes(out).hex() == want
    import uuid
    name = bytes(range(16))
    assert orc.take_str(orc.lib.orc_uuid_v3_x500(_buf(name), 16)) == str(uuid.uuid3(uuid.NAMESPACE_X500, name.decode("latin1"))) or True
    import hashlib
    d = bytearray(hashlib.md5(uuid.NAMESPACE_X500.bytes + name).digest())
    d[6] = (d[6] & 0x0F) | 0x30
    d[8] = (d[8] & 0x3F) | 0x80
    assert orc.take_str(orc.lib.orc_uuid_v3_x500(_buf(name), 16)) == str(uuid.UUID(bytes=bytes(d)))
