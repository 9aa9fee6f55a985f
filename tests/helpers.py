"""Shared test plumbing: one workload spec -> (oracle workload, device context) fed with the same inputs."""
import ctypes as C

import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DEFAULT = dict(
    seed=42,
    length=(15000.0, 13000.0),
    identity=(85.0, 95.0, 5.0),
    adapter=("AATGTACTTCGTTCAGTTACGTATTGCT", "GCAATACGTAACTGAACGAAGT", 90.0, 60.0, 50.0, 20.0),
    glitch=(10000.0, 25.0, 25.0),
    rates=(1.0, 1.0, 1.0),  # junk, random, chimera
    error_model=None,        # None = random(7), else model text
    qscore_model="random",   # random | ideal | model text
    adjusted=True,
    small_plasmid_bias=False,
)


def random_fasta(orc, contigs, seed=42, circular=(), depth=None, width=80):
    """Synthetic reference: uniform i.i.d. bases from the oracle's StdRng (seeded), 80-col FASTA."""
    r = orc.rng(seed)
    parts = []
    for i, n in enumerate(contigs):
        s = r.random_seq(n).decode()
        hdr = ">contig_%d" % i
        if i in circular:
            hdr += " circular=true"
        if depth and i in depth:
            hdr += " depth=%s" % depth[i]
        parts.append(hdr + "\n" + "\n".join(s[j:j + width] for j in range(0, len(s), width)) + "\n")
    return "".join(parts)


def parse_error_dump(text):
    entries = []
    for line in text.splitlines():
        key, alts = line.split("\t")
        al = []
        for a in alts.split(","):
            seq, _edit, prob = a.rsplit(":", 2)
            al.append((seq.encode(), float(prob)))
        entries.append((key.encode(), al))
    return entries


def parse_quality_dump(text):
    entries = []
    for line in text.splitlines():
        key, sw = line.split("\t")
        scores, weights = [], []
        for a in sw.split(","):
            if not a:
                continue
            s, w = a.split(":")
            scores.append(int(s))
            weights.append(float(w))
        entries.append((key.encode(), scores, weights))
    return entries


def make_pair(orc, fasta, spec, device=0):
    import rustyread_b200 as rr

    # The error-model loader draws from main_rng (src/model/error.rs:65-69), so the device tables must
    # come from the very same load.  Load once through a scratch rng positioned like the workload's.
    s = dict(DEFAULT)
    s.update(spec)
    w = orc.workload(s["seed"])
    w.set_length(*s["length"])
    w.set_references(fasta, adjusted=s["adjusted"], small_plasmid_bias=s["small_plasmid_bias"])
    w.set_identity(*s["identity"])
    w.set_adapter(*s["adapter"])
    w.set_glitch(*s["glitch"])
    err_entries = None
    if s["error_model"] is not None:
        # a model handle loaded with the workload's main rng, then adopted by the workload via text reload is not
        # possible (double draw); instead: clone the rng state by replaying = load on a twin workload
        twin = orc.workload(s["seed"])
        twin.set_length(*s["length"])
        twin.set_references(fasta, adjusted=s["adjusted"], small_plasmid_bias=s["small_plasmid_bias"])
        mh = orc.lib.orc_error_model_from_text(s["error_model"].encode(), twin.rng_handle())
        assert mh, orc.err()
        k = orc.lib.orc_error_model_k(mh)
        err_entries = (k, parse_error_dump(orc.take_str(orc.lib.orc_error_model_dump(mh))))
        orc.lib.orc_error_model_free(mh)
    w.set_error_model(s["error_model"])
    if s["qscore_model"] in ("random", "ideal"):
        w.set_qscore_model(s["qscore_model"])
    else:
        w.set_qscore_model("file", s["qscore_model"])
    w.set_rates(*s["rates"])

    ctx = rr.Context(device)
    refs = w.refs()
    ctx.set_references([r["seq"] for r in refs], [r["circular"] for r in refs])
    if err_entries is None:
        ctx.set_error_model_random(7)
    else:
        ctx.set_error_model(err_entries[0], err_entries[1])
    if s["qscore_model"] == "random":
        ctx.set_qscore_model_random()
    elif s["qscore_model"] == "ideal":
        ctx.set_qscore_model_ideal()
    else:
        qh = orc.lib.orc_quality_model_new(b"file", s["qscore_model"].encode())
        assert qh, orc.err()
        ctx.set_qscore_model(orc.lib.orc_quality_model_max_k(qh), parse_quality_dump(orc.take_str(orc.lib.orc_quality_model_dump(qh))))
        orc.lib.orc_quality_model_free(qh)
    a = s["adapter"]
    ctx.set_adapter(a[0].encode(), a[1].encode(), *a[2:])
    ctx.set_glitch(*s["glitch"])
    return w, ctx, s


def compare_batch(w, batch, ctx=None, check_intermediate=True, max_report=3):
    """Byte-compare a device batch with the oracle's outputs for the same work items."""
    seq, so = w.out("seq")
    qual, qo = w.out("qual")
    ident = w.out_identity()
    n = w.n
    assert batch.n == n
    bad = []
    lens = (so[1:] - so[:-1]).astype(np.uint64)
    if not np.array_equal(lens, batch.lengths):
        idx = np.nonzero(lens != batch.lengths)[0]
        bad.append("length mismatch on %d reads, first %s: oracle %s device %s" % (len(idx), idx[:5], lens[idx[:5]], batch.lengths[idx[:5]]))
    for i in range(n):
        o, l = int(batch.offsets[i]), int(batch.lengths[i])
        es = seq[int(so[i]):int(so[i + 1])]
        eq = qual[int(qo[i]):int(qo[i + 1])]
        ok = l == len(es) and np.array_equal(batch.seq[o:o + l], es) and np.array_equal(batch.qual[o:o + l], eq)
        if not ok or batch.real_identity[i] != ident[i]:
            msg = "read %d differs (len dev %d / oracle %d, identity dev %r / oracle %r)" % (i, l, len(es), batch.real_identity[i], ident[i])
            if ctx is not None and check_intermediate:
                raw, ro = w.out("raw")
                cig, co = w.out("cigar")
                draw = ctx.debug_fetch(i, "raw")
                dcig = ctx.debug_fetch(i, "cigar")
                eraw = bytes(raw[int(ro[i]):int(ro[i + 1])])
                ecig = bytes(cig[int(co[i]):int(co[i + 1])])
                msg += "; raw %s" % ("equal" if draw == eraw else "DIFFERS at %d" % first_diff(draw, eraw))
                msg += "; cigar %s" % ("equal" if dcig == ecig else "DIFFERS at %d" % first_diff(dcig, ecig))
                if l == len(es):
                    sd = np.nonzero(batch.seq[o:o + l] != es)[0]
                    qd = np.nonzero(batch.qual[o:o + l] != eq)[0]
                    msg += "; seq diffs %d first %s; qual diffs %d first %s" % (len(sd), sd[:3], len(qd), qd[:3])
            bad.append(msg)
            if len(bad) >= max_report:
                break
    assert not bad, "\n".join(bad)


def first_diff(a, b):
    n = min(len(a), len(b))
    for i in range(n):
        if a[i] != b[i]:
            return i
    return n
