"""Developer tool: BASELINE.json configs[2] at its full reference size on ONE GPU.

A synthetic 3.1 Gbp reference (24 linear contigs with human-like lengths, uniform ACGT, numpy seed 42) goes through the
product's host side (FASTA parse, depth adjustment, feeder) and the device library; `--quantity <cov>x` (default 1x =
3.1 Gbases; the CPU side of SURVEY §8d is timed at 1x as well, throughput is quantity-independent) is generated as ONE
reference batch in ~1 Gbase device chunks through rr_submit / rr_collect, and a sample of reads (the heaviest ones, every
chimera / junk / random read among the first thousand, and a random draw) is checked byte for byte against the CPU oracle
fed with the same reference and the same work items.

    python tools/human_scale.py [coverage] [out.json]
"""
import json
import os
import sys
import time

import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(here, "..", "tests"))
sys.path.insert(0, os.path.join(here, ".."))

REL = [248, 242, 198, 190, 181, 171, 159, 145, 138, 133, 135, 133, 114, 107, 102, 90, 83, 80, 58, 64, 46, 50, 156, 57]
TOTAL = 3_100_000_000


def synthetic_genome(total=TOTAL, seed=42, width=80):
    rs = np.random.RandomState(seed)
    nuc = np.frombuffer(b"ACGT", dtype=np.uint8)
    parts, lens = [], []
    scale = total / float(sum(REL))
    for i, r in enumerate(REL):
        n = int(r * scale)
        lens.append(n)
        rows = (n + width - 1) // width
        body = np.full((rows, width + 1), ord("\n"), dtype=np.uint8)
        flat = nuc[rs.randint(0, 4, size=rows * width, dtype=np.uint8)]
        body[:, :width] = flat.reshape(rows, width)
        text = body.reshape(-1)
        cut = rows * (width + 1) - (rows * width - n) - 1  # drop the padding bases of the last line (and its newline)
        parts.append(b">chr%d\n" % (i + 1))
        parts.append(text[:cut].tobytes())
        parts.append(b"\n")
    return b"".join(parts), lens


def main():
    cov = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    out_path = sys.argv[2] if len(sys.argv) > 2 else None
    total_bp = int(os.environ.get("RR_GENOME_BP", TOTAL))
    import oraclelib
    import rustyread_b200.capi as capi
    from rustyread_b200.simulate import chunk_bounds

    res = {"workload": "24 linear contigs, %d bp, uniform ACGT (numpy seed 42); --quantity %gx --seed 42 defaults, error_model=random qscore_model=random" % (total_bp, cov)}
    t = time.time()
    fasta, lens = synthetic_genome(total_bp)
    res["make_fasta_s"] = time.time() - t
    t = time.time()
    sim = capi.Simulation(fasta, seed=42)
    res["host_load_s"] = time.time() - t  # FASTA parse + adjust_depth (100 000 length draws, src/references.rs:95-132)
    t = time.time()
    ctx = capi.Context(0)
    sim.attach(ctx)
    res["device_attach_s"] = time.time() - t
    assert sim.genome_length == sum(lens), (sim.genome_length, sum(lens))
    t = time.time()
    descs = sim.feed(int(sim.genome_length * cov))
    res["feeder_s"] = time.time() - t
    res["reads"] = int(len(descs))
    cb = chunk_bounds(descs, 1_000_000_000)
    res["device_chunks"] = len(cb) - 1

    # ---- pipelined generation of the batch (host buffers out), wall clock ----
    def run_pipeline(keep):
        kept, bases, kms = {}, 0, []
        ctx.submit(descs[cb[0]:cb[1]])
        for c in range(len(cb) - 1):
            if c + 1 < len(cb) - 1:
                ctx.submit(descs[cb[c + 1]:cb[c + 2]])
            b = ctx.collect()
            kms.append(ctx.last_run_ms())
            bases += b.total_bases
            for i in keep:
                if cb[c] <= i < cb[c + 1]:
                    j = i - cb[c]
                    o, l = int(b.offsets[j]), int(b.lengths[j])
                    kept[i] = (b.seq[o:o + l].copy(), b.qual[o:o + l].copy(), float(b.real_identity[j]))
        return kept, bases, kms

    # sample: heaviest reads, special reads, random draw
    ln = np.where(descs["end"] > descs["start"], descs["end"] - descs["start"], 0).astype(np.float64)
    work = (1 - descs["identity"]) * ln
    rs = np.random.RandomState(7)
    special = [int(i) for i in np.nonzero((descs["type"] != capi.RR_REAL) | (descs["has_chimera"] == 1))[0][:40]]
    sample = sorted(set([int(i) for i in np.argsort(-work)[:12]] + special + [int(i) for i in rs.randint(0, len(descs), size=200)]))
    run_pipeline([])  # warm-up: arenas, pinned staging
    t = time.time()
    kept, bases, kms = run_pipeline(sample)
    wall = time.time() - t
    res["bases"] = int(bases)
    res["e2e_wall_s"] = wall
    res["e2e_gbases_per_s"] = bases / wall / 1e9
    res["kernels_ms_per_chunk"] = kms
    dev_ms = sum(k["total"] for k in kms)
    res["resident_gbases_per_s"] = bases / dev_ms / 1e6

    # ---- sampled parity against the oracle on the same 3.1 Gbp reference ----
    t = time.time()
    orc = oraclelib.load()
    w = orc.workload(42)
    w.set_length(15000.0, 13000.0)
    w.set_references(fasta, adjusted=True, small_plasmid_bias=False)
    w.set_identity(85.0, 95.0, 5.0)
    w.set_adapter("AATGTACTTCGTTCAGTTACGTATTGCT", "GCAATACGTAACTGAACGAAGT", 90.0, 60.0, 50.0, 20.0)
    w.set_glitch(10000.0, 25.0, 25.0)
    w.set_error_model(None)
    w.set_qscore_model("random")
    w.set_rates(1.0, 1.0, 1.0)
    res["oracle_load_s"] = time.time() - t
    # the feeder itself: the product's work items for this reference against the oracle's own feeder
    t = time.time()
    n2 = w.feed(int(sim.genome_length * cov))
    od, oident, _ = w.items()
    pd, pident = capi.descs_to_item_table(descs)
    res["feeder_matches_oracle"] = bool(n2 == len(descs) and np.array_equal(od, pd) and np.array_equal(oident, pident))
    res["oracle_feeder_s"] = time.time() - t
    d12, ident = capi.descs_to_item_table(descs[sample])
    w.load_items(d12, ident)
    w.generate(threads=os.cpu_count() or 8, keep_cigar_raw=False)
    seq, so = w.out("seq")
    qual, qo = w.out("qual")
    oid = w.out_identity()
    bad = 0
    for j, i in enumerate(sample):
        s, q, rid = kept[i]
        if not (np.array_equal(s, seq[int(so[j]):int(so[j + 1])]) and np.array_equal(q, qual[int(qo[j]):int(qo[j + 1])]) and rid == oid[j]):
            bad += 1
    res["parity_sample_reads"] = len(sample)
    res["parity_sample_bases"] = int(so[-1])
    res["parity_mismatching_reads"] = bad
    print(json.dumps(res, indent=1))
    if out_path:
        json.dump(res, open(out_path, "w"), indent=1)
    ctx.close()
    return 1 if (bad or not res["feeder_matches_oracle"]) else 0


if __name__ == "__main__":
    sys.exit(main())
