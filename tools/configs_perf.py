"""Developer tool: kernel times and throughput of the BASELINE.json configurations other than the bench line
(the bench is configs[1]; configs[2] at full reference size is tools/human_scale.py), plus the file-model paths.

    python tools/configs_perf.py [out.json]

Every line: reads, bases, per-kernel ms (best of 3 resident runs), resident Gbases/s and host-to-host Gbases/s (rr_generate)."""
import json
import os
import random
import sys
import time

here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(here, "..", "tests"))
sys.path.insert(0, os.path.join(here, ".."))
import numpy as np
import helpers
import oraclelib
import rustyread_b200.capi as capi
from test_gpu_parity import _synthetic_models

orc = oraclelib.load()
error_text, qscore_text = _synthetic_models()
ref5m = helpers.random_fasta(orc, [5_000_000])
rnd = random.Random(5)
plasmid_lens = [int(2000 * (25 ** rnd.random())) for _ in range(200)]
plasmids = helpers.random_fasta(orc, plasmid_lens, circular=set(range(200)))
cfg4 = dict(length=(50000.0, 40000.0), identity=(75.0, 90.0, 8.0), glitch=(2000.0, 50.0, 50.0), rates=(1.0, 1.0, 5.0))
# with the default --length the reference (and so the feeder) never terminates on an all-circular reference whose contigs are
# all shorter than a drawn fragment (fragments.rs:80-85); 4000,3000 keeps every draw of a 1-Gbase run below the largest plasmids
# (6000,5000 still draws one such fragment within ~170 k reads)
cfg5 = dict(rates=(10.0, 10.0, 1.0), small_plasmid_bias=True, length=(4000.0, 3000.0))
CASES = [
    ("configs[0] ref5m 20x defaults random/random", ref5m, {}, 5_000_000 * 20),
    ("configs[1] ref5m 200x defaults random/random (bench line)", ref5m, {}, 5_000_000 * 200),
    ("ref5m 600x defaults random/random (3 Gbases in one batch; resident only)", ref5m, {}, 5_000_000 * 600),
    ("configs[3] ref5m 200x --length 50000,40000 --identity 75,90,8 --glitches 2000,50,50 --chimera 5", ref5m, cfg4, 5_000_000 * 200),
    ("configs[4] 200 circular plasmids 2-50 kbp, 1 Gbase, --length 4000,3000 --small_plasmid_bias --junk_reads 10 --random_reads 10", plasmids, cfg5, 1_000_000_000),
    ("ref5m 200x synthetic badread-scale FILE models (12k 7-mers x <=25 alternatives, CIGAR windows up to 9)", ref5m,
     dict(error_model=error_text, qscore_model=qscore_text), 5_000_000 * 200),
    ("ref5m 200x random error model, ideal qscores", ref5m, dict(qscore_model="ideal"), 5_000_000 * 200),
]
out = []
only = os.environ.get("RR_CASES")  # comma separated indices
if only:
    CASES = [CASES[int(i)] for i in only.split(",")]
for name, fasta, spec, target in CASES:
    t_case = time.time()
    w, ctx, s = helpers.make_pair(orc, fasta, spec)
    n = w.feed(int(target))
    d, ident, _ = w.items()
    descs = capi.descs_from_oracle_items(d, ident)
    ctx.upload(descs)
    best = None
    for it in range(4):
        ctx.run()
        ms = ctx.last_run_ms()
        if it and (best is None or ms["total"] < best["total"]):
            best = dict(ms)
    bases = ctx.last_run_bases()
    e2e = float("nan")
    if "resident only" not in name:
        ctx.submit(descs)  # warm both batch slots (their pinned staging is allocated on first use)
        ctx.submit(descs)
        ctx.collect()
        ctx.collect()
        t = time.perf_counter()
        reps = 3
        ctx.submit(descs)
        for r in range(reps):
            if r + 1 < reps:
                ctx.submit(descs)
            ctx.collect()
        e2e = reps * bases / (time.perf_counter() - t)
    line = {"case": name, "reads": int(n), "bases": int(bases), "kernels_ms": best, "resident_gbases_per_s": bases / best["total"] / 1e6,
            "host_to_host_gbases_per_s": (e2e / 1e9) if e2e == e2e else None}
    line["case_wall_s"] = time.time() - t_case
    print(json.dumps(line), flush=True)
    out.append(line)
    ctx.close()
    del w, ctx
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
