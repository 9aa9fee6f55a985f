"""Developer tool: rr_run timing with the synthetic badread-scale FILE models (table-driven alternatives, 9-wide qscore windows)."""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import oraclelib, helpers
import rustyread_b200.capi as capi
from test_gpu_parity import _synthetic_models

orc = oraclelib.load()
error_text, qscore_text = _synthetic_models()
cov = float(sys.argv[1]) if len(sys.argv) > 1 else 100
fasta = helpers.random_fasta(orc, [5_000_000])
for name, spec in [("random/random", {}), ("file/file", dict(error_model=error_text, qscore_model=qscore_text)), ("random/ideal", dict(qscore_model="ideal"))]:
    w, ctx, s = helpers.make_pair(orc, fasta, spec)
    n = w.feed(int(5_000_000 * cov))
    d, ident, _ = w.items()
    descs = capi.descs_from_oracle_items(d, ident)
    ctx.upload(descs)
    for it in range(3):
        ctx.run()
    ms = ctx.last_run_ms()
    b = ctx.last_run_bases()
    print("%-14s %d reads: read %.1f ms qscore %.1f ms pack %.1f ms -> %.2f Gbases/s" % (name, n, ms["read_kernel"], ms["qscore"], ms["pack"], b / ms["total"] / 1e6))
    ctx.close()
