"""Developer tool: steady-state host-to-host throughput of a long rr_submit / rr_collect pipeline on one workload of
tools/configs_perf.py (the 3-batch figure there is dominated by the first and last batch).  usage: pipeline_perf.py <cfg1|cfg4> [reps=10]"""
import os, sys, time
here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(here, "..", "tests")); sys.path.insert(0, os.path.join(here, ".."))
import helpers, oraclelib
import rustyread_b200.capi as capi
orc = oraclelib.load()
which = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
spec = dict(length=(50000.0, 40000.0), identity=(75.0, 90.0, 8.0), glitch=(2000.0, 50.0, 50.0), rates=(1.0, 1.0, 5.0)) if which == "cfg4" else {}
w, ctx, s = helpers.make_pair(orc, helpers.random_fasta(orc, [5_000_000]), spec)
w.feed(1_000_000_000)
d, ident, _ = w.items()
descs = capi.descs_from_oracle_items(d, ident)
for _ in range(2):
    ctx.submit(descs); ctx.submit(descs); ctx.collect(); ctx.collect()
t = time.perf_counter()
bases = 0
ctx.submit(descs)
for r in range(reps):
    if r + 1 < reps:
        ctx.submit(descs)
    bases += ctx.collect().total_bases
dt = time.perf_counter() - t
print("%s: %d batches, %.1f ms per batch, %.2f Gbases/s host to host (RR_OVERLAP_RATIO=%s RR_ROOM_BLOCKS=%s)" % (
    which, reps, 1e3 * dt / reps, bases / dt / 1e9, os.environ.get("RR_OVERLAP_RATIO", "default"), os.environ.get("RR_ROOM_BLOCKS", "default")))
