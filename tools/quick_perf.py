"""Developer tool (not the bench): times rr_run on a synthetic workload, descriptors from the oracle feeder."""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import oraclelib, helpers
import rustyread_b200.capi as capi

orc = oraclelib.load()
if os.environ.get('RR_LIB'):
    capi.load_library(os.environ['RR_LIB'])
ref_len = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
cov = float(sys.argv[2]) if len(sys.argv) > 2 else 20
spec = {}
if len(sys.argv) > 3 and sys.argv[3] == "cfg4":
    spec = dict(length=(50000.0, 40000.0), identity=(75.0, 90.0, 8.0), glitch=(2000.0, 50.0, 50.0), rates=(1.0, 1.0, 5.0))
t = time.time()
fasta = helpers.random_fasta(orc, [ref_len])
w, ctx, s = helpers.make_pair(orc, fasta, spec)
n = w.feed(int(ref_len * cov))
d, ident, _ = w.items()
descs = capi.descs_from_oracle_items(d, ident)
print("setup %.1fs, %d reads" % (time.time() - t, n))
for it in range(3):
    t = time.time(); ctx.upload(descs); t1 = time.time(); ctx.run(); t2 = time.time()
    if os.environ.get('NODL'):
        class B: total_bases = ctx.last_run_bases()
        b = B()
    else:
        b = ctx.download()
    t3 = time.time()
    ms = ctx.last_run_ms()
    print("iter %d: upload %.1f ms run %.1f ms (dev total %.1f read %.1f qscore %.1f pack %.1f) download %.1f ms  bases %d -> %.2f Gbases/s kernel, %.2f e2e" % (
        it, (t1 - t) * 1e3, (t2 - t1) * 1e3, ms["total"], ms["read_kernel"], ms["qscore"], ms["pack"], (t3 - t2) * 1e3, b.total_bases,
        b.total_bases / ms["total"] / 1e6, b.total_bases / (t3 - t) / 1e9))
if os.environ.get("CPU"):
    for th in (1, 16):
        sec = w.generate(threads=th, keep_cigar_raw=False)
        print("oracle %d threads: %.2fs -> %.1f Mbases/s" % (th, sec, b.total_bases / sec / 1e6))
