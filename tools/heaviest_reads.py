"""Developer tool: the heaviest reads of a batch (by estimated work) timed one at a time: the critical path of a small batch."""
import os, sys
here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(here, "..", "tests")); sys.path.insert(0, os.path.join(here, ".."))
import numpy as np, oraclelib, helpers
import rustyread_b200.capi as capi
orc = oraclelib.load()
fasta = helpers.random_fasta(orc, [5_000_000])
w, ctx, s = helpers.make_pair(orc, fasta, {})
n = w.feed(int(5_000_000 * float(sys.argv[1])))
d, ident, _ = w.items()
descs = capi.descs_from_oracle_items(d, ident)
ln = np.where(descs['end'] > descs['start'], descs['end'] - descs['start'], 0).astype(np.float64) + 64
cl = np.where(descs['has_chimera'] == 1, np.abs(descs['cend'].astype(np.int64) - descs['cstart'].astype(np.int64)), 0)
work = 80 * (1 - descs['identity']) * (ln + cl) + ln + cl
share = work.sum() / 4736
order = np.argsort(-work)
for i in order[:24]:
    ctx.upload(descs[i:i + 1])
    r = []
    for it in range(3):
        ctx.run(); r.append(ctx.last_run_ms()["read_kernel"])
    b = ctx.download()
    print("read %6d type %d chim %d len %7d clen %7d identity %.3f  est work %.2f share -> %.2f ms alone, err_len %d, %.2f us/est-event" % (
        i, descs['type'][i], descs['has_chimera'][i], ln[i], cl[i], descs['identity'][i], work[i] / share, min(r), b.lengths[0], min(r) * 1e3 / ((1 - descs['identity'][i]) * (ln[i] + cl[i]))), flush=True)
