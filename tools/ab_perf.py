"""Developer tool: A/B timing of library builds on the same box.  usage: ab_perf.py cov lib1[@VAR=value,...] lib2 ...  (prints min / median read-kernel ms)"""
import os, subprocess, sys
cov = sys.argv[1]
here = os.path.dirname(os.path.abspath(__file__))
code = r'''
import sys, os
sys.path.insert(0, os.path.join(%r, "..", "tests")); sys.path.insert(0, os.path.join(%r, ".."))
import numpy as np, oraclelib, helpers
import rustyread_b200.capi as capi
if os.environ.get("RR_LIB"): capi.load_library(os.environ["RR_LIB"])
orc = oraclelib.load()
fasta = helpers.random_fasta(orc, [5_000_000])
w, ctx, s = helpers.make_pair(orc, fasta, {})
n = w.feed(int(5_000_000 * float(sys.argv[1])))
d, ident, _ = w.items()
descs = capi.descs_from_oracle_items(d, ident)
ctx.upload(descs)
r = []
for it in range(int(sys.argv[2])):
    ctx.run(); ms = ctx.last_run_ms(); r.append((ms["read_kernel"], ms["qscore"]))
r = r[2:]
rk = sorted(x[0] for x in r); q = sorted(x[1] for x in r)
print("read min %%.2f med %%.2f  qscore min %%.2f  (%%d runs)" %% (rk[0], rk[len(rk)//2], q[0], len(r)))
''' % (here, here)
for rnd in range(2):
    for lib in sys.argv[2:]:
        env = dict(os.environ)
        lib, _, settings = lib.partition("@")  # lib.so@VAR=value,VAR2=value: environment of that run
        for kv in filter(None, settings.split(",")):
            env[kv.split("=")[0]] = kv.split("=")[1]
        if lib != "-": env["RR_LIB"] = lib
        out = subprocess.run([sys.executable, "-c", code, cov, "10"], env=env, capture_output=True, text=True)
        print("%-40s %s" % (lib.split("/")[-1] + ("@" + settings if settings else ""), out.stdout.strip() or out.stderr.strip()[-300:]), flush=True)
