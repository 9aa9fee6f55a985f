"""Developer tool: randomised parity sweep on the GPU.  Every case draws its own reference (linear + circular contigs, some with
N runs, IUPAC codes and lower case), models (Error::random, or a small-k file model of random shape; qscore random / ideal / file)
and flags, runs the same work items through the device library and the CPU oracle and compares every read byte for byte
(tests/helpers.compare_batch).  usage: fuzz_parity.py [n_cases=24] [first_seed=0]"""
import os, random, sys, time
here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(here, "..", "tests")); sys.path.insert(0, os.path.join(here, ".."))
import helpers, oraclelib
import rustyread_b200.capi as capi
import test_gpu_parity as T

orc = oraclelib.load()
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 24
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
_, qscore_text = T._synthetic_models(k=3)
fails = 0
for seed in range(first, first + n_cases):
    rnd = random.Random(1000 + seed)
    contigs = [rnd.randint(20_000, 120_000)] + [rnd.randint(3_000, 30_000) for _ in range(rnd.randint(0, 3))]
    circ = {i for i in range(1, len(contigs)) if rnd.random() < 0.5}
    fasta = helpers.random_fasta(orc, contigs, seed=seed, circular=circ)
    if rnd.random() < 0.4:  # bytes outside ACGT
        lines = fasta.split("\n")
        for i in range(len(lines)):
            if lines[i] and lines[i][0] != ">" and rnd.random() < 0.1:
                r = rnd.random()
                lines[i] = "N" * len(lines[i]) if r < 0.3 else lines[i].lower() if r < 0.6 else "".join(rnd.choice("ACGTRYKMSWN") for _ in lines[i])
        fasta = "\n".join(lines)
    mean = rnd.choice([20000.0, 50000.0] if os.environ.get("FUZZ_LONG") else [800.0, 2500.0, 6000.0])  # FUZZ_LONG=1: ultra-long reads
    spec = dict(seed=seed, length=(mean, mean * rnd.choice([0.5, 0.9, 1.3])), identity=rnd.choice([(85.0, 95.0, 5.0), (70.0, 88.0, 9.0), (97.0, 100.0, 2.0), (99.0, 100.0, 0.5)]),
                glitch=rnd.choice([(10000.0, 25.0, 25.0), (300.0, 40.0, 10.0), (0.0, 25.0, 25.0), (150.0, 5.0, 60.0)]),
                rates=rnd.choice([(1.0, 1.0, 1.0), (10.0, 10.0, 20.0), (0.0, 0.0, 50.0)]), small_plasmid_bias=rnd.random() < 0.5)
    em = rnd.choice(["random", "packed4", "packed7e", "long3", "synth5"])
    if em == "packed4": spec["error_model"] = T._packed_edge_model(4, seed=seed)
    elif em == "packed7e": spec["error_model"] = T._packed_edge_model(7, seed=seed) if seed % 4 == 0 else T._packed_edge_model(5, seed=seed)
    elif em == "long3": spec["error_model"] = T._long_alt_model(3, seed=seed)
    elif em == "synth5": spec["error_model"] = T._synthetic_models(k=5, seed=seed)[0]
    spec["qscore_model"] = rnd.choice(["random", "ideal", qscore_text])
    t = time.time()
    try:
        target = rnd.choice([2_000_000, 5_000_000] if os.environ.get("FUZZ_LONG") else [150_000, 400_000])
        w, ctx, batch, n = T._run_pair(orc, fasta, spec, target)
        ctx.close()
        print("case %3d ok: %d reads, %d bases, contigs %s circ %s error %s qscore %s identity %s glitch %s (%.1f s)" % (
            seed, n, batch.total_bases, contigs, sorted(circ), em, spec["qscore_model"][:6].replace("\n", " "), spec["identity"], spec["glitch"], time.time() - t), flush=True)
    except Exception as e:
        msg = str(e)
        verdict = "FAIL" if "mismatch" in msg or "differ" in msg or "Assertion" in type(e).__name__ else "error"
        if verdict == "error" and "reference panics" in msg:
            # a panic of the reference itself (a slice out of range): legitimate for extreme flags, but the oracle has to fail on
            # the same read and on nothing before it
            import re
            r = int(re.search(r"read (\d+) ", msg).group(1))
            w2, ctx2, _ = helpers.make_pair(orc, fasta, spec)
            w2.feed(target)
            d2, i2, _ = w2.items()
            ctx2.close()
            try:
                w2.load_items(d2[:r], i2[:r]); w2.generate(threads=4)
                before_ok = True
            except Exception:
                before_ok = False
            try:
                w2.load_items(d2[r:r + 1], i2[r:r + 1]); w2.generate(threads=1)
                same = False
            except Exception as oe:
                same = "reference panics" in str(oe)
            verdict = "ok (both sides panic on read %d like the reference)" % r if (before_ok and same) else "FAIL (panic disagreement)"
        print("case %3d %s: %s" % (seed, verdict, msg[:300]), flush=True)
        fails += 0 if verdict.startswith("ok") else 1
print("done: %d cases, %d not ok" % (n_cases, fails))
sys.exit(1 if fails else 0)
