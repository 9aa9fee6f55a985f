"""Developer tool: BASELINE.json configs[2] through the CLI on N GPUs of one box: the 3.1 Gbp / 24-contig synthetic reference,
`--quantity <q>`, sharded over N ranks with the parallel writer, against the same command on one GPU (md5 of both files).
usage: human_scale_cli.py <n_gpus> [quantity=1x] [out.json]"""
import hashlib, json, os, subprocess, sys, time
here = os.path.dirname(os.path.abspath(__file__))
root = os.path.join(here, "..")
sys.path.insert(0, here)
import human_scale

n = int(sys.argv[1])
q = sys.argv[2] if len(sys.argv) > 2 else "1x"
out_json = sys.argv[3] if len(sys.argv) > 3 else None
bp = int(os.environ.get("RR_GENOME_BP", human_scale.TOTAL))
ref = "/dev/shm/rr_ref3g.fasta"
t = time.time()
fasta, lens = human_scale.synthetic_genome(bp)
open(ref, "wb").write(fasta)
del fasta
res = {"reference_bp": sum(lens), "contigs": len(lens), "quantity": q, "make_reference_s": time.time() - t, "runs": {}}


def md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


for ranks in ((n, 1) if not os.environ.get("RR_ONLY_N") else (n,)):
    out = "/dev/shm/rr_out_%d.fastq" % ranks
    base = [sys.executable, "-m", "rustyread_b200", "-v", "simulate", "--reference", ref, "--quantity", q, "--seed", "42", "--error_model", "random",
            "--qscore_model", "random", "--output", out]
    cmd = base if ranks == 1 else [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % ranks, "--master-addr", "127.0.0.1",
                                   "--master-port", "29571", "-m", "rustyread_b200"] + base[3:]
    t = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=root, env=dict(os.environ, RUSTYREAD_B200_TIMING="1"))
    dt = time.time() - t
    lines = r.stderr.strip().split("\n")
    wrote = [l for l in lines if l.startswith("wrote")]
    run = {"rc": r.returncode, "wall_s": dt, "timing": sorted(l for l in lines if "timing" in l and ("rank 0" in l or "rank 1:" in l or "rank 1 " in l or "timing (s)" in l))}
    if r.returncode == 0 and wrote:
        run["reads"], run["bases"] = int(wrote[-1].split()[1]), int(wrote[-1].split()[-2])
        run["gbases_per_s_whole_process"] = run["bases"] / dt / 1e9
        run["file_bytes"] = os.path.getsize(out)
        run["md5"] = md5(out)
    else:
        run["stderr_tail"] = r.stderr[-1500:]
    if os.path.exists(out):
        os.remove(out)
    res["runs"]["%d_gpu" % ranks] = run
    print(json.dumps({("%d_gpu" % ranks): run}), flush=True)
a, b = res["runs"]["%d_gpu" % n], res["runs"].get("1_gpu", {"md5": os.environ.get("RR_EXPECT_MD5")})
res["identical"] = bool(a.get("md5") and a.get("md5") == b.get("md5"))
os.remove(ref)
print(json.dumps(res, indent=1))
if out_json:
    json.dump(res, open(out_json, "w"), indent=1)
sys.exit(0 if res["identical"] else 1)
