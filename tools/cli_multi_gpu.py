"""Developer tool: wall clock of the sharded CLI (torchrun, one rank per GPU) with the single writer (shards gathered on
rank 0) and with the parallel writer (every rank pwrites its records in place), and a byte comparison of the two files.
usage: cli_multi_gpu.py <n_gpus> [quantity] [number_base_store]"""
import hashlib, os, subprocess, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
root = os.path.join(os.path.dirname(__file__), "..")
n = int(sys.argv[1])
q = sys.argv[2] if len(sys.argv) > 2 else "400x"
store = sys.argv[3] if len(sys.argv) > 3 else "1G"
ref = "/tmp/ref5m.fasta"
open(ref, "wb").write(bench.synthetic_fasta())
digests = {}
for mode in ("single", "parallel"):
    out = "/dev/shm/rr_%s.fastq" % mode
    env = dict(os.environ)
    if mode == "single":
        env["RUSTYREAD_B200_SINGLE_WRITER"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % n, "--master-addr", "127.0.0.1", "--master-port", "29561",
           "-m", "rustyread_b200", "-v", "simulate", "--reference", ref, "--quantity", q, "--seed", "42", "--error_model", "random", "--qscore_model", "random",
           "--output", out, "--number_base_store", store]
    t = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=root)
    dt = time.time() - t
    tail = [l for l in r.stderr.strip().split("\n") if l.startswith("wrote")]
    print(mode, "rc", r.returncode, tail[-1] if tail else r.stderr[-600:], "wall %.2fs" % dt, flush=True)
    if r.returncode == 0 and tail:
        bases = int(tail[-1].split()[-2])
        print("  -> %.2f Gbases/s whole job (incl. process start, reference parse, context creation on %d ranks)" % (bases / dt / 1e9, n))
        h = hashlib.md5()
        with open(out, "rb") as f:
            for blk in iter(lambda: f.read(1 << 24), b""):
                h.update(blk)
        digests[mode] = (h.hexdigest(), os.path.getsize(out))
        os.remove(out)
print(digests, "IDENTICAL" if len(digests) == 2 and digests["single"] == digests["parallel"] else "DIFFERENT")
