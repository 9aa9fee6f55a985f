"""Developer tool: wall-clock of the whole CLI (`python -m rustyread_b200 simulate`) on the bench reference."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
ref = "/tmp/ref5m.fasta"
open(ref, "wb").write(bench.synthetic_fasta())
q = sys.argv[1] if len(sys.argv) > 1 else "1000x"
for out in ("/dev/null", "/dev/shm/rr_reads.fastq"):
    t = time.time()
    r = subprocess.run([sys.executable, "-m", "rustyread_b200", "-v", "simulate", "--reference", ref, "--quantity", q, "--seed", "42", "--error_model", "random",
                        "--qscore_model", "random", "--output", out], capture_output=True, text=True, env=dict(os.environ, RUSTYREAD_B200_TIMING="1"), cwd=os.path.join(os.path.dirname(__file__), ".."))
    dt = time.time() - t
    print(out, "rc", r.returncode, r.stderr.strip()[-400:], "wall %.2fs" % dt)
    wrote = [l for l in r.stderr.splitlines() if l.startswith("wrote")]
    bases = int(wrote[-1].split()[-2]) if (r.returncode == 0 and wrote) else 0
    print("  -> %.2f Gbases/s whole process (incl. python start, reference parse, 100k length draws, context creation)" % (bases / dt / 1e9))
    if out != "/dev/null" and os.path.exists(out):
        print("  file size %.2f GB" % (os.path.getsize(out) / 1e9)); os.remove(out)
