"""Developer tool: host timeline of the double-buffered rr_submit / rr_collect loop (set RR_TRACE=1)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
import rustyread_b200 as rr
import rustyread_b200.capi as capi

cov = int(sys.argv[1]) if len(sys.argv) > 1 else 200
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
fasta = bench.synthetic_fasta()
sim = capi.Simulation(fasta, seed=42)
ctx = rr.Context(0)
sim.attach(ctx)
descs = sim.feed(sim.genome_length * cov)
for _ in range(3):
    b = ctx.generate(descs)
t0 = time.perf_counter()
ctx.submit(descs)
bases = 0
for step in range(steps):
    if step + 1 < steps:
        ctx.submit(descs)
    bases += ctx.collect().total_bases
dt = time.perf_counter() - t0
print("e2e %.2f Gbases/s, %.1f ms/step" % (bases / dt / 1e9, dt / steps * 1e3))
