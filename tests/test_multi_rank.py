"""world_size-2 (gloo, CPU) test of the sharded `simulate` driver: the host-side logic of the N > 1 path
(sharding, shard transport, single-writer FASTQ with main_rng kept in step on every rank).  The per-read
generator is stood in for by the CPU oracle here (there is no GPU on this box); the result must be the same
bytes as the oracle's own single-process simulate()."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
    import numpy as np
    import torch.distributed as dist
    import oraclelib, helpers
    import rustyread_b200.capi as capi
    from rustyread_b200 import simulate as S

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    orc = oraclelib.load()
    fasta = open(%(fasta)r, "rb").read()

    class OracleGenerator:  # test stand-in for the device: same interface as simulate.DeviceGenerator
        def __init__(self, sim):
            spec = dict(seed=9, length=(2500.0, 2200.0), rates=(5.0, 5.0, 10.0), glitch=(900.0, 25.0, 25.0))
            self.w, _ = None, None
            s = dict(helpers.DEFAULT); s.update(spec)
            w = orc.workload(s["seed"]); w.set_length(*s["length"]); w.set_references(fasta, adjusted=True)
            w.set_identity(*s["identity"]); w.set_adapter(*s["adapter"]); w.set_glitch(*s["glitch"]); w.set_error_model(None)
            w.set_qscore_model("random"); w.set_rates(*s["rates"])
            self.w = w
            self.fed = 0
        def __call__(self, descs):
            raise NotImplementedError
    # the oracle generates whole batches; pick this rank's range out of it
    class Gen:
        def __init__(self, sim):
            self.o = OracleGenerator(sim); self.sim = sim
        def run(self, base_loop, lo, hi, emitted_prev):
            w = self.o.w
            for _ in range(2 * emitted_prev):
                orc.lib.orc_rng_next_u64(w.rng_handle())   # the reference writer's u128 draws of the previous batch
            n = w.feed(base_loop); w.generate(threads=2, keep_cigar_raw=False)
            seq, so = w.out("seq"); qual, _ = w.out("qual"); ident = w.out_identity()
            lens = (so[1:] - so[:-1]).astype(np.uint64)
            a, b = int(so[lo]), int(so[hi])
            return ((so[lo:hi + 1] - so[lo]).astype(np.uint64), lens[lo:hi], seq[a:b].copy(), qual[a:b].copy(), ident[lo:hi].copy())

    # drive simulate() with a generate_fn that needs the batch context: wrap feed to capture it
    state = dict(base_loop=None, emitted_prev=0, bounds=None)
    gen_holder = {}
    def factory(sim):
        g = Gen(sim); gen_holder["g"] = g
        orig_feed = sim.feed
        def feed(base_loop):
            state["base_loop"] = base_loop
            d = orig_feed(base_loop)
            state["bounds"] = S.shard_bounds(d, world)
            return d
        sim.feed = feed
        orig_skip = sim.skip_names
        def call(descs):
            lo, hi = state["bounds"][rank], state["bounds"][rank + 1]
            out = g.run(state["base_loop"], lo, hi, state["emitted_prev"])
            return out
        return call
    # emitted count of the previous batch is needed by every rank's oracle: recompute it from lengths (len > 14)
    import io
    out = io.BytesIO() if rank == 0 else None
    class CountingSim: pass
    orig_simulate_fastq = capi.Simulation.fastq
    def fastq(self, b, first=0):
        t, r, bs = orig_simulate_fastq(self, b, first)
        state["emitted_batch"] = state.get("emitted_batch", 0) + r
        return t, r, bs
    capi.Simulation.fastq = fastq
    orig_skip = capi.Simulation.skip_names
    def skip(self, n):
        state["emitted_prev"] = n
        return orig_skip(self, n)
    capi.Simulation.skip_names = skip
    orig_feed_cls = capi.Simulation.feed
    def feed_cls(self, base_loop):
        if rank == 0:
            state["emitted_prev"] = state.pop("emitted_batch", 0)
        return orig_feed_cls(self, base_loop)
    capi.Simulation.feed = feed_cls
    reads, bases = S.simulate(fasta, "10x", out, generate_factory=factory, dist=dist, rank=rank, world=world, seed=9, length="2500,2200",
                              junk_reads=5.0, random_reads=5.0, chimera=10.0, glitches="900,25,25", error_model="random", qscore_model="random",
                              number_base_store="300K")
    if rank == 0:
        open(%(out)r, "wb").write(out.getvalue())
        print("READS", reads, bases)
    dist.destroy_process_group()
''')


def test_shard_bounds_cover_and_balance():
    import rustyread_b200.capi as capi
    from rustyread_b200.simulate import shard_bounds

    rnd = np.random.RandomState(1)
    d = np.zeros(1000, dtype=capi.DESC_DTYPE)
    d["end"] = rnd.gamma(1.3, 11000, size=1000).astype(np.uint64) + 1
    for world in (1, 2, 3, 8):
        b = shard_bounds(d, world)
        assert b[0] == 0 and b[-1] == 1000 and len(b) == world + 1 and all(x <= y for x, y in zip(b, b[1:]))
        loads = [int(d["end"][b[i]:b[i + 1]].sum()) for i in range(world)]
        assert max(loads) < 1.25 * (sum(loads) / world) + 200_000
    assert shard_bounds(d[:0], 4) == [0, 0, 0, 0, 0]


def test_flag_parsers():
    from rustyread_b200.simulate import parse_duo, parse_trio

    assert parse_duo("15000,13000") == (15000, 13000) and parse_trio("85,95,5") == (85, 95, 5)
    for bad in ("1", "1,2,3", "a,b", "1.5,2"):
        with pytest.raises(ValueError):
            parse_duo(bad)
    with pytest.raises(ValueError):
        parse_trio("1,2")


def test_world_size_2_simulate_matches_single_process(orc, tmp_path):
    fasta = helpers.random_fasta(orc, [70_000, 5_000], circular={1})
    fpath, opath, wpath = tmp_path / "ref.fa", tmp_path / "out.fq", tmp_path / "worker.py"
    fpath.write_text(fasta)
    wpath.write_text(WORKER % dict(root=ROOT, fasta=str(fpath), out=str(opath)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29517", str(wpath)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    params = dict(seed=9, length="2500,2200", quantity="10x", junk_reads=5, random_reads=5, chimera=10, glitches="900,25,25", threads=2,
                  number_base_store="300K")
    text, stats = orc.simulate(params, fasta, capture=True)
    got = opath.read_bytes().decode()
    assert "READS %d %d" % (stats["reads"], stats["bases"]) in r.stdout
    assert got == text
