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


def free_port():
    """A TCP port nobody listens on right now (a fixed one can still be in TIME_WAIT from the previous run of the suite)."""
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]

WORKER = textwrap.dedent('''
    import io, os, sys
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

    class OracleGenerator:
        """Stand-in for simulate.DeviceGenerator: generate_read of the CPU oracle on the work items it is given."""
        def __init__(self, sim):
            spec = dict(helpers.DEFAULT, seed=9, length=(2500.0, 2200.0), rates=(5.0, 5.0, 10.0), glitch=(900.0, 25.0, 25.0))
            w = orc.workload(spec["seed"])
            w.set_length(*spec["length"]); w.set_references(fasta, adjusted=True); w.set_identity(*spec["identity"])
            w.set_adapter(*spec["adapter"]); w.set_glitch(*spec["glitch"]); w.set_error_model(None); w.set_qscore_model("random")
            self.w = w
        def __call__(self, descs):
            w = self.w
            w.load_items(*capi.descs_to_item_table(descs))
            if len(descs) == 0:
                return (np.zeros(1, np.uint64), np.zeros(0, np.uint64), np.zeros(0, np.uint8), np.zeros(0, np.uint8), np.zeros(0, np.float64))
            w.generate(threads=2, keep_cigar_raw=False)
            seq, so = w.out("seq"); qual, _ = w.out("qual")
            return (so, (so[1:] - so[:-1]).astype(np.uint64), seq[:int(so[-1])], qual[:int(so[-1])], w.out_identity())

    parallel = %(parallel)r
    out = io.BytesIO() if rank == 0 else None
    reads, bases = S.simulate(fasta, "10x", out, generate_factory=OracleGenerator, dist=dist, rank=rank, world=world, seed=9, length="2500,2200",
                              junk_reads=5.0, random_reads=5.0, chimera=10.0, glitches="900,25,25", error_model="random", qscore_model="random",
                              number_base_store="300K", output_path=%(out)r if parallel else None, chunk_bases=%(chunk)d)
    if rank == 0:
        if not parallel:
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


def test_piece_bounds_are_round_major_and_bounded():
    """Parallel writer: rounds x world contiguous pieces covering the batch in order, each about chunk_bases at most."""
    import rustyread_b200.capi as capi
    from rustyread_b200.simulate import piece_bounds

    rs = np.random.RandomState(3)
    d = np.zeros(5000, dtype=capi.DESC_DTYPE)
    d["type"] = capi.RR_REAL
    d["start"] = rs.randint(0, 10 ** 6, size=len(d))
    d["end"] = d["start"] + rs.gamma(1.3, 11000, size=len(d)).astype(np.uint64)
    total = int((d["end"] - d["start"]).sum()) + 64 * len(d)
    for world, chunk in ((2, 10 ** 9), (3, 4_000_000), (8, 1_500_000)):
        rounds, b = piece_bounds(d, world, chunk)
        assert len(b) == rounds * world + 1 and b[0] == 0 and b[-1] == len(d) and all(x <= y for x, y in zip(b, b[1:]))
        assert rounds == max(1, -(-total // (world * chunk)))
        w = (d["end"] - d["start"]).astype(np.int64) + 64
        loads = [int(w[b[i]:b[i + 1]].sum()) for i in range(rounds * world)]
        assert max(loads) < 1.1 * total / (rounds * world) + 300_000
    assert piece_bounds(d[:0], 4, 10 ** 6) == (1, [0, 0, 0, 0, 0])


def test_flag_parsers():
    from rustyread_b200.simulate import parse_duo, parse_trio

    assert parse_duo("15000,13000") == (15000, 13000) and parse_trio("85,95,5") == (85, 95, 5)
    for bad in ("1", "1,2,3", "a,b", "1.5,2"):
        with pytest.raises(ValueError):
            parse_duo(bad)
    with pytest.raises(ValueError):
        parse_trio("1,2")


def test_compressed_inputs_are_sniffed_like_niffler(tmp_path):
    """Reference and model files may be plain, gz, bz2 or xz (niffler, src/simulate/mod.rs:46,96,121; Cargo.toml:38)."""
    import bz2
    import gzip
    import lzma

    from rustyread_b200.simulate import read_maybe_compressed

    text = b">c\nACGTACGT\n" * 50
    for name, data in (("plain", text), ("gz", gzip.compress(text)), ("bz2", bz2.compress(text)), ("xz", lzma.compress(text))):
        p = tmp_path / ("ref." + name)
        p.write_bytes(data)
        assert read_maybe_compressed(str(p)) == text


def test_found_model_searches_badread_models_along_sys_path(tmp_path):
    """cli::simulate::found_model (src/cli/simulate.rs:130-159): a file is itself; a name is badread/<type>_models/<name>.gz on sys.path."""
    import gzip

    from rustyread_b200.simulate import found_model, read_maybe_compressed

    f = tmp_path / "my_model.txt"
    f.write_text("x")
    assert found_model(str(f), "error") == str(f)
    site = tmp_path / "site-packages"
    (site / "badread" / "error_models").mkdir(parents=True)
    (site / "badread" / "qscore_models").mkdir(parents=True)
    (site / "badread" / "error_models" / "nanopore2020.gz").write_bytes(gzip.compress(b"ACGTACG,0.5;ACGTACC,0.5;\n"))
    got = found_model("nanopore2020", "error", search_path=["/nonexistent", str(site)])
    assert got == str(site / "badread" / "error_models" / "nanopore2020.gz")
    assert read_maybe_compressed(got).startswith(b"ACGTACG")
    # PathBuf::set_extension: "a.b" becomes "a.gz"
    (site / "badread" / "qscore_models" / "pacbio.gz").write_bytes(gzip.compress(b"q"))
    assert found_model("pacbio.2016", "qscore", search_path=[str(site)]).endswith("qscore_models/pacbio.gz")
    with pytest.raises(FileNotFoundError):
        found_model("nanopore2020", "qscore", search_path=[str(site)])
    with pytest.raises(FileNotFoundError):  # the CLI default without badread installed: the reference fails the same way
        found_model("nanopore2020", "error", search_path=[])


SEEDLESS_WORKER = textwrap.dedent('''
    import os, sys, time
    sys.path.insert(0, %(root)r)
    import torch.distributed as dist
    from rustyread_b200 import simulate as S
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    if rank == 1:
        time.sleep(1.2)  # the ranks read the clock in different seconds
    open(os.path.join(%(out)r, "seed_%%d" %% rank), "w").write(str(S.agree_on_seed(dist, rank, world)))
    dist.destroy_process_group()
''')


def test_ranks_agree_on_the_time_seed(tmp_path):
    """No --seed under torchrun: rank 0 reads the clock, everybody uses its value (src/simulate/mod.rs:31-36 has one process)."""
    wpath = tmp_path / "seedless.py"
    wpath.write_text(SEEDLESS_WORKER % dict(root=ROOT, out=str(tmp_path)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), str(wpath)]
    r = subprocess.run(cmd, env=dict(os.environ, MASTER_ADDR="127.0.0.1"), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    seeds = [int((tmp_path / ("seed_%d" % i)).read_text()) for i in range(2)]
    assert seeds[0] == seeds[1] and abs(seeds[0] - __import__("time").time()) < 120
    from rustyread_b200.simulate import agree_on_seed

    assert abs(agree_on_seed(None, 0, 1) - __import__("time").time()) < 5
    with pytest.raises(ValueError):
        agree_on_seed(None, 1, 2)


BAD_REFERENCE_WORKER = textwrap.dedent('''
    import io, os, sys
    sys.path.insert(0, %(root)r)
    import torch.distributed as dist
    from rustyread_b200 import simulate as S
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    try:
        S.simulate(b"not a fasta file" if rank == 0 else None, "10x", io.BytesIO() if rank == 0 else None, dist=dist, rank=rank, world=world, seed=1,
                   error_model="random", qscore_model="random", share_reference=True)
        msg = "no error"
    except RuntimeError as e:
        msg = "RuntimeError: " + str(e)
    open(os.path.join(%(out)r, "res_%%d" %% rank), "w").write(msg)
    dist.destroy_process_group()
''')


def test_shared_reference_load_failure_reaches_every_rank(tmp_path):
    """Rank 0 alone reads the reference; when that fails the other ranks get the error instead of waiting for a contig index."""
    wpath = tmp_path / "bad.py"
    wpath.write_text(BAD_REFERENCE_WORKER % dict(root=ROOT, out=str(tmp_path)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), str(wpath)]
    r = subprocess.run(cmd, env=dict(os.environ, MASTER_ADDR="127.0.0.1"), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for i in range(2):
        msg = (tmp_path / ("res_%d" % i)).read_text()
        assert msg.startswith("RuntimeError: rank 0 could not load the reference") and "no sequence found" in msg, msg


@pytest.mark.parametrize("parallel,world,chunk", [(False, 2, 10 ** 9), (True, 2, 10 ** 9), (True, 3, 40_000)])
def test_world_size_2_simulate_matches_single_process(orc, tmp_path, parallel, world, chunk):
    """Single writer (shards gathered on rank 0) and parallel writer (every rank formats its shard and pwrites it at its
    final offset, read-name draws kept in step through skip_names; with 3 ranks in several rounds of pieces per batch):
    same bytes as the oracle's single-process run."""
    fasta = helpers.random_fasta(orc, [70_000, 5_000], circular={1})
    fpath, opath, wpath = tmp_path / "ref.fa", tmp_path / "out.fq", tmp_path / "worker.py"
    fpath.write_text(fasta)
    wpath.write_text(WORKER % dict(root=ROOT, fasta=str(fpath), out=str(opath), parallel=parallel, chunk=chunk))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world, "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), str(wpath)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    params = dict(seed=9, length="2500,2200", quantity="10x", junk_reads=5, random_reads=5, chimera=10, glitches="900,25,25", threads=2,
                  number_base_store="300K")
    text, stats = orc.simulate(params, fasta, capture=True)
    got = opath.read_bytes().decode()
    assert "READS %d %d" % (stats["reads"], stats["bases"]) in r.stdout
    assert got == text
