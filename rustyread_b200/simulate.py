"""`rustyread simulate` driver over the C ABI: the batch loop of src/simulate/mod.rs:129-215, with the
`par_bridge().map(generate_read)` stage replaced by the device library and, optionally, sharded over the
ranks of a torch.distributed job (one process per GPU, no collective on the data path).

Sharding: every rank runs the same seeded serial feeder (it is O(10) draws per read), takes a
contiguous, length-balanced range of the batch's work items, generates it on its own GPU, and ships the packed
result to rank 0, which is the only rank that formats FASTQ (read names are drawn from main_rng in write
order, src/simulate/mod.rs:199-203).  The other ranks skip the same number of name draws so that the next
batch's feeder stays in step.  The file is byte-identical to the single-GPU (and to the reference
`--threads 1`) output.
"""
import numpy as np

from . import capi

FLAG_DEFAULTS = dict(  # src/cli/simulate.rs:165-238
    length="15000,13000", identity="85,95,5", error_model="nanopore2020", qscore_model="nanopore2020", seed=None,
    start_adapter="90,60", end_adapter="50,20", start_adapter_seq="AATGTACTTCGTTCAGTTACGTATTGCT",
    end_adapter_seq="GCAATACGTAACTGAACGAAGT", junk_reads=1.0, random_reads=1.0, chimera=1.0, glitches="10000,25,25",
    small_plasmid_bias=False, number_base_store=None,
)


def read_maybe_compressed(path):
    """Whole file, decompressed by magic number like niffler::get_reader does for the reference and the model files
    (src/simulate/mod.rs:46,96,121; Cargo.toml:38 enables gz, bz2 and lzma/xz)."""
    raw = open(path, "rb").read()
    if raw[:2] == b"\x1f\x8b":
        import gzip

        return gzip.decompress(raw)
    if raw[:3] == b"BZh":
        import bz2

        return bz2.decompress(raw)
    if raw[:6] == b"\xfd7zXZ\x00":
        import lzma

        return lzma.decompress(raw)
    return raw


def found_model(value, model_type, search_path=None):
    """cli::simulate::found_model (src/cli/simulate.rs:130-159): an existing file is taken as it is; anything else is a
    model NAME looked up as badread/<model_type>_models/<name>.gz along python's sys.path (where an installed badread
    keeps `nanopore2020` and friends).  Like PathBuf::set_extension, a dotted name loses its last extension for `.gz`.
    Raises FileNotFoundError (the reference: Cli::CantFoundModelPath) when the search fails."""
    import os
    import sys

    if os.path.isfile(value):
        return value
    stem = value
    base = os.path.basename(value)
    if "." in base.lstrip("."):  # set_extension replaces the text after the last dot of the file name
        stem = value[: len(value) - (len(base) - base.rindex("."))]
    for root in (sys.path if search_path is None else search_path):
        cand = os.path.join(root or ".", "badread", "%s_models" % model_type, stem + ".gz")
        if os.path.isfile(cand):
            return cand
    raise FileNotFoundError("can't find %s model %r: not a file, and no badread/%s_models/%s.gz along python's sys.path "
                            "(src/cli/simulate.rs:130-159); pass a model file, 'random' or (qscore) 'ideal'" % (model_type, value, model_type, stem))


def agree_on_seed(dist, rank, world):
    """No --seed: the reference seeds main_rng from the unix time in seconds (src/simulate/mod.rs:31-36).  Every rank of
    a multi-process run replays the same serial feeder, so they must all use ONE seed: rank 0 reads the clock and
    broadcasts it (ranks started around a second boundary would otherwise simulate different batches and the
    parallel writer would stitch records of different simulations together, or hang on unequal round counts)."""
    import time

    seed = int(time.time())
    if world > 1:
        if dist is None:
            raise ValueError("simulate: world > 1 without a torch.distributed handle and without --seed")
        import torch

        t = torch.tensor([seed if rank == 0 else 0], dtype=torch.int64)
        dist.broadcast(t, 0)
        seed = int(t.item())
    return seed


def parse_duo(text):  # cli::simulate::Duo, src/cli/simulate.rs:82-101
    parts = text.split(",")
    if len(parts) != 2 or not all(p.isdigit() for p in parts):
        raise ValueError("can't parse %r as <u64>,<u64>" % text)
    return tuple(int(p) for p in parts)


def parse_trio(text):  # cli::simulate::Trio, src/cli/simulate.rs:103-125
    parts = text.split(",")
    if len(parts) != 3 or not all(p.isdigit() for p in parts):
        raise ValueError("can't parse %r as <u64>,<u64>,<u64>" % text)
    return tuple(int(p) for p in parts)


def shard_bounds(descs, world):
    """Contiguous split of a batch into `world` ranges with about the same number of bases each."""
    n = len(descs)
    if world <= 1 or n == 0:
        return [0, n] + [n] * max(0, world - 1)
    w = (descs["end"].astype(np.int64) - descs["start"].astype(np.int64))
    w = np.where(descs["type"] == capi.RR_REAL, np.abs(w), descs["end"].astype(np.int64))
    cw = (descs["cend"].astype(np.int64) - descs["cstart"].astype(np.int64))
    w = w + np.where(descs["has_chimera"] == 1, np.abs(cw), 0) + 64
    csum = np.cumsum(w)
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(csum, csum[-1] * r / world)))
    bounds.append(n)
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds


class DeviceGenerator:
    """generate_fn backed by the CUDA library: descs -> (offsets, lengths, seq, qual, real_identity)."""

    def __init__(self, sim, device=0, ctx=None):
        if ctx is None:
            ctx = capi.Context(device)
            sim.attach(ctx)
        self.ctx = ctx  # a context that already holds the session's references and models

    def batch(self, descs):
        """The library-owned batch (pinned arenas, valid until the next call); None for an empty shard."""
        return self.ctx.generate(descs) if len(descs) else None

    def __call__(self, descs):
        if len(descs) == 0:
            return (np.zeros(1, np.uint64), np.zeros(0, np.uint64), np.zeros(0, np.uint8), np.zeros(0, np.uint8), np.zeros(0, np.float64))
        b = self.ctx.generate(descs)
        tb = b.total_bytes
        return (b.offsets.copy(), b.lengths.copy(), b.seq[:tb].copy(), b.qual[:tb].copy(), b.real_identity.copy())


def _send(dist, arrays, dst):
    import torch

    meta = torch.tensor([len(a) for a in arrays], dtype=torch.int64)
    dist.send(meta, dst)
    for a in arrays:
        if len(a):
            dist.send(torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).copy()), dst)


def _recv(dist, dtypes, src):
    import torch

    meta = torch.zeros(len(dtypes), dtype=torch.int64)
    dist.recv(meta, src)
    out = []
    for n, dt in zip(meta.tolist(), dtypes):
        nbytes = n * np.dtype(dt).itemsize
        t = torch.zeros(nbytes, dtype=torch.uint8)
        if nbytes:
            dist.recv(t, src)
        out.append(t.numpy().view(dt))
    return out


class _DeviceBytes:
    """A device address as something torch.as_tensor understands (no copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def shared_reference_context(sim, dist, rank, device, n_bytes):
    """One rank of a box read and parsed the FASTA; its device copy of the reference goes to the GPUs of the other ranks
    over NVLink (one NCCL broadcast of the concatenated contigs) instead of every rank parsing the file and uploading it
    over PCIe.  This is the only collective that moves reference or read data, and it runs once, before the first batch.
    Returns a device context with references and models attached."""
    import torch

    torch.cuda.set_device(device)
    ctx = capi.Context(device)
    group = dist.new_group(backend="nccl")
    if rank == 0:
        sim.attach(ctx)
        ptr, n = ctx.references_device()
        if n != n_bytes:
            raise RuntimeError("reference arena holds %d bytes, index says %d" % (n, n_bytes))
        t = torch.as_tensor(_DeviceBytes(ptr, n), device="cuda:%d" % device)
    else:
        t = torch.empty(n_bytes, dtype=torch.uint8, device="cuda:%d" % device)
    dist.broadcast(t, 0, group=group)
    torch.cuda.synchronize()
    if rank != 0:
        sim.attach(ctx, device_bases=t.data_ptr())
    del t
    dist.destroy_process_group(group)
    return ctx


def piece_bounds(descs, world, chunk_bases):
    """Split a reference batch into rounds x world contiguous pieces of about the same number of bases, each at most
    ~chunk_bases: in round j rank r takes piece j * world + r, so the file order (= canonical read order) is round-major
    and the offsets of a round are known as soon as the sizes of that round are."""
    n = len(descs)
    w = np.where(descs["type"] == capi.RR_REAL, np.abs(descs["end"].astype(np.int64) - descs["start"].astype(np.int64)), descs["end"].astype(np.int64))
    cw = np.abs(descs["cend"].astype(np.int64) - descs["cstart"].astype(np.int64))
    w = w + np.where(descs["has_chimera"] == 1, cw, 0) + 64
    csum = np.cumsum(w) if n else np.zeros(1, np.int64)
    total = int(csum[-1]) if n else 0
    rounds = max(1, -(-total // max(1, world * chunk_bases)))
    pieces = rounds * world
    bounds = [0] + [int(np.searchsorted(csum, total * i / pieces)) for i in range(1, pieces)] + [n]
    for i in range(1, len(bounds)):
        bounds[i] = min(n, max(bounds[i], bounds[i - 1]))
    return rounds, bounds


def _write_batch_in_place(sim, gen, descs, dist, rank, world, pw, chunk_bases):
    """Parallel writer of one reference batch: in every round this rank generates one piece, formats its records and
    writes them at their final byte offset of the shared output file.  Read names are drawn from main_rng in write
    order (src/simulate/mod.rs:199-203), so a rank first learns how many records the pieces before its own emit (reads
    of 14 bases or fewer are dropped, mod.rs:192-194), skips that many name draws, draws its own, and skips the rest:
    all main_rng copies leave the round in the same state.  One tiny all_gather per round, no data-path collective.
    On the device path the next round's piece is generated while this one is drained in ~32 MiB parts and written with
    writev straight from pinned memory."""
    import os
    import time

    import torch

    acc = pw.setdefault("acc", dict(generate=0.0, exchange=0.0, write=0.0))
    rounds, pb = piece_bounds(descs, world, chunk_bases)
    device = hasattr(gen, "ctx")
    mine = [(pb[j * world + rank], pb[j * world + rank + 1]) for j in range(rounds)]
    if device and mine[0][1] > mine[0][0]:
        gen.ctx.submit(descs[mine[0][0]:mine[0][1]])
    tot_reads = tot_bases = 0
    for j, (lo, hi) in enumerate(mine):
        t0 = time.perf_counter()
        if device and j + 1 < rounds and mine[j + 1][1] > mine[j + 1][0]:
            gen.ctx.submit(descs[mine[j + 1][0]:mine[j + 1][1]])  # generated while this round is drained
        b = None
        nbytes = recs = bases = 0
        if hi > lo:
            if device:
                _, _, lens, ident = gen.ctx.collect_begin()
            else:
                off, lens, seq, qual, ident = gen(descs[lo:hi])
                b = capi.batch_struct_from_arrays(off, lens, seq if len(seq) else np.zeros(1, np.uint8), qual if len(qual) else np.zeros(1, np.uint8), ident, sim.k)
            nbytes, recs, bases = sim.fastq_bytes(lens, ident, first=lo)
        t1 = time.perf_counter()
        t = torch.tensor([recs, nbytes, bases], dtype=torch.int64)
        got = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(got, t)
        got = [[int(x) for x in g.tolist()] for g in got]
        sim.skip_names(sum(g[0] for g in got[:rank]))
        t2 = time.perf_counter()
        pos = pw["pos"] + sum(g[1] for g in got[:rank])
        wrote = 0
        if hi > lo:
            if device:
                os.lseek(pw["fd"], pos, os.SEEK_SET)
                for first, part in gen.ctx.parts():
                    _, _, nb = sim.write_fastq(part.struct, pw["fd"], first=lo + first)
                    wrote += nb
            else:
                text, r2, _ = sim.fastq(b, first=lo)
                view = memoryview(text)
                while len(view):
                    n = os.pwrite(pw["fd"], view, pos + wrote)
                    view = view[n:]
                    wrote += n
        if wrote != nbytes:
            raise RuntimeError("parallel writer: %d bytes written, %d expected" % (wrote, nbytes))
        sim.skip_names(sum(g[0] for g in got[rank + 1:]))
        pw["pos"] += sum(g[1] for g in got)
        tot_reads += sum(g[0] for g in got)
        tot_bases += sum(g[2] for g in got)
        t3 = time.perf_counter()
        acc["generate"] += t1 - t0
        acc["exchange"] += t2 - t1
        acc["write"] += t3 - t2
    return tot_reads, tot_bases


def chunk_bounds(descs, chunk_bases):
    """Split a batch into contiguous device chunks of about `chunk_bases` bases (whole batch if it is smaller)."""
    n = len(descs)
    if n == 0:
        return [0, 0]
    w = np.where(descs["type"] == capi.RR_REAL, np.abs(descs["end"].astype(np.int64) - descs["start"].astype(np.int64)), descs["end"].astype(np.int64)) + 64
    csum = np.cumsum(w)
    k = max(1, int(round(csum[-1] / max(1, chunk_bases))))
    bounds = [0] + [int(np.searchsorted(csum, csum[-1] * i / k)) for i in range(1, k)] + [n]
    return sorted(set(bounds))


def _simulate_single_gpu(sim, total, limit, output, device, chunk_bases):
    """One GPU, pipelined: while chunk c is drained and written (writev from pinned memory), chunk c+1 is generated.
    Chunks are sub-ranges of ONE reference batch, so main_rng sees exactly the reference's order: all Fragments draws
    of the batch, then one u128 per emitted read in canonical order (src/simulate/mod.rs:163-203)."""
    import os
    import sys
    import time

    timing = os.environ.get("RUSTYREAD_B200_TIMING")  # developer aid: where the wall clock of a run goes (stderr)
    acc = dict(context=0.0, feed=0.0, submit=0.0, collect=0.0, write=0.0)
    t0 = time.perf_counter()
    ctx = capi.Context(device)
    sim.attach(ctx)
    acc["context"] = time.perf_counter() - t0
    try:
        fd = output.fileno()
        output.flush()
    except Exception:
        fd = None
    if fd is not None:
        # a regular output file is filled by several threads (rr_sim_set_writer_threads); this process is its only writer
        wt = os.environ.get("RUSTYREAD_B200_WRITER_THREADS")
        sim.set_writer_threads(int(wt) if wt else max(1, min(8, (os.cpu_count() or 2) // 2)))
    produced, n_reads, n_bases = 0, 0, 0
    while produced < total:
        base_loop = min(limit, total - produced)
        produced += base_loop
        t = time.perf_counter()
        descs = sim.feed(base_loop)
        acc["feed"] += time.perf_counter() - t
        cb = chunk_bounds(descs, chunk_bases)
        nchunks = len(cb) - 1
        if nchunks <= 0:
            continue
        t = time.perf_counter()
        ctx.submit(descs[cb[0]:cb[1]])
        acc["submit"] += time.perf_counter() - t
        for c in range(nchunks):
            if c + 1 < nchunks:
                t = time.perf_counter()
                ctx.submit(descs[cb[c + 1]:cb[c + 2]])
                acc["submit"] += time.perf_counter() - t
            t = time.perf_counter()
            for first, b in ctx.collect_parts():  # ~32 MiB parts through a small pinned ring; D2H of part p + 1 overlaps the write of part p
                t1 = time.perf_counter()
                acc["collect"] += t1 - t
                if fd is not None:
                    recs, bases, _ = sim.write_fastq(b.struct, fd, first=cb[c] + first)
                else:
                    text, recs, bases = sim.fastq(b.struct, first=cb[c] + first)
                    output.write(text)
                t = time.perf_counter()
                acc["write"] += t - t1
                n_reads += recs
                n_bases += bases
            acc["collect"] += time.perf_counter() - t
    t = time.perf_counter()
    ctx.close()
    acc["close"] = time.perf_counter() - t
    if timing:
        print("rustyread_b200 timing (s): " + ", ".join("%s %.3f" % kv for kv in acc.items()) + ", total %.3f" % (time.perf_counter() - t0), file=sys.stderr)
    return n_reads, n_bases


def simulate(fasta, quantity, output, generate_factory=None, dist=None, rank=0, world=1, device=0, chunk_bases=1_000_000_000, output_path=None,
             share_reference=False, **flags):
    """Run `rustyread simulate`.  `output` is a binary file object (only used on rank 0).  Returns (reads, bases).

    share_reference (world > 1, device path): only rank 0 needs `fasta`; the other ranks build their session from the contig
    index rank 0 broadcasts and receive the bases device to device (shared_reference_context).

    world > 1 with `output_path` (a regular file every rank can open): every rank formats the records of its own shard
    and writes them at their final offset with pwrite (parallel writer); otherwise the shards travel to rank 0, the
    single writer.  Both produce the same bytes."""
    f = dict(FLAG_DEFAULTS)
    f.update({k: v for k, v in flags.items() if v is not None})
    if f["seed"] is None:
        f["seed"] = agree_on_seed(dist, rank, world)  # src/simulate/mod.rs:31-36

    def model_text(v, kind, builtins):
        return v if v in builtins else read_maybe_compressed(found_model(v, kind))

    kw = dict(seed=f["seed"], length=parse_duo(f["length"]), identity=parse_trio(f["identity"]),
              start_adapter=parse_duo(f["start_adapter"]), end_adapter=parse_duo(f["end_adapter"]),
              start_adapter_seq=f["start_adapter_seq"], end_adapter_seq=f["end_adapter_seq"],
              junk_reads=f["junk_reads"], random_reads=f["random_reads"], chimera=f["chimera"],
              glitches=parse_trio(f["glitches"]), small_plasmid_bias=f["small_plasmid_bias"],
              error_model=model_text(f["error_model"], "error", ("random",)),
              qscore_model=model_text(f["qscore_model"], "qscore", ("random", "ideal")))
    share = bool(share_reference) and world > 1 and generate_factory is None
    import os
    import sys
    import time

    timing = os.environ.get("RUSTYREAD_B200_TIMING")  # developer aid: where the set-up time of this rank goes (stderr)
    t_setup = time.perf_counter()

    def lap(what):
        nonlocal t_setup
        if timing:
            now = time.perf_counter()
            print("rustyread_b200 timing rank %d: %s %.3f s" % (rank, what, now - t_setup), file=sys.stderr, flush=True)
            t_setup = now

    if share:
        # rank 0 parses; everybody else gets the contig index (ids, lengths, circular flags, header weights): the serial
        # feeder and the writer need nothing else of the reference, and main_rng is drawn from in the same order
        box = [None]
        if rank == 0:
            try:
                sim = capi.Simulation(fasta, **kw)
                box = [sim.contig_index()]
            except Exception as e:  # the other ranks must not wait for an index that never comes
                box = [RuntimeError("rank 0 could not load the reference: %s" % e)]  # a plain exception: it has to survive pickling
        lap("parse reference + models (rank 0)")
        dist.broadcast_object_list(box, 0)
        if isinstance(box[0], Exception):
            raise box[0]
        if rank != 0:
            sim = capi.Simulation(None, index=box[0], **kw)
        lap("contig index received, session ready")
        shared_ctx = shared_reference_context(sim, dist, rank, device, sum(c[1] for c in box[0]))
        lap("device context + reference broadcast")
        generate_factory = lambda s: DeviceGenerator(s, device, ctx=shared_ctx)
    else:
        sim = capi.Simulation(fasta, **kw)
        lap("parse reference + models")
    total = sim.quantity(quantity)
    limit = sim.quantity(f["number_base_store"]) if f["number_base_store"] else total
    if world == 1 and generate_factory is None:
        return _simulate_single_gpu(sim, total, limit, output, device, chunk_bases)
    gen = (generate_factory or (lambda s: DeviceGenerator(s, device)))(sim)
    produced, n_reads, n_bases = 0, 0, 0
    pw = None
    if world > 1 and output_path is not None:
        import os

        if rank == 0:
            if output is not None:
                output.flush()
            os.close(os.open(output_path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644))
        dist.barrier()
        pw = dict(fd=os.open(output_path, os.O_WRONLY), pos=0)
    DT = (np.uint64, np.uint64, np.uint8, np.uint8, np.float64)
    while produced < total:  # src/simulate/mod.rs:153-214
        base_loop = min(limit, total - produced)
        produced += base_loop
        descs = sim.feed(base_loop)
        bounds = shard_bounds(descs, world)
        lo, hi = bounds[rank], bounds[rank + 1]
        if world > 1 and output_path is not None:
            r, b = _write_batch_in_place(sim, gen, descs, dist, rank, world, pw, chunk_bases)
            n_reads += r
            n_bases += b
            continue
        mine = gen(descs[lo:hi])
        if world > 1:
            import torch

            if rank == 0:
                shards = [mine] + [_recv(dist, DT, r) for r in range(1, world)]
            else:
                _send(dist, mine, 0)
                shards = None
        else:
            shards = [mine]
        emitted = 0
        if rank == 0:
            for r, (off, ln, seq, qual, ident) in enumerate(shards):
                if len(ln) == 0:
                    continue
                b = capi.batch_struct_from_arrays(off, ln, seq if len(seq) else np.zeros(1, np.uint8), qual if len(qual) else np.zeros(1, np.uint8), ident, sim.k)
                text, recs, bases = sim.fastq(b, first=bounds[r])
                output.write(text)
                emitted += recs
                n_bases += bases
            n_reads += emitted
        if world > 1:
            import torch

            t = torch.tensor([emitted], dtype=torch.int64)
            dist.broadcast(t, 0)
            if rank != 0:
                sim.skip_names(int(t.item()))
    if pw is not None:
        import os
        import sys

        os.close(pw["fd"])
        dist.barrier()
        if os.environ.get("RUSTYREAD_B200_TIMING") and "acc" in pw:  # developer aid: where the wall clock of this rank went
            print("rustyread_b200 timing rank %d (s): " % rank + ", ".join("%s %.3f" % kv for kv in pw["acc"].items()), file=sys.stderr)
    elif rank == 0:
        output.flush()
    return n_reads, n_bases
