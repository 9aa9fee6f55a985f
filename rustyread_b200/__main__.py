"""`python -m rustyread_b200 simulate ...` — the `rustyread simulate` command line (src/cli/mod.rs:10-28,
src/cli/simulate.rs:162-239) over the B200 library.  Under torchrun every rank drives its own GPU."""
import argparse
import os
import sys

from . import simulate as sim_mod


def main(argv=None):
    ap = argparse.ArgumentParser(prog="rustyread", description="A long read simulator based on badread idea (B200 hot path)")
    ap.add_argument("-t", "--threads", type=int, default=None, help="accepted for compatibility; the per-read work runs on the GPU")
    ap.add_argument("-v", "--verbosity", action="count", default=0)
    sub = ap.add_subparsers(dest="cmd", required=True)
    s = sub.add_parser("simulate", help="Generate fake long read")
    s.add_argument("--reference", required=True)
    s.add_argument("--output", default=None)
    s.add_argument("--quantity", required=True)
    d = sim_mod.FLAG_DEFAULTS
    for name in ("length", "identity", "error_model", "qscore_model", "start_adapter", "end_adapter", "start_adapter_seq", "end_adapter_seq", "glitches"):
        s.add_argument("--" + name, default=d[name])
    s.add_argument("--seed", type=int, default=None)
    for name in ("junk_reads", "random_reads", "chimera"):
        s.add_argument("--" + name, type=float, default=d[name])
    s.add_argument("--small_plasmid_bias", action="store_true")
    s.add_argument("--number_base_store", default=None)
    a = ap.parse_args(argv)
    import time

    t_start = time.perf_counter()

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    from .placement import bind_to_gpu_numa
    bind_to_gpu_numa(local)  # keep the pinned drain buffers on the GPU's socket
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("gloo")  # shard results travel host to host; the data path itself has no collective
    # Several ranks: rank 0 alone reads and parses the reference; the others get its index and receive the bases device to
    # device over NVLink (simulate.shared_reference_context).  RUSTYREAD_B200_NO_SHARED_REFERENCE=1: every rank reads the file.
    share = world > 1 and not os.environ.get("RUSTYREAD_B200_NO_SHARED_REFERENCE")
    if share:
        import torch

        share = torch.cuda.is_available() and dist.is_nccl_available()
    if os.environ.get("RUSTYREAD_B200_TIMING"):
        print("rustyread_b200 timing rank %d: imports + process group %.3f s" % (rank, time.perf_counter() - t_start), file=sys.stderr, flush=True)
    fasta = sim_mod.read_maybe_compressed(a.reference) if (rank == 0 or not share) else None  # niffler: plain, gz, bz2 or xz by magic number
    if os.environ.get("RUSTYREAD_B200_TIMING"):
        print("rustyread_b200 timing rank %d: reference read at %.3f s" % (rank, time.perf_counter() - t_start), file=sys.stderr, flush=True)
    out = (open(a.output, "w+b") if a.output else sys.stdout.buffer) if rank == 0 else None  # read-write: the threaded writer maps the file
    flags = {k: getattr(a, k) for k in d}
    chunk = int(os.environ.get("RUSTYREAD_B200_CHUNK_BASES", "1000000000"))  # device chunk size inside one reference batch
    # several ranks + a regular output file: every rank writes its own records in place (RUSTYREAD_B200_SINGLE_WRITER=1: gather on rank 0)
    par = a.output if (world > 1 and a.output and not os.environ.get("RUSTYREAD_B200_SINGLE_WRITER")) else None
    reads, bases = sim_mod.simulate(fasta, a.quantity, out, dist=dist, rank=rank, world=world, device=local, chunk_bases=chunk, output_path=par, share_reference=share, **flags)
    if rank == 0 and a.verbosity:
        print("wrote %d reads, %d bases" % (reads, bases), file=sys.stderr)
    if os.environ.get("RUSTYREAD_B200_TIMING"):
        print("rustyread_b200 timing rank %d: main() %.3f s" % (rank, time.perf_counter() - t_start), file=sys.stderr)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
