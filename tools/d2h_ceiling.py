"""Developer tool: aggregate pinned D2H bandwidth with N ranks copying concurrently (the e2e drain's platform ceiling)."""
import os, time, torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
for _ in range(2):
    h.copy_(d, non_blocking=True); torch.cuda.synchronize()
if world > 1: dist.barrier()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(10):
    h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t
x = torch.tensor([dt], device="cuda", dtype=torch.float64)
if world > 1: dist.all_reduce(x, op=dist.ReduceOp.MAX)
if rank == 0: print("N=%d: %.1f GB/s per GPU, %.1f GB/s aggregate" % (world, 10 * n / x.item() / 1e9, world * 10 * n / x.item() / 1e9))
if world > 1: dist.destroy_process_group()
