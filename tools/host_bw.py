"""Developer tool: host memory bandwidth of the box (threads x numpy copies of 256 MiB), to tell a DRAM limit from a PCIe-side
limit when the pinned D2H of several GPUs saturates.  usage: host_bw.py [threads=16] [seconds=3]"""
import sys, threading, time
import numpy as np
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 16
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
N = 256 << 20
bufs = [(np.ones(N, np.uint8), np.empty(N, np.uint8)) for _ in range(nt)]
for a, b in bufs: np.copyto(b, a)
count = [0] * nt
stop = time.time() + secs
def work(i):
    a, b = bufs[i]
    while time.time() < stop:
        np.copyto(b, a); count[i] += 1
t0 = time.time()
th = [threading.Thread(target=work, args=(i,)) for i in range(nt)]
[t.start() for t in th]; [t.join() for t in th]
dt = time.time() - t0
print("%d threads: %.1f GB/s copied (read + write = %.1f GB/s of DRAM traffic)" % (nt, sum(count) * N / dt / 1e9, 2 * sum(count) * N / dt / 1e9))
