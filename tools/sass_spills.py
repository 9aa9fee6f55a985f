"""Developer tool: LDL / STL instructions of rr_read_kernel by source line (static count, from nvdisasm -g)."""
import collections, os, re, subprocess, sys, tempfile
lib = sys.argv[1]
tmp = tempfile.mkdtemp()
subprocess.check_call("cd %s && cuobjdump -xelf all %s >/dev/null" % (tmp, os.path.abspath(lib)), shell=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run("nvdisasm -g -c %s/%s" % (tmp, cub), shell=True, capture_output=True, text=True).stdout.split("\n")
infn = False; line = None; cnt = collections.Counter(); tot = 0; n_ins = 0
for l in dis:
    if l.startswith(".text."):
        infn = "rr_read_kernel" in l; continue
    if not infn: continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m: line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m:
        n_ins += 1
        op = m.group(1)
        if re.search(r"\b(LDL|STL)\b", op.replace(".", " ")):
            cnt[line] += 1; tot += 1
print("instructions %d, LDL/STL %d" % (n_ins, tot))
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 10**9)
sel = sum(v for k, v in cnt.items() if k and k[0] == "rr_read_kernel.cuh" and lo <= k[1] <= hi)
print("in rr_read_kernel.cuh lines %d-%d: %d" % (lo, hi, sel))
for k, v in sorted(cnt.items(), key=lambda x: -x[1])[:25]:
    print(k, v)
