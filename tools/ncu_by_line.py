"""Developer tool: per-source-line executed instructions / stall samples from an ncu report
(ncu --page source --csv) joined with nvdisasm -g line info of the cubin."""
import csv, re, collections, sys, subprocess, os, tempfile
rep, lib = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
tmp = tempfile.mkdtemp()
subprocess.check_call("cd %s && cuobjdump -xelf all %s >/dev/null" % (tmp, os.path.abspath(lib)), shell=True)
cub = [f for f in os.listdir(tmp) if f.startswith("rr_kernels.")][0]
dis = subprocess.run("nvdisasm -g -c %s/%s" % (tmp, cub), shell=True, capture_output=True, text=True).stdout.split("\n")
src = subprocess.run("ncu -i %s --page source --csv" % rep, shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(src.split("\n")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(H)]
ia, ie, isamp, isrc = H.index("Address"), H.index("Instructions Executed"), H.index("# Samples"), H.index("Source")
def norm(s):
    s = re.sub(r"/\*.*?\*/", "", s); s = s.strip().rstrip(";").strip()
    s = re.sub(r"`\([^)]*\)", "X", s); s = re.sub(r"0x[0-9a-f]+", "N", s)
    return re.sub(r"\s+", " ", s)
ncu_ops = [norm(r[isrc]) for r in data]
# sections
secs = []; cur = None; name = None; line = None; KSEL = os.environ.get("KSEL", "")
for l in dis:
    if l.startswith(".text."):
        name = l[6:].rstrip(":"); cur = []; secs.append((name, cur)); line = None; continue
    if cur is None: continue
    if l.startswith(".section") or l.startswith("//----"):
        if l.startswith(".section"): cur = None
        continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m: line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m: cur.append((norm(m.group(1)), line))
agg = collections.defaultdict(lambda: [0, 0, 0]); fn = collections.defaultdict(lambda: [0, 0, 0])
tot_e = sum(int(r[ie]) for r in data); tot_s = sum(int(r[isamp]) for r in data)
used = 0
for name, ins in secs:
    if len(ins) < 8: continue
    sig = [x[0].split(" ")[0] for x in ins[:12]]
    pos = None
    for p in range(0, len(ncu_ops) - len(ins) + 1):
        if [o.split(" ")[0] for o in ncu_ops[p:p + 12]] == sig:
            # verify a longer stretch
            if all(ncu_ops[p + j].split(" ")[0] == ins[j][0].split(" ")[0] for j in range(0, len(ins), max(1, len(ins) // 50))):
                pos = p; break
    if pos is None: continue
    used += len(ins)
    for j, (op, ln) in enumerate(ins):
        r = data[pos + j]; e = int(r[ie]); s = int(r[isamp])
        agg[ln][0] += e; agg[ln][1] += s; agg[ln][2] += 1
        fn[name][0] += e; fn[name][1] += s; fn[name][2] += 1
if os.environ.get("LISTING"):  # annotated SASS listing: executions per SASS instruction, source line, opcode text
    with open(os.environ["LISTING"], "w") as f:
        for name, ins in secs:
            if len(ins) < 8: continue
            sig = [x[0].split(" ")[0] for x in ins[:12]]
            for p in range(0, len(ncu_ops) - len(ins) + 1):
                if [o.split(" ")[0] for o in ncu_ops[p:p + 12]] == sig:
                    for j, (op, ln) in enumerate(ins):
                        r = data[p + j]
                        f.write("%6d %12d %7d  %-22s %5d  %s\n" % (j, int(r[ie]), int(r[isamp]), ln[0] if ln else "?", ln[1] if ln else 0, r[isrc].strip()))
                    break
print("ncu rows %d mapped %d ; total executed %d samples %d" % (len(data), used, tot_e, tot_s))
for k, v in sorted(fn.items(), key=lambda x: -x[1][0]):
    print("FUNC %-90s exec %5.2f%% samples %5.2f%% sass %d" % (k[:90], 100 * v[0] / tot_e, 100 * v[1] / tot_s, v[2]))
for k, v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print("%-24s line %4d  exec %5.2f%%  samples %5.2f%%  sass %d" % (k[0] if k else "?", k[1] if k else 0, 100 * v[0] / tot_e, 100 * v[1] / tot_s, v[2]))
