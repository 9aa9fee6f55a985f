"""Developer tool: per-kernel summary of an `ncu --csv --metrics ...` launch list (one row per launch x metric).
usage: launch_summary.py launches.csv out_summary.json [traffic.json]"""
import collections, csv, json, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) >= 15 and r[0].isdigit()]
per = collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows:
    per[r[4].split("(")[0]][r[12]].append(float(r[14].replace(",", "")))
tot = sum(sum(m.get("gpu__time_duration.sum", [])) for m in per.values())
out = {}
for k, m in sorted(per.items(), key=lambda kv: -sum(kv[1].get("gpu__time_duration.sum", [0]))):
    t = m.get("gpu__time_duration.sum", [])
    mean = lambda name: (sum(m[name]) / len(m[name])) if m.get(name) else None
    out[k] = {"launches": len(t), "mean_ms": sum(t) / len(t) / 1e6 if t else None, "share": sum(t) / tot if tot else None,
              "dram_read_bytes_per_launch": mean("dram__bytes_read.sum"), "dram_write_bytes_per_launch": mean("dram__bytes_write.sum"),
              "warp_instructions_per_launch": mean("smsp__inst_executed.sum"), "issue_active_pct": mean("smsp__issue_active.avg.pct_of_peak_sustained_active")}
json.dump(out, open(sys.argv[2], "w"), indent=1)
if len(sys.argv) > 3:
    k = [x for x in out if "rr_read_kernel" in x][0]
    json.dump({"kernel": "rr_read_kernel", "workload": "bench.py default (ref5m 200x)",
               "dram_bytes_per_launch": out[k]["dram_read_bytes_per_launch"] + out[k]["dram_write_bytes_per_launch"],
               "dram_read_bytes_per_launch": out[k]["dram_read_bytes_per_launch"], "dram_write_bytes_per_launch": out[k]["dram_write_bytes_per_launch"],
               "warp_instructions_per_launch": out[k]["warp_instructions_per_launch"],
               "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,... --clock-control none (%s)" % sys.argv[1]},
              open(sys.argv[3], "w"), indent=1)
print(json.dumps(out, indent=1)[:1500])
