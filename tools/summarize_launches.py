"""ncu launch list (--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv) ->
per-kernel table (count, total time, share, DRAM traffic) and, with --json OUT, the per-launch DRAM traffic of
every kernel that bench.py reports as `roofline.traffic`.
   python tools/summarize_launches.py LAUNCHES.csv [--last N] [--json OUT.json]"""
import argparse
import collections
import csv
import json
import re

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--last", type=int, default=0, help="only the last N launches (e.g. one DDIM step = 225)")
ap.add_argument("--json", default=None)
a = ap.parse_args()
rows = [r for r in csv.reader(open(a.csv, errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, mi, vi, ui, idi = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
launches = collections.OrderedDict()
for r in rows:
    if r is hdr or r[ki] == "Kernel Name":
        continue
    d = launches.setdefault(r[idi], {"name": re.sub(r"\(.*", "", r[ki]).replace("void mmdm::", "").replace("void ", "").replace("mmdm::", "").strip()})
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    if r[mi] == "gpu__time_duration.sum":
        unit = r[ui].strip()
        v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
    d[r[mi]] = v
items = list(launches.values())
if a.last:
    items = items[-a.last:]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in items:
    g = agg[d["name"]]
    g[0] += 1
    g[1] += d.get("gpu__time_duration.sum", 0.0)
    g[2] += d.get("dram__bytes_read.sum", 0.0)
    g[3] += d.get("dram__bytes_write.sum", 0.0)
tot = sum(v[1] for v in agg.values())
print("| kernel | launches | total us | share | DRAM read MB | DRAM write MB | DRAM GB/s |\n|---|---:|---:|---:|---:|---:|---:|")
for k, (n, t, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f}% | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {(rd + wr) / 1e3 / max(t, 1e-9):.0f} |")
print(f"| **total** | {sum(v[0] for v in agg.values())} | {tot:.1f} | 100% | | | |")
if a.json:
    out = {k: {"launches": n, "us_per_launch": t / n, "dram_bytes_per_launch": (rd + wr) / n} for k, (n, t, rd, wr) in agg.items()}
    gem = [(n, t, rd + wr) for k, (n, t, rd, wr) in agg.items() if k.startswith("gemm_bf16")]
    if gem:
        n = sum(g[0] for g in gem)
        out["gemm_class"] = {"launches": n, "us_per_launch": sum(g[1] for g in gem) / n,
                             "dram_bytes_per_launch": sum(g[2] for g in gem) / n}
    json.dump(out, open(a.json, "w"), indent=1)
