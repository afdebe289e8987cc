"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel table (count, total, share)."""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    if r is hdr or len(r) <= vi or r[ki] == "Kernel Name":
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui].strip().replace("nsecond", "ns").replace("usecond", "us").replace("msecond", "ms"), 1e-3)
    name = re.sub(r"\(.*", "", r[ki]).replace("void mmdm::", "").strip()
    agg[name][0] += 1
    agg[name][1] += v * scale
tot = sum(v[1] for v in agg.values())
print(f"| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f}% |")
print(f"| **total** | {sum(v[0] for v in agg.values())} | {tot:.1f} | 100% |")
