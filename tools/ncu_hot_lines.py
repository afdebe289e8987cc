"""Attribute ncu warp-stall samples (source page, SASS view) to CUDA source lines using nvdisasm line info.
   python tools/ncu_hot_lines.py <report.ncu-rep> <cubin> <mangled kernel name substring> [top]"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
sass = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
# instruction index -> (file, line) for the kernel's text section
lines, cur, inside = [], None, False
for ln in sass:
    if ln.startswith("//---------------------") and ".text." in ln:
        inside = kname in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines.append(cur)
if rep.endswith(".csv"):   # a source page exported beforehand: ncu -i REP --page source --csv [-k regex:NAME] > file.csv
    out = open(rep).read()
else:
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stalls = [h for h in hdr if h.startswith("stall_") and "Not" not in h]
agg = defaultdict(lambda: defaultdict(float))
n_inst = defaultdict(float)
if len(data) != len(lines):
    print(f"warning: {len(data)} profiled instructions vs {len(lines)} disassembled", file=sys.stderr)
for i, r in enumerate(data):
    key = lines[i] if i < len(lines) else None
    for s in stalls:
        try:
            agg[key][s] += float(r[ix[s]])
        except ValueError:
            pass
    try:
        n_inst[key] += float(r[ix["Instructions Executed"]])
    except ValueError:
        pass
tot = sum(sum(v.values()) for v in agg.values())
print(f"total samples {tot:.0f}")
for key, v in sorted(agg.items(), key=lambda kv: -sum(kv[1].values()))[:top]:
    n = sum(v.values())
    best = sorted(v.items(), key=lambda kv: -kv[1])[:3]
    print(f"{n:7.0f} {100 * n / tot:5.1f}%  {key}  inst {n_inst[key]:.0f}  " + ", ".join(f"{s[6:]} {c:.0f}" for s, c in best if c > 0))
