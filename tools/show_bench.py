"""pretty-print the JSON line(s) of a bench.py log:  python tools/show_bench.py LOG"""
import json
import sys

for line in open(sys.argv[1]):
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    if d.get("impl") == "reference":
        print("reference arm:", d["value"], d["unit"], d["cpu_baseline"])
        continue
    print(f'value {d["value"]:.2f} {d["unit"]}  ms/ddim-step {d["ms_per_denoise_step"]:.2f}  e2e {d["e2e"]["value"]:.2f}  '
          f'model {d["model_tflops"]:.0f} TF/s  launches {d["gpu_launches"]}  clocks {d["clocks"]}')
    r = d["roofline"]
    print(f'roofline GEMM: {r["achieved"]:.0f} / {r["peak"]} TF/s = {r["frac"]:.3f}   share {r["share_of_step"]:.2f}')
    for key in ("roofline_all_gemms", "roofline_resid"):
        x = d.get(key)
        if x:
            print(f'{key}: {x["achieved"]:.0f} / {x["peak"]} {x["unit"]} = {x["frac"]:.3f}   share {x["share_of_step"]:.2f}')
    m = d.get("roofline_mix")
    if m:
        print(f'roofline mix : {m["achieved"]:.0f} / {m["peak"]} GB/s = {m["frac"]:.3f}   {m["avg_launch_us"]:.0f} us/launch')
    for k in d["kernels"]:
        tf = f'{k["tflops"]:.0f} TF/s' if k["tflops"] else ""
        print(f'  {k["name"]:14s} {k["ms_per_ddim_step"]:8.3f} ms {100 * k["share"]:5.1f}%  {tf}')
    if d.get("cpu_baseline"):
        print("cpu_baseline:", d["cpu_baseline"])
