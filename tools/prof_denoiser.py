"""BASELINE config[2]: the interaction denoiser alone (two persons x 262-dim, batch 64, 300 frames), one forward =
one denoise step without CFG. Per-kernel-class device times (CUDA events around every launch) and the tensor-pipe
roofline of the GEMM / attention classes:   python tools/prof_denoiser.py [--batch 64] [--frames 300] [--iters 5]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from mixermdm_b200 import modules

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--frames", type=int, default=300)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
den = modules.in2INDenoiser(262, mode="interaction", dropout=0.1, **bench.FULL_DIMS["den"]).eval()
g = torch.Generator().manual_seed(1)
with torch.no_grad():
    for p in den.parameters():
        if p.abs().sum() == 0:
            p.normal_(0, 0.02, generator=g)
den = den.to(dev)
B, T = a.batch, a.frames
x = torch.randn(B, T, 524, generator=g).to(dev)
cond = torch.randn(B, 2304, generator=g).to(dev)
t = torch.full((B,), 500, dtype=torch.int64, device=dev)
with torch.no_grad():
    for _ in range(3):
        out = den(x, t, cond=cond)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        out = den(x, t, cond=cond)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    hd = den._standalone.hd if getattr(den, "_standalone", None) is not None else None
    prof = None
    if hd is not None:
        hd.profile(True)
        out = den(x, t, cond=cond)
        torch.cuda.synchronize()
        prof = hd.profile_read()
        hd.profile(False)
# 133.55 GFLOP per sample and forward (no CFG doubling)
flop = B * 133.55e9 * T / 300
peaks = bench.measured_peaks()
res = {"config": f"interaction denoiser alone, batch {B} x {T} frames, one forward", "ms_per_forward": ms,
       "model_tflops": flop / (ms * 1e-3) / 1e12, "finite": bool(torch.isfinite(out).all())}
if prof:
    tot = sum(p["ms"] for p in prof)
    res["kernels"] = [{"name": p["name"], "launches": p["launches"], "ms": p["ms"], "share": p["ms"] / tot,
                       "tflops": p["flops"] / (p["ms"] * 1e-3) / 1e12 if p["flops"] > 0 and p["ms"] > 0 else None,
                       "frac_of_sustained_peak": (p["flops"] / (p["ms"] * 1e-3) / 1e12 / peaks["tflops_sustained"])
                       if p["flops"] > 0 and p["ms"] > 0 else None}
                      for p in sorted(prof, key=lambda q: -q["ms"])]
print(json.dumps(res))
