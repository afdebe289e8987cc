"""Runs a few DDIM steps of the full-size model (BASELINE config[1] by default) — the command profiled by ncu:
   python tools/run_steps.py [--batch 32] [--frames 300] [--steps 2] [--mode 4] [--graph]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--frames", type=int, default=300)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--mode", type=int, default=4)
ap.add_argument("--graph", action="store_true", help="run a whole 50-step graph loop instead of eager steps")
a = ap.parse_args()
dev = torch.device("cuda", 0)
mix = bench.build_model(dev, a.mode)
eng = mix.engine(2 * a.batch, a.frames, dev)
g = torch.Generator().manual_seed(0)
cond = torch.randn(a.batch, 6144, generator=g).to(dev)
noise = torch.randn(a.batch, a.frames, 524, generator=g).to(dev)
if a.graph:
    out = eng.sample(cond, noise)
else:
    eng.set_condition(cond)
    eng.set_state(noise, noise, 49)
    for _ in range(a.steps):
        eng.step()
torch.cuda.synchronize()
print("ok, launches:", eng.hd.launches())
