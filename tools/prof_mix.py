"""Stand-alone timing of the fused mix/CFG/centre/DDIM kernel (BASELINE config[3]: granularity sweep, batch 128):
   python tools/prof_mix.py [--batch 128] [--frames 300] [--iters 20]"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mixermdm_b200 import _lib, schedule
from mixermdm_b200.engine import MixstepArgs

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--frames", type=int, default=300)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--dense", action="store_true", help="separate out1a/out2a and x1/x2 buffers instead of the engine's "
                                                     "row-interleaved pairs")
a = ap.parse_args()
lib = _lib.load()
B, T, dev = a.batch, a.frames, "cuda"
sched = schedule.SpacedSchedule(schedule.space_timesteps(1000, "ddim50"), schedule.get_named_beta_schedule("cosine", 1000))
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g)
if a.dense:
    out1a, out2a = r(2 * B, T, 524), r(2 * B, T, 524)
    x1, x2 = r(B, T, 524), r(B, T, 524)
else:   # the engine's layout: [out1a row | out2a row], [x1 row | x2 row]
    preds, state = r(2 * B, T, 2, 524), r(B, T, 2, 524)
    out1a, out2a, x1, x2 = preds[:, :, 0], preds[:, :, 1], state[:, :, 0], state[:, :, 1]
fin = torch.empty(B, T, 524, device=dev)
coef = torch.from_numpy(sched.ddim_coefficients()).to(dev)
step = torch.tensor([17], dtype=torch.int32, device=dev)
m1, s1, m2, s2 = r(262) * 0.1, r(262).abs() + 0.5, r(262) * 0.1, r(262).abs() + 0.5
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for mode in (1, 2, 3, 4):
    nw = 1 if mode <= 2 else 23
    pooled = mode in (1, 3)
    rows = 2 * B * 2 * (1 if pooled else T)
    w = torch.sigmoid(r(rows, 24))
    args = MixstepArgs()
    args.B, args.T, args.mixing_mode, args.ncfg, args.mix_only, args.center, args.has_force = B, T, mode, 2, 0, 1, 0
    args.force_val, args.cfg_scale, args.nsteps = 0.0, 3.5, 50
    for name, t in [("step_ptr", step), ("coef", coef), ("weights", w), ("out1a", out1a), ("out2a", out2a), ("x1", x1),
                    ("x2", x2), ("final_out", fin), ("mean1", m1), ("std1", s1), ("mean2", m2), ("std2", s2)]:
        setattr(args, name, t.data_ptr())
    times = []
    for it in range(a.iters + 3):
        flush.zero_()                                  # L2 flush between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.mmdm_mix_cfg_ddim(C.byref(args), _lib.cur_stream()), "mix")
        e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    nbytes = B * (T * (2 * 2 * 524 * 4 + 2 * 524 * 4 + 2 * 524 * 4) + 2 * 2 * nw * 4 * (1 if pooled else T))
    print(f"mode {mode}: {ms * 1e3:8.1f} us   {nbytes / ms / 1e6:8.1f} GB/s   ({nbytes / 1e6:.1f} MB algorithmic)")
