"""clock64 timeline of block 0 of the fused mix kernel (mmdm_debug_mix_trace): per warp and frame the time spent
waiting for the frame's rows and working on them (incl. re-arming the slot), and the waits at the consumer barriers.
   python tools/mix_trace.py [--batch 128] [--mode 4]"""
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
ap.add_argument("--mode", type=int, default=4)
a = ap.parse_args()
lib = _lib.load()
B, T, dev, mode = a.batch, a.frames, "cuda", a.mode
sched = schedule.SpacedSchedule(schedule.space_timesteps(1000, "ddim50"), schedule.get_named_beta_schedule("cosine", 1000))
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g)  # noqa: E731
preds, state = r(2 * B, T, 2, 524), r(B, T, 2, 524)   # the engine's layout: [out1a row | out2a row], [x1 row | x2 row]
out1a, out2a, x1, x2 = preds[:, :, 0], preds[:, :, 1], state[:, :, 0], state[:, :, 1]
fin = torch.empty(B, T, 524, device=dev)
coef = torch.from_numpy(sched.ddim_coefficients()).to(dev)
step = torch.tensor([17], dtype=torch.int32, device=dev)
m1, s1, m2, s2 = r(262) * 0.1, r(262).abs() + 0.5, r(262) * 0.1, r(262).abs() + 0.5
pooled = mode in (1, 3)
w = torch.sigmoid(r(2 * B * 2 * (1 if pooled else T), 24))
args = MixstepArgs()
args.B, args.T, args.mixing_mode, args.ncfg, args.mix_only, args.center, args.has_force = B, T, mode, 2, 0, 1, 0
args.force_val, args.cfg_scale, args.nsteps = 0.0, 3.5, 50
for name, t in [("step_ptr", step), ("coef", coef), ("weights", w), ("out1a", out1a), ("out2a", out2a), ("x1", x1),
                ("x2", x2), ("final_out", fin), ("mean1", m1), ("std1", s1), ("mean2", m2), ("std2", s2)]:
    setattr(args, name, t.data_ptr())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    _lib.check(lib.mmdm_mix_cfg_ddim(C.byref(args), _lib.cur_stream()), "mix")
torch.cuda.synchronize()
NE = 512
trace = torch.zeros(16 * NE * 4, dtype=torch.int64, device=dev)
lib.mmdm_debug_mix_trace.argtypes = [C.c_void_p]
assert lib.mmdm_debug_mix_trace(C.c_void_p(trace.data_ptr())) == 0
flush.zero_()
_lib.check(lib.mmdm_mix_cfg_ddim(C.byref(args), _lib.cur_stream()), "mix")
torch.cuda.synchronize()
lib.mmdm_debug_mix_trace(C.c_void_p(0))
tr = trace.cpu().numpy().reshape(16, NE, 4)
t0 = tr[:, :, 1][tr[:, :, 1] > 0].min()
tend = tr[:, :, 1:].max()
print(f"block 0 active for {tend - t0} cycles")
kinds = {1: "P1 frame", 2: "P2 frame", 4: "sync after P1", 5: "sync before P2"}
for wp in range(16):
    ev = tr[wp][tr[wp, :, 1] > 0]
    if len(ev) == 0:
        continue
    line = [f"warp {wp:2d}: {len(ev):3d} events"]
    for kind in (1, 2, 4, 5):
        e = ev[(ev[:, 0] >> 32) == kind]
        if len(e) == 0:
            continue
        a_ = (e[:, 2] - e[:, 1])
        if kind in (4, 5):
            line.append(f"{kinds[kind]}: n={len(e)} wait mean {a_.mean():.0f} max {a_.max()}")
        else:
            b_ = (e[:, 3] - e[:, 2])
            line.append(f"{kinds[kind]}: n={len(e)} {'acquire' if kind == 3 else 'wait'} mean {a_.mean():.0f} max {a_.max()} | "
                        f"{'issue' if kind == 3 else 'work'} mean {b_.mean():.0f} max {b_.max()}")
    print("  ".join(line))
# timeline of warps 0 (does the cluster exchange) and 1: first 60 events
for wp in (0, 1):
    print(f"--- warp {wp} timeline (tag kind:index, start, wait/acquire, work/issue)")
    ev = tr[wp][tr[wp, :, 1] > 0][:60]
    for e in ev:
        print(f"   {kinds[int(e[0] >> 32)]:16s} {int(e[0] & 0xffffffff):5d}  @{e[1] - t0:8d}  {e[2] - e[1]:7d}  {max(e[3] - e[2], 0):7d}")
