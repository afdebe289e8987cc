"""clock64 timeline of block 0 of the fused mix kernel (mmdm_debug_mix_trace): per consumer warp and frame the time
spent waiting for the rows and working on them, per producer group the time to acquire slots and to issue the copies.
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
kinds = {1: "P1 frame", 2: "P2 frame", 3: "producer group", 4: "sync after P1", 5: "sync before P2", 6: "slot acquire"}
for wp in range(16):
    ev = tr[wp][tr[wp, :, 1] > 0]
    if len(ev) == 0:
        continue
    line = [f"warp {wp:2d}: {len(ev):3d} events"]
    for kind in (1, 2, 3, 4, 5, 6):
        e = ev[(ev[:, 0] >> 32) == kind]
        if len(e) == 0:
            continue
        a_ = (e[:, 2] - e[:, 1])
        if kind == 6:
            b_ = (e[:, 3] - e[:, 2])
            line.append(f"own items: n={len(e)} wait for slot mean {a_.mean():.0f} max {a_.max()} | issue copies mean {b_.mean():.0f} max {b_.max()}")
        elif kind in (4, 5):
            line.append(f"{kinds[kind]}: n={len(e)} wait mean {a_.mean():.0f} max {a_.max()}")
        else:
            b_ = (e[:, 3] - e[:, 2])
            line.append(f"{kinds[kind]}: n={len(e)} {'acquire' if kind == 3 else 'wait'} mean {a_.mean():.0f} max {a_.max()} | "
                        f"{'issue' if kind == 3 else 'work'} mean {b_.mean():.0f} max {b_.max()}")
    print("  ".join(line))
# ring hand-over: item c (traced for producer warp 14 of block 0: the even items) reuses the slot of item c - NS; how long after that item's release did
# the producer hold the slot again (re-arm delay), and how long after the re-arm did the consumer see the data?
NS = 25
rel, ready = {}, {}
stream = []   # item index in ring order -> (phase kind, ticket)
per = T // 4 + (1 if T % 4 else 0)
for wp in range(14):
    for e in tr[wp][tr[wp, :, 1] > 0]:
        kind = int(e[0] >> 32)
        if kind in (1, 2):
            rel[(kind, int(e[0] & 0xffffffff))] = int(e[3])
            ready[(kind, int(e[0] & 0xffffffff))] = int(e[2])
nfr = min(per, T)   # block 0 = rank 0
S = max(k[1] for k in rel) // nfr + 1
def item_of(kind, ticket):
    k, fl = ticket // nfr, ticket % nfr
    if kind == 1:
        return k * nfr + max(k - 1, 0) * nfr + fl
    return min(k + 2, S) * nfr + k * nfr + fl
by_item = {item_of(k, t): (k, t) for (k, t) in rel}
acq = {}
for wp in (14, 15):
    for e in tr[wp][tr[wp, :, 1] > 0]:
        if int(e[0] >> 32) == 6 and e[2] > 0:
            acq[int(e[0] & 0xffffffff)] = (int(e[1]), int(e[2]))
d_rearm, d_data, d_wait = [], [], []
for item, (t_loc, t_acq) in sorted(acq.items()):
    prev = by_item.get(item - NS)
    cur = by_item.get(item)
    if prev is not None:
        d_rearm.append(t_acq - rel[prev])
        d_wait.append(t_acq - t_loc)
    if cur is not None:
        d_data.append(ready[cur] - t_acq)
if d_rearm:
    print(f"re-arm delay (slot released -> producer holds it again): mean {np.mean(d_rearm):.0f} median {np.median(d_rearm):.0f} "
          f"min {np.min(d_rearm)} max {np.max(d_rearm)} cycles; producer waited for the slot {np.mean(d_wait):.0f} cycles on average")
    print(f"re-arm -> consumer sees the data: mean {np.mean(d_data):.0f} median {np.median(d_data):.0f} min {np.min(d_data)} max {np.max(d_data)}")
# timeline of warp 1 (a plain consumer) and producer warp 14: first 40 events
for wp in (1, 14):
    print(f"--- warp {wp} timeline (tag kind:index, start, wait/acquire, work/issue)")
    ev = tr[wp][tr[wp, :, 1] > 0][:60]
    for e in ev:
        print(f"   {kinds[int(e[0] >> 32)]:16s} {int(e[0] & 0xffffffff):5d}  @{e[1] - t0:8d}  {e[2] - e[1]:7d}  {max(e[3] - e[2], 0):7d}")
