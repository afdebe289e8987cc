"""Per-role clock64 timeline of one attention CTA (block 0): builds an instrumented copy of the library
(-DMMDM_ATTN_TRACE, attention.cu only) and prints, per query-tile job j, when each warp role passed its events.
   python tools/attn_trace.py build      (here, needs nvcc)
   MMDM_LIB_PATH=mixermdm_b200/libmixermdm_b200_trace.so python tools/attn_trace.py run [hd] [S]   (on the GPU)"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TRACE_LIB = os.path.join(ROOT, "mixermdm_b200", "libmixermdm_b200_trace.so")

if sys.argv[1] == "build":
    from mixermdm_b200 import _build
    _build.build()
    objdir = os.path.join(_build.CSRC, "_obj")
    tobj = os.path.join(objdir, "attention_trace.o")
    subprocess.check_call([_build._nvcc()] + _build.NVCC_FLAGS + ["-DMMDM_ATTN_TRACE", "-c",
                          os.path.join(_build.CSRC, "attention.cu"), "-o", tobj])
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in _build.SOURCES if s != "attention.cu"] + [tobj]
    subprocess.check_call([_build._nvcc(), "-shared", "-o", TRACE_LIB] + objs +
                          ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-Xcompiler", "-fPIC"])
    print(TRACE_LIB)
    sys.exit(0)

import torch
from mixermdm_b200 import _lib, ops

hd = int(sys.argv[2]) if len(sys.argv) > 2 else 128
S = int(sys.argv[3]) if len(sys.argv) > 3 else 128
T, H = 300, 8
D = H * hd
M, Tp = S * T, 304
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
qk = (torch.randn(M, 2 * D, device=dev, generator=g) * 0.5).to(torch.bfloat16)
vt = (torch.randn(S * H, hd, Tp, device=dev, generator=g) * 0.5).to(torch.bfloat16)
att = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
lib = _lib.load()
trace = torch.zeros(12 * 64 * 16, dtype=torch.int64, device=dev)
for _ in range(3):
    ops.attention(qk[:, :D], qk[:, D:], vt, att, S, T, H, hd, 0)
torch.cuda.synchronize()
assert lib.mmdm_debug_attn_trace(C.c_void_p(trace.data_ptr())) == 0
ops.attention(qk[:, :D], qk[:, D:], vt, att, S, T, H, hd, 0)
torch.cuda.synchronize()
tr = trace.cpu().view(12, 64, 16)
t0 = int(tr[tr > 0].min())
SM = ["blk_begin", "S_in_regs", "prefetch_tried", "exp_done", "P_published", "wait_O", "O_ready", "tile_done", "max_done", "exp_half", "mid_try_done", "end_try_done"]
MMA = ["S_begin", "K_ok", "S_issued", "PV_begin", "P_ok", "V_ok", "PV_issued", "-"]
NB = (T + 1 + 63) // 64
print("per-warp S_in_regs / exp_done / P_published of blocks 1..8 (softmax warps 0-7)")
for n in range(1, 9):
    print(f"  block {n}: " + "  ".join(f"w{w}:{int(tr[w, n, 1]) - t0}/{int(tr[w, n, 3]) - t0}/{int(tr[w, n, 4]) - t0}" for w in range(8)))
for n in range(0, 24):
    print(f"--- block {n} (tile {n // NB} of the warpgroup, kb {n % NB})")
    for w in (0, 4, 8, 9):
        nm = SM if w < 8 else MMA
        ev = [(int(tr[w, n, e]) - t0, nm[e]) for e in range(len(nm)) if tr[w, n, e] > 0]
        ev.sort()
        print(f"  warp {w}: " + "  ".join(f"{nn}@{tt}" for tt, nn in ev))
