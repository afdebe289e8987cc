"""Stand-alone launches of the hot kernels at BASELINE config[1] shapes (M = 38400 tokens) for ncu captures and
quick CUDA-event timings:  python tools/prof_ops.py [qkv|proj|ffn1|ffn2|attn|ln|all] [--iters N]"""
import argparse
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mixermdm_b200 import ops
from mixermdm_b200._lib import EPI_BF16, EPI_BF16_LN, EPI_GELU_BF16, EPI_GELU_BF16_LN, EPI_RESID, EPI_RESID_LN

ap = argparse.ArgumentParser()
ap.add_argument("what", nargs="?", default="all")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--S", type=int, default=128)
args = ap.parse_args()

S, T, D, F, H = args.S, 300, 1024, 2048, 8
M, hd, Tp = S * T, D // H, 304
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rb = lambda *s: (torch.randn(*s, device=dev, generator=g) * 0.5).to(torch.bfloat16)
x = rb(M, D)
w_qkv, w_o, w1, w2 = rb(3 * D, D) / 16, rb(D, D) / 16, rb(F, D) / 16, rb(D, F) / 16
b3, b1, bd, bf = [torch.randn(n, device=dev, generator=g) * 0.1 for n in (3 * D, D, D, F)]
qk = torch.empty(M, 2 * D, dtype=torch.bfloat16, device=dev)
vt = torch.zeros(S * H, hd, Tp, dtype=torch.bfloat16, device=dev)
att = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
h = torch.randn(M, D, device=dev, generator=g)
mid = torch.empty(M, F, dtype=torch.bfloat16, device=dev)
mod = torch.randn(S, 4 * D, device=dev, generator=g) * 0.1
xn = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
xf = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
stats = torch.zeros(D // 64, M, 2, device=dev)
gb3 = torch.randn(S, 2 * 3 * D, device=dev, generator=g) * 0.1
gbf = torch.randn(S, 2 * F, device=dev, generator=g) * 0.1
rowa = torch.zeros(M, 4, dtype=torch.int32, device=dev)
moff = torch.randn(M, device=dev, generator=g) * 0.1
moff_out = torch.zeros(M, device=dev)
lp0 = dict(scale0=mod[:, :D], out0=xn, stats=stats, T=T, rowa=rowa)                                   # no per-row offsets
lp1 = dict(scale0=mod[:, :D], out0=xn, stats=stats, T=T, rowa=rowa, moff_in=moff, moff_out=moff_out)
lp2 = dict(scale0=mod[:, :D], out0=xn, scale1=mod[:, 2 * D:3 * D], out1=xf, stats=stats, T=T, rowa=rowa, moff_in=moff,
           moff_out=moff_out)

runs = {
    "proj_ln0": (lambda: ops.linear_bf16(x, w_o, mode=EPI_RESID_LN, out=h, bias=b1, resid=h, ln_produce=lp0), 2.0 * M * D * D),
    "proj_ln": (lambda: ops.linear_bf16(x, w_o, mode=EPI_RESID_LN, out=h, bias=b1, resid=h, ln_produce=lp1), 2.0 * M * D * D),
    "proj_ln2": (lambda: ops.linear_bf16(x, w_o, mode=EPI_RESID_LN, out=h, bias=b1, resid=h, ln_produce=lp2), 2.0 * M * D * D),
    "ffn2_ln": (lambda: ops.linear_bf16(mid, w2, mode=EPI_RESID_LN, out=h, bias=bd, resid=h, ln_produce=lp1), 2.0 * M * D * F),
    "qkv_ln": (lambda: ops.linear_bf16(x, w_qkv, mode=EPI_BF16_LN, out=qk, vt=vt, vt_col0=2 * D, T=T, H=H, hd=hd, Tp=Tp,
                                       ln_consume=dict(stats=stats, g=gb3[:, :3 * D], b=gb3[:, 3 * D:], rowa=rowa)), 2.0 * M * 3 * D * D),
    "ffn1_ln": (lambda: ops.linear_bf16(x, w1, mode=EPI_GELU_BF16_LN, out=mid, T=T,
                                        ln_consume=dict(stats=stats, g=gbf[:, :F], b=gbf[:, F:], rowa=rowa)), 2.0 * M * F * D),
    "qkv": (lambda: ops.linear_bf16(x, w_qkv, mode=EPI_BF16, out=qk, bias=b3, vt=vt, vt_col0=2 * D, T=T, H=H, hd=hd, Tp=Tp), 2.0 * M * 3 * D * D),
    "proj": (lambda: ops.linear_bf16(x, w_o, mode=EPI_RESID, out=h, bias=b1, resid=h), 2.0 * M * D * D),
    "ffn1": (lambda: ops.linear_bf16(x, w1, mode=EPI_GELU_BF16, out=mid, bias=bf), 2.0 * M * F * D),
    "ffn1_plain": (lambda: ops.linear_bf16(x, w1, mode=EPI_BF16, out=mid, bias=bf), 2.0 * M * F * D),
    "ffn2": (lambda: ops.linear_bf16(mid, w2, mode=EPI_RESID, out=h, bias=bd, resid=h), 2.0 * M * D * F),
    "attn": (lambda: ops.attention(qk[:, :D], qk[:, D:], vt, att, S, T, H, hd, 0), 4.0 * S * H * T * (T + 1) * hd),
    "ln": (lambda: ops.ln_modulate(h, T, mod[:, :2 * D], out0=xn, mod1=mod[:, 2 * D:], out1=xf), 0.0),
    "ln1": (lambda: ops.ln_modulate(h, T, mod[:, :2 * D], out0=xn), -1.0),
}
names = list(runs) if args.what == "all" else [args.what]
runs["qkv"][0]()  # make qk/vt meaningful for attention
for n in names:
    fn, fl = runs[n]
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    extra = (f"{fl / ms / 1e9:8.1f} TFLOP/s" if fl > 0 else
             f"{(M * D * 4 + (2 if fl == 0 else 1) * M * D * 2) / ms / 1e6:8.1f} GB/s")
    print(f"{n:5s} {ms * 1e3:9.1f} us  {extra}")
