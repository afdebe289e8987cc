"""Sweep of the two LayerNorm code paths (folded into the GEMMs / stand-alone kernel) over network widths, head sizes,
frame counts and batch sizes: both denoisers, outputs compared with each other (two bf16 roundings of one function).
   python tools/fold_sweep.py"""
import itertools
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from helpers import build_mixer, relerr  # noqa: E402

DIMS = [dict(den=dict(latent_dim=256, ff_size=512, num_layers=2, num_heads=2), mix=dict(latent_dim=512, ff_size=1024, n_blocks=2, n_heads=8)),
        dict(den=dict(latent_dim=512, ff_size=1024, num_layers=3, num_heads=4), mix=dict(latent_dim=512, ff_size=1024, n_blocks=2, n_heads=8)),
        dict(den=dict(latent_dim=512, ff_size=2048, num_layers=2, num_heads=8), mix=dict(latent_dim=512, ff_size=1024, n_blocks=2, n_heads=8)),
        dict(den=dict(latent_dim=1024, ff_size=2048, num_layers=2, num_heads=8), mix=dict(latent_dim=512, ff_size=1024, n_blocks=2, n_heads=8))]
worst = 0.0
for di, dims in enumerate(DIMS):
    mix, _ = build_mixer(dims, seed=11 + di)
    mix = mix.to("cuda")
    for T, B in itertools.product([256, 263, 300, 511], [1, 2, 3, 5]):
        g = torch.Generator().manual_seed(1000 * di + 10 * T + B)
        x = torch.randn(B, T, 524, generator=g).cuda()
        cond = torch.randn(B, 2304, generator=g).cuda()
        t = torch.randint(0, 1000, (B,), generator=g).cuda()
        res = {}
        for flag in ("0", "1"):
            os.environ["MMDM_LN_FOLD"] = flag
            mix.denoiser1._standalone = None
            mix.denoiser2._standalone = None
            res[flag] = (mix.denoiser2(x, t, cond=cond).clone(),
                         mix.denoiser1(x[..., :262].contiguous(), t, cond=cond[:, :768].contiguous()).clone())
        torch.cuda.synchronize()
        e2, e1 = relerr(res["1"][0], res["0"][0]), relerr(res["1"][1], res["0"][1])
        ok = all(torch.isfinite(r).all() for pair in res.values() for r in pair)
        worst = max(worst, e1, e2)
        flagstr = "" if (ok and max(e1, e2) < 1.5e-2) else "   <-- CHECK"
        print(f"D={dims['den']['latent_dim']} H={dims['den']['num_heads']} F={dims['den']['ff_size']} T={T} B={B}: "
              f"interaction {e2:.2e} individual {e1:.2e}{flagstr}")
print(f"worst folded-vs-kernel difference {worst:.2e}")
