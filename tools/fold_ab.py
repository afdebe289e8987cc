"""A/B of the LN fold against the stand-alone LayerNorm kernel on the teacher-forced reference chain: per-step errors of
the denoisers and of Mixer.forward against the fp32 oracle. Run twice: MMDM_LN_FOLD=1 / 0."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from helpers import GOLDEN, build_mixer, build_oracle, default_schedule, relerr  # noqa: E402
from oracle import mixer_oracle, synth  # noqa: E402

DEV = "cuda:0"
ch = np.load(os.path.join(GOLDEN, "reference_chain.npz"))
meta = json.load(open(os.path.join(GOLDEN, "reference_chain.json")))
B, T = meta["B"], meta["T"]
H = meta["dims"]["den"]["num_heads"]
cond, _ = synth.inputs(B, T, seed=meta["inputs_seed"])
mix, sd = build_mixer(meta["dims"], mixing_mode=meta["mixing_mode"], seed=meta["weights_seed"])
mix = mix.to(DEV)
sdf = {k: v.float() for k, v in sd.items()}
sched = default_schedule()
orc = build_oracle(sd, meta["dims"], mixing_mode=meta["mixing_mode"])
eng = mix.engine(2 * B, T, torch.device(DEV), sched=sched, cfg_scale=3.5, center=True)
eng.set_condition(cond.to(DEV))
x1s = torch.from_numpy(ch["x1"].astype(np.float32))
x2s = torch.from_numpy(ch["x2"].astype(np.float32))
cc = torch.cat([cond, torch.zeros_like(cond)])
rows = []
for i in range(0, 50, int(os.environ.get("STRIDE", "1"))):
    idx = 49 - i
    t = torch.full((2 * B,), int(sched.timestep_map[idx]), dtype=torch.long)
    xx1, xx2 = torch.cat([x1s[i], x1s[i]]), torch.cat([x2s[i], x2s[i]])
    d1 = mix.denoiser1(xx1[..., :262].contiguous().to(DEV), t.to(DEV), cond=cc[:, 2304:3072].contiguous().to(DEV))
    d2 = mix.denoiser2(xx2.to(DEV), t.to(DEV), cond=cc[:, :2304].contiguous().to(DEV))
    net, *_ = eng.mixer_forward(xx1.to(DEV), xx2.to(DEV), t.to(DEV), cc.to(DEV))
    with torch.no_grad():
        r1 = mixer_oracle.denoiser(sdf, "denoiser1.", "individual", xx1[..., :262], t, cc[:, 2304:3072], H)
        r2 = mixer_oracle.denoiser(sdf, "denoiser2.", "interaction", xx2, t, cc[:, :2304], H)
        ref_net = orc.mixer(xx1, xx2, t, cc)
    eng.set_condition(cond.to(DEV))
    eng.set_state(x1s[i].to(DEV), x2s[i].to(DEV), idx)
    final = torch.zeros(B, T, 524, device=DEV)
    eng.step(final)
    s1, s2 = eng.get_state()
    torch.cuda.synchronize()
    e_state = max(relerr(s1, x1s[i + 1]), relerr(s2, x2s[i + 1])) if idx > 0 else relerr(final, ch["final"])
    rows.append((idx, relerr(d1, r1), relerr(d2, r2), relerr(net, ref_net), e_state))
a = np.array(rows)
print("fold =", os.environ.get("MMDM_LN_FOLD", "1"))
print("den1 mean %.2e max %.2e | den2 mean %.2e max %.2e | mixer mean %.2e max %.2e (idx %d)" % (
    a[:, 1].mean(), a[:, 1].max(), a[:, 2].mean(), a[:, 2].max(), a[:, 3].mean(), a[:, 3].max(), a[a[:, 3].argmax(), 0]))
print("mixer per step:", " ".join("%.1f" % (1e3 * v) for v in a[:, 3]))
print("state mean %.2e max %.2e (idx %d); per step (x1e-3): %s" % (a[:, 4].mean(), a[:, 4].max(), a[a[:, 4].argmax(), 0],
                                                                 " ".join("%.1f" % (1e3 * v) for v in a[:, 4])))
