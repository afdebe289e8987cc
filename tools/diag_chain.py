import json, os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from helpers import GOLDEN, build_mixer, build_oracle, default_schedule, relerr
from oracle import mixer_oracle, synth
DEV = "cuda"
ch = np.load(os.path.join(GOLDEN, "reference_chain.npz"))
meta = json.load(open(os.path.join(GOLDEN, "reference_chain.json")))
B, T = meta["B"], meta["T"]
cond, _ = synth.inputs(B, T, seed=meta["inputs_seed"])
mix, sd = build_mixer(meta["dims"], mixing_mode=4, seed=0)
mix = mix.to(DEV)
sched = default_schedule()
orc = build_oracle(sd, meta["dims"], mixing_mode=4)
sdf = {k: v.float() for k, v in sd.items()}
H = meta["dims"]["den"]["num_heads"]
x1s = torch.from_numpy(ch["x1"].astype(np.float32)); x2s = torch.from_numpy(ch["x2"].astype(np.float32))
cc = torch.cat([cond, torch.zeros_like(cond)])
for i in range(0, 50, 3):
    idx = 49 - i
    t = torch.full((2 * B,), int(sched.timestep_map[idx]), dtype=torch.long)
    xx1, xx2 = torch.cat([x1s[i], x1s[i]]), torch.cat([x2s[i], x2s[i]])
    net = mix(xx1.to(DEV), t.to(DEV), cond=cc.to(DEV), mask=None, x2=xx2.to(DEV))
    d1 = mix.denoiser1(xx1[..., :262].contiguous().to(DEV), t.to(DEV), cond=cc[:, 2304:3072].contiguous().to(DEV))
    d2 = mix.denoiser2(xx2.to(DEV), t.to(DEV), cond=cc[:, :2304].contiguous().to(DEV))
    with torch.no_grad():
        r = orc.mixer(xx1, xx2, t, cc)
        r1 = mixer_oracle.denoiser(sdf, "denoiser1.", "individual", xx1[..., :262], t, cc[:, 2304:3072], H)
        r2 = mixer_oracle.denoiser(sdf, "denoiser2.", "interaction", xx2, t, cc[:, :2304], H)
    print(f"idx {idx:2d} t {int(t[0]):3d}: mixer {relerr(net, r):.2e}  den_ind {relerr(d1, r1):.2e}  den_int {relerr(d2, r2):.2e}  |x| {float(xx2.abs().mean()):.2f}")
