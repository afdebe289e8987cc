"""shared fixtures for the parity tests"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import mixer_oracle, synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden():
    arrays = np.load(os.path.join(GOLDEN, "reference_small.npz"))
    with open(os.path.join(GOLDEN, "reference_small.json")) as f:
        meta = json.load(f)
    return arrays, meta


def default_schedule():
    from mixermdm_b200 import schedule
    return schedule.SpacedSchedule(schedule.space_timesteps(1000, "ddim50"),
                                   schedule.get_named_beta_schedule("cosine", 1000))


def build_mixer(dims, mixing_mode=4, align=True, store_influence=False, seed=0, mode="eval_intermediate",
                force_influence_val=None):
    """our drop-in modules with the synthetic weights of oracle/synth.py; returns (mixer, state_dict)"""
    from mixermdm_b200 import modules
    kw = dict(dropout=0.1, **dims["den"])
    d1 = modules.in2INDenoiser(262, mode="individual", **kw)
    d2 = modules.in2INDenoiser(262, mode="interaction", **kw)
    mix = modules.Mixer(d1, d2, nfeats=262, text_dim=768, mixing_mode=mixing_mode, store_influence=store_influence,
                        mode=mode, align=align, normalizers=synth.normalizer_stats(),
                        force_influence_val=force_influence_val, **dims["mix"]).eval()
    sd = synth.synthetic_state_dict(mix.state_dict(), seed=seed)
    mix.load_state_dict(sd, strict=True)
    return mix, sd


def build_oracle(sd, dims, mixing_mode=4, align=True, center=True, force=None):
    return mixer_oracle.Oracle(sd, dims, synth.normalizer_stats(), mixing_mode=mixing_mode, align=align,
                               center=center, sched=default_schedule(), force=force)


def _t(x):
    if isinstance(x, torch.Tensor):
        return x.detach().float().cpu()
    return torch.from_numpy(np.asarray(x)).float()


def relerr(a, b):
    a, b = _t(a), _t(b)
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()
