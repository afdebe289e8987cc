"""TEST INFRASTRUCTURE ONLY — drives the UNMODIFIED reference (pabloruizponce/MixerMDM, /root/reference) on CPU.

Used (a) to pin oracle/mixer_oracle.py against the real reference and (b) to generate the golden vectors under
tests/golden/ (oracle/make_goldens.py). It imports the reference, it does not copy it; /root/reference only
exists in the build container, so nothing here is reachable from `-m gpu` tests, smoke() or bench.py.

Import blockers handled without touching the reference:
  * `clip`, `aitviewer.renderables.lines`, `yacs.config` are stubbed in sys.modules;
  * the normalisers np.load("./data/...") relative to the CWD -> a scratch CWD with synthetic stats.
"""
import contextlib
import os
import sys
import tempfile
import types

import numpy as np
import torch

REF_SRC = "/root/reference/src"


def available():
    return os.path.isdir(REF_SRC)


def _stub(name, **attrs):
    if name in sys.modules:
        return
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m


def normalizer_stats(seed=0):
    """synthetic (mean, std) pairs: [0] InterHuman/global (interaction model), [1] HumanML3D (individual model)"""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(2):
        mean = (0.1 * rng.normal(size=262)).astype(np.float32)
        std = rng.uniform(0.5, 1.5, size=262).astype(np.float32)
        out.append((mean, std))
    return out


@contextlib.contextmanager
def reference_env(stats_seed=0):
    """chdir into a scratch dir holding the four normaliser files and make the reference importable."""
    if not available():
        raise RuntimeError("reference sources not present at /root/reference")
    _stub("clip")
    _stub("aitviewer")
    _stub("aitviewer.renderables")
    _stub("aitviewer.renderables.lines", Lines=object)
    _stub("yacs")
    _stub("yacs.config", CfgNode=type("CfgNode", (dict,), {}))
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    (gm, gs), (hm, hs) = normalizer_stats(stats_seed)
    old = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "data", "HumanML3D"))
        np.save(os.path.join(d, "data", "global_mean.npy"), gm)
        np.save(os.path.join(d, "data", "global_std.npy"), gs)
        np.save(os.path.join(d, "data", "HumanML3D", "mean_ih_new.npy"), hm)
        np.save(os.path.join(d, "data", "HumanML3D", "std_ih_new.npy"), hs)
        os.chdir(d)
        try:
            yield
        finally:
            os.chdir(old)


def build_reference(dims, mixing_mode=4, align=True, store_influence=True, stats_seed=0, mode="eval_intermediate"):
    """dims: dict(den=dict(latent_dim, ff_size, num_layers, num_heads), mix=dict(latent_dim, ff_size, n_blocks, n_heads)).
    Returns (mixer, diffusion, cfg_model) built from the reference's own classes."""
    with reference_env(stats_seed):
        from models.in2in import in2INDenoiser
        from models.mixermdm import Mixer
        from models.utils.cfg_sampler import ClassifierFreeSampleModelX2
        from models.utils.gaussian_diffusion import (LossType, MixerDiffusion, ModelMeanType, ModelVarType,
                                                     get_named_beta_schedule, space_timesteps)
        kw = dict(dropout=0.1, **dims["den"])
        d1 = in2INDenoiser(262, mode="individual", **kw)
        d2 = in2INDenoiser(262, mode="interaction", **kw)
        mix = Mixer(d1, d2, nfeats=262, text_dim=768, mixing_mode=mixing_mode, store_influence=store_influence,
                    mode=mode, align=align, **dims["mix"]).eval()
        if store_influence:
            mix.history_influence_i1, mix.history_influence_i2 = [], []
        diff = MixerDiffusion(use_timesteps=space_timesteps(1000, dims.get("sampling", "ddim50")),
                              betas=get_named_beta_schedule("cosine", 1000), model_mean_type=ModelMeanType.START_X,
                              model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE,
                              rescale_timesteps=False)
        cfgm = ClassifierFreeSampleModelX2(mix, 3.5)
    return mix, diff, cfgm


def reference_modules():
    """the reference's own motion-space functions, for unit pinning of the oracle"""
    with reference_env():
        from utils import alignment, quaternion, rotation_conversions
    return alignment, quaternion, rotation_conversions
