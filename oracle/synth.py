"""TEST INFRASTRUCTURE ONLY — version-independent synthetic weights and inputs (numpy PCG64 streams), so that
the reference (in the build container), the oracle and the CUDA path (on the GPU box) see identical numbers."""
import zlib

import numpy as np
import torch

SMALL_DIMS = dict(den=dict(latent_dim=256, ff_size=512, num_layers=2, num_heads=2),
                  mix=dict(latent_dim=512, ff_size=1024, n_blocks=2, n_heads=8))
FULL_DIMS = dict(den=dict(latent_dim=1024, ff_size=2048, num_layers=8, num_heads=8),
                 mix=dict(latent_dim=512, ff_size=1024, n_blocks=4, n_heads=8))

# parameters the reference zero-initialises (zero_module: models/utils/layers.py:11,96,112; in2in.py:398);
# left at zero every network would output exactly 0 and parity would be vacuous (SURVEY.md §8a note 1)
_ZERO_INIT = ("emb_layers.1.", "ffn.linear2.", "out.linear.")


def _rng(seed, name):
    return np.random.default_rng([seed, zlib.crc32(name.encode())])


def synthetic_state_dict(template, seed=0):
    """template: a state dict (for names/shapes). Buffers named '*.pe' are kept; everything else is redrawn."""
    out = {}
    for name, t in template.items():
        if name.endswith(".pe") or name == "pe":
            out[name] = t.clone()
            continue
        g = _rng(seed, name)
        shape = tuple(t.shape)
        if any(z in name for z in _ZERO_INIT):
            a = g.normal(0.0, 0.02, size=shape)
        elif len(shape) >= 2:
            bound = 1.0 / np.sqrt(shape[-1])
            a = g.uniform(-bound, bound, size=shape)
        else:
            a = g.normal(0.0, 0.02, size=shape)
        out[name] = torch.from_numpy(a.astype(np.float32))
    return out


def normalizer_stats(seed=0):
    """((mean1,std1) HumanML3D / individual model, (mean2,std2) InterHuman / interaction model) — same numbers
    as oracle/ref_harness.normalizer_stats writes to disk for the reference"""
    rng = np.random.default_rng(seed)
    pairs = []
    for _ in range(2):
        mean = (0.1 * rng.normal(size=262)).astype(np.float32)
        std = rng.uniform(0.5, 1.5, size=262).astype(np.float32)
        pairs.append((mean, std))
    glob, hml = pairs
    return hml, glob


def inputs(B, T, seed=0):
    """cond (B,6144) unit Gaussian; per-sample seeded noise (B,T,524) — independent of the batch partition"""
    cond = np.stack([_rng(seed, f"cond{b}").normal(size=6144) for b in range(B)]).astype(np.float32)
    noise = np.stack([_rng(seed, f"noise{b}").normal(size=(T, 524)) for b in range(B)]).astype(np.float32)
    return torch.from_numpy(cond), torch.from_numpy(noise)


def sample_noise(global_index, T, seed=0):
    return torch.from_numpy(_rng(seed, f"noise{global_index}").normal(size=(T, 524)).astype(np.float32))


def sample_cond(global_index, seed=0):
    return torch.from_numpy(_rng(seed, f"cond{global_index}").normal(size=6144).astype(np.float32))
