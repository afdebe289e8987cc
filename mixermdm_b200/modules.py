"""Drop-in mirror of the reference's Python model API for the denoising loop.

Module names, constructor arguments, `forward` signatures and — above all — the state-dict layout follow
pabloruizponce/MixerMDM (`src/models/in2in.py:358-462`, `src/models/mixermdm.py:605-811`,
`src/models/utils/{layers,blocks,influence,utils,cfg_sampler}.py`), so reference checkpoints load with
`load_state_dict(strict=True)`. The leaf modules here are *parameter containers only*: no arithmetic is done
in PyTorch. `forward` of the path-level modules (in2INDenoiser, Mixer, ClassifierFreeSampleModelX2) hands
device pointers to libmixermdm_b200.so (hand-written sm_100a kernels) and raises if it is unavailable.
"""
import math

import numpy as np
import torch
from torch import nn

from . import engine as _engine


def zero_module(module):
    """reference: models/utils/utils.py:91-97 (zero-initialised linears)"""
    for p in module.parameters():
        p.detach().zero_()
    return module


class PositionalEncoding(nn.Module):
    """sinusoidal table, persistent buffer `pe` (max_len, d_model) — reference models/utils/utils.py:24-39"""

    def __init__(self, d_model, dropout=0.0, max_len=5000):
        super().__init__()
        pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        freq = torch.exp(torch.arange(0, d_model, 2).float() * (-np.log(10000.0) / d_model))
        table = torch.zeros(max_len, d_model)
        table[:, 0::2] = torch.sin(pos * freq)
        table[:, 1::2] = torch.cos(pos * freq)
        self.register_buffer("pe", table)
        self.dropout = nn.Dropout(p=dropout)


class TimestepEmbedder(nn.Module):
    """time_embed = Linear, SiLU, Linear over pe[t] — reference models/utils/utils.py:41-55"""

    def __init__(self, latent_dim, sequence_pos_encoder):
        super().__init__()
        self.latent_dim = latent_dim
        self.sequence_pos_encoder = sequence_pos_encoder
        self.time_embed = nn.Sequential(nn.Linear(latent_dim, latent_dim), nn.SiLU(), nn.Linear(latent_dim, latent_dim))


class AdaLN(nn.Module):
    """emb_layers = (SiLU, zero Linear D->2D); LayerNorm without affine — reference models/utils/layers.py:3-25"""

    def __init__(self, latent_dim, embed_dim=None):
        super().__init__()
        embed_dim = latent_dim if embed_dim is None else embed_dim
        self.emb_layers = nn.Sequential(nn.SiLU(), zero_module(nn.Linear(embed_dim, 2 * latent_dim, bias=True)))
        self.norm = nn.LayerNorm(latent_dim, elementwise_affine=False, eps=1e-6)


class VanillaSelfAttention(nn.Module):
    """reference models/utils/layers.py:28-45"""

    def __init__(self, latent_dim, num_head, dropout, embed_dim=None):
        super().__init__()
        self.num_head = num_head
        self.norm = AdaLN(latent_dim, embed_dim)
        self.attention = nn.MultiheadAttention(latent_dim, num_head, dropout=dropout, batch_first=True,
                                               add_zero_attn=True)


class VanillaCrossAttention(nn.Module):
    """reference models/utils/layers.py:68-88"""

    def __init__(self, latent_dim, xf_latent_dim, num_head, dropout, embed_dim=None):
        super().__init__()
        self.num_head = num_head
        self.norm = AdaLN(latent_dim, embed_dim)
        self.xf_norm = AdaLN(xf_latent_dim, embed_dim)
        self.attention = nn.MultiheadAttention(latent_dim, num_head, kdim=xf_latent_dim, vdim=xf_latent_dim,
                                               dropout=dropout, batch_first=True, add_zero_attn=True)


class FFN(nn.Module):
    """reference models/utils/layers.py:91-106"""

    def __init__(self, latent_dim, ffn_dim, dropout, embed_dim=None):
        super().__init__()
        self.norm = AdaLN(latent_dim, embed_dim)
        self.linear1 = nn.Linear(latent_dim, ffn_dim, bias=True)
        self.linear2 = zero_module(nn.Linear(ffn_dim, latent_dim, bias=True))
        self.activation = nn.GELU()
        self.dropout = nn.Dropout(dropout)


class FinalLayer(nn.Module):
    """reference models/utils/layers.py:109-116"""

    def __init__(self, latent_dim, out_dim):
        super().__init__()
        self.linear = zero_module(nn.Linear(latent_dim, out_dim, bias=True))


class TransformerBlockDoubleCond(nn.Module):
    """SA -> [CA] -> FFN container; `ca_block` exists in every mode (reference models/utils/blocks.py:30-63)"""

    def __init__(self, mode, latent_dim=512, num_heads=8, ff_size=1024, dropout=0., cond_abl=False, **kargs):
        super().__init__()
        self.latent_dim, self.num_heads, self.dropout, self.cond_abl, self.mode = latent_dim, num_heads, dropout, cond_abl, mode
        self.sa_block = VanillaSelfAttention(latent_dim, num_heads, dropout)
        self.ca_block = VanillaCrossAttention(latent_dim, latent_dim, num_heads, dropout, latent_dim)
        self.ffn = FFN(latent_dim, ff_size, dropout, latent_dim)


class InfluenceBlockCross(nn.Module):
    """reference models/utils/influence.py:6-48"""

    def __init__(self, latent_dim=512, num_heads=8, ff_size=1024, dropout=0., **kargs):
        super().__init__()
        self.latent_dim, self.num_heads, self.dropout = latent_dim, num_heads, dropout
        self.sa_block = VanillaSelfAttention(latent_dim, num_heads, dropout)
        self.ca_block = VanillaCrossAttention(latent_dim, latent_dim, num_heads, dropout, latent_dim)
        self.ffn = FFN(latent_dim, ff_size, dropout, latent_dim)


class in2INDenoiser(nn.Module):
    """Two-stream (interaction) / single-stream (individual) motion transformer.
    reference: models/in2in.py:358-462 — same constructor, same parameters, same forward(x, timesteps, mask, cond)."""

    def __init__(self, input_feats, mode, latent_dim=512, num_frames=240, ff_size=1024, num_layers=8, num_heads=8,
                 dropout=0.1, activation="gelu", **kargs):
        super().__init__()
        if mode not in ("individual", "interaction"):
            raise ValueError("Mode not recognized")   # dual_* modes are outside the MixerMDM path
        self.num_frames, self.latent_dim, self.ff_size = num_frames, latent_dim, ff_size
        self.num_layers, self.num_heads, self.dropout = num_layers, num_heads, dropout
        self.activation, self.input_feats, self.time_embed_dim, self.mode = activation, input_feats, latent_dim, mode
        self.text_emb_dim = 768
        self.text_dim = 768
        self.sequence_pos_encoder = PositionalEncoding(latent_dim, dropout=0)
        self.embed_timestep = TimestepEmbedder(latent_dim, self.sequence_pos_encoder)
        self.motion_embed = nn.Linear(input_feats, latent_dim)
        self.text_embed = nn.Linear(self.text_emb_dim, latent_dim)
        self.blocks = nn.ModuleList([
            TransformerBlockDoubleCond(num_heads=num_heads, latent_dim=latent_dim, dropout=dropout, ff_size=ff_size,
                                       mode=mode) for _ in range(num_layers)])
        self.out = zero_module(FinalLayer(latent_dim, input_feats))
        self._standalone = None   # lazily built single-network engine for direct forward() calls

    def forward(self, x, timesteps, mask=None, cond=None):
        """x: (B,T,262) [individual] or (B,T,524) [interaction]; timesteps: (B,) original-scale indices into the
        PE table; cond: (B,768) / (B,2304). Returns the normalised-space x0 prediction, same shape as x."""
        if self._standalone is None or not self._standalone.matches(self, x):
            self._standalone = None   # parameters changed (load_state_dict, in-place update, .to()): repack
            self._standalone = _engine.StandaloneDenoiser(self, x)
        return self._standalone.forward(x, timesteps, cond, mask)   # mask: (B, T, P) float, mask[..., 0] > 0.5 = real frame

    def __getstate__(self):   # the engine holds ctypes pointers: copies / pickles rebuild it on first use
        d = dict(self.__dict__)
        d["_standalone"] = None
        return d


class Influence(nn.Module):
    """Mixer weight network (reference models/utils/influence.py:50-126). Runs inside Mixer.forward."""

    def __init__(self, input_shape, n_blocks, n_heads, ff_size, mode):
        super().__init__()
        self.input_shape, self.n_blocks, self.n_heads, self.ff_size, self.mode = input_shape, n_blocks, n_heads, ff_size, mode
        self.blocks = nn.ModuleList([InfluenceBlockCross(latent_dim=input_shape, num_heads=n_heads, ff_size=ff_size)
                                     for _ in range(n_blocks)])
        if mode in (1, 2):
            self.out = nn.Linear(input_shape, 1)
        elif mode in (3, 4):
            self.out = nn.Linear(input_shape, 23)
        else:
            raise ValueError("Mode not recognized")


def history_view(w, mixing_mode):
    """The kernels write the weights expanded to the 262 features. The reference stores them that way for the
    joint-wise modes 3/4, but as (Be, T, 1) for the global / temporal modes 1/2 (models/mixermdm.py:739-796)."""
    return w[..., :1] if mixing_mode in (1, 2) else w


class Mixer(nn.Module):
    """Per-step body of MixerMDM (reference models/mixermdm.py:605-811): both frozen denoisers, alignment, the
    Influence network and the weighted mix. forward() keeps the reference signature."""

    def __init__(self, denoiser1, denoiser2, nfeats, latent_dim, ff_size, text_dim, n_blocks, n_heads, mixing_mode,
                 store_influence=False, force_influence_val=None, mode="train", align=True, normalizers=None):
        super().__init__()
        self.nfeats, self.latent_dim, self.ff_size, self.text_dim = nfeats, latent_dim, ff_size, text_dim
        self.n_blocks, self.n_heads, self.mixing_mode, self.mode, self.align = n_blocks, n_heads, mixing_mode, mode, align
        self.denoiser1 = denoiser1
        self.denoiser2 = denoiser2
        self.force_influence_val = force_influence_val
        self.influence = Influence(input_shape=latent_dim, n_blocks=n_blocks, n_heads=n_heads, ff_size=ff_size,
                                   mode=mixing_mode)
        self.store_influence = store_influence
        if store_influence:
            self.history_influence_i1 = []
            self.history_influence_i2 = []
        if mode == "eval":
            self.history_out1, self.history_out2, self.history_out_influenced = [], [], []
        self.sequence_pos_encoder = PositionalEncoding(latent_dim, dropout=0)
        self.embed_timestep = TimestepEmbedder(latent_dim, self.sequence_pos_encoder)
        self.motion_embed = nn.Linear(nfeats, latent_dim)
        self.text_embed = nn.Linear(text_dim, latent_dim)
        # (mean, std) pairs: normalizer_model1 = HumanML3D stats (individual), normalizer_model2 = InterHuman stats
        self.normalizer_model1, self.normalizer_model2 = _engine.load_normalizers(normalizers)
        self._engine = None

    def engine(self, batch_rows, frames, device, sched=None, cfg_scale=None, center=None):
        """the CUDA engine bound to a workspace for `batch_rows` network rows (CFG-doubled batch) of `frames`;
        rebuilt when a sampler asks for another schedule / guidance scale / centring flag"""
        eng = self._engine
        device = _engine.canonical_device(device)
        stale = (eng is None or eng.device != device or
                 eng.fingerprint != _engine.weights_fingerprint(self))   # live parameters changed since packing
        if eng is not None and not stale:
            stale = ((sched is not None and not eng.sched.same_as(sched)) or
                     (cfg_scale is not None and float(cfg_scale) != eng.cfg_scale) or
                     (center is not None and bool(center) != eng.center))
        if stale:
            kw = {}
            if sched is not None:
                kw["sched"] = sched
            elif eng is not None:
                kw["sched"] = eng.sched
            kw["cfg_scale"] = float(cfg_scale) if cfg_scale is not None else (eng.cfg_scale if eng else 3.5)
            kw["center"] = bool(center) if center is not None else (eng.center if eng else True)
            self._engine = _engine.MixerEngine(self, device, **kw)
        self._engine.bind((batch_rows + 1) // 2, frames)
        return self._engine

    def invalidate_engine(self):
        """drop the packed weight copies now. Not needed for correctness: engine() compares a fingerprint of the live
        parameters (identity + version counters) with the one taken at pack time and repacks on any difference."""
        self._engine = None

    def __getstate__(self):   # the engine holds ctypes pointers: copies / pickles rebuild it on first use
        d = dict(self.__dict__)
        d["_engine"] = None
        return d

    def forward(self, x1, timesteps, cond=None, mask=None, x2=None):
        """x1/x2: (Be,T,524) state in individual / interaction space, timesteps (Be,), cond (Be,6144).
        eval modes return out_influenced (Be,T,524); train mode returns (out_influenced, out1, out2)."""
        _check_mask(self, mask)
        eng = self.engine(x1.shape[0], x1.shape[1], x1.device)
        out, out1, out2, w1, w2 = eng.mixer_forward(x1, x2, timesteps, cond,
                                                    want_aux=(self.mode != "eval_intermediate"),
                                                    want_influence=self.store_influence, mask=mask)
        if self.store_influence:
            self.history_influence_i1.append(history_view(w1, self.mixing_mode))
            self.history_influence_i2.append(history_view(w2, self.mixing_mode))
        if self.mode == "train":
            return out, out1, out2
        if self.mode == "eval":
            self.history_out1.append(out1)
            self.history_out2.append(out2)
            self.history_out_influenced.append(out)
        return out


def _check_mask(mixer, mask):
    """Frame masks reach the three networks' attention layers (in2in.py:433-435, influence.py:107-116). With
    Mixer(align=True) the reference hands the (B, T, 2) mask to align_trajectories, which builds a (B, 2) length tensor
    and indexes one frame per batch row with it (utils/alignment.py:86-89): a shape error for B != 2. There is nothing
    to be identical to, so that combination raises here too."""
    if mask is not None and mixer.align:
        raise NotImplementedError("mask with Mixer(align=True): the reference's masked trajectory alignment does not run "
                                  "(utils/alignment.py:86-89); use align=False or mask=None")


class ClassifierFreeSampleModelX2(nn.Module):
    """CFG wrapper for the two-stream sampler (reference models/utils/cfg_sampler.py:31-56):
    s * model(cond) + (1 - s) * model(cond = 0) on a doubled batch."""

    def __init__(self, model, cfg_scale):
        super().__init__()
        self.model = model
        self.s = cfg_scale

    def forward(self, x, x2, timesteps, cond=None, mask=None):
        _check_mask(self.model, mask)
        B = x.shape[0]
        eng = self.model.engine(2 * B, x.shape[1], x.device)
        return eng.cfg_forward(x, x2, timesteps, cond, self.s, store_influence=self.model.store_influence,
                               owner=self.model, mask=mask)
