"""Top-level `MixerMDM` with the reference's constructor, attributes, state-dict layout and forward contracts
(reference models/mixermdm.py:18-602). The denoising loop runs in the CUDA engine; everything that is not on
that path (CLIP text tower, the sub-models' text encoders, discriminators) exists here as parameter containers so
that reference checkpoints load with `load_state_dict(strict=True)` after the loaders' key surgery
(scripts/infer/mixermdm.py:169-177), and so that callers find the attributes they read.

Text conditioning: the reference tokenises with OpenAI `clip` and runs the frozen ViT-L/14@336 text tower plus three
2-layer encoders once per batch (models/mixermdm.py:283-356). Here the tower and the encoders run in the CUDA library
(csrc/text.cu through text.TextEngine); the modules below only hold their parameters. Tokenisation is the one host
step: `clip.tokenize` when the package is installed, or any callable set as `model.tokenizer`; a precomputed
(B, 6144) condition can still be passed as batch["cond"].
"""
import os
from collections import OrderedDict

import torch
from torch import nn

from . import text as text_engine
from .diffusion import LossType, MixerDiffusion, ModelMeanType, ModelVarType, get_named_beta_schedule, space_timesteps
from .modules import (ClassifierFreeSampleModelX2, FFN, Mixer, PositionalEncoding, TimestepEmbedder,
                      VanillaSelfAttention, in2INDenoiser)


class AttrDict(dict):
    """minimal stand-in for yacs.CfgNode (attribute access + `in`), used when yacs is not installed"""
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None   # hasattr(), copy.deepcopy() and pickle rely on AttributeError

    def __setattr__(self, k, v):
        self[k] = v


def get_config(config_file):
    """reference utils/configs.py:26-41 — YAML -> frozen node; string 'None' is literal-evaluated like yacs does"""
    import yaml
    with open(config_file) as f:
        raw = yaml.safe_load(f)

    def conv(x):
        if isinstance(x, dict):
            return AttrDict({k: conv(v) for k, v in x.items()})
        if isinstance(x, str) and x == "None":
            return None
        return x
    return conv(raw)


# ------------------------------------------------------------------------------------------------
# parameter containers off the hot path
# ------------------------------------------------------------------------------------------------
class QuickGELU(nn.Module):
    """parameter-free placeholder that keeps the `mlp.gelu` slot of CLIP's block layout"""


class _ClipResBlock(nn.Module):
    """OpenAI CLIP ResidualAttentionBlock layout: attn / ln_1 / mlp.{c_fc,gelu,c_proj} / ln_2 (parameters only; the
    arithmetic runs in csrc/text.cu)"""

    def __init__(self, d, heads):
        super().__init__()
        self.attn = nn.MultiheadAttention(d, heads)
        self.ln_1 = nn.LayerNorm(d)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d, 4 * d)), ("gelu", QuickGELU()), ("c_proj", nn.Linear(4 * d, d))]))
        self.ln_2 = nn.LayerNorm(d)


class _ClipTextTransformer(nn.Module):
    def __init__(self, width=768, layers=12, heads=12):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.Sequential(*[_ClipResBlock(width, heads) for _ in range(layers)])


class _ClipTextTower:
    """mixin building the frozen CLIP ViT-L/14 text-tower members with the reference's attribute names"""

    def _build_clip(self):
        self.token_embedding = nn.Embedding(49408, 768)
        self.clip_transformer = _ClipTextTransformer(768, 12, 12)
        self.positional_embedding = nn.Parameter(torch.empty(77, 768).normal_(std=0.01))
        self.ln_final = nn.LayerNorm(768)
        self.dtype = torch.float32
        self.tokenizer = None          # callable(list[str]) -> (B, 77) ints; default: clip.tokenize(..., truncate=True)
        for m in (self.token_embedding, self.clip_transformer, self.ln_final):
            for p in m.parameters():
                p.requires_grad = False

    def _tokens(self, raw_text):
        tok = self.tokenizer or text_engine.default_tokenizer
        return torch.as_tensor(tok(raw_text))

    def _encode_text(self, key, encoders, tokens):
        """run the CUDA text path: this object's CLIP tower, then `encoders` [(TransformerEncoder, LayerNorm), ...]"""
        cache = self.__dict__.setdefault("_text_engines", {})
        from .engine import canonical_device
        device = canonical_device(self.positional_embedding.device)
        eng = cache.get(key)
        if eng is None or eng.device != device or eng.stale():
            cache[key] = None
            eng = cache[key] = text_engine.TextEngine(self, encoders, device)
        return eng.encode(tokens)


def _text_encoder():
    layer = nn.TransformerEncoderLayer(d_model=768, nhead=8, dim_feedforward=2048, dropout=0.1, activation="gelu",
                                       batch_first=True)
    return nn.TransformerEncoder(layer, num_layers=2), nn.LayerNorm(768)


class in2INDiffusion(nn.Module):
    """container for `decoder.net_individual` / `decoder.net_interaction` (reference models/in2in.py:137-207)"""

    def __init__(self, cfg, mode, sampling_strategy="ddim50"):
        super().__init__()
        self.cfg, self.mode = cfg, mode
        kw = dict(latent_dim=cfg.LATENT_DIM, ff_size=cfg.FF_SIZE, num_layers=cfg.NUM_LAYERS, num_heads=cfg.NUM_HEADS,
                  dropout=cfg.DROPOUT, activation=cfg.ACTIVATION)
        if mode == "interaction":
            self.net_interaction = in2INDenoiser(cfg.INPUT_DIM, mode="interaction", **kw)
        elif mode == "individual":
            self.net_individual = in2INDenoiser(cfg.INPUT_DIM, mode="individual", **kw)
        else:
            raise ValueError("Mode not recognized")   # "dual" (DualMDM) is outside the MixerMDM path


class in2IN(nn.Module, _ClipTextTower):
    """pre-trained sub-model wrapper: denoiser + its own text encoder (reference models/in2in.py:11-135)"""

    def __init__(self, cfg, mode):
        super().__init__()
        self.cfg, self.latent_dim, self.mode = cfg, cfg.LATENT_DIM, mode
        self.decoder = in2INDiffusion(cfg, mode, sampling_strategy=cfg.get("STRATEGY", "ddim50"))
        if mode == "interaction":
            self.clipTransEncoder_interaction, self.clip_ln_interaction = _text_encoder()
        else:
            self.clipTransEncoder_individual, self.clip_ln_individual = _text_encoder()
        self._build_clip()

    def _text_encoders(self, mode):
        if mode == "individual":
            return [(self.clipTransEncoder_individual, self.clip_ln_individual)]
        if mode == "interaction":
            return [(self.clipTransEncoder_interaction, self.clip_ln_interaction)]
        raise ValueError("Mode not recognized")

    def text_process(self, batch, mode, text_name="text", out_name="cond"):
        """reference models/in2in.py:109-135: CLIP tower -> this sub-model's text encoder -> end-of-text row"""
        batch[out_name] = self._encode_text(mode, self._text_encoders(mode), self._tokens(batch[text_name]))[0]
        return batch


class TransformerBlockSimple(nn.Module):
    """reference models/utils/blocks.py:66-89"""

    def __init__(self, latent_dim=512, num_heads=8, ff_size=1024, dropout=0., cond_abl=False, **kargs):
        super().__init__()
        self.sa_block = VanillaSelfAttention(latent_dim, num_heads, dropout)
        self.ffn = FFN(latent_dim, ff_size, dropout, latent_dim)


class DiscriminatorTransfomer(nn.Module):
    """training-only adversary; kept because it is part of the checkpoint (reference models/utils/discriminators.py)"""

    def __init__(self, input_feats, latent_dim, num_frames, ff_size, num_layers, num_heads, dropout=0.1,
                 activation="gelu", **kargs):
        super().__init__()
        self.sequence_pos_encoder = PositionalEncoding(latent_dim, dropout=0)
        self.embed_timestep = TimestepEmbedder(latent_dim, self.sequence_pos_encoder)
        self.motion_embed = nn.Linear(input_feats, latent_dim)
        self.text_embed = nn.Linear(768, latent_dim)
        self.blocks = nn.ModuleList([TransformerBlockSimple(num_heads=num_heads, latent_dim=latent_dim, dropout=dropout,
                                                            ff_size=ff_size) for _ in range(num_layers)])
        self.out = nn.Linear(latent_dim, 1)

    def forward(self, *a, **k):
        raise NotImplementedError("the discriminators are only used by the reference's training step")


# ------------------------------------------------------------------------------------------------
class MixerMDM(nn.Module, _ClipTextTower):
    def __init__(self, cfg, num_frames=300, sampling_strategy="ddim50", store_influence=True, align=True,
                 normalizers=None, load_submodel_checkpoints=True):
        super().__init__()
        self.cfg = cfg
        self.cfg_model1 = get_config(cfg.MODEL1) if isinstance(cfg.MODEL1, str) else cfg.MODEL1
        self.cfg_model2 = get_config(cfg.MODEL2) if isinstance(cfg.MODEL2, str) else cfg.MODEL2
        self.align = align
        self.store_influence = store_influence
        if self.cfg_model1.NAME != "in2INind" or self.cfg_model2.NAME != "in2IN":
            raise NotImplementedError("mixermdm_b200 accelerates the default pairing MODEL1=in2INind, MODEL2=in2IN "
                                      "(MDM / InterGen back-ends stay on the reference)")
        self.model1 = in2IN(self.cfg_model1, mode="individual")
        self.model2 = in2IN(self.cfg_model2, mode="interaction")
        if load_submodel_checkpoints:
            for model, c in ((self.model1, self.cfg_model1), (self.model2, self.cfg_model2)):
                ck = c.get("CHECKPOINT")
                if ck:   # like the reference (models/mixermdm.py:43-59) a configured but missing file is an error
                    if not os.path.exists(ck):
                        raise FileNotFoundError(f"sub-model checkpoint {ck!r} (cfg.CHECKPOINT) not found; pass "
                                                "load_submodel_checkpoints=False to build without it")
                    model.load_state_dict(torch.load(ck, map_location="cpu"), strict=True)
        for m in (self.model1, self.model2):
            for p in m.parameters():
                p.requires_grad = False
            m.eval()
        self.denoiser1 = self.model1.decoder.net_individual
        self.denoiser2 = self.model2.decoder.net_interaction
        self.cond_pre_func1 = self.model1.text_process
        self.cond_pre_func2 = self.model2.text_process

        if "GENERATOR" in cfg and "DISCRIMINATOR" in cfg:
            g, d = cfg.GENERATOR, cfg.DISCRIMINATOR
            self.g_nfeats, self.g_latent_dim, self.g_ff_size = g.INPUT_DIM, g.LATENT_DIM, g.FF_SIZE
            self.g_num_layers, self.g_num_heads, self.g_dropout = g.NUM_LAYERS, g.NUM_HEADS, g.DROPOUT
            self.d_nfeats, self.d_latent_dim, self.d_ff_size = d.INPUT_DIM, d.LATENT_DIM, d.FF_SIZE
            self.d_num_layers, self.d_num_heads, self.d_dropout = d.NUM_LAYERS, d.NUM_HEADS, d.DROPOUT
            self.nfeats = self.g_nfeats
            gen = (self.g_nfeats, self.g_latent_dim, self.g_ff_size, self.g_num_layers, self.g_num_heads)
            dis = (self.d_nfeats, self.d_latent_dim, self.d_ff_size, self.d_num_layers, self.d_num_heads, self.d_dropout)
        else:
            self.nfeats, self.latent_dim, self.ff_size = cfg.INPUT_DIM, cfg.LATENT_DIM, cfg.FF_SIZE
            self.num_layers, self.num_heads, self.dropout = cfg.NUM_LAYERS, cfg.NUM_HEADS, cfg.DROPOUT
            gen = (self.nfeats, self.latent_dim, self.ff_size, self.num_layers, self.num_heads)
            dis = gen + (self.dropout,)
        self.activation = cfg.ACTIVATION
        self.motion_rep = cfg.MOTION_REP
        self.cfg_mixing_weight = cfg.CFG_WEIGHT
        self.text_dim = 768
        self.mixing_mode = cfg.MIXING_MODE
        self.diffusion_steps = cfg.DIFFUSION_STEPS
        self.beta_scheduler = cfg.BETA_SCHEDULER
        self.sampling_strategy = sampling_strategy
        self.betas = get_named_beta_schedule(self.beta_scheduler, self.diffusion_steps)
        self.diffusion = MixerDiffusion(use_timesteps=space_timesteps(self.diffusion_steps, [self.diffusion_steps]),
                                        betas=self.betas, model_mean_type=ModelMeanType.START_X,
                                        model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE,
                                        rescale_timesteps=False)
        self.sampler = cfg.SAMPLER
        self.force_influnce_val = cfg.FORCE_INFLUENCE_VAL
        self.mixing = Mixer(denoiser1=self.denoiser1, denoiser2=self.denoiser2, nfeats=gen[0], latent_dim=gen[1],
                            ff_size=gen[2], text_dim=self.text_dim, n_blocks=gen[3], n_heads=gen[4],
                            mixing_mode=self.mixing_mode, store_influence=self.store_influence,
                            force_influence_val=self.force_influnce_val, align=self.align, normalizers=normalizers)
        self.discriminator_i = DiscriminatorTransfomer(input_feats=dis[0], latent_dim=dis[1], num_frames=num_frames,
                                                       ff_size=dis[2], num_layers=dis[3], num_heads=dis[4],
                                                       dropout=dis[5], activation=self.activation)
        self.discriminator_I = DiscriminatorTransfomer(input_feats=dis[0] * 2, latent_dim=dis[1], num_frames=num_frames,
                                                       ff_size=dis[2], num_layers=dis[3], num_heads=dis[4],
                                                       dropout=dis[5], activation=self.activation)
        self._build_clip()
        # the reference shares the CLIP tower with model2 (mixermdm.py:229-233)
        self.model2.token_embedding = self.token_embedding
        self.model2.clip_transformer = self.clip_transformer
        self.model2.ln_final = self.ln_final
        self.clipTransEncoder, self.clip_ln = _text_encoder()

    # -- conditioning -----------------------------------------------------------------------------
    def text_process(self, batch, text_name="text", out_name="cond"):
        """reference models/mixermdm.py:283-317: CLIP tower -> clipTransEncoder -> clip_ln -> end-of-text row"""
        batch[out_name] = self._encode_text("mixer", [(self.clipTransEncoder, self.clip_ln)],
                                            self._tokens(batch[text_name]))[0]
        return batch

    def generate_cond(self, batch):
        """(B, 8*768): [interaction | inter-ind1 | inter-ind2 | ind-ind1 | ind-ind2 | mix-inter | mix-ind1 | mix-ind2]
        (reference mixermdm.py:283-356). A precomputed batch["cond"] short-circuits the text path."""
        if "cond" in batch and torch.is_tensor(batch["cond"]):
            return batch["cond"]
        inter = "text_interaction" if "text_interaction" in batch else "text"
        # The reference makes eight text_process calls (five through the sub-models, three of its own), i.e. eight CLIP
        # tower passes. Sequences are independent and model2 shares the tower with this module (mixermdm.py:229-233), so
        # the same eight vectors come from two calls: model1's tower on [individual1; individual2] with its encoder, and
        # the shared tower on [individual1; individual2; interaction] feeding model2's encoder and clipTransEncoder
        # (three calls if model2's own positional_embedding parameter differs from this module's).
        tok = self.tokenizer or text_engine.default_tokenizer
        t1, t2, tI = (torch.as_tensor(tok(batch[k])) for k in ("text_individual1", "text_individual2", inter))
        B = t1.shape[0]
        self.model1.tokenizer = self.model2.tokenizer = self.tokenizer
        (ind,) = self.model1._encode_text("individual", self.model1._text_encoders("individual"), torch.cat([t1, t2]))
        t3 = torch.cat([t1, t2, tI])
        if torch.equal(self.model2.positional_embedding, self.positional_embedding):
            inter_out, mix_out = self._encode_text(
                "generate_cond", self.model2._text_encoders("interaction") + [(self.clipTransEncoder, self.clip_ln)], t3)
        else:   # model2 keeps its own positional_embedding parameter: if it differs, its tower input differs too
            (inter_out,) = self.model2._encode_text("interaction", self.model2._text_encoders("interaction"), t3)
            (mix_out,) = self._encode_text("mixer", [(self.clipTransEncoder, self.clip_ln)], t3)
        batch["cond_individual_individual1"], batch["cond_individual_individual2"] = ind[:B], ind[B:]
        batch["cond_interaction_individual1"], batch["cond_interaction_individual2"] = inter_out[:B], inter_out[B:2 * B]
        batch["cond_interaction"] = inter_out[2 * B:]
        batch["cond_influence_individual1"], batch["cond_influence_individual2"] = mix_out[:B], mix_out[B:2 * B]
        batch["cond_influence_interaction"] = mix_out[2 * B:]
        return torch.cat([batch[k] for k in ("cond_interaction", "cond_interaction_individual1",
                                             "cond_interaction_individual2", "cond_individual_individual1",
                                             "cond_individual_individual2", "cond_influence_interaction",
                                             "cond_influence_individual1", "cond_influence_individual2")], dim=1)

    # -- sampling -----------------------------------------------------------------------------------
    def _run(self, batch, mixer_mode, noise=None):
        self.mixing.mode = mixer_mode
        cond = self.generate_cond(batch)
        B = cond.shape[0]
        T = int(batch["motion_lens"][0])
        self.mixing.history_influence_i1, self.mixing.history_influence_i2 = [], []
        if mixer_mode == "eval":
            self.mixing.history_out1, self.mixing.history_out2, self.mixing.history_out_influenced = [], [], []
        self.diffusion_test = MixerDiffusion(use_timesteps=space_timesteps(self.diffusion_steps, self.sampling_strategy),
                                             betas=self.betas, model_mean_type=ModelMeanType.START_X,
                                             model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE,
                                             rescale_timesteps=False)
        self.cfg_model = ClassifierFreeSampleModelX2(self.mixing, self.cfg_mixing_weight)
        return self.diffusion_test.ddim_sample_loop(self.cfg_model, (B, T, self.nfeats * 2), noise=noise,
                                                    clip_denoised=False, progress=True,
                                                    model_kwargs={"mask": None, "cond": cond}, x_start=None)

    def forward(self, batch):
        """full denoising chain, Mixer mode "eval" (reference mixermdm.py:490-548)"""
        out = self._run(batch, "eval", noise=batch.get("noise"))
        m = self.mixing
        return {"output": out, "influence_i1": m.history_influence_i1, "influence_i2": m.history_influence_i2,
                "out1": m.history_out1, "out2": m.history_out2, "out_influenced": m.history_out_influenced}

    def forward_test(self, batch):
        """same chain without the per-step prediction histories (reference mixermdm.py:550-602)"""
        out = self._run(batch, "eval_intermediate", noise=batch.get("noise"))
        m = self.mixing
        return {"output": out, "influence_i1": m.history_influence_i1, "influence_i2": m.history_influence_i2}

    def compute_loss(self, *a, **k):
        raise NotImplementedError("training (compute_loss, discriminators) stays on the reference's PyTorch path")
