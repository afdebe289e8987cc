"""TEST INFRASTRUCTURE ONLY - CPU restatement of the text-conditioning path for parity tests.

The reference (models/mixermdm.py:283-356, models/in2in.py:109-135) runs a frozen CLIP text tower that lives in a
third-party package absent from /root/reference: OpenAI `clip` 1.0 (environment.yaml:49), model "ViT-L/14@336px".
Its published text transformer (clip/model.py: QuickGELU, ResidualAttentionBlock, Transformer,
CLIP.build_attention_mask) is restated below from the architecture description: pre-LN residual blocks
x = x + attn(ln_1(x), causal mask); x = x + c_proj(QuickGELU(c_fc(ln_2(x)))), nn.MultiheadAttention on (L, N, D)
inputs, LayerNorm eps 1e-5. Everything around it (token + positional embedding, ln_final, the 2-layer
nn.TransformerEncoder, LayerNorm, end-of-text row, the order of the eight vectors) is the reference's OWN code:
oracle/make_goldens.py text() calls its unmodified `generate_cond` on holder objects whose CLIP members are the
modules below, with `clip.tokenize` replaced by `stub_tokenize`; tests/golden/reference_text.npz holds the result.
Parity of the CLIP blocks themselves is therefore pinned to this restatement, not to the (absent) package."""
import zlib
from collections import OrderedDict

import numpy as np
import torch
from torch import nn

CTX, VOCAB, WIDTH, LAYERS, HEADS = 77, 49408, 768, 12, 12
SOT, EOT = 49406, 49407


def stub_tokenize(texts, context_length=CTX, truncate=True):
    """deterministic stand-in for clip.tokenize: <sot> word-hash ids <eot>, zero padded; the end-of-text id is the
    largest, so argmax finds it exactly as in the reference (models/mixermdm.py:315)"""
    if isinstance(texts, str):
        texts = [texts]
    out = torch.zeros(len(texts), context_length, dtype=torch.long)
    for i, t in enumerate(texts):
        ids = [SOT] + [1 + zlib.crc32(w.encode()) % 49000 for w in t.lower().split()] + [EOT]
        if len(ids) > context_length:
            ids = ids[:context_length]
            ids[-1] = EOT
        out[i, :len(ids)] = torch.tensor(ids)
    return out


class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model, n_head, attn_mask=None):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = nn.LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)), ("gelu", QuickGELU()),
                                              ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = nn.LayerNorm(d_model)
        self.attn_mask = attn_mask

    def forward(self, x):
        m = self.attn_mask.to(dtype=x.dtype, device=x.device) if self.attn_mask is not None else None
        y = self.ln_1(x)
        x = x + self.attn(y, y, y, need_weights=False, attn_mask=m)[0]
        return x + self.mlp(self.ln_2(x))


class Transformer(nn.Module):
    def __init__(self, width=WIDTH, layers=LAYERS, heads=HEADS, attn_mask=None):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask) for _ in range(layers)])

    def forward(self, x):
        return self.resblocks(x)


def build_attention_mask(ctx=CTX):
    return torch.full((ctx, ctx), float("-inf")).triu_(1)


def text_encoder_pair():
    """the reference's own construction (models/mixermdm.py:241-256, models/in2in.py:24-52)"""
    layer = nn.TransformerEncoderLayer(d_model=768, nhead=8, dim_feedforward=2048, dropout=0.1, activation="gelu",
                                       batch_first=True)
    return nn.TransformerEncoder(layer, num_layers=2), nn.LayerNorm(768)


def attach_clip(holder, layers=LAYERS):
    holder.token_embedding = nn.Embedding(VOCAB, WIDTH)
    holder.clip_transformer = Transformer(WIDTH, layers, HEADS, build_attention_mask())
    holder.positional_embedding = nn.Parameter(torch.zeros(CTX, WIDTH))
    holder.ln_final = nn.LayerNorm(WIDTH)
    holder.dtype = torch.float32


def text_state_dict(template, seed=0):
    """per-name seeded synthetic values (same stream as oracle/synth.py); LayerNorm gains are centred on 1"""
    from oracle import synth
    sd = synth.synthetic_state_dict(template, seed=seed)
    for k in sd:
        leaf = k.split(".")[-2] if "." in k else ""
        if k.endswith(".weight") and (leaf.startswith("ln_") or leaf.startswith("norm") or leaf.startswith("clip_ln")):
            sd[k] = sd[k] + 1.0
    return sd


def text_keys(state_dict):
    """the keys of a MixerMDM state dict that belong to the text path"""
    pats = ("token_embedding.", "clip_transformer.", "positional_embedding", "ln_final.", "clipTransEncoder", "clip_ln")
    return [k for k in state_dict if any(p in k for p in pats)]


def text_process(holder, enc, ln, tokens):
    """restatement of MixerMDM.text_process (models/mixermdm.py:283-317) for given token ids"""
    with torch.no_grad():
        x = holder.token_embedding(tokens) + holder.positional_embedding
        x = holder.clip_transformer(x.permute(1, 0, 2)).permute(1, 0, 2)
        out = ln(enc(holder.ln_final(x)))
        return out[torch.arange(x.shape[0]), tokens.argmax(dim=-1)]


class _Holder(nn.Module):
    pass


def build_text_oracle(seed=0):
    """module tree with the real model's attribute names (= state-dict keys) for the text path, synthetic weights"""
    m = _Holder()
    m.model1, m.model2 = _Holder(), _Holder()
    attach_clip(m.model1)
    m.model1.clipTransEncoder_individual, m.model1.clip_ln_individual = text_encoder_pair()
    attach_clip(m)
    m.model2.token_embedding, m.model2.clip_transformer = m.token_embedding, m.clip_transformer
    # (like the reference, models/mixermdm.py:229-233: model2 borrows the token table, the blocks and ln_final but keeps
    #  its own positional_embedding parameter)
    m.model2.positional_embedding = nn.Parameter(torch.zeros(CTX, WIDTH))
    m.model2.ln_final = m.ln_final
    m.model2.clipTransEncoder_interaction, m.model2.clip_ln_interaction = text_encoder_pair()
    m.clipTransEncoder, m.clip_ln = text_encoder_pair()
    sd = text_state_dict(m.state_dict(), seed=seed)
    m.load_state_dict(sd, strict=True)
    return m.eval(), sd


def generate_cond(m, prompts, tokenize=stub_tokenize):
    """restatement of MixerMDM.generate_cond (models/mixermdm.py:319-356): the eight vectors in the reference's order"""
    inter = "text_interaction" if "text_interaction" in prompts else "text"
    t1, t2, tI = (tokenize(prompts[k]) for k in ("text_individual1", "text_individual2", inter))
    e1 = (m.model1.clipTransEncoder_individual, m.model1.clip_ln_individual)
    e2 = (m.model2.clipTransEncoder_interaction, m.model2.clip_ln_interaction)
    e3 = (m.clipTransEncoder, m.clip_ln)
    parts = [text_process(m.model2, *e2, tI), text_process(m.model2, *e2, t1), text_process(m.model2, *e2, t2),
             text_process(m.model1, *e1, t1), text_process(m.model1, *e1, t2),
             text_process(m, *e3, tI), text_process(m, *e3, t1), text_process(m, *e3, t2)]
    return torch.cat(parts, dim=1)
