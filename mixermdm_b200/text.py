"""Host side of the text-conditioning path (reference: MixerMDM.text_process / generate_cond,
models/mixermdm.py:283-356; in2IN.text_process, models/in2in.py:109-135): packs the frozen CLIP text tower and the
2-layer text encoders of the (unchanged) module tree for `mmdm_text_encode` and forwards token ids to it.
Tokenisation is the only host step (OpenAI `clip.tokenize`, a pure-Python BPE; any callable with the same contract
can be plugged in). No arithmetic of the path runs in PyTorch and there is no CPU fallback."""
import ctypes as C

import torch

from . import _lib
from ._lib import MmdmError, check, cur_stream

CTX = 77


class ClipBlock(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "ln1_g", "ln1_b", "attn_in_w", "attn_in_b", "attn_out_w", "attn_out_b", "ln2_g", "ln2_b", "fc_w", "fc_b",
        "proj_w", "proj_b")]


class TextTower(C.Structure):
    _fields_ = [("width", C.c_int32), ("layers", C.c_int32), ("heads", C.c_int32), ("ctx", C.c_int32),
                ("vocab", C.c_int32), ("tok_emb", C.c_void_p), ("pos_emb", C.c_void_p),
                ("blocks", C.POINTER(ClipBlock)), ("lnf_g", C.c_void_p), ("lnf_b", C.c_void_p)]


class EncLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "in_w", "in_b", "out_w", "out_b", "norm1_g", "norm1_b", "lin1_w", "lin1_b", "lin2_w", "lin2_b", "norm2_g",
        "norm2_b")]


class TextEncoder(C.Structure):
    _fields_ = [("layers", C.c_int32), ("heads", C.c_int32), ("ff", C.c_int32), ("layers_", C.POINTER(EncLayer)),
                ("ln_g", C.c_void_p), ("ln_b", C.c_void_p)]


def default_tokenizer(texts):
    """clip.tokenize(texts, truncate=True) -> (B, 77) int tensor; the `clip` package supplies the BPE vocabulary"""
    try:
        import clip
    except ImportError as e:
        raise MmdmError("text prompts need a tokenizer: install OpenAI `clip` (its BPE tokenizer is pure Python), set "
                        "`model.tokenizer = callable(list[str]) -> (B, 77) int tensor`, or pass a precomputed "
                        "batch['cond'] (B, 6144)") from e
    return clip.tokenize(texts, truncate=True)


class _Pack:
    def __init__(self, device):
        self.device, self.keep = device, []

    def bf16(self, t):
        t = t.detach().to(device=self.device, dtype=torch.bfloat16).contiguous()
        self.keep.append(t)
        return C.c_void_p(t.data_ptr())

    def f32(self, t):
        t = t.detach().to(device=self.device, dtype=torch.float32).contiguous()
        self.keep.append(t)
        return C.c_void_p(t.data_ptr())


class TextEngine:
    """One CLIP tower plus the encoders that read its output (packed once, repacked when the parameters change)."""

    def __init__(self, tower, encoders, device):
        """tower: object with token_embedding / clip_transformer.resblocks / positional_embedding / ln_final;
        encoders: list of (nn.TransformerEncoder container, nn.LayerNorm container)"""
        from .engine import weights_fingerprint
        self.lib = _lib.load()
        from .engine import canonical_device
        self.device = canonical_device(device)
        if self.device.type != "cuda":
            raise MmdmError("the text path runs on CUDA devices only (sm_100a); there is no CPU path")
        pk = _Pack(self.device)
        blocks = list(tower.clip_transformer.resblocks)
        arr = (ClipBlock * len(blocks))()
        for i, b in enumerate(blocks):
            a = arr[i]
            a.ln1_g, a.ln1_b = pk.f32(b.ln_1.weight), pk.f32(b.ln_1.bias)
            a.attn_in_w, a.attn_in_b = pk.bf16(b.attn.in_proj_weight), pk.f32(b.attn.in_proj_bias)
            a.attn_out_w, a.attn_out_b = pk.bf16(b.attn.out_proj.weight), pk.f32(b.attn.out_proj.bias)
            a.ln2_g, a.ln2_b = pk.f32(b.ln_2.weight), pk.f32(b.ln_2.bias)
            a.fc_w, a.fc_b = pk.bf16(b.mlp.c_fc.weight), pk.f32(b.mlp.c_fc.bias)
            a.proj_w, a.proj_b = pk.bf16(b.mlp.c_proj.weight), pk.f32(b.mlp.c_proj.bias)
        tw = TextTower()
        tw.width = tower.token_embedding.weight.shape[1]
        tw.layers, tw.heads = len(blocks), blocks[0].attn.num_heads
        tw.ctx, tw.vocab = tower.positional_embedding.shape[0], tower.token_embedding.weight.shape[0]
        tw.tok_emb, tw.pos_emb = pk.f32(tower.token_embedding.weight), pk.f32(tower.positional_embedding)
        tw.blocks = C.cast(arr, C.POINTER(ClipBlock))
        tw.lnf_g, tw.lnf_b = pk.f32(tower.ln_final.weight), pk.f32(tower.ln_final.bias)
        encs = (TextEncoder * len(encoders))()
        ff_max = 4 * tw.width
        for e, (enc, ln) in enumerate(encoders):
            layers = list(enc.layers)
            la = (EncLayer * len(layers))()
            for i, l in enumerate(layers):
                a = la[i]
                a.in_w, a.in_b = pk.bf16(l.self_attn.in_proj_weight), pk.f32(l.self_attn.in_proj_bias)
                a.out_w, a.out_b = pk.bf16(l.self_attn.out_proj.weight), pk.f32(l.self_attn.out_proj.bias)
                a.norm1_g, a.norm1_b = pk.f32(l.norm1.weight), pk.f32(l.norm1.bias)
                a.lin1_w, a.lin1_b = pk.bf16(l.linear1.weight), pk.f32(l.linear1.bias)
                a.lin2_w, a.lin2_b = pk.bf16(l.linear2.weight), pk.f32(l.linear2.bias)
                a.norm2_g, a.norm2_b = pk.f32(l.norm2.weight), pk.f32(l.norm2.bias)
            pk.keep.append(la)
            encs[e].layers, encs[e].heads = len(layers), layers[0].self_attn.num_heads
            encs[e].ff = layers[0].linear1.out_features
            encs[e].layers_ = C.cast(la, C.POINTER(EncLayer))
            encs[e].ln_g, encs[e].ln_b = pk.f32(ln.weight), pk.f32(ln.bias)
            ff_max = max(ff_max, encs[e].ff)
        pk.keep += [arr, encs]
        self.tower, self.encs, self.n_enc, self.keep, self.ff_max = tw, encs, len(encoders), pk.keep, ff_max
        self.ws = None
        self.lib.mmdm_text_workspace_bytes.restype = C.c_size_t
        mods = [tower.token_embedding, tower.clip_transformer, tower.ln_final] + [m for pair in encoders for m in pair]
        self._mods, self._pos = mods, tower.positional_embedding
        self.fingerprint = self._fingerprint(weights_fingerprint)

    def _fingerprint(self, fn):
        return hash(tuple(fn(m) for m in self._mods) + (self._pos.data_ptr(), self._pos._version))

    def stale(self):
        from .engine import weights_fingerprint
        return self._fingerprint(weights_fingerprint) != self.fingerprint

    def encode(self, tokens):
        """tokens: (nseq, 77) integer tensor (host or device) -> list of n_enc (nseq, 768) fp32 device tensors"""
        tokens = torch.as_tensor(tokens)
        if tokens.dim() != 2 or tokens.shape[1] != self.tower.ctx:
            raise ValueError(f"tokens: expected (nseq, {self.tower.ctx}), got {tuple(tokens.shape)}")
        nseq = tokens.shape[0]
        tok = tokens.to(device=self.device, dtype=torch.int32).contiguous()
        eot = tokens.argmax(dim=-1).to(device=self.device, dtype=torch.int32).contiguous()   # end-of-text position
        need = self.lib.mmdm_text_workspace_bytes(nseq, self.tower.ctx, self.tower.width, self.ff_max)
        if self.ws is None or self.ws.numel() < need + 1024:
            self.ws = torch.empty(need + 1024, dtype=torch.uint8, device=self.device)
        base = (self.ws.data_ptr() + 1023) // 1024 * 1024
        outs = [torch.empty(nseq, self.tower.width, dtype=torch.float32, device=self.device) for _ in range(self.n_enc)]
        ptrs = (C.c_void_p * self.n_enc)(*[o.data_ptr() for o in outs])
        with torch.cuda.device(self.device):
            check(self.lib.mmdm_text_encode(C.byref(self.tower), self.encs, self.n_enc, C.c_void_p(tok.data_ptr()),
                                            C.c_void_p(eot.data_ptr()), nseq, ptrs, self.tower.width, C.c_void_p(base),
                                            C.c_size_t(need), cur_stream()), "mmdm_text_encode")
        return outs
