"""Host side of the engine: packs the weights of the (unchanged) module tree into the layouts the kernels want,
owns the device workspace (a torch tensor used as raw memory) and forwards calls to the C ABI.
PyTorch is plumbing here (device memory, dtype casts at load time, streams); no arithmetic of the path runs in it."""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib, schedule
from ._lib import MmdmError, check, cur_stream, ptr

NFEATS = 262
KPAD = 264
TEXT_DIM = 768
NET_INDIVIDUAL, NET_INTERACTION, NET_INFLUENCE = 0, 1, 2


# ------------------------------------------------------------------------------------------------
# ctypes mirrors of include/mixermdm_b200.h
# ------------------------------------------------------------------------------------------------
class BlockWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "sa_in_w", "sa_in_b", "sa_out_w", "sa_out_b", "ca_q_w", "ca_q_b", "ca_kv_w", "ca_kv_b", "ca_out_w",
        "ca_out_b", "ffn_w1", "ffn_b1", "ffn_w2", "ffn_b2")]


class NetWeights(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("D", C.c_int32), ("F", C.c_int32), ("L", C.c_int32), ("H", C.c_int32),
        ("ada0_w", C.c_void_p), ("ada0_b", C.c_void_p), ("ada0_n", C.c_int32),
        ("ada1_w", C.c_void_p), ("ada1_b", C.c_void_p), ("ada1_n", C.c_int32),
        ("embed_w", C.c_void_p), ("embed_b", C.c_void_p),
        ("pe", C.c_void_p), ("pe_rows", C.c_int32),
        ("out_w", C.c_void_p), ("out_b", C.c_void_p),
        ("time_w0", C.c_void_p), ("time_b0", C.c_void_p), ("time_w2", C.c_void_p), ("time_b2", C.c_void_p),
        ("text_w", C.c_void_p), ("text_b", C.c_void_p),
        ("blocks", C.POINTER(BlockWeights)),
        ("comp0_w", C.c_void_p), ("comp0_b", C.c_void_p), ("comp0_n", C.c_int32),
        ("comp1_w", C.c_void_p), ("comp1_b", C.c_void_p), ("comp1_n", C.c_int32),
    ]


class Config(C.Structure):
    _fields_ = [
        ("nfeats", C.c_int32), ("kpad", C.c_int32), ("text_dim", C.c_int32), ("nsteps", C.c_int32),
        ("mixing_mode", C.c_int32), ("align", C.c_int32), ("center", C.c_int32), ("has_force", C.c_int32),
        ("force_val", C.c_float), ("cfg_scale", C.c_float),
        ("timestep_map", C.POINTER(C.c_int32)), ("coef", C.c_void_p),
        ("head_w", C.c_void_p), ("head_b", C.c_void_p),
        ("mean1", C.c_void_p), ("std1", C.c_void_p), ("mean2", C.c_void_p), ("std2", C.c_void_p),
    ]


class MixstepArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("T", C.c_int32), ("mixing_mode", C.c_int32), ("ncfg", C.c_int32),
        ("mix_only", C.c_int32), ("center", C.c_int32), ("has_force", C.c_int32),
        ("force_val", C.c_float), ("cfg_scale", C.c_float),
        ("step_ptr", C.c_void_p), ("nsteps", C.c_int32), ("coef", C.c_void_p),
        ("weights", C.c_void_p),
        ("out1a", C.c_void_p), ("out2a", C.c_void_p), ("x1", C.c_void_p), ("x2", C.c_void_p),
        ("final_out", C.c_void_p), ("mix_out", C.c_void_p),
        ("mean1", C.c_void_p), ("std1", C.c_void_p), ("mean2", C.c_void_p), ("std2", C.c_void_p),
        ("hist_i1", C.c_void_p), ("hist_i2", C.c_void_p),
    ]


class ProfileEntry(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("launches", C.c_int64), ("ms", C.c_double), ("flops", C.c_double),
                ("bytes", C.c_double)]


# ------------------------------------------------------------------------------------------------
# normaliser statistics (reference utils/utils.py:44-82 loads them relative to the CWD)
# ------------------------------------------------------------------------------------------------
class MotionStats:
    """(mean, std) of one motion space; attribute names follow MotionNormalizerTorch*."""

    def __init__(self, mean, std):
        self.motion_mean = torch.as_tensor(np.asarray(mean, dtype=np.float32))
        self.motion_std = torch.as_tensor(np.asarray(std, dtype=np.float32))

    # Caller-side helpers with the reference's names (utils/utils.py:49-61, 69-82). Inside the sampling loop both are
    # folded into the CUDA kernels (output-layer epilogue, fused mix kernel); these serve code that holds a
    # `mixing.normalizer_model*` and converts a tensor outside the loop.
    def forward(self, x):
        return (x - self.motion_mean.to(x.device)) / self.motion_std.to(x.device)

    def backward(self, x):
        return x * self.motion_std.to(x.device) + self.motion_mean.to(x.device)


def load_normalizers(normalizers=None):
    """returns (normalizer_model1 [HumanML3D stats, individual model], normalizer_model2 [InterHuman stats]).
    `normalizers` may be ((mean1, std1), (mean2, std2)); otherwise the reference's files are read from ./data."""
    if normalizers is not None:
        (m1, s1), (m2, s2) = normalizers
        return MotionStats(m1, s1), MotionStats(m2, s2)
    paths = ["./data/HumanML3D/mean_ih_new.npy", "./data/HumanML3D/std_ih_new.npy", "./data/global_mean.npy",
             "./data/global_std.npy"]
    missing = [p for p in paths if not os.path.exists(p)]
    if missing:
        raise FileNotFoundError(f"normaliser statistics not found (reference layout, relative to CWD): {missing}")
    m1, s1, m2, s2 = [np.load(p) for p in paths]
    return MotionStats(m1, s1), MotionStats(m2, s2)


# ------------------------------------------------------------------------------------------------
# weight packing
# ------------------------------------------------------------------------------------------------
class _Packer:
    def __init__(self, device):
        self.device = device
        self.keep = []
        # composite conditioning tables of the LN fold; MMDM_LN_FOLD=0: none, the engine then runs the
        # stand-alone LayerNorm kernel at every frame count
        self.fold = os.environ.get("MMDM_LN_FOLD", "1") != "0"

    def bf16(self, t):
        t = t.detach().to(device=self.device, dtype=torch.bfloat16).contiguous()
        self.keep.append(t)
        return t

    def f32(self, t):
        t = t.detach().to(device=self.device, dtype=torch.float32).contiguous()
        self.keep.append(t)
        return t

    def p(self, t):
        return C.c_void_p(t.data_ptr())

    def composite(self, blocks, sites, D):
        """Composite conditioning table of the LN fold for `sites` = [(norm getter, linear getter)], per block.
        Load-time weight folding (like folding a BatchNorm into a convolution): for the Linear (W, c) behind an AdaLN
        whose modulation is [scale; shift] = Wada E + bada, the rows [W Wsc; W Wsh] and constants
        [W (1 + bsc); W bsh + c], evaluated in fp64 on the bf16 values the kernels multiply with, stored as bf16
        pairs [hi | lo] along K. Returns ((sum 2N, 2D) bf16, (sum 2N) fp32)."""
        rows, consts = [], []
        rnd = lambda t: t.detach().to(device=self.device, dtype=torch.bfloat16).double()
        for blk in blocks:
            for norm_of, lin_of in sites:
                ada = norm_of(blk).emb_layers[1]
                wada, bada = rnd(ada.weight), ada.bias.detach().to(self.device).double()
                lw, lb = lin_of(blk)
                W, c = rnd(lw), lb.detach().to(self.device).double()
                rows += [W @ wada[:D], W @ wada[D:]]
                consts += [W @ (1.0 + bada[:D]), W @ bada[D:] + c]
        full = torch.cat(rows, 0)
        hi = full.to(torch.bfloat16)
        lo = (full - hi.double()).to(torch.bfloat16)
        table = torch.cat([hi, lo], 1).contiguous()
        const = torch.cat(consts, 0).float().contiguous()
        self.keep += [table, const]
        return table, const

    def pack_net(self, kind, net, blocks, with_out):
        """net: module with motion_embed/text_embed/embed_timestep/sequence_pos_encoder; blocks: ModuleList"""
        D = net.motion_embed.out_features
        att0 = blocks[0].sa_block.attention
        H = att0.num_heads
        F = blocks[0].ffn.linear1.out_features
        L = len(blocks)
        w = NetWeights()
        w.kind, w.D, w.F, w.L, w.H = kind, D, F, L, H

        def ada_cat(getters):
            ws, bs = [], []
            for blk in blocks:
                for g in getters:
                    lin = g(blk).emb_layers[1]
                    ws.append(lin.weight)
                    bs.append(lin.bias)
            return self.bf16(torch.cat([x.detach() for x in ws], 0)), self.f32(torch.cat([x.detach() for x in bs], 0))

        if kind == NET_INFLUENCE:
            g0 = [lambda b: b.sa_block.norm]
            g1 = [lambda b: b.ca_block.norm, lambda b: b.ca_block.xf_norm, lambda b: b.ffn.norm]
        elif kind == NET_INTERACTION:
            g0 = [lambda b: b.sa_block.norm, lambda b: b.ffn.norm]
            g1 = [lambda b: b.ca_block.norm, lambda b: b.ca_block.xf_norm]
        else:
            g0 = [lambda b: b.sa_block.norm, lambda b: b.ffn.norm]
            g1 = []
        a0w, a0b = ada_cat(g0)
        w.ada0_w, w.ada0_b, w.ada0_n = self.p(a0w), self.p(a0b), a0w.shape[0]
        if g1:
            a1w, a1b = ada_cat(g1)
            w.ada1_w, w.ada1_b, w.ada1_n = self.p(a1w), self.p(a1b), a1w.shape[0]
        else:
            w.ada1_w, w.ada1_b, w.ada1_n = None, None, 0
        # LN fold (include/mixermdm_b200.h, "composite conditioning tables"): the Linear behind every AdaLN site,
        # in the same per-block order as g0 / g1. The Influence xf_norm has no entry (its rows, m_I, come from no GEMM
        # of the block), so it is absent from c1 there.
        sa_lin = lambda b: (b.sa_block.attention.in_proj_weight, b.sa_block.attention.in_proj_bias)
        ffn_lin = lambda b: (b.ffn.linear1.weight, b.ffn.linear1.bias)
        caq_lin = lambda b: (b.ca_block.attention.in_proj_weight[:D], b.ca_block.attention.in_proj_bias[:D])
        cakv_lin = lambda b: (b.ca_block.attention.in_proj_weight[D:], b.ca_block.attention.in_proj_bias[D:])
        if kind == NET_INFLUENCE:
            c0 = [(g0[0], sa_lin)]
            c1 = [(g1[0], caq_lin), (g1[2], ffn_lin)]
        elif kind == NET_INTERACTION:
            c0 = [(g0[0], sa_lin), (g0[1], ffn_lin)]
            c1 = [(g1[0], caq_lin), (g1[1], cakv_lin)]
        else:
            c0 = [(g0[0], sa_lin), (g0[1], ffn_lin)]
            c1 = []
        if self.fold and D % 64 == 0:
            cw, cb = self.composite(blocks, c0, D)
            w.comp0_w, w.comp0_b, w.comp0_n = self.p(cw), self.p(cb), cw.shape[0]
            if c1:
                cw, cb = self.composite(blocks, c1, D)
                w.comp1_w, w.comp1_b, w.comp1_n = self.p(cw), self.p(cb), cw.shape[0]
        ew = torch.zeros(D, KPAD, dtype=torch.float32)
        ew[:, :NFEATS] = net.motion_embed.weight.detach().float().cpu()
        w.embed_w, w.embed_b = self.p(self.bf16(ew)), self.p(self.f32(net.motion_embed.bias))
        pe = self.f32(net.sequence_pos_encoder.pe)
        w.pe, w.pe_rows = self.p(pe), pe.shape[0]
        if with_out:
            w.out_w, w.out_b = self.p(self.bf16(net.out.linear.weight)), self.p(self.f32(net.out.linear.bias))
        else:
            w.out_w, w.out_b = None, None
        te = net.embed_timestep.time_embed
        w.time_w0, w.time_b0 = self.p(self.f32(te[0].weight)), self.p(self.f32(te[0].bias))
        w.time_w2, w.time_b2 = self.p(self.f32(te[2].weight)), self.p(self.f32(te[2].bias))
        w.text_w, w.text_b = self.p(self.f32(net.text_embed.weight)), self.p(self.f32(net.text_embed.bias))
        arr = (BlockWeights * L)()
        for l, blk in enumerate(blocks):
            bw = arr[l]
            sa = blk.sa_block.attention
            bw.sa_in_w, bw.sa_in_b = self.p(self.bf16(sa.in_proj_weight)), self.p(self.f32(sa.in_proj_bias))
            bw.sa_out_w, bw.sa_out_b = self.p(self.bf16(sa.out_proj.weight)), self.p(self.f32(sa.out_proj.bias))
            if kind != NET_INDIVIDUAL:
                ca = blk.ca_block.attention
                bw.ca_q_w, bw.ca_q_b = self.p(self.bf16(ca.in_proj_weight[:D])), self.p(self.f32(ca.in_proj_bias[:D]))
                bw.ca_kv_w, bw.ca_kv_b = self.p(self.bf16(ca.in_proj_weight[D:])), self.p(self.f32(ca.in_proj_bias[D:]))
                bw.ca_out_w, bw.ca_out_b = self.p(self.bf16(ca.out_proj.weight)), self.p(self.f32(ca.out_proj.bias))
            bw.ffn_w1, bw.ffn_b1 = self.p(self.bf16(blk.ffn.linear1.weight)), self.p(self.f32(blk.ffn.linear1.bias))
            bw.ffn_w2, bw.ffn_b2 = self.p(self.bf16(blk.ffn.linear2.weight)), self.p(self.f32(blk.ffn.linear2.bias))
        self.keep.append(arr)
        w.blocks = C.cast(arr, C.POINTER(BlockWeights))
        return w


def weights_fingerprint(module):
    """Identity + version of every parameter and buffer under `module`. The engine runs on PACKED copies (bf16 casts,
    concatenated AdaLN matrices) made once; the reference always reads the live parameters. Comparing this value
    before each forward makes load_state_dict(), in-place updates (optimizer steps, .copy_()) and .to()/.half()
    rebuild the packed copies instead of silently running on stale weights."""
    acc = []
    for t in list(module.parameters()) + list(module.buffers()):
        acc.append((t.data_ptr(), t._version, t.dtype, t.device))
    return hash(tuple(acc))


def _absent_net(kind):
    w = NetWeights()
    w.kind, w.D, w.F, w.L, w.H = kind, 0, 0, 0, 0
    return w


class _Handle:
    """owns one mmdm_handle + its workspace"""

    def __init__(self, cfg, nets, keep, device):
        self.lib = _lib.load()
        if device.type != "cuda":
            raise MmdmError("mixermdm_b200 runs on CUDA devices only (sm_100a); there is no CPU path")
        self.device = device
        self.keep = keep
        self.cfg = cfg
        arr = (NetWeights * 3)(*nets)
        self.keep.append(arr)
        self.h = C.c_void_p()
        self.lib.mmdm_workspace_bytes.restype = C.c_size_t
        self.lib.mmdm_launch_count.restype = C.c_int64
        with torch.cuda.device(device):
            check(self.lib.mmdm_create(C.byref(cfg), arr, C.byref(self.h)), "mmdm_create")   # device of the handle
        self.ws = None
        self.B = self.T = 0

    def call(self, name, *args):
        """invoke handle entry point `name`(h, *args); the C side makes the handle's device current for the call"""
        check(getattr(self.lib, name)(self.h, *args), name, self.h)

    def set_key_mask(self, mask, frames=None):
        """frame mask for the attention layers of the following calls (reference: the `mask` argument of
        in2INDenoiser / Influence, (B, T, P) float with 1 = real frame; they read mask[..., 0]). None clears it."""
        if mask is None:
            self.call("mmdm_set_key_mask", None, 0, self.stream())
            return
        m = mask[..., 0] if mask.dim() == 3 else mask
        if m.dim() != 2 or (frames is not None and m.shape[1] != frames):
            raise ValueError(f"mask: expected (B, {frames}) or (B, {frames}, P), got {tuple(mask.shape)}")
        m = m.detach().to(device=self.device, dtype=torch.float32).contiguous()
        self.call("mmdm_set_key_mask", ptr(m), int(m.shape[0]), self.stream())
        self._mask_keep = m   # (the words are built on the stream: keep the source alive until then)

    def stream(self):
        """the caller's current torch stream ON THE HANDLE'S DEVICE (not on whatever device happens to be current)"""
        return cur_stream(self.device)

    def bind(self, B, T):
        if self.ws is not None and B == self.B and T == self.T:
            return
        need = self.lib.mmdm_workspace_bytes(self.h, B, T)
        if need == 0:
            raise MmdmError(f"invalid workspace request B={B} T={T}")
        if self.ws is None or self.ws.numel() < need:
            self.ws = None
            self.ws = torch.empty(need + 1024, dtype=torch.uint8, device=self.device)
        base = (self.ws.data_ptr() + 1023) // 1024 * 1024
        self.call("mmdm_bind_workspace", C.c_void_p(base), C.c_size_t(need), B, T)
        self.B, self.T = B, T

    def launches(self):
        return int(self.lib.mmdm_launch_count(self.h))

    def graph_captures(self):
        self.lib.mmdm_graph_captures.restype = C.c_int64
        return int(self.lib.mmdm_graph_captures(self.h))

    def profile(self, on):
        self.call("mmdm_profile_enable", int(on))

    def profile_read(self):
        arr = (ProfileEntry * 64)()
        n = self.lib.mmdm_profile_read(self.h, arr, 64)
        return [dict(name=arr[i].name.decode(), launches=arr[i].launches, ms=arr[i].ms, flops=arr[i].flops,
                     bytes=arr[i].bytes) for i in range(n)]

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.h.value:
                self.lib.mmdm_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass


def canonical_device(device):
    """torch.device with an explicit index ("cuda" and "cuda:0" must not look like two devices to the engine caches)"""
    device = torch.device(device)
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def _f32c(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def _expect(t, shape, name):
    """the C ABI takes raw pointers: a wrong shape would be silently wrong rows or an out-of-bounds read"""
    if tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")


def _denoiser_call(hd, which, x, timesteps, cond, mask=None):
    """in2INDenoiser.forward through the handle. Like the reference (models/in2in.py:414-424, slices of cond) a cond
    wider than the network's 768 / 2304 columns is accepted and its leading columns are used."""
    Bx, T = x.shape[0], x.shape[1]
    persons = 2 if which == NET_INTERACTION else 1
    ncond = (3 if which == NET_INTERACTION else 1) * TEXT_DIM
    _expect(x, (Bx, T, persons * NFEATS), "x")
    _expect(timesteps, (Bx,), "timesteps")
    if cond.dim() != 2 or cond.shape[0] != Bx or cond.shape[1] < ncond:
        raise ValueError(f"cond: expected ({Bx}, >={ncond}), got {tuple(cond.shape)}")
    x, cond = _f32c(x, hd.device), _f32c(cond[:, :ncond], hd.device)
    ts = timesteps.detach().to(device=hd.device, dtype=torch.int64).contiguous()
    out = torch.empty_like(x)
    hd.set_key_mask(mask, T)
    hd.call("mmdm_denoiser_forward", which, ptr(x), ptr(ts), ptr(cond), ptr(out), Bx, hd.stream())
    return out


class MixerEngine:
    """engine for one `Mixer` module (both denoisers + Influence + fused mix / sampler step)"""

    def __init__(self, mixer, device, sched=None, cfg_scale=3.5, center=True):
        device = canonical_device(device)
        pk = _Packer(device)
        nets = [pk.pack_net(NET_INDIVIDUAL, mixer.denoiser1, mixer.denoiser1.blocks, True),
                pk.pack_net(NET_INTERACTION, mixer.denoiser2, mixer.denoiser2.blocks, True),
                pk.pack_net(NET_INFLUENCE, mixer, mixer.influence.blocks, False)]
        if sched is None:   # the reference's inference default: "ddim50" of 1000 cosine steps
            sched = schedule.SpacedSchedule(schedule.space_timesteps(1000, "ddim50"),
                                            schedule.get_named_beta_schedule("cosine", 1000))
        self.sched = sched
        self.center = bool(center)
        cfg = Config()
        cfg.nfeats, cfg.kpad, cfg.text_dim, cfg.nsteps = NFEATS, KPAD, TEXT_DIM, self.sched.num_timesteps
        cfg.mixing_mode = int(mixer.mixing_mode)
        cfg.align = int(bool(mixer.align))
        cfg.center = int(bool(center))
        cfg.has_force = int(mixer.force_influence_val is not None)
        cfg.force_val = float(mixer.force_influence_val or 0.0)
        cfg.cfg_scale = float(cfg_scale)
        tmap = (C.c_int32 * cfg.nsteps)(*self.sched.timestep_map)
        pk.keep.append(tmap)
        cfg.timestep_map = C.cast(tmap, C.POINTER(C.c_int32))
        self.coef = pk.f32(torch.from_numpy(self.sched.ddim_coefficients()))   # rewritten in place by set_sampler()
        cfg.coef = pk.p(self.coef)
        self.sampler, self.sigma = "ddim", None
        self.head_w, self.head_b = pk.bf16(mixer.influence.out.weight), pk.f32(mixer.influence.out.bias)
        cfg.head_w, cfg.head_b = pk.p(self.head_w), pk.p(self.head_b)
        n1, n2 = mixer.normalizer_model1, mixer.normalizer_model2
        self.mean1, self.std1 = pk.f32(n1.motion_mean), pk.f32(n1.motion_std)
        self.mean2, self.std2 = pk.f32(n2.motion_mean), pk.f32(n2.motion_std)
        cfg.mean1, cfg.std1, cfg.mean2, cfg.std2 = pk.p(self.mean1), pk.p(self.std1), pk.p(self.mean2), pk.p(self.std2)
        self.hd = _Handle(cfg, nets, pk.keep, device)
        self.device = device
        self.fingerprint = weights_fingerprint(mixer)
        self.nsteps = cfg.nsteps
        self.mixing_mode = cfg.mixing_mode
        self.cfg_scale = float(cfg_scale)

    # -- plumbing ---------------------------------------------------------------------------------
    @property
    def lib(self):
        return self.hd.lib

    def bind(self, B, T):
        self.hd.bind(B, T)

    def _check_batch(self, B):
        """The reference's centring calls torch.cross without `dim` (utils/alignment.py:198): for a batch of exactly
        3 the legacy rule picks the BATCH axis, so every sample is turned by +-90 degrees about X depending on the
        other two samples' hips. That cross-sample artefact is not reproduced; rather than return something the
        reference would not, a centred sampler step on 3 samples is refused (pad the batch to 4 or split it)."""
        if B == 3 and self.center:
            raise MmdmError("batch of exactly 3 samples: the reference's center_motion takes torch.cross over the batch "
                            "axis there (legacy dim rule) and mixes samples; this engine does not reproduce that. "
                            "Use another batch size (e.g. pad to 4).")

    def _check_state(self, x, name):
        if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
            raise TypeError(f"{name}: expected a contiguous fp32 CUDA tensor")

    # -- sampler ----------------------------------------------------------------------------------
    def set_condition(self, cond):
        cond = _f32c(cond, self.device)
        assert cond.shape == (self.hd.B, 8 * TEXT_DIM), f"cond must be ({self.hd.B}, {8 * TEXT_DIM})"
        self.hd.call("mmdm_set_condition", ptr(cond), self.hd.stream())

    def set_state(self, x1, x2, step_index):
        x1, x2 = _f32c(x1, self.device), _f32c(x2, self.device)
        for x in (x1, x2):
            if tuple(x.shape) != (self.hd.B, self.hd.T, 2 * NFEATS):
                raise ValueError(f"state must be ({self.hd.B}, {self.hd.T}, {2 * NFEATS}), got {tuple(x.shape)}")
        self.hd.call("mmdm_set_state", ptr(x1), ptr(x2), int(step_index), self.hd.stream())

    def get_state(self):
        shape = (self.hd.B, self.hd.T, 2 * NFEATS)
        x1 = torch.empty(shape, device=self.device)
        x2 = torch.empty(shape, device=self.device)
        self.hd.call("mmdm_get_state", ptr(x1), ptr(x2), self.hd.stream())
        return x1, x2

    def step(self, final_out=None):
        if final_out is not None:
            self._check_state(final_out, "final_out")
        self._check_batch(self.hd.B)
        self.hd.call("mmdm_step", ptr(final_out), self.hd.stream())

    def set_sampler(self, kind, var_type="fixed_small"):
        """"ddim" (eta = 0, the reference's inference path) or "ddpm" (ancestral update): swaps the per-step coefficient
        table the fused step reads on the device; captured step graphs stay valid (same pointer)."""
        key = (kind, var_type if kind == "ddpm" else None)
        if key == (self.sampler, getattr(self, "_var_type", None)):
            return
        if kind == "ddim":
            table, self.sigma = self.sched.ddim_coefficients(), None
        elif kind == "ddpm":
            table, self.sigma = self.sched.ddpm_coefficients(var_type)
        else:
            raise ValueError(f"sampler must be 'ddim' or 'ddpm', got {kind!r}")
        with torch.cuda.device(self.device):
            self.coef.copy_(torch.from_numpy(table).to(self.device))
            torch.cuda.current_stream().synchronize()
        self.sampler, self._var_type = key

    def add_state_noise(self, noise1, noise2, scale):
        """state += scale * noise on both streams (the sigma_t z term of a DDPM step)"""
        for n in (noise1, noise2):
            self._check_state(n, "noise")
            _expect(n, (self.hd.B, self.hd.T, 2 * NFEATS), "noise")
        self.hd.call("mmdm_add_state_noise", ptr(noise1), ptr(noise2), C.c_float(float(scale)), self.hd.stream())

    def sample_ddpm(self, cond, noise, step_noise=None, generator=None, var_type="fixed_small", mask=None):
        """whole ancestral (DDPM) loop: the fused step with posterior-mean coefficients + sigma_t z after every step but
        the last. step_noise: optional (nsteps - 1, B, T, 524) tensor of the z's (index i is used after DDIM-table
        index nsteps - 1 - i); otherwise drawn with torch.randn(generator=...). Both state streams receive the same z,
        as they start from the same noise. Returns the last step's prediction like sample()."""
        B, T = noise.shape[0], noise.shape[1]
        _expect(noise, (B, T, 2 * NFEATS), "noise")
        _expect(cond, (B, 8 * TEXT_DIM), "cond")
        self._check_batch(B)
        self.bind(B, T)
        prev = (self.sampler, getattr(self, "_var_type", None))
        self.set_sampler("ddpm", var_type)
        try:
            noise = _f32c(noise, self.device)
            self.hd.set_key_mask(mask, T)
            self.set_condition(cond)
            self.set_state(noise, noise, self.nsteps - 1)
            out = torch.empty(B, T, 2 * NFEATS, device=self.device)
            for i in range(self.nsteps - 1, -1, -1):
                self.step(out if i == 0 else None)
                if i > 0:
                    k = self.nsteps - 1 - i
                    z = (_f32c(step_noise[k], self.device) if step_noise is not None else
                         torch.randn(B, T, 2 * NFEATS, device=self.device, generator=generator))
                    self.add_state_noise(z, z, float(self.sigma[i]))
        finally:
            self.set_sampler(prev[0], prev[1] or "fixed_small")
        return out

    def sample(self, cond, noise, out=None, store_influence=False, use_graph=True, store_outputs=False, mask=None):
        """whole DDIM loop; cond (B,6144), noise (B,T,524) device fp32 -> (B,T,524) real-space prediction.
        store_influence: also return the per-step expanded weights, 2 x (nsteps, 2B, T, 262).
        store_outputs:   also return out1 / out2 / out_influenced of every step, 3 x (nsteps, 2B, T, 524)."""
        B, T = noise.shape[0], noise.shape[1]
        _expect(noise, (B, T, 2 * NFEATS), "noise")
        _expect(cond, (B, 8 * TEXT_DIM), "cond")
        self._check_batch(B)
        self.bind(B, T)
        cond = _f32c(cond, self.device)
        noise = _f32c(noise, self.device)
        if out is not None:
            self._check_state(out, "out")
            _expect(out, (B, T, 2 * NFEATS), "out")
        if out is None:
            out = torch.empty(B, T, 2 * NFEATS, device=self.device)
        h1 = h2 = o1 = o2 = om = None
        if store_influence:
            h1 = torch.empty(self.nsteps, 2 * B, T, NFEATS, device=self.device)
            h2 = torch.empty(self.nsteps, 2 * B, T, NFEATS, device=self.device)
        if store_outputs:
            o1, o2, om = [torch.empty(self.nsteps, 2 * B, T, 2 * NFEATS, device=self.device) for _ in range(3)]
        self.hd.set_key_mask(mask, T)
        self.hd.call("mmdm_sample", ptr(cond), ptr(noise), ptr(out), ptr(h1), ptr(h2), ptr(o1), ptr(o2), ptr(om),
                     int(use_graph), self.hd.stream())
        res = [out]
        if store_influence:
            res += [h1, h2]
        if store_outputs:
            res += [o1, o2, om]
        return res[0] if len(res) == 1 else tuple(res)

    # -- module-level entry points ------------------------------------------------------------------
    def mixer_forward(self, x1, x2, timesteps, cond, want_aux=False, want_influence=False, cfg_combine=False,
                      cfg_scale=None, mask=None):
        Bx, T = x1.shape[0], x1.shape[1]
        _expect(x1, (Bx, T, 2 * NFEATS), "x1")
        _expect(x2, (Bx, T, 2 * NFEATS), "x2")
        _expect(cond, (Bx, 8 * TEXT_DIM), "cond")
        _expect(timesteps, (Bx,), "timesteps")
        x1, x2, cond = _f32c(x1, self.device), _f32c(x2, self.device), _f32c(cond, self.device)
        ts = timesteps.detach().to(device=self.device, dtype=torch.int64).contiguous()
        rows = Bx // 2 if cfg_combine else Bx
        out = torch.empty(rows, T, 2 * NFEATS, device=self.device)
        o1 = torch.empty(Bx, T, 2 * NFEATS, device=self.device) if want_aux else None
        o2 = torch.empty(Bx, T, 2 * NFEATS, device=self.device) if want_aux else None
        w1 = torch.empty(Bx, T, NFEATS, device=self.device) if want_influence else None
        w2 = torch.empty(Bx, T, NFEATS, device=self.device) if want_influence else None
        self.hd.set_key_mask(mask, T)
        self.hd.call("mmdm_mixer_forward", ptr(x1), ptr(x2), ptr(ts), ptr(cond), ptr(out), ptr(o1), ptr(o2), ptr(w1),
                     ptr(w2), Bx, int(cfg_combine), C.c_float(self.cfg_scale if cfg_scale is None else cfg_scale),
                     self.hd.stream())
        return out, o1, o2, w1, w2

    def cfg_forward(self, x, x2, timesteps, cond, scale, store_influence=False, owner=None, mask=None):
        """ClassifierFreeSampleModelX2.forward: doubled batch, zero condition for the second half, guidance."""
        xx = torch.cat([x, x], 0)
        xx2 = torch.cat([x2, x2], 0)
        tt = torch.cat([timesteps, timesteps], 0)
        cc = torch.cat([cond, torch.zeros_like(cond)], 0)
        out, _, _, w1, w2 = self.mixer_forward(xx, xx2, tt, cc, want_influence=store_influence, cfg_combine=True,
                                               cfg_scale=scale, mask=mask)   # (mask rows serve both halves: bx % B)
        if store_influence and owner is not None:
            from .modules import history_view
            owner.history_influence_i1.append(history_view(w1, self.mixing_mode))
            owner.history_influence_i2.append(history_view(w2, self.mixing_mode))
        return out

    def denoiser_forward(self, which, x, timesteps, cond, mask=None):
        return _denoiser_call(self.hd, which, x, timesteps, cond, mask)


class StandaloneDenoiser:
    """engine holding a single in2INDenoiser (direct module.forward calls, BASELINE config 3)"""

    def __init__(self, den, x):
        device = canonical_device(x.device)
        pk = _Packer(device)
        self.kind = NET_INTERACTION if den.mode == "interaction" else NET_INDIVIDUAL
        nets = [_absent_net(0), _absent_net(1), _absent_net(2)]
        nets[self.kind] = pk.pack_net(self.kind, den, den.blocks, True)
        cfg = Config()
        cfg.nfeats, cfg.kpad, cfg.text_dim, cfg.nsteps = NFEATS, KPAD, TEXT_DIM, 1
        cfg.mixing_mode, cfg.align, cfg.center, cfg.has_force = 4, 0, 0, 0
        cfg.cfg_scale = 1.0
        tmap = (C.c_int32 * 1)(0)
        pk.keep.append(tmap)
        cfg.timestep_map = C.cast(tmap, C.POINTER(C.c_int32))
        dummy = pk.f32(torch.zeros(8))
        cfg.coef = pk.p(dummy)
        self.hd = _Handle(cfg, nets, pk.keep, device)
        self.device = device
        self.fingerprint = weights_fingerprint(den)
        self.rows = 0
        self.T = 0

    def matches(self, den, x):
        return canonical_device(x.device) == self.device and weights_fingerprint(den) == self.fingerprint

    def forward(self, x, timesteps, cond, mask=None):
        Bx, T = x.shape[0], x.shape[1]
        if Bx > self.rows or T != self.T:
            self.hd.bind((Bx + 1) // 2, T)
            self.rows, self.T = 2 * ((Bx + 1) // 2), T
        return _denoiser_call(self.hd, self.kind, x, timesteps, cond, mask)
