"""TEST INFRASTRUCTURE ONLY — CPU restatement (torch fp32, functional, state-dict driven) of the MixerMDM
denoising loop of pabloruizponce/MixerMDM. It is the checker for the CUDA path (tests/, __graft_entry__.smoke,
bench.py's cpu_baseline / --impl reference legs) and is never imported by the product package.

Pinned against the unmodified reference: tests/test_oracle_cpu.py runs both on identical weights and
inputs (reference imported from /root/reference through oracle/ref_harness.py) and tests/golden/*.npz hold
outputs of the reference itself (oracle/make_goldens.py). The reference ships no tests or golden vectors of its
own, so those two are the pin.

Every function cites the reference lines it restates (paths relative to /root/reference/src).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

NF = 262


# ------------------------------------------------------------------------------------------------
# transformer pieces
# ------------------------------------------------------------------------------------------------
def ada_ln(sd, pre, h, emb):
    """models/utils/layers.py:15-25 — LN(eps=1e-6, no affine) * (1 + scale) + shift, [scale|shift] = Linear(SiLU(emb))"""
    mod = F.linear(F.silu(emb), sd[pre + "emb_layers.1.weight"], sd[pre + "emb_layers.1.bias"])
    scale, shift = mod.chunk(2, dim=-1)
    return F.layer_norm(h, (h.shape[-1],), eps=1e-6) * (1 + scale[:, None]) + shift[:, None]


def mha_zero_attn(sd, pre, q_in, kv_in, heads, pad=None):
    """nn.MultiheadAttention(batch_first, add_zero_attn=True) (models/utils/layers.py:33-45,74-88;
    torch F.multi_head_attention_forward): packed in-projection, one extra key/value row of zeros appended
    AFTER the projection, logits scaled by 1/sqrt(hd), out-projection. pad: optional key_padding_mask (B, Tk) bool,
    True = ignore the key; torch pads it with False for the zero key, so a fully padded row attends to that key."""
    w, b = sd[pre + "in_proj_weight"], sd[pre + "in_proj_bias"]
    D = q_in.shape[-1]
    q = F.linear(q_in, w[:D], b[:D])
    k = F.linear(kv_in, w[D:2 * D], b[D:2 * D])
    v = F.linear(kv_in, w[2 * D:], b[2 * D:])
    B, Tq, _ = q.shape
    Tk = k.shape[1]
    hd = D // heads
    q = q.view(B, Tq, heads, hd).transpose(1, 2)
    k = k.view(B, Tk, heads, hd).transpose(1, 2)
    v = v.view(B, Tk, heads, hd).transpose(1, 2)
    zeros = torch.zeros(B, heads, 1, hd, dtype=q.dtype)
    k = torch.cat([k, zeros], dim=2)
    v = torch.cat([v, zeros], dim=2)
    logits = q @ k.transpose(-1, -2) / math.sqrt(hd)
    if pad is not None:
        padz = torch.cat([pad, torch.zeros(B, 1, dtype=torch.bool)], dim=1)
        logits = logits.masked_fill(padz[:, None, None, :], float("-inf"))
    att = torch.softmax(logits, dim=-1) @ v
    att = att.transpose(1, 2).reshape(B, Tq, D)
    return F.linear(att, sd[pre + "out_proj.weight"], sd[pre + "out_proj.bias"])


def ffn(sd, pre, h, emb):
    """models/utils/layers.py:100-106 — linear2(GELU_erf(linear1(AdaLN(h)))) (dropout inactive in eval)"""
    x = ada_ln(sd, pre + "norm.", h, emb)
    x = F.gelu(F.linear(x, sd[pre + "linear1.weight"], sd[pre + "linear1.bias"]))
    return F.linear(x, sd[pre + "linear2.weight"], sd[pre + "linear2.bias"])


def block(sd, pre, x, y, emb_sa, emb_ca, emb_ffn, heads, cross, pad=None):
    """models/utils/blocks.py:49-63 and models/utils/influence.py:36-48 (they differ only in which embedding
    feeds the FFN's AdaLN)"""
    xs = ada_ln(sd, pre + "sa_block.norm.", x, emb_sa)
    h = x + mha_zero_attn(sd, pre + "sa_block.attention.", xs, xs, heads, pad)
    if cross:
        qn = ada_ln(sd, pre + "ca_block.norm.", h, emb_ca)
        kn = ada_ln(sd, pre + "ca_block.xf_norm.", y, emb_ca)
        h = h + mha_zero_attn(sd, pre + "ca_block.attention.", qn, kn, heads, pad)
    return h + ffn(sd, pre + "ffn.", h, emb_ffn)


def time_embed(sd, pre, t):
    """models/utils/utils.py:41-55 — MLP over rows of the sinusoidal table picked by the ORIGINAL-scale timestep"""
    e = sd[pre + "sequence_pos_encoder.pe"][t]
    e = F.linear(e, sd[pre + "embed_timestep.time_embed.0.weight"], sd[pre + "embed_timestep.time_embed.0.bias"])
    return F.linear(F.silu(e), sd[pre + "embed_timestep.time_embed.2.weight"], sd[pre + "embed_timestep.time_embed.2.bias"])


def n_layers(sd, pre):
    return 1 + max(int(k[len(pre + "blocks."):].split(".")[0]) for k in sd if k.startswith(pre + "blocks."))


def _pad_of(mask):
    """models/in2in.py:410-411,433-435 / influence.py:107-116: mask (B, T, P) float, frame kept where mask[..., 0] > 0.5"""
    return None if mask is None else ~(mask[..., 0] > 0.5)


def denoiser(sd, pre, mode, x, t, cond, heads, mask=None):
    """models/in2in.py:401-462 — both streams read the PREVIOUS layer's states; weights shared between streams"""
    T = x.shape[1]
    pad = _pad_of(mask)
    te = time_embed(sd, pre, t)
    text = lambda c: F.linear(c, sd[pre + "text_embed.weight"], sd[pre + "text_embed.bias"])  # noqa: E731
    pe = sd[pre + "sequence_pos_encoder.pe"][:T]
    embed = lambda m: F.linear(m, sd[pre + "motion_embed.weight"], sd[pre + "motion_embed.bias"]) + pe  # noqa: E731
    L = n_layers(sd, pre)
    if mode == "individual":
        e1 = te + text(cond[:, :768])
        h = embed(x[..., :NF])
        for l in range(L):
            h = block(sd, f"{pre}blocks.{l}.", h, None, e1, None, e1, heads, cross=False, pad=pad)
        return F.linear(h, sd[pre + "out.linear.weight"], sd[pre + "out.linear.bias"])
    e = te + text(cond[:, :768])
    e1 = te + text(cond[:, 768:1536])
    e2 = te + text(cond[:, 1536:2304])
    ha, hb = embed(x[..., :NF]), embed(x[..., NF:])
    for l in range(L):
        p = f"{pre}blocks.{l}."
        na = block(sd, p, ha, hb, e1, e, e1, heads, cross=True, pad=pad)
        nb = block(sd, p, hb, ha, e2, e, e2, heads, cross=True, pad=pad)
        ha, hb = na, nb
    out = lambda h: F.linear(h, sd[pre + "out.linear.weight"], sd[pre + "out.linear.bias"])  # noqa: E731
    return torch.cat([out(ha), out(hb)], dim=-1)


def influence(sd, pre, m_i, m_I, cond_i, cond_I, heads, mode, mask=None):
    """models/utils/influence.py:93-126 (the time pooling of modes 1 / 3 runs over padded frames too)"""
    h = m_i
    pad = _pad_of(mask)
    for l in range(n_layers(sd, pre)):
        h = block(sd, f"{pre}blocks.{l}.", h, m_I, cond_i, cond_I, cond_I, heads, cross=True, pad=pad)
    if mode in (1, 3):
        h = h.mean(dim=1)
    return torch.sigmoid(F.linear(h, sd[pre + "out.weight"], sd[pre + "out.bias"]))


# ------------------------------------------------------------------------------------------------
# motion-space math
# ------------------------------------------------------------------------------------------------
def _cross(a, b):
    return torch.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                        a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                        a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], dim=-1)


def qrot(q, v):
    """utils/quaternion.py:54-73"""
    qv = q[..., 1:]
    uv = _cross(qv, v)
    uuv = _cross(qv, uv)
    return v + 2 * (q[..., :1] * uv + uuv)


def qbetween(v0, v1):
    """utils/quaternion.py:386-396 (+ qnormalize :28-30)"""
    v = _cross(v0, v1)
    w = torch.sqrt((v0 ** 2).sum(-1, keepdim=True) * (v1 ** 2).sum(-1, keepdim=True)) + (v0 * v1).sum(-1, keepdim=True) + 1e-8
    q = torch.cat([w, v], dim=-1)
    return q / torch.sqrt((q ** 2).sum(-1, keepdim=True))


def _sin_half_over_angle(half, ang):
    small = ang.abs() < 1e-6
    safe = torch.where(small, torch.ones_like(ang), ang)
    return torch.where(small, 0.5 - (ang * ang) / 48, torch.sin(half) / safe)


def rot6d_roundtrip(d6):
    """ih_to_smpl then smpl_to_ih on the rotation block (utils/alignment.py:11-67):
    rotation_6d_to_matrix (utils/rotation_conversions.py:513-535, interleaved layout, F.normalize eps 1e-12) ->
    matrix_to_quaternion (:98-120) -> quaternion_to_axis_angle (:482-510) -> *-1 -> *-1 ->
    axis_angle_to_quaternion (:450-479) -> quaternion_to_matrix (:38-66) -> matrix_to_rotation_6d (:541-571)."""
    a1, a2 = d6[..., [0, 2, 4]], d6[..., [1, 3, 5]]
    b1 = a1 / a1.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    b2 = a2 - (b1 * a2).sum(-1, keepdim=True) * b1
    b2 = b2 / b2.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    b3 = _cross(b1, b2)
    m00, m11, m22 = b1[..., 0], b2[..., 1], b3[..., 2]
    sp = lambda x: torch.where(x > 0, torch.sqrt(x.clamp_min(0)), torch.zeros_like(x))  # noqa: E731
    cs = lambda a, b: torch.where((a < 0) != (b < 0), -a, a)  # noqa: E731
    qw = 0.5 * sp(1 + m00 + m11 + m22)
    qx = cs(0.5 * sp(1 + m00 - m11 - m22), b3[..., 1] - b2[..., 2])
    qy = cs(0.5 * sp(1 - m00 + m11 - m22), b1[..., 2] - b3[..., 0])
    qz = cs(0.5 * sp(1 - m00 - m11 + m22), b2[..., 0] - b1[..., 1])
    qv = torch.stack([qx, qy, qz], dim=-1)
    nv = qv.norm(dim=-1, keepdim=True)
    half = torch.atan2(nv, qw[..., None])
    ang = 2 * half
    aa = qv / _sin_half_over_angle(half, ang)
    aa = (aa * -1) * -1
    ang2 = aa.norm(dim=-1, keepdim=True)
    half2 = 0.5 * ang2
    s2 = _sin_half_over_angle(half2, ang2)
    r = torch.cos(half2)[..., 0]
    i, j, k = (aa * s2).unbind(-1)
    two_s = 2.0 / (r * r + i * i + j * j + k * k)
    r00, r01, r02 = 1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r)
    r10, r11, r12 = two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r)
    return torch.stack([r00, r10, r01, r11, r02, r12], dim=-1)


def roundtrip_motion(m):
    """(.., 262) -> (.., 262) with the 21x6 rotation block passed through the SMPL round trip; fc kept"""
    rot = rot6d_roundtrip(m[..., 132:258].reshape(*m.shape[:-1], 21, 6)).reshape(*m.shape[:-1], 126)
    return torch.cat([m[..., :132], rot, m[..., 258:]], dim=-1)


def align_pair(inter, indiv):
    """ih_to_smpl x2 -> align_motions(inter, indiv) -> smpl_to_ih x2 (models/mixermdm.py:702-712;
    utils/alignment.py:69-158 with mask=None). Returns (inter', indiv'): inter only gets the rotation round
    trip; indiv is translated/rotated onto inter's root trajectory and LOSES its foot contacts (201-d bug)."""
    B, T = inter.shape[:2]
    inter_o = roundtrip_motion(inter)
    ind = roundtrip_motion(indiv)
    p1 = inter[..., :66].reshape(B, T, 22, 3)
    p2 = ind[..., :66].reshape(B, T, 22, 3)
    v2 = ind[..., 66:132].reshape(B, T, 22, 3)
    p2 = p2 + (p1[:, 0, 0] - p2[:, 0, 0])[:, None, None]
    a = p1[:, -1, 0] - p1[:, 0, 0]
    b = p2[:, -1, 0] - p2[:, 0, 0]
    a = a * torch.tensor([1.0, 0.0, 1.0])
    b = b * torch.tensor([1.0, 0.0, 1.0])
    a = a / torch.sqrt((a ** 2).sum(1, keepdim=True) + 1e-8)
    b = b / torch.sqrt((b ** 2).sum(1, keepdim=True) + 1e-8)
    q = qbetween(b, a)[:, None, None].expand(B, T, 22, 4)
    p2 = qrot(q, p2)
    p2 = p2 + (p1[:, 0, 0] - p2[:, 0, 0])[:, None, None]
    v2 = qrot(q, v2)
    ind_o = torch.cat([p2.reshape(B, T, 66), v2.reshape(B, T, 66), ind[..., 132:258], torch.zeros(B, T, 4)], dim=-1)
    return inter_o, ind_o


def center_person(m):
    """ih_to_smpl -> center_motion -> smpl_to_ih (models/utils/gaussian_diffusion.py:2043-2049;
    utils/alignment.py:161-222): floor at min-y, frame-0 root to the xz origin, frame-0 hips facing +Z
    (FACE_JOINT_INDX[:2] = (2, 1), utils/paramUtil.py:89); foot contacts dropped -> 0."""
    B, T = m.shape[:2]
    mm = roundtrip_motion(m)
    p = mm[..., :66].reshape(B, T, 22, 3).clone()
    v = mm[..., 66:132].reshape(B, T, 22, 3)
    floor = p[..., 1].reshape(B, -1).min(dim=1).values
    p[..., 1] = p[..., 1] - floor[:, None, None]
    root0 = p[:, 0]
    p = p - (root0[:, 0] * torch.tensor([1.0, 0.0, 1.0]))[:, None, None]
    across = root0[:, 2] - root0[:, 1]
    across = across / torch.sqrt((across ** 2).sum(-1))[:, None]
    fwd = _cross(torch.tensor([0.0, 1.0, 0.0]).expand(B, 3), across)
    fwd = fwd / torch.sqrt((fwd ** 2).sum(-1))[:, None]
    q = qbetween(fwd, torch.tensor([0.0, 0.0, 1.0]).expand(B, 3))[:, None, None].expand(B, T, 22, 4)
    p, v = qrot(q, p), qrot(q, v)
    return torch.cat([p.reshape(B, T, 66), v.reshape(B, T, 66), mm[..., 132:258], torch.zeros(B, T, 4)], dim=-1)


def expand_weights(w, T, mode):
    """models/mixermdm.py:739-786 — 1 / T / 23 / Tx23 weights -> (.., T, 262)"""
    if mode in (1, 3):
        w = w[:, None].expand(-1, T, -1)
    if mode in (1, 2):
        return w
    j = w[..., :22].repeat_interleave(3, dim=-1)
    r = w[..., :21].repeat_interleave(6, dim=-1)
    f = w[..., 22:].expand(*w.shape[:-1], 4)
    return torch.cat([j, j, r, f], dim=-1)


# ------------------------------------------------------------------------------------------------
# the step
# ------------------------------------------------------------------------------------------------
class Oracle:
    def __init__(self, sd, dims, stats, mixing_mode=4, align=True, center=True, cfg_scale=3.5, force=None,
                 sched=None):
        """sd: Mixer state dict (keys as in the reference: denoiser1.*, denoiser2.*, influence.*, motion_embed.*,
        text_embed.*, embed_timestep.*, sequence_pos_encoder.pe); stats: ((mean1,std1),(mean2,std2)) =
        (HumanML3D / individual, InterHuman / interaction)."""
        self.sd = {k: v.detach().float().cpu() for k, v in sd.items()}
        self.h_den, self.h_mix = dims["den"]["num_heads"], dims["mix"]["n_heads"]
        (m1, s1), (m2, s2) = stats
        self.m1, self.s1, self.m2, self.s2 = [torch.as_tensor(np.asarray(a), dtype=torch.float32) for a in (m1, s1, m2, s2)]
        self.mode, self.align, self.center, self.s, self.force = mixing_mode, align, center, cfg_scale, force
        self.sched = sched

    def mixer(self, x1, x2, t, cond, detail=False, mask=None):
        """models/mixermdm.py:660-811 (eval_intermediate): returns out_influenced (Be,T,524). mask: (Be, T, 2) float
        frame mask handed to the three networks; the reference's align_trajectories cannot take it (mask.squeeze().sum(dim=1)
        is (B, 2) for this mask and is then used as one index per batch row, utils/alignment.py:86-89: shape error for
        B != 2), so a mask requires align=False here as it does there."""
        assert mask is None or not self.align
        sd = self.sd
        T = x1.shape[1]
        te = time_embed(sd, "", t)
        tx = lambda c: F.linear(c, sd["text_embed.weight"], sd["text_embed.bias"])  # noqa: E731
        cI, ci1, ci2 = te + tx(cond[:, 3840:4608]), te + tx(cond[:, 4608:5376]), te + tx(cond[:, 5376:6144])
        o11 = denoiser(sd, "denoiser1.", "individual", x1[..., :NF], t, cond[:, 2304:3072], self.h_den, mask)
        o12 = denoiser(sd, "denoiser1.", "individual", x1[..., NF:], t, cond[:, 3072:3840], self.h_den, mask)
        o2 = denoiser(sd, "denoiser2.", "interaction", x2, t, cond[:, :2304], self.h_den, mask)
        o11, o12 = o11 * self.s1 + self.m1, o12 * self.s1 + self.m1
        o21, o22 = o2[..., :NF] * self.s2 + self.m2, o2[..., NF:] * self.s2 + self.m2
        raw = (torch.cat([o11, o12], -1), torch.cat([o21, o22], -1))
        if self.align:
            o21, o11 = align_pair(o21, o11)
            o22, o12 = align_pair(o22, o12)
        pe = sd["sequence_pos_encoder.pe"][:T]
        emb = lambda m: F.linear(m, sd["motion_embed.weight"], sd["motion_embed.bias"]) + pe  # noqa: E731
        w1 = influence(sd, "influence.", emb(o11), emb(o21), ci1, cI, self.h_mix, self.mode, mask)
        w2 = influence(sd, "influence.", emb(o12), emb(o22), ci2, cI, self.h_mix, self.mode, mask)
        w1, w2 = expand_weights(w1, T, self.mode), expand_weights(w2, T, self.mode)
        if self.force is not None:
            w1, w2 = torch.ones_like(w1) * self.force, torch.ones_like(w2) * self.force
        out = torch.cat([o21 + w1 * (o11 - o21), o22 + w2 * (o12 - o22)], dim=-1)
        if detail:
            return dict(out=out, out1=torch.cat([o11, o12], -1), out2=torch.cat([o21, o22], -1),
                        w1=w1.expand(-1, T, NF) if w1.shape[-1] == 1 else w1,
                        w2=w2.expand(-1, T, NF) if w2.shape[-1] == 1 else w2, raw1=raw[0], raw2=raw[1])
        return out

    def guided(self, x1, x2, t, cond, mask=None):
        """models/utils/cfg_sampler.py:38-56"""
        B = x1.shape[0]
        out = self.mixer(torch.cat([x1, x1]), torch.cat([x2, x2]), torch.cat([t, t]),
                         torch.cat([cond, torch.zeros_like(cond)]),
                         mask=None if mask is None else torch.cat([mask, mask]))
        return self.s * out[:B] + (1 - self.s) * out[B:]

    def process_xstart(self, g, last):
        """models/utils/gaussian_diffusion.py:2031-2062 — returns (pred_xstart [stream 1], pred_xstart2 [stream 2])"""
        if last:            # t == 0: neither centring result nor normalisation is applied
            return g.clone(), g.clone()
        a, b = g[..., :NF], g[..., NF:]
        if self.center:
            a, b = center_person(a), center_person(b)
        x1 = torch.cat([(a - self.m1) / self.s1, (b - self.m1) / self.s1], dim=-1)
        x2 = torch.cat([(g[..., :NF] - self.m2) / self.s2, (g[..., NF:] - self.m2) / self.s2], dim=-1)
        return x1, x2

    def ddim_step(self, x1, x2, i, cond, mask=None):
        """models/utils/gaussian_diffusion.py:1902-1965 with eta = 0; i = index into the re-spaced schedule"""
        B = x1.shape[0]
        t = torch.full((B,), int(self.sched.timestep_map[i]), dtype=torch.long)
        g = self.guided(x1, x2, t, cond, mask)
        p1, p2 = self.process_xstart(g, last=(i == 0))
        c = self.sched.ddim_coefficients()[i]
        cr, crm1, ca, ce = [torch.tensor(float(v), dtype=torch.float32) for v in c]
        n1 = p1 * ca + ce * ((cr * x1 - p1) / crm1)
        n2 = p2 * ca + ce * ((cr * x2 - p2) / crm1)
        return dict(sample=n1, sample2=n2, pred_xstart=p1, pred_xstart2=p2, guided=g)

    def ddpm_step(self, x1, x2, i, cond, z):
        """Ancestral update of GaussianDiffusion.p_sample (models/utils/gaussian_diffusion.py:632-692) over the two-stream
        prediction: mean = coef1 x0 + coef2 x_t (q_posterior_mean_variance :452-478), + exp(0.5 logvar) z for t != 0,
        FIXED_SMALL variance (:520-538). The reference's MixerDiffusion cannot execute p_sample (SURVEY 8a note 2), so
        this restates the inherited formula on MixerDiffusion.p_mean_variance's two pred_xstart streams."""
        B = x1.shape[0]
        t = torch.full((B,), int(self.sched.timestep_map[i]), dtype=torch.long)
        g = self.guided(x1, x2, t, cond)
        p1, p2 = self.process_xstart(g, last=(i == 0))
        c1 = torch.tensor(float(self.sched.posterior_mean_coef1[i]), dtype=torch.float32)
        c2 = torch.tensor(float(self.sched.posterior_mean_coef2[i]), dtype=torch.float32)
        sg = torch.tensor(float(np.exp(0.5 * self.sched.posterior_log_variance_clipped[i])), dtype=torch.float32)
        nz = 0.0 if i == 0 else 1.0
        return dict(sample=c1 * p1 + c2 * x1 + nz * sg * z, sample2=c1 * p2 + c2 * x2 + nz * sg * z,
                    pred_xstart=p1, pred_xstart2=p2, guided=g)

    def sample(self, noise, cond, steps=None, trace=None):
        """models/utils/gaussian_diffusion.py:1769-1899 — both streams start from the same noise"""
        x1, x2 = noise.clone(), noise.clone()
        n = self.sched.num_timesteps
        out = None
        for i in list(range(n))[::-1][:steps]:
            out = self.ddim_step(x1, x2, i, cond)
            if trace is not None:
                trace.append(out)
            x1, x2 = out["sample"], out["sample2"]
        return out["pred_xstart2"]
