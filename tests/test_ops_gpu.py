"""Operator parity on the GPU: each C-ABI operator against a plain torch fp32 evaluation of the same op on
the same bf16-rounded operands (tolerances are bf16-output rounding + fp32 accumulation order)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from mixermdm_b200 import ops as o
    from mixermdm_b200 import _lib
    _lib.load()
    return o


def _rand(shape, seed, scale=1.0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).cuda()


def _relerr(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 128), (300, 1024, 1024), (1200, 3072, 1024), (4864, 1024, 256),
                                   (1000, 262, 1024), (777, 1536, 512), (600, 512, 264), (128, 2048, 2048)])
def test_linear_bf16_bias(ops, M, N, K):
    from mixermdm_b200._lib import EPI_BF16
    a = _rand((M, K), 1, dtype=torch.bfloat16)
    w = _rand((N, K), 2, 1.0 / math.sqrt(K), dtype=torch.bfloat16)
    b = _rand((N,), 3)
    out = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
    ops.linear_bf16(a, w, mode=EPI_BF16, out=out, bias=b)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + b
    assert _relerr(out, ref) < 6e-3
    assert (out.float() - ref).abs().max().item() < 0.06


# (900, 512, 1024): fewer tiles than CTA pairs. (4864, 512, 1024): 76 / 38 pair tiles on 74 pairs -> the last wave is cut
# into 256 x 64 quarter tiles (FFN-1: 2 tiles -> 8 quarters; FFN-2 has no short last wave: 38 tiles).
# (4864, 1024, 1024): both GEMMs end in quarter tiles (76 tiles each).
@pytest.mark.parametrize("M,D,F", [(900, 512, 1024), (4864, 512, 1024), (4864, 1024, 1024)])
def test_linear_gelu_resid_f32(ops, M, D, F):
    from mixermdm_b200._lib import EPI_GELU_BF16, EPI_RESID, EPI_F32
    x = _rand((M, D), 4, dtype=torch.bfloat16)
    w1 = _rand((F, D), 5, 1 / math.sqrt(D), dtype=torch.bfloat16)
    b1 = _rand((F,), 6)
    w2 = _rand((D, F), 7, 1 / math.sqrt(F), dtype=torch.bfloat16)
    b2 = _rand((D,), 8)
    h = _rand((M, D), 9)
    mid = torch.empty(M, F, dtype=torch.bfloat16, device="cuda")
    ops.linear_bf16(x, w1, mode=EPI_GELU_BF16, out=mid, bias=b1)
    ref_mid = torch.nn.functional.gelu(x.float() @ w1.float().t() + b1)
    assert _relerr(mid, ref_mid) < 6e-3
    h_out = h.clone()
    h_bf = torch.empty(M, D, dtype=torch.bfloat16, device="cuda")
    ops.linear_bf16(mid, w2, mode=EPI_RESID, out=h_out, bias=b2, resid=h_out, out_bf16=h_bf)
    ref = h + mid.float() @ w2.float().t() + b2
    torch.cuda.synchronize()
    assert _relerr(h_out, ref) < 1e-5
    assert _relerr(h_bf, ref) < 4e-3
    o32 = torch.empty(M, F, dtype=torch.float32, device="cuda")
    ops.linear_bf16(x, w1, mode=EPI_F32, out=o32, bias=b1)
    assert _relerr(o32, x.float() @ w1.float().t() + b1) < 1e-5


def _ln_ref(h, scale, shift, T):
    """AdaLN body in fp32 (reference models/utils/layers.py:15-25), scale / shift rows per sequence s = row // T"""
    seq = torch.arange(h.shape[0], device=h.device) // T
    xn = torch.nn.functional.layer_norm(h, (h.shape[1],), eps=1e-6)
    return xn * (1 + scale[seq]) + shift[seq]


# LN fold: a residual GEMM that also emits bf16(h (1 + scale)) + row sums, then the consuming GEMM that finishes the
# normalisation in its epilogue, against LayerNorm + modulation + Linear in fp32. (8, 300, 1024, .): CTA-pair kernel with
# quarter tiles in the producer (2400 rows: 10 pair tiles x 4 = 40 < 74 -> none; 4864-row case below has them);
# (3, 24, 128, .): single-CTA 128-wide tiles, several sequences per warp.
@pytest.mark.parametrize("S,T,D,N2", [(4, 37, 256, 768), (8, 300, 1024, 3072), (16, 304, 1024, 2048), (3, 24, 128, 384),
                                      (16, 300, 512, 1024)])
@pytest.mark.parametrize("dual", [False, True])
def test_ln_fold_resid_producer_and_consumers(ops, S, T, D, N2, dual):
    from mixermdm_b200._lib import EPI_RESID_LN, EPI_BF16_LN, EPI_GELU_BF16_LN
    M = S * T
    a = _rand((M, D), 31, dtype=torch.bfloat16)
    w = _rand((D, D), 32, 1 / math.sqrt(D), dtype=torch.bfloat16)
    b = _rand((D,), 33)
    h = _rand((M, D), 34) * 2.0 + 0.4 + 0.3 * _rand((M, 1), 35)   # rows with a mean of a few tenths of their spread
    mods = _rand((S, 4 * D), 36, 0.3)   # [scale0 | shift0 | scale1 | shift1]
    sc0, sh0, sc1, sh1 = mods[:, :D], mods[:, D:2 * D], mods[:, 2 * D:3 * D], mods[:, 3 * D:]
    stats = torch.full((D // 64, M, 2), float("nan"), device="cuda")
    hg0 = torch.empty(M, D, dtype=torch.bfloat16, device="cuda")
    hg1 = torch.empty(M, D, dtype=torch.bfloat16, device="cuda") if dual else None
    h_out = h.clone()
    # finished per-row values (-mean, 1 / rstd as bf16 pairs; rstd), written by the row pass behind the GEMM
    rowa = torch.zeros(M, 4, dtype=torch.int32, device="cuda")
    lp = dict(scale0=sc0, out0=hg0, stats=stats, T=T, rowa=rowa)
    # per-row offsets (dual cases): operand and row sums are those of h - m_r; ANY m_r must give the same consumer output,
    # a good estimate of the row mean keeps the operand's bf16 rounding free of it. Here: 0.8 x the mean + a constant.
    moff = None
    if dual:
        moff = 0.8 * (h + a.float() @ w.float().t() + b).mean(1) + 0.05
        moff_out = torch.zeros(M, device="cuda")
        lp.update(scale1=sc1, out1=hg1, moff_in=moff.contiguous(), moff_out=moff_out)
    ops.linear_bf16(a, w, mode=EPI_RESID_LN, out=h_out, bias=b, resid=h_out, ln_produce=lp)
    torch.cuda.synchronize()
    ref_h = h + a.float() @ w.float().t() + b
    assert _relerr(h_out, ref_h) < 1e-5
    seq = torch.arange(M, device="cuda") // T
    ref_c = ref_h - moff[:, None] if moff is not None else ref_h   # what operand and row sums describe
    assert _relerr(hg0, ref_c * (1 + sc0[seq])) < 4e-3
    ref_stats = torch.stack([ref_c.view(M, D // 64, 64).sum(-1).t(), (ref_c ** 2).view(M, D // 64, 64).sum(-1).t()], -1)
    assert (stats - ref_stats).abs().max().item() < 2e-3 and _relerr(stats[..., 1], ref_stats[..., 1]) < 1e-5
    if moff is not None:
        assert (moff_out - ref_h.mean(1)).abs().max().item() < 1e-5
    ref_rstd = (ref_h.var(1, unbiased=False) + 1e-6).rsqrt()
    assert _relerr(rowa[:, 3].contiguous().view(torch.float32), ref_rstd) < 1e-5
    def pair(w):   # (low, high) bf16 halves of a 32-bit word as fp32
        return ((w & 0xFFFF) << 16).view(torch.float32), (w & -65536).view(torch.float32)
    mh, _ = pair(rowa[:, 0].contiguous())
    ml, sh = pair(rowa[:, 1].contiguous())
    _, sl = pair(rowa[:, 2].contiguous())
    assert _relerr(-(mh + ml), ref_c.mean(1)) < 1e-4 and _relerr(sh + sl, 1 / ref_rstd) < 1e-4
    w2 = _rand((N2, D), 37, 1 / math.sqrt(D), dtype=torch.bfloat16)
    c2 = _rand((N2,), 38)
    cases = [(hg0, sc0, sh0)] + ([(hg1, sc1, sh1)] if dual else [])
    for hg, sc, sh in cases:
        gb = torch.cat([(1 + sc) @ w2.float().t(), sh @ w2.float().t() + c2], 1).contiguous()
        lc = dict(stats=stats, g=gb[:, :N2], b=gb[:, N2:], rowa=rowa)
        ref = _ln_ref(ref_h, sc, sh, T) @ w2.float().t() + c2
        out = torch.empty(M, N2, dtype=torch.bfloat16, device="cuda")
        ops.linear_bf16(hg, w2, mode=EPI_BF16_LN, out=out, T=T, ln_consume=lc)
        assert _relerr(out, ref) < 8e-3
        ops.linear_bf16(hg, w2, mode=EPI_GELU_BF16_LN, out=out, T=T, ln_consume=lc)
        assert _relerr(out, torch.nn.functional.gelu(ref)) < 8e-3
    if D == 1024 and N2 == 3072:   # QKV form: the last third of the columns goes, transposed per head, to V^T
        H, hd, Tp = 8, 128, (T + 7) // 8 * 8
        qk = torch.zeros(M, 2 * D, dtype=torch.bfloat16, device="cuda")
        vt = torch.zeros(S * H, hd, Tp, dtype=torch.bfloat16, device="cuda")
        gb = torch.cat([(1 + sc0) @ w2.float().t(), sh0 @ w2.float().t() + c2], 1).contiguous()
        ops.linear_bf16(hg0, w2, mode=EPI_BF16_LN, out=qk, T=T, H=H, hd=hd, Tp=Tp, vt=vt, vt_col0=2 * D,
                        ln_consume=dict(stats=stats, g=gb[:, :N2], b=gb[:, N2:], rowa=rowa))
        ref = _ln_ref(ref_h, sc0, sh0, T) @ w2.float().t() + c2
        assert _relerr(qk, ref[:, :2 * D]) < 8e-3
        ref_vt = ref[:, 2 * D:].view(S, T, H, hd).permute(0, 2, 3, 1).reshape(S * H, hd, T)
        assert _relerr(vt[:, :, :T], ref_vt) < 8e-3


def test_ln_fold_quarter_tiles_and_determinism(ops):
    """4864 rows x 1024 columns = 76 pair tiles on 74 pairs: the producer's last wave runs as 256 x 64 quarter tiles, whose
    statistics slots must hold exactly what whole tiles would have written (bitwise: the row sums of a row do not depend
    on which tile shape covered it), and a rerun is bitwise identical."""
    from mixermdm_b200._lib import EPI_RESID_LN
    M, D, T = 4864, 1024, 304
    a = _rand((M, D), 41, dtype=torch.bfloat16)
    w = _rand((D, D), 42, 1 / math.sqrt(D), dtype=torch.bfloat16)
    h = _rand((M, D), 43)
    sc = _rand((M // T, D), 44, 0.3)
    res = []
    for r0 in (0, 2432):   # the second call covers rows [2432, 4864) with whole tiles only (40 pair tiles)
        stats = torch.zeros(D // 64, M, 2, device="cuda")
        hg = torch.zeros(M, D, dtype=torch.bfloat16, device="cuda")
        h_out = h.clone()
        ops.linear_bf16(a[r0:], w, mode=EPI_RESID_LN, out=h_out[r0:], resid=h_out[r0:],
                        ln_produce=dict(scale0=sc[r0 // T:], out0=hg[r0:], stats=stats[:, r0:], T=T))
        res.append((stats, hg, h_out))
    torch.cuda.synchronize()
    ref_h = h + a.float() @ w.float().t()
    assert _relerr(res[0][2], ref_h) < 1e-5
    ref_stats = torch.stack([ref_h.view(M, 16, 64).sum(-1).t(), (ref_h ** 2).view(M, 16, 64).sum(-1).t()], -1)
    assert _relerr(res[0][0], ref_stats) < 1e-5
    # rows [4608, 4864) x columns [512, 1024) are quarter tiles in the first call (tiles 74, 75 of 76) and whole tiles in
    # the second: identical bits either way
    assert torch.equal(res[0][0][:, 2432:], res[1][0][:, 2432:])
    assert torch.equal(res[0][1][2432:], res[1][1][2432:])
    assert torch.equal(res[0][2][2432:], res[1][2][2432:])


def test_ln_fold_embed_producer_with_cfg_duplicate(ops):
    from mixermdm_b200._lib import EPI_EMBED_PE_LN
    S, T, D, nf, kpad = 6, 37, 256, 262, 264     # S unique sequences, written again for the unconditional half
    Mu = S * T
    x = torch.zeros(Mu, kpad, dtype=torch.bfloat16, device="cuda")
    x[:, :nf] = _rand((Mu, nf), 50, dtype=torch.bfloat16)
    w = torch.zeros(D, kpad, dtype=torch.bfloat16, device="cuda")
    w[:, :nf] = _rand((D, nf), 51, 0.06, dtype=torch.bfloat16)
    b = _rand((D,), 52)
    pe = _rand((64, D), 53)
    sc = _rand((2 * S, D), 54, 0.3)
    out = torch.zeros(2 * Mu, D, device="cuda")
    hg = torch.zeros(2 * Mu, D, dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros(D // 64, 2 * Mu, 2, device="cuda")
    ops.linear_bf16(x, w, mode=EPI_EMBED_PE_LN, out=out, bias=b, pe=pe, T=T, dup_row_offset=Mu,
                    ln_produce=dict(scale0=sc, out0=hg, stats=stats, T=T, dup_seq=S))
    torch.cuda.synchronize()
    ref = x.float() @ w.float().t() + b + pe[torch.arange(Mu, device="cuda") % T]
    ref2 = torch.cat([ref, ref], 0)
    assert _relerr(out, ref2) < 1e-5
    seq = torch.arange(2 * Mu, device="cuda") // T
    assert _relerr(hg, ref2 * (1 + sc[seq])) < 4e-3
    ref_stats = torch.stack([ref2.view(2 * Mu, D // 64, 64).sum(-1).t(), (ref2 ** 2).view(2 * Mu, D // 64, 64).sum(-1).t()], -1)
    assert _relerr(stats, ref_stats) < 1e-5
    # rows >= `rows` only get h (the Influence network's m_I half)
    out.zero_(); hg.zero_(); stats.zero_()
    ops.linear_bf16(x, w, mode=EPI_EMBED_PE_LN, out=out[:Mu], bias=b, pe=pe, T=T,
                    ln_produce=dict(scale0=sc, out0=hg, stats=stats, T=T, rows=3 * T))
    torch.cuda.synchronize()
    assert _relerr(out[:Mu], ref) < 1e-5
    assert _relerr(hg[:3 * T], (ref * (1 + sc[seq[:Mu]]))[:3 * T]) < 4e-3
    assert hg[3 * T:].abs().max().item() == 0 and stats[:, 3 * T:].abs().max().item() == 0


def test_linear_embed_pe_and_final(ops):
    from mixermdm_b200._lib import EPI_EMBED_PE, EPI_FINAL
    S, T, D, nf, kpad = 6, 37, 256, 262, 264
    M = S * T
    x = torch.zeros(M, kpad, dtype=torch.bfloat16, device="cuda")
    x[:, :nf] = _rand((M, nf), 10, dtype=torch.bfloat16)
    w = torch.zeros(D, kpad, dtype=torch.bfloat16, device="cuda")
    w[:, :nf] = _rand((D, nf), 11, 0.06, dtype=torch.bfloat16)
    b = _rand((D,), 12)
    pe = _rand((64, D), 13)
    out = torch.zeros(2 * M, D, device="cuda")
    ops.linear_bf16(x, w, mode=EPI_EMBED_PE, out=out, bias=b, pe=pe, T=T, dup_row_offset=M)
    ref = x.float() @ w.float().t() + b + pe[:T].repeat(S, 1)
    torch.cuda.synchronize()
    assert _relerr(out[:M], ref) < 1e-5 and _relerr(out[M:], ref) < 1e-5
    # final layer: scatter to (S/2, T, 2*nf) with de-normalisation
    hfin = _rand((M, D), 14, dtype=torch.bfloat16)
    wo = _rand((nf, D), 15, 1 / 16.0, dtype=torch.bfloat16)
    bo = _rand((nf,), 16)
    std = _rand((nf,), 17).abs() + 0.5
    mean = _rand((nf,), 18)
    o = torch.zeros(S // 2, T, 2 * nf, device="cuda")
    ops.linear_bf16(hfin, wo, mode=EPI_FINAL, out=o, bias=bo, std=std, mean=mean, nfeats=nf, T=T)
    r = ((hfin.float() @ wo.float().t() + bo) * std + mean).view(S // 2, 2, T, nf).permute(0, 2, 1, 3).reshape(S // 2, T, 2 * nf)
    torch.cuda.synchronize()
    assert _relerr(o, r) < 1e-5


@pytest.mark.parametrize("D", [512, 1024])
def test_ln_modulate(ops, D):
    S, T = 5, 33
    h = _rand((S * T, D), 20, 3.0) + 0.7
    mod0 = _rand((S, 4 * D), 21, 0.3)
    mod1 = _rand((S, 2 * D), 22, 0.3)
    o0, o1 = ops.ln_modulate(h, T, mod0[:, D:3 * D], mod1=mod1)
    ln = torch.nn.functional.layer_norm(h, (D,), eps=1e-6)
    sc0, sh0 = mod0[:, D:2 * D], mod0[:, 2 * D:3 * D]
    r0 = ln.view(S, T, D) * (1 + sc0[:, None]) + sh0[:, None]
    r1 = ln.view(S, T, D) * (1 + mod1[:, None, :D]) + mod1[:, None, D:]
    torch.cuda.synchronize()
    assert _relerr(o0, r0.reshape(-1, D)) < 4e-3
    assert _relerr(o1, r1.reshape(-1, D)) < 4e-3


def _ref_attention(q, k, v, scale, pad=None):
    # q,k,v: (S,H,T,hd) fp32; zero-attention sink appended (logit 0, value 0); pad: (S,T) bool, True = key ignored
    logits = torch.einsum("shqd,shkd->shqk", q, k) * scale
    if pad is not None:
        logits = logits.masked_fill(pad[:, None, None, :], float("-inf"))
    zeros = torch.zeros(*logits.shape[:-1], 1, device=q.device)
    p = torch.softmax(torch.cat([logits, zeros], -1), -1)[..., :-1]
    return torch.einsum("shqk,shkd->shqd", p, v)


@pytest.mark.parametrize("T,H,hd,S,kv_xor", [(300, 8, 128, 4, 0), (300, 8, 64, 4, 0), (300, 2, 128, 4, 1),
                                             (77, 2, 64, 2, 0), (128, 2, 128, 2, 1), (255, 1, 64, 2, 0),
                                             (303, 2, 128, 2, 0), (303, 2, 64, 4, 1), (24, 4, 64, 100, 0), (159, 2, 128, 2, 0),
                                             (160, 2, 64, 2, 1), (300, 8, 128, 40, 1), (300, 8, 64, 56, 0),
                                             # block / tile boundaries of the 64-key, 128-row tiling
                                             (1, 2, 64, 2, 0), (63, 2, 128, 2, 0), (64, 2, 64, 2, 1), (127, 1, 128, 6, 0),
                                             (191, 2, 64, 2, 0), (256, 2, 128, 2, 1), (319, 2, 128, 3, 0),
                                             # longer than the resident K/V ring: blocks stream through it
                                             (320, 2, 128, 2, 0), (400, 2, 64, 4, 1), (512, 2, 128, 2, 0),
                                             (512, 8, 64, 40, 1), (700, 1, 128, 4, 0)])
def test_attention_zero_sink(ops, T, H, hd, S, kv_xor):
    from mixermdm_b200._lib import EPI_BF16
    D = H * hd
    M = S * T
    Tp = (T + 7) // 8 * 8
    x = _rand((M, D), 30, dtype=torch.bfloat16)
    w = _rand((3 * D, D), 31, 1.5 / math.sqrt(D), dtype=torch.bfloat16)
    b = _rand((3 * D,), 32, 0.1)
    qkv = torch.zeros(M, 2 * D, dtype=torch.bfloat16, device="cuda")
    vt = torch.zeros(S * H, hd, Tp, dtype=torch.bfloat16, device="cuda")
    # in-proj GEMM writes Q|K row-major and V transposed per (sequence, head)
    ops.linear_bf16(x, w, mode=EPI_BF16, out=qkv, bias=b, vt=vt, vt_col0=2 * D, T=T, H=H, hd=hd, Tp=Tp)
    ref_qkv = x.float() @ w.float().t() + b
    torch.cuda.synchronize()
    assert _relerr(qkv, ref_qkv[:, :2 * D]) < 6e-3
    v_ref = ref_qkv[:, 2 * D:].view(S, T, H, hd).permute(0, 2, 3, 1)  # S,H,hd,T
    assert _relerr(vt.view(S, H, hd, Tp)[..., :T], v_ref) < 6e-3
    out = torch.zeros(M, D, dtype=torch.bfloat16, device="cuda")
    ops.attention(qkv[:, :D], qkv[:, D:], vt, out, S, T, H, hd, kv_xor)
    torch.cuda.synchronize()
    q = qkv[:, :D].float().view(S, T, H, hd).permute(0, 2, 1, 3)
    k = qkv[:, D:].float().view(S, T, H, hd).permute(0, 2, 1, 3)
    v = vt.float().view(S, H, hd, Tp)[..., :T].permute(0, 1, 3, 2)
    if kv_xor:
        idx = torch.arange(S, device="cuda") ^ kv_xor
        k, v = k[idx], v[idx]
    ref = _ref_attention(q, k, v, 1.0 / math.sqrt(hd)).permute(0, 2, 1, 3).reshape(M, D)
    assert _relerr(out, ref) < 1e-2
    assert (out.float() - ref).abs().max().item() < 0.05


@pytest.mark.parametrize("T,H,hd,S,kv_xor,kind", [(300, 2, 128, 6, 0, "prefix"), (300, 2, 64, 6, 1, "prefix"),
                                                  (300, 2, 128, 4, 1, "random"), (77, 2, 64, 4, 0, "random"),
                                                  (400, 1, 128, 4, 0, "prefix"), (130, 2, 64, 4, 0, "edge")])
def test_attention_key_padding_mask(ops, T, H, hd, S, kv_xor, kind):
    """key_padding_mask of nn.MultiheadAttention(add_zero_attn=True) (reference models/utils/layers.py:36-44): masked
    keys get no weight, the zero-attention key always does - so a row whose keys are all masked returns exactly 0.
    Masked keys carry large logits here: they must neither leak nor set the softmax reference. Two sequences (persons)
    share a mask row, as in the engine."""
    D = H * hd
    Tp = (T + 7) // 8 * 8
    g = torch.Generator(device="cuda").manual_seed(11)
    q = torch.randn(S, T, H, hd, device="cuda", generator=g).to(torch.bfloat16)
    k = torch.randn(S, T, H, hd, device="cuda", generator=g).to(torch.bfloat16)
    v = torch.randn(S, T, H, hd, device="cuda", generator=g).to(torch.bfloat16)
    rows = S // 2
    keep = torch.ones(rows, T, device="cuda")
    if kind == "prefix":
        lens = [T, 1, T // 2, 65, 0, 64][:rows]
        for r, n in enumerate(lens):
            keep[r, n:] = 0
    elif kind == "random":
        keep = (torch.rand(rows, T, device="cuda", generator=g) > 0.4).float()
        keep[0] = 0                      # everything masked: only the zero key is left
    else:
        keep[0, 63:65] = 0               # holes across a 64-key block boundary
        keep[1, :64] = 0                 # the whole first block masked (it normally sets the running maximum)
    pad_rows = ~(keep > 0.5)
    pad = pad_rows.repeat_interleave(2, dim=0)            # sequence s -> mask row s // 2
    k = torch.where(pad[:, :, None, None], 6.0 * k.float(), k.float()).to(torch.bfloat16)   # masked keys: big logits
    qk = torch.cat([q.reshape(S * T, D), k.reshape(S * T, D)], 1).contiguous()
    vt = torch.zeros(S * H, hd, Tp, dtype=torch.bfloat16, device="cuda")
    vt[..., :T] = v.permute(0, 2, 3, 1).reshape(S * H, hd, T)
    out = torch.full((S * T, D), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.attention(qk[:, :D], qk[:, D:], vt, out, S, T, H, hd, kv_xor, key_mask=keep, persons=2)
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    if kv_xor:
        idx = torch.arange(S, device="cuda") ^ kv_xor
        kf, vf = kf[idx], vf[idx]                         # (the partner sequence belongs to the same mask row)
    ref = _ref_attention(qf, kf, vf, 1.0 / math.sqrt(hd), pad).permute(0, 2, 1, 3).reshape(S * T, D)
    assert torch.isfinite(out.float()).all()
    assert _relerr(out, ref) < 1e-2
    dead = pad.all(dim=1)
    if dead.any():
        assert out.view(S, T, D)[dead].abs().max().item() == 0.0


@pytest.mark.parametrize("T,hd", [(300, 128), (300, 64), (512, 128)])
def test_attention_online_rescale(ops, T, hd):
    """Keys whose logits grow along the sequence: later 64-key blocks exceed the running maximum by far more than the
    lazy-rescale threshold (2^8), so the in-place rescale of O and of the row sum is exercised in every row."""
    S, H = 3, 2
    D = H * hd
    Tp = (T + 7) // 8 * 8
    g = torch.Generator(device="cuda").manual_seed(5)
    q = (3.0 * torch.randn(S, T, H, hd, device="cuda", generator=g)).to(torch.bfloat16)
    ramp = (0.1 + 4.0 * torch.arange(T, device="cuda") / T).view(1, T, 1, 1)
    k = (torch.randn(S, T, H, hd, device="cuda", generator=g) * ramp).to(torch.bfloat16)
    v = torch.randn(S, T, H, hd, device="cuda", generator=g).to(torch.bfloat16)
    qk = torch.cat([q.reshape(S * T, D), k.reshape(S * T, D)], 1).contiguous()
    vt = torch.zeros(S * H, hd, Tp, dtype=torch.bfloat16, device="cuda")
    vt[..., :T] = v.permute(0, 2, 3, 1).reshape(S * H, hd, T)
    out = torch.zeros(S * T, D, dtype=torch.bfloat16, device="cuda")
    ops.attention(qk[:, :D], qk[:, D:], vt, out, S, T, H, hd, 0)
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    logits = torch.einsum("shqd,shkd->shqk", qf, kf) / math.sqrt(hd)
    blockmax = torch.stack([logits[..., i:i + 64].amax(-1) for i in range(0, T, 64)], -1).clamp_min(0)
    jumps = (blockmax[..., 1:] - blockmax.cummax(-1).values[..., :-1]).amax(-1)
    assert (jumps > 8 * math.log(2)).float().mean().item() > 0.5      # the rescale path really runs in most rows
    ref = _ref_attention(qf, kf, vf, 1.0 / math.sqrt(hd)).permute(0, 2, 1, 3).reshape(S * T, D)
    assert torch.isfinite(out.float()).all()
    assert _relerr(out, ref) < 1e-2


def test_small_linear(ops):
    a = _rand((7, 768), 40)
    w = _rand((512, 768), 41, 0.03)
    b = _rand((512,), 42)
    o = ops.small_linear(a, w, b, act=1)
    r = torch.nn.functional.silu(a @ w.t() + b)
    torch.cuda.synchronize()
    assert _relerr(o, r) < 1e-5


@pytest.mark.parametrize("shape", [(1, 299, 524), (3, 300, 524), (2, 5, 524), (1, 1, 524), (4, 77, 264)])
@pytest.mark.parametrize("sigma", [1.0, 2.0])
def test_smooth_motion_matches_scipy(shape, sigma):
    """Post-processing of the reference's inference script (scripts/infer/mixermdm.py:131):
    gaussian_filter1d(motion, sigma, axis=time, mode='nearest'); scipy is the oracle."""
    from scipy.ndimage import gaussian_filter1d
    from mixermdm_b200 import postprocess
    x = _rand(shape, 7)
    y = postprocess.smooth_motion(x, sigma)
    torch.cuda.synchronize()
    ref = gaussian_filter1d(x.cpu().numpy(), sigma, axis=1, mode="nearest")
    assert y.shape == x.shape
    assert (y.cpu() - torch.from_numpy(ref)).abs().max().item() < 2e-6
    # the script's own call: one sample viewed as (T, 2, 262)
    if shape[2] == 524:
        one = postprocess.motion_to_persons(x[0], sigma=sigma)
        ref1 = gaussian_filter1d(x[0].cpu().numpy().reshape(shape[1], 2, -1), sigma, axis=0, mode="nearest")
        assert one.shape == (shape[1], 2, 262)
        assert (one.cpu() - torch.from_numpy(ref1)).abs().max().item() < 2e-6
