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
