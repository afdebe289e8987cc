"""Operator-level wrappers over the C ABI (one per reference library call replaced). Tensors are torch CUDA
tensors used purely as device-memory handles; all arithmetic happens in libmixermdm_b200.so."""
import ctypes as C

import torch

from . import _lib
from ._lib import (EPI_BF16, EPI_EMBED_PE, EPI_F32, EPI_FINAL, EPI_GELU_BF16, EPI_RESID, Epilogue, check,
                   cur_stream, ptr)


def _req(t, dtype, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype):
        raise TypeError(f"{name}: expected CUDA tensor of dtype {dtype}")
    return t


def linear_bf16(a, w, *, mode, out, bias=None, resid=None, out_bf16=None, vt=None, vt_col0=-1, T=1, H=1, hd=1,
                Tp=1, pe=None, dup_row_offset=0, std=None, mean=None, nfeats=0, persons_per_row=2,
                ln_produce=None, ln_consume=None):
    """out = epilogue(a @ w.T). a: (M,K) bf16, w: (N,K) bf16 (nn.Linear layout), rows may be strided.
    ln_produce (EPI_RESID_LN / EPI_EMBED_PE_LN): dict(scale0, out0, stats, T[, scale1, out1, rows, dup_seq]) - scale*: (S, >=N)
    fp32 views of the AdaLN scale rows, out*: (M', N) bf16, stats: (N // 64, M', 2) fp32.
    ln_consume (EPI_BF16_LN / EPI_GELU_BF16_LN): dict(stats, g, b) - g / b: (S, N) fp32 views with one row pitch; pass T."""
    lib = _lib.load()
    _req(a, torch.bfloat16, "a")
    _req(w, torch.bfloat16, "w")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and a.stride(1) == 1 and w.stride(1) == 1
    ep = Epilogue()
    ep.mode = mode
    ep.ldo = out.stride(0) if mode != EPI_FINAL else 0
    ep.bias = ptr(bias).value if bias is not None else None
    ep.out = ptr(out).value
    ep.vt_col0 = vt_col0
    ep.T, ep.H, ep.hd, ep.Tp = T, H, hd, Tp
    ep.vt = ptr(vt).value if vt is not None else None
    ep.resid = ptr(resid).value if resid is not None else None
    ep.out_bf16 = ptr(out_bf16).value if out_bf16 is not None else None
    ep.pe = ptr(pe).value if pe is not None else None
    ep.pe_ld = pe.stride(0) if pe is not None else 0
    ep.dup_row_offset = dup_row_offset
    ep.std_ = ptr(std).value if std is not None else None
    ep.mean = ptr(mean).value if mean is not None else None
    ep.nfeats = nfeats
    ep.persons_per_row = persons_per_row
    if ln_produce is not None:
        lp = ln_produce
        ep.lnp_scale0, ep.lnp_ld0, ep.lnp_out0 = ptr(lp["scale0"]).value, lp["scale0"].stride(0), ptr(lp["out0"]).value
        if lp.get("scale1") is not None:
            ep.lnp_scale1, ep.lnp_ld1, ep.lnp_out1 = ptr(lp["scale1"]).value, lp["scale1"].stride(0), ptr(lp["out1"]).value
        ep.lnp_stats, ep.lnp_pitch = ptr(lp["stats"]).value, lp["stats"].stride(0) // 2
        ep.lnp_T, ep.lnp_rows, ep.lnp_dup_seq = lp["T"], lp.get("rows", 0), lp.get("dup_seq", 0)
        if lp.get("rowa") is not None:   # (M', 4) int32
            ep.lnp_rowa = ptr(lp["rowa"]).value
        if lp.get("moff_in") is not None:   # (M',) fp32 per-row offsets
            ep.lnp_moff_in = ptr(lp["moff_in"]).value
        if lp.get("moff_out") is not None:
            ep.lnp_moff_out = ptr(lp["moff_out"]).value
    if ln_consume is not None:
        lc = ln_consume
        assert lc["g"].stride(0) == lc["b"].stride(0)
        ep.lnc_stats, ep.lnc_pitch = ptr(lc["stats"]).value, lc["stats"].stride(0) // 2
        ep.lnc_g, ep.lnc_b, ep.lnc_ld = ptr(lc["g"]).value, ptr(lc["b"]).value, lc["g"].stride(0)
        if lc.get("rowa") is not None:
            ep.lnc_rowa = ptr(lc["rowa"]).value
    with torch.cuda.device(a.device):   # kernels launch on the tensors' device, whatever the current one is
        check(lib.mmdm_linear_bf16(ptr(a), C.c_int(a.stride(0)), ptr(w), C.c_int(w.stride(0)), M, N, K, C.byref(ep),
                                   cur_stream()), "mmdm_linear_bf16")
    return out


def ln_modulate(h, T, mod0, out0=None, mod1=None, out1=None):
    """AdaLN body: bf16(LN(h)*(1+scale)+shift) with per-sequence (scale|shift) rows in mod (pitch = stride(0))."""
    lib = _lib.load()
    _req(h, torch.float32, "h")
    M, D = h.shape
    if out0 is None:
        out0 = torch.empty(M, D, dtype=torch.bfloat16, device=h.device)
    if mod1 is not None and out1 is None:
        out1 = torch.empty(M, D, dtype=torch.bfloat16, device=h.device)
    with torch.cuda.device(h.device):
        check(lib.mmdm_ln_modulate(ptr(h), M, D, T, ptr(mod0), C.c_int(mod0.stride(0)), ptr(out0),
                                   ptr(mod1), C.c_int(mod1.stride(0) if mod1 is not None else 0), ptr(out1),
                                   cur_stream()), "mmdm_ln_modulate")
    return (out0, out1) if mod1 is not None else out0


def attention(q, k, vt, out, S, T, H, hd, kv_xor=0, key_mask=None, persons=1):
    """q/k: 2-D bf16 views (rows = s*T+t) whose columns are the H*hd projected features; vt: (S*H, hd, Tp).
    key_mask: optional (rows, T) float tensor, > 0.5 = key attendable; sequence s reads row (s // persons) % rows."""
    lib = _lib.load()
    with torch.cuda.device(q.device):
        if key_mask is None:
            check(lib.mmdm_attention(ptr(q), C.c_int(q.stride(0)), ptr(k), C.c_int(k.stride(0)), ptr(vt),
                                     C.c_int(vt.stride(1)), ptr(out), C.c_int(out.stride(0)), S, T, H, hd, kv_xor,
                                     cur_stream()), "mmdm_attention")
            return out
        m = key_mask.detach().to(device=q.device, dtype=torch.float32).contiguous()
        assert m.dim() == 2 and m.shape[1] == T
        words = torch.empty(S, (T + 64) // 64, dtype=torch.int64, device=q.device)
        check(lib.mmdm_build_key_mask(ptr(m), int(m.shape[0]), S, T, persons, ptr(words), cur_stream()),
              "mmdm_build_key_mask")
        check(lib.mmdm_attention_masked(ptr(q), C.c_int(q.stride(0)), ptr(k), C.c_int(k.stride(0)), ptr(vt),
                                        C.c_int(vt.stride(1)), ptr(out), C.c_int(out.stride(0)), S, T, H, hd, kv_xor,
                                        ptr(words), cur_stream()), "mmdm_attention_masked")
    return out


def small_linear(a, w, b=None, act=0):
    lib = _lib.load()
    _req(a, torch.float32, "a")
    R, K = a.shape
    N = w.shape[0]
    out = torch.empty(R, N, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        check(lib.mmdm_small_linear(ptr(a), C.c_int(a.stride(0)), ptr(w.contiguous()), ptr(b), ptr(out), N, R, N, K,
                                    act, cur_stream()), "mmdm_small_linear")
    return out
