"""ctypes binding of include/mixermdm_b200.h. The library is required: there is no CPU or PyTorch fallback."""
import ctypes as C
import os

from . import _build

_LIB = None
ABI_VERSION = 4   # MMDM_ABI_VERSION of include/mixermdm_b200.h


class MmdmError(RuntimeError):
    pass


class Epilogue(C.Structure):
    """mirror of `mmdm_epilogue` (include/mixermdm_b200.h)"""
    _fields_ = [
        ("mode", C.c_int32), ("ldo", C.c_int32),
        ("bias", C.c_void_p), ("out", C.c_void_p),
        ("vt_col0", C.c_int32), ("T", C.c_int32), ("H", C.c_int32), ("hd", C.c_int32), ("Tp", C.c_int32),
        ("vt", C.c_void_p),
        ("resid", C.c_void_p), ("out_bf16", C.c_void_p),
        ("pe", C.c_void_p), ("pe_ld", C.c_int32), ("dup_row_offset", C.c_int64),
        ("std_", C.c_void_p), ("mean", C.c_void_p), ("nfeats", C.c_int32), ("persons_per_row", C.c_int32),
        ("lnp_scale0", C.c_void_p), ("lnp_ld0", C.c_int32), ("lnp_out0", C.c_void_p),
        ("lnp_scale1", C.c_void_p), ("lnp_ld1", C.c_int32), ("lnp_out1", C.c_void_p),
        ("lnp_stats", C.c_void_p), ("lnp_pitch", C.c_int64), ("lnp_T", C.c_int32), ("lnp_rows", C.c_int32),
        ("lnp_dup_seq", C.c_int32),
        ("lnc_stats", C.c_void_p), ("lnc_pitch", C.c_int64),
        ("lnc_g", C.c_void_p), ("lnc_b", C.c_void_p), ("lnc_ld", C.c_int32),
        ("lnp_rowa", C.c_void_p), ("lnc_rowa", C.c_void_p), ("lnp_moff_in", C.c_void_p), ("lnp_moff_out", C.c_void_p),
    ]


EPI_BF16, EPI_GELU_BF16, EPI_RESID, EPI_EMBED_PE, EPI_FINAL, EPI_F32 = range(6)
EPI_BF16_LN, EPI_GELU_BF16_LN, EPI_RESID_LN, EPI_EMBED_PE_LN = 8, 9, 10, 11


def lib_path():
    return _build.LIB_PATH


def load():
    """dlopen the in-tree library; raises if it has not been built (no fallback by design)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise MmdmError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(the CUDA extension is mandatory, there is no fallback path)")
    lib = C.CDLL(path)
    lib.mmdm_last_error.restype = C.c_char_p
    lib.mmdm_last_error.argtypes = [C.c_void_p]
    lib.mmdm_abi_version.restype = C.c_int
    if lib.mmdm_abi_version() != ABI_VERSION:
        raise MmdmError(f"{path}: ABI version {lib.mmdm_abi_version()} != {ABI_VERSION}; rebuild the library")
    _LIB = lib
    return lib


def check(rc, what="", handle=None):
    """raise on a non-zero return code with the message kept for `handle` (None: the thread's handle-less calls)"""
    if rc != 0:
        msg = load().mmdm_last_error(handle).decode("utf-8", "replace")
        raise MmdmError(f"{what} failed (rc={rc}): {msg}")


def ptr(t):
    """device pointer of a torch tensor (or None)"""
    return None if t is None else C.c_void_p(t.data_ptr())


def cur_stream(device=None):
    """the current torch stream of `device` (default: the current device) as a cudaStream_t"""
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
