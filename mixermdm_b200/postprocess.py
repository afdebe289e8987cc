"""Post-processing of a sampled motion, the step after the sampler in the reference's inference script
(reference: src/scripts/infer/mixermdm.py:124-132): `output[0]` (T, 524) is viewed as (T, 2, 262) and smoothed along
time with `scipy.ndimage.gaussian_filter1d(x, 1, axis=0, mode='nearest')`. Here the blur runs on the GPU
(`mmdm_smooth_motion`, csrc/motion.cu) for a whole batch before anything is copied to the host."""
import ctypes as C

import numpy as np
import torch

from . import _lib


def gaussian_taps(sigma, truncate=4.0):
    """scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius) with radius = int(truncate * sigma + 0.5), float64"""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1, dtype=np.float64)
    phi = np.exp(-0.5 / (float(sigma) * float(sigma)) * x ** 2)
    return (phi / phi.sum()), radius


def smooth_motion(output, sigma=1.0, truncate=4.0):
    """output: (B, T, F) or (T, F) fp32 CUDA tensor (F % 4 == 0) -> smoothed copy of the same shape."""
    if not output.is_cuda:
        raise _lib.MmdmError("smooth_motion runs on the GPU only (there is no CPU fallback)")
    x = output.detach().to(torch.float32).contiguous()
    x3 = x if x.dim() == 3 else x.unsqueeze(0)
    if x3.dim() != 3:
        raise ValueError(f"expected (B, T, F) or (T, F), got {tuple(output.shape)}")
    B, T, F = x3.shape
    taps, radius = gaussian_taps(sigma, truncate)
    w = (C.c_float * len(taps))(*[float(v) for v in taps])
    out = torch.empty_like(x3)
    lib = _lib.load()
    with torch.cuda.device(x3.device):
        _lib.check(lib.mmdm_smooth_motion(_lib.ptr(x3), _lib.ptr(out), B, T, F, w, radius, _lib.cur_stream()),
                   "smooth_motion")
    return out.view_as(x)


def motion_to_persons(output, smooth=True, sigma=1.0):
    """`generate_loop`'s output format (scripts/infer/mixermdm.py:124-132): (B, T, 524) -> (B, T, 2, 262), blurred."""
    y = smooth_motion(output, sigma) if smooth else output
    return y.reshape(*y.shape[:-1], 2, y.shape[-1] // 2)
