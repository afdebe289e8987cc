"""Sampler entry point with the reference's interface: `MixerDiffusion(...).ddim_sample_loop(model, shape, ...)`
(reference models/utils/gaussian_diffusion.py:1434-1463, 1769-2091). The 50-step loop runs as one CUDA-graph
replayed step on the device; the tables are built once on the host in float64 (schedule.py)."""
import enum

import numpy as np
import torch

from . import schedule
from .modules import _check_mask
from .schedule import get_named_beta_schedule, space_timesteps  # noqa: F401  (re-exported, reference names)


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()


class ModelVarType(enum.Enum):
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


class LossType(enum.Enum):
    MSE = enum.auto()
    RESCALED_MSE = enum.auto()
    KL = enum.auto()
    RESCALED_KL = enum.auto()


class MixerDiffusion:
    """Two-stream DDIM sampler of MixerMDM. Only the configuration the reference's inference path uses is
    accepted: x0-prediction (START_X), fixed variance, eta = 0, no clipping, no guidance callbacks."""

    def __init__(self, use_timesteps, align=True, *, betas, model_mean_type=ModelMeanType.START_X,
                 model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE, rescale_timesteps=False):
        if model_mean_type != ModelMeanType.START_X:
            raise NotImplementedError("MixerDiffusion: only ModelMeanType.START_X is on the MixerMDM path")
        if rescale_timesteps:
            raise NotImplementedError("MixerDiffusion: rescale_timesteps=False only")
        self.align = align
        self.use_timesteps = set(use_timesteps)
        self.model_mean_type, self.model_var_type, self.loss_type = model_mean_type, model_var_type, loss_type
        self.rescale_timesteps = rescale_timesteps
        self._betas_full = np.array(betas, dtype=np.float64)
        s = schedule.SpacedSchedule(self.use_timesteps, self._betas_full)
        self.schedule = s
        self.timestep_map = s.timestep_map
        self.original_num_steps = s.original_num_steps
        self.betas = s.betas
        self.num_timesteps = s.num_timesteps
        self.alphas_cumprod = s.alphas_cumprod
        self.alphas_cumprod_prev = s.alphas_cumprod_prev
        self.sqrt_recip_alphas_cumprod = s.sqrt_recip_alphas_cumprod
        self.sqrt_recipm1_alphas_cumprod = s.sqrt_recipm1_alphas_cumprod

    # ---------------------------------------------------------------------------------------------
    def _engine_for(self, model, B, T, device):
        from .modules import ClassifierFreeSampleModelX2, Mixer
        if not isinstance(model, ClassifierFreeSampleModelX2) or not isinstance(model.model, Mixer):
            raise TypeError("ddim_sample_loop expects ClassifierFreeSampleModelX2(Mixer, scale) from mixermdm_b200")
        mixer = model.model
        eng = mixer.engine(2 * B, T, device, sched=self.schedule, cfg_scale=float(model.s), center=self.align)
        return mixer, eng

    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                      model_kwargs=None, device=None, progress=False, skip_timesteps=0, init_image=None,
                      randomize_class=False, cond_fn_with_grad=False, dump_steps=None, const_noise=False,
                      generator=None, step_noise=None):
        """Ancestral (DDPM) sampling over the same two-stream step (GaussianDiffusion.p_sample_loop's signature,
        gaussian_diffusion.py:694-760; update of p_sample :683-691: mean = coef1 x0 + coef2 x_t, + sigma_t z for
        t != 0). The reference's MixerDiffusion inherits p_sample but cannot run it (its p_mean_variance override takes
        the second stream, SURVEY 8a note 2), so this variant is derived and its parity is formula-level (tested)."""
        if dump_steps is not None or const_noise or skip_timesteps or init_image is not None:
            raise NotImplementedError()
        if clip_denoised or denoised_fn is not None or cond_fn is not None or cond_fn_with_grad or randomize_class:
            raise NotImplementedError("clip_denoised=False, no denoised_fn/cond_fn: the MixerMDM inference path")
        model_kwargs = model_kwargs or {}
        mask = model_kwargs.get("mask")
        cond = model_kwargs["cond"]
        if device is None:
            device = cond.device if noise is None else noise.device
        B, T, D = shape
        assert D == 524
        if noise is None:
            noise = torch.randn(*shape, device=device, generator=generator)
        var = {ModelVarType.FIXED_SMALL: "fixed_small", ModelVarType.FIXED_LARGE: "fixed_large"}.get(self.model_var_type)
        if var is None:
            raise NotImplementedError("MixerDiffusion.p_sample_loop: fixed variances only")
        mixer, eng = self._engine_for(model, B, T, torch.device(device))
        _check_mask(mixer, mask)
        return eng.sample_ddpm(cond, noise, step_noise=step_noise, generator=generator, var_type=var, mask=mask)

    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                         model_kwargs=None, device=None, progress=False, eta=0.0, skip_timesteps=0, init_image=None,
                         randomize_class=False, cond_fn_with_grad=False, dump_steps=None, const_noise=False,
                         x_start=None):
        """returns pred_xstart2 of the last step, (B,T,524) in real (de-normalised) space."""
        if dump_steps is not None or const_noise:
            raise NotImplementedError()
        if clip_denoised or denoised_fn is not None or cond_fn is not None or cond_fn_with_grad or randomize_class:
            raise NotImplementedError("clip_denoised=False, no denoised_fn/cond_fn: the MixerMDM inference path")
        if eta != 0.0 or skip_timesteps or init_image is not None or x_start is not None:
            raise NotImplementedError("eta=0, full schedule, no inpainting: the MixerMDM inference path")
        model_kwargs = model_kwargs or {}
        mask = model_kwargs.get("mask")
        cond = model_kwargs["cond"]
        if device is None:
            device = cond.device if noise is None else noise.device
        B, T, D = shape
        assert D == 524
        if noise is None:
            noise = torch.randn(*shape, device=device)
        mixer, eng = self._engine_for(model, B, T, torch.device(device))
        _check_mask(mixer, mask)
        want_out = (mixer.mode == "eval")
        res = eng.sample(cond, noise, store_influence=mixer.store_influence, store_outputs=want_out, mask=mask)
        if not (mixer.store_influence or want_out):
            return res
        res = list(res)
        out = res.pop(0)
        if mixer.store_influence:
            h1, h2 = res.pop(0), res.pop(0)
            from .modules import history_view
            mixer.history_influence_i1.extend(list(history_view(h1, mixer.mixing_mode).unbind(0)))
            mixer.history_influence_i2.extend(list(history_view(h2, mixer.mixing_mode).unbind(0)))
        if want_out:
            o1, o2, om = res
            mixer.history_out1.extend(list(o1.unbind(0)))
            mixer.history_out2.extend(list(o2.unbind(0)))
            mixer.history_out_influenced.extend(list(om.unbind(0)))
        return out
