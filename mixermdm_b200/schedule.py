"""Host-side diffusion tables (float64 numpy, computed once per sampler — the reference rebuilds them on every
forward). Follows models/utils/gaussian_diffusion.py: cosine betas (:236-279), DDIM striding (:1279-1332), the
re-spaced betas of MixerDiffusion.__init__ (:1436-1463) and the GaussianDiffusion tables (:331-382). The kernels
receive the per-step scalars as fp32, cast exactly like `_extract_into_tensor` (:1264-1277) does."""
import math

import numpy as np


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps):
    n = int(num_diffusion_timesteps)
    if schedule_name == "linear":
        scale = 1000 / n
        return np.linspace(scale * 0.0001, scale * 0.02, n, dtype=np.float64)
    if schedule_name == "cosine":
        def abar(t):
            return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        return np.array([min(1 - abar((i + 1) / n) / abar(i / n), 0.999) for i in range(n)])
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def space_timesteps(num_timesteps, section_counts):
    """subset of original timesteps to keep; "ddimN" = fixed integer stride giving exactly N steps"""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == want:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    per, extra = divmod(num_timesteps, len(section_counts))
    start, steps = 0, []
    for i, count in enumerate(section_counts):
        size = per + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        steps += [start + round(cur) for cur in _accumulate(stride, count)]
        start += size
    return set(steps)


def _accumulate(stride, count):
    cur, out = 0.0, []
    for _ in range(count):
        out.append(cur)
        cur += stride
    return out


class SpacedSchedule:
    """tables of the re-spaced process + the map back to original timesteps"""

    def __init__(self, use_timesteps, betas):
        base_ac = np.cumprod(1.0 - np.array(betas, dtype=np.float64), axis=0)
        use = set(use_timesteps)
        self.timestep_map, new_betas, last = [], [], 1.0
        for i, ac in enumerate(base_ac):
            if i in use:
                new_betas.append(1 - ac / last)
                last = ac
                self.timestep_map.append(i)
        self.original_num_steps = len(base_ac)
        self.betas = np.array(new_betas, dtype=np.float64)
        assert self.betas.ndim == 1 and (self.betas > 0).all() and (self.betas <= 1).all()
        self.num_timesteps = int(self.betas.shape[0])
        alphas = 1.0 - self.betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = self.betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:])) \
            if self.num_timesteps > 1 else np.log(self.posterior_variance)
        self.posterior_mean_coef1 = self.betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)

    def same_as(self, other):
        return self.timestep_map == other.timestep_map and np.array_equal(self.betas, other.betas)

    def ddim_coefficients(self):
        """(nsteps, 4) fp32: sqrt_recip_ab, sqrt_recipm1_ab, sqrt(ab_prev), sqrt(1 - ab_prev) with eta = 0.
        The last two are formed from the fp32-cast alpha_bar_prev in fp32, as ddim_sample does (:1938-1956)."""
        c = np.zeros((self.num_timesteps, 4), dtype=np.float32)
        c[:, 0] = self.sqrt_recip_alphas_cumprod.astype(np.float32)
        c[:, 1] = self.sqrt_recipm1_alphas_cumprod.astype(np.float32)
        ab_prev = self.alphas_cumprod_prev.astype(np.float32)
        c[:, 2] = np.sqrt(ab_prev)
        c[:, 3] = np.sqrt(np.float32(1.0) - ab_prev - np.float32(0.0))
        return c

    def ddpm_coefficients(self, var_type="fixed_small"):
        """Ancestral (DDPM) update x_{t-1} = coef1 x0 + coef2 x_t + sigma_t z in the kernels' table format
        (reference GaussianDiffusion.q_posterior_mean_variance :452-478, p_mean_variance :520-538, p_sample :683-691).
        The fused step forms A = c2 - c3 / c1 and B = c3 / c1 * c0 from its row (c0, c1, c2, c3) and applies
        x_prev = A x0 + B x_t, so the row (1, 1, coef1 + coef2, coef2) gives A = coef1, B = coef2.
        Returns ((nsteps, 4) fp32 table, (nsteps,) float64 sigma_t = exp(0.5 log variance); sigma is unused at t = 0)."""
        c = np.zeros((self.num_timesteps, 4), dtype=np.float32)
        c[:, 0] = 1.0
        c[:, 1] = 1.0
        c[:, 3] = self.posterior_mean_coef2.astype(np.float32)
        # (the kernel forms c2 - c3 in double: it gets coef1 back to half an fp32 ulp of coef1 + coef2 <= 1)
        c[:, 2] = (self.posterior_mean_coef1.astype(np.float32).astype(np.float64) + c[:, 3].astype(np.float64))
        if var_type == "fixed_small":
            logvar = self.posterior_log_variance_clipped
        elif var_type == "fixed_large":
            logvar = np.log(np.append(self.posterior_variance[1], self.betas[1:]))
        else:
            raise NotImplementedError("learned variances are not on the MixerMDM path")
        return c, np.exp(0.5 * logvar)
