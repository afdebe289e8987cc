"""TEST INFRASTRUCTURE ONLY — regenerates tests/golden/*.npz from the UNMODIFIED reference (CPU, build container).

    python oracle/make_goldens.py

Everything is derived from fixed numpy seeds (oracle/synth.py), so the fixtures only store the reference's
OUTPUTS (plus the state-dict key list and the schedule tables). Small geometry (oracle/synth.SMALL_DIMS,
B=2, T=24) keeps the files small; full-size parity is checked live against the oracle on the GPU box.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness, synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
B, T = 2, 24


CHAIN_B, CHAIN_T = 1, 16


def chain():
    """tests/golden/reference_chain.npz: the reference's own 50-step trajectory for teacher-forced per-step parity.
    After every ddim_sample the two state streams are rounded to fp16-representable values BEFORE they are fed to the
    next step, so the stored fp16 tensors are exactly the inputs the reference saw (and every output is the next
    input): x1[i], x2[i] = state entering DDIM index 49 - i (i = 0: the noise), x1[50], x2[50] = state after index 0,
    final = pred_xstart2 of index 0 (the sampler's return value), fp32."""
    dims = synth.SMALL_DIMS
    cond, noise = synth.inputs(CHAIN_B, CHAIN_T, seed=3)
    mix, diff, cfgm = ref_harness.build_reference(dims, mixing_mode=4, store_influence=False)
    mix.load_state_dict(synth.synthetic_state_dict(mix.state_dict(), seed=0), strict=True)
    q = lambda t: t.to(torch.float16).to(torch.float32)  # noqa: E731
    x1 = x2 = q(noise)
    xs1, xs2 = [x1.numpy().astype(np.float16)], [x2.numpy().astype(np.float16)]
    final = None
    with torch.no_grad(), ref_harness.reference_env():
        for idx in range(diff.num_timesteps - 1, -1, -1):
            tt = torch.full((CHAIN_B,), idx, dtype=torch.long)
            o = diff.ddim_sample(cfgm, x1, x2, tt, clip_denoised=False, model_kwargs={"mask": None, "cond": cond})
            x1, x2 = q(o["sample"]), q(o["sample2"])
            xs1.append(x1.numpy().astype(np.float16))
            xs2.append(x2.numpy().astype(np.float16))
            if idx == 0:
                final = o["pred_xstart2"].numpy()
    np.savez_compressed(os.path.join(OUT, "reference_chain.npz"), x1=np.stack(xs1), x2=np.stack(xs2), final=final)
    with open(os.path.join(OUT, "reference_chain.json"), "w") as f:
        json.dump({"B": CHAIN_B, "T": CHAIN_T, "dims": dims, "weights_seed": 0, "inputs_seed": 3, "mixing_mode": 4,
                   "generator": "oracle/make_goldens.py chain()", "torch": torch.__version__}, f, indent=1)
    print("chain:", os.path.getsize(os.path.join(OUT, "reference_chain.npz")) // 1024, "KiB")


TEXT_PROMPTS = {
    "text": ["two people greet each other and shake hands", "one person pushes the other who stumbles backwards"],
    "text_individual1": ["a person walks forward and extends the right hand", "a person steps in and pushes with both arms"],
    "text_individual2": ["a person stands still then raises an arm to shake hands", "a person is pushed and takes three steps back"],
}


def build_text_reference(seed=0):
    """Holder objects carrying the reference's own text methods (MixerMDM.generate_cond / text_process,
    in2IN.text_process - imported, unmodified) over the restated CLIP modules of oracle/text_oracle.py, with the
    attribute names (= state-dict keys) of the real model. Returns (mixermdm_holder, state_dict)."""
    from torch import nn
    from oracle import text_oracle as to
    with ref_harness.reference_env():
        import clip as clip_stub
        clip_stub.tokenize = to.stub_tokenize
        from models.in2in import in2IN
        from models.mixermdm import MixerMDM

    def holder(cls):
        obj = cls.__new__(cls)
        nn.Module.__init__(obj)
        return obj
    m = holder(MixerMDM)
    m.model1, m.model2 = holder(in2IN), holder(in2IN)
    to.attach_clip(m.model1)
    m.model1.clipTransEncoder_individual, m.model1.clip_ln_individual = to.text_encoder_pair()
    to.attach_clip(m)
    m.model2.token_embedding, m.model2.clip_transformer = m.token_embedding, m.clip_transformer
    # (like the reference, models/mixermdm.py:229-233: model2 borrows the token table, the blocks and ln_final but keeps
    #  its own positional_embedding parameter)
    m.model2.positional_embedding = nn.Parameter(torch.zeros(to.CTX, to.WIDTH))
    m.model2.ln_final, m.model2.dtype = m.ln_final, m.dtype
    m.model2.clipTransEncoder_interaction, m.model2.clip_ln_interaction = to.text_encoder_pair()
    m.clipTransEncoder, m.clip_ln = to.text_encoder_pair()
    m.cond_pre_func1, m.cond_pre_func2 = m.model1.text_process, m.model2.text_process
    sd = to.text_state_dict(m.state_dict(), seed=seed)
    m.load_state_dict(sd, strict=True)
    return m.eval(), sd


def text():
    """tests/golden/reference_text.npz: the (B, 6144) condition the reference's generate_cond builds from prompts"""
    m, _ = build_text_reference(seed=0)
    batch = {k: list(v) for k, v in TEXT_PROMPTS.items()}
    with torch.no_grad(), ref_harness.reference_env():
        cond = m.generate_cond(batch)
    np.savez_compressed(os.path.join(OUT, "reference_text.npz"), cond=cond.numpy())
    with open(os.path.join(OUT, "reference_text.json"), "w") as f:
        json.dump({"prompts": TEXT_PROMPTS, "weights_seed": 0, "tokenizer": "oracle.text_oracle.stub_tokenize",
                   "generator": "oracle/make_goldens.py text()", "torch": torch.__version__}, f, indent=1)
    print("text:", cond.shape, float(cond.abs().mean()))


MASK_LENS = (24, 15)


def masked():
    """tests/golden/reference_masked.npz: the three networks with a frame mask (generate_src_mask's (B, T, 2) layout,
    models/mixermdm.py:358-372): both denoisers alone and Mixer.forward with align=False in modes 2 and 4 (with
    align=True the reference's masked trajectory alignment raises, utils/alignment.py:86-89)."""
    dims = synth.SMALL_DIMS
    cond, noise = synth.inputs(B, T, seed=0)
    t_mid = torch.full((B,), 500, dtype=torch.long)
    x1, x2 = noise, 0.9 * noise.flip(0)
    mask = torch.ones(B, T, 2)
    for i, n in enumerate(MASK_LENS):
        mask[i, n:] = 0
    arrays = {"mask": mask.numpy()}
    for mode in (2, 4):
        mix, _, _ = ref_harness.build_reference(dims, mixing_mode=mode, store_influence=True)
        mix.load_state_dict(synth.synthetic_state_dict(mix.state_dict(), seed=0), strict=True)
        mix.align = False
        with torch.no_grad(), ref_harness.reference_env():
            arrays[f"mixer_out_mode{mode}"] = mix(x1, t_mid, cond=cond, mask=mask, x2=x2).numpy()
            arrays[f"mixer_w1_mode{mode}"] = mix.history_influence_i1[-1].expand(B, T, 262).numpy().copy()
            if mode == 4:
                arrays["den_individual"] = mix.denoiser1(x1[..., :262], t_mid, mask=mask, cond=cond[:, 2304:3072]).numpy()
                arrays["den_interaction"] = mix.denoiser2(x2, t_mid, mask=mask, cond=cond[:, :2304]).numpy()
                try:
                    mix.align = True
                    mix(torch.cat([x1, x1]), torch.cat([t_mid, t_mid]), cond=torch.cat([cond, cond]),
                        mask=torch.cat([mask, mask]), x2=torch.cat([x2, x2]))
                    arrays["align_with_mask_runs"] = np.array(1)
                except Exception as e:   # noqa: BLE001  (recorded: the reason masks need align=False)
                    arrays["align_with_mask_runs"] = np.array(0)
                    print("reference Mixer(align=True) with a mask:", type(e).__name__, str(e)[:120])
    np.savez_compressed(os.path.join(OUT, "reference_masked.npz"), **arrays)
    with open(os.path.join(OUT, "reference_masked.json"), "w") as f:
        json.dump({"B": B, "T": T, "dims": dims, "weights_seed": 0, "inputs_seed": 0, "lens": list(MASK_LENS),
                   "generator": "oracle/make_goldens.py masked()", "torch": torch.__version__}, f, indent=1)
    print("masked:", sorted(arrays))


def schedule_tables():
    """tests/golden/reference_schedule.npz: the posterior tables of the reference's re-spaced diffusion object (the
    DDPM side of GaussianDiffusion.__init__, gaussian_diffusion.py:355-382), float64"""
    _, diff, _ = ref_harness.build_reference(synth.SMALL_DIMS, mixing_mode=4, store_influence=False)
    names = ("betas", "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1",
             "posterior_mean_coef2")
    np.savez_compressed(os.path.join(OUT, "reference_schedule.npz"), **{n: np.asarray(getattr(diff, n)) for n in names})
    print("schedule:", {n: np.asarray(getattr(diff, n)).shape for n in names})


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--schedule-only" in sys.argv:
        schedule_tables()
        return
    if "--masked-only" in sys.argv:
        masked()
        return
    if "--chain-only" in sys.argv:
        chain()
        return
    if "--text-only" in sys.argv:
        text()
        return
    dims = synth.SMALL_DIMS
    meta = {"B": B, "T": T, "dims": dims, "weights_seed": 0, "inputs_seed": 0,
            "generator": "oracle/make_goldens.py", "torch": torch.__version__}
    cond, noise = synth.inputs(B, T, seed=0)
    t_mid = torch.full((B,), 500, dtype=torch.long)
    x1, x2 = noise, 0.9 * noise.flip(0)
    arrays = {}
    for mode in (1, 2, 3, 4):
        mix, diff, cfgm = ref_harness.build_reference(dims, mixing_mode=mode, store_influence=True)
        sd = synth.synthetic_state_dict(mix.state_dict(), seed=0)
        mix.load_state_dict(sd, strict=True)
        if mode == 4:
            meta["state_dict_keys"] = {k: list(v.shape) for k, v in mix.state_dict().items()}
            meta["timestep_map"] = list(map(int, diff.timestep_map))
            arrays["alphas_cumprod"] = diff.alphas_cumprod
            arrays["sqrt_recip_alphas_cumprod"] = diff.sqrt_recip_alphas_cumprod
            arrays["sqrt_recipm1_alphas_cumprod"] = diff.sqrt_recipm1_alphas_cumprod
            arrays["alphas_cumprod_prev"] = diff.alphas_cumprod_prev
        with torch.no_grad(), ref_harness.reference_env():
            out = mix(x1, t_mid, cond=cond, mask=None, x2=x2)
            arrays[f"mixer_out_mode{mode}"] = out.numpy()
            arrays[f"mixer_w1_mode{mode}"] = mix.history_influence_i1[-1].expand(B, T, 262).numpy().copy()
            if mode == 4:
                arrays["den_individual"] = mix.denoiser1(x1[..., :262], t_mid, cond=cond[:, 2304:3072]).numpy()
                arrays["den_interaction"] = mix.denoiser2(x2, t_mid, cond=cond[:, :2304]).numpy()
                mix.align = False
                arrays["mixer_out_mode4_noalign"] = mix(x1, t_mid, cond=cond, mask=None, x2=x2).numpy()
                mix.align = True
                # one DDIM step in the middle of the schedule and the last one (index 0)
                for idx in (30, 0):
                    tt = torch.full((B,), idx, dtype=torch.long)
                    o = diff.ddim_sample(cfgm, x1, x2, tt, clip_denoised=False, model_kwargs={"mask": None, "cond": cond})
                    arrays[f"step{idx}_sample"] = o["sample"].numpy()
                    arrays[f"step{idx}_sample2"] = o["sample2"].numpy()
                    arrays[f"step{idx}_pred_xstart2"] = o["pred_xstart2"].numpy()
            # whole 50-step loop from noise
            mix.history_influence_i1, mix.history_influence_i2 = [], []
            loop = diff.ddim_sample_loop(cfgm, (B, T, 524), noise=noise.clone(), clip_denoised=False, progress=False,
                                         model_kwargs={"mask": None, "cond": cond})
            arrays[f"loop_mode{mode}"] = loop.numpy()
            if mode == 4:
                arrays["loop_w1_first"] = mix.history_influence_i1[0].numpy()
                arrays["loop_w1_last"] = mix.history_influence_i1[-1].numpy()
    # motion-space unit vectors from the reference's own functions
    alignment, quaternion, rc = ref_harness.reference_modules()
    g = np.random.default_rng(7)
    m_a = torch.from_numpy(g.normal(size=(3, 9, 262)).astype(np.float32))
    m_b = torch.from_numpy(g.normal(size=(3, 9, 262)).astype(np.float32))
    with torch.no_grad():
        sa, sb = alignment.ih_to_smpl(m_a), alignment.ih_to_smpl(m_b)
        _, al = alignment.align_motions(sa, sb, None)
        arrays["unit_in_a"], arrays["unit_in_b"] = m_a.numpy(), m_b.numpy()
        arrays["unit_roundtrip_a"] = alignment.smpl_to_ih(sa).numpy()
        arrays["unit_aligned_b"] = alignment.smpl_to_ih(al).numpy()
        # batch of 4 for centring: the reference's torch.cross without dim picks a wrong axis when B == 3
        m_c = torch.from_numpy(g.normal(size=(4, 9, 262)).astype(np.float32))
        arrays["unit_in_c"] = m_c.numpy()
        arrays["unit_centered_c"] = alignment.smpl_to_ih(alignment.center_motion(alignment.ih_to_smpl(m_c))).numpy()
    np.savez_compressed(os.path.join(OUT, "reference_small.npz"), **arrays)
    with open(os.path.join(OUT, "reference_small.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", len(arrays), "arrays;", os.path.getsize(os.path.join(OUT, "reference_small.npz")) // 1024, "KiB")
    chain()
    text()
    schedule_tables()
    masked()


if __name__ == "__main__":
    main()
