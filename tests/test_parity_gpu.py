"""GPU parity: the CUDA path (through the drop-in modules / C ABI) against the reference's golden outputs and
the CPU oracle. Tolerance for everything that passes through the bf16 tensor-core GEMMs: 1e-2 relative
(BASELINE.json north_star); fp32-only kernels: 1e-4 in the north star, gated here at 1e-5 / 2e-5 (measured 3e-7 for the
fused mix / centre / DDIM kernel, 4e-6 for the alignment kernel)."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import build_mixer, build_oracle, default_schedule, load_golden, relerr
from oracle import mixer_oracle, synth

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL_BF16 = 1e-2
# Whole-loop tolerance (stated, north_star "final joint positions within a stated tolerance"): 50 chained steps of a
# RANDOM-INIT network are not contractive and guidance multiplies differences by 3.5/2.5 every step, so bf16
# rounding (~3e-3 per step) grows to a few 1e-2; fp32 reference vs fp32 oracle already differ by ~1e-4..2e-3 here.
# The per-step tests with the reference's own x_t fed in are the 1e-2 gate.
LOOP_TOL = 1e-1


def report(name, err):
    print(f"[parity] {name}: rel L2 err = {err:.3e}")
    return err


@pytest.fixture(scope="module")
def golden():
    return load_golden()


def _inputs(meta):
    cond, noise = synth.inputs(meta["B"], meta["T"], seed=meta["inputs_seed"])
    t = torch.full((meta["B"],), 500, dtype=torch.long)
    return cond.to(DEV), noise.to(DEV), t.to(DEV), noise.to(DEV), (0.9 * noise.flip(0)).to(DEV)


def test_library_is_the_product_path():
    from mixermdm_b200 import _lib
    lib = _lib.load()
    assert lib.mmdm_abi_version() == _lib.ABI_VERSION == 4


def test_denoiser_forward_vs_reference_golden(golden):
    arr, meta = golden
    cond, noise, t, x1, x2 = _inputs(meta)
    mix, _ = build_mixer(meta["dims"])
    mix = mix.to(DEV)
    out_i = mix.denoiser1(x1[..., :262].contiguous(), t, cond=cond[:, 2304:3072].contiguous())
    out_I = mix.denoiser2(x2, t, cond=cond[:, :2304].contiguous())
    torch.cuda.synchronize()
    assert report("denoiser individual", relerr(out_i, arr["den_individual"])) < TOL_BF16
    assert report("denoiser interaction", relerr(out_I, arr["den_interaction"])) < TOL_BF16


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
def test_mixer_forward_vs_reference_golden(golden, mode):
    arr, meta = golden
    cond, noise, t, x1, x2 = _inputs(meta)
    mix, _ = build_mixer(meta["dims"], mixing_mode=mode, store_influence=True)
    mix = mix.to(DEV)
    out = mix(x1, t, cond=cond, mask=None, x2=x2)
    torch.cuda.synchronize()
    assert out.shape == (meta["B"], meta["T"], 524)
    assert report(f"mixer forward mode {mode}", relerr(out, arr[f"mixer_out_mode{mode}"])) < TOL_BF16
    assert report(f"influence weights mode {mode}", relerr(mix.history_influence_i1[-1], arr[f"mixer_w1_mode{mode}"])) < TOL_BF16


def test_mixer_forward_noalign_and_train_mode(golden):
    arr, meta = golden
    cond, noise, t, x1, x2 = _inputs(meta)
    mix, sd = build_mixer(meta["dims"], align=False, mode="train")
    mix = mix.to(DEV)
    out, out1, out2 = mix(x1, t, cond=cond, mask=None, x2=x2)
    torch.cuda.synchronize()
    assert relerr(out, arr["mixer_out_mode4_noalign"]) < TOL_BF16
    d = build_oracle(sd, meta["dims"], align=False).mixer(x1.cpu(), x2.cpu(), t.cpu(), cond.cpu(), detail=True)
    assert relerr(out1, d["out1"]) < TOL_BF16 and relerr(out2, d["out2"]) < TOL_BF16


def test_cfg_wrapper_vs_oracle(golden):
    from mixermdm_b200.modules import ClassifierFreeSampleModelX2
    arr, meta = golden
    cond, noise, t, x1, x2 = _inputs(meta)
    mix, sd = build_mixer(meta["dims"])
    mix = mix.to(DEV)
    g = ClassifierFreeSampleModelX2(mix, 3.5)(x1, x2, t, cond=cond, mask=None)
    torch.cuda.synchronize()
    ref = build_oracle(sd, meta["dims"]).guided(x1.cpu(), x2.cpu(), t.cpu(), cond.cpu())
    assert report("cfg wrapper", relerr(g, ref)) < 2 * TOL_BF16      # guidance amplifies the cond/uncond difference by 3.5 / 2.5


@pytest.mark.parametrize("idx", [30, 0])
def test_ddim_step_vs_reference_golden(golden, idx):
    arr, meta = golden
    cond, noise, t, x1, x2 = _inputs(meta)
    mix, _ = build_mixer(meta["dims"])
    mix = mix.to(DEV)
    eng = mix.engine(2 * meta["B"], meta["T"], torch.device(DEV), sched=default_schedule(), cfg_scale=3.5, center=True)
    eng.set_condition(cond)
    eng.set_state(x1, x2, idx)
    final = torch.zeros(meta["B"], meta["T"], 524, device=DEV)
    eng.step(final)
    s1, s2 = eng.get_state()
    torch.cuda.synchronize()
    if idx == 0:
        assert report("ddim step 0 result", relerr(final, arr["step0_pred_xstart2"])) < 2 * TOL_BF16
    else:
        assert report("ddim step 30 stream1", relerr(s1, arr[f"step{idx}_sample"])) < TOL_BF16
        assert report("ddim step 30 stream2", relerr(s2, arr[f"step{idx}_sample2"])) < TOL_BF16


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
def test_full_loop_vs_reference_golden(golden, mode):
    from mixermdm_b200.diffusion import MixerDiffusion, get_named_beta_schedule, space_timesteps
    from mixermdm_b200.modules import ClassifierFreeSampleModelX2
    arr, meta = golden
    cond, noise, *_ = _inputs(meta)
    mix, _ = build_mixer(meta["dims"], mixing_mode=mode, store_influence=(mode == 4))
    mix = mix.to(DEV)
    diff = MixerDiffusion(use_timesteps=space_timesteps(1000, "ddim50"), betas=get_named_beta_schedule("cosine", 1000))
    out = diff.ddim_sample_loop(ClassifierFreeSampleModelX2(mix, 3.5), (meta["B"], meta["T"], 524), noise=noise.clone(),
                                clip_denoised=False, progress=False, model_kwargs={"mask": None, "cond": cond})
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    err = relerr(out, arr[f"loop_mode{mode}"])
    print(f"mode {mode}: 50-step loop relative error vs reference = {err:.3e}")
    assert err < LOOP_TOL
    if mode == 4:
        assert len(mix.history_influence_i1) == 50
        assert relerr(mix.history_influence_i1[0], arr["loop_w1_first"]) < TOL_BF16


def test_graph_and_eager_loops_agree_bitwise(golden):
    arr, meta = golden
    cond, noise, *_ = _inputs(meta)
    mix, _ = build_mixer(meta["dims"])
    mix = mix.to(DEV)
    eng = mix.engine(2 * meta["B"], meta["T"], torch.device(DEV))
    a = eng.sample(cond, noise, use_graph=True).clone()
    b = eng.sample(cond, noise, use_graph=False).clone()
    c = eng.sample(cond, noise, use_graph=True).clone()
    torch.cuda.synchronize()
    assert torch.equal(a, b) and torch.equal(a, c)


def test_batch_partition_invariance_bitwise(golden):
    """samples are independent: a batch of 4 equals two batches of 2 bit for bit (what multi-GPU sharding relies on)"""
    _, meta = golden
    mix, _ = build_mixer(meta["dims"])
    mix = mix.to(DEV)
    cond, noise = synth.inputs(4, 24, seed=3)
    cond, noise = cond.to(DEV), noise.to(DEV)
    eng = mix.engine(8, 24, torch.device(DEV))
    full = eng.sample(cond, noise).clone()
    lo = eng.sample(cond[:2].contiguous(), noise[:2].contiguous()).clone()
    hi = eng.sample(cond[2:].contiguous(), noise[2:].contiguous()).clone()
    torch.cuda.synchronize()
    assert torch.equal(full, torch.cat([lo, hi], 0))


# ---- fp32 kernels against the oracle's motion math -------------------------------------------------
def test_align_kernel_vs_oracle():
    from mixermdm_b200 import _lib
    lib = _lib.load()
    g = np.random.default_rng(5)
    Be, T = 6, 37
    o1 = torch.from_numpy(g.normal(size=(Be, T, 524)).astype(np.float32)).to(DEV)
    o2 = torch.from_numpy(g.normal(size=(Be, T, 524)).astype(np.float32)).to(DEV)
    o1a, o2a = torch.empty_like(o1), torch.empty_like(o2)
    ei = torch.empty(Be * 2 * T, 264, dtype=torch.bfloat16, device=DEV)
    eI = torch.empty_like(ei)
    _lib.check(lib.mmdm_align_motions(_lib.ptr(o1), _lib.ptr(o2), _lib.ptr(o1a), _lib.ptr(o2a), _lib.ptr(ei),
                                      _lib.ptr(eI), Be, T, 264, 1, _lib.cur_stream()), "align")
    torch.cuda.synchronize()
    for p in range(2):
        inter, ind = mixer_oracle.align_pair(o2.cpu()[..., p * 262:(p + 1) * 262], o1.cpu()[..., p * 262:(p + 1) * 262])
        assert report("align: interaction stream", relerr(o2a[..., p * 262:(p + 1) * 262], inter)) < 2e-5
        assert report("align: individual stream", relerr(o1a[..., p * 262:(p + 1) * 262], ind)) < 2e-5
    rows = o1a.view(Be, T, 2, 262).permute(0, 2, 1, 3).reshape(Be * 2 * T, 262)
    assert relerr(ei[:, :262], rows) < 4e-3 and ei[:, 262:].float().abs().max() == 0


# (4, 45): ragged split of the frames over the 4 CTAs of a cluster (12/12/12/9). (40, 300): BASELINE frame count and
# more samples than clusters fit on the device, so the persistent loop and the ring wrap across samples are covered.
# (100, 64): three to four samples per cluster (33 clusters are resident), i.e. every segment transition of the
# software-pipelined frame stream (P1 -> P1, P1 -> P2 of the previous sample, P2 -> P1 two samples on, P2 -> P2).
@pytest.mark.parametrize("shape", [(4, 45), (40, 300), (3, 2), (100, 64)])
@pytest.mark.parametrize("mode", [1, 2, 3, 4])
@pytest.mark.parametrize("final", [False, True])
def test_mix_cfg_ddim_kernel_vs_oracle(mode, final, shape):
    from mixermdm_b200 import _lib
    from mixermdm_b200.engine import MixstepArgs
    lib = _lib.load()
    g = np.random.default_rng(100 + mode)
    B, T = shape
    if B == 3:
        B = 2   # (B == 3 is the reference's torch.cross landmine, SURVEY 8a note 7); T = 2: CTAs without frames
    nw = 1 if mode in (1, 2) else 23
    f32 = lambda *s: torch.from_numpy(g.normal(size=s).astype(np.float32))  # noqa: E731
    hfin = f32(2 * B * 2 * T, 512).to(torch.bfloat16)
    head_w, head_b = f32(nw, 512) * 0.05, f32(nw) * 0.1
    out1a, out2a = f32(2 * B, T, 524), f32(2 * B, T, 524)
    x1, x2 = f32(B, T, 524), f32(B, T, 524)
    (m1, s1), (m2, s2) = synth.normalizer_stats()
    sched = default_schedule()
    step = 0 if final else 17
    # ---- oracle: head -> expand -> mix -> guidance -> process_xstart -> DDIM
    h = hfin.float().view(2 * B, 2, T, 512)
    hp = h.mean(dim=2) if mode in (1, 3) else h
    w = torch.sigmoid(hp @ head_w.t() + head_b)
    outs = []
    for c in range(2):
        per = []
        for p in range(2):
            wp = mixer_oracle.expand_weights(w[c * B:(c + 1) * B, p], T, mode)
            a = out1a[c * B:(c + 1) * B, :, p * 262:(p + 1) * 262]
            b = out2a[c * B:(c + 1) * B, :, p * 262:(p + 1) * 262]
            per.append(b + wp * (a - b))
        outs.append(torch.cat(per, -1))
    gref = 3.5 * outs[0] + (1 - 3.5) * outs[1]
    orc = mixer_oracle.Oracle({}, dict(den=dict(num_heads=1), mix=dict(n_heads=1)), ((m1, s1), (m2, s2)), sched=sched)
    p1, p2 = orc.process_xstart(gref, last=final)
    cr, crm1, ca, ce = [torch.tensor(float(v)) for v in sched.ddim_coefficients()[step]]
    n1 = p1 * ca + ce * ((cr * x1 - p1) / crm1)
    n2 = p2 * ca + ce * ((cr * x2 - p2) / crm1)
    # ---- kernel
    dev = lambda t: t.to(DEV).contiguous()  # noqa: E731
    wdev = torch.zeros(w.shape[0] * w.shape[1] * (1 if mode in (1, 3) else T), 24)
    wdev[:, :nw] = w.reshape(-1, nw)
    keep = dict(weights=dev(wdev), out1a=dev(out1a), out2a=dev(out2a),
                x1=dev(x1), x2=dev(x2), final=torch.zeros(B, T, 524, device=DEV),
                coef=dev(torch.from_numpy(sched.ddim_coefficients())), step=torch.tensor([step], dtype=torch.int32, device=DEV),
                m1=dev(torch.from_numpy(m1)), s1=dev(torch.from_numpy(s1)), m2=dev(torch.from_numpy(m2)),
                s2=dev(torch.from_numpy(s2)), h1=torch.zeros(50, 2 * B, T, 262, device=DEV),
                h2=torch.zeros(50, 2 * B, T, 262, device=DEV))
    a = MixstepArgs()
    a.B, a.T, a.mixing_mode, a.ncfg, a.mix_only, a.center, a.has_force = B, T, mode, 2, 0, 1, 0
    a.force_val, a.cfg_scale, a.nsteps = 0.0, 3.5, 50
    for name, key in [("step_ptr", "step"), ("coef", "coef"), ("weights", "weights"),
                      ("out1a", "out1a"), ("out2a", "out2a"), ("x1", "x1"), ("x2", "x2"), ("final_out", "final"),
                      ("mean1", "m1"), ("std1", "s1"), ("mean2", "m2"), ("std2", "s2"), ("hist_i1", "h1"), ("hist_i2", "h2")]:
        setattr(a, name, keep[key].data_ptr())
    _lib.check(lib.mmdm_mix_cfg_ddim(C.byref(a), _lib.cur_stream()), "mix")
    torch.cuda.synchronize()
    if final:
        assert report(f"mix kernel mode {mode} {shape} final", relerr(keep["final"], gref)) < 1e-5
    else:
        assert report(f"mix kernel mode {mode} {shape} x1", relerr(keep["x1"], n1)) < 1e-5
        assert report(f"mix kernel mode {mode} {shape} x2", relerr(keep["x2"], n2)) < 1e-5
    slot = 50 - 1 - step
    wexp = mixer_oracle.expand_weights(w[:, 0], T, mode).expand(2 * B, T, 262)
    assert relerr(keep["h1"][slot], wexp) < 1e-4


def test_top_level_mixermdm_forward_contract(golden):
    """MixerMDM.forward / forward_test (reference mixermdm.py:490-602): dict keys, history lengths/shapes and the
    sampled motion against the reference's golden loop output."""
    from mixermdm_b200.mixermdm import AttrDict, MixerMDM
    arr, meta = golden
    dd, dm = meta["dims"]["den"], meta["dims"]["mix"]
    den = AttrDict(NAME="in2IN", NUM_LAYERS=dd["num_layers"], NUM_HEADS=dd["num_heads"], DROPOUT=0.1, INPUT_DIM=262,
                   LATENT_DIM=dd["latent_dim"], FF_SIZE=dd["ff_size"], ACTIVATION="gelu", CHECKPOINT=None, STRATEGY="ddim50")
    cfg = AttrDict(NAME="MixerMDM", MODEL1=AttrDict(den, NAME="in2INind"), MODEL2=den, ACTIVATION="gelu",
                   MOTION_REP="global", CFG_WEIGHT=3.5, MIXING_MODE=4, DIFFUSION_STEPS=1000, BETA_SCHEDULER="cosine",
                   SAMPLER="uniform", FORCE_INFLUENCE_VAL=None,
                   GENERATOR=AttrDict(NUM_LAYERS=dm["n_blocks"], NUM_HEADS=dm["n_heads"], DROPOUT=0.1, INPUT_DIM=262,
                                      LATENT_DIM=dm["latent_dim"], FF_SIZE=dm["ff_size"]),
                   DISCRIMINATOR=AttrDict(NUM_LAYERS=1, NUM_HEADS=4, DROPOUT=0.1, INPUT_DIM=262, LATENT_DIM=64, FF_SIZE=128))
    model = MixerMDM(cfg, normalizers=synth.normalizer_stats())
    sd = synth.synthetic_state_dict(model.mixing.state_dict(), seed=meta["weights_seed"])
    model.mixing.load_state_dict(sd, strict=True)       # denoisers are shared objects: model1/2 see the same tensors
    model = model.to(DEV)
    cond, noise, *_ = _inputs(meta)
    B, T = meta["B"], meta["T"]
    batch = {"cond": cond, "noise": noise.clone(), "motion_lens": torch.full((B,), T, dtype=torch.long)}
    out = model.forward_test(dict(batch))
    torch.cuda.synchronize()
    assert set(out) == {"output", "influence_i1", "influence_i2"}
    assert len(out["influence_i1"]) == 50 and out["influence_i1"][0].shape == (2 * B, T, 262)
    assert report("MixerMDM.forward_test loop", relerr(out["output"], arr["loop_mode4"])) < LOOP_TOL
    full = model.forward(dict(batch))
    torch.cuda.synchronize()
    assert set(full) == {"output", "influence_i1", "influence_i2", "out1", "out2", "out_influenced"}
    assert len(full["out1"]) == len(full["out2"]) == len(full["out_influenced"]) == 50
    assert full["out_influenced"][0].shape == (2 * B, T, 524)
    assert torch.equal(full["output"], out["output"])                    # eager eval-mode loop == graph loop
    # first step of the history against the oracle's Mixer.forward on the CFG-doubled batch
    orc = build_oracle(sd, meta["dims"])
    tt = torch.full((2 * B,), 980, dtype=torch.long)
    x = torch.cat([noise, noise]).cpu()
    cc = torch.cat([cond, torch.zeros_like(cond)]).cpu()
    with torch.no_grad():
        d = orc.mixer(x, x, tt, cc, detail=True)
    assert report("eval history out1[0]", relerr(full["out1"][0], d["out1"])) < TOL_BF16
    assert report("eval history out2[0]", relerr(full["out2"][0], d["out2"])) < TOL_BF16
    assert report("eval history out_influenced[0]", relerr(full["out_influenced"][0], d["out"])) < TOL_BF16


def test_sampler_rejects_what_the_reference_path_never_uses(golden):
    from mixermdm_b200.diffusion import MixerDiffusion, get_named_beta_schedule, space_timesteps
    from mixermdm_b200.modules import ClassifierFreeSampleModelX2
    _, meta = golden
    mix, _ = build_mixer(meta["dims"])
    diff = MixerDiffusion(use_timesteps=space_timesteps(1000, "ddim50"), betas=get_named_beta_schedule("cosine", 1000))
    cfgm = ClassifierFreeSampleModelX2(mix.to(DEV), 3.5)
    kw = dict(model_kwargs={"mask": None, "cond": torch.zeros(2, 6144, device=DEV)})
    with pytest.raises(NotImplementedError):
        diff.ddim_sample_loop(cfgm, (2, 24, 524), clip_denoised=True, **kw)
    with pytest.raises(NotImplementedError):
        diff.ddim_sample_loop(cfgm, (2, 24, 524), clip_denoised=False, eta=0.5, **kw)


# Per-step tolerances of the teacher-forced sweep (measured over the 50 steps, tools/diag_chain.py):
#   denoiser outputs (north star: 1e-2 relative for bf16): 2.7e-3 .. 3.1e-3 at every step.
#   Mixer.forward 5.3e-3 .. 1.7e-2: the alignment turns the individual prediction by an angle taken from two frames
#   of the interaction prediction (utils/alignment.py:69-105), which amplifies a 3e-3 perturbation of the inputs -
#   an fp32 perturbation of that size does the same to the reference.
#   Guidance extrapolates: g = 3.5 c - 2.5 u adds the errors of the two halves with weights 3.5 and 2.5 while |g| stays
#   ~|c| (worst case 6x the error of Mixer.forward); the DDIM update is a_t * prediction + b_t * x with a_t growing
#   from ~0 at the noisy end (state error 4e-4 at index 30 of the small goldens) to ~0.8 at the end, so the state
#   inherits the guided error there: below 1e-2 for every index >= 2, 1.9e-2 at index 1.
#   The engine has two LayerNorm code paths: the stand-alone kernel (what this geometry runs by default) and AdaLN folded
#   into the GEMMs (the default from 256 frames on, forced here with MMDM_LN_FOLD=2). The sweep runs in both with the same
#   gates (tools/fold_ab.py: denoisers 2.79e-3 / 3.03e-3 mean in both; Mixer.forward mean 7.0e-3 / 7.2e-3, worst step
#   1.69e-2 / 1.74e-2; state after the last step 1.92e-2 / 2.23e-2). Without the per-row operand offsets of the folded
#   path (gemm.cuh: lnp_moff_in) the last two were 2.3e-2 and 3.2e-2: rounding the operand before the row mean is
#   removed costs accuracy exactly at the ill-conditioned steps.
TOL_MIXER_STEP = 2e-2
TOL_MIXER_MEAN = 1e-2
TOL_GUIDED_STATE = 2.5e-2


@pytest.mark.parametrize("ln_fold", ["kernel", "folded"])
def test_every_ddim_step_teacher_forced_vs_reference_chain(ln_fold, monkeypatch):
    """All 50 DDIM indices, each fed the REFERENCE's own x_t (tests/golden/reference_chain.npz: the reference's
    trajectory, rounded to fp16 between steps so that the stored inputs are exactly what it saw).
    Gates per step: (a) both denoisers' outputs against the fp32 oracle (itself pinned to this chain on the CPU) at the
    north star's 1e-2; (b) Mixer.forward on the CFG-doubled batch (denoisers + alignment + Influence + mix) at
    TOL_MIXER_STEP; (c) the state the step produces against the reference's next x_t at TOL_GUIDED_STATE."""
    import json
    import os
    from helpers import GOLDEN
    monkeypatch.setenv("MMDM_LN_FOLD", "2" if ln_fold == "folded" else "0")   # read when the engine is built
    tol_state = TOL_GUIDED_STATE
    ch = np.load(os.path.join(GOLDEN, "reference_chain.npz"))
    with open(os.path.join(GOLDEN, "reference_chain.json")) as f:
        meta = json.load(f)
    B, T = meta["B"], meta["T"]
    H = meta["dims"]["den"]["num_heads"]
    cond, _ = synth.inputs(B, T, seed=meta["inputs_seed"])
    mix, sd = build_mixer(meta["dims"], mixing_mode=meta["mixing_mode"], seed=meta["weights_seed"])
    mix = mix.to(DEV)
    sdf = {k: v.float() for k, v in sd.items()}
    sched = default_schedule()
    orc = build_oracle(sd, meta["dims"], mixing_mode=meta["mixing_mode"])
    eng = mix.engine(2 * B, T, torch.device(DEV), sched=sched, cfg_scale=3.5, center=True)
    eng.set_condition(cond.to(DEV))
    x1s = torch.from_numpy(ch["x1"].astype(np.float32))
    x2s = torch.from_numpy(ch["x2"].astype(np.float32))
    cc = torch.cat([cond, torch.zeros_like(cond)])
    worst = dict(den=0.0, mixer=0.0, state=0.0)
    mixer_sum = 0.0
    final = torch.zeros(B, T, 524, device=DEV)
    for i in range(50):
        idx = 49 - i
        t = torch.full((2 * B,), int(sched.timestep_map[idx]), dtype=torch.long)
        xx1, xx2 = torch.cat([x1s[i], x1s[i]]), torch.cat([x2s[i], x2s[i]])
        d1 = mix.denoiser1(xx1[..., :262].contiguous().to(DEV), t.to(DEV), cond=cc[:, 2304:3072].contiguous().to(DEV))
        d2 = mix.denoiser2(xx2.to(DEV), t.to(DEV), cond=cc[:, :2304].contiguous().to(DEV))
        net, *_ = eng.mixer_forward(xx1.to(DEV), xx2.to(DEV), t.to(DEV), cc.to(DEV))
        with torch.no_grad():
            r1 = mixer_oracle.denoiser(sdf, "denoiser1.", "individual", xx1[..., :262], t, cc[:, 2304:3072], H)
            r2 = mixer_oracle.denoiser(sdf, "denoiser2.", "interaction", xx2, t, cc[:, :2304], H)
            ref_net = orc.mixer(xx1, xx2, t, cc)
        e_den = max(relerr(d1, r1), relerr(d2, r2))
        e_net = relerr(net, ref_net)
        eng.set_condition(cond.to(DEV))   # (the explicit-batch forwards above re-used the handle's conditioning tables)
        eng.set_state(x1s[i].to(DEV), x2s[i].to(DEV), idx)
        eng.step(final)
        s1, s2 = eng.get_state()
        torch.cuda.synchronize()
        if idx > 0:   # (fp16 storage of the golden adds <= 5e-4)
            e = max(relerr(s1, x1s[i + 1]), relerr(s2, x2s[i + 1]))
        else:         # the last step returns the raw guided prediction
            e = relerr(final, ch["final"])
        worst = dict(den=max(worst["den"], e_den), mixer=max(worst["mixer"], e_net), state=max(worst["state"], e))
        mixer_sum += e_net
        assert e_den < TOL_BF16, f"DDIM index {idx}: denoiser outputs off by {e_den:.3e}"
        assert e_net < TOL_MIXER_STEP, f"DDIM index {idx}: Mixer.forward off by {e_net:.3e}"
        assert e < tol_state, f"DDIM index {idx}: state off by {e:.3e}"
    report("worst per-step denoiser output over the 50 teacher-forced steps", worst["den"])
    report("worst per-step Mixer.forward over the 50 teacher-forced steps", worst["mixer"])
    assert report("mean per-step Mixer.forward over the 50 teacher-forced steps", mixer_sum / 50) < TOL_MIXER_MEAN
    report("worst per-step state vs the reference chain", worst["state"])


@pytest.mark.parametrize("mode", [2, 4])
def test_frame_mask_vs_reference_golden(golden, mode):
    """key-padding masks (in2in.py:410-435, influence.py:107-116) through the engine: both denoisers, Mixer.forward with
    align=False and the CFG wrapper, against outputs of the unmodified reference; align=True + mask raises as the
    reference does; a later unmasked call is unaffected (the mask is per call)."""
    import os
    from mixermdm_b200.modules import ClassifierFreeSampleModelX2
    arr, meta = golden
    ref = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_masked.npz"))
    cond, noise, t, x1, x2 = _inputs(meta)
    mask = torch.from_numpy(ref["mask"]).to(DEV)
    mix, sd = build_mixer(meta["dims"], mixing_mode=mode, align=False, store_influence=True)
    mix = mix.to(DEV)
    out = mix(x1, t, cond=cond, mask=mask, x2=x2)
    torch.cuda.synchronize()
    assert report(f"masked mixer mode {mode}", relerr(out, ref[f"mixer_out_mode{mode}"])) < TOL_BF16
    w1 = mix.history_influence_i1[-1].expand(meta["B"], meta["T"], 262)
    assert report(f"masked weights mode {mode}", relerr(w1, ref[f"mixer_w1_mode{mode}"])) < TOL_BF16
    if mode == 4:
        a = mix.denoiser1(x1[..., :262], t, mask=mask, cond=cond[:, 2304:3072])
        b = mix.denoiser2(x2, t, mask=mask, cond=cond[:, :2304])
        assert report("masked individual denoiser", relerr(a, ref["den_individual"])) < TOL_BF16
        assert report("masked interaction denoiser", relerr(b, ref["den_interaction"])) < TOL_BF16
        b0 = mix.denoiser2(x2, t, mask=None, cond=cond[:, :2304])          # the mask does not stick to the handle
        assert report("unmasked call after a masked one", relerr(b0, arr["den_interaction"])) < TOL_BF16
        # CFG wrapper: doubled batch, one mask row per sample
        orc = build_oracle(sd, meta["dims"], mixing_mode=mode, align=False)
        g = ClassifierFreeSampleModelX2(mix, 3.5)(x1, x2, t, cond=cond, mask=mask)
        want = 3.5 * orc.mixer(x1.cpu(), x2.cpu(), t.cpu(), cond.cpu(), mask=mask.cpu()) - \
            2.5 * orc.mixer(x1.cpu(), x2.cpu(), t.cpu(), torch.zeros_like(cond).cpu(), mask=mask.cpu())
        assert report("masked CFG wrapper", relerr(g, want)) < 2 * TOL_BF16
        aligned, _ = build_mixer(meta["dims"], mixing_mode=mode, align=True)
        with pytest.raises(NotImplementedError):
            aligned.to(DEV)(x1, t, cond=cond, mask=mask, x2=x2)
        # sampler: one masked DDIM step against the oracle, and the graph-replayed masked loop == eager masked steps
        orc.sched = default_schedule()
        eng = mix.engine(2 * meta["B"], meta["T"], torch.device(DEV), sched=orc.sched, cfg_scale=3.5, center=True)
        eng.hd.set_key_mask(mask, meta["T"])
        eng.set_condition(cond)
        eng.set_state(x1, x2, 30)
        eng.step()
        s1, s2 = eng.get_state()
        o = orc.ddim_step(x1.cpu(), x2.cpu(), 30, cond.cpu(), mask=mask.cpu())
        assert report("masked ddim step stream2", relerr(s2, o["sample2"])) < TOL_BF16
        loop_g = eng.sample(cond, noise, mask=mask).clone()
        loop_e = eng.sample(cond, noise, mask=mask, use_graph=False).clone()
        loop_0 = eng.sample(cond, noise)
        assert torch.equal(loop_g, loop_e) and not torch.equal(loop_g, loop_0)


def test_ddpm_step_formula_level(golden):
    """The DDPM variant of the fused update (north star: "DDIM/DDPM update"). The reference's MixerDiffusion cannot run
    p_sample, so parity is formula-level: (1) against the oracle's restatement of p_sample on the two-stream prediction
    (bf16 networks: guided-state tolerance); (2) exactly, against the x0 implied by this engine's own DDIM update of the
    same state: x_ddpm = coef1 x0 + coef2 x_t + sigma z (fp32: 1e-4); (3) the loop entry point is reproducible."""
    arr, meta = golden
    cond, noise, t, x1, x2 = _inputs(meta)
    mix, sd = build_mixer(meta["dims"])
    mix = mix.to(DEV)
    sched = default_schedule()
    eng = mix.engine(2 * meta["B"], meta["T"], torch.device(DEV), sched=sched, cfg_scale=3.5, center=True)
    idx = 30
    eng.set_condition(cond)
    eng.set_state(x1, x2, idx)
    eng.step()
    d1, d2 = eng.get_state()
    eng.set_sampler("ddpm")
    eng.set_state(x1, x2, idx)
    eng.step()
    m1, m2 = eng.get_state()
    z = torch.randn(x1.shape, device=DEV, generator=torch.Generator(device=DEV).manual_seed(5))
    sigma = float(eng.sigma[idx])
    eng.add_state_noise(z, z, sigma)
    s1, s2 = eng.get_state()
    eng.set_sampler("ddim")
    torch.cuda.synchronize()
    # (2) exact relation between the two updates of the same prediction
    cr, crm1, ca, ce = [float(v) for v in sched.ddim_coefficients()[idx]]
    A, Bc = ca - ce / crm1, ce / crm1 * cr
    c1, c2 = float(sched.posterior_mean_coef1[idx]), float(sched.posterior_mean_coef2[idx])
    for d, m, s_, x in ((d1, m1, s1, x1), (d2, m2, s2, x2)):
        x0 = (d.double() - Bc * x.double()) / A
        want = c1 * x0 + c2 * x.double()
        assert report("ddpm mean vs own ddim x0", relerr(m, want.float())) < 1e-4
        assert report("ddpm noise term", relerr(s_ - m, sigma * z)) < 1e-5
    # (1) oracle restatement of p_sample
    orc = build_oracle(sd, meta["dims"])
    orc.sched = sched
    o = orc.ddpm_step(x1.cpu(), x2.cpu(), idx, cond.cpu(), z.cpu())
    assert report("ddpm step stream1 vs oracle", relerr(s1, o["sample"])) < TOL_GUIDED_STATE
    assert report("ddpm step stream2 vs oracle", relerr(s2, o["sample2"])) < TOL_GUIDED_STATE
    # (3) loop entry point: reproducible with given noises, ends on the last step's prediction, leaves DDIM selected
    from mixermdm_b200.diffusion import MixerDiffusion, get_named_beta_schedule, space_timesteps
    from mixermdm_b200.modules import ClassifierFreeSampleModelX2
    diff = MixerDiffusion(use_timesteps=space_timesteps(1000, "ddim50"), betas=get_named_beta_schedule("cosine", 1000))
    zs = torch.randn(49, *noise.shape, device=DEV, generator=torch.Generator(device=DEV).manual_seed(6))
    kw = dict(noise=noise.clone(), clip_denoised=False, model_kwargs={"mask": None, "cond": cond}, step_noise=zs)
    cfgm = ClassifierFreeSampleModelX2(mix, 3.5)
    a = diff.p_sample_loop(cfgm, tuple(noise.shape), **kw)
    b = diff.p_sample_loop(cfgm, tuple(noise.shape), **kw)
    assert torch.isfinite(a).all() and torch.equal(a, b)
    ref = diff.ddim_sample_loop(cfgm, tuple(noise.shape), noise=noise.clone(), clip_denoised=False,
                                model_kwargs={"mask": None, "cond": cond})
    assert report("ddim loop after a ddpm loop", relerr(ref, arr["loop_mode4"])) < LOOP_TOL
    assert not torch.equal(a, ref)


def test_long_sequences_stream_through_the_attention_ring(golden):
    """T > 319: the K/V blocks of a sequence no longer fit the attention kernel's shared-memory ring and stream
    through it (the round-1 kernel stopped at T = 303; the reference's positional table goes to 5000)."""
    _, meta = golden
    T = 400
    cond, noise = synth.inputs(2, T, seed=4)
    t = torch.full((2,), 500, dtype=torch.long)
    mix, sd = build_mixer(meta["dims"])
    mix = mix.to(DEV)
    out_I = mix.denoiser2(noise.to(DEV), t.to(DEV), cond=cond[:, :2304].contiguous().to(DEV))
    out = mix(noise.to(DEV), t.to(DEV), cond=cond.to(DEV), mask=None, x2=(0.9 * noise.flip(0)).to(DEV))
    torch.cuda.synchronize()
    with torch.no_grad():
        ref_I = mixer_oracle.denoiser({k: v.float() for k, v in sd.items()}, "denoiser2.", "interaction", noise, t,
                                      cond[:, :2304], meta["dims"]["den"]["num_heads"])
        ref = build_oracle(sd, meta["dims"]).mixer(noise, 0.9 * noise.flip(0), t, cond)
    assert report("interaction denoiser, T=400", relerr(out_I, ref_I)) < TOL_BF16
    assert report("Mixer.forward, T=400", relerr(out, ref)) < TOL_BF16


# ---- full-size geometry (BASELINE configs): D=1024, 8 layers, T=300 --------------------------------------------
@pytest.fixture(scope="module")
def full_model():
    mix, sd = build_mixer(synth.FULL_DIMS, seed=2)
    return mix.to(DEV), sd


def test_full_size_step_vs_oracle(full_model):
    """config[0] geometry (B=1, 300 frames): one guided Mixer evaluation + DDIM update against the CPU oracle"""
    mix, sd = full_model
    cond, noise = synth.inputs(1, 300, seed=9)
    x1, x2 = noise, 0.8 * noise
    idx = 37
    eng = mix.engine(2, 300, torch.device(DEV), sched=default_schedule(), cfg_scale=3.5, center=True)
    eng.set_condition(cond.to(DEV))
    eng.set_state(x1.to(DEV), x2.to(DEV), idx)
    eng.step()
    s1, s2 = eng.get_state()
    t = torch.full((1,), default_schedule().timestep_map[idx], dtype=torch.long)
    out, *_ = eng.mixer_forward(torch.cat([x1, x1]).to(DEV), torch.cat([x2, x2]).to(DEV), torch.cat([t, t]).to(DEV),
                                torch.cat([cond, torch.zeros_like(cond)]).to(DEV))
    torch.cuda.synchronize()
    orc = build_oracle(sd, synth.FULL_DIMS)
    with torch.no_grad():
        ref = orc.ddim_step(x1, x2, idx, cond)
        ref_mix = orc.mixer(torch.cat([x1, x1]), torch.cat([x2, x2]), torch.cat([t, t]), torch.cat([cond, torch.zeros_like(cond)]))
    assert report("full-size Mixer.forward (Be=2)", relerr(out, ref_mix)) < TOL_BF16
    assert report("full-size DDIM step stream1", relerr(s1, ref["sample"])) < TOL_BF16
    assert report("full-size DDIM step stream2", relerr(s2, ref["sample2"])) < TOL_BF16


def test_full_size_interaction_denoiser_batch64_properties(full_model):
    """config[2] geometry: interaction denoiser alone, batch 64 x 300 frames. Size-independent checks: rows of a
    big batch equal the same rows run in a small batch (bitwise), and a small slice matches the oracle."""
    mix, sd = full_model
    g = torch.Generator().manual_seed(4)
    x = torch.randn(64, 300, 524, generator=g)
    cond = torch.randn(64, 2304, generator=g)
    t = torch.full((64,), 500, dtype=torch.long)
    big = mix.denoiser2(x.to(DEV), t.to(DEV), cond=cond.to(DEV))
    small = mix.denoiser2(x[10:12].contiguous().to(DEV), t[10:12].to(DEV), cond=cond[10:12].contiguous().to(DEV))
    torch.cuda.synchronize()
    assert torch.isfinite(big).all()
    assert torch.equal(big[10:12], small)
    with torch.no_grad():
        ref = mixer_oracle.denoiser({k: v.float() for k, v in sd.items()}, "denoiser2.", "interaction", x[10:11], t[10:11], cond[10:11], 8)
    assert report("full-size interaction denoiser", relerr(big[10:11], ref)) < TOL_BF16


def test_full_size_batch32_loop_properties(full_model):
    """config[1] geometry (B=32 x 300): finite output, run-to-run determinism, and batch-partition invariance
    (rows 0..15 of the B=32 run == a B=16 run of the same samples; row 21 == a B=1 run of that sample), all bitwise.
    Together with test_full_size_step_vs_oracle (B=1 against the oracle) this ties every row of the B=32 benchmark
    configuration to the oracle."""
    mix, _ = full_model
    cond, noise = synth.inputs(32, 300, seed=21)
    cond, noise = cond.to(DEV), noise.to(DEV)
    eng = mix.engine(64, 300, torch.device(DEV))
    a = eng.sample(cond, noise).clone()
    b = eng.sample(cond, noise).clone()
    half = eng.sample(cond[:16].contiguous(), noise[:16].contiguous()).clone()
    one = eng.sample(cond[21:22].contiguous(), noise[21:22].contiguous()).clone()
    torch.cuda.synchronize()
    assert torch.isfinite(a).all()
    assert torch.equal(a, b)
    assert torch.equal(a[:16], half)
    assert torch.equal(a[21:22], one)


@pytest.mark.parametrize("B,T", [(3, 257), (2, 400), (5, 256), (1, 300)])
def test_ln_fold_vs_layernorm_kernel_ragged_full_size(full_model, B, T, monkeypatch):
    """The two LayerNorm code paths of the engine on the full-size interaction denoiser at shapes whose row count is no
    multiple of the 256-row tiles (1542 / 1600 / 2560 rows; two sequences inside a tile at every offset): folded into
    the GEMMs (aux k step of the CTA-pair kernel, T >= 256) against the stand-alone kernel, and one of them against the
    oracle. Both are bf16 roundings of the same function: each within the north star's 1e-2 of the oracle, hence within
    2e-2 of each other (measured ~4e-3)."""
    mix, sd = full_model
    g = torch.Generator().manual_seed(100 + T)
    x = torch.randn(B, T, 524, generator=g)
    cond = torch.randn(B, 2304, generator=g)
    t = torch.full((B,), 321, dtype=torch.long)
    outs = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("MMDM_LN_FOLD", flag)
        mix.denoiser2._standalone = None   # repack: the flag is read when an engine is built
        outs[flag] = mix.denoiser2(x.to(DEV), t.to(DEV), cond=cond.to(DEV)).clone()
    mix.denoiser2._standalone = None
    torch.cuda.synchronize()
    assert torch.isfinite(outs["1"]).all()
    assert report(f"folded vs LayerNorm kernel, B={B} T={T}", relerr(outs["1"], outs["0"])) < 2e-2
    with torch.no_grad():
        ref = mixer_oracle.denoiser({k: v.float() for k, v in sd.items()}, "denoiser2.", "interaction", x[:1], t[:1], cond[:1], 8)
    assert report(f"folded vs oracle, T={T}", relerr(outs["1"][:1], ref)) < TOL_BF16
    assert report(f"LayerNorm kernel vs oracle, T={T}", relerr(outs["0"][:1], ref)) < TOL_BF16
    if B == 1:   # the individual denoiser on one sample: 300 rows, folded GEMMs on the single-CTA kernels (scalar epilogue)
        for flag in ("0", "1"):
            monkeypatch.setenv("MMDM_LN_FOLD", flag)
            mix.denoiser1._standalone = None
            outs[flag] = mix.denoiser1(x[..., :262].contiguous().to(DEV), t.to(DEV), cond=cond[:, :768].contiguous().to(DEV)).clone()
        mix.denoiser1._standalone = None
        with torch.no_grad():
            ref = mixer_oracle.denoiser({k: v.float() for k, v in sd.items()}, "denoiser1.", "individual", x[..., :262], t, cond[:, :768], 8)
        assert report("individual denoiser, 300 rows, folded vs oracle", relerr(outs["1"], ref)) < TOL_BF16
        assert report("individual denoiser, 300 rows, LayerNorm kernel vs oracle", relerr(outs["0"], ref)) < TOL_BF16


@pytest.mark.parametrize("D,H,F", [(256, 2, 512), (512, 8, 2048)])
def test_ln_fold_vs_layernorm_kernel_other_widths(D, H, F, monkeypatch):
    """the folded path at other network widths / head sizes (128 and 64) than the BASELINE's, 263 frames x 3 samples
    (CTA-pair kernels with 4 / 8 statistics slots); tools/fold_sweep.py runs the whole grid"""
    dims = dict(den=dict(latent_dim=D, ff_size=F, num_layers=2, num_heads=H), mix=synth.SMALL_DIMS["mix"])
    mix, sd = build_mixer(dims, seed=5)
    mix = mix.to(DEV)
    g = torch.Generator().manual_seed(D)
    x, cond = torch.randn(3, 263, 524, generator=g), torch.randn(3, 2304, generator=g)
    t = torch.tensor([10, 500, 990])
    outs = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("MMDM_LN_FOLD", flag)
        mix.denoiser2._standalone = None
        outs[flag] = mix.denoiser2(x.to(DEV), t.to(DEV), cond=cond.to(DEV)).clone()
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = mixer_oracle.denoiser({k: v.float() for k, v in sd.items()}, "denoiser2.", "interaction", x, t, cond, H)
    assert report(f"D={D}: folded vs LayerNorm kernel", relerr(outs["1"], outs["0"])) < 1e-2
    assert report(f"D={D}: folded vs oracle", relerr(outs["1"], ref)) < TOL_BF16
    assert report(f"D={D}: LayerNorm kernel vs oracle", relerr(outs["0"], ref)) < TOL_BF16


@pytest.mark.parametrize("mode", [2, 4])
def test_full_size_mixer_forward_batch4_vs_oracle(mode):
    """full geometry (D=1024, L=8, T=300), Be=4 rows, temporal and spatio-temporal weights: Mixer.forward (aligned
    predictions, Influence, mix) against the CPU oracle"""
    mix, sd = build_mixer(synth.FULL_DIMS, mixing_mode=mode, seed=2)
    mix = mix.to(DEV)
    cond, noise = synth.inputs(4, 300, seed=13)
    x1, x2 = noise, 0.85 * noise.flip(0)
    t = torch.tensor([40, 500, 760, 980], dtype=torch.long)
    out = mix(x1.to(DEV), t.to(DEV), cond=cond.to(DEV), mask=None, x2=x2.to(DEV))
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = build_oracle(sd, synth.FULL_DIMS, mixing_mode=mode).mixer(x1, x2, t, cond)
    assert report(f"full-size Mixer.forward Be=4 mode {mode}", relerr(out, ref)) < TOL_BF16


def test_format_motion_matches_numpy():
    """evaluation/datasets.py:121-134: reshape to (B,T,2,262), MotionNormalizer.backward, zero padding to max_length"""
    from mixermdm_b200.generation import format_generated
    g = np.random.default_rng(5)
    x = torch.from_numpy(g.normal(size=(3, 37, 524)).astype(np.float32))
    mean, std = g.normal(size=262).astype(np.float32), (g.random(262) + 0.5).astype(np.float32)
    for m, s in ((None, None), (mean, std)):
        y = format_generated(x.to(DEV), 64, m, s)
        torch.cuda.synchronize()
        ref = x.numpy().reshape(3, 37, 2, 262)
        if m is not None:
            ref = ref * std + mean
        ref = np.concatenate([ref, np.zeros((3, 64 - 37, 2, 262), np.float32)], axis=1)
        assert y.shape == (3, 64, 2, 262)
        assert np.abs(y.cpu().numpy() - ref).max() < 1e-6
    with pytest.raises(ValueError):
        format_generated(x.to(DEV), 10)


def test_batched_generator_equals_one_prompt_per_call(golden):
    """Packing prompts of equal length into one sampler call (generation.BatchedGenerator) returns bit for bit what the
    reference's batch_size=1 evaluation loop gets from one forward_test per prompt (samples are independent)."""
    from mixermdm_b200.generation import BatchedGenerator, MotionRequest, format_generated
    from mixermdm_b200.mixermdm import AttrDict, MixerMDM
    _, meta = golden
    dd, dm = meta["dims"]["den"], meta["dims"]["mix"]
    den = AttrDict(NAME="in2IN", NUM_LAYERS=dd["num_layers"], NUM_HEADS=dd["num_heads"], DROPOUT=0.1, INPUT_DIM=262,
                   LATENT_DIM=dd["latent_dim"], FF_SIZE=dd["ff_size"], ACTIVATION="gelu", CHECKPOINT=None, STRATEGY="ddim50")
    cfg = AttrDict(NAME="MixerMDM", MODEL1=AttrDict(den, NAME="in2INind"), MODEL2=den, ACTIVATION="gelu",
                   MOTION_REP="global", CFG_WEIGHT=3.5, MIXING_MODE=4, DIFFUSION_STEPS=1000, BETA_SCHEDULER="cosine",
                   SAMPLER="uniform", FORCE_INFLUENCE_VAL=None,
                   GENERATOR=AttrDict(NUM_LAYERS=dm["n_blocks"], NUM_HEADS=dm["n_heads"], DROPOUT=0.1, INPUT_DIM=262,
                                      LATENT_DIM=dm["latent_dim"], FF_SIZE=dm["ff_size"]),
                   DISCRIMINATOR=AttrDict(NUM_LAYERS=1, NUM_HEADS=4, DROPOUT=0.1, INPUT_DIM=262, LATENT_DIM=64, FF_SIZE=128))
    model = MixerMDM(cfg, normalizers=synth.normalizer_stats(), store_influence=False)
    model.mixing.load_state_dict(synth.synthetic_state_dict(model.mixing.state_dict(), seed=meta["weights_seed"]), strict=True)
    model = model.to(DEV)
    g = torch.Generator().manual_seed(11)
    lens, reps = [24, 17, 24, 24, 17], [1, 2, 4, 1, 1]      # T=17: 2 + 1 = 3 rows, padded to 4 inside the generator
    reqs = [MotionRequest(torch.randn(6144, generator=g).to(DEV), T, r, torch.randn(r, T, 524, generator=g).to(DEV))
            for T, r in zip(lens, reps)]
    (m1, s1), _ = synth.normalizer_stats()
    gen = BatchedGenerator(model, max_length=32, normalizer=(m1, s1), max_batch=8)
    res = gen.generate(reqs)
    assert gen.calls == 2                                   # lengths {24, 17}: two sampler calls for five prompts
    for r, got in zip(reqs, res):
        one = model.forward_test({"cond": r.cond.unsqueeze(0).expand(r.repeats, -1).contiguous(), "noise": r.noise.clone(),
                                  "motion_lens": torch.full((r.repeats,), r.motion_len, dtype=torch.long)})["output"]
        want = format_generated(one, 32, m1, s1)
        torch.cuda.synchronize()
        assert got.shape == (r.repeats, 32, 2, 262)
        assert torch.equal(got, want)
