"""GPU suite: the drop-in's guard rails — stale packed weights, over-stepping the schedule, wrong shapes at the
ctypes boundary, the B == 3 refusal, per-handle error messages, launching for a model on a non-current device."""
import ctypes as C

import pytest
import torch

from helpers import build_mixer, default_schedule, load_golden, relerr
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def golden():
    return load_golden()


def _inputs(meta, B=None):
    B = B or meta["B"]
    cond, noise = synth.inputs(B, meta["T"], seed=meta["inputs_seed"])
    t = torch.full((B,), 500, dtype=torch.long)
    return cond.to(DEV), t.to(DEV), noise.to(DEV), (0.9 * noise.flip(0)).to(DEV)


def test_load_state_dict_after_forward_uses_the_new_weights(golden):
    """forward -> load_state_dict(new) -> forward must equal a fresh model built from `new` (the reference reads live
    parameters; here the packed bf16 copies are rebuilt when the parameter fingerprint changes)"""
    _, meta = golden
    cond, t, x1, x2 = _inputs(meta)
    mix, sd0 = build_mixer(meta["dims"], seed=0)
    mix = mix.to(DEV)
    out0 = mix(x1, t, cond=cond, mask=None, x2=x2).clone()
    den0 = mix.denoiser2(x2, t, cond=cond[:, :2304].contiguous()).clone()
    fresh, sd1 = build_mixer(meta["dims"], seed=7)
    fresh = fresh.to(DEV)
    want = fresh(x1, t, cond=cond, mask=None, x2=x2)
    want_den = fresh.denoiser2(x2, t, cond=cond[:, :2304].contiguous())
    mix.load_state_dict(sd1, strict=True)
    got = mix(x1, t, cond=cond, mask=None, x2=x2)
    got_den = mix.denoiser2(x2, t, cond=cond[:, :2304].contiguous())
    torch.cuda.synchronize()
    assert relerr(out0, want) > 1e-2 and relerr(den0, want_den) > 1e-2     # the two weight sets really differ
    assert torch.equal(got, want) and torch.equal(got_den, want_den)
    with torch.no_grad():                                                  # in-place update (an optimizer step)
        mix.influence.out.bias.add_(0.5)
        fresh.influence.out.bias.add_(0.5)
    assert torch.equal(mix(x1, t, cond=cond, mask=None, x2=x2), fresh(x1, t, cond=cond, mask=None, x2=x2))


def test_stepping_past_the_schedule_is_an_error(golden):
    from mixermdm_b200._lib import MmdmError
    _, meta = golden
    cond, t, x1, x2 = _inputs(meta)
    mix, _ = build_mixer(meta["dims"])
    mix = mix.to(DEV)
    eng = mix.engine(2 * meta["B"], meta["T"], torch.device(DEV), sched=default_schedule(), cfg_scale=3.5, center=True)
    eng.set_condition(cond)
    with pytest.raises(MmdmError, match="no step left"):
        eng.step()                       # no state yet
    eng.set_state(x1, x2, 1)
    eng.step()
    eng.step()                           # index 0: the last one
    with pytest.raises(MmdmError, match="no step left"):
        eng.step()
    # the message is the handle's own
    assert b"no step left" in eng.lib.mmdm_last_error(eng.hd.h)
    # an explicit-batch forward in between re-uses the conditioning tables: stepping without a new set_condition is refused
    eng.set_state(x1, x2, 5)
    mix(torch.cat([x1, x1]), torch.cat([t, t]), cond=torch.cat([cond, cond]), mask=None, x2=torch.cat([x2, x2]))
    with pytest.raises(MmdmError, match="no condition set"):
        eng.step()
    eng.set_condition(cond)
    eng.step()


def test_wrong_shapes_raise_instead_of_reading_out_of_bounds(golden):
    _, meta = golden
    cond, t, x1, x2 = _inputs(meta)
    mix, _ = build_mixer(meta["dims"])
    mix = mix.to(DEV)
    with pytest.raises(ValueError):
        mix(x1[..., :500].contiguous(), t, cond=cond, mask=None, x2=x2)
    with pytest.raises(ValueError):
        mix(x1, t[:1], cond=cond, mask=None, x2=x2)
    with pytest.raises(ValueError):
        mix(x1, t, cond=cond[:, :1000].contiguous(), mask=None, x2=x2)
    with pytest.raises(ValueError):
        mix.denoiser2(x2, t, cond=cond[:, :768].contiguous())
    # like the reference, the interaction denoiser takes the leading 2304 columns of a wider cond
    a = mix.denoiser2(x2, t, cond=cond)
    b = mix.denoiser2(x2, t, cond=cond[:, :2304].contiguous())
    assert torch.equal(a, b)


def test_batch_of_three_is_refused(golden):
    from mixermdm_b200._lib import MmdmError
    from mixermdm_b200.diffusion import MixerDiffusion, get_named_beta_schedule, space_timesteps
    from mixermdm_b200.modules import ClassifierFreeSampleModelX2
    _, meta = golden
    cond, noise = synth.inputs(3, meta["T"], seed=1)
    mix, _ = build_mixer(meta["dims"])
    mix = mix.to(DEV)
    diff = MixerDiffusion(use_timesteps=space_timesteps(1000, "ddim50"), betas=get_named_beta_schedule("cosine", 1000))
    with pytest.raises(MmdmError, match="exactly 3"):
        diff.ddim_sample_loop(ClassifierFreeSampleModelX2(mix, 3.5), (3, meta["T"], 524), noise=noise.to(DEV),
                              clip_denoised=False, model_kwargs={"mask": None, "cond": cond.to(DEV)})


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_model_on_a_non_current_device(golden):
    """a model on cuda:1 while cuda:0 is current: kernels, streams and kernel attributes follow the model's device"""
    _, meta = golden
    cond, t, x1, x2 = _inputs(meta)
    mix0, _ = build_mixer(meta["dims"])
    mix1, _ = build_mixer(meta["dims"])
    mix0, mix1 = mix0.to("cuda:0"), mix1.to("cuda:1")
    torch.cuda.set_device(0)
    a = mix0(x1.to("cuda:0"), t.to("cuda:0"), cond=cond.to("cuda:0"), mask=None, x2=x2.to("cuda:0"))
    b = mix1(x1.to("cuda:1"), t.to("cuda:1"), cond=cond.to("cuda:1"), mask=None, x2=x2.to("cuda:1"))
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    assert b.device == torch.device("cuda:1") and torch.equal(a.cpu(), b.cpu())


def test_programmatic_dependent_launch_is_bitwise_neutral(tmp_path):
    """MMDM_PDL=1 (opt-in: the CTA-pair GEMM and the attention kernel start before their predecessor has finished and wait
    after their prologue) must not change a single bit of a graph-replayed sampler loop at a geometry that runs those
    kernels (256 frames, LN fold on). The switch is read once per process, hence two child processes."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = (
        "import sys, torch\n"
        f"sys.path[:0] = [{root!r}, {os.path.join(root, 'tests')!r}]\n"
        "from helpers import build_mixer, default_schedule\n"
        "from oracle import synth\n"
        "mix, _ = build_mixer(synth.SMALL_DIMS, seed=3)\n"
        "mix = mix.to('cuda')\n"
        "cond, noise = synth.inputs(2, 256, seed=5)\n"
        "eng = mix.engine(4, 256, torch.device('cuda', 0), sched=default_schedule(), cfg_scale=3.5, center=True)\n"
        "out = eng.sample(cond.cuda(), noise.cuda())\n"
        "torch.cuda.synchronize()\n"
        "torch.save(out.cpu(), sys.argv[1])\n")
    outs = []
    for flag in ("0", "1"):
        path = str(tmp_path / f"pdl{flag}.pt")
        env = dict(os.environ, MMDM_PDL=flag)
        r = subprocess.run([sys.executable, "-c", script, path], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(torch.load(path))
    assert torch.isfinite(outs[0]).all()
    assert torch.equal(outs[0], outs[1])
