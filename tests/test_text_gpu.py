"""GPU suite: the text-conditioning path (CLIP text tower + 2-layer text encoders, csrc/text.cu) through the drop-in
`MixerMDM.generate_cond` / `text_process` against the golden produced by the reference's own code, and its kernels
against torch."""
import json
import math
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, relerr
from oracle import synth, text_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model():
    from mixermdm_b200.mixermdm import AttrDict, MixerMDM
    den = AttrDict(NAME="in2IN", NUM_LAYERS=1, NUM_HEADS=2, DROPOUT=0.1, INPUT_DIM=262, LATENT_DIM=128, FF_SIZE=256,
                   ACTIVATION="gelu", CHECKPOINT=None, STRATEGY="ddim50")
    cfg = AttrDict(NAME="MixerMDM", MODEL1=AttrDict(den, NAME="in2INind"), MODEL2=den, ACTIVATION="gelu",
                   MOTION_REP="global", CFG_WEIGHT=3.5, MIXING_MODE=4, DIFFUSION_STEPS=1000, BETA_SCHEDULER="cosine",
                   SAMPLER="uniform", FORCE_INFLUENCE_VAL=None,
                   GENERATOR=AttrDict(NUM_LAYERS=1, NUM_HEADS=8, DROPOUT=0.1, INPUT_DIM=262, LATENT_DIM=512, FF_SIZE=1024),
                   DISCRIMINATOR=AttrDict(NUM_LAYERS=1, NUM_HEADS=4, DROPOUT=0.1, INPUT_DIM=262, LATENT_DIM=64, FF_SIZE=128))
    return MixerMDM(cfg, normalizers=synth.normalizer_stats(), store_influence=False)


@pytest.fixture(scope="module")
def text_model():
    with open(os.path.join(GOLDEN, "reference_text.json")) as f:
        meta = json.load(f)
    model = _model()
    sd = model.state_dict()
    keys = text_oracle.text_keys(sd)
    new = text_oracle.text_state_dict({k: sd[k] for k in keys}, seed=meta["weights_seed"])
    missing, unexpected = model.load_state_dict(new, strict=False)
    assert not unexpected and not any(k in keys for k in missing)
    model.tokenizer = text_oracle.stub_tokenize
    return model.to(DEV).eval(), meta


def test_generate_cond_vs_reference_golden(text_model):
    model, meta = text_model
    gold = np.load(os.path.join(GOLDEN, "reference_text.npz"))["cond"]
    batch = {k: list(v) for k, v in meta["prompts"].items()}
    cond = model.generate_cond(batch)
    torch.cuda.synchronize()
    assert cond.shape == (2, 6144) and cond.is_cuda
    names = ["interaction", "interaction_individual1", "interaction_individual2", "individual_individual1",
             "individual_individual2", "influence_interaction", "influence_individual1", "influence_individual2"]
    for i, n in enumerate(names):
        e = relerr(cond[:, 768 * i:768 * (i + 1)], gold[:, 768 * i:768 * (i + 1)])
        print(f"[parity] cond_{n}: rel L2 err = {e:.3e}")
        assert e < 1e-2, n


def test_text_process_entry_points_agree_with_generate_cond(text_model):
    """the per-field entry points the reference's callers use (cond_pre_func1/2, MixerMDM.text_process) give the same
    vectors as the batched generate_cond (sequences are independent)"""
    model, meta = text_model
    batch = {k: list(v) for k, v in meta["prompts"].items()}
    cond = model.generate_cond(dict(batch))
    b = dict(batch)
    model.cond_pre_func1(b, "individual", "text_individual2", "a")
    model.cond_pre_func2(b, "interaction", "text", "b")
    model.text_process(b, "text_individual1", "c")
    torch.cuda.synchronize()
    assert relerr(b["a"], cond[:, 768 * 4:768 * 5]) < 1e-5
    assert relerr(b["b"], cond[:, :768]) < 1e-5
    assert relerr(b["c"], cond[:, 768 * 6:768 * 7]) < 1e-5


def test_forward_test_from_prompts(text_model):
    """prompts in, motion out: generate_cond feeds the sampler exactly like a precomputed batch['cond']"""
    model, meta = text_model
    batch = {k: list(v) for k, v in meta["prompts"].items()}
    g = torch.Generator().manual_seed(3)
    noise = torch.randn(2, 20, 524, generator=g).to(DEV)
    batch["motion_lens"] = torch.full((2,), 20, dtype=torch.long)
    batch["noise"] = noise.clone()
    out = model.forward_test(batch)["output"]
    cond = model.generate_cond({k: list(v) for k, v in meta["prompts"].items()})
    out2 = model.forward_test({"cond": cond, "motion_lens": batch["motion_lens"], "noise": noise.clone()})["output"]
    torch.cuda.synchronize()
    assert out.shape == (2, 20, 524) and torch.isfinite(out).all() and torch.equal(out, out2)


def test_evaluation_set_from_prompts_mixed_lengths_and_graph_cache(text_model):
    """The evaluation generators' contract (evaluation/datasets.py:64-160) on the real model: prompts of different
    lengths are packed per length, every motion equals what one forward_test per prompt gives, and alternating
    between lengths re-uses the cached step graphs instead of capturing again."""
    from mixermdm_b200.generation import BatchedGenerator, MotionRequest, generate_evaluation_set
    model, meta = text_model
    P = meta["prompts"]
    items = []
    for i, T in enumerate([20, 28, 20, 28]):
        j = i % 2
        items.append((("n",), (P["text"][j],), None, None, torch.tensor([T]), (P["text_individual1"][j],),
                      (P["text_individual2"][j],)))
    gen, mm = generate_evaluation_set(model, items, mm_idxs=[0], mm_num_repeats=2, max_length=32, extended=True)
    assert len(gen) == 4 and len(mm) == 1 and mm[0]["mm_motions"].shape == (2, 32, 2, 262)
    assert gen[1]["motion1"].shape == (32, 262) and np.isfinite(gen[1]["motion1"]).all()
    assert (gen[0]["motion1"][20:] == 0).all() and (gen[1]["motion2"][28:] == 0).all()
    # bitwise: one request with fixed noise through the batched path == the same prompt alone through forward_test
    g = torch.Generator().manual_seed(5)
    reqs = [MotionRequest(None, T, 1, torch.randn(1, T, 524, generator=g).to(DEV), P["text"][j],
                          P["text_individual1"][j], P["text_individual2"][j]) for T, j in [(20, 0), (28, 1), (20, 1)]]
    eng = model.mixing.engine(4, 20, torch.device(DEV))
    outs = BatchedGenerator(model, max_length=32, max_batch=8).generate(reqs)
    before = eng.hd.graph_captures()
    for r, got in zip(reqs, outs):
        one = model.forward_test({"text": [r.text], "text_individual1": [r.text_individual1],
                                  "text_individual2": [r.text_individual2], "noise": r.noise.clone(),
                                  "motion_lens": torch.tensor([r.motion_len])})["output"]
        torch.cuda.synchronize()
        assert torch.equal(got[0, :r.motion_len].reshape(r.motion_len, 524), one[0])
    again = BatchedGenerator(model, max_length=32, max_batch=8).generate(reqs)
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(outs, again))
    eng2 = model.mixing.engine(4, 20, torch.device(DEV))
    assert eng2 is eng and eng.hd.graph_captures() - before <= 2      # B=1 x {20, 28} captured once each, then re-used
