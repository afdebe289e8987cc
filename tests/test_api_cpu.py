"""CPU suite: the drop-in boundary — state-dict layout, C-ABI symbol export, host tables, loud failure without CUDA,
and the world_size-2 (gloo) sharding logic."""
import ctypes
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

from helpers import ROOT, build_mixer, load_golden
from oracle import ref_harness, synth


def test_mixer_state_dict_layout_matches_reference_golden():
    _, meta = load_golden()
    mix, _ = build_mixer(meta["dims"])
    ours = {k: list(v.shape) for k, v in mix.state_dict().items()}
    assert ours == meta["state_dict_keys"]          # same names, same order, same shapes as the reference Mixer


def test_positional_table_matches_reference_buffer():
    if not ref_harness.available():
        pytest.skip("reference not mounted")
    from mixermdm_b200.modules import PositionalEncoding
    with ref_harness.reference_env():
        from models.utils.utils import PositionalEncoding as RefPE
        ref = RefPE(512, dropout=0).pe
    assert torch.equal(PositionalEncoding(512).pe, ref)


@pytest.mark.skipif(not ref_harness.available(), reason="reference sources only exist in the build container")
def test_discriminator_and_denoiser_layouts_match_live_reference():
    from mixermdm_b200.mixermdm import DiscriminatorTransfomer
    from mixermdm_b200.modules import in2INDenoiser
    with ref_harness.reference_env():
        from models.in2in import in2INDenoiser as RefDen
        from models.utils.discriminators import DiscriminatorTransfomer as RefDisc
        rd = RefDisc(input_feats=262, latent_dim=256, num_frames=300, ff_size=512, num_layers=2, num_heads=4)
        rn = RefDen(262, mode="individual", latent_dim=256, ff_size=512, num_layers=2, num_heads=2)
    od = DiscriminatorTransfomer(input_feats=262, latent_dim=256, num_frames=300, ff_size=512, num_layers=2, num_heads=4)
    on = in2INDenoiser(262, mode="individual", latent_dim=256, ff_size=512, num_layers=2, num_heads=2)
    for r, o in ((rd, od), (rn, on)):
        assert [(k, tuple(v.shape)) for k, v in r.state_dict().items()] == [(k, tuple(v.shape)) for k, v in o.state_dict().items()]


def test_library_exports_every_declared_symbol():
    from mixermdm_b200 import _build, _lib
    if not os.path.exists(_build.LIB_PATH):
        _build.build()
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "mixermdm_b200.h")).read()
    names = sorted(set(re.findall(r"\b(mmdm_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mixermdm_b200.h but not exported"
    assert lib.mmdm_abi_version() == 4
    assert ctypes.sizeof(_lib.Epilogue) % 8 == 0


def test_topology_of_top_level_model_and_loud_failure_without_cuda():
    from mixermdm_b200.mixermdm import AttrDict, MixerMDM
    den = AttrDict(NAME="in2IN", NUM_LAYERS=1, NUM_HEADS=2, DROPOUT=0.1, INPUT_DIM=262, LATENT_DIM=128, FF_SIZE=256,
                   ACTIVATION="gelu", CHECKPOINT=None, STRATEGY="ddim50")
    ind = AttrDict(den, NAME="in2INind")
    cfg = AttrDict(NAME="MixerMDM", MODEL1=ind, MODEL2=den, ACTIVATION="gelu", MOTION_REP="global", CFG_WEIGHT=3.5,
                   MIXING_MODE=4, DIFFUSION_STEPS=1000, BETA_SCHEDULER="cosine", SAMPLER="uniform",
                   FORCE_INFLUENCE_VAL=None,
                   GENERATOR=AttrDict(NUM_LAYERS=1, NUM_HEADS=8, DROPOUT=0.1, INPUT_DIM=262, LATENT_DIM=512, FF_SIZE=1024),
                   DISCRIMINATOR=AttrDict(NUM_LAYERS=1, NUM_HEADS=4, DROPOUT=0.1, INPUT_DIM=262, LATENT_DIM=64, FF_SIZE=128))
    m = MixerMDM(cfg, normalizers=synth.normalizer_stats())
    keys = list(m.state_dict().keys())
    for prefix in ("model1.decoder.net_individual.", "model2.decoder.net_interaction.", "denoiser1.", "denoiser2.",
                   "mixing.denoiser1.", "mixing.influence.blocks.0.", "mixing.influence.out.", "discriminator_i.",
                   "discriminator_I.", "clipTransEncoder.layers.1.", "clip_ln.", "clip_transformer.resblocks.11.",
                   "token_embedding.", "ln_final.", "model1.clipTransEncoder_individual.", "model2.clip_ln_interaction."):
        assert any(k.startswith(prefix) for k in keys), prefix
    assert "positional_embedding" in keys and m.mixing_mode == 4 and m.mixing.mode == "train"
    assert m.mixing.denoiser1 is m.denoiser1 is m.model1.decoder.net_individual
    if not torch.cuda.is_available():
        from mixermdm_b200._lib import MmdmError
        batch = {"cond": torch.zeros(1, 6144), "motion_lens": torch.tensor([16])}
        with pytest.raises((MmdmError, RuntimeError, AssertionError)):
            m.forward_test(batch)          # no CPU fallback: the CUDA engine is the only path
    with pytest.raises(NotImplementedError):
        m.compute_loss({})


def test_shard_bounds_cover_batch_exactly():
    from mixermdm_b200.distributed import shard_bounds
    for total in (0, 1, 7, 32, 1024):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, T, outdir):
    import torch.distributed as dist
    from mixermdm_b200.distributed import sample_sharded
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    cond = torch.arange(total * 6144, dtype=torch.float32).reshape(total, 6144) / 1000.0

    def fake_loop(c, n):          # stands in for the per-rank DDIM loop: any per-sample function of (cond, noise)
        return n * 2.0 + c[:, :1, None] + n.flip(1).cumsum(1) * 0.01
    out = sample_sharded(fake_loop, cond, T, base_seed=17)
    np.save(os.path.join(outdir, f"rank{rank}.npy"), out.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [5, 8])
def test_sharded_sampling_equals_single_process_bitwise(total):
    import torch.multiprocessing as mp
    from mixermdm_b200.distributed import sample_sharded, seeded_noise
    T = 6
    cond = torch.arange(total * 6144, dtype=torch.float32).reshape(total, 6144) / 1000.0
    ref = sample_sharded(lambda c, n: n * 2.0 + c[:, :1, None] + n.flip(1).cumsum(1) * 0.01, cond, T, base_seed=17)
    assert torch.equal(seeded_noise([3], T, 17)[0], seeded_noise(range(5), T, 17)[3])
    import tempfile
    ctx = mp.get_context("spawn")
    port = _free_port()
    with tempfile.TemporaryDirectory() as d:
        procs = [ctx.Process(target=_worker, args=(r, 2, port, total, T, d)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=180)
            assert p.exitcode == 0
        for r in range(2):
            got = torch.from_numpy(np.load(os.path.join(d, f"rank{r}.npy")))
            assert torch.equal(got, ref)      # both ranks hold the full batch, identical to the 1-process result


def test_gaussian_taps_match_scipy_impulse_response():
    """postprocess.gaussian_taps restates scipy's kernel (scripts/infer/mixermdm.py:131 uses sigma = 1)."""
    import numpy as np
    from scipy.ndimage import gaussian_filter1d
    from mixermdm_b200.postprocess import gaussian_taps
    for sigma in (0.5, 1.0, 2.0):
        taps, radius = gaussian_taps(sigma)
        imp = np.zeros(4 * radius + 1)
        imp[2 * radius] = 1.0
        ref = gaussian_filter1d(imp, sigma, mode="nearest")[radius:3 * radius + 1]
        assert radius == int(4 * sigma + 0.5) and len(taps) == 2 * radius + 1
        assert np.abs(taps - ref).max() < 1e-15


def _mix_stream_reference(S, nfr, has2):
    """the fused mix kernel's frame stream, spelled out: P1(0), [P1(k), P2(k-1)] for k = 1..S-1, P2(S-1)"""
    items = [(1, 0, f) for f in range(nfr)]
    for k in range(1, S + 1):
        if k < S:
            items += [(1, k, f) for f in range(nfr)]
        if has2:
            items += [(2, k - 1, f) for f in range(nfr)]
    return items


def test_mix_kernel_stream_arithmetic_matches_enumeration():
    """csrc/mix_stream.cuh through mmdm_debug_mix_stream (host code, no GPU): decoding an item, its index from
    (phase, sample, frame), its ring slot / lap (reciprocal multiply) and the item that takes over its slot
    ring_slots positions later, for every item of many stream shapes - incl. one-slot rings, one sample per
    cluster, no phase 2 and 3+ samples, where all four segment transitions occur."""
    import ctypes as C
    from mixermdm_b200 import _lib
    lib = _lib.load()
    out = (C.c_int32 * 10)()
    for S in (1, 2, 3, 4, 5):
        for nfr in (1, 2, 3, 12, 25, 76):
            for has2 in (0, 1):
                ref = _mix_stream_reference(S, nfr, has2)
                for NS in sorted({1, 2, min(nfr, 25), nfr}):
                    if NS > nfr:
                        continue
                    for c, (ph, n, f) in enumerate(ref):
                        assert lib.mmdm_debug_mix_stream(S, nfr, has2, NS, c, out) == 0
                        assert (out[0], out[1], out[2]) == (ph, n, f), (S, nfr, has2, NS, c)
                        assert (out[3], out[4]) == (c % NS, c // NS)
                        assert out[5] == c
                        if c + NS < len(ref):
                            assert out[6] == 1 and (out[7], out[8], out[9]) == ref[c + NS], (S, nfr, has2, NS, c)
                        else:
                            assert out[6] == 0
    assert lib.mmdm_debug_mix_stream(2, 4, 1, 5, 0, out) != 0          # ring deeper than a sample's frames: rejected
    assert lib.mmdm_debug_mix_stream(2, 4, 1, 4, 16, out) != 0         # item outside the stream


def test_attrdict_behaves_like_a_config_node():
    import copy
    from mixermdm_b200.mixermdm import AttrDict
    cfg = AttrDict(A=1, SUB=AttrDict(B=2))
    assert cfg.A == 1 and cfg.SUB.B == 2 and "A" in cfg
    assert not hasattr(cfg, "MISSING")               # AttributeError, not KeyError
    with pytest.raises(AttributeError):
        cfg.MISSING
    c2 = copy.deepcopy(cfg)
    c2.SUB.B = 3
    assert cfg.SUB.B == 2 and c2.SUB.B == 3


def test_weight_fingerprint_tracks_parameter_changes_and_modules_copy():
    """the packed-weight cache is keyed on this fingerprint (stale weights after load_state_dict)"""
    import copy
    from mixermdm_b200.engine import weights_fingerprint
    _, meta = load_golden()
    mix, sd = build_mixer(meta["dims"])
    f0 = weights_fingerprint(mix)
    assert weights_fingerprint(mix) == f0
    mix.load_state_dict(sd)                                   # same values, but copy_() bumps the version counters
    f1 = weights_fingerprint(mix)
    assert f1 != f0
    with torch.no_grad():
        next(mix.influence.parameters()).add_(1.0)            # in-place update (an optimizer step)
    f2 = weights_fingerprint(mix)
    assert f2 != f1
    mix.double()                                              # _apply replaces the storages
    assert weights_fingerprint(mix) != f2
    mix._engine = object()                                    # stands in for the ctypes-holding engine
    m2 = copy.deepcopy(mix)
    assert m2._engine is None and m2.denoiser1._standalone is None
    assert list(m2.state_dict().keys()) == list(mix.state_dict().keys())


def test_motion_stats_forward_backward_and_history_view():
    from mixermdm_b200.engine import MotionStats
    from mixermdm_b200.modules import history_view
    mean, std = np.linspace(-1, 1, 262, dtype=np.float32), np.linspace(0.5, 1.5, 262, dtype=np.float32)
    n = MotionStats(mean, std)
    x = torch.randn(2, 5, 262)
    assert torch.allclose(n.backward(n.forward(x)), x, atol=1e-5)
    assert torch.allclose(n.forward(x), (x - torch.from_numpy(mean)) / torch.from_numpy(std))
    w = torch.rand(4, 5, 262)
    assert history_view(w, 1).shape == (4, 5, 1) and history_view(w, 2).shape == (4, 5, 1)
    assert history_view(w, 3).shape == (4, 5, 262) and history_view(w, 4) is w


def test_missing_submodel_checkpoint_is_an_error():
    from mixermdm_b200.mixermdm import AttrDict, MixerMDM
    den = AttrDict(NAME="in2IN", NUM_LAYERS=1, NUM_HEADS=2, DROPOUT=0.1, INPUT_DIM=262, LATENT_DIM=128, FF_SIZE=256,
                   ACTIVATION="gelu", CHECKPOINT="/nonexistent/in2in.ckpt", STRATEGY="ddim50")
    ind = AttrDict(den, NAME="in2INind")
    cfg = AttrDict(NAME="MixerMDM", MODEL1=ind, MODEL2=den, ACTIVATION="gelu", MOTION_REP="global", CFG_WEIGHT=3.5,
                   MIXING_MODE=4, DIFFUSION_STEPS=1000, BETA_SCHEDULER="cosine", SAMPLER="uniform",
                   FORCE_INFLUENCE_VAL=None, INPUT_DIM=262, LATENT_DIM=512, FF_SIZE=1024, NUM_LAYERS=1, NUM_HEADS=8,
                   DROPOUT=0.1)
    with pytest.raises(FileNotFoundError):
        MixerMDM(cfg, normalizers=synth.normalizer_stats())
    MixerMDM(cfg, normalizers=synth.normalizer_stats(), load_submodel_checkpoints=False)


def test_last_error_is_per_thread_without_a_handle():
    from mixermdm_b200 import _lib
    lib = _lib.load()
    assert lib.mmdm_mix_cfg_ddim(None, None) != 0                 # null argument: fails before touching a device
    assert b"null argument" in lib.mmdm_last_error(None)
    import threading
    seen = []
    t = threading.Thread(target=lambda: seen.append(lib.mmdm_last_error(None)))
    t.start(); t.join()
    assert seen == [b""]                                          # another thread has its own (empty) slot
