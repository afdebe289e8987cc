"""Host logic of the batched generation caller (reference: evaluation/datasets.py:17-163): request grouping, call
packing and result order, with a stub model and a CPU stand-in for the GPU formatter."""
import pytest
import torch

from mixermdm_b200.generation import BatchedGenerator, MotionRequest, plan_batches


class _StubModel:
    """forward_test returns output[b, t, f] = cond[b, 0] + t (+ noise), so every row is traceable to its request"""

    def __init__(self):
        self.batches = []

    def forward_test(self, batch):
        cond, T = batch["cond"], int(batch["motion_lens"][0])
        assert batch["motion_lens"].shape[0] == cond.shape[0] and bool((batch["motion_lens"] == T).all())
        self.batches.append((cond.shape[0], T))
        out = cond[:, :1, None] + torch.arange(T, dtype=torch.float32)[None, :, None] + torch.zeros(1, 1, 524)
        if "noise" in batch:
            out = out + batch["noise"]
        return {"output": out}


def _cpu_formatter(output, max_length, mean, std):
    B, T, F = output.shape
    out = torch.zeros(B, max_length, 2, F // 2)
    y = output.reshape(B, T, 2, F // 2)
    if mean is not None:
        y = y * torch.as_tensor(std) + torch.as_tensor(mean)
    out[:, :T] = y
    return out


def _req(tag, T, repeats=1, noise=False):
    cond = torch.full((6144,), float(tag))
    n = torch.full((repeats, T, 524), 0.25) if noise else None
    return MotionRequest(cond, T, repeats, n)


def test_plan_groups_by_length_and_respects_max_batch():
    reqs = [_req(0, 40), _req(1, 77), _req(2, 40, repeats=3), _req(3, 40), _req(4, 77), _req(5, 40, repeats=2)]
    plan = plan_batches(reqs, max_batch=4)
    assert plan == [(40, [0, 2]), (40, [3, 5]), (77, [1, 4])]
    assert plan_batches(reqs, max_batch=32) == [(40, [0, 2, 3, 5]), (77, [1, 4])]
    with pytest.raises(ValueError):
        plan_batches([_req(0, 40, repeats=5)], max_batch=4)
    with pytest.raises(ValueError):
        plan_batches([_req(0, 0)], max_batch=4)


def test_results_come_back_in_request_order_padded_and_denormalised():
    reqs = [_req(10, 40), _req(20, 77, repeats=2), _req(30, 40, repeats=3), _req(40, 5)]
    model = _StubModel()
    mean, std = torch.arange(262, dtype=torch.float32), torch.full((262,), 2.0)
    gen = BatchedGenerator(model, max_length=100, normalizer=(mean, std), max_batch=8, formatter=_cpu_formatter)
    res = gen.generate(reqs)
    assert model.batches == [(4, 40), (2, 77), (1, 5)] and gen.calls == 3      # 7 rows in 3 sampler calls, not 4
    for r, out in zip(reqs, res):
        assert out.shape == (r.repeats, 100, 2, 262)
        t = torch.arange(r.motion_len, dtype=torch.float32)
        want = (float(r.cond[0]) + t)[None, :, None, None] * 2.0 + mean
        assert torch.equal(out[:, :r.motion_len], want.expand(r.repeats, -1, 2, -1))
        assert bool((out[:, r.motion_len:] == 0).all())


def test_noise_travels_with_its_request_and_mixed_batches_are_rejected():
    model = _StubModel()
    gen = BatchedGenerator(model, max_length=50, formatter=_cpu_formatter)
    res = gen.generate([_req(1, 50, noise=True), _req(2, 50, repeats=2, noise=True)])
    assert torch.allclose(res[1][:, 3, 0, 0], torch.full((2,), 2 + 3 + 0.25))
    with pytest.raises(ValueError):
        gen.generate([_req(1, 50, noise=True), _req(2, 50)])


def test_gpu_formatter_refuses_cpu_tensors():
    from mixermdm_b200 import _lib
    from mixermdm_b200.generation import format_generated
    with pytest.raises(_lib.MmdmError):
        format_generated(torch.zeros(1, 4, 524), 8)


class _StubTextModel:
    """forward_test on text batches: output[b] = hash of the three prompts of row b (so rows are traceable)"""

    def __init__(self):
        self.calls = []

    def forward_test(self, batch):
        T = int(batch["motion_lens"][0])
        rows = len(batch["text"])
        assert len(batch["text_individual1"]) == rows == len(batch["text_individual2"]) == batch["motion_lens"].shape[0]
        assert rows != 3                                     # the generator never issues a 3-row sampler call
        self.calls.append((rows, T))
        tag = torch.tensor([float(len(a) + 10 * len(b) + 100 * len(c)) for a, b, c in
                            zip(batch["text"], batch["text_individual1"], batch["text_individual2"])])
        return {"output": tag[:, None, None] + torch.zeros(rows, T, 524)}


def test_evaluation_set_adapter_follows_the_reference_contract():
    """evaluation/datasets.py:64-160: DataLoader(batch_size=1) items in, the reference's dict lists out"""
    from mixermdm_b200.generation import generate_evaluation_set
    items = []
    for i, (T, txt) in enumerate([(24, "aa"), (17, "bbbb"), (24, "c"), (24, "dddddd"), (17, "ee")]):
        items.append((("name%d" % i,), (txt,), None, None, torch.tensor([T]), (txt + "1",), (txt + "22",)))
    model = _StubTextModel()
    generated, mm = generate_evaluation_set(model, items, mm_idxs=[1], mm_num_repeats=2, max_length=32,
                                            extended=True, max_batch=8, formatter=_cpu_formatter)
    assert sorted(model.calls) == [(4, 17), (4, 24)]      # T=17: 2 + 1 = 3 rows padded to 4; T=24: 3 prompts + ... = 3 -> 4
    assert len(generated) == 5 and len(mm) == 1
    for it, g in zip(items, generated):
        txt = it[1][0]
        tag = len(txt) + 10 * (len(txt) + 1) + 100 * (len(txt) + 2)
        T = int(it[4][0])
        assert g["text"] == txt and g["text_individual1"] == txt + "1" and int(g["motion_lens"]) == T
        assert g["motion1"].shape == (32, 262) and g["motion2"].shape == (32, 262)
        assert (g["motion1"][:T] == tag).all() and (g["motion1"][T:] == 0).all()
    assert mm[0]["mm_motions"].shape == (2, 32, 2, 262) and mm[0]["text"] == "bbbb"
