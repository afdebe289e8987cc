"""Batched caller of the sampler: what the reference's evaluation generators do around `model.forward_test`
(reference: src/evaluation/datasets.py:17-163 `EvaluationDatasetInterHuman`, :385-485 `EvaluationDatasetHumanML3D`).

The reference walks the data set with `batch_size=1`: one prompt (repeated `mm_num_repeats` times for the
MultiModality samples) per `forward_test` call, `T = motion_lens[0]` frames, then reshapes the result to
(B, T, 2, 262), de-normalises it (`normalize=True`) and pads it with zero frames to the data set's `max_length`.
Samples are independent, so here requests with the same length are packed into ONE sampler call (the loop's GEMMs
then see B*T*4 rows instead of T*4), and the output formatting runs on the GPU (`mmdm_format_motion`)."""
import collections

import torch

from . import _lib

MotionRequest = collections.namedtuple("MotionRequest", "cond motion_len repeats noise text text_individual1 text_individual2",
                                      defaults=(1, None, None, None, None))
MotionRequest.__doc__ = """One prompt of the evaluation loop.
cond: (6144,) or (repeats, 6144) condition vector(s) (`MixerMDM.generate_cond`), or None when the prompt is given as
text (`text`, and for the extended data sets `text_individual1` / `text_individual2`: the model encodes them);
motion_len: frames to generate (`motion_lens[0]`); repeats: copies of the prompt (`mm_num_repeats` for MultiModality
samples, else 1); noise: optional (repeats, motion_len, 524) initial noise (otherwise the model draws it)."""


def format_generated(output, max_length, mean=None, std=None):
    """(B, T, 524) CUDA tensor -> (B, max_length, 2, 262): `normalizer.backward` (x * std + mean per person) when
    mean/std are given, zero frames beyond T (evaluation/datasets.py:121-134)."""
    if not output.is_cuda:
        raise _lib.MmdmError("format_generated runs on the GPU only (there is no CPU fallback)")
    x = output.detach().to(torch.float32).contiguous()
    B, T, F = x.shape
    if max_length < T:
        raise ValueError(f"max_length {max_length} is shorter than the generated motion ({T} frames)")
    nf = F // 2
    out = torch.empty(B, max_length, F, dtype=torch.float32, device=x.device)
    m = s = None
    if mean is not None:
        m = torch.as_tensor(mean, dtype=torch.float32).to(x.device).contiguous()
        s = torch.as_tensor(std, dtype=torch.float32).to(x.device).contiguous()
        if m.numel() != nf or s.numel() != nf:
            raise ValueError(f"mean/std must hold {nf} values")
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.mmdm_format_motion(_lib.ptr(x), _lib.ptr(out), B, T, int(max_length), nf, _lib.ptr(m),
                                          _lib.ptr(s), _lib.cur_stream()), "format_motion")
    return out.view(B, max_length, 2, nf)


def plan_batches(requests, max_batch):
    """Groups request indices by motion length (first-seen order) and cuts every group into sampler calls of at most
    `max_batch` rows; a request's repeats stay together. Returns [(motion_len, [request index, ...]), ...]."""
    groups = collections.OrderedDict()
    for i, r in enumerate(requests):
        if r.repeats < 1 or r.motion_len < 1:
            raise ValueError(f"request {i}: repeats and motion_len must be positive")
        if r.repeats > max_batch:
            raise ValueError(f"request {i}: {r.repeats} repeats exceed max_batch={max_batch}")
        groups.setdefault(int(r.motion_len), []).append(i)
    plan = []
    for T, idxs in groups.items():
        cur, rows = [], 0
        for i in idxs:
            if rows + requests[i].repeats > max_batch and cur:
                plan.append((T, cur))
                cur, rows = [], 0
            cur.append(i)
            rows += requests[i].repeats
        if cur:
            plan.append((T, cur))
    return plan


class BatchedGenerator:
    """`EvaluationDataset*`'s generation loop with request batching.

    model: a `MixerMDM` (anything with `forward_test(batch) -> {"output": (B, T, 524)}` taking `batch["cond"]`,
    `batch["motion_lens"]` and optionally `batch["noise"]`); max_length: the data set's padded length;
    normalizer: optional (mean, std) of `MotionNormalizer` for `normalize=True`."""

    def __init__(self, model, max_length=300, normalizer=None, max_batch=32, formatter=format_generated):
        self.model = model
        self.max_length = int(max_length)
        self.normalizer = normalizer
        self.max_batch = int(max_batch)
        self.formatter = formatter
        self.calls = 0

    def generate(self, requests):
        """-> list (one entry per request, in request order) of (repeats, max_length, 2, 262) tensors."""
        requests = [r if isinstance(r, MotionRequest) else MotionRequest(*r) for r in requests]
        results = [None] * len(requests)
        mean, std = self.normalizer if self.normalizer is not None else (None, None)
        for T, idxs in plan_batches(requests, self.max_batch):
            conds, noises, have_noise = [], [], [requests[i].noise is not None for i in idxs]
            have_cond = [requests[i].cond is not None for i in idxs]
            if any(have_noise) and not all(have_noise):
                raise ValueError("either every request of a batch carries its initial noise or none does")
            if any(have_cond) and not all(have_cond):
                raise ValueError("either every request of a batch carries its condition vector or every one its text")
            texts = {"text": [], "text_individual1": [], "text_individual2": []}
            for i in idxs:
                r = requests[i]
                if r.cond is not None:
                    c = r.cond if r.cond.dim() == 2 else r.cond.unsqueeze(0).expand(r.repeats, -1)
                    if c.shape[0] != r.repeats:
                        raise ValueError(f"request {i}: cond has {c.shape[0]} rows for {r.repeats} repeats")
                    conds.append(c)
                else:
                    if r.text is None:
                        raise ValueError(f"request {i}: neither a condition vector nor a text prompt")
                    # not extended data sets carry one description: like the reference's non-extended callers, it
                    # stands for the individual ones too
                    for k in texts:
                        texts[k] += [getattr(r, k) if getattr(r, k) is not None else r.text] * r.repeats
                if r.noise is not None:
                    if tuple(r.noise.shape[:2]) != (r.repeats, T):
                        raise ValueError(f"request {i}: noise must be (repeats, motion_len, 524)")
                    noises.append(r.noise)
            cond = torch.cat(conds, dim=0) if conds else None
            noise = torch.cat(noises, dim=0) if noises else None
            rows = cond.shape[0] if cond is not None else len(texts["text"])
            pad = 1 if rows == 3 else 0
            if pad:
                # a sampler call on exactly 3 rows is refused (the reference's center_motion mixes the samples of a
                # batch of 3, engine.MixerEngine._check_batch): samples are independent, so one repeated row is added
                # and dropped again
                cond = torch.cat([cond, cond[-1:]], dim=0) if cond is not None else None
                noise = torch.cat([noise, noise[-1:]], dim=0) if noise is not None else None
                texts = {k: v + v[-1:] for k, v in texts.items()}
            batch = {"motion_lens": torch.full((rows + pad,), T, dtype=torch.long)}
            if cond is not None:
                batch["cond"] = cond
            else:
                batch.update(texts)
            if noise is not None:
                batch["noise"] = noise
            out = self.model.forward_test(batch)["output"][:rows]
            self.calls += 1
            formatted = self.formatter(out, self.max_length, mean, std)
            row = 0
            for i in idxs:
                results[i] = formatted[row:row + requests[i].repeats]
                row += requests[i].repeats
        return results


def generate_evaluation_set(model, items, mm_idxs=(), mm_num_repeats=1, max_length=300, normalizer=None, extended=False,
                            max_batch=32, formatter=format_generated):
    """The generation loop of `EvaluationDatasetInterHuman.__init__` (reference evaluation/datasets.py:64-160) with
    request batching. items: what its DataLoader(batch_size=1) yields - (name, text, motion1, motion2, motion_lens
    [, text_individual1, text_individual2]) with 1-tuples of strings and a (1,) length tensor; the ground-truth motions
    are ignored. mm_idxs: indices generated `mm_num_repeats` times for the MultiModality metric. Returns
    (generated_motions, mm_generated_motions): lists of dicts with the reference's keys, motions as numpy arrays
    (max_length, 262) / (repeats, max_length, 2, 262), `normalizer` = (mean, std) for its normalize=True."""
    mm = set(int(i) for i in mm_idxs)
    reqs, meta = [], []
    for i, data in enumerate(items):
        text, lens = data[1], data[4]
        ti1, ti2 = (data[5], data[6]) if extended else (None, None)
        T = int(lens[0])
        reqs.append(MotionRequest(None, T, mm_num_repeats if i in mm else 1, None, text[0],
                                  ti1[0] if extended else None, ti2[0] if extended else None))
        meta.append((text[0], lens[0], ti1[0] if extended else None, ti2[0] if extended else None, i in mm))
    outs = BatchedGenerator(model, max_length=max_length, normalizer=normalizer, max_batch=max_batch,
                            formatter=formatter).generate(reqs)
    generated, mm_generated = [], []
    for (text, length, ti1, ti2, is_mm), out in zip(meta, outs):
        arr = out.detach().cpu().numpy()
        sub = {"motion1": arr[0, :, 0], "motion2": arr[0, :, 1], "motion_lens": length, "text": text}
        mm_sub = {"mm_motions": arr, "motion_lens": length, "text": text}
        if extended:
            for d in (sub, mm_sub):
                d["text_individual1"], d["text_individual2"] = ti1, ti2
        generated.append(sub)
        if is_mm:
            mm_generated.append(mm_sub)
    return generated, mm_generated
