#!/usr/bin/env python
"""Benchmark of the MixerMDM denoising loop (BASELINE.json metric: motion samples/sec, 300 frames, 50 DDIM steps).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4|5]

One "step" = one complete 50-step DDIM sampling loop over the batch.
  config 2 (default at N=1): BASELINE configs[1], B=32 x 300 frames on one GPU, spatio-temporal Mixer, bf16 GEMMs.
  config 5 (default at N>1): BASELINE configs[4], global B=1024 split N ways (strong scaling), one NCCL all_gather.
  config 3 / 4             : the interaction denoiser alone at B=64 / the fused mix+DDIM kernel sweep at B=128; they
                             are also carried as sub-records ("config3", "config4") of the default N=1 line.
Prints ONE JSON line (rank 0).
`value`   : samples/s with cond + noise already resident in HBM, timed with CUDA events, max over ranks.
`e2e`     : the same loop through the public API (distributed.sample_sharded -> MixerDiffusion.ddim_sample_loop) with
            pinned HOST inputs: H2D of cond+noise and D2H of the result inside the timed region.
`roofline`: the dominant kernel of the step, the QKV-class projection GEMM (tensor-bound) - algorithmic FLOPs / CUDA-event
            time per launch, measured in a separate eager pass with events around every launch of the same step
            (per-launch events cannot be recorded inside a CUDA-graph replay). `roofline_all_gemms`: the same over every
            tcgen05 GEMM launch of the step (the quantity rounds 1 / 2a reported as `roofline`); `roofline_resid`: the
            HBM-bound residual GEMMs against the HBM peak; `roofline_mix`: the fused HBM-bound mix/sampler kernel.
`cpu_baseline` / `--impl reference`: the CPU port of the reference path (oracle/, torch fp32 on all host cores; the
            Python reference itself cannot travel to the GPU box) on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

FULL_DIMS = dict(den=dict(latent_dim=1024, ff_size=2048, num_layers=8, num_heads=8),
                 mix=dict(latent_dim=512, ff_size=1024, n_blocks=4, n_heads=8))
NSTEPS = 50
METRIC = "motion samples/sec (300f, 50 DDIM steps)"
# algorithmic work per sample per DDIM step (CFG included); interaction denoiser alone, no CFG
GFLOP_PER_SAMPLE_STEP = 478.5
GFLOP_INTERACTION_FWD = 133.55
CONFIG5_GLOBAL_BATCH = 1024


def synthetic_stats():
    rng = np.random.default_rng(0)
    pairs = []
    for _ in range(2):
        pairs.append(((0.1 * rng.normal(size=262)).astype(np.float32), rng.uniform(0.5, 1.5, size=262).astype(np.float32)))
    glob, hml = pairs
    return hml, glob


def _reinit_zero_modules(module):
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in module.parameters():
            if p.abs().sum() == 0:
                p.normal_(0, 0.02, generator=g)


def build_model(device, mixing_mode=4, dims=FULL_DIMS):
    """random-init weights of the named architecture (seed 0); zero-initialised linears re-drawn N(0, 0.02^2)
    (seed 1) so the networks do not output exact zeros (SURVEY.md 8a note 1)."""
    from mixermdm_b200 import modules
    torch.manual_seed(0)
    kw = dict(dropout=0.1, **dims["den"])
    d1 = modules.in2INDenoiser(262, mode="individual", **kw)
    d2 = modules.in2INDenoiser(262, mode="interaction", **kw)
    mix = modules.Mixer(d1, d2, nfeats=262, text_dim=768, mixing_mode=mixing_mode, store_influence=False,
                        mode="eval_intermediate", align=True, normalizers=synthetic_stats(), **dims["mix"]).eval()
    _reinit_zero_modules(mix)
    return mix.to(device)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(batch, frames, mode):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed ncu launch list of the
    config-2 workload (tools/summarize_launches.py --json). ncu cannot run inside the bench; the file is valid only for
    the configuration that was profiled (batch 32 x 300 frames, mode 4), otherwise `traffic` is null."""
    if not (batch == 32 and frames == 300 and mode == 4):
        return {}, None
    for name in ("r02b_kernel_traffic.json", "r02_kernel_traffic.json", "r01_kernel_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            with open(path) as f:
                return json.load(f), "profiles/" + name
    return {}, None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], tflops_burst=p["bf16_tflops"], tflops_sustained=p["bf16_tflops_sustained"],
                    source="MEASURED_PEAKS.json")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------
# CPU port of the reference path (oracle/) - reported baseline and the --impl reference arm
# ------------------------------------------------------------------------------------------------
_CPU = {}


def cpu_port(mixing_mode=4):
    """the CPU oracle of the full-size model, built ONCE per process, on every host core. torchrun exports
    OMP_NUM_THREADS=1 to its workers, so the thread count is set explicitly (and reported)."""
    if mixing_mode not in _CPU:
        from oracle import mixer_oracle
        from mixermdm_b200 import schedule
        torch.set_num_threads(os.cpu_count() or 1)
        mix = build_model("cpu", mixing_mode)
        sd = {k: v for k, v in mix.state_dict().items()}
        sched = schedule.SpacedSchedule(schedule.space_timesteps(1000, "ddim50"),
                                        schedule.get_named_beta_schedule("cosine", 1000))
        _CPU[mixing_mode] = mixer_oracle.Oracle(sd, FULL_DIMS, synthetic_stats(), mixing_mode=mixing_mode, sched=sched)
    return _CPU[mixing_mode]


def cpu_port_sample(batch, ddim_steps, T, mixing_mode=4, warm=False):
    """wall seconds of `ddim_steps` consecutive DDIM steps (CFG-doubled like the reference) of `batch` samples"""
    orc = cpu_port(mixing_mode)
    g = torch.Generator().manual_seed(123)
    cond = torch.randn(batch, 6144, generator=g)
    x = torch.randn(batch, T, 524, generator=g)
    with torch.no_grad():
        if warm:
            orc.ddim_step(x, x, NSTEPS - 1, cond)
        t0 = time.perf_counter()
        x1, x2 = x, x
        for i in range(ddim_steps):
            o = orc.ddim_step(x1, x2, NSTEPS - 1 - i, cond)
            x1, x2 = o["sample"], o["sample2"]
        return time.perf_counter() - t0


def run_reference_arm(args, rank, world):
    """The reference's CPU path (oracle port; kind "port") on all host cores. One timed step = a BOUNDED SAMPLE of the
    workload: ref_batch samples x ref_ddim_steps DDIM steps. `ms_per_step` is that measured time; `value` scales it to
    the metric's unit (samples per 50-step loop per second) - the CPU cost is linear in samples and DDIM steps (every
    step runs the same three networks), the factor is stated."""
    if rank != 0:
        return
    cfg_id = args.config or (2 if world == 1 else 5)
    cpu_port(args.mode)
    cpu_port_sample(args.ref_batch, 1, args.frames, args.mode)           # warm-up of the allocator / thread pool
    per = []
    for i in range(args.warmup + args.steps):
        dt = cpu_port_sample(args.ref_batch, args.ref_ddim_steps, args.frames, args.mode)
        if i >= args.warmup:
            per.append(dt)
    s_sample = statistics.mean(per)
    s_per_sample_step = s_sample / (args.ref_batch * args.ref_ddim_steps)
    value = 1.0 / (s_per_sample_step * NSTEPS)
    threads = torch.get_num_threads()
    sample = (f"batch {args.ref_batch} x {args.frames} frames x {args.ref_ddim_steps} of {NSTEPS} DDIM steps per timed step "
              f"(CFG-doubled), fp32 torch CPU port of the reference path")
    B = args.batch if cfg_id == 2 else CONFIG5_GLOBAL_BATCH // world
    conf = workload_config(cfg_id, B, args.frames, args.mode, world)
    conf["cpu_sample"] = {"batch": args.ref_batch, "ddim_steps": args.ref_ddim_steps}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_sample * 1e3,
        "higher_is_better": True, "scaling": conf["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": conf,
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "extrapolation": {"measured": f"{args.ref_batch} samples x {args.ref_ddim_steps} DDIM steps in ms_per_step",
                          "samples_per_timed_step": args.ref_batch,
                          "ddim_step_factor": NSTEPS / args.ref_ddim_steps,
                          "value_formula": "ref_batch / (ms_per_step / 1e3 * 50 / ref_ddim_steps)"},
        "ms_per_denoise_step_per_sample": s_per_sample_step * 1e3,
    }
    emit(line)


def workload_config(cfg_id, B, T, mode, world):
    if cfg_id == 5:
        return {"workload": f"BASELINE configs[4]: MixerMDM batch {CONFIG5_GLOBAL_BATCH} x {T} frames batch-sharded over "
                            f"{world} GPU(s) ({B} per GPU), {NSTEPS} DDIM steps (ddim50), both denoisers (D=1024, L=8) + "
                            f"Mixer (mode {mode}), CFG 3.5, one final NCCL all_gather",
                "baseline_config": 5, "batch_per_gpu": B, "global_batch": CONFIG5_GLOBAL_BATCH, "frames": T,
                "ddim_steps": NSTEPS, "mixing_mode": mode, "scaling": "strong",
                "parallelism": f"dp{world} (contiguous batch shards, distributed.sample_sharded, final NCCL all_gather)",
                "l2": "per-step working set (GBs of activations) exceeds the 126 MB L2; no explicit flush"}
    return {"workload": f"BASELINE configs[1]: MixerMDM batch {B} x {T} frames per GPU, {NSTEPS} DDIM steps (ddim50), "
                        f"interaction + individual denoisers (D=1024, L=8) + spatio-temporal Mixer (mode {mode}), CFG 3.5",
            "baseline_config": 2, "batch_per_gpu": B, "global_batch": B * world, "frames": T,
            "ddim_steps": NSTEPS, "mixing_mode": mode, "scaling": "weak",
            "parallelism": f"dp{world} (contiguous batch shards, distributed.sample_sharded, final NCCL all_gather)",
            "l2": "per-step working set (~1.3 GB activations at B=32) exceeds the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------
# BASELINE configs[2]: the interaction denoiser alone (two persons x 262-dim, batch 64, 300 frames)
# ------------------------------------------------------------------------------------------------
def bench_config3(dev, batch=64, frames=300, iters=10, warmup=3):
    from mixermdm_b200 import modules
    torch.manual_seed(0)
    den = modules.in2INDenoiser(262, mode="interaction", dropout=0.1, **FULL_DIMS["den"]).eval()
    _reinit_zero_modules(den)
    den = den.to(dev)
    g = torch.Generator().manual_seed(1)
    B, T = batch, frames
    x = torch.randn(B, T, 524, generator=g).to(dev)
    cond = torch.randn(B, 2304, generator=g).to(dev)
    t = torch.full((B,), 500, dtype=torch.int64, device=dev)
    with torch.no_grad():
        for _ in range(warmup):
            out = den(x, t, cond=cond)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            out = den(x, t, cond=cond)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        hd = den._standalone.hd
        hd.profile(True)
        out = den(x, t, cond=cond)
        torch.cuda.synchronize()
        prof = hd.profile_read()
        hd.profile(False)
    peaks = measured_peaks()
    flop = B * GFLOP_INTERACTION_FWD * 1e9 * T / 300
    tot = sum(p["ms"] for p in prof)
    tens = [p for p in prof if p["flops"] > 0]
    t_ms, t_fl = sum(p["ms"] for p in tens), sum(p["flops"] for p in tens)
    ach = t_fl / (t_ms * 1e-3) / 1e12 if t_ms > 0 else 0.0
    return {
        "workload": f"BASELINE configs[2]: interaction denoiser alone, batch {B} x {T} frames, one forward (no CFG), "
                    f"M = {2 * B * T} tokens", "ms_per_forward": ms, "model_tflops": flop / (ms * 1e-3) / 1e12,
        "samples_per_s_one_step": B / (ms * 1e-3), "finite": bool(torch.isfinite(out).all()),
        "l2": "activations (~160 MB per tensor) exceed L2; back-to-back forwards, no explicit flush",
        "roofline": {"bound": "tensor", "kernel": "tcgen05 GEMM + attention launches of one forward", "achieved": ach,
                     "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["tflops_sustained"],
                     "traffic": None, "launches": sum(p["launches"] for p in tens),
                     "peak_source": peaks["source"] + " bf16_tflops_sustained, of measured"},
        "kernels": [{"name": p["name"], "launches": p["launches"], "ms": p["ms"], "share": p["ms"] / tot,
                     "tflops": p["flops"] / (p["ms"] * 1e-3) / 1e12 if p["flops"] > 0 and p["ms"] > 0 else None,
                     "frac_of_sustained_peak": (p["flops"] / (p["ms"] * 1e-3) / 1e12 / peaks["tflops_sustained"])
                     if p["flops"] > 0 and p["ms"] > 0 else None}
                    for p in sorted(prof, key=lambda q: -q["ms"])],
    }


# ------------------------------------------------------------------------------------------------
# BASELINE configs[3]: Mixer granularity sweep, the fused mix + CFG + centre + DDIM kernel alone, batch 128
# ------------------------------------------------------------------------------------------------
def bench_config4(dev, batch=128, frames=300, iters=20, warmup=3):
    from mixermdm_b200 import _lib, schedule
    from mixermdm_b200.engine import MixstepArgs
    lib = _lib.load()
    B, T = batch, frames
    sched = schedule.SpacedSchedule(schedule.space_timesteps(1000, "ddim50"), schedule.get_named_beta_schedule("cosine", 1000))
    g = torch.Generator(device=dev).manual_seed(0)

    def r(*s):
        return torch.randn(*s, device=dev, generator=g)
    # the engine's layout: [out1a row | out2a row], [x1 row | x2 row]
    preds, state = r(2 * B, T, 2, 524), r(B, T, 2, 524)
    out1a, out2a, x1, x2 = preds[:, :, 0], preds[:, :, 1], state[:, :, 0], state[:, :, 1]
    fin = torch.empty(B, T, 524, device=dev)
    coef = torch.from_numpy(sched.ddim_coefficients()).to(dev)
    step = torch.tensor([17], dtype=torch.int32, device=dev)
    m1, s1, m2, s2 = r(262) * 0.1, r(262).abs() + 0.5, r(262) * 0.1, r(262).abs() + 0.5
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peaks = measured_peaks()
    modes = []
    names = {1: "global", 2: "temporal", 3: "spatial", 4: "spatio-temporal"}
    for mode in (1, 2, 3, 4):
        nw = 1 if mode <= 2 else 23
        pooled = mode in (1, 3)
        rows = 2 * B * 2 * (1 if pooled else T)
        w = torch.sigmoid(r(rows, 24))
        a = MixstepArgs()
        a.B, a.T, a.mixing_mode, a.ncfg, a.mix_only, a.center, a.has_force = B, T, mode, 2, 0, 1, 0
        a.force_val, a.cfg_scale, a.nsteps = 0.0, 3.5, 50
        for name, tt in [("step_ptr", step), ("coef", coef), ("weights", w), ("out1a", out1a), ("out2a", out2a),
                         ("x1", x1), ("x2", x2), ("final_out", fin), ("mean1", m1), ("std1", s1), ("mean2", m2),
                         ("std2", s2)]:
            setattr(a, name, tt.data_ptr())
        times = []
        for it in range(iters + warmup):
            flush.zero_()                                  # L2 flush between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.mmdm_mix_cfg_ddim(C.byref(a), _lib.cur_stream()), "mix")
            e1.record()
            torch.cuda.synchronize()
            if it >= warmup:
                times.append(e0.elapsed_time(e1))
        ms = float(np.median(times))
        # as built (weights materialised by the head GEMM): T*17136 B per sample + the weights
        built = B * (T * (2 * 2 * 524 * 4 + 2 * 524 * 4 + 2 * 524 * 4) + 2 * 2 * nw * 4 * (1 if pooled else T))
        # SURVEY 8(d) canonical accounting (head fused, bf16 h rows read instead of fp32 weights): T*20864 B per sample
        canon = B * T * 20864
        modes.append({"mode": mode, "granularity": names[mode], "us": ms * 1e3,
                      "bytes_as_built": built, "gbs_as_built": built / ms / 1e6,
                      "frac_as_built": built / ms / 1e6 / peaks["hbm_gbs"],
                      "bytes_canonical": canon, "gbs_canonical": canon / ms / 1e6,
                      "frac_canonical": canon / ms / 1e6 / peaks["hbm_gbs"]})
    m4 = modes[-1]
    return {
        "workload": f"BASELINE configs[3]: Mixer granularity sweep, fused mix+CFG+centre+DDIM kernel alone, batch {B} x {T} "
                    f"frames, L2 flushed (256 MB write) between timed launches, median of {iters}",
        "roofline": {"bound": "hbm", "kernel": "mixstep_kernel<23> (mode 4)", "achieved": m4["gbs_as_built"],
                     "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": m4["frac_as_built"], "traffic": None,
                     "bytes_per_launch": m4["bytes_as_built"], "avg_launch_us": m4["us"],
                     "accounting": "bytes the kernel as built must move (weights materialised by the head GEMM); the "
                                   "SURVEY 8(d) canonical T*20864 B figure is in modes[].gbs_canonical",
                     "peak_source": peaks["source"] + " hbm_gbs, of measured"},
        "modes": modes,
    }


# ------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line: libraries that write to file descriptor 1 (NCCL's version banner does) are
    pointed at stderr for the rest of the process; emit() writes the line to the original descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=0, choices=[0, 2, 3, 4, 5],
                    help="BASELINE config (1-based as in this file's docstring); 0 = 2 on one GPU, 5 on several")
    ap.add_argument("--batch", type=int, default=32, help="samples per GPU (config 2)")
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--mode", type=int, default=4)
    ap.add_argument("--profile-steps", type=int, default=2, help="DDIM steps of the eager per-kernel timing pass")
    ap.add_argument("--cpu-batch", type=int, default=2, help="samples of the cpu_baseline sample")
    ap.add_argument("--cpu-ddim-steps", type=int, default=4, help="DDIM steps of the cpu_baseline sample")
    ap.add_argument("--ref-batch", type=int, default=4, help="samples per timed step of the --impl reference arm")
    ap.add_argument("--ref-ddim-steps", type=int, default=2, help="DDIM steps per timed step of the reference arm")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed loops of the e2e leg (0: min(steps, 5))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-subconfigs", action="store_true", help="skip the config3 / config4 sub-records")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the CUDA extension is the only product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    if args.config in (3, 4):
        if rank == 0:
            sub = bench_config3(dev) if args.config == 3 else bench_config4(dev)
            rl = sub["roofline"]
            line = {"metric": "interaction denoiser forward, ms" if args.config == 3 else "fused mix+DDIM kernel, GB/s (mode 4)",
                    "value": sub["ms_per_forward"] if args.config == 3 else rl["achieved"],
                    "unit": "ms" if args.config == 3 else "GB/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                    "ms_per_step": sub["ms_per_forward"] if args.config == 3 else rl["avg_launch_us"] / 1e3,
                    "higher_is_better": args.config == 4, "scaling": "weak", "vs_baseline": None,
                    "dtype": "bf16" if args.config == 3 else "f32", "data": "synthetic",
                    "config": {"workload": sub["workload"]}, "roofline": rl}
            line.update({k: v for k, v in sub.items() if k not in ("workload", "roofline")})
            emit(line)
        return

    dist = None
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION prints it there) off it
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    from mixermdm_b200 import distributed as mdist
    from mixermdm_b200.diffusion import MixerDiffusion, get_named_beta_schedule, space_timesteps
    from mixermdm_b200.modules import ClassifierFreeSampleModelX2

    cfg_id = args.config or (2 if world == 1 else 5)
    T = args.frames
    if cfg_id == 5:
        Bg = CONFIG5_GLOBAL_BATCH
        lo, hi = mdist.shard_bounds(Bg, rank, world)
        B = hi - lo
    else:
        B = args.batch
        Bg = B * world
        lo, hi = rank * B, (rank + 1) * B
    mix = build_model(dev, args.mode)
    cfgm = ClassifierFreeSampleModelX2(mix, 3.5)
    diff = MixerDiffusion(use_timesteps=space_timesteps(1000, "ddim50"), betas=get_named_beta_schedule("cosine", 1000))
    eng = mix.engine(2 * B, T, dev, sched=diff.schedule, cfg_scale=3.5, center=True)

    # inputs: the full batch of conditions (every rank holds it, as sample_sharded expects) and this rank's noise,
    # seeded per GLOBAL sample index so the result does not depend on the partition
    gcpu = torch.Generator(device="cpu").manual_seed(123)
    cond_h = torch.randn(Bg, 6144, generator=gcpu).pin_memory()
    noise_h = mdist.seeded_noise(range(lo, hi), T, base_seed=1000).pin_memory()
    cond_d, noise_d = cond_h.to(dev), noise_h.to(dev)

    def local_loop(c, n):
        return diff.ddim_sample_loop(cfgm, tuple(n.shape), noise=n, clip_denoised=False, progress=False,
                                     model_kwargs={"mask": None, "cond": c})

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def one_loop_resident():
        return mdist.sample_sharded(local_loop, cond_d, T, noise=noise_d, device=dev)

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        one_loop_resident()
    launches0 = eng.hd.launches()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    total_ms = timed(one_loop_resident, args.steps)
    clock_info = clocks.stop() if rank == 0 else None
    launches = eng.hd.launches() - launches0
    ms_per_loop = total_ms / args.steps
    value = Bg / (ms_per_loop / 1e3)

    # ---- end-to-end through the public API with host buffers ---------------------------------------
    res_h = torch.empty(Bg, T, 524).pin_memory() if rank == 0 else None

    def one_loop_e2e():
        c = cond_h.to(dev, non_blocking=True)
        n = noise_h.to(dev, non_blocking=True)
        out = mdist.sample_sharded(local_loop, c, T, noise=n, device=dev)
        if rank == 0:
            res_h.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_steps = args.e2e_steps or min(args.steps, 5)
    one_loop_e2e()
    e2e_ms = timed(one_loop_e2e, e2e_steps) / e2e_steps
    e2e_value = Bg / (e2e_ms / 1e3)
    h2d = cond_h.numel() * 4 + noise_h.numel() * 4
    d2h = Bg * T * 524 * 4

    # ---- per-kernel device times (eager pass, CUDA events around every launch on the launch stream) ---
    prof = None
    if rank == 0:
        eng.set_condition(cond_d[lo:hi])
        eng.set_state(noise_d, noise_d, NSTEPS - 1)
        eng.step()                                  # warm, unprofiled
        eng.hd.profile(True)
        for _ in range(args.profile_steps):
            eng.step()
        torch.cuda.synchronize()
        prof = eng.hd.profile_read()
        eng.hd.profile(False)
    barrier()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    traffic, traffic_file = ncu_traffic(B, T, args.mode)
    tot_ms = sum(p["ms"] for p in prof)
    gem = [p for p in prof if p["name"].startswith("gemm_")]
    gem_ms = sum(p["ms"] for p in gem)
    gem_fl = sum(p["flops"] for p in gem)
    gem_n = sum(p["launches"] for p in gem)
    achieved_tf = gem_fl / (gem_ms * 1e-3) / 1e12 if gem_ms > 0 else 0.0
    tsrc = (f"ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, committed launch list {traffic_file} "
            f"(not this run)") if traffic else None
    roofline_all = {"bound": "tensor", "kernel": "gemm_bf16_tn_kernel (all tcgen05 GEMMs of the step, HBM-bound residual GEMMs "
                                                 "and the M = 128 conditioning GEMMs included)",
                    "achieved": achieved_tf, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                    "frac": achieved_tf / peaks["tflops_sustained"],
                    "traffic": traffic.get("gemm_class", {}).get("dram_bytes_per_launch"), "traffic_source": tsrc,
                    "launches": gem_n, "avg_launch_us": 1e3 * gem_ms / max(gem_n, 1),
                    "flops_per_launch": gem_fl / max(gem_n, 1), "share_of_step": gem_ms / tot_ms if tot_ms else None}
    # the dominant kernel of the step: the QKV-class projections (self-attention in-proj, cross-attention q / kv)
    qkv = [p for p in prof if p["name"] == "gemm_qkv"]
    q_ms, q_fl, q_n = (qkv[0]["ms"], qkv[0]["flops"], qkv[0]["launches"]) if qkv else (0.0, 0.0, 0)
    q_tf = q_fl / (q_ms * 1e-3) / 1e12 if q_ms > 0 else 0.0
    q_kernel = "gemm_bf16_tn_2cta_kernel<8>" if "gemm_bf16_tn_2cta_kernel<8>" in traffic else "gemm_bf16_tn_2cta_kernel<0>"
    roofline = {"bound": "tensor", "kernel": f"{q_kernel} (gemm_qkv: the QKV-class projections, the largest share of the step)",
                "achieved": q_tf, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": q_tf / peaks["tflops_sustained"],
                "traffic": traffic.get(q_kernel, {}).get("dram_bytes_per_launch"), "traffic_source": tsrc,
                "peak_source": peaks["source"] + " bf16_tflops_sustained (kernel timed inside a long step), of measured",
                "launches": q_n, "avg_launch_us": 1e3 * q_ms / max(q_n, 1), "flops_per_launch": q_fl / max(q_n, 1),
                "share_of_step": q_ms / tot_ms if tot_ms else None}
    # the HBM-bound GEMMs: out-projection / FFN-2 + residual (+ the next AdaLN site's bf16 operands with the LN fold)
    res = [p for p in prof if p["name"] in ("gemm_proj", "gemm_ffn2")]
    r_ms, r_by, r_n = sum(p["ms"] for p in res), sum(p["bytes"] for p in res), sum(p["launches"] for p in res)
    r_kernel = "gemm_bf16_tn_2cta_kernel<10>" if "gemm_bf16_tn_2cta_kernel<10>" in traffic else "gemm_bf16_tn_2cta_kernel<2>"
    roofline_resid = None
    if r_ms > 0:
        r_gbs = r_by / (r_ms * 1e-3) / 1e9
        roofline_resid = {"bound": "hbm", "kernel": f"{r_kernel} (gemm_proj + gemm_ffn2: residual GEMMs)", "achieved": r_gbs,
                          "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": r_gbs / peaks["hbm_gbs"],
                          "traffic": traffic.get(r_kernel, {}).get("dram_bytes_per_launch"), "traffic_source": tsrc,
                          "bytes_per_launch": r_by / max(r_n, 1), "launches": r_n, "avg_launch_us": 1e3 * r_ms / max(r_n, 1),
                          "share_of_step": r_ms / tot_ms if tot_ms else None,
                          "accounting": "both GEMM operands once + fp32 residual row in + fp32 row out + bf16 operand(s) out"}
    mixp = [p for p in prof if p["name"] == "mix_cfg_ddim"]
    roofline_mix = None
    if mixp and mixp[0]["ms"] > 0:
        gbs = mixp[0]["bytes"] / (mixp[0]["ms"] * 1e-3) / 1e9
        roofline_mix = {"bound": "hbm", "kernel": "mixstep_kernel", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": gbs / peaks["hbm_gbs"],
                        "traffic": traffic.get("mixstep_kernel<23>", {}).get("dram_bytes_per_launch"),
                        "bytes_per_launch": mixp[0]["bytes"] / mixp[0]["launches"],
                        "avg_launch_us": 1e3 * mixp[0]["ms"] / mixp[0]["launches"]}
    kernels = sorted(({"name": p["name"], "launches_per_ddim_step": p["launches"] / args.profile_steps,
                       "ms_per_ddim_step": p["ms"] / args.profile_steps, "share": p["ms"] / tot_ms if tot_ms else None,
                       "tflops": (p["flops"] / (p["ms"] * 1e-3) / 1e12) if p["ms"] > 0 and p["flops"] > 0 else None,
                       "frac_of_sustained_peak": (p["flops"] / (p["ms"] * 1e-3) / 1e12 / peaks["tflops_sustained"])
                       if p["ms"] > 0 and p["flops"] > 0 else None,
                       "gbs": (p["bytes"] / (p["ms"] * 1e-3) / 1e9) if p["ms"] > 0 and p["bytes"] > 0 else None}
                      for p in prof), key=lambda d: -d["ms_per_ddim_step"])

    sub3 = sub4 = None
    if world == 1 and not args.no_subconfigs:
        del eng
        mix._engine = None
        torch.cuda.empty_cache()
        sub3 = bench_config3(dev)
        sub4 = bench_config4(dev)

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_port(args.mode)
        dt = cpu_port_sample(args.cpu_batch, args.cpu_ddim_steps, T, args.mode, warm=True)
        per = dt / (args.cpu_batch * args.cpu_ddim_steps)
        cpu_baseline = {"value": 1.0 / (per * NSTEPS), "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"batch {args.cpu_batch} x {T} frames x {args.cpu_ddim_steps} of {NSTEPS} DDIM steps "
                                  f"(CFG-doubled) in {dt:.1f} s, fp32 torch CPU port of the reference path; value = "
                                  f"samples per second at 50 steps per sample (cost is linear in both)",
                        "ms_per_denoise_step_per_sample": per * 1e3}

    conf = workload_config(cfg_id, B, T, args.mode, world)
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_loop, "higher_is_better": True,
        "scaling": conf["scaling"], "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": conf,
        "ms_per_denoise_step": ms_per_loop / NSTEPS,
        "model_tflops": Bg * GFLOP_PER_SAMPLE_STEP * NSTEPS / 1e3 / (ms_per_loop / 1e3),
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms, "timed_loops": e2e_steps,
                "api": "distributed.sample_sharded(MixerDiffusion.ddim_sample_loop(ClassifierFreeSampleModelX2(Mixer), ...))"},
        "gpu_launches": launches,
        "clocks": clock_info,
        "roofline": roofline,
        "roofline_all_gemms": roofline_all,
        "roofline_resid": roofline_resid,
        "roofline_mix": roofline_mix,
        "kernels": kernels,
        "config3": sub3,
        "config4": sub4,
        "cpu_baseline": cpu_baseline,
    }
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
