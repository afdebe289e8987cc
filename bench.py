#!/usr/bin/env python
"""Benchmark of the MixerMDM denoising loop (BASELINE.json metric: motion samples/sec, 300 frames, 50 DDIM steps).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--frames T]

One "step" = one complete 50-step DDIM sampling loop over a batch of B samples per GPU (BASELINE config[1]:
B=32 x 300 frames, spatio-temporal Mixer, bf16 tensor-core GEMMs, fp32 state). Prints ONE JSON line (rank 0).
`value`   : samples/s with cond + noise already resident in HBM, timed with CUDA events, max over ranks.
`e2e`     : the same loop through the public API (MixerDiffusion.ddim_sample_loop) with pinned HOST inputs:
            H2D of cond+noise and D2H of the result inside the timed region.
`roofline`: dominant kernel class (tcgen05 GEMMs) — algorithmic FLOPs / CUDA-event time per launch, measured
            in a separate eager pass with events around every launch of the same step (per-launch events cannot
            be recorded inside a CUDA-graph replay); `roofline_mix` is the fused HBM-bound mix/sampler kernel.
`cpu_baseline` / `--impl reference`: the CPU port of the reference path (oracle/, torch fp32 on the host cores;
            the Python reference itself cannot travel to the GPU box) on a bounded sample of the same workload.
"""
import argparse
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
# algorithmic work per sample per DDIM step (CFG included)
GFLOP_PER_SAMPLE_STEP = 478.5


def synthetic_stats():
    rng = np.random.default_rng(0)
    pairs = []
    for _ in range(2):
        pairs.append(((0.1 * rng.normal(size=262)).astype(np.float32), rng.uniform(0.5, 1.5, size=262).astype(np.float32)))
    glob, hml = pairs
    return hml, glob


def build_model(device, mixing_mode=4, dims=FULL_DIMS):
    """random-init weights of the named architecture (seed 0); zero-initialised linears re-drawn N(0, 0.02^2)
    (seed 1) so the networks do not output exact zeros (SURVEY.md §8a note 1)."""
    from mixermdm_b200 import modules
    torch.manual_seed(0)
    kw = dict(dropout=0.1, **dims["den"])
    d1 = modules.in2INDenoiser(262, mode="individual", **kw)
    d2 = modules.in2INDenoiser(262, mode="interaction", **kw)
    mix = modules.Mixer(d1, d2, nfeats=262, text_dim=768, mixing_mode=mixing_mode, store_influence=False,
                        mode="eval_intermediate", align=True, normalizers=synthetic_stats(), **dims["mix"]).eval()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in mix.parameters():
            if p.abs().sum() == 0:
                p.normal_(0, 0.02, generator=g)
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


def ncu_traffic(args):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed ncu launch list of this
    same workload (profiles/r01_launches_v5.csv -> profiles/r01_kernel_traffic.json, tools/summarize_launches.py).
    Only valid for the configuration that was profiled (batch 32 x 300 frames, mode 4)."""
    path = os.path.join(ROOT, "profiles", "r01_kernel_traffic.json")
    if not (os.path.exists(path) and args.batch == 32 and args.frames == 300 and args.mode == 4):
        return {}
    with open(path) as f:
        return json.load(f)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], tflops_burst=p["bf16_tflops"], tflops_sustained=p["bf16_tflops_sustained"],
                    source="MEASURED_PEAKS.json")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------
# CPU port of the reference path (oracle/) — reported baseline and the --impl reference arm
# ------------------------------------------------------------------------------------------------
def cpu_port_time(ddim_steps, T, threads=None, mixing_mode=4):
    """seconds per DDIM step of the CPU port at B=1 (CFG-doubled like the reference), fp32, all host cores"""
    from oracle import mixer_oracle
    from mixermdm_b200 import schedule
    if threads:
        torch.set_num_threads(threads)
    mix = build_model("cpu", mixing_mode)
    sd = {k: v for k, v in mix.state_dict().items()}
    sched = schedule.SpacedSchedule(schedule.space_timesteps(1000, "ddim50"), schedule.get_named_beta_schedule("cosine", 1000))
    orc = mixer_oracle.Oracle(sd, FULL_DIMS, synthetic_stats(), mixing_mode=mixing_mode, sched=sched)
    g = torch.Generator().manual_seed(123)
    cond = torch.randn(1, 6144, generator=g)
    x = torch.randn(1, T, 524, generator=g)
    with torch.no_grad():
        orc.ddim_step(x, x, NSTEPS - 1, cond)          # warm-up
        t0 = time.perf_counter()
        x1, x2 = x, x
        for i in range(ddim_steps):
            o = orc.ddim_step(x1, x2, NSTEPS - 1 - i, cond)
            x1, x2 = o["sample"], o["sample2"]
        dt = (time.perf_counter() - t0) / ddim_steps
    return dt, torch.get_num_threads()


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    per = []
    threads = None
    for i in range(args.warmup + args.steps):
        dt, threads = cpu_port_time(args.ref_ddim_steps, args.frames)
        if i >= args.warmup:
            per.append(dt)
    s_per_step = statistics.mean(per)
    loop_s = s_per_step * NSTEPS
    value = 1.0 / loop_s
    sample = (f"B=1 x {args.frames} frames, {args.ref_ddim_steps} of {NSTEPS} DDIM steps per timed step "
              f"(CFG-doubled), extrapolated to the 50-step loop")
    line = {
        "impl": "reference", "metric": "motion samples/sec (300f, 50 DDIM steps)", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": loop_s * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ms_per_denoise_step": s_per_step * 1e3,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": f"MixerMDM batch {args.batch} x {args.frames} frames per GPU, {NSTEPS} DDIM steps (ddim50), "
                        f"interaction + individual denoisers (D=1024, L=8) + spatio-temporal Mixer (mode {args.mode}), CFG 3.5",
            "batch_per_gpu": args.batch, "global_batch": args.batch * world, "frames": args.frames,
            "ddim_steps": NSTEPS, "mixing_mode": args.mode, "parallelism": f"dp{world} (batch sharded, final NCCL all_gather)",
            "l2": "per-step working set (~1.3 GB activations at B=32) exceeds the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="samples per GPU")
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--mode", type=int, default=4)
    ap.add_argument("--profile-steps", type=int, default=2, help="DDIM steps of the eager per-kernel timing pass")
    ap.add_argument("--cpu-ddim-steps", type=int, default=6, help="DDIM steps of the CPU baseline sample")
    ap.add_argument("--ref-ddim-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    from mixermdm_b200.diffusion import MixerDiffusion, get_named_beta_schedule, space_timesteps
    from mixermdm_b200.modules import ClassifierFreeSampleModelX2

    B, T = args.batch, args.frames
    mix = build_model(dev, args.mode)
    cfgm = ClassifierFreeSampleModelX2(mix, 3.5)
    diff = MixerDiffusion(use_timesteps=space_timesteps(1000, "ddim50"), betas=get_named_beta_schedule("cosine", 1000))
    eng = mix.engine(2 * B, T, dev, sched=diff.schedule, cfg_scale=3.5, center=True)

    # per-sample seeded inputs: global sample index = rank*B + b, independent of the partition
    g = torch.Generator(device="cpu")
    conds, noises = [], []
    for b in range(B):
        g.manual_seed(1000 + rank * B + b)
        conds.append(torch.randn(6144, generator=g))
        noises.append(torch.randn(T, 524, generator=g))
    cond_h = torch.stack(conds).pin_memory()
    noise_h = torch.stack(noises).pin_memory()
    cond_d, noise_d = cond_h.to(dev), noise_h.to(dev)
    out_d = torch.empty(B, T, 524, device=dev)
    gathered = [torch.empty_like(out_d) for _ in range(world)] if world > 1 else None

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def one_loop_resident():
        eng.sample(cond_d, noise_d, out=out_d, use_graph=True)
        if dist is not None:
            dist.all_gather(gathered, out_d)

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
    value = world * B / (ms_per_loop / 1e3)

    # ---- end-to-end through the public API with host buffers ---------------------------------------
    res_h = torch.empty(B, T, 524).pin_memory()

    def one_loop_e2e():
        c = cond_h.to(dev, non_blocking=True)
        n = noise_h.to(dev, non_blocking=True)
        out = diff.ddim_sample_loop(cfgm, (B, T, 524), noise=n, clip_denoised=False, progress=False,
                                    model_kwargs={"mask": None, "cond": c})
        if dist is not None:
            dist.all_gather(gathered, out)
        res_h.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    one_loop_e2e()
    e2e_ms = timed(one_loop_e2e, args.steps) / args.steps
    e2e_value = world * B / (e2e_ms / 1e3)
    h2d = cond_h.numel() * 4 + noise_h.numel() * 4
    d2h = res_h.numel() * 4

    # ---- per-kernel device times (eager pass, CUDA events around every launch on the launch stream) ---
    prof = None
    if rank == 0:
        eng.set_condition(cond_d)
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
    traffic = ncu_traffic(args)
    tot_ms = sum(p["ms"] for p in prof)
    gem = [p for p in prof if p["name"].startswith("gemm_")]
    gem_ms = sum(p["ms"] for p in gem)
    gem_fl = sum(p["flops"] for p in gem)
    gem_n = sum(p["launches"] for p in gem)
    achieved_tf = gem_fl / (gem_ms * 1e-3) / 1e12 if gem_ms > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "gemm_bf16_tn_kernel (all tcgen05 GEMMs of the step)",
                "achieved": achieved_tf, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                "frac": achieved_tf / peaks["tflops_sustained"],
                "traffic": traffic.get("gemm_class", {}).get("dram_bytes_per_launch"),
                "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the GEMM launches of one DDIM "
                                  "step (profiles/r01_launches_v5.csv)" if traffic else None,
                "peak_source": peaks["source"] + " bf16_tflops_sustained (kernel timed inside a long step), of measured",
                "launches": gem_n, "avg_launch_us": 1e3 * gem_ms / max(gem_n, 1),
                "flops_per_launch": gem_fl / max(gem_n, 1), "share_of_step": gem_ms / tot_ms if tot_ms else None}
    mixp = [p for p in prof if p["name"] == "mix_cfg_ddim"]
    roofline_mix = None
    if mixp and mixp[0]["ms"] > 0:
        gbs = mixp[0]["bytes"] / (mixp[0]["ms"] * 1e-3) / 1e9
        roofline_mix = {"bound": "hbm", "kernel": "mixstep_kernel", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": gbs / peaks["hbm_gbs"],
                        "traffic": traffic.get("mixstep_kernel<23>", {}).get("dram_bytes_per_launch"), "bytes_per_launch": mixp[0]["bytes"] / mixp[0]["launches"],
                        "avg_launch_us": 1e3 * mixp[0]["ms"] / mixp[0]["launches"]}
    kernels = sorted(({"name": p["name"], "launches_per_ddim_step": p["launches"] / args.profile_steps,
                       "ms_per_ddim_step": p["ms"] / args.profile_steps, "share": p["ms"] / tot_ms if tot_ms else None,
                       "tflops": (p["flops"] / (p["ms"] * 1e-3) / 1e12) if p["ms"] > 0 and p["flops"] > 0 else None}
                      for p in prof), key=lambda d: -d["ms_per_ddim_step"])

    cpu_baseline = None
    if not args.no_cpu_baseline:
        dt, threads = cpu_port_time(args.cpu_ddim_steps, T, mixing_mode=args.mode)
        cpu_baseline = {"value": 1.0 / (dt * NSTEPS), "unit": "samples/s", "cores": threads, "kind": "port",
                        "sample": f"B=1 x {T} frames, {args.cpu_ddim_steps} of {NSTEPS} DDIM steps (CFG-doubled), fp32 torch "
                                  f"CPU port of the reference path, extrapolated to the 50-step loop",
                        "ms_per_denoise_step": dt * 1e3}

    line = {
        "metric": "motion samples/sec (300f, 50 DDIM steps)", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_loop, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, world),
        "ms_per_denoise_step": ms_per_loop / NSTEPS,
        "model_tflops": world * B * GFLOP_PER_SAMPLE_STEP * NSTEPS / 1e3 / (ms_per_loop / 1e3),
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms, "api": "MixerDiffusion.ddim_sample_loop(ClassifierFreeSampleModelX2(Mixer), ...)"},
        "gpu_launches": launches,
        "clocks": clock_info,
        "roofline": roofline,
        "roofline_mix": roofline_mix,
        "kernels": kernels,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
