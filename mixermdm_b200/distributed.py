"""Batch sharding for multi-GPU sampling. Samples are independent, so a batch of B prompts is split
into contiguous slices, one per rank (one process per GPU), the 50 DDIM steps run without any communication and
one NCCL all_gather over NVLink assembles the (B, T, 524) result. Noise is seeded per GLOBAL sample index, so the
union of the shards equals the single-GPU result bit for bit, whatever the partition."""
import torch


def shard_bounds(total, rank, world):
    """contiguous [lo, hi) slice of `total` samples owned by `rank`; earlier ranks take the remainder"""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def seeded_noise(global_indices, T, base_seed=0, device="cpu", feats=524):
    """(len(idx), T, 524) standard normal noise; sample i always draws from Generator(base_seed + i)"""
    out = []
    for i in global_indices:
        g = torch.Generator(device="cpu").manual_seed(int(base_seed) + int(i))
        out.append(torch.randn(T, feats, generator=g))
    if not out:
        return torch.empty(0, T, feats, device=device)
    return torch.stack(out).to(device)


def gather_batch(local, total, group=None):
    """all_gather of ragged contiguous shards back into a (total, ...) tensor, identical on every rank"""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = [shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world)]
    cap = max(sizes)
    buf = local.new_zeros((cap,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)


def sample_sharded(sample_fn, cond, T, base_seed=0, group=None, device=None, noise=None):
    """cond: (B, 6144) full batch of conditions (same on every rank). sample_fn(cond_slice, noise_slice) -> (b, T, 524)
    runs the local DDIM loop (e.g. `lambda c, n: diffusion.ddim_sample_loop(cfg_model, n.shape, noise=n, ...)`).
    noise: optional (hi - lo, T, 524) initial noise of THIS rank's shard (e.g. `seeded_noise(range(lo, hi), T, seed)`
    kept resident on the device); by default it is drawn here from the per-sample generators.
    Returns the gathered (B, T, 524) tensor."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    B = cond.shape[0]
    lo, hi = shard_bounds(B, rank, world)
    device = device if device is not None else cond.device
    if noise is None:
        noise = seeded_noise(range(lo, hi), T, base_seed, device)
    elif noise.shape[0] != hi - lo or noise.shape[1] != T:
        raise ValueError(f"noise: expected this rank's shard ({hi - lo}, {T}, 524), got {tuple(noise.shape)}")
    if hi > lo:
        local = sample_fn(cond[lo:hi].to(device).contiguous(), noise)
    else:
        local = torch.empty(0, T, 524, device=device)
    if world == 1:
        return local
    return gather_batch(local, B, group)
