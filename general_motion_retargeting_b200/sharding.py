"""Clip sharding for multi-GPU runs: clips are independent (the reference runs one clip per
pool worker, scripts/smplx_to_robot_dataset.py:241-242), so rank r of W owns a contiguous
range of clips and there is no collective on the solve path.  Only the timing / result
hand-over of a caller needs communication (max over ranks, optional gather)."""
from __future__ import annotations

from typing import List, Tuple


def clip_shard(num_clips: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of rank `rank`: sizes differ by at most one, earlier ranks take the extras."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(num_clips, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def all_shards(num_clips: int, world: int) -> List[Tuple[int, int]]:
    return [clip_shard(num_clips, r, world) for r in range(world)]


def bucket_by_robot(robot_of_clip) -> dict:
    """Mixed-robot batches (BASELINE.json configs[4]: clip i -> robot i mod 5): a launch is robot-uniform (one
    constant block per CTA), so the host groups the clip indices by robot; each bucket is retargeted by its own
    `GeneralMotionRetargeting(src, robot)` and the results are scattered back by index.  Returns
    {robot: ascending list of clip indices}."""
    out: dict = {}
    for i, r in enumerate(robot_of_clip):
        out.setdefault(r, []).append(i)
    return out


def max_over_ranks(value: float) -> float:
    """Max of a host scalar over the default process group (no-op without one)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_clips(local, num_clips: int):
    """Concatenate per-rank results [c_r, ...] in rank order on every rank (off the timed path)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [e - b for b, e in all_shards(num_clips, world)]
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat([o[:n] for o, n in zip(out, sizes)], dim=0)
