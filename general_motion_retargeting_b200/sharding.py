"""Clip sharding for multi-GPU runs: clips are independent (the reference runs one clip per
pool worker, scripts/smplx_to_robot_dataset.py:241-242), so every rank owns a set of clips and
there is no collective on the solve path.  Only the timing / result hand-over of a caller needs
communication (max over ranks, optional gather).

Two assignments: contiguous ranges (`clip_shard`), and a hardness-aware deal (`lpt_shard`): a
step ends when the slowest rank ends, and what makes a rank slow is how many HARD clips it drew
(clips that start far from the robot's initial orientation settle on joint limits and need 2-3x
the IK steps, DESIGN.md "Clip scheduling"), so the clips are dealt to the ranks in order of a
cheap hardness proxy - every rank gets the same number of clips and the same share of the hard
ones."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def clip_shard(num_clips: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of rank `rank`: sizes differ by at most one, earlier ranks take the extras."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(num_clips, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def all_shards(num_clips: int, world: int) -> List[Tuple[int, int]]:
    return [clip_shard(num_clips, r, world) for r in range(world)]


def hardness_proxy(quat0: np.ndarray, human_root: int, root_rot_offset: Sequence[float], qpos0_quat: Sequence[float]) -> np.ndarray:
    """Rotation angle (rad) between each clip's first root target and the robot's initial root orientation - the
    criterion the library's own clip ordering uses (gmr_order_kernel, csrc/gmr_kernels.cu).  quat0 [C,nh,4] wxyz: frame 0."""
    q = np.asarray(quat0, np.float64)[:, human_root]
    o = np.asarray(root_rot_offset, np.float64)
    a0, a1, a2, a3 = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    t = np.stack([a0 * o[0] - a1 * o[1] - a2 * o[2] - a3 * o[3], a0 * o[1] + a1 * o[0] + a2 * o[3] - a3 * o[2],
                  a0 * o[2] - a1 * o[3] + a2 * o[0] + a3 * o[1], a0 * o[3] + a1 * o[2] - a2 * o[1] + a3 * o[0]], -1)
    r = np.asarray(qpos0_quat, np.float64)
    dot = np.abs(t @ r) / np.maximum(np.linalg.norm(t, axis=-1) * np.linalg.norm(r), 1e-30)
    return 2.0 * np.arccos(np.clip(dot, 0.0, 1.0))


def lpt_shard(hardness: Sequence[float], world: int) -> List[np.ndarray]:
    """Deal clips to `world` ranks in decreasing order of `hardness`, boustrophedon (0..W-1, W-1..0, ...): every rank gets
    floor/ceil(C / W) clips and an equal share of every hardness band.  Returns one ascending index array per rank;
    deterministic (ties broken by clip index), so every rank can compute the same assignment without communication."""
    h = np.asarray(hardness, np.float64)
    if world < 1:
        raise ValueError("bad world size")
    order = np.lexsort((np.arange(h.size), -h))
    pos = np.arange(h.size)
    rnd, k = pos // world, pos % world
    rank_of = np.where(rnd % 2 == 0, k, world - 1 - k)
    return [np.sort(order[rank_of == r]) for r in range(world)]


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
