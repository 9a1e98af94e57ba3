"""Dataset-scale entry: one process, all visible GPUs, several robots.

The reference converts a dataset with `mp.Pool.starmap(process_file, ...)` - one file per worker process, one
robot per run (scripts/smplx_to_robot_dataset.py:241-242, scripts/bvh_to_robot_dataset.py:166-179).  Here the
caller hands over whole buckets of clips, one bucket per (source format, robot) pair, as host arrays; the clips
of every bucket are dealt to the GPUs by a hardness proxy (`sharding.lpt_shard`: a step ends when the slowest
GPU ends, and what makes a GPU slow is its share of the hard clips), one host thread per GPU runs its share of
ALL buckets as one mixed-robot call (`retarget_mixed`, pinned host memory used in place), and the results come
back in the callers' clip order.  No collective: clips are independent (BASELINE.json configs[4]).
"""
from __future__ import annotations

import threading
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .motion_retarget import GeneralMotionRetargeting, retarget_mixed
from .sharding import all_shards, hardness_proxy, lpt_shard

Job = Tuple[str, str, np.ndarray, np.ndarray, Optional[np.ndarray]]       # (src_human, tgt_robot, pos, quat, heights)


def plan_shards(jobs: Sequence[Job], n_devices: int, shard: str = "lpt") -> List[List[np.ndarray]]:
    """[job][device] -> ascending clip indices.  Pure host logic (tested without a GPU)."""
    if shard not in ("lpt", "contiguous"):
        raise ValueError("shard must be 'lpt' or 'contiguous'")
    plans = []
    for src, robot, pos, quat, heights in jobs:
        C = int(np.shape(pos)[0])
        if shard == "contiguous" or n_devices == 1 or C == 0:
            plans.append([np.arange(b, e) for b, e in all_shards(C, n_devices)])
            continue
        g = GeneralMotionRetargeting.__new__(GeneralMotionRetargeting)      # tables only: no device, no library
        from .ik_config import compile_task_table
        from .params import load_pack
        rob, cfg, _ = load_pack(src, robot)
        table = compile_task_table(rob, cfg)
        hard = hardness_proxy(np.asarray(quat)[:, 0], table.root_idx, table.rot_off[table.root_idx], rob.qpos0[3:7])
        plans.append(lpt_shard(hard, n_devices))
        del g
    return plans


def retarget_clips_multi_gpu(jobs: Sequence[Job], devices: Optional[Sequence[int]] = None, precision: str = "f64",
                             shard: str = "lpt", return_info: bool = False):
    """Retarget several buckets of clips on all GPUs of the node from ONE process.

    jobs: [(src_human, tgt_robot, pos [C,T,nh,3], quat [C,T,nh,4], heights [C] or None), ...] with host arrays
    (numpy or CPU torch tensors; at most 8 buckets).  devices: CUDA device indices (default: all visible).
    Returns one float32 array [C,T,nq] per job, in the job's clip order; with return_info also the iteration
    counts [C,T,2] per job.  Clip c of a job is solved exactly as `GeneralMotionRetargeting(src, robot,
    heights[c])` + `retarget(frame)` per frame, whichever GPU it lands on."""
    import torch
    if devices is None:
        devices = list(range(torch.cuda.device_count()))
    devices = list(devices)
    if not devices:
        raise RuntimeError("no CUDA device visible")
    plans = plan_shards(jobs, len(devices), shard)
    retargeters = [[GeneralMotionRetargeting(src, robot, device=d) for (src, robot, *_rest) in jobs] for d in devices]
    results = [np.zeros((int(np.shape(pos)[0]), int(np.shape(pos)[1]), retargeters[0][k]._robot.nq), np.float32)
               for k, (_, _, pos, _, _) in enumerate(jobs)]
    infos = [np.zeros((int(np.shape(pos)[0]), int(np.shape(pos)[1]), 2), np.int32) for (_, _, pos, _, _) in jobs] if return_info else None
    errors: List[BaseException] = []

    def work(di: int, dev: int) -> None:
        try:
            torch.cuda.set_device(dev)
            buckets, where = [], []
            for k, (src, robot, pos, quat, heights) in enumerate(jobs):
                ids = plans[k][di]
                if ids.size == 0:
                    continue
                p = torch.from_numpy(np.ascontiguousarray(np.asarray(pos)[ids], np.float32)).pin_memory()
                q = torch.from_numpy(np.ascontiguousarray(np.asarray(quat)[ids], np.float32)).pin_memory()
                h = None if heights is None else torch.from_numpy(np.ascontiguousarray(np.asarray(heights)[ids], np.float32))
                buckets.append((retargeters[di][k], p, q, h)); where.append((k, ids))
            if not buckets:
                return
            out = retarget_mixed(buckets, precision=precision, return_info=return_info, device=dev)
            outs, its = out if return_info else (out, None)
            for j, (k, ids) in enumerate(where):
                results[k][ids] = outs[j].numpy()
                if return_info:
                    infos[k][ids] = its[j].numpy()
        except BaseException as e:            # surfaced in the calling thread
            errors.append(e)

    threads = [threading.Thread(target=work, args=(di, dev), name=f"gmr-gpu{dev}") for di, dev in enumerate(devices)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return (results, infos) if return_info else results
