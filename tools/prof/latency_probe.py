#!/usr/bin/env python
"""Separates the two regimes of the solve kernel on one GPU:
  * lone-warp latency  : one clip per SM (C = #SMs)                -> us per solve of an isolated warp
  * balanced throughput: every warp slot runs a copy of the SAME clip -> us per convoy round, no tail
  * the benchmark mix  : 4096 different clips                      -> what the imbalance costs
Prints one JSON line per case."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from general_motion_retargeting_b200 import GeneralMotionRetargeting, params
from general_motion_retargeting_b200.ik_config import compile_task_table
from general_motion_retargeting_b200.synthetic import make_clips

T = int(os.environ.get("PROBE_T", "100"))
robot, cfg, _ = params.load_pack("smplx", "unitree_g1")
table = compile_task_table(robot, cfg)
gmr = GeneralMotionRetargeting("smplx", "unitree_g1", device=0)
base = make_clips(robot, table, range(4096), T=T)

def run(tag, pos, quat, h, prec):
    dp, dq, dh = (torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (pos, quat, h))
    q, it, _ = gmr.retarget_batch(dp, dq, dh, return_info=True, precision=prec)
    torch.cuda.synchronize()
    s = it.sum(-1).sum(-1).cpu().numpy()
    for _ in range(2):
        gmr.retarget_batch(dp, dq, dh, precision=prec)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    n = 3
    for _ in range(n):
        gmr.retarget_batch(dp, dq, dh, precision=prec)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(json.dumps({"case": tag, "precision": prec, "clips": int(pos.shape[0]), "frames": T, "ms": ms,
                      "frames_per_s": pos.shape[0] * T / ms * 1e3, "solves_mean": float(s.mean()), "solves_max": int(s.max()),
                      "us_per_solve_of_longest_clip": ms * 1e3 / s.max(), "us_per_mean_solve": ms * 1e3 / s.mean()}), flush=True)

for prec in ("f32", "f64"):
    slots = 148 * (28 if prec == "f32" else 14)
    run("lone_warp_one_clip_per_sm", base.pos[:148], base.quat[:148], base.heights[:148], prec)
    rep = lambda x, n: np.repeat(x[:1], n, axis=0)
    run("lone_warp_same_clip", rep(base.pos, 148), rep(base.quat, 148), rep(base.heights, 148), prec)
    for w in (2, 4, 8, 16):
        run("same_clip_%d_warps_per_sm" % w, rep(base.pos, 148 * w), rep(base.quat, 148 * w), rep(base.heights, 148 * w), prec)
    run("balanced_all_slots_same_clip", rep(base.pos, slots), rep(base.quat, slots), rep(base.heights, slots), prec)
    run("balanced_2x_slots_same_clip", rep(base.pos, 2 * slots), rep(base.quat, 2 * slots), rep(base.heights, 2 * slots), prec)
    run("benchmark_mix_4096", base.pos, base.quat, base.heights, prec)
