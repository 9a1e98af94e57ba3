#!/usr/bin/env python
"""Smallest instances of every launch shape of the library, for compute-sanitizer
(memcheck / racecheck / synccheck / initcheck); tools/prof/sanitize.sh runs them.

    sanitize_cases.py twophase|plain|mixed|stream|dataset|host [f32|f64]

twophase needs GMR_PARTITION=1 in the environment so that ~300 clips already take the
frame-0 / classify / partitioned path (launch() in csrc/gmr_kernels.cu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from general_motion_retargeting_b200 import GeneralMotionRetargeting, params, retarget_mixed
from general_motion_retargeting_b200.ik_config import compile_task_table
from general_motion_retargeting_b200.synthetic import make_clips

case = sys.argv[1]
prec = sys.argv[2] if len(sys.argv) > 2 else "f64"


def problem(src, robot_name, ids, T, stress=False):
    robot, cfg, _ = params.load_pack(src, robot_name)
    table = compile_task_table(robot, cfg)
    clips = make_clips(robot, table, ids, T=T, src_human=src, stress=stress)
    return GeneralMotionRetargeting(src, robot_name, device=0), clips


def dev(clips):
    return tuple(torch.from_numpy(x).cuda() for x in (clips.pos, clips.quat, clips.heights))


if case == "twophase":
    g, clips = problem("smplx", "unitree_g1", range(320), 16)
    q, it, err = g.retarget_batch(*dev(clips), precision=prec, return_info=True)
    torch.cuda.synchronize()
    print("twophase", tuple(q.shape), "solves/frame", float(it.sum()) / (320 * 16))
elif case == "plain":
    g, clips = problem("smplx", "unitree_g1", range(40), 6, stress=True)     # active bounds from frame 0 on
    q, it, err = g.retarget_batch(*dev(clips), precision=prec, return_info=True)
    torch.cuda.synchronize()
    print("plain", tuple(q.shape), "solves/frame", float(it.sum()) / (40 * 6))
elif case == "mixed":
    buckets = []
    for k, name in enumerate(["unitree_g1", "booster_t1", "stanford_toddy"]):
        g, clips = problem("smplx", name, range(k, 24 + k), 6)
        buckets.append((g,) + dev(clips))
    outs = retarget_mixed(buckets, precision=prec)
    torch.cuda.synchronize()
    print("mixed", [tuple(o.shape) for o in outs])
elif case == "stream":
    g, clips = problem("smplx", "unitree_g1", range(1), 6)
    names = g.human_body_names
    for t in range(6):
        frame = {n: (clips.pos[0, t, i].astype(np.float64), clips.quat[0, t, i].astype(np.float64)) for i, n in enumerate(names)}
        q = g.retarget(frame)
    print("stream", q.shape, g.error1(), g.error2())
elif case == "dataset":
    g, clips = problem("smplx", "unitree_g1", range(12), 8)
    lengths = np.array([8, 7, 6, 5, 8, 8, 3, 8, 8, 1, 8, 8], np.int32)
    m = g.retarget_dataset(*dev(clips), lengths=lengths, precision=prec)
    print("dataset", len(m), m[1]["root_pos"].shape)
elif case == "host":
    g, clips = problem("smplx", "unitree_g1", range(64), 6)
    q = g.retarget_batch(clips.pos, clips.quat, clips.heights, precision=prec)
    print("host", q.shape)
else:
    raise SystemExit("unknown case " + case)
