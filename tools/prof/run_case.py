#!/usr/bin/env python
"""One retarget_batch launch of a chosen shape, for ncu captures.
   run_case.py <precision> <clips> <frames> [same]   ("same": every clip is a copy of clip 0)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from general_motion_retargeting_b200 import GeneralMotionRetargeting, params
from general_motion_retargeting_b200.ik_config import compile_task_table
from general_motion_retargeting_b200.synthetic import make_clips
prec, C, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
same = len(sys.argv) > 4 and sys.argv[4] == "same"
robot, cfg, _ = params.load_pack("smplx", "unitree_g1")
table = compile_task_table(robot, cfg)
gmr = GeneralMotionRetargeting("smplx", "unitree_g1", device=0)
first = int(os.environ.get("CASE_CLIP", "0"))
clips = make_clips(robot, table, range(first, first + (1 if same else C)), T=T, device="cuda")   # same draws as on the CPU, seconds instead of a minute
rep = (lambda x: np.repeat(x[:1], C, axis=0)) if same else (lambda x: x)
dp, dq, dh = (torch.from_numpy(np.ascontiguousarray(rep(x))).cuda() for x in (clips.pos, clips.quat, clips.heights))
for _ in range(3):
    q = gmr.retarget_batch(dp, dq, dh, precision=prec)
torch.cuda.synchronize()
print("ok", q.shape)
