#!/usr/bin/env python
"""Times one batch shape once per precision; scheduler knobs come from the environment (GMR_*).
   PROBE_C clips (4096), PROBE_T frames (100), PROBE_ROBOT (unitree_g1), PROBE_SRC (smplx), PROBE_STRESS (0/1: unreachable targets)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from general_motion_retargeting_b200 import GeneralMotionRetargeting, params
from general_motion_retargeting_b200.ik_config import compile_task_table
from general_motion_retargeting_b200.synthetic import make_clips
T = int(os.environ.get("PROBE_T", "100")); C = int(os.environ.get("PROBE_C", "4096"))
robot_name, src = os.environ.get("PROBE_ROBOT", "unitree_g1"), os.environ.get("PROBE_SRC", "smplx")
stress = os.environ.get("PROBE_STRESS", "0") == "1"
robot, cfg, _ = params.load_pack(src, robot_name)
table = compile_task_table(robot, cfg)
gmr = GeneralMotionRetargeting(src, robot_name, device=0)
b = make_clips(robot, table, range(C), T=T, src_human=src, stress=stress, device="cuda")
dp, dq, dh = (torch.from_numpy(x).cuda() for x in (b.pos, b.quat, b.heights))
out = {"mix": f"{robot_name}/{src} {C}x{T}" + (" stress" if stress else ""), "env": {k: v for k, v in os.environ.items() if k.startswith("GMR_") and k != "GMR_B200_LIB"}}
for prec in sys.argv[1:] or ["f32", "f64"]:
    for _ in range(2): gmr.retarget_batch(dp, dq, dh, precision=prec)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(3): gmr.retarget_batch(dp, dq, dh, precision=prec)
    e1.record(); torch.cuda.synchronize()
    out[prec] = round(e0.elapsed_time(e1) / 3, 2)
print(json.dumps(out), flush=True)
