#!/usr/bin/env python
"""Times the benchmark mix (4096 different G1 clips x T frames) once per precision; knobs come from the environment."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from general_motion_retargeting_b200 import GeneralMotionRetargeting, params
from general_motion_retargeting_b200.ik_config import compile_task_table
from general_motion_retargeting_b200.synthetic import make_clips
T = int(os.environ.get("PROBE_T", "100")); C = int(os.environ.get("PROBE_C", "4096"))
robot, cfg, _ = params.load_pack("smplx", "unitree_g1")
table = compile_task_table(robot, cfg)
gmr = GeneralMotionRetargeting("smplx", "unitree_g1", device=0)
cache = "/tmp/mix_%d_%d.npz" % (C, T)
if os.path.exists(cache):
    z = np.load(cache); pos, quat, h = z["pos"], z["quat"], z["h"]
else:
    b = make_clips(robot, table, range(C), T=T, device="cuda"); pos, quat, h = b.pos, b.quat, b.heights
    np.savez(cache, pos=pos, quat=quat, h=h)
dp, dq, dh = (torch.from_numpy(x).cuda() for x in (pos, quat, h))
out = {"env": {k: v for k, v in os.environ.items() if k.startswith("GMR_")}}
for prec in sys.argv[1:] or ["f32", "f64"]:
    for _ in range(2): gmr.retarget_batch(dp, dq, dh, precision=prec)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(3): gmr.retarget_batch(dp, dq, dh, precision=prec)
    e1.record(); torch.cuda.synchronize()
    out[prec] = round(e0.elapsed_time(e1) / 3, 2)
print(json.dumps(out), flush=True)
