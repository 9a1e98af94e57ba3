#!/usr/bin/env python
"""Per-call device times of the benchmark mix under a few variants (same box, same process):
   on_error raise/ignore, lie_eps 1e-10 / 2.2e-15.  PROBE_T / PROBE_C as in mix_case.py."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from general_motion_retargeting_b200 import GeneralMotionRetargeting, params
from general_motion_retargeting_b200.ik_config import compile_task_table
from general_motion_retargeting_b200.synthetic import make_clips
T = int(os.environ.get("PROBE_T", "300")); C = int(os.environ.get("PROBE_C", "4096")); N = int(os.environ.get("PROBE_N", "6"))
robot, cfg, _ = params.load_pack("smplx", "unitree_g1")
table = compile_task_table(robot, cfg)
b = make_clips(robot, table, range(C), T=T, device="cuda")
dp, dq, dh = (torch.from_numpy(x).cuda() for x in (b.pos, b.quat, b.heights))
for prec in sys.argv[1:] or ["f64"]:
    for lie in (None, 2.220446049250313e-15):
        gmr = GeneralMotionRetargeting("smplx", "unitree_g1", device=0, lie_eps=lie)
        for mode in ("raise", "ignore"):
            for _ in range(2): gmr.retarget_batch(dp, dq, dh, precision=prec, on_error=mode)
            torch.cuda.synchronize()
            ts = []
            for _ in range(N):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); gmr.retarget_batch(dp, dq, dh, precision=prec, on_error=mode); e1.record()
                torch.cuda.synchronize(); ts.append(round(e0.elapsed_time(e1), 2))
            print(json.dumps({"prec": prec, "lie": lie or 1e-10, "on_error": mode, "ms": ts}), flush=True)
