#!/usr/bin/env python
"""Dense SMs on NORMAL clips only (one clip per warp slot, all different): does a CTA-wide rendezvous per factorisation
(GMR_CONVOY=1: the warps of an SM walk through the code together, one instruction stream per SM instead of 16) pay off?"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from general_motion_retargeting_b200 import GeneralMotionRetargeting, params
from general_motion_retargeting_b200.ik_config import compile_task_table
from general_motion_retargeting_b200.synthetic import make_clips
T = int(os.environ.get("PROBE_T", "100")); prec = sys.argv[1] if len(sys.argv) > 1 else "f64"
robot, cfg, _ = params.load_pack("smplx", "unitree_g1")
table = compile_task_table(robot, cfg)
g = GeneralMotionRetargeting("smplx", "unitree_g1", device=0)
clips = make_clips(robot, table, range(4096), T=T, device="cuda")
dp, dq, dh = (torch.from_numpy(x).cuda() for x in (clips.pos, clips.quat, clips.heights))
err = np.load("/tmp/convoy_ids.npy") if os.path.exists("/tmp/convoy_ids.npy") else None
if err is None:
    q, it, e = g.retarget_batch(dp, dq, dh, return_info=True, precision=prec)
    tot = it.sum(-1).sum(-1).cpu().numpy()
    err = np.argsort(tot)[:148 * 16]
    np.save("/tmp/convoy_ids.npy", err)
n = 148 * (16 if prec == "f64" else 28)
ii = torch.from_numpy(np.sort(err[:n])).cuda()
rp, rq, rh = dp[ii].contiguous(), dq[ii].contiguous(), dh[ii].contiguous()
for _ in range(2): g.retarget_batch(rp, rq, rh, precision=prec)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(3): g.retarget_batch(rp, rq, rh, precision=prec)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"convoy": os.environ.get("GMR_CONVOY", "0"), "clips": int(ii.numel()), "ms": round(e0.elapsed_time(e1) / 3, 3)}))
