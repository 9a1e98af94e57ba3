#!/usr/bin/env python
"""Does it matter whether the warps of one SM run the SAME clip (lock step: one instruction stream per SM) or DIFFERENT clips
(k instruction streams per SM)?  Same multiset of clips, same k per SM, two placements.  Isolates instruction-fetch / divergence
effects from occupancy effects (tools/prof/slow_curve.py measures the lock-step case only)."""
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
clips = make_clips(robot, table, range(1024), T=T, device="cuda")
dp, dq, dh = (torch.from_numpy(x).cuda() for x in (clips.pos, clips.quat, clips.heights))
q, it, err = g.retarget_batch(dp, dq, dh, return_info=True, precision=prec)
tot = it.sum(-1).sum(-1).cpu().numpy()
order = np.argsort(-tot)
for name, pool in (("slow", order[:16]), ("normal", order[500:516])):
    for k in (4, 8, 16):
        ids_same = np.array([pool[b % len(pool)] for w in range(k) for b in range(148)])          # clip index p = w * 148 + b
        ids_diff = np.array([pool[(w + b) % len(pool)] for w in range(k) for b in range(148)])
        res = {}
        for tag, ids in (("same_clip_per_sm", ids_same), ("different_clips_per_sm", ids_diff)):
            ii = torch.from_numpy(ids).cuda()
            rp, rq, rh = dp[ii].contiguous(), dq[ii].contiguous(), dh[ii].contiguous()
            for _ in range(2): g.retarget_batch(rp, rq, rh, precision=prec)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record(); g.retarget_batch(rp, rq, rh, precision=prec); e1.record(); torch.cuda.synchronize()
            res[tag] = round(e0.elapsed_time(e1), 3)
        print(json.dumps({"clips": name, "per_sm": k, "solves_max": int(tot[pool].max()), "solves_mean": float(tot[pool].mean()), **res}), flush=True)
