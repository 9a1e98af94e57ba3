#!/usr/bin/env python
"""Speed of ONE clip as a function of how many copies share an SM (k copies per SM on all SMs): the machine curve behind the
partition of the two-phase schedule.  Measured for a normal clip and for the slowest clip of the benchmark set."""
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
clips = make_clips(robot, table, range(512), T=T, device="cuda")
dp, dq, dh = (torch.from_numpy(x).cuda() for x in (clips.pos, clips.quat, clips.heights))
q, it, err = g.retarget_batch(dp, dq, dh, return_info=True, precision=prec)
tot = it.sum(-1).sum(-1).cpu().numpy()
cs, cn = int(np.argmax(tot)), int(np.argsort(tot)[len(tot) // 2])
for name, c in (("normal", cn), ("slow", cs)):
    for k in (1, 2, 4, 6, 8, 10, 12, 16):
        n = 148 * k
        rp, rq, rh = dp[c:c+1].repeat(n, 1, 1, 1), dq[c:c+1].repeat(n, 1, 1, 1), dh[c:c+1].repeat(n)
        for _ in range(2): g.retarget_batch(rp, rq, rh, precision=prec)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); g.retarget_batch(rp, rq, rh, precision=prec); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(json.dumps({"clip": name, "id": c, "solves": int(tot[c]), "copies_per_sm": k, "ms": round(ms, 3), "us_per_solve": round(ms * 1e3 / tot[c], 2),
                          "sm_solves_per_ms": round(k * tot[c] / ms, 1)}), flush=True)
