#!/usr/bin/env python
"""Only the suspects (error > 1 after frame 0) of the benchmark batch, scheduled as in the full batch (GMR_FORCE_SCHED=1,
8 per sparse SM) but with the dense SMs EMPTY: separates what the placement costs a slow clip from what the rest of the chip
costs it.  GMR_NO_STEAL=1 keeps them on their sparse SMs; without it they spread out at their first segment boundary."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from general_motion_retargeting_b200 import GeneralMotionRetargeting, params
from general_motion_retargeting_b200.ik_config import compile_task_table
from general_motion_retargeting_b200.synthetic import make_clips
T = int(os.environ.get("PROBE_T", "300")); prec = sys.argv[1] if len(sys.argv) > 1 else "f64"
robot, cfg, _ = params.load_pack("smplx", "unitree_g1")
table = compile_task_table(robot, cfg)
g = GeneralMotionRetargeting("smplx", "unitree_g1", device=0)
clips = make_clips(robot, table, range(4096), T=T, device="cuda")
dp, dq, dh = (torch.from_numpy(x).cuda() for x in (clips.pos, clips.quat, clips.heights))
q, it, err = g.retarget_batch(dp, dq, dh, return_info=True, precision=prec)
sus = torch.nonzero(err[:, 0, 1] > 1.0).flatten()
tot = it.sum(-1).sum(-1)
# pad with clips of length 1 (they end with launch A's frame 0): the batch keeps the full batch's launch geometry
npad = 2400
ids = torch.cat([sus, torch.arange(npad, device="cuda") % 4096])
rp, rq, rh = dp[ids].contiguous(), dq[ids].contiguous(), dh[ids].contiguous()
lengths = np.concatenate([np.full(sus.numel(), T), np.ones(npad)]).astype(np.int32)
run = lambda: g.retarget_dataset(rp, rq, rh, lengths=lengths, precision=prec, as_numpy=False)
for _ in range(2): run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(); run(); e1.record(); torch.cuda.synchronize()
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("GMR_")}, "suspects": int(sus.numel()), "max_solves": int(tot[sus].max()),
                  "ms": round(e0.elapsed_time(e1), 2)}))
