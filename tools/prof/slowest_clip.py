import sys, json
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from general_motion_retargeting_b200 import GeneralMotionRetargeting, params
from general_motion_retargeting_b200.ik_config import compile_task_table
from general_motion_retargeting_b200.synthetic import make_clips
robot, cfg, _ = params.load_pack("smplx", "unitree_g1")
table = compile_task_table(robot, cfg)
g = GeneralMotionRetargeting("smplx", "unitree_g1", device=0)
clips = make_clips(robot, table, range(4096), T=300, device="cuda")
dp, dq, dh = (torch.from_numpy(x).cuda() for x in (clips.pos, clips.quat, clips.heights))
for prec in ("f64", "f32"):
    q, it, err = g.retarget_batch(dp, dq, dh, return_info=True, precision=prec)
    tot = it.sum(-1).sum(-1).cpu().numpy()
    order = np.argsort(-tot)
    print(prec, "top solves", tot[order[:8]].tolist(), "clips", order[:8].tolist(), "mean", tot.mean(), "n>1800", int((tot > 1800).sum()), "n>1300", int((tot > 1300).sum()))
    # lone time of the slowest clips: each replicated on 148 SMs
    for c in order[:3]:
        rp, rq, rh = dp[c:c+1].repeat(148, 1, 1, 1), dq[c:c+1].repeat(148, 1, 1, 1), dh[c:c+1].repeat(148)
        for _ in range(2): g.retarget_batch(rp, rq, rh, precision=prec)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); g.retarget_batch(rp, rq, rh, precision=prec); e1.record(); torch.cuda.synchronize()
        print("   clip", int(c), "solves", int(tot[c]), "lone ms", round(e0.elapsed_time(e1), 2))
    for _ in range(2): g.retarget_batch(dp, dq, dh, precision=prec)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); g.retarget_batch(dp, dq, dh, precision=prec); e1.record(); torch.cuda.synchronize()
    print("   full batch ms", round(e0.elapsed_time(e1), 2))
