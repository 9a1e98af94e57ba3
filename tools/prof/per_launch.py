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
ts = []
for i in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.retarget_batch(dp, dq, dh, precision="f64"); e1.record(); torch.cuda.synchronize()
    ts.append(round(e0.elapsed_time(e1), 1))
print("per-launch ms (with a sync after each):", ts)
ts = []
evs = [torch.cuda.Event(enable_timing=True) for _ in range(9)]
evs[0].record()
for i in range(8):
    g.retarget_batch(dp, dq, dh, precision="f64"); evs[i + 1].record()
torch.cuda.synchronize()
print("per-launch ms (back to back):", [round(evs[i].elapsed_time(evs[i + 1]), 1) for i in range(8)])
