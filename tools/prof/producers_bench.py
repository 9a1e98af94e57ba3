#!/usr/bin/env python
"""Throughput of the human-frame producer kernels against the HBM roofline (they move ~1 KB per frame and do a
few hundred flops), with the NumPy restatement of the reference's loader arithmetic timed beside them on a
bounded sample.  One JSON line per producer."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from general_motion_retargeting_b200 import producers
from oracle import producers_oracle as P
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_golden_producers import LAFAN_BONES, LAFAN_PARENTS, SMPLX_NAMES, SMPLX_PARENTS

peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm = peaks.get("hbm_gbs", 6650.0)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)

def timeit(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

# ---- BVH: 4096 clips x 300 frames ----------------------------------------------------------------------------
F, J = 4096 * 300, 22
bvh_bodies = ["Hips", "Spine2", "LeftUpLeg", "RightUpLeg", "LeftLeg", "RightLeg", "LeftFootMod", "RightFootMod",
              "LeftArm", "RightArm", "LeftForeArm", "RightForeArm", "LeftHand", "RightHand"]
q = torch.randn((F, J, 4), device=dev, generator=g); q = q / q.norm(dim=-1, keepdim=True)
p = torch.randn((F, J, 3), device=dev, generator=g) * 20
ms = timeit(lambda: producers.bvh_frames(q, p, LAFAN_PARENTS, LAFAN_BONES, bvh_bodies))
nh = len(bvh_bodies)
lanes = 19                                      # joints on the chains of the 14 bodies (+ toes)
alg = F * (lanes * 7 + nh * 7) * 4
S = 20000
t0 = time.perf_counter(); P.bvh_frames(q[:S].cpu().numpy(), p[:S].cpu().numpy(), LAFAN_PARENTS, list(range(nh)), list(range(nh))); cpu = S / (time.perf_counter() - t0)
print(json.dumps({"producer": "bvh (utils/lafan1.py:17-35)", "frames": F, "ms": ms, "frames_per_s": F / ms * 1e3,
                  "roofline": {"bound": "hbm", "achieved": alg / ms / 1e6, "peak": hbm, "unit": "GB/s", "frac": alg / ms / 1e6 / hbm,
                               "algorithmic_bytes_per_frame": alg // F},
                  "cpu_baseline": {"value": cpu, "unit": "frames/s", "cores": 1, "kind": "port", "sample": f"{S} frames, vectorised NumPy restatement"}}), flush=True)
del q, p
# ---- SMPL-X: 120 fps -> 30 fps, 1024 clips x 1200 source frames ------------------------------------------------
F, NJ, NJo = 1024 * 1200, 55, 127
go = torch.randn((F, 3), device=dev, generator=g); fp = torch.randn((F, NJ, 3), device=dev, generator=g) * 0.5; jt = torch.randn((F, NJo, 3), device=dev, generator=g)
bodies = ["pelvis", "spine3", "left_hip", "right_hip", "left_knee", "right_knee", "left_foot", "right_foot", "left_shoulder",
          "right_shoulder", "left_elbow", "right_elbow", "left_wrist", "right_wrist"]
ms = timeit(lambda: producers.smplx_frames(go, fp, jt, SMPLX_PARENTS, SMPLX_NAMES, bodies, 120.0, 30.0))
Fo = F // 4
lanes = 20
alg = Fo * (2 * lanes * 3 + 2 * 14 * 3 + 14 * 7) * 4
S = 8000
a = (go[:S].cpu().numpy(), fp[:S].cpu().numpy(), jt[:S].cpu().numpy())
t0 = time.perf_counter(); P.smplx_frames(*a, SMPLX_PARENTS, 120.0, 30.0, list(range(22))); cpu = (S // 4) / (time.perf_counter() - t0)
print(json.dumps({"producer": "smplx 120->30 fps (utils/smpl.py:127-196)", "frames_in": F, "frames_out": Fo, "ms": ms, "frames_out_per_s": Fo / ms * 1e3,
                  "roofline": {"bound": "hbm", "achieved": alg / ms / 1e6, "peak": hbm, "unit": "GB/s", "frac": alg / ms / 1e6 / hbm,
                               "algorithmic_bytes_per_output_frame": alg // Fo},
                  "cpu_baseline": {"value": cpu, "unit": "output frames/s", "cores": 1, "kind": "port", "sample": f"{S} source frames, vectorised NumPy restatement"}}), flush=True)
