#!/usr/bin/env python
"""Per-frame latency of the live-stream entry (GeneralMotionRetargeting.retarget -> gmr_stream_retarget)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from general_motion_retargeting_b200 import GeneralMotionRetargeting, params
from general_motion_retargeting_b200.ik_config import compile_task_table
from general_motion_retargeting_b200.synthetic import make_clips
robot, cfg, _ = params.load_pack("smplx", "unitree_g1")
table = compile_task_table(robot, cfg)
clips = make_clips(robot, table, [0], T=300)
gmr = GeneralMotionRetargeting("smplx", "unitree_g1", actual_human_height=float(clips.heights[0]), device=0)
frames = [{n: (clips.pos[0, t, i], clips.quat[0, t, i]) for i, n in enumerate(table.human_names)} for t in range(300)]
lat = []
for rep in range(3):
    gmr.setup_retarget_configuration()
    for t, f in enumerate(frames):
        t0 = time.perf_counter(); gmr.retarget(f); dt = time.perf_counter() - t0
        if rep > 0 and t > 0: lat.append(dt)
lat = np.array(lat) * 1e6
# the C call alone (no dict packing)
import ctypes as C
pos = np.ascontiguousarray(clips.pos[0]); quat = np.ascontiguousarray(clips.quat[0]); q = np.empty(robot.nq)
gmr.setup_retarget_configuration(); gmr.retarget(frames[0])
raw = []
for t in range(1, 300):
    t0 = time.perf_counter()
    gmr._lib.gmr_stream_retarget(gmr._stream.ptr, pos[t].ctypes.data, quat[t].ctypes.data, 0, q.ctypes.data, None, None, None)
    raw.append(time.perf_counter() - t0)
raw = np.array(raw) * 1e6
print(json.dumps({"case": "live stream, unitree_g1/smplx, 299 warm frames", "python_retarget_us": {"median": float(np.median(lat)), "p95": float(np.percentile(lat, 95))},
                  "c_abi_call_us": {"median": float(np.median(raw)), "p95": float(np.percentile(raw, 95))},
                  "frames_per_s_single_stream": float(1e6 / np.median(lat)), "reference_published_fps": "35-70 (README.md:215-221)"}))
