#!/usr/bin/env python
"""Launches every kernel of the library other than the solve kernel once at a representative size (for ncu):
gmr_order_kernel (via a batch larger than the warp slots, two-phase off), gmr_classify_kernel (two-phase),
gmr_finalize_kernel (retarget_dataset), gmr_bvh_kernel / gmr_smplx_kernel (producers)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
from general_motion_retargeting_b200 import GeneralMotionRetargeting, params, producers
from general_motion_retargeting_b200.ik_config import compile_task_table
from general_motion_retargeting_b200.synthetic import make_clips
from make_golden_producers import LAFAN_BONES, LAFAN_PARENTS, SMPLX_NAMES, SMPLX_PARENTS
robot, cfg, _ = params.load_pack("smplx", "unitree_g1")
table = compile_task_table(robot, cfg)
gmr = GeneralMotionRetargeting("smplx", "unitree_g1", device=0)
clips = make_clips(robot, table, range(4096), T=20, device="cuda")
m = gmr.retarget_dataset(clips.pos, clips.quat, clips.heights, as_numpy=False)          # two-phase: classify + finalize
gmr.retarget_batch(torch.from_numpy(clips.pos[:, :8]).cuda(), torch.from_numpy(clips.quat[:, :8]).cuda(), torch.from_numpy(clips.heights).cuda())  # T < 16: order kernel
dev = torch.device("cuda", 0); g = torch.Generator(device=dev).manual_seed(1)
F, J = 4096 * 300, 22
q = torch.randn((F, J, 4), device=dev, generator=g); q = q / q.norm(dim=-1, keepdim=True); p = torch.randn((F, J, 3), device=dev, generator=g) * 20
producers.bvh_frames(q, p, LAFAN_PARENTS, LAFAN_BONES, GeneralMotionRetargeting("bvh", "unitree_g1", device=0).human_body_names)
F = 1024 * 1200
go = torch.randn((F, 3), device=dev, generator=g); fp = torch.randn((F, 55, 3), device=dev, generator=g) * 0.5; jt = torch.randn((F, 127, 3), device=dev, generator=g)
producers.smplx_frames(go, fp, jt, SMPLX_PARENTS, SMPLX_NAMES, gmr.human_body_names, 120.0, 30.0)
torch.cuda.synchronize(); print("ok")
