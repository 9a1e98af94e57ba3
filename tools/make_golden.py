#!/usr/bin/env python
"""Generate tests/golden/*.npz.

Two kinds of fixtures:
 (1) outputs of the REFERENCE's own code run in this container (/root/reference), used to pin
     the oracle: `KinematicsModel.forward_kinematics` (kinematics_model.py:213-246) and the
     target preprocessing methods of `GeneralMotionRetargeting` (motion_retarget.py:209-270),
     the latter executed unmodified with the absent `mink`/`mujoco`/`rich` modules stubbed
     (those three methods never touch them);
 (2) float64 oracle traces on the synthetic inputs, used as known-answer vectors for the CUDA
     path on the GPU box (where /root/reference does not exist).
Usage: python tools/make_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

from general_motion_retargeting_b200 import params  # noqa: E402
from general_motion_retargeting_b200.ik_config import compile_task_table  # noqa: E402
from general_motion_retargeting_b200.synthetic import make_clips  # noqa: E402
from general_motion_retargeting_b200._native import LIE_EPS_ROUND1  # noqa: E402


def load_reference_modules():
    """Import the reference's kinematics_model / motion_retarget without its package __init__."""
    pkg = types.ModuleType("general_motion_retargeting")
    pkg.__path__ = [os.path.join(REF, "general_motion_retargeting")]
    sys.modules["general_motion_retargeting"] = pkg
    for stub in ("mink", "mujoco", "rich"):
        if stub not in sys.modules:
            m = types.ModuleType(stub)
            if stub == "rich":
                m.print = print
            sys.modules[stub] = m

    def load(name):
        spec = importlib.util.spec_from_file_location(f"general_motion_retargeting.{name}",
                                                      os.path.join(REF, "general_motion_retargeting", f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        return mod

    load("params")
    load("torch_utils")
    return load("kinematics_model"), load("motion_retarget")


def golden_reference_fk(km_mod):
    import torch
    out = {}
    rng = np.random.default_rng(123)
    for robot, rel in params.ROBOT_XML_REL.items():
        xml = os.path.join(REF, "assets", rel)
        try:
            km = km_mod.KinematicsModel(xml, "cpu")
        except AssertionError as e:          # pm01: nested <include>, kinematics_model.py:105
            print(f"  reference KinematicsModel cannot parse {robot}: {e!r}")
            continue
        lo, hi = km.get_dof_limits()
        T = 6
        dof = (lo + (hi - lo) * torch.from_numpy(rng.uniform(0.1, 0.9, (T, km.num_dof))).float()).float()
        root_pos = torch.from_numpy(rng.uniform(-1, 1, (T, 3))).float()
        q = rng.normal(size=(T, 4)); q /= np.linalg.norm(q, axis=-1, keepdims=True)
        root_rot_xyzw = torch.from_numpy(q[:, [1, 2, 3, 0]]).float()
        bp, br = km.forward_kinematics(root_pos, root_rot_xyzw, dof)
        out[f"{robot}.body_names"] = np.array(km.body_names)
        out[f"{robot}.parent"] = km.parent_indices.numpy()
        out[f"{robot}.dof"] = dof.numpy()
        out[f"{robot}.root_pos"] = root_pos.numpy()
        out[f"{robot}.root_quat_wxyz"] = q.astype(np.float32)
        out[f"{robot}.body_pos"] = bp.numpy()
        out[f"{robot}.body_rot_xyzw"] = br.numpy()
        print(f"  {robot}: {len(km.body_names)} bodies, {km.num_dof} dof")
    np.savez_compressed(os.path.join(OUT, "reference_fk.npz"), **out)


def golden_reference_preprocess(mr_mod):
    """scale_human_data / offset_human_data / offset_human_data_to_ground of the reference class,
    called on an instance created without __init__ (they only use their arguments)."""
    import json
    from scipy.spatial.transform import Rotation as R
    out = {}
    G = mr_mod.GeneralMotionRetargeting
    for src, robot in (("smplx", "unitree_g1"), ("bvh", "booster_t1"), ("smplx", "hightorque_hi")):
        with open(os.path.join(REF, "general_motion_retargeting", "ik_configs", params.IK_CONFIG_REL[src][robot])) as f:
            cfg = json.load(f)
        m, c, _ = params.load_pack(src, robot)
        tt = compile_task_table(m, c)
        clips = make_clips(m, tt, [7], T=3, src_human=src)
        height = float(clips.heights[0])
        ratio = height / cfg["human_height_assumption"]
        table = {k: v * ratio for k, v in cfg["human_scale_table"].items()}
        ground = cfg["ground_height"] * np.array([0, 0, 1])
        pos_off, rot_off = {}, {}
        for frame_name, entry in cfg["ik_match_table1"].items():
            body, pw, rw, po, ro = entry
            if pw != 0 or rw != 0:
                pos_off[body] = np.array(po) - ground
                rot_off[body] = R.from_quat(ro, scalar_first=True)
        g = object.__new__(G)
        res = np.zeros((3, tt.nh, 7)); res_g = np.zeros((3, tt.nh, 7))
        for t in range(3):
            hd = {n: (clips.pos[0, t, i].astype(np.float64), clips.quat[0, t, i].astype(np.float64)) for i, n in enumerate(tt.human_names)}
            hd["extra_body_not_in_table"] = (np.zeros(3), np.array([1.0, 0, 0, 0]))
            hd = g.to_numpy(hd)
            s = g.scale_human_data(hd, cfg["human_root_name"], table)
            o = g.offset_human_data(s, pos_off, rot_off)
            og = g.offset_human_data_to_ground(o)
            for i, n in enumerate(tt.human_names):
                res[t, i, :3], res[t, i, 3:] = o[n][0], o[n][1]
                res_g[t, i, :3], res_g[t, i, 3:] = og[n][0], og[n][1]
        key = f"{src}_{robot}"
        out[key + ".pos"], out[key + ".quat"], out[key + ".height"] = clips.pos[0], clips.quat[0], np.array(height)
        out[key + ".targets"], out[key + ".targets_ground"] = res, res_g
        print(f"  preprocessing golden for {key}")
    np.savez_compressed(os.path.join(OUT, "reference_preprocess.npz"), **out)


def golden_oracle_traces():
    from oracle import native
    out = {}
    for src, robot, ids, T, stress in (("smplx", "unitree_g1", [0, 1, 15], 40, False),
                                       ("bvh", "booster_t1", [3, 4], 30, False),
                                       ("smplx", "hightorque_hi", [5], 30, False),
                                       ("smplx", "stanford_toddy", [2, 9], 25, True),
                                       ("smplx", "kuavo_s45", [6], 25, False)):
        m, c, _ = params.load_pack(src, robot)
        tt = compile_task_table(m, c)
        clips = make_clips(m, tt, ids, T=T, src_human=src, stress=stress)
        # the Lie threshold is a parameter (GmrModelDesc.lie_eps): traces for the default (1e-10) and for round 1's 10 eps
        q, it, err = native.retarget_batch(m, tt, clips.pos, clips.quat, clips.ratio(tt))
        q1, it1, err1 = native.retarget_batch(m, tt, clips.pos, clips.quat, clips.ratio(tt), lie_eps=LIE_EPS_ROUND1)
        key = f"{src}_{robot}"
        out[key + ".pos"], out[key + ".quat"], out[key + ".heights"] = clips.pos, clips.quat, clips.heights
        out[key + ".qpos"], out[key + ".iters"], out[key + ".err"] = q, it, err
        out[key + ".qpos_eps10"], out[key + ".iters_eps10"], out[key + ".err_eps10"] = q1, it1, err1
        print(f"  oracle trace {key}: solves/frame {it.sum(-1).mean():.2f}")
    np.savez_compressed(os.path.join(OUT, "oracle_traces.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if "--oracle-only" not in sys.argv:
        km_mod, mr_mod = load_reference_modules()
        print("reference FK:"); golden_reference_fk(km_mod)
        print("reference preprocessing:"); golden_reference_preprocess(mr_mod)
    print("oracle traces:"); golden_oracle_traces()
