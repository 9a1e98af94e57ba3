#!/usr/bin/env python
"""Golden fixtures for the human-frame producers, made by running the REFERENCE's own
loader code in this container (/root/reference) on synthetic inputs:

 * LAFAN1/BVH: a synthetic 22-bone BVH file is written to /tmp and pushed through the reference's
   `load_lafan1_file` (utils/lafan1.py:8-41 -> lafan_vendor/extract.py:read_bvh, lafan_vendor/utils.py:quat_fk);
 * SMPL-X: `get_smplx_data_offline_fast` (utils/smpl.py:109-196) with mocked `smplx_output` / `body_model`
   objects (the licensed body model itself is not needed after its forward pass), once with 120 -> 30 fps
   resampling and once without.
The package __init__ of the reference imports mink; the modules are loaded individually with the absent third-party
packages (smplx, mink, mujoco, rich) stubbed.  Output: tests/golden/reference_producers.npz
"""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "reference_producers.npz")

LAFAN_BONES = ["Hips", "LeftUpLeg", "LeftLeg", "LeftFoot", "LeftToe", "RightUpLeg", "RightLeg", "RightFoot", "RightToe",
               "Spine", "Spine1", "Spine2", "Neck", "Head", "LeftShoulder", "LeftArm", "LeftForeArm", "LeftHand",
               "RightShoulder", "RightArm", "RightForeArm", "RightHand"]
LAFAN_PARENTS = [-1, 0, 1, 2, 3, 0, 5, 6, 7, 0, 9, 10, 11, 12, 11, 14, 15, 16, 11, 18, 19, 20]
SMPLX_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 15, 15, 15,
                 20, 25, 26, 20, 28, 29, 20, 31, 32, 20, 34, 35, 20, 37, 38,
                 21, 40, 41, 21, 43, 44, 21, 46, 47, 21, 49, 50, 21, 52, 53]
SMPLX_NAMES = ["pelvis", "left_hip", "right_hip", "spine1", "left_knee", "right_knee", "spine2", "left_ankle", "right_ankle",
               "spine3", "left_foot", "right_foot", "neck", "left_collar", "right_collar", "head", "left_shoulder",
               "right_shoulder", "left_elbow", "right_elbow", "left_wrist", "right_wrist", "jaw", "left_eye_smplhf",
               "right_eye_smplhf"] + [f"{s}_{f}{i}" for s in ("left", "right") for f in ("index", "middle", "pinky", "ring", "thumb") for i in (1, 2, 3)]


def load_reference_utils():
    pkg = types.ModuleType("general_motion_retargeting"); pkg.__path__ = [os.path.join(REF, "general_motion_retargeting")]
    sys.modules["general_motion_retargeting"] = pkg
    up = types.ModuleType("general_motion_retargeting.utils"); up.__path__ = [os.path.join(REF, "general_motion_retargeting", "utils")]
    sys.modules["general_motion_retargeting.utils"] = up
    lv = types.ModuleType("general_motion_retargeting.utils.lafan_vendor")
    lv.__path__ = [os.path.join(REF, "general_motion_retargeting", "utils", "lafan_vendor")]
    sys.modules["general_motion_retargeting.utils.lafan_vendor"] = lv
    smplx = types.ModuleType("smplx"); jn = types.ModuleType("smplx.joint_names"); jn.JOINT_NAMES = SMPLX_NAMES
    smplx.joint_names = jn; sys.modules["smplx"] = smplx; sys.modules["smplx.joint_names"] = jn

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, "general_motion_retargeting", rel))
        mod = importlib.util.module_from_spec(spec); sys.modules[name] = mod; spec.loader.exec_module(mod); return mod

    load("general_motion_retargeting.utils.lafan_vendor.utils", "utils/lafan_vendor/utils.py")
    load("general_motion_retargeting.utils.lafan_vendor.extract", "utils/lafan_vendor/extract.py")
    return load("general_motion_retargeting.utils.lafan1", "utils/lafan1.py"), load("general_motion_retargeting.utils.smpl", "utils/smpl.py")


def write_bvh(path, rng, nframes):
    children = {i: [j for j, p in enumerate(LAFAN_PARENTS) if p == i] for i in range(len(LAFAN_BONES))}
    offsets = rng.uniform(-25, 25, (len(LAFAN_BONES), 3)); offsets[0] = 0
    lines = ["HIERARCHY"]

    def emit(i, ind):
        pad = "  " * ind
        lines.append(f"{pad}{'ROOT' if i == 0 else 'JOINT'} {LAFAN_BONES[i]}")
        lines.append(pad + "{")
        lines.append(f"{pad}  OFFSET {offsets[i, 0]:.6f} {offsets[i, 1]:.6f} {offsets[i, 2]:.6f}")
        lines.append(f"{pad}  CHANNELS 6 Xposition Yposition Zposition Zrotation Yrotation Xrotation")
        if not children[i]:
            lines.extend([f"{pad}  End Site", pad + "  {", f"{pad}    OFFSET 0.000000 5.000000 0.000000", pad + "  }"])
        for c in children[i]:
            emit(c, ind + 1)
        lines.append(pad + "}")

    emit(0, 0)
    lines += ["MOTION", f"Frames: {nframes}", "Frame Time: 0.033333"]
    for f in range(nframes):
        row = []
        for i in range(len(LAFAN_BONES)):
            p = offsets[i] + (rng.uniform(-100, 100, 3) if i == 0 else rng.uniform(-0.5, 0.5, 3))
            r = rng.uniform(-80, 80, 3) if i else rng.uniform(-180, 180, 3)
            row += [f"{v:.6f}" for v in list(p) + list(r)]
        lines.append(" ".join(row))
    open(path, "w").write("\n".join(lines) + "\n")


def main():
    lafan1, smpl = load_reference_utils()
    rng = np.random.default_rng(77)
    out = {}
    # ---- LAFAN1 --------------------------------------------------------------------------------------------
    from general_motion_retargeting.utils.lafan_vendor.extract import read_bvh
    path = "/tmp/gmr_synth.bvh"
    write_bvh(path, rng, 9)
    frames, height = lafan1.load_lafan1_file(path)
    data = read_bvh(path)
    out["bvh.quats"] = np.asarray(data.quats, np.float64); out["bvh.pos"] = np.asarray(data.pos, np.float64)
    out["bvh.parents"] = np.asarray(data.parents, np.int32); out["bvh.bones"] = np.array(data.bones)
    names = sorted(frames[0].keys())
    out["bvh.out_names"] = np.array(names)
    out["bvh.out_pos"] = np.array([[np.asarray(fr[n][0], np.float64) for n in names] for fr in frames])
    out["bvh.out_quat"] = np.array([[np.asarray(fr[n][1], np.float64) for n in names] for fr in frames])
    out["bvh.height"] = np.array(height)
    # ---- SMPL-X --------------------------------------------------------------------------------------------
    import torch
    for tag, F, fps in (("smplx120", 41, 120.0), ("smplx30", 7, 30.0)):
        go = rng.normal(0, 1.2, (F, 3)); fp = rng.normal(0, 0.6, (F, 55, 3)); fp[:, 0] = go
        fp[2, 5] = 1e-4 * rng.normal(size=3)            # small-angle branch of from_rotvec
        fp[3, 4] = fp[4, 4] + 1e-3                       # nearly equal neighbours: the lerp branch of slerp
        jt = rng.normal(0, 1.0, (F, 127, 3))
        so = types.SimpleNamespace(global_orient=torch.from_numpy(go).float()[:, None, :], full_pose=torch.from_numpy(fp.reshape(F, -1)).float(),
                                   joints=torch.from_numpy(jt).float())
        bm = types.SimpleNamespace(parents=np.array(SMPLX_PARENTS))
        sd = {"mocap_frame_rate": np.array(fps), "pose_body": np.zeros((F, 63))}
        frames, afps = smpl.get_smplx_data_offline_fast(sd, bm, so, tgt_fps=30)
        out[f"{tag}.global_orient"] = go.astype(np.float32); out[f"{tag}.full_pose"] = fp.astype(np.float32); out[f"{tag}.joints"] = jt.astype(np.float32)
        out[f"{tag}.parents"] = np.array(SMPLX_PARENTS, np.int32); out[f"{tag}.names"] = np.array(SMPLX_NAMES); out[f"{tag}.src_fps"] = np.array(fps)
        out[f"{tag}.aligned_fps"] = np.array(afps)
        out[f"{tag}.out_pos"] = np.array([[np.asarray(fr[n][0], np.float64) for n in SMPLX_NAMES] for fr in frames])
        out[f"{tag}.out_quat"] = np.array([[np.asarray(fr[n][1], np.float64) for n in SMPLX_NAMES] for fr in frames])
        print(tag, "frames in", F, "out", len(frames), "aligned fps", afps)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; bvh frames", len(out["bvh.out_pos"]), "height", height)


if __name__ == "__main__":
    main()
