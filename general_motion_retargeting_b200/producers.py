"""Human-frame producers on the GPU: what the reference's loaders compute after parsing
a file, for whole batches of frames, emitted directly in the layout `retarget_batch` / `retarget_dataset` take
(`pos [F,nh,3]`, `quat [F,nh,4]` wxyz, bodies ordered as `GeneralMotionRetargeting.human_body_names`).

 * `bvh_frames`   — reference utils/lafan1.py:17-35 (`quat_fk`, Y-up -> Z-up, cm -> m, `LeftFootMod`/`RightFootMod`)
                    from the arrays `read_bvh` returns (`Anim.quats`, `Anim.pos`, `Anim.parents`, `Anim.bones`).
 * `smplx_frames` — reference utils/smpl.py:127-196 (30 fps resampling with SLERP / linear interpolation, global
                    joint-orientation chain) from the SMPL-X body model's outputs (`global_orient`, `full_pose`,
                    `joints`) and its `parents`.
File parsing (BVH text, npz) and the licensed SMPL-X body-model forward stay on the host / out of scope.
There is no CPU fallback: the functions need the CUDA library and a device.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

from . import _native

# the synthesised bodies of utils/lafan1.py:32-33: (position joint, orientation joint)
BVH_SYNTHESISED = {"LeftFootMod": ("LeftFoot", "LeftToe"), "RightFootMod": ("RightFoot", "RightToe")}
BVH_HUMAN_HEIGHT = 1.75            # utils/lafan1.py:39


def _check(lib, rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {lib.gmr_last_error().decode()}")


def _i32(a):
    return np.ascontiguousarray(np.asarray(a), np.int32)


def bvh_frames(quats, pos, parents: Sequence[int], bones: Sequence[str], body_names: Sequence[str], device=None):
    """quats [F,J,4] wxyz / pos [F,J,3] cm local (any array-like or CUDA tensor) -> (pos [F,nh,3], quat [F,nh,4])
    float32 CUDA tensors for `body_names` (KeyError for an unknown bone, like the reference's dict lookup)."""
    import torch
    lib = _native.load_library()
    bones = list(bones)
    pj, rj = [], []
    for n in body_names:
        a, b = BVH_SYNTHESISED.get(n, (n, n))
        if a not in bones or b not in bones:
            raise KeyError(n)
        pj.append(bones.index(a)); rj.append(bones.index(b))
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(dev):
        q = torch.as_tensor(quats).to(dev, torch.float32).contiguous()
        p = torch.as_tensor(pos).to(dev, torch.float32).contiguous()
        F, J = int(q.shape[0]), int(q.shape[1])
        if tuple(q.shape) != (F, J, 4) or tuple(p.shape) != (F, J, 3) or len(parents) != J:
            raise ValueError("expected quats [F,J,4], pos [F,J,3] and J parents")
        nh = len(body_names)
        out_p = torch.empty((F, nh, 3), dtype=torch.float32, device=dev)
        out_q = torch.empty((F, nh, 4), dtype=torch.float32, device=dev)
        par, pj_, rj_ = _i32(parents), _i32(pj), _i32(rj)
        rc = lib.gmr_produce_bvh_frames(q.data_ptr(), p.data_ptr(), par.ctypes.data, F, J, pj_.ctypes.data, rj_.ctypes.data, nh,
                                        out_p.data_ptr(), out_q.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        _check(lib, rc, "gmr_produce_bvh_frames")
    return out_p, out_q


def smplx_frames(global_orient, full_pose, joints, parents: Sequence[int], joint_names: Sequence[str],
                 body_names: Sequence[str], src_fps: float, tgt_fps: float = 30.0, device=None) -> Tuple[object, object, float]:
    """global_orient [F,3], full_pose [F,NJ,3] (or [F,NJ*3]) axis-angle, joints [F,>=NJ,3] -> (pos [F',nh,3],
    quat [F',nh,4], aligned_fps): F' = F // int(src_fps / tgt_fps) frames when tgt_fps < src_fps (utils/smpl.py:127-176),
    else F' = F."""
    import torch
    lib = _native.load_library()
    names = list(joint_names)
    bj = []
    for n in body_names:
        if n not in names:
            raise KeyError(n)
        bj.append(names.index(n))
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(dev):
        go = torch.as_tensor(global_orient).to(dev, torch.float32).reshape(-1, 3).contiguous()
        F = int(go.shape[0])
        fp = torch.as_tensor(full_pose).to(dev, torch.float32).reshape(F, -1, 3).contiguous()
        jt = torch.as_tensor(joints).to(dev, torch.float32).reshape(F, -1, 3).contiguous()
        NJ, NJo = int(fp.shape[1]), int(jt.shape[1])
        if len(parents) != NJ:
            raise ValueError(f"full_pose has {NJ} joints but {len(parents)} parents were given")
        if tgt_fps < src_fps:
            Fo = F // int(src_fps / tgt_fps)
            aligned = Fo / F * src_fps
        else:
            Fo, aligned = F, tgt_fps
        nh = len(body_names)
        out_p = torch.empty((Fo, nh, 3), dtype=torch.float32, device=dev)
        out_q = torch.empty((Fo, nh, 4), dtype=torch.float32, device=dev)
        par, bj_ = _i32(list(parents)[:NJ]), _i32(bj)
        rc = lib.gmr_produce_smplx_frames(go.data_ptr(), fp.data_ptr(), jt.data_ptr(), par.ctypes.data, F, NJ,
                                          NJo, Fo, bj_.ctypes.data, nh, out_p.data_ptr(), out_q.data_ptr(),
                                          torch.cuda.current_stream(dev).cuda_stream)
        _check(lib, rc, "gmr_produce_smplx_frames")
    return out_p, out_q, aligned
