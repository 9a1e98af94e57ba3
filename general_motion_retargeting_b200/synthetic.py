"""Deterministic synthetic human-keypoint clips (SURVEY.md §8d "synthetic input generator").

There are no datasets in the build/bench environment (SMPL-X body models are licensed,
LAFAN1 needs a download), so benchmark and parity inputs are synthesised: a smooth,
in-limit robot trajectory is pushed through forward kinematics, the pose of every
table-1 robot frame is mapped *backwards* through the reference's target preprocessing
(inverse of `offset_human_data`, then of `scale_human_data`; reference
general_motion_retargeting/motion_retarget.py:209-250) and perturbed with noise.  The
result has the exact format the reference's loaders emit (per body: position in metres,
world Z-up; quaternion wxyz) and is reachable up to the noise, like real mocap after
scaling.

clip `i` uses `numpy.random.Generator(PCG64(seed0 + i))`; the draw order below is part of
the contract (tests/golden was generated with it).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np

from .ik_config import TaskTable
from .mjcf import RobotModel

SEED0 = 20260000
FPS = 30.0


# ---- small batched quaternion helpers (wxyz, arrays [...,4]) ---------------------------
def _qmul(a, b):
    aw, ax, ay, az = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bw, bx, by, bz = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ], axis=-1)


def _qconj(q):
    return q * np.array([1.0, -1.0, -1.0, -1.0])


def _qrot(q, v):
    w = q[..., 0:1]
    u = q[..., 1:4]
    t = 2.0 * np.cross(u, v)
    return v + w * t + np.cross(u, t)


def _axis_angle(axis, angle):
    half = 0.5 * angle
    return np.concatenate([np.cos(half)[..., None], axis * np.sin(half)[..., None]], axis=-1)


def batched_fk(robot: RobotModel, qpos: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """World pose of every body for qpos[N,nq] → (xpos[N,nb,3], xquat[N,nb,4])."""
    N = qpos.shape[0]
    nb = robot.nbody
    xpos = np.zeros((N, nb, 3))
    xquat = np.zeros((N, nb, 4))
    for b in range(nb):
        p = int(robot.parent[b])
        if p < 0:
            xpos[:, b] = qpos[:, 0:3]
            q = qpos[:, 3:7]
            xquat[:, b] = q / np.linalg.norm(q, axis=-1, keepdims=True)
            continue
        xpos[:, b] = xpos[:, p] + _qrot(xquat[:, p], np.broadcast_to(robot.body_pos[b], (N, 3)))
        q = _qmul(xquat[:, p], np.broadcast_to(robot.body_quat[b], (N, 4)))
        j = int(robot.body_hinge[b])
        if j >= 0:
            q = _qmul(q, _axis_angle(np.broadcast_to(robot.hinge_axis[j], (N, 3)), qpos[:, 7 + j]))
        xquat[:, b] = q / np.linalg.norm(q, axis=-1, keepdims=True)
    return xpos, xquat


@dataclass
class ClipBatch:
    pos: np.ndarray       # [C,T,nh,3] float32
    quat: np.ndarray      # [C,T,nh,4] float32 wxyz
    heights: np.ndarray   # [C] float32 actual_human_height
    qpos_gen: np.ndarray  # [C,T,nq] float64 trajectory the targets were generated from


def make_clips(robot: RobotModel, table: TaskTable, clip_ids: Sequence[int], T: int = 300,
               src_human: str = "smplx", stress: bool = False, seed0: int = SEED0,
               fixed_height: Optional[float] = None) -> ClipBatch:
    clip_ids = list(clip_ids)
    C = len(clip_ids)
    nhinge, nh = robot.nhinge, table.nh
    t = np.arange(T) / FPS

    lo = np.where(robot.hinge_limited, robot.hinge_lo, -np.pi)
    hi = np.where(robot.hinge_limited, robot.hinge_hi, np.pi)
    mid, half = 0.5 * (lo + hi), 0.5 * (hi - lo)
    pos_sigma, rot_sigma = (0.05, 0.3) if stress else (0.005, 0.02)

    qpos = np.zeros((C, T, robot.nq))
    heights = np.zeros(C)
    noise_p = np.zeros((C, T, nh, 3))
    noise_axis = np.zeros((C, T, nh, 3))
    noise_ang = np.zeros((C, T, nh))
    sign = np.zeros((C, nh))
    for ci, cid in enumerate(clip_ids):
        rng = np.random.Generator(np.random.PCG64(seed0 + int(cid)))
        heights[ci] = rng.uniform(1.55, 1.95)
        a = rng.uniform(0.0, 1.0, (nhinge, 3)) / 3.0
        f = rng.uniform(0.2, 1.5, (nhinge, 3))
        ph = rng.uniform(0.0, 2 * np.pi, (nhinge, 3))
        s = (a[:, :, None] * np.sin(2 * np.pi * f[:, :, None] * t[None, None, :] + ph[:, :, None])).sum(1)
        q = mid[:, None] + 0.35 * half[:, None] * s
        q = np.clip(q, (lo + 0.02)[:, None], (hi - 0.02)[:, None])
        qpos[ci, :, 7:] = q.T
        # root: xy walk (two sinusoids per axis, total amplitude <= 1 m) about a random start
        start = rng.uniform(-2.0, 2.0, 2)
        wa = rng.uniform(0.0, 0.5, (2, 2))
        wf = rng.uniform(0.05, 0.4, (2, 2))
        wp = rng.uniform(0.0, 2 * np.pi, (2, 2))
        xy = start[:, None] + (wa[:, :, None] * np.sin(2 * np.pi * wf[:, :, None] * t + wp[:, :, None])).sum(1)
        zf, zp = rng.uniform(0.2, 1.5), rng.uniform(0.0, 2 * np.pi)
        z = robot.qpos0[2] + 0.05 * np.sin(2 * np.pi * zf * t + zp)
        yaw0, ya, yf, yp = rng.uniform(-np.pi, np.pi), rng.uniform(0.0, 1.0), rng.uniform(0.05, 0.3), rng.uniform(0.0, 2 * np.pi)
        yaw = yaw0 + ya * np.sin(2 * np.pi * yf * t + yp)
        rp_f = rng.uniform(0.2, 1.0, 2)
        rp_p = rng.uniform(0.0, 2 * np.pi, 2)
        roll = 0.15 * np.sin(2 * np.pi * rp_f[0] * t + rp_p[0])
        pitch = 0.15 * np.sin(2 * np.pi * rp_f[1] * t + rp_p[1])
        ez = np.array([0.0, 0.0, 1.0]); ey = np.array([0.0, 1.0, 0.0]); ex = np.array([1.0, 0.0, 0.0])
        qr = _qmul(_qmul(_axis_angle(np.broadcast_to(ez, (T, 3)), yaw),
                         _axis_angle(np.broadcast_to(ey, (T, 3)), pitch)),
                   _axis_angle(np.broadcast_to(ex, (T, 3)), roll))
        qr = _qmul(qr, np.broadcast_to(robot.qpos0[3:7], (T, 4)))
        qpos[ci, :, 0:2] = xy.T
        qpos[ci, :, 2] = z
        qpos[ci, :, 3:7] = qr
        noise_p[ci] = rng.normal(0.0, pos_sigma, (T, nh, 3))
        ax = rng.normal(0.0, 1.0, (T, nh, 3))
        noise_axis[ci] = ax / np.linalg.norm(ax, axis=-1, keepdims=True)
        noise_ang[ci] = rng.normal(0.0, rot_sigma, (T, nh))
        sign[ci] = 1.0 - 2.0 * rng.integers(0, 2, nh)
    if src_human == "bvh":
        heights[:] = 1.75              # reference utils/lafan1.py:39
    if fixed_height is not None:
        heights[:] = fixed_height

    # robot frame bound to each human body by table 1 (unique per human body)
    frame_of_human = -np.ones(nh, np.int64)
    for k in range(table.nt):
        if table.in1[k] or not table.use1:
            frame_of_human[table.task_human[k]] = table.task_body[k]
    if np.any(frame_of_human < 0):
        raise ValueError("every human body needs a table-1 robot frame to synthesise targets")

    pos = np.zeros((C, T, nh, 3), np.float32)
    quat = np.zeros((C, T, nh, 4), np.float32)
    chunk = max(1, 65536 // max(T, 1))
    for c0 in range(0, C, chunk):
        c1 = min(C, c0 + chunk)
        n = (c1 - c0) * T
        xpos, xquat = batched_fk(robot, qpos[c0:c1].reshape(n, -1))
        pf = xpos[:, frame_of_human]          # [n,nh,3]
        Rf = xquat[:, frame_of_human]         # [n,nh,4]
        # inverse of offset_human_data: q_h = R_f * rot_off^-1 ; p' = p_f - R_f(pos_off)
        qh = _qmul(Rf, np.broadcast_to(_qconj(table.rot_off), Rf.shape))
        pp = pf - _qrot(Rf, np.broadcast_to(table.pos_off, pf.shape))
        # inverse of scale_human_data with s = scale * height / height_assumption
        ratio = np.repeat(heights[c0:c1], T) / table.height_assumption          # [n]
        s = table.scale[None, :] * ratio[:, None]                                # [n,nh]
        root = pp[:, table.root_idx] / s[:, table.root_idx, None]                # [n,3]
        ph_ = (pp - pp[:, table.root_idx][:, None, :]) / s[:, :, None] + root[:, None, :]
        ph_[:, table.root_idx] = root
        # noise
        ph_ = ph_ + noise_p[c0:c1].reshape(n, nh, 3)
        qn = _axis_angle(noise_axis[c0:c1].reshape(n, nh, 3), noise_ang[c0:c1].reshape(n, nh))
        qh = _qmul(qh, qn)
        qh = qh / np.linalg.norm(qh, axis=-1, keepdims=True)
        qh = qh * np.repeat(sign[c0:c1], T, axis=0)[:, :, None]
        pos[c0:c1] = ph_.reshape(c1 - c0, T, nh, 3).astype(np.float32)
        quat[c0:c1] = qh.reshape(c1 - c0, T, nh, 4).astype(np.float32)
    return ClipBatch(pos=pos, quat=quat, heights=heights.astype(np.float32), qpos_gen=qpos)
