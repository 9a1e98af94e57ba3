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

Clip `i` draws its parameters from `numpy.random.Generator(PCG64(seed0 + i))` in the order
written below (part of the contract: tests/golden was generated with it).  The arithmetic
after the draws (trajectories, FK, inverse mapping) runs in float64 torch on `device`
("cpu" by default; bench.py uses the GPU so that 4096 x 300 frames take a second).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from .ik_config import TaskTable
from .mjcf import RobotModel

SEED0 = 20260000
FPS = 30.0


# ---- batched quaternion helpers (wxyz, tensors [...,4]) ------------------------------------
def _qmul(a, b):
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ], dim=-1)


def _qconj(q):
    return q * torch.tensor([1.0, -1.0, -1.0, -1.0], dtype=q.dtype, device=q.device)


def _qrot(q, v):
    w = q[..., 0:1]
    u = q[..., 1:4]
    t = 2.0 * torch.cross(u, v.expand_as(u), dim=-1)
    return v + w * t + torch.cross(u, t, dim=-1)


def _axis_angle(axis, angle):
    half = 0.5 * angle
    return torch.cat([torch.cos(half)[..., None], axis * torch.sin(half)[..., None]], dim=-1)


def batched_fk(robot: RobotModel, qpos: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """World pose of every body for qpos[N,nq] → (xpos[N,nb,3], xquat[N,nb,4])."""
    dev, dt = qpos.device, qpos.dtype
    bpos = torch.as_tensor(robot.body_pos, dtype=dt, device=dev)
    bquat = torch.as_tensor(robot.body_quat, dtype=dt, device=dev)
    axis = torch.as_tensor(robot.hinge_axis, dtype=dt, device=dev)
    xpos = [None] * robot.nbody
    xquat = [None] * robot.nbody
    for b in range(robot.nbody):
        p = int(robot.parent[b])
        if p < 0:
            xpos[b] = qpos[:, 0:3]
            q = qpos[:, 3:7]
            xquat[b] = q / q.norm(dim=-1, keepdim=True)
            continue
        xpos[b] = xpos[p] + _qrot(xquat[p], bpos[b])
        q = _qmul(xquat[p], bquat[b].expand_as(xquat[p]))
        j = int(robot.body_hinge[b])
        if j >= 0:
            q = _qmul(q, _axis_angle(axis[j].expand(qpos.shape[0], 3), qpos[:, 7 + j]))
        xquat[b] = q / q.norm(dim=-1, keepdim=True)
    return torch.stack(xpos, dim=1), torch.stack(xquat, dim=1)


@dataclass
class ClipBatch:
    pos: np.ndarray       # [C,T,nh,3] float32
    quat: np.ndarray      # [C,T,nh,4] float32 wxyz
    heights: np.ndarray   # [C] float32 actual_human_height
    qpos_gen: np.ndarray  # [C,T,nq] float64 trajectory the targets were generated from

    def ratio(self, table: TaskTable) -> np.ndarray:
        """Per-clip height ratio exactly as the batched entry computes it (float32)."""
        return (self.heights.astype(np.float64) / float(table.height_assumption)).astype(np.float32)


def make_clips(robot: RobotModel, table: TaskTable, clip_ids: Sequence[int], T: int = 300,
               src_human: str = "smplx", stress: bool = False, seed0: int = SEED0,
               fixed_height: Optional[float] = None, device: str = "cpu") -> ClipBatch:
    clip_ids = list(clip_ids)
    C = len(clip_ids)
    nhinge, nh = robot.nhinge, table.nh
    pos_sigma, rot_sigma = (0.05, 0.3) if stress else (0.005, 0.02)

    # ---- per-clip random draws (host, numpy PCG64) ------------------------------------------
    heights = np.zeros(C)
    ha = np.zeros((C, nhinge, 3)); hf = np.zeros((C, nhinge, 3)); hp = np.zeros((C, nhinge, 3))
    start = np.zeros((C, 2)); wa = np.zeros((C, 2, 2)); wf = np.zeros((C, 2, 2)); wp = np.zeros((C, 2, 2))
    zpar = np.zeros((C, 2)); ypar = np.zeros((C, 4)); rpf = np.zeros((C, 2)); rpp = np.zeros((C, 2))
    noise_p = np.zeros((C, T, nh, 3), np.float32)
    noise_axis = np.zeros((C, T, nh, 3), np.float32)
    noise_ang = np.zeros((C, T, nh), np.float32)
    sign = np.zeros((C, nh))
    for ci, cid in enumerate(clip_ids):
        rng = np.random.Generator(np.random.PCG64(seed0 + int(cid)))
        heights[ci] = rng.uniform(1.55, 1.95)
        ha[ci] = rng.uniform(0.0, 1.0, (nhinge, 3)) / 3.0
        hf[ci] = rng.uniform(0.2, 1.5, (nhinge, 3))
        hp[ci] = rng.uniform(0.0, 2 * np.pi, (nhinge, 3))
        start[ci] = rng.uniform(-2.0, 2.0, 2)
        wa[ci] = rng.uniform(0.0, 0.5, (2, 2))
        wf[ci] = rng.uniform(0.05, 0.4, (2, 2))
        wp[ci] = rng.uniform(0.0, 2 * np.pi, (2, 2))
        zpar[ci] = [rng.uniform(0.2, 1.5), rng.uniform(0.0, 2 * np.pi)]
        ypar[ci] = [rng.uniform(-np.pi, np.pi), rng.uniform(0.0, 1.0), rng.uniform(0.05, 0.3), rng.uniform(0.0, 2 * np.pi)]
        rpf[ci] = rng.uniform(0.2, 1.0, 2)
        rpp[ci] = rng.uniform(0.0, 2 * np.pi, 2)
        noise_p[ci] = rng.normal(0.0, pos_sigma, (T, nh, 3))
        ax = rng.normal(0.0, 1.0, (T, nh, 3))
        noise_axis[ci] = ax / np.linalg.norm(ax, axis=-1, keepdims=True)
        noise_ang[ci] = rng.normal(0.0, rot_sigma, (T, nh))
        sign[ci] = 1.0 - 2.0 * rng.integers(0, 2, nh)
    if src_human == "bvh":
        heights[:] = 1.75              # reference utils/lafan1.py:39
    if fixed_height is not None:
        heights[:] = fixed_height
    heights32 = heights.astype(np.float32)

    # robot frame bound to each human body by table 1 (unique per human body)
    frame_of_human = -np.ones(nh, np.int64)
    for k in range(table.nt):
        if table.in1[k] or not table.use1:
            frame_of_human[table.task_human[k]] = table.task_body[k]
    if np.any(frame_of_human < 0):
        raise ValueError("every human body needs a table-1 robot frame to synthesise targets")

    # ---- trajectories, FK and the inverse target mapping (float64 torch) ----------------------
    dev = torch.device(device)
    f64 = torch.float64
    tt = lambda a: torch.as_tensor(np.asarray(a), dtype=f64, device=dev)  # noqa: E731
    t = torch.arange(T, dtype=f64, device=dev) / FPS
    lo = np.where(robot.hinge_limited, robot.hinge_lo, -np.pi)
    hi = np.where(robot.hinge_limited, robot.hinge_hi, np.pi)
    mid, half = tt(0.5 * (lo + hi)), tt(0.5 * (hi - lo))
    two_pi = 2 * np.pi
    pos = np.zeros((C, T, nh, 3), np.float32)
    quat = np.zeros((C, T, nh, 4), np.float32)
    qpos_all = np.zeros((C, T, robot.nq))
    rot_off_c = _qconj(tt(table.rot_off))
    pos_off = tt(table.pos_off)
    scale = tt(table.scale)
    q0 = tt(robot.qpos0)
    fidx = torch.as_tensor(frame_of_human, device=dev)
    chunk = max(1, 131072 // max(T, 1))
    for c0 in range(0, C, chunk):
        c1 = min(C, c0 + chunk)
        n = c1 - c0
        s = (tt(ha[c0:c1])[..., None] * torch.sin(two_pi * tt(hf[c0:c1])[..., None] * t + tt(hp[c0:c1])[..., None])).sum(2)  # [n,nhinge,T]
        qh_ = mid[None, :, None] + 0.35 * half[None, :, None] * s
        qh_ = torch.minimum(torch.maximum(qh_, tt(lo + 0.02)[None, :, None]), tt(hi - 0.02)[None, :, None])
        qpos = torch.zeros((n, T, robot.nq), dtype=f64, device=dev)
        qpos[:, :, 7:] = qh_.transpose(1, 2)
        xy = tt(start[c0:c1])[..., None] + (tt(wa[c0:c1])[..., None] * torch.sin(two_pi * tt(wf[c0:c1])[..., None] * t + tt(wp[c0:c1])[..., None])).sum(2)
        qpos[:, :, 0:2] = xy.transpose(1, 2)
        zp = tt(zpar[c0:c1]); yp = tt(ypar[c0:c1]); rf = tt(rpf[c0:c1]); rp = tt(rpp[c0:c1])
        qpos[:, :, 2] = q0[2] + 0.05 * torch.sin(two_pi * zp[:, 0:1] * t + zp[:, 1:2])
        yaw = yp[:, 0:1] + yp[:, 1:2] * torch.sin(two_pi * yp[:, 2:3] * t + yp[:, 3:4])
        roll = 0.15 * torch.sin(two_pi * rf[:, 0:1] * t + rp[:, 0:1])
        pitch = 0.15 * torch.sin(two_pi * rf[:, 1:2] * t + rp[:, 1:2])
        ex = torch.tensor([1.0, 0.0, 0.0], dtype=f64, device=dev).expand(n, T, 3)
        ey = torch.tensor([0.0, 1.0, 0.0], dtype=f64, device=dev).expand(n, T, 3)
        ez = torch.tensor([0.0, 0.0, 1.0], dtype=f64, device=dev).expand(n, T, 3)
        qr = _qmul(_qmul(_axis_angle(ez, yaw), _axis_angle(ey, pitch)), _axis_angle(ex, roll))
        qpos[:, :, 3:7] = _qmul(qr, q0[3:7].expand(n, T, 4))
        xpos, xquat = batched_fk(robot, qpos.reshape(n * T, -1))
        pf = xpos[:, fidx]                    # [N,nh,3]
        Rf = xquat[:, fidx]                   # [N,nh,4]
        # inverse of offset_human_data: q_h = R_f * rot_off^-1 ; p' = p_f - R_f(pos_off)
        qh = _qmul(Rf, rot_off_c.expand_as(Rf))
        pp = pf - _qrot(Rf, pos_off.expand_as(pf))
        # inverse of scale_human_data with s = scale * ratio (ratio as the float32 the solver receives)
        ratio = (heights32[c0:c1].astype(np.float64) / float(table.height_assumption)).astype(np.float32).astype(np.float64)
        sc = scale[None, :] * tt(np.repeat(ratio, T))[:, None]                    # [N,nh]
        rootp = pp[:, table.root_idx] / sc[:, table.root_idx, None]               # [N,3]
        ph_ = (pp - pp[:, table.root_idx][:, None, :]) / sc[:, :, None] + rootp[:, None, :]
        ph_[:, table.root_idx] = rootp
        ph_ = ph_ + tt(noise_p[c0:c1]).reshape(n * T, nh, 3)
        qn = _axis_angle(tt(noise_axis[c0:c1]).reshape(n * T, nh, 3), tt(noise_ang[c0:c1]).reshape(n * T, nh))
        qh = _qmul(qh, qn)
        qh = qh / qh.norm(dim=-1, keepdim=True)
        qh = qh * tt(np.repeat(sign[c0:c1], T, axis=0))[:, :, None]
        pos[c0:c1] = ph_.reshape(n, T, nh, 3).to(torch.float32).cpu().numpy()
        quat[c0:c1] = qh.reshape(n, T, nh, 4).to(torch.float32).cpu().numpy()
        qpos_all[c0:c1] = qpos.cpu().numpy()
    return ClipBatch(pos=pos, quat=quat, heights=heights32, qpos_gen=qpos_all)
