"""CPU restatement (NumPy, float64) of the reference's human-frame producers — TEST INFRASTRUCTURE ONLY: only
tests/ and bench tooling import it; the product path (general_motion_retargeting_b200/producers.py) runs the CUDA
kernels.  Pinned against the reference's own code through tests/golden/reference_producers.npz
(tools/make_golden_producers.py).

 * bvh_frames: utils/lafan1.py:17-35 — `quat_fk` (lafan_vendor/utils.py:88-103), the Y-up -> Z-up rotation
   [[1,0,0],[0,0,-1],[0,1,0]], centimetres -> metres, the synthesised LeftFootMod / RightFootMod bodies
   (position of the foot, orientation of the toe).
 * smplx_frames: utils/smpl.py:109-196 — optional resampling to 30 fps (rotations: the file's own `slerp`
   :77-104 between from_rotvec'd neighbours, positions: linear), then the global joint orientation chain
   rot_i = rot_parent(i) * from_rotvec(pose_i) and (joint position, quaternion wxyz) per joint.
"""
import numpy as np


def qmul(a, b):
    aw, ax, ay, az = np.moveaxis(a, -1, 0); bw, bx, by, bz = np.moveaxis(b, -1, 0)
    return np.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw], -1)


def qrot(q, v):
    u = q[..., 1:]
    t = 2.0 * np.cross(u, v)
    return v + q[..., :1] * t + np.cross(u, t)


def bvh_frames(lrot, lpos, parents, pos_joint, rot_joint):
    """lrot [F,J,4] wxyz local, lpos [F,J,3] cm local -> (pos [F,nh,3] m Z-up, quat [F,nh,4] wxyz) for the bodies
    whose position comes from joint pos_joint[b] and orientation from joint rot_joint[b]."""
    lrot = np.asarray(lrot, np.float64); lpos = np.asarray(lpos, np.float64)
    J = lrot.shape[1]
    gr = [None] * J; gp = [None] * J
    gr[0], gp[0] = lrot[:, 0], lpos[:, 0]
    for i in range(1, J):                                   # quat_fk, lafan_vendor/utils.py:97-100
        p = parents[i]
        gp[i] = qrot(gr[p], lpos[:, i]) + gp[p]
        gr[i] = qmul(gr[p], lrot[:, i])
    gr = np.stack(gr, 1); gp = np.stack(gp, 1)
    Rm = np.array([[1.0, 0, 0], [0, 0, -1.0], [0, 1.0, 0]])
    rq = np.array([np.sqrt(0.5), np.sqrt(0.5), 0.0, 0.0])  # R.from_matrix(Rm).as_quat(scalar_first=True)
    pos = gp[:, pos_joint] @ Rm.T / 100.0                   # lafan1.py:28
    quat = qmul(np.broadcast_to(rq, gr[:, rot_joint].shape), gr[:, rot_joint])   # lafan1.py:27
    return pos, quat


def from_rotvec(r):
    """scipy Rotation.from_rotvec -> quaternion wxyz (small-angle series below 1e-3 rad, as scipy)."""
    r = np.asarray(r, np.float64)
    ang = np.linalg.norm(r, axis=-1, keepdims=True)
    small = ang <= 1e-3
    a2 = ang * ang
    scale = np.where(small, 0.5 - a2 / 48 + a2 * a2 / 3840, np.sin(ang / 2) / np.where(small, 1.0, ang))
    return np.concatenate([np.cos(ang / 2), scale * r], -1)


def slerp(q1, q2, t):
    """utils/smpl.py:77-104 on wxyz quaternions (arrays [...,4], t [...,1])."""
    q1 = q1 / np.linalg.norm(q1, axis=-1, keepdims=True); q2 = q2 / np.linalg.norm(q2, axis=-1, keepdims=True)
    dot = np.sum(q1 * q2, -1, keepdims=True)
    q2 = np.where(dot < 0, -q2, q2); dot = np.abs(dot)
    lin = q1 + t * (q2 - q1)
    th0 = np.arccos(np.clip(dot, -1, 1)); th = th0 * t
    s0 = np.cos(th) - dot * np.sin(th) / np.where(np.sin(th0) == 0, 1.0, np.sin(th0))
    s1 = np.sin(th) / np.where(np.sin(th0) == 0, 1.0, np.sin(th0))
    q = np.where(dot > 0.9995, lin, s0 * q1 + s1 * q2)
    return q / np.linalg.norm(q, axis=-1, keepdims=True)   # R.from_quat normalises


def smplx_frames(global_orient, full_pose, joints, parents, src_fps, tgt_fps, body_joint):
    """global_orient [F,3], full_pose [F,NJ,3] rotvecs, joints [F,>=NJ,3] -> (pos [F',nh,3], quat [F',nh,4] wxyz,
    aligned_fps) for the joints body_joint[b]."""
    go = np.asarray(global_orient, np.float64); fp = np.asarray(full_pose, np.float64).copy(); jt = np.asarray(joints, np.float64)
    F, NJ = fp.shape[0], fp.shape[1]
    fp[:, 0] = go
    q = from_rotvec(fp)                                                    # [F,NJ,4]
    if tgt_fps < src_fps:
        skip = int(src_fps / tgt_fps)
        Fo = F // skip
        tt = np.linspace(0, F - 1, Fo)
        i1 = np.floor(tt).astype(int); i2 = np.minimum(i1 + 1, F - 1); al = (tt - i1)[:, None, None]
        q = slerp(q[i1], q[i2], al)
        w = q[..., :1]
        q = np.where(w < 0, -q, q)                                          # as_rotvec()/from_rotvec() round trip: canonical sign
        jt = jt[i1] + al * (jt[i2] - jt[i1])                               # interp1d(kind="linear")
        aligned = Fo / F * src_fps
    else:
        aligned = tgt_fps
    g = [None] * NJ
    g[0] = q[:, 0]
    for i in range(1, NJ):
        g[i] = qmul(g[parents[i]], q[:, i])
    g = np.stack(g, 1)
    return jt[:, body_joint], g[:, body_joint], aligned
