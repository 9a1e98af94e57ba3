import ctypes as C
import os

import numpy as np

from general_motion_retargeting_b200 import params
from general_motion_retargeting_b200._native import build_desc
from general_motion_retargeting_b200.ik_config import compile_task_table
from general_motion_retargeting_b200.synthetic import make_clips

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu", "libgmr_emu.so")
_emu = None


def problem(src, robot):
    m, cfg, pack = params.load_pack(src, robot)
    return m, compile_task_table(m, cfg), pack


def emu_retarget_batch(robot, table, pos, quat, ratio, bits=64, flags=0, qpos_init=None, max_iter=10, nthreads=0, lie_eps=0.0):
    """Lane-serial host build of the kernel body (tests/emu) — a debugging aid, not a product path."""
    global _emu
    if _emu is None:
        _emu = C.CDLL(EMU)
        _emu.gmr_emu_retarget_batch.restype = C.c_int
    Cn, T = pos.shape[:2]
    desc, keep = build_desc(robot, table, max_iter=max_iter, lie_eps=lie_eps)
    qpos = np.zeros((Cn, T, robot.nq)); iters = np.zeros((Cn, T, 2), np.int32); err = np.zeros((Cn, T, 2))
    tg = np.zeros((Cn, T, table.nh, 7)); refac = C.c_int64(0)
    pos = np.ascontiguousarray(pos, np.float32); quat = np.ascontiguousarray(quat, np.float32)
    ratio = None if ratio is None else np.ascontiguousarray(ratio, np.float32)
    qi = None if qpos_init is None else np.ascontiguousarray(qpos_init, np.float64)
    rc = _emu.gmr_emu_retarget_batch(
        C.byref(desc), C.c_void_p(pos.ctypes.data), C.c_void_p(quat.ctypes.data),
        None if ratio is None else C.c_void_p(ratio.ctypes.data), Cn, T,
        None if qi is None else C.c_void_p(qi.ctypes.data), C.c_void_p(qpos.ctypes.data),
        C.c_void_p(iters.ctypes.data), C.c_void_p(err.ctypes.data), C.c_void_p(tg.ctypes.data), flags, nthreads, bits,
        C.byref(refac))
    assert rc == 0, rc
    return qpos, iters, err, tg, refac.value


def compare(q, it, q_ref, it_ref):
    """(agreement rate of the per-frame iteration counts, max|dq| over all frames, max|dq| over
    frames with an identical iteration history = frames of a clip up to and including the last
    frame before its first count mismatch).  A different count is a different number of IK
    steps; its effect persists through the warm start for as long as the clip takes to
    re-converge, so frames after a mismatch are not 'identical inputs and iteration counts'."""
    same = (it == it_ref).all(-1)
    dq = np.abs(q - q_ref).max(-1)
    clean = np.logical_and.accumulate(same, axis=1)
    return float(same.mean()), float(dq.max()), float(dq[clean].max()) if clean.any() else 0.0


def emu_retarget_batch_ex(robot, table, pos, quat, ratio, lengths=None, bits=64, flags=0, qpos_init=None, warm_state=None,
                          status=None):
    """Emulator with the GmrBatchExtra members: returns (qpos, iters, local_body_pos, lowest_z, warm_state);
    `status` (int32 [C], zeroed by the caller) receives the per-clip GMR_STATUS_* words."""
    from general_motion_retargeting_b200._native import GmrBatchExtra
    global _emu
    if _emu is None:
        _emu = C.CDLL(EMU)
        _emu.gmr_emu_retarget_batch.restype = C.c_int
    Cn, T = pos.shape[:2]
    desc, keep = build_desc(robot, table)
    qpos = np.zeros((Cn, T, robot.nq)); iters = np.zeros((Cn, T, 2), np.int32)
    lbp = np.zeros((Cn, T, robot.nbody, 3), np.float32); low = np.zeros(Cn, np.float32)
    warm = np.zeros((Cn, 4), np.uint32) if warm_state is None else np.ascontiguousarray(warm_state, np.uint32).copy()
    pos = np.ascontiguousarray(pos, np.float32); quat = np.ascontiguousarray(quat, np.float32)
    ratio = np.ascontiguousarray(ratio, np.float32)
    ln = None if lengths is None else np.ascontiguousarray(lengths, np.int32)
    qi = None if qpos_init is None else np.ascontiguousarray(qpos_init, np.float64)
    ex = GmrBatchExtra(None if ln is None else ln.ctypes.data, lbp.ctypes.data, low.ctypes.data, warm.ctypes.data,
                       None if status is None else status.ctypes.data)
    rc = _emu.gmr_emu_retarget_batch_ex(
        C.byref(desc), C.c_void_p(pos.ctypes.data), C.c_void_p(quat.ctypes.data), C.c_void_p(ratio.ctypes.data), Cn, T,
        None if qi is None else C.c_void_p(qi.ctypes.data), C.c_void_p(qpos.ctypes.data), C.c_void_p(iters.ctypes.data),
        None, None, C.byref(ex), flags, 0, bits)
    assert rc == 0, rc
    return qpos, iters, lbp, low, warm


def oracle_body_positions(robot, pack, qpos):
    """(local_body_pos [T,nb,3] = FK with an identity root, world body_pos [T,nb,3]) from the float64 oracle FK."""
    from oracle.gmr_oracle import OracleRetargeter
    o = OracleRetargeter(robot, pack["ik_config"])
    T = qpos.shape[0]
    local = np.zeros((T, robot.nbody, 3)); world = np.zeros((T, robot.nbody, 3))
    for t in range(T):
        o.qpos[:] = qpos[t]; o._fk(); world[t] = o.xpos
        o.qpos[0:3] = 0.0; o.qpos[3:7] = [1.0, 0.0, 0.0, 0.0]; o._fk(); local[t] = o.xpos
    return local, world
