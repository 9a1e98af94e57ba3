"""ctypes loader for oracle/liboracle.so (CPU checker; test infrastructure only).

Reuses the product's GmrModelDesc ctypes mirror (same C struct, include/gmr_b200.h); the
arithmetic lives entirely in oracle/gmr_oracle.cpp.
"""
from __future__ import annotations

import ctypes as C
import os
import pathlib
import subprocess

import numpy as np

from general_motion_retargeting_b200._native import GmrModelDesc, build_desc

HERE = pathlib.Path(__file__).parent
LIB = HERE / "liboracle.so"
_lib = None
FLAG_OFFSET_TO_GROUND = 1
FLAG_STABLE_LIE = 0x10000     # oracle-only: cancellation-free SO3/SE3 Jacobian coefficients (gmr_oracle.py STABLE_LIE)


def build(force: bool = False) -> None:
    if force or not LIB.is_file() or LIB.stat().st_mtime < (HERE / "gmr_oracle.cpp").stat().st_mtime:
        subprocess.check_call(["make", "-C", str(HERE), "-B" if force else "-s"])


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB.is_file():
            build()
        _lib = C.CDLL(str(LIB))
        _lib.gmr_oracle_retarget_batch.restype = C.c_int
        _lib.gmr_oracle_retarget_batch.argtypes = [
            C.POINTER(GmrModelDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32, C.c_int32]
    return _lib


def retarget_batch(robot, table, pos, quat, ratio=None, qpos_init=None, flags=0, nthreads=0, precision_bits=64,
                   damping=0.5, lie_eps=0.0):
    """pos[C,T,nh,3], quat[C,T,nh,4] float32; ratio[C] float32 or None → (qpos[C,T,nq] f64, iters[C,T,2], err[C,T,2])."""
    lib = load()
    pos = np.ascontiguousarray(pos, np.float32)
    quat = np.ascontiguousarray(quat, np.float32)
    Cn, T = pos.shape[0], pos.shape[1]
    assert pos.shape == (Cn, T, table.nh, 3) and quat.shape == (Cn, T, table.nh, 4)
    desc, keep = build_desc(robot, table, damping=damping, lie_eps=lie_eps)
    r = None if ratio is None else np.ascontiguousarray(ratio, np.float32)
    qi = None if qpos_init is None else np.ascontiguousarray(qpos_init, np.float64)
    qpos = np.zeros((Cn, T, robot.nq), np.float64)
    iters = np.zeros((Cn, T, 2), np.int32)
    err = np.zeros((Cn, T, 2), np.float64)
    rc = lib.gmr_oracle_retarget_batch(
        C.byref(desc), pos.ctypes.data, quat.ctypes.data, None if r is None else r.ctypes.data, Cn, T,
        None if qi is None else qi.ctypes.data, qpos.ctypes.data, iters.ctypes.data, err.ctypes.data,
        flags, nthreads, precision_bits)
    if rc != 0:
        raise RuntimeError(f"gmr_oracle_retarget_batch failed: {rc}")
    del keep
    return qpos, iters, err
