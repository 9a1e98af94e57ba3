"""ctypes binding of include/gmr_b200.h.

`GmrModelDesc` mirrors the C struct field for field; `build_desc` fills it from a compiled
(RobotModel, TaskTable).  `load_library()` loads csrc/libgmr_b200.so and FAILS LOUDLY when
it is missing — there is no CPU fallback in the product path.
"""
from __future__ import annotations

import ctypes as C
import os
import pathlib
from typing import List, Tuple

import numpy as np

from .ik_config import TaskTable
from .mjcf import RobotModel

HERE = pathlib.Path(__file__).parent
LIB_PATH = HERE / "csrc" / "libgmr_b200.so"

GMR_MAX_BODY, GMR_MAX_HINGE, GMR_MAX_HUMAN, GMR_MAX_TASK = 64, 32, 32, 32
GMR_FLAG_OFFSET_TO_GROUND = 1
GMR_STATUS_BAD_INPUT, GMR_STATUS_AS_CAP, GMR_STATUS_NONFINITE = 1, 2, 4
GMR_STATUS_FATAL = GMR_STATUS_BAD_INPUT | GMR_STATUS_NONFINITE
# mink.lie.utils.get_epsilon(float64) as recalled from upstream (not stated in the reference tree); round 1 assumed
# LIE_EPS_ROUND1.  A model parameter, see GmrModelDesc.lie_eps.
LIE_EPS_DEFAULT = 1e-10
LIE_EPS_ROUND1 = 2.220446049250313e-15

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)
_pb = C.POINTER(C.c_uint8)


class GmrModelDesc(C.Structure):
    _fields_ = [
        ("nbody", C.c_int32), ("nhinge", C.c_int32), ("nhuman", C.c_int32), ("ntask", C.c_int32),
        ("body_parent", _pi), ("body_pos", _pd), ("body_quat", _pd), ("body_hinge", _pi),
        ("hinge_axis", _pd), ("hinge_lo", _pd), ("hinge_hi", _pd), ("hinge_limited", _pb),
        ("qpos0", _pd),
        ("human_root", C.c_int32),
        ("human_scale", _pd), ("human_pos_off", _pd), ("human_rot_off", _pd), ("human_foot", _pb),
        ("task_body", _pi), ("task_human", _pi), ("task_w1", _pd), ("task_w2", _pd),
        ("task_in1", _pb), ("task_in2", _pb),
        ("use_stage1", C.c_int32), ("use_stage2", C.c_int32),
        ("damping", C.c_double), ("lm_damping", C.c_double), ("limit_gain", C.c_double),
        ("tol", C.c_double), ("timestep", C.c_double), ("max_iter", C.c_int32),
        ("lie_eps", C.c_double),
    ]


def build_desc(robot: RobotModel, table: TaskTable, damping: float = 0.5, lm_damping: float = 1.0,
               limit_gain: float = 0.95, tol: float = 1e-3, max_iter: int = 10,
               lie_eps: float = 0.0) -> Tuple[GmrModelDesc, List[np.ndarray]]:
    """Returns (desc, keepalive): the arrays in `keepalive` back the struct's pointers.
    `lie_eps`: mink's small-angle threshold (0 = LIE_EPS_DEFAULT, see include/gmr_b200.h)."""
    keep: List[np.ndarray] = []

    def arr(a, dtype):
        x = np.ascontiguousarray(np.asarray(a), dtype=dtype)
        keep.append(x)
        return x

    def pd(a):
        return arr(a, np.float64).ctypes.data_as(_pd)

    def pi(a):
        return arr(a, np.int32).ctypes.data_as(_pi)

    def pb(a):
        return arr(a, np.uint8).ctypes.data_as(_pb)

    d = GmrModelDesc()
    d.nbody, d.nhinge, d.nhuman, d.ntask = robot.nbody, robot.nhinge, table.nh, table.nt
    d.body_parent = pi(robot.parent)
    d.body_pos = pd(robot.body_pos)
    d.body_quat = pd(robot.body_quat)
    d.body_hinge = pi(robot.body_hinge)
    d.hinge_axis = pd(robot.hinge_axis)
    d.hinge_lo = pd(robot.hinge_lo)
    d.hinge_hi = pd(robot.hinge_hi)
    d.hinge_limited = pb(robot.hinge_limited)
    d.qpos0 = pd(robot.qpos0)
    d.human_root = table.root_idx
    d.human_scale = pd(table.scale)
    d.human_pos_off = pd(table.pos_off)
    d.human_rot_off = pd(table.rot_off)
    d.human_foot = pb(table.foot_mask)
    d.task_body = pi(table.task_body)
    d.task_human = pi(table.task_human)
    d.task_w1 = pd(table.w1)
    d.task_w2 = pd(table.w2)
    d.task_in1 = pb(table.in1)
    d.task_in2 = pb(table.in2)
    d.use_stage1, d.use_stage2 = int(table.use1), int(table.use2)
    d.damping, d.lm_damping, d.limit_gain = float(damping), float(lm_damping), float(limit_gain)
    d.tol, d.timestep, d.max_iter = float(tol), float(robot.timestep), int(max_iter)
    d.lie_eps = float(lie_eps) if lie_eps and lie_eps > 0 else LIE_EPS_DEFAULT
    return d, keep


class GmrBatchExtra(C.Structure):
    """Mirror of the C struct (device pointers as integers; 0 = NULL)."""
    _fields_ = [("lengths", C.c_void_p), ("local_body_pos", C.c_void_p), ("lowest_z", C.c_void_p),
                ("warm_state", C.c_void_p), ("status", C.c_void_p)]


class GmrBatchDesc(C.Structure):
    """Mirror of the C struct: one robot-uniform bucket of a mixed-robot launch (device pointers as integers)."""
    _fields_ = [("model", C.c_void_p), ("pos", C.c_void_p), ("quat", C.c_void_p), ("ratio", C.c_void_p),
                ("C", C.c_int32), ("T", C.c_int32), ("qpos_init", C.c_void_p), ("qpos_out", C.c_void_p),
                ("iters_out", C.c_void_p), ("err_out", C.c_void_p)]


class NativeLibraryMissing(RuntimeError):
    pass


_LIB = None


def load_library() -> C.CDLL:
    """Load libgmr_b200.so (built by __graft_entry__.build() / csrc/build.sh)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("GMR_B200_LIB", str(LIB_PATH))
    if not os.path.isfile(path):
        raise NativeLibraryMissing(
            f"{path} not found: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; "
            f"g.build()'` (or general_motion_retargeting_b200/csrc/build.sh). There is no CPU fallback.")
    lib = C.CDLL(path)
    vp, f32p, f64p, i32p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
    lib.gmr_model_create.argtypes = [C.POINTER(GmrModelDesc), C.c_int, C.POINTER(vp)]
    lib.gmr_model_create.restype = C.c_int
    lib.gmr_model_destroy.argtypes = [vp]
    lib.gmr_model_destroy.restype = C.c_int
    lib.gmr_retarget_batch.argtypes = [vp, f32p, f32p, f32p, C.c_int32, C.c_int32, f32p, f32p, i32p, f32p, f32p,
                                       C.c_uint32, vp]
    lib.gmr_retarget_batch.restype = C.c_int
    lib.gmr_retarget_batch_f64.argtypes = [vp, f32p, f32p, f32p, C.c_int32, C.c_int32, f64p, f64p, i32p, f64p, f64p,
                                           C.c_uint32, vp]
    lib.gmr_retarget_batch_f64.restype = C.c_int
    lib.gmr_retarget_batch_f64_ex.argtypes = [vp, f32p, f32p, f32p, C.c_int32, C.c_int32, f64p, f64p, i32p, f64p, f64p,
                                              C.POINTER(GmrBatchExtra), C.c_uint32, vp]
    lib.gmr_retarget_batch_f64_ex.restype = C.c_int
    lib.gmr_retarget_batch_host.argtypes = [vp, f32p, f32p, f32p, C.c_int32, C.c_int32, f32p, f32p, i32p, f32p,
                                            C.c_uint32]
    lib.gmr_retarget_batch_host.restype = C.c_int
    lib.gmr_retarget_batch_host_ex.argtypes = [vp, f32p, f32p, f32p, C.c_int32, C.c_int32, f32p, f32p, i32p, f32p, i32p,
                                               C.c_uint32]
    lib.gmr_retarget_batch_host_ex.restype = C.c_int
    lib.gmr_retarget_batch_ex.argtypes = [vp, f32p, f32p, f32p, C.c_int32, C.c_int32, f32p, f32p, i32p, f32p, f32p,
                                          C.POINTER(GmrBatchExtra), C.c_uint32, vp]
    lib.gmr_retarget_batch_ex.restype = C.c_int
    lib.gmr_finalize_motion.argtypes = [vp, f32p, f32p, i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        f32p, f32p, f32p, vp]
    lib.gmr_finalize_motion.restype = C.c_int
    lib.gmr_stream_create.argtypes = [vp, C.c_double, C.POINTER(vp)]
    lib.gmr_stream_create.restype = C.c_int
    lib.gmr_stream_destroy.argtypes = [vp]
    lib.gmr_stream_destroy.restype = C.c_int
    lib.gmr_stream_reset.argtypes = [vp, vp]
    lib.gmr_stream_reset.restype = C.c_int
    lib.gmr_stream_retarget.argtypes = [vp, vp, vp, C.c_uint32, vp, vp, vp, vp]
    lib.gmr_stream_retarget.restype = C.c_int
    lib.gmr_produce_bvh_frames.argtypes = [f32p, f32p, i32p, C.c_int32, C.c_int32, i32p, i32p, C.c_int32, f32p, f32p, vp]
    lib.gmr_produce_bvh_frames.restype = C.c_int
    lib.gmr_produce_smplx_frames.argtypes = [f32p, f32p, f32p, i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, i32p, C.c_int32,
                                             f32p, f32p, vp]
    lib.gmr_produce_smplx_frames.restype = C.c_int
    lib.gmr_retarget_multi.argtypes = [C.POINTER(GmrBatchDesc), C.c_int32, C.c_uint32, vp]
    lib.gmr_retarget_multi.restype = C.c_int
    lib.gmr_launch_count.argtypes = []
    lib.gmr_launch_count.restype = C.c_int64
    lib.gmr_last_error.argtypes = []
    lib.gmr_last_error.restype = C.c_char_p
    lib.gmr_kernel_info.argtypes = [vp, C.c_int32] + [C.POINTER(C.c_int32)] * 5
    lib.gmr_kernel_info.restype = C.c_int
    _LIB = lib
    return lib


EXPORTED_SYMBOLS = [
    "gmr_model_create", "gmr_model_destroy", "gmr_retarget_batch", "gmr_retarget_batch_f64",
    "gmr_retarget_batch_host", "gmr_retarget_batch_host_ex", "gmr_launch_count", "gmr_last_error", "gmr_kernel_info",
    "gmr_retarget_batch_ex", "gmr_retarget_batch_f64_ex", "gmr_finalize_motion",
    "gmr_stream_create", "gmr_stream_destroy", "gmr_stream_reset", "gmr_stream_retarget",
    "gmr_produce_bvh_frames", "gmr_produce_smplx_frames", "gmr_retarget_multi", "gmr_debug_trace",
]
