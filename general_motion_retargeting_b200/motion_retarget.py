"""`GeneralMotionRetargeting` — drop-in for the reference class of the same name
(reference general_motion_retargeting/motion_retarget.py:10-270) with the per-frame IK solve
running in the sm_100a CUDA library (csrc/libgmr_b200.so) instead of mink/MuJoCo/DAQP.

Kept from the reference: the constructor signature (:13-21), ``retarget(human_data,
offset_to_ground=False) -> qpos[nq]`` float64 ``[x y z qw qx qy qz hinge…]`` (:139-185),
``update_targets`` (:117-136), ``error1``/``error2`` (:188-200), the helper methods
``to_numpy``/``scale_human_data``/``offset_human_data``/``offset_human_data_to_ground``
(:203-270) and the public attributes callers read (``xml_file``, ``scaled_human_data``,
``ik_match_table1/2``, ``human_scale_table``, ``max_iter`` …).

Added: ``retarget_batch(pos, quat, heights)`` — C clips × T frames in one launch, clip ``c``
being exactly ``GMR(src, robot, heights[c])`` followed by ``for t: retarget(frame[c, t])``;
this is what scripts/smplx_to_robot_dataset.py:84-87 and scripts/bvh_to_robot_dataset.py:96-103
call instead of their per-frame loop.

There is no CPU fallback: constructing the class without the built CUDA library raises.
"""
from __future__ import annotations

import copy
import ctypes as C
import types
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _native
from .ik_config import IKConfig, TaskTable, compile_task_table
from .mjcf import RobotModel
from .params import IK_CONFIG_DICT, ROBOT_XML_DICT, load_pack

FLAG_OFFSET_TO_GROUND = 1
FLAG_NO_SOLVE = 2
FLAG_COMPUTE_F64 = 4


def _is_torch(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch"


class RetargetFailure(RuntimeError):
    """Raised by the batched entries when the kernel reports a fatal per-clip status: what the reference signals per
    clip with an exception out of `retarget()` (scipy / mink on non-finite targets, `assert dq is not None` in
    mink.solve_ik — motion_retarget.py:147-150), which scripts/smplx_to_robot_dataset.py:62-76,96-100 catches per
    file.  `clip_ids` / `frames` / `status` name the clips, `result` holds what the call would have returned (every
    other clip is valid; a failed clip's frames from `frames[i]` on are not written)."""

    def __init__(self, clip_ids, frames, status, result=None):
        self.clip_ids, self.frames, self.status, self.result = list(clip_ids), list(frames), status, result
        what = ", ".join(f"clip {c} at frame {f}" for c, f in list(zip(self.clip_ids, self.frames))[:8])
        more = "" if len(self.clip_ids) <= 8 else f" (+{len(self.clip_ids) - 8} more)"
        super().__init__(f"retargeting failed for {len(self.clip_ids)} clip(s): non-finite keypoints / zero-norm "
                         f"quaternion or no QP solution: {what}{more}")


_SOLVERS = ("daqp", "quadprog", "proxqp", "osqp", "cvxopt", "ecos", "scs", "clarabel", "highs", "piqp", "qpalm")


class _Handle:
    """One gmr_model_create() handle per (instance, device)."""

    def __init__(self, lib, desc, device: int):
        self.lib = lib
        self.ptr = C.c_void_p()
        rc = lib.gmr_model_create(C.byref(desc), device, C.byref(self.ptr))
        if rc != 0:
            raise RuntimeError(f"gmr_model_create failed ({rc}): {lib.gmr_last_error().decode()}")

    def __del__(self):
        try:
            if self.ptr:
                self.lib.gmr_model_destroy(self.ptr)
        except Exception:
            pass


class _Stream:
    """One gmr_stream_create() handle: the persistent configuration of a live retargeter."""

    def __init__(self, lib, handle: _Handle, ratio: float):
        self.lib, self.handle = lib, handle          # keeps the model handle alive
        self.ptr = C.c_void_p()
        rc = lib.gmr_stream_create(handle.ptr, float(ratio), C.byref(self.ptr))
        if rc != 0:
            raise RuntimeError(f"gmr_stream_create failed ({rc}): {lib.gmr_last_error().decode()}")

    def __del__(self):
        try:
            if self.ptr:
                self.lib.gmr_stream_destroy(self.ptr)
        except Exception:
            pass


class GeneralMotionRetargeting:
    """General Motion Retargeting (GMR), B200-native solve."""

    def __init__(
        self,
        src_human: str,
        tgt_robot: str,
        actual_human_height: Optional[float] = None,
        solver: str = "daqp",
        damping: float = 5e-1,
        verbose: bool = False,
        device: Union[int, str, None] = None,
        precision: str = "f64",
        lie_eps: Optional[float] = None,
    ) -> None:
        # the robot model and IK config (KeyError on unknown names, like the reference's dict lookups).  xml_file is the
        # MJCF path callers hand to KinematicsModel (smplx_to_robot_dataset.py:94); the model itself comes from the
        # compiled pack, which load_pack() has checked against that very file when a checkout is present.
        self.xml_file = str(ROBOT_XML_DICT[tgt_robot])
        self._robot, cfg, _pack = load_pack(src_human, tgt_robot)
        self.model_source = str(_pack.get("loaded_from", ""))
        self._cfg: IKConfig = cfg
        if verbose:
            print("Use robot model: ", self.xml_file)
            print("Use IK config: ", IK_CONFIG_DICT[src_human][tgt_robot])
            print("Loaded from: ", self.model_source, "(source sha256", str(_pack.get("source_sha256", "?"))[:12] + ")")
        if solver not in _SOLVERS:
            # the reference hands `solver` to qpsolvers, which raises for a name it does not know; every supported
            # solver returns the same unique minimiser of the strictly convex QP, computed here exactly on the GPU
            raise ValueError(f"unknown QP solver '{solver}' (qpsolvers names: {', '.join(_SOLVERS)})")
        if precision not in ("f32", "f64"):
            raise ValueError("precision must be 'f32' or 'f64'")
        self.precision = precision
        # mink's small-angle threshold of SO3/SE3 log and jlog (GmrModelDesc.lie_eps; None = 1e-10, the value recalled
        # for mink's float64 get_epsilon; round 1 assumed 10 eps = 2.2e-15 — DESIGN.md §5 records what moves between them)
        self.lie_eps = float(lie_eps) if lie_eps else 0.0
        self.src_human, self.tgt_robot = src_human, tgt_robot

        # height ratio (motion_retarget.py:36-43)
        if actual_human_height is not None:
            ratio = actual_human_height / cfg.human_height_assumption
        else:
            ratio = 1.0
        self._ratio = float(ratio)
        self.human_scale_table = {k: v * ratio for k, v in cfg.human_scale_table.items()}

        self.ik_match_table1 = copy.deepcopy(cfg.ik_match_table1)
        self.ik_match_table2 = copy.deepcopy(cfg.ik_match_table2)
        self.human_root_name = cfg.human_root_name
        self.robot_root_name = cfg.robot_root_name
        self.use_ik_match_table1 = cfg.use_ik_match_table1
        self.use_ik_match_table2 = cfg.use_ik_match_table2
        self.ground = cfg.ground_height * np.array([0, 0, 1])
        self.max_iter = 10
        self.solver = solver          # validated above; the QP's unique minimiser is computed exactly on the GPU
        self.damping = damping

        self._table: TaskTable = compile_task_table(self._robot, cfg)
        self.model = self._robot
        self.tasks1 = [{"frame_name": self._table.task_frames[k], "position_cost": float(self._table.w1[k, 0]),
                        "orientation_cost": float(self._table.w1[k, 1])}
                       for k in range(self._table.nt) if self._table.in1[k]]
        self.tasks2 = [{"frame_name": self._table.task_frames[k], "position_cost": float(self._table.w2[k, 0]),
                        "orientation_cost": float(self._table.w2[k, 1])}
                       for k in range(self._table.nt) if self._table.in2[k]]
        self.pos_offsets1 = {n: self._table.pos_off[i].copy() for i, n in enumerate(self._table.human_names)}
        self.rot_offsets1 = {n: self._table.rot_off[i].copy() for i, n in enumerate(self._table.human_names)}

        self._lib = _native.load_library()          # raises NativeLibraryMissing: no CPU fallback
        self._handles: Dict[int, _Handle] = {}
        self._stream: Optional[_Stream] = None
        self._stream_qpos: Optional[np.ndarray] = None
        self._device = self._resolve_device(device)

        self.setup_retarget_configuration()

    # ------------------------------------------------------------------ plumbing ------------
    @staticmethod
    def _resolve_device(device) -> int:
        if device is None:
            try:
                import torch
                return int(torch.cuda.current_device()) if torch.cuda.is_available() else 0
            except Exception:
                return 0
        if isinstance(device, int):
            return device
        s = str(device)
        return int(s.split(":")[1]) if ":" in s else 0

    def _handle(self, device: int) -> _Handle:
        h = self._handles.get(device)
        if h is None:
            desc, keep = _native.build_desc(self._robot, self._table, damping=self.damping, max_iter=self.max_iter, lie_eps=self.lie_eps)
            h = _Handle(self._lib, desc, device)
            del keep
            self._handles[device] = h
        return h

    def _invalidate_handles(self) -> None:
        self._stream = None
        self._handles = {}

    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self._lib.gmr_last_error().decode()}")

    # ------------------------------------------------------------------ reference API --------
    def setup_retarget_configuration(self):
        """mink.Configuration(model): qpos = qpos0 (motion_retarget.py:74-75)."""
        self._qpos = np.array(self._robot.qpos0, dtype=np.float64)
        self.configuration = types.SimpleNamespace(q=self._qpos, model=self._robot,
                                                   data=types.SimpleNamespace(qpos=self._qpos))
        self.scaled_human_data = None
        self._last_frame: Optional[Tuple[np.ndarray, np.ndarray]] = None
        self._last_ground = False
        self.last_iters = (0, 0)
        self.last_errors = (0.0, 0.0)

    def _pack_frame(self, human_data) -> Tuple[np.ndarray, np.ndarray]:
        names = self._table.human_names
        pos = np.empty((1, 1, len(names), 3), np.float32)
        quat = np.empty((1, 1, len(names), 4), np.float32)
        for i, n in enumerate(names):
            p, q = human_data[n]          # KeyError for a missing table body, as in the reference (:129/:241)
            pos[0, 0, i] = np.asarray(p, dtype=np.float64)
            quat[0, 0, i] = np.asarray(q, dtype=np.float64)
        return pos, quat

    def _run_single(self, pos: np.ndarray, quat: np.ndarray, flags: int):
        """One frame through the live-stream entry (gmr_stream_retarget): the configuration and the solver's
        working sets stay on the device between calls, one captured CUDA graph launch per frame (float64)."""
        nq, nh = self._robot.nq, self._table.nh
        if self._stream is None:
            self._stream = _Stream(self._lib, self._handle(self._device), self._ratio)
            self._stream_qpos = None
        if self._stream_qpos is None or not np.array_equal(self._stream_qpos, self._qpos):
            # fresh instance, setup_retarget_configuration(), or the caller edited configuration.q
            q0 = np.ascontiguousarray(self._qpos, np.float64)
            self._check(self._lib.gmr_stream_reset(self._stream.ptr, q0.ctypes.data), "gmr_stream_reset")
            self._stream_qpos = q0.copy()
        q = np.empty(nq, np.float64)
        it = np.zeros(2, np.int32)
        err = np.zeros(2, np.float64)
        tg = np.empty((nh, 7), np.float64)
        rc = self._lib.gmr_stream_retarget(self._stream.ptr, pos.ctypes.data, quat.ctypes.data, flags, q.ctypes.data,
                                           it.ctypes.data, err.ctypes.data, tg.ctypes.data)
        self._check(rc, "gmr_stream_retarget")
        if not (flags & FLAG_NO_SOLVE):
            self._stream_qpos = q.copy()
        return q, it, err, tg

    def _store_targets(self, tg: np.ndarray) -> None:
        self.scaled_human_data = {n: [tg[i, 0:3].copy(), tg[i, 3:7].copy()]
                                  for i, n in enumerate(self._table.human_names)}

    def update_targets(self, human_data, offset_to_ground=False):
        human_data = self.to_numpy(human_data)
        pos, quat = self._pack_frame(human_data)
        flags = FLAG_NO_SOLVE | (FLAG_OFFSET_TO_GROUND if offset_to_ground else 0)
        _, _, err, tg = self._run_single(pos, quat, flags)
        self._last_frame, self._last_ground = (pos, quat), bool(offset_to_ground)
        self.last_errors = (float(err[0]), float(err[1]))
        self._store_targets(tg)

    def retarget(self, human_data, offset_to_ground=False):
        human_data = self.to_numpy(human_data)
        pos, quat = self._pack_frame(human_data)
        flags = FLAG_OFFSET_TO_GROUND if offset_to_ground else 0
        q, it, err, tg = self._run_single(pos, quat, flags)
        self._last_frame, self._last_ground = (pos, quat), bool(offset_to_ground)
        self._qpos[:] = q
        self.last_iters = (int(it[0]), int(it[1]))
        self.last_errors = (float(err[0]), float(err[1]))
        self._store_targets(tg)
        return self._qpos.copy()

    def _errors_now(self) -> Tuple[float, float]:
        if self._last_frame is None:
            raise RuntimeError("target not set: call update_targets() or retarget() first")   # mink TargetNotSet
        flags = FLAG_NO_SOLVE | (FLAG_OFFSET_TO_GROUND if self._last_ground else 0)
        _, _, err, _ = self._run_single(self._last_frame[0], self._last_frame[1], flags)
        return float(err[0]), float(err[1])

    def error1(self):
        return self._errors_now()[0]

    def error2(self):
        return self._errors_now()[1]

    # helper methods of the reference, same semantics (host-side, numpy; the batched path does
    # the same arithmetic inside the kernel)
    def to_numpy(self, human_data):
        for body_name in human_data.keys():
            human_data[body_name] = [np.asarray(human_data[body_name][0]), np.asarray(human_data[body_name][1])]
        return human_data

    def scale_human_data(self, human_data, human_root_name, human_scale_table):
        root_pos, root_quat = human_data[human_root_name]
        scaled_root_pos = human_scale_table[human_root_name] * root_pos
        out = {human_root_name: (scaled_root_pos, root_quat)}
        for body_name in human_data.keys():
            if body_name not in human_scale_table or body_name == human_root_name:
                continue
            out[body_name] = ((human_data[body_name][0] - root_pos) * human_scale_table[body_name] + scaled_root_pos,
                              human_data[body_name][1])
        return out

    @staticmethod
    def _qmul(a, b):
        return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                         a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                         a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                         a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])

    @staticmethod
    def _qrot(q, v):
        u = q[1:4]
        t = 2.0 * np.cross(u, v)
        return v + q[0] * t + np.cross(u, t)

    def offset_human_data(self, human_data, pos_offsets, rot_offsets):
        out = {}
        for body_name in human_data.keys():
            pos, quat = human_data[body_name]
            q = np.asarray(quat, np.float64)
            q = q / np.linalg.norm(q)
            r = np.asarray(rot_offsets[body_name], np.float64)
            updated = self._qmul(q, r / np.linalg.norm(r))
            updated = updated / np.linalg.norm(updated)
            out[body_name] = [np.asarray(pos, np.float64) + self._qrot(updated, np.asarray(pos_offsets[body_name], np.float64)), updated]
        return out

    def offset_human_data_to_ground(self, human_data):
        lowest = np.inf
        for body_name in human_data.keys():
            if "Foot" not in body_name and "foot" not in body_name:
                continue
            if human_data[body_name][0][2] < lowest:
                lowest = human_data[body_name][0][2]
        return {b: [p - np.array([0, 0, lowest]) + np.array([0, 0, 0.1]), q] for b, (p, q) in human_data.items()}

    # ------------------------------------------------------------------ batched entry ---------
    @property
    def human_body_names(self) -> List[str]:
        """Order of the ``nh`` axis of the batched inputs (= ``human_scale_table`` keys)."""
        return list(self._table.human_names)

    def pack_clips(self, clips: Sequence[Sequence[dict]]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """list of reference-format clips (list of per-frame dicts) → (pos, quat, lengths), padded to
        the longest clip by repeating the last frame."""
        names = self._table.human_names
        lengths = np.array([len(c) for c in clips], np.int32)
        T = int(lengths.max()) if len(clips) else 0
        pos = np.zeros((len(clips), T, len(names), 3), np.float32)
        quat = np.zeros((len(clips), T, len(names), 4), np.float32)
        quat[..., 0] = 1.0
        for ci, clip in enumerate(clips):
            for t, frame in enumerate(clip):
                for i, n in enumerate(names):
                    p, q = frame[n]
                    pos[ci, t, i] = p
                    quat[ci, t, i] = q
            if 0 < len(clip) < T:
                pos[ci, len(clip):] = pos[ci, len(clip) - 1]
                quat[ci, len(clip):] = quat[ci, len(clip) - 1]
        return pos, quat, lengths

    def _check_batch_shapes(self, pos, quat, heights, qpos_init, lengths=None):
        """ValueError for anything the kernel would otherwise read out of bounds (the C ABI takes raw pointers)."""
        nq, nh = self._robot.nq, self._table.nh
        if len(pos.shape) != 4 or len(quat.shape) != 4 or tuple(pos.shape[2:]) != (nh, 3) or tuple(quat.shape[2:]) != (nh, 4) \
                or tuple(pos.shape[:2]) != tuple(quat.shape[:2]):
            raise ValueError(f"expected pos [C,T,{nh},3] and quat [C,T,{nh},4], got {tuple(pos.shape)} and {tuple(quat.shape)}")
        Cn = int(pos.shape[0])
        if heights is not None and tuple(np.shape(heights)) != (Cn,):
            raise ValueError(f"heights must have shape ({Cn},), got {tuple(np.shape(heights))}")
        if qpos_init is not None and tuple(np.shape(qpos_init)) != (Cn, nq):
            raise ValueError(f"qpos_init must have shape ({Cn}, {nq}), got {tuple(np.shape(qpos_init))}")
        if lengths is not None and tuple(np.shape(lengths)) != (Cn,):
            raise ValueError(f"lengths must have shape ({Cn},), got {tuple(np.shape(lengths))}")
        return Cn, int(pos.shape[1])

    @staticmethod
    def _raise_on_status(status: np.ndarray, result, on_error: str):
        """status [C] int32 (GMR_STATUS_* | frame << 8) -> RetargetFailure for the fatal bits."""
        bad = np.nonzero(status & _native.GMR_STATUS_FATAL)[0]
        if on_error == "raise" and bad.size:
            raise RetargetFailure(bad.tolist(), (status[bad] >> 8).tolist(), status, result)

    def retarget_batch(self, pos, quat=None, heights=None, qpos_init=None, offset_to_ground: bool = False,
                       return_info: bool = False, precision: Optional[str] = None, out=None, on_error: str = "raise"):
        """Retarget C clips of T frames.

        pos [C,T,nh,3] metres world Z-up, quat [C,T,nh,4] wxyz, bodies ordered as
        ``human_body_names``; ``heights`` [C] = actual_human_height per clip (None: the instance's
        own height ratio).  torch CUDA tensors are solved in place on their device and stream
        (no host copies) and a torch tensor is returned; numpy arrays go through the library's
        host-buffer pipeline and numpy is returned.  A list of reference-format clips is packed
        first.  `out` (host path): preallocated float32 [C,T,nq] array, e.g. pinned memory.
        `precision`: arithmetic of the kernel, "f64" (default: the reference's own precision) or
        "f32" (APPROXIMATE: ~1.3x faster, but outside the 1e-3 rad parity gate on ill-conditioned
        clips).
        `on_error`: what to do when a clip fails the way the reference raises (non-finite keypoints,
        zero-norm quaternion, no QP solution): "raise" (default) raises `RetargetFailure` naming the
        clips (the result for all other clips is in `.result`; this synchronises the stream),
        "status" appends the per-clip status words [C] int32 to the return value, "ignore" does
        neither and stays asynchronous on the torch path.
        Returns qpos [C,T,nq]: torch float64/float32 for CUDA inputs, numpy float32 for host
        inputs; with return_info also (iters [C,T,2] int32, err [C,T,2])."""
        if quat is None:
            pos, quat, _ = self.pack_clips(pos)
        precision = precision or self.precision
        if precision not in ("f32", "f64"):
            raise ValueError("precision must be 'f32' or 'f64'")
        if on_error not in ("raise", "status", "ignore"):
            raise ValueError("on_error must be 'raise', 'status' or 'ignore'")
        flags = FLAG_OFFSET_TO_GROUND if offset_to_ground else 0
        nq, nh = self._robot.nq, self._table.nh
        Cn, T = self._check_batch_shapes(pos, quat, heights, qpos_init)

        if _is_torch(pos):
            import torch
            if not pos.is_cuda:
                raise ValueError("torch inputs must be CUDA tensors (pass numpy arrays for host buffers)")
            dev = pos.device
            h = self._handle(dev.index if dev.index is not None else torch.cuda.current_device())
            f64 = precision == "f64"
            dt = torch.float64 if f64 else torch.float32
            with torch.cuda.device(dev):
                d_pos = pos.to(torch.float32).contiguous()
                d_quat = quat.to(dev, torch.float32).contiguous()
                if heights is None:
                    d_ratio = torch.full((Cn,), self._ratio, dtype=torch.float32, device=dev)
                else:
                    # float32(float64(height) / assumption): one definition for every entry point
                    d_ratio = (torch.as_tensor(heights, device=dev).to(torch.float64) / float(self._cfg.human_height_assumption)).to(torch.float32).contiguous()
                d_init = None if qpos_init is None else torch.as_tensor(qpos_init, device=dev).to(dt).contiguous()
                d_q = torch.empty((Cn, T, nq), dtype=dt, device=dev)
                d_it = torch.zeros((Cn, T, 2), dtype=torch.int32, device=dev) if return_info else None
                d_err = torch.zeros((Cn, T, 2), dtype=dt, device=dev) if return_info else None
                d_st = torch.zeros((Cn,), dtype=torch.int32, device=dev) if on_error != "ignore" else None
                ex = _native.GmrBatchExtra(None, None, None, None, None if d_st is None else d_st.data_ptr())
                fn = self._lib.gmr_retarget_batch_f64_ex if f64 else self._lib.gmr_retarget_batch_ex
                rc = fn(h.ptr, d_pos.data_ptr(), d_quat.data_ptr(), d_ratio.data_ptr(), Cn, T,
                        None if d_init is None else d_init.data_ptr(), d_q.data_ptr(),
                        None if d_it is None else d_it.data_ptr(), None if d_err is None else d_err.data_ptr(),
                        None, C.byref(ex), flags, torch.cuda.current_stream(dev).cuda_stream)
                self._check(rc, "gmr_retarget_batch")
                res = (d_q, d_it, d_err) if return_info else d_q
                if on_error == "raise":
                    if bool(d_st.any()):                                  # one 1-byte read-back; synchronises
                        self._raise_on_status(d_st.cpu().numpy(), res, on_error)
                elif on_error == "status":
                    res = (res + (d_st,)) if return_info else (d_q, d_st)
            return res

        # host buffers: float32 in / float32 out, arithmetic in the requested precision
        if precision == "f64":
            flags |= FLAG_COMPUTE_F64
        h = self._handle(self._device)
        a_pos = np.ascontiguousarray(pos, np.float32)
        a_quat = np.ascontiguousarray(quat, np.float32)
        if heights is None:
            a_ratio = np.full((Cn,), self._ratio, np.float32)
        else:
            a_ratio = np.ascontiguousarray(np.asarray(heights, np.float64) / float(self._cfg.human_height_assumption), np.float32)
        a_init = None if qpos_init is None else np.ascontiguousarray(qpos_init, np.float32)
        if out is not None:
            if out.dtype != np.float32 or tuple(out.shape) != (Cn, T, nq) or not out.flags.c_contiguous:
                raise ValueError(f"out must be a C-contiguous float32 array of shape {(Cn, T, nq)}")
            qpos = out
        else:
            qpos = np.empty((Cn, T, nq), np.float32)
        iters = np.zeros((Cn, T, 2), np.int32) if return_info else None
        err = np.zeros((Cn, T, 2), np.float32) if return_info else None
        status = np.zeros((Cn,), np.int32) if on_error != "ignore" else None
        rc = self._lib.gmr_retarget_batch_host_ex(
            h.ptr, a_pos.ctypes.data, a_quat.ctypes.data, a_ratio.ctypes.data, Cn, T,
            None if a_init is None else a_init.ctypes.data, qpos.ctypes.data,
            None if iters is None else iters.ctypes.data, None if err is None else err.ctypes.data,
            None if status is None else status.ctypes.data, flags)
        self._check(rc, "gmr_retarget_batch_host")
        res = (qpos, iters, err) if return_info else qpos
        if on_error == "raise":
            self._raise_on_status(status, res, on_error)
        elif on_error == "status":
            res = (res + (status,)) if return_info else (qpos, status)
        return res


    # ------------------------------------------------------------------ dataset entry ---------
    def retarget_dataset(self, pos, quat=None, heights=None, lengths=None, height_adjust: bool = True,
                         root_origin_offset: bool = True, precision: Optional[str] = None, fps: float = 30.0,
                         as_numpy: bool = True, on_error: str = "raise"):
        """What `process_file` of scripts/smplx_to_robot_dataset.py:78-146 (and bvh_to_robot_dataset.py:96-157)
        computes for a whole batch of files in two launches: the per-frame IK (as `retarget_batch`) with the
        post-solve forward kinematics fused in (`local_body_pos`, the clip-wide lowest body height), then one
        elementwise pass that splits qpos into `root_pos` (height-adjusted, re-origined to the first frame),
        `root_rot` (xyzw) and `dof_pos`.  `height_adjust`/`root_origin_offset` are the scripts' HEIGHT_ADJUST and
        ROOT_ORIGIN_OFFSET switches (True/True in the SMPL-X script, False/False in the BVH script).
        `lengths` [C]: frames per clip for ragged batches (padding frames are not solved).
        `on_error`: "raise" raises `RetargetFailure` for clips the reference would raise on; "skip" mirrors the
        scripts' per-file try/except (smplx_to_robot_dataset.py:62-76,96-100): a failed clip's entry is None.
        Returns a list of C motion dicts in the reference's pkl layout
        (fps, root_pos [T,3], root_rot [T,4] xyzw, dof_pos [T,ndof], local_body_pos [T,nbody,3], link_body_list)."""
        import torch
        if quat is None:
            pos, quat, lengths = self.pack_clips(pos)
        precision = precision or self.precision
        nq, nh, nb = self._robot.nq, self._table.nh, self._robot.nbody
        dev = pos.device if _is_torch(pos) and pos.is_cuda else torch.device("cuda", self._device)
        h = self._handle(dev.index if dev.index is not None else torch.cuda.current_device())
        with torch.cuda.device(dev):
            d_pos = torch.as_tensor(pos).to(dev, torch.float32).contiguous()
            d_quat = torch.as_tensor(quat).to(dev, torch.float32).contiguous()
            Cn, T = self._check_batch_shapes(d_pos, d_quat, heights, None, lengths)
            if on_error not in ("raise", "skip"):
                raise ValueError("on_error must be 'raise' or 'skip'")
            if heights is None:
                d_ratio = torch.full((Cn,), self._ratio, dtype=torch.float32, device=dev)
            else:
                d_ratio = (torch.as_tensor(heights, device=dev).to(torch.float64) / float(self._cfg.human_height_assumption)).to(torch.float32).contiguous()
            d_len = None if lengths is None else torch.as_tensor(np.asarray(lengths), device=dev).to(torch.int32).contiguous()
            d_q = torch.zeros((Cn, T, nq), dtype=torch.float32, device=dev)
            d_lbp = torch.zeros((Cn, T, nb, 3), dtype=torch.float32, device=dev)
            d_low = torch.zeros((Cn,), dtype=torch.float32, device=dev)
            d_st = torch.zeros((Cn,), dtype=torch.int32, device=dev)
            ex = _native.GmrBatchExtra(None if d_len is None else d_len.data_ptr(), d_lbp.data_ptr(), d_low.data_ptr(), None,
                                       d_st.data_ptr())
            st = torch.cuda.current_stream(dev).cuda_stream
            flags = FLAG_COMPUTE_F64 if precision == "f64" else 0
            rc = self._lib.gmr_retarget_batch_ex(h.ptr, d_pos.data_ptr(), d_quat.data_ptr(), d_ratio.data_ptr(), Cn, T, None,
                                                 d_q.data_ptr(), None, None, None, C.byref(ex), flags, st)
            self._check(rc, "gmr_retarget_batch_ex")
            d_rp = torch.empty((Cn, T, 3), dtype=torch.float32, device=dev)
            d_rr = torch.empty((Cn, T, 4), dtype=torch.float32, device=dev)
            d_dof = torch.empty((Cn, T, nq - 7), dtype=torch.float32, device=dev)
            rc = self._lib.gmr_finalize_motion(h.ptr, d_q.data_ptr(), d_low.data_ptr(), None if d_len is None else d_len.data_ptr(),
                                               Cn, T, int(height_adjust), int(root_origin_offset),
                                               d_rp.data_ptr(), d_rr.data_ptr(), d_dof.data_ptr(), st)
            self._check(rc, "gmr_finalize_motion")
            status = d_st.cpu().numpy()
            failed = (status & _native.GMR_STATUS_FATAL) != 0
            if failed.any() and on_error == "raise":
                self._raise_on_status(status, None, "raise")
            if not as_numpy:
                return {"status": d_st, "fps": fps, "root_pos": d_rp, "root_rot": d_rr, "dof_pos": d_dof, "local_body_pos": d_lbp,
                        "lowest_z": d_low, "qpos": d_q, "link_body_list": list(self._robot.body_names), "lengths": d_len}
            rp, rr, dof, lbp = (x.cpu().numpy() for x in (d_rp, d_rr, d_dof, d_lbp))
        n = [T] * Cn if lengths is None else [int(x) for x in np.asarray(lengths)]
        return [None if failed[c] else
                {"fps": fps, "root_pos": rp[c, :n[c]], "root_rot": rr[c, :n[c]], "dof_pos": dof[c, :n[c]],
                 "local_body_pos": lbp[c, :n[c]], "link_body_list": list(self._robot.body_names)} for c in range(Cn)]

    @staticmethod
    def save_motion_pkls(motions: Sequence[dict], paths: Sequence[str], workers: int = 8) -> None:
        """pickle.dump of each motion dict to its path (scripts/smplx_to_robot_dataset.py:143-146), on a
        small thread pool: the files are independent and the work is I/O."""
        import os
        import pickle
        from concurrent.futures import ThreadPoolExecutor

        def one(args):
            m, p = args
            if m is None:          # a clip skipped by retarget_dataset(on_error="skip")
                return
            d = os.path.dirname(p)
            if d:
                os.makedirs(d, exist_ok=True)
            with open(p, "wb") as f:
                pickle.dump(m, f)

        with ThreadPoolExecutor(max_workers=max(1, workers)) as ex:
            list(ex.map(one, zip(motions, paths)))


def retarget_mixed(buckets, precision: str = "f64", return_info: bool = False, device=None, out=None):
    """Mixed-robot batches (BASELINE.json configs[4]) in ONE call: `buckets` is a list of
    ``(retargeter, pos [C,T,nh,3], quat [C,T,nh,4], heights [C] or None)``, one entry per robot (group the clips with
    `sharding.bucket_by_robot`).  The arrays are torch CUDA tensors on one device, or torch CPU tensors in PINNED memory
    (``.pin_memory()``): those are used in place - the kernels read the keypoints over the host link and write qpos
    into pinned output tensors - and the call returns when the outputs are complete.  The device's SMs are divided
    among the buckets by work, so a bucket's slow clips overlap the other buckets' bulk instead of every bucket paying
    its own tail.  Each clip is solved exactly as by `retargeter.retarget_batch`.  `out`: optional list of preallocated
    float32 [C,T,nq] tensors, one per bucket, on the inputs' side (pinned for host buckets) - pinning gigabytes per call
    costs more than the solve.  Returns the list of qpos tensors ([C,T,nq] float32, on the inputs' side), with
    `return_info` also the per-frame iteration counts."""
    import torch
    if precision not in ("f32", "f64"):
        raise ValueError("precision must be 'f32' or 'f64'")
    if not buckets:
        return []
    lib = buckets[0][0]._lib
    p0 = buckets[0][1]
    if not _is_torch(p0):
        raise ValueError("retarget_mixed takes torch tensors (CUDA, or CPU in pinned memory)")
    host = not p0.is_cuda
    if host:
        dev = torch.device("cuda", GeneralMotionRetargeting._resolve_device(device if device is not None else buckets[0][0]._device))
    else:
        dev = p0.device
    descs = (_native.GmrBatchDesc * len(buckets))()
    keep, outs, iters = [], [], []
    with torch.cuda.device(dev):
        for k, (g, pos, quat, heights) in enumerate(buckets):
            if not (_is_torch(pos) and _is_torch(quat)):
                raise ValueError("retarget_mixed takes torch tensors")
            if host:
                if pos.is_cuda or quat.is_cuda or not (pos.is_pinned() and quat.is_pinned()):
                    raise ValueError(f"bucket {k}: host buckets must be CPU tensors in pinned memory (.pin_memory())")
            elif not (pos.is_cuda and pos.device == dev and quat.is_cuda and quat.device == dev):
                raise ValueError("retarget_mixed takes all buckets on one device")
            nq, nh = g._robot.nq, g._table.nh
            Cn, T = int(pos.shape[0]), int(pos.shape[1])
            if tuple(pos.shape[2:]) != (nh, 3) or tuple(quat.shape) != (Cn, T, nh, 4):
                raise ValueError(f"bucket {k}: expected pos [C,T,{nh},3] and quat [C,T,{nh},4]")
            if pos.dtype != torch.float32 or quat.dtype != torch.float32 or not pos.is_contiguous() or not quat.is_contiguous():
                if host:
                    raise ValueError(f"bucket {k}: pinned host buckets must be contiguous float32")
                pos = pos.to(torch.float32).contiguous(); quat = quat.to(torch.float32).contiguous()
            if heights is None:
                d_ratio = torch.full((Cn,), g._ratio, dtype=torch.float32, device=dev)
            else:
                d_ratio = (torch.as_tensor(heights).to(dev).to(torch.float64) / float(g._cfg.human_height_assumption)).to(torch.float32).contiguous()
            if out is not None:
                d_q = out[k]
                ok = _is_torch(d_q) and d_q.dtype == torch.float32 and tuple(d_q.shape) == (Cn, T, nq) and d_q.is_contiguous() and \
                    ((not d_q.is_cuda and d_q.is_pinned()) if host else (d_q.is_cuda and d_q.device == dev))
                if not ok:
                    raise ValueError(f"out[{k}] must be a contiguous float32 tensor of shape {(Cn, T, nq)} on the inputs' side")
            elif host:
                d_q = torch.empty((Cn, T, nq), dtype=torch.float32).pin_memory()
            else:
                d_q = torch.empty((Cn, T, nq), dtype=torch.float32, device=dev)
            if host:
                d_it = torch.zeros((Cn, T, 2), dtype=torch.int32).pin_memory() if return_info else None
            else:
                d_it = torch.zeros((Cn, T, 2), dtype=torch.int32, device=dev) if return_info else None
            h = g._handle(dev.index if dev.index is not None else torch.cuda.current_device())
            keep += [pos, quat, d_ratio, h]
            # pinned host memory is addressable by the device at the same address (unified virtual addressing)
            descs[k] = _native.GmrBatchDesc(h.ptr.value, pos.data_ptr(), quat.data_ptr(), d_ratio.data_ptr(), Cn, T, None,
                                            d_q.data_ptr(), None if d_it is None else d_it.data_ptr(), None)
            outs.append(d_q); iters.append(d_it)
        flags = FLAG_COMPUTE_F64 if precision == "f64" else 0
        stream = torch.cuda.current_stream(dev)
        rc = lib.gmr_retarget_multi(descs, len(buckets), flags, stream.cuda_stream)
        if rc != 0:
            raise RuntimeError(f"gmr_retarget_multi failed ({rc}): {lib.gmr_last_error().decode()}")
        if host:
            stream.synchronize()
    return (outs, iters) if return_info else outs
