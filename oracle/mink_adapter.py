"""Opportunistic REAL-reference oracle (test infrastructure — NOT a product path).

When `mink`, `mujoco` (and the QP backend the reference selects, `daqp` through `qpsolvers`)
are importable together with the reference package `general_motion_retargeting` — e.g. after
the driver dropped an install under baseline/_ref/ — this module runs the UNMODIFIED reference
loop on batched synthetic clips:

    retargeter = GeneralMotionRetargeting(src_human, tgt_robot, actual_human_height=h)    # motion_retarget.py:13-21
    for frame in clip: qpos = retargeter.retarget(frame)                                   # motion_retarget.py:139-185

one clip per worker process, as scripts/smplx_to_robot_dataset.py:241-242 does with
`mp.Pool.starmap(process_file, ...)`.  Nothing of the reference is re-implemented here: the
adapter only converts the packed arrays of this repository ([C,T,nh,3] / [C,T,nh,4], bodies in
`human_scale_table` order) into the reference's per-frame dicts and collects what comes back.
Iteration counts are observed, not computed: `mink.solve_ik` is wrapped by a counter inside the
worker (the reference code itself is untouched), so count = IK steps per stage − 1, the
`num_iter` of motion_retarget.py:152-161,171-183.

None of these modules exist in the build container or the GPU image, so
`available()` is normally False and the C++/NumPy restatement (oracle/gmr_oracle.*) stays the
checker — "parity unpinned for A6–A12".  With them present, tests/test_mink_adapter.py pins the
restatement AND the CUDA kernel against the real thing, and `bench.py --impl reference` times
it (`kind: "reference"`).
"""
from __future__ import annotations

import importlib
import multiprocessing as mp
import os
import pathlib
import sys
import time
from typing import List, Optional, Sequence, Tuple

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
_SEARCH = [ROOT / "baseline" / "_ref", ROOT / "oracle" / "_ref"]
_status: Optional[Tuple[bool, str]] = None


def _extend_path() -> None:
    """baseline/_ref (pip --target install of the reference and its dependencies) first, then a
    reference checkout named by GMR_REFERENCE_ROOT or at /root/reference (package only)."""
    cands = [p for p in _SEARCH if p.is_dir()]
    env = os.environ.get("GMR_REFERENCE_ROOT")
    for p in ([pathlib.Path(env)] if env else []) + [pathlib.Path("/root/reference")]:
        if (p / "general_motion_retargeting" / "motion_retarget.py").is_file():
            cands.append(p)
    for p in cands:
        s = str(p)
        if s not in sys.path:
            sys.path.append(s)


def available() -> Tuple[bool, str]:
    """(True, versions) when the unmodified reference can run in this interpreter, else (False, why)."""
    global _status
    if _status is not None:
        return _status
    _extend_path()
    missing = []
    vers = []
    for name in ("mujoco", "mink", "qpsolvers", "daqp"):
        try:
            m = importlib.import_module(name)
            vers.append(f"{name} {getattr(m, '__version__', '?')}")
        except Exception as e:                       # ImportError, or a binary wheel that does not load here
            missing.append(f"{name} ({type(e).__name__})")
    if missing:
        _status = (False, "not importable: " + ", ".join(missing))
        return _status
    try:
        ref = importlib.import_module("general_motion_retargeting")
        if not hasattr(ref, "GeneralMotionRetargeting"):
            raise ImportError("package has no GeneralMotionRetargeting")
        where = os.path.dirname(getattr(ref, "__file__", "?"))
    except Exception as e:
        _status = (False, f"reference package general_motion_retargeting not importable: {type(e).__name__}: {e}")
        return _status
    _status = (True, ", ".join(vers) + f"; reference package at {where}")
    return _status


# ---- worker ---------------------------------------------------------------------------------
_W = {}


def _worker_init(src: str, robot: str, names: Sequence[str], offset_to_ground: bool) -> None:
    _extend_path()
    ref = importlib.import_module("general_motion_retargeting")
    mr = importlib.import_module("general_motion_retargeting.motion_retarget")
    _W.update(cls=ref.GeneralMotionRetargeting, mr=mr, src=src, robot=robot, names=list(names), otg=bool(offset_to_ground))


def _one_clip(args):
    """The reference, unmodified: a fresh retargeter per clip, frames in order (warm start)."""
    pos, quat, height = args                        # [T,nh,3] f64, [T,nh,4] f64, float
    cls, mr, names = _W["cls"], _W["mr"], _W["names"]
    g = cls(_W["src"], _W["robot"], actual_human_height=float(height))
    # observe the IK steps per stage: solve_ik(configuration, tasks, dt, solver, damping) — motion_retarget.py:147,156,166,176
    real = mr.mink.solve_ik
    calls = [0, 0]

    def counted(configuration, tasks, *a, **k):
        calls[0 if tasks is g.tasks1 else 1] += 1
        return real(configuration, tasks, *a, **k)

    class _Mink:                                     # module proxy: only solve_ik is intercepted
        def __getattr__(self, n):
            return counted if n == "solve_ik" else getattr(_W["mink"], n)

    _W["mink"] = mr.mink
    T = pos.shape[0]
    nq = int(g.model.nq)
    qpos = np.zeros((T, nq))
    iters = np.zeros((T, 2), np.int32)
    err = np.zeros((T, 2))
    mr.mink = _Mink()
    try:
        for t in range(T):
            calls[0] = calls[1] = 0
            frame = {n: (pos[t, i].copy(), quat[t, i].copy()) for i, n in enumerate(names)}
            qpos[t] = g.retarget(frame, offset_to_ground=_W["otg"])
            iters[t] = [max(calls[0] - 1, 0), max(calls[1] - 1, 0)]
            err[t] = [g.error1() if g.use_ik_match_table1 else 0.0, g.error2() if g.use_ik_match_table2 else 0.0]
    finally:
        mr.mink = _W["mink"]
    return qpos, iters, err


def retarget_batch(src: str, robot: str, names: Sequence[str], pos: np.ndarray, quat: np.ndarray, heights: np.ndarray,
                   offset_to_ground: bool = False, processes: int = 0):
    """pos [C,T,nh,3], quat [C,T,nh,4] (wxyz), heights [C] → (qpos [C,T,nq] f64, iters [C,T,2] i32, err [C,T,2] f64, seconds).
    `names`: the nh body names in array order.  Raises RuntimeError when the reference cannot run."""
    ok, why = available()
    if not ok:
        raise RuntimeError("the unmodified reference is unavailable: " + why)
    C = pos.shape[0]
    work = [(np.asarray(pos[c], np.float64), np.asarray(quat[c], np.float64), float(heights[c])) for c in range(C)]
    nproc = processes or min(os.cpu_count() or 1, max(C, 1))
    t0 = time.perf_counter()
    if nproc <= 1 or C <= 1:
        _worker_init(src, robot, names, offset_to_ground)
        res = [_one_clip(w) for w in work]
    else:
        ctx = mp.get_context("spawn")                # MuJoCo / torch state must not be forked
        with ctx.Pool(nproc, initializer=_worker_init, initargs=(src, robot, list(names), offset_to_ground)) as pool:
            res = pool.map(_one_clip, work, chunksize=1)
    dt = time.perf_counter() - t0
    qpos = np.stack([r[0] for r in res]) if res else np.zeros((0, 0, 0))
    iters = np.stack([r[1] for r in res]) if res else np.zeros((0, 0, 2), np.int32)
    err = np.stack([r[2] for r in res]) if res else np.zeros((0, 0, 2))
    return qpos, iters, err, dt
