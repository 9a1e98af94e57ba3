"""CPU ORACLE (test infrastructure — NOT a product path).

Float64 NumPy restatement of the reference's per-frame two-stage IK,
`GeneralMotionRetargeting.retarget` (reference
general_motion_retargeting/motion_retarget.py:139-185), including the third-party
arithmetic it calls.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this file; the product (general_motion_retargeting_b200)
never does and fails loudly without its CUDA library.

PARITY STATUS: **parity unpinned for A6-A12.**  mink, mujoco, daqp and qpsolvers are not
vendored in /root/reference and not installed in this image, and the
reference ships no tests or golden vectors.  What IS pinned against code of
the reference run in this container (tests/golden/, tools/make_golden.py):
  * A2/A3/A4 target preprocessing against the reference's own
    `scale_human_data` / `offset_human_data` / `offset_human_data_to_ground`
    (motion_retarget.py:209-270) executed unmodified with mink/mujoco stubbed out;
  * A6 forward kinematics against the reference's `KinematicsModel.forward_kinematics`
    (kinematics_model.py:213-246) for the 7 robots it can parse.
The remaining steps restate the published algorithms of the un-vendored dependencies
(all unpinned in requirements.txt:7-9 / setup.py:15-19):
  * mink (kevinzakka/mink; `Configuration`, `FrameTask`, `Task.compute_qp_objective`,
    `ConfigurationLimit`, `solve_ik`, `lie.SO3/SE3`),
  * MuJoCo 3.x (`mj_kinematics`, `mj_jacBody`, `mj_integratePos`, `mju_quatIntegrate`),
  * DAQP through qpsolvers: the exact minimiser of a strictly convex box-constrained QP,
    which is unique, so any exact active-set method reproduces it to solver tolerance.
and are pinned only by self-consistency tests (Jacobian vs finite differences of the error
under the integration convention, log/jlog vs scipy Rotation + numerical differentiation,
QP vs brute-force KKT enumeration and scipy BVLS).

Row labels A0..A13 refer to SURVEY.md §8(a).
"""
from __future__ import annotations

import copy
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

MJ_MINVAL = 1e-15          # mjMINVAL
LIMIT_GAIN = 0.95          # mink ConfigurationLimit default gain
LM_DAMPING = 1.0           # motion_retarget.py:88,106
TASK_GAIN = 1.0            # mink FrameTask default gain


# --------------------------------------------------------------------------------------
# quaternion / rotation helpers (wxyz), following MuJoCo's engine_util_spatial.c
# --------------------------------------------------------------------------------------
def quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """mju_mulQuat."""
    return np.array([
        a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
        a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
        a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
        a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0],
    ])


def quat_conj(q: np.ndarray) -> np.ndarray:
    return np.array([q[0], -q[1], -q[2], -q[3]])


def quat_normalize(q: np.ndarray) -> np.ndarray:
    """mju_normalize4: a (near-)zero quaternion becomes identity."""
    n = math.sqrt(float(q @ q))
    if n < MJ_MINVAL:
        return np.array([1.0, 0.0, 0.0, 0.0])
    return q / n


def quat_to_mat(q: np.ndarray) -> np.ndarray:
    """mju_quat2Mat."""
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def quat_rotate(q: np.ndarray, v: np.ndarray) -> np.ndarray:
    """mju_rotVecQuat."""
    return quat_to_mat(q) @ v


def axis_angle_to_quat(axis: np.ndarray, angle: float) -> np.ndarray:
    """mju_axisAngle2Quat."""
    if angle == 0.0:
        return np.array([1.0, 0.0, 0.0, 0.0])
    s = math.sin(0.5 * angle)
    return np.array([math.cos(0.5 * angle), axis[0] * s, axis[1] * s, axis[2] * s])


def skew(v: np.ndarray) -> np.ndarray:
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


# --------------------------------------------------------------------------------------
# Lie-group pieces, following mink/lie/so3.py and mink/lie/se3.py
# --------------------------------------------------------------------------------------
# mink.lie.utils.get_epsilon(float64): threshold of every small-angle branch below.  mink is not vendored and
# /root/reference never states the value; upstream (jaxlie-derived lie/utils.py) is recalled as 1e-10 for float64,
# round 1 assumed 10 * machine epsilon = 2.2e-15.  It is a module-level PARAMETER so that the parity matrix can be
# run with both (tests/test_lie_eps.py); the C++ port and the CUDA kernel take the same value through
# GmrModelDesc.lie_eps.
LIE_EPS_DEFAULT = 1e-10
LIE_EPS_ROUND1 = float(np.finfo(np.float64).eps) * 10.0
_EPS64 = LIE_EPS_DEFAULT


def set_lie_eps(eps: float) -> None:
    global _EPS64
    _EPS64 = float(eps)


def so3_log(q: np.ndarray) -> np.ndarray:
    """SO3.log of a unit quaternion (wxyz) → rotation vector with |ω| <= π (A7)."""
    w = q[0]
    norm_sq = float(q[1:] @ q[1:])
    use_taylor = norm_sq < _EPS64
    norm_safe = 1.0 if use_taylor else math.sqrt(norm_sq)
    w_safe = w if use_taylor else 1.0
    atan_n_over_w = math.atan2(-norm_safe if w < 0 else norm_safe, abs(w))
    if use_taylor:
        atan_factor = 2.0 / w_safe - 2.0 / 3.0 * norm_sq / w_safe ** 3
    elif abs(w) < _EPS64:
        atan_factor = (1.0 if w > 0.0 else -1.0) * math.pi / norm_safe
    else:
        atan_factor = 2.0 * atan_n_over_w / norm_safe
    return atan_factor * q[1:]


# mink evaluates J_l^{-1} and Barfoot's Q with closed forms that cancel catastrophically for
# tiny angles (1 - cos θ, θ - sin θ …): for θ ~ 1e-6 rad the coefficient A below carries an
# absolute error ~1e-16/θ^4 and the Jacobian entries ~1e-4.  LITERAL = the formulas as mink
# writes them (the faithful restatement, default).  STABLE_LIE = True switches to
# mathematically identical, well-conditioned forms (half-angle cotangent, power series below
# 0.25 rad); tests use it to show which part of a GPU-vs-oracle difference is the reference's
# own rounding noise.
STABLE_LIE = False


def so3_ljacinv(omega: np.ndarray) -> np.ndarray:
    """SO3 inverse left Jacobian J_l^{-1}(ω) (A9)."""
    theta = math.sqrt(float(omega @ omega))
    if theta < _EPS64:
        t2 = theta * theta
        A = (1.0 / 12.0) * (1.0 + t2 / 60.0 * (1.0 + t2 / 42.0 * (1.0 + t2 / 40.0)))
    elif STABLE_LIE:
        half = 0.5 * theta
        A = (1.0 - half * math.cos(half) / math.sin(half)) / (theta * theta)
    else:
        A = (1.0 / theta ** 2) * (1.0 - (theta * math.sin(theta) / (2.0 * (1.0 - math.cos(theta)))))
    S = skew(omega)
    return np.eye(3) - 0.5 * S + A * (S @ S)


def se3_log(q: np.ndarray, t: np.ndarray) -> np.ndarray:
    """SE3.log → [ρ(3); ω(3)] (translation part first) (A7)."""
    omega = so3_log(q)
    theta_sq = float(omega @ omega)
    S = skew(omega)
    S2 = S @ S
    if theta_sq < _EPS64:
        V_inv = np.eye(3) - 0.5 * S + S2 / 12.0
    else:
        theta = math.sqrt(theta_sq)
        half = 0.5 * theta
        V_inv = np.eye(3) - 0.5 * S + ((1.0 - 0.5 * theta * math.cos(half) / math.sin(half)) / theta_sq) * S2
    return np.concatenate([V_inv @ t, omega])


def _se3_getQ(c: np.ndarray) -> np.ndarray:
    """Barfoot's Q(ρ, ω) block of the SE(3) left Jacobian (mink lie/se3.py `_getQ`)."""
    theta_sq = float(c[3:] @ c[3:])
    A = 0.5
    if theta_sq < _EPS64:
        B = 1.0 / 6.0 + theta_sq / 120.0
        C = -1.0 / 24.0 + theta_sq / 720.0
        D = 1.0 / 120.0          # limit of the closed form below (Barfoot's 4th coefficient)
    elif STABLE_LIE and theta_sq < 0.0625:
        t2 = theta_sq
        B = 1.0 / 6.0 - t2 * (1.0 / 120.0 - t2 * (1.0 / 5040.0 - t2 / 362880.0))
        C = -1.0 / 24.0 + t2 * (1.0 / 720.0 - t2 * (1.0 / 40320.0 - t2 / 3628800.0))
        D = 1.0 / 120.0 - t2 * (1.0 / 2520.0 - t2 * (1.0 / 120960.0 - t2 / 9979200.0))
    else:
        theta = math.sqrt(theta_sq)
        s, co = math.sin(theta), math.cos(theta)
        B = (theta - s) / (theta_sq * theta)
        C = (1.0 - theta_sq / 2.0 - co) / (theta_sq * theta_sq)
        D = (2.0 * theta - 3.0 * s + theta * co) / (2.0 * theta_sq * theta_sq * theta)
    V = skew(c[:3])
    W = skew(c[3:])
    VW = V @ W
    WV = VW.T
    WVW = WV @ W
    VWW = VW @ W
    return A * V + B * (WV + VW + WVW) - C * (VWW - VWW.T - 3.0 * WVW) + D * (WVW @ W + W @ WVW)


def se3_ljacinv(xi: np.ndarray) -> np.ndarray:
    theta = xi[3:]
    if float(theta @ theta) < _EPS64:
        return np.eye(6)
    Q = _se3_getQ(xi)
    Jinv = so3_ljacinv(theta)
    out = np.zeros((6, 6))
    out[:3, :3] = Jinv
    out[:3, 3:] = -Jinv @ Q @ Jinv
    out[3:, 3:] = Jinv
    return out


def se3_jlog(q: np.ndarray, t: np.ndarray) -> np.ndarray:
    """SE3.jlog = J_r^{-1}(log T) = J_l^{-1}(-log T)."""
    return se3_ljacinv(-se3_log(q, t))


# --------------------------------------------------------------------------------------
# exact box-constrained QP (stands in for qpsolvers → DAQP, A11)
# --------------------------------------------------------------------------------------
def solve_box_qp(H: np.ndarray, c: np.ndarray, lo: np.ndarray, hi: np.ndarray,
                 max_iter: Optional[int] = None) -> Tuple[np.ndarray, int]:
    """argmin ½xᵀHx + cᵀx s.t. lo <= x <= hi (±inf = unbounded), H SPD.

    Primal active-set method (Nocedal & Wright alg. 16.3 specialised to bounds) started
    from clip(0, lo, hi).  Returns (x, number of active bounds at the solution)."""
    n = c.shape[0]
    x = np.clip(np.zeros(n), lo, hi)
    W = np.zeros(n, dtype=np.int8)            # 0 free, -1 pinned at lo, +1 pinned at hi
    if max_iter is None:
        max_iter = 10 * n + 10
    for _ in range(max_iter):
        free = W == 0
        xs = x.copy()
        if free.any():
            rhs = -(c[free] + H[np.ix_(free, ~free)] @ x[~free])
            L = np.linalg.cholesky(H[np.ix_(free, free)])
            xs[free] = np.linalg.solve(L.T, np.linalg.solve(L, rhs))
        p = xs - x
        alpha, blk, side = 1.0, -1, 0
        for i in np.nonzero(free)[0]:
            if p[i] > 0.0 and np.isfinite(hi[i]):
                a = (hi[i] - x[i]) / p[i]
                if a < alpha:
                    alpha, blk, side = a, i, +1
            elif p[i] < 0.0 and np.isfinite(lo[i]):
                a = (lo[i] - x[i]) / p[i]
                if a < alpha:
                    alpha, blk, side = a, i, -1
        if blk >= 0:
            alpha = max(alpha, 0.0)
            x = x + alpha * p
            x[blk] = hi[blk] if side > 0 else lo[blk]
            W[blk] = side
            continue
        x = xs
        g = H @ x + c
        lam = np.where(W < 0, g, np.where(W > 0, -g, 0.0))     # KKT multipliers, want >= 0
        act = np.nonzero(W)[0]
        if act.size == 0 or lam[act].min() >= -1e-12 * max(1.0, float(np.abs(g).max())):
            return x, int(act.size)
        W[act[np.argmin(lam[act])]] = 0
    raise RuntimeError("solve_box_qp: active-set iteration limit reached")


# --------------------------------------------------------------------------------------
# the retargeter (A0-A13)
# --------------------------------------------------------------------------------------
class OracleRetargeter:
    """Float64 restatement of `GeneralMotionRetargeting`.

    `robot` is any object with the flat-tree fields of
    general_motion_retargeting_b200.mjcf.RobotModel (parent, body_pos, body_quat,
    body_hinge, hinge_axis, hinge_lo/hi/limited, qpos0, timestep, body_names);
    `ik_config` is the reference's JSON dict (ik_configs/*.json)."""

    def __init__(self, robot, ik_config: dict, actual_human_height: Optional[float] = None,
                 damping: float = 5e-1, max_iter: int = 10) -> None:
        self.robot = robot
        ik_config = copy.deepcopy(ik_config)
        # motion_retarget.py:36-43
        ratio = (actual_human_height / ik_config["human_height_assumption"]
                 if actual_human_height is not None else 1.0)
        for key in ik_config["human_scale_table"].keys():
            ik_config["human_scale_table"][key] = ik_config["human_scale_table"][key] * ratio
        # :47-59
        self.ik_match_table1 = ik_config["ik_match_table1"]
        self.ik_match_table2 = ik_config["ik_match_table2"]
        self.human_root_name = ik_config["human_root_name"]
        self.robot_root_name = ik_config["robot_root_name"]
        self.use_ik_match_table1 = ik_config["use_ik_match_table1"]
        self.use_ik_match_table2 = ik_config["use_ik_match_table2"]
        self.human_scale_table = ik_config["human_scale_table"]
        self.ground = ik_config["ground_height"] * np.array([0.0, 0.0, 1.0])
        self.max_iter = max_iter
        self.damping = damping

        # :74-114  tasks are (body id, position_cost, orientation_cost, human body name)
        self.tasks1: List[dict] = []
        self.tasks2: List[dict] = []
        self.human_body_to_task1: Dict[str, dict] = {}
        self.human_body_to_task2: Dict[str, dict] = {}
        self.pos_offsets1: Dict[str, np.ndarray] = {}
        self.rot_offsets1: Dict[str, np.ndarray] = {}
        names = list(robot.body_names)
        for table, tasks, h2t, keep_off in (
            (self.ik_match_table1, self.tasks1, self.human_body_to_task1, True),
            (self.ik_match_table2, self.tasks2, self.human_body_to_task2, False),
        ):
            for frame_name, entry in table.items():
                body_name, pos_weight, rot_weight, pos_offset, rot_offset = entry
                if pos_weight != 0 or rot_weight != 0:
                    task = {"body": names.index(frame_name), "frame": frame_name,
                            "cost": np.array([pos_weight] * 3 + [rot_weight] * 3, np.float64),
                            "target": None}
                    h2t[body_name] = task
                    if keep_off:
                        self.pos_offsets1[body_name] = np.array(pos_offset, np.float64) - self.ground
                        q = np.array(rot_offset, np.float64)
                        self.rot_offsets1[body_name] = q / np.linalg.norm(q)   # R.from_quat normalises
                    tasks.append(task)

        # mink.Configuration(model): qpos = qpos0, then update() (:75)
        self.nv = 6 + len(robot.hinge_axis)
        self.dt = float(robot.timestep)                       # model.opt.timestep (:146)
        self.qpos = np.array(robot.qpos0, np.float64).copy()
        self._fk()
        self.scaled_human_data = None
        # statistics of the last retarget() call
        self.last_iters = [0, 0]
        self.last_errors = [0.0, 0.0]
        self.last_active = 0

    # ---- A6: mj_kinematics ---------------------------------------------------------
    def _fk(self) -> None:
        r = self.robot
        nb = len(r.body_names)
        self.xpos = np.zeros((nb, 3))
        self.xquat = np.zeros((nb, 4))
        self.xmat = np.zeros((nb, 3, 3))
        self.xaxis = np.zeros((len(r.hinge_axis), 3))
        for b in range(nb):
            p = int(r.parent[b])
            if p < 0:
                xpos = self.qpos[0:3].copy()
                xquat = quat_normalize(self.qpos[3:7])
            else:
                xpos = self.xpos[p] + self.xmat[p] @ r.body_pos[b]
                xquat = quat_mul(self.xquat[p], r.body_quat[b])
                j = int(r.body_hinge[b])
                if j >= 0:
                    self.xaxis[j] = quat_rotate(xquat, r.hinge_axis[j])     # before the joint rotation
                    qloc = axis_angle_to_quat(r.hinge_axis[j], self.qpos[7 + j])
                    xquat = quat_mul(xquat, qloc)
                    # joint pos = 0 in every supported MJCF: no off-centre correction
                xquat = quat_normalize(xquat)
            self.xpos[b] = xpos
            self.xquat[b] = xquat
            self.xmat[b] = quat_to_mat(xquat)

    # ---- A8: mj_jacBody rotated into the body frame (mink get_frame_jacobian) --------
    def _frame_jacobian(self, body: int) -> np.ndarray:
        r = self.robot
        jacp = np.zeros((3, self.nv))
        jacr = np.zeros((3, self.nv))
        point = self.xpos[body]
        b = body
        while b >= 0:
            j = int(r.body_hinge[b])
            if j >= 0:
                ax = self.xaxis[j]
                jacr[:, 6 + j] = ax
                jacp[:, 6 + j] = np.cross(ax, point - self.xpos[b])       # anchor = body origin
            if int(r.parent[b]) < 0:                                      # free joint
                jacp[:, 0:3] = np.eye(3)
                Rroot = self.xmat[b]
                for k in range(3):
                    jacr[:, 3 + k] = Rroot[:, k]
                    jacp[:, 3 + k] = np.cross(Rroot[:, k], point - self.xpos[b])
            b = int(r.parent[b])
        Rt = self.xmat[body].T
        return np.vstack([Rt @ jacp, Rt @ jacr])

    # ---- A7: FrameTask.compute_error -------------------------------------------------
    def _task_error(self, task: dict) -> np.ndarray:
        if task["target"] is None:
            raise RuntimeError(f"target not set for frame task '{task['frame']}'")   # mink TargetNotSet
        tq, tp = task["target"]
        b = task["body"]
        qb_inv = quat_conj(self.xquat[b])
        # T_b^{-1} T_t
        q_bt = quat_mul(qb_inv, tq)
        t_bt = quat_rotate(qb_inv, tp) + (-quat_rotate(qb_inv, self.xpos[b]))
        return se3_log(q_bt, t_bt)

    # ---- A9: FrameTask.compute_jacobian ----------------------------------------------
    def _task_jacobian(self, task: dict) -> np.ndarray:
        tq, tp = task["target"]
        b = task["body"]
        jac = self._frame_jacobian(b)
        tq_inv = quat_conj(tq)
        q_tb = quat_mul(tq_inv, self.xquat[b])
        t_tb = quat_rotate(tq_inv, self.xpos[b]) + (-quat_rotate(tq_inv, tp))
        return -se3_jlog(q_tb, t_tb) @ jac

    # ---- A9/A10: Task.compute_qp_objective + build_ik ---------------------------------
    def _build_qp(self, tasks: Sequence[dict]):
        nv = self.nv
        H = np.eye(nv) * self.damping
        c = np.zeros(nv)
        for task in tasks:
            J = self._task_jacobian(task)
            minus_gain_error = -TASK_GAIN * self._task_error(task)
            Wt = np.diag(task["cost"])
            WJ = Wt @ J
            We = Wt @ minus_gain_error
            mu = LM_DAMPING * float(We @ We)
            H += WJ.T @ WJ + mu * np.eye(nv)
            c += -We @ WJ
        r = self.robot
        lo = np.full(nv, -np.inf)
        hi = np.full(nv, np.inf)
        for j in range(len(r.hinge_axis)):
            if r.hinge_limited[j]:
                hi[6 + j] = LIMIT_GAIN * (r.hinge_hi[j] - self.qpos[7 + j])
                lo[6 + j] = -LIMIT_GAIN * (self.qpos[7 + j] - r.hinge_lo[j])
        return H, c, lo, hi

    # ---- A11: solve_ik ----------------------------------------------------------------
    def _solve_ik(self, tasks: Sequence[dict]) -> np.ndarray:
        H, c, lo, hi = self._build_qp(tasks)
        dq, nact = solve_box_qp(H, c, lo, hi)
        self.last_active = max(self.last_active, nact)
        return dq / self.dt

    # ---- A12: integrate_inplace (mj_integratePos + update) -----------------------------
    def _integrate(self, v: np.ndarray) -> None:
        dt = self.dt
        self.qpos[0:3] += dt * v[0:3]
        w = v[3:6].copy()
        nrm = math.sqrt(float(w @ w))
        if nrm < MJ_MINVAL:
            axis, nrm = np.array([1.0, 0.0, 0.0]), 0.0
        else:
            axis = w / nrm
        qrot = axis_angle_to_quat(axis, dt * nrm)
        self.qpos[3:7] = quat_mul(quat_normalize(self.qpos[3:7]), qrot)
        self.qpos[7:] += dt * v[6:]
        self._fk()

    # ---- error1 / error2 (:188-200) -----------------------------------------------------
    def _error(self, tasks: Sequence[dict]) -> float:
        return float(np.linalg.norm(np.concatenate([self._task_error(t) for t in tasks])))

    def error1(self) -> float:
        return self._error(self.tasks1)

    def error2(self) -> float:
        return self._error(self.tasks2)

    # ---- A1-A5 -------------------------------------------------------------------------
    @staticmethod
    def to_numpy(human_data):
        for body_name in human_data.keys():
            human_data[body_name] = [np.asarray(human_data[body_name][0], np.float64),
                                     np.asarray(human_data[body_name][1], np.float64)]
        return human_data

    def scale_human_data(self, human_data, human_root_name, human_scale_table):
        root_pos, root_quat = human_data[human_root_name]
        scaled_root_pos = human_scale_table[human_root_name] * root_pos
        out = {human_root_name: (scaled_root_pos, root_quat)}
        for body_name in human_data.keys():
            if body_name not in human_scale_table or body_name == human_root_name:
                continue
            local = (human_data[body_name][0] - root_pos) * human_scale_table[body_name]
            out[body_name] = (local + scaled_root_pos, human_data[body_name][1])
        return out

    def offset_human_data(self, human_data, pos_offsets, rot_offsets):
        out = {}
        for body_name in human_data.keys():
            pos, quat = human_data[body_name]
            q = np.asarray(quat, np.float64)
            q = q / np.linalg.norm(q)                       # R.from_quat normalises
            updated_quat = quat_mul(q, rot_offsets[body_name])
            updated_quat = updated_quat / np.linalg.norm(updated_quat)
            global_pos_offset = quat_rotate(updated_quat, pos_offsets[body_name])
            out[body_name] = [pos + global_pos_offset, updated_quat]
        return out

    def offset_human_data_to_ground(self, human_data):
        lowest = np.inf
        for body_name in human_data.keys():
            if "Foot" not in body_name and "foot" not in body_name:
                continue
            if human_data[body_name][0][2] < lowest:
                lowest = human_data[body_name][0][2]
        out = {}
        for body_name in human_data.keys():
            pos, quat = human_data[body_name]
            out[body_name] = [pos - np.array([0.0, 0.0, lowest]) + np.array([0.0, 0.0, 0.1]), quat]
        return out

    def update_targets(self, human_data, offset_to_ground: bool = False) -> None:
        human_data = self.to_numpy(human_data)
        human_data = self.scale_human_data(human_data, self.human_root_name, self.human_scale_table)
        human_data = self.offset_human_data(human_data, self.pos_offsets1, self.rot_offsets1)
        if offset_to_ground:
            human_data = self.offset_human_data_to_ground(human_data)
        self.scaled_human_data = human_data
        if self.use_ik_match_table1:
            for body_name, task in self.human_body_to_task1.items():
                pos, rot = human_data[body_name]
                task["target"] = (np.asarray(rot, np.float64), np.asarray(pos, np.float64))
        if self.use_ik_match_table2:
            for body_name, task in self.human_body_to_task2.items():
                pos, rot = human_data[body_name]
                task["target"] = (np.asarray(rot, np.float64), np.asarray(pos, np.float64))

    # ---- A13 ---------------------------------------------------------------------------
    def _stage(self, tasks: Sequence[dict]) -> Tuple[int, float]:
        curr_error = self._error(tasks)
        self._integrate(self._solve_ik(tasks))
        next_error = self._error(tasks)
        nsolve = 1
        num_iter = 0
        while curr_error - next_error > 0.001 and num_iter < self.max_iter:
            curr_error = next_error
            self._integrate(self._solve_ik(tasks))
            next_error = self._error(tasks)
            num_iter += 1
            nsolve += 1
        return nsolve, next_error

    def retarget(self, human_data, offset_to_ground: bool = False) -> np.ndarray:
        self.update_targets(human_data, offset_to_ground)
        self.last_iters = [0, 0]
        self.last_errors = [0.0, 0.0]
        self.last_active = 0
        if self.use_ik_match_table1:
            self.last_iters[0], self.last_errors[0] = self._stage(self.tasks1)
        if self.use_ik_match_table2:
            self.last_iters[1], self.last_errors[1] = self._stage(self.tasks2)
        return self.qpos.copy()


def retarget_clip(robot, ik_config: dict, human_names: Sequence[str], pos: np.ndarray, quat: np.ndarray,
                  actual_human_height: Optional[float], damping: float = 0.5,
                  offset_to_ground: bool = False):
    """Run one clip (pos[T,nh,3], quat[T,nh,4] wxyz, bodies named `human_names`) through a
    fresh OracleRetargeter.  Returns qpos[T,nq], iters[T,2], err[T,2]."""
    o = OracleRetargeter(robot, ik_config, actual_human_height, damping)
    T = pos.shape[0]
    qpos = np.zeros((T, o.qpos.shape[0]))
    iters = np.zeros((T, 2), np.int32)
    err = np.zeros((T, 2))
    for t in range(T):
        frame = {n: (pos[t, i].astype(np.float64), quat[t, i].astype(np.float64))
                 for i, n in enumerate(human_names)}
        qpos[t] = o.retarget(frame, offset_to_ground)
        iters[t] = o.last_iters
        err[t] = o.last_errors
    return qpos, iters, err
