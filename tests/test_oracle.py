"""Pins for the CPU oracle (oracle/gmr_oracle.py, oracle/gmr_oracle.cpp).

The reference's IK dependencies (mink/mujoco/daqp) are absent, so the oracle is pinned by
 (a) outputs of the reference's OWN code where it can run here (tests/golden/reference_*.npz),
 (b) independent mathematics: scipy Rotation for SO(3), finite differences for every Jacobian,
     brute-force KKT enumeration and scipy BVLS for the box QP,
 (c) agreement of the two independent restatements (NumPy vs C++)."""
import itertools
import os

import numpy as np
import pytest
from scipy.optimize import lsq_linear
from scipy.spatial.transform import Rotation as R

import oracle.gmr_oracle as O
from helpers import problem
from general_motion_retargeting_b200.synthetic import make_clips

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_so3_log_matches_scipy():
    rng = np.random.default_rng(0)
    for _ in range(200):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        rv = R.from_quat(q, scalar_first=True).as_rotvec()
        np.testing.assert_allclose(O.so3_log(q), rv, atol=1e-12)
    # tiny angle and sign invariance
    q = np.array([1.0, 1e-9, -2e-9, 0.5e-9]); q /= np.linalg.norm(q)
    np.testing.assert_allclose(O.so3_log(q), 2 * q[1:], rtol=1e-9)
    np.testing.assert_allclose(O.so3_log(-q), O.so3_log(q), atol=1e-15)


def _se3_exp(xi):
    rho, om = xi[:3], xi[3:]
    th = np.linalg.norm(om)
    S = O.skew(om)
    if th < 1e-9:
        Rm, V = np.eye(3) + S, np.eye(3) + 0.5 * S
    else:
        Rm = np.eye(3) + np.sin(th) / th * S + (1 - np.cos(th)) / th ** 2 * S @ S
        V = np.eye(3) + (1 - np.cos(th)) / th ** 2 * S + (th - np.sin(th)) / th ** 3 * S @ S
    return R.from_matrix(Rm).as_quat(scalar_first=True), V @ rho


def test_se3_log_inverts_exp_and_jlog_is_its_derivative():
    rng = np.random.default_rng(1)
    for scale in (1.5, 0.3, 0.02):
        xi = rng.normal(size=6) * scale
        q, t = _se3_exp(xi)
        np.testing.assert_allclose(O.se3_log(q, t), xi, atol=1e-10)
        # jlog = J_r^{-1}: log(T * exp(d)) ~ log(T) + jlog(T) d
        J = O.se3_jlog(q, t)
        Jn = np.zeros((6, 6))
        h = 1e-6
        Rm = O.quat_to_mat(q)
        for k in range(6):
            d = np.zeros(6); d[k] = h
            qd, td = _se3_exp(d)
            qp, tp = O.quat_mul(q, qd), Rm @ td + t
            dm = np.zeros(6); dm[k] = -h
            qd2, td2 = _se3_exp(dm)
            qm, tm = O.quat_mul(q, qd2), Rm @ td2 + t
            Jn[:, k] = (O.se3_log(qp, tp) - O.se3_log(qm, tm)) / (2 * h)
        np.testing.assert_allclose(J, Jn, atol=2e-8)


def _oracle_at_random_config(src, robot, seed):
    m, tt, pack = problem(src, robot)
    clips = make_clips(m, tt, [seed], T=2, src_human=src)
    o = O.OracleRetargeter(m, pack["ik_config"], float(clips.heights[0]))
    rng = np.random.default_rng(seed)
    o.qpos[7:] = np.clip(rng.normal(0, 0.4, m.nhinge), m.hinge_lo + 0.05, m.hinge_hi - 0.05)
    q = rng.normal(size=4); o.qpos[3:7] = q / np.linalg.norm(q)
    o.qpos[0:3] = rng.uniform(-1, 1, 3)
    o._fk()
    frame = {n: (clips.pos[0, 1, i].astype(float), clips.quat[0, 1, i].astype(float)) for i, n in enumerate(tt.human_names)}
    o.update_targets(frame)
    return m, o


@pytest.mark.parametrize("src,robot", [("smplx", "unitree_g1"), ("bvh", "booster_t1"), ("smplx", "engineai_pm01")])
def test_task_jacobian_is_derivative_of_error_under_integration(src, robot):
    """J = -jlog(T_tb) J_b must equal d e(q (+) dq)/d dq with MuJoCo's integration convention
    (free joint: world translation, body-local rotation, right-multiplied)."""
    m, o = _oracle_at_random_config(src, robot, 3)
    q0 = o.qpos.copy()
    h = 1e-6
    for task in o.tasks1[:4] + o.tasks1[-3:]:
        J = o._task_jacobian(task)
        Jn = np.zeros_like(J)
        for k in range(o.nv):
            for sgn in (+1, -1):
                o.qpos[:] = q0; o._fk()
                v = np.zeros(o.nv); v[k] = sgn * h / o.dt
                o._integrate(v)
                e = o._task_error(task)
                Jn[:, k] += sgn * e / (2 * h)
        o.qpos[:] = q0; o._fk()
        np.testing.assert_allclose(J, Jn, atol=5e-8)


def test_box_qp_against_bruteforce_kkt_and_bvls():
    rng = np.random.default_rng(5)
    for trial in range(60):
        n = int(rng.integers(2, 7))
        A = rng.normal(size=(n + 2, n))
        H = A.T @ A + 0.1 * np.eye(n)
        c = rng.normal(size=n) * 3
        lo = -rng.uniform(0, 1, n); hi = rng.uniform(0, 1, n)
        if trial % 5 == 0:
            lo[0] = 0.0                      # bound exactly at the start point
        if trial % 7 == 0:
            lo[1], hi[1] = -np.inf, np.inf   # an unbounded variable (floating base)
        x, nact = O.solve_box_qp(H, c, lo, hi)
        # brute force: every assignment of {free, lo, hi}; keep the feasible KKT point
        best = None
        for assign in itertools.product((0, -1, 1), repeat=n):
            if any((a == -1 and not np.isfinite(lo[i])) or (a == 1 and not np.isfinite(hi[i])) for i, a in enumerate(assign)):
                continue
            a = np.array(assign)
            xs = np.where(a < 0, lo, np.where(a > 0, hi, 0.0))
            xs = np.where(np.isfinite(xs), xs, 0.0)
            F = a == 0
            if F.any():
                xs[F] = np.linalg.solve(H[np.ix_(F, F)], -(c[F] + H[np.ix_(F, ~F)] @ xs[~F]))
            if np.any(xs < lo - 1e-12) or np.any(xs > hi + 1e-12):
                continue
            g = H @ xs + c
            if np.all(g[a < 0] >= -1e-10) and np.all(g[a > 0] <= 1e-10):
                best = xs
                break
        assert best is not None
        np.testing.assert_allclose(x, best, atol=1e-9)
        # BVLS on the Cholesky factor: min 1/2 |L^T x + L^-1 c|^2
        L = np.linalg.cholesky(H)
        res = lsq_linear(L.T, -np.linalg.solve(L, c), bounds=(lo, hi), method="bvls", tol=1e-14)
        np.testing.assert_allclose(x, res.x, atol=1e-7)


def test_preprocessing_matches_reference_code():
    """scale_human_data / offset_human_data / offset_human_data_to_ground vs the reference's own
    methods (motion_retarget.py:209-270) run by tools/make_golden.py."""
    g = np.load(os.path.join(GOLD, "reference_preprocess.npz"))
    for key, (src, robot) in {"smplx_unitree_g1": ("smplx", "unitree_g1"), "bvh_booster_t1": ("bvh", "booster_t1"),
                              "smplx_hightorque_hi": ("smplx", "hightorque_hi")}.items():
        m, tt, pack = problem(src, robot)
        for ground, name in ((False, ".targets"), (True, ".targets_ground")):
            o = O.OracleRetargeter(m, pack["ik_config"], float(g[key + ".height"]))
            for t in range(3):
                frame = {n: (g[key + ".pos"][t, i].astype(float), g[key + ".quat"][t, i].astype(float)) for i, n in enumerate(tt.human_names)}
                frame["extra_body_not_in_table"] = (np.zeros(3), np.array([1.0, 0, 0, 0]))
                o.update_targets(frame, offset_to_ground=ground)
                assert set(o.scaled_human_data) == set(tt.human_names)       # extras are dropped (:219)
                for i, n in enumerate(tt.human_names):
                    p, q = o.scaled_human_data[n]
                    ref = g[key + name][t, i]
                    np.testing.assert_allclose(p, ref[:3], atol=1e-12)
                    assert min(np.abs(q - ref[3:]).max(), np.abs(q + ref[3:]).max()) < 1e-12


def test_retarget_loop_shape_and_warm_start():
    m, tt, pack = problem("smplx", "unitree_g1")
    clips = make_clips(m, tt, [0], T=6)
    q, it, err = O.retarget_clip(m, pack["ik_config"], tt.human_names, clips.pos[0], clips.quat[0], float(clips.heights[0]))
    assert it.min() >= 1 and it.max() <= 11                 # 1 unconditional + <= 10 conditional per stage
    assert it[0, 0] == 11                                    # first frame starts from qpos0 far away: hits the cap
    assert (it[2:].sum(1) <= 6).all()                        # warm-started frames converge quickly
    assert (q[:, 7:] >= m.hinge_lo - 1e-9).all() and (q[:, 7:] <= m.hinge_hi + 1e-9).all()
    # a missing table body raises KeyError like the reference (:129 / :241)
    o = O.OracleRetargeter(m, pack["ik_config"], 1.7)
    frame = {n: (clips.pos[0, 0, i].astype(float), clips.quat[0, 0, i].astype(float)) for i, n in enumerate(tt.human_names)}
    del frame["left_foot"]
    with pytest.raises(KeyError):
        o.retarget(frame)


@pytest.mark.parametrize("src,robot,stress", [("smplx", "unitree_g1", False), ("bvh", "booster_t1", False),
                                              ("smplx", "stanford_toddy", True), ("smplx", "kuavo_s45", False)])
def test_cpp_port_matches_numpy_oracle(built, src, robot, stress):
    from oracle import native
    m, tt, pack = problem(src, robot)
    clips = make_clips(m, tt, [1, 2], T=8, src_human=src, stress=stress)
    qc, itc, errc = native.retarget_batch(m, tt, clips.pos, clips.quat, clips.ratio(tt))
    for c in range(2):
        height = float(clips.ratio(tt)[c]) * tt.height_assumption
        qn, itn, errn = O.retarget_clip(m, pack["ik_config"], tt.human_names, clips.pos[c], clips.quat[c], height)
        np.testing.assert_array_equal(itn, itc[c])
        np.testing.assert_allclose(qn, qc[c], atol=1e-9)
        np.testing.assert_allclose(errn, errc[c], atol=1e-9)


def test_oracle_traces_fixture_is_reproducible(built):
    """tests/golden/oracle_traces.npz (the known-answer vectors of the GPU tests) = today's oracle."""
    from oracle import native
    g = np.load(os.path.join(GOLD, "oracle_traces.npz"))
    from general_motion_retargeting_b200._native import LIE_EPS_ROUND1
    for key in sorted({k.rsplit(".", 1)[0] for k in g.files}):
        src, robot = key.split("_", 1)
        m, tt, _ = problem(src, robot)
        ratio = (g[key + ".heights"].astype(np.float64) / float(tt.height_assumption)).astype(np.float32)
        for suffix, eps in (("", 0.0), ("_eps10", LIE_EPS_ROUND1)):           # both values of the Lie threshold
            q, it, err = native.retarget_batch(m, tt, g[key + ".pos"], g[key + ".quat"], ratio, lie_eps=eps)
            np.testing.assert_array_equal(it, g[key + ".iters" + suffix])
            np.testing.assert_allclose(q, g[key + ".qpos" + suffix], atol=1e-9)
