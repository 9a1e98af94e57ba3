"""The kernel's algorithm (csrc/gmr_solver.cuh), built lane-serially for the host (tests/emu),
against the float64 oracle.  This exercises every phase of the warp solver — composite
"spring inertia" assembly, register-row Cholesky, warm-started active set, flattened
stage loop — on CPU; the GPU tests repeat the comparison on the real kernel."""
import numpy as np
import pytest

from conftest import ALL_PAIRS
from helpers import compare, emu_retarget_batch, problem
from general_motion_retargeting_b200.synthetic import make_clips


@pytest.mark.parametrize("src,robot", ALL_PAIRS)
def test_emulated_kernel_matches_oracle_f64(built, src, robot):
    from oracle import native
    m, tt, _ = problem(src, robot)
    clips = make_clips(m, tt, range(4), T=25, src_human=src)
    q, it, err, tg, _ = emu_retarget_batch(m, tt, clips.pos, clips.quat, clips.ratio(tt), bits=64)
    # (1) the literal restatement of mink's formulas: the gate of BASELINE.json (1e-3 rad).  mink's
    #     closed forms for J_l^-1 / Q cancel for tiny rotation errors, which alone moves qpos by ~1e-4.
    q_ref, it_ref, err_ref = native.retarget_batch(m, tt, clips.pos, clips.quat, clips.ratio(tt))
    agree, dq_all, dq_clean = compare(q, it, q_ref, it_ref)
    assert agree == 1.0
    assert dq_all < 1e-3, dq_all
    # (2) the same oracle with cancellation-free coefficients: float64 kernel arithmetic is exact to rounding
    q_st, it_st, err_st = native.retarget_batch(m, tt, clips.pos, clips.quat, clips.ratio(tt), flags=native.FLAG_STABLE_LIE)
    agree, dq_all, dq_clean = compare(q, it, q_st, it_st)
    assert agree == 1.0
    assert dq_all < 1e-9, dq_all
    np.testing.assert_allclose(err, err_st, atol=1e-9)
    assert np.abs(np.linalg.norm(q[..., 3:7], axis=-1) - 1).max() < 1e-12
    assert (q[..., 7:] >= m.hinge_lo - 1e-7).all() and (q[..., 7:] <= m.hinge_hi + 1e-7).all()


@pytest.mark.parametrize("src,robot", [("smplx", "unitree_g1"), ("bvh", "booster_t1"), ("smplx", "hightorque_hi")])
def test_emulated_kernel_f32_within_tolerance(built, src, robot):
    """BASELINE.json tolerance: max |dqpos| <= 1e-3 rad on identical inputs and iteration counts."""
    from oracle import native
    m, tt, _ = problem(src, robot)
    clips = make_clips(m, tt, range(16), T=40, src_human=src)
    q_ref, it_ref, _ = native.retarget_batch(m, tt, clips.pos, clips.quat, clips.ratio(tt))
    q, it, err, tg, _ = emu_retarget_batch(m, tt, clips.pos, clips.quat, clips.ratio(tt), bits=32)
    agree, dq_all, dq_clean = compare(q, it, q_ref, it_ref)
    assert agree > 0.995
    assert dq_clean < 1e-3, dq_clean


@pytest.mark.parametrize("src,robot", [("smplx", "stanford_toddy"), ("bvh", "booster_t1"), ("smplx", "unitree_g1")])
def test_active_set_path_on_unreachable_targets(built, src, robot):
    """Stress clips (5 cm / 0.3 rad noise) drive joints into their limits: the warm-started
    primal active set must land on the same (unique) optimum as the oracle's solver."""
    from oracle import native
    m, tt, _ = problem(src, robot)
    clips = make_clips(m, tt, range(6), T=20, src_human=src, stress=True)
    q_ref, it_ref, _ = native.retarget_batch(m, tt, clips.pos, clips.quat, clips.ratio(tt))
    q, it, err, tg, nfac = emu_retarget_batch(m, tt, clips.pos, clips.quat, clips.ratio(tt), bits=64)
    agree, dq_all, dq_clean = compare(q, it, q_ref, it_ref)
    assert nfac > it.sum()                       # some solves needed more than one factorisation
    assert agree > 0.99 and dq_clean < 1e-3
    q_st, it_st, _ = native.retarget_batch(m, tt, clips.pos, clips.quat, clips.ratio(tt), flags=native.FLAG_STABLE_LIE)
    agree, dq_all, dq_clean = compare(q, it, q_st, it_st)
    assert agree > 0.99 and dq_clean < 1e-8, dq_clean
    on_limit = (np.abs(q[..., 7:] - m.hinge_lo) < 1e-6) | (np.abs(q[..., 7:] - m.hinge_hi) < 1e-6)
    assert on_limit.any() or (q_ref[..., 7:] - m.hinge_lo).min() < 1e-3 or (m.hinge_hi - q_ref[..., 7:]).min() < 1e-3


def test_targets_out_no_solve_qpos_init_and_ground(built):
    from oracle import native
    import oracle.gmr_oracle as O
    m, tt, pack = problem("smplx", "unitree_g1")
    clips = make_clips(m, tt, [4], T=10)
    ratio = clips.ratio(tt)
    # targets_out = scaled_human_data of the oracle
    q, it, err, tg, _ = emu_retarget_batch(m, tt, clips.pos, clips.quat, ratio, bits=64)
    o = O.OracleRetargeter(m, pack["ik_config"], float(ratio[0]) * tt.height_assumption)
    frame = {n: (clips.pos[0, 3, i].astype(float), clips.quat[0, 3, i].astype(float)) for i, n in enumerate(tt.human_names)}
    o.update_targets(frame)
    for i, n in enumerate(tt.human_names):
        np.testing.assert_allclose(tg[0, 3, i, :3], o.scaled_human_data[n][0], atol=1e-12)
        np.testing.assert_allclose(tg[0, 3, i, 3:], o.scaled_human_data[n][1], atol=1e-12)
    # splitting a clip in two calls chained through qpos_init == one call (warm start is the only state)
    qa, ita, _, _, _ = emu_retarget_batch(m, tt, clips.pos[:, :6], clips.quat[:, :6], ratio, bits=64)
    qb, itb, _, _, _ = emu_retarget_batch(m, tt, clips.pos[:, 6:], clips.quat[:, 6:], ratio, bits=64, qpos_init=qa[:, -1])
    np.testing.assert_allclose(np.concatenate([qa, qb], 1), q, atol=1e-9)
    np.testing.assert_array_equal(np.concatenate([ita, itb], 1), it)
    # NO_SOLVE: qpos untouched, errors = error1()/error2() at that configuration
    qn, itn, errn, _, _ = emu_retarget_batch(m, tt, clips.pos[:, 6:7], clips.quat[:, 6:7], ratio, bits=64, qpos_init=qa[:, -1], flags=2)
    np.testing.assert_allclose(qn[:, 0], qa[:, -1], atol=0)
    assert (itn == 0).all()
    o2 = O.OracleRetargeter(m, pack["ik_config"], float(ratio[0]) * tt.height_assumption)
    o2.qpos[:] = qa[0, -1]; o2._fk()
    frame = {n: (clips.pos[0, 6, i].astype(float), clips.quat[0, 6, i].astype(float)) for i, n in enumerate(tt.human_names)}
    o2.update_targets(frame)
    np.testing.assert_allclose(errn[0, 0], [o2.error1(), o2.error2()], atol=1e-9)
    # offset_to_ground flag
    qg_ref, itg_ref, _ = native.retarget_batch(m, tt, clips.pos, clips.quat, ratio, flags=1 | native.FLAG_STABLE_LIE)
    qg, itg, _, _, _ = emu_retarget_batch(m, tt, clips.pos, clips.quat, ratio, bits=64, flags=1)
    np.testing.assert_array_equal(itg, itg_ref)
    np.testing.assert_allclose(qg, qg_ref, atol=1e-9)
    assert np.abs(qg - q).max() > 1e-3


def test_single_stage_and_iteration_cap(built):
    from oracle import native
    m, tt, _ = problem("smplx", "kuavo_s45")         # use_ik_match_table2 = false
    clips = make_clips(m, tt, [0, 1], T=12)
    q, it, err, _, _ = emu_retarget_batch(m, tt, clips.pos, clips.quat, clips.ratio(tt), bits=64)
    assert (it[..., 1] == 0).all() and (it[..., 0] >= 1).all() and it.max() <= 11
    # max_iter = 0: exactly one solve per enabled stage
    m, tt, _ = problem("smplx", "unitree_g1")
    clips = make_clips(m, tt, [0], T=5)
    q, it, err, _, _ = emu_retarget_batch(m, tt, clips.pos, clips.quat, clips.ratio(tt), bits=64, max_iter=0)
    assert (it == 1).all()


def test_fk_epilogue_ragged_lengths_and_warm_state(built):
    """GmrBatchExtra: local_body_pos / lowest_z (the dataset scripts' post-solve FK, fused), ragged clip
    lengths, and the working-set hand-over that lets a clip be fed in pieces."""
    from helpers import emu_retarget_batch_ex, oracle_body_positions
    m, tt, pack = problem("smplx", "unitree_g1")
    clips = make_clips(m, tt, [2, 3, 5], T=12)
    ratio = clips.ratio(tt)
    lengths = np.array([12, 7, 1], np.int32)
    q, it, lbp, low, _ = emu_retarget_batch_ex(m, tt, clips.pos, clips.quat, ratio, lengths=lengths)
    q_full, it_full, _, _, _ = emu_retarget_batch(m, tt, clips.pos, clips.quat, ratio, bits=64)
    for c, n in enumerate(lengths):
        np.testing.assert_array_equal(q[c, :n], q_full[c, :n])        # a shorter clip is a prefix of the longer one
        assert (q[c, n:] == 0).all() and (it[c, n:] == 0).all()         # padding frames are not touched
        local, world = oracle_body_positions(m, pack, q[c, :n])
        np.testing.assert_allclose(lbp[c, :n], local, atol=2e-6)
        assert abs(low[c] - world[..., 2].min()) < 2e-6
    # with joints resting on their limits, splitting a clip needs qpos AND the working sets to be identical
    m, tt, _ = problem("smplx", "unitree_g1")
    clips = make_clips(m, tt, [878], T=16)               # a clip that settles with ~10 joints on their limits
    ratio = clips.ratio(tt)
    q_all, it_all, _, _, w_all = emu_retarget_batch_ex(m, tt, clips.pos, clips.quat, ratio)
    qa, ita, _, _, wa = emu_retarget_batch_ex(m, tt, clips.pos[:, :9], clips.quat[:, :9], ratio)
    qb, itb, _, _, wb = emu_retarget_batch_ex(m, tt, clips.pos[:, 9:], clips.quat[:, 9:], ratio, qpos_init=qa[:, -1], warm_state=wa)
    np.testing.assert_allclose(np.concatenate([qa, qb], 1), q_all, atol=1e-12)
    np.testing.assert_array_equal(np.concatenate([ita, itb], 1), it_all)
    np.testing.assert_array_equal(wb, w_all)
    assert w_all.any()                                               # some bound was active at the end


def test_compiled_in_limits_are_reported_not_truncated(built, tmp_path):
    """A chain of more than GMR_MAXD (10) hinges, or more than 32 hinges, is refused with GMR_ELIMIT (-4) by the
    same gmr_fill_consts that gmr_model_create runs; a 10-deep arm is accepted and solved."""
    import ctypes as C
    import helpers
    from general_motion_retargeting_b200._native import build_desc
    from general_motion_retargeting_b200.ik_config import IKConfig, compile_task_table
    from general_motion_retargeting_b200.mjcf import load_mjcf

    def chain_robot(n):
        body = ""
        for i in reversed(range(n)):
            body = f'<body name="l{i}" pos="0 0 0.1"><joint name="j{i}" axis="{int(i % 2 == 0)} {int(i % 2 == 1)} 0" range="-1 1"/>{body}</body>'
        xml = f'<mujoco><compiler angle="radian"/><worldbody><body name="base" pos="0 0 1"><freejoint/>{body}</body></worldbody></mujoco>'
        p = tmp_path / f"chain{n}.xml"
        p.write_text(xml)
        m = load_mjcf(str(p))
        ent = lambda frame, human: {frame: [human, 10, 5, [0, 0, 0], [1, 0, 0, 0]]}
        table = {**ent("base", "root"), **ent(f"l{n - 1}", "tip")}
        cfg = IKConfig.from_dict({"robot_root_name": "base", "human_root_name": "root", "ground_height": 0.0,
                                  "human_height_assumption": 1.8, "use_ik_match_table1": True, "use_ik_match_table2": True,
                                  "human_scale_table": {"root": 1.0, "tip": 1.0}, "ik_match_table1": table, "ik_match_table2": table})
        return m, compile_task_table(m, cfg)

    helpers.emu_retarget_batch(*chain_robot(2), np.zeros((1, 1, 2, 3), np.float32), np.tile([1, 0, 0, 0], (1, 1, 2, 1)).astype(np.float32), None)
    emu = helpers._emu
    for n, want in ((10, 0), (11, -4)):
        m, tt = chain_robot(n)
        desc, keep = build_desc(m, tt)
        pos = np.zeros((1, 2, 2, 3), np.float32); pos[..., 1, 2] = 1.0 + 0.1 * n
        quat = np.tile(np.array([1, 0, 0, 0], np.float32), (1, 2, 2, 1))
        q = np.zeros((1, 2, m.nq))
        rc = emu.gmr_emu_retarget_batch(C.byref(desc), C.c_void_p(pos.ctypes.data), C.c_void_p(quat.ctypes.data), None, 1, 2, None,
                                        C.c_void_p(q.ctypes.data), None, None, None, 0, 1, 64, None)
        assert rc == want, (n, rc)
        if want == 0:
            assert np.isfinite(q).all() and abs(np.linalg.norm(q[0, -1, 3:7]) - 1) < 1e-12


@pytest.mark.parametrize("bits", [64, 32])
def test_lane_blocks_do_not_depend_on_the_lane_order(built, bits):
    """compute-sanitizer is closed on the GPU pool (profiles/r2_compute_sanitizer_closed.txt), so the intra-warp hazards of the
    lane blocks are checked here: the emulator visits the 32 lanes of every block in a given order, and the kernel's rule -
    inside one block a lane never reads shared memory another lane writes in the same block - means the result cannot depend
    on it.  Normal and stress clips (active sets, several factorisations per solve), FK epilogue and ragged lengths included;
    identity, reversed and two shuffled orders must agree bit for bit."""
    import ctypes as C
    import helpers
    from helpers import emu_retarget_batch_ex
    m, tt, _ = problem("smplx", "unitree_g1")
    a = make_clips(m, tt, range(3), T=10)
    b = make_clips(m, tt, range(3), T=10, stress=True)
    pos, quat = np.concatenate([a.pos, b.pos]), np.concatenate([a.quat, b.quat])
    ratio = np.concatenate([a.ratio(tt), b.ratio(tt)])
    lengths = np.array([10, 7, 10, 10, 9, 10], np.int32)
    rng = np.random.default_rng(5)
    orders = [np.arange(32), np.arange(32)[::-1], rng.permutation(32), rng.permutation(32)]
    results = []
    try:
        for o in orders:
            emu_retarget_batch(m, tt, pos[:1], quat[:1], ratio[:1], bits=bits)          # makes sure the library is loaded
            arr = (C.c_int * 32)(*[int(x) for x in o])
            assert helpers._emu.gmr_emu_set_lane_order(arr) == 0
            q, it, err, tg, nfac = emu_retarget_batch(m, tt, pos, quat, ratio, bits=bits, nthreads=1)
            ex = emu_retarget_batch_ex(m, tt, pos, quat, ratio, lengths=lengths, bits=bits)
            results.append((q, it, err, tg, nfac) + tuple(ex))
    finally:
        helpers._emu.gmr_emu_set_lane_order((C.c_int * 32)(*range(32)))
    assert results[0][4] > results[0][1].sum()                                           # the active-set path was exercised
    for r in results[1:]:
        for x, y in zip(results[0], r):
            np.testing.assert_array_equal(np.asarray(x), np.asarray(y))
    assert helpers._emu.gmr_emu_set_lane_order((C.c_int * 32)(*([0] * 32))) == -1         # not a permutation


@pytest.mark.parametrize("src,robot", [("smplx", "unitree_g1"), ("bvh", "booster_t1"), ("smplx", "hightorque_hi"), ("smplx", "stanford_toddy")])
@pytest.mark.parametrize("bits", [64, 32])
def test_shared_memory_layout_is_aligned_and_bank_friendly(src, robot, bits):
    """The per-warp working set: every region 16-byte aligned (128-bit accesses), regions disjoint, the records
    that consecutive lanes touch at strides that are ODD in 16-byte units (a quarter-warp's 128-bit access is then one
    wavefront), the task-block region large enough to be the FK scan's second pose buffer, the factor rows inside the union -
    and, for G1 in float64, 16 warps per CTA within the 227 KB a B200 SM offers (the occupancy the headline numbers rest on)."""
    import ctypes as C
    from general_motion_retargeting_b200._native import build_desc
    import helpers
    m, tt, _ = helpers.problem(src, robot)
    c0 = make_clips(m, tt, [0], T=1)
    helpers.emu_retarget_batch(m, tt, c0.pos, c0.quat, c0.ratio(tt), bits=bits)   # loads the library
    desc, keep = build_desc(m, tt)
    out = (C.c_int32 * 16)()
    assert helpers._emu.gmr_emu_layout(C.byref(desc), bits, out) == 0
    gs_var, o_u, o_xq, o_xp, o_sd, o_tg, o_in, wel, PX, SD, TG, MT, rs, rowh, cbytes, esz = list(out)
    per16 = 16 // esz                                                       # elements per 16-byte unit
    for off in (gs_var, o_u, o_xq, o_xp, o_sd, o_tg, o_in, wel):
        assert off % per16 == 0, "region not 16-byte aligned"
    for stride in (PX, MT, rs) + ((SD, TG) if bits == 64 else ()):
        assert stride % per16 == 0 and (stride // per16) % 2 == 1, f"stride {stride} is not odd in 16-byte units"
    nb, nh, nt, nhum = m.nbody, m.nhinge, tt.nt, tt.nh
    assert o_u == gs_var and o_xp == o_xq + 4
    assert o_xq - o_u >= max(MT * nt, PX * nb)                               # task blocks; also the scan's second pose buffer
    assert o_sd >= o_xq + PX * nb and o_sd >= o_u + rowh + rs * nh - rowh    # poses and factor rows end inside the union
    assert o_tg >= o_sd + SD * nh and o_in >= o_tg + TG * nhum and wel > o_in
    if robot == "unitree_g1" and bits == 64:
        assert 16 + cbytes + 16 * wel * esz <= 232448, "G1 float64 no longer fits 16 warps per CTA"
