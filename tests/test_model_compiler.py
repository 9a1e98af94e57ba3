"""Host logic: MJCF compiler, IK-config compiler, registries, committed model packs."""
import json
import os

import numpy as np
import pytest

from conftest import ALL_PAIRS, REFERENCE_ROOT, needs_reference
from general_motion_retargeting_b200 import params
from general_motion_retargeting_b200.ik_config import IKConfig, compile_task_table
from general_motion_retargeting_b200.mjcf import MjcfError, load_mjcf

GOLD = os.path.join(os.path.dirname(__file__), "golden")

# robot -> (bodies, hinges)
SIZES = {"unitree_g1": (38, 29), "booster_t1": (32, 21), "booster_t1_4dof": (26, 21), "stanford_toddy": (33, 22),
         "fourier_n1": (29, 23), "engineai_pm01": (29, 24), "hightorque_hi": (26, 25), "kuavo_s45": (29, 28)}


@pytest.mark.parametrize("src,robot", ALL_PAIRS)
def test_pack_loads_and_has_expected_sizes(src, robot):
    m, cfg, pack = params.load_pack(src, robot)
    assert (m.nbody, m.nhinge) == SIZES[robot]
    assert m.nq == 7 + m.nhinge and m.nv == 6 + m.nhinge
    assert m.parent[0] == -1 and all(0 <= m.parent[b] < b for b in range(1, m.nbody))
    np.testing.assert_allclose(np.linalg.norm(m.body_quat, axis=1), 1.0, atol=1e-12)
    np.testing.assert_allclose(np.linalg.norm(m.hinge_axis, axis=1), 1.0, atol=1e-12)
    assert m.hinge_limited.all() and (m.hinge_lo < m.hinge_hi).all()
    assert m.timestep == 0.002
    tt = compile_task_table(m, cfg)
    assert tt.nt in (13, 14, 15) and tt.nh == tt.nt
    assert tt.use1 and tt.use2 == (robot != "kuavo_s45")
    # every human body has exactly one robot frame, offsets are unit quaternions
    assert sorted(tt.task_human.tolist()) == list(range(tt.nh))
    np.testing.assert_allclose(np.linalg.norm(tt.rot_off, axis=1), 1.0, atol=1e-12)


def test_unknown_names_raise_keyerror_like_the_reference():
    with pytest.raises(KeyError):
        params.load_pack("smplx", "no_such_robot")
    with pytest.raises(KeyError):
        params.load_pack("bvh", "kuavo_s45")        # no bvh config for kuavo (params.py:28-35)


@needs_reference
@pytest.mark.parametrize("src,robot", ALL_PAIRS)
def test_committed_pack_equals_fresh_compile(src, robot):
    fresh = params.compile_pack(src, robot, REFERENCE_ROOT)
    with open(params.pack_path(src, robot)) as f:
        committed = json.load(f)
    assert json.loads(json.dumps(fresh)) == committed


@needs_reference
def test_nested_include_and_class_defaults():
    # engineai_pm01/pm_v2.xml -> xml/serial_pm_v2.xml -> xml/serial_links.xml; oblique axes
    m = load_mjcf(os.path.join(REFERENCE_ROOT, "assets", "engineai_pm01", "pm_v2.xml"))
    assert m.body_names[0] == "LINK_BASE" and m.nhinge == 24
    np.testing.assert_allclose(m.hinge_axis[0], [0, 0.965926, -0.258819], atol=1e-6)
    # fourier_n1: class default range="0 0" overridden per joint
    n1 = load_mjcf(os.path.join(REFERENCE_ROOT, "assets", "fourier_n1", "n1_mocap.xml"))
    assert n1.hinge_limited.all()
    j = n1.hinge_names.index("left_hip_pitch_joint")
    assert (n1.hinge_lo[j], n1.hinge_hi[j]) == (-2.617, 2.617)


def test_reference_kinematics_model_tree_matches(tmp_path):
    """Body order / parents / FK against the reference's own KinematicsModel (golden made by
    tools/make_golden.py from reference kinematics_model.py:101-163,213-246)."""
    from oracle.gmr_oracle import OracleRetargeter
    g = np.load(os.path.join(GOLD, "reference_fk.npz"))
    robots = sorted({k.split(".")[0] for k in g.files})
    assert len(robots) == 7                                   # pm01 cannot be parsed by the reference class
    for robot in robots:
        src = "bvh" if robot == "booster_t1_4dof" else "smplx"
        m, cfg, pack = params.load_pack(src, robot)
        assert list(g[f"{robot}.body_names"]) == m.body_names
        np.testing.assert_array_equal(g[f"{robot}.parent"], m.parent)
        o = OracleRetargeter(m, pack["ik_config"])
        for t in range(g[f"{robot}.dof"].shape[0]):
            o.qpos[0:3] = g[f"{robot}.root_pos"][t]
            o.qpos[3:7] = g[f"{robot}.root_quat_wxyz"][t]
            o.qpos[7:] = g[f"{robot}.dof"][t]
            o._fk()
            np.testing.assert_allclose(o.xpos, g[f"{robot}.body_pos"][t], atol=2e-5)   # reference FK is float32
            ref_q = g[f"{robot}.body_rot_xyzw"][t][:, [3, 0, 1, 2]]
            dots = np.abs(np.sum(o.xquat * ref_q, axis=1))
            np.testing.assert_allclose(dots, 1.0, atol=2e-5)


def test_mjcf_rejects_unsupported(tmp_path):
    p = tmp_path / "bad.xml"
    p.write_text('<mujoco><worldbody><body name="r"><freejoint/><body name="a"><joint type="slide" axis="0 0 1"/></body></body></worldbody></mujoco>')
    with pytest.raises(MjcfError):
        load_mjcf(str(p))
    p.write_text('<mujoco><worldbody><body name="r"><body name="a"/></body></worldbody></mujoco>')
    with pytest.raises(MjcfError):
        load_mjcf(str(p))
    p.write_text('<mujoco><compiler angle="degree"/><worldbody><body name="r" pos="0 0 1"><freejoint/>'
                 '<body name="a" euler="0 0 90"><joint axis="0 0 2" range="-90 90"/></body></body></worldbody></mujoco>')
    m = load_mjcf(str(p))
    np.testing.assert_allclose(m.hinge_lo, [-np.pi / 2]); np.testing.assert_allclose(m.hinge_axis, [[0, 0, 1]])
    np.testing.assert_allclose(m.body_quat[1], [np.cos(np.pi / 4), 0, 0, np.sin(np.pi / 4)], atol=1e-12)
    np.testing.assert_allclose(m.qpos0, [0, 0, 1, 1, 0, 0, 0, 0])


def test_ik_config_edge_cases():
    m, cfg, pack = params.load_pack("smplx", "unitree_g1")
    d = cfg.to_dict()
    # a scale-table body without a table-1 entry -> KeyError (motion_retarget.py:241)
    d2 = json.loads(json.dumps(d)); d2["ik_match_table1"]["torso_link"][1] = 0; d2["ik_match_table1"]["torso_link"][2] = 0
    with pytest.raises(KeyError):
        compile_task_table(m, IKConfig.from_dict(d2))
    # two robot frames bound to one human body
    d3 = json.loads(json.dumps(d)); d3["ik_match_table1"]["left_knee_link"][0] = "left_hip"
    with pytest.raises((ValueError, KeyError)):
        compile_task_table(m, IKConfig.from_dict(d3))
    # unknown robot frame
    d4 = json.loads(json.dumps(d)); d4["ik_match_table1"]["not_a_body"] = d4["ik_match_table1"].pop("torso_link")
    with pytest.raises(KeyError):
        compile_task_table(m, IKConfig.from_dict(d4))
    # zero-weight entry in table 2 only drops the stage-2 task
    d5 = json.loads(json.dumps(d)); d5["ik_match_table2"]["torso_link"][1] = 0; d5["ik_match_table2"]["torso_link"][2] = 0
    tt = compile_task_table(m, IKConfig.from_dict(d5))
    assert tt.in1.all() and tt.in2.sum() == tt.nt - 1
    # table-2 offsets are dead data (motion_retarget.py:121): the compiled offsets are table 1's
    k = list(cfg.ik_match_table1).index("left_toe_link")
    h = tt.task_human[k]
    np.testing.assert_allclose(tt.pos_off[h], [0.0, 0.02, 0.0])
