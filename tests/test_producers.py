"""Human-frame producers: the CPU restatement against outputs of the reference's own
loader code (tests/golden/reference_producers.npz, made by tools/make_golden_producers.py), and — on the GPU —
the CUDA kernels against both."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_producers.npz")


def quat_close(a, b, atol):
    """equal up to the sign of each quaternion"""
    d = np.minimum(np.abs(a - b).max(-1), np.abs(a + b).max(-1))
    assert d.max() < atol, d.max()


def bvh_case(g):
    bones = list(g["bvh.bones"])
    names = ["Hips", "Spine2", "LeftUpLeg", "RightUpLeg", "LeftLeg", "RightLeg", "LeftFootMod", "RightFootMod",
             "LeftArm", "RightArm", "LeftForeArm", "RightForeArm", "LeftHand", "RightHand"]          # bvh_to_g1.json
    pos_joint = [bones.index({"LeftFootMod": "LeftFoot", "RightFootMod": "RightFoot"}.get(n, n)) for n in names]
    rot_joint = [bones.index({"LeftFootMod": "LeftToe", "RightFootMod": "RightToe"}.get(n, n)) for n in names]
    out_names = list(g["bvh.out_names"])
    sel = [out_names.index(n) for n in names]
    return names, pos_joint, rot_joint, g["bvh.out_pos"][:, sel], g["bvh.out_quat"][:, sel]


def test_oracle_bvh_producer_matches_reference():
    from oracle import producers_oracle as P
    g = np.load(GOLD)
    names, pj, rj, ref_pos, ref_quat = bvh_case(g)
    pos, quat = P.bvh_frames(g["bvh.quats"], g["bvh.pos"], g["bvh.parents"], pj, rj)
    np.testing.assert_allclose(pos, ref_pos, atol=1e-9)
    quat_close(quat, ref_quat, 1e-9)


@pytest.mark.parametrize("tag", ["smplx120", "smplx30"])
def test_oracle_smplx_producer_matches_reference(tag):
    from oracle import producers_oracle as P
    g = np.load(GOLD)
    body_joint = list(range(22))
    pos, quat, afps = P.smplx_frames(g[f"{tag}.global_orient"], g[f"{tag}.full_pose"], g[f"{tag}.joints"], g[f"{tag}.parents"],
                                     float(g[f"{tag}.src_fps"]), 30, body_joint)
    assert afps == pytest.approx(float(g[f"{tag}.aligned_fps"]))
    assert pos.shape == g[f"{tag}.out_pos"][:, :22].shape
    np.testing.assert_allclose(pos, g[f"{tag}.out_pos"][:, :22], atol=1e-6)          # reference joints are float32
    quat_close(quat, g[f"{tag}.out_quat"][:, :22], 2e-6)


@pytest.mark.gpu
def test_gpu_bvh_producer_matches_reference_and_feeds_the_solver():
    import torch
    from general_motion_retargeting_b200 import GeneralMotionRetargeting, producers
    g = np.load(GOLD)
    names, pj, rj, ref_pos, ref_quat = bvh_case(g)
    gmr = GeneralMotionRetargeting("bvh", "unitree_g1", actual_human_height=producers.BVH_HUMAN_HEIGHT, device=0)
    assert gmr.human_body_names == names
    pos, quat = producers.bvh_frames(g["bvh.quats"], g["bvh.pos"], g["bvh.parents"], list(g["bvh.bones"]), gmr.human_body_names)
    np.testing.assert_allclose(pos.cpu().numpy(), ref_pos, atol=2e-5)            # float32 kernel, metres
    quat_close(quat.cpu().numpy(), ref_quat, 5e-6)
    # the producer's output is the batched entry's input: same qpos as the per-frame dict API of the reference
    F = pos.shape[0]
    q_batch = gmr.retarget_batch(pos[None], quat[None], torch.full((1,), producers.BVH_HUMAN_HEIGHT, device="cuda")).cpu().numpy()[0]
    for t in range(F):
        frame = {n: (pos[t, i].cpu().numpy(), quat[t, i].cpu().numpy()) for i, n in enumerate(names)}
        np.testing.assert_allclose(gmr.retarget(frame), q_batch[t], atol=1e-12)
    with pytest.raises(KeyError):
        producers.bvh_frames(g["bvh.quats"], g["bvh.pos"], g["bvh.parents"], list(g["bvh.bones"]), ["NoSuchBone"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["smplx120", "smplx30"])
def test_gpu_smplx_producer_matches_reference(tag):
    from general_motion_retargeting_b200 import GeneralMotionRetargeting, producers
    g = np.load(GOLD)
    gmr = GeneralMotionRetargeting("smplx", "unitree_g1", device=0)
    names = list(g[f"{tag}.names"])
    sel = [names.index(n) for n in gmr.human_body_names]
    pos, quat, afps = producers.smplx_frames(g[f"{tag}.global_orient"], g[f"{tag}.full_pose"], g[f"{tag}.joints"], g[f"{tag}.parents"],
                                             names, gmr.human_body_names, float(g[f"{tag}.src_fps"]), 30.0)
    assert afps == pytest.approx(float(g[f"{tag}.aligned_fps"]))
    np.testing.assert_allclose(pos.cpu().numpy(), g[f"{tag}.out_pos"][:, sel], atol=2e-6)
    quat_close(quat.cpu().numpy(), g[f"{tag}.out_quat"][:, sel], 2e-5)           # float32 chain of up to 6 rotations
    # hands need more than 32 joints on their chains only together with the rest of the body: a wrist-only request works,
    # asking for every joint reports the limit instead of computing something else
    with pytest.raises(RuntimeError):
        producers.smplx_frames(g[f"{tag}.global_orient"], g[f"{tag}.full_pose"], g[f"{tag}.joints"], g[f"{tag}.parents"],
                               names, names[:33], float(g[f"{tag}.src_fps"]), 30.0)
