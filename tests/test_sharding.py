"""Multi-rank path on CPU: contiguous clip shards + gloo (world_size 2) for the hand-over."""
import os
import subprocess
import sys

import numpy as np
import pytest

from general_motion_retargeting_b200.sharding import all_shards, clip_shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shards_partition_the_clips():
    for C in (0, 1, 7, 4096, 65536, 65537):
        for W in (1, 2, 3, 4, 8):
            sh = all_shards(C, W)
            assert sh[0][0] == 0 and sh[-1][1] == C
            assert all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
            sizes = [e - b for b, e in sh]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        clip_shard(10, 2, 2)


def test_bucket_by_robot_groups_and_preserves_order():
    from general_motion_retargeting_b200.sharding import bucket_by_robot
    robots = ["g1", "t1", "toddy", "n1", "pm01"]
    b = bucket_by_robot([robots[i % 5] for i in range(23)])
    assert sorted(b) == sorted(robots) and sum(len(v) for v in b.values()) == 23
    assert b["g1"] == [0, 5, 10, 15, 20] and b["pm01"] == [4, 9, 14, 19]


def test_lpt_shard_balances_counts_and_hard_clips():
    from general_motion_retargeting_b200.sharding import hardness_proxy, lpt_shard
    rng = np.random.default_rng(3)
    for C, W in ((4096, 8), (65536, 4), (1001, 3), (5, 8), (0, 2)):
        h = rng.uniform(0, np.pi, C)
        sh = lpt_shard(h, W)
        allidx = np.concatenate(sh) if C else np.zeros(0, int)
        assert sorted(allidx.tolist()) == list(range(C))                       # a partition
        sizes = [len(x) for x in sh]
        assert max(sizes) - min(sizes) <= 1
        if C >= 1000:
            hard = [int((h[x] > 2.3).sum()) for x in sh]                       # the library's own "hard clip" band
            assert max(hard) - min(hard) <= 1
            assert max(h[x].sum() for x in sh) - min(h[x].sum() for x in sh) < 0.01 * h.sum() / W + np.pi
    # deterministic, ties by index
    assert [x.tolist() for x in lpt_shard([1.0, 1.0, 1.0, 1.0], 2)] == [[0, 3], [1, 2]]
    # the proxy is the rotation angle between root target and initial orientation
    q = np.zeros((3, 2, 4)); q[:, :, 0] = 1.0
    q[1, 1] = [np.cos(0.5), 0, 0, np.sin(0.5)]; q[2, 1] = [-np.cos(1.2), 0, -np.sin(1.2), 0]
    ang = hardness_proxy(q, 1, [1, 0, 0, 0], [1, 0, 0, 0])
    assert np.allclose(ang, [0.0, 1.0, 2.4], atol=1e-12)


WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["GMR_ROOT"]); sys.path.insert(0, os.path.join(os.environ["GMR_ROOT"], "tests"))
import numpy as np, torch, torch.distributed as dist
from helpers import problem, emu_retarget_batch
from general_motion_retargeting_b200.synthetic import make_clips
from general_motion_retargeting_b200.sharding import clip_shard, gather_clips, max_over_ranks
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
m, tt, _ = problem("smplx", "unitree_g1")
C, T = 5, 6
b, e = clip_shard(C, rank, world)
clips = make_clips(m, tt, range(b, e), T=T)
q, it, err, tg, _ = emu_retarget_batch(m, tt, clips.pos, clips.quat, clips.ratio(tt), bits=64)
full = gather_clips(torch.from_numpy(q), C)
t = max_over_ranks(float(rank + 1))
if rank == 0:
    np.save(os.environ["GMR_OUT"], full.numpy())
    assert t == float(world)
dist.destroy_process_group()
'''


def test_two_ranks_over_gloo_equal_one_process(built, tmp_path):
    from helpers import emu_retarget_batch, problem
    from general_motion_retargeting_b200.synthetic import make_clips
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = tmp_path / "q.npy"
    env = dict(os.environ, GMR_ROOT=ROOT, GMR_OUT=str(out), OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    m, tt, _ = problem("smplx", "unitree_g1")
    clips = make_clips(m, tt, range(5), T=6)
    q, *_ = emu_retarget_batch(m, tt, clips.pos, clips.quat, clips.ratio(tt), bits=64)
    np.testing.assert_array_equal(np.load(out), q)


def test_dataset_plan_deals_every_bucket_over_the_devices():
    """dataset.plan_shards: the host logic of the one-process / all-GPUs entry (no GPU needed)."""
    from helpers import problem
    from general_motion_retargeting_b200.dataset import plan_shards
    from general_motion_retargeting_b200.synthetic import make_clips
    jobs = []
    for src, robot, n in (("smplx", "unitree_g1", 41), ("bvh", "booster_t1", 7), ("smplx", "engineai_pm01", 0)):
        m, tt, _ = problem(src, robot)
        c = make_clips(m, tt, range(n), T=2, src_human=src)
        jobs.append((src, robot, c.pos, c.quat, c.heights))
    for W in (1, 2, 8):
        for shard in ("lpt", "contiguous"):
            plans = plan_shards(jobs, W, shard)
            assert len(plans) == 3 and all(len(p) == W for p in plans)
            for p, (_, _, pos, _, _) in zip(plans, jobs):
                assert sorted(np.concatenate(p).tolist()) == list(range(pos.shape[0]))
                assert max(len(x) for x in p) - min(len(x) for x in p) <= 1
    with pytest.raises(ValueError):
        plan_shards(jobs, 2, "random")
