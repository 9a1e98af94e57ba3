#!/usr/bin/env python
"""BASELINE.json configs[4] at one GPU's share: 8192 mixed-robot clips (clip i -> robot i mod 5 of
g1/t1/toddy/n1/pm01, smplx mapping) x 300 frames, bucketed by robot on the host, one launch sequence per bucket.
Also configs[2] (booster_t1 / bvh, 8192 clips) and configs[3] (hightorque_hi / smplx, 8192 clips).  float64 kernel.
One JSON line per configuration; parity of a slice of every robot against the float64 oracle."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from general_motion_retargeting_b200 import GeneralMotionRetargeting, params
from general_motion_retargeting_b200.ik_config import compile_task_table
from general_motion_retargeting_b200.sharding import bucket_by_robot
from general_motion_retargeting_b200.synthetic import make_clips
from oracle import native

T = int(os.environ.get("SWEEP_T", "300"))

def prepare(src, robot, ids):
    m, cfg, _ = params.load_pack(src, robot)
    tt = compile_task_table(m, cfg)
    clips = make_clips(m, tt, ids, T=T, src_human=src, device="cuda")
    g = GeneralMotionRetargeting(src, robot, device=0)
    dev = [torch.from_numpy(x).cuda() for x in (clips.pos, clips.quat, clips.heights)]
    return m, tt, clips, g, dev

def run(label, buckets):
    """buckets: list of (src, robot, clip ids)"""
    prep = [prepare(s, r, ids) for s, r, ids in buckets]
    outs = None
    for _ in range(2):
        outs = [g.retarget_batch(*dev, return_info=True) for (_, _, _, g, dev) in prep]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    n = 3
    # one stream per robot bucket: the buckets' tails overlap with the other buckets' bulk instead of adding up
    streams = [torch.cuda.Stream() for _ in prep] if len(prep) > 1 else [torch.cuda.current_stream()]
    for _ in range(n):
        for st, (_, _, _, g, dev) in zip(streams, prep):
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                g.retarget_batch(*dev)
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    frames = sum(len(ids) for _, _, ids in buckets) * T
    par = {}
    for (m, tt, clips, g, dev), (q, it, err), (src, robot, ids) in zip(prep, outs, buckets):
        S = 48
        q_ref, it_ref, _ = native.retarget_batch(m, tt, clips.pos[:S], clips.quat[:S], clips.ratio(tt)[:S])
        same = (it[:S].cpu().numpy() == it_ref).all(-1)
        dq = np.abs(q[:S].double().cpu().numpy() - q_ref).max(-1)
        par[robot] = {"clips": len(ids), "solves_per_frame": float(it.sum().item() / (len(ids) * T)),
                      "iteration_agreement": float(same.mean()), "max_abs_dqpos": float(dq.max())}
    print(json.dumps({"config": label, "clips": frames // T, "frames": T, "ms": ms, "frames_per_s": frames / ms * 1e3, "precision": "f64",
                      "per_robot": par}), flush=True)

def run_multi(label, buckets):
    """the same buckets through ONE gmr_retarget_multi launch"""
    from general_motion_retargeting_b200 import retarget_mixed
    prep = [prepare(s, r, ids) for s, r, ids in buckets]
    bk = [(g, dev[0], dev[1], dev[2]) for (_, _, _, g, dev) in prep]
    for _ in range(2):
        outs, its = retarget_mixed(bk, return_info=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    n = 3
    for _ in range(n):
        retarget_mixed(bk)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    frames = sum(len(ids) for _, _, ids in buckets) * T
    par = {}
    for (m, tt, clips, g, dev), q, it, (src, robot, ids) in zip(prep, outs, its, buckets):
        S = 48
        q_ref, it_ref, _ = native.retarget_batch(m, tt, clips.pos[:S], clips.quat[:S], clips.ratio(tt)[:S])
        same = (it[:S].cpu().numpy() == it_ref).all(-1)
        par[robot] = {"clips": len(ids), "iteration_agreement": float(same.mean()), "max_abs_dqpos": float(np.abs(q[:S].double().cpu().numpy() - q_ref).max())}
    print(json.dumps({"config": label, "clips": frames // T, "frames": T, "ms": ms, "frames_per_s": frames / ms * 1e3, "precision": "f64", "per_robot": par}), flush=True)


robots5 = ["unitree_g1", "booster_t1", "stanford_toddy", "fourier_n1", "engineai_pm01"]
b = bucket_by_robot([robots5[i % 5] for i in range(8192)])
run("configs[4] share of one GPU: 8192 mixed-robot clips (g1/t1/toddy/n1/pm01, smplx), bucketed by robot", [("smplx", r, b[r]) for r in robots5])
run_multi("configs[4] share of one GPU, ONE mixed-robot launch (gmr_retarget_multi)", [("smplx", r, b[r]) for r in robots5])
run("configs[2]: booster_t1, bvh mapping, 8192 clips", [("bvh", "booster_t1", list(range(8192)))])
run("configs[3]: hightorque_hi, smplx mapping, 8192 clips", [("smplx", "hightorque_hi", list(range(8192)))])
