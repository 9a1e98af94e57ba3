#!/usr/bin/env python
"""Headline benchmark: retargeted frames/s (Unitree G1, SMPL-X mapping) — BASELINE.json.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--precision f32|f64]

One "step" = one pass of the hot path (the batched two-stage IK of
GeneralMotionRetargeting.retarget, reference motion_retarget.py:139-185) over one batch of
synthetic clips: BASELINE.json configs[1], 4096 clips x 300 frames of unitree_g1 per GPU
(weak scaling: every rank solves its own 4096 clips, no collective on the solve path).
Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "retargeted frames/sec (G1 29-DoF)"       # BASELINE.json's metric, quoted on configs[1]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("GMR_BENCH_PRECISION", "f64"), choices=["f32", "f64"])
    ap.add_argument("--clips", type=int, default=4096, help="clips per GPU")
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--robot", default="unitree_g1")
    ap.add_argument("--src", default="smplx")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-f32", action="store_true", help="skip the float32 fast-mode side measurement")
    return ap.parse_args()


# ---- algorithmic work (dense convention; the figure roofline.achieved uses) ----
def flop_model(robot, table):
    nv, nb = robot.nv, robot.nbody

    def per_stage(mask, use):
        if not use:
            return None
        nt = int(mask.sum())
        m = 6 * nt
        chain = sum(6 + len(robot.chain_hinges(int(b))) for b, on in zip(table.task_body, mask) if on)
        f_solve = m * nv * (nv + 1) + 2 * m * nv + nv ** 3 / 3.0 + 2 * nv * nv + 105 * chain + 400 * nt + 90 * nb
        f_err = 90 * nb + 150 * nt
        return f_solve, f_err

    return per_stage(table.in1, table.use1), per_stage(table.in2, table.use2)


def flops_of_run(robot, table, iters):
    """Σ_frames n_solve·F_solve + (n_solve + 1)·F_err per enabled stage; iters [C,T,2]."""
    s1, s2 = flop_model(robot, table)
    total = 0.0
    nframes = iters.shape[0] * iters.shape[1]
    for k, s in enumerate((s1, s2)):
        if s is None:
            continue
        n = float(iters[..., k].sum())
        total += n * s[0] + (n + nframes) * s[1]
    return total


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def load_problem(args):
    from general_motion_retargeting_b200 import params
    from general_motion_retargeting_b200.ik_config import compile_task_table
    robot, cfg, _ = params.load_pack(args.src, args.robot)
    return robot, compile_task_table(robot, cfg)


def workload_config(args, table):
    """The `config` object of BOTH arms (identical by construction: the driver compares them)."""
    C, T = args.clips, args.frames
    in_mb = C * T * table.nh * 7 * 4 / 1e6
    return {"workload": f"{args.robot} ({args.src} mapping), {C} synthetic clips x {T} frames per GPU, "
                        "sequential two-stage IK warm-started per clip (BASELINE.json configs[1])",
            "clips_per_gpu": C, "frames": T, "sharding": "clips, contiguous ranges per rank, no collective",
            "l2": f"inputs ~{in_mb:.0f} MB per step {'>' if in_mb > 126 else '<= (reduced-size run: L2 not flushed)'} 126 MB L2"}


def cpu_reference_run(robot, table, clips, nthreads=0, src=None, robot_name=None):
    """The reference's CPU path for this workload on all host threads.  With mink/mujoco/daqp and the reference
    package importable (oracle/mink_adapter.py, e.g. from baseline/_ref/) the UNMODIFIED reference loop runs, one clip
    per worker process (kind = "reference"); otherwise the float64 C++ oracle port stands in (kind = "port").
    Returns (qpos, iters, seconds, kind, note)."""
    from oracle import mink_adapter
    ok, why = mink_adapter.available()
    if ok and src is not None:
        q, it, _, dt = mink_adapter.retarget_batch(src, robot_name, table.human_names, clips.pos, clips.quat, clips.heights)
        return q, it, dt, "reference", "unmodified reference loop (" + why + ")"
    from oracle import native
    t0 = time.perf_counter()
    q, it, err = native.retarget_batch(robot, table, clips.pos, clips.quat, clips.ratio(table), nthreads=nthreads)
    dt = time.perf_counter() - t0
    return q, it, dt, "port", "float64 C++ oracle port (real reference unavailable: " + why + ")"


def main_reference(args):
    """CPU arm on the SAME workload as the GPU arm (same generator, robot, clip ids 0..C-1, T).  One step = one pass over a
    bounded, rotating sample of those clips (step k takes the next `sample` clip ids, wrapping around), sized from a
    short probe so that warm-up + K steps take about two minutes whatever the host and whichever implementation runs."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from general_motion_retargeting_b200.synthetic import make_clips
    robot, table = load_problem(args)
    cores = os.cpu_count() or 1
    C, T = args.clips, args.frames
    sub = lambda b, ids: type(b)(pos=b.pos[ids], quat=b.quat[ids], heights=b.heights[ids], qpos_gen=b.qpos_gen[ids])
    run = lambda b: cpu_reference_run(robot, table, b, src=args.src, robot_name=args.robot)
    # probe: the first 2 clips per core (at most 64) give the rate of whichever implementation is present
    n_probe = min(C, max(2 * cores, 8), 64)
    probe = make_clips(robot, table, range(n_probe), T=T, src_human=args.src)
    _, _, dt, kind, note = run(probe)
    rate = n_probe * T / dt
    budget_s = float(os.environ.get("GMR_REF_BUDGET_S", "120"))
    passes = max(args.steps + max(args.warmup, 0), 1)
    sample = int(min(C, max(cores, rate * budget_s / passes / T)))
    n_gen = min(C, sample * passes)                       # only the clips some pass will touch are generated
    clips = probe if n_gen <= n_probe else make_clips(robot, table, range(n_gen), T=T, src_human=args.src)
    n_gen = clips.pos.shape[0]
    ids_of = lambda k: [(k * sample + i) % n_gen for i in range(sample)]
    k = 0
    for _ in range(max(args.warmup, 0)):
        run(sub(clips, ids_of(k))); k += 1
    times = []
    for _ in range(args.steps):
        _, _, dt, _, _ = run(sub(clips, ids_of(k))); k += 1
        times.append(dt)
    tot = sum(times)
    value = sample * T * args.steps / tot
    line = {
        "metric": METRIC if args.robot == "unitree_g1" else f"retargeted frames/sec ({args.robot})", "value": value, "unit": "frames/s", "impl": "reference",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, table),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": f"each step = {sample} of the workload's {C} clips x {T} frames (rotating window over clip ids "
                                   f"0..{n_gen - 1}), {args.steps} steps; {note}; {cores} host threads"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "published_reference_fps_single_process": "35-70 (README.md:215-221, desktop CPUs)",
    }
    print(json.dumps(line))


def main_ours(args):
    import torch
    import torch.distributed as dist

    from general_motion_retargeting_b200 import GeneralMotionRetargeting, _native
    from general_motion_retargeting_b200.synthetic import make_clips

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its "NCCL version ..." banner on stdout when the communicator is created; stdout must carry
        # the one JSON line only, so fd 1 points at stderr while the communicator comes up
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    robot, table = load_problem(args)
    C, T = args.clips, args.frames
    # rank r owns clips [r*C, (r+1)*C): contiguous clip ranges, no data-path collective
    clips = make_clips(robot, table, range(rank * C, (rank + 1) * C), T=T, src_human=args.src, device=str(dev))
    gmr = GeneralMotionRetargeting(args.src, args.robot, device=local)
    lib = _native.load_library()

    d_pos = torch.from_numpy(clips.pos).to(dev)
    d_quat = torch.from_numpy(clips.quat).to(dev)
    d_h = torch.from_numpy(clips.heights).to(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def maxreduce(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # one untimed run with iteration counts: algorithmic flops of the workload + parity sample
    q_info, it_info, _ = gmr.retarget_batch(d_pos, d_quat, d_h, return_info=True, precision=args.precision)
    torch.cuda.synchronize(dev)
    iters = it_info.cpu().numpy()
    flops_step = flops_of_run(robot, table, iters)
    q_gpu_sample = q_info[: min(C, 4096)].double().cpu().numpy()
    out_elem_bytes = q_info.element_size()      # float64 qpos from the float64 kernel, float32 from the float32 one
    del q_info, it_info

    for _ in range(args.warmup):
        gmr.retarget_batch(d_pos, d_quat, d_h, precision=args.precision)
    barrier()

    # ---- device-resident timing: K steps, CUDA events on the launch stream ----------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.gmr_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for k in range(args.steps):
        gmr.retarget_batch(d_pos, d_quat, d_h, precision=args.precision)
        ev[k + 1].record()
    barrier()
    launches = lib.gmr_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = maxreduce(ev[0].elapsed_time(ev[-1]))
    kernel_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    kernel_ms_avg = float(np.mean(kernel_ms))
    frames_step_all = world * C * T
    value = frames_step_all * args.steps / (total_ms * 1e-3)

    # ---- float32 fast mode, same workload (reported beside the headline, never instead of it) -----
    f32_side = None
    if args.precision == "f64" and not args.no_f32:
        q32, it32, _ = gmr.retarget_batch(d_pos, d_quat, d_h, return_info=True, precision="f32")
        for _ in range(2):
            gmr.retarget_batch(d_pos, d_quat, d_h, precision="f32")
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            gmr.retarget_batch(d_pos, d_quat, d_h, precision="f32")
        e1.record()
        barrier()
        ms32 = maxreduce(e0.elapsed_time(e1))
        it32 = it32.cpu().numpy()
        same = (it32 == iters).all(-1)
        dq = (q32[: q_gpu_sample.shape[0]].double().cpu().numpy() - q_gpu_sample)
        dq = np.abs(dq).max(-1)
        prefix = np.logical_and.accumulate(same[: dq.shape[0]], axis=1)
        f32_side = {"value": frames_step_all * args.steps / (ms32 * 1e-3), "unit": "frames/s", "ms_per_step": ms32 / args.steps,
                    "vs_f64_kernel": {"iteration_count_agreement": float(same.mean()),
                                      "max_abs_dqpos_identical_history": float(dq[prefix].max()),
                                      "max_abs_dqpos_all": float(dq.max()), "p999_abs_dqpos": float(np.quantile(dq, 0.999))}}
        del q32

    # ---- end to end through the public API with HOST (pinned) buffers ----------------------
    e2e = None
    if not args.no_e2e:
        p_pos = torch.from_numpy(clips.pos).pin_memory()
        p_quat = torch.from_numpy(clips.quat).pin_memory()
        p_out = torch.empty((C, T, robot.nq), dtype=torch.float32).pin_memory()
        n_pos, n_quat, n_out = p_pos.numpy(), p_quat.numpy(), p_out.numpy()
        for _ in range(max(1, min(args.warmup, 2))):
            gmr.retarget_batch(n_pos, n_quat, clips.heights, out=n_out, precision=args.precision)
        n_out.fill(np.nan)                    # the timed steps must produce every value they are credited with
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            gmr.retarget_batch(n_pos, n_quat, clips.heights, out=n_out, precision=args.precision)
        torch.cuda.synchronize(dev)
        e2e_s = maxreduce(time.perf_counter() - t0)
        e2e = {"value": frames_step_all * args.steps / e2e_s, "unit": "frames/s",
               "h2d_bytes_per_step": int(n_pos.nbytes + n_quat.nbytes + clips.heights.nbytes),
               "d2h_bytes_per_step": int(n_out.nbytes), "ms_per_step": 1e3 * e2e_s / args.steps,
               "api": "GeneralMotionRetargeting.retarget_batch(numpy arrays in pinned host memory, float32) -> "
                      "gmr_retarget_batch_host_ex; the kernel reads the keypoints over the host link (TMA bulk copies from "
                      "the mapped arrays, one frame ahead) and writes qpos into the caller's array: the bytes below cross "
                      "the link inside the timed region, every step; wall clock around the calls, max over ranks"}
        e2e_check = float(np.abs(n_out[: q_gpu_sample.shape[0]].astype(np.float64) - q_gpu_sample).max())
    else:
        e2e_check = None

    # ---- CPU baseline + parity on a bounded sample (rank 0, N = 1 only) ----------------------
    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = min(C, max(64 * cores, 64))      # ~10 s of CPU work on the box's host cores
        sub = type(clips)(pos=clips.pos[:sample], quat=clips.quat[:sample], heights=clips.heights[:sample],
                          qpos_gen=clips.qpos_gen[:sample])
        q_ref, it_ref, dt, kind, note = cpu_reference_run(robot, table, sub)      # the port: the parity sample needs seconds, not hours
        cpu_baseline = {"value": sample * T / dt, "unit": "frames/s", "cores": cores, "kind": kind,
                        "sample": f"first {sample} clips x {T} frames of the same workload, {note}, "
                                  f"{cores} host threads, {dt:.1f} s"}
        same = (iters[:sample] == it_ref).all(-1)
        dq = np.abs(q_gpu_sample[:sample] - q_ref).max(-1)
        prefix = np.logical_and.accumulate(same, axis=1)
        parity = {"frames": int(same.size), "iteration_count_agreement": float(same.mean()),
                  "max_abs_dqpos_all": float(dq.max()),
                  "max_abs_dqpos_identical_history": float(dq[prefix].max()) if prefix.any() else None,
                  "p999_abs_dqpos": float(np.quantile(dq, 0.999)), "vs": "float64 CPU oracle (oracle/gmr_oracle.cpp)"}
        # the real thing, when it can run here (oracle/mink_adapter.py): one clip per core, first 60 frames
        from oracle import mink_adapter
        ok, why = mink_adapter.available()
        if ok:
            n, tt = min(cores, C, 16), min(T, 60)
            q_m, it_m, _, dt_m = mink_adapter.retarget_batch(args.src, args.robot, table.human_names, clips.pos[:n, :tt],
                                                            clips.quat[:n, :tt], clips.heights[:n])
            same = (iters[:n, :tt] == it_m).all(-1)
            dq = np.abs(q_gpu_sample[:n, :tt] - q_m).max(-1)
            prefix = np.logical_and.accumulate(same, axis=1)
            parity["vs_unmodified_reference"] = {
                "frames": int(same.size), "iteration_count_agreement": float(same.mean()), "max_abs_dqpos_all": float(dq.max()),
                "max_abs_dqpos_identical_history": float(dq[prefix].max()) if prefix.any() else None,
                "frames_per_s": n * tt / dt_m, "what": why}
        else:
            parity["vs_unmodified_reference"] = {"unavailable": why}

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        sm_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
        sm_obs = (clocks or {}).get("sm_mhz") or sm_max
        nsm = torch.cuda.get_device_properties(dev).multi_processor_count
        lanes = 128 if args.precision == "f32" else 64
        peak_tf = nsm * lanes * 2 * sm_max * 1e6 / 1e12
        achieved_tf = flops_step / (kernel_ms_avg * 1e-3) / 1e12
        s1, s2 = flop_model(robot, table)
        bytes_step = C * T * (table.nh * 7 * 4 + robot.nq * out_elem_bytes)
        # DRAM bytes of one launch of this exact configuration from an `ncu --set full` capture
        # (profiles/traffic.json, written by tools/prof/summarize.py); None when no capture matches
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                for rec in json.load(f):
                    if (rec["robot"], rec["src"], rec["clips"], rec["frames"], rec["precision"]) == \
                            (args.robot, args.src, C, T, args.precision):
                        traffic, traffic_src = rec["dram_bytes_per_launch"], rec["source"]
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        line = {
            "metric": METRIC if args.robot == "unitree_g1" else f"retargeted frames/sec ({args.robot})", "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": workload_config(args, table),
            "solves_per_frame": float(iters.sum() / (C * T)),
            "gpu_launches": int(launches),
            "clocks": clocks,
            "e2e": e2e,
            "roofline": {"bound": "fp32_fma" if args.precision == "f32" else "fp64_fma",
                         "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "peak_basis": f"{nsm} SMs x {lanes} lanes x 2 x {sm_max:.0f} MHz (nominal at max clock; "
                                       f"median clock observed under load {sm_obs:.0f} MHz)",
                         "flops_per_step": flops_step, "kernel_ms": kernel_ms_avg,
                         "F_solve_F_err_stage1": s1, "F_solve_F_err_stage2": s2,
                         "hbm": {"algorithmic_bytes_per_step": bytes_step,
                                 "achieved_gbs": bytes_step / (kernel_ms_avg * 1e-3) / 1e9,
                                 "peak_gbs": hbm_peak, "frac": bytes_step / (kernel_ms_avg * 1e-3) / 1e9 / hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}},
            "cpu_baseline": cpu_baseline,
            "parity": parity,
            "e2e_vs_device_max_abs_diff": e2e_check,
            "f32_fast_mode": f32_side,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
