#!/usr/bin/env python
"""Per-clip timeline of one batch (gmr_debug_trace): where the step's time goes between the slow chain and the
dense bulk.  PROBE_T / PROBE_C / precision as argv[1]."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from general_motion_retargeting_b200 import GeneralMotionRetargeting, params, _native
from general_motion_retargeting_b200.ik_config import compile_task_table
from general_motion_retargeting_b200.synthetic import make_clips
T = int(os.environ.get("PROBE_T", "300")); C = int(os.environ.get("PROBE_C", "4096")); prec = sys.argv[1] if len(sys.argv) > 1 else "f64"
robot, cfg, _ = params.load_pack("smplx", "unitree_g1")
table = compile_task_table(robot, cfg)
gmr = GeneralMotionRetargeting("smplx", "unitree_g1", device=0)
b = make_clips(robot, table, range(C), T=T, device="cuda", stress=os.environ.get("PROBE_STRESS", "0") == "1")
dp, dq, dh = (torch.from_numpy(x).cuda() for x in (b.pos, b.quat, b.heights))
lib = _native.load_library()
lib.gmr_debug_trace.argtypes = [ctypes.c_void_p]; lib.gmr_debug_trace.restype = None
for _ in range(2): gmr.retarget_batch(dp, dq, dh, precision=prec)
tr = torch.zeros((2 * C, 4), dtype=torch.int64, device="cuda")
lib.gmr_debug_trace(tr.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); _, its, _ = gmr.retarget_batch(dp, dq, dh, precision=prec, return_info=True); e1.record(); torch.cuda.synchronize()
its = its.sum(-1).cpu().numpy()
its_sum = np.concatenate([its[:, 0], its[:, 1:].sum(1)])          # solves of frame 0 | of the continuing launch
lib.gmr_debug_trace(None)
t = tr.cpu().numpy()
print(json.dumps({"ms": e0.elapsed_time(e1)}))
for name, rows in (("first launch (from frame 0)", t[:C]), ("continuing launch", t[C:])):
    ok = rows[:, 1] > 0
    if not ok.any(): continue
    r = rows[ok]; t0 = r[:, 0].min()
    st, en = (r[:, 0] - t0) / 1e6, (r[:, 1] - t0) / 1e6
    run_ms = r[:, 2] / 1e6; fac = r[:, 3] & 0xffffffff; nseg = (r[:, 3] >> 32) & 0xffff; sm = (r[:, 3] >> 48) & 0xffff
    solves = its_sum[:C][ok] if name.startswith("first") else its_sum[C:][ok]
    dur = en - st
    print(f"== {name}: clips {ok.sum()}, span {en.max():.2f} ms, solves {solves.sum()}, factorisations {fac.sum()} ({fac.sum() / max(solves.sum(), 1):.3f} per solve)")
    # SM classes by how many clips they served
    per_sm = np.bincount(sm, minlength=sm.max() + 1)
    print("   clips per SM: min %d median %d max %d" % (per_sm[per_sm > 0].min(), np.median(per_sm[per_sm > 0]), per_sm.max()))
    q = np.percentile(en, [50, 90, 99, 100]); print("   finish-time percentiles 50/90/99/100: " + " ".join(f"{x:.1f}" for x in q))
    top = np.argsort(-dur)[:8]
    print("   longest clips: " + ", ".join(f"{dur[i]:.1f}ms(running {run_ms[i]:.1f},{solves[i]}s,{fac[i]}f,{nseg[i]}seg,start {st[i]:.1f},sm{sm[i]})" for i in top))
    print(f"   waiting in the rings (span - running): median {np.median(dur - run_ms):.2f} ms, p99 {np.percentile(dur - run_ms, 99):.2f}, max {(dur - run_ms).max():.2f}")
    # busy warps over time
    edges = np.linspace(0, en.max(), 21)
    busy = [(np.minimum(en, edges[k + 1]) - np.maximum(st, edges[k])).clip(0).sum() / (edges[k + 1] - edges[k]) for k in range(20)]
    print("   busy warps per 5% slice: " + " ".join(f"{x:.0f}" for x in busy))
    us = run_ms * 1e3 / np.maximum(solves, 1)
    slow = fac > 1.2 * solves
    print(f"   us per solve: all median {np.median(us):.1f}; clips with >1.2 factorisations/solve: n={slow.sum()} median {np.median(us[slow]) if slow.any() else 0:.1f} us, finish median {np.median(en[slow]) if slow.any() else 0:.1f} max {en[slow].max() if slow.any() else 0:.1f}")
    if not name.startswith("first"):
        slowc = fac > 1.2 * solves
        print(f"   slow clips by the CTA of their last segment: CTA < 59: {(slowc & (sm < 59)).sum()}, CTA >= 59: {(slowc & (sm >= 59)).sum()}; segments: median {np.median(nseg[slowc]) if slowc.any() else 0}")
    late = np.argsort(-en)[:8]
    print("   last to finish: " + ", ".join(f"{en[i]:.1f}ms(start {st[i]:.1f},running {run_ms[i]:.1f},{solves[i]}s,sm{sm[i]})" for i in late))
np.save(os.environ.get("TRACE_OUT", "/tmp/trace.npy"), t)
