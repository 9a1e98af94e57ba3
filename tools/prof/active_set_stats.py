#!/usr/bin/env python
"""Active-set statistics of chosen clips from the lane-serial emulator built with -DGMR_STATS (CPU):
   g++ -O2 -std=c++17 -fPIC -shared -pthread -DGMR_STATS -o /tmp/libgmr_emu_stats.so tests/emu/gmr_emu.cpp"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers
from general_motion_retargeting_b200.synthetic import make_clips
helpers.EMU = "/tmp/libgmr_emu_stats.so"
m, tt, _ = helpers.problem("smplx", "unitree_g1")
ids = [int(x) for x in sys.argv[1:]] or [2762, 1672, 0]
T = int(os.environ.get("PROBE_T", "300"))
for cid in ids:
    clips = make_clips(m, tt, [cid], T=T)
    q, it, err, tg, refac = helpers.emu_retarget_batch(m, tt, clips.pos, clips.quat, clips.ratio(tt), bits=64)
    out = (C.c_longlong * 16)()
    helpers._emu.gmr_emu_stats(out, 1)
    s = list(out)
    print(f"clip {cid}: solves {s[0]}, factorisations {s[1]} ({s[1]/s[0]:.2f}/solve), blocked passes {s[2]} ({s[2]/s[0]:.2f}/solve), "
          f"CHECK passes {s[3]} ({s[3]/s[0]:.2f}/solve), checks that released {s[4]} ({s[4]/s[0]:.2f}/solve, {s[5]} bounds), "
          f"pinned at end: mean {s[6]/s[0]:.2f}, solves ending with pins {s[7]/s[0]:.1%}, warm set already optimal {s[8]/s[0]:.1%}")
