#!/usr/bin/env python
"""Compile every registered (source, robot) pair of a reference checkout into a model pack
(general_motion_retargeting_b200/data/packs/<src>_to_<robot>.json).

Usage: python tools/compile_packs.py [--root /root/reference]
The packs are derived data (flat kinematic tree + parsed IK tables); they are committed so
that the solver, tests and bench run where the reference checkout is absent (the GPU box).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from general_motion_retargeting_b200 import params  # noqa: E402


def _compact(obj, indent=0):
    """JSON with one line per leaf list (keeps the packs diff-able but small)."""
    return json.dumps(obj, separators=(",", ":"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--root", default=None)
    args = ap.parse_args()
    os.makedirs(params.PACK_ROOT, exist_ok=True)
    for src, robots in params.IK_CONFIG_REL.items():
        for robot in robots:
            pack = params.compile_pack(src, robot, args.root)
            out = params.pack_path(src, robot)
            with open(out, "w") as f:
                f.write(_compact(pack))
                f.write("\n")
            r = pack["robot"]
            print(f"{out.name}: nbody={len(r['body_names'])} nhinge={len(r['hinge_names'])} "
                  f"limited={sum(r['hinge_limited'])} dt={r['timestep']}")


if __name__ == "__main__":
    main()
