#!/bin/bash
# Developer loop: rebuild the CUDA library (register / spill summary), the lane-serial emulator, and run the emulator tests.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
(cd "$ROOT/general_motion_retargeting_b200/csrc" && bash build.sh 2>&1 | grep -E "error|gmr_retarget_kernelI|spill|Used" | grep -E -A2 "gmr_retarget_kernelI(dd|df|ff)Li" | grep -E "error|spill|Used" || true)
g++ -O2 -std=c++17 -fPIC -shared -pthread -Wno-unknown-pragmas -o "$ROOT/tests/emu/libgmr_emu.so" "$ROOT/tests/emu/gmr_emu.cpp"
cd "$ROOT" && timeout 1200 python -m pytest tests/test_emulator.py -x -q 2>&1 | tail -3
