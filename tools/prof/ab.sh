#!/bin/bash
# Same-box A/B of library builds on the benchmark mix: tools/prof/ab.sh libA.so libB.so ...   (files under csrc/; each is
# measured twice, interleaved, PROBE_T=300 unless set)
export PROBE_T=${PROBE_T:-300}
for rep in 1 2; do
  for lib in "$@"; do
    echo "== $lib $(GMR_B200_LIB=$PWD/general_motion_retargeting_b200/csrc/$lib python tools/prof/mix_case.py ${AB_PREC:-f64 f32} 2>&1 | tail -1)"
  done
done
