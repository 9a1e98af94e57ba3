#!/bin/bash
# Scheduler settings across workloads the defaults were NOT tuned on (DESIGN.md "Clip scheduling"): every line is one process.
# Usage: tools/prof/mix_table.sh > gpurun_out/mix_table.jsonl
run() { env "$@" python tools/prof/mix_case.py f64 2>/dev/null | tail -1; }
for mix in "PROBE_ROBOT=unitree_g1 PROBE_SRC=smplx PROBE_C=4096 PROBE_T=300" \
           "PROBE_ROBOT=unitree_g1 PROBE_SRC=smplx PROBE_C=8192 PROBE_T=300" \
           "PROBE_ROBOT=unitree_g1 PROBE_SRC=smplx PROBE_C=4096 PROBE_T=60" \
           "PROBE_ROBOT=unitree_g1 PROBE_SRC=smplx PROBE_C=4096 PROBE_T=300 PROBE_STRESS=1" \
           "PROBE_ROBOT=booster_t1 PROBE_SRC=bvh PROBE_C=8192 PROBE_T=300" \
           "PROBE_ROBOT=hightorque_hi PROBE_SRC=smplx PROBE_C=8192 PROBE_T=300" \
           "PROBE_ROBOT=kuavo_s45 PROBE_SRC=smplx PROBE_C=4096 PROBE_T=300"; do
  for knobs in ${MIX_KNOBS:-"GMR_NOTE=default" "GMR_PARTITION=0" "GMR_PARTITION=6 GMR_PARTITION_PCT=60" "GMR_PARTITION=7 GMR_PARTITION_PCT=55" "GMR_PARTITION=10" "GMR_SEGMENT=50"}; do
    run $mix $knobs
  done
done
