#!/bin/bash
# One GPU call of the kernel development loop: parity tests, same-box A/B of the benchmark mix against a base build,
# per-clip timelines of both, neighbouring scheduler settings, and ncu captures (balanced batch, lone slow clip).
# Usage: tools/prof/ab_full.sh <out-dir under gpurun_out> [base-lib under csrc/]
OUT=gpurun_out/$1; BASE=${2:-libgmr_base.so}
mkdir -p $OUT
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3 > $OUT/pytest.log
AB_PREC=f64 bash tools/prof/ab.sh $BASE libgmr_b200.so > $OUT/ab.log 2>&1
python tools/prof/timeline.py f64 > $OUT/timeline_new.log 2>&1
GMR_B200_LIB=$PWD/general_motion_retargeting_b200/csrc/$BASE python tools/prof/timeline.py f64 > $OUT/timeline_base.log 2>&1
for k in "GMR_PARTITION=7 GMR_PARTITION_PCT=55" "GMR_PARTITION=6 GMR_PARTITION_PCT=60" "GMR_PARTITION=10"; do
  env $k PROBE_T=300 python tools/prof/mix_case.py f64 2>&1 | tail -1 >> $OUT/knobs.log
done
if [ -z "$NO_NCU" ]; then
  ncu --set full --clock-control none --import-source on -k regex:gmr_retarget -s 2 -c 1 -f -o $OUT/bal python tools/prof/run_case.py f64 2368 60 same > $OUT/bal.log 2>&1
  CASE_CLIP=135 ncu --set full --clock-control none --import-source on -k regex:gmr_retarget -s 2 -c 1 -f -o $OUT/slow python tools/prof/run_case.py f64 148 60 same > $OUT/slow.log 2>&1
fi
cat $OUT/pytest.log $OUT/ab.log $OUT/knobs.log; head -c 1800 $OUT/timeline_new.log; echo; head -c 1800 $OUT/timeline_base.log
