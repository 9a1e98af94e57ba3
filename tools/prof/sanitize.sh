#!/bin/bash
# compute-sanitizer over the smallest instance of every launch shape (tools/prof/sanitize_cases.py).
# Usage: tools/prof/sanitize.sh <outdir>      (run on a GPU box; logs are copied to profiles/ by hand)
out=${1:-gpurun_out/sanitizer}
mkdir -p "$out"
CS=/usr/local/cuda/bin/compute-sanitizer
run() {   # tool case precision [env...]
  local tool=$1 cs=$2 prec=$3; shift 3
  local log="$out/${tool}_${cs}_${prec}.log"
  env "$@" timeout 900 $CS --tool $tool --print-limit 20 --error-exitcode 9 \
      python tools/prof/sanitize_cases.py $cs $prec > "$log" 2>&1
  echo "$tool $cs $prec exit=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$log" | tail -1)"
}
for prec in f64 f32; do
  run memcheck  twophase $prec GMR_PARTITION=1
  run memcheck  plain    $prec
  run racecheck twophase $prec GMR_PARTITION=1
  run racecheck plain    $prec
  run synccheck twophase $prec GMR_PARTITION=1
done
run memcheck  mixed   f64
run racecheck mixed   f64
run memcheck  stream  f64
run racecheck stream  f64
run memcheck  dataset f64
run racecheck dataset f64
run memcheck  host    f64
run initcheck plain   f64
