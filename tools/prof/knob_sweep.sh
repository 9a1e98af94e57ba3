#!/bin/bash
# Scheduler knob sweep on one workload: tools/prof/knob_sweep.sh "<mix env>" "<knobs 1>" "<knobs 2>" ...   (one process per line)
mix=$1; shift
for k in "$@"; do env $mix $k python tools/prof/mix_case.py f64 2>&1 | tail -1; done
