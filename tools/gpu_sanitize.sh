#!/bin/bash
# compute-sanitizer over the kernels: memcheck on everything, racecheck + synccheck on the shared-memory
# kernels.  Usage (GPU box, repo root): bash tools/gpu_sanitize.sh [tag]
TAG=${1:-san}
OUT=gpurun_out/$TAG
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # name tool parts...
  local name=$1 tool=$2; shift 2
  echo "== $tool: $*"
  timeout 1500 $CS --tool $tool --error-exitcode 9 --print-limit 20 python tools/sanitize_workload.py "$@" > $OUT/${name}.log 2>&1
  echo "$name rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload ok|Error|error" $OUT/${name}.log | head -8
}
run memcheck_all memcheck smoke config5 occupancy multi gap dreamer host
run racecheck_env racecheck smoke config5 occupancy multi
run racecheck_policy racecheck gap dreamer
run synccheck_all synccheck smoke config5 occupancy multi gap dreamer
ls -la $OUT
