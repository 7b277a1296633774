#!/bin/bash
# compute-sanitizer, the wide sweep: memcheck over EVERY GPU test (the drop-in blocks and the full-size tests
# included), initcheck and synccheck over all but the full-size ones, racecheck by kernel family.
# Usage (under gpurun): bash scripts/gpu_sanitizer2.sh [tag]
TAG=${1:-san2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
run() { # name, tool, seconds, pytest arguments...
  local name=$1 tool=$2 secs=$3; shift 3
  echo "== $tool: $name"
  timeout $secs $CS --tool $tool --print-limit 8 --error-exitcode 9 python -m pytest "$@" > $OUT/${tool}_$name.log 2>&1
  echo "exit $?" >> $OUT/${tool}_$name.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|^exit" $OUT/${tool}_$name.log | tail -4
}
run all memcheck 1500 tests -m gpu -q --timeout=900 -k "not full_size"
run full_size memcheck 1800 tests -m gpu -q --timeout=1500 -k "full_size"
run all initcheck 900 tests -m gpu -q --timeout=800 -k "not full_size"
run all synccheck 900 tests -m gpu -q --timeout=800 -k "not full_size"
run v4 racecheck 600 tests/test_parity_gpu.py -m gpu -q --timeout=500 -k "v4_streaming or random_cuts or shared_tuner or sliding_window"
run spectrum racecheck 600 tests/test_parity_gpu.py -m gpu -q --timeout=500 -k "spectrum"
run stage racecheck 600 tests/test_parity_gpu.py -m gpu -q --timeout=500 -k "stage_ or palette or audio_format or design_on_device"
