#!/bin/bash
# compute-sanitizer over the hot path: memcheck on smoke() and on the ragged / golden-chain /
# u8 / pipelined parity tests (all three kernel families), racecheck and synccheck on smoke().
# Usage (under gpurun): bash scripts/gpu_sanitizer.sh [tag]
TAG=${1:-san}
OUT=gpurun_out/$TAG
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
run() { # name, tool, command...
  local name=$1 tool=$2; shift 2
  echo "== $tool: $name"
  timeout 280 $CS --tool $tool --print-limit 20 --error-exitcode 9 "$@" > $OUT/${tool}_$name.log 2>&1
  echo "exit $?" >> $OUT/${tool}_$name.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|smoke ok|^exit" $OUT/${tool}_$name.log | tail -4
}
run smoke memcheck python -c "import __graft_entry__ as g; g.smoke()"
run parity memcheck python -m pytest tests/test_parity_gpu.py -q -x --timeout=250 -k "ragged or golden_chain or u8_ingest or pipelined or v3_ragged"
# round 2: the streaming channel kernel (cuts with one / two receivers per warp, ragged ends), the sliding-window
# audio FIR, the spectrum kernels (v4 / v3 / v2, both hops), the shared upload
run v4 memcheck python -m pytest tests/test_parity_gpu.py -q -x --timeout=270 -k "v4_streaming or v4_is_what or sliding_window or random_cuts or shared_tuner"
run spectrum memcheck python -m pytest tests/test_parity_gpu.py -q -x --timeout=270 -k "spectrum_8192 or spectrum_rows or share_one_upload or pipelined_by_stream"
run v4 synccheck python -m pytest tests/test_parity_gpu.py -q -x --timeout=270 -k "v4_is_what"
run smoke synccheck python -c "import __graft_entry__ as g; g.smoke()"
run smoke racecheck python -c "import __graft_entry__ as g; g.smoke()"
