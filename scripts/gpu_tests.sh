#!/bin/bash
# Quick GPU-box visit: parity tests + smoke only.  Usage: bash scripts/gpu_tests.sh [tag] [pytest args]
TAG=${1:-t}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 "$@" 2>&1 | tail -60 | tee $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.log
