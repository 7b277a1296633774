#!/bin/bash
TAG=${1:-fft4}; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_blocks_gpu.py -q -x --timeout=600 -k "spectrum or palette or upload" 2>&1 | tail -6 | tee $OUT/pytest.log
for e in "WR_FFT_V4=1" "WR_FFT_V4=0" "$@"; do
env $e timeout 600 python bench.py --workload cfg4 --no-cpu-baseline --no-e2e 2>$OUT/err.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('[$e] cfg4 value %.0f step %.4f ms frac %.3f parity %s' % (d['value'], d['ms_per_step'], r['frac'], d['parity']['max_error_over_frame_peak']))"
done
tail -3 $OUT/err.log
