#!/bin/bash
# The default bench line at N GPUs as the driver launches it.  Usage (under gpurun --gpus N): bash scripts/gpu_scale.sh <tag> <N>
TAG=${1:-scale}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1; nproc > $OUT/nproc.txt; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> $OUT/nproc.txt
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N > $OUT/bench_n$N.json ) 2> $OUT/bench_n$N.err
echo rc=$?; tail -4 $OUT/bench_n$N.err
python - $OUT/bench_n$N.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
def show(n,r):
    e=r.get('e2e') or {}
    print(n,'value %.0f step %.4f ms frac %.4f' % (r['value'], r['ms_per_step'], r['roofline']['frac']), {k:round(v) for k,v in e.items() if k.endswith('value') and isinstance(v,(int,float))})
show('main',d)
for k,v in d.get('configs',{}).items(): show(k,v)
print(d.get('clocks'), d.get('placement'))
PY
