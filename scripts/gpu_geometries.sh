#!/bin/bash
# Per-geometry throughput of the channel path: 64 receivers on one 2.4 MSPS tuner, 102400-frame blocks,
# device-resident, for the instantiated geometries (v3) and a few that fall to v2 / v1.
# Usage (under gpurun): bash scripts/gpu_geometries.sh <tag>
TAG=${1:-geo}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python - <<'PY' 2>&1 | tee $OUT/geometries.txt
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from webradio_b200 import capi, synth
fs, F, R = 2400000, 102400, 64
print("64 receivers (FM) on one 2.4 MSPS tuner, 102400-frame blocks, device-resident, CUDA events, 200 blocks")
print("%-22s %-8s %12s %14s" % ("channel FIR / decim", "kernels", "us per block", "MS/s (sum rx)"))
for n1, d1, n2, d2 in [(64, 10, 64, 5), (64, 8, 64, 4), (127, 50, 64, 1), (127, 40, 64, 5), (255, 50, 64, 1),
                       (64, 16, 64, 3), (96, 20, 64, 2), (128, 25, 64, 2), (255, 25, 64, 2), (63, 5, 64, 10), (31, 7, 32, 7)]:
    bank = capi.Bank(1, R, F, n1, d1, n2, d2)
    t1 = synth.windowed_sinc(n1, 0.4 / d1); t2 = synth.windowed_sinc(n2, 0.4 / d2)
    ifs = synth.receiver_ifs(R, fs)
    for r in range(R):
        bank.set_taps(r, 0, t1); bank.set_taps(r, 1, t2); bank.set_if(r, int(ifs[r]), fs); bank.set_mode(r, 1)
    st = torch.cuda.ExternalStream(bank.stream())
    with torch.cuda.stream(st):
        iq = [torch.from_numpy(synth.lattice_noise(F, stream=i)).cuda() for i in range(8)]
        m2 = F // d1 // d2
        out = torch.zeros(R, m2, device="cuda")
        for i in range(10):
            bank.process_device(iq[i % 8].data_ptr(), F, F, out.data_ptr(), m2, st.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        n = 200
        for i in range(n):
            bank.process_device(iq[i % 8].data_ptr(), F, F, out.data_ptr(), m2, st.cuda_stream)
        e1.record(st)
        e1.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    print("%-22s v%-7d %12.2f %14.0f" % (f"{n1} taps / {d1}", bank.variant_in_use(), us, R * F / us))
    bank.close()
PY
