"""Debug: v4 over several device-resident blocks at several sizes; counts receivers whose audio is all zero."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from webradio_b200 import capi, synth
from oracle import wro

def run(R, F, nblocks=3):
    w = synth.WORKLOADS["cfg3"]
    rng = np.random.default_rng(1)
    t1 = (rng.uniform(-1, 1, 255) / 255 * 4).astype(np.float32)
    t2 = (rng.uniform(-1, 1, 64) / 64 * 4).astype(np.float32)
    bank = capi.Bank(R, R, F, 255, 50, 64, 1)
    ifs = synth.receiver_ifs(R, w["fs"])
    for r in range(R):
        bank.set_taps(r, 0, t1); bank.set_taps(r, 1, t2); bank.set_if(r, int(ifs[r]), w["fs"]); bank.set_mode(r, 0); bank.set_stream(r, r)
    m2 = F // 50
    st = torch.cuda.ExternalStream(bank.stream())
    torch.cuda.set_stream(st)
    picks = sorted(set([0, 1, R // 2, R - 1]))
    orx = {r: wro.Rx(w["fs"], int(ifs[r]), t1, 50, 0, t2, 1) for r in picks}
    g = torch.Generator(device="cuda").manual_seed(5)
    for b in range(nblocks):
        d_iq = (torch.randint(0, 256, (R, F, 2), device="cuda", generator=g, dtype=torch.int16).float() - 128.0) / 128.0
        d_audio = torch.full((R, m2), 7.0, device="cuda")
        bank.process_device(d_iq.data_ptr(), F, F, d_audio.data_ptr(), m2, st.cuda_stream)
        torch.cuda.synchronize()
        a = d_audio.cpu().numpy()
        zero = int((np.abs(a).max(axis=1) == 0).sum()); seven = int((a == 7.0).all(axis=1).sum())
        ok = [bool(np.array_equal(a[r].view(np.uint32), orx[r].process(d_iq[r].cpu().numpy().ravel()).view(np.uint32))) for r in picks]
        print(f"R={R} F={F} block {b}: variant {bank.variant_in_use()} zero-rows {zero} untouched-rows {seven} picks ok {ok}", flush=True)
    bank.close()

for R, F in [(160, 102400), (1024, 25600)]:
    run(R, F)
