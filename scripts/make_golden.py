#!/usr/bin/env python
"""Generates tests/golden/*.npz from the REAL reference (oracle/_ref/libwr_ref.so, i.e.
mikestir/webradio's unmodified src/dsp + src/io/spectrumsink.cxx compiled by oracle/Makefile).

Run in the build container (where /root/reference is mounted):
    make -C oracle ref && python scripts/make_golden.py
The fixtures pin the plain-C oracle port (tests/test_oracle_golden.py) wherever the
reference itself is not available (e.g. on the GPU box).
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "harness"))
import graphlib as G  # noqa: E402
from webradio_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def lattice_u8(nframes, stream, start):
    x = synth.lattice_noise(nframes, stream=stream, start=start)
    b = np.round(x * 128.0 + 128.0).astype(np.uint8)
    assert np.array_equal(((b.astype(np.float32) - 128.0) / 128.0).astype(np.float32), x)
    return b


def chain_case(name, fs, frames, if_hz, mode, d1, d2, taps1=None, taps2=None, blocks=3, stream=0,
               events=None, pb1=80000, pb2=8000):
    with G.Graph("ref", fs, frames) as g:
        g.add_receiver(if_hz=if_hz, ch_passband=pb1, ch_rate=0, ch_decim=d1, mode=mode,
                       au_passband=pb2, au_rate=0, au_decim=d2)
        assert g.start()
        if taps1 is not None:
            g.set_taps(0, 0, taps1)
        if taps2 is not None:
            g.set_taps(0, 1, taps2)
        t1, t2 = g.get_taps(0, 0), g.get_taps(0, 1)
        ins, outs = [], {s: [] for s in G.STAGES}
        for b in range(blocks):
            for ev in (events or {}).get(b, []):
                if ev[0] == "if":
                    g.set_if(0, ev[1])
                else:
                    assert g.set_mode(0, ev[1])
            u8 = lattice_u8(frames, stream, b * frames)
            ins.append(u8)
            assert g.run(((u8.astype(np.float32) - 128.0) / 128.0).astype(np.float32))
            for s in G.STAGES:
                outs[s].append(g.get(0, s))
    ev_flat = np.array([[b, 0 if e[0] == "if" else 1, e[1] if e[0] == "if" else synth.MODE_NAMES.index(e[1])]
                        for b, lst in sorted((events or {}).items()) for e in lst], dtype=np.int64).reshape(-1, 3)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), fs=fs, frames=frames, if_hz=if_hz,
                        mode=synth.MODE_NAMES.index(mode), d1=d1, d2=d2, taps1=t1, taps2=t2,
                        iq_u8=np.stack(ins), events=ev_flat,
                        mixed_last=outs["mixed"][-1], channel=np.stack(outs["channel"]),
                        demod=np.stack(outs["demod"]), audio=np.stack(outs["audio"]))
    print("wrote", name)


def hamming_lowpass(n, cutoff):
    """Plain windowed-sinc taps for the non-power-of-two cases (tap VALUES are just an input to
    LowPass::process; the reference cannot design these itself, lowpass.cxx:39)."""
    k = np.arange(n) - (n - 1) / 2.0
    h = 2 * cutoff * np.sinc(2 * cutoff * k) * (0.54 - 0.46 * np.cos(2 * np.pi * np.arange(n) / (n - 1)))
    return h.astype(np.float32)


def main():
    os.makedirs(OUT, exist_ok=True)
    tbl = G.ref_sintable()
    pick = np.array([0, 1, 2, 3, 100, 8191, 8192, 16383, 16384, 16385, 24576, 32767, 32768, 32769,
                     40000, 49151, 49152, 49153, 60000, 65534, 65535])
    designs = {}
    for fs, rate, pb in [(2400000, 240000, 80000), (2400000, 240000, 200000), (2400000, 48000, 12500),
                         (2048000, 256000, 100000), (240000, 48000, 8000), (2400000, 240000, 1200000)]:
        with G.Graph("ref", fs, 64 * (fs // rate)) as g:
            g.add_receiver(ch_passband=pb, ch_rate=rate, au_rate=0, au_decim=1)
            assert g.start()
            designs[f"design_{fs}_{pb}"] = g.get_taps(0, 0)
    np.savez_compressed(os.path.join(OUT, "tables.npz"),
                        sintable_sha256=np.frombuffer(hashlib.sha256(tbl.tobytes()).digest(), dtype=np.uint8),
                        sintable_idx=pick, sintable_val=tbl[pick], **designs)
    print("wrote tables")

    fs = 2400000
    chain_case("chain_fm_default", fs, 8000, 100000, "FM", 10, 5)
    chain_case("chain_am_default", fs, 8000, -345678, "AM", 10, 5, stream=1)
    chain_case("chain_usb_127_d50", fs, 6400, 612345, "USB", 50, 1, taps1=hamming_lowpass(127, 0.005), stream=2)
    chain_case("chain_am_255_d50", fs, 6400, -1000001, "AM", 50, 1, taps1=hamming_lowpass(255, 0.005), stream=3)
    chain_case("chain_events", fs, 4000, 50000, "LSB", 10, 5, blocks=5, stream=4,
               events={1: [("if", -250000)], 2: [("mode", "FM")], 3: [("mode", "USB"), ("if", 7)], 4: [("mode", "AM")]})
    chain_case("chain_short_blocks", fs, 40, 123456, "FM", 10, 2, blocks=8, stream=5)

    for n in (512, 8192):
        F = 2 * n + n // 2
        with G.Graph("ref", fs, F) as g:
            g.add_spectrum(n)
            assert g.start()
            ins, rows = [], []
            for b in range(2):
                u8 = lattice_u8(F, 6, b * F)
                # add a strong tone so the spectrum has dynamic range
                x = ((u8.astype(np.float32) - 128.0) / 128.0).astype(np.float32) * np.float32(0.25)
                ph = 2 * np.pi * 0.1234 * np.arange(b * F, (b + 1) * F)
                x[0::2] += np.float32(0.5) * np.cos(ph).astype(np.float32)
                x[1::2] += np.float32(0.5) * np.sin(ph).astype(np.float32)
                ins.append(x)
                assert g.run(x)
                rows.append(g.spectrum(n))
        np.savez_compressed(os.path.join(OUT, f"spectrum_{n}.npz"), n=n, frames=F,
                            iq=np.stack(ins).astype(np.float16 if False else np.float32), db=np.stack(rows))
        print("wrote spectrum", n)


if __name__ == "__main__":
    main()
