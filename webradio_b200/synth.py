"""Seedable synthetic IQ generators and the BASELINE.json workload table.

Two generators, both deterministic so the oracle and the GPU path see
bit-identical input:
  * lattice_noise: counter-hash -> u8 -> (b-128)/128, i.e. exactly the sample lattice the
    reference's RTL-SDR tuner produces (reference src/io/rtlsdrtuner.cxx:104-108);
  * structured: complex carriers at the receivers' IFs, AM or FM modulated, plus lattice noise.
Host-side numpy only: this is input synthesis, not part of the DSP path.
"""
import numpy as np

AM, FM, USB, LSB = 0, 1, 2, 3
MODE_NAMES = ["AM", "FM", "USB", "LSB"]


def _hash32(x):
    """Counter-based integer hash (lowbias32 finaliser) on uint32 arrays."""
    x = x.astype(np.uint32, copy=True)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def lattice_noise(nframes, seed=0xB200, stream=0, start=0):
    """Interleaved IQ float32[2*nframes] on the RTL-SDR lattice (b-128)/128, b in 0..255."""
    key = stream_key(seed, stream)
    idx = (np.arange(2 * start, 2 * (start + nframes), dtype=np.uint64)
           & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    h = _hash32(idx ^ np.uint32(key))
    b = (h >> np.uint32(24)).astype(np.float32)
    return ((b - np.float32(128.0)) / np.float32(128.0)).astype(np.float32)


def stream_key(seed, stream):
    """The per-stream hash key of lattice_noise (shared by the numpy and the torch generator)."""
    return int(_hash32(np.array([(seed * 0x9E3779B1 + stream * 0x85EBCA77 + 1) & 0xFFFFFFFF],
                                dtype=np.uint32))[0])


def lattice_u8_torch(nframes, streams, start=0, seed=0xB200, device="cuda", chunk_elems=1 << 26):
    """The bytes behind lattice_noise for several streams at once, generated ON THE DEVICE:
    uint8 tensor [len(streams), nframes, 2], bit-identical to
    ((lattice_noise(nframes, seed, s, start) * 128) + 128) for every s in `streams` (bench.py checks a
    slice against the numpy generator in every run).  int64 arithmetic masked to 32 bits stands in
    for uint32; the products wrap, which leaves their low 32 bits right."""
    import torch
    M = 0xFFFFFFFF
    T = len(streams)
    out = torch.empty((T, nframes, 2), dtype=torch.uint8, device=device)
    idx = ((torch.arange(2 * start, 2 * (start + nframes), dtype=torch.int64, device=device)) & M)
    per = max(1, chunk_elems // max(1, 2 * nframes))

    def h32(x):
        x = x ^ (x >> 16)
        x = (x * 0x7FEB352D) & M
        x = x ^ (x >> 15)
        x = (x * 0x846CA68B) & M
        return x ^ (x >> 16)

    for t0 in range(0, T, per):
        keys = torch.tensor([stream_key(seed, s) for s in streams[t0:t0 + per]], dtype=torch.int64, device=device)
        h = h32(idx[None, :] ^ keys[:, None])
        out[t0:t0 + per] = (h >> 24).to(torch.uint8).view(-1, nframes, 2)
    return out


def u8_to_f32_torch(u8):
    """RtlSdrTuner's conversion (reference src/io/rtlsdrtuner.cxx:106) on a torch tensor; exact."""
    import torch
    return ((u8.to(torch.float32) - 128.0) / 128.0).contiguous()


def receiver_ifs(n_rx, fs):
    """if_r = round((r - R/2 + 0.5) * Fs * 0.8 / R) Hz."""
    r = np.arange(n_rx, dtype=np.float64)
    return np.round((r - n_rx / 2 + 0.5) * fs * 0.8 / n_rx).astype(np.int32)


def structured(nframes, fs, ifs, modes, start=0, amp=0.5, noise_db=-30.0, seed=0xB200, stream=0,
               fm_dev=5000.0, tone=1000.0):
    """Sum of modulated carriers (one per receiver IF) + lattice noise, float32 interleaved IQ.

    Amplitude is split across carriers so the sum stays within [-1, 1].
    """
    n = np.arange(start, start + nframes, dtype=np.float64)
    t = n / fs
    sig = np.zeros(nframes, dtype=np.complex128)
    k = max(1, len(ifs))
    a = amp / k
    for f0, m in zip(ifs, modes):
        if m == FM:
            ph = 2 * np.pi * f0 * t + (fm_dev / tone) * np.sin(2 * np.pi * tone * t)
            sig += a * np.exp(1j * ph)
        else:
            env = 1.0 + 0.5 * np.sin(2 * np.pi * tone * t)
            sig += a * env * np.exp(2j * np.pi * f0 * t)
    out = np.empty(2 * nframes, dtype=np.float32)
    out[0::2] = sig.real
    out[1::2] = sig.imag
    out += np.float32(10.0 ** (noise_db / 20.0)) * lattice_noise(nframes, seed, stream, start)
    return out


def windowed_sinc(n, cutoff):
    """Hamming windowed-sinc low-pass, n taps, cutoff as a fraction of the sample rate.  Tap VALUES
    are an input of the FIR; the reference cannot design non-power-of-two lengths itself
    (FIR_LENGTH is a compile-time 64, reference src/dsp/lowpass.cxx:39)."""
    k = np.arange(n) - (n - 1) / 2.0
    h = 2 * cutoff * np.sinc(2 * cutoff * k) * (0.54 - 0.46 * np.cos(2 * np.pi * np.arange(n) / max(n - 1, 1)))
    return h.astype(np.float32)


# Integer-legal variants of BASELINE.json configs (the reference rejects
# non-integer rate ratios, reference src/dsp/dspblock.cxx:119-130).
WORKLOADS = {
    # cfg1a: the shipped operating point (reference src/main.cxx:74-75, src/radio.cxx:78-81)
    "cfg1": dict(fs=2400000, frames=102400, n_rx=1, n_streams=1, n1=64, d1=10, pb1=80000,
                 n2=64, d2=5, pb2=8000, modes="FM",
                 desc="single FM receiver, 2.4 MSPS -> 240 k -> 48 k, 64/64 taps (reference CPU case)"),
    # cfg1b: BASELINE configs[0] as written says 2.048 MSPS -> 48 kHz, a ratio of 42.67 that the
    # reference rejects (dspblock.cxx:126-130); SURVEY.md 8d's integer-legal form of it:
    # 2.048 M -> /8 -> 256 k (pass-band 100 kHz) -> FM -> /4 -> 64 k (pass-band 15 kHz)
    "cfg1b": dict(fs=2048000, frames=102400, n_rx=1, n_streams=1, n1=64, d1=8, pb1=100000,
                  n2=64, d2=4, pb2=15000, modes="FM",
                  desc="single WBFM receiver, synthetic 2.048 MSPS IQ -> 256 k -> 64 k audio, 64/64 taps "
                       "(integer-legal form of BASELINE configs[0])"),
    "cfg2": dict(fs=2400000, frames=102400, n_rx=64, n_streams=1, n1=127, d1=50, pb1=12500,
                 n2=64, d2=1, pb2=3000, modes="FM",
                 desc="64 NBFM receivers on one 2.4 MSPS tuner, 127-tap FIR, decim 50"),
    "cfg3": dict(fs=2400000, frames=102400, n_rx=1024, n_streams=1024, n1=255, d1=50, pb1=12500,
                 n2=64, d2=1, pb2=3000, modes="AM",
                 desc="1024 independent AM streams, 255-tap FIR, decim 50, 2.4 MSPS float2 IQ"),
    "cfg5": dict(fs=10000000, frames=409600, n_rx=1024, n_streams=16, n1=127, d1=40, pb1=12500,
                 n2=64, d2=5, pb2=3000, modes="mixed",
                 desc="16 tuners x 64 mixed FM/AM/USB/LSB receivers per GPU, 10 MSPS -> 250 k -> 50 k"),
}


def workload_modes(w):
    n = w["n_rx"]
    if w["modes"] == "mixed":
        # r mod 4 -> AM, FM, USB, LSB
        return (np.arange(n) % 4).astype(np.int32)
    return np.full(n, MODE_NAMES.index(w["modes"]), dtype=np.int32)


def workload_ifs(w):
    per = w["n_rx"] // w["n_streams"]
    return np.tile(receiver_ifs(per, w["fs"]), w["n_streams"]).astype(np.int32)
