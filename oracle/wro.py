"""ctypes front-end for oracle/libwr_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module, and only as the checker or as the timed CPU baseline.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libwr_oracle.so")

AM, FM, USB, LSB = 0, 1, 2, 3
MODES = {"AM": AM, "FM": FM, "USB": USB, "LSB": LSB}

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_lib = None


def build(force=False):
    if force or not os.path.exists(SO) or (
            os.path.getmtime(SO) < max(os.path.getmtime(os.path.join(HERE, f))
                                       for f in ("wr_oracle.c", "wr_oracle.h", "shim/fftw3.h"))):
        subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    return SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(SO)
        L.wro_sintable.argtypes = [_fp]
        L.wro_phase_step.restype = C.c_int32
        L.wro_phase_step.argtypes = [C.c_int, C.c_uint]
        L.wro_lowpass_design.argtypes = [C.c_uint, C.c_uint, C.c_uint, _fp]
        L.wro_spectrum_window.argtypes = [C.c_uint, _fp]
        L.wro_mix.argtypes = [_fp, C.POINTER(C.c_uint32), C.c_int32, _fp, C.c_size_t, _fp]
        L.wro_fir_create.restype = C.c_void_p
        L.wro_fir_create.argtypes = [C.c_uint, _fp, C.c_uint, C.c_uint]
        L.wro_fir_set_taps.argtypes = [C.c_void_p, _fp, C.c_uint]
        L.wro_fir_process.restype = C.c_size_t
        L.wro_fir_process.argtypes = [C.c_void_p, _fp, C.c_size_t, _fp]
        L.wro_fir_destroy.argtypes = [C.c_void_p]
        L.wro_demod.argtypes = [C.c_int, _fp, _fp, C.c_size_t, _fp]
        L.wro_libm_atan2f.argtypes = [_fp, _fp, C.c_size_t, _fp]
        L.wro_libm_atan2f.restype = None
        L.wro_waterfall_index.argtypes = [_fp, C.c_size_t, C.c_void_p]
        L.wro_waterfall_index.restype = None
        L.wro_lame_scale.argtypes = [_fp, C.c_size_t, _fp]
        L.wro_lame_scale.restype = None
        L.wro_rtlsdr_convert.argtypes = [C.c_void_p, C.c_size_t, _fp]
        L.wro_rtlsdr_convert.restype = None
        L.wro_rx_create.restype = C.c_void_p
        L.wro_rx_create.argtypes = [C.c_uint, C.c_int, _fp, C.c_uint, C.c_uint, C.c_int,
                                    _fp, C.c_uint, C.c_uint]
        L.wro_rx_set_if.argtypes = [C.c_void_p, C.c_int]
        L.wro_rx_set_mode.argtypes = [C.c_void_p, C.c_int]
        L.wro_rx_set_taps.argtypes = [C.c_void_p, C.c_int, _fp, C.c_uint]
        L.wro_rx_process.restype = C.c_size_t
        L.wro_rx_process.argtypes = [C.c_void_p, _fp, C.c_size_t, _fp, _fp, _fp, _fp]
        L.wro_rx_destroy.argtypes = [C.c_void_p]
        L.wro_spectrum_create.restype = C.c_void_p
        L.wro_spectrum_create.argtypes = [C.c_uint, C.c_uint]
        L.wro_spectrum_process.restype = C.c_size_t
        L.wro_spectrum_process.argtypes = [C.c_void_p, _fp, C.c_size_t, _fp, C.c_size_t]
        L.wro_spectrum_get.argtypes = [C.c_void_p, _fp]
        L.wro_spectrum_get_bins.argtypes = [C.c_void_p, _fp]
        L.wro_spectrum_destroy.argtypes = [C.c_void_p]
        L.wro_bench.restype = C.c_double
        L.wro_bench.argtypes = [C.c_uint, C.c_size_t, C.c_uint, C.c_uint, _ip, _ip,
                                _fp, C.c_uint, C.c_uint, _fp, C.c_uint, C.c_uint,
                                _fp, C.c_uint, C.c_uint, C.c_uint, _fp]
        _lib = L
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_fp)


def sintable():
    out = np.empty(65536, np.float32)
    lib().wro_sintable(out.ctypes.data_as(_fp))
    return out


def phase_step(if_hz, fs):
    return int(lib().wro_phase_step(int(if_hz), int(fs)))


def lowpass_design(n, passband, fs):
    out = np.empty(n, np.float32)
    lib().wro_lowpass_design(n, passband, fs, out.ctypes.data_as(_fp))
    return out


def spectrum_window(n):
    out = np.empty(n, np.float32)
    lib().wro_spectrum_window(n, out.ctypes.data_as(_fp))
    return out


def mix(table, phase, step, iq):
    """Returns (mixed, new_phase)."""
    t, tp = _f(table)
    a, ap = _f(iq)
    out = np.empty_like(a)
    ph = C.c_uint32(phase)
    lib().wro_mix(tp, C.byref(ph), step, ap, a.size // 2, out.ctypes.data_as(_fp))
    return out, ph.value


class Fir:
    def __init__(self, channels, taps, decim):
        t, tp = _f(taps)
        self.ch, self.decim = channels, decim
        self.h = lib().wro_fir_create(channels, tp, t.size, decim)

    def set_taps(self, taps):
        t, tp = _f(taps)
        lib().wro_fir_set_taps(self.h, tp, t.size)

    def process(self, x):
        a, ap = _f(x)
        nframes = a.size // self.ch
        out = np.empty((nframes // self.decim) * self.ch, np.float32)
        lib().wro_fir_process(self.h, ap, nframes, out.ctypes.data_as(_fp))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            lib().wro_fir_destroy(self.h)
            self.h = None


def waterfall_index(db):
    """Palette index per bin as the browser computes it (waterfallhandler.cxx:62-68, waterfall.js:92-109)."""
    a, ap = _f(db)
    out = np.empty(a.size, np.uint8)
    lib().wro_waterfall_index(ap, a.size, out.ctypes.data)
    return out


def lame_scale(x):
    """MP3Encoder::encode's sample conversion (mp3encoder.cxx:66-73)."""
    a, ap = _f(x)
    out = np.empty(a.size, np.float32)
    lib().wro_lame_scale(ap, a.size, out.ctypes.data_as(_fp))
    return out.reshape(np.shape(x))


def rtlsdr_convert(u8):
    """RtlSdrTuner's byte-to-sample conversion (rtlsdrtuner.cxx:106)."""
    b = np.ascontiguousarray(u8, dtype=np.uint8)
    out = np.empty(b.size, np.float32)
    lib().wro_rtlsdr_convert(b.ctypes.data, b.size, out.ctypes.data_as(_fp))
    return out.reshape(b.shape)


def libm_atan2f(y, x):
    """atan2f of the C library installed on this box (what reference demodulator.cxx:97 calls)."""
    ya, yp = _f(y)
    xa, xp = _f(x)
    assert ya.size == xa.size
    out = np.empty(ya.size, np.float32)
    lib().wro_libm_atan2f(yp, xp, ya.size, out.ctypes.data_as(_fp))
    return out


def demod(mode, prev, iq):
    """prev: 2-element float32 array, updated in place."""
    a, ap = _f(iq)
    out = np.empty(a.size // 2, np.float32)
    assert prev.dtype == np.float32 and prev.size == 2
    rc = lib().wro_demod(mode, prev.ctypes.data_as(_fp), ap, a.size // 2, out.ctypes.data_as(_fp))
    if rc != 0:
        raise ValueError("bad mode")
    return out


class Rx:
    """One receiver chain with carried state."""

    def __init__(self, fs, if_hz, taps1, d1, mode, taps2, d2):
        t1, p1 = _f(taps1)
        t2, p2 = _f(taps2)
        self.d1, self.d2 = d1, d2
        self.h = lib().wro_rx_create(fs, if_hz, p1, t1.size, d1, MODES.get(mode, mode), p2, t2.size, d2)

    def set_if(self, hz):
        lib().wro_rx_set_if(self.h, hz)

    def set_mode(self, mode):
        lib().wro_rx_set_mode(self.h, MODES.get(mode, mode))

    def set_taps(self, which, taps):
        t, tp = _f(taps)
        lib().wro_rx_set_taps(self.h, which, tp, t.size)

    def process(self, iq, stages=False):
        a, ap = _f(iq)
        n = a.size // 2
        n1 = n // self.d1
        audio = np.empty(n1 // self.d2, np.float32)
        if stages:
            mixed = np.empty(2 * n, np.float32)
            chan = np.empty(2 * n1, np.float32)
            dem = np.empty(n1, np.float32)
            lib().wro_rx_process(self.h, ap, n, mixed.ctypes.data_as(_fp), chan.ctypes.data_as(_fp),
                                 dem.ctypes.data_as(_fp), audio.ctypes.data_as(_fp))
            return {"mixed": mixed, "channel": chan, "demod": dem, "audio": audio}
        lib().wro_rx_process(self.h, ap, n, None, None, None, audio.ctypes.data_as(_fp))
        return audio

    def __del__(self):
        if getattr(self, "h", None):
            lib().wro_rx_destroy(self.h)
            self.h = None


class Spectrum:
    def __init__(self, n, hop=None):
        self.n = n
        self.hop = hop or n
        self.h = lib().wro_spectrum_create(n, self.hop)
        if not self.h:
            raise ValueError("fft size must be a power of two, 0 < hop <= n")

    def process(self, iq, rows=True):
        a, ap = _f(iq)
        nframes = a.size // 2
        max_rows = nframes // self.hop + 2
        out = np.empty((max_rows, self.n), np.float32) if rows else None
        done = lib().wro_spectrum_process(self.h, ap, nframes,
                                          out.ctypes.data_as(_fp) if rows else None, max_rows)
        return out[:done] if rows else done

    def get(self):
        out = np.empty(self.n, np.float32)
        lib().wro_spectrum_get(self.h, out.ctypes.data_as(_fp))
        return out

    def bins(self):
        out = np.empty(2 * self.n, np.float32)
        lib().wro_spectrum_get_bins(self.h, out.ctypes.data_as(_fp))
        return out.view(np.complex64)

    def __del__(self):
        if getattr(self, "h", None):
            lib().wro_spectrum_destroy(self.h)
            self.h = None


def bench(fs, nframes, if_hz, modes, taps1, d1, taps2, d2, iq, n_streams=1, nthreads=1,
          warmup=1, blocks=4):
    """Seconds for `blocks` blocks of all receivers (max over worker threads)."""
    ifs = np.ascontiguousarray(if_hz, dtype=np.int32)
    md = np.ascontiguousarray(modes, dtype=np.int32)
    t1, p1 = _f(taps1)
    t2, p2 = _f(taps2)
    a, ap = _f(iq)
    assert a.size == 2 * nframes * n_streams
    chk = C.c_float(0)
    s = lib().wro_bench(fs, nframes, ifs.size, n_streams, ifs.ctypes.data_as(_ip),
                        md.ctypes.data_as(_ip), p1, t1.size, d1, p2, t2.size, d2, ap,
                        nthreads, warmup, blocks, C.byref(chk))
    return float(s), float(chk.value)
