"""ctypes front-end for tests/harness/graph_harness.cxx (TEST DRIVER).

The same C ABI is exported by two builds of that one source:
  * oracle/_ref/libwr_ref.so               -- the unmodified reference blocks (CPU)
  * tests/harness/libwr_blocks_harness.so  -- webradio_b200's GPU-backed drop-in blocks
  * tests/harness/libwr_blocks_harness_mock.so -- the same drop-in blocks over a CPU stand-in for
    the device entry points (tests/harness/mock_capi.cxx): host logic only, for the CPU suite
so a test can build the reference's receiver graph (reference src/radio.cxx:62-90)
on either and compare stage by stage.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libwr_ref.so")
BLOCKS_SO = os.path.join(ROOT, "tests", "harness", "libwr_blocks_harness.so")
MOCK_SO = os.path.join(ROOT, "tests", "harness", "libwr_blocks_harness_mock.so")
REF_O0_SO = os.path.join(ROOT, "oracle", "_ref", "libwr_ref_O0.so")
_PATHS = {"ref": REF_SO, "ref_O0": REF_O0_SO, "blocks": BLOCKS_SO, "mock": MOCK_SO}

MODES = {"AM": 0, "FM": 1, "USB": 2, "LSB": 3}
STAGES = {"mixed": 0, "channel": 1, "demod": 2, "audio": 3}

_fp = C.POINTER(C.c_float)


def _bind(path):
    # RTLD_LOCAL: both builds define the same C++ class names (DspBlock, LowPass, ...)
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    lib.wrh_graph_create.restype = C.c_void_p
    lib.wrh_graph_create.argtypes = [C.c_uint, C.c_uint]
    lib.wrh_graph_add_receiver.restype = C.c_int
    lib.wrh_graph_add_receiver.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_uint, C.c_uint,
                                           C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_uint]
    lib.wrh_graph_add_spectrum.argtypes = [C.c_void_p, C.c_uint]
    lib.wrh_graph_start.argtypes = [C.c_void_p]
    lib.wrh_graph_run.argtypes = [C.c_void_p, _fp]
    lib.wrh_graph_get.restype = C.c_long
    lib.wrh_graph_get.argtypes = [C.c_void_p, C.c_int, C.c_int, _fp, C.c_long]
    lib.wrh_graph_set_if.argtypes = [C.c_void_p, C.c_int, C.c_int]
    if hasattr(lib, "wrh_graph_detach"):
        lib.wrh_graph_detach.argtypes = [C.c_void_p, C.c_int]
        lib.wrh_graph_attach.argtypes = [C.c_void_p, C.c_int]
    if hasattr(lib, "wrh_graph_restart"):
        lib.wrh_graph_restart.argtypes = [C.c_void_p]
    lib.wrh_graph_set_mode.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
    lib.wrh_graph_get_mode.argtypes = [C.c_void_p, C.c_int]
    lib.wrh_graph_set_passband.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint]
    lib.wrh_graph_set_taps.argtypes = [C.c_void_p, C.c_int, C.c_int, _fp, C.c_uint]
    lib.wrh_graph_get_taps.argtypes = [C.c_void_p, C.c_int, C.c_int, _fp, C.c_uint]
    lib.wrh_graph_rates.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint)]
    lib.wrh_graph_spectrum.argtypes = [C.c_void_p, _fp]
    lib.wrh_graph_profile.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_ulonglong),
                                      C.POINTER(C.c_ulonglong)]
    lib.wrh_graph_destroy.argtypes = [C.c_void_p]
    lib.wrh_set_quiet.argtypes = [C.c_int]
    if hasattr(lib, "wrh_ref_sintable"):
        lib.wrh_ref_sintable.argtypes = [_fp]
    return lib


_libs = {}


def load(which):
    """which: 'ref', 'blocks' or 'mock'."""
    path = _PATHS[which]
    if path not in _libs:
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        _libs[path] = _bind(path)
    return _libs[path]


def have(which):
    return os.path.exists(_PATHS[which])


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_fp)


class Graph:
    """One tuner stream feeding N receivers (+ optional SpectrumSink)."""

    def __init__(self, which, fs, block_frames):
        self.lib = load(which)
        self.h = self.lib.wrh_graph_create(fs, block_frames)
        self.block_frames = block_frames
        self.caps = []

    def add_receiver(self, if_hz=0, ch_passband=80000, ch_rate=240000, ch_decim=0, mode="AM",
                     au_passband=8000, au_rate=48000, au_decim=0, capture=0xF):
        idx = self.lib.wrh_graph_add_receiver(self.h, if_hz, ch_passband, ch_rate, ch_decim,
                                              MODES[mode] if isinstance(mode, str) else mode,
                                              au_passband, au_rate, au_decim, capture)
        if idx < 0:
            raise RuntimeError("add_receiver failed")
        self.caps.append(capture)
        return idx

    def add_spectrum(self, fft_size):
        if self.lib.wrh_graph_add_spectrum(self.h, fft_size) != 0:
            raise RuntimeError("add_spectrum failed")

    def start(self):
        return self.lib.wrh_graph_start(self.h) == 0

    def run(self, iq):
        a, p = _f32(iq)
        assert a.size == 2 * self.block_frames, (a.size, self.block_frames)
        return self.lib.wrh_graph_run(self.h, p) == 0

    def get(self, rx, stage):
        s = STAGES[stage] if isinstance(stage, str) else stage
        n = self.lib.wrh_graph_get(self.h, rx, s, None, 0)
        if n < 0:
            raise KeyError((rx, stage))
        out = np.empty(n, dtype=np.float32)
        self.lib.wrh_graph_get(self.h, rx, s, out.ctypes.data_as(_fp), n)
        return out

    def detach(self, rx):
        """Receiver::setFrontEnd(NULL) on a live pipeline."""
        return self.lib.wrh_graph_detach(self.h, rx) == 0

    def attach(self, rx):
        """Receiver::setFrontEnd(frontEnd) on a live pipeline; False if the chain did not start."""
        return self.lib.wrh_graph_attach(self.h, rx) == 0

    def restart(self):
        """DspSource::stop() followed by start()."""
        return self.lib.wrh_graph_restart(self.h) == 0

    def set_if(self, rx, hz):
        return self.lib.wrh_graph_set_if(self.h, rx, hz)

    def set_mode(self, rx, mode):
        return self.lib.wrh_graph_set_mode(self.h, rx, mode.encode()) == 0

    def get_mode(self, rx):
        return self.lib.wrh_graph_get_mode(self.h, rx)

    def set_passband(self, rx, which, hz):
        return self.lib.wrh_graph_set_passband(self.h, rx, which, hz)

    def set_taps(self, rx, which, taps):
        a, p = _f32(taps)
        if self.lib.wrh_graph_set_taps(self.h, rx, which, p, a.size) != 0:
            raise RuntimeError("set_taps failed")

    def get_taps(self, rx, which):
        n = self.lib.wrh_graph_get_taps(self.h, rx, which, None, 0)
        out = np.empty(n, dtype=np.float32)
        self.lib.wrh_graph_get_taps(self.h, rx, which, out.ctypes.data_as(_fp), n)
        return out

    def rates(self, rx):
        r = (C.c_uint * 8)()
        self.lib.wrh_graph_rates(self.h, rx, r)
        return list(r)

    def spectrum(self, n):
        out = np.empty(n, dtype=np.float32)
        got = self.lib.wrh_graph_spectrum(self.h, out.ctypes.data_as(_fp))
        assert got == n, (got, n)
        return out

    def profile(self, rx):
        ns = (C.c_ulonglong * 4)()
        fr = (C.c_ulonglong * 4)()
        self.lib.wrh_graph_profile(self.h, rx, ns, fr)
        return list(ns), list(fr)

    def close(self):
        if self.h:
            self.lib.wrh_graph_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def ref_sintable():
    lib = load("ref")
    out = np.empty(65536, dtype=np.float32)
    lib.wrh_ref_sintable(out.ctypes.data_as(_fp))
    return out
