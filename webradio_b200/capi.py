"""ctypes binding of libwebradio_b200.so (include/webradio_b200.h).

This is plumbing for tests and bench.py: every call goes straight through the C ABI, which is
the drop-in boundary.  There is no Python or CPU implementation behind it -- if the shared
library is missing, importing fails loudly; if there is no CUDA device, the create calls fail.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WEBRADIO_B200_LIB") or os.path.join(HERE, "libwebradio_b200.so")   # the override is for A/B runs of two builds

AM, FM, USB, LSB = 0, 1, 2, 3
MODES = {"AM": AM, "FM": FM, "USB": USB, "LSB": LSB}
RESET_PHASE, RESET_CHANNEL, RESET_DEMOD, RESET_AUDIO = 1, 2, 4, 8
STAGE_CHANNEL, STAGE_DEMOD = 1, 2
AUDIO_FLOAT, AUDIO_LAME = 0, 1

_fp = C.POINTER(C.c_float)

# every symbol include/webradio_b200.h declares (tests/test_abi.py checks the export list)
SYMBOLS = [
    "wr_version", "wr_last_error", "wr_device_count", "wr_phase_step", "wr_build_sintable",
    "wr_lowpass_design", "wr_lo_compress_check", "wr_lo3_compress_check", "wr_plan_runs", "wr_atan2f_host",
    "wr_bank_create", "wr_bank_destroy", "wr_bank_set_sintable", "wr_rx_set_stream",
    "wr_rx_set_phase_step", "wr_rx_set_taps", "wr_bank_design_taps", "wr_rx_get_taps", "wr_rx_set_mode", "wr_rx_reset", "wr_rx_set_phase", "wr_rx_get_phase", "wr_rx_set_lookback", "wr_rx_get_lookback",
    "wr_bank_process", "wr_bank_process_device", "wr_bank_submit", "wr_bank_wait",
    "wr_bank_process_u8", "wr_bank_process_device_u8", "wr_bank_submit_u8",
    "wr_bank_run_device_steps_u8", "wr_bank_run_host_steps_u8",
    "wr_bank_pipeline_depth", "wr_bank_set_handover", "wr_bank_run_device_steps", "wr_bank_run_host_steps", "wr_bank_stream", "wr_bank_sync", "wr_bank_keep_channel",
    "wr_bank_read_stage", "wr_bank_set_audio_format", "wr_bank_set_variant", "wr_bank_variant_in_use", "wr_bank_launch_count", "wr_bank_set_timing",
    "wr_bank_kernel_times",
    "wr_stage_create", "wr_stage_destroy", "wr_stage_mix", "wr_stage_fir_config", "wr_stage_fir",
    "wr_stage_fir_reset", "wr_stage_demod", "wr_stage_atan2f", "wr_stage_palette",
    "wr_spectrum_create", "wr_spectrum_destroy", "wr_spectrum_process", "wr_spectrum_process_device",
    "wr_spectrum_get", "wr_spectrum_get_palette", "wr_spectrum_launch_count", "wr_spectrum_sync",
    "wr_upload_create", "wr_upload_destroy", "wr_upload_capacity", "wr_upload_device", "wr_upload_begin", "wr_upload_finish",
    "wr_bank_process_upload", "wr_spectrum_process_upload", "wr_spectrum_reserve", "wr_host_alloc", "wr_host_free",
    "wr_rx_get_history", "wr_rx_set_history",
]

_lib = None


class WrError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()); "
                          "webradio_b200 has no fallback implementation")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, u, i, sz = C.c_void_p, C.c_uint, C.c_int, C.c_size_t
    L.wr_version.restype = C.c_char_p
    L.wr_last_error.restype = C.c_char_p
    L.wr_phase_step.restype = C.c_int32
    L.wr_phase_step.argtypes = [i, u]
    L.wr_build_sintable.argtypes = [_fp]
    L.wr_lowpass_design.argtypes = [u, u, u, _fp]
    L.wr_lo_compress_check.argtypes = [_fp]
    L.wr_lo3_compress_check.argtypes = [_fp]
    L.wr_plan_runs.argtypes = [u, u, u, u, u] + [C.POINTER(C.c_uint)] * 6
    L.wr_bank_create.restype = vp
    L.wr_bank_create.argtypes = [i, u, u, u, u, u, u, u]
    L.wr_bank_destroy.argtypes = [vp]
    L.wr_bank_set_sintable.argtypes = [vp, _fp]
    L.wr_rx_set_stream.argtypes = [vp, u, u]
    L.wr_rx_set_phase_step.argtypes = [vp, u, C.c_int32]
    L.wr_rx_set_taps.argtypes = [vp, u, i, _fp, u]
    L.wr_rx_set_mode.argtypes = [vp, u, i]
    L.wr_bank_design_taps.argtypes = [vp, i, C.POINTER(C.c_uint), u]
    L.wr_rx_get_taps.argtypes = [vp, u, i, _fp, u]
    L.wr_rx_reset.argtypes = [vp, u, u]
    L.wr_rx_set_phase.argtypes = [vp, u, C.c_uint32]
    L.wr_rx_get_phase.argtypes = [vp, u, C.POINTER(C.c_uint32)]
    L.wr_rx_set_lookback.argtypes = [vp, u, C.POINTER(C.c_float)]
    L.wr_rx_get_lookback.argtypes = [vp, u, C.POINTER(C.c_float)]
    L.wr_bank_process.argtypes = [vp, vp, u, vp, sz]
    L.wr_bank_process_device.argtypes = [vp, vp, sz, u, vp, sz, vp]
    L.wr_bank_submit.argtypes = [vp, vp, u, vp, sz]
    L.wr_bank_wait.argtypes = [vp]
    L.wr_bank_pipeline_depth.argtypes = [vp]
    pp = C.POINTER(C.c_void_p)
    L.wr_bank_run_device_steps.argtypes = [vp, pp, u, sz, u, pp, u, sz, u, u]
    L.wr_bank_run_host_steps.argtypes = [vp, pp, u, u, pp, u, sz, u, u, i]
    L.wr_bank_process_u8.argtypes = L.wr_bank_process.argtypes
    L.wr_bank_process_device_u8.argtypes = L.wr_bank_process_device.argtypes
    L.wr_bank_submit_u8.argtypes = L.wr_bank_submit.argtypes
    L.wr_bank_run_device_steps_u8.argtypes = L.wr_bank_run_device_steps.argtypes
    L.wr_bank_run_host_steps_u8.argtypes = L.wr_bank_run_host_steps.argtypes
    L.wr_bank_stream.restype = vp
    L.wr_bank_stream.argtypes = [vp]
    L.wr_bank_sync.argtypes = [vp]
    L.wr_bank_keep_channel.argtypes = [vp, i]
    L.wr_bank_read_stage.restype = C.c_long
    L.wr_bank_read_stage.argtypes = [vp, u, i, _fp, sz]
    L.wr_bank_set_variant.argtypes = [vp, i]
    L.wr_bank_set_handover.argtypes = [vp, i]
    L.wr_bank_variant_in_use.argtypes = [vp]
    L.wr_bank_launch_count.restype = C.c_ulonglong
    L.wr_bank_launch_count.argtypes = [vp]
    L.wr_bank_kernel_times.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_ulonglong)]
    L.wr_bank_set_timing.argtypes = [vp, i]
    L.wr_stage_create.restype = vp
    L.wr_stage_create.argtypes = [i]
    L.wr_stage_destroy.argtypes = [vp]
    L.wr_stage_mix.argtypes = [vp, _fp, C.POINTER(C.c_uint32), C.c_int32, _fp, u, _fp]
    L.wr_stage_fir_config.argtypes = [vp, u, _fp, u]
    L.wr_stage_fir.argtypes = [vp, _fp, u, u, _fp]
    L.wr_stage_fir_reset.argtypes = [vp]
    L.wr_stage_demod.argtypes = [vp, i, _fp, _fp, u, _fp]
    L.wr_stage_atan2f.argtypes = [vp, _fp, _fp, u, _fp]
    L.wr_stage_palette.argtypes = [vp, _fp, u, vp]
    L.wr_spectrum_get_palette.argtypes = [vp, u, vp]
    L.wr_bank_set_audio_format.argtypes = [vp, i]
    L.wr_atan2f_host.argtypes = [_fp, _fp, sz, _fp]
    L.wr_atan2f_host.restype = None
    L.wr_spectrum_create.restype = vp
    L.wr_spectrum_create.argtypes = [i, u, u, u, u]
    L.wr_spectrum_destroy.argtypes = [vp]
    L.wr_spectrum_process.restype = C.c_long
    L.wr_spectrum_process.argtypes = [vp, vp, u, vp, sz]
    L.wr_spectrum_process_device.restype = C.c_long
    L.wr_spectrum_process_device.argtypes = [vp, vp, sz, u, vp, sz, vp]
    L.wr_spectrum_get.argtypes = [vp, u, _fp]
    L.wr_spectrum_launch_count.restype = C.c_ulonglong
    L.wr_spectrum_launch_count.argtypes = [vp]
    L.wr_spectrum_sync.argtypes = [vp]
    L.wr_upload_create.restype = vp
    L.wr_upload_create.argtypes = [i, sz]
    L.wr_upload_destroy.argtypes = [vp]
    L.wr_upload_capacity.restype = sz
    L.wr_upload_capacity.argtypes = [vp]
    L.wr_upload_device.argtypes = [vp]
    L.wr_upload_begin.argtypes = [vp, vp, u]
    L.wr_upload_finish.argtypes = [vp]
    L.wr_bank_process_upload.argtypes = [vp, vp, u, vp, sz]
    L.wr_spectrum_process_upload.restype = C.c_long
    L.wr_spectrum_process_upload.argtypes = [vp, vp, u]
    L.wr_spectrum_reserve.argtypes = [vp, u]
    L.wr_host_alloc.restype = vp
    L.wr_host_alloc.argtypes = [sz]
    L.wr_host_free.argtypes = [vp]
    L.wr_rx_get_history.argtypes = [vp, u, i, _fp, u]
    L.wr_rx_set_history.argtypes = [vp, u, i, _fp, u]
    _lib = L
    return L


def last_error():
    return lib().wr_last_error().decode(errors="replace")


def _check(rc, what):
    if rc < 0:
        raise WrError(f"{what} failed ({rc}): {last_error()}")
    return rc


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_fp)


def phase_step(if_hz, fs):
    return int(lib().wr_phase_step(int(if_hz), int(fs)))


def build_sintable():
    out = np.empty(65536, np.float32)
    lib().wr_build_sintable(out.ctypes.data_as(_fp))
    return out


def lowpass_design(n, passband, fs):
    out = np.empty(n, np.float32)
    _check(lib().wr_lowpass_design(n, passband, fs, out.ctypes.data_as(_fp)), "wr_lowpass_design")
    return out


class Bank:
    """n_receivers fused receiver chains fed by n_streams tuner streams on one GPU."""

    def __init__(self, n_streams, n_receivers, max_frames, n1, d1, n2, d2, device=0):
        self.L = lib()
        self.T, self.R, self.max_frames = n_streams, n_receivers, max_frames
        self.n1, self.d1, self.n2, self.d2 = n1, d1, n2, d2
        self.h = self.L.wr_bank_create(device, n_streams, n_receivers, max_frames, n1, d1, n2, d2)
        if not self.h:
            raise WrError("wr_bank_create failed: " + last_error())

    def close(self):
        if getattr(self, "h", None):
            self.L.wr_bank_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_sintable(self, table):
        t, p = _f32(table)
        _check(self.L.wr_bank_set_sintable(self.h, p), "wr_bank_set_sintable")

    def set_stream(self, rx, stream):
        _check(self.L.wr_rx_set_stream(self.h, rx, stream), "wr_rx_set_stream")

    def set_phase_step(self, rx, step):
        _check(self.L.wr_rx_set_phase_step(self.h, rx, step), "wr_rx_set_phase_step")

    def set_if(self, rx, if_hz, fs):
        self.set_phase_step(rx, phase_step(if_hz, fs))

    def set_taps(self, rx, stage, coeff):
        c, p = _f32(coeff)
        _check(self.L.wr_rx_set_taps(self.h, rx, stage, p, c.size), "wr_rx_set_taps")

    def design_taps(self, stage, passbands_hz, fs):
        """LowPass::setPassband for every receiver at once, designed on the device."""
        pb = np.ascontiguousarray(passbands_hz, dtype=np.uint32)
        assert pb.size == self.R
        _check(self.L.wr_bank_design_taps(self.h, stage, pb.ctypes.data_as(C.POINTER(C.c_uint)), int(fs)),
               "wr_bank_design_taps")

    def get_taps(self, rx, stage):
        n = self.n2 if stage else self.n1
        out = np.empty(n, np.float32)
        _check(self.L.wr_rx_get_taps(self.h, rx, stage, out.ctypes.data_as(_fp), n), "wr_rx_get_taps")
        return out

    def set_mode(self, rx, mode):
        _check(self.L.wr_rx_set_mode(self.h, rx, MODES.get(mode, mode)), "wr_rx_set_mode")

    def reset(self, rx, flags):
        _check(self.L.wr_rx_reset(self.h, rx, flags), "wr_rx_reset")

    def set_phase(self, rx, phase):
        _check(self.L.wr_rx_set_phase(self.h, rx, phase), "wr_rx_set_phase")

    def get_phase(self, rx):
        v = C.c_uint32(0)
        _check(self.L.wr_rx_get_phase(self.h, rx, C.byref(v)), "wr_rx_get_phase")
        return v.value

    def set_lookback(self, rx, prev_iq):
        """The FM discriminator's look-back sample (prev_i, prev_q); applies at the next block."""
        v = (C.c_float * 2)(float(prev_iq[0]), float(prev_iq[1]))
        _check(self.L.wr_rx_set_lookback(self.h, rx, v), "wr_rx_set_lookback")

    def get_history(self, rx, stage):
        """The FIR history a receiver carries: stage 0 the last n1-1 mixed IQ frames, stage 1 the last n2-1 demodulated samples."""
        n = (self.n2 - 1) if stage else 2 * (self.n1 - 1)
        out = np.empty(max(n, 1), np.float32)
        _check(self.L.wr_rx_get_history(self.h, rx, stage, out.ctypes.data_as(_fp), n), "wr_rx_get_history")
        return out[:n]

    def set_history(self, rx, stage, hist):
        a, p = _f32(hist)
        _check(self.L.wr_rx_set_history(self.h, rx, stage, p, a.size), "wr_rx_set_history")

    def process_upload(self, upload, nframes):
        m2 = self.out_frames(nframes)
        out = np.zeros((self.R, max(m2, 1)), np.float32)
        _check(self.L.wr_bank_process_upload(self.h, upload.h, nframes, out.ctypes.data, out.shape[1]), "wr_bank_process_upload")
        return out[:, :m2]

    def get_lookback(self, rx):
        v = (C.c_float * 2)()
        _check(self.L.wr_rx_get_lookback(self.h, rx, v), "wr_rx_get_lookback")
        return np.array([v[0], v[1]], np.float32)

    def out_frames(self, nframes):
        return nframes // self.d1 // self.d2

    def process(self, iq, nframes=None):
        """iq: float32 [n_streams, nframes, 2] (or flat).  Returns audio [n_receivers, M2]."""
        a, _ = _f32(iq)
        if nframes is None:
            nframes = a.size // (2 * self.T)
        assert a.size == 2 * self.T * nframes
        m2 = self.out_frames(nframes)
        out = np.zeros((self.R, max(m2, 1)), np.float32)
        _check(self.L.wr_bank_process(self.h, a.ctypes.data, nframes, out.ctypes.data, out.shape[1]),
               "wr_bank_process")
        return out[:, :m2]

    def process_u8(self, iq_u8, nframes=None):
        """iq_u8: uint8 [n_streams, nframes, 2], raw RTL-SDR bytes (reference rtlsdrtuner.cxx:104-108)."""
        a = np.ascontiguousarray(iq_u8, dtype=np.uint8)
        if nframes is None:
            nframes = a.size // (2 * self.T)
        assert a.size == 2 * self.T * nframes
        m2 = self.out_frames(nframes)
        out = np.zeros((self.R, max(m2, 1)), np.float32)
        _check(self.L.wr_bank_process_u8(self.h, a.ctypes.data, nframes, out.ctypes.data, out.shape[1]),
               "wr_bank_process_u8")
        return out[:, :m2]

    def process_device(self, iq_ptr, stream_stride, nframes, audio_ptr, audio_stride, cuda_stream=None, u8=False):
        f = self.L.wr_bank_process_device_u8 if u8 else self.L.wr_bank_process_device
        _check(f(self.h, iq_ptr, stream_stride, nframes, audio_ptr, audio_stride, cuda_stream), "wr_bank_process_device")

    def submit(self, iq_ptr, nframes, audio_ptr, audio_stride, u8=False):
        f = self.L.wr_bank_submit_u8 if u8 else self.L.wr_bank_submit
        _check(f(self.h, iq_ptr, nframes, audio_ptr, audio_stride), "wr_bank_submit")

    def wait(self):
        _check(self.L.wr_bank_wait(self.h), "wr_bank_wait")

    def run_device_steps(self, iq_ptrs, stride, nframes, audio_ptrs, audio_stride, first, steps, u8=False):
        a = (C.c_void_p * len(iq_ptrs))(*iq_ptrs)
        o = (C.c_void_p * len(audio_ptrs))(*audio_ptrs)
        f = self.L.wr_bank_run_device_steps_u8 if u8 else self.L.wr_bank_run_device_steps
        _check(f(self.h, a, len(iq_ptrs), stride, nframes, o, len(audio_ptrs), audio_stride, first, steps),
               "wr_bank_run_device_steps")

    def run_host_steps(self, iq_ptrs, nframes, audio_ptrs, audio_stride, first, steps, pipelined=True, u8=False):
        a = (C.c_void_p * len(iq_ptrs))(*iq_ptrs)
        o = (C.c_void_p * len(audio_ptrs))(*audio_ptrs)
        f = self.L.wr_bank_run_host_steps_u8 if u8 else self.L.wr_bank_run_host_steps
        _check(f(self.h, a, len(iq_ptrs), nframes, o, len(audio_ptrs), audio_stride, first, steps, int(pipelined)),
               "wr_bank_run_host_steps")

    def pipeline_depth(self):
        return self.L.wr_bank_pipeline_depth(self.h)

    def stream(self):
        return self.L.wr_bank_stream(self.h)

    def sync(self):
        _check(self.L.wr_bank_sync(self.h), "wr_bank_sync")

    def keep_channel(self, keep=True):
        _check(self.L.wr_bank_keep_channel(self.h, int(keep)), "wr_bank_keep_channel")

    def read_stage(self, rx, stage, nframes):
        m1 = nframes // self.d1
        n = 2 * m1 if stage == STAGE_CHANNEL else m1
        out = np.empty(max(n, 1), np.float32)
        got = _check(self.L.wr_bank_read_stage(self.h, rx, stage, out.ctypes.data_as(_fp), n), "wr_bank_read_stage")
        return out[:got]

    def set_audio_format(self, fmt):
        _check(self.L.wr_bank_set_audio_format(self.h, fmt), "wr_bank_set_audio_format")

    def set_handover(self, scheme):
        """0 = CUDA events, 1 = counter in HBM in / event out (default), 2 = counter in / audio stored by the kernel into the pinned buffer."""
        return _check(self.L.wr_bank_set_handover(self.h, scheme), "wr_bank_set_handover")

    def set_variant(self, v):
        _check(self.L.wr_bank_set_variant(self.h, v), "wr_bank_set_variant")

    def variant_in_use(self):
        return int(self.L.wr_bank_variant_in_use(self.h))

    def launch_count(self):
        return int(self.L.wr_bank_launch_count(self.h))

    def set_timing(self, on=True):
        _check(self.L.wr_bank_set_timing(self.h, int(on)), "wr_bank_set_timing")

    def kernel_times(self):
        """(total ms in the fused channel kernel, total ms in the audio kernel, blocks timed)."""
        ms = (C.c_double * 2)()
        n = C.c_ulonglong(0)
        _check(self.L.wr_bank_kernel_times(self.h, ms, C.byref(n)), "wr_bank_kernel_times")
        return float(ms[0]), float(ms[1]), int(n.value)


class Stage:
    """Strict single-stage blocks (one kernel per call, host buffers)."""

    def __init__(self, device=0):
        self.L = lib()
        self.h = self.L.wr_stage_create(device)
        if not self.h:
            raise WrError("wr_stage_create failed: " + last_error())

    def close(self):
        if getattr(self, "h", None):
            self.L.wr_stage_destroy(self.h)
            self.h = None

    __del__ = close

    def mix(self, phase, step, iq, table=None):
        a, ap = _f32(iq)
        out = np.empty_like(a)
        ph = C.c_uint32(phase)
        tp = None
        if table is not None:
            t, tp = _f32(table)
        _check(self.L.wr_stage_mix(self.h, tp, C.byref(ph), step, ap, a.size // 2, out.ctypes.data_as(_fp)),
               "wr_stage_mix")
        return out, ph.value

    def fir_config(self, channels, coeff):
        c, p = _f32(coeff)
        self.ch = channels
        _check(self.L.wr_stage_fir_config(self.h, channels, p, c.size), "wr_stage_fir_config")

    def fir(self, x, decim):
        a, ap = _f32(x)
        nframes = a.size // self.ch
        out = np.empty(max(1, (nframes // decim) * self.ch), np.float32)
        _check(self.L.wr_stage_fir(self.h, ap, nframes, decim, out.ctypes.data_as(_fp)), "wr_stage_fir")
        return out[:(nframes // decim) * self.ch]

    def fir_reset(self):
        _check(self.L.wr_stage_fir_reset(self.h), "wr_stage_fir_reset")

    def demod(self, mode, prev, iq):
        a, ap = _f32(iq)
        out = np.empty(max(1, a.size // 2), np.float32)
        assert prev.dtype == np.float32 and prev.size == 2
        _check(self.L.wr_stage_demod(self.h, MODES.get(mode, mode), prev.ctypes.data_as(_fp), ap, a.size // 2,
                                     out.ctypes.data_as(_fp)), "wr_stage_demod")
        return out[:a.size // 2]

    def palette(self, db):
        a, ap = _f32(db)
        out = np.empty(max(1, a.size), np.uint8)
        _check(self.L.wr_stage_palette(self.h, ap, a.size, out.ctypes.data), "wr_stage_palette")
        return out[:a.size]

    def atan2f(self, y, x):
        ya, yp = _f32(y)
        xa, xp = _f32(x)
        assert ya.size == xa.size
        out = np.empty(max(1, ya.size), np.float32)
        _check(self.L.wr_stage_atan2f(self.h, yp, xp, ya.size, out.ctypes.data_as(_fp)), "wr_stage_atan2f")
        return out[:ya.size]


def atan2f_host(y, x):
    """Host twin of the kernels' atan2f (restatement of glibc's, webradio_b200/csrc/wr_atan2f.h)."""
    ya, yp = _f32(y)
    xa, xp = _f32(x)
    assert ya.size == xa.size
    out = np.empty(ya.size, np.float32)
    lib().wr_atan2f_host(yp, xp, ya.size, out.ctypes.data_as(_fp))
    return out


class Upload:
    """One tuner block carried to the device once and shared by a bank and a spectrum sink."""

    def __init__(self, max_frames, device=0):
        self.L = lib()
        self.h = self.L.wr_upload_create(device, max_frames)
        if not self.h:
            raise WrError("wr_upload_create failed: " + last_error())
        self.keep = None

    def close(self):
        if getattr(self, "h", None):
            self.L.wr_upload_destroy(self.h)
            self.h = None

    __del__ = close

    def begin(self, iq):
        a, _ = _f32(iq)
        self.keep = a            # the host buffer must outlive the copies
        _check(self.L.wr_upload_begin(self.h, a.ctypes.data, a.size // 2), "wr_upload_begin")
        return a.size // 2

    def finish(self):
        _check(self.L.wr_upload_finish(self.h), "wr_upload_finish")


class Spectrum:
    def __init__(self, fft_size, hop=None, n_streams=1, max_frames=None, device=0):
        self.L = lib()
        self.N, self.hop, self.T = fft_size, hop or fft_size, n_streams
        self.max_frames = max_frames or 64 * fft_size
        self.h = self.L.wr_spectrum_create(device, fft_size, self.hop, n_streams, self.max_frames)
        if not self.h:
            raise WrError("wr_spectrum_create failed: " + last_error())

    def close(self):
        if getattr(self, "h", None):
            self.L.wr_spectrum_destroy(self.h)
            self.h = None

    __del__ = close

    def process(self, iq, rows=True):
        """iq: [n_streams, nframes, 2].  Returns rows [n_streams, nrows, N] (or nrows if rows=False)."""
        a, _ = _f32(iq)
        nframes = a.size // (2 * self.T)
        max_rows = (nframes + self.N) // self.hop + 1
        out = np.empty((self.T, max_rows, self.N), np.float32) if rows else None
        n = _check(self.L.wr_spectrum_process(self.h, a.ctypes.data, nframes,
                                              out.ctypes.data if rows else None, max_rows * self.N),
                   "wr_spectrum_process")
        return out[:, :n, :] if rows else n

    def process_device(self, iq_ptr, stride, nframes, rows_ptr, row_stride, cuda_stream=None):
        return _check(self.L.wr_spectrum_process_device(self.h, iq_ptr, stride, nframes, rows_ptr, row_stride,
                                                        cuda_stream), "wr_spectrum_process_device")

    def process_upload(self, upload, nframes):
        return _check(self.L.wr_spectrum_process_upload(self.h, upload.h, nframes), "wr_spectrum_process_upload")

    def reserve(self, max_frames):
        _check(self.L.wr_spectrum_reserve(self.h, max_frames), "wr_spectrum_reserve")
        self.max_frames = max(self.max_frames, max_frames)

    def get(self, stream=0):
        out = np.empty(self.N, np.float32)
        _check(self.L.wr_spectrum_get(self.h, stream, out.ctypes.data_as(_fp)), "wr_spectrum_get")
        return out

    def get_palette(self, stream=0):
        """The latest row as the browser's 256-entry palette index (waterfall.js:92-109)."""
        out = np.empty(self.N, np.uint8)
        _check(self.L.wr_spectrum_get_palette(self.h, stream, out.ctypes.data), "wr_spectrum_get_palette")
        return out

    def launch_count(self):
        return int(self.L.wr_spectrum_launch_count(self.h))

    def sync(self):
        _check(self.L.wr_spectrum_sync(self.h), "wr_spectrum_sync")
