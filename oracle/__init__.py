"""ctypes front end of the CPU ORACLE (test infrastructure, NOT product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this package.  It wraps oracle/libwam_oracle.so, the float64 C restatement of the reference's
FSKCore / filters / RingBuffer / CRC16 / XModem packet path (see wam_oracle.h for the file:line map).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libwam_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "wam_oracle.c")
    hdr = os.path.join(_HERE, "wam_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(_LIB_PATH) for f in (src, hdr)
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libwam_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class FSKConfigStruct(C.Structure):
    """Same layout as wamo_fsk_config / wam_fsk_config."""

    _fields_ = [
        ("sampleRate", C.c_double),
        ("baudRate", C.c_double),
        ("markFrequency", C.c_double),
        ("spaceFrequency", C.c_double),
        ("preamblePattern", C.POINTER(C.c_uint8)),
        ("preambleLength", C.c_int32),
        ("sfdPattern", C.POINTER(C.c_uint8)),
        ("sfdLength", C.c_int32),
        ("startBits", C.c_int32),
        ("stopBits", C.c_int32),
        ("parity", C.c_int32),
        ("syncThreshold", C.c_double),
        ("agcEnabled", C.c_int32),
        ("preFilterBandwidth", C.c_double),
        ("adaptiveThreshold", C.c_int32),
    ]


class StatusStruct(C.Structure):
    _fields_ = [
        ("ready", C.c_int32),
        ("frameStarted", C.c_int32),
        ("globalSampleCounter", C.c_double),
        ("receivedBitsLength", C.c_double),
        ("byteBufferLength", C.c_double),
        ("demodulationCalls", C.c_double),
        ("syncDetections", C.c_double),
        ("silenceThreshold", C.c_double),
        ("totalSamplesProcessed", C.c_double),
        ("eodEvents", C.c_double),
        ("errorEvents", C.c_double),
        ("configuredEvents", C.c_double),
    ]


class XmodemRxState(C.Structure):
    _fields_ = [("expectedSequence", C.c_int32), ("retries", C.c_int32), ("done", C.c_int32),
                ("dataLen", C.c_int32), ("packetsReceived", C.c_int32), ("packetsDropped", C.c_int32)]


class PktResult(C.Structure):
    _fields_ = [
        ("status", C.c_int32),
        ("sequence", C.c_int32),
        ("length", C.c_int32),
        ("payloadOffset", C.c_int32),
        ("crcReceived", C.c_int32),
        ("crcComputed", C.c_int32),
        ("bytesConsumed", C.c_int32),
    ]


DEFAULT_FSK_CONFIG = dict(  # src/modems/fsk.ts:19-33
    sampleRate=48000,
    baudRate=1200,
    markFrequency=1650,
    spaceFrequency=1850,
    preamblePattern=[0x55, 0x55],
    sfdPattern=[0x7E],
    startBits=1,
    stopBits=1,
    parity="none",
    syncThreshold=0.85,
    agcEnabled=True,
    preFilterBandwidth=800,
    adaptiveThreshold=True,
)
_PARITY = {"none": 0, "even": 1, "odd": 2}


def make_config_struct(cfg: dict):
    """dict (FSKConfig field names) → (struct, keepalive)."""
    full = {**DEFAULT_FSK_CONFIG, **cfg}
    pre = (C.c_uint8 * max(1, len(full["preamblePattern"])))(*[b & 0xFF for b in full["preamblePattern"]])
    sfd = (C.c_uint8 * max(1, len(full["sfdPattern"])))(*[b & 0xFF for b in full["sfdPattern"]])
    s = FSKConfigStruct(
        float(full["sampleRate"]), float(full["baudRate"]), float(full["markFrequency"]),
        float(full["spaceFrequency"]), C.cast(pre, C.POINTER(C.c_uint8)), len(full["preamblePattern"]),
        C.cast(sfd, C.POINTER(C.c_uint8)), len(full["sfdPattern"]), int(full["startBits"]),
        int(full["stopBits"]), _PARITY[full["parity"]], float(full["syncThreshold"]),
        1 if full["agcEnabled"] else 0, float(full["preFilterBandwidth"]), 1 if full["adaptiveThreshold"] else 0,
    )
    return s, (pre, sfd)


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    vp, dp, fp, u8p, i32p = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_int32)
    L.wamo_fsk_new.restype = vp
    L.wamo_fsk_free.argtypes = [vp]
    L.wamo_fsk_configure.argtypes = [vp, C.POINTER(FSKConfigStruct)]
    L.wamo_fsk_modulate.restype = C.c_long
    L.wamo_fsk_modulate.argtypes = [vp, u8p, C.c_long, fp, C.c_long]
    L.wamo_fsk_demodulate.restype = C.c_long
    L.wamo_fsk_demodulate.argtypes = [vp, fp, C.c_long, u8p, C.c_long]
    L.wamo_fsk_reset.argtypes = [vp]
    L.wamo_fsk_status_get.argtypes = [vp, C.POINTER(StatusStruct)]
    L.wamo_fsk_params.argtypes = [vp, dp]
    L.wamo_fsk_set_prefilter_tap.argtypes = [vp, fp, C.c_long]
    L.wamo_fsk_set_decim_tap.argtypes = [vp, dp, dp, C.c_long]
    L.wamo_fsk_decim_tap_count.restype = C.c_long
    L.wamo_fsk_decim_tap_count.argtypes = [vp]
    L.wamo_iir_new.restype = vp
    L.wamo_iir_new.argtypes = [dp, C.c_int, dp, C.c_int, C.POINTER(C.c_int)]
    L.wamo_iir_free.argtypes = [vp]
    L.wamo_iir_process.restype = C.c_double
    L.wamo_iir_process.argtypes = [vp, C.c_double]
    L.wamo_iir_process_buffer.argtypes = [vp, fp, fp, C.c_long]
    L.wamo_iir_reset.argtypes = [vp]
    L.wamo_iir_coefficients.restype = C.c_int
    L.wamo_iir_coefficients.argtypes = [vp, dp, dp]
    L.wamo_fir_new.restype = vp
    L.wamo_fir_new.argtypes = [dp, C.c_int]
    L.wamo_fir_free.argtypes = [vp]
    L.wamo_fir_process.restype = C.c_double
    L.wamo_fir_process.argtypes = [vp, C.c_double]
    L.wamo_fir_process_buffer.argtypes = [vp, fp, fp, C.c_long]
    L.wamo_fir_reset.argtypes = [vp]
    for n in ("lowpass", "highpass"):
        getattr(L, f"wamo_design_butterworth_{n}").argtypes = [C.c_double, C.c_double, dp, dp]
        f = getattr(L, f"wamo_design_sinc_{n}")
        f.restype = C.c_int
        f.argtypes = [C.c_double, C.c_double, C.c_int, dp]
    L.wamo_design_butterworth_bandpass.argtypes = [C.c_double, C.c_double, C.c_double, dp, dp]
    L.wamo_design_sinc_bandpass.restype = C.c_int
    L.wamo_design_sinc_bandpass.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, dp]
    L.wamo_ring_new.restype = vp
    L.wamo_ring_new.argtypes = [C.c_int, C.c_double]
    L.wamo_ring_free.argtypes = [vp]
    L.wamo_ring_put.argtypes = [vp, C.c_double]
    L.wamo_ring_get.restype = C.c_int
    L.wamo_ring_get.argtypes = [vp, C.c_double, dp]
    L.wamo_ring_length.restype = C.c_double
    L.wamo_ring_length.argtypes = [vp]
    L.wamo_ring_clear.argtypes = [vp]
    L.wamo_crc16.restype = C.c_uint16
    L.wamo_crc16.argtypes = [u8p, C.c_long]
    L.wamo_xmodem_serialize.restype = C.c_long
    L.wamo_xmodem_serialize.argtypes = [C.c_int, u8p, C.c_long, u8p, C.c_long]
    L.wamo_xmodem_check.argtypes = [u8p, C.c_long, C.c_int, C.POINTER(PktResult)]
    L.wamo_xmodem_receive.restype = C.c_long
    L.wamo_xmodem_receive.argtypes = [u8p, C.c_long, C.c_int, C.POINTER(XmodemRxState), u8p, C.c_int, i32p, u8p, C.c_long]
    L.wamo_fsk_batch_demodulate.restype = C.c_int
    L.wamo_fsk_batch_demodulate.argtypes = [C.POINTER(FSKConfigStruct), i32p, C.c_long, fp, C.c_long, C.c_long,
                                            u8p, C.c_long, i32p, C.POINTER(StatusStruct), C.c_int]
    _lib = L
    return L


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _f32(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f64(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class NotConfigured(RuntimeError):
    pass


class FSKCore:
    """Oracle twin of the reference FSKCore (src/modems/fsk.ts:82-494)."""

    def __init__(self):
        self._h = lib().wamo_fsk_new()
        self._keep = None
        self._cfg = None

    def __del__(self):
        try:
            if self._h:
                lib().wamo_fsk_free(self._h)
                self._h = None
        except Exception:
            pass

    def configure(self, cfg: dict | None = None):
        self._cfg = {**DEFAULT_FSK_CONFIG, **(cfg or {})}
        s, self._keep = make_config_struct(self._cfg)
        lib().wamo_fsk_configure(self._h, C.byref(s))

    def getConfig(self):
        return dict(self._cfg)

    def modulateData(self, data) -> np.ndarray:
        data = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
        n = lib().wamo_fsk_modulate(self._h, _u8(data) if len(data) else None, len(data), None, 0)
        if n < 0:
            raise NotConfigured("FSK modulator not configured")
        out = np.zeros(n, dtype=np.float32)
        lib().wamo_fsk_modulate(self._h, _u8(data) if len(data) else None, len(data), _f32(out), n)
        return out

    def demodulateData(self, samples: np.ndarray, tap: np.ndarray | None = None) -> bytes:
        """In-place on `samples` (float32, C-contiguous) when AGC is on, like the reference."""
        assert samples.dtype == np.float32 and samples.flags.c_contiguous
        cap = len(samples) // 8 + 16
        out = np.zeros(cap, dtype=np.uint8)
        if tap is not None:
            lib().wamo_fsk_set_prefilter_tap(self._h, _f32(tap), len(tap))
        n = lib().wamo_fsk_demodulate(self._h, _f32(samples), len(samples), _u8(out), cap)
        if tap is not None:
            lib().wamo_fsk_set_prefilter_tap(self._h, None, 0)
        if n < 0:
            raise NotConfigured("FSK demodulator not configured")
        return bytes(out[:n])

    def demodulateTapped(self, samples: np.ndarray):
        """demodulateData plus the decimated-rate internals: (bytes, filteredPhaseDiff[k], amplitude[k]) — fsk.ts:246-264."""
        assert samples.dtype == np.float32 and samples.flags.c_contiguous
        nd = len(samples) // 2 + 2
        f, a = np.zeros(nd), np.zeros(nd)
        lib().wamo_fsk_set_decim_tap(self._h, _f64(f), _f64(a), nd)
        out = self.demodulateData(samples)
        n = lib().wamo_fsk_decim_tap_count(self._h)
        lib().wamo_fsk_set_decim_tap(self._h, None, None, 0)
        return out, f[:n], a[:n]

    def reset(self):
        lib().wamo_fsk_reset(self._h)

    def getStatus(self) -> dict:
        st = StatusStruct()
        lib().wamo_fsk_status_get(self._h, C.byref(st))
        return {k: getattr(st, k) for k, _ in StatusStruct._fields_}

    def params(self) -> dict:
        out = np.zeros(8)
        lib().wamo_fsk_params(self._h, _f64(out))
        keys = ["samplesPerBit", "downsampledSamplesPerBit", "bitsPerByte", "nbits", "centerFreq",
                "syncRingCapacity", "ampRingCapacity", "samplesForEOD"]
        return dict(zip(keys, out.tolist()))


class IIRFilter:
    ERRORS = {1: "Feedforward coefficients (b) cannot be empty",
              2: "Feedback coefficients (a) cannot be empty",
              3: "First feedback coefficient (a[0]) cannot be zero"}

    def __init__(self, b, a):
        b = np.asarray(b, dtype=np.float64)
        a = np.asarray(a, dtype=np.float64)
        err = C.c_int(0)
        self._h = lib().wamo_iir_new(_f64(b) if len(b) else None, len(b), _f64(a) if len(a) else None, len(a), C.byref(err))
        if not self._h:
            raise ValueError(self.ERRORS[err.value])
        self._nb, self._na = len(b), len(a)

    def __del__(self):
        try:
            if self._h:
                lib().wamo_iir_free(self._h)
        except Exception:
            pass

    def process(self, x: float) -> float:
        return lib().wamo_iir_process(self._h, float(x))

    def processBuffer(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.zeros_like(x)
        lib().wamo_iir_process_buffer(self._h, _f32(x), _f32(out), len(x))
        return out

    def reset(self):
        lib().wamo_iir_reset(self._h)

    def getCoefficients(self):
        b = np.zeros(self._nb)
        a = np.zeros(self._na)
        lib().wamo_iir_coefficients(self._h, _f64(b), _f64(a))
        return {"b": b, "a": a}


class FIRFilter:
    def __init__(self, taps):
        self._taps = np.asarray(taps, dtype=np.float64).copy()
        self._h = lib().wamo_fir_new(_f64(self._taps) if len(self._taps) else None, len(self._taps))

    def __del__(self):
        try:
            if self._h:
                lib().wamo_fir_free(self._h)
        except Exception:
            pass

    def process(self, x: float) -> float:
        return lib().wamo_fir_process(self._h, float(x))

    def processBuffer(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.zeros_like(x)
        lib().wamo_fir_process_buffer(self._h, _f32(x), _f32(out), len(x))
        return out

    def reset(self):
        lib().wamo_fir_reset(self._h)

    def getCoefficients(self):
        return self._taps.copy()


class FilterDesign:
    @staticmethod
    def butterworthLowpass(fc, fs):
        b, a = np.zeros(3), np.zeros(3)
        lib().wamo_design_butterworth_lowpass(fc, fs, _f64(b), _f64(a))
        return {"b": b, "a": a}

    @staticmethod
    def butterworthHighpass(fc, fs):
        b, a = np.zeros(3), np.zeros(3)
        lib().wamo_design_butterworth_highpass(fc, fs, _f64(b), _f64(a))
        return {"b": b, "a": a}

    @staticmethod
    def butterworthBandpass(f0, bw, fs):
        b, a = np.zeros(3), np.zeros(3)
        lib().wamo_design_butterworth_bandpass(f0, bw, fs, _f64(b), _f64(a))
        return {"b": b, "a": a}

    @staticmethod
    def sincLowpass(fc, fs, numTaps):
        out = np.zeros(numTaps + 2)
        n = lib().wamo_design_sinc_lowpass(fc, fs, numTaps, _f64(out))
        return out[:n].copy()

    @staticmethod
    def sincHighpass(fc, fs, numTaps):
        out = np.zeros(numTaps + 2)
        n = lib().wamo_design_sinc_highpass(fc, fs, numTaps, _f64(out))
        return out[:n].copy()

    @staticmethod
    def sincBandpass(f0, bw, fs, numTaps):
        out = np.zeros(numTaps + 2)
        n = lib().wamo_design_sinc_bandpass(f0, bw, fs, numTaps, _f64(out))
        return out[:n].copy()


class RingBuffer:
    def __init__(self, kind: str, size: float):
        self._h = lib().wamo_ring_new({"u8": 0, "f32": 1}[kind], float(size))

    def __del__(self):
        try:
            if self._h:
                lib().wamo_ring_free(self._h)
        except Exception:
            pass

    def put(self, *values):
        for v in values:
            lib().wamo_ring_put(self._h, float(v))

    def get(self, index):
        v = C.c_double(0)
        st = lib().wamo_ring_get(self._h, float(index), C.byref(v))
        if st < 0:
            raise IndexError("Index out of bounds")
        return None if st == 1 else v.value

    @property
    def length(self):
        return lib().wamo_ring_length(self._h)

    def clear(self):
        lib().wamo_ring_clear(self._h)


def crc16(data) -> int:
    a = np.frombuffer(bytes(data), dtype=np.uint8)
    return int(lib().wamo_crc16(_u8(a) if len(a) else None, len(a)))


def xmodem_serialize(sequence: int, payload) -> bytes:
    p = np.frombuffer(bytes(payload), dtype=np.uint8)
    out = np.zeros(len(p) + 6, dtype=np.uint8)
    n = lib().wamo_xmodem_serialize(sequence, _u8(p) if len(p) else None, len(p), _u8(out), len(out))
    if n == -1:
        raise ValueError(f"Invalid sequence: {sequence}. Must be 1-255.")
    if n == -2:
        raise ValueError(f"Payload too large: {len(p)}. Max 255 bytes.")
    return bytes(out[:n])


PKT_STATUS = ["OK", "DUPLICATE", "NO_SOH", "EOT", "INCOMPLETE", "BAD_COMPLEMENT", "BAD_CRC", "UNEXPECTED_SEQ"]


def xmodem_check(data, expected_sequence: int = 1) -> dict:
    a = np.frombuffer(bytes(data), dtype=np.uint8)
    r = PktResult()
    lib().wamo_xmodem_check(_u8(a) if len(a) else None, len(a), expected_sequence, C.byref(r))
    return {k: getattr(r, k) for k, _ in PktResult._fields_}


RX_STATE_FIELDS = [k for k, _ in XmodemRxState._fields_]


def xmodem_receive(data, state: dict | None = None, max_retries: int = 10, reply_cap: int = 64, data_cap: int = 1 << 16):
    """One burst through the receive side of XModemTransport (xmodem.ts:232-321).
    Returns (new_state dict, replies bytes, n_replies, consumed, payload bytes appended by this burst)."""
    a = np.frombuffer(bytes(data), dtype=np.uint8)
    st = XmodemRxState(1, 0, 0, 0, 0, 0)
    if state is not None:
        for k in RX_STATE_FIELDS:
            setattr(st, k, int(state[k]))
    base = st.dataLen
    replies = np.zeros(max(reply_cap, 1), dtype=np.uint8)
    out = np.zeros(base + data_cap, dtype=np.uint8)
    nrep = C.c_int32(0)
    consumed = lib().wamo_xmodem_receive(_u8(a) if len(a) else None, len(a), max_retries, C.byref(st), _u8(replies),
                                         reply_cap, C.byref(nrep), _u8(out), len(out))
    new = {k: getattr(st, k) for k in RX_STATE_FIELDS}
    return new, bytes(replies[: min(nrep.value, reply_cap)]), nrep.value, int(consumed), bytes(out[base: st.dataLen])


def batch_demodulate(cfgs: list[dict], cfg_index, samples: np.ndarray, n_threads: int = 1, want_status=True):
    """samples: float32 [n_streams, n_samples], mutated in place (AGC).  Returns (list[bytes], status list)."""
    assert samples.dtype == np.float32 and samples.flags.c_contiguous and samples.ndim == 2
    ns, n = samples.shape
    structs = (FSKConfigStruct * len(cfgs))()
    keep = []
    for i, c in enumerate(cfgs):
        s, k = make_config_struct(c)
        structs[i] = s
        keep.append(k)
    idx = np.ascontiguousarray(cfg_index, dtype=np.int32) if cfg_index is not None else None
    cap = n // 8 + 16
    out = np.zeros((ns, cap), dtype=np.uint8)
    out_len = np.zeros(ns, dtype=np.int32)
    st = (StatusStruct * ns)() if want_status else None
    lib().wamo_fsk_batch_demodulate(structs, idx.ctypes.data_as(C.POINTER(C.c_int32)) if idx is not None else None,
                                    ns, _f32(samples), samples.strides[0] // 4, n, _u8(out), cap,
                                    out_len.ctypes.data_as(C.POINTER(C.c_int32)), st, n_threads)
    res = [bytes(out[i, : out_len[i]]) for i in range(ns)]
    status = [{k: getattr(st[i], k) for k, _ in StatusStruct._fields_} for i in range(ns)] if want_status else None
    return res, status
