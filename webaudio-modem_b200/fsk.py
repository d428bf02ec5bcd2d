"""Host-side mirror of the reference's modulator interface for the FSK path.

`FSKCore` keeps the reference's IModulator surface (src/core.ts:88-117; src/modems/fsk.ts:82-494):
configure / getConfig / modulateData / demodulateData / reset / isReady / getStatus /
getSignalQuality / on / off / emit, the same FSKConfig field names and defaults, the same
"not configured" errors and the same 'configured' / 'eod' / 'error' events — but every sample is
processed by libwam.so on the GPU through the C ABI in include/wam.h.  `FSKBatch` is the new
batched entry point: thousands of independent streams with device-resident streaming state.

The reference's host language is TypeScript; no Node toolchain exists in this image, so this
Python class is the executable mirror used by the parity tests, and host/ holds the (unexecuted)
TypeScript + N-API binding described in INTEGRATION.md.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable

import numpy as np

from . import _lib as L

# src/modems/fsk.ts:19-33
DEFAULT_FSK_CONFIG = dict(
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
# README-only aliases (README.md:33-38) accepted for convenience; the code's names win.
_ALIASES = {"baud": "baudRate", "markFreq": "markFrequency", "spaceFreq": "spaceFrequency"}
_PARITY = {"none": 0, "even": 1, "odd": 2}


def normalize_config(cfg: dict | None) -> dict:
    cfg = dict(cfg or {})
    for k, v in list(cfg.items()):
        if k in _ALIASES:
            cfg.setdefault(_ALIASES[k], v)
            del cfg[k]
    return {**DEFAULT_FSK_CONFIG, **cfg}  # fsk.ts:134 — spread over defaults, no validation


def config_struct(cfg: dict):
    pre = (C.c_uint8 * max(1, len(cfg["preamblePattern"])))(*[int(b) & 0xFF for b in cfg["preamblePattern"]])
    sfd = (C.c_uint8 * max(1, len(cfg["sfdPattern"])))(*[int(b) & 0xFF for b in cfg["sfdPattern"]])
    s = L.FSKConfigStruct(
        float(cfg["sampleRate"]), float(cfg["baudRate"]), float(cfg["markFrequency"]), float(cfg["spaceFrequency"]),
        C.cast(pre, C.POINTER(C.c_uint8)), len(cfg["preamblePattern"]),
        C.cast(sfd, C.POINTER(C.c_uint8)), len(cfg["sfdPattern"]),
        int(cfg["startBits"]), int(cfg["stopBits"]), _PARITY[cfg["parity"]], float(cfg["syncThreshold"]),
        1 if cfg["agcEnabled"] else 0, float(cfg["preFilterBandwidth"]), 1 if cfg["adaptiveThreshold"] else 0,
    )
    return s, (pre, sfd)


def _status_dict(st: L.StatusStruct) -> dict:
    return {
        "ready": bool(st.ready),
        "frameStarted": bool(st.frameStarted),
        "globalSampleCounter": int(st.globalSampleCounter),
        "receivedBitsLength": st.receivedBitsLength,
        "byteBufferLength": int(st.byteBufferLength),
        "demodulationCalls": int(st.demodulationCalls),
        "syncDetections": int(st.syncDetections),
        "silenceThreshold": st.silenceThreshold,
        "totalSamplesProcessed": int(st.totalSamplesProcessed),
        "eodEvents": int(st.eodEvents),
        "errorEvents": int(st.errorEvents),  # device-side error flags (output overflow, pipeline time-out); 0 = none
    }


class EventEmitter:
    """src/core.ts EventEmitter: on / off / emit."""

    def __init__(self):
        self._listeners: dict[str, list[Callable]] = {}

    def on(self, name: str, cb: Callable):
        self._listeners.setdefault(name, []).append(cb)

    def off(self, name: str, cb: Callable | None = None):
        if cb is None:
            self._listeners.pop(name, None)
        elif name in self._listeners and cb in self._listeners[name]:
            self._listeners[name].remove(cb)

    def emit(self, name: str, event=None):
        for cb in list(self._listeners.get(name, [])):
            cb(event)


class FSKCore(EventEmitter):
    """GPU-backed drop-in for the reference FSKCore (one stream)."""

    name = "FSK"
    type = "FSK"

    def __init__(self, device: int = 0):
        super().__init__()
        self._device = device
        self._h = C.c_void_p()
        self._config = None
        self._keep = None
        self._eod_seen = 0
        self._lib = L.lib()
        # created lazily so that `new FSKCore()` never touches CUDA before configure()

    def _ensure(self):
        if not self._h:
            L.check(self._lib.wam_fsk_create(self._device, C.byref(self._h)))

    def dispose(self):
        if self._h:
            self._lib.wam_fsk_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.dispose()
        except Exception:
            pass

    # -- configuration (fsk.ts:133-157) -------------------------------------------------------
    def configure(self, config: dict | None = None):
        self._ensure()
        self._config = normalize_config(config)
        s, self._keep = config_struct(self._config)
        L.check(self._lib.wam_fsk_configure(self._h, C.byref(s)))
        self.emit("configured")

    def getConfig(self) -> dict:
        return dict(self._config) if self._config is not None else {}

    def isReady(self) -> bool:
        return bool(self._h) and bool(self._lib.wam_fsk_is_ready(self._h))

    # -- modulateData (fsk.ts:377-424) ---------------------------------------------------------
    def modulateData(self, data) -> np.ndarray:
        if not self.isReady():
            raise RuntimeError("FSK modulator not configured")
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        n = self._lib.wam_fsk_modulate_size(self._h, len(buf))
        L.check(n)
        out = np.zeros(n, dtype=np.float32)
        n_out = C.c_long(0)
        L.check(self._lib.wam_fsk_modulate(
            self._h, buf.ctypes.data_as(C.POINTER(C.c_uint8)) if len(buf) else None, len(buf),
            out.ctypes.data_as(C.POINTER(C.c_float)), n, C.byref(n_out)))
        return out

    # -- demodulateData (fsk.ts:190-222) --------------------------------------------------------
    def demodulateData(self, samples: np.ndarray) -> np.ndarray:
        """`samples` (float32, contiguous) is mutated in place when AGC is on, like the reference."""
        if not self.isReady():
            raise RuntimeError("FSK demodulator not configured")
        if not (isinstance(samples, np.ndarray) and samples.dtype == np.float32 and samples.flags.c_contiguous):
            raise TypeError("samples must be a C-contiguous float32 numpy array (Float32Array)")
        try:
            cap = max(16, len(samples) // 8 + 16)
            out = np.zeros(cap, dtype=np.uint8)
            n_out = C.c_long(0)
            L.check(self._lib.wam_fsk_demodulate(
                self._h, samples.ctypes.data_as(C.POINTER(C.c_float)), len(samples),
                out.ctypes.data_as(C.POINTER(C.c_uint8)), cap, C.byref(n_out)))
            st = self._status()
            new_eod = int(st.eodEvents) - self._eod_seen
            self._eod_seen = int(st.eodEvents)
            for _ in range(new_eod):
                self.emit("eod")  # fsk.ts:289
            return out[: n_out.value].copy()
        except L.WamError as e:  # fsk.ts:218-221: errors become an event and an empty result
            self.emit("error", {"data": e})
            return np.zeros(0, dtype=np.uint8)

    def reset(self):  # fsk.ts:464-469
        if self._h:
            L.check(self._lib.wam_fsk_reset(self._h))

    def _status(self) -> L.StatusStruct:
        st = L.StatusStruct()
        L.check(self._lib.wam_fsk_status_get(self._h, C.byref(st)))
        return st

    def getStatus(self) -> dict:  # fsk.ts:481-493
        if not self._h:
            return {"ready": False, "frameStarted": False, "globalSampleCounter": 0, "receivedBitsLength": 0,
                    "byteBufferLength": 0, "demodulationCalls": 0, "syncDetections": 0, "silenceThreshold": 0.01,
                    "totalSamplesProcessed": 0, "eodEvents": 0, "errorEvents": 0}
        return _status_dict(self._status())

    def getSignalQuality(self) -> dict:  # fsk.ts:471-479 — the reference returns zeros
        return {"snr": 0, "ber": 0, "eyeOpening": 0, "phaseJitter": 0, "frequencyOffset": 0}


class FSKBatch:
    """n_streams independent FSKCore instances on one GPU (wam_fsk_batch_*).

    configs: one dict (all streams) or a list of dicts plus `cfg_index[stream]`.
    """

    def __init__(self, n_streams: int, configs, cfg_index=None, device: int = 0):
        self._lib = L.lib()
        if isinstance(configs, dict) or configs is None:
            configs = [configs or {}]
        self.configs = [normalize_config(c) for c in configs]
        self.n_streams = int(n_streams)
        self.device = device
        structs = (L.FSKConfigStruct * len(self.configs))()
        self._keep = []
        for i, c in enumerate(self.configs):
            s, k = config_struct(c)
            structs[i] = s
            self._keep.append(k)
        idx = None
        if cfg_index is not None:
            idx = np.ascontiguousarray(cfg_index, dtype=np.int32)
            assert idx.shape == (self.n_streams,)
        self._h = C.c_void_p()
        L.check(self._lib.wam_fsk_batch_create(device, self.n_streams, structs, len(self.configs),
                                               idx.ctypes.data_as(C.POINTER(C.c_int32)) if idx is not None else None,
                                               C.byref(self._h)))

    def close(self):
        if self._h:
            self._lib.wam_fsk_batch_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def out_capacity(self, n_samples: int) -> int:
        return int(L.check(self._lib.wam_fsk_batch_out_capacity(self._h, n_samples)))

    def reset(self):
        L.check(self._lib.wam_fsk_batch_reset(self._h))

    def renew(self, stream: int = 0):
        """new FSKCore() + configure() on every stream (stream-ordered)."""
        L.check(self._lib.wam_fsk_batch_renew(self._h, stream or None))

    def launch_count(self) -> int:
        return int(self._lib.wam_fsk_batch_launch_count(self._h))

    def debug_fast_band(self, scale: float):
        """Test hook: widen the fast kernel's doubt band (wam_fsk_batch_debug_fast_band)."""
        L.check(self._lib.wam_fsk_batch_debug_fast_band(self._h, float(scale)))

    def fast_stats(self) -> dict:
        """Counters of the mixed-precision fast path (wam_fsk_batch_fast_stats)."""
        st = L.FastStats()
        L.check(self._lib.wam_fsk_batch_fast_stats(self._h, C.byref(st)))
        return {k: int(getattr(st, k)) for k, _ in L.FastStats._fields_}

    # -- host buffers --------------------------------------------------------------------------
    def demodulate(self, samples: np.ndarray, writeback_agc: bool = False, flags: int = 0):
        """samples float32 [n_streams, n]; returns (out uint8 [n_streams, cap], out_len int32 [n_streams])."""
        assert samples.dtype == np.float32 and samples.ndim == 2 and samples.shape[0] == self.n_streams
        n = samples.shape[1]
        assert n == 0 or samples.strides[1] == 4
        cap = self.out_capacity(n)
        out = np.zeros((self.n_streams, cap), dtype=np.uint8)
        out_len = np.zeros(self.n_streams, dtype=np.int32)
        L.check(self._lib.wam_fsk_batch_demodulate(
            self._h, samples.ctypes.data if n else None, max(samples.strides[0] // 4, n), n, out.ctypes.data, cap,
            out_len.ctypes.data,
            (L.WAM_BATCH_WRITEBACK_AGC if writeback_agc else 0) | flags))
        return out, out_len

    def debug_fast_windows(self, group: int = 0):
        """Verification windows of the last fast call: (n_classes, list of dict(cls, stream, slab, result)); the last two
        classes are the stages of an end-of-data check."""
        nc = C.c_int(0)
        cap = L.check(self._lib.wam_fsk_batch_debug_fast_windows(self._h, group, C.byref(nc), None, None, 0))
        counts = np.zeros(nc.value, dtype=np.int32)
        items = np.zeros(nc.value * cap * 3, dtype=np.int32)
        L.check(self._lib.wam_fsk_batch_debug_fast_windows(self._h, group, C.byref(nc), counts.ctypes.data, items.ctypes.data,
                                                           items.size))
        out = []
        for c in range(nc.value):
            for i in range(min(int(counts[c]), cap)):
                k = (c * cap + i) * 3
                out.append(dict(cls=c, stream=int(items[k]), slab=int(items[k + 1]), result=int(items[k + 2])))
        return nc.value, out

    def demodulate_pcm16(self, samples: np.ndarray, flags: int = 0):
        """samples int16 [n_streams, n] (16-bit PCM, value / 32768); same result as demodulate(samples / 32768)."""
        assert samples.dtype == np.int16 and samples.ndim == 2 and samples.shape[0] == self.n_streams
        n = samples.shape[1]
        assert n == 0 or samples.strides[1] == 2
        cap = self.out_capacity(n)
        out = np.zeros((self.n_streams, cap), dtype=np.uint8)
        out_len = np.zeros(self.n_streams, dtype=np.int32)
        L.check(self._lib.wam_fsk_batch_demodulate_pcm16(
            self._h, samples.ctypes.data if n else None, max(samples.strides[0] // 2, n), n, out.ctypes.data, cap,
            out_len.ctypes.data, flags))
        return out, out_len

    def demodulate_ragged(self, samples: np.ndarray, n_valid, flags: int = 0) -> list[bytes]:
        """samples float32 [n_streams, n_max]; stream s receives demodulateData(samples[s, :n_valid[s]]), a negative
        n_valid[s] means the stream is not called in this round.  Returns the bytes completed per stream."""
        assert samples.dtype == np.float32 and samples.ndim == 2 and samples.shape[0] == self.n_streams
        assert samples.strides[1] == 4 or samples.shape[1] == 0
        n = samples.shape[1]
        nv = np.ascontiguousarray(n_valid, dtype=np.int32)
        assert nv.shape == (self.n_streams,)
        cap = self.out_capacity(n)
        out = np.zeros((self.n_streams, cap), dtype=np.uint8)
        out_len = np.zeros(self.n_streams, dtype=np.int32)
        L.check(self._lib.wam_fsk_batch_demodulate_ragged(
            self._h, samples.ctypes.data if n else None, max(samples.strides[0] // 4, n), n, nv.ctypes.data, out.ctypes.data, cap,
            out_len.ctypes.data, flags))
        return [bytes(out[i, : out_len[i]]) for i in range(self.n_streams)]

    def demodulate_bytes(self, samples: np.ndarray, writeback_agc: bool = False, flags: int = 0) -> list[bytes]:
        out, out_len = self.demodulate(samples, writeback_agc, flags)
        return [bytes(out[i, : out_len[i]]) for i in range(self.n_streams)]

    # -- device buffers (raw pointers, e.g. torch tensors' data_ptr()) --------------------------
    def demodulate_device(self, d_samples: int, stride: int, n: int, d_out: int, out_stride: int, d_out_len: int,
                          stream: int = 0, d_tap: int = 0, flags: int = 0):
        L.check(self._lib.wam_fsk_batch_demodulate_device(self._h, d_samples, stride, n, d_out, out_stride, d_out_len,
                                                          d_tap or None, stream or None, flags))

    def modulate(self, data: np.ndarray, data_len=None) -> tuple[np.ndarray, np.ndarray]:
        """data uint8 [n_streams, nbytes] → (samples float32 [n_streams, total], out_len)."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        assert data.ndim == 2 and data.shape[0] == self.n_streams
        nbytes = data.shape[1]
        c = self.configs[0]
        spb = int(c["sampleRate"] // c["baudRate"])
        bpb = 8 + c["startBits"] + c["stopBits"] + (0 if c["parity"] == "none" else 1)
        total_bytes = len(c["preamblePattern"]) + len(c["sfdPattern"]) + nbytes
        total = total_bytes * bpb * spb + (2 * spb if total_bytes > 0 else 0) + bpb * spb
        out = np.zeros((self.n_streams, total), dtype=np.float32)
        out_len = np.zeros(self.n_streams, dtype=np.int32)
        dl = None
        if data_len is not None:
            dl = np.ascontiguousarray(data_len, dtype=np.int32)
        L.check(self._lib.wam_fsk_batch_modulate(self._h, data.ctypes.data if nbytes else None, max(nbytes, 1),
                                                 dl.ctypes.data if dl is not None else None, nbytes,
                                                 out.ctypes.data, total, out_len.ctypes.data))
        return out, out_len

    def modulate_device(self, d_data: int, data_stride: int, nbytes: int, d_out: int, out_stride: int,
                        d_out_len: int = 0, d_data_len: int = 0, stream: int = 0):
        L.check(self._lib.wam_fsk_batch_modulate_device(self._h, d_data or None, data_stride, d_data_len or None, nbytes,
                                                        d_out, out_stride, d_out_len or None, stream or None))

    def status(self) -> list[dict]:
        st = (L.StatusStruct * self.n_streams)()
        L.check(self._lib.wam_fsk_batch_status(self._h, st))
        return [_status_dict(st[i]) for i in range(self.n_streams)]


class ChunkedModulator:
    """Mirror of src/webaudio/chunked-modulator.ts: modulates once, hands the signal out in slices."""

    def __init__(self, modulator):
        self.modulator = modulator
        self._signal = None
        self._pos = 0

    def startModulation(self, data):
        data = bytes(data)
        if not len(data):
            self._reset()
            return
        self._signal = self.modulator.modulateData(data)
        self._pos = 0

    def getNextSamples(self, sampleCount: int):
        if self._signal is None:
            return None
        remaining = len(self._signal) - self._pos
        if remaining <= 0:
            return None
        k = min(sampleCount, remaining)
        signal = self._signal[self._pos:self._pos + k].copy()
        self._pos += k
        if self._pos >= len(self._signal):
            total = len(self._signal)
            self._reset()
            return dict(signal=signal, isComplete=True, samplesConsumed=total, totalSamples=total)
        return dict(signal=signal, isComplete=False, samplesConsumed=self._pos, totalSamples=len(self._signal))

    def isModulating(self) -> bool:
        return self._signal is not None

    def getProgress(self) -> float:
        return self._pos / len(self._signal) if self._signal is not None else 0.0

    def cancel(self):
        self._reset()

    def _reset(self):
        self._signal = None
        self._pos = 0


class FSKSessionMux:
    """Host adapter for many concurrent block-wise callers (wam_fsk_mux_*): every session pushes its render quanta
    (FSKProcessor.process() delivers 128 samples per call, fsk-processor.ts:152-167), one flush() runs a single ragged
    GPU batch over the sessions that pushed and returns each session's decoded bytes.  Sessions that pushed nothing
    are not called."""

    def __init__(self, n_sessions: int, configs, cfg_index=None, max_block: int = 1024, device: int = 0):
        self._lib = L.lib()
        if isinstance(configs, dict) or configs is None:
            configs = [configs or {}]
        self.configs = [normalize_config(c) for c in configs]
        self.n_sessions = int(n_sessions)
        structs = (L.FSKConfigStruct * len(self.configs))()
        self._keep = []
        for i, c in enumerate(self.configs):
            st, k = config_struct(c)
            structs[i] = st
            self._keep.append(k)
        idx = None
        if cfg_index is not None:
            idx = np.ascontiguousarray(cfg_index, dtype=np.int32)
            assert idx.shape == (self.n_sessions,)
        self._h = C.c_void_p()
        L.check(self._lib.wam_fsk_mux_create(device, self.n_sessions, structs, len(self.configs),
                                             idx.ctypes.data_as(C.POINTER(C.c_int32)) if idx is not None else None,
                                             max_block, C.byref(self._h)))
        self._cap = int(L.check(self._lib.wam_fsk_mux_out_capacity(self._h)))
        self._out = np.zeros((self.n_sessions, self._cap), dtype=np.uint8)
        self._out_len = np.zeros(self.n_sessions, dtype=np.int32)

    def close(self):
        if self._h:
            self._lib.wam_fsk_mux_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def push(self, session: int, samples: np.ndarray):
        a = np.ascontiguousarray(samples, dtype=np.float32)
        L.check(self._lib.wam_fsk_mux_push(self._h, session, a.ctypes.data if a.size else None, a.size))

    def pending(self, session: int) -> int:
        return int(L.check(self._lib.wam_fsk_mux_pending(self._h, session)))

    def flush(self) -> list[bytes]:
        L.check(self._lib.wam_fsk_mux_flush(self._h, self._out.ctypes.data, self._cap, self._out_len.ctypes.data))
        return [bytes(self._out[i, : self._out_len[i]]) for i in range(self.n_sessions)]

    def flush_raw(self):
        """flush() without building Python objects: (out uint8 [n_sessions, cap], out_len int32 [n_sessions]) views."""
        L.check(self._lib.wam_fsk_mux_flush(self._h, self._out.ctypes.data, self._cap, self._out_len.ctypes.data))
        return self._out, self._out_len

    # -- send half: one ChunkedModulator per session (src/webaudio/chunked-modulator.ts:31-87) -----------
    def send(self, session: int, data):
        """startModulation(data) for one session; the signal exists after the next modulate()."""
        a = np.frombuffer(bytes(data), dtype=np.uint8)
        L.check(self._lib.wam_fsk_mux_send(self._h, session, a.ctypes.data if len(a) else None, len(a)))

    def modulate(self):
        """modulateData() of every session that queued a payload, one batched GPU call."""
        L.check(self._lib.wam_fsk_mux_modulate(self._h))

    def is_modulating(self, session: int) -> bool:
        return bool(L.check(self._lib.wam_fsk_mux_is_modulating(self._h, session)))

    def pull(self, session: int, sample_count: int = 128):
        """getNextSamples(sampleCount): dict(signal, isComplete, samplesConsumed, totalSamples) or None."""
        out = np.zeros(max(sample_count, 1), dtype=np.float32)
        res = L.ChunkResult()
        rc = L.check(self._lib.wam_fsk_mux_pull(self._h, session, out.ctypes.data, sample_count, C.byref(res)))
        if rc == 0:
            return None
        return dict(signal=out[:res.samples].copy(), isComplete=bool(res.isComplete), samplesConsumed=int(res.samplesConsumed),
                    totalSamples=int(res.totalSamples))

    def status(self) -> list[dict]:
        st = (L.StatusStruct * self.n_sessions)()
        L.check(self._lib.wam_fsk_batch_status(self._lib.wam_fsk_mux_batch(self._h), st))
        return [_status_dict(st[i]) for i in range(self.n_sessions)]

