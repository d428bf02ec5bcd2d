"""Host-side mirror of src/dsp/filters.ts: IIRFilter, FIRFilter, FilterDesign, FilterFactory.

Designs are host math inside libwam.so (wam_design_*); sample processing runs on the GPU
(wam_iir_process_batch: time-chunked linear-recurrence scan; wam_fir_process_batch: shared-memory
staged FIR).  Filter state (the reference's circular x/y histories) lives in a small float64 vector
carried between calls.  Outputs are float32, i.e. the reference's processBuffer() contract
(filters.ts:81-87); process(x) returns that float32 value as a Python float.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

_dp = C.POINTER(C.c_double)


def _d(a):
    return a.ctypes.data_as(_dp)


class IIRFilter:
    def __init__(self, b, a, device: int = 0):
        # constructor checks and messages — filters.ts:19-21
        if b is None or len(b) == 0:
            raise ValueError("Feedforward coefficients (b) cannot be empty")
        if a is None or len(a) == 0:
            raise ValueError("Feedback coefficients (a) cannot be empty")
        if a[0] == 0:
            raise ValueError("First feedback coefficient (a[0]) cannot be zero")
        self._b = np.array(b, dtype=np.float64)  # copies (filters.ts:23-24)
        self._a = np.array(a, dtype=np.float64)
        if self._a[0] != 1:  # filters.ts:30-39
            a0 = self._a[0]
            self._b = self._b / a0
            self._a[1:] = self._a[1:] / a0
            self._a[0] = 1
        self._device = device
        self._lib = L.lib()
        self.reset()

    def reset(self):  # filters.ts:92-98
        n = int(self._lib.wam_iir_state_size(len(self._b), len(self._a)))
        self._state = np.zeros(max(n, 1), dtype=np.float64)

    def processBuffer(self, x) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.zeros_like(x)
        if len(x) == 0:
            return out
        L.check(self._lib.wam_iir_process_batch(self._device, _d(self._b), len(self._b), _d(self._a), len(self._a),
                                                x.ctypes.data, out.ctypes.data, len(x), len(x), 1,
                                                self._state.ctypes.data))
        return out

    def process(self, x: float) -> float:
        return float(self.processBuffer(np.array([x], dtype=np.float32))[0])

    def getCoefficients(self):
        return {"b": self._b.copy(), "a": self._a.copy()}


class FIRFilter:
    def __init__(self, coefficients, device: int = 0):
        self._c = np.array(coefficients, dtype=np.float64)
        self._device = device
        self._lib = L.lib()
        self.reset()

    def reset(self):  # filters.ts:156-159
        self._state = np.zeros(max(len(self._c) - 1, 1), dtype=np.float64)

    def processBuffer(self, x) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.zeros_like(x)
        if len(x) == 0:
            return out
        L.check(self._lib.wam_fir_process_batch(self._device, _d(self._c) if len(self._c) else None, len(self._c),
                                                x.ctypes.data, out.ctypes.data, len(x), len(x), 1,
                                                self._state.ctypes.data))
        return out

    def process(self, x: float) -> float:
        return float(self.processBuffer(np.array([x], dtype=np.float32))[0])

    def getCoefficients(self):
        return self._c.copy()


def iir_process_batch(b, a, x: np.ndarray, state: np.ndarray | None = None, device: int = 0) -> np.ndarray:
    """x float32 [n_streams, n] → y float32, every stream filtered by the same (b, a)."""
    b = np.ascontiguousarray(b, dtype=np.float64)
    a = np.ascontiguousarray(a, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float32)
    assert x.ndim == 2
    out = np.zeros_like(x)
    L.check(L.lib().wam_iir_process_batch(device, _d(b) if len(b) else None, len(b), _d(a) if len(a) else None, len(a),
                                          x.ctypes.data, out.ctypes.data, x.shape[1], x.shape[1], x.shape[0],
                                          state.ctypes.data if state is not None else None))
    return out


def fir_process_batch(taps, x: np.ndarray, state: np.ndarray | None = None, device: int = 0) -> np.ndarray:
    taps = np.ascontiguousarray(taps, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float32)
    assert x.ndim == 2
    out = np.zeros_like(x)
    L.check(L.lib().wam_fir_process_batch(device, _d(taps) if len(taps) else None, len(taps), x.ctypes.data,
                                          out.ctypes.data, x.shape[1], x.shape[1], x.shape[0],
                                          state.ctypes.data if state is not None else None))
    return out


class FilterDesign:
    """src/dsp/filters.ts:172-315 (host math in libwam.so, float64, the reference's operation order)."""

    @staticmethod
    def butterworthLowpass(cutoffFreq, sampleRate):
        b, a = np.zeros(3), np.zeros(3)
        L.lib().wam_design_butterworth_lowpass(cutoffFreq, sampleRate, _d(b), _d(a))
        return {"b": b, "a": a}

    @staticmethod
    def butterworthHighpass(cutoffFreq, sampleRate):
        b, a = np.zeros(3), np.zeros(3)
        L.lib().wam_design_butterworth_highpass(cutoffFreq, sampleRate, _d(b), _d(a))
        return {"b": b, "a": a}

    @staticmethod
    def butterworthBandpass(centerFreq, bandwidth, sampleRate):
        b, a = np.zeros(3), np.zeros(3)
        L.lib().wam_design_butterworth_bandpass(centerFreq, bandwidth, sampleRate, _d(b), _d(a))
        return {"b": b, "a": a}

    @staticmethod
    def sincLowpass(cutoffFreq, sampleRate, numTaps):
        out = np.zeros(numTaps + 2)
        n = L.lib().wam_design_sinc_lowpass(cutoffFreq, sampleRate, numTaps, _d(out))
        return out[:n].copy()

    @staticmethod
    def sincHighpass(cutoffFreq, sampleRate, numTaps):
        out = np.zeros(numTaps + 2)
        n = L.lib().wam_design_sinc_highpass(cutoffFreq, sampleRate, numTaps, _d(out))
        return out[:n].copy()

    @staticmethod
    def sincBandpass(centerFreq, bandwidth, sampleRate, numTaps):
        out = np.zeros(numTaps + 2)
        n = L.lib().wam_design_sinc_bandpass(centerFreq, bandwidth, sampleRate, numTaps, _d(out))
        return out[:n].copy()


class FilterFactory:
    """src/dsp/filters.ts:320-368"""

    @staticmethod
    def createIIRLowpass(cutoffFreq, sampleRate):
        c = FilterDesign.butterworthLowpass(cutoffFreq, sampleRate)
        return IIRFilter(c["b"], c["a"])

    @staticmethod
    def createIIRHighpass(cutoffFreq, sampleRate):
        c = FilterDesign.butterworthHighpass(cutoffFreq, sampleRate)
        return IIRFilter(c["b"], c["a"])

    @staticmethod
    def createIIRBandpass(centerFreq, bandwidth, sampleRate):
        c = FilterDesign.butterworthBandpass(centerFreq, bandwidth, sampleRate)
        return IIRFilter(c["b"], c["a"])

    @staticmethod
    def createFIRLowpass(cutoffFreq, sampleRate, numTaps=51):
        return FIRFilter(FilterDesign.sincLowpass(cutoffFreq, sampleRate, numTaps))

    @staticmethod
    def createFIRHighpass(cutoffFreq, sampleRate, numTaps=51):
        return FIRFilter(FilterDesign.sincHighpass(cutoffFreq, sampleRate, numTaps))

    @staticmethod
    def createFIRBandpass(centerFreq, bandwidth, sampleRate, numTaps=51):
        return FIRFilter(FilterDesign.sincBandpass(centerFreq, bandwidth, sampleRate, numTaps))
