"""ctypes binding of libwam.so (include/wam.h).  Fails loudly if the library is missing:
there is no Python/CPU fallback for any compute entry point."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WAM_LIB") or os.path.join(HERE, "libwam.so")  # WAM_LIB: A/B experiments only

WAM_OK = 0
WAM_E_INVALID = -1
WAM_E_NOT_CONFIGURED = -2
WAM_E_CUDA = -3
WAM_E_NOMEM = -4
WAM_E_CAPACITY = -5
WAM_E_UNSUPPORTED = -6
WAM_E_FILTER_B_EMPTY = -10
WAM_E_FILTER_A_EMPTY = -11
WAM_E_FILTER_A0_ZERO = -12
WAM_E_PKT_SEQUENCE = -20
WAM_E_PKT_PAYLOAD = -21

WAM_BATCH_WRITEBACK_AGC = 1
WAM_BATCH_TAP_PREFILTER = 2
WAM_BATCH_DEBUG_GENERIC_SM = 4
WAM_BATCH_NO_PIPELINE = 8
WAM_BATCH_NO_TMA = 16
WAM_BATCH_NO_SLABS = 32
WAM_BATCH_FORCE_SLABS = 64
WAM_BATCH_EXACT_ONLY = 128
WAM_BATCH_FORCE_FAST = 256
WAM_BATCH_FAST_UNGUARDED = 512
WAM_BATCH_TAP_FAST_DECISION = 1024


class WamError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


class FSKConfigStruct(C.Structure):
    _fields_ = [
        ("sampleRate", C.c_double), ("baudRate", C.c_double),
        ("markFrequency", C.c_double), ("spaceFrequency", C.c_double),
        ("preamblePattern", C.POINTER(C.c_uint8)), ("preambleLength", C.c_int32),
        ("sfdPattern", C.POINTER(C.c_uint8)), ("sfdLength", C.c_int32),
        ("startBits", C.c_int32), ("stopBits", C.c_int32), ("parity", C.c_int32),
        ("syncThreshold", C.c_double), ("agcEnabled", C.c_int32),
        ("preFilterBandwidth", C.c_double), ("adaptiveThreshold", C.c_int32),
    ]


class StatusStruct(C.Structure):
    _fields_ = [
        ("ready", C.c_int32), ("frameStarted", C.c_int32),
        ("globalSampleCounter", C.c_double), ("receivedBitsLength", C.c_double),
        ("byteBufferLength", C.c_double), ("demodulationCalls", C.c_double),
        ("syncDetections", C.c_double), ("silenceThreshold", C.c_double),
        ("totalSamplesProcessed", C.c_double), ("eodEvents", C.c_double),
        ("errorEvents", C.c_double), ("configuredEvents", C.c_double),
    ]


class PktResult(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("sequence", C.c_int32), ("length", C.c_int32),
        ("payloadOffset", C.c_int32), ("crcReceived", C.c_int32), ("crcComputed", C.c_int32),
        ("bytesConsumed", C.c_int32),
    ]


class FastStats(C.Structure):
    """wam_fast_stats (include/wam.h)"""
    _fields_ = [("fast_calls", C.c_int64), ("flagged_last_call", C.c_int64), ("flagged_streams", C.c_int64),
                ("doubtful_samples", C.c_int64), ("flag_causes", C.c_uint32), ("error_flags", C.c_uint32),
                ("windows_confirmed", C.c_int64), ("windows_refuted", C.c_int64), ("windows_dropped", C.c_int64),
                ("carried_settled", C.c_int64), ("carried_corrected", C.c_int64)]


class ChunkResult(C.Structure):
    """wam_chunk_result (include/wam.h)"""
    _fields_ = [("samples", C.c_long), ("isComplete", C.c_int), ("samplesConsumed", C.c_long), ("totalSamples", C.c_long)]


class XmodemRxState(C.Structure):
    """wam_xmodem_rx_state (include/wam.h)"""
    _fields_ = [("expectedSequence", C.c_int32), ("retries", C.c_int32), ("done", C.c_int32),
                ("dataLen", C.c_int32), ("packetsReceived", C.c_int32), ("packetsDropped", C.c_int32)]


# every symbol include/wam.h declares: name -> (restype, argtypes)
_vp, _dp, _fp = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_float)
_u8p, _i32p, _u16p = C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_uint16)
_cfgp, _stp, _pktp = C.POINTER(FSKConfigStruct), C.POINTER(StatusStruct), C.POINTER(PktResult)
SYMBOLS = {
    "wam_version": (C.c_int, []),
    "wam_last_error": (C.c_char_p, []),
    "wam_error_string": (C.c_char_p, [C.c_int]),
    "wam_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "wam_fsk_default_config": (None, [_cfgp]),
    "wam_fsk_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "wam_fsk_destroy": (C.c_int, [_vp]),
    "wam_fsk_configure": (C.c_int, [_vp, _cfgp]),
    "wam_fsk_is_ready": (C.c_int, [_vp]),
    "wam_fsk_modulate_size": (C.c_long, [_vp, C.c_long]),
    "wam_fsk_modulate": (C.c_int, [_vp, _u8p, C.c_long, _fp, C.c_long, C.POINTER(C.c_long)]),
    "wam_fsk_demodulate": (C.c_int, [_vp, _fp, C.c_long, _u8p, C.c_long, C.POINTER(C.c_long)]),
    "wam_fsk_reset": (C.c_int, [_vp]),
    "wam_fsk_status_get": (C.c_int, [_vp, _stp]),
    "wam_fsk_batch_create": (C.c_int, [C.c_int, C.c_long, _cfgp, C.c_int, _i32p, C.POINTER(_vp)]),
    "wam_fsk_batch_destroy": (C.c_int, [_vp]),
    "wam_fsk_batch_reset": (C.c_int, [_vp]),
    "wam_fsk_batch_renew": (C.c_int, [_vp, _vp]),
    "wam_fsk_batch_out_capacity": (C.c_long, [_vp, C.c_long]),
    "wam_fsk_batch_demodulate": (C.c_int, [_vp, _vp, C.c_long, C.c_long, _vp, C.c_long, _vp, C.c_uint32]),
    "wam_fsk_batch_demodulate_device": (C.c_int, [_vp, _vp, C.c_long, C.c_long, _vp, C.c_long, _vp, _vp, _vp, C.c_uint32]),
    "wam_fsk_batch_demodulate_ragged": (C.c_int, [_vp, _vp, C.c_long, C.c_long, _vp, _vp, C.c_long, _vp, C.c_uint32]),
    "wam_fsk_batch_demodulate_ragged_device": (C.c_int, [_vp, _vp, C.c_long, C.c_long, _vp, _vp, C.c_long, _vp, _vp, C.c_uint32]),
    "wam_fsk_mux_create": (C.c_int, [C.c_int, C.c_long, _cfgp, C.c_int, _i32p, C.c_long, C.POINTER(_vp)]),
    "wam_fsk_mux_destroy": (C.c_int, [_vp]),
    "wam_fsk_mux_push": (C.c_int, [_vp, C.c_long, _vp, C.c_long]),
    "wam_fsk_mux_pending": (C.c_long, [_vp, C.c_long]),
    "wam_fsk_mux_out_capacity": (C.c_long, [_vp]),
    "wam_fsk_mux_flush": (C.c_int, [_vp, _vp, C.c_long, _vp]),
    "wam_fsk_mux_batch": (_vp, [_vp]),
    "wam_fsk_batch_status": (C.c_int, [_vp, _stp]),
    "wam_fsk_batch_launch_count": (C.c_long, [_vp]),
    "wam_fsk_batch_fast_stats": (C.c_int, [_vp, C.POINTER(FastStats)]),
    "wam_fsk_batch_debug_fast_band": (C.c_int, [_vp, C.c_double]),
    "wam_fsk_batch_demodulate_pcm16": (C.c_int, [_vp, _vp, C.c_long, C.c_long, _vp, C.c_long, _vp, C.c_uint32]),
    "wam_host_bind_near_device": (C.c_int, [C.c_int]),
    "wam_fsk_batch_debug_fast_windows": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_long]),
    "wam_iir_scratch_bytes": (C.c_size_t, [C.c_long, C.c_long]),
    "wam_iir_process_batch_device": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, _vp, C.c_long, C.c_long, C.c_long, _vp, _vp, C.c_size_t, _vp]),
    "wam_fir_process_batch_device": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_long, C.c_long, C.c_long, _vp, _vp, _vp]),
    "wam_fsk_mux_send": (C.c_int, [_vp, C.c_long, _vp, C.c_long]),
    "wam_fsk_mux_modulate": (C.c_int, [_vp]),
    "wam_fsk_mux_is_modulating": (C.c_int, [_vp, C.c_long]),
    "wam_fsk_mux_pull": (C.c_int, [_vp, C.c_long, _vp, C.c_long, C.POINTER(ChunkResult)]),
    "wam_awgn_add_device": (C.c_int, [_vp, C.c_long, C.c_long, C.c_long, _vp, C.c_ulonglong, C.c_uint, _vp]),
    "wam_fsk_batch_debug_phase_cycles": (C.c_int, [_vp, C.c_int, _dp, _dp, C.c_long]),
    "wam_fsk_batch_modulate": (C.c_int, [_vp, _vp, C.c_long, _vp, C.c_long, _vp, C.c_long, _vp]),
    "wam_fsk_batch_modulate_device": (C.c_int, [_vp, _vp, C.c_long, _vp, C.c_long, _vp, C.c_long, _vp, _vp]),
    "wam_crc16": (C.c_uint16, [_u8p, C.c_long]),
    "wam_xmodem_serialize": (C.c_long, [C.c_int, _u8p, C.c_long, _u8p, C.c_long]),
    "wam_xmodem_batch_check": (C.c_int, [C.c_int, _vp, C.c_long, _vp, _vp, C.c_long, _pktp]),
    "wam_xmodem_batch_check_device": (C.c_int, [_vp, C.c_long, _vp, _vp, C.c_long, _vp, _vp]),
    "wam_xmodem_batch_receive": (C.c_int, [C.c_int, _vp, C.c_long, _vp, C.c_long, C.c_int, _vp, _vp, C.c_int, _vp, _vp, _vp, C.c_long]),
    "wam_xmodem_batch_receive_device": (C.c_int, [_vp, C.c_long, _vp, C.c_long, C.c_int, _vp, _vp, C.c_int, _vp, _vp, _vp, C.c_long, _vp]),
    "wam_crc16_batch": (C.c_int, [C.c_int, _vp, C.c_long, _vp, C.c_long, _vp]),
    "wam_design_butterworth_lowpass": (None, [C.c_double, C.c_double, _dp, _dp]),
    "wam_design_butterworth_highpass": (None, [C.c_double, C.c_double, _dp, _dp]),
    "wam_design_butterworth_bandpass": (None, [C.c_double, C.c_double, C.c_double, _dp, _dp]),
    "wam_design_sinc_lowpass": (C.c_int, [C.c_double, C.c_double, C.c_int, _dp]),
    "wam_design_sinc_highpass": (C.c_int, [C.c_double, C.c_double, C.c_int, _dp]),
    "wam_design_sinc_bandpass": (C.c_int, [C.c_double, C.c_double, C.c_double, C.c_int, _dp]),
    "wam_iir_state_size": (C.c_long, [C.c_int, C.c_int]),
    "wam_iir_process_batch": (C.c_int, [C.c_int, _dp, C.c_int, _dp, C.c_int, _vp, _vp, C.c_long, C.c_long, C.c_long, _vp]),
    "wam_fir_process_batch": (C.c_int, [C.c_int, _dp, C.c_int, _vp, _vp, C.c_long, C.c_long, C.c_long, _vp]),
    "wam_debug_fastmath": (C.c_int, [C.c_int, _dp, _dp, C.c_long, _dp, _dp, _dp]),
    "wam_host_alloc": (C.c_int, [C.POINTER(_vp), C.c_size_t]),
    "wam_host_free": (C.c_int, [_vp]),
}

_lib = None


def lib():
    """Load libwam.so (built by webaudio-modem_b200/build.py).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m __graft_entry__` / build.py (nvcc, sm_100a). "
            "There is no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc: int) -> int:
    if rc < 0:
        msg = lib().wam_last_error().decode("utf-8", "replace")
        raise WamError(rc, msg or lib().wam_error_string(rc).decode())
    return rc
