"""ctypes front end of oracle/fastmodel.c — the CPU model of the float32 fast demodulator with decision-level doubt
tracking (test infrastructure; see the header of fastmodel.c)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import FSKConfigStruct, make_config_struct

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libwam_fastmodel.so")
CAUSES = ("vote_start", "vote_data", "vote_stop", "sync", "eod", "range")


class Params(C.Structure):
    _fields_ = [("eps0", C.c_double), ("kappa", C.c_double), ("eps_amp", C.c_double), ("bc_delta", C.c_double),
                ("form", C.c_int32), ("unguarded", C.c_int32)]


class Result(C.Structure):
    _fields_ = [("n_out", C.c_int32), ("flag", C.c_int32), ("first_cause", C.c_int32), ("first_index", C.c_int32),
                ("n_cause", C.c_int32 * 6), ("n_doubt_samples", C.c_int32), ("n_dec", C.c_int32), ("n_bc", C.c_int32),
                ("started", C.c_int32), ("syncDetections", C.c_double), ("eodEvents", C.c_double), ("gsc", C.c_double),
                ("sil_thr", C.c_double), ("err_max", C.c_double), ("err_rms", C.c_double), ("ratio_max", C.c_double),
                ("n_wrong_bits", C.c_int32), ("n_wrong_undoubted", C.c_int32),
                ("w_index", C.c_double), ("w_amp", C.c_double), ("w_S", C.c_double), ("w_err", C.c_double),
                ("w_band", C.c_double), ("w_pd", C.c_double), ("w_oamp", C.c_double), ("w_F", C.c_double)]


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("fastmodel.c", "wam_oracle.c", "wam_oracle.h")]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(f) > os.path.getmtime(_LIB_PATH) for f in srcs)
    if force or stale:
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-D_GNU_SOURCE",
                               "-Wno-unused-function", "-shared", "-o", _LIB_PATH, srcs[0], "-lm", "-lpthread"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
    return _lib


DEFAULT_PARAMS = dict(eps0=1e-6, kappa=3e-7, eps_amp=1e-4, bc_delta=2e-6, form=1, unguarded=0)


def run(cfgs: list[dict], cfg_index, samples: np.ndarray, compare: bool = False, n_threads: int = 1, **params):
    """samples float32 [n_streams, n]; returns (list[bytes], list[Result])."""
    assert samples.dtype == np.float32 and samples.flags.c_contiguous and samples.ndim == 2
    ns, n = samples.shape
    structs = (FSKConfigStruct * len(cfgs))()
    keep = []
    for i, c in enumerate(cfgs):
        s, k = make_config_struct(c)
        structs[i] = s
        keep.append(k)
    idx = np.ascontiguousarray(cfg_index, dtype=np.int32) if cfg_index is not None else None
    p = Params(**{**DEFAULT_PARAMS, **params})
    cap = n // 8 + 16
    out = np.zeros((ns, cap), dtype=np.uint8)
    res = (Result * ns)()
    lib().fm_batch(structs, idx.ctypes.data_as(C.POINTER(C.c_int32)) if idx is not None else None, C.c_long(ns),
                   samples.ctypes.data_as(C.POINTER(C.c_float)), C.c_long(samples.strides[0] // 4), C.c_long(n),
                   C.byref(p), out.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_long(cap), res, 1 if compare else 0,
                   n_threads)
    return [bytes(out[i, : min(res[i].n_out, cap)]) for i in range(ns)], list(res)
