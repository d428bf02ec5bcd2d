"""Experiment (CPU, oracle): distance until a speculative FSKCore run started mid-stream has the state of the true run
(oracle/merge_exp.c).  Config 3 style (48 kHz / 1200 Bd, 128-byte frames back to back, 0..2000-sample gaps, +6 dB) and
config 4 style (44.1 kHz / 1200 Bd, parity even, tone offset, +9 dB) streams, plus V.21 300 Bd.  JSON lines."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as O  # noqa: E402
import siggen  # noqa: E402

so = os.path.join(ROOT, "oracle", "libwam_merge_exp.so")
subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=c11", "-ffp-contract=off", "-D_GNU_SOURCE", "-Wno-unused-function", "-shared",
                       "-o", so, os.path.join(ROOT, "oracle", "merge_exp.c"), "-lm", "-lpthread"])
lib = C.CDLL(so)


class Res(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("merged_b", "merged_close", "merged_exact", "gsc_guess_ok", "bytes_true", "bytes_same")]


def run(name, cfg, n, payload, snr, warmup, n_streams, freq_off=0.0, max_gap=2000):
    rows = []
    for s in range(n_streams):
        x, _ = siggen.multi_frame_stream(cfg, n, payload, snr, seed=900 + s, max_gap=max_gap, freq_offset_hz=freq_off * ((s % 5) - 2) / 2)
        st, keep = O.make_config_struct(cfg)
        for T in (n // 4, n // 2, (5 * n) // 8):
            T -= T % 2
            r = Res()
            lib.merge_experiment(C.byref(st), x.ctypes.data_as(C.POINTER(C.c_float)), C.c_long(n), C.c_long(T), C.c_long(warmup), C.byref(r))
            rows.append((r.merged_b, r.merged_close, r.merged_exact, r.gsc_guess_ok, r.bytes_true, r.bytes_same))
    a = np.array(rows, dtype=np.int64)

    def q(col):
        v = a[:, col]
        ok = v[v >= 0]
        return dict(never=int((v < 0).sum()), median=int(np.median(ok)) if len(ok) else None,
                    p90=int(np.percentile(ok, 90)) if len(ok) else None, max=int(ok.max()) if len(ok) else None)

    out = dict(workload=name, trials=len(rows), warmup_samples=warmup, samples_per_stream=n,
               state_machine_equal=q(0), dsp_within_1e12=q(1), dsp_bitwise=q(2),
               gsc_guess_ok=int(a[:, 3].sum()), bytes_after_merge=int(a[:, 4].sum()), bytes_after_merge_identical=int(a[:, 5].sum()))
    print(json.dumps(out), flush=True)
    return out


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    k = 4 if quick else 12
    run("config 3 style: 48 kHz / 1200 Bd, 128 B frames, gaps 0..2000, +6 dB", {}, 400000, 128, 6.0, 16384, k)
    run("config 4 style: 44.1 kHz / 1200 Bd, parity even, tone offset, 64 B frames, +9 dB",
        dict(sampleRate=44100, baudRate=1200, parity="even"), 300000, 64, 9.0, 16384, k, freq_off=20.0)
    run("V.21 ch2 300 Bd, 25 B frames, gaps 0..2000, +6 dB", siggen.V21_CH2, 400000, 25, 6.0, 16384, k)
    run("config 3 style, short warm-up 2048", {}, 400000, 128, 6.0, 2048, k)
    run("config 3 style, low SNR 0 dB", {}, 400000, 128, 0.0, 16384, k)
