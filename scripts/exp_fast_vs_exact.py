"""Experiment (GPU): the mixed-precision fast demodulator against the float64 kernel and the oracle.

Part A — oracle subset: S_A streams of BASELINE config 2 generated on the host; oracle bytes/counters vs GPU exact,
GPU fast (guarded) and GPU fast unguarded (float32 results kept for flagged streams), per SNR class.
Part B — full config 2 on the device (65,536 streams): exact vs fast vs unguarded on every stream, CUDA-event timing.
Prints JSON lines; `--out` appends them to a file under profiles/.
"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def status_key(st):
    return (float(st["syncDetections"]), float(st["eodEvents"]), float(st["globalSampleCounter"]), float(st["frameStarted"]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams-a", type=int, default=4096)
    ap.add_argument("--streams-b", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--out", type=str, default="")
    ap.add_argument("--skip-a", action="store_true")
    a = ap.parse_args()
    import torch

    wam = importlib.import_module("webaudio-modem_b200")
    L = importlib.import_module("webaudio-modem_b200._lib")
    dev = torch.device("cuda", 0)
    lines = []

    def emit(d):
        print(json.dumps(d), flush=True)
        lines.append(d)

    if not a.skip_a:
        import oracle as O

        S = a.streams_a
        x, cfg_index, snr, payloads = bench.generate_on_host(S, seed=7)
        t0 = time.time()
        want, ost = O.batch_demodulate([bench.CFG_CH1, bench.CFG_CH2], cfg_index, x.copy(),
                                       n_threads=a.threads or bench.host_cores())
        t_or = time.time() - t0
        res = {}
        for name, fl in (("exact", L.WAM_BATCH_EXACT_ONLY), ("fast", L.WAM_BATCH_FORCE_FAST),
                         ("unguarded", L.WAM_BATCH_FORCE_FAST | L.WAM_BATCH_FAST_UNGUARDED)):
            b = wam.FSKBatch(S, [bench.CFG_CH1, bench.CFG_CH2], cfg_index)
            got = b.demodulate_bytes(x.copy(), flags=fl)
            st = b.status()
            fs = b.fast_stats()
            bad = [i for i in range(S) if got[i] != want[i] or status_key(st[i]) != status_key(ost[i])]
            per = {}
            for i in bad:
                per[str(snr[i])] = per.get(str(snr[i]), 0) + 1
            res[name] = dict(mismatch=len(bad), per_snr=per, fast_stats=fs)
            b.close()
        emit(dict(part="A", streams=S, oracle_s=round(t_or, 2), bytes=sum(len(w) for w in want), **res))

    # ---- part B
    S = a.streams_b
    x, cfg_index, snr, payloads = bench.generate_on_device(wam, torch, dev, S, seed=1000)
    cap = None
    outs = {}
    for name, fl in (("exact", L.WAM_BATCH_EXACT_ONLY), ("fast", 0), ("unguarded", L.WAM_BATCH_FAST_UNGUARDED)):
        b = wam.FSKBatch(S, [bench.CFG_CH1, bench.CFG_CH2], cfg_index)
        cap = b.out_capacity(bench.N_SAMPLES)
        d_out = torch.zeros((S, cap), dtype=torch.uint8, device=dev)
        d_len = torch.zeros(S, dtype=torch.int32, device=dev)
        sp = torch.cuda.current_stream().cuda_stream
        times = []
        for it in range(a.iters + 2):
            b.renew(sp)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            b.demodulate_device(x.data_ptr(), bench.N_SAMPLES, bench.N_SAMPLES, d_out.data_ptr(), cap, d_len.data_ptr(),
                                stream=sp, flags=fl)
            e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                times.append(e0.elapsed_time(e1))
        st = b.status()
        fs = b.fast_stats()
        outs[name] = (d_out.cpu().numpy(), d_len.cpu().numpy(), st, fs, times)
        b.close()
        del d_out, d_len
    eo, el, est, _, et = outs["exact"]
    total = S * bench.N_SAMPLES
    for name in ("fast", "unguarded"):
        go, gl, gst, fs, gt = outs[name]
        bad = []
        for i in range(S):
            if gl[i] != el[i] or (gl[i] and not np.array_equal(go[i, :gl[i]], eo[i, :el[i]])) or \
                    status_key(gst[i]) != status_key(est[i]):
                bad.append(i)
        per = {}
        for i in bad:
            per[str(snr[i])] = per.get(str(snr[i]), 0) + 1
        emit(dict(part="B", kernel=name, streams=S, differs_from_exact=len(bad), per_snr=per, fast_stats=fs,
                  ms=[round(t, 3) for t in gt], ms_min=round(min(gt), 3),
                  gsamples_per_s=round(total / min(gt) / 1e6, 1),
                  exact_ms_min=round(min(et), 3), exact_gsamples_per_s=round(total / min(et) / 1e6, 1)))
    if a.out:
        with open(a.out, "a") as f:
            for d in lines:
                f.write(json.dumps(d) + "\n")


if __name__ == "__main__":
    main()
