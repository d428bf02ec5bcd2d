"""Experiment (GPU): soundness of the fast kernel's doubt band.  S streams of BASELINE config 2 (host-generated) run
through the fast kernel's TAP variant (filteredPhaseDiff + band per decimated sample) and through the oracle with its
decimated-rate tap; on streams whose bytes and counters agree, reports max |F_fast - F_oracle| / band (must stay < 1),
the error statistics and the hard bits that differ outside the band (must be 0).  JSON lines; --out appends."""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=512)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--out", type=str, default="")
    a = ap.parse_args()
    import torch
    import oracle as O

    wam = importlib.import_module("webaudio-modem_b200")
    L = importlib.import_module("webaudio-modem_b200._lib")
    dev = torch.device("cuda", 0)
    S, N = a.streams, bench.N_SAMPLES
    x, cfg_index, snr, payloads = bench.generate_on_host(S, seed=a.seed)
    cfgs = [bench.CFG_CH1, bench.CFG_CH2]
    dx = torch.from_numpy(x).to(dev)
    b = wam.FSKBatch(S, cfgs, cfg_index)
    cap = b.out_capacity(N)
    d_out = torch.zeros((S, cap), dtype=torch.uint8, device=dev)
    d_len = torch.zeros(S, dtype=torch.int32, device=dev)
    d_tap = torch.zeros((S, N), dtype=torch.float32, device=dev)
    b.demodulate_device(dx.data_ptr(), N, N, d_out.data_ptr(), cap, d_len.data_ptr(), d_tap=d_tap.data_ptr(),
                        flags=L.WAM_BATCH_FORCE_FAST | L.WAM_BATCH_FAST_UNGUARDED | L.WAM_BATCH_TAP_FAST_DECISION |
                        L.WAM_BATCH_NO_SLABS)
    torch.cuda.synchronize()
    fs = b.fast_stats()
    st = b.status()
    out, ln, tap = d_out.cpu().numpy(), d_len.cpu().numpy(), d_tap.cpu().numpy()
    rows = {}
    for i in range(S):
        m = O.FSKCore()
        m.configure(cfgs[cfg_index[i]])
        want, oF, oA = m.demodulateTapped(x[i].copy())
        ost = m.getStatus()
        got = bytes(out[i, :ln[i]])
        same = got == want and float(st[i]["syncDetections"]) == float(ost["syncDetections"]) and \
            float(st[i]["eodEvents"]) == float(ost["eodEvents"])
        r = rows.setdefault(float(snr[i]), dict(streams=0, differ=0, ratio_max=0.0, err_max=0.0, err_sq=0.0, n=0,
                                                wrong_bits=0, wrong_undoubted=0, doubt=0))
        r["streams"] += 1
        if not same:
            r["differ"] += 1
            continue
        n = min(len(oF), N // 2)
        F, band = tap[i, 0:2 * n:2].astype(np.float64), tap[i, 1:2 * n:2].astype(np.float64)
        err = np.abs(F - oF[:n])
        r["ratio_max"] = max(r["ratio_max"], float(np.max(err / band)))
        r["err_max"] = max(r["err_max"], float(np.max(err)))
        r["err_sq"] += float(np.sum(err * err)); r["n"] += n
        wrong = (F > 0) != (oF[:n] > 0)
        r["wrong_bits"] += int(np.sum(wrong))
        r["wrong_undoubted"] += int(np.sum(wrong & ~(np.abs(F) < band)))
        r["doubt"] += int(np.sum(np.abs(F) < band))
    lines = []
    for s in sorted(rows):
        r = rows[s]
        d = dict(snr=s, streams=r["streams"], differ=r["differ"], ratio_max=round(r["ratio_max"], 4), err_max=r["err_max"],
                 err_rms=(r["err_sq"] / max(r["n"], 1)) ** 0.5, wrong_bits=r["wrong_bits"],
                 wrong_undoubted=r["wrong_undoubted"], doubt_per_stream=r["doubt"] / max(r["streams"] - r["differ"], 1))
        lines.append(d)
        print(json.dumps(d), flush=True)
    tot = dict(streams=S, differ=sum(r["differ"] for r in rows.values()),
               ratio_max=max(r["ratio_max"] for r in rows.values()),
               wrong_undoubted=sum(r["wrong_undoubted"] for r in rows.values()), fast_stats=fs)
    lines.append(tot)
    print(json.dumps(tot), flush=True)
    if a.out:
        with open(a.out, "a") as f:
            for d in lines:
                f.write(json.dumps(d) + "\n")


if __name__ == "__main__":
    main()
