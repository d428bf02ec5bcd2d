"""Experiment (CPU): float32 fast-path model vs the float64 oracle on config-2 style streams.
Reports, per SNR class: float32 error of filteredPhaseDiff, doubtful samples, flagged streams by cause, and
soundness (every stream whose bytes / counters differ from the oracle must be flagged)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as O  # noqa: E402
from oracle import fastmodel as FM  # noqa: E402
import siggen  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--per-class", type=int, default=32)
    ap.add_argument("--n", type=int, default=48000)
    ap.add_argument("--baud", type=int, default=300)
    ap.add_argument("--payload", type=int, default=25)
    ap.add_argument("--threads", type=int, default=8)
    ap.add_argument("--form", type=int, default=1)
    ap.add_argument("--eps0", type=float, default=1e-6)
    ap.add_argument("--kappa", type=float, default=3e-7)
    ap.add_argument("--unguarded", type=int, default=0)
    ap.add_argument("--eps-amp", type=float, default=1e-6)
    ap.add_argument("--bc-delta", type=float, default=2e-6)
    ap.add_argument("--snr", type=str, default="-15:31:3")
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    lo, hi, st = (int(v) for v in a.snr.split(":"))
    snrs = np.arange(lo, hi, st, dtype=np.float64)
    cfg = dict(siggen.V21_CH2) if a.baud == 300 else dict(baudRate=a.baud)
    snr = np.repeat(snrs, a.per_class)
    t0 = time.time()
    x, _ = siggen.noisy_streams(cfg, len(snr), a.n, a.payload, snr, seed=a.seed)
    t1 = time.time()
    want, ost = O.batch_demodulate([cfg], None, x.copy(), n_threads=a.threads)
    t2 = time.time()
    got, res = FM.run([cfg], None, x, compare=True, n_threads=a.threads, form=a.form, eps0=a.eps0, kappa=a.kappa,
                      unguarded=a.unguarded, eps_amp=a.eps_amp, bc_delta=a.bc_delta)
    t3 = time.time()
    print(f"gen {t1-t0:.1f}s oracle {t2-t1:.1f}s model {t3-t2:.1f}s", file=sys.stderr)
    rows = []
    for ci, s in enumerate(snrs):
        sl = slice(ci * a.per_class, (ci + 1) * a.per_class)
        r = res[sl]
        mism = [i for i in range(sl.start, sl.stop) if got[i] != want[i] or res[i].syncDetections != ost[i]["syncDetections"]
                or res[i].gsc != ost[i]["globalSampleCounter"] or res[i].eodEvents != ost[i]["eodEvents"]
                or res[i].started != ost[i]["frameStarted"]]
        same = [i for i in range(sl.start, sl.stop) if i not in mism]
        unsound = [i for i in mism if not res[i].flag]
        causes = np.sum([[q.n_cause[k] > 0 for k in range(6)] for q in r], axis=0)
        row = dict(snr=float(s), streams=a.per_class, mismatch=len(mism), flagged=sum(q.flag for q in r),
                   unsound=len(unsound), causes={FM.CAUSES[k]: int(causes[k]) for k in range(6) if causes[k]},
                   doubt_per_stream=float(np.mean([q.n_doubt_samples for q in r])),
                   bc_per_stream=float(np.mean([q.n_bc for q in r])),
                   err_rms=float(np.median([res[i].err_rms for i in same])) if same else None,
                   err_max=float(np.max([res[i].err_max for i in same])) if same else None,
                   ratio_max=float(np.max([res[i].ratio_max for i in same])) if same else None,
                   wrong_bits=int(sum(res[i].n_wrong_bits for i in same)),
                   wrong_undoubted=int(sum(res[i].n_wrong_undoubted for i in same)),
                   bytes=int(sum(len(w) for w in want[sl])))
        rows.append(row)
        print(json.dumps(row))
    tot = dict(streams=len(snr), mismatch=sum(r["mismatch"] for r in rows), flagged=sum(r["flagged"] for r in rows),
               unsound=sum(r["unsound"] for r in rows))
    print(json.dumps(tot))


if __name__ == "__main__":
    main()
