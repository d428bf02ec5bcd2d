"""Soak (GPU): config 2 at full size over many seeds — the timed fast path against the float64 kernels on every stream
(decoded bytes, lengths, status counters), with the fast path's statistics and the call time per seed.
usage: python scripts/soak_config2_seeds.py [--seeds 1000:1016] [--out profiles/r02_soak_seeds.jsonl]"""
import argparse, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # (sets CUDA_DEVICE_MAX_CONNECTIONS before the CUDA context exists)
import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--seeds", default="1000:1016")
ap.add_argument("--streams", type=int, default=65536)
ap.add_argument("--out", default="")
a = ap.parse_args()
lo, hi = (int(v) for v in a.seeds.split(":"))
wam = importlib.import_module("webaudio-modem_b200")
L = importlib.import_module("webaudio-modem_b200._lib")
dev = torch.device("cuda", 0)
S = a.streams
KEYS = ("syncDetections", "eodEvents", "globalSampleCounter", "frameStarted", "receivedBitsLength", "silenceThreshold")
lines = []
for seed in range(lo, hi):
    x, cfg_index, snr, payloads = bench.generate_on_device(wam, torch, dev, S, seed=seed)
    res = {}
    for name, fl in (("fast", 0), ("exact", L.WAM_BATCH_EXACT_ONLY)):
        b = wam.FSKBatch(S, [bench.CFG_CH1, bench.CFG_CH2], cfg_index)
        cap = b.out_capacity(bench.N_SAMPLES)
        d_out = torch.zeros((S, cap), dtype=torch.uint8, device=dev)
        d_len = torch.zeros(S, dtype=torch.int32, device=dev)
        sp = torch.cuda.current_stream().cuda_stream
        ms = []
        for it in range(3):
            b.renew(sp)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            b.demodulate_device(x.data_ptr(), bench.N_SAMPLES, bench.N_SAMPLES, d_out.data_ptr(), cap, d_len.data_ptr(), stream=sp, flags=fl)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        st = b.status()
        res[name] = (d_out.cpu().numpy(), d_len.cpu().numpy(), [[float(s[k]) for k in KEYS] for s in st], min(ms[1:]), b.fast_stats())
        b.close()
    fo, fl_, fs, fms, fstat = res["fast"]
    eo, el, es, ems, _ = res["exact"]
    bad = [i for i in range(S) if fl_[i] != el[i] or bytes(fo[i, :fl_[i]]) != bytes(eo[i, :el[i]]) or fs[i][:5] != es[i][:5]
           or abs(fs[i][5] - es[i][5]) > 1e-5 * abs(es[i][5])]
    d = {"seed": seed, "streams": S, "differing": len(bad), "first": bad[:5], "decoded_bytes": int(el.sum()), "fast_ms": round(fms, 3),
         "float64_ms": round(ems, 3), "windows": fstat["windows_confirmed"] + fstat["windows_refuted"], "refuted": fstat["windows_refuted"],
         "rerun_streams": fstat["flagged_last_call"], "flag_causes": fstat["flag_causes"], "error_flags": fstat["error_flags"]}
    print(json.dumps(d), flush=True)
    lines.append(d)
    del x
if a.out:
    with open(a.out, "w") as f:
        for d in lines:
            f.write(json.dumps(d) + "\n")
print(json.dumps({"seeds": len(lines), "total_differing": sum(d["differing"] for d in lines),
                  "mean_fast_ms": round(float(np.mean([d["fast_ms"] for d in lines])), 3),
                  "median_fast_ms": round(float(np.median([d["fast_ms"] for d in lines])), 3),
                  "max_fast_ms": max(d["fast_ms"] for d in lines)}))
