"""One (or a few) device-resident config-2 demodulation calls, for ncu captures: python scripts/run_config2_once.py [--flags F] [--iters N] [--streams S]"""
import argparse, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
ap = argparse.ArgumentParser()
ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--streams", type=int, default=65536)
ap.add_argument("--seed", type=int, default=1000)
a = ap.parse_args()
import torch
wam = importlib.import_module("webaudio-modem_b200")
dev = torch.device("cuda", 0)
x, cfg_index, snr, payloads = bench.generate_on_device(wam, torch, dev, a.streams, seed=a.seed)
b = wam.FSKBatch(a.streams, [bench.CFG_CH1, bench.CFG_CH2], cfg_index)
cap = b.out_capacity(bench.N_SAMPLES)
d_out = torch.zeros((a.streams, cap), dtype=torch.uint8, device=dev)
d_len = torch.zeros(a.streams, dtype=torch.int32, device=dev)
sp = torch.cuda.current_stream().cuda_stream
evs = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
evs[0].record()
for it in range(a.iters):
    b.renew(sp)
    b.demodulate_device(x.data_ptr(), bench.N_SAMPLES, bench.N_SAMPLES, d_out.data_ptr(), cap, d_len.data_ptr(), stream=sp, flags=a.flags)
    evs[it + 1].record()
torch.cuda.synchronize()
print("ms per call", [round(evs[i].elapsed_time(evs[i + 1]), 2) for i in range(a.iters)])
print("done", b.fast_stats(), b.launch_count())
for g in (0, 1):
    nc, w = b.debug_fast_windows(g)
    print("group", g, "windows per class", [sum(1 for x in w if x["cls"] == c) for c in range(nc)],
          "not confirmed / stages:", [(x["cls"], x["stream"], x["slab"], hex(x["result"])) for x in w if x["result"] != 1 or x["cls"] >= nc - 2])
