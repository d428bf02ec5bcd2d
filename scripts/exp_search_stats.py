"""Experiment: how the frame searches of config 2 are distributed (needs a -DWAM_SEARCH_STATS build as WAM_LIB)."""
import ctypes as C, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
wam = importlib.import_module("webaudio-modem_b200")
dev = torch.device("cuda", 0)
N = 48000
S = 65536
x, cfg_index, snr, payloads = bench.generate_on_device(wam, torch, dev, S, seed=1000)
raw = C.CDLL(os.environ["WAM_LIB"])
levels = sorted(set(snr.tolist()))
ONLY = [float(a) for a in sys.argv[1:]]
per = S // len(levels)
for li, level in enumerate(levels):
    if ONLY and level not in ONLY: continue
    lo, hi = li * per, (li + 1) * per
    ci = int(cfg_index[lo])
    assert (cfg_index[lo:hi] == ci).all()
    b = wam.FSKBatch(per, [[bench.CFG_CH1, bench.CFG_CH2][ci]], None)
    cap = b.out_capacity(N)
    out = torch.zeros((per, cap), dtype=torch.uint8, device=dev); ln = torch.zeros(per, dtype=torch.int32, device=dev)
    raw.wam_debug_search_stats(None, 1)
    b.demodulate_device(x[lo:hi].data_ptr(), N, N, out.data_ptr(), cap, ln.data_ptr(), flags=wam._lib.WAM_BATCH_NO_PIPELINE)
    torch.cuda.synchronize()
    st = (C.c_ulonglong * 52)()
    raw.wam_debug_search_stats(st, 0)
    st = np.array(list(st), dtype=np.float64)
    warps = per // 32
    print(f"snr {level:+4.0f}: warp calls/warp {st[0]/warps:7.1f} (check periods per stream {N/2/20:.0f}), lane searches/stream {st[1]/per:7.1f}, "
          f"lanes/call {st[1]/max(st[0],1):5.1f}\n   calls by active lanes 1..32: " + " ".join(f"{int(h/warps)}" for h in st[5:37]) +
          "\n   calls per warp by time (tenths of the second): " + " ".join(f"{int(h/warps)}" for h in st[40:52]))
    b.close()
