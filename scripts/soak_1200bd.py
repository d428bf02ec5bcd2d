"""Soak (GPU): the fast path against the float64 kernels at another operating point — 48 kHz / 1200 Bd (FSKCore's default
configuration), 32,768 streams x 16,384 samples, one 30-byte frame per stream at a random offset, AWGN -15..+30 dB, in
two calls of 8,192 samples so that state, rings and open readings carry over.  Every stream: bytes and counters equal.
usage: python scripts/soak_1200bd.py [--seeds 1:9]"""
import argparse, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: F401  (sets CUDA_DEVICE_MAX_CONNECTIONS before the CUDA context exists)
import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--seeds", default="1:9")
ap.add_argument("--out", default="")
a = ap.parse_args()
lo, hi = (int(v) for v in a.seeds.split(":"))
wam = importlib.import_module("webaudio-modem_b200")
L = importlib.import_module("webaudio-modem_b200._lib")
lib = wam.lib()
dev = torch.device("cuda", 0)
S, N, PAY = 32768, 16384, 30
KEYS = ("syncDetections", "eodEvents", "globalSampleCounter", "frameStarted", "receivedBitsLength")
lines = []
for seed in range(lo, hi):
    rng = np.random.Generator(np.random.Philox(seed))
    payload = rng.integers(0, 256, (S, PAY), dtype=np.uint8)
    mb = wam.FSKBatch(S, {})
    d_data = torch.from_numpy(payload).to(dev)
    frames = torch.zeros((S, N), dtype=torch.float32, device=dev)
    sp = torch.cuda.current_stream().cuda_stream
    mb.modulate_device(d_data.data_ptr(), PAY, PAY, frames.data_ptr(), N, stream=sp)
    torch.cuda.synchronize()
    mb.close()
    off = torch.from_numpy(rng.integers(0, 2500, S)).to(dev)
    idx = torch.arange(N, device=dev)[None, :] - off[:, None]
    x = torch.gather(frames, 1, idx.clamp(min=0)) * (idx >= 0)
    del frames, idx
    snr = np.resize(np.arange(-15.0, 31.0, 3.0), S)
    sigma = torch.from_numpy(np.sqrt(0.5 / 10.0 ** (snr / 10.0)).astype(np.float32)).to(dev)
    x = x.contiguous()
    assert lib.wam_awgn_add_device(x.data_ptr(), N, S, N, sigma.data_ptr(), 0xB200 + seed, 0, sp) == 0
    torch.cuda.synchronize()
    res = {}
    for name, fl in (("fast", L.WAM_BATCH_FORCE_FAST), ("exact", L.WAM_BATCH_EXACT_ONLY)):
        b = wam.FSKBatch(S, {})
        got = [b""] * S
        for h in (0, 1):
            n = N // 2
            cap = b.out_capacity(n)
            d_out = torch.zeros((S, cap), dtype=torch.uint8, device=dev)
            d_len = torch.zeros(S, dtype=torch.int32, device=dev)
            xs = x[:, h * n:(h + 1) * n].contiguous()
            b.demodulate_device(xs.data_ptr(), n, n, d_out.data_ptr(), cap, d_len.data_ptr(), stream=sp, flags=fl)
            torch.cuda.synchronize()
            ho, hl = d_out.cpu().numpy(), d_len.cpu().numpy()
            got = [g + bytes(ho[i, :hl[i]]) for i, g in enumerate(got)]
        st = b.status()
        res[name] = (got, [tuple(float(s[k]) for k in KEYS) for s in st], b.fast_stats())
        b.close()
    bad = [i for i in range(S) if res["fast"][0][i] != res["exact"][0][i] or res["fast"][1][i] != res["exact"][1][i]]
    fs = res["fast"][2]
    d = {"seed": seed, "streams": S, "differing": len(bad), "first": bad[:5], "decoded_bytes": sum(len(g) for g in res["exact"][0]),
         "frames_intact": sum(1 for i in range(S) if res["exact"][0][i] == payload[i].tobytes()),
         "fast_calls": fs["fast_calls"], "windows": fs["windows_confirmed"] + fs["windows_refuted"], "refuted": fs["windows_refuted"],
         "rerun_streams_last_call": fs["flagged_last_call"], "carried_settled": fs["carried_settled"], "carried_corrected": fs["carried_corrected"],
         "flag_causes": fs["flag_causes"], "error_flags": fs["error_flags"]}
    print(json.dumps(d), flush=True)
    lines.append(d)
if a.out:
    with open(a.out, "w") as f:
        for d in lines:
            f.write(json.dumps(d) + "\n")
print(json.dumps({"seeds": len(lines), "total_differing": sum(d["differing"] for d in lines)}))
