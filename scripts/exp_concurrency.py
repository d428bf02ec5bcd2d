"""Experiment: does the two-config batch overlap its two launches?  Compare against one config."""
import importlib, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
wam = importlib.import_module("webaudio-modem_b200")
S, N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 48000
dev = torch.device("cuda", 0)
x = (torch.randn((S, N), device=dev) * 0.3).contiguous()
CH1 = dict(baudRate=300, markFrequency=980, spaceFrequency=1180)
CH2 = dict(baudRate=300, markFrequency=1650, spaceFrequency=1850)
for name, cfgs, idx in (("one config", [CH2], None), ("two configs", [CH1, CH2], np.repeat([0, 1], S // 2).astype(np.int32))):
    b = wam.FSKBatch(S, cfgs, idx)
    cap = b.out_capacity(N)
    out = torch.zeros((S, cap), dtype=torch.uint8, device=dev); ln = torch.zeros(S, dtype=torch.int32, device=dev)
    for st_name, st in (("legacy", 0), ("torch stream", torch.cuda.Stream())):
        sp = st.cuda_stream if st else 0
        for _ in range(2):
            b.renew(sp); b.demodulate_device(x.data_ptr(), N, N, out.data_ptr(), cap, ln.data_ptr(), stream=sp)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            b.renew(sp); b.demodulate_device(x.data_ptr(), N, N, out.data_ptr(), cap, ln.data_ptr(), stream=sp)
        torch.cuda.synchronize()
        print(f"{name:12s} {st_name:12s} {(time.perf_counter()-t0)/3*1e3:8.2f} ms/step")
    b.close()
