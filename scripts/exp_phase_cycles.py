"""Experiment: per-phase SM cycles of the demod kernel (clock64 instrumentation) at several occupancies."""
import ctypes as C, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
wam = importlib.import_module("webaudio-modem_b200")
dev = torch.device("cuda", 0)
N = 48000
for S in [int(a) for a in sys.argv[1:]] or [4736, 65536]:
    x, cfg_index, snr, payloads = bench.generate_on_device(wam, torch, dev, S, seed=1000)
    if os.environ.get('ONE_CFG'):
        b = wam.FSKBatch(S, [bench.CFG_CH2], None)
    else:
        b = wam.FSKBatch(S, [bench.CFG_CH1, bench.CFG_CH2], cfg_index)
    cap = b.out_capacity(N)
    out = torch.zeros((S, cap), dtype=torch.uint8, device=dev); ln = torch.zeros(S, dtype=torch.int32, device=dev)
    for _ in range(2):
        b.renew(0); b.demodulate_device(x.data_ptr(), N, N, out.data_ptr(), cap, ln.data_ptr())
    wam.lib().wam_fsk_batch_debug_phase_cycles(b._h, 1, None, None, 0)
    b.renew(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); b.demodulate_device(x.data_ptr(), N, N, out.data_ptr(), cap, ln.data_ptr()); e1.record()
    torch.cuda.synchronize()
    o = np.zeros(4)
    warps = (S + 31) // 32
    pc = np.zeros((warps, 4))
    dp = C.POINTER(C.c_double)
    wam.lib().wam_fsk_batch_debug_phase_cycles(b._h, 0, o.ctypes.data_as(dp), pc.ctypes.data_as(dp), warps)
    lv = snr[::32][:warps]
    for level in sorted(set(lv.tolist())):
        m = pc[lv == level] / N
        print(f"   snr {level:+5.0f} dB: cycles/sample A1 {m[:,0].mean():5.0f} A2 {m[:,1].mean():5.0f} B {m[:,2].mean():5.0f} other {m[:,3].mean():4.0f} total {m.sum(1).mean():6.0f} (max {m.sum(1).max():6.0f})")
    per = o / warps / N
    print(f"S={S} warps/SM={warps/148:.1f} kernel {e0.elapsed_time(e1):.2f} ms | cycles per warp-sample: A1 {per[0]:.0f} A2 {per[1]:.0f} B {per[2]:.0f} other {per[3]:.0f} total {per.sum():.0f}")
    b.close(); del x
