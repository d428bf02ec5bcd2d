"""Experiment: busy SM cycles per input sample of the three roles of fsk_demod_pipe_kernel (build libwam with
-DWAM_PHASE_TIMING).  usage: exp_pipe_cycles.py [n_streams] [baud] [seconds]"""
import ctypes as C, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as bc  # (frame_samples / modulate_rows now take a Ctx: adapt before use)
wam = bc.wam
S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
baud = int(sys.argv[2]) if len(sys.argv) > 2 else 1200
secs = float(sys.argv[3]) if len(sys.argv) > 3 else 5.0
cfg = dict(baudRate=baud)
fs = 48000
N = int(secs * fs) // 32 * 32
rng = np.random.default_rng(0)
flen = bc.frame_samples(cfg, 64)
one = bc.modulate_rows(cfg, rng.integers(0, 256, (S, 64), dtype=np.uint8), flen)
reps = (N + flen + 999) // (flen + 1000) + 1
x = torch.nn.functional.pad(one, (0, 1000)).repeat(1, reps)[:, :N].contiguous()
x.add_(torch.randn_like(x), alpha=float(np.sqrt(0.5 / 10 ** 0.9)))
for flags, name in ((0, "pipe"), (8, "fused")):
    b = wam.FSKBatch(S, cfg)
    cap = b.out_capacity(N)
    out = torch.zeros((S, cap), dtype=torch.uint8, device=bc.dev); ln = torch.zeros(S, dtype=torch.int32, device=bc.dev)
    b.demodulate_device(x.data_ptr(), N, N, out.data_ptr(), cap, ln.data_ptr(), flags=flags)
    wam.lib().wam_fsk_batch_debug_phase_cycles(b._h, 1, None, None, 0)
    b.renew(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); b.demodulate_device(x.data_ptr(), N, N, out.data_ptr(), cap, ln.data_ptr(), flags=flags); e1.record()
    torch.cuda.synchronize()
    o = np.zeros(4)
    wam.lib().wam_fsk_batch_debug_phase_cycles(b._h, 0, o.ctypes.data_as(C.POINTER(C.c_double)), None, 0)
    ctas = (S + 31) // 32
    per = o / ctas / N
    ms = e0.elapsed_time(e1)
    print(f"{name}: S={S} baud={baud} kernel {ms:.2f} ms = {ms*1e-3*1.965e9/N:.0f} clk/sample | busy cycles per sample: "
          f"A1 {per[0]:.0f} A2 {per[1]:.0f} B {per[2]:.0f} {'lifetime' if name=='pipe' else 'other'} {per[3]:.0f}; bytes {int(ln.sum())}")
    b.close()
