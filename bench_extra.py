#!/usr/bin/env python
"""bench_extra.py — device-resident throughput of the other kernels on the path at BASELINE-like sizes
(not the driver's contract: that is bench.py).  One JSON line per kernel with a roofline fraction.

  modulate      config 5 shape: 134-byte XModem packets at 48 kHz / 1200 Bd (55,280 samples each)
  xmodem_check  1,000,000 demodulated packets: SOH scan + seq/~seq/len + warp-level CRC-16
  crc16         1,000,000 blocks of 128 bytes
  iir_scan      time-chunked linear-recurrence scan, 1,024 streams x 1,048,576 samples (host API incl. copies
                is not timed: kernels only, via CUDA events around the three launches)
"""
import ctypes as C
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

wam = importlib.import_module("webaudio-modem_b200")
lib = wam.lib()
dev = torch.device("cuda", 0)
PEAK, PEAK_SRC = bench.measured_peak_gbs()


def timed(fn, warm=3, steps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def line(name, unit, per_step_units, ms, bytes_per_step, extra=None):
    d = {"kernel": name, "value": per_step_units / (ms * 1e-3) / 1e6, "unit": unit, "ms_per_step": ms,
         "roofline": {"bound": "hbm", "achieved": bytes_per_step / (ms * 1e-3) / 1e9, "peak": PEAK, "unit": "GB/s",
                      "frac": bytes_per_step / (ms * 1e-3) / 1e9 / PEAK, "peak_source": PEAK_SRC}}
    d.update(extra or {})
    print(json.dumps(d), flush=True)


def bench_modulate(n_packets=32768):
    rng = np.random.Generator(np.random.Philox(5))
    payload = rng.integers(0, 256, (n_packets, 134), dtype=np.uint8)
    d_data = torch.from_numpy(payload).to(dev)
    total = 55280
    out = torch.empty((n_packets, total), dtype=torch.float32, device=dev)
    b = wam.FSKBatch(n_packets, {})
    sp = torch.cuda.current_stream().cuda_stream
    ms = timed(lambda: b.modulate_device(d_data.data_ptr(), 134, 134, out.data_ptr(), total, stream=sp))
    line("fsk_modulate_fused_kernel", "Msamples/s", n_packets * total, ms, n_packets * total * 4.0,
         {"workload": f"{n_packets} x 134-byte packets, 48 kHz / 1200 Bd, {total} samples each (4 B/sample written)",
          "note": "write-only kernel: the peak is the measured COPY bandwidth (read + write); torch.fill_ on the same buffer "
                  "writes 7.5 TB/s on this box, so a fraction slightly above 1 is possible"})
    b.close()


def bench_frames(n=1_000_000):
    rng = np.random.Generator(np.random.Philox(6))
    rows = rng.integers(0, 256, (n, 136), dtype=np.uint8)
    rows[:, 0] = 1; rows[:, 1] = (np.arange(n) % 255 + 1).astype(np.uint8); rows[:, 2] = 255 - rows[:, 1]; rows[:, 3] = 128
    d_rows = torch.from_numpy(rows).to(dev)
    d_len = torch.full((n,), 134, dtype=torch.int32, device=dev)
    d_seq = torch.from_numpy((np.arange(n) % 255 + 1).astype(np.int32)).to(dev)
    d_res = torch.empty((n, 7), dtype=torch.int32, device=dev)
    sp = torch.cuda.current_stream().cuda_stream
    ms = timed(lambda: lib.wam_xmodem_batch_check_device(d_rows.data_ptr(), 136, d_len.data_ptr(), d_seq.data_ptr(), n,
                                                          d_res.data_ptr(), sp))
    line("xmodem_check_kernel", "Mpackets/s", n, ms, n * (134 + 28.0), {"workload": f"{n} packets of 134 bytes (128-byte payload)"})


def bench_mux(n_sessions=4096, block=128, ticks=200):
    """Session multiplexer (wam_fsk_mux_*): every session pushes one 128-sample render quantum per tick
    (FSKProcessor.process(), 2.67 ms of audio at 48 kHz), one flush per tick = H2D + one ragged batch + D2H."""
    import time
    rng = np.random.default_rng(7)
    mux = wam.FSKSessionMux(n_sessions, {}, max_block=block)
    x = (rng.standard_normal((n_sessions, block)) * 0.1).astype(np.float32)
    lib_ = wam.lib()
    push = lib_.wam_fsk_mux_push
    t_push = t_flush = 0.0
    for t in range(ticks + 20):
        t0 = time.perf_counter()
        for s in range(0, n_sessions):
            push(mux._h, s, x[s].ctypes.data, block)
        t1 = time.perf_counter()
        mux.flush_raw()
        t2 = time.perf_counter()
        if t >= 20:
            t_push += t1 - t0
            t_flush += t2 - t1
    ms_flush = 1e3 * t_flush / ticks
    d = {"kernel": "wam_fsk_mux_flush (H2D + ragged demodulate + D2H)", "n_sessions": n_sessions, "block": block,
         "ms_per_flush": ms_flush, "value": n_sessions * block / (ms_flush * 1e-3) / 1e6, "unit": "Msamples/s",
         "audio_ms_per_tick": 1e3 * block / 48000.0, "realtime_factor": (1e3 * block / 48000.0) / ms_flush,
         "ms_per_tick_pushes_python_loop": 1e3 * t_push / ticks}
    print(json.dumps(d), flush=True)
    mux.close()


def bench_filters(n_streams=65536, n=48000):
    """src/dsp/filters.ts at config-2 size, device-resident: the 2nd-order IIR of FilterFactory (time-chunked scan with
    decoupled look-back, one pass: 4 B read + 4 B written per sample) and the 51-tap sinc FIR (float64 multiply-adds:
    bound by the FP64 pipe, the HBM fraction is reported all the same)."""
    F = importlib.import_module("webaudio-modem_b200.filters")
    x = torch.empty((n_streams, n), dtype=torch.float32, device=dev).normal_()
    y = torch.empty_like(x)
    sp = torch.cuda.current_stream().cuda_stream
    c = F.FilterDesign.butterworthLowpass(300.0, 48000.0)
    b = np.ascontiguousarray(c["b"], dtype=np.float64); a = np.ascontiguousarray(c["a"], dtype=np.float64)
    sb = int(lib.wam_iir_scratch_bytes(n, n_streams))
    scratch = torch.empty(sb, dtype=torch.uint8, device=dev)
    state = torch.zeros((n_streams, 4), dtype=torch.float64, device=dev)

    def iir():
        rc = lib.wam_iir_process_batch_device(b.ctypes.data, len(b), a.ctypes.data, len(a), x.data_ptr(), y.data_ptr(), n, n, n_streams,
                                              state.data_ptr(), scratch.data_ptr(), sb, sp)
        assert rc == 0, lib.wam_last_error()
    ms = timed(iir, steps=5)
    # spot check against the host-API path (itself checked against the reference algorithm in tests/test_gpu_filters.py)
    state.zero_(); iir(); torch.cuda.synchronize()
    want = F.iir_process_batch(b, a, x[:4].cpu().numpy())
    err = float(np.max(np.abs(y[:4].cpu().numpy() - want)))
    line("iir_scan_kernel<2,2>", "Msamples/s", n_streams * n, ms, n_streams * n * 8.0,
         {"workload": f"{n_streams} streams x {n} samples, Butterworth low-pass 300 Hz (biquad), 8 B/sample",
          "max_abs_diff_vs_host_api_first_4_streams": err})
    taps = np.ascontiguousarray(F.FilterDesign.sincLowpass(1000.0, 48000.0, 51), dtype=np.float64)
    d_taps = torch.from_numpy(taps).to(dev)

    def fir():
        rc = lib.wam_fir_process_batch_device(d_taps.data_ptr(), len(taps), x.data_ptr(), y.data_ptr(), n, n, n_streams, None, None, sp)
        assert rc == 0, lib.wam_last_error()
    ms = timed(fir, steps=5)
    want = F.fir_process_batch(taps, x[:4].cpu().numpy())
    err = float(np.max(np.abs(y[:4].cpu().numpy() - want)))
    line("fir_kernel (51 taps)", "Msamples/s", n_streams * n, ms, n_streams * n * 8.0,
         {"workload": f"{n_streams} streams x {n} samples, 51-tap sinc low-pass, 8 B/sample",
          "fp64_pipe": {"dfma_per_sample": 51, "tflops": 2.0 * 51 * n_streams * n / (ms * 1e-3) / 1e12,
                        "note": "51 float64 multiply-adds per sample: bound by the FP64 pipe (~40 TFLOP/s), not by HBM"},
          "max_abs_diff_vs_host_api_first_4_streams": err})
    del x, y


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "filters":
        bench_filters(int(sys.argv[2]) if len(sys.argv) > 2 else 65536)
        return
    bench_filters()
    bench_modulate()
    bench_frames()
    bench_mux(4096)
    bench_mux(32768)


if __name__ == "__main__":
    main()
