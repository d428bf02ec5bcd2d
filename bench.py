#!/usr/bin/env python
"""bench.py — FSK demodulation throughput on B200 (BASELINE.json metric) + CPU reference arm.

Workload (N=1): BASELINE config 2 — ITU-T V.21 both channels (980/1180 and 1650/1850 Hz), 300 Bd,
48 kHz, 65,536 independent 1 s streams per GPU (12.58 GB float32), one 25-byte frame per stream at
a random offset, AWGN swept -15..+30 dB in 3 dB steps (4096 streams per level).  Synthetic data
generated on the device (own modulator kernel + torch Philox noise).  N>1: every rank runs the same
per-GPU workload on its own streams (weak scaling, no data-path collective).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--streams S] [--impl reference]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  `value` = demodulated Msamples/s with inputs resident in HBM;
`e2e` = the same metric through the HOST-buffer C-ABI call (H2D + D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import ctypes
import importlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 48000
N_SAMPLES = 48000
PAYLOAD = 25
CFG_CH1 = dict(baudRate=300, markFrequency=980, spaceFrequency=1180)
CFG_CH2 = dict(baudRate=300, markFrequency=1650, spaceFrequency=1850)
SNR_LEVELS = list(range(-15, 31, 3))  # 16 levels
BYTES_PER_SAMPLE = 4.0                # algorithmic HBM bytes per demodulated input sample
WORKLOAD = "config2: V.21 ch1+ch2 300 Bd 48 kHz, {s} x 1 s streams/GPU, 25 B frame, AWGN -15..+30 dB"


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of the demod kernel over ONE step of this workload (a step is a
    series of time-slab launches), from the committed ncu launch list `profiles/r01_demod_slabs.csv`
    (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:fsk_demod_exact`
    of `bench.py --steps 1 --warmup 0`), as (mean bytes per launch, launches), or (None, 0)."""
    import csv
    p = os.path.join(ROOT, "profiles", "r01_demod_slabs.csv")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
    try:
        tot, ids = 0.0, set()
        for r in csv.reader(open(p)):
            if len(r) < 15 or not r[0].isdigit() or "fsk_demod_exact_kernel" not in r[4]:
                continue
            if r[12] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(r[14].replace(",", "")) * scale[r[13]]
                ids.add(r[0])
        return (tot / len(ids), len(ids)) if ids else (None, 0)
    except Exception:
        return None, 0


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------
# synthetic workload
# ---------------------------------------------------------------------------------------------
def stream_plan(n_streams: int, seed: int):
    """Per-stream channel (first half ch1, second half ch2), SNR level, start offset, payload."""
    rng = np.random.Generator(np.random.Philox(seed))
    cfg_index = np.zeros(n_streams, dtype=np.int32)
    cfg_index[n_streams // 2:] = 1
    per = max(1, n_streams // len(SNR_LEVELS))
    snr = np.array([SNR_LEVELS[min(i // per, len(SNR_LEVELS) - 1)] for i in range(n_streams)], dtype=np.float64)
    offsets = rng.integers(0, 1280, n_streams).astype(np.int64)
    payloads = rng.integers(0, 256, (n_streams, PAYLOAD), dtype=np.uint8)
    return cfg_index, snr, offsets, payloads


def generate_on_device(wam, torch, dev, n_streams, seed):
    """x[n_streams, N_SAMPLES] float32 on the device: frame at offset + AWGN over the whole second."""
    cfg_index, snr, offsets, payloads = stream_plan(n_streams, seed)
    x = torch.zeros((n_streams, N_SAMPLES), dtype=torch.float32, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(0xB200 + seed)
    half = n_streams // 2
    chunk = 4096
    for cfg, lo_all, hi_all in ((CFG_CH1, 0, half), (CFG_CH2, half, n_streams)):
        for lo in range(lo_all, hi_all, chunk):
            hi = min(hi_all, lo + chunk)
            rows = hi - lo
            mb = wam.FSKBatch(rows, cfg, device=dev.index)
            d_data = torch.from_numpy(payloads[lo:hi].copy()).to(dev)
            frames = torch.zeros((rows, N_SAMPLES), dtype=torch.float32, device=dev)
            mb.modulate_device(d_data.data_ptr(), PAYLOAD, PAYLOAD, frames.data_ptr(), N_SAMPLES,
                               stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            mb.close()
            off = torch.from_numpy(offsets[lo:hi]).to(dev)
            idx = torch.arange(N_SAMPLES, device=dev)[None, :] - off[:, None]
            valid = idx >= 0
            shifted = torch.gather(frames, 1, idx.clamp_(min=0)) * valid
            sigma = torch.from_numpy(np.sqrt(0.5 / (10.0 ** (snr[lo:hi] / 10.0)))).to(dev, torch.float32)
            noise = torch.randn((rows, N_SAMPLES), generator=gen, device=dev, dtype=torch.float32)
            x[lo:hi] = shifted + noise * sigma[:, None]
            del frames, idx, valid, shifted, noise
    return x, cfg_index, snr, payloads


def generate_on_host(n_streams, seed):
    """Same statistics on the CPU through the oracle modulator (for --impl reference / cpu_baseline)."""
    import oracle as O

    cfg_index, snr, offsets, payloads = stream_plan(n_streams, seed)
    rng = np.random.Generator(np.random.Philox(seed + 1))
    x = np.zeros((n_streams, N_SAMPLES), dtype=np.float32)
    mods = []
    for cfg in (CFG_CH1, CFG_CH2):
        m = O.FSKCore()
        m.configure(cfg)
        mods.append(m)
    for s in range(n_streams):
        sig = mods[cfg_index[s]].modulateData(payloads[s].tobytes())
        n = min(len(sig), N_SAMPLES - int(offsets[s]))
        x[s, offsets[s]:offsets[s] + n] = sig[:n]
        sigma = np.sqrt(0.5 / (10.0 ** (snr[s] / 10.0)))
        x[s] += (rng.standard_normal(N_SAMPLES) * sigma).astype(np.float32)
    return x, cfg_index, snr, payloads


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [t.strip() for t in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            busy = [v for v in sm if v > 0]
            out.update(sm_mhz=statistics.median(busy or sm), sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------
# CPU reference arm (oracle port of the reference FSKCore, all host threads)
# ---------------------------------------------------------------------------------------------
def cpu_reference_run(n_sample_streams, steps, warmup, threads):
    import oracle as O

    O.build()
    x, cfg_index, snr, payloads = generate_on_host(n_sample_streams, seed=1)
    times = []
    for it in range(warmup + steps):
        xi = x.copy()
        t0 = time.perf_counter()
        res, _ = O.batch_demodulate([CFG_CH1, CFG_CH2], cfg_index, xi, n_threads=threads, want_status=False)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = n_sample_streams * N_SAMPLES
    bits = sum(len(r) for r in res) * 8
    cpu_reference_run.last = (x, cfg_index, res)  # for the GPU-vs-oracle check of bench's cpu_baseline leg
    return total, times, bits


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    threads = cores
    n_streams = max(32, min(4096, threads * 32))  # bounded sample: about 1.5 s of CPU work per step
    total, times, bits = cpu_reference_run(n_streams, args.steps, min(args.warmup, 1), threads)
    t = sum(times)
    value = total * len(times) / t / 1e6
    line = {
        "impl": "reference",
        "metric": "fsk_demod_msamples_per_s", "value": value, "unit": "Msamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(s=65536), "sample": f"{n_streams} streams x {N_SAMPLES} samples per step"},
        "decoded_bits_per_s": bits / (t / len(times)),
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": threads, "kind": "port",
                         "sample": f"{n_streams} streams x 1 s of the same workload per step, C float64 port of the "
                                   f"reference FSKCore (oracle/), one pthread per host core"},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line (NCCL_DEBUG=VERSION prints there)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    build = importlib.import_module("webaudio-modem_b200.build")
    build.build()
    wam = importlib.import_module("webaudio-modem_b200")

    S = args.streams
    x, cfg_index, snr, payloads = generate_on_device(wam, torch, dev, S, seed=1000 + rank)
    batch = wam.FSKBatch(S, [CFG_CH1, CFG_CH2], cfg_index, device=local_rank)
    cap = batch.out_capacity(N_SAMPLES)
    d_out = torch.zeros((S, cap), dtype=torch.uint8, device=dev)
    d_len = torch.zeros(S, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    def step():
        batch.renew(sp)
        batch.demodulate_device(x.data_ptr(), N_SAMPLES, N_SAMPLES, d_out.data_ptr(), cap, d_len.data_ptr(), stream=sp,
                                flags=args.demod_flags)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region: exactly K steps, CUDA events on the launching stream ----------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = batch.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    ev0.record(stream)
    for k in range(args.steps):
        batch.renew(sp)
        kev[k][0].record(stream)
        batch.demodulate_device(x.data_ptr(), N_SAMPLES, N_SAMPLES, d_out.data_ptr(), cap, d_len.data_ptr(), stream=sp,
                                flags=args.demod_flags)
        kev[k][1].record(stream)
    ev1.record(stream)
    barrier()
    launches = batch.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    kernel_ms = [a.elapsed_time(b) for a, b in kev]
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total_max = float(t.item())

    # results of the last step
    lens = d_len.cpu().numpy()
    outs = d_out.cpu().numpy()
    # full-size self-check, untimed: the same call as ONE launch per group set (no time slabs) gives the same bytes
    parity = {}
    if rank == 0 and not (args.demod_flags & wam._lib.WAM_BATCH_NO_SLABS):
        d_out2 = torch.zeros_like(d_out)
        d_len2 = torch.zeros_like(d_len)
        batch.renew(sp)
        batch.demodulate_device(x.data_ptr(), N_SAMPLES, N_SAMPLES, d_out2.data_ptr(), cap, d_len2.data_ptr(), stream=sp,
                                flags=args.demod_flags | wam._lib.WAM_BATCH_NO_SLABS)
        torch.cuda.synchronize()
        lens2 = d_len2.cpu().numpy()
        outs2 = d_out2.cpu().numpy()
        ok_same = bool((lens2 == lens).all()) and all(
            bytes(outs2[s, :lens[s]]) == bytes(outs[s, :lens[s]]) for s in range(S))
        parity["time_slabs_equal_one_pass"] = {"streams": int(S), "identical": ok_same}
        if not ok_same:  # reported, not fatal
            print("bench.py: WARNING: time-slab launches and the one-pass launch decoded different bytes", file=sys.stderr)
        del d_out2, d_len2
    decoded_bytes = int(lens.sum())
    ok = np.array([lens[s] >= PAYLOAD and bytes(outs[s, :PAYLOAD]) == payloads[s].tobytes() for s in range(S)])
    hi_snr = snr >= 6
    frac_ok_hi = float(ok[hi_snr].mean()) if hi_snr.any() else None

    samples_per_step = S * N_SAMPLES * world
    value = samples_per_step * args.steps / (ms_total_max * 1e-3) / 1e6

    # ---- e2e: HOST buffers through the C ABI, H2D/D2H inside the timed region ----------------
    e2e = None
    if not args.no_e2e:
        hx = torch.empty((S, N_SAMPLES), dtype=torch.float32, pin_memory=True)
        hx.copy_(x)
        del x
        torch.cuda.empty_cache()
        h_out = torch.zeros((S, cap), dtype=torch.uint8, pin_memory=True)
        h_len = torch.zeros(S, dtype=torch.int32, pin_memory=True)
        lib = wam.lib()

        def e2e_step():
            batch.renew(0)
            rc = lib.wam_fsk_batch_demodulate(batch._h, hx.data_ptr(), N_SAMPLES, N_SAMPLES, h_out.data_ptr(), cap,
                                              h_len.data_ptr(), 0)
            if rc != 0:
                raise RuntimeError(lib.wam_last_error().decode())

        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        td = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
        dt = float(td.item())
        assert int(h_len.numpy().sum()) == decoded_bytes, "e2e result differs from the device-resident run"
        e2e = {"value": samples_per_step * e2e_steps / dt / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": int(S * N_SAMPLES * 4), "d2h_bytes_per_step": int(S * cap + S * 4),
               "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps,
               "api": "wam_fsk_batch_demodulate (host buffers, pinned)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak_gbs()
    k_ms = statistics.mean(kernel_ms)
    achieved = S * N_SAMPLES * BYTES_PER_SAMPLE / (k_ms * 1e-3) / 1e9
    lps = launches / args.steps
    traffic, traffic_launches = ncu_traffic_bytes() if S == 65536 else (None, 0)
    line = {
        "metric": "fsk_demod_msamples_per_s", "value": value, "unit": "Msamples/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(s=S), "streams_per_gpu": S, "samples_per_stream": N_SAMPLES,
                   "l2": "inputs (12.58 GB/GPU at 65536 streams) exceed the 126 MB L2; no flush needed",
                   "timed_step": "renew state (configure) + demodulate, inputs resident in HBM"},
        "decoded_bits_per_s": decoded_bytes * 8 * world / (ms_total_max * 1e-3 / args.steps),
        "frame_ok_frac_snr_ge_6dB": frac_ok_hi,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_unit": "bytes per launch (mean over the launches of one step, ncu)",
                     "traffic_launches": traffic_launches,
                     "peak_source": peak_src, "kernel": "fsk_demod_exact_kernel",
                     "launches_per_step": lps,
                     "achieved_per_launch_bytes": S * N_SAMPLES * BYTES_PER_SAMPLE / max(lps, 1.0),
                     "kernel_ms_per_launch": k_ms / max(lps, 1.0), "kernel_ms": k_ms,
                     "note": "a step is one demodulate call = a series of overlapping time-slab launches of "
                             "fsk_demod_exact_kernel (2048 samples per stream each, two streams, DESIGN.md 5.1); "
                             "kernel_ms = CUDA-event time of the whole series, achieved = algorithmic 4 B/input sample "
                             "x samples of the call / kernel_ms (= per-launch bytes / per-launch share of that time); "
                             "the kernel is instruction-issue/latency bound (float64, reference-faithful), see "
                             "profiles/r01_notes.md for issue-slot utilisation"},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "parity": parity,
    }
    if not args.no_cpu and world >= 1:
        cores = host_cores()
        n_cpu = max(32, min(2048, cores * 4))
        total, times, bits = cpu_reference_run(n_cpu, 1, 0, cores)
        line["cpu_baseline"] = {"value": total / times[0] / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port",
                                "sample": f"{n_cpu} streams x 1 s of the same workload, C float64 port of the reference "
                                          f"FSKCore (oracle/), one pthread per host core"}
        # the same sample through the CUDA path (checker use of the oracle, outside every timed region)
        hx, hcfg, want = cpu_reference_run.last
        chk = wam.FSKBatch(n_cpu, [CFG_CH1, CFG_CH2], hcfg, device=dev.index)
        got = chk.demodulate_bytes(hx.copy())
        chk.close()
        same = sum(1 for g, w in zip(got, want) if g == w)
        line["parity"]["oracle_sample"] = {"streams": n_cpu, "identical": same, "decoded_bytes": sum(len(w) for w in want)}
        if same != n_cpu:  # reported, not fatal: the line above carries the count
            print(f"bench.py: WARNING: GPU bytes differ from the oracle on {n_cpu - same} of {n_cpu} sample streams",
                  file=sys.stderr)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--streams", type=int, default=65536, help="streams per GPU (config 2: 65536)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--demod-flags", type=int, default=0, help="A/B experiments: WAM_BATCH_* flags for the device-resident step (16 = no TMA)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
