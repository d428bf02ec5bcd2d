#!/usr/bin/env python
"""bench.py — FSK demodulation throughput on B200 (BASELINE.json metric) + CPU reference arm.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|3|4|5] [--impl reference]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

--config 2 (default, the configuration BASELINE.json's metric is quoted on): ITU-T V.21 both channels (980/1180 and
  1650/1850 Hz), 300 Bd, 48 kHz, 65,536 independent 1 s streams per GPU (12.58 GB float32), one 25-byte frame per
  stream at a random offset, AWGN -15..+30 dB in 3 dB steps (4096 streams per level).  N > 1: every rank runs the
  same per-GPU workload on its own streams (weak scaling, no data-path collective).
--config 3: 1200 Bd at 48 kHz, 16,384 streams x 60 s (188.7 GB: time slabs of 2.5 s, state carried on the device),
  128-byte frames back to back with 0..2000-sample gaps, +6 dB.  N > 1: the streams are sharded (strong scaling).
--config 4: 1,024 streams x 10 min at 44.1 kHz / 1200 Bd, parity 'even', tones shifted -20..+20 Hz per stream,
  +9 dB; sharded over the ranks.
--config 5: 1,000,000 XModem packets (134 bytes -> 55,280 samples at 48 kHz / 1200 Bd): serialize -> modulate ->
  AWGN (-15..+30 dB) -> demodulate -> SOH / seq / len / CRC-16 check, all inside the timed region; sharded over the
  ranks, 125,000 packets per pass.

One JSON line on stdout (rank 0).  `value` = whole-job rate with inputs resident in HBM; `e2e` = the same metric
through the HOST-buffer C-ABI call (H2D + D2H inside the timed region).  Synthetic data (own modulator kernel + Philox
noise) is generated on the device outside the timed regions, except config 5 where it is the workload.
Parity is part of the run and FATAL (exit code 1, "value": null): on config 2 the timed (mixed-precision) result is
compared with the float64 kernels on every stream and with the CPU oracle on a 4,096-stream sample; configs 3 / 5
compare an oracle subset (256 streams / 10,000 packets).  --no-parity-fatal keeps the numbers for experiments.
"""
from __future__ import annotations

import argparse
import hashlib
import importlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

# before anything creates a CUDA context: one hardware work queue per stream of the fast path's float64 checks
# (webaudio-modem_b200/__init__.py says why)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

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
ORACLE_SAMPLE_STREAMS = 4096          # config 2 oracle check: 256 streams per SNR level


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def kernel_source_hash():
    """sha256 over the CUDA sources the library is built from (what the committed ncu numbers belong to)."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "webaudio-modem_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".inl")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic_bytes():
    """DRAM bytes (read + write) of the demodulator launches of ONE config 2 step from the committed ncu launch list
    profiles/r02_demod_traffic.json = {"source_hash", "bytes_per_step", "launches"}; None when the sources have changed
    since (a stale figure is not reported)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_demod_traffic.json")) as f:
            d = json.load(f)
        if d.get("source_hash") != kernel_source_hash():
            return None, 0, "profiles/r02_demod_traffic.json belongs to other kernel sources"
        return float(d["bytes_per_step"]) / max(int(d["launches"]), 1), int(d["launches"]), "ncu dram__bytes_read+write, " + d.get("how", "")
    except Exception as e:  # noqa: BLE001
        return None, 0, f"no committed ncu traffic ({e.__class__.__name__})"


PCM_FULL_SCALE = 32.0  # config 2 as 16-bit PCM: +-32 spans the -15 dB streams' noise peaks


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------
# synthetic workload, config 2
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
def cpu_reference_run(n_sample_streams, steps, warmup, threads, seed=1):
    import oracle as O

    O.build()
    x, cfg_index, snr, payloads = generate_on_host(n_sample_streams, seed=seed)
    times = []
    for it in range(warmup + steps):
        xi = x.copy()
        t0 = time.perf_counter()
        res, st = O.batch_demodulate([CFG_CH1, CFG_CH2], cfg_index, xi, n_threads=threads, want_status=True)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = n_sample_streams * N_SAMPLES
    bits = sum(len(r) for r in res) * 8
    cpu_reference_run.last = (x, cfg_index, res, st)  # for the GPU-vs-oracle check of bench's cpu_baseline leg
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
# GPU arm: shared plumbing
# ---------------------------------------------------------------------------------------------
class Ctx:
    """torch / torch.distributed plumbing of one rank."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.args, self.torch, self.dist = args, torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback); use --impl reference for the CPU arm")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        build = importlib.import_module("webaudio-modem_b200.build")
        build.build()
        self.wam = importlib.import_module("webaudio-modem_b200")
        self.L = importlib.import_module("webaudio-modem_b200._lib")
        self.lib = self.wam.lib()
        # host threads and the staging buffers they allocate live next to this rank's GPU (NUMA); the CPU legs put
        # the full mask back (all_cpus)
        self.all_cpus = os.sched_getaffinity(0)
        self.near_cpus = 0 if args.no_affinity else int(self.lib.wam_host_bind_near_device(self.local_rank))
        self.stream = torch.cuda.current_stream()
        self.sp = self.stream.cuda_stream
        self.peak, self.peak_src = measured_peak_gbs()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v: float) -> float:
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v: float) -> float:
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def finish(self, line, parity_ok: bool):
        if self.rank == 0:
            if not parity_ok:
                line["parity_failed"] = True
                if not self.args.no_parity_fatal:
                    line["value_unverified"] = line.get("value")
                    line["value"] = None
            print(json.dumps(line), flush=True)
        if self.world > 1:
            self.dist.destroy_process_group()
        return 0 if (parity_ok or self.args.no_parity_fatal) else 1

    def roof(self, bytes_, ms, **extra):
        a = bytes_ / (ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": a, "peak": self.peak, "unit": "GB/s", "frac": a / self.peak,
                "peak_source": self.peak_src, **extra}


def status_key(st):
    return tuple(float(st[k]) for k in ("syncDetections", "eodEvents", "globalSampleCounter", "frameStarted"))


# ---------------------------------------------------------------------------------------------
# config 2
# ---------------------------------------------------------------------------------------------
def run_config2(c: Ctx):
    args, torch, wam, L = c.args, c.torch, c.wam, c.L
    S = args.streams
    x, cfg_index, snr, payloads = generate_on_device(wam, torch, c.dev, S, seed=1000 + c.rank)
    batch = wam.FSKBatch(S, [CFG_CH1, CFG_CH2], cfg_index, device=c.local_rank)
    cap = batch.out_capacity(N_SAMPLES)
    d_out = torch.zeros((S, cap), dtype=torch.uint8, device=c.dev)
    d_len = torch.zeros(S, dtype=torch.int32, device=c.dev)
    sp = c.sp

    def demod(flags, out=d_out, ln=d_len):
        batch.demodulate_device(x.data_ptr(), N_SAMPLES, N_SAMPLES, out.data_ptr(), cap, ln.data_ptr(), stream=sp, flags=flags)

    for _ in range(max(args.warmup, 3)):
        batch.renew(sp)
        demod(args.demod_flags)
    c.barrier()

    # ---- timed region: exactly K steps, CUDA events on the launching stream ----------------
    sampler = ClockSampler(c.local_rank)
    if c.rank == 0:
        sampler.start()
    l0 = batch.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    c.barrier()
    ev0.record(c.stream)
    for k in range(args.steps):
        batch.renew(sp)
        kev[k][0].record(c.stream)
        demod(args.demod_flags)
        kev[k][1].record(c.stream)
    ev1.record(c.stream)
    c.barrier()
    launches = batch.launch_count() - l0
    clocks = sampler.stop() if c.rank == 0 else None
    ms_total_max = c.max_over_ranks(ev0.elapsed_time(ev1))
    kernel_ms = [a.elapsed_time(b) for a, b in kev]
    fast = batch.fast_stats()

    lens = d_len.cpu().numpy()
    outs = d_out.cpu().numpy()
    gst = batch.status()
    decoded_bytes = int(lens.sum())
    ok = np.array([lens[s] >= PAYLOAD and bytes(outs[s, :PAYLOAD]) == payloads[s].tobytes() for s in range(S)])
    hi_snr = snr >= 6
    frac_ok_hi = float(ok[hi_snr].mean()) if hi_snr.any() else None

    # ---- parity, untimed: (1) the float64 kernels alone on every stream of this rank --------------
    parity, parity_ok = {}, True
    d_out2, d_len2 = torch.zeros_like(d_out), torch.zeros_like(d_len)
    batch.renew(sp)
    demod(L.WAM_BATCH_EXACT_ONLY, d_out2, d_len2)
    torch.cuda.synchronize()
    lens2, outs2, est = d_len2.cpu().numpy(), d_out2.cpu().numpy(), batch.status()
    differ = [s for s in range(S) if lens2[s] != lens[s] or bytes(outs2[s, :lens[s]]) != bytes(outs[s, :lens[s]])
              or status_key(est[s]) != status_key(gst[s])]
    n_differ = int(c.sum_over_ranks(len(differ)))
    parity["timed_path_vs_float64_kernels"] = {"streams": int(S * c.world), "differing": n_differ,
                                               "compared": "decoded bytes, lengths, syncDetections, eodEvents, "
                                                           "globalSampleCounter, frameStarted of every stream"}
    parity_ok = parity_ok and n_differ == 0
    del d_out2, d_len2

    samples_per_step = S * N_SAMPLES * c.world
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

        def e2e_step():
            batch.renew(0)
            rc = c.lib.wam_fsk_batch_demodulate(batch._h, hx.data_ptr(), N_SAMPLES, N_SAMPLES, h_out.data_ptr(), cap,
                                                h_len.data_ptr(), 0)
            if rc != 0:
                raise RuntimeError(c.lib.wam_last_error().decode())

        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        e2e_step()
        c.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = c.max_over_ranks(time.perf_counter() - t0)
        e2e_same = int(h_len.numpy().sum()) == decoded_bytes and bool((h_len.numpy() == lens).all())
        parity["e2e_equals_device_run"] = {"identical": bool(e2e_same)}
        parity_ok = parity_ok and e2e_same
        e2e = {"value": samples_per_step * e2e_steps / dt / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": int(S * N_SAMPLES * 4), "d2h_bytes_per_step": int(S * cap + S * 4),
               "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps,
               "api": "wam_fsk_batch_demodulate (host buffers, pinned)",
               "host_cpus_near_gpu": c.near_cpus}

        # ---- the same streams as 16-bit PCM (full scale = +-32 so that the -15 dB noise fits): half the PCIe bytes.
        # Checked against a device-resident float64 run on the identical widened samples.
        x = torch.empty((S, N_SAMPLES), dtype=torch.float32, device=c.dev)
        x.copy_(hx)
        x.mul_(32768.0 / PCM_FULL_SCALE).round_().clamp_(-32768.0, 32767.0)
        d_pcm = x.to(torch.int16)
        hp = hx.view(torch.int16).view(-1)[: S * N_SAMPLES].view(S, N_SAMPLES)  # reuse the pinned pages
        hp.copy_(d_pcm)
        x.copy_(d_pcm)
        x.mul_(1.0 / 32768.0)
        del d_pcm
        batch.renew(sp)
        demod(L.WAM_BATCH_EXACT_ONLY)
        torch.cuda.synchronize()
        pcm_len = d_len.cpu().numpy().copy()
        pcm_out = d_out.cpu().numpy().copy()
        del x

        def pcm_step():
            batch.renew(0)
            rc = c.lib.wam_fsk_batch_demodulate_pcm16(batch._h, hp.data_ptr(), N_SAMPLES, N_SAMPLES, h_out.data_ptr(), cap,
                                                      h_len.data_ptr(), 0)
            if rc != 0:
                raise RuntimeError(c.lib.wam_last_error().decode())

        pcm_step()
        c.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            pcm_step()
        torch.cuda.synchronize()
        dt = c.max_over_ranks(time.perf_counter() - t0)
        ho, hl = h_out.numpy(), h_len.numpy()
        pcm_same = bool((hl == pcm_len).all()) and all(
            bytes(ho[i, : hl[i]]) == bytes(pcm_out[i, : pcm_len[i]]) for i in range(S))
        parity["e2e_pcm16_equals_device_run_on_widened_samples"] = {"identical": bool(pcm_same), "streams": S,
                                                                   "decoded_bytes": int(pcm_len.sum())}
        parity_ok = parity_ok and pcm_same
        e2e["pcm16"] = {"value": samples_per_step * e2e_steps / dt / 1e6, "unit": "Msamples/s",
                        "h2d_bytes_per_step": int(S * N_SAMPLES * 2), "d2h_bytes_per_step": int(S * cap + S * 4),
                        "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps,
                        "api": "wam_fsk_batch_demodulate_pcm16 (int16 host buffers, pinned; sample = pcm / 32768)",
                        "pcm_full_scale": PCM_FULL_SCALE}

    if c.rank != 0:
        return c.finish(None, parity_ok)

    k_ms = statistics.mean(kernel_ms)
    achieved = S * N_SAMPLES * BYTES_PER_SAMPLE / (k_ms * 1e-3) / 1e9
    lps = launches / args.steps
    traffic, traffic_launches, traffic_src = ncu_traffic_bytes() if S == 65536 else (None, 0, "other stream count")
    used_fast = fast["fast_calls"] > 0
    line = {
        "metric": "fsk_demod_msamples_per_s", "value": value, "unit": "Msamples/s",
        "n_gpus": c.world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32+f64" if used_fast else "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(s=S), "streams_per_gpu": S, "samples_per_stream": N_SAMPLES,
                   "l2": "inputs (12.58 GB/GPU at 65536 streams) exceed the 126 MB L2; no flush needed",
                   "timed_step": "renew state (configure) + demodulate, inputs resident in HBM"},
        "decoded_bits_per_s": decoded_bytes * 8 * c.world / (ms_total_max * 1e-3 / args.steps),
        "frame_ok_frac_snr_ge_6dB": frac_ok_hi,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": c.peak, "unit": "GB/s", "frac": achieved / c.peak,
                     "traffic": traffic, "traffic_unit": "bytes per launch (mean over the demodulator launches of one step)",
                     "traffic_launches": traffic_launches, "traffic_source": traffic_src,
                     "peak_source": c.peak_src,
                     "kernel": "fsk_demod_fast_kernel" if used_fast else "fsk_demod_exact_kernel",
                     "launches_per_step": lps,
                     "achieved_per_launch_bytes": S * N_SAMPLES * BYTES_PER_SAMPLE / max(lps, 1.0),
                     "kernel_ms_per_launch": k_ms / max(lps, 1.0), "kernel_ms": k_ms,
                     "note": "a step is one demodulate call: float32 kernel with certified decisions in time-slab launches "
                             "(one per configuration group and slab, overlapping on two streams per group), float64 runs "
                             "of the windows around doubtful decisions, small copy kernels; kernel_ms = CUDA-event time "
                             "of the whole call, achieved = 4 B per input sample x samples of the call / kernel_ms; the "
                             "kernel is instruction-issue / latency bound, see profiles/r02_notes.md"},
        "fast_path": {**fast, "doubtful_decision_windows_per_step": fast["windows_confirmed"] + fast["windows_refuted"],
                      "streams_rerun_whole_call_per_step": fast["flagged_last_call"],
                      "note": "mixed precision: float32 DSP behind a float64 AGC, every state-machine decision certified "
                              "against the float32 error band; doubtful decisions are re-decided in float64"},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "parity": parity,
    }
    if not args.no_cpu:
        os.sched_setaffinity(0, c.all_cpus)
        cores = host_cores()
        n_cpu = ORACLE_SAMPLE_STREAMS if not args.quick else 512
        total, times, bits = cpu_reference_run(n_cpu, 1, 0, cores, seed=1)
        line["cpu_baseline"] = {"value": total / times[0] / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port",
                                "sample": f"{n_cpu} streams x 1 s of the same workload ({n_cpu // 16} per SNR level), C "
                                          f"float64 port of the reference FSKCore (oracle/), one pthread per host core"}
        # (2) the same sample through the CUDA path — fast kernel forced, and the float64 kernels — against the oracle
        hx2, hcfg, want, ost = cpu_reference_run.last
        dx = torch.from_numpy(hx2).to(c.dev)
        res = {}
        for name, fl in (("fast_path", L.WAM_BATCH_FORCE_FAST), ("float64_kernels", L.WAM_BATCH_EXACT_ONLY)):
            chk = wam.FSKBatch(n_cpu, [CFG_CH1, CFG_CH2], hcfg, device=c.dev.index)
            ccap = chk.out_capacity(N_SAMPLES)
            co = torch.zeros((n_cpu, ccap), dtype=torch.uint8, device=c.dev)
            cl = torch.zeros(n_cpu, dtype=torch.int32, device=c.dev)
            chk.demodulate_device(dx.data_ptr(), N_SAMPLES, N_SAMPLES, co.data_ptr(), ccap, cl.data_ptr(), stream=sp, flags=fl)
            torch.cuda.synchronize()
            con, cln, cst = co.cpu().numpy(), cl.cpu().numpy(), chk.status()
            same = sum(1 for i in range(n_cpu) if bytes(con[i, :cln[i]]) == want[i] and status_key(cst[i]) == status_key(ost[i]))
            res[name] = {"identical": same, "fast_calls": chk.fast_stats()["fast_calls"]}
            chk.close()
            parity_ok = parity_ok and same == n_cpu
        line["parity"]["oracle_sample"] = {"streams": n_cpu, "decoded_bytes": sum(len(w) for w in want), **res,
                                           "compared": "decoded bytes and status counters of every sample stream"}
    return c.finish(line, parity_ok)


# ---------------------------------------------------------------------------------------------
# configs 3 and 4: long streams in time slabs
# ---------------------------------------------------------------------------------------------
def frame_samples(wam, cfg, nbytes):
    cc = wam.normalize_config(cfg)
    spb = int(cc["sampleRate"] // cc["baudRate"])
    bpb = 8 + cc["startBits"] + cc["stopBits"] + (0 if cc["parity"] == "none" else 1)
    tb = len(cc["preamblePattern"]) + len(cc["sfdPattern"]) + nbytes
    return tb * bpb * spb + 2 * spb + bpb * spb


def modulate_rows(c: Ctx, cfg, payloads: np.ndarray, total: int):
    """payloads uint8 [rows, nbytes] -> float32 [rows, total] on the device (chunks of <= 32768 rows)."""
    torch, wam = c.torch, c.wam
    rows, nbytes = payloads.shape
    out = torch.zeros((rows, total), dtype=torch.float32, device=c.dev)
    for lo in range(0, rows, 32768):
        hi = min(rows, lo + 32768)
        mb = wam.FSKBatch(hi - lo, cfg, device=c.local_rank)
        d = torch.from_numpy(payloads[lo:hi].copy()).to(c.dev)
        mb.modulate_device(d.data_ptr(), nbytes, nbytes, out[lo:hi].data_ptr(), total, stream=c.sp)
        torch.cuda.synchronize()
        mb.close()
    return out


def run_long(c: Ctx, which: int):
    """Configs 3 / 4: every stream is one frame + gap repeated in time with fresh noise per slab; the streams of the
    job are sharded over the ranks.  Timed: the demodulate calls of all slabs (synthesis of a slab is not)."""
    args, torch, wam, L = c.args, c.torch, c.wam, c.L
    import oracle as O

    scale = args.scale
    if which == 3:
        cfg = {}
        n_total, seconds, payload_bytes, snr_db, slab_seconds, sub_streams = 16384, 60.0 * scale, 128, 6.0, 2.5, 256
        name = f"config3: 16384 streams x {seconds:g} s, 48 kHz / 1200 Bd, 128 B frames back to back, 0..2000-sample gaps, +6 dB"
        tone_off = np.zeros(n_total, dtype=np.int64)
    else:
        cfg = dict(sampleRate=44100, baudRate=1200, parity="even")
        n_total, seconds, payload_bytes, snr_db, slab_seconds, sub_streams = 1024, 600.0 * scale, 64, 9.0, 30.0, 64
        name = (f"config4: 1024 streams x {seconds:g} s, 44.1 kHz / 1200 Bd, parity even, tone offset -20..+20 Hz, "
                f"random payloads, +9 dB")
        tone_off = np.random.Generator(np.random.Philox(44)).integers(-20, 21, n_total)
    from importlib import import_module
    shard_range = import_module("webaudio-modem_b200.shard").shard_range
    lo, hi = shard_range(n_total, c.rank, c.world)
    n_streams = hi - lo
    fs = int(wam.normalize_config(cfg)["sampleRate"])
    rng = np.random.Generator(np.random.Philox(3))
    payloads = rng.integers(0, 256, (n_total, payload_bytes), dtype=np.uint8)[lo:hi]
    gaps = rng.integers(0, 2001, n_total)[lo:hi]
    tone_off = tone_off[lo:hi]
    flen = frame_samples(wam, cfg, payload_bytes)
    period = torch.from_numpy((flen + gaps).astype(np.int64)).to(c.dev)
    one = torch.zeros((n_streams, flen + 2000), dtype=torch.float32, device=c.dev)
    full = wam.normalize_config(cfg)
    for o in np.unique(tone_off):
        idx = np.nonzero(tone_off == o)[0]
        mcfg = dict(cfg, markFrequency=full["markFrequency"] + int(o), spaceFrequency=full["spaceFrequency"] + int(o))
        one[torch.from_numpy(idx).to(c.dev), :flen] = modulate_rows(c, mcfg, payloads[idx], flen)
    total = int(seconds * fs) // 32 * 32
    slab = int(slab_seconds * fs) // 32 * 32
    sigma = torch.full((n_streams,), float(np.sqrt(0.5 / 10.0 ** (snr_db / 10.0))), dtype=torch.float32, device=c.dev)
    batch = wam.FSKBatch(n_streams, cfg, device=c.local_rank)
    cap = batch.out_capacity(slab)
    d_out = torch.zeros((n_streams, cap), dtype=torch.uint8, device=c.dev)
    d_len = torch.zeros(n_streams, dtype=torch.int32, device=c.dev)
    x = torch.empty((n_streams, slab), dtype=torch.float32, device=c.dev)
    ar = torch.arange(slab, device=c.dev)[None, :]
    sub = min(sub_streams, n_streams) if c.rank == 0 else 0
    sub_slabs = 2 if which == 3 else 1

    def synth(pos, n, seq):
        idx = (ar[:, :n] + pos) % period[:, None]
        x[:, :n] = torch.gather(one, 1, idx)
        del idx
        L.check(c.lib.wam_awgn_add_device(x.data_ptr(), slab, n_streams, n, sigma.data_ptr(), 0xC0F3 + which, seq, c.sp))

    def one_pass(keep):
        """All slabs of the job once; returns (ms of the demodulate calls, decoded bytes, kept (input, bytes) of the subset)."""
        batch.renew(c.sp)
        ms, decoded, pos, seq = 0.0, 0, 0, 0
        kept_x, kept_bytes = [], [b""] * sub
        while pos < total:
            n = min(slab, total - pos)
            synth(pos, n, seq)
            if keep and seq < sub_slabs and sub:
                kept_x.append(x[:sub, :n].cpu().numpy().copy())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(c.stream)
            batch.demodulate_device(x.data_ptr(), slab, n, d_out.data_ptr(), cap, d_len.data_ptr(), stream=c.sp,
                                    flags=args.demod_flags)
            e1.record(c.stream)
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
            lens = d_len.cpu().numpy()
            decoded += int(lens.sum())
            if keep and seq < sub_slabs and sub:
                o = d_out[:sub].cpu().numpy()
                kept_bytes = [kb + bytes(o[i, :lens[i]]) for i, kb in enumerate(kept_bytes)]
            pos += n
            seq += 1
        return ms, decoded, kept_x, kept_bytes

    for _ in range(args.warmup if args.warmup < 3 else 1):  # a pass is seconds long: one warm-up pass
        one_pass(False)
    c.barrier()
    sampler = ClockSampler(c.local_rank)
    if c.rank == 0:
        sampler.start()
    l0 = batch.launch_count()
    t_wall = time.perf_counter()
    passes = [one_pass(k == 0) for k in range(args.steps)]
    c.barrier()
    wall = time.perf_counter() - t_wall
    launches = batch.launch_count() - l0
    clocks = sampler.stop() if c.rank == 0 else None
    ms_step = c.max_over_ranks(statistics.mean(p[0] for p in passes))
    decoded = c.sum_over_ranks(passes[-1][1])
    syncs = c.sum_over_ranks(sum(s["syncDetections"] for s in batch.status()))
    samples = n_total * total
    fast = batch.fast_stats()

    # ---- parity: the subset's first slabs through the oracle (same float32 input, copied back from the device) ----
    parity, parity_ok, cpu_baseline = {}, True, None
    if c.rank == 0 and sub and not args.no_cpu:
        kept_x, kept_bytes = passes[0][2], passes[0][3]
        xs = np.concatenate(kept_x, axis=1)
        t0 = time.perf_counter()
        os.sched_setaffinity(0, c.all_cpus)
        cores = host_cores()
        # one oracle batch per kept slab is not possible (state carries): run stream by stream over the concatenation
        res, _ = O.batch_demodulate([wam.normalize_config(cfg)], None, np.ascontiguousarray(xs), n_threads=cores, want_status=False)
        dt = time.perf_counter() - t0
        same = sum(1 for i in range(sub) if res[i] == kept_bytes[i])
        parity["oracle_subset"] = {"streams": sub, "samples_per_stream": int(xs.shape[1]), "identical": same,
                                   "decoded_bytes": sum(len(r) for r in res)}
        parity_ok = same == sub
        cpu_baseline = {"value": sub * xs.shape[1] / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port",
                        "sample": f"{sub} streams x {xs.shape[1]} samples of this workload, C float64 port of the reference "
                                  f"FSKCore (oracle/), one pthread per host core"}
    # ---- e2e: two slabs through the host-buffer call ----
    e2e = None
    if not args.no_e2e:
        n = min(slab, total)
        synth(0, n, 0)
        hx = torch.empty((n_streams, n), dtype=torch.float32, pin_memory=True)
        hx.copy_(x[:, :n])
        h_out = torch.zeros((n_streams, cap), dtype=torch.uint8, pin_memory=True)
        h_len = torch.zeros(n_streams, dtype=torch.int32, pin_memory=True)
        batch.renew(0)
        reps = 2
        c.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            rc = c.lib.wam_fsk_batch_demodulate(batch._h, hx.data_ptr(), n, n, h_out.data_ptr(), cap, h_len.data_ptr(), 0)
            if rc != 0:
                raise RuntimeError(c.lib.wam_last_error().decode())
        torch.cuda.synchronize()
        dt = c.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": n_total * n * reps / dt / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(n_streams * n * 4),
               "d2h_bytes_per_step": int(n_streams * cap + n_streams * 4), "steps": reps,
               "api": "wam_fsk_batch_demodulate (host buffers, pinned), one time slab per call"}
    batch.close()
    if c.rank != 0:
        return c.finish(None, parity_ok)
    line = {
        "metric": "fsk_demod_msamples_per_s", "value": samples / (ms_step * 1e-3) / 1e6, "unit": "Msamples/s",
        "n_gpus": c.world, "steps": args.steps, "warmup": 1, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32+f64" if fast["fast_calls"] else "f64", "data": "synthetic",
        "config": {"workload": name, "streams": n_total, "streams_per_gpu": n_streams, "seconds_per_stream": seconds,
                   "slab_seconds": slab_seconds, "l2": "a slab of the rank's streams exceeds the 126 MB L2",
                   "timed_step": "the demodulate calls of all time slabs of one pass (state carried on the device); the "
                                 "synthesis of each slab (gather + Philox AWGN kernel) is outside the timed regions",
                   "wall_s_incl_synthesis": wall},
        "decoded_bits_per_s": decoded * 8 / (ms_step * 1e-3), "sync_detections": syncs,
        "roofline": c.roof(samples * 4.0 / c.world, ms_step, kernel="fsk_demod_pipe_kernel / fsk_demod_fast_kernel",
                           note="4 B per input sample of this rank / time of its demodulate calls"),
        "fast_path": fast, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "parity": parity,
    }
    if cpu_baseline:
        line["cpu_baseline"] = cpu_baseline
    return c.finish(line, parity_ok)


# ---------------------------------------------------------------------------------------------
# config 5: XModem packets end to end
# ---------------------------------------------------------------------------------------------
def run_config5(c: Ctx):
    args, torch, wam, L = c.args, c.torch, c.wam, c.L
    import oracle as O

    n_job = int(1_000_000 * args.scale)
    lo, hi = importlib.import_module("webaudio-modem_b200.shard").shard_range(n_job, c.rank, c.world)
    n_rank = hi - lo
    chunk = min(125_000, n_rank)
    cfg = {}
    total = frame_samples(wam, cfg, 134)        # 55,280
    stride = (total + 2048 + 31) // 32 * 32     # noise behind the frame so that the last stop bit is followed by samples
    rng = np.random.Generator(np.random.Philox(5))
    payload_all = rng.integers(0, 256, (chunk, 128), dtype=np.uint8)  # the same payloads in every pass of `chunk` packets
    seq = (np.arange(chunk) % 255 + 1).astype(np.int32)
    pk = np.zeros((chunk, 134), dtype=np.uint8)
    pk[:, 0] = 1; pk[:, 1] = seq; pk[:, 2] = 255 - seq; pk[:, 3] = 128; pk[:, 4:132] = payload_all
    crc = wam.crc16_batch(payload_all, np.full(chunk, 128, dtype=np.int32), device=c.local_rank)
    pk[:, 132] = crc >> 8; pk[:, 133] = crc & 0xFF
    snr = np.array(SNR_LEVELS, dtype=np.float64)[(np.arange(chunk) * len(SNR_LEVELS)) // chunk]
    x = torch.zeros((chunk, stride), dtype=torch.float32, device=c.dev)
    d_pk = torch.from_numpy(pk).to(c.dev)
    sig = torch.from_numpy(np.sqrt(0.5 / 10.0 ** (snr / 10.0))).to(c.dev, torch.float32)
    mods = []
    for a in range(0, chunk, 32768):
        b_ = min(chunk, a + 32768)
        mods.append((a, b_, wam.FSKBatch(b_ - a, cfg, device=c.local_rank)))
    batch = wam.FSKBatch(chunk, cfg, device=c.local_rank)
    cap = batch.out_capacity(stride)
    d_out = torch.zeros((chunk, cap), dtype=torch.uint8, device=c.dev)
    d_len = torch.zeros(chunk, dtype=torch.int32, device=c.dev)
    d_seq = torch.from_numpy(seq).to(c.dev)
    d_res = torch.zeros((chunk, 7), dtype=torch.int32, device=c.dev)
    n_chunks = (n_rank + chunk - 1) // chunk
    stage_ms = {"modulate": 0.0, "awgn": 0.0, "demodulate": 0.0, "check": 0.0}

    def pipeline(seqno, timed):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        evs[0].record(c.stream)
        x.zero_()  # silence around the frame (the modulator writes the frame's samples only)
        for a, b_, mb in mods:
            mb.modulate_device(d_pk[a:b_].data_ptr(), 134, 134, x[a:b_].data_ptr(), stride, stream=c.sp)
        evs[1].record(c.stream)
        L.check(c.lib.wam_awgn_add_device(x.data_ptr(), stride, chunk, stride, sig.data_ptr(), 0x55AA, seqno, c.sp))
        evs[2].record(c.stream)
        batch.renew(c.sp)
        batch.demodulate_device(x.data_ptr(), stride, stride, d_out.data_ptr(), cap, d_len.data_ptr(), stream=c.sp,
                                flags=args.demod_flags)
        evs[3].record(c.stream)
        rc = c.lib.wam_xmodem_batch_check_device(d_out.data_ptr(), cap, d_len.data_ptr(), d_seq.data_ptr(), chunk,
                                                 d_res.data_ptr(), c.sp)
        assert rc == 0
        evs[4].record(c.stream)
        if timed:
            torch.cuda.synchronize()
            for k, nm in enumerate(("modulate", "awgn", "demodulate", "check")):
                stage_ms[nm] += evs[k].elapsed_time(evs[k + 1])
        return evs[0], evs[4]

    # the tail of every row (behind the frame) must be zero before the noise is added: the modulator writes the frame only
    pipeline(0, False)
    torch.cuda.synchronize()
    c.barrier()
    sampler = ClockSampler(c.local_rank)
    if c.rank == 0:
        sampler.start()
    l0 = batch.launch_count() + sum(m[2].launch_count() for m in mods)
    ms_steps = []
    for k in range(args.steps):
        ms = 0.0
        for ch in range(n_chunks):
            e0, e1 = pipeline(k * n_chunks + ch + 1, True)
            ms += e0.elapsed_time(e1)
        ms_steps.append(ms)
    c.barrier()
    launches = batch.launch_count() + sum(m[2].launch_count() for m in mods) - l0
    clocks = sampler.stop() if c.rank == 0 else None
    ms_step = c.max_over_ranks(statistics.mean(ms_steps))
    res = d_res.cpu().numpy()
    ok = res[:, 0] == 0
    by_snr = {int(s): float(ok[snr == s].mean()) for s in SNR_LEVELS if (snr == s).any()}
    outs = d_out.cpu().numpy()
    lens = d_len.cpu().numpy()
    good = np.nonzero(ok)[0]
    off = res[good, 3]
    payload_same = all(bytes(outs[i, o:o + 128]) == payload_all[i].tobytes() for i, o in zip(good[:20000], off[:20000]))
    fast = batch.fast_stats()

    # ---- parity: 10,000 packets of the last pass through the oracle (same float32 input) ----
    parity, parity_ok, cpu_baseline = {"ok_packets_carry_the_sent_payload": bool(payload_same)}, bool(payload_same), None
    if c.rank == 0 and not args.no_cpu:
        sub = min(10_000 if not args.quick else 1024, chunk)
        pick = np.linspace(0, chunk - 1, sub).astype(np.int64)  # all SNR levels
        xs = x[torch.from_numpy(pick).to(c.dev)].cpu().numpy()
        os.sched_setaffinity(0, c.all_cpus)
        cores = host_cores()
        t0 = time.perf_counter()
        want, _ = O.batch_demodulate([wam.normalize_config(cfg)], None, xs, n_threads=cores, want_status=False)
        dt = time.perf_counter() - t0
        same = 0
        for j, i in enumerate(pick):
            got = bytes(outs[i, :lens[i]])
            chk = O.xmodem_check(want[j], int(seq[i]))
            same += int(got == want[j] and chk["status"] == int(res[i, 0]) and (chk["status"] != 0 or chk["payloadOffset"] == int(res[i, 3])))
        parity["oracle_subset"] = {"packets": sub, "identical": same,
                                   "compared": "demodulated bytes, packet check status and payload offset"}
        parity_ok = parity_ok and same == sub
        cpu_baseline = {"value": sub / dt, "unit": "packets/s (demodulate + check only)", "cores": cores, "kind": "port",
                        "sample": f"{sub} packets of this workload (all SNR levels), C float64 port of the reference FSKCore + "
                                  f"XModem receive checks (oracle/), one pthread per host core"}
    # ---- e2e: packets in host memory -> (H2D) modulate, AWGN, demodulate, check -> flags in host memory ----
    e2e = None
    if not args.no_e2e:
        h_pk = torch.from_numpy(pk).pin_memory()
        h_res = torch.zeros((chunk, 7), dtype=torch.int32).pin_memory()
        c.barrier()
        t0 = time.perf_counter()
        reps = 2
        for r in range(reps):
            d_pk.copy_(h_pk, non_blocking=True)
            pipeline(1000 + r, False)
            h_res.copy_(d_res, non_blocking=True)
            torch.cuda.synchronize()
        dt = c.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": chunk * c.world * reps / dt, "unit": "packets/s", "h2d_bytes_per_step": int(pk.nbytes),
               "d2h_bytes_per_step": int(chunk * 7 * 4), "steps": reps,
               "api": "packet bytes in pinned host memory -> wam_fsk_batch_modulate_device / wam_awgn_add_device / "
                      "wam_fsk_batch_demodulate_device / wam_xmodem_batch_check_device -> check results in host memory"}
    batch.close()
    for m in mods:
        m[2].close()
    if c.rank != 0:
        return c.finish(None, parity_ok)
    samples = n_job * stride
    line = {
        "metric": "xmodem_packets_per_s", "value": n_job / (ms_step * 1e-3), "unit": "packets/s",
        "n_gpus": c.world, "steps": args.steps, "warmup": 1, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32+f64" if fast["fast_calls"] else "f64", "data": "synthetic",
        "config": {"workload": f"config5: {n_job} XModem packets, 134 B -> {total} samples @48 kHz/1200 Bd, AWGN -15..+30 dB, "
                               f"modulate -> AWGN -> demodulate -> SOH/seq/len/CRC-16 check",
                   "packets_per_gpu": n_rank, "packets_per_pass": chunk, "samples_per_packet_row": stride,
                   "l2": "a pass (28.7 GB at 125,000 packets) exceeds the 126 MB L2",
                   "timed_step": "all four stages of all passes of the rank, CUDA events"},
        "msamples_per_s_pipeline": samples / (ms_step * 1e-3) / 1e6,
        "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
        "crc_ok_frac_by_snr_db": by_snr,
        "roofline": c.roof(n_rank * stride * 16.0, ms_step, kernel="modulate + awgn + demodulate",
                           note="16 B per sample as built: modulated float32 written once, read and rewritten by the AWGN "
                                "kernel, read once by the demodulator (SURVEY 8d's 8 B assume the noise is added in the "
                                "demodulator's load)"),
        "fast_path": fast, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "parity": parity,
    }
    if cpu_baseline:
        line["cpu_baseline"] = cpu_baseline
    return c.finish(line, parity_ok)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5])
    ap.add_argument("--streams", type=int, default=65536, help="config 2: streams per GPU")
    ap.add_argument("--scale", type=float, default=1.0, help="configs 3-5: fraction of the duration / packet count")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-affinity", action="store_true", help="leave the ranks' CPU affinity alone")
    ap.add_argument("--demod-flags", type=int, default=0, help="A/B experiments: WAM_BATCH_* flags for the timed calls (128 = float64 kernels only)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--quick", action="store_true", help="smaller oracle samples (experiments)")
    ap.add_argument("--no-parity-fatal", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 5 if args.config == 2 else 1
    if args.impl == "reference":
        return run_reference(args)
    c = Ctx(args)
    if args.config == 2:
        return run_config2(c)
    if args.config in (3, 4):
        return run_long(c, args.config)
    return run_config5(c)


if __name__ == "__main__":
    sys.exit(main())
