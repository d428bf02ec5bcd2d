#!/usr/bin/env python
"""bench_configs.py — BASELINE.json configs 3, 4 and 5 on one GPU's share of the work (not the driver's
contract: that is bench.py on config 2).  One JSON line per config.

  config3   1200 Bd at 48 kHz, 16,384 streams x 60 s (188.7 GB of float32: does not fit in HBM at once), processed
            as time slabs with the per-stream state carried on the device between slabs.
  config4   long-stream stress: 1,024 streams x 10 min at 44.1 kHz / 1200 Bd, parity 'even' (integral sync ring,
            SURVEY R10), both tones shifted by a per-stream constant of -20..+20 Hz at modulation time;
            `--share G` runs the 1/G slice one GPU of G would get (streams are sharded, no collective).
  config5   end-to-end physical layer for XModem packets: serialize -> modulate -> AWGN -> demodulate ->
            locate SOH, check seq/~seq/len and CRC-16; 1,000,000 packets over 8 GPUs = 125,000 per GPU.

Signals are synthesised on the device with this repo's modulator kernel plus torch Philox noise; synthesis is
outside the timed regions.  Long streams are built from one period per stream (frame + gap) repeated in time,
with fresh noise per slab.  `--scale f` shortens the durations (config 3/4) or the packet count (config 5) for
quick runs; the JSON says what was run.
"""
from __future__ import annotations

import argparse
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


def roof(bytes_, ms):
    a = bytes_ / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": a, "peak": PEAK, "unit": "GB/s", "frac": a / PEAK, "peak_source": PEAK_SRC}


class Timer:
    def __init__(self):
        self.ms = {}

    def region(self, name):
        t = self

        class R:
            def __enter__(self):
                self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                self.e0.record()

            def __exit__(self, *a):
                self.e1.record()
                torch.cuda.synchronize()
                t.ms[name] = t.ms.get(name, 0.0) + self.e0.elapsed_time(self.e1)

        return R()


def modulate_rows(cfg, payloads: np.ndarray, total: int) -> torch.Tensor:
    """payloads uint8 [rows, nbytes] -> float32 [rows, total] on the device (chunks of <= 32768 rows)."""
    rows, nbytes = payloads.shape
    out = torch.zeros((rows, total), dtype=torch.float32, device=dev)
    sp = torch.cuda.current_stream().cuda_stream
    for lo in range(0, rows, 32768):
        hi = min(rows, lo + 32768)
        mb = wam.FSKBatch(hi - lo, cfg)
        d = torch.from_numpy(payloads[lo:hi].copy()).to(dev)
        mb.modulate_device(d.data_ptr(), nbytes, nbytes, out[lo:hi].data_ptr(), total, stream=sp)
        torch.cuda.synchronize()
        mb.close()
    return out


def frame_samples(cfg, nbytes):
    c = wam.normalize_config(cfg)
    spb = int(c["sampleRate"] // c["baudRate"])
    bpb = 8 + c["startBits"] + c["stopBits"] + (0 if c["parity"] == "none" else 1)
    tb = len(c["preamblePattern"]) + len(c["sfdPattern"]) + nbytes
    return tb * bpb * spb + 2 * spb + bpb * spb


def long_streams(name, cfg, demod_cfg, n_streams, seconds, payload_bytes, snr_db, slab_seconds, mod_groups=None, seed=3):
    """Periodic frames + AWGN, demodulated slab by slab with carried state.  mod_groups: list of
    (stream index array, modulator config) when streams are modulated with different tones."""
    fs = int(wam.normalize_config(cfg)["sampleRate"])
    rng = np.random.Generator(np.random.Philox(seed))
    payloads = rng.integers(0, 256, (n_streams, payload_bytes), dtype=np.uint8)
    gaps = rng.integers(0, 2001, n_streams)
    flen = frame_samples(cfg, payload_bytes)
    period = torch.from_numpy((flen + gaps).astype(np.int64)).to(dev)
    one = torch.zeros((n_streams, flen + 2000), dtype=torch.float32, device=dev)
    if mod_groups is None:
        mod_groups = [(np.arange(n_streams), cfg)]
    for idx, mcfg in mod_groups:
        if len(idx):
            one[torch.from_numpy(idx).to(dev), :flen] = modulate_rows(mcfg, payloads[idx], flen)
    total = int(seconds * fs)
    slab = int(slab_seconds * fs) // 32 * 32
    sigma = float(np.sqrt(0.5 / 10.0 ** (snr_db / 10.0)))
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    batch = wam.FSKBatch(n_streams, demod_cfg)
    cap = batch.out_capacity(slab)
    d_out = torch.zeros((n_streams, cap), dtype=torch.uint8, device=dev)
    d_len = torch.zeros(n_streams, dtype=torch.int32, device=dev)
    x = torch.empty((n_streams, slab), dtype=torch.float32, device=dev)
    sp = torch.cuda.current_stream().cuda_stream
    T = Timer()
    decoded = 0
    good_frames = 0
    pos = 0
    ar = torch.arange(slab, device=dev)[None, :]
    while pos < total:
        n = min(slab, total - pos)
        with T.region("synth"):
            idx = (ar[:, :n] + pos) % period[:, None]
            x[:, :n] = torch.gather(one, 1, idx)
            x[:, :n].add_(torch.randn((n_streams, n), generator=gen, device=dev, dtype=torch.float32), alpha=sigma)
            del idx
        with T.region("demod"):
            batch.demodulate_device(x.data_ptr(), slab, n, d_out.data_ptr(), cap, d_len.data_ptr(), stream=sp)
        lens = d_len.cpu().numpy()
        decoded += int(lens.sum())
        pos += n
    st = batch.status()
    syncs = sum(s["syncDetections"] for s in st)
    batch.close()
    samples = n_streams * total
    ms = T.ms["demod"]
    frames_sent = float((total / (flen + gaps)).sum())
    return {"config": name, "metric": "fsk_demod_msamples_per_s", "value": samples / (ms * 1e-3) / 1e6, "unit": "Msamples/s",
            "ms_demod": ms, "ms_synth_untimed": T.ms["synth"], "n_streams": n_streams, "seconds_per_stream": seconds,
            "samples": samples, "slab_seconds": slab_seconds, "decoded_bits_per_s": decoded * 8 / (ms * 1e-3),
            "decoded_bytes": decoded, "payload_bytes_sent": int(frames_sent * payload_bytes),
            "sync_detections": syncs, "frames_sent": int(frames_sent), "snr_db": snr_db,
            "roofline": roof(samples * 4.0, ms), "warps": (n_streams + 31) // 32,
            "note": "one thread walks one stream in time: with fewer than ~14 warps per SM the kernel runs at the "
                    "dependency-latency floor of a single warp (DESIGN.md 5.1), not at the config 2 rate"}


def config3(scale):
    cfg = {}  # 48 kHz / 1200 Bd defaults (fsk.ts:19-33)
    return long_streams(f"config3: 16384 streams x {60 * scale:g} s, 48 kHz / 1200 Bd, 128 B frames back to back, +6 dB",
                        cfg, cfg, 16384, 60 * scale, 128, 6.0, slab_seconds=2.5)


def config4(scale, share):
    base = dict(sampleRate=44100, baudRate=1200, parity="even")
    n = 1024 // share
    rng = np.random.Generator(np.random.Philox(44))
    off = rng.integers(-20, 21, n)
    groups = []
    for o in np.unique(off):
        c = dict(base, markFrequency=1650 + int(o), spaceFrequency=1850 + int(o))
        groups.append((np.nonzero(off == o)[0], c))
    return long_streams(f"config4: {n} streams (1024 / {share} GPUs) x {600 * scale:g} s, 44.1 kHz / 1200 Bd, parity even, "
                        f"tone offset -20..+20 Hz, +9 dB", base, base, n, 600 * scale, 64, 9.0, slab_seconds=30.0,
                        mod_groups=groups)


def config5(scale):
    n = int(125000 * scale)
    rng = np.random.Generator(np.random.Philox(5))
    payload = rng.integers(0, 256, (n, 128), dtype=np.uint8)
    seq = (np.arange(n) % 255 + 1).astype(np.int32)
    pk = np.zeros((n, 134), dtype=np.uint8)
    pk[:, 0] = 1; pk[:, 1] = seq; pk[:, 2] = 255 - seq; pk[:, 3] = 128; pk[:, 4:132] = payload
    crc = wam.crc16_batch(payload, np.full(n, 128, dtype=np.int32))
    pk[:, 132] = crc >> 8; pk[:, 133] = crc & 0xFF
    cfg = {}
    total = frame_samples(cfg, 134)  # 55,280
    stride = (total + 2048 + 31) // 32 * 32  # tail of noise so the last byte's stop bit is followed by samples
    snr = np.array(bench.SNR_LEVELS, dtype=np.float64)[(np.arange(n) * len(bench.SNR_LEVELS)) // n]
    T = Timer()
    x = torch.zeros((n, stride), dtype=torch.float32, device=dev)
    d_pk = torch.from_numpy(pk).to(dev)
    sp = torch.cuda.current_stream().cuda_stream
    mods = []
    for lo in range(0, n, 32768):
        hi = min(n, lo + 32768)
        mods.append((lo, hi, wam.FSKBatch(hi - lo, cfg)))
    for rep in range(2):  # first pass warms up
        T.ms.pop("modulate", None)
        with T.region("modulate"):
            for lo, hi, mb in mods:
                mb.modulate_device(d_pk[lo:hi].data_ptr(), 134, 134, x[lo:hi].data_ptr(), stride, stream=sp)
    for _, _, mb in mods:
        mb.close()
    gen = torch.Generator(device=dev)
    gen.manual_seed(55)
    sig = torch.from_numpy(np.sqrt(0.5 / 10.0 ** (snr / 10.0))).to(dev, torch.float32)
    with T.region("awgn"):
        for lo in range(0, n, 8192):
            hi = min(n, lo + 8192)
            x[lo:hi].addcmul_(torch.randn((hi - lo, stride), generator=gen, device=dev, dtype=torch.float32), sig[lo:hi, None])
    batch = wam.FSKBatch(n, cfg)
    cap = batch.out_capacity(stride)
    d_out = torch.zeros((n, cap), dtype=torch.uint8, device=dev)
    d_len = torch.zeros(n, dtype=torch.int32, device=dev)
    d_seq = torch.from_numpy(seq).to(dev)
    d_res = torch.zeros((n, 7), dtype=torch.int32, device=dev)
    for rep in range(2):  # first pass warms up; the input is not modified (no AGC write-back)
        batch.renew(sp)
        T.ms.pop("demodulate", None); T.ms.pop("check", None)
        with T.region("demodulate"):
            batch.demodulate_device(x.data_ptr(), stride, stride, d_out.data_ptr(), cap, d_len.data_ptr(), stream=sp)
        with T.region("check"):
            rc = lib.wam_xmodem_batch_check_device(d_out.data_ptr(), cap, d_len.data_ptr(), d_seq.data_ptr(), n, d_res.data_ptr(), sp)
            assert rc == 0
    res = d_res.cpu().numpy()
    ok = res[:, 0] == 0
    by_snr = {int(s): float(ok[snr == s].mean()) for s in bench.SNR_LEVELS if (snr == s).any()}
    # a packet that passed must carry the sent payload
    outs = d_out.cpu().numpy()
    good = np.nonzero(ok)[0]
    off = res[good, 3]
    same = all(bytes(outs[i, o:o + 128]) == payload[i].tobytes() for i, o in zip(good[:20000], off[:20000]))
    batch.close()
    ms = T.ms["modulate"] + T.ms["demodulate"] + T.ms["check"]
    samples = n * stride
    return {"config": f"config5: {n} XModem packets (1,000,000 / 8 GPUs x {scale:g}), 134 B -> {total} samples @48 kHz/1200 Bd, "
                      f"AWGN -15..+30 dB", "metric": "packets_per_s", "value": n / (ms * 1e-3), "unit": "packets/s",
            "ms": {k: round(v, 3) for k, v in T.ms.items()}, "samples_per_packet_row": stride,
            "msamples_per_s_pipeline": samples / (ms * 1e-3) / 1e6,
            "crc_ok_frac_by_snr_db": by_snr, "ok_packets_payload_identical": bool(same),
            "roofline": roof(samples * 8.0, ms),
            "roofline_note": "8 B per sample: modulated float32 written once, read once (SURVEY 8d, stage-separated)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["config5", "config3", "config4"])
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--share", type=int, default=8, help="config 4: number of GPUs the 1,024 streams are sharded over")
    a = ap.parse_args()
    for w in a.which:
        if w == "config3":
            r = config3(a.scale)
        elif w == "config4":
            r = config4(a.scale, a.share)
        elif w == "config5":
            r = config5(a.scale)
        else:
            raise SystemExit(f"unknown config {w}")
        print(json.dumps(r), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
