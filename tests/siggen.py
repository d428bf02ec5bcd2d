"""Synthetic FSK test signals (test infrastructure).  Frames come from the ORACLE modulator so
that GPU-vs-oracle demodulation tests do not depend on the GPU modulator; noise is numpy Philox."""
from __future__ import annotations

import numpy as np

import oracle as O

V21_CH1 = dict(baudRate=300, markFrequency=980, spaceFrequency=1180)
V21_CH2 = dict(baudRate=300, markFrequency=1650, spaceFrequency=1850)
BELL103 = dict(baudRate=300, markFrequency=1070, spaceFrequency=1270)  # lower tone = mark


def modulate(cfg: dict, payload: bytes) -> np.ndarray:
    m = O.FSKCore()
    m.configure(cfg)
    return m.modulateData(payload)


def noisy_streams(cfg: dict, n_streams: int, n_samples: int, payload_len: int, snr_db, seed: int,
                  max_offset: int = 1280, amplitude: float = 1.0):
    """One frame per stream at a random offset, AWGN over the whole stream.
    snr_db: scalar or per-stream array (signal power 0.5*amplitude^2 over noise power, full band)."""
    rng = np.random.Generator(np.random.Philox(seed))
    snr = np.broadcast_to(np.asarray(snr_db, dtype=np.float64), (n_streams,))
    x = np.zeros((n_streams, n_samples), dtype=np.float32)
    payloads = []
    for s in range(n_streams):
        p = rng.integers(0, 256, payload_len, dtype=np.uint8).tobytes()
        payloads.append(p)
        sig = modulate(cfg, p) * amplitude
        off = int(rng.integers(0, max_offset + 1))
        n = min(len(sig), n_samples - off)
        if n > 0:
            x[s, off:off + n] = sig[:n]
        sigma = np.sqrt(0.5 * amplitude * amplitude / (10.0 ** (snr[s] / 10.0)))
        x[s] += (rng.standard_normal(n_samples) * sigma).astype(np.float32)
    return x, payloads


def multi_frame_stream(cfg: dict, n_samples: int, payload_len: int, snr_db: float, seed: int, max_gap: int = 2000,
                       freq_offset_hz: float = 0.0):
    """Back-to-back frames with random gaps + AWGN (configs 3/4 style)."""
    rng = np.random.Generator(np.random.Philox(seed))
    mcfg = dict(cfg)
    if freq_offset_hz:
        full = {**O.DEFAULT_FSK_CONFIG, **cfg}
        mcfg["markFrequency"] = full["markFrequency"] + freq_offset_hz
        mcfg["spaceFrequency"] = full["spaceFrequency"] + freq_offset_hz
    x = np.zeros(n_samples, dtype=np.float32)
    pos = int(rng.integers(0, max_gap + 1))
    payloads = []
    while pos < n_samples:
        p = rng.integers(0, 256, payload_len, dtype=np.uint8).tobytes()
        sig = modulate(mcfg, p)
        n = min(len(sig), n_samples - pos)
        x[pos:pos + n] = sig[:n]
        payloads.append(p)
        pos += len(sig) + int(rng.integers(0, max_gap + 1))
    sigma = np.sqrt(0.5 / (10.0 ** (snr_db / 10.0)))
    x += (rng.standard_normal(n_samples) * sigma).astype(np.float32)
    return x, payloads
