"""GPU tests of the mixed-precision fast path (fsk_demod_fast.cuh + fast_host.inl): float32 kernel with certified
decisions, float64 checks of the doubtful ones.  Whatever the fast path does, bytes and counters must be the oracle's
(= the reference FSKCore's, src/modems/fsk.ts:190-375) on every stream."""
import numpy as np
import pytest

import siggen

pytestmark = pytest.mark.gpu

KEYS = ("syncDetections", "eodEvents", "globalSampleCounter", "frameStarted", "receivedBitsLength")


def _lib(gpu_wam):
    import importlib

    return importlib.import_module("webaudio-modem_b200._lib")


class DeviceBatch:
    """FSKBatch driven through the device-buffer entry point (the fast path needs aligned device rows)."""

    def __init__(self, wam, configs, cfg_index, n_streams):
        import torch

        self.torch = torch
        self.dev = torch.device("cuda", 0)
        self.b = wam.FSKBatch(n_streams, configs, cfg_index)
        self.n = n_streams

    def run(self, x: np.ndarray, flags: int, tap: bool = False):
        torch = self.torch
        n = x.shape[1]
        dx = torch.from_numpy(np.ascontiguousarray(x)).to(self.dev)
        cap = self.b.out_capacity(n)
        d_out = torch.zeros((self.n, max(cap, 1)), dtype=torch.uint8, device=self.dev)
        d_len = torch.zeros(self.n, dtype=torch.int32, device=self.dev)
        d_tap = torch.zeros((self.n, n), dtype=torch.float32, device=self.dev) if tap else None
        self.b.demodulate_device(dx.data_ptr(), n, n, d_out.data_ptr(), max(cap, 1), d_len.data_ptr(),
                                 d_tap=d_tap.data_ptr() if tap else 0, flags=flags)
        torch.cuda.synchronize()
        out, ln = d_out.cpu().numpy(), d_len.cpu().numpy()
        got = [bytes(out[i, :ln[i]]) for i in range(self.n)]
        return (got, d_tap.cpu().numpy()) if tap else got


def _oracle(O, configs, cfg_index, x, chunks=None):
    want, status = [], []
    for i in range(x.shape[0]):
        m = O.FSKCore()
        m.configure(configs[cfg_index[i]] if cfg_index is not None else configs[0])
        xi = x[i].copy()
        if chunks is None:
            want.append(m.demodulateData(xi))
        else:
            out, pos = b"", 0
            for c in chunks:
                out += m.demodulateData(xi[pos:pos + c])
                pos += c
            want.append(out)
        status.append(m.getStatus())
    return want, status


def _check(got, gst, want, ost):
    bad = [i for i in range(len(want)) if got[i] != want[i] or any(float(gst[i][k]) != float(ost[i][k]) for k in KEYS)]
    assert not bad, f"{len(bad)} streams differ from the oracle, first {bad[:5]}"


def _v21_batch(n_streams, seed, n=48000, interleaved=False):
    cfgs = [siggen.V21_CH1, siggen.V21_CH2]
    # the fast path wants every configuration group's streams contiguous (one TMA descriptor per group)
    idx = np.arange(n_streams, dtype=np.int32) % 2 if interleaved else (np.arange(n_streams) >= n_streams // 2).astype(np.int32)
    snr = np.resize(np.arange(-15.0, 31.0, 3.0), n_streams)
    x = np.zeros((n_streams, n), dtype=np.float32)
    for c in (0, 1):
        sel = np.nonzero(idx == c)[0]
        xs, _ = siggen.noisy_streams(cfgs[c], len(sel), n, 25, snr[sel], seed=seed + c)
        x[sel] = xs
    return cfgs, idx, x


def test_fast_path_equals_oracle(gpu_wam, oracle):
    L = _lib(gpu_wam)
    cfgs, idx, x = _v21_batch(256, seed=11)
    want, ost = _oracle(oracle, cfgs, idx, x)
    db = DeviceBatch(gpu_wam, cfgs, idx, 256)
    got = db.run(x, L.WAM_BATCH_FORCE_FAST)
    fs = db.b.fast_stats()
    assert fs["fast_calls"] == 1 and fs["error_flags"] == 0
    _check(got, db.b.status(), want, ost)
    # the same through the float64 kernels only
    de = DeviceBatch(gpu_wam, cfgs, idx, 256)
    assert de.run(x, L.WAM_BATCH_EXACT_ONLY) == want and de.b.fast_stats()["fast_calls"] == 0


@pytest.mark.parametrize("scale", [300.0, 30000.0])
def test_wide_doubt_band_keeps_results(gpu_wam, oracle, scale):
    """A band hundreds of times wider than calibrated flags many decisions: windows are checked in float64 (some are
    refuted or cannot be formed and their streams re-run over the whole call) and nothing changes."""
    L = _lib(gpu_wam)
    cfgs, idx, x = _v21_batch(192, seed=23)
    want, ost = _oracle(oracle, cfgs, idx, x)
    db = DeviceBatch(gpu_wam, cfgs, idx, 192)
    db.b.debug_fast_band(scale)
    got = db.run(x, L.WAM_BATCH_FORCE_FAST)
    fs = db.b.fast_stats()
    assert fs["fast_calls"] == 1
    assert fs["windows_confirmed"] + fs["windows_refuted"] + fs["flagged_last_call"] > 0, fs
    _check(got, db.b.status(), want, ost)


def test_streaming_calls_carry_state(gpu_wam, oracle):
    """Four calls on the same streams (state, rings and doubt tracking carried between fast calls), then the
    float64 kernel and the fast kernel alternate on one batch."""
    L = _lib(gpu_wam)
    cfgs, idx, x = _v21_batch(96, seed=31)
    chunks = [12000, 12000, 12000, 12000]
    want, ost = _oracle(oracle, cfgs, idx, x, chunks)
    for plan in ([L.WAM_BATCH_FORCE_FAST] * 4,
                 [L.WAM_BATCH_FORCE_FAST, L.WAM_BATCH_EXACT_ONLY, L.WAM_BATCH_FORCE_FAST, L.WAM_BATCH_EXACT_ONLY]):
        db = DeviceBatch(gpu_wam, cfgs, idx, 96)
        got = [b""] * 96
        pos = 0
        for c, fl in zip(chunks, plan):
            part = db.run(x[:, pos:pos + c], fl)
            got = [g + p for g, p in zip(got, part)]
            pos += c
        _check(got, db.b.status(), want, ost)
        assert db.b.fast_stats()["fast_calls"] == sum(1 for f in plan if f == L.WAM_BATCH_FORCE_FAST)


@pytest.mark.parametrize("scale", [1.0, 300.0])
def test_open_readings_are_settled_at_the_next_call(gpu_wam, oracle, scale):
    """Streams that end a fast call with float32 readings still open (running vote, silent run, ring bits inside the
    doubt band) keep their last slabs; the next call — fast or float64 — begins by re-reading them in float64.  With a
    wide band nearly every stream carries; results stay the oracle's and nothing is left on record as unverifiable."""
    L = _lib(gpu_wam)
    cfgs, idx, x = _v21_batch(160, seed=41)
    chunks = [9600, 19200, 9600, 9600]
    want, ost = _oracle(oracle, cfgs, idx, x, chunks)
    for plan in ([L.WAM_BATCH_FORCE_FAST] * 4,
                 [L.WAM_BATCH_FORCE_FAST, L.WAM_BATCH_FORCE_FAST, L.WAM_BATCH_EXACT_ONLY, L.WAM_BATCH_FORCE_FAST]):
        db = DeviceBatch(gpu_wam, cfgs, idx, 160)
        db.b.debug_fast_band(scale)
        got = [b""] * 160
        pos = 0
        for c, fl in zip(chunks, plan):
            part = db.run(x[:, pos:pos + c], fl)
            got = [g + p for g, p in zip(got, part)]
            pos += c
        fs = db.b.fast_stats()
        _check(got, db.b.status(), want, ost)
        assert fs["error_flags"] == 0, fs
        if scale > 1.0:
            assert fs["carried_settled"] > 0, fs
            # the float64 re-reading agrees with the open float32 readings nearly always (a doubtful sample is wrong about
            # once in 300): the usual outcome is "cleared", not "corrected"
            assert fs["carried_corrected"] * 20 <= fs["carried_settled"], fs
    # reset() and renew() drop what was carried
    db = DeviceBatch(gpu_wam, cfgs, idx, 160)
    db.b.debug_fast_band(300.0)
    db.run(x[:, :9600], L.WAM_BATCH_FORCE_FAST)
    db.b.renew()
    db.b.debug_fast_band(300.0)
    got = db.run(x, L.WAM_BATCH_FORCE_FAST)
    want1, ost1 = _oracle(oracle, cfgs, idx, x)
    _check(got, db.b.status(), want1, ost1)


def test_interleaved_groups_fall_back_to_float64(gpu_wam, oracle):
    L = _lib(gpu_wam)
    cfgs, idx, x = _v21_batch(64, seed=37, n=24000, interleaved=True)
    want, ost = _oracle(oracle, cfgs, idx, x)
    db = DeviceBatch(gpu_wam, cfgs, idx, 64)
    got = db.run(x, L.WAM_BATCH_FORCE_FAST)
    _check(got, db.b.status(), want, ost)
    assert db.b.fast_stats()["fast_calls"] == 0


def test_unaligned_call_closes_fast_path(gpu_wam, oracle):
    L = _lib(gpu_wam)
    cfgs, idx, x = _v21_batch(64, seed=41, n=24000)
    chunks = [8000, 1000, 15000]  # 1000 is not a whole number of 32-sample tiles
    want, ost = _oracle(oracle, cfgs, idx, x, chunks)
    db = DeviceBatch(gpu_wam, cfgs, idx, 64)
    got = [b""] * 64
    pos = 0
    for c in chunks:
        part = db.run(x[:, pos:pos + c], L.WAM_BATCH_FORCE_FAST)
        got = [g + p for g, p in zip(got, part)]
        pos += c
    _check(got, db.b.status(), want, ost)
    assert db.b.fast_stats()["fast_calls"] == 1  # only the first call qualified


@pytest.mark.parametrize("cfg,payload_len,snr,fast", [
    ({}, 128, 6.0, 1),                                   # config 3: 1200 Bd, back-to-back frames with gaps
    (dict(agcEnabled=False), 64, 12.0, 1),
    (dict(parity="even"), 64, 9.0, 0),                   # sync ring capacity 1430.0000000000002: float64 kernels only
    (dict(baudRate=300, markFrequency=1070, spaceFrequency=1270), 20, 3.0, 1),
    (dict(baudRate=600, markFrequency=1300, spaceFrequency=1700), 40, 6.0, 1),
])
def test_multi_frame_streams(gpu_wam, oracle, cfg, payload_len, snr, fast):
    L = _lib(gpu_wam)
    n = 96000
    xs = [siggen.multi_frame_stream(cfg, n, payload_len, snr, seed=100 + s)[0] for s in range(64)]
    x = np.stack(xs)
    want, ost = _oracle(oracle, [cfg], None, x)
    db = DeviceBatch(gpu_wam, [cfg], None, 64)
    got = db.run(x, L.WAM_BATCH_FORCE_FAST)
    assert db.b.fast_stats()["fast_calls"] == fast
    _check(got, db.b.status(), want, ost)
    assert sum(len(w) for w in want) > 0


def test_clean_signals_with_digital_silence(gpu_wam, oracle):
    """Zero padding (exact zeros) around clean frames: every sample is doubtful there, the checks must carry it."""
    L = _lib(gpu_wam)
    cfg = siggen.V21_CH2
    n = 32768
    x = np.zeros((32, n), dtype=np.float32)
    for s in range(32):
        sig = siggen.modulate(cfg, bytes(range(s, s + 8)))
        x[s, 64 * s:64 * s + len(sig)] = sig[: n - 64 * s]
    want, ost = _oracle(oracle, [cfg], None, x)
    db = DeviceBatch(gpu_wam, [cfg], None, 32)
    got = db.run(x, L.WAM_BATCH_FORCE_FAST)
    _check(got, db.b.status(), want, ost)
    assert all(len(w) == 8 for w in want)


def test_doubt_band_covers_the_float32_error(gpu_wam, oracle):
    """TAP variant: filteredPhaseDiff and the doubt band per decimated sample against the oracle's float64 values."""
    L = _lib(gpu_wam)
    cfgs, idx, x = _v21_batch(64, seed=53)
    db = DeviceBatch(gpu_wam, cfgs, idx, 64)
    got, tap = db.run(x, L.WAM_BATCH_FORCE_FAST | L.WAM_BATCH_FAST_UNGUARDED | L.WAM_BATCH_TAP_FAST_DECISION |
                      L.WAM_BATCH_NO_SLABS, tap=True)
    worst = 0.0
    for i in range(64):
        m = oracle.FSKCore()
        m.configure(cfgs[idx[i]])
        want, oF, _ = m.demodulateTapped(x[i].copy())
        if got[i] != want:
            continue  # resets differ from here on: the per-sample comparison is meaningless
        k = min(len(oF), x.shape[1] // 2)
        F, band = tap[i, 0:2 * k:2].astype(np.float64), tap[i, 1:2 * k:2].astype(np.float64)
        err = np.abs(F - oF[:k])
        worst = max(worst, float(np.max(err / band)))
        wrong = (F > 0) != (oF[:k] > 0)
        assert not np.any(wrong & ~(np.abs(F) < band)), f"stream {i}: a hard bit differs outside the doubt band"
    assert 0.0 < worst < 1.0, worst


def test_time_slabs_next_to_a_foreign_kernel(gpu_wam, oracle):
    """Time-slab launches hand their streams over through a spin on a flag; the spin is bounded, so foreign kernels
    holding SM slots may slow the call down or — at worst — end it with WAM_ERR_SLAB_TIMEOUT on record, never hang
    it.  Both the float64 kernel and the fast kernel, each next to a stream of large matmuls."""
    import torch

    L = _lib(gpu_wam)
    cfgs, idx, x = _v21_batch(128, seed=61, n=32768)
    want, ost = _oracle(oracle, cfgs, idx, x)
    side = torch.cuda.Stream()
    a = torch.randn((8192, 8192), device="cuda", dtype=torch.float16)
    for flags in (L.WAM_BATCH_EXACT_ONLY | L.WAM_BATCH_NO_PIPELINE | L.WAM_BATCH_FORCE_SLABS, L.WAM_BATCH_FORCE_FAST):
        db = DeviceBatch(gpu_wam, cfgs, idx, 128)
        with torch.cuda.stream(side):
            for _ in range(40):
                a = (a @ a).clamp_(-1, 1)
        got = db.run(x, flags)
        torch.cuda.synchronize()
        st = db.b.status()
        if db.b.fast_stats()["error_flags"] & 4:
            continue  # hand-over timed out: reported, not silent
        _check(got, st, want, ost)


def test_local_repair_of_a_data_bit_voted_the_other_way(gpu_wam):
    """BASELINE config 2 at full size, a seed on which the float64 check of a doubtful data-bit vote disagrees with the
    fast pass while all timing agrees: the byte is corrected in place (no whole-call re-run) and every stream equals the
    float64 kernels' result."""
    import os
    import sys

    import torch

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    L = _lib(gpu_wam)
    dev = torch.device("cuda", 0)
    S = 65536
    x, cfg_index, _, _ = bench.generate_on_device(gpu_wam, torch, dev, S, seed=2043)
    res = {}
    for name, fl in (("fast", 0), ("exact", L.WAM_BATCH_EXACT_ONLY)):
        b = gpu_wam.FSKBatch(S, [bench.CFG_CH1, bench.CFG_CH2], cfg_index)
        cap = b.out_capacity(bench.N_SAMPLES)
        d_out = torch.zeros((S, cap), dtype=torch.uint8, device=dev)
        d_len = torch.zeros(S, dtype=torch.int32, device=dev)
        b.demodulate_device(x.data_ptr(), bench.N_SAMPLES, bench.N_SAMPLES, d_out.data_ptr(), cap, d_len.data_ptr(), flags=fl)
        torch.cuda.synchronize()
        st = b.status()
        res[name] = (d_out.cpu().numpy(), d_len.cpu().numpy(), [tuple(float(s[k]) for k in KEYS) for s in st])
        if name == "fast":
            fs = b.fast_stats()
            repaired = [w for g in (0, 1) for w in b.debug_fast_windows(g)[1] if w["result"] & (1 << 29)]
        b.close()
    fo, fl_, fst = res["fast"]
    eo, el, est = res["exact"]
    assert fs["fast_calls"] == 1 and fs["error_flags"] == 0
    assert (fl_ == el).all() and fst == est
    assert all(bytes(fo[i, :fl_[i]]) == bytes(eo[i, :el[i]]) for i in range(S))
    assert repaired, "this seed is known to need a local repair"
    assert fs["windows_refuted"] == 0 and fs["flagged_last_call"] == 0
