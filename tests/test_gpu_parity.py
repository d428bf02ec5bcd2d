"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.
Bar: decoded bytes, counters and CRC results bit-exact; samples within 1e-4 (BASELINE.json)."""
import numpy as np
import pytest

import siggen

pytestmark = pytest.mark.gpu

TOL = 1e-4  # north-star tolerance for modulated / filtered samples (absolute and relative)


def oracle_run(O, cfg, x, chunk=None):
    m = O.FSKCore()
    m.configure(cfg)
    x = x.copy()
    if chunk is None:
        out = m.demodulateData(x)
    else:
        out = b"".join(m.demodulateData(x[i:i + chunk]) for i in range(0, len(x), chunk))
    return out, m.getStatus(), x


STATUS_KEYS = ["frameStarted", "globalSampleCounter", "receivedBitsLength", "syncDetections", "eodEvents",
               "demodulationCalls", "totalSamplesProcessed"]


def assert_status_equal(gs, os_, ctx=""):
    for k in STATUS_KEYS:
        assert float(gs[k]) == float(os_[k]), f"{ctx} status {k}: gpu {gs[k]} oracle {os_[k]}"
    assert gs["silenceThreshold"] == pytest.approx(os_["silenceThreshold"], rel=1e-9, abs=1e-15), ctx


@pytest.mark.parametrize("cfg,payload", [
    ({}, b"AB"),
    ({}, b"Hello, World!"),
    (dict(baudRate=300), b"Hello"),
    (siggen.BELL103, b"Hello, World!"),
    (dict(baudRate=300, markFrequency=1270, spaceFrequency=1070), b"Hello, World!"),  # R9: decodes nothing
    (siggen.V21_CH1, b"Hello, World!"),
    (siggen.V21_CH2, b"Hello, World!"),
    (dict(markFrequency=2125, spaceFrequency=2295), b"\x55\x7e\x48\x55\x7e"),
    (dict(parity="even"), b"parity"),
    (dict(parity="odd", stopBits=1), b"odd"),
    (dict(sampleRate=44100), b"AB"),                      # fractional sync-ring capacity (R10)
    (dict(sampleRate=44100, parity="even"), b"ABCD"),
    (dict(agcEnabled=False), b"no agc"),
    ({}, b""),
])
def test_single_stream_roundtrip(gpu_wam, oracle, cfg, payload):
    sig = siggen.modulate(cfg, payload)
    want, ost, omut = oracle_run(oracle, cfg, sig)
    m = gpu_wam.FSKCore()
    m.configure(cfg)
    x = sig.copy()
    got = bytes(m.demodulateData(x))
    assert got == want
    assert_status_equal(m.getStatus(), ost)
    # AGC mutates the caller's buffer in place (fsk.ts:55): same float32 values
    np.testing.assert_allclose(x, omut, rtol=TOL, atol=TOL)
    assert np.mean(x != omut) < 1e-3


def test_config1_polarity(gpu_wam, oracle):
    """BASELINE config 1: 'Hello, World!' 48 kHz / 300 Bd.  Bell 103 as named (1270/1070) decodes
    nothing in the reference; with the lower tone as mark it decodes exactly."""
    for mark, space, want in ((1270, 1070, b""), (1070, 1270, b"Hello, World!")):
        cfg = dict(baudRate=300, markFrequency=mark, spaceFrequency=space)
        sig = siggen.modulate(cfg, b"Hello, World!")
        assert len(sig) == 27520
        m = gpu_wam.FSKCore()
        m.configure(cfg)
        assert bytes(m.demodulateData(sig.copy())) == want
        assert oracle_run(oracle, cfg, sig)[0] == want


@pytest.mark.parametrize("chunk", [32, 64, 128, 256, 1000])
def test_chunked_equals_whole(gpu_wam, oracle, chunk):
    sig = siggen.modulate({}, b"AB")
    m = gpu_wam.FSKCore()
    m.configure({})
    got = b"".join(bytes(m.demodulateData(sig[i:i + chunk].copy())) for i in range(0, len(sig), chunk))
    want, ost, _ = oracle_run(oracle, {}, sig, chunk)
    assert got == want == b"AB"
    assert_status_equal(m.getStatus(), ost)


def test_three_messages_with_gaps(gpu_wam, oracle):
    cfg = {}
    parts = []
    for p in (b"AB", b"Hel", b"lo"):
        parts += [siggen.modulate(cfg, p), np.zeros(500, dtype=np.float32)]
    sig = np.concatenate(parts)
    m = gpu_wam.FSKCore()
    m.configure(cfg)
    events = []
    m.on("eod", lambda e: events.append("eod"))
    got = b"".join(bytes(m.demodulateData(sig[i:i + 128].copy())) for i in range(0, len(sig), 128))
    want, ost, _ = oracle_run(oracle, cfg, sig, 128)
    assert got == want == b"ABHello"
    assert len(events) == ost["eodEvents"]
    assert_status_equal(m.getStatus(), ost)


def test_reset_and_reconfigure(gpu_wam, oracle):
    sig = siggen.modulate({}, b"xyz")
    g = gpu_wam.FSKCore()
    o = oracle.FSKCore()
    for core in (g, o):
        core.configure({})
    for step in range(3):
        a = bytes(g.demodulateData(sig.copy()))
        b = o.demodulateData(sig.copy())
        assert a == b
        assert_status_equal(g.getStatus(), o.getStatus(), f"step {step}")
        if step == 0:
            g.reset(); o.reset()
        else:
            g.configure(dict(syncThreshold=0.7)); o.configure(dict(syncThreshold=0.7))
    assert not gpu_wam.FSKCore().isReady()
    with pytest.raises(RuntimeError, match="not configured"):
        gpu_wam.FSKCore().demodulateData(np.zeros(4, dtype=np.float32))
    with pytest.raises(RuntimeError, match="not configured"):
        gpu_wam.FSKCore().modulateData(b"x")


@pytest.mark.parametrize("cfg", [siggen.V21_CH1, siggen.V21_CH2, {}])
def test_batch_awgn_sweep_matches_oracle(gpu_wam, oracle, cfg):
    """Miniature of BASELINE config 2: AWGN swept -15..+30 dB in 3 dB steps; bytes + counters exact."""
    n_streams, n = 256, 48000 if cfg.get("baudRate") == 300 else 16000
    snr = np.repeat(np.arange(-15, 31, 3), n_streams // 16)
    x, _ = siggen.noisy_streams(cfg, n_streams, n, 25 if cfg.get("baudRate") == 300 else 30, snr, seed=0xB200)
    want, ost = oracle.batch_demodulate([cfg], None, x.copy(), n_threads=8)
    b = gpu_wam.FSKBatch(n_streams, cfg)
    got = b.demodulate_bytes(x)
    gst = b.status()
    bad = [i for i in range(n_streams) if got[i] != want[i]]
    assert not bad, f"{len(bad)} streams differ, first {bad[:5]}"
    for i in range(n_streams):
        assert_status_equal(gst[i], ost[i], f"stream {i}")
    assert sum(len(w) for w in want) > 0


@pytest.mark.parametrize("preamble,force_fused", [([0x55] * 4, True), ([0x55] * 4, False), ([0x55, 0x55, 0xAA], True)])
def test_long_sync_template_falls_back_to_offset_tables(gpu_wam, oracle, preamble, force_fused):
    """The frame search compares against ONE template passed by value (up to 80 words = 2560 compared samples).
    A 4-byte preamble at 300 Bd needs (50 - 1) * 80 = 3920 samples: both kernels must take the per-offset tables in
    global memory instead and still agree with the oracle; a 3-byte preamble (3120 samples = 98 words) likewise."""
    cfg = dict(baudRate=300, preamblePattern=preamble)
    n_streams, n = 64, 56000
    snr = np.repeat(np.array([-6.0, 0.0, 6.0, 30.0]), n_streams // 4)
    x, _ = siggen.noisy_streams(cfg, n_streams, n, 20, snr, seed=77)
    want, ost = oracle.batch_demodulate([cfg], None, x.copy(), n_threads=8)
    b = gpu_wam.FSKBatch(n_streams, cfg)
    got = b.demodulate_bytes(x, flags=gpu_wam._lib.WAM_BATCH_NO_PIPELINE if force_fused else 0)
    gst = b.status()
    bad = [i for i in range(n_streams) if got[i] != want[i]]
    assert not bad, f"{len(bad)} streams differ, first {bad[:5]}"
    for i in range(n_streams):
        assert_status_equal(gst[i], ost[i], f"stream {i}")
    assert sum(st["syncDetections"] for st in ost) > 0 and sum(len(w) for w in want) > 0


def test_batch_two_configs_and_streaming_state(gpu_wam, oracle):
    """V.21 ch1 + ch2 in one batch (config 2 layout), fed in 3 unequal slabs: the device-resident
    state must carry across calls exactly like one FSKCore per stream."""
    n_streams, n = 64, 48000
    x1, _ = siggen.noisy_streams(siggen.V21_CH1, n_streams // 2, n, 25, 9.0, seed=1)
    x2, _ = siggen.noisy_streams(siggen.V21_CH2, n_streams // 2, n, 25, 3.0, seed=2)
    x = np.ascontiguousarray(np.concatenate([x1, x2]))
    idx = np.array([0] * (n_streams // 2) + [1] * (n_streams // 2), dtype=np.int32)
    want, ost = oracle.batch_demodulate([siggen.V21_CH1, siggen.V21_CH2], idx, x.copy(), n_threads=8)
    b = gpu_wam.FSKBatch(n_streams, [siggen.V21_CH1, siggen.V21_CH2], idx)
    got = [b""] * n_streams
    for lo, hi in ((0, 10001), (10001, 30003), (30003, n)):
        part = b.demodulate_bytes(np.ascontiguousarray(x[:, lo:hi]))
        got = [g + p for g, p in zip(got, part)]
    assert got == want
    gst = b.status()
    for i in range(n_streams):
        for k in ["frameStarted", "globalSampleCounter", "syncDetections", "eodEvents"]:
            assert float(gst[i][k]) == float(ost[i][k]), (i, k)


def test_multi_frame_long_stream(gpu_wam, oracle):
    """Config 3 style: 1200 Bd, back-to-back 128-byte frames with gaps, +6 dB AWGN."""
    n_streams, n = 32, 48000 * 4
    xs = [siggen.multi_frame_stream({}, n, 128, 6.0, seed=100 + s)[0] for s in range(n_streams)]
    x = np.ascontiguousarray(np.stack(xs))
    want, ost = oracle.batch_demodulate([{}], None, x.copy(), n_threads=8)
    b = gpu_wam.FSKBatch(n_streams, {})
    assert b.demodulate_bytes(x) == want
    gst = b.status()
    for i in range(n_streams):
        for k in ["frameStarted", "globalSampleCounter", "syncDetections", "eodEvents"]:
            assert float(gst[i][k]) == float(ost[i][k]), (i, k)


def test_44k_fractional_ring_quirk(gpu_wam, oracle):
    """Config 4 style at 44.1 kHz: default framing (capacity 1227.6: degenerate ring, first frame only)
    and parity:'even' (capacity 1287.0), with a frequency offset.  GPU == oracle on both."""
    for cfg in (dict(sampleRate=44100), dict(sampleRate=44100, parity="even")):
        xs = [siggen.multi_frame_stream(cfg, 44100 * 2, 16, 20.0, seed=7 + s, freq_offset_hz=(-20 + 10 * s))[0]
              for s in range(5)]
        x = np.ascontiguousarray(np.stack(xs))
        want, ost = oracle.batch_demodulate([cfg], None, x.copy(), n_threads=5)
        b = gpu_wam.FSKBatch(5, cfg)
        got = b.demodulate_bytes(x)
        assert got == want, cfg
        gst = b.status()
        for i in range(5):
            for k in ["frameStarted", "globalSampleCounter", "receivedBitsLength", "syncDetections", "eodEvents"]:
                assert float(gst[i][k]) == float(ost[i][k]), (cfg, i, k)


def test_modulate_matches_oracle(gpu_wam, oracle):
    for cfg, payload in (({}, bytes(range(64))), (siggen.V21_CH1, b"Hello, World!"), (dict(parity="odd", stopBits=2), b"\x00\xff\x55"),
                         (dict(sampleRate=44100, markFrequency=1633.5, spaceFrequency=1870.25), b"frac"), ({}, b"")):
        want = siggen.modulate(cfg, payload)
        m = gpu_wam.FSKCore()
        m.configure(cfg)
        got = m.modulateData(payload)
        assert got.shape == want.shape
        np.testing.assert_allclose(got, want, rtol=0, atol=TOL)
        assert np.max(np.abs(got - want), initial=0) < 5e-6  # float32 sinpif on an exact phase


def test_batch_modulate_then_demodulate(gpu_wam, oracle):
    rng = np.random.default_rng(5)
    data = rng.integers(0, 256, (40, 20), dtype=np.uint8)
    b = gpu_wam.FSKBatch(40, {})
    sig, out_len = b.modulate(data)
    for s in range(40):
        want = siggen.modulate({}, data[s].tobytes())
        assert out_len[s] == len(want)
        np.testing.assert_allclose(sig[s], want, rtol=0, atol=TOL)
    got = b.demodulate_bytes(sig)
    assert got == [data[s].tobytes() for s in range(40)]


def test_event_driven_state_machine_equals_generic(gpu_wam, oracle):
    """The event-driven tile state machine and the per-sample one (debug flag) must agree exactly,
    including counters, on noisy multi-frame streams with many resets."""
    L = gpu_wam._lib
    n_streams, n = 64, 48000 * 2
    xs = [siggen.multi_frame_stream({}, n, 24, float(s % 8) * 2 - 4, seed=900 + s, max_gap=3000)[0] for s in range(n_streams)]
    x = np.ascontiguousarray(np.stack(xs))
    a = gpu_wam.FSKBatch(n_streams, {})
    b = gpu_wam.FSKBatch(n_streams, {})
    got_a, got_b = [b""] * n_streams, [b""] * n_streams
    for lo, hi in ((0, 30001), (30001, 30002), (30002, n)):  # odd slab lengths exercise the decimator parity
        pa = a.demodulate_bytes(np.ascontiguousarray(x[:, lo:hi]))
        pb = b.demodulate_bytes(np.ascontiguousarray(x[:, lo:hi]), flags=L.WAM_BATCH_DEBUG_GENERIC_SM)
        got_a = [u + v for u, v in zip(got_a, pa)]
        got_b = [u + v for u, v in zip(got_b, pb)]
    assert got_a == got_b
    sa, sb = a.status(), b.status()
    for i in range(n_streams):
        assert sa[i] == sb[i], i
    want, ost = oracle.batch_demodulate([{}], None, x.copy(), n_threads=8)
    assert got_a == want
    for i in range(n_streams):
        for k in ["frameStarted", "globalSampleCounter", "receivedBitsLength", "syncDetections", "eodEvents"]:
            assert float(sa[i][k]) == float(ost[i][k]), (i, k)


@pytest.mark.parametrize("cfg,n,payload_len,max_gap", [({}, 48000 * 2, 24, 3000), (siggen.V21_CH1, 48000 * 3, 6, 12000),
                                                    (dict(sampleRate=44100, parity="even"), 44100 * 2, 16, 2000)])
def test_pipelined_kernel_equals_fused_and_oracle(gpu_wam, oracle, cfg, n, payload_len, max_gap):
    """Few streams run on the warp-specialised pipeline (fsk_demod_pipe.cuh: A1 / A2 / B in three warps with
    roll-back on resetState()), many streams on the fused kernel.  Same input through both (WAM_BATCH_NO_PIPELINE
    forces the fused one) and through the oracle: bytes, counters and the complete carried state must agree.  The
    streams mix noisy multi-frame traffic (bad-start-bit resets), exact silence (EOD resets every 0.7 byte times)
    and odd slab lengths (decimator parity across calls)."""
    L = gpu_wam._lib
    n_streams = 48
    xs = []
    for s in range(n_streams):
        x = siggen.multi_frame_stream(cfg, n, payload_len, float(s % 8) * 3 - 6, seed=4000 + s, max_gap=max_gap)[0]
        if s % 3 == 0:  # exact silence in the middle: the AGC holds, amplitudes fall below the threshold -> EODs
            x[n // 3: n // 3 + n // 4] = 0.0
        if s % 5 == 0:  # noise-free stream: clean frames and exact-zero gaps
            x = siggen.multi_frame_stream(cfg, n, payload_len, 200.0, seed=5000 + s, max_gap=max_gap)[0]
        xs.append(x)
    x = np.ascontiguousarray(np.stack(xs))
    a = gpu_wam.FSKBatch(n_streams, cfg)
    b = gpu_wam.FSKBatch(n_streams, cfg)
    got_a, got_b = [b""] * n_streams, [b""] * n_streams
    cuts = (0, 20001, 20002, 20002 + 4096, n)
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        pa = a.demodulate_bytes(np.ascontiguousarray(x[:, lo:hi]))
        pb = b.demodulate_bytes(np.ascontiguousarray(x[:, lo:hi]), flags=L.WAM_BATCH_NO_PIPELINE)
        got_a = [u + v for u, v in zip(got_a, pa)]
        got_b = [u + v for u, v in zip(got_b, pb)]
    assert got_a == got_b
    sa, sb = a.status(), b.status()
    for i in range(n_streams):
        assert sa[i] == sb[i], i
        assert sa[i]["errorEvents"] == 0
    want, ost = oracle.batch_demodulate([cfg], None, x.copy(), n_threads=8)
    assert got_a == want
    for i in range(n_streams):
        for k in ["frameStarted", "globalSampleCounter", "receivedBitsLength", "syncDetections", "eodEvents"]:
            assert float(sa[i][k]) == float(ost[i][k]), (i, k)
    assert sum(s["eodEvents"] for s in sa) > 100 and sum(s["syncDetections"] for s in sa) > 100


@pytest.mark.parametrize("force_fused", [False, True])
def test_ragged_batches_match_one_fskcore_per_stream(gpu_wam, oracle, force_fused):
    """wam_fsk_batch_demodulate_ragged: every stream receives its own number of samples per round (0, odd counts,
    the maximum) or is not called at all (n_valid < 0) — the shape of a server multiplexing independent sessions.
    Each stream must behave like its own FSKCore fed the same sequence of demodulateData() calls: bytes, state and
    the per-stream debug counters.  Run through the few-stream pipeline and through the fused kernel."""
    L = gpu_wam._lib
    flags = L.WAM_BATCH_NO_PIPELINE if force_fused else 0
    cfg = {}
    n_streams, total = 40, 48000 * 2
    rng = np.random.default_rng(99)
    xs = np.stack([siggen.multi_frame_stream(cfg, total, 20, float(s % 5) * 4 - 2, seed=700 + s, max_gap=2500)[0]
                   for s in range(n_streams)])
    cores = []
    for s in range(n_streams):
        c = oracle.FSKCore()
        c.configure(cfg)
        cores.append(c)
    b = gpu_wam.FSKBatch(n_streams, cfg)
    pos = np.zeros(n_streams, dtype=np.int64)
    want = [b""] * n_streams
    got = [b""] * n_streams
    calls = np.zeros(n_streams)
    for rnd in range(14):
        n_max = int(rng.choice([128, 1000, 4097, 9000]))
        nv = np.zeros(n_streams, dtype=np.int32)
        buf = np.zeros((n_streams, n_max), dtype=np.float32)
        for s in range(n_streams):
            kind = rng.integers(0, 6)
            n = {0: -1, 1: 0, 2: n_max}.get(int(kind), int(rng.integers(1, n_max + 1)))
            n = min(n, total - int(pos[s]))
            nv[s] = n
            if n > 0:
                buf[s, :n] = xs[s, pos[s]:pos[s] + n]
            if n >= 0:
                want[s] += cores[s].demodulateData(xs[s, pos[s]:pos[s] + n].copy())
                calls[s] += 1
                pos[s] += n
        part = b.demodulate_ragged(buf, nv, flags=flags)
        for s in range(n_streams):
            if nv[s] < 0:
                assert part[s] == b""
            got[s] += part[s]
    assert got == want
    assert sum(len(w) for w in want) > 200
    st = b.status()
    for s in range(n_streams):
        o = cores[s].getStatus()
        assert st[s]["demodulationCalls"] == calls[s] == o["demodulationCalls"], s
        assert st[s]["totalSamplesProcessed"] == pos[s] == o["totalSamplesProcessed"], s
        for k in ["frameStarted", "globalSampleCounter", "receivedBitsLength", "syncDetections", "eodEvents"]:
            assert float(st[s][k]) == float(o[k]), (s, k)
        assert st[s]["errorEvents"] == 0
    # an empty round (n_max == 0) is a call with no samples on every stream that takes part
    part = b.demodulate_ragged(np.zeros((n_streams, 0), dtype=np.float32), np.zeros(n_streams, dtype=np.int32), flags=flags)
    assert part == [b""] * n_streams
    assert b.status()[0]["demodulationCalls"] == calls[0] + 1


def test_session_mux_matches_one_fskcore_per_session(gpu_wam, oracle):
    """wam_fsk_mux_*: sessions deliver 128-sample render quanta at their own pace (every tick, every other tick,
    two quanta per tick, pauses); one flush per tick.  Every session must decode exactly what its own FSKCore decodes
    when it is fed the same samples (one demodulateData() per flush over what the session pushed)."""
    cfgs = [{}, siggen.V21_CH2]
    n_sessions, total = 24, 48000
    idx = np.array([s % 2 for s in range(n_sessions)], dtype=np.int32)
    xs = [siggen.multi_frame_stream(cfgs[idx[s]], total, 8 if idx[s] == 0 else 3, 12.0, seed=300 + s, max_gap=1500)[0]
          for s in range(n_sessions)]
    cores = []
    for s in range(n_sessions):
        c = oracle.FSKCore()
        c.configure(cfgs[idx[s]])
        cores.append(c)
    mux = gpu_wam.FSKSessionMux(n_sessions, cfgs, idx, max_block=256)
    pos = [0] * n_sessions
    want, got = [b""] * n_sessions, [b""] * n_sessions
    tick = 0
    while min(pos) < total and tick < 2000:
        for s in range(n_sessions):
            quanta = {0: 1, 1: (tick + s) % 2, 2: 2, 3: 0 if (tick // 7) % 2 else 1}[s % 4]
            n_call = 0
            for _ in range(quanta):
                n = min(128, total - pos[s])
                if n <= 0:
                    break
                mux.push(s, xs[s][pos[s]:pos[s] + n])
                pos[s] += n
                n_call += n
            if n_call:
                want[s] += cores[s].demodulateData(xs[s][pos[s] - n_call:pos[s]].copy())
        assert mux.pending(2) in (0, 128, 256)
        part = mux.flush()
        got = [g + p for g, p in zip(got, part)]
        tick += 1
    assert got == want
    assert sum(len(w) for w in want) > 50
    st = mux.status()
    for s in range(n_sessions):
        o = cores[s].getStatus()
        for k in ["frameStarted", "globalSampleCounter", "syncDetections", "eodEvents", "demodulationCalls", "totalSamplesProcessed"]:
            assert float(st[s][k]) == float(o[k]), (s, k)
    with pytest.raises(gpu_wam.WamError):
        for _ in range(3):
            mux.push(0, np.zeros(128, dtype=np.float32))  # more than max_block between flushes
    mux.close()


def test_tma_staging_equals_cp_async_staging(gpu_wam, oracle):
    """Many-stream launches stage their input tiles with TMA (cp.async.bulk.tensor, 128-byte swizzle) when the rows of
    a configuration group are contiguous; WAM_BATCH_NO_TMA forces the per-lane cp.async path.  Same bytes and state,
    including ragged tails (n not a multiple of 32), a stream count that is not a multiple of 32 and several calls."""
    L = gpu_wam._lib
    n_streams, n = 32 * 23 + 5, 20000 + 13
    cfg = siggen.V21_CH2
    x, _ = siggen.noisy_streams(cfg, n_streams, n, 4, np.linspace(-6, 20, n_streams), seed=31)
    flags_fused = L.WAM_BATCH_NO_PIPELINE
    a = gpu_wam.FSKBatch(n_streams, cfg)
    b = gpu_wam.FSKBatch(n_streams, cfg)
    got_a, got_b = [b""] * n_streams, [b""] * n_streams
    for lo, hi in ((0, 7001), (7001, 7001 + 4096), (7001 + 4096, n)):
        part = np.ascontiguousarray(x[:, lo:hi])
        pa = a.demodulate_bytes(part, flags=flags_fused)
        pb = b.demodulate_bytes(part, flags=flags_fused | L.WAM_BATCH_NO_TMA)
        got_a = [u + v for u, v in zip(got_a, pa)]
        got_b = [u + v for u, v in zip(got_b, pb)]
    assert got_a == got_b
    assert a.status() == b.status()
    want, _ = oracle.batch_demodulate([cfg], None, x.copy(), n_threads=8)
    assert got_a == want and sum(len(w) for w in want) > 100


def test_dynamic_time_slabs_equal_one_pass_and_oracle(gpu_wam, oracle):
    """fsk_demod_slab_kernel: the call is cut into 64-tile work items claimed dynamically, a group's slabs run in order
    on whichever CTA is free, state carried through the per-stream arrays.  Forced here for a small batch (it is
    chosen by itself from ~38,000 streams): bytes, counters and carried state equal the one-pass kernel and the oracle,
    also across a second call."""
    L = gpu_wam._lib
    cfg = siggen.V21_CH2
    n_streams, n = 160, 40000  # 5 warp-groups x 20 slabs (last one partial: 1250 tiles)
    snr = np.repeat(np.array([-9.0, 0.0, 6.0, 12.0, 30.0]), n_streams // 5)
    x, _ = siggen.noisy_streams(cfg, n_streams, n + 9000, 20, snr, seed=4242)
    want, ost = oracle.batch_demodulate([cfg], None, x.copy(), n_threads=8)
    got = {}
    for name, flags in (("slabs", L.WAM_BATCH_NO_PIPELINE | L.WAM_BATCH_FORCE_SLABS), ("one pass", L.WAM_BATCH_NO_PIPELINE | L.WAM_BATCH_NO_SLABS)):
        b = gpu_wam.FSKBatch(n_streams, cfg)
        first = b.demodulate_bytes(np.ascontiguousarray(x[:, :n]), flags=flags)
        second = b.demodulate_bytes(np.ascontiguousarray(x[:, n:]), flags=flags)
        got[name] = [a + c for a, c in zip(first, second)]
        gst = b.status()
        for i in range(n_streams):
            assert_status_equal(gst[i], {**ost[i], "demodulationCalls": 2}, f"{name} stream {i}")
    assert got["slabs"] == got["one pass"]
    bad = [i for i in range(n_streams) if got["slabs"][i] != want[i]]
    assert not bad, f"{len(bad)} streams differ, first {bad[:5]}"
    assert sum(len(w) for w in want) > 0


@pytest.mark.parametrize("n", [16000, 16000 - 13])
def test_pcm16_entry_equals_widened_float_samples(gpu_wam, oracle, n):
    """wam_fsk_batch_demodulate_pcm16: int16 / 32768 widened on the device = demodulateData(Float32Array of pcm / 32768),
    bytes and counters exact against the oracle (row lengths with and without 16-byte rows)."""
    cfg = {}
    n_streams = 96
    snr = np.repeat(np.arange(-12, 36, 3), n_streams // 16)
    x, _ = siggen.noisy_streams(cfg, n_streams, 16000, 30, snr, seed=0x9C)
    x = x[:, :n]
    pcm = np.clip(np.round(x * (32768.0 / 8.0)), -32768, 32767).astype(np.int16)
    widened = (pcm.astype(np.float32) / np.float32(32768.0)).astype(np.float32)
    want, ost = oracle.batch_demodulate([cfg], None, widened.copy(), n_threads=8)
    b = gpu_wam.FSKBatch(n_streams, cfg)
    out, out_len = b.demodulate_pcm16(np.ascontiguousarray(pcm))
    got = [bytes(out[i, : out_len[i]]) for i in range(n_streams)]
    assert got == want
    gst = b.status()
    for i in range(n_streams):
        assert_status_equal(gst[i], ost[i], f"stream {i}")
    assert sum(len(w) for w in want) > 0
    # a second call continues the streams exactly like a second demodulateData()
    b2 = gpu_wam.FSKBatch(n_streams, cfg)
    h = (n // 2) // 32 * 32
    o1, l1 = b2.demodulate_pcm16(np.ascontiguousarray(pcm[:, :h]))
    o2, l2 = b2.demodulate_pcm16(np.ascontiguousarray(pcm[:, h:]))
    assert [bytes(o1[i, : l1[i]]) + bytes(o2[i, : l2[i]]) for i in range(n_streams)] == want


def test_host_bind_near_device_keeps_a_usable_mask(gpu_wam):
    import os
    before = os.sched_getaffinity(0)
    n = gpu_wam.lib().wam_host_bind_near_device(0)
    after = os.sched_getaffinity(0)
    try:
        assert n >= 0
        assert after <= before and len(after) > 0
        if n > 0:
            assert len(after) == n
    finally:
        os.sched_setaffinity(0, before)


@pytest.mark.parametrize("start_bits,stop_bits", [(5, 4), (16, 16), (3, 1)])
def test_wide_framing_long_frames(gpu_wam, oracle, start_bits, stop_bits):
    """startBits + stopBits > 2 (bits per byte 17 and 40): line bits of bytes beyond 240 of a frame (the modulator's
    bit-index division), and a demodulator that completes a byte every stop_pos + 1 decided bits whatever the framing
    says, so that the output capacity must not be sized by bits per byte."""
    cfg = dict(baudRate=1200, startBits=start_bits, stopBits=stop_bits)
    payload = bytes((i * 37 + 11) & 0xFF for i in range(300))
    want = siggen.modulate(cfg, payload)
    m = gpu_wam.FSKCore()
    m.configure(cfg)
    got = m.modulateData(payload)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=0, atol=TOL)
    # batch modulator (bit table kernel + sample kernel, other chunk size) on the same frame
    b = gpu_wam.FSKBatch(3, cfg)
    sig, out_len = b.modulate(np.frombuffer(payload, dtype=np.uint8)[None, :].repeat(3, axis=0))
    assert out_len.tolist() == [len(want)] * 3
    np.testing.assert_allclose(sig[1], want, rtol=0, atol=TOL)
    # demodulated bytes and counters equal to the reference algorithm's on the same samples
    x = np.concatenate([want, np.zeros(4000, dtype=np.float32)]).astype(np.float32)
    exp, ost = oracle.batch_demodulate([cfg], None, x[None, :].copy(), n_threads=1)
    rx = gpu_wam.FSKBatch(1, cfg)
    out = rx.demodulate_bytes(x[None, :].copy())
    assert out[0] == exp[0]
    st = rx.status()[0]
    assert_status_equal(st, ost[0], "wide framing")
    assert st["errorEvents"] == 0
