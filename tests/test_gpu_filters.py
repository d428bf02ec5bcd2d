"""filters.ts on the GPU: the reference's expectations through the product API, and direct
GPU-vs-oracle comparisons (time-chunked IIR scan over long streams, smem FIR, state carry)."""
import numpy as np
import pytest

import filter_cases

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.mark.parametrize("case", filter_cases.ALL, ids=lambda f: f.__name__)
def test_reference_filter_expectations(gpu_wam, case):
    case(gpu_wam)


def test_designs_equal_oracle(gpu_wam, oracle):
    G, O = gpu_wam.FilterDesign, oracle.FilterDesign
    for args in ((300, 48000), (1200, 48000), (1000, 44100)):
        for n in ("butterworthLowpass", "butterworthHighpass"):
            a, b = getattr(G, n)(*args), getattr(O, n)(*args)
            np.testing.assert_array_equal(a["b"], b["b"]); np.testing.assert_array_equal(a["a"], b["a"])
    a, b = G.butterworthBandpass(1750, 2600, 48000), O.butterworthBandpass(1750, 2600, 48000)
    np.testing.assert_array_equal(a["b"], b["b"]); np.testing.assert_array_equal(a["a"], b["a"])
    for taps in (51, 50, 7):
        np.testing.assert_array_equal(G.sincLowpass(1000, 44100, taps), O.sincLowpass(1000, 44100, taps))
        np.testing.assert_array_equal(G.sincHighpass(1000, 44100, taps), O.sincHighpass(1000, 44100, taps))
        np.testing.assert_array_equal(G.sincBandpass(1500, 400, 44100, taps), O.sincBandpass(1500, 400, 44100, taps))


@pytest.mark.parametrize("n", [1, 127, 128, 129, 4095, 4096, 4097, 100000, 600000])
def test_iir_time_chunked_scan_matches_oracle(gpu_wam, oracle, n):
    """Long streams: chunked linear-recurrence scan (128-sample chunks, 32-chunk warp spans)."""
    filt = importlib_filters(gpu_wam)
    rng = np.random.default_rng(n)
    x = rng.standard_normal((3, n)).astype(np.float32)
    for coeffs in (oracle.FilterDesign.butterworthLowpass(300, 48000),      # poles at radius 0.97
                   oracle.FilterDesign.butterworthBandpass(1750, 2600, 48000),
                   {"b": np.array([0.2, 0.1, 0.05, 0.3]), "a": np.array([2.0, -0.6, 0.4, -0.1, 0.05])}):
        y = filt.iir_process_batch(coeffs["b"], coeffs["a"], x)
        for s in range(3):
            want = oracle.IIRFilter(coeffs["b"], coeffs["a"]).processBuffer(x[s])
            np.testing.assert_allclose(y[s], want, rtol=TOL, atol=TOL)
            assert np.max(np.abs(y[s] - want)) < 1e-6 * max(1.0, np.max(np.abs(want)))


def test_iir_state_carry_across_calls(gpu_wam, oracle):
    filt = importlib_filters(gpu_wam)
    c = oracle.FilterDesign.butterworthLowpass(1200, 48000)
    rng = np.random.default_rng(3)
    x = rng.standard_normal((4, 20000)).astype(np.float32)
    state = np.zeros((4, 4), dtype=np.float64)
    parts = [filt.iir_process_batch(c["b"], c["a"], np.ascontiguousarray(x[:, lo:hi]), state)
             for lo, hi in ((0, 1), (1, 5000), (5000, 5003), (5003, 20000))]
    y = np.concatenate(parts, axis=1)
    for s in range(4):
        np.testing.assert_allclose(y[s], oracle.IIRFilter(c["b"], c["a"]).processBuffer(x[s]), rtol=TOL, atol=1e-6)


@pytest.mark.parametrize("ntaps", [1, 5, 51, 255, 1024])
def test_fir_matches_oracle(gpu_wam, oracle, ntaps):
    filt = importlib_filters(gpu_wam)
    rng = np.random.default_rng(ntaps)
    taps = rng.standard_normal(ntaps) / ntaps
    x = rng.standard_normal((5, 3001)).astype(np.float32)
    y = filt.fir_process_batch(taps, x)
    for s in range(5):
        want = oracle.FIRFilter(taps).processBuffer(x[s])
        np.testing.assert_allclose(y[s], want, rtol=TOL, atol=1e-6)


def importlib_filters(wam):
    import importlib

    return importlib.import_module("webaudio-modem_b200.filters")


def test_more_streams_than_a_grid_dimension(gpu_wam, oracle):
    """70,000 streams in one call (round 1 stopped at 65,535: streams were a grid's y dimension), IIR with carried state
    and FIR, spot-checked against the oracle."""
    filt = importlib_filters(gpu_wam)
    n_streams, n = 70000, 300
    rng = np.random.default_rng(70)
    x = rng.standard_normal((n_streams, n)).astype(np.float32)
    c = oracle.FilterDesign.butterworthBandpass(1750, 800, 48000)
    state = np.zeros((n_streams, 4), dtype=np.float64)
    y1 = filt.iir_process_batch(c["b"], c["a"], np.ascontiguousarray(x[:, :100]), state)
    y2 = filt.iir_process_batch(c["b"], c["a"], np.ascontiguousarray(x[:, 100:]), state)
    y = np.concatenate([y1, y2], axis=1)
    taps = oracle.FilterDesign.sincLowpass(1000, 48000, 51)
    z = filt.fir_process_batch(taps, x)
    for s in (0, 1, 32767, 65534, 65535, 65536, 69999):
        np.testing.assert_allclose(y[s], oracle.IIRFilter(c["b"], c["a"]).processBuffer(x[s]), rtol=TOL, atol=1e-6)
        np.testing.assert_allclose(z[s], oracle.FIRFilter(taps).processBuffer(x[s]), rtol=TOL, atol=1e-6)
