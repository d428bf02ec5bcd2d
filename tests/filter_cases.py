"""The reference's filter expectations (tests/dsp/filters.node.test.ts, filters-advanced.node.test.ts)
as implementation-agnostic checks: `F` is a module-like object exposing IIRFilter, FIRFilter,
FilterDesign, FilterFactory (the oracle package or the GPU-backed product package)."""
import numpy as np
import pytest

FS = 44100


def tone_mix(freqs, amps, fs, dur):
    t = np.arange(int(fs * dur)) / fs
    return sum(a * np.sin(2 * np.pi * f * t) for f, a in zip(freqs, amps)).astype(np.float32)


def band_power(x, fs, lo, hi):
    spec = np.abs(np.fft.rfft(x.astype(np.float64))) ** 2
    f = np.fft.rfftfreq(len(x), 1 / fs)
    return spec[(f >= lo) & (f <= hi)].sum()


def check_iir_basics(F):
    f = F.IIRFilter([1, 2, 1], [1, -0.5, 0.25])                      # filters :124-133
    c = f.getCoefficients()
    assert list(c["b"]) == [1, 2, 1] and list(c["a"]) == [1, -0.5, 0.25]
    c = F.IIRFilter([2, 4, 2], [2, -1, 0.5]).getCoefficients()       # :135-143 a0 normalisation
    assert c["a"][0] == 1 and c["b"][0] == 1 and c["a"][1] == -0.5 and c["b"][1] == 2
    for b, a, msg in (([], [1], "Feedforward coefficients"), ([1], [], "Feedback coefficients"),
                      ([1], [0, 1], "cannot be zero")):                # advanced :115-125
        with pytest.raises(ValueError, match=msg):
            F.IIRFilter(b, a)
    c["b"][0] = 99                                                    # :415-425 copies
    assert f.getCoefficients()["b"][0] == 1


def check_iir_behaviour(F):
    f = F.IIRFilter([0.1, 0.2, 0.1], [1, -0.5, 0.2])
    imp = np.zeros(100, dtype=np.float32); imp[0] = 1
    r = f.processBuffer(imp)                                          # :145-160 stability
    assert abs(r[-1]) < 1e-6 and np.all(np.isfinite(r))
    f.reset()                                                         # :162-176, advanced :402-421
    r2 = f.processBuffer(imp)
    np.testing.assert_array_equal(r, r2)
    assert len(f.processBuffer(np.zeros(0, dtype=np.float32))) == 0   # :407-413
    big = F.IIRFilter([0.1, 0.2, 0.1], [1, -0.5, 0.2]).processBuffer(np.full(64, 1e6, dtype=np.float32))
    assert np.all(np.isfinite(big))                                   # :391-405
    # chunked processing == whole (streaming state), sample-by-sample process() too
    x = tone_mix([300, 5000], [1, 0.5], FS, 0.05)
    whole = F.IIRFilter([0.1, 0.2, 0.1], [1, -0.5, 0.2]).processBuffer(x)
    g = F.IIRFilter([0.1, 0.2, 0.1], [1, -0.5, 0.2])
    parts = np.concatenate([g.processBuffer(x[i:i + 333]) for i in range(0, len(x), 333)])
    np.testing.assert_allclose(parts, whole, rtol=0, atol=1e-6)
    h = F.IIRFilter([0.1, 0.2, 0.1], [1, -0.5, 0.2])
    single = np.array([h.process(float(v)) for v in x[:40]])
    np.testing.assert_allclose(single, whole[:40], rtol=0, atol=1e-6)


def check_fir(F):
    taps = [0.1, 0.2, 0.4, 0.2, 0.1]
    f = F.FIRFilter(taps)
    imp = np.zeros(10, dtype=np.float32); imp[0] = 1
    r = f.processBuffer(imp)                                          # :190-206
    np.testing.assert_allclose(r[:5], taps, atol=1e-5)
    np.testing.assert_allclose(r[5:], 0, atol=1e-10)
    x1 = tone_mix([500], [1], FS, 0.01); x2 = tone_mix([1500], [0.7], FS, 0.01)   # :208-229 linearity
    y1 = F.FIRFilter(taps).processBuffer(x1); y2 = F.FIRFilter(taps).processBuffer(x2)
    y12 = F.FIRFilter(taps).processBuffer((x1 + x2).astype(np.float32))
    np.testing.assert_allclose(y12, y1 + y2, atol=1e-5)
    # group delay (N-1)/2 for a symmetric design — advanced :281-307
    c = F.FilterDesign.sincLowpass(2000, FS, 51)
    imp = np.zeros(200, dtype=np.float32); imp[0] = 1
    assert int(np.argmax(np.abs(F.FIRFilter(c).processBuffer(imp)))) == 25
    g = F.FIRFilter(taps)                                             # streaming state
    x = tone_mix([700, 3000], [1, 1], FS, 0.02)
    parts = np.concatenate([g.processBuffer(x[i:i + 97]) for i in range(0, len(x), 97)])
    np.testing.assert_allclose(parts, F.FIRFilter(taps).processBuffer(x), atol=1e-6)
    g.reset()
    np.testing.assert_array_equal(g.processBuffer(imp), F.FIRFilter(taps).processBuffer(imp))


def check_designs(F):
    d = F.FilterDesign.butterworthLowpass(1000, FS)                   # advanced :311-324
    assert d["a"][0] == 1 and len(d["b"]) == 3 and len(d["a"]) == 3
    assert abs(d["b"].sum() / d["a"].sum() - 1) < 1e-5
    for fc in (1, 20000, 22050):                                      # advanced :326-337
        F.FilterDesign.butterworthLowpass(fc, FS)
    c = F.FilterDesign.sincLowpass(1000, FS, 51)                      # :300-323, advanced :339-362
    assert len(c) == 51
    np.testing.assert_allclose(c[:25], c[::-1][:25], atol=1e-10)
    assert np.argmax(np.abs(c)) == 25 and abs(c[12]) > abs(c[0])
    assert len(F.FilterDesign.sincLowpass(1000, FS, 50)) == 51        # :344-347
    assert len(F.FilterDesign.sincHighpass(1000, FS, 51)) == 51       # :325-342
    assert len(F.FilterDesign.sincBandpass(1500, 400, FS, 51)) == 51


def check_band_selectivity(F):
    x = tone_mix([500, 5000], [1, 1], FS, 0.1)
    lp = F.FilterFactory.createIIRLowpass(1000, FS).processBuffer(x)  # :235-253, :353-372
    assert band_power(lp, FS, 0, 800) > 10 * band_power(lp, FS, 4000, 6000)
    hp = F.FilterFactory.createIIRHighpass(2000, FS).processBuffer(x)  # :255-272
    assert band_power(hp, FS, 4000, 6000) > 10 * band_power(hp, FS, 0, 800)
    y = tone_mix([500, 1500, 5000], [1, 1, 1], FS, 0.1)
    bp = F.FilterFactory.createIIRBandpass(1500, 400, FS).processBuffer(y)  # :274-295
    mid = band_power(bp, FS, 1300, 1700)
    assert mid > band_power(bp, FS, 0, 800) and mid > band_power(bp, FS, 4000, 6000)
    z = tone_mix([500, 2000], [1, 1], FS, 0.1)
    fl = F.FilterFactory.createFIRLowpass(1000, FS).processBuffer(z)  # :374-389, :300-323
    assert band_power(fl, FS, 0, 800) > band_power(fl, FS, 1500, 3000)
    fh = F.FilterFactory.createFIRHighpass(1000, FS).processBuffer(z)
    assert band_power(fh, FS, 1500, 3000) > band_power(fh, FS, 0, 800)
    fb = F.FilterFactory.createFIRBandpass(1500, 400, FS).processBuffer(y)
    assert band_power(fb, FS, 1300, 1700) > band_power(fb, FS, 4000, 6000)
    # -3 dB at the cutoff — advanced :196-217
    t = tone_mix([1000], [1], FS, 0.2)
    out = F.FilterFactory.createIIRLowpass(1000, FS).processBuffer(t)
    gain = np.sqrt(np.mean(out[2000:].astype(np.float64) ** 2) / np.mean(t[2000:].astype(np.float64) ** 2))
    assert abs(20 * np.log10(gain) + 3.0) < 0.5


ALL = [check_iir_basics, check_iir_behaviour, check_fir, check_designs, check_band_selectivity]
