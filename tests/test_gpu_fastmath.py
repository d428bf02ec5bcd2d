"""Error bounds of the device float64 primitives used by the discriminator (fastmath.cuh)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def run(wam, y, x):
    y = np.ascontiguousarray(y, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    o1, o2, o3 = np.zeros_like(x), np.zeros_like(x), np.zeros_like(x)
    dp = C.POINTER(C.c_double)
    rc = wam.lib().wam_debug_fastmath(0, y.ctypes.data_as(dp), x.ctypes.data_as(dp), len(x), o1.ctypes.data_as(dp),
                                      o2.ctypes.data_as(dp), o3.ctypes.data_as(dp))
    assert rc == 0, wam.lib().wam_last_error()
    return o1, o2, o3


def test_fast_atan2_sqrt_rcp_accuracy(gpu_wam):
    rng = np.random.default_rng(1)
    n = 1 << 20
    ang = rng.uniform(-np.pi, np.pi, n)
    mag = 10.0 ** rng.uniform(-12, 3, n)
    y, x = mag * np.sin(ang), mag * np.cos(ang)
    a, s, r = run(gpu_wam, y, x)
    ref = np.arctan2(y, x)
    err = np.abs(a - ref)
    ulp = np.spacing(np.maximum(np.abs(ref), 1e-3))
    assert np.max(err / ulp) <= 6.0, np.max(err / ulp)
    assert np.max(err) < 2e-15
    xs = np.abs(x)
    assert np.max(np.abs(s - np.sqrt(xs)) / np.spacing(np.sqrt(xs))) <= 2.0
    assert np.max(np.abs(r - 1.0 / x) / np.spacing(np.abs(1.0 / x))) <= 2.0


def test_fast_atan2_special_cases(gpu_wam):
    y = np.array([0.0, 0.0, -0.0, 0.0, -0.0, 1.0, -1.0, 1e-300, 1e300, 1.0, 3e-310, 0.5, -0.5, 1.0, 1e-200])
    x = np.array([0.0, 1.0, 1.0, -1.0, -1.0, 0.0, 0.0, 1e-300, 1e300, 1.0, 3e-310, -0.0, -0.0, 1e-200, 1.0])
    a, _, _ = run(gpu_wam, y, x)
    ref = np.arctan2(y, x)
    np.testing.assert_allclose(a, ref, rtol=0, atol=1e-15)
    assert np.all(np.signbit(a) == np.signbit(ref))
    # every table interval boundary and the diagonal
    k = np.arange(0, 65)
    t = np.concatenate([k / 64.0, (k + 0.5) / 64.0, (k + 0.4999) / 64.0])
    for sx, sy in ((1, 1), (-1, 1), (1, -1), (-1, -1)):
        a, _, _ = run(gpu_wam, sy * t, sx * np.ones_like(t))
        np.testing.assert_allclose(a, np.arctan2(sy * t, sx * np.ones_like(t)), rtol=0, atol=1e-15)
        a, _, _ = run(gpu_wam, sy * np.ones_like(t), sx * t)
        np.testing.assert_allclose(a, np.arctan2(sy * np.ones_like(t), sx * t), rtol=0, atol=1e-15)
