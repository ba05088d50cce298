"""Device build of the emulated libm vs the host's glibc, bit for bit, on millions of arguments (the exhaustive sweep of
the same source runs on the host in tests/test_math_kat.py)."""
import numpy as np
import pytest
import torch

from urban_road_filter_b200 import api

pytestmark = pytest.mark.gpu


def _run(which, a, b=None):
    lib = api.load_library()
    a = np.ascontiguousarray(a, np.float32)
    out = np.empty_like(a)
    bb = np.ascontiguousarray(b, np.float32) if b is not None else None
    rc = lib.urf_test_math(0, which, a.ctypes.data, bb.ctypes.data if bb is not None else None, out.ctypes.data, a.size)
    assert rc == 0
    return out


def _same(x, y):
    return np.array_equal(x.view(np.uint32), y.view(np.uint32)) or np.all((x.view(np.uint32) == y.view(np.uint32)) | (np.isnan(x) & np.isnan(y)))


def test_device_libm_matches_glibc():
    assert torch.cuda.is_available()
    import ctypes as C
    libm = C.CDLL("libm.so.6")
    rng = np.random.default_rng(0)
    n = 1 << 22
    u = np.concatenate([rng.uniform(-1, 1, n), np.linspace(-1, 1, 100001), [1.0, -1.0, 0.0, 1.5, np.nan, 0.975, 0.5, -0.5]]).astype(np.float32)
    bits = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32).view(np.float32)

    def host(fn, *args):
        f = getattr(libm, fn)
        f.restype = C.c_float
        f.argtypes = [C.c_float] * len(args)
        return np.array([f(*[float(v) for v in t]) for t in zip(*args)], np.float32)

    # full arrays through numpy-vectorised ctypes would be slow; sample 200k for the host side
    sel = rng.choice(u.size, 200000, replace=False)
    assert _same(_run(0, u)[sel], host("asinf", u[sel]))
    assert _same(_run(1, u)[sel], host("acosf", u[sel]))
    y = rng.uniform(-100, 100, n).astype(np.float32)
    x = rng.uniform(-100, 100, n).astype(np.float32)
    y[::5] *= 1e-6
    sel = rng.choice(n, 200000, replace=False)
    assert _same(_run(2, y, x)[sel], host("atan2f", y[sel], x[sel]))
    sel = rng.choice(n, 100000, replace=False)
    assert _same(_run(3, bits)[sel], host("atanf", bits[sel]))
