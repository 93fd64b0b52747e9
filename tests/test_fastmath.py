"""The solver kernels' lean fp64 exp / sqrt / reciprocal / division (csrc/kernels/fastmath.cuh) against libm / IEEE
arithmetic (the oracle's side of the same probe).  Tolerance: 2 ulp, written here; the solvers' own parity bar is the
reference's 1e-5 W/m2 on fluxes (tests/test_allsky_parity.py)."""
import ctypes

import numpy as np
import pytest


def _probe(lib, x):
    n = x.size
    outs = [np.empty(n) for _ in range(4)]
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.cdll.rrtmgpb_fastmath_probe(ctypes.c_int(n), P(x), *[P(o) for o in outs])
    if lib.backend != "cpu-oracle":
        lib.cdll.rrtmgpb_sync()
    return outs


def _inputs():
    rng = np.random.default_rng(11)
    x = np.concatenate([
        -np.abs(rng.standard_cauchy(20000)) * 3.0,        # the solvers' exp arguments: -tau*k, -tau/mu0
        -10.0 ** rng.uniform(-300, 2.8, 20000),           # tiny ... -630
        10.0 ** rng.uniform(-12, 2.8, 10000),             # positive (sqrt / rcp / div operands)
        np.array([-2.0e9, -3.0e9, -1.0e15, -1.0e200, -746.0, -709.0]),  # |x*log2(e)| beyond the int range: must still flush to 0
        np.array([-708.0, -707.9, -1e-320 - 1e-300, -0.0 - 1e-17, -1.0, -0.5, 1e-12, 2.220446049250313e-12, 1.0, 4.0]),
    ])
    return np.ascontiguousarray(x[(np.abs(x) > 1e-300) & (np.abs(x) < 1e300)])


def _ulps(a, b):
    return np.abs(a - b) / np.spacing(np.maximum(np.abs(b), np.finfo(float).tiny))


def test_oracle_probe_is_libm(oracle_lib):
    x = _inputs()
    e, s, r, d = _probe(oracle_lib, x)
    np.testing.assert_array_equal(s, np.sqrt(np.abs(x)))
    np.testing.assert_array_equal(r, 1.0 / x)
    assert np.max(_ulps(e, np.exp(x))) <= 1.0


@pytest.mark.gpu
def test_fastmath_within_2ulp(oracle_lib, cuda_lib):
    x = _inputs()
    ref = _probe(oracle_lib, x)
    got = _probe(cuda_lib, x)
    sel = x >= -708.0  # below: flushed to 0 by design (true values are subnormal)
    assert np.all(got[0][x < -708.001] == 0.0)  # however negative (the power of two must not wrap around)
    fin = np.abs(x) < 1e100  # (x*x + 1)/x of the probe overflows beyond; those arguments are there for exp only
    for name, g, r, m in (("exp", got[0], ref[0], sel), ("sqrt", got[1], ref[1], fin),
                          ("rcp", got[2], ref[2], fin), ("div", got[3], ref[3], fin)):
        u = _ulps(g[m], r[m])
        print(f"fastmath {name}: max {np.max(u):.3f} ulp, mean {np.mean(u):.4f} ulp over {u.size} arguments")
        assert np.max(u) <= 2.0, (name, float(np.max(u)), x[m][np.argmax(u)])
