"""rte_compute_Planck_source_1D / _2D (rte/kernels/mo_gas_optics_utils.F90:36-95; api
rte/kernels/api/mo_gas_optics_utils.F90:7-36).  Three pins: (1) a numpy statement of B_nu :36-41 evaluated with the same
association, (2) the analytic limit - the spectral integral of B_nu over a fine wavenumber grid is sigma T^4 / pi,
(3) CUDA vs oracle at 1e-13 relative (libdevice exp vs glibc exp, <= 1 ulp each, amplified by 1/(exp(x) - 1))."""
import numpy as np

import refcases as rc
from rte_rrtmgp_b200.abi import fzeros

H, C, KB = 6.626075540e-34, 2.99792458e8, 1.380649e-23   # mo_gas_optics_constants.F90:11,22-23


def _b_nu(T, nu):
    nu100 = nu * 100.0
    return 100.0 * 2.0 * H * (nu100 * nu100 * nu100) * (C * C) / (np.exp((H * C * nu * 100.0) / (KB * T)) - 1.0)


def _inputs():
    rng = np.random.default_rng(12)
    ncol, nlay, nnu = 37, 19, 23
    nus = np.sort(rng.uniform(10.0, 3200.0, nnu))
    dnus = rng.uniform(1.0, 40.0, nnu)
    T2 = np.asfortranarray(rng.uniform(160.0, 340.0, (ncol, nlay)))
    return ncol, nlay, nnu, nus, dnus, T2


def _run(lib, device, ncol, nlay, nnu, nus, dnus, T2):
    d = lambda a: rc.dev(a, device)
    s2 = fzeros((ncol, nlay, nnu), device=device)
    lib.rte_compute_Planck_source_2D(ncol, nlay, nnu, d(nus), d(dnus), d(T2), s2)
    s1 = fzeros((ncol, nnu), device=device)
    lib.rte_compute_Planck_source_1D(ncol, nnu, d(nus), d(dnus), d(np.ascontiguousarray(T2[:, 0])), s1)
    lib.sync()
    return rc.host(s2), rc.host(s1)


def test_planck_source_1d_2d(backend):
    lib, device = backend
    ncol, nlay, nnu, nus, dnus, T2 = _inputs()
    s2, s1 = _run(lib, device, ncol, nlay, nnu, nus, dnus, T2)
    ref2 = _b_nu(T2[:, :, None], nus[None, None, :]) * dnus[None, None, :]
    np.testing.assert_allclose(s2, ref2, rtol=1e-13)
    np.testing.assert_allclose(s1, ref2[:, 0, :], rtol=1e-13)
    if device is not None:   # CUDA vs the oracle
        import oracle

        o2, o1 = _run(oracle.lib(), None, ncol, nlay, nnu, nus, dnus, T2)
        np.testing.assert_allclose(s2, o2, rtol=1e-13)
        np.testing.assert_allclose(s1, o1, rtol=1e-13)


def test_planck_source_integrates_to_stefan_boltzmann(backend):
    lib, device = backend
    nnu = 6000
    edges = np.linspace(0.01, 12000.0, nnu + 1)   # cm-1; > 99.999% of the energy at 200-320 K
    nus, dnus = 0.5 * (edges[1:] + edges[:-1]), np.diff(edges)
    T = np.array([200.0, 255.0, 288.0, 320.0])
    d = lambda a: rc.dev(a, device)
    s1 = fzeros((T.size, nnu), device=device)
    lib.rte_compute_Planck_source_1D(T.size, nnu, d(nus), d(dnus), d(T), s1)
    lib.sync()
    sigma = 2.0 * np.pi**5 * KB**4 / (15.0 * H**3 * C**2)
    np.testing.assert_allclose(rc.host(s1).sum(axis=1), sigma * T**4 / np.pi, rtol=2e-5)
