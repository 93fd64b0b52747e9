"""Restatement of the reference's tests/rte_sw_solver_unit_tests.F90 at kernel level:
thin scattering atmospheres (tau in {1e-4,1e-2}, ssa = 1-tau, g in {.85,.65}), delta-scaled,
mu0 in {1, .5}; direct beam vs Beer-Lambert-Bouguer within 20 spacings (:121-130), column
subsetting, vertical flip, linearity in the TOA flux (:139-211)."""
import numpy as np
import pytest

import refcases as rc

NCOL, NLAY = 8, 16
G = np.array([0.85, 0.65])
TAU = np.array([1.0e-4, 1.0e-2])
SSA = 1.0 - np.array([1.0e-4, 1.0e-2])
ALB = np.zeros((NCOL, 1), order="F")
TOA = np.ones((NCOL, 1), order="F")


def _solve(lib, device, prob, mu0, toa=TOA, top_at_1=True):
    return rc.sw_2stream_broadband(lib, device, prob, np.full(NCOL, mu0), toa, ALB, ALB, top_at_1)


@pytest.mark.parametrize("mu0", [1.0, 0.5])
def test_thin_scattering(backend, mu0):
    lib, device = backend
    tol = 2.0 if device is None else 8.0  # GPU: libdevice exp/FMA vs glibc; reference notes GPU needs up to 20
    prob = rc.thin_scattering(lib, device, TAU, SSA, G, NLAY)
    up, dn, dr = _solve(lib, device, prob, mu0)
    # direct beam at the surface (:121-130), tol 20 spacings as in the reference
    beer = TOA[:, 0] * mu0 * np.exp(-np.sum(prob["tau"][:, :, 0], axis=1) / mu0)
    assert rc.allclose(dr[:, NLAY], beer, tol=20.0), rc.max_spacings(dr[:, NLAY], beer)
    assert np.all(dn >= dr) and np.all(up >= 0)
    # subsetting (:172-176)
    for s, e in ((0, 3), (3, 8)):
        sub = {k: np.asfortranarray(v[s:e]) for k, v in prob.items()}
        u, d, _ = rc.sw_2stream_broadband(lib, device, sub, np.full(e - s, mu0), np.asfortranarray(TOA[s:e]),
                                          np.asfortranarray(ALB[s:e]), np.asfortranarray(ALB[s:e]), True)
        assert rc.allclose(u, up[s:e], tol=tol) and rc.allclose(d, dn[s:e], tol=tol)
    # vertical flip (:181-199)
    u, d, _ = _solve(lib, device, rc.vr(prob), mu0, top_at_1=False)
    assert rc.allclose(u[:, ::-1], up, tol=tol) and rc.allclose(d[:, ::-1], dn, tol=tol)
    # linear in TOA flux (:204-211)
    u, d, _ = _solve(lib, device, prob, mu0, toa=np.asfortranarray(TOA * 2.0))
    assert rc.allclose(u / 2.0, up, tol=tol) and rc.allclose(d / 2.0, dn, tol=tol)


def test_night_columns_give_zero_flux(backend):
    """mu0 <= 0: the reference masks sources (:1120-1125); the direct beam underflows to 0 below TOA."""
    lib, device = backend
    prob = rc.thin_scattering(lib, device, TAU, SSA, G, NLAY)
    mu0 = np.array([1.0, -0.2, 0.5, 0.0, 1.0, -1.0, 0.3, 0.9])
    up, dn, dr = rc.sw_2stream_broadband(lib, device, prob, mu0, TOA, ALB, ALB, True)
    night = mu0 <= 0
    assert np.all(up[night] == 0.0)
    assert np.all(dr[night, 1:] == 0.0) and np.all(dn[night, 1:] == 0.0)
    assert np.all(up[~night][:, 0] > 0.0)  # reflected sunlight leaves the top; the surface is black
