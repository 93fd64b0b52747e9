"""By-band flux reductions (rte/extensions/mo_fluxes_byband.F90:159-218) and heating rates
(rte/extensions/mo_heating_rates.F90:34-117) - SURVEY 8f rank 3.  Oracle vs a numpy statement of the same loops
(bit-exact: same summation order), CUDA vs oracle (bit-exact for the reductions; 1e-14 relative for the heating
rates, whose expression nvcc may contract into FMAs)."""
import ctypes as C

import numpy as np
import pytest

from rte_rrtmgp_b200.frontend import Context

NCOL, NLEV, BANDS = 37, 23, [(1, 3), (4, 4), (5, 11), (12, 16)]
NGPT = 16


def _ptr(a):
    from rte_rrtmgp_b200.frontend import _addr
    return C.c_void_p(_addr(a))


def _i(v):
    return C.byref(C.c_int(v))


def _fluxes(seed=0):
    rng = np.random.default_rng(seed)
    up = np.asfortranarray(rng.uniform(0, 30, (NCOL, NLEV, NGPT)))
    dn = np.asfortranarray(rng.uniform(0, 30, (NCOL, NLEV, NGPT)))
    lims = np.asfortranarray(np.array(BANDS, dtype=np.int32).T)
    return up, dn, lims


def _byband(lib, device):
    ctx = Context(lib, device)
    up, dn, lims = _fluxes()
    d_up, d_dn, d_l = ctx.put(up), ctx.put(dn), ctx.put(lims)
    nb = len(BANDS)
    out = {k: ctx.zeros((NCOL, NLEV, nb)) for k in ("up", "dn", "net_full", "net_pre")}
    c = ctx.c
    c.rte_sum_byband(_i(NCOL), _i(NLEV), _i(NGPT), _i(nb), _ptr(d_l), _ptr(d_up), _ptr(out["up"]))
    c.rte_sum_byband(_i(NCOL), _i(NLEV), _i(NGPT), _i(nb), _ptr(d_l), _ptr(d_dn), _ptr(out["dn"]))
    c.rte_net_byband_full(_i(NCOL), _i(NLEV), _i(NGPT), _i(nb), _ptr(d_l), _ptr(d_dn), _ptr(d_up), _ptr(out["net_full"]))
    c.net_byband_precalc(_i(NCOL), _i(NLEV), _i(nb), _ptr(out["dn"]), _ptr(out["up"]), _ptr(out["net_pre"]))
    return {k: ctx.get(v) for k, v in out.items()}


def _heating(lib, device, top_at_1):
    ctx = Context(lib, device)
    rng = np.random.default_rng(5)
    nlay = NLEV - 1
    p = np.sort(rng.uniform(100.0, 101000.0, (NCOL, NLEV)), axis=1)
    if not top_at_1:
        p = p[:, ::-1]
    up, dn = rng.uniform(0, 400, (NCOL, NLEV)), rng.uniform(0, 400, (NCOL, NLEV))
    dr = dn * rng.uniform(0, 1, (NCOL, NLEV))
    mu0 = np.full((NCOL, nlay), 0.5) + rng.uniform(0, 0.4, (NCOL, nlay))
    # the sun sets inside some columns: mu0 = 0 in the layers nearest the surface
    for c in range(0, NCOL, 2):
        k = int(rng.integers(2, nlay - 2))
        if top_at_1:
            mu0[c, k:] = 0.0
        else:
            mu0[c, :nlay - k] = 0.0
    f = np.asfortranarray
    args = [ctx.put(f(a)) for a in (up, dn, dr, p, mu0)]
    hr, hrs = ctx.zeros((NCOL, nlay)), ctx.zeros((NCOL, nlay))
    ctx.c.rrtmgpb_heating_rate(NCOL, nlay, _ptr(args[0]), _ptr(args[1]), _ptr(args[3]), _ptr(hr))
    ctx.c.rrtmgpb_heating_rate_solar_varmu0(NCOL, nlay, _ptr(args[0]), _ptr(args[1]), _ptr(args[2]), _ptr(args[3]),
                                            _ptr(args[4]), _ptr(hrs))
    return ctx.get(hr), ctx.get(hrs), (up, dn, dr, p, mu0)


def test_oracle_byband_is_the_reference_loop(oracle_lib):
    got = _byband(oracle_lib, None)
    up, dn, _ = _fluxes()
    for ib, (a, b) in enumerate(BANDS):
        s_up, s_dn, net = up[:, :, a - 1].copy(), dn[:, :, a - 1].copy(), dn[:, :, a - 1] - up[:, :, a - 1]
        for g in range(a, b):
            s_up, s_dn = s_up + up[:, :, g], s_dn + dn[:, :, g]
            net = net + dn[:, :, g] - up[:, :, g]
        np.testing.assert_array_equal(got["up"][:, :, ib], s_up)
        np.testing.assert_array_equal(got["dn"][:, :, ib], s_dn)
        np.testing.assert_array_equal(got["net_full"][:, :, ib], net)
    np.testing.assert_array_equal(got["net_pre"], got["dn"] - got["up"])
    # summing the bands recovers the broadband flux to rounding
    np.testing.assert_allclose(got["up"].sum(axis=2), up.sum(axis=2), rtol=1e-14)


@pytest.mark.parametrize("top_at_1", [True, False])
def test_oracle_heating_rates(oracle_lib, top_at_1):
    hr, hrs, (up, dn, dr, p, mu0) = _heating(oracle_lib, None, top_at_1)
    grav, cp = 9.80665, 1004.64
    ref = (up[:, 1:] - up[:, :-1] - dn[:, 1:] + dn[:, :-1]) * grav / (cp * (p[:, 1:] - p[:, :-1]))
    np.testing.assert_array_equal(hr, ref)
    # expected result from the Fortran semantics of :93-116, stated independently with numpy
    eps = np.finfo(float).eps
    exp = ref.copy()
    nlay = mu0.shape[1]
    assert (mu0 < eps).any()
    masked_min = np.where(mu0 > 0, mu0, np.inf)
    masked_max = np.where(mu0 > 0, mu0, -np.inf)
    has = (mu0 > 0).any(axis=1)
    if (mu0[:, nlay - 1] < eps).any():   # minloc(mu0, mask=mu0>0, dim=2) + 1   (1-based; 0 when the mask is empty)
        last = np.where(has, masked_min.argmin(axis=1) + 1, 0) + 1
    else:                                # maxloc(...) - 1
        last = np.where(has, masked_max.argmax(axis=1) + 1, 0) - 1
    for c in range(mu0.shape[0]):
        il = int(last[c])
        if 1 < il < nlay:
            l = il - 1
            exp[c, l] = (up[c, l + 1] - up[c, l] - dn[c, l + 1] + dn[c, l] + dr[c, l + 1] - dr[c, l]) * grav / (cp * (p[c, l + 1] - p[c, l]))
    np.testing.assert_array_equal(hrs, exp)
    assert (hrs != hr).any()


@pytest.mark.gpu
def test_cuda_byband_bit_exact(oracle_lib, cuda_lib):
    ref, got = _byband(oracle_lib, None), _byband(cuda_lib, "cuda:0")
    for k in ref:
        np.testing.assert_array_equal(got[k], ref[k], err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize("top_at_1", [True, False])
def test_cuda_heating_rates(oracle_lib, cuda_lib, top_at_1):
    r0, r1, _ = _heating(oracle_lib, None, top_at_1)
    g0, g1, _ = _heating(cuda_lib, "cuda:0", top_at_1)
    np.testing.assert_allclose(g0, r0, rtol=1e-14, atol=0)
    np.testing.assert_allclose(g1, r1, rtol=1e-14, atol=0)
    assert np.array_equal(g1 != g0, r1 != r0)
