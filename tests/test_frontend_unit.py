"""The reference's solver unit tests driven through the FRONTEND mirror (rte_lw / rte_sw with
ty_fluxes_broadband), i.e. as tests/rte_lw_solver_unit_tests.F90:94-226 and
tests/rte_sw_solver_unit_tests.F90:98-211 call them: net-flux variants, increments by a transparent medium,
error strings.  Oracle backend on CPU; CUDA backend under -m gpu."""
import numpy as np
import pytest

import refcases as rc
from rte_rrtmgp_b200.frontend import Context, FluxesBroadband, OpticalProps, SourceFuncLW, rte_lw, rte_sw

NCOL, NLAY = 8, 16
SFC_T = np.array([285.0] * 4 + [310.0] * 4)
TOTAL_TAU = np.array([0.1, 1.0, 10.0, 50.0] * 2)
LIMS = np.array([[1], [1]], dtype=np.int32)
WVN_LW = np.array([[0.0], [3250.0]])
WVN_SW = np.array([[3250.0], [1.0e5]])


def _lw_problem(ctx, top_at_1=True):
    prob = rc.gray_rad_equil(SFC_T, TOTAL_TAU, NLAY, top_at_1)
    atmos = OpticalProps(ctx, "1scl", NCOL, NLAY, LIMS, WVN_LW, top_at_1=top_at_1, name="Gray atmosphere")
    atmos.tau = ctx.put(prob["tau"])
    src = SourceFuncLW(ctx, NCOL, NLAY, 1)
    src.lay_source, src.lev_source = ctx.put(prob["lay_source"]), ctx.put(prob["lev_source"])
    src.sfc_source, src.sfc_source_Jac = ctx.put(prob["sfc_source"]), ctx.put(prob["sfc_source_Jac"])
    return atmos, src, ctx.put(np.ones((1, NCOL), order="F"))


def test_rte_lw_net_flux_variants_and_olr(backend):
    lib, device = backend
    ctx = Context(lib, device)
    tol = 8.0 if device is None else 16.0
    atmos, src, emis = _lw_problem(ctx)
    up, dn, net = (ctx.zeros((NCOL, NLAY + 1)) for _ in range(3))
    rte_lw(ctx, atmos, src, emis, FluxesBroadband(flux_up=up, flux_dn=dn, flux_net=net))
    up_h, dn_h, net_h = ctx.get(up), ctx.get(dn), ctx.get(net)
    assert rc.allclose(up_h[:, 0], rc.gray_rad_equil_olr(SFC_T, TOTAL_TAU), tol=tol)
    assert rc.allclose(net_h, dn_h - up_h)  # :113-118
    net2 = ctx.zeros((NCOL, NLAY + 1))
    rte_lw(ctx, atmos, src, emis, FluxesBroadband(flux_net=net2))  # net only (:122-128)
    assert rc.allclose(ctx.get(net2), dn_h - up_h)
    up2, dn2 = ctx.zeros((NCOL, NLAY + 1)), ctx.zeros((NCOL, NLAY + 1))
    rte_lw(ctx, atmos, src, emis, FluxesBroadband(flux_up=up2, flux_dn=dn2))  # up/down only (:132-138)
    assert rc.allclose(ctx.get(dn2) - ctx.get(up2), net_h)


def test_rte_lw_increment_with_transparent_and_two_stream_props(backend):
    lib, device = backend
    ctx = Context(lib, device)
    tol = 2.0 if device is None else 8.0  # spacings; GPU: FMA contraction / libdevice vs glibc
    atmos, src, emis = _lw_problem(ctx)
    up, dn = ctx.zeros((NCOL, NLAY + 1)), ctx.zeros((NCOL, NLAY + 1))
    fl = FluxesBroadband(flux_up=up, flux_dn=dn)
    rte_lw(ctx, atmos, src, emis, fl)
    ref_up, ref_dn = ctx.get(up), ctx.get(dn)
    for kind in ("1scl", "2str"):  # mo_comparisons.F90:164-218 increment_with_*
        transparent = OpticalProps.like(ctx, kind, NCOL, NLAY, atmos)
        transparent.increment(atmos)
        atmos.validate()
        rte_lw(ctx, atmos, src, emis, fl)
        assert rc.allclose(ctx.get(up), ref_up, tol) and rc.allclose(ctx.get(dn), ref_dn, tol)
    # 2-stream properties with ssa = g = 0 through the rescaled solver (:194-210), with a Jacobian
    sw_atmos = OpticalProps.like(ctx, "2str", NCOL, NLAY, atmos)
    sw_atmos.tau = atmos.tau
    jac = ctx.zeros((NCOL, NLAY + 1))
    rte_lw(ctx, sw_atmos, src, emis, fl, flux_up_Jac=jac)
    assert rc.allclose(ctx.get(up), ref_up, tol) and rc.allclose(ctx.get(dn), ref_dn, tol)
    # three Gauss angles integrate the same gray problem to a close (not identical) answer: up to ~5% for the
    # optically thick columns
    rte_lw(ctx, atmos, src, emis, fl, n_gauss_angles=3)
    d3 = np.max(np.abs(ctx.get(up) - ref_up) / ref_up)
    assert 1e-4 < d3 < 0.06


def test_rte_lw_error_strings(backend):
    lib, device = backend
    ctx = Context(lib, device)
    atmos, src, emis = _lw_problem(ctx)
    up = ctx.zeros((NCOL, NLAY + 1))
    with pytest.raises(RuntimeError, match="no space allocated for fluxes"):
        rte_lw(ctx, atmos, src, emis, FluxesBroadband())
    with pytest.raises(RuntimeError, match="can't use two-stream methods with only absorption optical depth"):
        rte_lw(ctx, atmos, src, emis, FluxesBroadband(flux_up=up), use_2stream=1)
    with pytest.raises(RuntimeError, match="sfc_emis has values < 0 or > 1"):
        rte_lw(ctx, atmos, src, ctx.put(np.full((1, NCOL), 1.5, order="F")), FluxesBroadband(flux_up=up))
    bad = OpticalProps.like(ctx, "1scl", NCOL, NLAY, atmos, name="Gray atmosphere")
    bad.tau = ctx.put(-np.ones((NCOL, NLAY, 1), order="F"))
    with pytest.raises(RuntimeError, match="tau values out of range"):
        rte_lw(ctx, bad, src, emis, FluxesBroadband(flux_up=up))
    ctx.config_checks(True, False)  # rte_config_checks(values = false): the same call now passes
    rte_lw(ctx, bad, src, emis, FluxesBroadband(flux_up=up))
    ctx.config_checks(True, True)


@pytest.mark.parametrize("mu0", [1.0, 0.5])
def test_rte_sw_net_flux_variants_and_direct_beam(backend, mu0):
    lib, device = backend
    ctx = Context(lib, device)
    prob = rc.thin_scattering(lib, device, np.array([1e-4, 1e-2]), 1.0 - np.array([1e-4, 1e-2]), np.array([0.85, 0.65]), NLAY)
    atmos = OpticalProps(ctx, "2str", NCOL, NLAY, LIMS, WVN_SW, top_at_1=True, name="Gray SW atmosphere")
    atmos.tau, atmos.ssa, atmos.g = ctx.put(prob["tau"]), ctx.put(prob["ssa"]), ctx.put(prob["g"])
    alb = ctx.put(np.zeros((1, NCOL), order="F"))
    toa = ctx.put(np.ones((NCOL, 1), order="F"))
    mu0_arr = ctx.put(np.full(NCOL, mu0))
    up, dn, dr, net = (ctx.zeros((NCOL, NLAY + 1)) for _ in range(4))
    rte_sw(ctx, atmos, mu0_arr, toa, alb, alb, FluxesBroadband(flux_up=up, flux_dn=dn, flux_dn_dir=dr, flux_net=net))
    up_h, dn_h, dr_h, net_h = ctx.get(up), ctx.get(dn), ctx.get(dr), ctx.get(net)
    beer = mu0 * np.exp(-np.sum(prob["tau"][:, :, 0], axis=1) / mu0)
    assert rc.allclose(dr_h[:, NLAY], beer, tol=20.0)  # rte_sw_solver_unit_tests.F90:121-130
    assert rc.allclose(net_h, dn_h - up_h)
    net2 = ctx.zeros((NCOL, NLAY + 1))
    rte_sw(ctx, atmos, mu0_arr, toa, alb, alb, FluxesBroadband(flux_net=net2))
    assert rc.allclose(ctx.get(net2), dn_h - up_h)
    transparent = OpticalProps.like(ctx, "2str", NCOL, NLAY, atmos)
    transparent.increment(atmos)
    rte_sw(ctx, atmos, mu0_arr, toa, alb, alb, FluxesBroadband(flux_up=up, flux_dn=dn))
    # incrementing by a transparent medium perturbs ssa and g in the last bit ((t*w*g)/(t*w) != g exactly), so
    # the fluxes agree to rounding, not to 2 spacings (the reference checks the properties, not the fluxes)
    np.testing.assert_allclose(ctx.get(up), up_h, rtol=1e-8, atol=1e-18)  # nearly conservative scattering: ill-conditioned
    np.testing.assert_allclose(ctx.get(dn), dn_h, rtol=1e-8, atol=1e-18)
    with pytest.raises(RuntimeError, match="one or more mu0 < -1 or > 1"):
        rte_sw(ctx, atmos, ctx.put(np.full(NCOL, 1.5)), toa, alb, alb, FluxesBroadband(flux_up=up))
