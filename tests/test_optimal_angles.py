"""ty_gas_optics_rrtmgp%compute_optimal_angles (rrtmgp/frontend/mo_gas_optics_rrtmgp.F90:1503-1562):
secant of the LW transport angle from the column transmissivity, and its use as rte_lw's lw_Ds."""
import numpy as np
import pytest

from rte_rrtmgp_b200 import synthetic as syn
from rte_rrtmgp_b200.allsky import AllSky
from rte_rrtmgp_b200.frontend import Context, FluxesBroadband, rte_lw


def _ctx(kind):
    if kind == "oracle":
        import oracle

        return Context(oracle.lib(), None)
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import rte_rrtmgp_b200

    return Context(rte_rrtmgp_b200.lib(), "cuda:0")


def _lw_state(ctx, ncol, nlay, kd):
    sky = AllSky(ctx, ncol, nlay, kd, None, do_clouds=False, fused=False)
    lw = sky.lw
    lw.go.gas_optics(sky.p_lay, sky.p_lev, sky.t_lay, sky.vmr, lw.atmos, t_sfc=lw.t_sfc, sources=lw.sources,
                     tlev=sky.t_lev)
    return sky


def test_oracle_matches_the_formula():
    ctx = _ctx("oracle")
    ncol, nlay = 7, 16
    kd = syn.make_kdist("lw", gpt_per_band=3)
    sky = _lw_state(ctx, ncol, nlay, kd)
    ds = sky.lw.go.compute_optimal_angles(sky.lw.atmos)
    tau = ctx.get(sky.lw.atmos.tau)
    fit = kd.extra["optimal_angle_fit"]
    t = np.zeros((ncol, kd.ngpt))
    for l in range(nlay):  # the reference's summation order
        t = t + tau[:, l, :]
    band = kd.gpoint_bands - 1
    want = fit[0, band][None, :] * np.exp(-t) + fit[1, band][None, :]
    np.testing.assert_allclose(ctx.get(ds), want, rtol=2e-16, atol=0)
    assert np.all(ctx.get(ds) >= 1.0)  # rte_lw rejects secants below 1 (mo_rte_lw.F90:229-231)


def test_error_strings():
    ctx = _ctx("oracle")
    kd = syn.make_kdist("lw", gpt_per_band=2)
    sky = _lw_state(ctx, 4, 8, kd)
    with pytest.raises(RuntimeError, match="optimal_angles different dimension"):
        sky.lw.go.compute_optimal_angles(sky.lw.atmos, ctx.zeros((5, kd.ngpt)))
    other = syn.make_kdist("lw", gpt_per_band=3)
    sky2 = _lw_state(ctx, 4, 8, other)
    with pytest.raises(RuntimeError, match="different spectral discretization"):
        sky.lw.go.compute_optimal_angles(sky2.lw.atmos)


@pytest.mark.gpu
def test_cuda_matches_oracle_and_drives_rte_lw():
    ncol, nlay = 130, 72
    kd = syn.make_kdist("lw", gpt_per_band=4)
    res = {}
    for kind in ("oracle", "cuda"):
        ctx = _ctx(kind)
        sky = _lw_state(ctx, ncol, nlay, kd)
        lw = sky.lw
        ds = lw.go.compute_optimal_angles(lw.atmos)
        rte_lw(ctx, lw.atmos, lw.sources, lw.emis_sfc, lw.fluxes, lw_Ds=ds)
        res[kind] = (ctx.get(ds), ctx.get(lw.flux_up), ctx.get(lw.flux_dn))
    # secants: exp() of the two libraries differs by <= 1 ulp; fluxes: the repo's regression tolerance is 1e-5 W/m2
    np.testing.assert_allclose(res["cuda"][0], res["oracle"][0], rtol=1e-14, atol=0)
    for a, b in zip(res["cuda"][1:], res["oracle"][1:]):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-9)
