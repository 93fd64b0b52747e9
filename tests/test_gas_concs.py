"""ty_gas_concs mirror (rte/frontend/gas-optics-template/mo_gas_concentrations.F90): scalar / profile / field storage,
get_vmr broadcast (:433-504) and the error strings of set_vmr / get_vmr; the all-sky workload driven through it
equals the workload driven by the dense vmr(ncol,nlay,ngas) array."""
import numpy as np
import pytest

from rte_rrtmgp_b200 import synthetic as syn
from rte_rrtmgp_b200.frontend import Context, GasConcs


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def ctx(request):
    if request.param == "oracle":
        import oracle

        return Context(oracle.lib(), None)
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import rte_rrtmgp_b200

    return Context(rte_rrtmgp_b200.lib(), "cuda:0")


def test_get_vmr_broadcasts_scalar_profile_and_field(ctx):
    ncol, nlay = 37, 11  # ragged on purpose
    rng = np.random.default_rng(5)
    prof = rng.uniform(0.0, 1.0, nlay)
    field = np.asfortranarray(rng.uniform(0.0, 1.0, (ncol, nlay)))
    gc = GasConcs(ctx, ["h2o", "CO2 ", "o3"])
    gc.set_vmr("co2", 348.0e-6)
    gc.set_vmr("o3", prof)
    gc.set_vmr("H2O", field)
    vmr = ctx.zeros((ncol, nlay, 3))
    gc.fill_vmr(["h2o", "co2", "o3"], vmr)
    got = ctx.get(vmr)
    # bit-exact: get_vmr only copies
    assert np.array_equal(got[:, :, 0], field)
    assert np.array_equal(got[:, :, 1], np.full((ncol, nlay), 348.0e-6))
    assert np.array_equal(got[:, :, 2], np.broadcast_to(prof[None, :], (ncol, nlay)))
    # setting a gas again replaces its storage kind (set_vmr_scalar :154-166)
    gc.set_vmr("o3", 1.0e-6)
    gc.fill_vmr(["o3"], vmr)
    assert np.array_equal(ctx.get(vmr)[:, :, 0], np.full((ncol, nlay), 1.0e-6))


def test_error_strings(ctx):
    gc = GasConcs(ctx, ["h2o", "co2"])
    with pytest.raises(RuntimeError, match="name not provided at initialization"):
        gc.set_vmr("ch4", 1.0e-6)
    with pytest.raises(RuntimeError, match="concentrations should be >= 0, <= 1"):
        gc.set_vmr("co2", 1.5)
    out = ctx.zeros((4, 3))
    with pytest.raises(RuntimeError, match="gas ch4 not found"):
        gc.get_vmr("ch4", out, 4, 3)
    with pytest.raises(RuntimeError, match="concentration hasn't been set"):
        gc.get_vmr("co2", out, 4, 3)
    gc.set_vmr("h2o", np.zeros((4, 3), order="F"))
    with pytest.raises(RuntimeError, match=r"wrong size \(ncol\)"):
        gc.get_vmr("h2o", ctx.zeros((5, 3)), 5, 3)
    with pytest.raises(RuntimeError, match=r"wrong size \(nlay\)"):
        gc.get_vmr("h2o", ctx.zeros((4, 2)), 4, 2)
    with pytest.raises(RuntimeError, match=r"different dimension \(ncol\)"):
        gc.set_vmr("co2", np.zeros((6, 3), order="F"))


def test_allsky_vmr_from_gas_concs_equals_dense(ctx):
    from rte_rrtmgp_b200.allsky import AllSky

    ncol, nlay = 12, 16
    kd = syn.make_kdist("lw", gpt_per_band=2)
    sky = AllSky(ctx, ncol, nlay, kd, None, do_clouds=False)
    dense = syn.allsky_gas_vmrs(syn.compute_profiles(300.0, ncol, nlay))
    assert np.array_equal(ctx.get(sky.vmr), dense)
