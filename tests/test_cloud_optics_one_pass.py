"""The one-pass cloud optics (rrtmgpb_cloud_optics_from_tables: masks + compute_cld_from_table x2 + liquid/ice combination,
mo_cloud_optics_rrtmgp.F90:334-424) against the reference's kernel-by-kernel sequence, for 1scl and 2str outputs,
on the oracle (where the one-pass entry IS the sequence: bit-identical) and on CUDA (vs the oracle, 1e-13 relative)."""
import numpy as np
import pytest

from rte_rrtmgp_b200 import synthetic as syn
from rte_rrtmgp_b200.frontend import CloudOptics, Context, OpticalProps


def _clouds(lib, device, kind, one_pass, ncol=37, nlay=60, lw=True, delta_scale=None):
    ctx = Context(lib, device)
    kd = syn.make_kdist("lw" if lw else "sw", gpt_per_band=2)
    prof = syn.perturbed_profiles(ncol, nlay, seed=21, top_at_1=True)
    lut = syn.make_cloud_lut(kd)
    cl = syn.compute_clouds(prof, lut)
    rng = np.random.default_rng(3)  # vary the particle sizes over the table range where there is cloud
    rel = np.where(cl["lwp"] > 0, rng.uniform(lut.radliq_lwr, lut.radliq_upr, cl["lwp"].shape), 0.0)
    dei = np.where(cl["iwp"] > 0, rng.uniform(lut.diamice_lwr, lut.diamice_upr, cl["iwp"].shape), 0.0)
    co = CloudOptics(ctx, lut)
    op = OpticalProps.like(ctx, kind, ncol, nlay, co)
    ctx.c.rrtmgpb_cloud_optics_one_pass(1 if one_pass else 0)
    try:
        args = (ctx.put(cl["lwp"]), ctx.put(cl["iwp"]), ctx.put(np.asfortranarray(rel)), ctx.put(np.asfortranarray(dei)), op)
        if delta_scale == "fused":      # cloud_optics + delta_scale in one call
            co.cloud_optics(*args, delta_scale=True)
        else:
            co.cloud_optics(*args)
            if delta_scale == "separate":  # the reference driver's two calls (rrtmgp_allsky.F90:350-352)
                op.delta_scale()
    finally:
        ctx.c.rrtmgpb_cloud_optics_one_pass(1)
    out = {"tau": ctx.get(op.tau)}
    if kind == "2str":
        out["ssa"], out["g"] = ctx.get(op.ssa), ctx.get(op.g)
    return out


@pytest.mark.parametrize("kind", ["1scl", "2str"])
def test_oracle_one_pass_is_the_reference_sequence(oracle_lib, kind):
    a, b = _clouds(oracle_lib, None, kind, True), _clouds(oracle_lib, None, kind, False)
    assert np.any(a["tau"] > 0)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize("lw", [True, False])
@pytest.mark.parametrize("kind", ["1scl", "2str"])
def test_cuda_one_pass_and_sequence_match_oracle(oracle_lib, cuda_lib, kind, lw):
    ref = _clouds(oracle_lib, None, kind, False, lw=lw)
    for one_pass in (True, False):
        got = _clouds(cuda_lib, "cuda:0", kind, one_pass, lw=lw)
        for k in ref:
            np.testing.assert_allclose(got[k], ref[k], rtol=1e-13, atol=1e-300, err_msg=f"{k} one_pass={one_pass}")


def test_oracle_delta_scaled_is_the_two_calls(oracle_lib):
    a = _clouds(oracle_lib, None, "2str", True, lw=False, delta_scale="fused")
    b = _clouds(oracle_lib, None, "2str", False, lw=False, delta_scale="separate")
    plain = _clouds(oracle_lib, None, "2str", True, lw=False)
    assert np.any(a["tau"] != plain["tau"])  # the scaling did something
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize("one_pass", [True, False])
def test_cuda_delta_scaled_matches_oracle(oracle_lib, cuda_lib, one_pass):
    ref = _clouds(oracle_lib, None, "2str", False, lw=False, delta_scale="separate")
    got = _clouds(cuda_lib, "cuda:0", "2str", one_pass, lw=False, delta_scale="fused")
    for k in ref:
        np.testing.assert_allclose(got[k], ref[k], rtol=1e-13, atol=1e-300, err_msg=k)
