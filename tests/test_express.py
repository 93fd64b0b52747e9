"""Express path: rrtmgpb_rte_lw_express / _sw_express - state in, broadband fluxes out, no
(ncol, nlay, ngpt) array.  On the oracle the entry point IS the reference call sequence (gas optics, clouds%increment,
rte_lw / rte_sw); the CUDA implementation (column chunks x band groups through an L2-sized scratch, register solvers
in accumulate mode) must match it within the reference's flux tolerance, 1e-5 W/m2
(examples/compare-to-reference.py:56-61), on every all-sky test configuration - and for every chunking: chunk widths
that do not divide ncol (odd last chunk), several bands per launch, 1-3 grid rows."""
import os

import numpy as np
import pytest

from rte_rrtmgp_b200 import synthetic as syn
from rte_rrtmgp_b200.allsky import AllSky
from rte_rrtmgp_b200.frontend import Context

FLUX_ATOL = 1.0e-5


@pytest.fixture(scope="module")
def kdists():
    return syn.make_kdist("lw"), syn.make_kdist("sw")


def _run(lib, device, ncol, nlay, kd_lw, kd_sw, express, **kw):
    a = AllSky(Context(lib, device), ncol, nlay, kd_lw, kd_sw, express=express, **kw)
    a.step()
    return a.fluxes_host()


class _Env:
    def __init__(self, **kv):
        self.kv = {k: str(v) for k, v in kv.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_oracle_express_is_the_reference_sequence(oracle_lib, kdists):
    kd_lw, kd_sw = kdists
    a = _run(oracle_lib, None, 12, 30, kd_lw, kd_sw, express=True)
    b = _run(oracle_lib, None, 12, 30, kd_lw, kd_sw, express=False, fused=False)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize("nc,bands,rows", [(2368, 1, 2), (16, 1, 1), (22, 3, 3), (64, 16, 2)])
@pytest.mark.parametrize("ncol,nlay", [(24, 72), (37, 60), (130, 72), (21, 78), (19, 96), (18, 150)])
def test_express_replicated_profile(oracle_lib, cuda_lib, kdists, ncol, nlay, nc, bands, rows):
    """(18, 150): beyond the register solvers' layer range (144) - the express entry falls back to one launch per chunk."""
    kd_lw, kd_sw = kdists
    ref = _run(oracle_lib, None, ncol, nlay, kd_lw, kd_sw, express=False, fused=False)
    with _Env(RRTMGPB_EXPRESS_NC=nc, RRTMGPB_EXPRESS_BANDS=bands, RRTMGPB_EXPRESS_ROWS=rows):
        got = _run(cuda_lib, "cuda:0", ncol, nlay, kd_lw, kd_sw, express=True)
    for k in ref:
        assert np.max(np.abs(got[k] - ref[k])) <= FLUX_ATOL, k


@pytest.mark.gpu
@pytest.mark.parametrize("top_at_1", [True, False])
def test_express_distinct_columns_clear_and_cloudy(oracle_lib, cuda_lib, kdists, top_at_1):
    kd_lw, kd_sw = kdists
    ncol, nlay = 96, 60
    prof = syn.perturbed_profiles(ncol, nlay, seed=1234, top_at_1=top_at_1)
    for clouds in (True, False):
        ref = _run(oracle_lib, None, ncol, nlay, kd_lw, kd_sw, express=False, fused=False, profiles=prof, do_clouds=clouds)
        with _Env(RRTMGPB_EXPRESS_NC=40, RRTMGPB_EXPRESS_BANDS=2, RRTMGPB_EXPRESS_ROWS=2):
            got = _run(cuda_lib, "cuda:0", ncol, nlay, kd_lw, kd_sw, express=True, profiles=prof, do_clouds=clouds)
        for k in ref:
            assert np.max(np.abs(got[k] - ref[k])) <= FLUX_ATOL, (k, clouds)


@pytest.mark.gpu
def test_express_reduced_and_ragged_kdists(oracle_lib, cuda_lib):
    cases = [(syn.make_kdist("lw", ngpt=128), syn.make_kdist("sw", ngpt=112)),
             (syn.make_kdist("lw", band_sizes=[3, 17, 16, 20, 1, 2, 37, 5, 16, 16, 7, 8, 9, 10, 11, 12], seed=5),
              syn.make_kdist("sw", band_sizes=[16, 1, 33, 4, 6, 16, 18, 2, 3, 5, 7, 16, 16, 9], seed=6))]
    for kd_lw, kd_sw in cases:
        ref = _run(oracle_lib, None, 40, 72, kd_lw, kd_sw, express=False, fused=False)
        with _Env(RRTMGPB_EXPRESS_NC=24, RRTMGPB_EXPRESS_BANDS=3, RRTMGPB_EXPRESS_ROWS=2):
            got = _run(cuda_lib, "cuda:0", 40, 72, kd_lw, kd_sw, express=True)
        for k in ref:
            assert np.max(np.abs(got[k] - ref[k])) <= FLUX_ATOL, k


@pytest.mark.gpu
def test_express_equals_plane_path_on_the_gpu_and_checks_inputs(cuda_lib, kdists):
    from rte_rrtmgp_b200.frontend import FluxesBroadband, rte_lw_express

    kd_lw, kd_sw = kdists
    a = AllSky(Context(cuda_lib, "cuda:0"), 200, 72, kd_lw, kd_sw, express=True)
    b = AllSky(Context(cuda_lib, "cuda:0"), 200, 72, kd_lw, kd_sw, express=False)
    a.step(); b.step()
    fa, fb = a.fluxes_host(), b.fluxes_host()
    for k in fa:   # same kernels, same per-g-point arithmetic: only the broadband sums are associated by band
        np.testing.assert_allclose(fa[k], fb[k], rtol=1e-12, atol=1e-9, err_msg=k)
    with pytest.raises(RuntimeError, match="rte_lw: no space allocated for fluxes"):
        rte_lw_express(a.ctx, a.lw.go, a.p_lay, a.p_lev, a.t_lay, a.lw.t_sfc, a.vmr, a.lw.emis_sfc, FluxesBroadband())
    bad = a.ctx.put(np.full((kd_lw.nbnd, 200), 1.5, order="F"))
    with pytest.raises(RuntimeError, match="rte_lw: sfc_emis has values < 0 or > 1"):
        rte_lw_express(a.ctx, a.lw.go, a.p_lay, a.p_lev, a.t_lay, a.lw.t_sfc, a.vmr, bad, a.lw.fluxes)
