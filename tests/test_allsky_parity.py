"""GPU parity of the whole hot path: one all-sky LW+SW iteration (reference loop body
examples/all-sky/rrtmgp_allsky.F90:332-409) through the C-ABI on CUDA vs the CPU oracle on identical
seeded inputs.  Tolerance: the reference's regression threshold for fluxes, 1e-5 W/m2 absolute
(examples/compare-to-reference.py:56-61); intermediates are compared relatively."""
import numpy as np
import pytest

from rte_rrtmgp_b200 import synthetic as syn
from rte_rrtmgp_b200.allsky import AllSky
from rte_rrtmgp_b200.frontend import Context

FLUX_ATOL = 1.0e-5  # W/m2, examples/compare-to-reference.py:58


@pytest.fixture(scope="module")
def kdists():
    return syn.make_kdist("lw"), syn.make_kdist("sw")


def _run(lib, device, ncol, nlay, kd_lw, kd_sw, profiles=None, do_clouds=True, fused=False):
    """Oracle runs use the reference call sequence (fused=False); CUDA runs are parametrised over both."""
    a = AllSky(Context(lib, device), ncol, nlay, kd_lw, kd_sw, do_clouds=do_clouds, profiles=profiles, fused=fused)
    a.step()
    return a


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("ncol,nlay", [(24, 72), (37, 60), (130, 72), (21, 78), (19, 96), (26, 137), (20, 150)])
def test_allsky_replicated_profile(oracle_lib, cuda_lib, kdists, ncol, nlay, variant, fused):
    """variant 0: register / warp-systolic solvers (8 lanes per column up to 80 layers, 16 lanes up to 144; 150 layers falls
    back to the tile kernels); 1: tile solvers."""
    kd_lw, kd_sw = kdists
    cuda_lib.cdll.rrtmgpb_set_solver_variant(variant)
    g = _run(cuda_lib, "cuda:0", ncol, nlay, kd_lw, kd_sw, fused=fused)
    c = _run(oracle_lib, None, ncol, nlay, kd_lw, kd_sw)
    fg, fc = g.fluxes_host(), c.fluxes_host()
    for k in fc:
        assert np.max(np.abs(fg[k] - fc[k])) <= FLUX_ATOL, k
    # API-visible intermediates (atmos%tau/ssa/g, lw_sources) relative to the oracle
    for name, ga, ca in (("lw tau", g.lw.atmos.tau, c.lw.atmos.tau), ("lay_source", g.lw.sources.lay_source, c.lw.sources.lay_source),
                         ("lev_source", g.lw.sources.lev_source, c.lw.sources.lev_source),
                         ("sw tau", g.sw.atmos.tau, c.sw.atmos.tau), ("sw ssa", g.sw.atmos.ssa, c.sw.atmos.ssa),
                         ("sw g", g.sw.atmos.g, c.sw.atmos.g)):
        x, y = g.ctx.get(ga), c.ctx.get(ca)
        np.testing.assert_allclose(x, y, rtol=1e-12, atol=1e-300, err_msg=name)
    cuda_lib.cdll.rrtmgpb_set_solver_variant(0)


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("top_at_1", [True, False])
def test_allsky_distinct_columns(oracle_lib, cuda_lib, kdists, top_at_1, fused):
    """RFMIP-like stand-in: every column different (no broadcast table access), both orientations."""
    kd_lw, kd_sw = kdists
    ncol, nlay = 96, 60
    prof = syn.perturbed_profiles(ncol, nlay, seed=1234, top_at_1=top_at_1)
    g = _run(cuda_lib, "cuda:0", ncol, nlay, kd_lw, kd_sw, profiles=prof, fused=fused)
    c = _run(oracle_lib, None, ncol, nlay, kd_lw, kd_sw, profiles=prof)
    assert g.lw.atmos.top_at_1 == top_at_1 == c.lw.atmos.top_at_1
    fg, fc = g.fluxes_host(), c.fluxes_host()
    for k in fc:
        assert np.max(np.abs(fg[k] - fc[k])) <= FLUX_ATOL, k


@pytest.mark.gpu
def test_clear_sky_and_reduced_kdist(oracle_lib, cuda_lib):
    """BASELINE configs 3/4 shapes at test size: clear sky; reduced 128/112 g-point k-distributions."""
    kd_lw, kd_sw = syn.make_kdist("lw", ngpt=128), syn.make_kdist("sw", ngpt=112)
    for clouds in (False, True):
        for fused in (False, True):
            g = _run(cuda_lib, "cuda:0", 40, 72, kd_lw, kd_sw, do_clouds=clouds, fused=fused)
            c = _run(oracle_lib, None, 40, 72, kd_lw, kd_sw, do_clouds=clouds)
            fg, fc = g.fluxes_host(), c.fluxes_host()
            for k in fc:
                assert np.max(np.abs(fg[k] - fc[k])) <= FLUX_ATOL, (k, clouds, fused)


@pytest.mark.gpu
@pytest.mark.parametrize("distinct", [False, True])
def test_ragged_bands_and_partial_minor_intervals(oracle_lib, cuda_lib, distinct):
    """Bands of 1..37 g-points (odd sizes -> 64-bit table loads; > 16 -> several register chunks) and minor
    contributors covering part of a band: the layouts the g-point-fastest gas-optics kernels must not assume away."""
    kd_lw = syn.make_kdist("lw", band_sizes=[3, 17, 16, 20, 1, 2, 37, 5, 16, 16, 7, 8, 9, 10, 11, 12], seed=5)
    kd_sw = syn.make_kdist("sw", band_sizes=[16, 1, 33, 4, 6, 16, 18, 2, 3, 5, 7, 16, 16, 9], seed=6)
    ncol, nlay = 70, 60
    prof = syn.perturbed_profiles(ncol, nlay, seed=99, top_at_1=True) if distinct else None
    c = _run(oracle_lib, None, ncol, nlay, kd_lw, kd_sw, profiles=prof)
    for fused in (False, True):
        g = _run(cuda_lib, "cuda:0", ncol, nlay, kd_lw, kd_sw, profiles=prof, fused=fused)
        fg, fc = g.fluxes_host(), c.fluxes_host()
        for k in fc:
            assert np.max(np.abs(fg[k] - fc[k])) <= FLUX_ATOL, (k, fused)
        for name, ga, ca in (("lw tau", g.lw.atmos.tau, c.lw.atmos.tau), ("lev_source", g.lw.sources.lev_source, c.lw.sources.lev_source),
                             ("sfc_source", g.lw.sources.sfc_source, c.lw.sources.sfc_source),
                             ("sw tau", g.sw.atmos.tau, c.sw.atmos.tau), ("sw ssa", g.sw.atmos.ssa, c.sw.atmos.ssa)):
            np.testing.assert_allclose(g.ctx.get(ga), c.ctx.get(ca), rtol=1e-12, atol=1e-300, err_msg=f"{name} fused={fused}")


def test_oracle_fused_entry_equals_reference_sequence(oracle_lib, kdists):
    """On the oracle the fused entry point is literally the reference sequence: results must be bit-identical."""
    kd_lw, kd_sw = kdists
    a = _run(oracle_lib, None, 12, 30, kd_lw, kd_sw, fused=True).fluxes_host()
    b = _run(oracle_lib, None, 12, 30, kd_lw, kd_sw, fused=False).fluxes_host()
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)


@pytest.mark.gpu
def test_smoke_entry():
    import __graft_entry__ as entry

    assert entry.smoke(ncol=24, nlay=72, verbose=False) <= FLUX_ATOL


@pytest.mark.gpu
def test_full_size_replicated_profile_properties(oracle_lib, cuda_lib, kdists):
    """BASELINE config 2 at its full size (65,536 columns x 72 layers, 256 + 224 g-points), checked through properties
    that do not need a CPU run of that size: (1) the profile is replicated and clouds repeat with period 3 in the
    column index (rrtmgp_allsky.F90:636-656), so every column must equal - bit for bit - the column of its class
    among the first three; (2) the first 48 columns match the oracle run on 48 columns within the flux tolerance."""
    import torch

    free, _ = torch.cuda.mem_get_info()
    if free < 100 * 2**30:
        pytest.skip("needs ~80 GB of device memory")
    kd_lw, kd_sw = kdists
    ncol, nlay = 65536, 72
    g = _run(cuda_lib, "cuda:0", ncol, nlay, kd_lw, kd_sw, fused=True)
    fg = g.fluxes_host()
    for k, v in fg.items():
        assert np.all(np.isfinite(v)), k
        for cls in range(3):
            same = v[cls::3]
            assert np.array_equal(same, np.broadcast_to(same[0], same.shape)), (k, cls)
    c = _run(oracle_lib, None, 48, nlay, kd_lw, kd_sw)
    fc = c.fluxes_host()
    for k in fc:
        assert np.max(np.abs(fg[k][:48] - fc[k])) <= FLUX_ATOL, k
    del g
    torch.cuda.empty_cache()
