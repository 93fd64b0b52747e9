"""The lanes-along-g-points mapping of the tau kernels (csrc/kernels/gas_optics_gfast.cuh: tau_band_rows - taken by blocks
whose cells do not share table rows: unrelated neighbouring columns, BASELINE config 3) against the cells-per-thread mapping
and the oracle: distinct columns with and without clouds, both orientations, a partial last block, the fused entry and the
extern symbol rrtmgp_compute_tau_absorption (ABI instantiation), mixed blocks (replicated columns beside distinct ones), and
a ragged k-distribution (irregular bands must keep the old mapping).  Same expressions in both mappings: tau, ssa, g agree
BIT FOR BIT."""
import numpy as np
import pytest

from rte_rrtmgp_b200 import synthetic as syn
from rte_rrtmgp_b200.allsky import AllSky
from rte_rrtmgp_b200.frontend import Context

FLUX_ATOL = 1.0e-5


def _run(lib, device, ncol, nlay, kd_lw, kd_sw, profiles, do_clouds, fused):
    a = AllSky(Context(lib, device), ncol, nlay, kd_lw, kd_sw, do_clouds=do_clouds, profiles=profiles, fused=fused)
    a.step()
    return a


def _planes(a):
    return {"lw tau": a.ctx.get(a.lw.atmos.tau), "sw tau": a.ctx.get(a.sw.atmos.tau), "sw ssa": a.ctx.get(a.sw.atmos.ssa),
            "sw g": a.ctx.get(a.sw.atmos.g)}


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("do_clouds", [False, True])
@pytest.mark.parametrize("top_at_1", [True, False])
def test_rows_path_distinct_columns(oracle_lib, cuda_lib, top_at_1, do_clouds, fused):
    kd_lw, kd_sw = syn.make_kdist("lw"), syn.make_kdist("sw")
    ncol, nlay = 300, 60   # 18,000 cells: 70 full blocks and a partial one
    prof = syn.perturbed_profiles(ncol, nlay, seed=77, top_at_1=top_at_1)
    c = _run(oracle_lib, None, ncol, nlay, kd_lw, kd_sw, prof, do_clouds, False)
    try:
        cuda_lib.cdll.rrtmgpb_set_gas_optics_rows_path(0)
        g0 = _run(cuda_lib, "cuda:0", ncol, nlay, kd_lw, kd_sw, prof, do_clouds, fused)
        cuda_lib.cdll.rrtmgpb_set_gas_optics_rows_path(1)
        g1 = _run(cuda_lib, "cuda:0", ncol, nlay, kd_lw, kd_sw, prof, do_clouds, fused)
    finally:
        cuda_lib.cdll.rrtmgpb_set_gas_optics_rows_path(-1)
    p0, p1, pc = _planes(g0), _planes(g1), _planes(c)
    for k in pc:
        np.testing.assert_allclose(p1[k], pc[k], rtol=1e-12, atol=1e-300, err_msg=k)
        assert np.array_equal(p0[k], p1[k]), f"{k}: the two mappings differ (max {np.max(np.abs(p0[k] - p1[k])):.3e})"
    f1, fc = g1.fluxes_host(), c.fluxes_host()
    for k in fc:
        assert np.max(np.abs(f1[k] - fc[k])) <= FLUX_ATOL, k


@pytest.mark.gpu
def test_rows_path_mixed_and_ragged(oracle_lib, cuda_lib):
    """Half the columns replicated (blocks keep the cells-per-thread mapping), half distinct (blocks switch); and a ragged
    k-distribution, whose irregular bands never switch."""
    nlay = 60
    d = syn.perturbed_profiles(512, nlay, seed=5, top_at_1=True)
    mixed = {k: np.asfortranarray(np.concatenate([np.repeat(v[:1], 512, axis=0), v], axis=0)) for k, v in d.items()}
    cases = [(syn.make_kdist("lw"), syn.make_kdist("sw"), mixed, 1024),
             (syn.make_kdist("lw", band_sizes=[3, 17, 16, 20, 1, 2, 37, 5, 16, 16, 7, 8, 9, 10, 11, 12], seed=5),
              syn.make_kdist("sw", band_sizes=[16, 1, 33, 4, 6, 16, 18, 2, 3, 5, 7, 16, 16, 9], seed=6), d, 512)]
    for kd_lw, kd_sw, prof, ncol in cases:
        c = _run(oracle_lib, None, ncol, nlay, kd_lw, kd_sw, prof, False, False)
        try:
            cuda_lib.cdll.rrtmgpb_set_gas_optics_rows_path(1)
            g = _run(cuda_lib, "cuda:0", ncol, nlay, kd_lw, kd_sw, prof, False, True)
        finally:
            cuda_lib.cdll.rrtmgpb_set_gas_optics_rows_path(-1)
        pg, pc = _planes(g), _planes(c)
        for k in pc:
            np.testing.assert_allclose(pg[k], pc[k], rtol=1e-12, atol=1e-300, err_msg=k)


@pytest.mark.gpu
def test_rows_path_automatic_selection(oracle_lib, cuda_lib):
    """Default setting (-1, RRTMGPB_TAU_ROWS unset): the fused entry samples how often neighbouring cells fall into different
    T / p bins and the NEXT call picks the kernel instantiation from that.  Whatever it picks - distinct columns after
    replicated ones, replicated after distinct - every step must give the same planes as the pinned mappings."""
    kd_lw, kd_sw = syn.make_kdist("lw"), syn.make_kdist("sw")
    ncol, nlay = 512, 60
    prof = syn.perturbed_profiles(ncol, nlay, seed=21, top_at_1=True)
    cuda_lib.cdll.rrtmgpb_set_gas_optics_rows_path(0)
    try:
        ref_d = _planes(_run(cuda_lib, "cuda:0", ncol, nlay, kd_lw, kd_sw, prof, False, True))
        ref_r = _planes(_run(cuda_lib, "cuda:0", ncol, nlay, kd_lw, kd_sw, None, True, True))
    finally:
        cuda_lib.cdll.rrtmgpb_set_gas_optics_rows_path(-1)
    d = AllSky(Context(cuda_lib, "cuda:0"), ncol, nlay, kd_lw, kd_sw, do_clouds=False, profiles=prof, fused=True)
    r = AllSky(Context(cuda_lib, "cuda:0"), ncol, nlay, kd_lw, kd_sw, do_clouds=True, fused=True)
    # the kernel-by-kernel sequence: rrtmgp_compute_tau_absorption takes the same sample in its per-cell pre-pass
    ds = AllSky(Context(cuda_lib, "cuda:0"), ncol, nlay, kd_lw, kd_sw, do_clouds=False, profiles=prof, fused=False)
    for sky, ref in ((d, ref_d), (d, ref_d), (r, ref_r), (r, ref_r), (d, ref_d), (d, ref_d), (ds, ref_d), (ds, ref_d), (ds, ref_d)):
        sky.step()
        got = _planes(sky)
        for k in ref:
            assert np.array_equal(got[k], ref[k]), k
