"""Aerosol optics (ty_aerosol_optics_rrtmgp_merra, rrtmgp/frontend/mo_aerosol_optics_rrtmgp_merra.F90:233-578) and the
BASELINE config-5 path (all-sky LW two-stream + aerosols).

CPU tests check the oracle restatement against an independent numpy evaluation of the same table lookup and the
frontend's error strings; GPU tests compare CUDA with the oracle through the C frontend.  Flux tolerance: the
reference's regression threshold 1e-5 W/m2 (examples/compare-to-reference.py:56-61)."""
import numpy as np
import pytest

from rte_rrtmgp_b200 import synthetic as syn
from rte_rrtmgp_b200.allsky import AllSky
from rte_rrtmgp_b200.frontend import AerosolOptics, Context, OpticalProps

FLUX_ATOL = 1.0e-5


def _random_aerosols(ncol, nlay, lut, seed=3):
    rng = np.random.default_rng(seed)
    typ = rng.integers(0, 8, (ncol, nlay)).astype(np.int32)
    lo, hi = lut.merra_aero_bin_lims[0, 0], lut.merra_aero_bin_lims[1, -1]
    size = rng.uniform(lo, hi, (ncol, nlay))
    size[0, 0] = lut.merra_aero_bin_lims[1, 0]  # exactly on a shared bin edge: the LAST matching bin wins (:457-462)
    mass = rng.uniform(0.0, 1e-4, (ncol, nlay))
    rh = rng.uniform(0.0, 1.0, (ncol, nlay))
    rh[0, 1], rh[1, 0], rh[1, 1] = 0.0, 1.0, lut.aero_rh[5]  # below/at the first node, above the last, on a node
    typ[0, 0], typ[0, 1], typ[1, 0], typ[1, 1] = 2, 3, 4, 6
    f = np.asfortranarray
    return f(typ), f(size), f(mass), f(rh)


def _numpy_aerosol(lut, typ, size, mass, rh):
    """Independent evaluation: vectorised bin search / rh bracket, tables in the loader's (nval, nrh, ...) form."""
    nb = lut.nbnd
    lims, arh = lut.merra_aero_bin_lims, lut.aero_rh
    ibin = np.zeros(typ.shape, dtype=int)
    for i in range(lims.shape[1]):
        ibin = np.where((size >= lims[0, i]) & (size <= lims[1, i]), i, ibin)
    i2 = np.searchsorted(arh, rh, side="left")  # first node >= rh  (the reference's while rh > aero_rh(irh2))
    i1 = np.maximum(0, i2 - 1)
    i2 = np.minimum(len(arh) - 1, i2)
    w = np.where(i1 == i2, 0.0, (rh - arh[i1]) / np.where(i1 == i2, 1.0, arh[i2] - arh[i1]))
    t = np.zeros(typ.shape + (nb,)); ts = np.zeros_like(t); tsg = np.zeros_like(t)
    for b in range(nb):
        vals = np.zeros((3,) + typ.shape)
        for v in range(3):
            lin = lambda tab: tab[i1] + w * (tab[i2] - tab[i1])
            per_type = {
                1: lut.aero_dust_tbl[v, :, b][ibin],
                2: lut.aero_salt_tbl[v, :, :, b][i1, ibin] + w * (lut.aero_salt_tbl[v, :, :, b][i2, ibin] - lut.aero_salt_tbl[v, :, :, b][i1, ibin]),
                3: lin(lut.aero_sulf_tbl[v, :, b]), 4: lin(lut.aero_bcar_rh_tbl[v, :, b]),
                5: np.full(typ.shape, lut.aero_bcar_tbl[v, b]), 6: lin(lut.aero_ocar_rh_tbl[v, :, b]),
                7: np.full(typ.shape, lut.aero_ocar_tbl[v, b]),
            }
            for k, arr in per_type.items():
                vals[v] = np.where(typ == k, arr, vals[v])
        t[..., b] = mass * vals[0] * (typ > 0)
        ts[..., b] = t[..., b] * vals[1]
        tsg[..., b] = ts[..., b] * vals[2]
    return t, ts, tsg


@pytest.mark.parametrize("kind", ["1scl", "2str"])
def test_oracle_aerosol_optics_matches_numpy(oracle_lib, kind):
    kd = syn.make_kdist("sw" if kind == "2str" else "lw")
    lut = syn.make_aerosol_lut(kd)
    ctx = Context(oracle_lib, None)
    ncol, nlay = 9, 7
    typ, size, mass, rh = _random_aerosols(ncol, nlay, lut)
    ao = AerosolOptics(ctx, lut)
    op = OpticalProps.like(ctx, kind, ncol, nlay, ao)
    ao.aerosol_optics(typ, size, mass, rh, op)
    t, ts, tsg = _numpy_aerosol(lut, typ, size, mass, rh)
    eps = np.finfo(np.float64).eps
    if kind == "1scl":
        np.testing.assert_allclose(op.tau, t - ts, rtol=1e-14, atol=0)
    else:
        np.testing.assert_allclose(op.tau, t, rtol=1e-14, atol=0)
        np.testing.assert_allclose(op.ssa, ts / np.maximum(eps, t), rtol=1e-13, atol=0)
        np.testing.assert_allclose(op.g, tsg / np.maximum(eps, ts), rtol=1e-13, atol=0)
        assert np.all(op.tau[typ == 0] == 0) and np.all(op.ssa[typ == 0] == 0) and np.all(op.g[typ == 0] == 0)


def test_aerosol_optics_error_strings(oracle_lib):
    """Messages of mo_aerosol_optics_rrtmgp_merra.F90:297-357."""
    kd = syn.make_kdist("lw")
    lut = syn.make_aerosol_lut(kd)
    ctx = Context(oracle_lib, None)
    ctx.config_checks(True, True)
    ncol, nlay = 4, 5
    typ, size, mass, rh = _random_aerosols(ncol, nlay, lut)
    ao = AerosolOptics(ctx, lut)
    op = OpticalProps.like(ctx, "1scl", ncol, nlay, ao)
    bad = typ.copy(order="F"); bad[2, 2] = 9
    with pytest.raises(RuntimeError, match="aerosol type is out of bounds"):
        ao.aerosol_optics(bad, size, mass, rh, op)
    bad = size.copy(order="F"); bad[typ > 0] = 100.0
    with pytest.raises(RuntimeError, match="requested aerosol size is out of bounds"):
        ao.aerosol_optics(typ, bad, mass, rh, op)
    bad = rh.copy(order="F"); bad[typ > 0] = 1.5
    with pytest.raises(RuntimeError, match="relative humidity fraction is out of bounds"):
        ao.aerosol_optics(typ, size, mass, bad, op)
    with pytest.raises(RuntimeError, match="optical_props have wrong extents"):
        ao.aerosol_optics(typ, size, mass, rh, OpticalProps.like(ctx, "1scl", ncol + 1, nlay, ao))
    with pytest.raises(RuntimeError, match="must be requested by band not g-points"):
        ao.aerosol_optics(typ, size, mass, rh, OpticalProps.like(ctx, "1scl", ncol, nlay, syn_go(kd)))


class syn_go:
    """spectral description with g-points (stand-in for a gas-optics object in OpticalProps.like)."""

    def __init__(self, kd):
        self.band_lims_gpt, self.band_lims_wvn = kd.band_lims_gpt, kd.band_lims_wvn


def test_allsky_aerosols_oracle_changes_fluxes(oracle_lib):
    """Aerosols are present in odd columns only (rrtmgp_allsky.F90:717): those columns' fluxes change, the others don't."""
    kd_lw, kd_sw = syn.make_kdist("lw", ngpt=32), syn.make_kdist("sw", ngpt=28)
    ctx = Context(oracle_lib, None)
    a = AllSky(ctx, 6, 72, kd_lw, kd_sw, do_aerosols=True); a.step()
    b = AllSky(ctx, 6, 72, kd_lw, kd_sw, do_aerosols=False); b.step()
    fa, fb = a.fluxes_host(), b.fluxes_host()
    for k in fa:
        d = np.max(np.abs(fa[k] - fb[k]), axis=1)
        # 1-based even columns carry no aerosol: incrementing by tau = 0 only re-rounds ssa and g
        assert np.all(d[1::2] < 1e-10), k
        assert np.all(d[0::2] > 1e-6), k


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["1scl", "2str"])
def test_cuda_aerosol_optics_vs_oracle(oracle_lib, cuda_lib, kind):
    kd = syn.make_kdist("sw" if kind == "2str" else "lw")
    lut = syn.make_aerosol_lut(kd)
    ncol, nlay = 131, 33
    typ, size, mass, rh = _random_aerosols(ncol, nlay, lut)
    res = []
    for lib, dev in ((oracle_lib, None), (cuda_lib, "cuda:0")):
        ctx = Context(lib, dev)
        ao = AerosolOptics(ctx, lut)
        op = OpticalProps.like(ctx, kind, ncol, nlay, ao)
        ao.aerosol_optics(ctx.put(typ), ctx.put(size), ctx.put(mass), ctx.put(rh), op)
        res.append([ctx.get(x) for x in (op.tau, op.ssa, op.g) if x is not None])
    for c, g in zip(*res):
        np.testing.assert_allclose(g, c, rtol=1e-14, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("lw_2stream", [False, True, "quirk"])
@pytest.mark.parametrize("ncol,nlay", [(26, 72), (35, 60)])
def test_allsky_with_aerosols(oracle_lib, cuda_lib, ncol, nlay, lw_2stream):
    """BASELINE config 5 at test size: clouds + aerosols; LW either no-scattering or two-stream (g-point fluxes
    summed with rte_sum_broadband).  LW two-stream is compared in both level-source modes: per g-point (the CUDA
    library's default = the reference's accelerator kernels, accel/mo_rte_solver_kernels.F90:958-962; the oracle is
    switched to it) and "quirk" (the serial default kernels' sequence association, mo_rte_solver_kernels.F90:422 = the
    oracle's default; the CUDA library is switched to it)."""
    kd_lw, kd_sw = syn.make_kdist("lw"), syn.make_kdist("sw")
    runs = []
    per_gpt = 0 if lw_2stream == "quirk" else 1
    lw_2stream = bool(lw_2stream)
    for lib, dev in ((oracle_lib, None), (cuda_lib, "cuda:0")):
        lib.cdll.rrtmgpb_set_lw_2stream_lev_source_per_gpt(per_gpt)
        try:
            a = AllSky(Context(lib, dev), ncol, nlay, kd_lw, kd_sw, do_aerosols=True, lw_2stream=lw_2stream)
            a.step()
            lib.sync()
        finally:  # defaults: oracle = the serial reference kernels' behaviour, CUDA = per g-point
            lib.cdll.rrtmgpb_set_lw_2stream_lev_source_per_gpt(0 if dev is None else 1)
        runs.append(a)
    c, g = runs
    fc, fg = c.fluxes_host(), g.fluxes_host()
    for k in fc:
        assert np.max(np.abs(fg[k] - fc[k])) <= FLUX_ATOL, k
    if lw_2stream:
        x, y = g.ctx.get(g.lw.gpt_flux_up), c.ctx.get(c.lw.gpt_flux_up)
        assert np.max(np.abs(x - y)) <= FLUX_ATOL


@pytest.mark.parametrize("lw_2stream", [False, True])
def test_oracle_fused_with_aerosols_equals_reference_sequence(oracle_lib, lw_2stream):
    """gas_optics(..., increment_by=clouds, increment_by2=aerosols) on the oracle is literally the reference sequence
    (gas optics, clouds%increment, aerosols%increment): bit-identical fluxes."""
    kd_lw, kd_sw = syn.make_kdist("lw", ngpt=32), syn.make_kdist("sw", ngpt=28)
    ctx = Context(oracle_lib, None)
    res = []
    for fused in (True, False):
        a = AllSky(ctx, 9, 40, kd_lw, kd_sw, do_aerosols=True, lw_2stream=lw_2stream, fused=fused)
        a.step()
        res.append(a.fluxes_host())
    for k in res[0]:
        np.testing.assert_array_equal(res[0][k], res[1][k], err_msg=k)


@pytest.mark.gpu
def test_cuda_fused_and_sequence_with_aerosols(oracle_lib, cuda_lib):
    """CUDA: the fused pass with both increments and the kernel-by-kernel sequence both match the oracle."""
    kd_lw, kd_sw = syn.make_kdist("lw"), syn.make_kdist("sw")
    c = AllSky(Context(oracle_lib, None), 30, 72, kd_lw, kd_sw, do_aerosols=True, fused=False)
    c.step()
    fc = c.fluxes_host()
    for fused in (True, False):
        g = AllSky(Context(cuda_lib, "cuda:0"), 30, 72, kd_lw, kd_sw, do_aerosols=True, fused=fused)
        g.step()
        fg = g.fluxes_host()
        for k in fc:
            assert np.max(np.abs(fg[k] - fc[k])) <= FLUX_ATOL, (k, fused)
        for name, ga, ca in (("lw tau", g.lw.atmos.tau, c.lw.atmos.tau), ("sw tau", g.sw.atmos.tau, c.sw.atmos.tau),
                             ("sw ssa", g.sw.atmos.ssa, c.sw.atmos.ssa), ("sw g", g.sw.atmos.g, c.sw.atmos.g)):
            np.testing.assert_allclose(g.ctx.get(ga), c.ctx.get(ca), rtol=1e-12, atol=1e-300, err_msg=f"{name} fused={fused}")
