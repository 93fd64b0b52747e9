"""Pins the C oracle's spectral kernels (the ones whose reference golden vectors live in the un-vendored rrtmgp-data
tarball) against a SECOND restatement written independently in vectorised numpy from the Fortran
(tests/numpy_gas_optics.py).  The oracle's parity build (-O2 -ffp-contract=off) must agree BIT FOR BIT: index outputs
with array_equal, floating-point outputs with array_equal too (same IEEE operations in the same association).
Runs on CPU (oracle only)."""
import numpy as np
import pytest

import gas_optics_calls as gc
import numpy_gas_optics as npg
from rte_rrtmgp_b200 import synthetic as syn

CASES = [  # kind, kwargs of make_kdist, ncol, nlay, top_at_1, seed
    ("lw", dict(gpt_per_band=4, seed=3), 9, 14, False, 1),
    ("lw", dict(band_sizes=[3, 17, 16, 20, 1, 2, 37, 5, 16, 16, 7, 8, 9, 10, 11, 12], seed=5), 7, 11, True, 2),
    ("sw", dict(gpt_per_band=3, seed=4), 8, 12, True, 3),
    ("sw", dict(band_sizes=[16, 1, 33, 4, 6, 16, 18, 2, 3, 5, 7, 16, 16, 9], seed=6), 6, 13, False, 4),
    ("lw", dict(ngpt=256), 5, 9, False, 5),
]


@pytest.mark.parametrize("kind,kw,ncol,nlay,top_at_1,seed", CASES)
def test_numpy_transcription_equals_c_oracle(oracle_lib, kind, kw, ncol, nlay, top_at_1, seed):
    kd = syn.make_kdist(kind, **kw)
    x = gc.profile(kd, ncol, nlay, seed, top_at_1=top_at_1)
    it_c = gc.interpolation(oracle_lib, None, kd, x["play"], x["tlay"], x["col_gas"], keep_device=True)
    it_n = npg.interpolation(kd, x["play"], x["tlay"], x["col_gas"])
    for k in ("jtemp", "jpress", "jeta", "tropo"):
        assert np.array_equal(np.asarray(it_c[k]).astype(np.int64), np.asarray(it_n[k]).astype(np.int64)), k
    for k in ("col_mix", "fmajor", "fminor"):
        assert np.array_equal(it_c[k], it_n[k]), k
    tau_c = gc.tau_absorption(oracle_lib, None, kd, x["play"], x["tlay"], x["col_gas"], it_c)
    tau_n = npg.compute_tau_absorption(kd, it_n, x["play"], x["tlay"], x["col_gas"], np.zeros_like(tau_c))
    assert tau_c.max() > 0
    assert np.array_equal(tau_c, tau_n)
    if kind == "sw":
        r_c = gc.tau_rayleigh(oracle_lib, None, kd, x["col_dry"], x["col_gas"], it_c)
        r_n = npg.compute_tau_rayleigh(kd, it_n, x["col_dry"], x["col_gas"])
        assert np.array_equal(r_c, r_n)
    else:
        sfc_lay = nlay if top_at_1 else 1
        out_c = gc.planck_source(oracle_lib, None, kd, x["tlay"], x["tlev"], x["tsfc"], sfc_lay, it_c)
        out_n = npg.compute_planck_source(kd, it_n, x["tlay"], x["tlev"], x["tsfc"], sfc_lay)
        for a, b, name in zip(out_c, out_n, ("sfc_src", "lay_src", "lev_src", "sfc_source_Jac")):
            assert np.array_equal(a, b), name


def test_layer_limits_follow_minloc_maxloc_for_non_monotonic_pressure(oracle_lib):
    """:274-285: the lower/upper minor-gas layer ranges come from minloc/maxloc of play under the tropo mask - for a
    pressure profile that is NOT monotonic this differs from the per-layer predicate, and the oracle must follow the
    Fortran (the numpy transcription implements minloc/maxloc literally)."""
    kd = syn.make_kdist("lw", gpt_per_band=2, seed=9)
    x = gc.profile(kd, 6, 12, seed=11, monotonic=False)
    x["play"][0, 0], x["play"][0, -1] = 9.0e4, 20.0  # orientation flag comes from column 1 (:274)
    it_c = gc.interpolation(oracle_lib, None, kd, x["play"], x["tlay"], x["col_gas"], keep_device=True)
    it_n = npg.interpolation(kd, x["play"], x["tlay"], x["col_gas"])
    tau_c = gc.tau_absorption(oracle_lib, None, kd, x["play"], x["tlay"], x["col_gas"], it_c)
    tau_n = npg.compute_tau_absorption(kd, it_n, x["play"], x["tlay"], x["col_gas"], np.zeros_like(tau_c))
    assert np.array_equal(tau_c, tau_n)


def test_cloud_lut_numpy_equals_c_oracle(oracle_lib):
    kdl = syn.make_kdist("sw", gpt_per_band=1)
    lut = syn.make_cloud_lut(kdl)
    rng = np.random.default_rng(3)
    ncol, nlay = 11, 7
    nsteps = lut.extliq.shape[0]
    step = (lut.radliq_upr - lut.radliq_lwr) / (nsteps - 1)
    re = np.asfortranarray(rng.uniform(lut.radliq_lwr, lut.radliq_upr, (ncol, nlay)))
    re[0, 0], re[1, 1] = lut.radliq_lwr, lut.radliq_upr  # both ends of the table (index clamp :46)
    lwp = np.asfortranarray(rng.uniform(0.0, 50.0, (ncol, nlay)))
    mask = np.asfortranarray(rng.random((ncol, nlay)) < 0.7)
    got = gc.cld_from_table(oracle_lib, None, mask, lwp, re, nsteps, step, lut.radliq_lwr, lut.extliq, lut.ssaliq, lut.asyliq)
    ref = npg.compute_cld_from_table(mask, lwp, re, nsteps, step, lut.radliq_lwr, lut.extliq, lut.ssaliq, lut.asyliq)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)


# ---------------------------------------------------------------------------------------------------------------------
# The two-stream flux solvers: the oracle (oracle/rte_solver_ref.c) against tests/numpy_solvers.py, an independent numpy
# transcription of mo_rte_solver_kernels.F90 (lw_solver_2stream, sw_solver_2stream, adding) - bit for bit.
# ---------------------------------------------------------------------------------------------------------------------
import numpy_solvers as nps  # noqa: E402
from rte_rrtmgp_b200.abi import fzeros  # noqa: E402


def _sw_case(ncol, nlay, ngpt, seed):
    rng = np.random.default_rng(seed)
    mu0 = np.asfortranarray(np.repeat(rng.uniform(-0.2, 1.0, ncol)[:, None], nlay, axis=1))
    mu0[1] = rng.uniform(0.05, 1.0, nlay)                  # a column whose mu0 varies with height
    mu0[2, nlay // 2:] = -0.1                              # the sun sets part of the way down the column
    tau = np.asfortranarray(10.0 ** rng.uniform(-7, 2.0, (ncol, nlay, ngpt)))
    tau[0, 0, 0] = 0.0
    return dict(tau=tau, ssa=np.asfortranarray(rng.uniform(0, 1.0, (ncol, nlay, ngpt))),
                g=np.asfortranarray(rng.uniform(-0.5, 0.95, (ncol, nlay, ngpt))), mu0=mu0,
                adir=np.asfortranarray(rng.uniform(0, 0.7, (ncol, ngpt))), adif=np.asfortranarray(rng.uniform(0, 0.7, (ncol, ngpt))),
                inc=np.asfortranarray(10.0 * rng.random((ncol, ngpt))), dif=np.asfortranarray(2.0 * rng.random((ncol, ngpt))))


@pytest.mark.parametrize("top_at_1", [True, False])
@pytest.mark.parametrize("bb,bc", [(True, False), (True, True), (False, False), (False, True)])
def test_numpy_sw_solver_2stream_equals_c_oracle(oracle_lib, top_at_1, bb, bc):
    x = _sw_case(11, 23, 5, seed=41)
    ncol, nlay, ngpt = x["tau"].shape
    gup, gdn, gdr = (fzeros((ncol, nlay + 1, ngpt)) for _ in range(3))
    bup, bdn, bdr = (fzeros((ncol, nlay + 1)) for _ in range(3))
    oracle_lib.rte_sw_solver_2stream(ncol, nlay, ngpt, top_at_1, x["tau"], x["ssa"], x["g"], x["mu0"], x["adir"], x["adif"], x["inc"],
                                     gup, gdn, gdr, bc, x["dif"], bb, bup, bdn, bdr)
    got = (bup, bdn, bdr) if bb else (gup, gdn, gdr)
    ref = nps.sw_solver_2stream(top_at_1, x["tau"], x["ssa"], x["g"], x["mu0"], x["adir"], x["adif"], x["inc"], bc, x["dif"], bb)
    for a, b, n in zip(got, ref, ("up", "dn", "dir")):
        assert np.max(np.abs(b)) > 0
        assert np.array_equal(a, b), f"{n}: max diff {np.max(np.abs(a - b)):.3e}"


@pytest.mark.parametrize("top_at_1", [True, False])
@pytest.mark.parametrize("lev_per_gpt", [False, True])
def test_numpy_lw_solver_2stream_equals_c_oracle(oracle_lib, top_at_1, lev_per_gpt):
    """Both readings of lev_source: the serial kernel as written (every g-point sees g-point 1's level source,
    mo_rte_solver_kernels.F90:422) and per g-point (the accelerator kernels)."""
    rng = np.random.default_rng(43)
    ncol, nlay, ngpt = 9, 21, 4
    tau = np.asfortranarray(10.0 ** rng.uniform(-9.5, 1.5, (ncol, nlay, ngpt)))   # some below the 1e-8 source cut (:947)
    ssa = np.asfortranarray(rng.uniform(0, 1.0, (ncol, nlay, ngpt)))
    g = np.asfortranarray(rng.uniform(-0.4, 0.9, (ncol, nlay, ngpt)))
    lev = np.asfortranarray(50.0 + 100.0 * rng.random((ncol, nlay + 1, ngpt)))
    lay = np.asfortranarray(0.5 * (lev[:, 1:] + lev[:, :-1]))
    emis = np.asfortranarray(0.8 + 0.2 * rng.random((ncol, ngpt)))
    sfc = np.asfortranarray(100.0 + 50.0 * rng.random((ncol, ngpt)))
    inc = np.asfortranarray(5.0 * rng.random((ncol, ngpt)))
    gup, gdn = (fzeros((ncol, nlay + 1, ngpt)) for _ in range(2))
    oracle_lib.cdll.rrtmgpb_set_lw_2stream_lev_source_per_gpt(1 if lev_per_gpt else 0)
    try:
        oracle_lib.rte_lw_solver_2stream(ncol, nlay, ngpt, top_at_1, tau, ssa, g, lay, lev, emis, sfc, inc, gup, gdn)
    finally:
        oracle_lib.cdll.rrtmgpb_set_lw_2stream_lev_source_per_gpt(0)
    fu, fd = nps.lw_solver_2stream(top_at_1, tau, ssa, g, lay, lev, emis, sfc, inc, lev_per_gpt)
    assert np.array_equal(gup, fu), f"up: max diff {np.max(np.abs(gup - fu)):.3e}"
    assert np.array_equal(gdn, fd), f"dn: max diff {np.max(np.abs(gdn - fd)):.3e}"


@pytest.mark.parametrize("top_at_1", [True, False])
@pytest.mark.parametrize("nmus,bb,jac,resc", [(1, True, False, False), (1, False, False, False), (3, True, True, False),
                                              (2, False, True, False), (1, True, True, True), (2, False, False, True)])
def test_numpy_lw_solver_noscat_equals_c_oracle(oracle_lib, top_at_1, nmus, bb, jac, resc):
    """Every flag combination of rte_lw_solver_noscat: broadband / g-point fluxes, 1-3 quadrature angles, surface-temperature
    Jacobian, Tang rescaling (with its orientation-asymmetric second sweep, mo_rte_solver_kernels.F90:801-804 vs :835-838)."""
    rng = np.random.default_rng(47)
    ncol, nlay, ngpt = 8, 19, 4
    tau = np.asfortranarray(10.0 ** rng.uniform(-7, 1.5, (ncol, nlay, ngpt)))     # both branches of the weighting factor (:652-656)
    lev = np.asfortranarray(50.0 + 100.0 * rng.random((ncol, nlay + 1, ngpt)))
    lay = np.asfortranarray(0.5 * (lev[:, 1:] + lev[:, :-1]) + rng.uniform(-1, 1, (ncol, nlay, ngpt)))
    emis = np.asfortranarray(0.8 + 0.2 * rng.random((ncol, ngpt)))
    sfc = np.asfortranarray(100.0 + 50.0 * rng.random((ncol, ngpt)))
    sjac = np.asfortranarray(rng.random((ncol, ngpt)))
    inc = np.asfortranarray(5.0 * rng.random((ncol, ngpt)))
    ssa = np.asfortranarray(rng.uniform(0, 0.9, (ncol, nlay, ngpt)))
    g = np.asfortranarray(rng.uniform(-0.2, 0.9, (ncol, nlay, ngpt)))
    Ds = np.asfortranarray(np.stack([np.full((ncol, ngpt), 1.0 / m) for m in (0.61, 0.25, 0.79)[:nmus]], axis=2))
    wts = np.array([1.0, 0.23, 0.77][:nmus]) if nmus > 1 else np.array([1.0])
    gup, gdn = fzeros((ncol, nlay + 1, ngpt)), fzeros((ncol, nlay + 1, ngpt))
    bup, bdn, fj = (fzeros((ncol, nlay + 1)) for _ in range(3))
    oracle_lib.rte_lw_solver_noscat(ncol, nlay, ngpt, top_at_1, nmus, Ds, wts, tau, lay, lev, emis, sfc, inc, gup, gdn, bb, bup, bdn,
                                    jac, sjac, fj, resc, ssa, g)
    ref = nps.lw_solver_noscat(top_at_1, Ds, wts, tau, lay, lev, emis, sfc, inc, bb, jac, sjac, resc, ssa, g)
    pairs = [(bup, ref[2], "bb up"), (bdn, ref[3], "bb dn")] if bb else [(gup, ref[0], "up"), (gdn, ref[1], "dn")]
    if jac:
        pairs.append((fj, ref[4], "jacobian"))
    for a, b, n in pairs:
        assert np.max(np.abs(b)) > 0, n
        assert np.array_equal(a, b), f"{n}: max diff {np.max(np.abs(a - b)):.3e}"


# ---------------------------------------------------------------------------------------------------------------------
# Optical-properties arithmetic: oracle/rte_optical_props_ref.c against tests/numpy_optical_props.py - bit for bit.
# ---------------------------------------------------------------------------------------------------------------------
import numpy_optical_props as npo  # noqa: E402
import test_optical_props_parity as opp  # noqa: E402  (its operand generator and its by-reference caller)


@pytest.mark.parametrize("bybnd", [False, True])
@pytest.mark.parametrize("k2", opp.KINDS)
@pytest.mark.parametrize("k1", opp.KINDS)
@pytest.mark.parametrize("nmom1,nmom2", [(4, 2), (2, 5), (3, 3)])
def test_numpy_increments_equal_c_oracle(oracle_lib, k1, k2, bybnd, nmom1, nmom2):
    if "nstream" not in (k1, k2) and (nmom1, nmom2) != (3, 3):
        pytest.skip("moment counts only matter for n-stream operands")
    rng = np.random.default_rng(1000 + 100 * opp.KINDS.index(k1) + 10 * opp.KINDS.index(k2) + bybnd)
    op1 = opp._props(rng, k1, opp.NGPT, nmom1)
    op2 = opp._props(rng, k2, opp.NBND if bybnd else opp.NGPT, nmom2)
    got = opp._call(oracle_lib, None, k1, k2, bybnd, nmom1, nmom2, op1, op2)
    ref = npo.increment_bybnd(k1, k2, op1, op2, opp.LIMS) if bybnd else npo.increment(k1, k2, op1, op2)
    for a, b, n in zip(got, ref, ("tau", "ssa", "g/p")):
        assert np.array_equal(a, b), f"{k1} += {k2} bybnd={bybnd}: {n} (max diff {np.max(np.abs(a - b)):.3e})"


def test_numpy_delta_scaling_equals_c_oracle(oracle_lib):
    rng = np.random.default_rng(7)
    tau, ssa, g = opp._props(rng, "2stream", opp.NGPT, 0)
    f = np.asfortranarray(rng.uniform(0.0, 0.9, tau.shape))
    f[rng.random(f.shape) < 0.05] = 1.0   # (1 - f) -> 0: the max(eps, .) guard (:69)
    t1, s1, g1 = (a.copy(order="F") for a in (tau, ssa, g))
    oracle_lib.rte_delta_scale_2str_k(opp.NCOL, opp.NLAY, opp.NGPT, t1, s1, g1)
    t2, s2, g2 = (a.copy(order="F") for a in (tau, ssa, g))
    oracle_lib.rte_delta_scale_2str_f_k(opp.NCOL, opp.NLAY, opp.NGPT, t2, s2, g2, f)
    for a, b, n in zip((t1, s1, g1), npo.delta_scale_2str(tau, ssa, g), ("tau", "ssa", "g")):
        assert np.array_equal(a, b), n
    for a, b, n in zip((t2, s2, g2), npo.delta_scale_2str_f(tau, ssa, g, f), ("tau_f", "ssa_f", "g_f")):
        assert np.array_equal(a, b), n


# ---------------------------------------------------------------------------------------------------------------------
# Frontend-resident loops (SURVEY 8a'): oracle/glue_ref.c, oracle/rte_misc_ref.c against tests/numpy_glue.py - bit for bit.
# ---------------------------------------------------------------------------------------------------------------------
import ctypes as C  # noqa: E402

import numpy_glue as npgl  # noqa: E402
from rte_rrtmgp_b200.abi import _ptr  # noqa: E402


def _P(a):
    return C.c_void_p(_ptr(a).value)


def test_numpy_glue_equals_c_oracle(oracle_lib):
    rng = np.random.default_rng(53)
    ncol, nlay, ngpt = 13, 17, 6
    f = lambda *s: np.asfortranarray(rng.random(s))
    c = oracle_lib.cdll
    # col_dry
    plev = np.asfortranarray(np.sort(rng.uniform(10.0, 1.0e5, (ncol, nlay + 1)), axis=1))
    q = np.asfortranarray(10.0 ** rng.uniform(-6, -1.7, (ncol, nlay)))
    col_dry = fzeros((ncol, nlay))
    c.rrtmgpb_get_col_dry(ncol, nlay, _P(q), _P(plev), _P(col_dry))
    assert np.array_equal(col_dry, npgl.get_layer_number(q, plev))
    # level temperatures
    play = np.asfortranarray(0.5 * (plev[:, 1:] + plev[:, :-1]))
    tlay = np.asfortranarray(rng.uniform(180.0, 320.0, (ncol, nlay)))
    tlev = fzeros((ncol, nlay + 1))
    c.rrtmgpb_interpolate_tlev(ncol, nlay, _P(play), _P(plev), _P(tlay), _P(tlev))
    assert np.array_equal(tlev, npgl.interpolate_tlev(play, plev, tlay))
    # absorption + Rayleigh
    ta = np.asfortranarray(10.0 ** rng.uniform(-9, 1, (ncol, nlay, ngpt)))
    tr = np.asfortranarray(10.0 ** rng.uniform(-9, 0, (ncol, nlay, ngpt)))
    ta[0, 0, :], tr[0, 0, :] = 0.0, 0.0          # t = 0: the ssa = 0 branch (:1995-1997)
    for kind in (1, 2):
        tau, ssa, g = (fzeros((ncol, nlay, ngpt)) for _ in range(3))
        g += 7.0
        c.rrtmgpb_combine_abs_and_rayleigh(ncol, nlay, ngpt, kind, _P(ta), _P(tr), _P(tau), _P(ssa), _P(g))
        rt, rs, rg = npgl.combine_abs_and_rayleigh(ta, tr, kind == 2)
        assert np.array_equal(tau, rt)
        if kind == 2:
            assert np.array_equal(ssa, rs) and np.array_equal(g, rg)
    # liquid + ice cloud properties
    lt, it = f(ncol, nlay, ngpt) * 5, f(ncol, nlay, ngpt) * 3
    lt[1, 1, :], it[1, 1, :] = 0.0, 0.0          # cloud-free cell: max(epsilon, .) guards (:417-418)
    lts, its = np.asfortranarray(lt * rng.random(lt.shape)), np.asfortranarray(it * rng.random(it.shape))
    ltsg, itsg = np.asfortranarray(lts * rng.uniform(-0.5, 0.9, lt.shape)), np.asfortranarray(its * rng.uniform(-0.5, 0.9, it.shape))
    for kind in (1, 2):
        tau, ssa, g = (fzeros((ncol, nlay, ngpt)) for _ in range(3))
        c.rrtmgpb_cloud_combine(ncol, nlay, ngpt, kind, _P(lt), _P(lts), _P(ltsg), _P(it), _P(its), _P(itsg), _P(tau), _P(ssa), _P(g))
        rt, rs, rg = npgl.cloud_combine(lt, lts, ltsg, it, its, itsg, kind == 2)
        assert np.array_equal(tau, rt)
        if kind == 2:
            assert np.array_equal(ssa, rs) and np.array_equal(g, rg)
    # broadband reductions
    fdn, fup = np.asfortranarray(300 * rng.random((ncol, nlay + 1, ngpt))), np.asfortranarray(300 * rng.random((ncol, nlay + 1, ngpt)))
    bb, net = fzeros((ncol, nlay + 1)), fzeros((ncol, nlay + 1))
    oracle_lib.rte_sum_broadband(ncol, nlay + 1, ngpt, fdn, bb)
    oracle_lib.rte_net_broadband_full(ncol, nlay + 1, ngpt, fdn, fup, net)
    assert np.array_equal(bb, npgl.sum_broadband(fdn))
    assert np.array_equal(net, npgl.net_broadband_full(fdn, fup))
