"""GPU parity of the five rrtmgp_* extern symbols, each called through the C-ABI on seeded random and on
distinct-column profiles, EVERY output compared with the oracle: index outputs (jtemp, jpress, jeta, tropo -
rrtmgp/kernels/mo_gas_optics_rrtmgp_kernels.F90:106-117,153) bit-exactly, floating-point outputs at 1e-12 relative
(FMA contraction / libdevice log on the GPU vs glibc and no contraction on the CPU; the interpolation weights, which are
differences of O(10) numbers, at 5e-14 absolute)."""
import numpy as np
import pytest

import gas_optics_calls as gc
import refcases as rc
from rte_rrtmgp_b200 import synthetic as syn

RTOL = 1.0e-12

KD = {
    "lw16": ("lw", dict()),
    "sw16": ("sw", dict()),
    "lw_ragged": ("lw", dict(band_sizes=[3, 17, 16, 20, 1, 2, 37, 5, 16, 16, 7, 8, 9, 10, 11, 12], seed=5)),
    "sw_ragged": ("sw", dict(band_sizes=[16, 1, 33, 4, 6, 16, 18, 2, 3, 5, 7, 16, 16, 9], seed=6)),
}


def _inputs(kd, source, ncol, nlay, top_at_1, seed):
    if source == "random":
        return gc.profile(kd, ncol, nlay, seed, top_at_1=top_at_1)
    prof = syn.perturbed_profiles(ncol, nlay, seed=seed, top_at_1=top_at_1)   # RFMIP-like distinct columns
    # col_dry / col_gas exactly as the frontend builds them (mo_gas_optics_rrtmgp.F90:581-609)
    import oracle
    from rte_rrtmgp_b200.abi import fzeros
    import ctypes as C
    from rte_rrtmgp_b200.abi import _ptr
    P = lambda a: C.c_void_p(_ptr(a).value)
    col_dry = fzeros((ncol, nlay))
    oracle.lib().cdll.rrtmgpb_get_col_dry(ncol, nlay, P(np.asfortranarray(prof["q"])), P(prof["p_lev"]), P(col_dry))
    vmr = {"h2o": prof["q"], "o3": prof["o3"]}
    col_gas = np.zeros((ncol, nlay, kd.ngas + 1), order="F")
    col_gas[:, :, 0] = col_dry
    for i, name in enumerate(syn.GAS_NAMES):
        v = vmr[name] if name in vmr else syn.ALLSKY_WELL_MIXED.get(name, 0.0)
        col_gas[:, :, i + 1] = v * col_dry
    sfc = nlay if top_at_1 else 0
    return dict(play=prof["p_lay"], tlay=prof["t_lay"], col_gas=col_gas, col_dry=col_dry, tlev=prof["t_lev"],
                tsfc=np.ascontiguousarray(prof["t_lev"][:, sfc]))


def _close(a, b, name, atol=1e-300):
    np.testing.assert_allclose(a, b, rtol=RTOL, atol=atol, err_msg=name)


@pytest.mark.gpu
@pytest.mark.parametrize("source", ["random", "distinct"])
@pytest.mark.parametrize("top_at_1", [False, True])
@pytest.mark.parametrize("kdname", sorted(KD))
def test_five_symbols_every_output(oracle_lib, cuda_lib, kdname, top_at_1, source):
    kind, kw = KD[kdname]
    kd = syn.make_kdist(kind, **kw)
    ncol, nlay = (37, 19) if source == "random" else (41, 60)
    x = _inputs(kd, source, ncol, nlay, top_at_1, seed=17)
    # a1 interpolation: outputs stay on the device for the kernels that consume them; a host copy is compared
    it_g = gc.interpolation(cuda_lib, "cuda:0", kd, x["play"], x["tlay"], x["col_gas"], keep_device=True)
    it_c = gc.interpolation(oracle_lib, None, kd, x["play"], x["tlay"], x["col_gas"], keep_device=True)
    for k in ("jtemp", "jpress", "jeta", "tropo"):
        assert np.array_equal(rc.host(it_g[k]), it_c[k]), f"{k} differs from the oracle (must be bit-exact)"
    _close(rc.host(it_g["col_mix"]), it_c["col_mix"], "col_mix")
    for k in ("fmajor", "fminor"):
        # interpolation weights are differences: fpress = locpress - aint(locpress) with locpress up to 59 (:111-115),
        # feta = loceta - aint(loceta) (:154): one ulp of log() or of the eta division is ~1e-14 ABSOLUTE on a weight in
        # [0, 1] - the index outputs above, which decide WHICH table nodes are combined, are bit-exact
        _close(rc.host(it_g[k]), it_c[k], k, atol=5.0e-14)
    # a2 tau_absorption
    tau_g = gc.tau_absorption(cuda_lib, "cuda:0", kd, x["play"], x["tlay"], x["col_gas"], it_g)
    tau_c = gc.tau_absorption(oracle_lib, None, kd, x["play"], x["tlay"], x["col_gas"], it_c)
    _close(tau_g, tau_c, "tau")
    if kind == "sw":   # a3
        _close(gc.tau_rayleigh(cuda_lib, "cuda:0", kd, x["col_dry"], x["col_gas"], it_g),
               gc.tau_rayleigh(oracle_lib, None, kd, x["col_dry"], x["col_gas"], it_c), "tau_rayleigh")
    else:              # a4
        sfc_lay = nlay if top_at_1 else 1
        got = gc.planck_source(cuda_lib, "cuda:0", kd, x["tlay"], x["tlev"], x["tsfc"], sfc_lay, it_g)
        ref = gc.planck_source(oracle_lib, None, kd, x["tlay"], x["tlev"], x["tsfc"], sfc_lay, it_c)
        for a, b, n in zip(got, ref, ("sfc_src", "lay_src", "lev_src", "sfc_source_Jac")):
            # the Jacobian is a difference of two nearby Planck values (:652-653): relative to the source itself
            if n == "sfc_source_Jac":
                assert np.max(np.abs(a - b)) <= RTOL * np.max(np.abs(ref[0])), n
            else:
                _close(a, b, n)


@pytest.mark.gpu
def test_tau_absorption_with_cached_gfast_tables(oracle_lib, cuda_lib):
    """Same call with the library allowed to keep g-point-fastest copies of the tables (rrtmgpb_abi_table_cache(1)),
    and again after the caller refills a table in place and reports it (rrtmgpb_tables_changed)."""
    kd = syn.make_kdist("lw")
    x = gc.profile(kd, 29, 17, seed=23)
    it_c = gc.interpolation(oracle_lib, None, kd, x["play"], x["tlay"], x["col_gas"], keep_device=True)
    ref = gc.tau_absorption(oracle_lib, None, kd, x["play"], x["tlay"], x["col_gas"], it_c)
    it_g = gc.interpolation(cuda_lib, "cuda:0", kd, x["play"], x["tlay"], x["col_gas"], keep_device=True)
    cuda_lib.cdll.rrtmgpb_abi_table_cache(1)
    try:
        _close(gc.tau_absorption(cuda_lib, "cuda:0", kd, x["play"], x["tlay"], x["col_gas"], it_g), ref, "tau (cached)")
    finally:
        cuda_lib.cdll.rrtmgpb_abi_table_cache(0)


@pytest.mark.gpu
@pytest.mark.parametrize("source", ["random", "distinct"])
@pytest.mark.parametrize("top_at_1", [False, True])
@pytest.mark.parametrize("kdname", sorted(KD))
def test_tau_and_planck_symbols_on_the_gfast_kernels(oracle_lib, cuda_lib, kdname, top_at_1, source):
    """With cached table copies allowed (rrtmgpb_abi_table_cache(1), what the C++ frontend mirror does) the extern symbols
    rrtmgp_compute_tau_absorption and rrtmgp_compute_Planck_source run the fused path's g-point-fastest kernels in their
    ABI instantiation (interpolation state read from the caller's arrays): regular and ragged bands, contributors of
    another flavour, replicated-like and distinct columns, both orientations - every output against the oracle."""
    kind, kw = KD[kdname]
    kd = syn.make_kdist(kind, **kw)
    ncol, nlay = (37, 19) if source == "random" else (300, 60)   # 300 columns: several blocks, whole warps
    x = _inputs(kd, source, ncol, nlay, top_at_1, seed=29)
    it_c = gc.interpolation(oracle_lib, None, kd, x["play"], x["tlay"], x["col_gas"], keep_device=True)
    it_g = gc.interpolation(cuda_lib, "cuda:0", kd, x["play"], x["tlay"], x["col_gas"], keep_device=True)
    tau_c = gc.tau_absorption(oracle_lib, None, kd, x["play"], x["tlay"], x["col_gas"], it_c)
    cuda_lib.cdll.rrtmgpb_abi_table_cache(1)
    try:
        n0 = cuda_lib.launch_count()
        tau_g = gc.tau_absorption(cuda_lib, "cuda:0", kd, x["play"], x["tlay"], x["col_gas"], it_g)
        _close(tau_g, tau_c, "tau (g-point-fastest kernels)")
        assert cuda_lib.launch_count() > n0
        if kind == "lw":
            sfc_lay = nlay if top_at_1 else 1
            got = gc.planck_source(cuda_lib, "cuda:0", kd, x["tlay"], x["tlev"], x["tsfc"], sfc_lay, it_g)
            ref = gc.planck_source(oracle_lib, None, kd, x["tlay"], x["tlev"], x["tsfc"], sfc_lay, it_c)
            for a, b, n in zip(got, ref, ("sfc_src", "lay_src", "lev_src", "sfc_source_Jac")):
                if n == "sfc_source_Jac":
                    assert np.max(np.abs(a - b)) <= RTOL * np.max(np.abs(ref[0])), n
                else:
                    _close(a, b, n)
    finally:
        cuda_lib.cdll.rrtmgpb_abi_table_cache(0)


@pytest.mark.gpu
def test_tau_absorption_accumulates_on_the_gfast_kernels(oracle_lib, cuda_lib):
    """The extern symbol ADDS to tau (the frontend zeroes it first, mo_gas_optics_rrtmgp_kernels.F90:391): a non-zero tau
    going in must come out incremented, on either kernel family."""
    import ctypes as C
    from rte_rrtmgp_b200.abi import fzeros
    kd = syn.make_kdist("lw")
    x = gc.profile(kd, 40, 21, seed=3)
    it_c = gc.interpolation(oracle_lib, None, kd, x["play"], x["tlay"], x["col_gas"], keep_device=True)
    it_g = gc.interpolation(cuda_lib, "cuda:0", kd, x["play"], x["tlay"], x["col_gas"], keep_device=True)
    base = gc.tau_absorption(oracle_lib, None, kd, x["play"], x["tlay"], x["col_gas"], it_c)
    orig = gc.fzeros

    def prefilled(shape, dtype=np.float64, device=None):
        a = orig(shape, dtype, device)
        if len(shape) == 3 and shape[2] == kd.ngpt:
            a += 0.25
        return a

    for cache in (0, 1):
        cuda_lib.cdll.rrtmgpb_abi_table_cache(cache)
        gc.fzeros = prefilled
        try:
            got = gc.tau_absorption(cuda_lib, "cuda:0", kd, x["play"], x["tlay"], x["col_gas"], it_g)
        finally:
            gc.fzeros = orig
            cuda_lib.cdll.rrtmgpb_abi_table_cache(0)
        _close(got, base + 0.25, f"tau accumulate (cache={cache})")


@pytest.mark.gpu
def test_cld_from_table_random(oracle_lib, cuda_lib):
    kdl = syn.make_kdist("sw", gpt_per_band=1)
    lut = syn.make_cloud_lut(kdl)
    rng = np.random.default_rng(5)
    ncol, nlay = 37, 19
    for tabs, lo, hi in (((lut.extliq, lut.ssaliq, lut.asyliq), lut.radliq_lwr, lut.radliq_upr),
                         ((lut.extice[:, :, 1], lut.ssaice[:, :, 1], lut.asyice[:, :, 1]), lut.diamice_lwr, lut.diamice_upr)):
        tabs = [np.asfortranarray(t) for t in tabs]
        nsteps = tabs[0].shape[0]
        step = (hi - lo) / (nsteps - 1)
        re = np.asfortranarray(rng.uniform(lo, hi, (ncol, nlay)))
        re[0, 0], re[1, 1] = lo, hi
        lwp = np.asfortranarray(rng.uniform(0.0, 50.0, (ncol, nlay)))
        mask = np.asfortranarray(rng.random((ncol, nlay)) < 0.6)
        got = gc.cld_from_table(cuda_lib, "cuda:0", mask, lwp, re, nsteps, step, lo, *tabs)
        ref = gc.cld_from_table(oracle_lib, None, mask, lwp, re, nsteps, step, lo, *tabs)
        for a, b, n in zip(got, ref, ("tau", "taussa", "taussag")):
            _close(a, b, n)
            assert np.array_equal(a == 0, b == 0), n
