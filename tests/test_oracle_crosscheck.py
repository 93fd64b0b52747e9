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
