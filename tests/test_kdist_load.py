"""k-distribution ingestion: the C++ restatement of ty_gas_optics_rrtmgp%load / init_abs_coeffs
(rrtmgp/frontend/mo_gas_optics_rrtmgp.F90:938-1381, 1568-1946) applied to a synthetic ON-DISK variable set
(mo_optics_utils_rrtmgp.F90:102-183: 19 absorbers, string-valued minor-gas tables, (gpt, eta, p, T) coefficient arrays).
The raw set is built FROM a kernel-layout KDist, so load(raw, all gases) must give that KDist back exactly; with a
minor-only gas withheld the expected tables follow from filtering the contributors by hand; with a key species withheld
the reference's error string must come back.  The loader is host-only logic: the same source runs in the oracle build."""
import numpy as np
import pytest

from rte_rrtmgp_b200 import synthetic as syn
from rte_rrtmgp_b200.frontend import load_kdist_raw

FIELDS = ["flavor", "gpoint_flavor", "band_lims_gpt", "band_lims_wvn", "gpoint_bands", "press_ref_log", "temp_ref", "vmr_ref",
          "kmajor", "kminor_lower", "kminor_upper", "minor_limits_gpt_lower", "minor_limits_gpt_upper",
          "minor_scales_with_density_lower", "minor_scales_with_density_upper", "scale_by_complement_lower",
          "scale_by_complement_upper", "idx_minor_lower", "idx_minor_upper", "kminor_start_lower", "kminor_start_upper"]
SCALARS = ["ngas", "nflav", "neta", "npres", "ntemp", "nbnd", "ngpt", "idx_h2o", "temp_ref_min", "temp_ref_max", "temp_ref_delta",
           "press_ref_min", "press_ref_max"]


def _same_tables(got, want):
    for n in SCALARS:
        assert getattr(got, n) == getattr(want, n), n
    assert got.extra["nminorlower"] == want.extra["nminorlower"] and got.extra["nminorupper"] == want.extra["nminorupper"]
    for n in FIELDS:
        a, b = np.asarray(getattr(got, n)), np.asarray(getattr(want, n))
        assert a.shape == b.shape and np.array_equal(a, b), n
    for lu in ("lower", "upper"):  # "no scaling gas": 0 in the synthetic tables, -1 from string_loc_in_array (:1672); the kernels test > 0
        a, b = getattr(got, f"idx_minor_scaling_{lu}"), getattr(want, f"idx_minor_scaling_{lu}")
        assert np.array_equal(np.maximum(a, 0), np.maximum(b, 0)), lu
    # derived constants: computed from press_ref / temp_ref as :1320-1360 do
    np.testing.assert_allclose(got.press_ref_log_delta, want.press_ref_log_delta, rtol=1e-15)
    np.testing.assert_allclose(got.press_ref_trop_log, want.press_ref_trop_log, rtol=1e-15)


@pytest.mark.parametrize("kind,kw", [("lw", {}), ("sw", {}), ("lw", dict(band_sizes=[3, 17, 16, 20, 1, 2, 37, 5, 16, 16, 7, 8, 9, 10, 11, 12], seed=5)),
                                     ("sw", dict(ngpt=112))])
def test_load_raw_reproduces_kernel_tables(oracle_lib, kind, kw):
    kd = syn.make_kdist(kind, **kw)
    raw = syn.make_kdist_raw(kd)
    assert len(raw["gas_names"]) == 19 and len(raw["minor_gases_lower"]) > kd.extra["nminorlower"]
    avail = [g.upper() + " " for g in reversed(syn.GAS_NAMES)] + ["cfc99"]   # host order / case / padding / unknown gases do not matter
    got = load_kdist_raw(oracle_lib, raw, avail)
    assert [g.strip().lower() for g in got.gas_names] == syn.GAS_NAMES
    _same_tables(got, kd)
    if kind == "lw":
        assert np.array_equal(got.planck_frac, kd.planck_frac) and np.array_equal(got.totplnk, kd.totplnk)
        assert got.totplnk_delta == (kd.temp_ref_max - kd.temp_ref_min) / (kd.totplnk.shape[0] - 1)   # :1029
        assert np.array_equal(got.extra["optimal_angle_fit"], kd.extra["optimal_angle_fit"])
    else:
        assert np.array_equal(got.krayl, kd.krayl)
        np.testing.assert_allclose(got.solar_source, kd.solar_source, rtol=1e-14)   # quiet + facular + sunspot terms (:788-792)
    # is_key (:1364-1372): exactly the gases that appear in a flavour
    want_key = [i + 1 in set(kd.flavor.ravel().tolist()) for i in range(kd.ngas)]
    assert got.extra["is_key"] == want_key


def test_withholding_a_minor_only_gas_removes_its_contributors(oracle_lib):
    kd = syn.make_kdist("lw")
    raw = syn.make_kdist_raw(kd)
    key = set(kd.flavor.ravel().tolist())
    minor_only = [g for i, g in enumerate(syn.GAS_NAMES) if i + 1 not in key and
                  (np.any(kd.idx_minor_lower == i + 1) or np.any(kd.idx_minor_upper == i + 1))]
    assert minor_only, "the synthetic k-distribution has no minor-only gas"
    drop = minor_only[0]
    avail = [g for g in syn.GAS_NAMES if g != drop]
    got = load_kdist_raw(oracle_lib, raw, avail)
    assert [g.strip().lower() for g in got.gas_names] == avail and got.ngas == kd.ngas - 1
    idrop = syn.GAS_NAMES.index(drop) + 1
    remap = {i: (i if i < idrop else i - 1) for i in range(kd.ngas + 1) if i != idrop}
    for lu in ("lower", "upper"):
        idx = getattr(kd, f"idx_minor_{lu}")[:kd.extra[f"nminor{lu}"]]
        keep = np.nonzero(idx != idrop)[0]
        lims = getattr(kd, f"minor_limits_gpt_{lu}")[:, keep]
        assert np.array_equal(getattr(got, f"minor_limits_gpt_{lu}"), lims)
        assert np.array_equal(getattr(got, f"idx_minor_{lu}"), [remap[int(v)] for v in idx[keep]])
        # contributor columns are compacted in order; kminor_start renumbered (:1884-1900)
        ks_old = getattr(kd, f"kminor_start_{lu}")[keep]
        ng = lims[1] - lims[0] + 1
        ks_new = 1 + np.concatenate([[0], np.cumsum(ng)[:-1]])
        assert np.array_equal(getattr(got, f"kminor_start_{lu}"), ks_new)
        kold, knew = getattr(kd, f"kminor_{lu}"), getattr(got, f"kminor_{lu}")
        assert knew.shape[2] == ng.sum()
        for a, b, n in zip(ks_old, ks_new, ng):
            assert np.array_equal(knew[:, :, b - 1:b - 1 + n], kold[:, :, a - 1:a - 1 + n])
        sc_old = getattr(kd, f"idx_minor_scaling_{lu}")[keep]
        sc_new = getattr(got, f"idx_minor_scaling_{lu}")
        # a scaling gas that is withheld becomes "none" (-1), :1672
        want = [remap[int(v)] if (v > 0 and v != idrop) else -1 for v in sc_old]
        assert np.array_equal(np.maximum(sc_new, 0), np.maximum(want, 0))
    # major tables and flavours are untouched apart from the gas renumbering
    assert np.array_equal(got.kmajor, kd.kmajor)
    assert np.array_equal(got.flavor, np.vectorize(lambda v: remap[int(v)])(kd.flavor))
    assert np.array_equal(got.vmr_ref, np.delete(kd.vmr_ref, idrop, axis=1))


def test_missing_key_species_is_reported_like_the_reference(oracle_lib):
    kd = syn.make_kdist("sw")
    raw = syn.make_kdist_raw(kd)
    with pytest.raises(RuntimeError, match=r"gas_optics: required gases .*H2O.* are not provided"):
        load_kdist_raw(oracle_lib, raw, [g for g in syn.GAS_NAMES if g != "h2o"])
    with pytest.raises(RuntimeError, match="rayl_lower and rayl_upper must have the same allocation status"):   # :1303-1306
        load_kdist_raw(oracle_lib, {**raw, "rayl_upper": None}, syn.GAS_NAMES)


def test_solar_variability_and_tsi(oracle_lib):
    kd = syn.make_kdist("sw")
    raw = syn.make_kdist_raw(kd)
    base = load_kdist_raw(oracle_lib, raw, syn.GAS_NAMES)
    q, f, s = raw["solar_source_quiet"], raw["solar_source_facular"], raw["solar_source_sunspot"]
    got = load_kdist_raw(oracle_lib, raw, syn.GAS_NAMES, mg_index=0.2, sb_index=500.0)
    np.testing.assert_array_equal(got.solar_source, q + (0.2 - 0.1495954) * f + (500.0 - 0.00066696) * s)
    got = load_kdist_raw(oracle_lib, raw, syn.GAS_NAMES, tsi=1000.0)
    np.testing.assert_allclose(got.solar_source.sum(), 1000.0, rtol=1e-14)
    np.testing.assert_allclose(got.solar_source / got.solar_source.sum(), base.solar_source / base.solar_source.sum(), rtol=1e-14)
    with pytest.raises(RuntimeError, match="mg_index out of range"):
        load_kdist_raw(oracle_lib, raw, syn.GAS_NAMES, mg_index=-1.0, sb_index=1.0)
    with pytest.raises(RuntimeError, match="tsi out of range"):
        load_kdist_raw(oracle_lib, raw, syn.GAS_NAMES, tsi=-5.0)


def test_product_library_loader_equals_the_oracle_build(oracle_lib):
    """kdist_load.cpp is host-only logic compiled into the product library too: both builds must agree exactly (runs
    without a GPU - nothing here touches the device)."""
    import rte_rrtmgp_b200

    prod = rte_rrtmgp_b200.lib()
    kd = syn.make_kdist("sw", ngpt=112)
    raw = syn.make_kdist_raw(kd, seed=3)
    avail = [g for g in syn.GAS_NAMES if g != "co"]
    a, b = load_kdist_raw(prod, raw, avail), load_kdist_raw(oracle_lib, raw, avail)
    for n in FIELDS + ["idx_minor_scaling_lower", "idx_minor_scaling_upper", "krayl", "solar_source"]:
        assert np.array_equal(getattr(a, n), getattr(b, n)), n


@pytest.mark.gpu
def test_allsky_on_tables_loaded_from_the_raw_set(oracle_lib, cuda_lib):
    """End to end: raw variable set -> load() for a host that does not carry every gas -> all-sky fluxes, CUDA vs oracle."""
    from rte_rrtmgp_b200.allsky import AllSky
    from rte_rrtmgp_b200.frontend import Context

    raws = [syn.make_kdist_raw(syn.make_kdist(k, ngpt=n)) for k, n in (("lw", 128), ("sw", 112))]
    kds_g = [load_kdist_raw(cuda_lib, r, syn.GAS_NAMES) for r in raws]
    kds_c = [load_kdist_raw(oracle_lib, r, syn.GAS_NAMES) for r in raws]
    for kd in kds_g + kds_c:   # AllSky addresses gases by position in syn.GAS_NAMES
        assert [g.strip().lower() for g in kd.gas_names] == syn.GAS_NAMES
        kd.gas_names = list(syn.GAS_NAMES)
    g = AllSky(Context(cuda_lib, "cuda:0"), 40, 72, kds_g[0], kds_g[1], fused=True)
    c = AllSky(Context(oracle_lib, None), 40, 72, kds_c[0], kds_c[1])
    g.step()
    c.step()
    fg, fc = g.fluxes_host(), c.fluxes_host()
    for k in fc:
        assert np.max(np.abs(fg[k] - fc[k])) <= 1.0e-5, k
