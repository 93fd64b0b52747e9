"""Single-precision builds (the reference's RTE_ENABLE_SP, rte/kernels/mo_rte_kind.F90:28-36: wp = c_float): the product
library rte_rrtmgp_b200/lib/librte_rrtmgp_b200_sp.so and the oracle's -DRTE_USE_SP build export the same 45 symbols on
float32 arrays.

Tolerances: the reference's own regression threshold for single-precision builds is 3.5e-1 W/m2 ABSOLUTE on fluxes
against double-precision results (examples/CMakeLists.txt:1-5, FAILURE_THRESHOLD; 7e-4 in double precision): the CUDA
single-precision fluxes are held to that against the DOUBLE-precision oracle on realistic magnitudes, and to a few
hundred float spacings of the field maximum against the single-precision oracle (same algorithm, libm expf vs the
device's, FMA contraction, the chunk-level scans' association)."""
import numpy as np
import pytest

import gas_optics_calls as gc
import refcases as rc
from rte_rrtmgp_b200 import synthetic as syn
from rte_rrtmgp_b200._abi_table import ABI
from rte_rrtmgp_b200.abi import fzeros

SP_VS_DP_FLUX_ATOL = 3.5e-1   # W/m2, examples/CMakeLists.txt:2
SP_RTOL = 3.0e-5              # of the field maximum, CUDA-SP vs oracle-SP (float eps = 6e-8; ~100-layer recurrences)


@pytest.fixture(scope="module")
def oracle_sp():
    import oracle

    return oracle.lib(sp=True)


@pytest.fixture(scope="module")
def cuda_sp():
    import rte_rrtmgp_b200

    return rte_rrtmgp_b200.lib_sp()


def _close(a, b, name, rtol=SP_RTOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.max(np.abs(b)), 1e-30)
    err = np.max(np.abs(a - b)) / scale
    assert np.all(np.isfinite(a)), name
    assert err <= rtol, f"{name}: rel err {err:.3e}"


# ------------------------------------------------------------------------------------------------ CPU
def test_sp_libraries_load_and_export_every_symbol(oracle_sp):
    import rte_rrtmgp_b200
    from rte_rrtmgp_b200.abi import KernelLib

    prod = KernelLib(rte_rrtmgp_b200.LIB_PATH_SP)   # loading needs no GPU; no compute call is made here
    assert prod.float_bytes == 4 and prod.np_float is np.float32 and prod.backend.startswith("cuda")
    assert oracle_sp.float_bytes == 4
    for name in ABI:
        assert prod.has(name), name
        assert oracle_sp.has(name), name
    assert rte_rrtmgp_b200.KernelLib(rte_rrtmgp_b200.LIB_PATH).float_bytes == 8


def _lw_problem(ncol, nlay, ngpt, seed):
    rng = np.random.default_rng(seed)
    tau = np.asfortranarray(10.0 ** rng.uniform(-5, 1.2, (ncol, nlay, ngpt)))
    lev = np.asfortranarray(20.0 + 90.0 * rng.random((ncol, nlay + 1, ngpt)))       # W/m2/sr-ish magnitudes
    lay = np.asfortranarray(0.5 * (lev[:, 1:] + lev[:, :-1]))
    return dict(tau=tau, lay=lay, lev=lev, emis=np.asfortranarray(0.85 + 0.15 * rng.random((ncol, ngpt))),
                sfc=np.asfortranarray(80 + 40 * rng.random((ncol, ngpt))), jac=np.asfortranarray(rng.random((ncol, ngpt))),
                inc=np.asfortranarray(2.0 * rng.random((ncol, ngpt))),
                ssa=np.asfortranarray(rng.uniform(0, 0.9, (ncol, nlay, ngpt))),
                g=np.asfortranarray(rng.uniform(-0.2, 0.9, (ncol, nlay, ngpt))))


def _run_lw(lib, device, x, top_at_1, nmus, bb, jac, resc=False):
    ncol, nlay, ngpt = x["tau"].shape
    F = lib.np_float
    d = lambda a: rc.dev(lib.cast(np.asfortranarray(a)), device)
    Ds = np.asfortranarray(np.stack([np.full((ncol, ngpt), 1.0 / m) for m in (0.61, 0.25, 0.79)[:nmus]], axis=2))
    wts = np.array([1.0, 0.23, 0.77][:nmus]) if nmus > 1 else np.array([1.0])
    gup, gdn = fzeros((ncol, nlay + 1, ngpt), F, device), fzeros((ncol, nlay + 1, ngpt), F, device)
    bup, bdn, fj = (fzeros((ncol, nlay + 1), F, device) for _ in range(3))
    lib.rte_lw_solver_noscat(ncol, nlay, ngpt, top_at_1, nmus, d(Ds), d(wts), d(x["tau"]), d(x["lay"]), d(x["lev"]),
                             d(x["emis"]), d(x["sfc"]), d(x["inc"]), gup, gdn, bb, bup, bdn, jac, d(x["jac"]), fj,
                             resc, d(x["ssa"]), d(x["g"]))
    lib.sync()
    out = {"bup": bup, "bdn": bdn} if bb else {"gup": gup, "gdn": gdn}
    if jac:
        out["jac"] = fj
    return {k: rc.host(v) for k, v in out.items()}


def test_oracle_sp_tracks_oracle_dp_on_lw_fluxes(oracle_sp, oracle_lib):
    x = _lw_problem(12, 40, 16, seed=2)
    sp = _run_lw(oracle_sp, None, x, True, 1, True, False)
    dp = _run_lw(oracle_lib, None, x, True, 1, True, False)
    for k in dp:
        assert sp[k].dtype == np.float32
        assert np.max(np.abs(sp[k].astype(np.float64) - dp[k])) <= SP_VS_DP_FLUX_ATOL, k
        _close(sp[k], dp[k], k, rtol=2e-5)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("top_at_1", [True, False])
@pytest.mark.parametrize("ncol,nlay", [(24, 72), (20, 41), (23, 60), (36, 96), (16, 137)])
def test_sp_lw_noscat(oracle_sp, oracle_lib, cuda_sp, ncol, nlay, top_at_1, variant):
    """ncol a multiple of 4: TMA tiles (64-byte rows, SWIZZLE_64B); otherwise the cp.async staging; 16-lane kernels above 80 layers."""
    x = _lw_problem(ncol, nlay, 16, seed=ncol + nlay)
    cuda_sp.cdll.rrtmgpb_set_solver_variant(variant)
    try:
        for nmus, bb, jac, resc in ((1, True, False, False), (3, True, True, False), (2, False, True, False), (1, True, True, True)):
            got = _run_lw(cuda_sp, "cuda:0", x, top_at_1, nmus, bb, jac, resc)
            ref = _run_lw(oracle_sp, None, x, top_at_1, nmus, bb, jac, resc)
            dp = _run_lw(oracle_lib, None, x, top_at_1, nmus, bb, jac, resc)
            for k in ref:
                assert got[k].dtype == np.float32
                _close(got[k], ref[k], f"{k} nmus={nmus} bb={bb} resc={resc}")
                assert np.max(np.abs(got[k].astype(np.float64) - dp[k])) <= SP_VS_DP_FLUX_ATOL, k
    finally:
        cuda_sp.cdll.rrtmgpb_set_solver_variant(0)


def _sw_problem(ncol, nlay, ngpt, seed):
    rng = np.random.default_rng(seed)
    mu0 = np.asfortranarray(np.repeat(rng.uniform(-0.2, 1.0, ncol)[:, None], nlay, axis=1))
    mu0[1] = rng.uniform(0.05, 1.0, nlay)
    return dict(tau=np.asfortranarray(10.0 ** rng.uniform(-5, 1.2, (ncol, nlay, ngpt))),
                ssa=np.asfortranarray(rng.uniform(0, 0.999, (ncol, nlay, ngpt))),
                g=np.asfortranarray(rng.uniform(-0.3, 0.9, (ncol, nlay, ngpt))), mu0=mu0,
                adir=np.asfortranarray(rng.uniform(0, 0.6, (ncol, ngpt))), adif=np.asfortranarray(rng.uniform(0, 0.6, (ncol, ngpt))),
                inc=np.asfortranarray(8.0 * rng.random((ncol, ngpt))), dif=np.asfortranarray(1.0 * rng.random((ncol, ngpt))))


def _run_sw(lib, device, x, top_at_1, bb, bc):
    ncol, nlay, ngpt = x["tau"].shape
    F = lib.np_float
    d = lambda a: rc.dev(lib.cast(a), device)
    if bb:
        decoy = fzeros((ncol, nlay + 1, ngpt), F, device)
        gup = gdn = gdr = decoy
    else:
        gup, gdn, gdr = (fzeros((ncol, nlay + 1, ngpt), F, device) for _ in range(3))
    bup, bdn, bdr = (fzeros((ncol, nlay + 1), F, device) for _ in range(3))
    lib.rte_sw_solver_2stream(ncol, nlay, ngpt, top_at_1, d(x["tau"]), d(x["ssa"]), d(x["g"]), d(x["mu0"]),
                              d(x["adir"]), d(x["adif"]), d(x["inc"]), gup, gdn, gdr, bc, d(x["dif"]), bb, bup, bdn, bdr)
    lib.sync()
    return [rc.host(o) for o in ((bup, bdn, bdr) if bb else (gup, gdn, gdr))]


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1, 3])
@pytest.mark.parametrize("top_at_1", [True, False])
@pytest.mark.parametrize("ncol,nlay", [(24, 72), (20, 41), (23, 60), (36, 96), (16, 137)])
def test_sp_sw_2stream(oracle_sp, oracle_lib, cuda_sp, ncol, nlay, top_at_1, variant):
    x = _sw_problem(ncol, nlay, 14, seed=ncol * nlay)
    cuda_sp.cdll.rrtmgpb_set_solver_variant(variant)
    try:
        for bb, bc in ((True, False), (False, True)):
            got = _run_sw(cuda_sp, "cuda:0", x, top_at_1, bb, bc)
            ref = _run_sw(oracle_sp, None, x, top_at_1, bb, bc)
            dp = _run_sw(oracle_lib, None, x, top_at_1, bb, bc)
            for a, b, c, n in zip(got, ref, dp, ("up", "dn", "dir")):
                assert a.dtype == np.float32
                # the two-stream coefficients divide by 1 - (k mu0)^2 (mo_rte_solver_kernels.F90:1071): where that is
                # small, single-precision rounding of k is amplified in both implementations alike
                _close(a, b, f"{n} bb={bb}", rtol=2.0e-4)
                # against double precision: the reference's threshold, or - on these adversarial random cells (ssa up to
                # 0.999, k*mu0 near 1) where single precision itself is worse than that - no worse than twice what
                # the single-precision ORACLE loses
                err_gpu, err_ref = np.max(np.abs(a.astype(np.float64) - c)), np.max(np.abs(b.astype(np.float64) - c))
                assert err_gpu <= max(SP_VS_DP_FLUX_ATOL, 2.0 * err_ref), (n, err_gpu, err_ref)
    finally:
        cuda_sp.cdll.rrtmgpb_set_solver_variant(0)


@pytest.mark.gpu
@pytest.mark.parametrize("top_at_1", [True, False])
def test_sp_lw_2stream(oracle_sp, cuda_sp, top_at_1):
    x = _lw_problem(20, 60, 10, seed=9)
    ncol, nlay, ngpt = x["tau"].shape
    # the two-stream source divides the level-source difference by tau*(gamma1 + gamma2) and then cancels the quotient
    # against itself (mo_rte_solver_kernels.F90:947-957): thin layers with large level-to-level jumps are noise in single
    # precision in ANY implementation, so this case uses a smooth source profile and tau >= 0.01
    rng = np.random.default_rng(10)
    x["tau"] = np.asfortranarray(10.0 ** rng.uniform(-2, 1.0, (ncol, nlay, ngpt)))
    x["lev"] = np.asfortranarray(40.0 + 60.0 * np.linspace(0, 1, nlay + 1)[None, :, None] + 0.2 * rng.random((ncol, nlay + 1, ngpt)))
    x["lay"] = np.asfortranarray(0.5 * (x["lev"][:, 1:] + x["lev"][:, :-1]))
    res = {}
    for name, lib, device in (("ref", oracle_sp, None), ("gpu", cuda_sp, "cuda:0")):
        lib.cdll.rrtmgpb_set_lw_2stream_lev_source_per_gpt(1)
        try:
            d = lambda a: rc.dev(lib.cast(a), device)
            gup, gdn = (fzeros((ncol, nlay + 1, ngpt), lib.np_float, device) for _ in range(2))
            lib.rte_lw_solver_2stream(ncol, nlay, ngpt, top_at_1, d(x["tau"]), d(x["ssa"]), d(x["g"]), d(x["lay"]), d(x["lev"]),
                                      d(x["emis"]), d(x["sfc"]), d(x["inc"]), gup, gdn)
            lib.sync()
            res[name] = [rc.host(gup), rc.host(gdn)]
        finally:
            lib.cdll.rrtmgpb_set_lw_2stream_lev_source_per_gpt(1 if name == "gpu" else 0)
    for a, b, n in zip(res["gpu"], res["ref"], ("up", "dn")):
        _close(a, b, n, rtol=1.0e-3)


@pytest.mark.gpu
def test_sp_optical_props_and_reductions(oracle_sp, cuda_sp):
    rng = np.random.default_rng(3)
    ncol, nlay, ngpt, nbnd = 21, 17, 11, 3
    lims = np.asfortranarray(np.array([[1, 4], [5, 5], [6, 11]], dtype=np.int32).T)
    t1, s1, g1 = (np.asfortranarray(rng.uniform(0.05, 2.0, (ncol, nlay, ngpt))), np.asfortranarray(rng.uniform(0.05, 0.95, (ncol, nlay, ngpt))),
                  np.asfortranarray(rng.uniform(-0.5, 0.9, (ncol, nlay, ngpt))))
    t2, s2, g2 = (np.asfortranarray(rng.uniform(0.0, 3.0, (ncol, nlay, nbnd))), np.asfortranarray(rng.uniform(0.05, 0.95, (ncol, nlay, nbnd))),
                  np.asfortranarray(rng.uniform(-0.5, 0.9, (ncol, nlay, nbnd))))
    flux = np.asfortranarray(50.0 * rng.random((ncol, nlay + 1, ngpt)))
    out = {}
    for name, lib, device in (("ref", oracle_sp, None), ("gpu", cuda_sp, "cuda:0")):
        d = lambda a: rc.dev(lib.cast(a.copy(order="F")), device)
        a, b, c = d(t1), d(s1), d(g1)
        lib.rte_inc_2stream_by_2stream_bybnd(ncol, nlay, ngpt, a, b, c, d(t2), d(s2), d(g2), nbnd, d(lims))
        lib.rte_delta_scale_2str_k(ncol, nlay, ngpt, a, b, c)
        bb = fzeros((ncol, nlay + 1), lib.np_float, device)
        lib.rte_sum_broadband(ncol, nlay + 1, ngpt, d(flux), bb)
        lib.sync()
        out[name] = [rc.host(v) for v in (a, b, c, bb)]
    for a, b, n in zip(out["gpu"], out["ref"], ("tau", "ssa", "g", "broadband")):
        assert a.dtype == np.float32
        _close(a, b, n, rtol=2.0e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["lw", "sw"])
def test_sp_gas_optics_symbols(oracle_sp, cuda_sp, kind):
    """interpolation -> tau_absorption -> Rayleigh / Planck source in single precision.  Index outputs are compared where
    the single-precision interpolation coordinate is not within rounding of a table node (there the two libm's may
    legitimately round to different sides; the interpolant is continuous across nodes, so tau still agrees)."""
    kd = syn.make_kdist(kind)
    ncol, nlay = 36, 24
    x = gc.profile(kd, ncol, nlay, seed=31, top_at_1=True)
    it_g = gc.interpolation(cuda_sp, "cuda:0", kd, x["play"], x["tlay"], x["col_gas"], keep_device=True)
    it_c = gc.interpolation(oracle_sp, None, kd, x["play"], x["tlay"], x["col_gas"], keep_device=True)
    for k in ("jtemp", "jpress", "tropo", "jeta"):
        same = np.mean(rc.host(it_g[k]) == it_c[k])
        assert same >= 0.995, f"{k}: only {same:.4f} of the indices agree"
    # tau from each library's OWN interpolation state (a Fortran host would chain them that way)
    tau_g = gc.tau_absorption(cuda_sp, "cuda:0", kd, x["play"], x["tlay"], x["col_gas"], it_g)
    tau_c = gc.tau_absorption(oracle_sp, None, kd, x["play"], x["tlay"], x["col_gas"], it_c)
    assert tau_g.dtype == np.float32
    rel = np.abs(tau_g.astype(np.float64) - tau_c) / np.maximum(np.abs(tau_c), 1e-30)
    assert np.quantile(rel, 0.999) <= 2.0e-4 and np.max(np.abs(tau_g.astype(np.float64) - tau_c)) <= 1e-3 * np.max(tau_c)
    if kind == "sw":
        _close(gc.tau_rayleigh(cuda_sp, "cuda:0", kd, x["col_dry"], x["col_gas"], it_g),
               gc.tau_rayleigh(oracle_sp, None, kd, x["col_dry"], x["col_gas"], it_c), "tau_rayleigh", rtol=1e-3)
    else:
        got = gc.planck_source(cuda_sp, "cuda:0", kd, x["tlay"], x["tlev"], x["tsfc"], nlay, it_g)
        ref = gc.planck_source(oracle_sp, None, kd, x["tlay"], x["tlev"], x["tsfc"], nlay, it_c)
        for a, b, n in zip(got[:3], ref[:3], ("sfc_src", "lay_src", "lev_src")):
            _close(a, b, n, rtol=1e-3)
