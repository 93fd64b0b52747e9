"""GPU parity, kernel by kernel, through the C-ABI: CUDA vs the CPU oracle on identical seeded random
inputs.  Covers every solver flag combination (broadband / g-point fluxes, Jacobians, 1-3 quadrature
angles, Tang rescaling, both vertical orientations, diffuse boundary condition, night columns), the
register and the tile kernel families, ragged sizes (ncol not a multiple of the tile, nlay not a multiple
of the chunk), and the host-pointer path (numpy arrays handed to the CUDA library are staged through the
device, as a Fortran host would use it).  Floating-point tolerance: 1e-11 relative to the field maximum
(libdevice exp / FMA contraction vs glibc / no-FMA); the north-star tolerance on fluxes is 1e-5."""
import numpy as np
import pytest

import refcases as rc
from rte_rrtmgp_b200.abi import fzeros

RTOL = 1.0e-11


def _close(a, b, name="", rtol=RTOL):
    scale = max(np.max(np.abs(b)), 1e-300)
    err = np.max(np.abs(a - b)) / scale
    assert err <= rtol, f"{name}: rel err {err:.3e}"


def _lw_inputs(ncol, nlay, ngpt, seed, scattering=False):
    rng = np.random.default_rng(seed)
    f = lambda *s: np.asfortranarray(rng.random(s))
    tau = np.asfortranarray(10.0 ** rng.uniform(-6, 1.5, (ncol, nlay, ngpt)))
    lev = np.asfortranarray(50.0 + 100.0 * rng.random((ncol, nlay + 1, ngpt)))
    lay = np.asfortranarray(0.5 * (lev[:, 1:] + lev[:, :-1]) + rng.uniform(-1, 1, (ncol, nlay, ngpt)))
    d = dict(tau=tau, lay=lay, lev=lev, emis=np.asfortranarray(0.8 + 0.2 * rng.random((ncol, ngpt))),
             sfc=np.asfortranarray(100 + 50 * rng.random((ncol, ngpt))), jac=f(ncol, ngpt),
             inc=np.asfortranarray(5.0 * rng.random((ncol, ngpt))))
    d["ssa"] = np.asfortranarray(rng.uniform(0, 0.9, (ncol, nlay, ngpt))) if scattering else tau
    d["g"] = np.asfortranarray(rng.uniform(-0.2, 0.9, (ncol, nlay, ngpt))) if scattering else tau
    return d


def _run_lw_noscat(lib, device, x, top_at_1, nmus, bb, jac, resc):
    ncol, nlay, ngpt = x["tau"].shape
    d = lambda a: rc.dev(a, device)
    Ds = np.asfortranarray(np.stack([np.full((ncol, ngpt), 1.0 / m) for m in (0.61, 0.25, 0.79)[:nmus]], axis=2))
    wts = np.array([1.0, 0.23, 0.77][:nmus]) if nmus > 1 else np.array([1.0])
    gup, gdn = fzeros((ncol, nlay + 1, ngpt), device=device), fzeros((ncol, nlay + 1, ngpt), device=device)
    bup, bdn, fj = (fzeros((ncol, nlay + 1), device=device) for _ in range(3))
    lib.rte_lw_solver_noscat(ncol, nlay, ngpt, top_at_1, nmus, d(Ds), d(wts), d(x["tau"]), d(x["lay"]), d(x["lev"]),
                             d(x["emis"]), d(x["sfc"]), d(x["inc"]), gup, gdn, bb, bup, bdn, jac, d(x["jac"]), fj,
                             resc, d(x["ssa"]), d(x["g"]))
    lib.sync()
    out = {"bup": bup, "bdn": bdn} if bb else {"gup": gup, "gdn": gdn}
    if jac:
        out["jac"] = fj
    return {k: rc.host(v) for k, v in out.items()}


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("top_at_1", [True, False])
@pytest.mark.parametrize("nmus,bb,jac,resc", [(1, True, False, False), (1, False, False, False), (3, True, True, False),
                                              (2, False, True, False), (1, True, True, True), (2, False, False, True)])
def test_lw_solver_noscat(oracle_lib, cuda_lib, variant, top_at_1, nmus, bb, jac, resc):
    cuda_lib.cdll.rrtmgpb_set_solver_variant(variant)
    x = _lw_inputs(21, 37, 5, seed=11, scattering=resc)
    ref = _run_lw_noscat(oracle_lib, None, x, top_at_1, nmus, bb, jac, resc)
    got = _run_lw_noscat(cuda_lib, "cuda:0", x, top_at_1, nmus, bb, jac, resc)
    cuda_lib.cdll.rrtmgpb_set_solver_variant(0)
    for k in ref:
        _close(got[k], ref[k], k)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("top_at_1", [True, False])
@pytest.mark.parametrize("per_gpt", [0, 1])
def test_lw_solver_2stream(oracle_lib, cuda_lib, variant, top_at_1, per_gpt):
    """per_gpt = 0: the reference's serial DEFAULT kernels' behaviour (every g-point uses g-point 1's level source,
    mo_rte_solver_kernels.F90:422; the oracle's default); 1: per-g-point level sources as in the reference's accel
    kernels (the CUDA library's default)."""
    x = _lw_inputs(19, 33, 4, seed=5, scattering=True)
    ncol, nlay, ngpt = x["tau"].shape

    def run(lib, device):
        lib.cdll.rrtmgpb_set_lw_2stream_lev_source_per_gpt(per_gpt)
        d = lambda a: rc.dev(a, device)
        gup, gdn = fzeros((ncol, nlay + 1, ngpt), device=device), fzeros((ncol, nlay + 1, ngpt), device=device)
        lib.rte_lw_solver_2stream(ncol, nlay, ngpt, top_at_1, d(x["tau"]), d(x["ssa"]), d(x["g"]), d(x["lay"]),
                                  d(x["lev"]), d(x["emis"]), d(x["sfc"]), d(x["inc"]), gup, gdn)
        lib.sync()
        lib.cdll.rrtmgpb_set_lw_2stream_lev_source_per_gpt(0 if device is None else 1)  # each library's default
        return rc.host(gup), rc.host(gdn)

    cuda_lib.cdll.rrtmgpb_set_solver_variant(variant)
    ref, got = run(oracle_lib, None), run(cuda_lib, "cuda:0")
    cuda_lib.cdll.rrtmgpb_set_solver_variant(0)
    # Toon's source terms divide level-source differences by tau*(gamma1+gamma2) (:951): for tau ~ 1e-6 the
    # cancellation amplifies last-bit differences (FMA contraction) to ~1e-10 relative
    _close(got[0], ref[0], "flux_up", rtol=2e-9)
    _close(got[1], ref[1], "flux_dn", rtol=2e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1])
def test_solvers_with_opaque_layers(oracle_lib, cuda_lib, variant):
    """Optical depths of 1e9 .. 1e13: -tau*D, -tau*k and -tau/mu0 times log2(e) exceed the int range of the lean exp's
    power-of-two extraction - the transmittance must still be 0 (libm underflows), not a wrapped-around huge value."""
    cuda_lib.cdll.rrtmgpb_set_solver_variant(variant)
    try:
        x = _lw_inputs(21, 37, 5, seed=12)
        x["tau"][3, 5, :] = 1.0e13
        x["tau"][7, 20, 2] = 3.0e9
        ref = _run_lw_noscat(oracle_lib, None, x, True, 1, True, False, False)
        got = _run_lw_noscat(cuda_lib, "cuda:0", x, True, 1, True, False, False)
        for k in ref:
            assert np.all(np.isfinite(got[k]))
            _close(got[k], ref[k], k)
        y = _sw_inputs(23, 41, 6, seed=4)
        y["tau"][2, 11, :] = 1.0e13
        y["tau"][9, 30, 1] = 5.0e9
        ncol, nlay, ngpt = y["tau"].shape
        res = {}
        for name, lib, device in (("ref", oracle_lib, None), ("gpu", cuda_lib, "cuda:0")):
            d = lambda a: rc.dev(a, device)
            decoy = fzeros((ncol, nlay + 1, ngpt), device=device)
            bup, bdn, bdr = (fzeros((ncol, nlay + 1), device=device) for _ in range(3))
            lib.rte_sw_solver_2stream(ncol, nlay, ngpt, True, d(y["tau"]), d(y["ssa"]), d(y["g"]), d(y["mu0"]), d(y["adir"]),
                                      d(y["adif"]), d(y["inc"]), decoy, decoy, decoy, False, d(y["dif"]), True, bup, bdn, bdr)
            lib.sync()
            res[name] = [rc.host(o) for o in (bup, bdn, bdr)]
        for a, b, n in zip(res["gpu"], res["ref"], ("up", "dn", "dir")):
            assert np.all(np.isfinite(a))
            _close(a, b, n)
    finally:
        cuda_lib.cdll.rrtmgpb_set_solver_variant(0)


@pytest.mark.gpu
def test_lw_solver_2stream_default_is_per_gpoint_level_source(oracle_lib, cuda_lib):
    """Out of the box the CUDA library indexes lev_source per g-point like the reference's accelerator kernels
    (accel/mo_rte_solver_kernels.F90:958-962); the serial kernel's g-point-1 quirk is opt-in."""
    x = _lw_inputs(19, 33, 4, seed=6, scattering=True)
    ncol, nlay, ngpt = x["tau"].shape
    out = {}
    for name, lib, device in (("ref", oracle_lib, None), ("gpu", cuda_lib, "cuda:0")):
        if device is None:
            lib.cdll.rrtmgpb_set_lw_2stream_lev_source_per_gpt(1)
        d = lambda a: rc.dev(a, device)
        gup, gdn = fzeros((ncol, nlay + 1, ngpt), device=device), fzeros((ncol, nlay + 1, ngpt), device=device)
        lib.rte_lw_solver_2stream(ncol, nlay, ngpt, True, d(x["tau"]), d(x["ssa"]), d(x["g"]), d(x["lay"]),
                                  d(x["lev"]), d(x["emis"]), d(x["sfc"]), d(x["inc"]), gup, gdn)
        lib.sync()
        if device is None:
            lib.cdll.rrtmgpb_set_lw_2stream_lev_source_per_gpt(0)
        out[name] = (rc.host(gup), rc.host(gdn))
    _close(out["gpu"][0], out["ref"][0], "flux_up", rtol=2e-9)
    _close(out["gpu"][1], out["ref"][1], "flux_dn", rtol=2e-9)


def _sw_inputs(ncol, nlay, ngpt, seed):
    rng = np.random.default_rng(seed)
    mu0 = np.asfortranarray(np.repeat(rng.uniform(-0.2, 1.0, ncol)[:, None], nlay, axis=1))
    mu0[1] = rng.uniform(0.05, 1.0, nlay)  # one column with mu0 varying with height (spherical correction)
    return dict(tau=np.asfortranarray(10.0 ** rng.uniform(-6, 1.5, (ncol, nlay, ngpt))),
                ssa=np.asfortranarray(rng.uniform(0, 1.0, (ncol, nlay, ngpt))),
                g=np.asfortranarray(rng.uniform(-0.3, 0.95, (ncol, nlay, ngpt))), mu0=mu0,
                adir=np.asfortranarray(rng.uniform(0, 0.6, (ncol, ngpt))), adif=np.asfortranarray(rng.uniform(0, 0.6, (ncol, ngpt))),
                inc=np.asfortranarray(10.0 * rng.random((ncol, ngpt))), dif=np.asfortranarray(2.0 * rng.random((ncol, ngpt))))


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("top_at_1", [True, False])
@pytest.mark.parametrize("bb,bc", [(True, False), (True, True), (False, False), (False, True)])
def test_sw_solver_2stream(oracle_lib, cuda_lib, variant, top_at_1, bb, bc):
    x = _sw_inputs(23, 41, 6, seed=3)
    ncol, nlay, ngpt = x["tau"].shape

    def run(lib, device):
        d = lambda a: rc.dev(a, device)
        if bb:  # the frontend aliases the three g-point outputs onto ONE decoy buffer (mo_rte_sw.F90:204-207)
            decoy = fzeros((ncol, nlay + 1, ngpt), device=device)
            gup = gdn = gdr = decoy
        else:
            gup, gdn, gdr = (fzeros((ncol, nlay + 1, ngpt), device=device) for _ in range(3))
        bup, bdn, bdr = (fzeros((ncol, nlay + 1), device=device) for _ in range(3))
        lib.rte_sw_solver_2stream(ncol, nlay, ngpt, top_at_1, d(x["tau"]), d(x["ssa"]), d(x["g"]), d(x["mu0"]),
                                  d(x["adir"]), d(x["adif"]), d(x["inc"]), gup, gdn, gdr, bc, d(x["dif"]), bb, bup, bdn, bdr)
        lib.sync()
        outs = (bup, bdn, bdr) if bb else (gup, gdn, gdr)
        if bb:
            assert float(np.max(np.abs(rc.host(decoy)))) == 0.0  # decoys are never written
        return [rc.host(o) for o in outs]

    cuda_lib.cdll.rrtmgpb_set_solver_variant(variant)
    ref, got = run(oracle_lib, None), run(cuda_lib, "cuda:0")
    cuda_lib.cdll.rrtmgpb_set_solver_variant(0)
    for a, b, n in zip(got, ref, ("up", "dn", "dir")):
        _close(a, b, n)


@pytest.mark.gpu
@pytest.mark.parametrize("top_at_1", [True, False])
@pytest.mark.parametrize("ncol,nlay", [(22, 90), (19, 90), (22, 137), (18, 144), (17, 112), (20, 81)])
def test_register_solvers_on_tall_columns(oracle_lib, cuda_lib, ncol, nlay, top_at_1):
    """80 < nlay <= 144 (ICON 90 layers, IFS 137): the register solvers run with 16 lanes per column (csrc/kernels/
    solver_reg.cuh) - every solver, broadband and g-point outputs, Jacobian, even ncol (TMA tiles, zero-filled padding
    rows) and odd ncol (cp.async fallback)."""
    assert cuda_lib.cdll.rrtmgpb_get_solver_variant() == 0
    x = _lw_inputs(ncol, nlay, 4, seed=21, scattering=True)
    for nmus, bb, jac in ((1, True, False), (2, False, True), (3, True, True)):
        ref = _run_lw_noscat(oracle_lib, None, x, top_at_1, nmus, bb, jac, False)
        got = _run_lw_noscat(cuda_lib, "cuda:0", x, top_at_1, nmus, bb, jac, False)
        for k in ref:
            _close(got[k], ref[k], f"lw noscat {k} nmus={nmus}")
    for lib, dev in ((oracle_lib, None), (cuda_lib, "cuda:0")):
        lib.cdll.rrtmgpb_set_lw_2stream_lev_source_per_gpt(1)
    try:
        res = {}
        for name, lib, device in (("ref", oracle_lib, None), ("gpu", cuda_lib, "cuda:0")):
            d = lambda a: rc.dev(a, device)
            gup, gdn = fzeros((ncol, nlay + 1, 4), device=device), fzeros((ncol, nlay + 1, 4), device=device)
            lib.rte_lw_solver_2stream(ncol, nlay, 4, top_at_1, d(x["tau"]), d(x["ssa"]), d(x["g"]), d(x["lay"]), d(x["lev"]),
                                      d(x["emis"]), d(x["sfc"]), d(x["inc"]), gup, gdn)
            lib.sync()
            res[name] = (rc.host(gup), rc.host(gdn))
    finally:
        oracle_lib.cdll.rrtmgpb_set_lw_2stream_lev_source_per_gpt(0)
    _close(res["gpu"][0], res["ref"][0], "lw 2stream up", rtol=2e-9)
    _close(res["gpu"][1], res["ref"][1], "lw 2stream dn", rtol=2e-9)
    y = _sw_inputs(ncol, nlay, 5, seed=22)
    for bb, bc in ((True, True), (False, False)):
        out = {}
        for name, lib, device in (("ref", oracle_lib, None), ("gpu", cuda_lib, "cuda:0")):
            d = lambda a: rc.dev(a, device)
            gup, gdn, gdr = (fzeros((ncol, nlay + 1, 5), device=device) for _ in range(3))
            bup, bdn, bdr = (fzeros((ncol, nlay + 1), device=device) for _ in range(3))
            lib.rte_sw_solver_2stream(ncol, nlay, 5, top_at_1, d(y["tau"]), d(y["ssa"]), d(y["g"]), d(y["mu0"]), d(y["adir"]),
                                      d(y["adif"]), d(y["inc"]), gup, gdn, gdr, bc, d(y["dif"]), bb, bup, bdn, bdr)
            lib.sync()
            out[name] = [rc.host(o) for o in ((bup, bdn, bdr) if bb else (gup, gdn, gdr))]
        for a, b, n in zip(out["gpu"], out["ref"], ("up", "dn", "dir")):
            _close(a, b, f"sw {n} bb={bb}")


@pytest.mark.gpu
@pytest.mark.parametrize("top_at_1", [True, False])
@pytest.mark.parametrize("ncol,nlay", [(22, 37), (24, 72), (26, 80), (22, 90), (18, 137)])
@pytest.mark.parametrize("nmus,bb,jac", [(1, True, False), (1, True, True), (2, False, True), (3, False, False)])
def test_lw_tang_rescaling_on_the_register_kernel(oracle_lib, cuda_lib, ncol, nlay, top_at_1, nmus, bb, jac):
    """do_rescaling = true (mo_rte_solver_kernels.F90:148-178, 753-844; the default LW treatment of scattering optical
    properties, mo_rte_lw.F90:395-423) with even ncol: TMA tiles, so the register kernel lw_rescl_reg_kernel runs (odd ncol
    takes the tile kernel: tests above).  Both orientations - the second downward sweep is orientation-asymmetric in the
    reference (:801-804 vs :835-838)."""
    x = _lw_inputs(ncol, nlay, 5, seed=31, scattering=True)
    ref = _run_lw_noscat(oracle_lib, None, x, top_at_1, nmus, bb, jac, True)
    got = _run_lw_noscat(cuda_lib, "cuda:0", x, top_at_1, nmus, bb, jac, True)
    for k in ref:
        _close(got[k], ref[k], k)


@pytest.mark.gpu
def test_sw_solver_noscat_and_reductions(oracle_lib, cuda_lib):
    x = _sw_inputs(17, 29, 5, seed=8)
    x["mu0"] = np.asfortranarray(np.abs(x["mu0"]) + 0.05)
    ncol, nlay, ngpt = x["tau"].shape
    res = {}
    for name, lib, device in (("ref", oracle_lib, None), ("gpu", cuda_lib, "cuda:0")):
        d = lambda a: rc.dev(a, device)
        fdir = fzeros((ncol, nlay + 1, ngpt), device=device)
        lib.rte_sw_solver_noscat(ncol, nlay, ngpt, False, d(x["tau"]), d(x["mu0"]), d(x["inc"]), fdir)
        bsum, bnet, bnet2 = (fzeros((ncol, nlay + 1), device=device) for _ in range(3))
        lib.rte_sum_broadband(ncol, nlay + 1, ngpt, fdir, bsum)
        half = d(0.5 * rc.host(fdir)) if device else np.asfortranarray(0.5 * fdir)
        lib.rte_net_broadband_full(ncol, nlay + 1, ngpt, fdir, half, bnet)
        lib.rte_net_broadband_precalc(ncol, nlay + 1, bsum, bnet, bnet2)
        lib.sync()
        res[name] = [rc.host(a) for a in (fdir, bsum, bnet, bnet2)]
    for a, b in zip(res["gpu"], res["ref"]):
        _close(a, b)


@pytest.mark.gpu
def test_host_pointers_are_staged(oracle_lib, cuda_lib):
    """A Fortran host hands the extern kernels HOST arrays; the CUDA library must stage them transparently."""
    x = _lw_inputs(9, 12, 3, seed=2)
    ref = _run_lw_noscat(oracle_lib, None, x, True, 1, True, True, False)
    got = _run_lw_noscat(cuda_lib, None, x, True, 1, True, True, False)  # numpy arrays into the CUDA library
    for k in ref:
        _close(got[k], ref[k], k)
    a = np.asfortranarray(np.random.default_rng(0).random((7, 5, 3)))
    b = a.copy(order="F")
    cuda_lib.rte_increment_1scalar_by_1scalar(7, 5, 3, a, b)
    np.testing.assert_array_equal(a, 2 * b)
    z = np.ones((4, 3, 2), order="F")
    cuda_lib.zero_array_3D(4, 3, 2, z)
    assert not z.any()
    cuda_lib.set_to_scalar_3D(4, 3, 2, z, 2.5)
    assert np.all(z == 2.5)


@pytest.mark.gpu
def test_glue_kernels(oracle_lib, cuda_lib):
    import ctypes as C

    from rte_rrtmgp_b200.abi import FLOAT, _ptr

    rng = np.random.default_rng(4)
    ncol, nlay, ngas, ngpt, nbnd = 13, 9, 4, 6, 3
    vmr = np.asfortranarray(rng.random((ncol, nlay, ngas)) * 1e-2)
    plev = np.asfortranarray(np.sort(rng.uniform(10, 1e5, (ncol, nlay + 1)), axis=1))
    play = np.asfortranarray(0.5 * (plev[:, 1:] + plev[:, :-1]))
    tlay = np.asfortranarray(rng.uniform(200, 300, (ncol, nlay)))
    tabs = np.asfortranarray(rng.random((ncol, nlay, ngpt)))
    tray = np.asfortranarray(rng.random((ncol, nlay, ngpt)) * 1e-2)
    lims = np.asfortranarray(np.array([[1, 3, 5], [2, 4, 6]], dtype=np.int32))
    band = np.asfortranarray(rng.random((nbnd, ncol)))
    P = lambda a: C.c_void_p(_ptr(a).value)
    out = {}
    for name, lib, device in (("ref", oracle_lib, None), ("gpu", cuda_lib, "cuda:0")):
        keep = []

        def d(a):  # keep every device temporary alive until the queue has drained
            keep.append(rc.dev(a, device))
            return keep[-1]

        c = lib.cdll
        col_dry = fzeros((ncol, nlay), device=device)
        c.rrtmgpb_get_col_dry(ncol, nlay, P(d(np.asfortranarray(vmr[:, :, 0]))), P(d(plev)), P(col_dry))
        col_gas = fzeros((ncol, nlay, ngas + 1), device=device)
        c.rrtmgpb_col_gas_from_vmr(ncol, nlay, ngas, P(d(vmr)), P(col_dry), P(col_gas))
        tlev = fzeros((ncol, nlay + 1), device=device)
        c.rrtmgpb_interpolate_tlev(ncol, nlay, P(d(play)), P(d(plev)), P(d(tlay)), P(tlev))
        tau, ssa, g = (fzeros((ncol, nlay, ngpt), device=device) for _ in range(3))
        c.rrtmgpb_combine_abs_and_rayleigh(ncol, nlay, ngpt, 2, P(d(tabs)), P(d(tray)), P(tau), P(ssa), P(g))
        exp = fzeros((ncol, ngpt), device=device)
        c.rrtmgpb_expand_and_transpose(ncol, nbnd, ngpt, P(d(lims)), P(d(band)), P(exp))
        lib.sync()
        out[name] = [rc.host(a) for a in (col_dry, col_gas, tlev, tau, ssa, g, exp)]
    for a, b in zip(out["gpu"], out["ref"]):
        _close(a, b)
    assert np.array_equal(out["gpu"][6], np.repeat(band.T, 2, axis=1))
