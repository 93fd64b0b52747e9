"""Problem builders and comparison helpers restated from the reference's own test programs:
tests/rte_lw_solver_unit_tests.F90, tests/rte_sw_solver_unit_tests.F90, tests/mo_comparisons.F90,
tests/mo_testing_utils.F90 (vr, increment_with_1scl...).  Kernel-level: arrays are Fortran-ordered.
"""
import numpy as np

from rte_rrtmgp_b200.abi import fzeros, to_device, to_host

SIGMA = 5.670374419e-8  # rte_lw_solver_unit_tests.F90:61
D_DIFF = 1.0 / 0.6096748751  # rte_lw_solver_unit_tests.F90:62; mo_rte_lw.F90:146
PI = float(np.arccos(-1.0))


def allclose(tst, ref, tol=2.0):
    """tests/mo_comparisons.F90:43-55: all(abs(tst-ref) <= tol*spacing(ref))"""
    tst, ref = np.asarray(tst), np.asarray(ref)
    return bool(np.all(np.abs(tst - ref) <= tol * np.spacing(np.abs(ref))))


def max_spacings(tst, ref):
    tst, ref = np.asarray(tst), np.asarray(ref)
    return float(np.max(np.abs(tst - ref) / np.spacing(np.abs(ref))))


def dev(a, device):
    return np.asfortranarray(a) if device is None else to_device(a, device)


def host(a):
    return to_host(a)


# ---------------- LW: gray radiative equilibrium (rte_lw_solver_unit_tests.F90:241-343) -------------
def gray_rad_equil_olr(T, tau):
    return (2.0 * SIGMA * T**4) / (2 + D_DIFF * tau)


def gray_rad_equil(sfc_t, total_tau, nlay, top_at_1):
    """Returns dict(tau(ncol,nlay,1), lay_source, lev_source(ncol,nlay+1,1), sfc_source, sfc_source_Jac(ncol,1))."""
    ncol = sfc_t.size
    tau = np.zeros((ncol, nlay, 1), order="F")
    tau[:, :, 0] = (total_tau / float(nlay))[:, None]
    olr = gray_rad_equil_olr(sfc_t, total_tau)
    lev = np.zeros((ncol, nlay + 1, 1), order="F")
    lay = np.zeros((ncol, nlay, 1), order="F")
    lev[:, 0, 0] = 0.5 / PI * olr
    for ilay in range(1, nlay + 1):
        lev[:, ilay, 0] = 0.5 / PI * olr * (1.0 + D_DIFF * np.sum(tau[:, :ilay, 0], axis=1))
        lay[:, ilay - 1, 0] = 0.5 * (lev[:, ilay, 0] + lev[:, ilay - 1, 0])
    if not top_at_1:
        lev = np.asfortranarray(lev[:, ::-1, :])
        lay = np.asfortranarray(lay[:, ::-1, :])
    sfc = np.zeros((ncol, 1), order="F")
    jac = np.zeros((ncol, 1), order="F")
    sfc[:, 0] = SIGMA / PI * sfc_t**4
    jac[:, 0] = 4.0 * SIGMA / PI * sfc_t**3
    return dict(tau=tau, lay_source=lay, lev_source=lev, sfc_source=sfc, sfc_source_Jac=jac, top_at_1=top_at_1)


def lw_noscat_broadband(lib, device, prob, sfc_emis_gpt, do_jacobians=False, Ds=None, nmus=1, weights=None,
                        rescale=None, inc_flux=None):
    """Kernel-level equivalent of rte_lw() for 1scl props + ty_fluxes_broadband
    (rte/frontend/mo_rte_lw.F90:329-378).  Returns flux_up, flux_dn[, flux_upJac] on the host."""
    tau = prob["tau"]
    ncol, nlay, ngpt = tau.shape
    if Ds is None:
        Ds = np.full((ncol, ngpt, nmus), D_DIFF, order="F")
    if weights is None:
        weights = np.array([1.0])
    if inc_flux is None:
        inc_flux = np.zeros((ncol, ngpt), order="F")
    d = lambda a: dev(a, device)
    bb_up = fzeros((ncol, nlay + 1), device=device)
    bb_dn = fzeros((ncol, nlay + 1), device=device)
    jac = fzeros((ncol, nlay + 1), device=device)
    tau_d = d(tau)
    if rescale is None:
        ssa_d, g_d, do_resc = tau_d, tau_d, False  # mo_rte_lw.F90:378 passes tau as ssa and g
    else:
        ssa_d, g_d, do_resc = d(rescale[0]), d(rescale[1]), True
    decoy = fzeros((ncol, nlay + 1), device=device)
    lib.rte_lw_solver_noscat(ncol, nlay, ngpt, bool(prob["top_at_1"]), nmus, d(Ds), np.asarray(weights, dtype=np.float64),
                             tau_d, d(prob["lay_source"]), d(prob["lev_source"]), d(sfc_emis_gpt),
                             d(prob["sfc_source"]), d(inc_flux), decoy, decoy, True, bb_up, bb_dn,
                             bool(do_jacobians), d(prob["sfc_source_Jac"]), jac, do_resc, ssa_d, g_d)
    lib.sync()
    out = (host(bb_up), host(bb_dn))
    return out + (host(jac),) if do_jacobians else out


# ---------------- SW: thin scattering atmospheres (rte_sw_solver_unit_tests.F90:226-272) -------------
def thin_scattering(lib, device, tau, ssa, g, nlay):
    """ncol = ntau*nssa*ng columns, vertically uniform, then delta-scaled.  Returns host dict."""
    ntau, nssa, ng = tau.size, ssa.size, g.size
    ncol = ntau * nssa * ng
    # Fortran implied-do constructors (:246, :254, :256)
    t_col = np.array([tau[i] for _j in range(nssa * ng) for i in range(ntau)])
    s_col = np.array([ssa[i] for _j in range(ng) for i in range(nssa) for _k in range(ntau)])
    g_col = np.array([g[i] for i in range(ng) for _k in range(ntau * ng)])
    T = np.zeros((ncol, nlay, 1), order="F")
    S = np.zeros((ncol, nlay, 1), order="F")
    G = np.zeros((ncol, nlay, 1), order="F")
    T[:, :, 0] = (t_col / float(nlay))[:, None]
    S[:, :, 0] = s_col[:, None]
    G[:, :, 0] = g_col[:, None]
    Td, Sd, Gd = dev(T, device), dev(S, device), dev(G, device)
    lib.rte_delta_scale_2str_k(ncol, nlay, 1, Td, Sd, Gd)
    lib.sync()
    return dict(tau=host(Td), ssa=host(Sd), g=host(Gd))


def sw_2stream_broadband(lib, device, prob, mu0, toa_flux, alb_dir_gpt, alb_dif_gpt, top_at_1, inc_flux_dif=None):
    """Kernel-level equivalent of rte_sw() for 2str props + ty_fluxes_broadband (mo_rte_sw.F90:197-361).
    The three g-point flux outputs alias ONE decoy buffer exactly as the frontend does (:204-207)."""
    tau = prob["tau"]
    ncol, nlay, ngpt = tau.shape
    d = lambda a: dev(a, device)
    mu0_bylay = np.asfortranarray(np.repeat(mu0[:, None], nlay, axis=1))
    up = fzeros((ncol, nlay + 1), device=device)
    dn = fzeros((ncol, nlay + 1), device=device)
    dr = fzeros((ncol, nlay + 1), device=device)
    decoy = fzeros((ncol, nlay + 1, ngpt), device=device)
    has_bc = inc_flux_dif is not None
    if inc_flux_dif is None:
        inc_flux_dif = np.zeros((ncol, ngpt), order="F")
    lib.rte_sw_solver_2stream(ncol, nlay, ngpt, bool(top_at_1), d(tau), d(prob["ssa"]), d(prob["g"]), d(mu0_bylay),
                              d(alb_dir_gpt), d(alb_dif_gpt), d(toa_flux), decoy, decoy, decoy, has_bc,
                              d(inc_flux_dif), True, up, dn, dr)
    lib.sync()
    return host(up), host(dn), host(dr)


def vr(prob):
    """tests/mo_testing_utils.F90 vr(): reverse the vertical ordering of every (ncol,nlay[,+1],ngpt) field."""
    out = dict(prob)
    for k in ("tau", "ssa", "g", "lay_source", "lev_source"):
        if k in prob:
            out[k] = np.asfortranarray(prob[k][:, ::-1, :])
    if "top_at_1" in prob:
        out["top_at_1"] = not prob["top_at_1"]
    return out
