"""A SECOND, independent restatement of the two-stream flux solvers, written in vectorised numpy straight from the Fortran
(rte/kernels/mo_rte_solver_kernels.F90) without looking at oracle/rte_solver_ref.c: lw_solver_2stream (:377-440 with
lw_two_stream :854-916, lw_source_2str :920-967), sw_solver_2stream (:503-609 with sw_dif_and_source :985-1127) and adding
(:1135-1245).  Columns are vectorised, layers and g-points are loops; every expression keeps the Fortran's association
(left to right for equal precedence), exp() goes through libm element by element (numpy's own vector exp is not bit-identical
to glibc's).  tests/test_oracle_crosscheck.py requires the C oracle's parity build (-O2 -ffp-contract=off) to agree with this
file BIT FOR BIT - test infrastructure only."""
import math

import numpy as np

_exp = np.frompyfunc(math.exp, 1, 1)
PI = math.acos(-1.0)                       # mo_rte_solver_kernels.F90:38
EPS = np.finfo(np.float64).eps
LW_DIFF_SEC = float(np.float32(1.66))      # :870 real literal without kind -> single precision, widened


def exp(x):
    return _exp(x).astype(np.float64)


def adding(top_at_1, albedo_sfc, rdif, tdif, src_dn, src_up, src_sfc, flux_dn_top):
    """:1135-1245.  rdif.. are (ncol, nlay); returns flux_up, flux_dn (ncol, nlay+1); flux_dn_top is the incident diffuse flux."""
    ncol, nlay = rdif.shape
    albedo = np.zeros((ncol, nlay + 1))
    src = np.zeros((ncol, nlay + 1))
    denom = np.zeros((ncol, nlay))
    flux_up = np.zeros((ncol, nlay + 1))
    flux_dn = np.zeros((ncol, nlay + 1))
    if top_at_1:
        albedo[:, nlay] = albedo_sfc
        src[:, nlay] = src_sfc
        for ilev in range(nlay - 1, -1, -1):   # Fortran ilev = nlay .. 1
            denom[:, ilev] = 1.0 / (1.0 - rdif[:, ilev] * albedo[:, ilev + 1])
            albedo[:, ilev] = rdif[:, ilev] + tdif[:, ilev] * tdif[:, ilev] * albedo[:, ilev + 1] * denom[:, ilev]
            src[:, ilev] = src_up[:, ilev] + tdif[:, ilev] * denom[:, ilev] * (src[:, ilev + 1] + albedo[:, ilev + 1] * src_dn[:, ilev])
        flux_dn[:, 0] = flux_dn_top
        flux_up[:, 0] = flux_dn[:, 0] * albedo[:, 0] + src[:, 0]
        for ilev in range(1, nlay + 1):
            flux_dn[:, ilev] = (tdif[:, ilev - 1] * flux_dn[:, ilev - 1] + rdif[:, ilev - 1] * src[:, ilev] + src_dn[:, ilev - 1]) * denom[:, ilev - 1]
            flux_up[:, ilev] = flux_dn[:, ilev] * albedo[:, ilev] + src[:, ilev]
    else:
        albedo[:, 0] = albedo_sfc
        src[:, 0] = src_sfc
        for ilev in range(nlay):
            denom[:, ilev] = 1.0 / (1.0 - rdif[:, ilev] * albedo[:, ilev])
            albedo[:, ilev + 1] = rdif[:, ilev] + tdif[:, ilev] * tdif[:, ilev] * albedo[:, ilev] * denom[:, ilev]
            src[:, ilev + 1] = src_up[:, ilev] + tdif[:, ilev] * denom[:, ilev] * (src[:, ilev] + albedo[:, ilev] * src_dn[:, ilev])
        flux_dn[:, nlay] = flux_dn_top
        flux_up[:, nlay] = flux_dn[:, nlay] * albedo[:, nlay] + src[:, nlay]
        for ilev in range(nlay - 1, -1, -1):
            flux_dn[:, ilev] = (tdif[:, ilev] * flux_dn[:, ilev + 1] + rdif[:, ilev] * src[:, ilev] + src_dn[:, ilev]) * denom[:, ilev]
            flux_up[:, ilev] = flux_dn[:, ilev] * albedo[:, ilev] + src[:, ilev]
    return flux_up, flux_dn


def lw_two_stream(tau, w0, g):
    """:854-916 for one g-point: (ncol, nlay) arrays."""
    gamma1 = LW_DIFF_SEC * (1.0 - 0.5 * w0 * (1.0 + g))
    gamma2 = LW_DIFF_SEC * 0.5 * w0 * (1.0 - g)
    k = np.sqrt(np.maximum((gamma1 - gamma2) * (gamma1 + gamma2), 1.0e-12))
    e = exp(-tau * k)
    e2 = e * e
    rt = 1.0 / (k * (1.0 + e2) + gamma1 * (1.0 - e2))
    rdif = rt * gamma2 * (1.0 - e2)
    tdif = rt * 2.0 * k * e
    return gamma1, gamma2, rdif, tdif


def lw_source_2str(top_at_1, sfc_emis, sfc_src, lay_source, lev_source, gamma1, gamma2, rdif, tdif, tau):
    """:920-967.  lev_source is (ncol, nlay+1) - WHICH g-point's plane the caller passes is the caller's business (:422)."""
    ncol, nlay = tau.shape
    source_up = np.zeros((ncol, nlay))
    source_dn = np.zeros((ncol, nlay))
    for ilay in range(nlay):
        top, bot = (lev_source[:, ilay], lev_source[:, ilay + 1]) if top_at_1 else (lev_source[:, ilay + 1], lev_source[:, ilay])
        t = tau[:, ilay]
        on = t > 1.0e-8
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            Z = (bot - top) / (t * (gamma1[:, ilay] + gamma2[:, ilay]))
            zup_top, zup_bot, zdn_top, zdn_bot = Z + top, Z + bot, -Z + top, -Z + bot
            su = PI * (zup_top - rdif[:, ilay] * zdn_top - tdif[:, ilay] * zup_bot)
            sd = PI * (zdn_bot - rdif[:, ilay] * zup_bot - tdif[:, ilay] * zdn_top)
        source_up[:, ilay] = np.where(on, su, 0.0)
        source_dn[:, ilay] = np.where(on, sd, 0.0)
    return source_dn, source_up, PI * sfc_emis * sfc_src


def lw_solver_2stream(top_at_1, tau, ssa, g, lay_source, lev_source, sfc_emis, sfc_src, inc_flux, lev_per_gpt):
    """:377-440.  lev_per_gpt False = the serial kernel as written: `lev_source` (rank 3) is passed whole to a rank-2 dummy, so
    EVERY g-point sees g-point 1's level source (:422); True = per g-point, like the accelerator kernels."""
    ncol, nlay, ngpt = tau.shape
    flux_up = np.zeros((ncol, nlay + 1, ngpt), order="F")
    flux_dn = np.zeros((ncol, nlay + 1, ngpt), order="F")
    for ig in range(ngpt):
        g1, g2, rdif, tdif = lw_two_stream(tau[:, :, ig], ssa[:, :, ig], g[:, :, ig])
        lev = lev_source[:, :, ig if lev_per_gpt else 0]
        sdn, sup, ssfc = lw_source_2str(top_at_1, sfc_emis[:, ig], sfc_src[:, ig], lay_source[:, :, ig], lev, g1, g2, rdif, tdif,
                                        tau[:, :, ig])
        fu, fd = adding(top_at_1, 1.0 - sfc_emis[:, ig], rdif, tdif, sdn, sup, ssfc, inc_flux[:, ig])
        flux_up[:, :, ig], flux_dn[:, :, ig] = fu, fd
    return flux_up, flux_dn


def sw_dif_and_source(top_at_1, mu0, sfc_albedo, tau, w0, g, flux_dir_top):
    """:985-1127 for one g-point.  Returns Rdif, Tdif, source_dn, source_up, source_sfc, flux_dn_dir (ncol, nlay+1)."""
    ncol, nlay = tau.shape
    min_k = 1.0e4 * EPS
    min_mu0 = math.sqrt(EPS)
    rdif = np.zeros((ncol, nlay)); tdif = np.zeros((ncol, nlay))
    source_up = np.zeros((ncol, nlay)); source_dn = np.zeros((ncol, nlay))
    fdir = np.zeros((ncol, nlay + 1))
    fdir[:, 0 if top_at_1 else nlay] = flux_dir_top
    lay_index = 0
    for j in range(nlay):
        if top_at_1:
            lay_index, inc, trans = j, j, j + 1
        else:
            lay_index = nlay - 1 - j
            inc, trans = lay_index + 1, lay_index
        tau_s, w0_s, g_s = tau[:, lay_index], w0[:, lay_index], g[:, lay_index]
        gamma1 = (8.0 - w0_s * (5.0 + 3.0 * g_s)) * 0.25
        gamma2 = 3.0 * (w0_s * (1.0 - g_s)) * 0.25
        k = np.sqrt(np.maximum((gamma1 - gamma2) * (gamma1 + gamma2), min_k))
        e = exp(-tau_s * k)
        e2 = e * e
        rt = 1.0 / (k * (1.0 + e2) + gamma1 * (1.0 - e2))
        rdif[:, lay_index] = rt * gamma2 * (1.0 - e2)
        tdif[:, lay_index] = rt * 2.0 * k * e
        mu0_s = np.maximum(min_mu0, mu0[:, lay_index])
        k_mu = k * mu0_s
        om = 1.0 - k_mu * k_mu
        rt = w0_s * rt / np.where(np.abs(om) >= EPS, om, EPS)
        gamma3 = (2.0 - 3.0 * mu0_s * g_s) * 0.25
        gamma4 = 1.0 - gamma3
        alpha1 = gamma1 * gamma4 + gamma2 * gamma3
        alpha2 = gamma1 * gamma3 + gamma2 * gamma4
        k_gamma3 = k * gamma3
        k_gamma4 = k * gamma4
        tnoscat = exp(-tau_s / mu0_s)
        rdir = rt * ((1.0 - k_mu) * (alpha2 + k_gamma3) - (1.0 + k_mu) * (alpha2 - k_gamma3) * e2 -
                     2.0 * (k_gamma3 - alpha2 * k_mu) * e * tnoscat)
        tdir = -rt * ((1.0 + k_mu) * (alpha1 + k_gamma4) * tnoscat - (1.0 - k_mu) * (alpha1 - k_gamma4) * e2 * tnoscat -
                      2.0 * (k_gamma4 + alpha1 * k_mu) * e)
        rdir = np.maximum(0.0, np.minimum(rdir, 1.0 - tnoscat))
        tdir = np.maximum(0.0, np.minimum(tdir, 1.0 - tnoscat - rdir))
        source_up[:, lay_index] = rdir * fdir[:, inc]
        source_dn[:, lay_index] = tdir * fdir[:, inc]
        fdir[:, trans] = tnoscat * fdir[:, inc]
    sfc_level = nlay if top_at_1 else 0          # dir_flux_trans of the last layer visited
    source_sfc = np.where(mu0[:, lay_index] > 0.0, fdir[:, sfc_level] * sfc_albedo, 0.0)
    night = mu0 <= 0.0
    source_up[night] = 0.0
    source_dn[night] = 0.0
    return rdif, tdif, source_dn, source_up, source_sfc, fdir


def sw_solver_2stream(top_at_1, tau, ssa, g, mu0, sfc_alb_dir, sfc_alb_dif, inc_flux_dir, has_dif_bc, inc_flux_dif, do_broadband):
    """:503-609.  Returns (flux_up, flux_dn, flux_dir): g-point arrays (ncol, nlay+1, ngpt), or broadband (ncol, nlay+1)."""
    ncol, nlay, ngpt = tau.shape
    top_layer = 0 if top_at_1 else nlay - 1
    if do_broadband:
        out = [np.zeros((ncol, nlay + 1)) for _ in range(3)]
    else:
        out = [np.zeros((ncol, nlay + 1, ngpt), order="F") for _ in range(3)]
    for ig in range(ngpt):
        dir_top = inc_flux_dir[:, ig] * mu0[:, top_layer]
        dn_top = inc_flux_dif[:, ig] if has_dif_bc else np.zeros(ncol)
        rdif, tdif, sdn, sup, ssfc, fdir = sw_dif_and_source(top_at_1, mu0, sfc_alb_dir[:, ig], tau[:, :, ig], ssa[:, :, ig],
                                                             g[:, :, ig], dir_top)
        fu, fd = adding(top_at_1, sfc_alb_dif[:, ig], rdif, tdif, sdn, sup, ssfc, dn_top)
        if do_broadband:
            out[0] = out[0] + fu
            out[1] = out[1] + fd + fdir
            out[2] = out[2] + fdir
        else:
            out[0][:, :, ig], out[1][:, :, ig], out[2][:, :, ig] = fu, fd + fdir, fdir
    return out


# ---------------------------------------------------------------------------------------------------------------------
# lw_solver_noscat (:51-376) with lw_source_noscat (:620-665), lw_transport_noscat_dn / _up (:671-741) and the Tang
# rescaling sweep lw_transport_1rescl (:753-844)
# ---------------------------------------------------------------------------------------------------------------------
def lw_source_noscat(top_at_1, lay_source, lev_source, tau, trans):
    ncol, nlay = tau.shape
    tau_thresh = math.sqrt(math.sqrt(EPS))
    with np.errstate(divide="ignore", invalid="ignore"):
        fact_big = (1.0 - trans) / tau - trans
    fact_small = tau * (0.5 + tau * (-1.0 / 3.0 + tau * 1.0 / 8.0))
    fact = np.where(tau > tau_thresh, fact_big, fact_small)
    source_inc = (1.0 - trans) * lev_source[:, 1:] + 2.0 * fact * (lay_source - lev_source[:, 1:])
    source_dec = (1.0 - trans) * lev_source[:, :-1] + 2.0 * fact * (lay_source - lev_source[:, :-1])
    return (source_inc, source_dec) if top_at_1 else (source_dec, source_inc)   # (source_dn, source_up)


def lw_solver_noscat_oneangle(top_at_1, D, weight, tau, lay_source, lev_source, sfc_emis, sfc_src, incident_flux, do_broadband,
                              do_jacobians, sfc_srcJac, do_rescaling, ssa, g):
    ncol, nlay, ngpt = tau.shape
    top_level, sfc_level = (0, nlay) if top_at_1 else (nlay, 0)
    flux_up = np.zeros((ncol, nlay + 1, ngpt), order="F"); flux_dn = np.zeros((ncol, nlay + 1, ngpt), order="F")
    bb_up = np.zeros((ncol, nlay + 1)); bb_dn = np.zeros((ncol, nlay + 1)); jac = np.zeros((ncol, nlay + 1))
    for ig in range(ngpt):
        dn = np.zeros((ncol, nlay + 1)); up = np.zeros((ncol, nlay + 1)); gj = np.zeros((ncol, nlay + 1))
        dn[:, top_level] = incident_flux[:, ig] / (PI * weight)
        if do_rescaling:
            ssal = ssa[:, :, ig]
            wb = ssal * (1.0 - g[:, :, ig]) * 0.5
            scale_tau = 1.0 - ssal + wb
            Cn = 0.4 * wb / scale_tau
            tau_loc = tau[:, :, ig] * D[:, ig][:, None] * scale_tau
            trans = exp(-tau_loc)
            An = 1.0 - trans * trans
        else:
            tau_loc = tau[:, :, ig] * D[:, ig][:, None]
            trans = exp(-tau_loc)
        source_dn, source_up = lw_source_noscat(top_at_1, lay_source[:, :, ig], lev_source[:, :, ig], tau_loc, trans)
        # lw_transport_noscat_dn
        if top_at_1:
            for ilev in range(1, nlay + 1):
                dn[:, ilev] = trans[:, ilev - 1] * dn[:, ilev - 1] + source_dn[:, ilev - 1]
        else:
            for ilev in range(nlay - 1, -1, -1):
                dn[:, ilev] = trans[:, ilev] * dn[:, ilev + 1] + source_dn[:, ilev]
        up[:, sfc_level] = dn[:, sfc_level] * (1.0 - sfc_emis[:, ig]) + sfc_emis[:, ig] * sfc_src[:, ig]
        if do_jacobians:
            gj[:, sfc_level] = sfc_emis[:, ig] * sfc_srcJac[:, ig]
        if do_rescaling:      # lw_transport_1rescl
            if top_at_1:
                for ilev in range(nlay - 1, -1, -1):
                    adj = Cn[:, ilev] * (An[:, ilev] * dn[:, ilev] - trans[:, ilev] * source_dn[:, ilev] - source_up[:, ilev])
                    up[:, ilev] = trans[:, ilev] * up[:, ilev + 1] + source_up[:, ilev] + adj
                    if do_jacobians:
                        gj[:, ilev] = trans[:, ilev] * gj[:, ilev + 1]
                for ilev in range(nlay):
                    adj = Cn[:, ilev] * (An[:, ilev] * up[:, ilev] - trans[:, ilev] * source_up[:, ilev] - source_dn[:, ilev])
                    dn[:, ilev + 1] = trans[:, ilev] * dn[:, ilev] + source_dn[:, ilev] + adj
            else:
                for ilev in range(nlay):
                    adj = Cn[:, ilev] * (An[:, ilev] * dn[:, ilev + 1] - trans[:, ilev] * source_dn[:, ilev] - source_up[:, ilev])
                    up[:, ilev + 1] = trans[:, ilev] * up[:, ilev] + source_up[:, ilev] + adj
                    if do_jacobians:
                        gj[:, ilev + 1] = trans[:, ilev] * gj[:, ilev]
                for ilev in range(nlay - 1, -1, -1):
                    adj = Cn[:, ilev] * (An[:, ilev] * up[:, ilev] - trans[:, ilev] * source_up[:, ilev] - source_dn[:, ilev])
                    dn[:, ilev] = trans[:, ilev] * dn[:, ilev + 1] + source_dn[:, ilev] + adj
        else:                 # lw_transport_noscat_up
            if top_at_1:
                for ilev in range(nlay - 1, -1, -1):
                    up[:, ilev] = trans[:, ilev] * up[:, ilev + 1] + source_up[:, ilev]
                    if do_jacobians:
                        gj[:, ilev] = trans[:, ilev] * gj[:, ilev + 1]
            else:
                for ilev in range(1, nlay + 1):
                    up[:, ilev] = trans[:, ilev - 1] * up[:, ilev - 1] + source_up[:, ilev - 1]
                    if do_jacobians:
                        gj[:, ilev] = trans[:, ilev - 1] * gj[:, ilev - 1]
        if do_broadband:
            bb_up = bb_up + up
            bb_dn = bb_dn + dn
        else:
            flux_dn[:, :, ig] = PI * weight * dn
            flux_up[:, :, ig] = PI * weight * up
        if do_jacobians:
            jac = jac + gj
    if do_broadband:
        bb_up = PI * weight * bb_up
        bb_dn = PI * weight * bb_dn
    if do_jacobians:
        jac = PI * weight * jac
    return flux_up, flux_dn, bb_up, bb_dn, jac


def lw_solver_noscat(top_at_1, Ds, weights, tau, lay_source, lev_source, sfc_emis, sfc_src, inc_flux, do_broadband, do_jacobians,
                     sfc_srcJac, do_rescaling, ssa, g):
    """:248-361.  Ds is (ncol, ngpt, nmus).  Returns flux_up, flux_dn (g-point) , broadband_up, broadband_dn, flux_upJac."""
    out = list(lw_solver_noscat_oneangle(top_at_1, Ds[:, :, 0], weights[0], tau, lay_source, lev_source, sfc_emis, sfc_src, inc_flux,
                                         do_broadband, do_jacobians, sfc_srcJac, do_rescaling, ssa, g))
    for imu in range(1, len(weights)):
        this = lw_solver_noscat_oneangle(top_at_1, Ds[:, :, imu], weights[imu], tau, lay_source, lev_source, sfc_emis, sfc_src,
                                         inc_flux, do_broadband, do_jacobians, sfc_srcJac, do_rescaling, ssa, g)
        if do_broadband:
            out[2] = out[2] + this[2]; out[3] = out[3] + this[3]
        else:
            out[0] = out[0] + this[0]; out[1] = out[1] + this[1]
        if do_jacobians:
            out[4] = out[4] + this[4]
    return out
