"""TEST INFRASTRUCTURE ONLY - a second, independently written restatement of the reference's gas-optics kernels.

Transcribed in vectorised numpy straight from the Fortran
(/root/reference/rrtmgp/kernels/mo_gas_optics_rrtmgp_kernels.F90: interpolation :37-170, compute_tau_absorption
:176-338, gas_optical_depths_major :345-396, gas_optical_depths_minor :402-501, compute_tau_rayleigh :506-565,
compute_Planck_source :568-710, interpolate1D/2D/3D :715-803; mo_cloud_optics_rrtmgp_kernels.F90:24-65), WITHOUT
looking at oracle/rrtmgp_gas_optics_ref.c: different author-pass, different language, different loop structure
(whole-array expressions instead of scalar loops).  tests/test_oracle_crosscheck.py requires the C oracle's parity
build (-O2 -ffp-contract=off) to agree with this file BIT FOR BIT on seeded inputs.  No Fortran compiler exists in the
image, so two independent restatements agreeing exactly is the strongest pin available for the spectral kernels.

Conventions: arrays carry the Fortran shapes, index VALUES are 1-based exactly as in the reference; `F(a)[i-1]`-style
offsets are applied at the point of use.  libm's log/exp are reached through `math` (numpy's SIMD log may differ from
glibc's in the last bit, the C oracle calls glibc).
"""
import math

import numpy as np

TINY = np.finfo(np.float64).tiny


def _aint(x):
    return np.trunc(x)


def interpolation(kd, play, tlay, col_gas):
    """:37-170.  kd needs flavor(2,nflav), press_ref_log, temp_ref, press_ref_log_delta, temp_ref_min, temp_ref_delta,
    press_ref_trop_log, vmr_ref(2,0:ngas,ntemp), neta, npres, ntemp."""
    ncol, nlay = play.shape
    nflav = kd.flavor.shape[1]
    neta, npres, ntemp = kd.neta, kd.npres, kd.ntemp
    press_ref_trop = math.exp(kd.press_ref_trop_log)                       # :99
    temp_ref_delta_inv = 1.0 / kd.temp_ref_delta                           # :100
    press_ref_log_delta_inv = 1.0 / kd.press_ref_log_delta                 # :102
    jtemp_ = ((tlay - (kd.temp_ref_min - kd.temp_ref_delta)) * temp_ref_delta_inv).astype(np.int64)  # INT(): toward zero
    jtemp = np.minimum(ntemp - 1, np.maximum(1, jtemp_))                   # :107
    # :108 reads temp_ref(jtemp_) with the UNclamped index; inside the table's validity range the two coincide, outside
    # the Fortran would read out of bounds - the index is held inside the table here
    ftemp = (tlay - kd.temp_ref[np.clip(jtemp_, 1, ntemp) - 1]) * temp_ref_delta_inv
    logp = np.vectorize(math.log)(play)
    locpress = 1.0 + (logp - kd.press_ref_log[0]) * press_ref_log_delta_inv   # :111
    jpress_aint = np.minimum(float(npres - 1), np.maximum(1.0, _aint(locpress)))
    jpress = jpress_aint.astype(np.int64)
    fpress = locpress - jpress_aint
    tropo = play > press_ref_trop                                          # :117
    itropo = np.where(tropo, 1, 2)

    jeta = np.zeros((2, ncol, nlay, nflav), dtype=np.int64)
    col_mix = np.zeros((2, ncol, nlay, nflav))
    fmajor = np.zeros((2, 2, 2, ncol, nlay, nflav))
    fminor = np.zeros((2, 2, ncol, nlay, nflav))
    for iflav in range(nflav):
        ig1, ig2 = int(kd.flavor[0, iflav]), int(kd.flavor[1, iflav])
        for itemp in (1, 2):
            jt = jtemp + itemp - 1
            ratio = kd.vmr_ref[itropo - 1, ig1, jt - 1] / kd.vmr_ref[itropo - 1, ig2, jt - 1]   # :127-128
            cm = col_gas[:, :, ig1] + ratio * col_gas[:, :, ig2]                                 # :129
            with np.errstate(divide="ignore", invalid="ignore"):
                eta = np.where(cm > 2.0 * TINY, col_gas[:, :, ig1] / cm, 0.5)                      # :147-151
            loceta = eta * float(neta - 1)
            jeta[itemp - 1, :, :, iflav] = np.minimum(loceta.astype(np.int64) + 1, neta - 1)    # :153
            feta = loceta - _aint(loceta)
            ftemp_term = float(2 - itemp) + float(2 * itemp - 3) * ftemp                         # :157
            f1 = (1.0 - feta) * ftemp_term
            f2 = feta * ftemp_term
            col_mix[itemp - 1, :, :, iflav] = cm
            fminor[0, itemp - 1, :, :, iflav] = f1
            fminor[1, itemp - 1, :, :, iflav] = f2
            fmajor[0, 0, itemp - 1, :, :, iflav] = (1.0 - fpress) * f1
            fmajor[1, 0, itemp - 1, :, :, iflav] = (1.0 - fpress) * f2
            fmajor[0, 1, itemp - 1, :, :, iflav] = fpress * f1
            fmajor[1, 1, itemp - 1, :, :, iflav] = fpress * f2
    return dict(jtemp=jtemp, jpress=jpress, tropo=tropo, jeta=jeta, col_mix=col_mix, fmajor=fmajor, fminor=fminor)


def _interp3d(k, scaling, fmajor, jeta, jtemp, jpress, gS, gE):
    """interpolate3D_byflav :765-803 for arrays of cells.  k(ntemp,neta,npres+1,ngpt); scaling (2,n); fmajor (2,2,2,n);
    jeta (2,n); jtemp, jpress (n) 1-based.  Returns (n, gE-gS+1)."""
    g = np.arange(gS - 1, gE)[None, :]
    jt, je1, je2, jp = jtemp[:, None] - 1, jeta[0][:, None] - 1, jeta[1][:, None] - 1, jpress[:, None] - 1
    f = lambda a, b, c: fmajor[a, b, c][:, None]
    return (scaling[0][:, None] *
            (f(0, 0, 0) * k[jt, je1, jp - 1, g] + f(1, 0, 0) * k[jt, je1 + 1, jp - 1, g] +
             f(0, 1, 0) * k[jt, je1, jp, g] + f(1, 1, 0) * k[jt, je1 + 1, jp, g]) +
            scaling[1][:, None] *
            (f(0, 0, 1) * k[jt + 1, je2, jp - 1, g] + f(1, 0, 1) * k[jt + 1, je2 + 1, jp - 1, g] +
             f(0, 1, 1) * k[jt + 1, je2, jp, g] + f(1, 1, 1) * k[jt + 1, je2 + 1, jp, g]))


def _interp2d(k, fminor, jeta, jtemp, kS, kE):
    """interpolate2D_byflav :741-760.  k(ntemp,neta,nk); fminor (2,2,n); jeta (2,n); jtemp (n).  Returns (n, kE-kS+1)."""
    g = np.arange(kS - 1, kE)[None, :]
    jt, je1, je2 = jtemp[:, None] - 1, jeta[0][:, None] - 1, jeta[1][:, None] - 1
    f = lambda a, b: fminor[a, b][:, None]
    return (f(0, 0) * k[jt, je1, g] + f(1, 0) * k[jt, je1 + 1, g] +
            f(0, 1) * k[jt + 1, je2, g] + f(1, 1) * k[jt + 1, je2 + 1, g])


def _by_cell_flavor(arr, iflav_cell):
    """arr(..., ncol, nlay, nflav), iflav_cell(ncol, nlay) 1-based -> (..., ncol*nlay) picking each cell's flavour."""
    ncol, nlay = iflav_cell.shape
    ic, il = np.meshgrid(np.arange(ncol), np.arange(nlay), indexing="ij")
    return arr[..., ic, il, iflav_cell - 1].reshape(arr.shape[:-3] + (ncol * nlay,))


def tau_major(kd, it, tau):
    """gas_optical_depths_major :345-396 (tau is incremented in place)."""
    ncol, nlay, _ = tau.shape
    itropo = np.where(it["tropo"], 1, 2)
    for ibnd in range(kd.band_lims_gpt.shape[1]):
        gS, gE = int(kd.band_lims_gpt[0, ibnd]), int(kd.band_lims_gpt[1, ibnd])
        iflav = kd.gpoint_flavor[itropo - 1, gS - 1]                       # :384, (ncol,nlay)
        res = _interp3d(kd.kmajor, _by_cell_flavor(it["col_mix"], iflav), _by_cell_flavor(it["fmajor"], iflav),
                        _by_cell_flavor(it["jeta"], iflav), it["jtemp"].reshape(-1), (it["jpress"] + itropo).reshape(-1),
                        gS, gE)
        tau[:, :, gS - 1:gE] = tau[:, :, gS - 1:gE] + res.reshape(ncol, nlay, gE - gS + 1)


def layer_limits(play, tropo):
    """:274-285: first/last layer (1-based, inclusive) of the lower and of the upper atmosphere per column;
    minloc/maxloc of an all-false mask give 0."""
    ncol, nlay = play.shape
    top_at_1 = play[0, 0] < play[0, nlay - 1]

    def minloc(mask):
        out = np.zeros(ncol, dtype=np.int64)
        for i in range(ncol):
            idx = np.nonzero(mask[i])[0]
            if idx.size:
                out[i] = idx[np.argmin(play[i, idx])] + 1
        return out

    def maxloc(mask):
        out = np.zeros(ncol, dtype=np.int64)
        for i in range(ncol):
            idx = np.nonzero(mask[i])[0]
            if idx.size:
                out[i] = idx[np.argmax(play[i, idx])] + 1
        return out

    lower, upper = np.zeros((ncol, 2), dtype=np.int64), np.zeros((ncol, 2), dtype=np.int64)
    if top_at_1:
        lower[:, 0], lower[:, 1] = minloc(tropo), nlay
        upper[:, 0], upper[:, 1] = 1, maxloc(~tropo)
    else:
        lower[:, 0], lower[:, 1] = 1, minloc(tropo)
        upper[:, 0], upper[:, 1] = maxloc(~tropo), nlay
    return lower, upper


def tau_minor(kd, which, it, play, tlay, col_gas, limits, tau):
    """gas_optical_depths_minor :402-501 for the 'lower' or 'upper' contributor set."""
    ncol, nlay, _ = tau.shape
    lims = getattr(kd, f"minor_limits_gpt_{which}")
    dens = getattr(kd, f"minor_scales_with_density_{which}")
    comp = getattr(kd, f"scale_by_complement_{which}")
    idx = getattr(kd, f"idx_minor_{which}")
    isc = getattr(kd, f"idx_minor_scaling_{which}")
    kstart = getattr(kd, f"kminor_start_{which}")
    kminor = getattr(kd, f"kminor_{which}")
    gpt_flv = kd.gpoint_flavor[0 if which == "lower" else 1]
    if not np.any(limits[:, 0] > 0):
        return
    lay1 = np.arange(1, nlay + 1)[None, :]
    inside = (limits[:, 0:1] > 0) & (lay1 >= limits[:, 0:1]) & (lay1 <= limits[:, 1:2])   # (ncol, nlay)
    for imnr in range(lims.shape[1]):
        scaling = col_gas[:, :, int(idx[imnr])].copy()                                   # :461
        if dens[imnr]:
            scaling = scaling * (0.01 * play / tlay)                                      # :467
            if int(isc[imnr]) > 0:
                vmr_fact = 1.0 / col_gas[:, :, 0]
                dry_fact = 1.0 / (1.0 + col_gas[:, :, kd.idx_h2o] * vmr_fact)
                if comp[imnr]:
                    scaling = scaling * (1.0 - col_gas[:, :, int(isc[imnr])] * vmr_fact * dry_fact)
                else:
                    scaling = scaling * (col_gas[:, :, int(isc[imnr])] * vmr_fact * dry_fact)
        gS, gE = int(lims[0, imnr]), int(lims[1, imnr])
        iflav = int(gpt_flv[gS - 1])                                                     # :487
        res = _interp2d(kminor, it["fminor"][:, :, :, :, iflav - 1].reshape(2, 2, -1),
                        it["jeta"][:, :, :, iflav - 1].reshape(2, -1), it["jtemp"].reshape(-1),
                        int(kstart[imnr]), int(kstart[imnr]) + (gE - gS)).reshape(ncol, nlay, gE - gS + 1)
        add = scaling[:, :, None] * res
        tau[:, :, gS - 1:gE] = np.where(inside[:, :, None], tau[:, :, gS - 1:gE] + add, tau[:, :, gS - 1:gE])


def compute_tau_absorption(kd, it, play, tlay, col_gas, tau):
    """:176-338"""
    lower, upper = layer_limits(play, it["tropo"])
    tau_major(kd, it, tau)
    tau_minor(kd, "lower", it, play, tlay, col_gas, lower, tau)
    tau_minor(kd, "upper", it, play, tlay, col_gas, upper, tau)
    return tau


def compute_tau_rayleigh(kd, it, col_dry, col_gas):
    """:506-565.  krayl(ntemp,neta,ngpt,2)."""
    ncol, nlay = col_dry.shape
    out = np.zeros((ncol, nlay, kd.krayl.shape[2]))
    itropo = np.where(it["tropo"], 1, 2)
    for ibnd in range(kd.band_lims_gpt.shape[1]):
        gS, gE = int(kd.band_lims_gpt[0, ibnd]), int(kd.band_lims_gpt[1, ibnd])
        iflav = kd.gpoint_flavor[itropo - 1, gS - 1]
        fm, je = _by_cell_flavor(it["fminor"], iflav), _by_cell_flavor(it["jeta"], iflav)
        jt = it["jtemp"].reshape(-1)
        k = np.zeros((ncol * nlay, gE - gS + 1))
        for a in (1, 2):
            sel = itropo.reshape(-1) == a
            if np.any(sel):
                k[sel] = _interp2d(kd.krayl[:, :, :, a - 1], fm[:, :, sel], je[:, sel], jt[sel], gS, gE)
        out[:, :, gS - 1:gE] = k.reshape(ncol, nlay, -1) * (col_gas[:, :, kd.idx_h2o] + col_dry)[:, :, None]   # :559
    return out


def _interp1d(val, offset, delta_r, table):
    """interpolate1D :715-737; table(nT, nbnd) -> (..., nbnd)"""
    val0 = (val - offset) * delta_r
    frac = val0 - _aint(val0)
    index = np.minimum(table.shape[0] - 1, np.maximum(1, val0.astype(np.int64) + 1))
    return table[index - 1] + frac[..., None] * (table[index] - table[index - 1])


def compute_planck_source(kd, it, tlay, tlev, tsfc, sfc_lay):
    """:568-710 -> sfc_src(ncol,ngpt), lay_src(ncol,nlay,ngpt), lev_src(ncol,nlay+1,ngpt), sfc_source_Jac(ncol,ngpt)"""
    ncol, nlay = tlay.shape
    ngpt = kd.planck_frac.shape[3]
    nbnd = kd.band_lims_gpt.shape[1]
    itropo = np.where(it["tropo"], 1, 2)
    pfrac = np.zeros((ncol, nlay, ngpt))
    one = np.ones((2, ncol * nlay))
    for ibnd in range(nbnd):
        gS, gE = int(kd.band_lims_gpt[0, ibnd]), int(kd.band_lims_gpt[1, ibnd])
        iflav = kd.gpoint_flavor[itropo - 1, gS - 1]
        pfrac[:, :, gS - 1:gE] = _interp3d(kd.planck_frac, one, _by_cell_flavor(it["fmajor"], iflav),
                                           _by_cell_flavor(it["jeta"], iflav), it["jtemp"].reshape(-1),
                                           (it["jpress"] + itropo).reshape(-1), gS, gE).reshape(ncol, nlay, -1)
    dr = 1.0 / kd.totplnk_delta
    band_of = np.zeros(ngpt, dtype=np.int64)
    for ibnd in range(nbnd):
        band_of[int(kd.band_lims_gpt[0, ibnd]) - 1:int(kd.band_lims_gpt[1, ibnd])] = ibnd
    p_sfc = _interp1d(tsfc, kd.temp_ref_min, dr, kd.totplnk)              # (ncol, nbnd)
    p_sfc1 = _interp1d(tsfc + 1.0, kd.temp_ref_min, dr, kd.totplnk)
    sfc_src = pfrac[:, sfc_lay - 1, :] * p_sfc[:, band_of]
    sfc_jac = pfrac[:, sfc_lay - 1, :] * (p_sfc1[:, band_of] - p_sfc[:, band_of])
    p_lay = _interp1d(tlay, kd.temp_ref_min, dr, kd.totplnk)               # (ncol, nlay, nbnd)
    lay_src = pfrac * p_lay[:, :, band_of]
    p_lev = _interp1d(tlev, kd.temp_ref_min, dr, kd.totplnk)               # (ncol, nlay+1, nbnd)
    lev_src = np.zeros((ncol, nlay + 1, ngpt))
    lev_src[:, 0] = pfrac[:, 0] * p_lev[:, 0, band_of]
    lev_src[:, 1:nlay] = np.sqrt(pfrac[:, :-1] * pfrac[:, 1:]) * p_lev[:, 1:nlay, band_of]
    lev_src[:, nlay] = pfrac[:, nlay - 1] * p_lev[:, nlay, band_of]
    return sfc_src, lay_src, lev_src, sfc_jac


def compute_cld_from_table(mask, lwp, re, nsteps, step_size, offset, tau_table, ssa_table, asy_table):
    """mo_cloud_optics_rrtmgp_kernels.F90:24-65; tables (nsteps, nbnd)"""
    index = np.minimum(np.floor((re - offset) / step_size).astype(np.int64) + 1, nsteps - 1)
    index = np.where(mask, index, 1)
    fint = (re - offset) / step_size - (index - 1).astype(np.float64)
    f = fint[:, :, None]
    i = index - 1
    t = lwp[:, :, None] * (tau_table[i] + f * (tau_table[i + 1] - tau_table[i]))
    ts = t * (ssa_table[i] + f * (ssa_table[i + 1] - ssa_table[i]))
    tsg = ts * (asy_table[i] + f * (asy_table[i + 1] - asy_table[i]))
    m = mask[:, :, None]
    return np.where(m, t, 0.0), np.where(m, ts, 0.0), np.where(m, tsg, 0.0)
