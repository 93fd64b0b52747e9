"""Seeded synthetic inputs with the reference's schemas.

The real k-distributions, cloud LUTs and RFMIP atmospheres live in un-vendored data tarballs
(rrtmgp-data v1.9.1, reference rrtmgp/CMakeLists.txt:18), so both the CUDA path and the oracle are
driven by tables generated here: same dimensions, same layouts as ty_gas_optics_rrtmgp holds them
AFTER load()/init_abs_coeffs (rrtmgp/frontend/mo_gas_optics_rrtmgp.F90:1151-1381: kmajor as
(ntemp,neta,npres+1,ngpt), krayl (ntemp,neta,ngpt,2), kminor (ntemp,neta,ncontrib), 1-based index
tables).  Physical realism is irrelevant to parity and to bytes moved; values are smooth and positive
so optical depths and fluxes land in realistic ranges.

Workload generators restate examples/all-sky/rrtmgp_allsky.F90:496-662 (compute_profiles,
compute_clouds) including its single-precision literals.
All arrays are numpy, Fortran-ordered, float64 / int32 / bool.
"""
from dataclasses import dataclass, field

import numpy as np

GAS_NAMES = ["h2o", "co2", "o3", "n2o", "co", "ch4", "o2", "n2"]  # rrtmgp_allsky.F90:49


def _f(a, dtype=np.float64):
    return np.asfortranarray(np.asarray(a, dtype=dtype))


@dataclass
class KDist:
    """Kernel-facing state of ty_gas_optics_rrtmgp (mo_gas_optics_rrtmgp.F90:46-155)."""

    is_lw: bool
    gas_names: list
    ngas: int
    nflav: int
    neta: int
    npres: int
    ntemp: int
    nbnd: int
    ngpt: int
    flavor: np.ndarray  # (2,nflav) int32, values 0..ngas
    gpoint_flavor: np.ndarray  # (2,ngpt) int32, 1-based
    band_lims_gpt: np.ndarray  # (2,nbnd) int32, 1-based inclusive
    band_lims_wvn: np.ndarray  # (2,nbnd)
    gpoint_bands: np.ndarray  # (ngpt) int32, 1-based
    press_ref: np.ndarray
    press_ref_log: np.ndarray
    temp_ref: np.ndarray
    press_ref_log_delta: float
    temp_ref_min: float
    temp_ref_max: float
    temp_ref_delta: float
    press_ref_min: float
    press_ref_max: float
    press_ref_trop_log: float
    vmr_ref: np.ndarray  # (2, 0:ngas, ntemp)
    kmajor: np.ndarray  # (ntemp,neta,npres+1,ngpt)
    kminor_lower: np.ndarray  # (ntemp,neta,nminorklower)
    kminor_upper: np.ndarray
    minor_limits_gpt_lower: np.ndarray  # (2,nminorlower) int32
    minor_limits_gpt_upper: np.ndarray
    minor_scales_with_density_lower: np.ndarray  # bool
    minor_scales_with_density_upper: np.ndarray
    scale_by_complement_lower: np.ndarray
    scale_by_complement_upper: np.ndarray
    idx_minor_lower: np.ndarray  # int32 index into col_gas (1..ngas)
    idx_minor_upper: np.ndarray
    idx_minor_scaling_lower: np.ndarray  # int32, 0 = none
    idx_minor_scaling_upper: np.ndarray
    kminor_start_lower: np.ndarray  # int32, 1-based
    kminor_start_upper: np.ndarray
    idx_h2o: int
    # LW
    planck_frac: np.ndarray = None  # (ntemp,neta,npres+1,ngpt)
    totplnk: np.ndarray = None  # (nPlanckTemp, nbnd)
    totplnk_delta: float = 0.0
    # SW
    krayl: np.ndarray = None  # (ntemp,neta,ngpt,2)
    solar_source: np.ndarray = None  # (ngpt)
    extra: dict = field(default_factory=dict)


def _planck_band_integrals(temps, wvn_lims):
    """Band-integrated Planck radiance [W m-2 sr-1] on a temperature grid (stand-in for totplnk)."""
    h, c, kb = 6.626075540e-34, 2.99792458e8, 1.380649e-23
    out = np.zeros((temps.size, wvn_lims.shape[1]))
    for b in range(wvn_lims.shape[1]):
        nu = np.linspace(wvn_lims[0, b], wvn_lims[1, b], 400) * 100.0  # m-1
        x = h * c * nu[None, :] / (kb * temps[:, None])
        B = 2.0 * h * c * c * nu[None, :] ** 3 / np.expm1(x)  # W m-2 sr-1 (m-1)-1
        out[:, b] = np.trapezoid(B, nu, axis=1)
    return out


def _smooth_table(rng, ntemp, neta, npres1, ngpt, lo, hi):
    """Smooth positive table 10**U(lo,hi) varying gently along T, eta, p with a per-g-point level."""
    t = np.linspace(0, 1, ntemp)[:, None, None, None]
    e = np.linspace(0, 1, neta)[None, :, None, None]
    p = np.linspace(0, 1, npres1)[None, None, :, None]
    a = rng.uniform(lo, hi, ngpt)[None, None, None, :]
    bt = rng.uniform(-0.5, 0.5, ngpt)[None, None, None, :]
    be = rng.uniform(-1.0, 1.0, ngpt)[None, None, None, :]
    bp = rng.uniform(-1.0, 1.0, ngpt)[None, None, None, :]
    noise = rng.uniform(-0.05, 0.05, (ntemp, neta, npres1, ngpt))
    return _f(10.0 ** (a + bt * t + be * e * e + bp * np.sin(3.0 * p) + noise))


def _minor(rng, nbnd, band_lims_gpt, ngas, ntemp, neta, max_per_band, lo, hi, partial=False):
    """partial=True: contributor intervals cover a random sub-range of their band (the kernels allow it;
    rrtmgp-data intervals always cover whole bands)."""
    lims, dens, comp, idx, idxs, start = [], [], [], [], [], []
    k = 1
    for b in range(nbnd):
        n = int(rng.integers(0, max_per_band + 1))
        for _ in range(n):
            gs, ge = int(band_lims_gpt[0, b]), int(band_lims_gpt[1, b])
            if partial and ge > gs:
                a, c = sorted(int(v) for v in rng.integers(gs, ge + 1, 2))
                gs, ge = a, c
            lims.append((gs, ge))
            d = bool(rng.integers(0, 2))
            dens.append(d)
            comp.append(bool(rng.integers(0, 2)))
            idx.append(int(rng.integers(1, ngas + 1)))
            idxs.append(int(rng.integers(0, ngas + 1)) if d else 0)
            start.append(k)
            k += ge - gs + 1
    nk = max(k - 1, 1)
    n = len(lims)
    if n == 0:  # keep arrays non-empty so every code path has valid addresses
        lims, dens, comp, idx, idxs, start = [(1, 0)], [False], [False], [1], [0], [1]
        n = 0
    kminor = _f(10.0 ** rng.uniform(lo, hi, (ntemp, neta, nk)))
    return dict(
        n=n,
        kminor=kminor,
        limits=_f(np.array(lims).T.reshape(2, -1), np.int32),
        dens=np.array(dens, dtype=np.bool_),
        comp=np.array(comp, dtype=np.bool_),
        idx=np.array(idx, dtype=np.int32),
        idxs=np.array(idxs, dtype=np.int32),
        start=np.array(start, dtype=np.int32),
    )


def make_kdist(kind="lw", ngpt=None, seed=42, gpt_per_band=None, nminor_per_band=4, band_sizes=None):
    """Synthetic k-distribution.  kind 'lw': 16 bands (256 or 128 g-points), Planck tables;
    kind 'sw': 14 bands (224 or 112 g-points), Rayleigh + solar source.  `gpt_per_band` shrinks it
    for unit tests; `band_sizes` (one entry per band) makes the bands ragged and the minor-contributor
    intervals partial."""
    rng = np.random.default_rng(seed + (0 if kind == "lw" else 1000))
    is_lw = kind == "lw"
    nbnd = 16 if is_lw else 14
    if gpt_per_band is None:
        gpt_per_band = 16 if ngpt is None else ngpt // nbnd
    sizes = np.full(nbnd, gpt_per_band) if band_sizes is None else np.asarray(band_sizes, dtype=int)
    assert sizes.shape == (nbnd,) and sizes.min() >= 1
    ngpt = int(sizes.sum())
    ngas, ntemp, npres, neta = len(GAS_NAMES), 14, 59, 9
    ends = np.cumsum(sizes)
    band_lims_gpt = _f(np.stack([ends - sizes + 1, ends]), np.int32)
    edges = np.linspace(10.0, 3250.0, nbnd + 1) if is_lw else np.linspace(820.0, 50000.0, nbnd + 1)
    band_lims_wvn = _f(np.stack([edges[:-1], edges[1:]]))
    gpoint_bands = np.repeat(np.arange(1, nbnd + 1), sizes).astype(np.int32)
    # key species per (pair, atmos layer, band); (0,0) is rewritten to (2,2) by create_flavor
    # (mo_gas_optics_rrtmgp.F90:1568-1576)
    choices = [(1, 2), (1, 3), (1, 4), (1, 6), (1, 0), (2, 3), (2, 0), (3, 0), (6, 2), (7, 0), (0, 0)]
    key_species = np.zeros((2, 2, nbnd), dtype=np.int32)
    for b in range(nbnd):
        for a in range(2):
            pr = choices[int(rng.integers(0, len(choices)))]
            key_species[:, a, b] = (2, 2) if pr == (0, 0) else pr
    flav = []
    for b in range(nbnd):  # create_flavor :1593-1626
        for a in range(2):
            pr = tuple(int(v) for v in key_species[:, a, b])
            if pr not in flav:
                flav.append(pr)
    flavor = _f(np.array(flav).T, np.int32)
    gpoint_flavor = np.zeros((2, ngpt), dtype=np.int32, order="F")
    for g in range(ngpt):  # create_gpoint_flavor
        for a in range(2):
            gpoint_flavor[a, g] = flav.index(tuple(int(v) for v in key_species[:, a, gpoint_bands[g] - 1])) + 1
    press_ref = np.exp(np.linspace(np.log(109663.31), np.log(1.0), npres))
    temp_ref = 160.0 + 15.0 * np.arange(ntemp)
    vmr_ref = np.zeros((2, ngas + 1, ntemp), order="F")
    base = np.array([1.0, 5e-3, 3.5e-4, 2e-6, 3e-7, 1e-7, 1.7e-6, 0.21, 0.78])
    for it in range(ntemp):
        for a in range(2):
            vmr_ref[a, :, it] = base * rng.uniform(0.5, 1.5, ngas + 1)
    vmr_ref[:, 0, :] = 1.0
    kmajor = _smooth_table(rng, ntemp, neta, npres + 1, ngpt, -26.0, -21.5)
    partial = band_sizes is not None
    lo = _minor(rng, nbnd, band_lims_gpt, ngas, ntemp, neta, nminor_per_band, -27.0, -23.0, partial)
    up = _minor(rng, nbnd, band_lims_gpt, ngas, ntemp, neta, max(nminor_per_band // 2, 1), -27.0, -23.0, partial)
    kd = KDist(
        is_lw=is_lw, gas_names=list(GAS_NAMES), ngas=ngas, nflav=flavor.shape[1], neta=neta, npres=npres,
        ntemp=ntemp, nbnd=nbnd, ngpt=ngpt, flavor=flavor, gpoint_flavor=gpoint_flavor,
        band_lims_gpt=band_lims_gpt, band_lims_wvn=band_lims_wvn, gpoint_bands=gpoint_bands,
        press_ref=press_ref, press_ref_log=np.log(press_ref), temp_ref=temp_ref,
        press_ref_log_delta=float((np.log(press_ref[-1]) - np.log(press_ref[0])) / (npres - 1)),
        temp_ref_min=float(temp_ref[0]), temp_ref_max=float(temp_ref[-1]),
        temp_ref_delta=float((temp_ref[-1] - temp_ref[0]) / (ntemp - 1)),
        press_ref_min=float(press_ref[-1]), press_ref_max=float(press_ref[0]),
        press_ref_trop_log=float(np.log(9948.431564193395)), vmr_ref=vmr_ref, kmajor=kmajor,
        kminor_lower=lo["kminor"], kminor_upper=up["kminor"],
        minor_limits_gpt_lower=lo["limits"], minor_limits_gpt_upper=up["limits"],
        minor_scales_with_density_lower=lo["dens"], minor_scales_with_density_upper=up["dens"],
        scale_by_complement_lower=lo["comp"], scale_by_complement_upper=up["comp"],
        idx_minor_lower=lo["idx"], idx_minor_upper=up["idx"],
        idx_minor_scaling_lower=lo["idxs"], idx_minor_scaling_upper=up["idxs"],
        kminor_start_lower=lo["start"], kminor_start_upper=up["start"], idx_h2o=1,
    )
    kd.extra["nminorlower"], kd.extra["nminorupper"] = lo["n"], up["n"]
    if is_lw:
        pf = rng.uniform(0.2, 1.0, (ntemp, neta, npres + 1, ngpt))
        if band_sizes is None:
            pf *= np.linspace(1.5, 0.5, gpt_per_band)[None, None, None, :].repeat(nbnd, axis=3).reshape(1, 1, 1, ngpt)
        for b in range(nbnd):
            s = slice(int(ends[b] - sizes[b]), int(ends[b]))
            pf[..., s] /= pf[..., s].sum(axis=3, keepdims=True)
        kd.planck_frac = _f(pf)
        tgrid = 160.0 + np.arange(196.0)
        kd.totplnk = _f(_planck_band_integrals(tgrid, band_lims_wvn))
        kd.totplnk_delta = 1.0
        # optimal_angle_fit(2,nbnd) (mo_gas_optics_rrtmgp.F90:1503-1562): secant = fit1*T_column + fit2 >= 1; own generator,
        # so that the tables above keep their values
        r2 = np.random.default_rng(seed + 7700)
        kd.extra["optimal_angle_fit"] = _f(np.stack([r2.uniform(0.0, 0.3, nbnd), r2.uniform(1.4, 1.7, nbnd)]))
    else:
        kd.krayl = _f(10.0 ** rng.uniform(-27.5, -26.0, (ntemp, neta, ngpt, 2)))
        ss = rng.uniform(0.2, 1.0, ngpt)
        kd.solar_source = ss / ss.sum() * 1360.9
    return kd


@dataclass
class CloudLUT:
    """ty_cloud_optics_rrtmgp LUT state (mo_cloud_optics_rrtmgp.F90:77-190), by band."""

    nbnd: int
    band_lims_wvn: np.ndarray
    radliq_lwr: float
    radliq_upr: float
    diamice_lwr: float
    diamice_upr: float
    extliq: np.ndarray  # (nsize_liq, nbnd)
    ssaliq: np.ndarray
    asyliq: np.ndarray
    extice: np.ndarray  # (nsize_ice, nbnd, nrghice)
    ssaice: np.ndarray
    asyice: np.ndarray
    icergh: int = 2  # rrtmgp_allsky.F90:214 set_ice_roughness(2)


def make_cloud_lut(kdist, seed=7):
    rng = np.random.default_rng(seed + (0 if kdist.is_lw else 1))
    nb, nl, ni, nr = kdist.nbnd, 20, 18, 3
    rl = np.linspace(2.5, 21.5, nl)[:, None]
    di = np.linspace(10.0, 180.0, ni)[:, None, None]
    ssa_hi = 0.6 if kdist.is_lw else 0.999
    return CloudLUT(
        nbnd=nb, band_lims_wvn=kdist.band_lims_wvn, radliq_lwr=2.5, radliq_upr=21.5, diamice_lwr=10.0,
        diamice_upr=180.0,
        extliq=_f(0.15 / rl * rng.uniform(0.8, 1.2, (1, nb))), ssaliq=_f(rng.uniform(0.3, ssa_hi, (nl, nb))),
        asyliq=_f(rng.uniform(0.7, 0.9, (nl, nb))),
        extice=_f(0.3 / di * rng.uniform(0.8, 1.2, (1, nb, nr))), ssaice=_f(rng.uniform(0.3, ssa_hi, (ni, nb, nr))),
        asyice=_f(rng.uniform(0.7, 0.9, (ni, nb, nr))),
    )


# ----------------------------------------------------------------------------------------------
def compute_profiles(SST, ncol, nlay):
    """examples/all-sky/rrtmgp_allsky.F90:496-587.  Layer 1 is at the surface (top_at_1 = False).
    The parameters g, Rd, p0, z_q1, z_q2, q_t, gamma, q_0 and the 0.608 / 1. literals are DEFAULT
    REAL (single precision) in the reference (:519-523,525,547-554) and widened to wp."""
    f32 = lambda v: float(np.float32(v))
    z_trop, z_top = 15000.0, 70.0e3
    g1, g2, g3, o3_min = 3.6478, 0.83209, 11.3515, 1e-13
    g, Rd, p0 = f32(9.79764), f32(287.04), f32(101480.0)
    z_q1, z_q2, q_t = f32(4.0e3), f32(7.5e3), f32(1.0e-8)
    gamma, q_0, c608 = f32(6.7e-3), f32(0.01864), f32(0.608)
    Tv0 = (1.0 + c608 * q_0) * SST
    half = nlay // 2
    if nlay % 2 == 0:   # the reference's own split (:514-517): half the layers below 15 km, half above
        z_lev = np.concatenate([[0.0], 2.0 * z_trop / nlay * np.arange(1, half + 1),
                                z_trop + 2.0 * (z_top - z_trop) / nlay * np.arange(1, half + 1)])
    else:               # odd layer counts (IFS 137): the extra layer goes to the upper half
        hi = nlay - half
        z_lev = np.concatenate([[0.0], z_trop / half * np.arange(1, half + 1), z_trop + (z_top - z_trop) / hi * np.arange(1, hi + 1)])
    z_lay = 0.5 * (z_lev[:nlay] + z_lev[1:nlay + 1])

    def prof(z):
        q = np.where(z > z_trop, q_t, q_0 * np.exp(-z / z_q1) * np.exp(-((z / z_q2) ** 2)))
        T = np.where(z > z_trop, SST - gamma * z_trop / (1.0 + c608 * q_0), SST - gamma * z / (1.0 + c608 * q))
        Tv = (1.0 + c608 * q) * T
        p = p0 * (Tv / Tv0) ** (g / (Rd * gamma))
        p = np.where(z > z_trop, p * np.exp(-((g * (z - z_trop)) / (Rd * Tv))), p)
        return p, T, q

    p_l, T_l, q_l = prof(z_lay)
    p_v, T_v, _ = prof(z_lev)
    p_hpa = p_l / 100.0
    o3 = np.maximum(o3_min, g1 * p_hpa**g2 * np.exp(-p_hpa / g3) * 1.0e-6)
    rep = lambda v: _f(np.repeat(v[None, :], ncol, axis=0))
    return dict(p_lay=rep(p_l), t_lay=rep(T_l), p_lev=rep(p_v), t_lev=rep(T_v), q=rep(q_l), o3=rep(o3))


# rrtmgp_allsky.F90:195-203: the well-mixed gases are scalars in gas_concs
ALLSKY_WELL_MIXED = {"co2": 348.0e-6, "ch4": 1650.0e-9, "n2o": 306.0e-9, "n2": 0.7808, "o2": 0.2095, "co": 0.0}


def allsky_gas_vmrs(prof):
    """rrtmgp_allsky.F90:195-203: vmr(ncol,nlay,ngas) in GAS_NAMES order."""
    ncol, nlay = prof["p_lay"].shape
    vmr = np.zeros((ncol, nlay, len(GAS_NAMES)), order="F")
    const = ALLSKY_WELL_MIXED
    for i, name in enumerate(GAS_NAMES):
        if name == "h2o":
            vmr[:, :, i] = prof["q"]
        elif name == "o3":
            vmr[:, :, i] = prof["o3"]
        else:
            vmr[:, :, i] = const[name]
    return vmr


def compute_clouds(prof, lut):
    """rrtmgp_allsky.F90:590-662.  icol is 1-based in mod(icol,3)."""
    p_lay, t_lay = prof["p_lay"], prof["t_lay"]
    ncol, nlay = p_lay.shape
    rel_val = float(np.float32(0.5)) * (lut.radliq_lwr + lut.radliq_upr)
    dei_val = float(np.float32(0.5)) * (lut.diamice_lwr + lut.diamice_upr)
    icol = np.arange(1, ncol + 1)[:, None]
    mask = (p_lay > 100.0 * 100.0) & (p_lay < 900.0 * 100.0) & (icol % 3 != 0)
    lwp = np.where(mask & (t_lay > 263.0), 10.0, 0.0)
    iwp = np.where(mask & (t_lay < 273.0), 10.0, 0.0)
    rel = np.where(lwp > 0.0, rel_val, 0.0)
    dei = np.where(iwp > 0.0, dei_val, 0.0)
    return dict(lwp=_f(lwp), iwp=_f(iwp), rel=_f(rel), dei=_f(dei))


# ----------------------------------------------------------------------------------------------
class AerosolLUT:
    """MERRA aerosol LUT with the loader's schema (mo_optics_utils_rrtmgp.F90:364-397): tables as load_lut()
    receives them, i.e. the rh-dependent ones shaped (nval, nrh, ...)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def make_aerosol_lut(kdist, seed=11, nrh=36, nbin=5):
    rng = np.random.default_rng(seed + (0 if kdist.is_lw else 1))
    nb, nval = kdist.nbnd, 3
    lims = np.array([[0.1, 1.0], [1.0, 1.8], [1.8, 3.0], [3.0, 6.0], [6.0, 10.0]][:nbin]).T  # (pair, nbin), microns
    rh = np.concatenate([np.linspace(0.0, 0.8, nrh - 10, endpoint=False), np.linspace(0.8, 0.99, 10)])
    ssa_hi = 0.5 if kdist.is_lw else 0.99

    def tbl(*shape):
        t = np.empty((nval,) + shape)
        t[0] = rng.uniform(50.0, 5000.0, shape)   # mass extinction coefficient (m2/kg)
        t[1] = rng.uniform(0.2, ssa_hi, shape)    # single-scattering albedo
        t[2] = rng.uniform(0.3, 0.85, shape)      # asymmetry
        return _f(t)

    def tbl_rh(*shape):  # smooth growth with humidity so interpolation is exercised
        base = tbl(*shape)  # (nval, ...)
        growth = 1.0 + np.array([1.5, 0.02, 0.1])[:, None] * rh[None, :] ** 2  # (nval, nrh)
        out = base[:, None, ...] * growth.reshape((nval, nrh) + (1,) * len(shape))
        out[1] = np.minimum(out[1], 0.999)
        return _f(out)

    return AerosolLUT(nbnd=nb, nval=nval, nrh=nrh, nbin=nbin, band_lims_wvn=kdist.band_lims_wvn,
                      merra_aero_bin_lims=_f(lims), aero_rh=_f(rh), aero_dust_tbl=tbl(nbin, nb),
                      aero_salt_tbl=tbl_rh(nbin, nb), aero_sulf_tbl=tbl_rh(nb), aero_bcar_tbl=tbl(nb),
                      aero_bcar_rh_tbl=tbl_rh(nb), aero_ocar_tbl=tbl(nb), aero_ocar_rh_tbl=tbl_rh(nb))


def get_relhum(p_lay, t_lay, vmr_h2o):
    """rrtmgp_allsky.F90:744-784 (m_h2o, m_dry from mo_gas_optics_constants.F90)."""
    m_h2o, m_dry = 0.018016, 0.028964
    mwd, t_ref, q_lay_min = m_h2o / m_dry, 273.16, 1.0e-7
    mmr = vmr_h2o * mwd
    q_lay = mmr / (1 + mmr)
    q_tmp = np.maximum(q_lay_min, q_lay)
    es_tmp = np.exp((17.67 * (t_lay - t_ref)) / (t_lay - 29.65))
    rh = (0.263 * p_lay * q_tmp) / es_tmp
    return _f(0.01 * rh)


def compute_aerosols(prof, col_offset=0):
    """rrtmgp_allsky.F90:664-738: sulfate (type 3) at 50-100 hPa, dust (type 1) at 700-900 hPa, in the columns
    with odd 1-based index (the reference's `is_even_column = mod(icol,2) /= 0`)."""
    p_lay = prof["p_lay"]
    ncol, nlay = p_lay.shape
    icol = np.arange(1 + col_offset, ncol + 1 + col_offset)[:, None]
    sel = icol % 2 != 0
    is_sulf = (p_lay > 50.0 * 100.0) & (p_lay < 100.0 * 100.0) & sel
    is_dust = (p_lay > 700.0 * 100.0) & (p_lay < 900.0 * 100.0) & sel & ~is_sulf
    aero_type = np.where(is_sulf, 3, np.where(is_dust, 1, 0)).astype(np.int32)
    aero_size = np.where(is_sulf, 0.2, np.where(is_dust, 0.5, 0.0))
    aero_mass = np.where(is_sulf, 1.0e-6, np.where(is_dust, 3.0e-5, 0.0))
    relhum = get_relhum(p_lay, prof["t_lay"], prof["q"])
    return dict(aero_type=np.asfortranarray(aero_type), aero_size=_f(aero_size), aero_mass=_f(aero_mass), relhum=relhum)


def perturbed_profiles(ncol, nlay, seed=1234, top_at_1=True):
    """RFMIP-like stand-in: `ncol` DISTINCT columns = the analytic profile with per-column
    SST in [285,310] K and humidity/ozone scaled by 0.5-2x, so neighbouring columns hit different table
    cells (the replicated profile is a pure broadcast).  Optionally flipped so layer 1 is the top."""
    rng = np.random.default_rng(seed)
    base = compute_profiles(300.0, 1, nlay)
    sst = rng.uniform(285.0, 310.0, ncol)
    out = {k: np.zeros((ncol,) + v.shape[1:], order="F") for k, v in base.items()}
    for k in out:
        out[k][:] = base[k][0]
    dT = (sst - 300.0)[:, None]
    out["t_lay"] = _f(out["t_lay"] + dT * (out["p_lay"] / out["p_lay"][:, :1]))
    out["t_lev"] = _f(out["t_lev"] + dT * (out["p_lev"] / out["p_lev"][:, :1]))
    out["q"] = _f(out["q"] * rng.uniform(0.5, 2.0, ncol)[:, None])
    out["o3"] = _f(out["o3"] * rng.uniform(0.5, 2.0, ncol)[:, None])
    ps = rng.uniform(0.93, 1.0, ncol)[:, None]  # surface-pressure spread (terrain)
    out["p_lay"] = _f(out["p_lay"] * ps)
    out["p_lev"] = _f(out["p_lev"] * ps)
    out["p_lev"][:, -1] = np.maximum(out["p_lev"][:, -1], 1.0 + 1e-6)
    if top_at_1:
        out = {k: _f(v[:, ::-1]) for k, v in out.items()}
    return out


# ----------------------------------------------------------------------------------------------
# Raw ("on-disk") k-distribution: the variable set of an rrtmgp-data file before ty_gas_optics_rrtmgp%load
# (rrtmgp/data-loading-examples/mo_optics_utils_rrtmgp.F90:102-183), synthesised so that load() restricted to the
# host's gases gives back exactly the tables of a KDist built by make_kdist().
# ----------------------------------------------------------------------------------------------
RAW_EXTRA_GASES = ["ccl4", "cfc11", "cfc12", "cfc22", "hfc143a", "hfc125", "hfc23", "hfc32", "hfc134a", "cf4", "no2"]


def make_kdist_raw(kd, seed=11):
    """dict of numpy arrays / string lists with the on-disk names and Fortran shapes (first index fastest):
    19 absorbers (the 8 of `kd` in their order, 11 others interleaved), key_species indexing THAT list, minor-contributor
    intervals of absent gases interleaved with kd's, string-valued minor_gases / scaling_gas, kmajor / plank_fraction as
    (gpt, mixing_fraction, pressure+1, temperature), rayl_lower/upper (gpt, mixing_fraction, temperature)."""
    rng = np.random.default_rng(seed)
    ngas = kd.ngas
    # --- absorber list: kd's gases keep their relative order (the reduced list must come out as kd.gas_names)
    slots = sorted(rng.choice(ngas + len(RAW_EXTRA_GASES), ngas, replace=False).tolist())
    names, host_pos, extra = [], {}, iter(RAW_EXTRA_GASES)
    for i in range(ngas + len(RAW_EXTRA_GASES)):
        if i in slots:
            g = kd.gas_names[slots.index(i)]
            host_pos[g] = len(names) + 1
            names.append(g.upper() if len(names) % 3 == 0 else g + "  ")   # case and blank padding must not matter
        else:
            names.append(next(extra))
    nabs = len(names)
    raw_of_host = [0] + [host_pos[g] for g in kd.gas_names]   # reduced index -> raw index
    # --- key_species(2, atmos_layer, bnd): kd's flavours; (2,2) is what load() makes of an on-disk (0,0)
    key_species = np.zeros((2, 2, kd.nbnd), dtype=np.int32, order="F")
    for b in range(kd.nbnd):
        g1 = int(kd.band_lims_gpt[0, b]) - 1
        for a in range(2):
            pr = kd.flavor[:, int(kd.gpoint_flavor[a, g1]) - 1]
            key_species[:, a, b] = (0, 0) if tuple(pr) == (2, 2) else [raw_of_host[int(v)] for v in pr]
    vmr_ref = np.asfortranarray(rng.uniform(1e-9, 1e-3, (2, nabs + 1, kd.ntemp)))
    for i in range(ngas + 1):
        vmr_ref[:, raw_of_host[i], :] = kd.vmr_ref[:, i, :]
    # --- minor contributors
    ident = [g if g != "h2o" else "h2o_frgn" for g in kd.gas_names] + ["h2o_self"] + RAW_EXTRA_GASES
    gas_minor = list(kd.gas_names) + ["h2o"] + RAW_EXTRA_GASES

    def minor(which):
        lims = getattr(kd, f"minor_limits_gpt_{which}")
        n = kd.extra[f"nminor{which}"]
        kmin = getattr(kd, f"kminor_{which}")          # (ntemp, neta, nk)
        out = dict(gases=[], scaling=[], lims=[], dens=[], comp=[], start=[], cols=[])
        k = 1

        def add(gas_ident, scal, gs, ge, dens, comp, cols):
            nonlocal k
            out["gases"].append(gas_ident); out["scaling"].append(scal); out["lims"].append((gs, ge))
            out["dens"].append(dens); out["comp"].append(comp); out["start"].append(k); out["cols"].append(cols)
            k += ge - gs + 1

        def absent():
            b = int(rng.integers(0, kd.nbnd))
            gs, ge = int(kd.band_lims_gpt[0, b]), int(kd.band_lims_gpt[1, b])
            add(RAW_EXTRA_GASES[int(rng.integers(0, len(RAW_EXTRA_GASES)))], "", gs, ge, bool(rng.integers(0, 2)), False,
                10.0 ** rng.uniform(-27, -23, (kd.ntemp, kd.neta, ge - gs + 1)))

        for i in range(n):
            while rng.random() < 0.35:
                absent()
            gs, ge = int(lims[0, i]), int(lims[1, i])
            g = kd.gas_names[int(getattr(kd, f"idx_minor_{which}")[i]) - 1]
            idn = g if g != "h2o" else ("h2o_self" if rng.random() < 0.5 else "h2o_frgn")
            isc = int(getattr(kd, f"idx_minor_scaling_{which}")[i])
            ks = int(getattr(kd, f"kminor_start_{which}")[i])
            add(idn.upper() if i % 4 == 0 else idn, kd.gas_names[isc - 1] if isc > 0 else "",
                gs, ge, bool(getattr(kd, f"minor_scales_with_density_{which}")[i]),
                bool(getattr(kd, f"scale_by_complement_{which}")[i]), kmin[:, :, ks - 1:ks - 1 + ge - gs + 1])
        absent()
        ncontrib = k - 1
        kraw = np.zeros((ncontrib, kd.neta, kd.ntemp), order="F")
        for s, c in zip(out["start"], out["cols"]):
            kraw[s - 1:s - 1 + c.shape[2]] = np.transpose(c, (2, 1, 0))
        return dict(kminor=kraw, gases=out["gases"], scaling=out["scaling"],
                    limits=_f(np.array(out["lims"]).T.reshape(2, -1), np.int32), dens=np.array(out["dens"], dtype=np.bool_),
                    comp=np.array(out["comp"], dtype=np.bool_), start=np.array(out["start"], dtype=np.int32))

    lo, up = minor("lower"), minor("upper")
    raw = dict(
        gas_names=names, key_species=key_species, bnd_limits_wavenumber=_f(kd.band_lims_wvn),
        bnd_limits_gpt=_f(kd.band_lims_gpt, np.int32), press_ref=np.array(kd.press_ref), temp_ref=np.array(kd.temp_ref),
        absorption_coefficient_ref_P=101325.0, absorption_coefficient_ref_T=296.0,
        press_ref_trop=float(np.exp(kd.press_ref_trop_log)), kminor_lower=lo["kminor"], kminor_upper=up["kminor"],
        gas_minor=gas_minor, identifier_minor=ident, minor_gases_lower=lo["gases"], minor_gases_upper=up["gases"],
        minor_limits_gpt_lower=lo["limits"], minor_limits_gpt_upper=up["limits"],
        minor_scales_with_density_lower=lo["dens"], minor_scales_with_density_upper=up["dens"],
        scale_by_complement_lower=lo["comp"], scale_by_complement_upper=up["comp"],
        scaling_gas_lower=lo["scaling"], scaling_gas_upper=up["scaling"],
        kminor_start_lower=lo["start"], kminor_start_upper=up["start"], vmr_ref=vmr_ref,
        kmajor=_f(np.transpose(kd.kmajor, (3, 1, 2, 0))),
    )
    if kd.is_lw:
        raw.update(totplnk=_f(kd.totplnk), plank_fraction=_f(np.transpose(kd.planck_frac, (3, 1, 2, 0))),
                   optimal_angle_fit=_f(kd.extra["optimal_angle_fit"]))
    else:
        raw.update(rayl_lower=_f(np.transpose(kd.krayl[:, :, :, 0], (2, 1, 0))),
                   rayl_upper=_f(np.transpose(kd.krayl[:, :, :, 1], (2, 1, 0))))
        # solar_source = quiet + (mg - 0.1495954) facular + (sb - 0.00066696) sunspot (mo_gas_optics_rrtmgp.F90:788-792)
        fac, spot = rng.uniform(0.0, 0.05, kd.ngpt) * kd.solar_source, rng.uniform(-0.05, 0.0, kd.ngpt) * kd.solar_source
        mg, sb = 0.1567652, 902.71260
        raw.update(solar_source_quiet=kd.solar_source - (mg - 0.1495954) * fac - (sb - 0.00066696) * spot,
                   solar_source_facular=fac, solar_source_sunspot=spot, tsi_default=1360.85767381726, mg_default=mg,
                   sb_default=sb)
    return raw
