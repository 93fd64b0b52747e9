"""Python view of the C frontend (include/rrtmgp_b200_frontend.h): thin ctypes wrappers so tests and the
benchmark read like the reference's Fortran programs (ty_optical_props_*, ty_source_func_lw,
ty_fluxes_broadband, ty_gas_optics_rrtmgp, ty_cloud_optics_rrtmgp, rte_lw, rte_sw).

All numerics live behind the C ABI; this module only owns memory (numpy on the host backend, torch on
CUDA) and marshals pointers.  Every call raises RuntimeError with the reference's error string when the
C frontend reports one (the Fortran functions return that string; empty = success).
"""
import ctypes as C

import numpy as np

from .abi import FLOAT, fzeros, to_device, to_host, _ptr

ERRLEN = 128
K1SCL, K2STR, KNSTR = 1, 2, 3
PF = C.POINTER(FLOAT)


class _OpStruct(C.Structure):
    _fields_ = [("kind", C.c_int), ("ncol", C.c_int), ("nlay", C.c_int), ("ngpt", C.c_int), ("nband", C.c_int),
                ("nmom", C.c_int), ("top_at_1", C.c_int), ("band_lims_gpt", C.c_void_p),
                ("band_lims_wvn", C.c_void_p), ("tau", C.c_void_p), ("ssa", C.c_void_p), ("g", C.c_void_p),
                ("p", C.c_void_p)]


class _SrcStruct(C.Structure):
    _fields_ = [("ncol", C.c_int), ("nlay", C.c_int), ("ngpt", C.c_int), ("lay_source", C.c_void_p),
                ("lev_source", C.c_void_p), ("sfc_source", C.c_void_p), ("sfc_source_Jac", C.c_void_p)]


class _FluxStruct(C.Structure):
    _fields_ = [("flux_up", C.c_void_p), ("flux_dn", C.c_void_p), ("flux_net", C.c_void_p),
                ("flux_dn_dir", C.c_void_p)]


class _KDistStruct(C.Structure):
    _fields_ = (
        [(n, C.c_int) for n in ("ngas", "nflav", "neta", "npres", "ntemp", "nbnd", "ngpt", "nminorlower",
                                "nminorklower", "nminorupper", "nminorkupper", "idx_h2o")]
        + [(n, C.c_void_p) for n in ("flavor", "gpoint_flavor", "band_lims_gpt", "gpoint_bands", "band_lims_wvn",
                                     "press_ref_log", "temp_ref", "vmr_ref")]
        + [(n, FLOAT) for n in ("press_ref_log_delta", "temp_ref_min", "temp_ref_max", "temp_ref_delta",
                                "press_ref_min", "press_ref_max", "press_ref_trop_log")]
        + [(n, C.c_void_p) for n in ("kmajor", "kminor_lower", "kminor_upper", "minor_limits_gpt_lower",
                                     "minor_limits_gpt_upper", "minor_scales_with_density_lower",
                                     "minor_scales_with_density_upper", "scale_by_complement_lower",
                                     "scale_by_complement_upper", "idx_minor_lower", "idx_minor_upper",
                                     "idx_minor_scaling_lower", "idx_minor_scaling_upper", "kminor_start_lower",
                                     "kminor_start_upper", "planck_frac", "totplnk")]
        + [("nPlanckTemp", C.c_int), ("totplnk_delta", FLOAT), ("krayl", C.c_void_p), ("solar_source", C.c_void_p)]
    )


class _CloudLutStruct(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("nbnd", "nsize_liq", "nsize_ice", "nrghice", "icergh")]
                + [("band_lims_gpt", C.c_void_p), ("band_lims_wvn", C.c_void_p)]
                + [(n, FLOAT) for n in ("radliq_lwr", "radliq_upr", "diamice_lwr", "diamice_upr")]
                + [(n, C.c_void_p) for n in ("extliq", "ssaliq", "asyliq", "extice", "ssaice", "asyice")])


class _AerosolLutStruct(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("nbnd", "nval", "nrh", "nbin")]
                + [(n, C.c_void_p) for n in ("band_lims_wvn", "merra_aero_bin_lims", "aero_rh", "aero_dust_tbl",
                                             "aero_salt_tbl", "aero_sulf_tbl", "aero_bcar_tbl", "aero_bcar_rh_tbl",
                                             "aero_ocar_tbl", "aero_ocar_rh_tbl")])


def _addr(x):
    return None if x is None else _ptr(x).value


def _check(rc, err):
    if rc:
        raise RuntimeError(err.value.decode())


class Context:
    """A kernel library + the memory space its arrays live in (device=None: host numpy)."""

    def __init__(self, lib, device=None):
        self.lib, self.device, self.c = lib, device, lib.cdll
        for fn in ("rrtmgpb_gas_optics_load", "rrtmgpb_cloud_optics_load", "rrtmgpb_aerosol_optics_load"):
            getattr(self.c, fn).restype = C.c_void_p

    def zeros(self, shape, dtype=np.float64):
        return fzeros(shape, dtype=dtype, device=self.device)

    def put(self, a):
        a = np.asfortranarray(a)
        return a if self.device is None else to_device(a, self.device)

    def get(self, a):
        """Host copy (always a fresh array, so results of successive calls never alias)."""
        self.lib.sync()
        return np.array(to_host(a), order="F", copy=True)

    def config_checks(self, extents=True, values=True):
        """rte_config_checks(), rte/frontend/mo_rte_config.F90:29-49"""
        self.c.rrtmgpb_rte_config_checks(int(extents), int(values))


class OpticalProps:
    """ty_optical_props_1scl / _2str / _nstr (rte/frontend/mo_optical_props.F90)."""

    def __init__(self, ctx, kind, ncol, nlay, band_lims_gpt, band_lims_wvn=None, nmom=0, top_at_1=True, name=""):
        self.ctx, self.kind, self.name = ctx, {"1scl": K1SCL, "2str": K2STR, "nstr": KNSTR}.get(kind, kind), name
        self.band_lims_gpt = np.asfortranarray(band_lims_gpt, dtype=np.int32)
        self.band_lims_wvn = None if band_lims_wvn is None else np.asfortranarray(band_lims_wvn, dtype=np.float64)
        self.nband = self.band_lims_gpt.shape[1]
        self.ngpt = int(self.band_lims_gpt.max())
        self.ncol, self.nlay, self.nmom, self.top_at_1 = ncol, nlay, nmom, top_at_1
        shp = (ncol, nlay, self.ngpt)
        self.tau = ctx.zeros(shp)
        self.ssa = ctx.zeros(shp) if self.kind != K1SCL else None
        self.g = ctx.zeros(shp) if self.kind == K2STR else None
        self.p = ctx.zeros((nmom,) + shp) if self.kind == KNSTR else None

    @classmethod
    def like(cls, ctx, kind, ncol, nlay, spectral, **kw):
        """alloc_*(ncol, nlay, spectral_desc): copy the spectral discretisation of another object."""
        return cls(ctx, kind, ncol, nlay, spectral.band_lims_gpt, spectral.band_lims_wvn, **kw)

    def struct(self):
        s = _OpStruct(self.kind, self.ncol, self.nlay, self.ngpt, self.nband, self.nmom, int(self.top_at_1),
                      self.band_lims_gpt.ctypes.data,
                      None if self.band_lims_wvn is None else self.band_lims_wvn.ctypes.data, _addr(self.tau),
                      _addr(self.ssa), _addr(self.g), _addr(self.p))
        return s

    def _sync_back(self, s):
        self.top_at_1 = bool(s.top_at_1)

    def set_top_at_1(self, v):
        self.top_at_1 = bool(v)

    def validate(self):
        err = C.create_string_buffer(ERRLEN)
        s = self.struct()
        _check(self.ctx.c.rrtmgpb_op_validate(C.byref(s), err), err)

    def delta_scale(self, forward=None):
        err = C.create_string_buffer(ERRLEN)
        s = self.struct()
        _check(self.ctx.c.rrtmgpb_op_delta_scale(C.byref(s), C.c_void_p(_addr(forward)), err), err)

    def increment(self, op_io):
        """call self%increment(op_io): op_io is incremented by self (mo_optical_props.F90:879)."""
        err = C.create_string_buffer(ERRLEN)
        a, b = self.struct(), op_io.struct()
        _check(self.ctx.c.rrtmgpb_op_increment(C.byref(a), C.byref(b), err), err)


class SourceFuncLW:
    """ty_source_func_lw (rte/frontend/mo_source_functions.F90:30-38)."""

    def __init__(self, ctx, ncol, nlay, ngpt):
        self.ctx, self.ncol, self.nlay, self.ngpt = ctx, ncol, nlay, ngpt
        self.lay_source = ctx.zeros((ncol, nlay, ngpt))
        self.lev_source = ctx.zeros((ncol, nlay + 1, ngpt))
        self.sfc_source = ctx.zeros((ncol, ngpt))
        self.sfc_source_Jac = ctx.zeros((ncol, ngpt))

    def struct(self):
        return _SrcStruct(self.ncol, self.nlay, self.ngpt, _addr(self.lay_source), _addr(self.lev_source),
                          _addr(self.sfc_source), _addr(self.sfc_source_Jac))


class FluxesBroadband:
    """ty_fluxes_broadband (rte/frontend/mo_fluxes.F90:47-54); None = pointer not associated."""

    def __init__(self, flux_up=None, flux_dn=None, flux_net=None, flux_dn_dir=None):
        self.flux_up, self.flux_dn, self.flux_net, self.flux_dn_dir = flux_up, flux_dn, flux_net, flux_dn_dir

    def struct(self):
        return _FluxStruct(_addr(self.flux_up), _addr(self.flux_dn), _addr(self.flux_net), _addr(self.flux_dn_dir))


def rte_lw(ctx, optical_props, sources, sfc_emis, fluxes, inc_flux=None, n_gauss_angles=0, use_2stream=-1,
           lw_Ds=None, flux_up_Jac=None):
    """rte_lw(), rte/frontend/mo_rte_lw.F90:79.  sfc_emis (nband,ncol)."""
    err = C.create_string_buffer(ERRLEN)
    o, s, f = optical_props.struct(), sources.struct(), fluxes.struct()
    _check(ctx.c.rrtmgpb_rte_lw(C.byref(o), C.byref(s), C.c_void_p(_addr(sfc_emis)), C.byref(f),
                                C.c_void_p(_addr(inc_flux)), int(n_gauss_angles), int(use_2stream),
                                C.c_void_p(_addr(lw_Ds)), C.c_void_p(_addr(flux_up_Jac)), err), err)


def rte_lw_bygpoint(ctx, optical_props, sources, sfc_emis, gpt_flux_up, gpt_flux_dn, inc_flux=None,
                    n_gauss_angles=0, use_2stream=-1, lw_Ds=None):
    err = C.create_string_buffer(ERRLEN)
    o, s = optical_props.struct(), sources.struct()
    _check(ctx.c.rrtmgpb_rte_lw_bygpoint(C.byref(o), C.byref(s), C.c_void_p(_addr(sfc_emis)),
                                         C.c_void_p(_addr(gpt_flux_up)), C.c_void_p(_addr(gpt_flux_dn)),
                                         C.c_void_p(_addr(inc_flux)), int(n_gauss_angles), int(use_2stream),
                                         C.c_void_p(_addr(lw_Ds)), err), err)


def rte_sw(ctx, atmos, mu0, inc_flux, sfc_alb_dir, sfc_alb_dif, fluxes, inc_flux_dif=None):
    """rte_sw() with mu0 by column, rte/frontend/mo_rte_sw.F90:56."""
    err = C.create_string_buffer(ERRLEN)
    o, f = atmos.struct(), fluxes.struct()
    _check(ctx.c.rrtmgpb_rte_sw(C.byref(o), C.c_void_p(_addr(mu0)), C.c_void_p(_addr(inc_flux)),
                                C.c_void_p(_addr(sfc_alb_dir)), C.c_void_p(_addr(sfc_alb_dif)), C.byref(f),
                                C.c_void_p(_addr(inc_flux_dif)), err), err)


def rte_sw_bygpoint(ctx, atmos, mu0, inc_flux, sfc_alb_dir, sfc_alb_dif, gpt_up, gpt_dn, gpt_dir, inc_flux_dif=None):
    err = C.create_string_buffer(ERRLEN)
    o = atmos.struct()
    _check(ctx.c.rrtmgpb_rte_sw_bygpoint(C.byref(o), C.c_void_p(_addr(mu0)), C.c_void_p(_addr(inc_flux)),
                                         C.c_void_p(_addr(sfc_alb_dir)), C.c_void_p(_addr(sfc_alb_dif)),
                                         C.c_void_p(_addr(gpt_up)), C.c_void_p(_addr(gpt_dn)),
                                         C.c_void_p(_addr(gpt_dir)), C.c_void_p(_addr(inc_flux_dif)), err), err)


def rte_lw_express(ctx, gas_optics, p_lay, p_lev, t_lay, t_sfc, vmr, sfc_emis, fluxes, clouds=None, col_dry=None,
                   tlev=None, n_gauss_angles=0):
    """Express path: gas_optics + clouds%increment + rte_lw -> broadband fluxes, no (ncol,nlay,ngpt) array
    (rrtmgpb_rte_lw_express, include/rrtmgp_b200_frontend.h).  vmr(ncol,nlay,ngas); sfc_emis(nband,ncol)."""
    err = C.create_string_buffer(ERRLEN)
    ncol, nlay = (int(v) for v in p_lay.shape)
    P = lambda x: C.c_void_p(_addr(x))
    f = fluxes.struct()
    o = clouds.struct() if clouds is not None else None
    _check(ctx.c.rrtmgpb_rte_lw_express(C.c_void_p(gas_optics.handle), ncol, nlay, P(p_lay), P(p_lev), P(t_lay), P(t_sfc),
                                        P(vmr), P(col_dry), P(tlev), C.byref(o) if o is not None else None, P(sfc_emis),
                                        int(n_gauss_angles), C.byref(f), err), err)


def rte_sw_express(ctx, gas_optics, p_lay, p_lev, t_lay, vmr, mu0, sfc_alb_dir, sfc_alb_dif, fluxes, clouds=None,
                   col_dry=None):
    """Express path, shortwave: gas_optics + clouds%increment (clouds already delta-scaled) + rte_sw -> broadband fluxes."""
    err = C.create_string_buffer(ERRLEN)
    ncol, nlay = (int(v) for v in p_lay.shape)
    P = lambda x: C.c_void_p(_addr(x))
    f = fluxes.struct()
    o = clouds.struct() if clouds is not None else None
    _check(ctx.c.rrtmgpb_rte_sw_express(C.c_void_p(gas_optics.handle), ncol, nlay, P(p_lay), P(p_lev), P(t_lay), P(vmr),
                                        P(col_dry), C.byref(o) if o is not None else None, P(mu0), P(sfc_alb_dir),
                                        P(sfc_alb_dif), C.byref(f), err), err)


# ---- McICA cloud sampling: rte/extensions/mo_cloud_sampling.F90 (C++ mirror: rrtmgpb_cloud_sampling_*) ----------------
def _sampled_mask(ctx, randoms, cloud_frac, overlap_param, cloud_mask):
    ngpt, nlay, ncol = (int(v) for v in randoms.shape)
    if cloud_mask is None:
        cloud_mask = ctx.zeros((ncol, nlay, ngpt), dtype=np.bool_)
    P = lambda x: C.c_void_p(_addr(x))
    err = C.create_string_buffer(ERRLEN)
    cf, m = [int(v) for v in cloud_frac.shape], [int(v) for v in cloud_mask.shape]
    if overlap_param is None:
        rc = ctx.c.rrtmgpb_cloud_sampling_mask_max_ran(ngpt, nlay, ncol, P(randoms), cf[0], cf[1], P(cloud_frac), m[0], m[1],
                                                       m[2], P(cloud_mask), err)
    else:
        op = [int(v) for v in overlap_param.shape]
        rc = ctx.c.rrtmgpb_cloud_sampling_mask_exp_ran(ngpt, nlay, ncol, P(randoms), cf[0], cf[1], P(cloud_frac), op[0], op[1],
                                                       P(overlap_param), m[0], m[1], m[2], P(cloud_mask), err)
    _check(rc, err)
    return cloud_mask


def sampled_mask_max_ran(ctx, randoms, cloud_frac, cloud_mask=None):
    """mo_cloud_sampling.F90:125-192: McICA mask for maximum-random overlap.  randoms(ngpt,nlay,ncol)."""
    return _sampled_mask(ctx, randoms, cloud_frac, None, cloud_mask)


def sampled_mask_exp_ran(ctx, randoms, cloud_frac, overlap_param, cloud_mask=None):
    """mo_cloud_sampling.F90:205-292: McICA mask for exponential-random overlap; overlap_param(ncol,nlay-1)."""
    return _sampled_mask(ctx, randoms, cloud_frac, overlap_param, cloud_mask)


def draw_samples(ctx, cloud_mask, clouds, clouds_sampled):
    """mo_cloud_sampling.F90:36-120: by-band cloud properties -> by-g-point properties, zero where the mask is false."""
    err = C.create_string_buffer(ERRLEN)
    a, b = clouds.struct(), clouds_sampled.struct()
    m = [int(v) for v in cloud_mask.shape]
    _check(ctx.c.rrtmgpb_cloud_sampling_draw_samples(m[0], m[1], m[2], C.c_void_p(_addr(cloud_mask)), C.byref(a), C.byref(b),
                                                     err), err)


class GasConcs:
    """ty_gas_concs (rte/frontend/gas-optics-template/mo_gas_concentrations.F90) through its C++ mirror
    (rrtmgpb_gc_*, include/rrtmgp_b200_frontend.h): concentrations by gas name, each stored as a scalar (1,1), a
    profile (1,nlay) or a field (ncol,nlay) - set_vmr_scalar/_1d/_2d :129-305 - and broadcast on demand by get_vmr
    (:433-504).  A host model that keeps its well-mixed gases as scalars (the all-sky example:
    rrtmgp_allsky.F90:195-203) only ever hands the fields of h2o and o3 to the device."""

    def __init__(self, ctx, gas_names):
        self.ctx = ctx
        ctx.c.rrtmgpb_gc_init.restype = C.c_void_p
        ctx.c.rrtmgpb_gc_set_vmr_scalar.argtypes = [C.c_void_p, C.c_char_p, FLOAT, C.c_char_p]
        names = (C.c_char_p * len(gas_names))(*[g.encode() for g in gas_names])
        err = C.create_string_buffer(ERRLEN)
        self.handle = ctx.c.rrtmgpb_gc_init(len(gas_names), names, err)
        if not self.handle:
            raise RuntimeError(err.value.decode())

    def set_vmr(self, gas, w):
        err = C.create_string_buffer(ERRLEN)
        h, name = C.c_void_p(self.handle), gas.encode()
        if np.isscalar(w):
            _check(self.ctx.c.rrtmgpb_gc_set_vmr_scalar(h, name, FLOAT(float(w)), err), err)
            return
        temporary = isinstance(w, np.ndarray) and self.ctx.device is not None
        if isinstance(w, np.ndarray):
            w = self.ctx.put(np.asfortranarray(w, dtype=np.float64))  # backend memory
        nd = w.ndim if isinstance(w, np.ndarray) else w.dim()
        if nd == 1:
            _check(self.ctx.c.rrtmgpb_gc_set_vmr_1d(h, name, int(w.shape[0]), C.c_void_p(_addr(w)), err), err)
        else:
            _check(self.ctx.c.rrtmgpb_gc_set_vmr_2d(h, name, int(w.shape[0]), int(w.shape[1]), C.c_void_p(_addr(w)), err), err)
        if temporary:  # the device copy of a host array dies here: the (stream-ordered) copy out of it has to be complete
            self.ctx.lib.sync()

    def get_vmr(self, gas, array, ncol, nlay):
        """get_vmr_2d, :433-504: array(ncol,nlay) <- the stored concentration, broadcast."""
        err = C.create_string_buffer(ERRLEN)
        _check(self.ctx.c.rrtmgpb_gc_get_vmr(C.c_void_p(self.handle), gas.encode(), ncol, nlay, C.c_void_p(_addr(array)), err),
               err)

    def fill_vmr(self, gas_names, vmr):
        """The loop of gas_optics over its gases, mo_gas_optics_rrtmgp.F90:540-545: vmr(:,:,igas) <- get_vmr(name)."""
        ncol, nlay = int(vmr.shape[0]), int(vmr.shape[1])
        for igas, name in enumerate(gas_names):
            self.get_vmr(name, vmr[:, :, igas], ncol, nlay)

    def __del__(self):
        try:
            self.ctx.c.rrtmgpb_gc_free(C.c_void_p(self.handle))
        except Exception:
            pass


class GasOptics:
    """ty_gas_optics_rrtmgp: load() copies the tables to backend memory once; gas_optics() dispatches on
    the source type like the Fortran generic (mo_gas_optics_rrtmgp.F90:220,337)."""

    def __init__(self, ctx, kd):
        self.ctx, self.kd = ctx, kd
        self._keep = []

        def h(a, dt):
            a = np.asfortranarray(a, dtype=dt)
            self._keep.append(a)
            return a.ctypes.data

        i32, f64, b8 = np.int32, np.float64, np.bool_
        s = _KDistStruct()
        s.ngas, s.nflav, s.neta, s.npres, s.ntemp, s.nbnd, s.ngpt = kd.ngas, kd.nflav, kd.neta, kd.npres, kd.ntemp, kd.nbnd, kd.ngpt
        s.nminorlower, s.nminorupper = kd.extra["nminorlower"], kd.extra["nminorupper"]
        s.nminorklower, s.nminorkupper = kd.kminor_lower.shape[2], kd.kminor_upper.shape[2]
        s.idx_h2o = kd.idx_h2o
        s.flavor, s.gpoint_flavor = h(kd.flavor, i32), h(kd.gpoint_flavor, i32)
        s.band_lims_gpt, s.gpoint_bands = h(kd.band_lims_gpt, i32), h(kd.gpoint_bands, i32)
        s.band_lims_wvn, s.press_ref_log = h(kd.band_lims_wvn, f64), h(kd.press_ref_log, f64)
        s.temp_ref, s.vmr_ref = h(kd.temp_ref, f64), h(kd.vmr_ref, f64)
        for n in ("press_ref_log_delta", "temp_ref_min", "temp_ref_max", "temp_ref_delta", "press_ref_min",
                  "press_ref_max", "press_ref_trop_log"):
            setattr(s, n, getattr(kd, n))
        s.kmajor, s.kminor_lower, s.kminor_upper = h(kd.kmajor, f64), h(kd.kminor_lower, f64), h(kd.kminor_upper, f64)
        for n in ("minor_limits_gpt", "idx_minor", "idx_minor_scaling", "kminor_start"):
            for lu in ("lower", "upper"):
                setattr(s, f"{n}_{lu}", h(getattr(kd, f"{n}_{lu}"), i32))
        for n in ("minor_scales_with_density", "scale_by_complement"):
            for lu in ("lower", "upper"):
                setattr(s, f"{n}_{lu}", h(getattr(kd, f"{n}_{lu}"), b8))
        if kd.is_lw:
            s.planck_frac, s.totplnk = h(kd.planck_frac, f64), h(kd.totplnk, f64)
            s.nPlanckTemp, s.totplnk_delta = kd.totplnk.shape[0], kd.totplnk_delta
        else:
            s.krayl, s.solar_source = h(kd.krayl, f64), h(kd.solar_source, f64)
        err = C.create_string_buffer(ERRLEN)
        self.handle = ctx.c.rrtmgpb_gas_optics_load(C.byref(s), err)
        if not self.handle:
            raise RuntimeError(err.value.decode())
        self.band_lims_gpt, self.band_lims_wvn = kd.band_lims_gpt, kd.band_lims_wvn
        self.ngpt, self.nband = kd.ngpt, kd.nbnd

    def source_is_internal(self):
        return bool(self.kd.is_lw)

    def source_is_external(self):
        return not self.kd.is_lw

    def gas_optics(self, p_lay, p_lev, t_lay, vmr, optical_props, t_sfc=None, sources=None, toa_src=None,
                   col_dry=None, tlev=None, fused=False, increment_by=None, increment_by2=None):
        """fused=False: the reference call sequence, kernel by kernel.  fused=True: one fused pass
        (rrtmgpb_gas_optics_*_fused) that can also fold in `increment_by%increment(optical_props)` and then
        `increment_by2%increment(optical_props)` (clouds, aerosols: rrtmgp_allsky.F90:376-377,392-399)."""
        ncol, nlay = optical_props.ncol, optical_props.nlay
        err = C.create_string_buffer(ERRLEN)
        o = optical_props.struct()
        P = lambda x: C.c_void_p(_addr(x))
        if fused:
            cl = increment_by.struct() if increment_by is not None else None
            clp = C.byref(cl) if cl is not None else None
            ae = increment_by2.struct() if increment_by2 is not None else None
            aep = C.byref(ae) if ae is not None else None
            if self.kd.is_lw:
                s = sources.struct()
                rc = self.ctx.c.rrtmgpb_gas_optics_int_fused(C.c_void_p(self.handle), ncol, nlay, P(p_lay), P(p_lev),
                                                             P(t_lay), P(t_sfc), P(vmr), C.byref(o), C.byref(s),
                                                             P(col_dry), P(tlev), clp, aep, err)
            else:
                rc = self.ctx.c.rrtmgpb_gas_optics_ext_fused(C.c_void_p(self.handle), ncol, nlay, P(p_lay), P(p_lev),
                                                             P(t_lay), P(vmr), C.byref(o), P(toa_src), P(col_dry),
                                                             clp, aep, err)
            _check(rc, err)
            optical_props._sync_back(o)
            return
        if increment_by is not None or increment_by2 is not None:
            raise ValueError("increment_by needs fused=True")
        if self.kd.is_lw:
            s = sources.struct()
            rc = self.ctx.c.rrtmgpb_gas_optics_int(C.c_void_p(self.handle), ncol, nlay, P(p_lay), P(p_lev), P(t_lay),
                                                   P(t_sfc), P(vmr), C.byref(o), C.byref(s), P(col_dry), P(tlev), err)
        else:
            rc = self.ctx.c.rrtmgpb_gas_optics_ext(C.c_void_p(self.handle), ncol, nlay, P(p_lay), P(p_lev), P(t_lay),
                                                   P(vmr), C.byref(o), P(toa_src), P(col_dry), err)
        _check(rc, err)
        optical_props._sync_back(o)

    def compute_optimal_angles(self, optical_props, optimal_angles=None):
        """ty_gas_optics_rrtmgp%compute_optimal_angles (mo_gas_optics_rrtmgp.F90:1503-1562): secant of the transport
        angle per (column, g-point) from the column transmissivity; feed it to rte_lw(..., lw_Ds=)."""
        err = C.create_string_buffer(ERRLEN)
        fit = self.kd.extra.get("optimal_angle_fit")
        if fit is not None and not getattr(self, "_oa_set", False):
            fit = np.asfortranarray(fit, dtype=np.float64)
            _check(self.ctx.c.rrtmgpb_gas_optics_set_optimal_angle_fit(C.c_void_p(self.handle),
                                                                       C.c_void_p(fit.ctypes.data), err), err)
            self._oa_set = True
        if optimal_angles is None:
            optimal_angles = self.ctx.zeros((optical_props.ncol, optical_props.ngpt))
        o = optical_props.struct()
        _check(self.ctx.c.rrtmgpb_gas_optics_compute_optimal_angles(
            C.c_void_p(self.handle), C.byref(o), int(optimal_angles.shape[0]), int(optimal_angles.shape[1]),
            C.c_void_p(_addr(optimal_angles)), err), err)
        return optimal_angles

    def __del__(self):
        try:
            self.ctx.c.rrtmgpb_gas_optics_free(C.c_void_p(self.handle))
        except Exception:
            pass


class CloudOptics:
    """ty_cloud_optics_rrtmgp in LUT form (mo_cloud_optics_rrtmgp.F90:77,256)."""

    def __init__(self, ctx, lut):
        self.ctx, self.lut, self._keep = ctx, lut, []

        def h(a):
            a = np.asfortranarray(a, dtype=np.float64)
            self._keep.append(a)
            return a.ctypes.data

        s = _CloudLutStruct()
        s.nbnd, s.nsize_liq, s.nsize_ice = lut.nbnd, lut.extliq.shape[0], lut.extice.shape[0]
        s.nrghice, s.icergh = lut.extice.shape[2], lut.icergh
        s.band_lims_gpt, s.band_lims_wvn = None, h(lut.band_lims_wvn)
        s.radliq_lwr, s.radliq_upr, s.diamice_lwr, s.diamice_upr = lut.radliq_lwr, lut.radliq_upr, lut.diamice_lwr, lut.diamice_upr
        for n in ("extliq", "ssaliq", "asyliq", "extice", "ssaice", "asyice"):
            setattr(s, n, h(getattr(lut, n)))
        err = C.create_string_buffer(ERRLEN)
        self.handle = ctx.c.rrtmgpb_cloud_optics_load(C.byref(s), err)
        if not self.handle:
            raise RuntimeError(err.value.decode())
        nb = lut.nbnd
        self.band_lims_gpt = np.asfortranarray(np.stack([np.arange(1, nb + 1), np.arange(1, nb + 1)]), dtype=np.int32)
        self.band_lims_wvn = lut.band_lims_wvn

    def cloud_optics(self, clwp, ciwp, reliq, dgice, optical_props, delta_scale=False):
        """delta_scale=True: cloud_optics() and optical_props%delta_scale() (rrtmgp_allsky.F90:350-352) in one pass."""
        err = C.create_string_buffer(ERRLEN)
        o = optical_props.struct()
        P = lambda x: C.c_void_p(_addr(x))
        _check(self.ctx.c.rrtmgpb_cloud_optics_delta_scaled(C.c_void_p(self.handle), optical_props.ncol,
                                                            optical_props.nlay, P(clwp), P(ciwp), P(reliq), P(dgice),
                                                            C.byref(o), int(bool(delta_scale)), err), err)

    def __del__(self):
        try:
            self.ctx.c.rrtmgpb_cloud_optics_free(C.c_void_p(self.handle))
        except Exception:
            pass


class AerosolOptics:
    """ty_aerosol_optics_rrtmgp_merra (mo_aerosol_optics_rrtmgp_merra.F90:62,99,233)."""

    def __init__(self, ctx, lut):
        self.ctx, self.lut, self._keep = ctx, lut, []

        def h(a):
            a = np.asfortranarray(a, dtype=np.float64)
            self._keep.append(a)
            return a.ctypes.data

        s = _AerosolLutStruct()
        s.nbnd, s.nval, s.nrh, s.nbin = lut.nbnd, lut.nval, lut.nrh, lut.nbin
        for n in ("band_lims_wvn", "merra_aero_bin_lims", "aero_rh", "aero_dust_tbl", "aero_salt_tbl", "aero_sulf_tbl",
                  "aero_bcar_tbl", "aero_bcar_rh_tbl", "aero_ocar_tbl", "aero_ocar_rh_tbl"):
            setattr(s, n, h(getattr(lut, n)))
        err = C.create_string_buffer(ERRLEN)
        self.handle = ctx.c.rrtmgpb_aerosol_optics_load(C.byref(s), err)
        if not self.handle:
            raise RuntimeError(err.value.decode())
        nb = lut.nbnd
        self.band_lims_gpt = np.asfortranarray(np.stack([np.arange(1, nb + 1), np.arange(1, nb + 1)]), dtype=np.int32)
        self.band_lims_wvn = lut.band_lims_wvn

    def aerosol_optics(self, aero_type, aero_size, aero_mass, relhum, optical_props):
        err = C.create_string_buffer(ERRLEN)
        o = optical_props.struct()
        P = lambda x: C.c_void_p(_addr(x))
        ncol, nlay = aero_type.shape  # ncol = size(aero_type,1), nlay = size(aero_type,2), :284-285
        _check(self.ctx.c.rrtmgpb_aerosol_optics(C.c_void_p(self.handle), int(ncol), int(nlay),
                                                 P(aero_type), P(aero_size), P(aero_mass), P(relhum), C.byref(o),
                                                 err), err)

    def __del__(self):
        try:
            self.ctx.c.rrtmgpb_aerosol_optics_free(C.c_void_p(self.handle))
        except Exception:
            pass


# ----------------------------------------------------------------------------------------------------
# k-distribution ingestion (include/rrtmgp_b200_kdist.h): on-disk variable set -> kernel tables.  The reduction runs in
# C++ (csrc/frontend/kdist_load.cpp); this is marshalling only.
# ----------------------------------------------------------------------------------------------------
_RAW_INT = ("ntemp", "npres", "nabsorbers", "nminorabsorbers", "nextabsorbers", "nmixingfracs", "nlayers", "nbnd", "ngpt",
            "nminor_absorber_intervals_lower", "nminor_absorber_intervals_upper", "ncontributors_lower",
            "ncontributors_upper", "ntemp_planck", "nfit_coeffs")


class _KDistRawStruct(C.Structure):
    _fields_ = ([(n, C.c_int) for n in _RAW_INT]
                + [(n, C.c_void_p) for n in ("gas_names", "key_species", "bnd_limits_wavenumber", "bnd_limits_gpt",
                                             "press_ref", "temp_ref")]
                + [(n, FLOAT) for n in ("absorption_coefficient_ref_P", "absorption_coefficient_ref_T", "press_ref_trop")]
                + [(n, C.c_void_p) for n in ("kminor_lower", "kminor_upper", "gas_minor", "identifier_minor",
                                             "minor_gases_lower", "minor_gases_upper", "minor_limits_gpt_lower",
                                             "minor_limits_gpt_upper", "minor_scales_with_density_lower",
                                             "minor_scales_with_density_upper", "scale_by_complement_lower",
                                             "scale_by_complement_upper", "scaling_gas_lower", "scaling_gas_upper",
                                             "kminor_start_lower", "kminor_start_upper", "vmr_ref", "kmajor", "rayl_lower",
                                             "rayl_upper", "totplnk", "plank_fraction", "optimal_angle_fit",
                                             "solar_source_quiet", "solar_source_facular", "solar_source_sunspot")]
                + [(n, FLOAT) for n in ("tsi_default", "mg_default", "sb_default")])


def load_kdist_raw(lib, raw, available_gases, mg_index=None, sb_index=None, tsi=None):
    """ty_gas_optics_rrtmgp%load on the on-disk variable set `raw` (dict, see synthetic.make_kdist_raw /
    mo_optics_utils_rrtmgp.F90:102-183) for the gases the host provides.  Returns a synthetic.KDist holding numpy copies
    of the reduced kernel-layout tables (feed it to GasOptics) - raises RuntimeError with the reference's message."""
    from .synthetic import KDist

    c = lib.cdll
    c.rrtmgpb_kdist_reduce.restype = C.c_void_p
    c.rrtmgpb_kdist_loaded_tables.restype = C.POINTER(_KDistStruct)
    c.rrtmgpb_kdist_loaded_gas_name.restype = C.c_char_p
    c.rrtmgpb_kdist_loaded_optimal_angle_fit.restype = C.c_void_p
    for fn in ("rrtmgpb_kdist_loaded_tables", "rrtmgpb_kdist_loaded_free", "rrtmgpb_kdist_loaded_optimal_angle_fit"):
        getattr(c, fn).argtypes = [C.c_void_p]
    c.rrtmgpb_kdist_loaded_gas_name.argtypes = [C.c_void_p, C.c_int]
    c.rrtmgpb_kdist_loaded_is_key.argtypes = [C.c_void_p, C.c_int]
    keep = []

    def arr(a, dt):
        a = np.asfortranarray(a, dtype=dt)
        keep.append(a)
        return a.ctypes.data

    def strs(lst):
        bufs = [C.create_string_buffer(s.encode()) for s in lst]
        tab = (C.c_char_p * max(len(lst), 1))(*[C.cast(b, C.c_char_p) for b in bufs])
        keep.extend([bufs, tab])
        return C.cast(tab, C.c_void_p)

    i32, f64, b8 = np.int32, np.float64, np.bool_
    s = _KDistRawStruct()
    km = np.asarray(raw["kmajor"])
    s.ngpt, s.nmixingfracs, s.npres, s.ntemp = km.shape[0], km.shape[1], km.shape[2] - 1, km.shape[3]
    s.nabsorbers, s.nextabsorbers = len(raw["gas_names"]), np.asarray(raw["vmr_ref"]).shape[1]
    s.nminorabsorbers, s.nlayers, s.nbnd = len(raw["gas_minor"]), 2, np.asarray(raw["bnd_limits_gpt"]).shape[1]
    s.nminor_absorber_intervals_lower, s.nminor_absorber_intervals_upper = len(raw["minor_gases_lower"]), len(raw["minor_gases_upper"])
    s.ncontributors_lower, s.ncontributors_upper = np.asarray(raw["kminor_lower"]).shape[0], np.asarray(raw["kminor_upper"]).shape[0]
    for n in ("gas_names", "gas_minor", "identifier_minor", "minor_gases_lower", "minor_gases_upper", "scaling_gas_lower",
              "scaling_gas_upper"):
        setattr(s, n, strs(raw[n]))
    for n, dt in (("key_species", i32), ("bnd_limits_wavenumber", f64), ("bnd_limits_gpt", i32), ("press_ref", f64),
                  ("temp_ref", f64), ("kminor_lower", f64), ("kminor_upper", f64), ("minor_limits_gpt_lower", i32),
                  ("minor_limits_gpt_upper", i32), ("minor_scales_with_density_lower", b8),
                  ("minor_scales_with_density_upper", b8), ("scale_by_complement_lower", b8),
                  ("scale_by_complement_upper", b8), ("kminor_start_lower", i32), ("kminor_start_upper", i32),
                  ("vmr_ref", f64), ("kmajor", f64)):
        setattr(s, n, arr(raw[n], dt))
    for n in ("absorption_coefficient_ref_P", "absorption_coefficient_ref_T", "press_ref_trop"):
        setattr(s, n, float(raw[n]))
    if raw.get("rayl_lower") is not None:
        s.rayl_lower = arr(raw["rayl_lower"], f64)
    if raw.get("rayl_upper") is not None:
        s.rayl_upper = arr(raw["rayl_upper"], f64)
    is_lw = "totplnk" in raw
    if is_lw:
        s.totplnk, s.plank_fraction = arr(raw["totplnk"], f64), arr(raw["plank_fraction"], f64)
        s.optimal_angle_fit = arr(raw["optimal_angle_fit"], f64)
        s.ntemp_planck, s.nfit_coeffs = np.asarray(raw["totplnk"]).shape[0], np.asarray(raw["optimal_angle_fit"]).shape[0]
    else:
        for n in ("solar_source_quiet", "solar_source_facular", "solar_source_sunspot"):
            setattr(s, n, arr(raw[n], f64))
        s.tsi_default, s.mg_default, s.sb_default = float(raw["tsi_default"]), float(raw["mg_default"]), float(raw["sb_default"])
    err = C.create_string_buffer(ERRLEN)
    avail = strs(list(available_gases))
    h = c.rrtmgpb_kdist_reduce(C.byref(s), len(available_gases), avail, err)
    if not h:
        raise RuntimeError(err.value.decode())
    try:
        if not is_lw and (mg_index is not None or sb_index is not None):   # set_solar_variability(mg, sb) :760-798
            c.rrtmgpb_kdist_set_solar_variability.argtypes = [C.c_void_p, FLOAT, FLOAT, FLOAT, C.c_char_p]
            _check(c.rrtmgpb_kdist_set_solar_variability(h, raw["mg_default"] if mg_index is None else mg_index,
                                                         raw["sb_default"] if sb_index is None else sb_index, -1.0, err), err)
        if not is_lw and tsi is not None:                                  # set_tsi(tsi) :800-835
            c.rrtmgpb_kdist_set_tsi.argtypes = [C.c_void_p, FLOAT, C.c_char_p]
            _check(c.rrtmgpb_kdist_set_tsi(h, tsi, err), err)
        t = c.rrtmgpb_kdist_loaded_tables(h).contents

        def get(ptr, shape, dt):
            n = int(np.prod(shape))
            if not ptr or n == 0:
                return np.zeros(shape, dtype=dt, order="F")
            buf = (C.c_byte * (n * np.dtype(dt).itemsize)).from_address(ptr)
            return np.asfortranarray(np.frombuffer(buf, dtype=dt).reshape(shape, order="F").copy(order="F"))

        nl, nu = t.nminorlower, t.nminorupper
        gas_names = [c.rrtmgpb_kdist_loaded_gas_name(h, i).decode() for i in range(t.ngas)]
        kd = KDist(
            is_lw=is_lw, gas_names=gas_names, ngas=t.ngas, nflav=t.nflav, neta=t.neta, npres=t.npres, ntemp=t.ntemp,
            nbnd=t.nbnd, ngpt=t.ngpt, flavor=get(t.flavor, (2, t.nflav), i32), gpoint_flavor=get(t.gpoint_flavor, (2, t.ngpt), i32),
            band_lims_gpt=get(t.band_lims_gpt, (2, t.nbnd), i32), band_lims_wvn=get(t.band_lims_wvn, (2, t.nbnd), f64),
            gpoint_bands=get(t.gpoint_bands, (t.ngpt,), i32), press_ref=np.array(raw["press_ref"], dtype=f64),
            press_ref_log=get(t.press_ref_log, (t.npres,), f64), temp_ref=get(t.temp_ref, (t.ntemp,), f64),
            press_ref_log_delta=t.press_ref_log_delta, temp_ref_min=t.temp_ref_min, temp_ref_max=t.temp_ref_max,
            temp_ref_delta=t.temp_ref_delta, press_ref_min=t.press_ref_min, press_ref_max=t.press_ref_max,
            press_ref_trop_log=t.press_ref_trop_log, vmr_ref=get(t.vmr_ref, (2, t.ngas + 1, t.ntemp), f64),
            kmajor=get(t.kmajor, (t.ntemp, t.neta, t.npres + 1, t.ngpt), f64),
            kminor_lower=get(t.kminor_lower, (t.ntemp, t.neta, max(t.nminorklower, 1)), f64),
            kminor_upper=get(t.kminor_upper, (t.ntemp, t.neta, max(t.nminorkupper, 1)), f64),
            minor_limits_gpt_lower=get(t.minor_limits_gpt_lower, (2, max(nl, 1)), i32),
            minor_limits_gpt_upper=get(t.minor_limits_gpt_upper, (2, max(nu, 1)), i32),
            minor_scales_with_density_lower=get(t.minor_scales_with_density_lower, (max(nl, 1),), b8),
            minor_scales_with_density_upper=get(t.minor_scales_with_density_upper, (max(nu, 1),), b8),
            scale_by_complement_lower=get(t.scale_by_complement_lower, (max(nl, 1),), b8),
            scale_by_complement_upper=get(t.scale_by_complement_upper, (max(nu, 1),), b8),
            idx_minor_lower=get(t.idx_minor_lower, (max(nl, 1),), i32), idx_minor_upper=get(t.idx_minor_upper, (max(nu, 1),), i32),
            idx_minor_scaling_lower=get(t.idx_minor_scaling_lower, (max(nl, 1),), i32),
            idx_minor_scaling_upper=get(t.idx_minor_scaling_upper, (max(nu, 1),), i32),
            kminor_start_lower=get(t.kminor_start_lower, (max(nl, 1),), i32),
            kminor_start_upper=get(t.kminor_start_upper, (max(nu, 1),), i32), idx_h2o=t.idx_h2o,
        )
        kd.extra["nminorlower"], kd.extra["nminorupper"] = nl, nu
        kd.extra["is_key"] = [bool(c.rrtmgpb_kdist_loaded_is_key(h, i)) for i in range(t.ngas)]
        if is_lw:
            kd.planck_frac = get(t.planck_frac, (t.ntemp, t.neta, t.npres + 1, t.ngpt), f64)
            kd.totplnk = get(t.totplnk, (t.nPlanckTemp, t.nbnd), f64)
            kd.totplnk_delta = t.totplnk_delta
            oaf = c.rrtmgpb_kdist_loaded_optimal_angle_fit(h)
            if oaf:
                kd.extra["optimal_angle_fit"] = get(oaf, (s.nfit_coeffs, t.nbnd), f64)
        else:
            kd.krayl = get(t.krayl, (t.ntemp, t.neta, t.ngpt, 2), f64) if t.krayl else None
            kd.solar_source = get(t.solar_source, (t.ngpt,), f64)
        return kd
    finally:
        c.rrtmgpb_kdist_loaded_free(h)


# ----------------------------------------------------------------------------------------------------
# The all-sky iteration on HOST buffers (rrtmgpb_allsky_stream_host, csrc/abi/allsky_stream.cu): marshalling only.
# ----------------------------------------------------------------------------------------------------
class _AllSkyHostInputs(C.Structure):
    _fields_ = ([("ncol", C.c_int), ("nlay", C.c_int)] + [(n, C.c_void_p) for n in ("p_lay", "p_lev", "t_lay", "t_lev")]
                + [("ngas", C.c_int), ("vmr_field", C.c_void_p), ("vmr_scalar", C.c_void_p)]
                + [(n, C.c_void_p) for n in ("lwp", "iwp", "rel", "dei", "t_sfc", "emis_sfc", "mu0", "sfc_alb_dir", "sfc_alb_dif")])


class _AllSkyHostFluxes(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("lw_flux_up", "lw_flux_dn", "sw_flux_up", "sw_flux_dn", "sw_flux_dir")]


def allsky_stream_host(lib, go_lw, go_sw, co_lw, co_sw, inputs, fluxes, gas_names, vmr_scalars, chunk_cols, express=False):
    """inputs / fluxes: dicts of HOST arrays in Fortran order - numpy arrays or (pinned) torch CPU tensors holding the
    transposed C-contiguous data, i.e. anything whose memory is (ncol, nlay) first-index-fastest.  Gas fields are the
    entries of `inputs` named like a gas; the other gases of `gas_names` take vmr_scalars[name]."""
    def addr(x):
        return None if x is None else _addr(x)

    ncol, nlay = inputs["shape"]
    fields = (C.c_void_p * len(gas_names))(*[addr(inputs.get(g)) for g in gas_names])
    scal = (FLOAT * len(gas_names))(*[float(vmr_scalars.get(g, 0.0)) for g in gas_names])
    s = _AllSkyHostInputs()
    s.ncol, s.nlay, s.ngas = int(ncol), int(nlay), len(gas_names)
    s.vmr_field, s.vmr_scalar = C.cast(fields, C.c_void_p), C.cast(scal, C.c_void_p)
    for n in ("p_lay", "p_lev", "t_lay", "t_lev", "lwp", "iwp", "rel", "dei", "t_sfc", "emis_sfc", "mu0", "sfc_alb_dir",
              "sfc_alb_dif"):
        setattr(s, n, addr(inputs.get(n)))
    f = _AllSkyHostFluxes()
    for n in ("lw_flux_up", "lw_flux_dn", "sw_flux_up", "sw_flux_dn", "sw_flux_dir"):
        setattr(f, n, addr(fluxes.get(n)))
    err = C.create_string_buffer(ERRLEN)
    h = lambda o: C.c_void_p(o.handle) if o is not None else None
    _check(lib.cdll.rrtmgpb_allsky_stream_host(h(go_lw), h(go_sw), h(co_lw), h(co_sw), C.byref(s), C.byref(f), int(chunk_cols),
                                               int(bool(express)), err), err)
