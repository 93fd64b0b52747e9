"""The all-sky LW+SW workload: the loop body of the reference's benchmark driver
examples/all-sky/rrtmgp_allsky.F90:332-409 (cloud optics -> gas optics -> increment -> rte_lw / rte_sw),
set up as in :193-311 (analytic RCEMIP-like profile replicated over ncol, ocean-ish boundary conditions,
clouds in 2/3 of the columns).  Backend-agnostic: `ctx` decides whether arrays are CUDA tensors driving
the product library or numpy arrays driving the CPU oracle (tests / CPU baseline only).
"""
import os

import numpy as np

from . import synthetic as syn
from .frontend import (AerosolOptics, CloudOptics, FluxesBroadband, GasConcs, GasOptics, OpticalProps, SourceFuncLW,
                       rte_lw, rte_lw_bygpoint, rte_lw_express, rte_sw, rte_sw_express)


class AllSky:
    def __init__(self, ctx, ncol, nlay, kd_lw=None, kd_sw=None, do_clouds=True, profiles=None, col_offset=0,
                 mu0=0.86, sfc_alb=0.06, emis=0.98, fused=None, do_aerosols=False, lw_2stream=False, express=False):
        """do_aerosols: rrtmgp_allsky.F90:233-236,664-738.  lw_2stream: LW with 2-stream optical properties and
        rte_lw(use_2stream=.true.) (BASELINE config 5); the solver then returns g-point fluxes, which are summed
        with rte_sum_broadband (the reference leaves the broadband arrays unfilled here)."""
        if fused is None:  # fused gas optics (+ cloud increment) by default; RRTMGPB_FUSED=0: the reference's kernel-by-kernel sequence
            fused = os.environ.get("RRTMGPB_FUSED", "1") == "1"
        self.ctx, self.ncol, self.nlay, self.fused = ctx, ncol, nlay, fused
        self.do_aerosols, self.lw_2stream = do_aerosols, lw_2stream
        # express: broadband fluxes straight from the state - atmos / sources (the (ncol,nlay,ngpt) arrays)
        # are never allocated; clouds only (no aerosols, no LW two-stream)
        self.express = express
        self.concurrent = os.environ.get("RRTMGPB_CONCURRENT_LW_SW", "0") == "1"
        self._side = None
        assert not (express and (do_aerosols or lw_2stream))
        prof = profiles if profiles is not None else syn.compute_profiles(300.0, ncol, nlay)
        self.host_inputs = {}
        put = ctx.put
        self.p_lay, self.p_lev = put(prof["p_lay"]), put(prof["p_lev"])
        self.t_lay, self.t_lev = put(prof["t_lay"]), put(prof["t_lev"])
        # gas_concs as the reference driver sets it up (rrtmgp_allsky.F90:195-203, 275-276): h2o and o3 are
        # (ncol,nlay) fields, the well-mixed gases scalars.  gas_optics expands it to vmr(ncol,nlay,ngas) on every
        # call (mo_gas_optics_rrtmgp.F90:540-545); here the expansion runs once per step, shared by LW and SW.
        self.h2o, self.o3 = put(prof["q"]), put(prof["o3"])
        self.gas_concs = GasConcs(ctx, syn.GAS_NAMES)
        for name, w in syn.ALLSKY_WELL_MIXED.items():
            self.gas_concs.set_vmr(name, w)
        self.vmr = ctx.zeros((ncol, nlay, len(syn.GAS_NAMES)))
        self.update_vmr()
        top_at_1 = bool(prof["p_lay"][0, 0] < prof["p_lay"][0, nlay - 1])
        sfc = nlay if top_at_1 else 0
        self.host_inputs.update(p_lay=prof["p_lay"], p_lev=prof["p_lev"], t_lay=prof["t_lay"], t_lev=prof["t_lev"], h2o=prof["q"], o3=prof["o3"])
        self.lw = self.sw = None
        if kd_lw is not None:
            lw = type("LW", (), {})()
            lw.go = GasOptics(ctx, kd_lw)
            lw_kind = "2str" if lw_2stream else "1scl"
            if not express:
                lw.atmos = OpticalProps.like(ctx, lw_kind, ncol, nlay, lw.go)
                lw.sources = SourceFuncLW(ctx, ncol, nlay, kd_lw.ngpt)
            lw.t_sfc = put(np.ascontiguousarray(prof["t_lev"][:, sfc]))  # rrtmgp_allsky.F90:296
            lw.emis_sfc = put(np.full((kd_lw.nbnd, ncol), emis, order="F"))
            lw.flux_up, lw.flux_dn = ctx.zeros((ncol, nlay + 1)), ctx.zeros((ncol, nlay + 1))
            lw.fluxes = FluxesBroadband(flux_up=lw.flux_up, flux_dn=lw.flux_dn)
            if do_clouds:
                lw.lut = syn.make_cloud_lut(kd_lw)
                lw.co = CloudOptics(ctx, lw.lut)
                lw.clouds = OpticalProps.like(ctx, lw_kind, ncol, nlay, lw.co)
            if do_aerosols:
                lw.alut = syn.make_aerosol_lut(kd_lw)
                lw.ao = AerosolOptics(ctx, lw.alut)
                lw.aerosols = OpticalProps.like(ctx, lw_kind, ncol, nlay, lw.ao)
            if lw_2stream:
                lw.gpt_flux_up = ctx.zeros((ncol, nlay + 1, kd_lw.ngpt))
                lw.gpt_flux_dn = ctx.zeros((ncol, nlay + 1, kd_lw.ngpt))
            self.lw = lw
        if kd_sw is not None:
            sw = type("SW", (), {})()
            sw.go = GasOptics(ctx, kd_sw)
            if not express:
                sw.atmos = OpticalProps.like(ctx, "2str", ncol, nlay, sw.go)
                sw.toa_flux = ctx.zeros((ncol, kd_sw.ngpt))
            sw.mu0 = put(np.full(ncol, mu0))
            sw.sfc_alb_dir = put(np.full((kd_sw.nbnd, ncol), sfc_alb, order="F"))
            sw.sfc_alb_dif = put(np.full((kd_sw.nbnd, ncol), sfc_alb, order="F"))
            sw.flux_up, sw.flux_dn, sw.flux_dir = (ctx.zeros((ncol, nlay + 1)) for _ in range(3))
            sw.fluxes = FluxesBroadband(flux_up=sw.flux_up, flux_dn=sw.flux_dn, flux_dn_dir=sw.flux_dir)
            if do_clouds:
                sw.lut = syn.make_cloud_lut(kd_sw)
                sw.co = CloudOptics(ctx, sw.lut)
                sw.clouds = OpticalProps.like(ctx, "2str", ncol, nlay, sw.co)
            if do_aerosols:
                sw.alut = syn.make_aerosol_lut(kd_sw)
                sw.ao = AerosolOptics(ctx, sw.alut)
                sw.aerosols = OpticalProps.like(ctx, "2str", ncol, nlay, sw.ao)
            self.sw = sw
        self.do_clouds = do_clouds
        if do_clouds:
            some = self.lw if self.lw is not None else self.sw
            # compute_clouds uses the 1-based GLOBAL column index in mod(icol,3); col_offset keeps the
            # pattern continuous when columns are sharded across ranks
            cl = syn.compute_clouds(prof, some.lut)
            if col_offset:
                cl = _shift_cloud_columns(prof, some.lut, col_offset)
            self.lwp, self.iwp, self.rel, self.dei = put(cl["lwp"]), put(cl["iwp"]), put(cl["rel"]), put(cl["dei"])
            self.host_inputs.update(cl)
        if do_aerosols:
            ae = syn.compute_aerosols(prof, col_offset)
            self.aero_type, self.aero_size = put(ae["aero_type"]), put(ae["aero_size"])
            self.aero_mass, self.relhum = put(ae["aero_mass"]), put(ae["relhum"])
            self.host_inputs.update(ae)

    def update_vmr(self):
        """gas_concs -> vmr(ncol,nlay,ngas): the get_vmr loop of gas_optics (mo_gas_optics_rrtmgp.F90:540-545)."""
        self.gas_concs.set_vmr("h2o", self.h2o)
        self.gas_concs.set_vmr("o3", self.o3)
        self.gas_concs.fill_vmr(syn.GAS_NAMES, self.vmr)

    # -- one iteration of the reference loop body, LW branch (rrtmgp_allsky.F90:340-381)
    def step_lw(self):
        lw = self.lw
        if self.do_clouds:
            lw.co.cloud_optics(self.lwp, self.iwp, self.rel, self.dei, lw.clouds)
        if self.do_aerosols:
            lw.ao.aerosol_optics(self.aero_type, self.aero_size, self.aero_mass, self.relhum, lw.aerosols)
        if self.express:
            rte_lw_express(self.ctx, lw.go, self.p_lay, self.p_lev, self.t_lay, lw.t_sfc, self.vmr, lw.emis_sfc, lw.fluxes,
                           clouds=lw.clouds if self.do_clouds else None, tlev=self.t_lev)
            return
        if self.fused:  # gas optics + clouds%increment(atmos) [+ aerosols%increment(atmos)] in one pass
            lw.go.gas_optics(self.p_lay, self.p_lev, self.t_lay, self.vmr, lw.atmos, t_sfc=lw.t_sfc,
                             sources=lw.sources, tlev=self.t_lev, fused=True,
                             increment_by=lw.clouds if self.do_clouds else None,
                             increment_by2=lw.aerosols if self.do_aerosols else None)
        else:
            lw.go.gas_optics(self.p_lay, self.p_lev, self.t_lay, self.vmr, lw.atmos, t_sfc=lw.t_sfc,
                             sources=lw.sources, tlev=self.t_lev)
            if self.do_clouds:
                lw.clouds.increment(lw.atmos)
            if self.do_aerosols:
                lw.aerosols.increment(lw.atmos)
        if self.lw_2stream:
            rte_lw_bygpoint(self.ctx, lw.atmos, lw.sources, lw.emis_sfc, lw.gpt_flux_up, lw.gpt_flux_dn, use_2stream=1)
            ngpt, nlev = lw.go.ngpt, self.nlay + 1
            self.ctx.lib.rte_sum_broadband(self.ncol, nlev, ngpt, lw.gpt_flux_up, lw.flux_up)
            self.ctx.lib.rte_sum_broadband(self.ncol, nlev, ngpt, lw.gpt_flux_dn, lw.flux_dn)
        else:
            rte_lw(self.ctx, lw.atmos, lw.sources, lw.emis_sfc, lw.fluxes)

    # -- SW branch (rrtmgp_allsky.F90:340-352,383-406)
    def step_sw(self):
        sw = self.sw
        if self.do_clouds and self.fused:  # cloud_optics + clouds%delta_scale() in one pass over the by-band arrays
            sw.co.cloud_optics(self.lwp, self.iwp, self.rel, self.dei, sw.clouds, delta_scale=True)
        elif self.do_clouds:
            sw.co.cloud_optics(self.lwp, self.iwp, self.rel, self.dei, sw.clouds)
            sw.clouds.delta_scale()
        if self.do_aerosols:
            sw.ao.aerosol_optics(self.aero_type, self.aero_size, self.aero_mass, self.relhum, sw.aerosols)
        if self.express:
            rte_sw_express(self.ctx, sw.go, self.p_lay, self.p_lev, self.t_lay, self.vmr, sw.mu0, sw.sfc_alb_dir,
                           sw.sfc_alb_dif, sw.fluxes, clouds=sw.clouds if self.do_clouds else None)
            return
        if self.fused:  # aerosols%delta_scale() does not depend on the gas optics: it moves in front of the fused pass
            if self.do_aerosols:
                sw.aerosols.delta_scale()
            sw.go.gas_optics(self.p_lay, self.p_lev, self.t_lay, self.vmr, sw.atmos, toa_src=sw.toa_flux, fused=True,
                             increment_by=sw.clouds if self.do_clouds else None,
                             increment_by2=sw.aerosols if self.do_aerosols else None)
        else:
            sw.go.gas_optics(self.p_lay, self.p_lev, self.t_lay, self.vmr, sw.atmos, toa_src=sw.toa_flux)
            if self.do_clouds:
                sw.clouds.increment(sw.atmos)
            if self.do_aerosols:
                sw.aerosols.delta_scale()
                sw.aerosols.increment(sw.atmos)
        rte_sw(self.ctx, sw.atmos, sw.mu0, sw.toa_flux, sw.sfc_alb_dir, sw.sfc_alb_dif, sw.fluxes)

    def step(self):
        self.update_vmr()
        if self.concurrent and self.lw is not None and self.sw is not None and self.ctx.device is not None:
            # the LW and the SW halves of the iteration are independent: issue them on two streams so that the tail of
            # one kernel overlaps the head of the other half's (the library launches on the calling thread's stream)
            import torch

            cur = torch.cuda.current_stream()
            if self._side is None:
                self._side = torch.cuda.Stream()
            self._side.wait_stream(cur)
            self.ctx.lib.set_stream(self._side.cuda_stream)
            self.step_sw()
            self.ctx.lib.set_stream(cur.cuda_stream)
            self.step_lw()
            cur.wait_stream(self._side)
            return
        if self.lw is not None:
            self.step_lw()
        if self.sw is not None:
            self.step_sw()

    def fluxes_host(self):
        out = {}
        if self.lw is not None:
            out["lw_flux_up"], out["lw_flux_dn"] = self.ctx.get(self.lw.flux_up), self.ctx.get(self.lw.flux_dn)
        if self.sw is not None:
            out["sw_flux_up"], out["sw_flux_dn"] = self.ctx.get(self.sw.flux_up), self.ctx.get(self.sw.flux_dn)
            out["sw_flux_dir"] = self.ctx.get(self.sw.flux_dir)
        return out


def _shift_cloud_columns(prof, lut, col_offset):
    ncol = prof["p_lay"].shape[0]
    big = {k: v for k, v in prof.items()}
    cl = syn.compute_clouds(big, lut)
    icol = np.arange(1 + col_offset, ncol + 1 + col_offset)[:, None]
    keep = icol % 3 != 0
    base = (prof["p_lay"] > 100.0 * 100.0) & (prof["p_lay"] < 900.0 * 100.0)
    rel_val = float(np.float32(0.5)) * (lut.radliq_lwr + lut.radliq_upr)
    dei_val = float(np.float32(0.5)) * (lut.diamice_lwr + lut.diamice_upr)
    mask = base & keep
    lwp = np.where(mask & (prof["t_lay"] > 263.0), 10.0, 0.0)
    iwp = np.where(mask & (prof["t_lay"] < 273.0), 10.0, 0.0)
    cl = dict(lwp=np.asfortranarray(lwp), iwp=np.asfortranarray(iwp),
              rel=np.asfortranarray(np.where(lwp > 0, rel_val, 0.0)),
              dei=np.asfortranarray(np.where(iwp > 0, dei_val, 0.0)))
    return cl
