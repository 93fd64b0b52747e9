"""Host-resident all-sky workloads streamed through one GPU (BASELINE config 4: more columns than fit on the device).

`HostAllSky` owns what a host model owns - the atmospheric state and the flux arrays in (pinned) host memory - plus the
loaded k-distributions and cloud tables on the device, and runs the reference's all-sky iteration
(examples/all-sky/rrtmgp_allsky.F90:332-409) through the library's host-buffer entry `rrtmgpb_allsky_stream_host`
(csrc/abi/allsky_stream.cu): column chunks, uploads and downloads overlapped with compute inside the library.
Python here only allocates and marshals."""
import numpy as np

from . import synthetic as syn
from .frontend import CloudOptics, Context, GasOptics, allsky_stream_host


def _pinned_f(a):
    """HOST array in Fortran order as a pinned torch tensor holding the transposed C-contiguous data."""
    import torch

    a = np.asfortranarray(a)
    t = torch.from_numpy(np.ascontiguousarray(a.T))
    return t.pin_memory() if torch.cuda.is_available() else t


class HostAllSky:
    def __init__(self, lib, ncol, nlay, kd_lw, kd_sw, chunk_cols, profiles=None, do_clouds=True, express=False, mu0=0.86,
                 sfc_alb=0.06, emis=0.98, pinned=True, device="cuda:0"):
        self.lib, self.ncol, self.nlay, self.chunk, self.express = lib, ncol, nlay, chunk_cols, express
        ctx = Context(lib, device)
        self.ctx = ctx
        prof = profiles if profiles is not None else syn.compute_profiles(300.0, ncol, nlay)
        top_at_1 = bool(prof["p_lay"][0, 0] < prof["p_lay"][0, nlay - 1])
        sfc = nlay if top_at_1 else 0
        put = _pinned_f if pinned else np.asfortranarray
        self.inputs = dict(shape=(ncol, nlay), p_lay=put(prof["p_lay"]), p_lev=put(prof["p_lev"]), t_lay=put(prof["t_lay"]),
                           t_lev=put(prof["t_lev"]), h2o=put(prof["q"]), o3=put(prof["o3"]))
        self.go_lw = GasOptics(ctx, kd_lw) if kd_lw is not None else None
        self.go_sw = GasOptics(ctx, kd_sw) if kd_sw is not None else None
        self.co_lw = self.co_sw = None
        if do_clouds:
            lut_lw = syn.make_cloud_lut(kd_lw) if kd_lw is not None else None
            lut_sw = syn.make_cloud_lut(kd_sw) if kd_sw is not None else None
            self.co_lw = CloudOptics(ctx, lut_lw) if lut_lw is not None else None
            self.co_sw = CloudOptics(ctx, lut_sw) if lut_sw is not None else None
            cl = syn.compute_clouds(prof, lut_lw if lut_lw is not None else lut_sw)
            self.inputs.update({k: put(cl[k]) for k in ("lwp", "iwp", "rel", "dei")})
        if kd_lw is not None:
            self.inputs.update(t_sfc=put(np.ascontiguousarray(prof["t_lev"][:, sfc])),
                               emis_sfc=put(np.full((kd_lw.nbnd, ncol), emis, order="F")))
        if kd_sw is not None:
            self.inputs.update(mu0=put(np.full(ncol, mu0)), sfc_alb_dir=put(np.full((kd_sw.nbnd, ncol), sfc_alb, order="F")),
                               sfc_alb_dif=put(np.full((kd_sw.nbnd, ncol), sfc_alb, order="F")))
        names = (["lw_flux_up", "lw_flux_dn"] if kd_lw is not None else []) + (["sw_flux_up", "sw_flux_dn", "sw_flux_dir"] if kd_sw is not None else [])
        self.fluxes = {n: put(np.zeros((ncol, nlay + 1), order="F")) for n in names}
        self.h2d_bytes = sum(self._nbytes(v) for k, v in self.inputs.items() if k != "shape")
        self.d2h_bytes = sum(self._nbytes(v) for v in self.fluxes.values())

    @staticmethod
    def _nbytes(a):
        return a.numel() * a.element_size() if hasattr(a, "numel") else a.nbytes

    def step(self):
        """One all-sky iteration over all columns; returns when the fluxes are in host memory."""
        allsky_stream_host(self.lib, self.go_lw, self.go_sw, self.co_lw, self.co_sw, self.inputs, self.fluxes, syn.GAS_NAMES,
                           syn.ALLSKY_WELL_MIXED, self.chunk, self.express)

    def fluxes_host(self):
        out = {}
        for k, v in self.fluxes.items():
            a = v.numpy().T if hasattr(v, "numpy") else v
            out[k] = np.asfortranarray(a)
        return out
