"""Callers of the five rrtmgp_* extern-ABI symbols (rrtmgp/kernels/api/mo_gas_optics_rrtmgp_kernels.F90:16,98,170,210;
api/mo_cloud_optics_rrtmgp_kernels.F90:18) on any KernelLib (oracle or CUDA), in the Fortran argument order.  Every call
returns host numpy arrays (Fortran shapes).  Floating-point arguments are cast to the library's working precision
(`lib.np_float`: float64, or float32 for the RTE_USE_SP builds)."""
import numpy as np

import refcases as rc
from rte_rrtmgp_b200.abi import fzeros


def profile(kd, ncol, nlay, seed, top_at_1=False, monotonic=True):
    """Seeded random (play, tlay, col_gas, col_dry, tlev, tsfc) inside the k-distribution's validity range."""
    rng = np.random.default_rng(seed)
    play = np.exp(rng.uniform(np.log(1.5), np.log(1.05e5), (ncol, nlay)))
    if monotonic:
        play = np.sort(play, axis=1) if top_at_1 else -np.sort(-play, axis=1)
    play = np.asfortranarray(play)
    tlay = np.asfortranarray(rng.uniform(165.0, 350.0, (ncol, nlay)))
    col_dry = np.asfortranarray(10.0 ** rng.uniform(21.0, 24.5, (ncol, nlay)))
    vmr = rng.uniform(0.0, 1.0, (ncol, nlay, kd.ngas)) * 10.0 ** rng.uniform(-7, -1.5, (1, 1, kd.ngas))
    vmr[rng.random(vmr.shape) < 0.03] = 0.0  # exactly absent gases (col_mix == 0 -> eta = 0.5 branch, :147-151)
    col_gas = np.zeros((ncol, nlay, kd.ngas + 1), order="F")
    col_gas[:, :, 0] = col_dry
    col_gas[:, :, 1:] = vmr * col_dry[:, :, None]
    tlev = np.asfortranarray(rng.uniform(165.0, 350.0, (ncol, nlay + 1)))
    tsfc = rng.uniform(240.0, 330.0, ncol)
    return dict(play=play, tlay=tlay, col_gas=col_gas, col_dry=col_dry, tlev=tlev, tsfc=tsfc)


def interpolation(lib, device, kd, play, tlay, col_gas, keep_device=False):
    ncol, nlay = play.shape
    nf = kd.nflav
    d = lambda a: rc.dev(lib.cast(a), device)
    out = dict(jtemp=fzeros((ncol, nlay), np.int32, device), jpress=fzeros((ncol, nlay), np.int32, device),
               tropo=fzeros((ncol, nlay), np.bool_, device), jeta=fzeros((2, ncol, nlay, nf), np.int32, device),
               col_mix=fzeros((2, ncol, nlay, nf), lib.np_float, device), fmajor=fzeros((2, 2, 2, ncol, nlay, nf), lib.np_float, device),
               fminor=fzeros((2, 2, ncol, nlay, nf), lib.np_float, device))
    lib.rrtmgp_interpolation(ncol, nlay, kd.ngas, nf, kd.neta, kd.npres, kd.ntemp, d(kd.flavor), d(kd.press_ref_log),
                             d(kd.temp_ref), kd.press_ref_log_delta, kd.temp_ref_min, kd.temp_ref_delta,
                             kd.press_ref_trop_log, d(kd.vmr_ref), d(play), d(tlay), d(col_gas), out["jtemp"],
                             out["fmajor"], out["fminor"], out["col_mix"], out["tropo"], out["jeta"], out["jpress"])
    lib.sync()
    return out if keep_device else {k: rc.host(v) for k, v in out.items()}


def tau_absorption(lib, device, kd, play, tlay, col_gas, it):
    """`it`: interpolation outputs living where `lib` expects them (keep_device=True)."""
    ncol, nlay = play.shape
    d = lambda a: rc.dev(lib.cast(a), device)
    tau = fzeros((ncol, nlay, kd.ngpt), lib.np_float, device)
    lib.rrtmgp_compute_tau_absorption(
        ncol, nlay, kd.nbnd, kd.ngpt, kd.ngas, kd.nflav, kd.neta, kd.npres, kd.ntemp, kd.extra["nminorlower"],
        kd.kminor_lower.shape[2], kd.extra["nminorupper"], kd.kminor_upper.shape[2], kd.idx_h2o, d(kd.gpoint_flavor),
        d(kd.band_lims_gpt), d(kd.kmajor), d(kd.kminor_lower), d(kd.kminor_upper), d(kd.minor_limits_gpt_lower),
        d(kd.minor_limits_gpt_upper), d(kd.minor_scales_with_density_lower), d(kd.minor_scales_with_density_upper),
        d(kd.scale_by_complement_lower), d(kd.scale_by_complement_upper), d(kd.idx_minor_lower), d(kd.idx_minor_upper),
        d(kd.idx_minor_scaling_lower), d(kd.idx_minor_scaling_upper), d(kd.kminor_start_lower), d(kd.kminor_start_upper),
        it["tropo"], it["col_mix"], it["fmajor"], it["fminor"], d(play), d(tlay), d(col_gas), it["jeta"], it["jtemp"],
        it["jpress"], tau)
    lib.sync()
    return rc.host(tau)


def tau_rayleigh(lib, device, kd, col_dry, col_gas, it):
    ncol, nlay = col_dry.shape
    d = lambda a: rc.dev(lib.cast(a), device)
    out = fzeros((ncol, nlay, kd.ngpt), lib.np_float, device)
    lib.rrtmgp_compute_tau_rayleigh(ncol, nlay, kd.nbnd, kd.ngpt, kd.ngas, kd.nflav, kd.neta, kd.npres, kd.ntemp,
                                    d(kd.gpoint_flavor), d(kd.band_lims_gpt), d(kd.krayl), kd.idx_h2o, d(col_dry),
                                    d(col_gas), it["fminor"], it["jeta"], it["tropo"], it["jtemp"], out)
    lib.sync()
    return rc.host(out)


def planck_source(lib, device, kd, tlay, tlev, tsfc, sfc_lay, it):
    ncol, nlay = tlay.shape
    d = lambda a: rc.dev(lib.cast(a), device)
    out = [fzeros((ncol, kd.ngpt), lib.np_float, device), fzeros((ncol, nlay, kd.ngpt), lib.np_float, device),
           fzeros((ncol, nlay + 1, kd.ngpt), lib.np_float, device), fzeros((ncol, kd.ngpt), lib.np_float, device)]
    lib.rrtmgp_compute_Planck_source(ncol, nlay, kd.nbnd, kd.ngpt, kd.nflav, kd.neta, kd.npres, kd.ntemp,
                                     kd.totplnk.shape[0], d(tlay), d(tlev), d(tsfc), sfc_lay, it["fmajor"], it["jeta"],
                                     it["tropo"], it["jtemp"], it["jpress"], d(kd.gpoint_bands), d(kd.band_lims_gpt),
                                     d(kd.planck_frac), kd.temp_ref_min, kd.totplnk_delta, d(kd.totplnk),
                                     d(kd.gpoint_flavor), *out)
    lib.sync()
    return [rc.host(o) for o in out]


def cld_from_table(lib, device, mask, lwp, re, nsteps, step, offset, tau_t, ssa_t, asy_t):
    ncol, nlay = lwp.shape
    nb = tau_t.shape[1]
    d = lambda a: rc.dev(lib.cast(a), device)
    out = [fzeros((ncol, nlay, nb), lib.np_float, device) for _ in range(3)]
    lib.rrtmgp_compute_cld_from_table(ncol, nlay, nb, d(mask), d(lwp), d(re), nsteps, step, offset, d(tau_t), d(ssa_t),
                                      d(asy_t), *out)
    lib.sync()
    return [rc.host(o) for o in out]
