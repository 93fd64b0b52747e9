"""A SECOND, independent restatement - vectorised numpy, written from the Fortran - of the frontend-resident loops the
library runs as kernels (SURVEY 8a'): get_layer_number / col_dry (rte/kernels/mo_gas_optics_utils.F90:127-152),
combine_abs_and_rayleigh (rrtmgp/frontend/mo_gas_optics_rrtmgp.F90:1954-2002), the level-temperature interpolation of
source() (:889-911), the liquid + ice combination of cloud_optics() (rrtmgp/frontend/mo_cloud_optics_rrtmgp.F90:392-424) and
the broadband reductions (rte/kernels/mo_fluxes_broadband_kernels.F90).  tests/test_oracle_crosscheck.py requires
oracle/glue_ref.c and oracle/rte_misc_ref.c to agree BIT FOR BIT."""
import numpy as np

M_DRY, M_H2O, AVOGAD, GRAV = 0.028964, 0.018016, 6.02214076e23, 9.80665   # mo_gas_optics_constants.F90
TINY, EPSILON = np.finfo(np.float64).tiny, np.finfo(np.float64).eps


def get_layer_number(vmr_h2o, plev):
    delta_plev = np.abs(plev[:, :-1] - plev[:, 1:])
    fact = 1.0 / (1.0 + vmr_h2o)
    m_air = (M_DRY + M_H2O * vmr_h2o) * fact
    return 10.0 * delta_plev * AVOGAD * fact / (1000.0 * m_air * 100.0 * GRAV)


def combine_abs_and_rayleigh(tau, tau_rayleigh, two_stream):
    t = tau + tau_rayleigh
    if not two_stream:
        return t, None, None
    with np.errstate(divide="ignore", invalid="ignore"):
        ssa = np.where(t > 2.0 * TINY, tau_rayleigh / t, 0.0)
    return t, ssa, np.zeros_like(t)


def interpolate_tlev(play, plev, tlay):
    ncol, nlay = play.shape
    tlev = np.zeros((ncol, nlay + 1))
    tlev[:, 0] = tlay[:, 0] + (plev[:, 0] - play[:, 0]) * (tlay[:, 1] - tlay[:, 0]) / (play[:, 1] - play[:, 0])
    tlev[:, nlay] = tlay[:, nlay - 1] + (plev[:, nlay] - play[:, nlay - 1]) * (tlay[:, nlay - 1] - tlay[:, nlay - 2]) / \
        (play[:, nlay - 1] - play[:, nlay - 2])
    for ilay in range(1, nlay):
        tlev[:, ilay] = (play[:, ilay - 1] * tlay[:, ilay - 1] * (plev[:, ilay] - play[:, ilay]) +
                         play[:, ilay] * tlay[:, ilay] * (play[:, ilay - 1] - plev[:, ilay])) / \
            (plev[:, ilay] * (play[:, ilay - 1] - play[:, ilay]))
    return tlev


def cloud_combine(ltau, ltaussa, ltaussag, itau, itaussa, itaussag, two_stream):
    if not two_stream:
        return (ltau - ltaussa) + (itau - itaussa), None, None
    tau = ltau + itau
    taussa = ltaussa + itaussa
    g = (ltaussag + itaussag) / np.maximum(EPSILON, taussa)
    ssa = taussa / np.maximum(EPSILON, tau)
    return tau, ssa, g


def sum_broadband(spectral_flux):
    out = np.zeros(spectral_flux.shape[:2])
    for ig in range(spectral_flux.shape[2]):      # g-points in order, like the serial loop
        out = out + spectral_flux[:, :, ig]
    return out


def net_broadband_full(flux_dn, flux_up):
    out = np.zeros(flux_dn.shape[:2])
    for ig in range(flux_dn.shape[2]):
        out = out + (flux_dn[:, :, ig] - flux_up[:, :, ig])
    return out
