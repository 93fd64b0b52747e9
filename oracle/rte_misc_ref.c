/* ORACLE - TEST INFRASTRUCTURE ONLY (see rte_solver_ref.c header for the rules).
 *
 * CPU restatement in plain C of
 *   /root/reference/rte/kernels/mo_fluxes_broadband_kernels.F90:32-128  (sum/net broadband)
 *   /root/reference/rte/kernels/mo_rte_util_array.F90:32-132            (zero_array, set_to_scalar)
 *   /root/reference/rte/kernels/mo_gas_optics_utils.F90:36-152          (B_nu, Planck 1D/2D,
 *                                                get_layer_mass, get_layer_number = col_dry)
 *   /root/reference/rte/kernels/mo_gas_optics_constants.F90:8-51        (constants)
 * Pinned by the flux-reduction checks in the restated solver unit tests (net = dn - up).
 */
#include <math.h>
#include <float.h>
#include <stddef.h>
#include "rte_kernels.h"
#include "oracle_ext.h"

/* mo_gas_optics_constants.F90:10-38 */
static const double boltzmann_k = 1.380649e-23;
static const double m_h2o = 0.018016;
static const double avogad = 6.02214076e23;
static const double planck_h = 6.626075540e-34;
static const double lightspeed = 2.99792458e8;
static double g_m_dry = 0.028964;
static double g_grav = 9.80665;
static double g_cp_dry = 1004.64;

/* init_constants :42-51 */
void rrtmgpb_init_constants(const Float* gravity, const Float* mol_weight_dry_air,
                            const Float* heat_capacity_dry_air) {
  if (gravity) g_grav = *gravity;
  if (mol_weight_dry_air) g_m_dry = *mol_weight_dry_air;
  if (heat_capacity_dry_air) g_cp_dry = *heat_capacity_dry_air;
}

/* sum_broadband :32-61 (sequential sum over g-points 1..ngpt) */
void rte_sum_broadband(const int* ncol, const int* nlev, const int* ngpt, const Float* spectral_flux,
                       Float* broadband_flux) {
  const size_t n2 = (size_t)*ncol * *nlev;
  for (size_t c = 0; c < n2; ++c) {
    Float s = 0;
    for (int ig = 0; ig < *ngpt; ++ig) s = s + spectral_flux[c + n2 * ig];
    broadband_flux[c] = s;
  }
}

/* net_broadband_full :66-102 */
void rte_net_broadband_full(const int* ncol, const int* nlev, const int* ngpt, const Float* spectral_flux_dn,
                            const Float* spectral_flux_up, Float* broadband_flux_net) {
  const size_t n2 = (size_t)*ncol * *nlev;
  for (size_t c = 0; c < n2; ++c) broadband_flux_net[c] = spectral_flux_dn[c] - spectral_flux_up[c];
  for (int ig = 1; ig < *ngpt; ++ig)
    for (size_t c = 0; c < n2; ++c) {
      const Float diff = spectral_flux_dn[c + n2 * ig] - spectral_flux_up[c + n2 * ig];
      broadband_flux_net[c] = broadband_flux_net[c] + diff;
    }
}

/* net_broadband_precalc :107-128 */
void rte_net_broadband_precalc(const int* ncol, const int* nlev, const Float* flux_dn, const Float* flux_up,
                               Float* broadband_flux_net) {
  const size_t n2 = (size_t)*ncol * *nlev;
  for (size_t c = 0; c < n2; ++c) broadband_flux_net[c] = flux_dn[c] - flux_up[c];
}

/* mo_rte_util_array.F90:32-132 */
static void fill(size_t n, Float* a, Float v) { for (size_t i = 0; i < n; ++i) a[i] = v; }
void zero_array_1D(const int* ni, Float* array) { fill((size_t)*ni, array, 0); }
void zero_array_2D(const int* ni, const int* nj, Float* array) { fill((size_t)*ni * *nj, array, 0); }
void zero_array_3D(const int* ni, const int* nj, const int* nk, Float* array) {
  fill((size_t)*ni * *nj * *nk, array, 0);
}
void zero_array_4D(const int* ni, const int* nj, const int* nk, const int* nl, Float* array) {
  fill((size_t)*ni * *nj * *nk * *nl, array, 0);
}
void set_to_scalar_1D(const int* ni, Float* array, const Float* value) { fill((size_t)*ni, array, *value); }
void set_to_scalar_2D(const int* ni, const int* nj, Float* array, const Float* value) {
  fill((size_t)*ni * *nj, array, *value);
}
void set_to_scalar_3D(const int* ni, const int* nj, const int* nk, Float* array, const Float* value) {
  fill((size_t)*ni * *nj * *nk, array, *value);
}
void set_to_scalar_4D(const int* ni, const int* nj, const int* nk, const int* nl, Float* array,
                      const Float* value) {
  fill((size_t)*ni * *nj * *nk * *nl, array, *value);
}

/* B_nu :36-41 */
static Float B_nu(Float T, Float nu) {
  const Float h = (Float)planck_h, c = (Float)lightspeed, kb = (Float)boltzmann_k;
  const Float nu100 = nu * (Float)100;
  return (Float)100 * (Float)2 * h * (nu100 * nu100 * nu100) * (c * c) /
         ((Float)exp((double)((h * c * nu * (Float)100) / (kb * T))) - (Float)1);
}
/* compute_Planck_source_2D :43-67 */
void rte_compute_Planck_source_2D(const int* ncol, const int* nlay, const int* nnu, const Float* nus,
                                  const Float* dnus, const Float* T, Float* source) {
  const size_t n2 = (size_t)*ncol * *nlay;
  for (int inu = 0; inu < *nnu; ++inu)
    for (size_t c = 0; c < n2; ++c) source[c + n2 * inu] = B_nu(T[c], nus[inu]) * dnus[inu];
}
/* compute_Planck_source_1D :69-93 */
void rte_compute_Planck_source_1D(const int* ncol, const int* nnu, const Float* nus, const Float* dnus,
                                  const Float* T, Float* source) {
  const size_t n1 = (size_t)*ncol;
  for (int inu = 0; inu < *nnu; ++inu)
    for (size_t c = 0; c < n1; ++c) source[c + n1 * inu] = B_nu(T[c], nus[inu]) * dnus[inu];
}

/* get_layer_mass :97-125 ; vmr(ngas,ncol,nlay), plev(ncol,nlay+1) */
void rrtmgpb_get_layer_mass(int ncol, int nlay, int ngas, const Float* vmr, const Float* plev,
                            const Float* mol_weights, Float m_dry, Float* layer_mass) {
  for (int ilay = 0; ilay < nlay; ++ilay)
    for (int icol = 0; icol < ncol; ++icol)
      for (int igas = 0; igas < ngas; ++igas) {
        const size_t k = (size_t)igas + (size_t)ngas * ((size_t)icol + (size_t)ncol * ilay);
        const Float dp = (Float)fabs((double)(plev[icol + (size_t)ncol * (ilay + 1)] - plev[icol + (size_t)ncol * ilay]));
        layer_mass[k] = vmr[k] * (mol_weights[igas] / m_dry) * dp / (Float)g_grav;
      }
}

/* get_layer_number ("col_dry") :127-152.  In the reference this is an array-valued Fortran FUNCTION
 * with a compiler-mangled name (api/mo_gas_optics_utils.F90:53-66); this is its C-callable form. */
void rrtmgpb_get_col_dry(int ncol, int nlay, const Float* vmr_h2o, const Float* plev, Float* col_dry) {
  for (int ilev = 0; ilev < nlay; ++ilev)
    for (int icol = 0; icol < ncol; ++icol) {
      const size_t k = (size_t)icol + (size_t)ncol * ilev;
      const Float delta_plev = (Float)fabs((double)(plev[k] - plev[k + ncol]));
      /* :147  `1.` is a default-real literal; exactly representable, so widening is exact */
      const Float fact = (Float)1 / ((Float)1.f + vmr_h2o[k]);
      const Float m_air = ((Float)g_m_dry + (Float)m_h2o * vmr_h2o[k]) * fact;
      col_dry[k] = (Float)10 * delta_plev * (Float)avogad * fact /
                   ((Float)1000 * m_air * (Float)100 * (Float)g_grav);
    }
}

/* ---- rte/extensions/mo_fluxes_byband.F90:159-218 ---- */
/* sum_byband :159-178 */
void rte_sum_byband(const int* ncol, const int* nlev, const int* ngpt, const int* nbnd, const int* band_lims,
                    const Float* spectral_flux, Float* byband_flux) {
  const size_t n2 = (size_t)*ncol * *nlev;
  (void)ngpt;
  for (int ib = 0; ib < *nbnd; ++ib)
    for (size_t c = 0; c < n2; ++c) {
      Float s = spectral_flux[c + n2 * (size_t)(band_lims[2 * ib] - 1)];
      for (int ig = band_lims[2 * ib] + 1; ig <= band_lims[2 * ib + 1]; ++ig) s = s + spectral_flux[c + n2 * (size_t)(ig - 1)];
      byband_flux[c + n2 * ib] = s;
    }
}
/* net_byband_full :184-208: net = net + dn - up, evaluated left to right */
void rte_net_byband_full(const int* ncol, const int* nlev, const int* ngpt, const int* nbnd, const int* band_lims,
                         const Float* spectral_flux_dn, const Float* spectral_flux_up, Float* byband_flux_net) {
  const size_t n2 = (size_t)*ncol * *nlev;
  (void)ngpt;
  for (int ib = 0; ib < *nbnd; ++ib)
    for (size_t c = 0; c < n2; ++c) {
      size_t o = c + n2 * (size_t)(band_lims[2 * ib] - 1);
      Float s = spectral_flux_dn[o] - spectral_flux_up[o];
      for (int ig = band_lims[2 * ib] + 1; ig <= band_lims[2 * ib + 1]; ++ig) {
        o = c + n2 * (size_t)(ig - 1);
        s = s + spectral_flux_dn[o] - spectral_flux_up[o];
      }
      byband_flux_net[c + n2 * ib] = s;
    }
}
/* net_byband_precalc :210-217 */
void net_byband_precalc(const int* ncol, const int* nlev, const int* nbnd, const Float* byband_flux_dn,
                        const Float* byband_flux_up, Float* byband_flux_net) {
  const size_t n = (size_t)*ncol * *nlev * *nbnd;
  for (size_t i = 0; i < n; ++i) byband_flux_net[i] = byband_flux_dn[i] - byband_flux_up[i];
}

/* ---- rte/extensions/mo_heating_rates.F90 ---- */
/* compute_heating_rate_general :34-64 */
void rrtmgpb_heating_rate(int ncol, int nlay, const Float* flux_up, const Float* flux_dn, const Float* p_lev,
                          Float* heating_rate) {
  const size_t nc = (size_t)ncol;
  const Float grav = (Float)g_grav, cp_dry = (Float)g_cp_dry;
  for (int l = 0; l < nlay; ++l)
    for (size_t c = 0; c < nc; ++c) {
      const size_t i = c + nc * l;
      heating_rate[i] = (flux_up[i + nc] - flux_up[i] - flux_dn[i + nc] + flux_dn[i]) * grav /
                        (cp_dry * (p_lev[i + nc] - p_lev[i]));
    }
}
/* compute_heating_rate_solar_varmu0 :66-117 */
void rrtmgpb_heating_rate_solar_varmu0(int ncol, int nlay, const Float* flux_up, const Float* flux_dn,
                                       const Float* flux_dir, const Float* p_lev, const Float* mu0, Float* heating_rate) {
  const size_t nc = (size_t)ncol, ncl = nc * nlay;
  const Float grav = (Float)g_grav, cp_dry = (Float)g_cp_dry;
  const Float eps = sizeof(Float) == 8 ? (Float)DBL_EPSILON : (Float)FLT_EPSILON;
  int any = 0, any_last = 0;
  rrtmgpb_heating_rate(ncol, nlay, flux_up, flux_dn, p_lev, heating_rate);
  for (size_t i = 0; i < ncl; ++i)
    if (mu0[i] < eps) { any = 1; if (i >= ncl - nc) any_last = 1; }
  if (!any) return; /* :85-86 */
  for (size_t c = 0; c < nc; ++c) {
    int loc = 0, ilay;
    Float best = 0;
    for (int l = 0; l < nlay; ++l) { /* minloc / maxloc over mu0 > 0, first occurrence, 1-based, 0 if none */
      const Float v = mu0[c + nc * l];
      if (!(v > (Float)0)) continue;
      if (loc == 0 || (any_last ? v < best : v > best)) { loc = l + 1; best = v; }
    }
    ilay = any_last ? loc + 1 : loc - 1; /* :102, :104 */
    if (ilay > 1 && ilay < nlay) {       /* :108 */
      const size_t i = c + nc * (size_t)(ilay - 1);
      heating_rate[i] = (flux_up[i + nc] - flux_up[i] - flux_dn[i + nc] + flux_dn[i] + flux_dir[i + nc] - flux_dir[i]) *
                        grav / (cp_dry * (p_lev[i + nc] - p_lev[i]));
    }
  }
}

/* ---- ssm/mo_optics_ssm_kernels.F90 ---- */
/* compute_tau :29-81 */
void ssm_compute_tau_absorption(const int* ncol, const int* nlay, const int* nnu, const int* ngas,
                                const Float* absorption_coeffs, const Float* play, const Float* pref,
                                const Float* layer_mass, Float* tau) {
  const size_t ncl = (size_t)*ncol * *nlay;
  for (int inu = 0; inu < *nnu; ++inu)
    for (size_t c = 0; c < ncl; ++c) {
      Float s = 0;
      for (int ig = 0; ig < *ngas; ++ig) s = s + layer_mass[ig + (size_t)*ngas * c] * absorption_coeffs[ig + (size_t)*ngas * inu];
      tau[c + ncl * inu] = (*pref > (Float)0) ? s * play[c] / *pref : s;
    }
}
/* compute_layer_mass :83-106 */
void ssm_compute_layer_mass(const int* ncol, const int* nlay, const int* ngas, const Float* vmr, const Float* plev,
                            const Float* mol_weights, const Float* m_dry, Float* layer_mass) {
  const size_t nc = (size_t)*ncol, ncl = nc * *nlay;
  for (size_t c = 0; c < ncl; ++c)
    for (int ig = 0; ig < *ngas; ++ig)
      layer_mass[ig + (size_t)*ngas * c] = vmr[ig + (size_t)*ngas * c] * (mol_weights[ig] / *m_dry) *
                                           fabs(plev[c + nc] - plev[c]) / (Float)g_grav;
}
