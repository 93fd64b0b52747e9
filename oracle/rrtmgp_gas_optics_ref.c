/* ORACLE - TEST INFRASTRUCTURE ONLY (see rte_solver_ref.c header for the rules).
 *
 * CPU restatement in plain C of the reference's DEFAULT RRTMGP gas-optics kernels,
 * /root/reference/rrtmgp/kernels/mo_gas_optics_rrtmgp_kernels.F90 (interpolation :37-170,
 * compute_tau_absorption :176-338 with gas_optical_depths_major :345-396 and _minor :402-501,
 * compute_tau_rayleigh :506-565, compute_Planck_source :568-710, interpolate1D :715-737,
 * interpolate2D_byflav :741-763, interpolate3D_byflav :765-803) and of
 * /root/reference/rrtmgp/kernels/mo_cloud_optics_rrtmgp_kernels.F90:24-65.
 *
 * PARITY UNPINNED: every golden vector for these kernels lives in the un-vendored rrtmgp-data
 * v1.9.1 tarball (reference rrtmgp/CMakeLists.txt:18); nothing in the reference tree pins them
 * offline.  Mitigation: property tests (tests/test_gas_optics_properties.py) and a second, independent numpy transcription
 * of the Fortran (tests/numpy_gas_optics.py) that this file must agree with bit for bit (tests/test_oracle_crosscheck.py).
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "rrtmgp_kernels.h"

#define MIN(a, b) (((a) < (b)) ? (a) : (b))
#define MAX(a, b) (((a) > (b)) ? (a) : (b))

/* interpolate3D_byflav :765-803.  k(ntemp,neta,npres+1,ngpt); jtemp/jeta/jpress 1-based; writes
 * res[gptS..gptE] (1-based, inclusive; res is indexed 0-based by g-point). */
static void interpolate3D_byflav(int neta, int npres, int ntemp, const Float scaling[2],
                                 const Float* fmajor /* (2,2,2) */, const Float* k, int gptS, int gptE,
                                 const int jeta[2], int jtemp, int jpress, Float* res) {
  const size_t s_eta = (size_t)ntemp, s_p = (size_t)ntemp * neta, s_g = (size_t)ntemp * neta * (npres + 1);
#define K4(t, e, p, g) k[(size_t)((t)-1) + s_eta * (size_t)((e)-1) + s_p * (size_t)((p)-1) + s_g * (size_t)((g)-1)]
#define FM(a, b, c) fmajor[((a)-1) + 2 * ((b)-1) + 4 * ((c)-1)]
  const int jeta1 = jeta[0], jeta2 = jeta[1];
  for (int igpt = gptS; igpt <= gptE; ++igpt) {
    res[igpt - 1] =
        scaling[0] * (FM(1, 1, 1) * K4(jtemp, jeta1, jpress - 1, igpt) +
                      FM(2, 1, 1) * K4(jtemp, jeta1 + 1, jpress - 1, igpt) +
                      FM(1, 2, 1) * K4(jtemp, jeta1, jpress, igpt) +
                      FM(2, 2, 1) * K4(jtemp, jeta1 + 1, jpress, igpt)) +
        scaling[1] * (FM(1, 1, 2) * K4(jtemp + 1, jeta2, jpress - 1, igpt) +
                      FM(2, 1, 2) * K4(jtemp + 1, jeta2 + 1, jpress - 1, igpt) +
                      FM(1, 2, 2) * K4(jtemp + 1, jeta2, jpress, igpt) +
                      FM(2, 2, 2) * K4(jtemp + 1, jeta2 + 1, jpress, igpt));
  }
#undef K4
#undef FM
}

/* interpolate2D_byflav :741-763.  k(ntemp,neta,nk); kS..kE are the 1-based third-index range;
 * res[0..kE-kS]. */
static void interpolate2D_byflav(int ntemp, int neta, const Float* fminor /* (2,2) */, const Float* k,
                                 int kS, int kE, const int jeta[2], int jtemp, Float* res) {
  const size_t s_eta = (size_t)ntemp, s_k = (size_t)ntemp * neta;
#define K3(t, e, g) k[(size_t)((t)-1) + s_eta * (size_t)((e)-1) + s_k * (size_t)((g)-1)]
  const int jeta1 = jeta[0], jeta2 = jeta[1];
  for (int ik = kS; ik <= kE; ++ik)
    res[ik - kS] = fminor[0] * K3(jtemp, jeta1, ik) + fminor[1] * K3(jtemp, jeta1 + 1, ik) +
                   fminor[2] * K3(jtemp + 1, jeta2, ik) + fminor[3] * K3(jtemp + 1, jeta2 + 1, ik);
#undef K3
}

/* interpolate1D :715-737.  table(n1, nres) */
static void interpolate1D(Float val, Float offset, Float delta_r, const Float* table, int n1, int nres,
                          Float* res, size_t res_stride) {
  const Float val0 = (val - offset) * delta_r;
  const Float frac = val0 - (Float)trunc((double)val0);
  int index = MIN(n1 - 1, MAX(1, (int)val0 + 1));
  for (int i = 0; i < nres; ++i) {
    const Float* t = table + (size_t)n1 * i;
    res[res_stride * i] = t[index - 1] + frac * (t[index] - t[index - 1]);
  }
}

/* interpolation :37-170 */
void rrtmgp_interpolation(const int* ncol_, const int* nlay_, const int* ngas_, const int* nflav_,
                          const int* neta_, const int* npres_, const int* ntemp_, const int* flavor,
                          const Float* press_ref_log, const Float* temp_ref,
                          const Float* press_ref_log_delta, const Float* temp_ref_min,
                          const Float* temp_ref_delta, const Float* press_ref_trop_log,
                          const Float* vmr_ref, const Float* play, const Float* tlay, const Float* col_gas,
                          int* jtemp, Float* fmajor, Float* fminor, Float* col_mix, Bool* tropo, int* jeta,
                          int* jpress) {
  const int ncol = *ncol_, nlay = *nlay_, ngas = *ngas_, nflav = *nflav_, neta = *neta_, npres = *npres_,
            ntemp = *ntemp_;
  const size_t ncl = (size_t)ncol * nlay;
  const Float tiny = (sizeof(Float) == 8) ? (Float)DBL_MIN : (Float)FLT_MIN;
  Float* ftemp = malloc(sizeof(Float) * ncl);
  Float* fpress = malloc(sizeof(Float) * ncl);
  const Float press_ref_trop = (Float)exp((double)*press_ref_trop_log); /* :99 */
  const Float temp_ref_delta_inv = (Float)1.0 / *temp_ref_delta;
  const Float press_ref_log_1 = press_ref_log[0];
  const Float press_ref_log_delta_inv = (Float)1.0 / *press_ref_log_delta;
  for (size_t c = 0; c < ncl; ++c) { /* :103-119 */
    const int jtemp_ = (int)((tlay[c] - (*temp_ref_min - *temp_ref_delta)) * temp_ref_delta_inv);
    jtemp[c] = MIN(ntemp - 1, MAX(1, jtemp_));
    /* :108 uses the UNCLAMPED index; clamp only the memory access so out-of-range T cannot fault */
    const int jt_mem = MIN(ntemp, MAX(1, jtemp_));
    ftemp[c] = (tlay[c] - temp_ref[jt_mem - 1]) * temp_ref_delta_inv;
    const Float locpress = (Float)1 + ((Float)log((double)play[c]) - press_ref_log_1) * press_ref_log_delta_inv;
    const Float jpress_aint = MIN((Float)(npres - 1), MAX((Float)1.0, (Float)trunc((double)locpress)));
    jpress[c] = (int)jpress_aint;
    fpress[c] = locpress - jpress_aint;
    tropo[c] = play[c] > press_ref_trop;
  }
  /* vmr_ref(2, 0:ngas, ntemp); col_gas(ncol,nlay,0:ngas) */
#define VMR_REF(it, ig, jt) vmr_ref[((it)-1) + 2 * ((size_t)(ig) + (size_t)(ngas + 1) * ((jt)-1))]
  for (int iflav = 0; iflav < nflav; ++iflav) { /* :121-168 */
    const int igas_1 = flavor[2 * iflav], igas_2 = flavor[2 * iflav + 1];
    for (int ilay = 0; ilay < nlay; ++ilay)
      for (int itemp = 1; itemp <= 2; ++itemp)
        for (int icol = 0; icol < ncol; ++icol) {
          const size_t c = (size_t)icol + (size_t)ncol * ilay;
          const int itropo = tropo[c] ? 1 : 2;
          const int jt = jtemp[c] + itemp - 1;
          const Float ratio_eta_half = VMR_REF(itropo, igas_1, jt) / VMR_REF(itropo, igas_2, jt);
          const Float cg1 = col_gas[c + ncl * igas_1], cg2 = col_gas[c + ncl * igas_2];
          const size_t cf = c + ncl * iflav; /* (icol,ilay,iflav) */
          const Float cm = cg1 + ratio_eta_half * cg2;
          col_mix[(itemp - 1) + 2 * cf] = cm;
          Float eta;
          if (cm > (Float)2 * tiny) eta = cg1 / cm; else eta = (Float)0.5; /* :147-151 */
          const Float loceta = eta * (Float)(neta - 1);
          jeta[(itemp - 1) + 2 * cf] = MIN((int)loceta + 1, neta - 1);
          const Float feta = loceta - (Float)trunc((double)loceta);
          const Float ftemp_term = ((Float)(2 - itemp) + (Float)(2 * itemp - 3) * ftemp[c]); /* :157 */
          Float* fmn = fminor + 4 * cf + 2 * (itemp - 1); /* fminor(2,2,col,lay,flav) */
          fmn[0] = ((Float)1 - feta) * ftemp_term;
          fmn[1] = feta * ftemp_term;
          Float* fmj = fmajor + 8 * cf + 4 * (itemp - 1); /* fmajor(2,2,2,col,lay,flav) */
          fmj[0] = ((Float)1 - fpress[c]) * fmn[0];
          fmj[1] = ((Float)1 - fpress[c]) * fmn[1];
          fmj[2] = fpress[c] * fmn[0];
          fmj[3] = fpress[c] * fmn[1];
        }
  }
#undef VMR_REF
  free(ftemp); free(fpress);
}

/* gas_optical_depths_major :345-396 */
static void gas_optical_depths_major(int ncol, int nlay, int nbnd, int ngpt, int nflav, int neta, int npres,
                                     int ntemp, const int* gpoint_flavor, const int* band_lims_gpt,
                                     const Float* kmajor, const Float* col_mix, const Float* fmajor,
                                     const int* jeta, const Bool* tropo, const int* jtemp, const int* jpress,
                                     Float* tau) {
  const size_t ncl = (size_t)ncol * nlay;
  Float* tau_major = malloc(sizeof(Float) * ngpt);
  (void)nflav;
  for (int ibnd = 0; ibnd < nbnd; ++ibnd) {
    const int gptS = band_lims_gpt[2 * ibnd], gptE = band_lims_gpt[2 * ibnd + 1];
    for (size_t c = 0; c < ncl; ++c) {
      const int itropo = tropo[c] ? 1 : 2;
      const int iflav = gpoint_flavor[(itropo - 1) + 2 * (gptS - 1)]; /* :384 band's first g-point */
      const size_t cf = c + ncl * (iflav - 1);
      interpolate3D_byflav(neta, npres, ntemp, col_mix + 2 * cf, fmajor + 8 * cf, kmajor, gptS, gptE,
                           jeta + 2 * cf, jtemp[c], jpress[c] + itropo, tau_major);
      for (int ig = gptS; ig <= gptE; ++ig) tau[c + ncl * (ig - 1)] = tau[c + ncl * (ig - 1)] + tau_major[ig - 1];
    }
  }
  free(tau_major);
}

/* gas_optical_depths_minor :402-501 */
static void gas_optical_depths_minor(int ncol, int nlay, int ngpt, int ngas, int nflav, int ntemp, int neta,
                                     int nminor, int nminork, int idx_h2o, const int* gpt_flv /* stride 2 */,
                                     const Float* kminor, const int* minor_limits_gpt,
                                     const Bool* minor_scales_with_density, const Bool* scale_by_complement,
                                     const int* idx_minor, const int* idx_minor_scaling,
                                     const int* kminor_start, const Float* play, const Float* tlay,
                                     const Float* col_gas, const Float* fminor, const int* jeta,
                                     const int* layer_limits /* (ncol,2) */, const int* jtemp, Float* tau) {
  const size_t ncl = (size_t)ncol * nlay;
  const Float PaTohPa = (Float)0.01;
  Float* tau_minor = malloc(sizeof(Float) * ngpt);
  (void)ngas; (void)nflav; (void)nminork;
  int any = 0;
  for (int i = 0; i < ncol; ++i) any |= (layer_limits[i] > 0);
  if (any) {
    for (int imnr = 0; imnr < nminor; ++imnr) {
      for (int icol = 0; icol < ncol; ++icol) {
        if (layer_limits[icol] > 0) {
          for (int ilay = layer_limits[icol]; ilay <= layer_limits[icol + ncol]; ++ilay) {
            const size_t c = (size_t)icol + (size_t)ncol * (ilay - 1);
            Float scaling = col_gas[c + ncl * idx_minor[imnr]];
            if (minor_scales_with_density[imnr]) {
              scaling = scaling * (PaTohPa * play[c] / tlay[c]);
              if (idx_minor_scaling[imnr] > 0) {
                const Float vmr_fact = (Float)1 / col_gas[c];
                const Float dry_fact = (Float)1 / ((Float)1 + col_gas[c + ncl * idx_h2o] * vmr_fact);
                if (scale_by_complement[imnr])
                  scaling = scaling * ((Float)1 - col_gas[c + ncl * idx_minor_scaling[imnr]] * vmr_fact * dry_fact);
                else
                  scaling = scaling * (col_gas[c + ncl * idx_minor_scaling[imnr]] * vmr_fact * dry_fact);
              }
            }
            const int gptS = minor_limits_gpt[2 * imnr], gptE = minor_limits_gpt[2 * imnr + 1];
            const int iflav = gpt_flv[2 * (gptS - 1)];
            const size_t cf = c + ncl * (iflav - 1);
            interpolate2D_byflav(ntemp, neta, fminor + 4 * cf, kminor, kminor_start[imnr],
                                 kminor_start[imnr] + (gptE - gptS), jeta + 2 * cf, jtemp[c], tau_minor);
            for (int ig = gptS; ig <= gptE; ++ig)
              tau[c + ncl * (ig - 1)] = tau[c + ncl * (ig - 1)] + scaling * tau_minor[ig - gptS];
          }
        }
      }
    }
  }
  free(tau_minor);
}

/* Fortran minloc/maxloc along dim 2 with a mask; returns 0 if the mask is all false. */
static int minloc_masked(int ncol, int nlay, const Float* play, const Bool* mask, int want_true, int icol) {
  int loc = 0; Float best = 0;
  for (int l = 0; l < nlay; ++l) {
    const size_t c = (size_t)icol + (size_t)ncol * l;
    if ((mask[c] != 0) == (want_true != 0) && (loc == 0 || play[c] < best)) { loc = l + 1; best = play[c]; }
  }
  return loc;
}
static int maxloc_masked(int ncol, int nlay, const Float* play, const Bool* mask, int want_true, int icol) {
  int loc = 0; Float best = 0;
  for (int l = 0; l < nlay; ++l) {
    const size_t c = (size_t)icol + (size_t)ncol * l;
    if ((mask[c] != 0) == (want_true != 0) && (loc == 0 || play[c] > best)) { loc = l + 1; best = play[c]; }
  }
  return loc;
}

/* compute_tau_absorption :176-338 */
void rrtmgp_compute_tau_absorption(
    const int* ncol_, const int* nlay_, const int* nbnd, const int* ngpt, const int* ngas, const int* nflav,
    const int* neta, const int* npres, const int* ntemp, const int* nminorlower, const int* nminorklower,
    const int* nminorupper, const int* nminorkupper, const int* idx_h2o, const int* gpoint_flavor,
    const int* band_lims_gpt, const Float* kmajor, const Float* kminor_lower, const Float* kminor_upper,
    const int* minor_limits_gpt_lower, const int* minor_limits_gpt_upper,
    const Bool* minor_scales_with_density_lower, const Bool* minor_scales_with_density_upper,
    const Bool* scale_by_complement_lower, const Bool* scale_by_complement_upper, const int* idx_minor_lower,
    const int* idx_minor_upper, const int* idx_minor_scaling_lower, const int* idx_minor_scaling_upper,
    const int* kminor_start_lower, const int* kminor_start_upper, const Bool* tropo, const Float* col_mix,
    const Float* fmajor, const Float* fminor, const Float* play, const Float* tlay, const Float* col_gas,
    const int* jeta, const int* jtemp, const int* jpress, Float* tau) {
  const int ncol = *ncol_, nlay = *nlay_;
  int* itropo_lower = malloc(sizeof(int) * 2 * ncol);
  int* itropo_upper = malloc(sizeof(int) * 2 * ncol);
  const int top_at_1 = play[0] < play[(size_t)ncol * (nlay - 1)]; /* :274 column 1 only */
  for (int i = 0; i < ncol; ++i) { /* :275-285 */
    if (top_at_1) {
      itropo_lower[i] = minloc_masked(ncol, nlay, play, tropo, 1, i);
      itropo_lower[i + ncol] = nlay;
      itropo_upper[i] = 1;
      itropo_upper[i + ncol] = maxloc_masked(ncol, nlay, play, tropo, 0, i);
    } else {
      itropo_lower[i] = 1;
      itropo_lower[i + ncol] = minloc_masked(ncol, nlay, play, tropo, 1, i);
      itropo_upper[i] = maxloc_masked(ncol, nlay, play, tropo, 0, i);
      itropo_upper[i + ncol] = nlay;
    }
  }
  gas_optical_depths_major(ncol, nlay, *nbnd, *ngpt, *nflav, *neta, *npres, *ntemp, gpoint_flavor,
                           band_lims_gpt, kmajor, col_mix, fmajor, jeta, tropo, jtemp, jpress, tau);
  gas_optical_depths_minor(ncol, nlay, *ngpt, *ngas, *nflav, *ntemp, *neta, *nminorlower, *nminorklower,
                           *idx_h2o, gpoint_flavor + 0, kminor_lower, minor_limits_gpt_lower,
                           minor_scales_with_density_lower, scale_by_complement_lower, idx_minor_lower,
                           idx_minor_scaling_lower, kminor_start_lower, play, tlay, col_gas, fminor, jeta,
                           itropo_lower, jtemp, tau);
  gas_optical_depths_minor(ncol, nlay, *ngpt, *ngas, *nflav, *ntemp, *neta, *nminorupper, *nminorkupper,
                           *idx_h2o, gpoint_flavor + 1, kminor_upper, minor_limits_gpt_upper,
                           minor_scales_with_density_upper, scale_by_complement_upper, idx_minor_upper,
                           idx_minor_scaling_upper, kminor_start_upper, play, tlay, col_gas, fminor, jeta,
                           itropo_upper, jtemp, tau);
  free(itropo_lower); free(itropo_upper);
}

/* compute_tau_rayleigh :506-565 ; krayl(ntemp,neta,ngpt,2) */
void rrtmgp_compute_tau_rayleigh(const int* ncol_, const int* nlay_, const int* nbnd, const int* ngpt_,
                                 const int* ngas, const int* nflav, const int* neta, const int* npres,
                                 const int* ntemp, const int* gpoint_flavor, const int* band_lims_gpt,
                                 const Float* krayl, const int* idx_h2o, const Float* col_dry,
                                 const Float* col_gas, const Float* fminor, const int* jeta, const Bool* tropo,
                                 const int* jtemp, Float* tau_rayleigh) {
  const int ncol = *ncol_, nlay = *nlay_, ngpt = *ngpt_;
  const size_t ncl = (size_t)ncol * nlay;
  Float* k = malloc(sizeof(Float) * ngpt);
  (void)ngas; (void)nflav; (void)npres;
  for (int ibnd = 0; ibnd < *nbnd; ++ibnd) {
    const int gptS = band_lims_gpt[2 * ibnd], gptE = band_lims_gpt[2 * ibnd + 1];
    for (size_t c = 0; c < ncl; ++c) {
      const int itropo = tropo[c] ? 1 : 2;
      const int iflav = gpoint_flavor[(itropo - 1) + 2 * (gptS - 1)];
      const size_t cf = c + ncl * (iflav - 1);
      interpolate2D_byflav(*ntemp, *neta, fminor + 4 * cf,
                           krayl + (size_t)(*ntemp) * (*neta) * ngpt * (itropo - 1), gptS, gptE,
                           jeta + 2 * cf, jtemp[c], k);
      for (int ig = gptS; ig <= gptE; ++ig)
        tau_rayleigh[c + ncl * (ig - 1)] = k[ig - gptS] * (col_gas[c + ncl * (*idx_h2o)] + col_dry[c]);
    }
  }
  free(k);
}

/* compute_Planck_source :568-710 */
void rrtmgp_compute_Planck_source(const int* ncol_, const int* nlay_, const int* nbnd_, const int* ngpt_,
                                  const int* nflav, const int* neta, const int* npres, const int* ntemp,
                                  const int* nPlanckTemp, const Float* tlay, const Float* tlev,
                                  const Float* tsfc, const int* sfc_lay, const Float* fmajor, const int* jeta,
                                  const Bool* tropo, const int* jtemp, const int* jpress,
                                  const int* gpoint_bands, const int* band_lims_gpt, const Float* pfracin,
                                  const Float* temp_ref_min, const Float* totplnk_delta, const Float* totplnk,
                                  const int* gpoint_flavor, Float* sfc_src, Float* lay_src, Float* lev_src,
                                  Float* sfc_source_Jac) {
  const int ncol = *ncol_, nlay = *nlay_, nbnd = *nbnd_, ngpt = *ngpt_;
  const size_t ncl = (size_t)ncol * nlay, nclp = (size_t)ncol * (nlay + 1);
  const Float delta_Tsurf = (Float)1.0;
  const Float one[2] = {(Float)1, (Float)1};
  Float* pfrac = malloc(sizeof(Float) * ncl * ngpt);
  Float* planck_function = malloc(sizeof(Float) * nclp * nbnd);
  Float* res = malloc(sizeof(Float) * ngpt);
  (void)nflav; (void)gpoint_bands;
  for (int ibnd = 0; ibnd < nbnd; ++ibnd) { /* :619-634 */
    const int gptS = band_lims_gpt[2 * ibnd], gptE = band_lims_gpt[2 * ibnd + 1];
    for (size_t c = 0; c < ncl; ++c) {
      const int itropo = tropo[c] ? 1 : 2;
      const int iflav = gpoint_flavor[(itropo - 1) + 2 * (gptS - 1)];
      const size_t cf = c + ncl * (iflav - 1);
      interpolate3D_byflav(*neta, *npres, *ntemp, one, fmajor + 8 * cf, pfracin, gptS, gptE, jeta + 2 * cf,
                           jtemp[c], jpress[c] + itropo, res);
      for (int ig = gptS; ig <= gptE; ++ig) pfrac[c + ncl * (ig - 1)] = res[ig - 1];
    }
  }
  const Float totplnk_delta_r = (Float)1.0 / *totplnk_delta;
  /* planck_function(ncol, nlay+1, nbnd): element (icol, ilev, ibnd) at icol + ncol*(ilev + (nlay+1)*ibnd) */
  for (int icol = 0; icol < ncol; ++icol) { /* :641-656 */
    interpolate1D(tsfc[icol], *temp_ref_min, totplnk_delta_r, totplnk, *nPlanckTemp, nbnd,
                  planck_function + icol, nclp);
    interpolate1D(tsfc[icol] + delta_Tsurf, *temp_ref_min, totplnk_delta_r, totplnk, *nPlanckTemp, nbnd,
                  planck_function + icol + ncol, nclp);
    for (int ibnd = 0; ibnd < nbnd; ++ibnd) {
      const int gptS = band_lims_gpt[2 * ibnd], gptE = band_lims_gpt[2 * ibnd + 1];
      for (int ig = gptS; ig <= gptE; ++ig) {
        const Float pf = pfrac[icol + (size_t)ncol * (*sfc_lay - 1) + ncl * (ig - 1)];
        sfc_src[icol + (size_t)ncol * (ig - 1)] = pf * planck_function[icol + nclp * ibnd];
        sfc_source_Jac[icol + (size_t)ncol * (ig - 1)] =
            pf * (planck_function[icol + ncol + nclp * ibnd] - planck_function[icol + nclp * ibnd]);
      }
    }
  }
  for (size_t c = 0; c < ncl; ++c) /* :658-663 (c = icol + ncol*ilay is also the slot in planck_function) */
    interpolate1D(tlay[c], *temp_ref_min, totplnk_delta_r, totplnk, *nPlanckTemp, nbnd, planck_function + c, nclp);
  for (int ibnd = 0; ibnd < nbnd; ++ibnd) { /* :668-678 */
    const int gptS = band_lims_gpt[2 * ibnd], gptE = band_lims_gpt[2 * ibnd + 1];
    for (int ig = gptS; ig <= gptE; ++ig)
      for (size_t c = 0; c < ncl; ++c)
        lay_src[c + ncl * (ig - 1)] = pfrac[c + ncl * (ig - 1)] * planck_function[c + nclp * ibnd];
  }
  for (size_t c = 0; c < nclp; ++c) /* :681-685 */
    interpolate1D(tlev[c], *temp_ref_min, totplnk_delta_r, totplnk, *nPlanckTemp, nbnd, planck_function + c, nclp);
  for (int ibnd = 0; ibnd < nbnd; ++ibnd) { /* :690-708 */
    const int gptS = band_lims_gpt[2 * ibnd], gptE = band_lims_gpt[2 * ibnd + 1];
    for (int ig = gptS; ig <= gptE; ++ig) {
      const Float* pf = pfrac + ncl * (ig - 1);
      Float* lev = lev_src + nclp * (ig - 1);
      const Float* pl = planck_function + nclp * ibnd;
      for (int icol = 0; icol < ncol; ++icol) lev[icol] = pf[icol] * pl[icol];
      for (int ilay = 1; ilay < nlay; ++ilay)
        for (int icol = 0; icol < ncol; ++icol) {
          const size_t c = (size_t)icol + (size_t)ncol * ilay;
          lev[c] = (Float)sqrt((double)(pf[c - ncol] * pf[c])) * pl[c];
        }
      for (int icol = 0; icol < ncol; ++icol) {
        const size_t c = (size_t)icol + (size_t)ncol * nlay;
        lev[c] = pf[c - ncol] * pl[c];
      }
    }
  }
  free(pfrac); free(planck_function); free(res);
}

/* compute_cld_from_table: mo_cloud_optics_rrtmgp_kernels.F90:24-65 */
void rrtmgp_compute_cld_from_table(const int* ncol, const int* nlay, const int* ngpt, const Bool* mask,
                                   const Float* lwp, const Float* re, const int* nsteps,
                                   const Float* step_size, const Float* offset, const Float* tau_table,
                                   const Float* ssa_table, const Float* asy_table, Float* tau, Float* taussa,
                                   Float* taussag) {
  const size_t ncl = (size_t)*ncol * *nlay;
  const int ns = *nsteps;
  for (int igpt = 0; igpt < *ngpt; ++igpt) {
    const Float *tt = tau_table + (size_t)ns * igpt, *st = ssa_table + (size_t)ns * igpt,
                *at = asy_table + (size_t)ns * igpt;
    for (size_t c = 0; c < ncl; ++c) {
      const size_t o = c + ncl * igpt;
      if (mask[c]) {
        const int index = MIN((int)floor((double)((re[c] - *offset) / *step_size)) + 1, ns - 1);
        const Float fint = (re[c] - *offset) / *step_size - (Float)(index - 1);
        const Float t = lwp[c] * (tt[index - 1] + fint * (tt[index] - tt[index - 1]));
        const Float ts = t * (st[index - 1] + fint * (st[index] - st[index - 1]));
        taussag[o] = ts * (at[index - 1] + fint * (at[index] - at[index - 1]));
        taussa[o] = ts;
        tau[o] = t;
      } else {
        tau[o] = 0; taussa[o] = 0; taussag[o] = 0;
      }
    }
  }
}
