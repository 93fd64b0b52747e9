/* ORACLE - TEST INFRASTRUCTURE ONLY (see rte_solver_ref.c header for the rules).
 *
 * CPU restatement in plain C of the O(ncol*nlay*ngpt) loops that the reference keeps in its
 * Fortran FRONTEND (SURVEY.md section 8a'), under the extension names of
 * include/rrtmgp_b200_ext.h, plus the host-memory version of the backend plumbing.
 * Each function cites the frontend lines it follows.  PARITY UNPINNED by reference golden data
 * (the frontend loops are only exercised by the reference's data-dependent regression tests); pinned instead by a second,
 * independent numpy transcription of the Fortran that must agree bit for bit (tests/numpy_glue.py, the numpy statements in
 * tests/test_aerosol_optics.py and tests/test_cloud_sampling.py; tests/test_oracle_crosscheck.py).
 */
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "oracle_ext.h"

#define MAXF(a, b) (((a) > (b)) ? (a) : (b))

const char* rrtmgpb_backend_name(void) { return "cpu-oracle"; }
int rrtmgpb_float_bytes(void) { return (int)sizeof(Float); }
void* rrtmgpb_mem_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void rrtmgpb_mem_free(void* p) { free(p); }
void rrtmgpb_mem_to_backend(void* d, const void* s, size_t n) { memcpy(d, s, n); }
void rrtmgpb_mem_to_host(void* d, const void* s, size_t n) { memcpy(d, s, n); }
void rrtmgpb_mem_copy(void* d, const void* s, size_t n) { memmove(d, s, n); }
void rrtmgpb_set_stream(void* s) { (void)s; }
void* rrtmgpb_get_stream(void) { return NULL; }
void rrtmgpb_set_device(int d) { (void)d; }
void rrtmgpb_sync(void) {}
long long rrtmgpb_launch_count(int reset) { (void)reset; return 0; }
void rrtmgpb_set_solver_variant(int v) { (void)v; }
int rrtmgpb_get_solver_variant(void) { return 0; }
void rrtmgpb_set_tma_staging(int on) { (void)on; }
void rrtmgpb_abi_table_cache(int on) { (void)on; }
void rrtmgpb_set_gas_optics_rows_path(int on) { (void)on; }
void rrtmgpb_tables_changed(const void* kmajor) { (void)kmajor; }
/* the checker's side of the fast-math probe: plain libm / IEEE arithmetic */
void rrtmgpb_fastmath_probe(int n, const double* x, double* e, double* s, double* r, double* d) {
  for (int i = 0; i < n; ++i) { e[i] = exp(x[i]); s[i] = sqrt(fabs(x[i])); r[i] = 1.0 / x[i]; d[i] = (x[i] * x[i] + 1.0) / x[i]; }
}
void rrtmgpb_profile_enable(int on) { (void)on; }
int rrtmgpb_profile_report(char* buf, size_t n) { if (buf && n) buf[0] = 0; return 0; }
void rrtmgpb_set_lw_2stream_lev_source_per_gpt(int on) { oracle_set_lw_2stream_lev_source_per_gpt(on); }

/* mo_gas_optics_rrtmgp.F90:594-609 */
void rrtmgpb_col_gas_from_vmr(int ncol, int nlay, int ngas, const Float* vmr, const Float* col_dry,
                              Float* col_gas) {
  const size_t ncl = (size_t)ncol * nlay;
  for (size_t c = 0; c < ncl; ++c) col_gas[c] = col_dry[c];
  for (int igas = 1; igas <= ngas; ++igas)
    for (size_t c = 0; c < ncl; ++c) col_gas[c + ncl * igas] = vmr[c + ncl * (igas - 1)] * col_dry[c];
}

/* combine_abs_and_rayleigh, mo_gas_optics_rrtmgp.F90:1954-2002 */
void rrtmgpb_combine_abs_and_rayleigh(int ncol, int nlay, int ngpt, int kind, const Float* tau_abs,
                                      const Float* tau_rayleigh, Float* tau, Float* ssa, Float* g) {
  const size_t n = (size_t)ncol * nlay * ngpt;
  const Float tiny = (sizeof(Float) == 8) ? (Float)DBL_MIN : (Float)FLT_MIN;
  if (kind == 1) {
    for (size_t i = 0; i < n; ++i) tau[i] = tau_abs[i] + tau_rayleigh[i];
    return;
  }
  for (size_t i = 0; i < n; ++i) {
    const Float t = tau_abs[i] + tau_rayleigh[i];
    if (t > (Float)2 * tiny) ssa[i] = tau_rayleigh[i] / t; else ssa[i] = 0;
    tau[i] = t;
  }
  for (size_t i = 0; i < n; ++i) g[i] = 0; /* :2002 zero_array(g) */
}

/* source(): tlev interpolation, mo_gas_optics_rrtmgp.F90:893-911 */
void rrtmgpb_interpolate_tlev(int ncol, int nlay, const Float* play, const Float* plev, const Float* tlay,
                              Float* tlev) {
#define A2(a, i, l) a[(size_t)(i) + (size_t)ncol * (size_t)(l)]
  for (int i = 0; i < ncol; ++i) {
    A2(tlev, i, 0) = A2(tlay, i, 0) + (A2(plev, i, 0) - A2(play, i, 0)) * (A2(tlay, i, 1) - A2(tlay, i, 0)) /
                                          (A2(play, i, 1) - A2(play, i, 0));
    A2(tlev, i, nlay) = A2(tlay, i, nlay - 1) + (A2(plev, i, nlay) - A2(play, i, nlay - 1)) *
                                                    (A2(tlay, i, nlay - 1) - A2(tlay, i, nlay - 2)) /
                                                    (A2(play, i, nlay - 1) - A2(play, i, nlay - 2));
  }
  for (int l = 1; l < nlay; ++l)
    for (int i = 0; i < ncol; ++i)
      A2(tlev, i, l) = (A2(play, i, l - 1) * A2(tlay, i, l - 1) * (A2(plev, i, l) - A2(play, i, l)) +
                        A2(play, i, l) * A2(tlay, i, l) * (A2(play, i, l - 1) - A2(plev, i, l))) /
                       (A2(plev, i, l) * (A2(play, i, l - 1) - A2(play, i, l)));
#undef A2
}

/* mo_gas_optics_rrtmgp.F90:405-411 */
void rrtmgpb_broadcast_by_gpt(int ncol, int ngpt, const Float* per_gpt, Float* out) {
  for (int g = 0; g < ngpt; ++g)
    for (int i = 0; i < ncol; ++i) out[(size_t)i + (size_t)ncol * g] = per_gpt[g];
}

/* expand_and_transpose, mo_rte_lw.F90:478-501 */
void rrtmgpb_expand_and_transpose(int ncol, int nband, int ngpt, const int* band_lims_gpt,
                                  const Float* arr_in, Float* arr_out) {
  (void)ngpt;
  for (int b = 0; b < nband; ++b)
    for (int i = 0; i < ncol; ++i)
      for (int g = band_lims_gpt[2 * b] - 1; g < band_lims_gpt[2 * b + 1]; ++g)
        arr_out[(size_t)i + (size_t)ncol * g] = arr_in[(size_t)b + (size_t)nband * i];
}

/* mo_rte_sw.F90:87-93 */
void rrtmgpb_broadcast_by_lay(int ncol, int nlay, const Float* per_col, Float* out) {
  for (int l = 0; l < nlay; ++l)
    for (int i = 0; i < ncol; ++i) out[(size_t)i + (size_t)ncol * l] = per_col[i];
}

/* ty_gas_concs%get_vmr_2d, rte/frontend/gas-optics-template/mo_gas_concentrations.F90:464-501 */
void rrtmgpb_gas_concs_get_vmr(int ncol, int nlay, int nc_conc, int nl_conc, const Float* conc, Float* array) {
  for (int l = 0; l < nlay; ++l)
    for (int i = 0; i < ncol; ++i) {
      Float v;
      if (nc_conc > 1) v = conc[(size_t)i + (size_t)ncol * l];  /* stored as 2D */
      else if (nl_conc > 1) v = conc[l];                        /* stored as 1D */
      else v = conc[0];                                         /* stored as scalar */
      array[(size_t)i + (size_t)ncol * l] = v;
    }
}

/* ty_gas_optics_rrtmgp%compute_optimal_angles, rrtmgp/frontend/mo_gas_optics_rrtmgp.F90:1540-1562 */
void rrtmgpb_compute_optimal_angles(int ncol, int nlay, int ngpt, int nband, const int* band_lims_gpt, const Float* tau,
                                    const Float* optimal_angle_fit, Float* optimal_angles) {
  const size_t ncl = (size_t)ncol * nlay;
  for (int i = 0; i < ncol; ++i)
    for (int g = 1; g <= ngpt; ++g) {
      int b = 0;
      while (b < nband - 1 && g > band_lims_gpt[2 * b + 1]) ++b; /* convert_gpt2band :1540-1542 */
      Float t = 0;
      for (int l = 0; l < nlay; ++l) t = t + tau[(size_t)i + (size_t)ncol * l + ncl * (size_t)(g - 1)]; /* :1552-1554 */
      const Float trans_total = exp(-t);                                                               /* :1555 */
      optimal_angles[(size_t)i + (size_t)ncol * (g - 1)] =
          optimal_angle_fit[2 * b] * trans_total + optimal_angle_fit[2 * b + 1];                       /* :1559-1560 */
    }
}

/* rte/extensions/mo_cloud_sampling.F90:160-190 (max-random) and :250-290 (exponential-random), one column at a time */
static void sampled_mask_ref(int exp_ran, int ncol, int nlay, int ngpt, const Float* randoms, const Float* cloud_frac,
                             const Float* overlap_param, Bool* cloud_mask) {
  const size_t ncl = (size_t)ncol * nlay;
  Float* local_rands = (Float*)malloc(sizeof(Float) * (size_t)(ngpt > 0 ? ngpt : 1));
  for (int icol = 0; icol < ncol; ++icol) {
    int fst = -1, lst = -1;
    for (int l = 0; l < nlay; ++l)
      if (cloud_frac[(size_t)icol + (size_t)ncol * l] > 0) { if (fst < 0) fst = l; lst = l; }
    for (int l = 0; l < nlay; ++l)
      for (int g = 0; g < ngpt; ++g) cloud_mask[(size_t)icol + (size_t)ncol * l + ncl * g] = 0;
    if (fst < 0) continue;
    for (int l = fst; l <= lst; ++l) {
      const Float frac = cloud_frac[(size_t)icol + (size_t)ncol * l];
      if (!(frac > 0)) continue;
      const Float* r = randoms + (size_t)ngpt * ((size_t)l + (size_t)nlay * icol);
      const int prev_cloudy = l > fst && cloud_frac[(size_t)icol + (size_t)ncol * (l - 1)] > 0;
      if (l == fst || !prev_cloudy) {
        for (int g = 0; g < ngpt; ++g) local_rands[g] = r[g];
      } else if (exp_ran) {
        const Float rho = overlap_param[(size_t)icol + (size_t)ncol * (l - 1)];
        const Float sq = sqrt((Float)1 - rho * rho);
        for (int g = 0; g < ngpt; ++g)
          local_rands[g] = rho * (local_rands[g] - (Float)0.5) + sq * (r[g] - (Float)0.5) + (Float)0.5;
      } /* max-random: same deviates as the cloudy layer above */
      for (int g = 0; g < ngpt; ++g)
        cloud_mask[(size_t)icol + (size_t)ncol * l + ncl * g] = local_rands[g] > ((Float)1 - frac);
    }
  }
  free(local_rands);
}
void rrtmgpb_sampled_mask_max_ran(int ncol, int nlay, int ngpt, const Float* randoms, const Float* cloud_frac,
                                  Bool* cloud_mask) {
  sampled_mask_ref(0, ncol, nlay, ngpt, randoms, cloud_frac, 0, cloud_mask);
}
void rrtmgpb_sampled_mask_exp_ran(int ncol, int nlay, int ngpt, const Float* randoms, const Float* cloud_frac,
                                  const Float* overlap_param, Bool* cloud_mask) {
  sampled_mask_ref(1, ncol, nlay, ngpt, randoms, cloud_frac, overlap_param, cloud_mask);
}
/* apply_cloud_mask, mo_cloud_sampling.F90:298-314 */
void rrtmgpb_apply_cloud_mask(int ncol, int nlay, int nbnd, int ngpt, const int* band_lims_gpt, const Bool* cloud_mask,
                              const Float* input_field, Float* sampled_field) {
  const size_t ncl = (size_t)ncol * nlay;
  (void)ngpt;
  for (int b = 0; b < nbnd; ++b)
    for (int g = band_lims_gpt[2 * b]; g <= band_lims_gpt[2 * b + 1]; ++g)
      for (size_t c = 0; c < ncl; ++c)
        sampled_field[c + ncl * (size_t)(g - 1)] = cloud_mask[c + ncl * (size_t)(g - 1)] ? input_field[c + ncl * (size_t)b] : (Float)0;
}

/* mo_cloud_optics_rrtmgp.F90:334-341 */
void rrtmgpb_cloud_masks(int ncol, int nlay, const Float* clwp, const Float* ciwp, Bool* liqmsk, Bool* icemsk) {
  const size_t ncl = (size_t)ncol * nlay;
  for (size_t c = 0; c < ncl; ++c) { liqmsk[c] = clwp[c] > 0; icemsk[c] = ciwp[c] > 0; }
}

/* mo_cloud_optics_rrtmgp.F90:399-424 */
void rrtmgpb_cloud_combine(int ncol, int nlay, int ngpt, int kind, const Float* ltau, const Float* ltaussa,
                           const Float* ltaussag, const Float* itau, const Float* itaussa,
                           const Float* itaussag, Float* tau, Float* ssa, Float* g) {
  const size_t n = (size_t)ncol * nlay * ngpt;
  const Float eps = (sizeof(Float) == 8) ? (Float)DBL_EPSILON : (Float)FLT_EPSILON;
  if (kind == 1) {
    for (size_t i = 0; i < n; ++i) tau[i] = (ltau[i] - ltaussa[i]) + (itau[i] - itaussa[i]);
    return;
  }
  for (size_t i = 0; i < n; ++i) {
    const Float t = ltau[i] + itau[i];
    const Float ts = ltaussa[i] + itaussa[i];
    g[i] = (ltaussag[i] + itaussag[i]) / MAXF(eps, ts);
    ssa[i] = ts / MAXF(eps, t);
    tau[i] = t;
  }
}

/* mo_rte_util_array_validation.F90 any_vals_less_than / any_vals_outside (masked and unmasked) */
int rrtmgpb_any_vals_less_than(size_t n, const Float* a, const Bool* mask, Float v) {
  for (size_t i = 0; i < n; ++i)
    if ((!mask || mask[i]) && a[i] < v) return 1;
  return 0;
}
int rrtmgpb_any_vals_outside(size_t n, const Float* a, const Bool* mask, Float lo, Float hi) {
  for (size_t i = 0; i < n; ++i)
    if ((!mask || mask[i]) && (a[i] < lo || a[i] > hi)) return 1;
  return 0;
}

/* extern void rrtmgp_compute_tau_absorption(...) is in rrtmgp_gas_optics_ref.c */
#include "rrtmgp_kernels.h"
#include "rte_kernels.h"
void rrtmgpb_compute_tau_absorption_assign(
    int ncol, int nlay, int nbnd, int ngpt, int ngas, int nflav, int neta, int npres, int ntemp,
    int nminorlower, int nminorklower, int nminorupper, int nminorkupper, int idx_h2o,
    const int* gpoint_flavor, const int* band_lims_gpt, const Float* kmajor, const Float* kminor_lower,
    const Float* kminor_upper, const int* minor_limits_gpt_lower, const int* minor_limits_gpt_upper,
    const Bool* minor_scales_with_density_lower, const Bool* minor_scales_with_density_upper,
    const Bool* scale_by_complement_lower, const Bool* scale_by_complement_upper,
    const int* idx_minor_lower, const int* idx_minor_upper, const int* idx_minor_scaling_lower,
    const int* idx_minor_scaling_upper, const int* kminor_start_lower, const int* kminor_start_upper,
    const Bool* tropo, const Float* col_mix, const Float* fmajor, const Float* fminor, const Float* play,
    const Float* tlay, const Float* col_gas, const int* jeta, const int* jtemp, const int* jpress, Float* tau) {
  /* the reference sequence: zero_array then accumulate (mo_gas_optics_rrtmgp.F90:679-706) */
  zero_array_3D(&ncol, &nlay, &ngpt, tau);
  rrtmgp_compute_tau_absorption(&ncol, &nlay, &nbnd, &ngpt, &ngas, &nflav, &neta, &npres, &ntemp,
                                &nminorlower, &nminorklower, &nminorupper, &nminorkupper, &idx_h2o,
                                gpoint_flavor, band_lims_gpt, kmajor, kminor_lower, kminor_upper,
                                minor_limits_gpt_lower, minor_limits_gpt_upper,
                                minor_scales_with_density_lower, minor_scales_with_density_upper,
                                scale_by_complement_lower, scale_by_complement_upper, idx_minor_lower,
                                idx_minor_upper, idx_minor_scaling_lower, idx_minor_scaling_upper,
                                kminor_start_lower, kminor_start_upper, tropo, col_mix, fmajor, fminor, play,
                                tlay, col_gas, jeta, jtemp, jpress, tau);
}

/* rrtmgpb_gas_optics_fused on the oracle = the reference's own sequence, kernel by kernel
 * (mo_gas_optics_rrtmgp.F90:578-745,893-928 and the driver's clouds%increment(atmos)). */
void rrtmgpb_gas_optics_fused(const rrtmgpb_gas_tables* t, int ncol, int nlay, const Float* play, const Float* plev,
                              const Float* tlay, const Float* vmr, const Float* col_dry, int op_kind, Float* tau,
                              Float* ssa, Float* g, int cld_kind, const Float* cld_tau, const Float* cld_ssa,
                              const Float* cld_g, int aer_kind, const Float* aer_tau, const Float* aer_ssa,
                              const Float* aer_g, const Float* tlev, const Float* tsfc, int sfc_lay, Float* sfc_src,
                              Float* lay_src, Float* lev_src, Float* sfc_source_Jac) {
  const size_t ncl = (size_t)ncol * nlay, nf = (size_t)t->nflav;
  const int ngpt = t->ngpt, nbnd = t->nbnd;
  Float* cd = malloc(sizeof(Float) * ncl);
  if (col_dry) memcpy(cd, col_dry, sizeof(Float) * ncl);
  else rrtmgpb_get_col_dry(ncol, nlay, vmr + ncl * (size_t)(t->idx_h2o - 1), plev, cd);
  Float* col_gas = malloc(sizeof(Float) * ncl * (t->ngas + 1));
  rrtmgpb_col_gas_from_vmr(ncol, nlay, t->ngas, vmr, cd, col_gas);
  int* jtemp = malloc(sizeof(int) * ncl); int* jpress = malloc(sizeof(int) * ncl);
  int* jeta = malloc(sizeof(int) * 2 * ncl * nf); Bool* tropo = malloc(sizeof(Bool) * ncl);
  Float* fmajor = malloc(sizeof(Float) * 8 * ncl * nf); Float* fminor = malloc(sizeof(Float) * 4 * ncl * nf);
  Float* col_mix = malloc(sizeof(Float) * 2 * ncl * nf);
  rrtmgp_interpolation(&ncol, &nlay, &t->ngas, &t->nflav, &t->neta, &t->npres, &t->ntemp, t->flavor, t->press_ref_log,
                       t->temp_ref, &t->press_ref_log_delta, &t->temp_ref_min, &t->temp_ref_delta,
                       &t->press_ref_trop_log, t->vmr_ref, play, tlay, col_gas, jtemp, fmajor, fminor, col_mix, tropo,
                       jeta, jpress);
  rrtmgpb_compute_tau_absorption_assign(
      ncol, nlay, nbnd, ngpt, t->ngas, t->nflav, t->neta, t->npres, t->ntemp, t->nminorlower, t->nminorklower,
      t->nminorupper, t->nminorkupper, t->idx_h2o, t->gpoint_flavor, t->band_lims_gpt, t->kmajor, t->kminor_lower,
      t->kminor_upper, t->minor_limits_gpt_lower, t->minor_limits_gpt_upper, t->minor_scales_with_density_lower,
      t->minor_scales_with_density_upper, t->scale_by_complement_lower, t->scale_by_complement_upper,
      t->idx_minor_lower, t->idx_minor_upper, t->idx_minor_scaling_lower, t->idx_minor_scaling_upper,
      t->kminor_start_lower, t->kminor_start_upper, tropo, col_mix, fmajor, fminor, play, tlay, col_gas, jeta, jtemp,
      jpress, tau);
  if (t->krayl) {
    Float* tr = malloc(sizeof(Float) * ncl * ngpt);
    rrtmgp_compute_tau_rayleigh(&ncol, &nlay, &nbnd, &ngpt, &t->ngas, &t->nflav, &t->neta, &t->npres, &t->ntemp,
                                t->gpoint_flavor, t->band_lims_gpt, t->krayl, &t->idx_h2o, cd, col_gas, fminor, jeta,
                                tropo, jtemp, tr);
    rrtmgpb_combine_abs_and_rayleigh(ncol, nlay, ngpt, op_kind, tau, tr, tau, ssa, g);
    free(tr);
  } else if (op_kind == 2) {
    zero_array_3D(&ncol, &nlay, &ngpt, ssa);
    zero_array_3D(&ncol, &nlay, &ngpt, g);
  }
  if (cld_kind == 1 && op_kind == 1) rte_inc_1scalar_by_1scalar_bybnd(&ncol, &nlay, &ngpt, tau, cld_tau, &nbnd, t->band_lims_gpt);
  if (cld_kind == 2 && op_kind == 1) rte_inc_1scalar_by_2stream_bybnd(&ncol, &nlay, &ngpt, tau, cld_tau, cld_ssa, &nbnd, t->band_lims_gpt);
  if (cld_kind == 1 && op_kind == 2) rte_inc_2stream_by_1scalar_bybnd(&ncol, &nlay, &ngpt, tau, ssa, cld_tau, &nbnd, t->band_lims_gpt);
  if (cld_kind == 2 && op_kind == 2) rte_inc_2stream_by_2stream_bybnd(&ncol, &nlay, &ngpt, tau, ssa, g, cld_tau, cld_ssa, cld_g, &nbnd, t->band_lims_gpt);
  if (aer_kind == 1 && op_kind == 1) rte_inc_1scalar_by_1scalar_bybnd(&ncol, &nlay, &ngpt, tau, aer_tau, &nbnd, t->band_lims_gpt);
  if (aer_kind == 2 && op_kind == 1) rte_inc_1scalar_by_2stream_bybnd(&ncol, &nlay, &ngpt, tau, aer_tau, aer_ssa, &nbnd, t->band_lims_gpt);
  if (aer_kind == 1 && op_kind == 2) rte_inc_2stream_by_1scalar_bybnd(&ncol, &nlay, &ngpt, tau, ssa, aer_tau, &nbnd, t->band_lims_gpt);
  if (aer_kind == 2 && op_kind == 2) rte_inc_2stream_by_2stream_bybnd(&ncol, &nlay, &ngpt, tau, ssa, g, aer_tau, aer_ssa, aer_g, &nbnd, t->band_lims_gpt);
  if (lay_src)
    rrtmgp_compute_Planck_source(&ncol, &nlay, &nbnd, &ngpt, &t->nflav, &t->neta, &t->npres, &t->ntemp, &t->nPlanckTemp,
                                 tlay, tlev, tsfc, &sfc_lay, fmajor, jeta, tropo, jtemp, jpress, t->gpoint_bands,
                                 t->band_lims_gpt, t->planck_frac, &t->temp_ref_min, &t->totplnk_delta, t->totplnk,
                                 t->gpoint_flavor, sfc_src, lay_src, lev_src, sfc_source_Jac);
  free(cd); free(col_gas); free(jtemp); free(jpress); free(jeta); free(tropo); free(fmajor); free(fminor); free(col_mix);
}

/* ---------------- aerosol optics: rrtmgp/frontend/mo_aerosol_optics_rrtmgp_merra.F90 ---------------- */
/* aerosol mask :343-347 */
void rrtmgpb_aerosol_mask(int ncol, int nlay, const int* type, Bool* aeromsk) {
  const size_t n = (size_t)ncol * nlay;
  for (size_t i = 0; i < n; ++i) aeromsk[i] = type[i] > 0;
}
/* any_int_vals_outside_2D :580-600 */
int rrtmgpb_any_int_vals_outside(size_t n, const int* a, int lo, int hi) {
  for (size_t i = 0; i < n; ++i)
    if (a[i] < lo || a[i] > hi) return 1;
  return 0;
}
/* linear_interp_aero_table :563-577 (1-based indices) */
static Float aero_lin(const Float* table, int index1, int index2, Float weight) {
  return table[index1 - 1] + weight * (table[index2 - 1] - table[index1 - 1]);
}
/* compute_all_from_table :436-559, loop order and per-band repeated searches as written */
static void compute_all_from_table(int ncol, int nlay, int nval, int nrh, int nbin, int nbnd, const int* type,
                                   const Float* size, const Float* mass, const Float* rh, const Float* lims,
                                   const Float* aero_rh, const Float* dust, const Float* salt, const Float* sulf,
                                   const Float* bcar_rh, const Float* bcar, const Float* ocar_rh, const Float* ocar,
                                   Float* tau, Float* taussa, Float* taussag) {
  const size_t ncl = (size_t)ncol * nlay;
  enum { EXT = 0, SSA = 1, G = 2 };
  int ibin = 1; /* the reference leaves ibin undefined when no bin matches; only reachable with checks off */
  for (int ibnd = 0; ibnd < nbnd; ++ibnd)
    for (size_t c = 0; c < ncl; ++c) {
      for (int i = 1; i <= nbin; ++i)
        if (size[c] >= lims[2 * (i - 1)] && size[c] <= lims[2 * (i - 1) + 1]) ibin = i;
      const int itype = type[c];
      int irh1 = 1, irh2 = 1;
      Float rdrh = 0;
      if (itype != 0) {
        irh2 = 1;
        while (rh[c] > aero_rh[irh2 - 1]) {
          irh2 = irh2 + 1;
          if (irh2 > nrh) break;
        }
        irh1 = (irh2 - 1 > 1) ? irh2 - 1 : 1;
        irh2 = (irh2 < nrh) ? irh2 : nrh;
        const Float drh0 = aero_rh[irh2 - 1] - aero_rh[irh1 - 1];
        const Float drh1 = rh[c] - aero_rh[irh1 - 1];
        rdrh = (irh1 == irh2) ? (Float)0 : drh1 / drh0;
      }
      const size_t o = c + ncl * ibnd;
      const size_t b3 = (size_t)nrh * nval * ibnd;                          /* (nrh,nval,nbnd) */
      const size_t b4 = (size_t)nrh * nval * ((ibin - 1) + (size_t)nbin * ibnd); /* (nrh,nval,nbin,nbnd) */
      const Float* rt = NULL;
      const Float* ft = NULL;
      switch (itype) {
        case 1: ft = dust + (size_t)nval * ((ibin - 1) + (size_t)nbin * ibnd); break;
        case 2: rt = salt + b4; break;
        case 3: rt = sulf + b3; break;
        case 4: rt = bcar_rh + b3; break;
        case 5: ft = bcar + (size_t)nval * ibnd; break;
        case 6: rt = ocar_rh + b3; break;
        case 7: ft = ocar + (size_t)nval * ibnd; break;
        default: break;
      }
      if (rt) {
        tau[o] = mass[c] * aero_lin(rt + (size_t)nrh * EXT, irh1, irh2, rdrh);
        taussa[o] = tau[o] * aero_lin(rt + (size_t)nrh * SSA, irh1, irh2, rdrh);
        taussag[o] = taussa[o] * aero_lin(rt + (size_t)nrh * G, irh1, irh2, rdrh);
      } else if (ft) {
        tau[o] = mass[c] * ft[EXT];
        taussa[o] = tau[o] * ft[SSA];
        taussag[o] = taussa[o] * ft[G];
      } else {
        tau[o] = 0; taussa[o] = 0; taussag[o] = 0;
      }
    }
}
/* aerosol_optics: table lookup (:370-380) then the combination (:385-418) */
void rrtmgpb_aerosol_optics_from_table(int ncol, int nlay, int nval, int nrh, int nbin, int nbnd, int kind,
                                       const int* type, const Float* size, const Float* mass, const Float* rh,
                                       const Float* bin_lims, const Float* aero_rh, const Float* dust_tbl,
                                       const Float* salt_tbl, const Float* sulf_tbl, const Float* bcar_rh_tbl,
                                       const Float* bcar_tbl, const Float* ocar_rh_tbl, const Float* ocar_tbl,
                                       Float* tau, Float* ssa, Float* g) {
  const size_t n = (size_t)ncol * nlay * nbnd;
  const Float eps = (sizeof(Float) == 8) ? (Float)DBL_EPSILON : (Float)FLT_EPSILON;
  Float* atau = (Float*)malloc(3 * n * sizeof(Float) + 8);
  Float *ataussa = atau + n, *ataussag = atau + 2 * n;
  compute_all_from_table(ncol, nlay, nval, nrh, nbin, nbnd, type, size, mass, rh, bin_lims, aero_rh, dust_tbl,
                         salt_tbl, sulf_tbl, bcar_rh_tbl, bcar_tbl, ocar_rh_tbl, ocar_tbl, atau, ataussa, ataussag);
  if (kind == 1) {
    for (size_t i = 0; i < n; ++i) tau[i] = atau[i] - ataussa[i];
  } else {
    for (size_t i = 0; i < n; ++i) {
      const Float t = atau[i], ts = ataussa[i];
      tau[i] = t;
      ssa[i] = ts / MAXF(eps, t);
      g[i] = ataussag[i] / MAXF(eps, ts);
    }
  }
  free(atau);
}

/* The checker's side of the one-pass cloud optics: literally the reference sequence (mo_cloud_optics_rrtmgp.F90:334-424). */
void rrtmgpb_cloud_optics_from_tables(int ncol, int nlay, int nbnd, int kind, const Float* clwp, const Float* ciwp,
                                      const Float* reliq, const Float* dgice, int liq_nsteps, Float liq_step_size,
                                      Float liq_offset, const Float* extliq, const Float* ssaliq, const Float* asyliq,
                                      int ice_nsteps, Float ice_step_size, Float ice_offset, const Float* extice,
                                      const Float* ssaice, const Float* asyice, Float* tau, Float* ssa, Float* g) {
  const size_t ncl = (size_t)ncol * nlay, n = ncl * nbnd;
  Bool* liqmsk = (Bool*)malloc(ncl ? ncl : 1);
  Bool* icemsk = (Bool*)malloc(ncl ? ncl : 1);
  Float* w = (Float*)malloc((n ? n : 1) * 6 * sizeof(Float));
  rrtmgpb_cloud_masks(ncol, nlay, clwp, ciwp, liqmsk, icemsk);
  rrtmgp_compute_cld_from_table(&ncol, &nlay, &nbnd, liqmsk, clwp, reliq, &liq_nsteps, &liq_step_size, &liq_offset, extliq,
                                ssaliq, asyliq, w, w + n, w + 2 * n);
  rrtmgp_compute_cld_from_table(&ncol, &nlay, &nbnd, icemsk, ciwp, dgice, &ice_nsteps, &ice_step_size, &ice_offset, extice,
                                ssaice, asyice, w + 3 * n, w + 4 * n, w + 5 * n);
  rrtmgpb_cloud_combine(ncol, nlay, nbnd, kind, w, w + n, w + 2 * n, w + 3 * n, w + 4 * n, w + 5 * n, tau, ssa, g);
  free(w); free(icemsk); free(liqmsk);
}
/* ... followed by clouds%delta_scale(): literally the two calls of the reference driver (rrtmgp_allsky.F90:350-352) */
void rte_delta_scale_2str_k(const int* ncol, const int* nlay, const int* ngpt, Float* tau, Float* ssa, Float* g);
void rrtmgpb_cloud_optics_from_tables_ds(int ncol, int nlay, int nbnd, int kind, const Float* clwp, const Float* ciwp,
                                         const Float* reliq, const Float* dgice, int liq_nsteps, Float liq_step_size,
                                         Float liq_offset, const Float* extliq, const Float* ssaliq, const Float* asyliq,
                                         int ice_nsteps, Float ice_step_size, Float ice_offset, const Float* extice,
                                         const Float* ssaice, const Float* asyice, Float* tau, Float* ssa, Float* g,
                                         int delta_scale) {
  rrtmgpb_cloud_optics_from_tables(ncol, nlay, nbnd, kind, clwp, ciwp, reliq, dgice, liq_nsteps, liq_step_size, liq_offset,
                                   extliq, ssaliq, asyliq, ice_nsteps, ice_step_size, ice_offset, extice, ssaice, asyice, tau,
                                   ssa, g);
  if (delta_scale && kind == 2) rte_delta_scale_2str_k(&ncol, &nlay, &nbnd, tau, ssa, g);
}


/* ---------------- express path: on the oracle, the reference call sequence on full arrays ---------------- */
int rrtmgpb_express_supported(int ncol, int nlay) { (void)ncol; (void)nlay; return 1; }
void rrtmgpb_express(const rrtmgpb_gas_tables* t, int ncol, int nlay, int top_at_1, const Float* play, const Float* plev,
                     const Float* tlay, const Float* tlev, const Float* tsfc, const Float* vmr, const Float* col_dry,
                     int cld_kind, const Float* cld_tau, const Float* cld_ssa, const Float* cld_g,
                     const Float* sfc_emis_or_alb_dir, const Float* sfc_alb_dif, const Float* mu0,
                     const Float* solar_source, int nmus, const Float* Ds_host, const Float* wts_host, Float* flux_up,
                     Float* flux_dn, Float* flux_dir) {
  const int sw = t->krayl != NULL, ngpt = t->ngpt, nbnd = t->nbnd, nlev = nlay + 1;
  const size_t ncl = (size_t)ncol * nlay, nclp = (size_t)ncol * nlev, ncg = (size_t)ncol * ngpt;
  const Bool top = top_at_1 != 0, yes = 1, no = 0;
  Float* tau = malloc(sizeof(Float) * ncl * ngpt);
  Float* decoy = malloc(sizeof(Float) * nclp);
  Float* zero = calloc(ncg, sizeof(Float));
  Float* bc_a = malloc(sizeof(Float) * ncg);
  rrtmgpb_expand_and_transpose(ncol, nbnd, ngpt, t->band_lims_gpt, sfc_emis_or_alb_dir, bc_a);
  if (sw) {
    Float* ssa = malloc(sizeof(Float) * ncl * ngpt); Float* g = malloc(sizeof(Float) * ncl * ngpt);
    Float* bc_b = malloc(sizeof(Float) * ncg); Float* toa = malloc(sizeof(Float) * ncg);
    Float* mu0l = malloc(sizeof(Float) * ncl);
    rrtmgpb_gas_optics_fused(t, ncol, nlay, play, plev, tlay, vmr, col_dry, 2, tau, ssa, g, cld_kind, cld_tau, cld_ssa,
                             cld_g, 0, NULL, NULL, NULL, NULL, NULL, 0, NULL, NULL, NULL, NULL);
    rrtmgpb_expand_and_transpose(ncol, nbnd, ngpt, t->band_lims_gpt, sfc_alb_dif, bc_b);
    rrtmgpb_broadcast_by_gpt(ncol, ngpt, solar_source, toa);
    rrtmgpb_broadcast_by_lay(ncol, nlay, mu0, mu0l);
    rte_sw_solver_2stream(&ncol, &nlay, &ngpt, &top, tau, ssa, g, mu0l, bc_a, bc_b, toa, decoy, decoy, decoy, &no, zero,
                          &yes, flux_up, flux_dn, flux_dir);
    free(ssa); free(g); free(bc_b); free(toa); free(mu0l);
  } else {
    Float* lay = malloc(sizeof(Float) * ncl * ngpt); Float* lev = malloc(sizeof(Float) * nclp * ngpt);
    Float* sfc = malloc(sizeof(Float) * ncg); Float* jac = malloc(sizeof(Float) * ncg);
    Float* Ds = malloc(sizeof(Float) * ncg * nmus);
    for (int imu = 0; imu < nmus; ++imu)
      for (size_t i = 0; i < ncg; ++i) Ds[i + ncg * imu] = Ds_host[imu];
    const int sfc_lay = top_at_1 ? nlay : 1;
    rrtmgpb_gas_optics_fused(t, ncol, nlay, play, plev, tlay, vmr, col_dry, 1, tau, NULL, NULL, cld_kind, cld_tau, cld_ssa,
                             cld_g, 0, NULL, NULL, NULL, tlev, tsfc, sfc_lay, sfc, lay, lev, jac);
    rte_lw_solver_noscat(&ncol, &nlay, &ngpt, &top, &nmus, Ds, wts_host, tau, lay, lev, bc_a, sfc, zero, decoy, decoy, &yes,
                         flux_up, flux_dn, &no, jac, decoy, &no, tau, tau);
    free(lay); free(lev); free(sfc); free(jac); free(Ds);
  }
  free(tau); free(decoy); free(zero); free(bc_a);
}
