/* ORACLE - TEST INFRASTRUCTURE ONLY.  Never linked into, loaded by, or called from the product
 * library (rte_rrtmgp_b200/lib/librte_rrtmgp_b200.so).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * CPU restatement in plain C of the reference's DEFAULT (serial Fortran) RTE solver kernels,
 * /root/reference/rte/kernels/mo_rte_solver_kernels.F90.  Same C symbols as the reference's
 * extern ABI (include/rte_kernels.h) so the two libraries are interchangeable behind it.
 * Loop order, association order, thresholds and the default-kernel quirks are kept
 * (compile with -O2 -ffp-contract=off so no FMA contraction reorders rounding).
 *
 * Parity pin: the solver paths are pinned by the reference's own analytic known-answer tests,
 * restated in tests/test_rte_lw_solver_unit.py and tests/test_rte_sw_solver_unit.py
 * (reference tests/rte_lw_solver_unit_tests.F90, tests/rte_sw_solver_unit_tests.F90).
 * lw_solver_2stream has no known-answer test in the reference ("parity unpinned" against the Fortran itself); it, the other
 * three solvers (every flag combination) and adding are additionally pinned by a second, independent numpy transcription of
 * the Fortran (tests/numpy_solvers.py) that this file must agree with bit for bit (tests/test_oracle_crosscheck.py).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include "rte_kernels.h"
#include "oracle_ext.h"

/* Fortran (ncol, n2) and (ncol, n2, n3) column-major, 0-based C indices */
#define I2(i, j) ((size_t)(i) + (size_t)ncol * (size_t)(j))
#define I3(i, j, k, n2) ((size_t)(i) + (size_t)ncol * ((size_t)(j) + (size_t)(n2) * (size_t)(k)))

/* mo_rte_solver_kernels.F90:38  pi = acos(-1._wp) */
static Float ref_pi(void) { return (Float)acos(-1.0); }

/* runtime switch for the lw_solver_2stream level-source quirk (see below); 0 = reference default */
static int g_lw2s_lev_source_per_gpt = 0;
void oracle_set_lw_2stream_lev_source_per_gpt(int on) { g_lw2s_lev_source_per_gpt = on; }

/* ------------------------------------------------------------------------------------------
 * lw_source_noscat: mo_rte_solver_kernels.F90:620-675 (Clough et al. 1992 eq 13, linear in tau)
 * 2D slices for ONE g-point; tau here is the slant path tau*D.
 */
static void lw_source_noscat(int ncol, int nlay, int top_at_1, const Float* lay_source,
                             const Float* lev_source, const Float* tau, const Float* trans,
                             Float* source_dn, Float* source_up) {
  /* :636 tau_thresh = sqrt(sqrt(epsilon(tau))) */
  const Float tau_thresh = (Float)sqrt(sqrt((double)(sizeof(Float) == 8 ? DBL_EPSILON : FLT_EPSILON)));
  Float* source_inc = top_at_1 ? source_dn : source_up; /* :638-644 */
  Float* source_dec = top_at_1 ? source_up : source_dn;
  for (int ilay = 0; ilay < nlay; ++ilay) {
    for (int icol = 0; icol < ncol; ++icol) {
      const Float t = tau[I2(icol, ilay)], tr = trans[I2(icol, ilay)];
      Float fact;
      if (t > tau_thresh) { /* :652-656 */
        fact = ((Float)1 - tr) / t - tr;
      } else {
        fact = t * ((Float)0.5 + t * (-(Float)1 / (Float)3 + t * (Float)1 / (Float)8));
      }
      const Float lay = lay_source[I2(icol, ilay)];
      const Float lev_lo = lev_source[I2(icol, ilay)], lev_hi = lev_source[I2(icol, ilay + 1)];
      /* :660-663 */
      source_inc[I2(icol, ilay)] = ((Float)1 - tr) * lev_hi + (Float)2 * fact * (lay - lev_hi);
      source_dec[I2(icol, ilay)] = ((Float)1 - tr) * lev_lo + (Float)2 * fact * (lay - lev_lo);
    }
  }
}

/* lw_transport_noscat_dn: :681-708 */
static void lw_transport_noscat_dn(int ncol, int nlay, int top_at_1, const Float* trans,
                                   const Float* source_dn, Float* radn_dn) {
  if (top_at_1) {
    for (int ilev = 1; ilev <= nlay; ++ilev)
      for (int i = 0; i < ncol; ++i)
        radn_dn[I2(i, ilev)] = trans[I2(i, ilev - 1)] * radn_dn[I2(i, ilev - 1)] + source_dn[I2(i, ilev - 1)];
  } else {
    for (int ilev = nlay - 1; ilev >= 0; --ilev)
      for (int i = 0; i < ncol; ++i)
        radn_dn[I2(i, ilev)] = trans[I2(i, ilev)] * radn_dn[I2(i, ilev + 1)] + source_dn[I2(i, ilev)];
  }
}

/* lw_transport_noscat_up: :710-745 */
static void lw_transport_noscat_up(int ncol, int nlay, int top_at_1, const Float* trans,
                                   const Float* source_up, Float* radn_up, int do_Jacobians,
                                   Float* radn_upJac) {
  if (top_at_1) {
    for (int ilev = nlay - 1; ilev >= 0; --ilev) {
      for (int i = 0; i < ncol; ++i)
        radn_up[I2(i, ilev)] = trans[I2(i, ilev)] * radn_up[I2(i, ilev + 1)] + source_up[I2(i, ilev)];
      if (do_Jacobians)
        for (int i = 0; i < ncol; ++i)
          radn_upJac[I2(i, ilev)] = trans[I2(i, ilev)] * radn_upJac[I2(i, ilev + 1)];
    }
  } else {
    for (int ilev = 1; ilev <= nlay; ++ilev) {
      for (int i = 0; i < ncol; ++i)
        radn_up[I2(i, ilev)] = trans[I2(i, ilev - 1)] * radn_up[I2(i, ilev - 1)] + source_up[I2(i, ilev - 1)];
      if (do_Jacobians)
        for (int i = 0; i < ncol; ++i)
          radn_upJac[I2(i, ilev)] = trans[I2(i, ilev - 1)] * radn_upJac[I2(i, ilev - 1)];
    }
  }
}

/* lw_transport_1rescl: :753-844 (Tang et al. 2018 adjustment; orientation-asymmetric as written) */
static void lw_transport_1rescl(int ncol, int nlay, int top_at_1, const Float* trans,
                                const Float* source_dn, const Float* source_up, Float* radn_up,
                                Float* radn_dn, const Float* An, const Float* Cn, int do_Jacobians,
                                Float* radn_up_Jac) {
  if (top_at_1) {
    for (int ilev = nlay - 1; ilev >= 0; --ilev) { /* :784-793 */
      for (int i = 0; i < ncol; ++i) {
        const Float adj = Cn[I2(i, ilev)] * (An[I2(i, ilev)] * radn_dn[I2(i, ilev)] -
                                             trans[I2(i, ilev)] * source_dn[I2(i, ilev)] - source_up[I2(i, ilev)]);
        radn_up[I2(i, ilev)] = trans[I2(i, ilev)] * radn_up[I2(i, ilev + 1)] + source_up[I2(i, ilev)] + adj;
      }
      if (do_Jacobians)
        for (int i = 0; i < ncol; ++i)
          radn_up_Jac[I2(i, ilev)] = trans[I2(i, ilev)] * radn_up_Jac[I2(i, ilev + 1)];
    }
    for (int ilev = 0; ilev < nlay; ++ilev) { /* :798-808 */
      for (int i = 0; i < ncol; ++i) {
        const Float adj = Cn[I2(i, ilev)] * (An[I2(i, ilev)] * radn_up[I2(i, ilev)] -
                                             trans[I2(i, ilev)] * source_up[I2(i, ilev)] - source_dn[I2(i, ilev)]);
        radn_dn[I2(i, ilev + 1)] = trans[I2(i, ilev)] * radn_dn[I2(i, ilev)] + source_dn[I2(i, ilev)] + adj;
      }
    }
  } else {
    for (int ilev = 0; ilev < nlay; ++ilev) { /* :816-826 */
      for (int i = 0; i < ncol; ++i) {
        const Float adj = Cn[I2(i, ilev)] * (An[I2(i, ilev)] * radn_dn[I2(i, ilev + 1)] -
                                             trans[I2(i, ilev)] * source_dn[I2(i, ilev)] - source_up[I2(i, ilev)]);
        radn_up[I2(i, ilev + 1)] = trans[I2(i, ilev)] * radn_up[I2(i, ilev)] + source_up[I2(i, ilev)] + adj;
      }
      if (do_Jacobians)
        for (int i = 0; i < ncol; ++i)
          radn_up_Jac[I2(i, ilev + 1)] = trans[I2(i, ilev)] * radn_up_Jac[I2(i, ilev)];
    }
    for (int ilev = nlay - 1; ilev >= 0; --ilev) { /* :832-842 */
      for (int i = 0; i < ncol; ++i) {
        const Float adj = Cn[I2(i, ilev)] * (An[I2(i, ilev)] * radn_up[I2(i, ilev)] -
                                             trans[I2(i, ilev)] * source_up[I2(i, ilev)] - source_dn[I2(i, ilev)]);
        radn_dn[I2(i, ilev)] = trans[I2(i, ilev)] * radn_dn[I2(i, ilev + 1)] + source_dn[I2(i, ilev)] + adj;
      }
    }
  }
}

/* lw_solver_noscat_oneangle: :51-240 */
static void lw_solver_noscat_oneangle(int ncol, int nlay, int ngpt, int top_at_1, const Float* D,
                                      Float weight, const Float* tau, const Float* lay_source,
                                      const Float* lev_source, const Float* sfc_emis,
                                      const Float* sfc_src, const Float* incident_flux,
                                      Float* flux_up, Float* flux_dn, int do_broadband,
                                      Float* broadband_up, Float* broadband_dn, int do_Jacobians,
                                      const Float* sfc_srcJac, Float* flux_upJac, int do_rescaling,
                                      const Float* ssa, const Float* g) {
  const Float pi = ref_pi();
  const size_t n2 = (size_t)ncol * nlay, n2p = (size_t)ncol * (nlay + 1);
  Float* tau_loc = malloc(sizeof(Float) * n2);
  Float* trans = malloc(sizeof(Float) * n2);
  Float* source_dn = malloc(sizeof(Float) * n2);
  Float* source_up = malloc(sizeof(Float) * n2);
  Float* An = malloc(sizeof(Float) * n2);
  Float* Cn = malloc(sizeof(Float) * n2);
  Float* loc_flux_up = malloc(sizeof(Float) * n2p);
  Float* loc_flux_dn = malloc(sizeof(Float) * n2p);
  Float* gpt_flux_Jac = malloc(sizeof(Float) * n2p);
  const int top_level = top_at_1 ? 0 : nlay;
  const int sfc_level = top_at_1 ? nlay : 0;

  if (do_broadband) { /* :125-128 */
    memset(broadband_up, 0, sizeof(Float) * n2p);
    memset(broadband_dn, 0, sizeof(Float) * n2p);
  }
  if (do_Jacobians) memset(flux_upJac, 0, sizeof(Float) * n2p);

  for (int igpt = 0; igpt < ngpt; ++igpt) {
    Float* gpt_flux_up = do_broadband ? loc_flux_up : flux_up + (size_t)igpt * n2p;
    Float* gpt_flux_dn = do_broadband ? loc_flux_dn : flux_dn + (size_t)igpt * n2p;
    for (int i = 0; i < ncol; ++i) /* :144 */
      gpt_flux_dn[I2(i, top_level)] = incident_flux[I2(i, igpt)] / (pi * weight);
    if (do_rescaling) { /* :148-178 */
      for (int ilay = 0; ilay < nlay; ++ilay) {
        for (int i = 0; i < ncol; ++i) {
          const Float ssal = ssa[I3(i, ilay, igpt, nlay)];
          const Float wb = ssal * ((Float)1 - g[I3(i, ilay, igpt, nlay)]) * (Float)0.5;
          const Float scaleTau = ((Float)1 - ssal + wb);
          Cn[I2(i, ilay)] = (Float)0.4 * wb / scaleTau;
          tau_loc[I2(i, ilay)] = tau[I3(i, ilay, igpt, nlay)] * D[I2(i, igpt)] * scaleTau;
        }
        for (int i = 0; i < ncol; ++i) {
          trans[I2(i, ilay)] = (Float)exp(-tau_loc[I2(i, ilay)]);
          An[I2(i, ilay)] = ((Float)1 - trans[I2(i, ilay)] * trans[I2(i, ilay)]);
        }
      }
    } else { /* :180-183 */
      for (int ilay = 0; ilay < nlay; ++ilay)
        for (int i = 0; i < ncol; ++i) {
          tau_loc[I2(i, ilay)] = tau[I3(i, ilay, igpt, nlay)] * D[I2(i, igpt)];
          trans[I2(i, ilay)] = (Float)exp(-tau_loc[I2(i, ilay)]);
        }
    }
    lw_source_noscat(ncol, nlay, top_at_1, lay_source + (size_t)igpt * n2,
                     lev_source + (size_t)igpt * n2p, tau_loc, trans, source_dn, source_up);
    lw_transport_noscat_dn(ncol, nlay, top_at_1, trans, source_dn, gpt_flux_dn);
    for (int i = 0; i < ncol; ++i) { /* :198-202 */
      const Float sfc_albedo = (Float)1 - sfc_emis[I2(i, igpt)];
      gpt_flux_up[I2(i, sfc_level)] =
          gpt_flux_dn[I2(i, sfc_level)] * sfc_albedo + sfc_emis[I2(i, igpt)] * sfc_src[I2(i, igpt)];
      if (do_Jacobians) gpt_flux_Jac[I2(i, sfc_level)] = sfc_emis[I2(i, igpt)] * sfc_srcJac[I2(i, igpt)];
    }
    if (do_rescaling)
      lw_transport_1rescl(ncol, nlay, top_at_1, trans, source_dn, source_up, gpt_flux_up, gpt_flux_dn,
                          An, Cn, do_Jacobians, gpt_flux_Jac);
    else
      lw_transport_noscat_up(ncol, nlay, top_at_1, trans, source_up, gpt_flux_up, do_Jacobians,
                             gpt_flux_Jac);
    if (do_broadband) { /* :216-218 */
      for (size_t k = 0; k < n2p; ++k) {
        broadband_up[k] = broadband_up[k] + gpt_flux_up[k];
        broadband_dn[k] = broadband_dn[k] + gpt_flux_dn[k];
      }
    } else { /* :223-224 */
      for (size_t k = 0; k < n2p; ++k) {
        gpt_flux_dn[k] = pi * weight * gpt_flux_dn[k];
        gpt_flux_up[k] = pi * weight * gpt_flux_up[k];
      }
    }
    if (do_Jacobians)
      for (size_t k = 0; k < n2p; ++k) flux_upJac[k] = flux_upJac[k] + gpt_flux_Jac[k];
  }
  if (do_broadband) { /* :233-236 */
    for (size_t k = 0; k < n2p; ++k) {
      broadband_up[k] = pi * weight * broadband_up[k];
      broadband_dn[k] = pi * weight * broadband_dn[k];
    }
  }
  if (do_Jacobians)
    for (size_t k = 0; k < n2p; ++k) flux_upJac[k] = pi * weight * flux_upJac[k];
  free(tau_loc); free(trans); free(source_dn); free(source_up); free(An); free(Cn);
  free(loc_flux_up); free(loc_flux_dn); free(gpt_flux_Jac);
}

/* lw_solver_noscat: :248-367 */
void rte_lw_solver_noscat(const int* ncol_, const int* nlay_, const int* ngpt_, const Bool* top_at_1,
                          const int* nmus_, const Float* Ds, const Float* weights, const Float* tau,
                          const Float* lay_source, const Float* lev_source, const Float* sfc_emis,
                          const Float* sfc_src, const Float* inc_flux, Float* flux_up, Float* flux_dn,
                          const Bool* do_broadband, Float* broadband_up, Float* broadband_dn,
                          const Bool* do_Jacobians, const Float* sfc_srcJac, Float* flux_upJac,
                          const Bool* do_rescaling, const Float* ssa, const Float* g) {
  const int ncol = *ncol_, nlay = *nlay_, ngpt = *ngpt_, nmus = *nmus_;
  const size_t n2p = (size_t)ncol * (nlay + 1), n3p = n2p * ngpt;
  lw_solver_noscat_oneangle(ncol, nlay, ngpt, *top_at_1, Ds, weights[0], tau, lay_source, lev_source,
                            sfc_emis, sfc_src, inc_flux, flux_up, flux_dn, *do_broadband, broadband_up,
                            broadband_dn, *do_Jacobians, sfc_srcJac, flux_upJac, *do_rescaling, ssa, g);
  if (nmus <= 1) return;
  Float *this_flux_up = flux_up, *this_flux_dn = flux_dn;
  Float *this_bb_up = broadband_up, *this_bb_dn = broadband_dn, *this_Jac = flux_upJac;
  if (*do_broadband) { /* :326-336 */
    this_bb_up = malloc(sizeof(Float) * n2p);
    this_bb_dn = malloc(sizeof(Float) * n2p);
  } else {
    this_flux_up = malloc(sizeof(Float) * n3p);
    this_flux_dn = malloc(sizeof(Float) * n3p);
  }
  if (*do_Jacobians) this_Jac = malloc(sizeof(Float) * n2p);
  for (int imu = 1; imu < nmus; ++imu) { /* :343-361 */
    lw_solver_noscat_oneangle(ncol, nlay, ngpt, *top_at_1, Ds + (size_t)imu * ncol * ngpt, weights[imu],
                              tau, lay_source, lev_source, sfc_emis, sfc_src, inc_flux, this_flux_up,
                              this_flux_dn, *do_broadband, this_bb_up, this_bb_dn, *do_Jacobians,
                              sfc_srcJac, this_Jac, *do_rescaling, ssa, g);
    if (*do_broadband) {
      for (size_t k = 0; k < n2p; ++k) {
        broadband_up[k] = broadband_up[k] + this_bb_up[k];
        broadband_dn[k] = broadband_dn[k] + this_bb_dn[k];
      }
    } else {
      for (size_t k = 0; k < n3p; ++k) {
        flux_up[k] = flux_up[k] + this_flux_up[k];
        flux_dn[k] = flux_dn[k] + this_flux_dn[k];
      }
    }
    if (*do_Jacobians)
      for (size_t k = 0; k < n2p; ++k) flux_upJac[k] = flux_upJac[k] + this_Jac[k];
  }
  if (*do_broadband) { free(this_bb_up); free(this_bb_dn); }
  else { free(this_flux_up); free(this_flux_dn); }
  if (*do_Jacobians) free(this_Jac);
}

/* ------------------------------------------------------------------------------------------
 * adding: :1135-1245 (Shonk & Hogan 2008); shared by LW and SW two-stream
 */
static void adding(int ncol, int nlay, int top_at_1, const Float* albedo_sfc, const Float* rdif,
                   const Float* tdif, const Float* src_dn, const Float* src_up, const Float* src_sfc,
                   Float* flux_up, Float* flux_dn) {
  const size_t n2 = (size_t)ncol * nlay, n2p = (size_t)ncol * (nlay + 1);
  Float* albedo = malloc(sizeof(Float) * n2p);
  Float* src = malloc(sizeof(Float) * n2p);
  Float* denom = malloc(sizeof(Float) * n2);
  if (top_at_1) {
    int ilev = nlay;
    for (int i = 0; i < ncol; ++i) { albedo[I2(i, ilev)] = albedo_sfc[i]; src[I2(i, ilev)] = src_sfc[i]; }
    for (ilev = nlay - 1; ilev >= 0; --ilev) { /* :1174-1186 */
      for (int i = 0; i < ncol; ++i) {
        denom[I2(i, ilev)] = (Float)1 / ((Float)1 - rdif[I2(i, ilev)] * albedo[I2(i, ilev + 1)]);
        albedo[I2(i, ilev)] = rdif[I2(i, ilev)] +
            tdif[I2(i, ilev)] * tdif[I2(i, ilev)] * albedo[I2(i, ilev + 1)] * denom[I2(i, ilev)];
        src[I2(i, ilev)] = src_up[I2(i, ilev)] +
            tdif[I2(i, ilev)] * denom[I2(i, ilev)] *
                (src[I2(i, ilev + 1)] + albedo[I2(i, ilev + 1)] * src_dn[I2(i, ilev)]);
      }
    }
    ilev = 0;
    for (int i = 0; i < ncol; ++i) /* :1190 */
      flux_up[I2(i, ilev)] = flux_dn[I2(i, ilev)] * albedo[I2(i, ilev)] + src[I2(i, ilev)];
    for (ilev = 1; ilev <= nlay; ++ilev) { /* :1196-1202 */
      for (int i = 0; i < ncol; ++i) {
        flux_dn[I2(i, ilev)] = (tdif[I2(i, ilev - 1)] * flux_dn[I2(i, ilev - 1)] +
                                rdif[I2(i, ilev - 1)] * src[I2(i, ilev)] + src_dn[I2(i, ilev - 1)]) *
                               denom[I2(i, ilev - 1)];
        flux_up[I2(i, ilev)] = flux_dn[I2(i, ilev)] * albedo[I2(i, ilev)] + src[I2(i, ilev)];
      }
    }
  } else {
    int ilev = 0;
    for (int i = 0; i < ncol; ++i) { albedo[I2(i, ilev)] = albedo_sfc[i]; src[I2(i, ilev)] = src_sfc[i]; }
    for (ilev = 0; ilev < nlay; ++ilev) { /* :1214-1226 */
      for (int i = 0; i < ncol; ++i) {
        denom[I2(i, ilev)] = (Float)1 / ((Float)1 - rdif[I2(i, ilev)] * albedo[I2(i, ilev)]);
        albedo[I2(i, ilev + 1)] = rdif[I2(i, ilev)] +
            tdif[I2(i, ilev)] * tdif[I2(i, ilev)] * albedo[I2(i, ilev)] * denom[I2(i, ilev)];
        src[I2(i, ilev + 1)] = src_up[I2(i, ilev)] +
            tdif[I2(i, ilev)] * denom[I2(i, ilev)] *
                (src[I2(i, ilev)] + albedo[I2(i, ilev)] * src_dn[I2(i, ilev)]);
      }
    }
    ilev = nlay;
    for (int i = 0; i < ncol; ++i) /* :1230 */
      flux_up[I2(i, ilev)] = flux_dn[I2(i, ilev)] * albedo[I2(i, ilev)] + src[I2(i, ilev)];
    for (ilev = nlay - 1; ilev >= 0; --ilev) { /* :1236-1243 */
      for (int i = 0; i < ncol; ++i) {
        flux_dn[I2(i, ilev)] = (tdif[I2(i, ilev)] * flux_dn[I2(i, ilev + 1)] +
                                rdif[I2(i, ilev)] * src[I2(i, ilev)] + src_dn[I2(i, ilev)]) *
                               denom[I2(i, ilev)];
        flux_up[I2(i, ilev)] = flux_dn[I2(i, ilev)] * albedo[I2(i, ilev)] + src[I2(i, ilev)];
      }
    }
  }
  free(albedo); free(src); free(denom);
}

/* lw_two_stream: :854-909 (Meador & Weaver 1980; Fu et al. 1997 coefficients) */
static void lw_two_stream(int ncol, int nlay, const Float* tau, const Float* w0, const Float* g,
                          Float* gamma1, Float* gamma2, Float* Rdif, Float* Tdif) {
  /* :870  LW_diff_sec = 1.66 is a DEFAULT-REAL (single precision) literal widened to wp */
  const Float LW_diff_sec = (Float)1.66f;
  for (int j = 0; j < nlay; ++j) {
    for (int i = 0; i < ncol; ++i) {
      const size_t ij = I2(i, j);
      gamma1[ij] = LW_diff_sec * ((Float)1 - (Float)0.5 * w0[ij] * ((Float)1 + g[ij]));
      gamma2[ij] = LW_diff_sec * (Float)0.5 * w0[ij] * ((Float)1 - g[ij]);
      const Float k = (Float)sqrt(fmax((double)((gamma1[ij] - gamma2[ij]) * (gamma1[ij] + gamma2[ij])), 1.e-12));
      const Float exp_minusktau = (Float)exp(-tau[ij] * k);
      const Float exp_minus2ktau = exp_minusktau * exp_minusktau;
      const Float RT_term =
          (Float)1 / (k * ((Float)1 + exp_minus2ktau) + gamma1[ij] * ((Float)1 - exp_minus2ktau));
      Rdif[ij] = RT_term * gamma2[ij] * ((Float)1 - exp_minus2ktau);
      Tdif[ij] = RT_term * (Float)2 * k * exp_minusktau;
    }
  }
}

/* lw_source_2str: :917-967 (Toon et al. 1989 eqs 26-27, as in ecRad) */
static void lw_source_2str(int ncol, int nlay, int top_at_1, const Float* sfc_emis, const Float* sfc_src,
                           const Float* lay_source, const Float* lev_source, const Float* gamma1,
                           const Float* gamma2, const Float* rdif, const Float* tdif, const Float* tau,
                           Float* source_dn, Float* source_up, Float* source_sfc) {
  const Float pi = ref_pi();
  (void)lay_source;
  for (int ilay = 0; ilay < nlay; ++ilay) {
    const Float* lev_top = top_at_1 ? lev_source + I2(0, ilay) : lev_source + I2(0, ilay + 1);
    const Float* lev_bot = top_at_1 ? lev_source + I2(0, ilay + 1) : lev_source + I2(0, ilay);
    for (int i = 0; i < ncol; ++i) {
      const size_t ij = I2(i, ilay);
      if (tau[ij] > (Float)1.0e-8) {
        const Float Z = (lev_bot[i] - lev_top[i]) / (tau[ij] * (gamma1[ij] + gamma2[ij]));
        const Float Zup_top = Z + lev_top[i];
        const Float Zup_bottom = Z + lev_bot[i];
        const Float Zdn_top = -Z + lev_top[i];
        const Float Zdn_bottom = -Z + lev_bot[i];
        source_up[ij] = pi * (Zup_top - rdif[ij] * Zdn_top - tdif[ij] * Zup_bottom);
        source_dn[ij] = pi * (Zdn_bottom - rdif[ij] * Zup_bottom - tdif[ij] * Zdn_top);
      } else {
        source_up[ij] = 0;
        source_dn[ij] = 0;
      }
    }
  }
  for (int i = 0; i < ncol; ++i) source_sfc[i] = pi * sfc_emis[i] * sfc_src[i];
}

/* lw_solver_2stream: :377-440.
 * QUIRK kept (reference default kernels): the rank-3 lev_source is passed whole to the rank-2
 * dummy of lw_source_2str (:422), so by sequence association EVERY g-point uses g-point 1's level
 * source.  The accel kernels index by g-point (accel/mo_rte_solver_kernels.F90:958-962);
 * oracle_set_lw_2stream_lev_source_per_gpt(1) selects that behaviour instead. */
void rte_lw_solver_2stream(const int* ncol_, const int* nlay_, const int* ngpt_, const Bool* top_at_1,
                           const Float* tau, const Float* ssa, const Float* g, const Float* lay_source,
                           const Float* lev_source, const Float* sfc_emis, const Float* sfc_src,
                           const Float* inc_flux, Float* flux_up, Float* flux_dn) {
  const int ncol = *ncol_, nlay = *nlay_, ngpt = *ngpt_;
  const size_t n2 = (size_t)ncol * nlay, n2p = (size_t)ncol * (nlay + 1);
  Float* Rdif = malloc(sizeof(Float) * n2);
  Float* Tdif = malloc(sizeof(Float) * n2);
  Float* gamma1 = malloc(sizeof(Float) * n2);
  Float* gamma2 = malloc(sizeof(Float) * n2);
  Float* source_dn = malloc(sizeof(Float) * n2);
  Float* source_up = malloc(sizeof(Float) * n2);
  Float* sfc_albedo = malloc(sizeof(Float) * ncol);
  Float* source_sfc = malloc(sizeof(Float) * ncol);
  const int top_level = *top_at_1 ? 0 : nlay;
  for (int igpt = 0; igpt < ngpt; ++igpt) {
    lw_two_stream(ncol, nlay, tau + igpt * n2, ssa + igpt * n2, g + igpt * n2, gamma1, gamma2, Rdif, Tdif);
    const Float* lev = g_lw2s_lev_source_per_gpt ? lev_source + igpt * n2p : lev_source;
    lw_source_2str(ncol, nlay, *top_at_1, sfc_emis + (size_t)igpt * ncol, sfc_src + (size_t)igpt * ncol,
                   lay_source + igpt * n2, lev, gamma1, gamma2, Rdif, Tdif, tau + igpt * n2, source_dn,
                   source_up, source_sfc);
    for (int i = 0; i < ncol; ++i) sfc_albedo[i] = (Float)1 - sfc_emis[I2(i, igpt)];
    for (int i = 0; i < ncol; ++i) flux_dn[igpt * n2p + I2(i, top_level)] = inc_flux[I2(i, igpt)];
    adding(ncol, nlay, *top_at_1, sfc_albedo, Rdif, Tdif, source_dn, source_up, source_sfc,
           flux_up + igpt * n2p, flux_dn + igpt * n2p);
  }
  free(Rdif); free(Tdif); free(gamma1); free(gamma2); free(source_dn); free(source_up);
  free(sfc_albedo); free(source_sfc);
}

/* sw_solver_noscat: :450-494 (direct beam only) */
void rte_sw_solver_noscat(const int* ncol_, const int* nlay_, const int* ngpt_, const Bool* top_at_1,
                          const Float* tau, const Float* mu0, const Float* inc_flux_dir, Float* flux_dir) {
  const int ncol = *ncol_, nlay = *nlay_, ngpt = *ngpt_;
  if (*top_at_1) {
    for (int igpt = 0; igpt < ngpt; ++igpt) {
      for (int i = 0; i < ncol; ++i)
        flux_dir[I3(i, 0, igpt, nlay + 1)] = inc_flux_dir[I2(i, igpt)] * mu0[I2(i, 0)];
      for (int ilev = 1; ilev <= nlay; ++ilev)
        for (int i = 0; i < ncol; ++i)
          flux_dir[I3(i, ilev, igpt, nlay + 1)] =
              flux_dir[I3(i, ilev - 1, igpt, nlay + 1)] *
              (Float)exp(-tau[I3(i, ilev - 1, igpt, nlay)] / mu0[I2(i, ilev - 1)]);
    }
  } else {
    for (int igpt = 0; igpt < ngpt; ++igpt) {
      for (int i = 0; i < ncol; ++i)
        flux_dir[I3(i, nlay, igpt, nlay + 1)] = inc_flux_dir[I2(i, igpt)] * mu0[I2(i, nlay - 1)];
      for (int ilev = nlay - 1; ilev >= 0; --ilev)
        for (int i = 0; i < ncol; ++i)
          flux_dir[I3(i, ilev, igpt, nlay + 1)] =
              flux_dir[I3(i, ilev + 1, igpt, nlay + 1)] *
              (Float)exp(-tau[I3(i, ilev, igpt, nlay)] / mu0[I2(i, ilev)]);
    }
  }
}

/* sw_dif_and_source: :985-1127 (Zdunkowski PIFM two-stream + direct-beam source) */
static void sw_dif_and_source(int ncol, int nlay, int top_at_1, const Float* mu0, const Float* sfc_albedo,
                              const Float* tau, const Float* w0, const Float* g, Float* Rdif, Float* Tdif,
                              Float* source_dn, Float* source_up, Float* source_sfc, Float* flux_dn_dir) {
  const Float eps = (sizeof(Float) == 8) ? (Float)DBL_EPSILON : (Float)FLT_EPSILON;
  const Float min_k = (Float)1.e4 * eps;    /* :1005 */
  const Float min_mu0 = (Float)sqrt(eps);   /* :1006 */
  int lay_index = 0;
  Float* dir_flux_trans = NULL;
  for (int j = 0; j < nlay; ++j) {
    Float* dir_flux_inc;
    if (top_at_1) {
      lay_index = j;
      dir_flux_inc = flux_dn_dir + I2(0, lay_index);
      dir_flux_trans = flux_dn_dir + I2(0, lay_index + 1);
    } else {
      lay_index = nlay - j - 1;
      dir_flux_inc = flux_dn_dir + I2(0, lay_index + 1);
      dir_flux_trans = flux_dn_dir + I2(0, lay_index);
    }
    for (int i = 0; i < ncol; ++i) {
      const size_t ij = I2(i, lay_index);
      const Float tau_s = tau[ij], w0_s = w0[ij], g_s = g[ij];
      const Float gamma1 = ((Float)8 - w0_s * ((Float)5 + (Float)3 * g_s)) * (Float).25; /* :1038 */
      const Float gamma2 = (Float)3 * (w0_s * ((Float)1 - g_s)) * (Float).25;            /* :1039 */
      const Float k = (Float)sqrt(fmax((double)((gamma1 - gamma2) * (gamma1 + gamma2)), (double)min_k));
      const Float exp_minusktau = (Float)exp(-tau_s * k);
      const Float exp_minus2ktau = exp_minusktau * exp_minusktau;
      Float RT_term = (Float)1 / (k * ((Float)1 + exp_minus2ktau) + gamma1 * ((Float)1 - exp_minus2ktau));
      Rdif[ij] = RT_term * gamma2 * ((Float)1 - exp_minus2ktau); /* :1055 */
      Tdif[ij] = RT_term * (Float)2 * k * exp_minusktau;         /* :1058 */
      const Float mu0_s = (min_mu0 > mu0[ij]) ? min_mu0 : mu0[ij]; /* :1065 */
      const Float k_mu = k * mu0_s;
      const Float om = (Float)1 - k_mu * k_mu;
      RT_term = w0_s * RT_term / ((Float)fabs((double)om) >= eps ? om : eps); /* :1071-1073 */
      const Float gamma3 = ((Float)2 - (Float)3 * mu0_s * g_s) * (Float).25;
      const Float gamma4 = (Float)1 - gamma3;
      const Float alpha1 = gamma1 * gamma4 + gamma2 * gamma3;
      const Float alpha2 = gamma1 * gamma3 + gamma2 * gamma4;
      const Float k_gamma3 = k * gamma3;
      const Float k_gamma4 = k * gamma4;
      const Float Tnoscat = (Float)exp(-tau_s / mu0_s);
      Float Rdir = RT_term * (((Float)1 - k_mu) * (alpha2 + k_gamma3) -
                              ((Float)1 + k_mu) * (alpha2 - k_gamma3) * exp_minus2ktau -
                              (Float)2.0 * (k_gamma3 - alpha2 * k_mu) * exp_minusktau * Tnoscat);
      Float Tdir = -RT_term * (((Float)1 + k_mu) * (alpha1 + k_gamma4) * Tnoscat -
                               ((Float)1 - k_mu) * (alpha1 - k_gamma4) * exp_minus2ktau * Tnoscat -
                               (Float)2.0 * (k_gamma4 + alpha1 * k_mu) * exp_minusktau);
      /* :1107-1108 */
      { const Float lim = (Float)1.0 - Tnoscat; Rdir = (Rdir < lim) ? Rdir : lim; Rdir = (Rdir > 0) ? Rdir : 0; }
      { const Float lim = (Float)1.0 - Tnoscat - Rdir; Tdir = (Tdir < lim) ? Tdir : lim; Tdir = (Tdir > 0) ? Tdir : 0; }
      source_up[ij] = Rdir * dir_flux_inc[i];
      source_dn[ij] = Tdir * dir_flux_inc[i];
      dir_flux_trans[i] = Tnoscat * dir_flux_inc[i];
    }
  }
  /* :1120-1125  (lay_index and dir_flux_trans keep their last-loop values: the surface layer) */
  for (int i = 0; i < ncol; ++i)
    source_sfc[i] = (mu0[I2(i, lay_index)] > 0) ? dir_flux_trans[i] * sfc_albedo[i] : (Float)0;
  for (size_t k = 0; k < (size_t)ncol * nlay; ++k)
    if (mu0[k] <= 0) { source_up[k] = 0; source_dn[k] = 0; }
}

/* sw_solver_2stream: :503-609 */
void rte_sw_solver_2stream(const int* ncol_, const int* nlay_, const int* ngpt_, const Bool* top_at_1,
                           const Float* tau, const Float* ssa, const Float* g, const Float* mu0,
                           const Float* sfc_alb_dir, const Float* sfc_alb_dif, const Float* inc_flux_dir,
                           Float* flux_up, Float* flux_dn, Float* flux_dir, const Bool* has_dif_bc,
                           const Float* inc_flux_dif, const Bool* do_broadband, Float* broadband_up,
                           Float* broadband_dn, Float* broadband_dir) {
  const int ncol = *ncol_, nlay = *nlay_, ngpt = *ngpt_;
  const size_t n2 = (size_t)ncol * nlay, n2p = (size_t)ncol * (nlay + 1);
  Float* Rdif = malloc(sizeof(Float) * n2);
  Float* Tdif = malloc(sizeof(Float) * n2);
  Float* source_up = malloc(sizeof(Float) * n2);
  Float* source_dn = malloc(sizeof(Float) * n2);
  Float* source_srf = malloc(sizeof(Float) * ncol);
  Float* loc_up = malloc(sizeof(Float) * n2p);
  Float* loc_dn = malloc(sizeof(Float) * n2p);
  Float* loc_dir = malloc(sizeof(Float) * n2p);
  const int top_level = *top_at_1 ? 0 : nlay;
  const int top_layer = *top_at_1 ? 0 : nlay - 1;
  if (*do_broadband) {
    memset(broadband_up, 0, sizeof(Float) * n2p);
    memset(broadband_dn, 0, sizeof(Float) * n2p);
    memset(broadband_dir, 0, sizeof(Float) * n2p);
  }
  for (int igpt = 0; igpt < ngpt; ++igpt) {
    Float* gup = *do_broadband ? loc_up : flux_up + igpt * n2p;
    Float* gdn = *do_broadband ? loc_dn : flux_dn + igpt * n2p;
    Float* gdir = *do_broadband ? loc_dir : flux_dir + igpt * n2p;
    for (int i = 0; i < ncol; ++i) /* :575 */
      gdir[I2(i, top_level)] = inc_flux_dir[I2(i, igpt)] * mu0[I2(i, top_layer)];
    for (int i = 0; i < ncol; ++i) /* :579-583 */
      gdn[I2(i, top_level)] = *has_dif_bc ? inc_flux_dif[I2(i, igpt)] : (Float)0;
    sw_dif_and_source(ncol, nlay, *top_at_1, mu0, sfc_alb_dir + (size_t)igpt * ncol, tau + igpt * n2,
                      ssa + igpt * n2, g + igpt * n2, Rdif, Tdif, source_dn, source_up, source_srf, gdir);
    adding(ncol, nlay, *top_at_1, sfc_alb_dif + (size_t)igpt * ncol, Rdif, Tdif, source_dn, source_up,
           source_srf, gup, gdn);
    if (*do_broadband) { /* :601-604 */
      for (size_t k = 0; k < n2p; ++k) {
        broadband_up[k] = broadband_up[k] + gup[k];
        broadband_dn[k] = broadband_dn[k] + gdn[k] + gdir[k];
        broadband_dir[k] = broadband_dir[k] + gdir[k];
      }
    } else { /* :606 */
      for (size_t k = 0; k < n2p; ++k) gdn[k] = gdn[k] + gdir[k];
    }
  }
  free(Rdif); free(Tdif); free(source_up); free(source_dn); free(source_srf);
  free(loc_up); free(loc_dn); free(loc_dir);
}
