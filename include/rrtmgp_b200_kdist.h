/* rrtmgp_b200_kdist.h - k-distribution ingestion without netCDF.
 *
 * The reference reads a k-distribution file into plain arrays (rrtmgp/data-loading-examples/
 * mo_optics_utils_rrtmgp.F90:41-246, load_gas_optics) and hands them to ty_gas_optics_rrtmgp%load
 * (rrtmgp/frontend/mo_gas_optics_rrtmgp.F90:938-1145), whose init_abs_coeffs (:1151-1381) REDUCES them to the gases the
 * host model provides and re-lays them out for the kernels:
 *     vmr_ref            -> (2, 0:ngas_red, ntemp)                        :1237-1245
 *     kminor_lower/upper -> reduce_minor_arrays, transposed to (ntemp, neta, ncontrib_red)   :1790-1907
 *     kmajor, plank_fraction -> (ntemp, neta, npres+1, ngpt)              :1299, 1024-1026
 *     rayl_lower/upper   -> krayl(ntemp, neta, ngpt, 2)                   :1308-1316
 *     key_species        -> create_key_species_reduce :1752, create_flavor :1598, create_gpoint_flavor :1930
 *     minor_gases / scaling_gas (strings) -> create_idx_minor :1637, create_idx_minor_scaling :1661
 *     press_ref_log, *_delta, *_min/_max, press_ref_trop_log              :1320-1358
 *     solar_source = quiet + (mg - 0.1495954) facular + (sb - 0.00066696) sunspot   :760-798 (no TSI renormalisation)
 * This header is that step for C / C++ / Python hosts: rrtmgpb_kdist_raw mirrors the ON-DISK variable set, array for
 * array, in the Fortran shapes read_field returns them in (first index fastest); rrtmgpb_kdist_reduce produces the
 * rrtmgpb_kdist that rrtmgpb_gas_optics_load (rrtmgp_b200_frontend.h) copies to the device.  Reading the bytes out of
 * an HDF5/netCDF-4 container is left to the host (no such reader exists in this image; numpy/.npz in the tests).
 * Strings are NUL-terminated; comparisons are case-insensitive on trimmed strings like string_loc_in_array
 * (rte/frontend/gas-optics-template/mo_gas_optics_util_string.F90:71-87).  All functions are host-only logic. */
#ifndef RRTMGP_B200_KDIST_H
#define RRTMGP_B200_KDIST_H
#include "rrtmgp_b200_frontend.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  /* dimensions (mo_optics_utils_rrtmgp.F90:102-120) */
  int ntemp, npres, nabsorbers, nminorabsorbers, nextabsorbers, nmixingfracs, nlayers /* atmos_layer = 2 */, nbnd, ngpt;
  int nminor_absorber_intervals_lower, nminor_absorber_intervals_upper, ncontributors_lower, ncontributors_upper;
  int ntemp_planck /* temperature_Planck; 0 for SW */, nfit_coeffs /* 0 for SW */;
  /* variables (:126-183) */
  const char* const* gas_names;            /* (absorber) */
  const int* key_species;                  /* (2, atmos_layer, bnd): 1-based indices into gas_names, 0 = none */
  const Float* bnd_limits_wavenumber;      /* (2, bnd) */
  const int* bnd_limits_gpt;               /* (2, bnd) */
  const Float *press_ref, *temp_ref;       /* (pressure), (temperature) */
  Float absorption_coefficient_ref_P, absorption_coefficient_ref_T, press_ref_trop;
  const Float *kminor_lower, *kminor_upper;   /* (contributors_*, mixing_fraction, temperature) */
  const char* const* gas_minor;            /* (minor_absorber) */
  const char* const* identifier_minor;     /* (minor_absorber) */
  const char* const* minor_gases_lower;    /* (minor_absorber_intervals_lower) */
  const char* const* minor_gases_upper;
  const int *minor_limits_gpt_lower, *minor_limits_gpt_upper;   /* (pair, intervals) */
  const Bool *minor_scales_with_density_lower, *minor_scales_with_density_upper;
  const Bool *scale_by_complement_lower, *scale_by_complement_upper;
  const char* const* scaling_gas_lower;
  const char* const* scaling_gas_upper;
  const int *kminor_start_lower, *kminor_start_upper;
  const Float* vmr_ref;                    /* (atmos_layer, absorber_ext, temperature) */
  const Float* kmajor;                     /* (gpt, mixing_fraction, pressure+1, temperature) */
  const Float *rayl_lower, *rayl_upper;    /* (gpt, mixing_fraction, temperature); both NULL when absent */
  /* internal sources (LW): NULL / 0 for SW */
  const Float* totplnk;                    /* (temperature_Planck, bnd) */
  const Float* plank_fraction;             /* (gpt, mixing_fraction, pressure+1, temperature) [sic] */
  const Float* optimal_angle_fit;          /* (fit_coeffs, bnd) */
  /* external sources (SW): NULL for LW */
  const Float *solar_source_quiet, *solar_source_facular, *solar_source_sunspot;   /* (gpt) */
  Float tsi_default, mg_default, sb_default;
} rrtmgpb_kdist_raw;

typedef struct rrtmgpb_kdist_loaded rrtmgpb_kdist_loaded;

/* ty_gas_optics_rrtmgp%load (load_int when raw->totplnk != NULL, else load_ext) for the `navailable` gases the host
 * provides (available_gases%gas_names).  Returns NULL and the reference's error string (e.g. "gas_optics: required
 * gases h2o are not provided", :1395) on failure. */
rrtmgpb_kdist_loaded* rrtmgpb_kdist_reduce(const rrtmgpb_kdist_raw* raw, int navailable,
                                           const char* const* available_gases, char* errmsg);
void rrtmgpb_kdist_loaded_free(rrtmgpb_kdist_loaded* kd);
/* The reduced, kernel-layout tables (HOST arrays owned by `kd`), ready for rrtmgpb_gas_optics_load(). */
const rrtmgpb_kdist* rrtmgpb_kdist_loaded_tables(const rrtmgpb_kdist_loaded* kd);
/* this%gas_names after the reduction (:1232): i in [0, ngas) */
const char* rrtmgpb_kdist_loaded_gas_name(const rrtmgpb_kdist_loaded* kd, int i);
/* this%is_key (:1364-1372) */
int rrtmgpb_kdist_loaded_is_key(const rrtmgpb_kdist_loaded* kd, int i);
/* optimal_angle_fit(nfit_coeffs, nbnd) as loaded (LW), NULL for SW */
const Float* rrtmgpb_kdist_loaded_optimal_angle_fit(const rrtmgpb_kdist_loaded* kd);
/* set_solar_variability (:760-798; tsi < 0: keep the integral) and set_tsi (:800-835) on the loaded SW tables;
 * 0 or 1 + the reference's message ("mg_index out of range", "sb_index out of range", "tsi out of range") */
int rrtmgpb_kdist_set_solar_variability(rrtmgpb_kdist_loaded* kd, Float mg_index, Float sb_index, Float tsi, char* errmsg);
int rrtmgpb_kdist_set_tsi(rrtmgpb_kdist_loaded* kd, Float tsi, char* errmsg);

#ifdef __cplusplus
}
#endif
#endif
