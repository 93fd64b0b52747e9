/* rrtmgp_b200_frontend.h - host-side mirror of the reference's Fortran frontend for the hot path.
 *
 * The reference's frontend is object-oriented Fortran (ty_optical_props_*, ty_source_func_lw,
 * ty_fluxes_broadband, ty_gas_optics_rrtmgp, ty_cloud_optics_rrtmgp, rte_lw, rte_sw).  It stays
 * the plugin API of a Fortran host.  For hosts without a Fortran toolchain - and
 * for this repo's tests and benchmark - the same call sequences are provided here in C++ behind a
 * plain-C interface: same names, same argument meaning, same error behaviour (a 128-character
 * message, empty = success; reference rte/frontend/mo_rte_lw.F90:107-108).
 *
 * Memory: every array pointer in these structs points to BACKEND memory (device memory for the CUDA
 * library; host memory for the CPU oracle build of the same frontend source) unless marked HOST.
 * The frontend never dereferences backend arrays itself - it sequences kernels of
 * rte_kernels.h / rrtmgp_kernels.h / rrtmgp_b200_ext.h, like the Fortran frontend does.
 */
#ifndef RRTMGP_B200_FRONTEND_H
#define RRTMGP_B200_FRONTEND_H

#include "rte_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define RRTMGPB_ERRLEN 128

enum { RRTMGPB_1SCL = 1, RRTMGPB_2STR = 2, RRTMGPB_NSTR = 3 };

/* ty_optical_props_{1scl,2str,nstr}: rte/frontend/mo_optical_props.F90:78-113,183-227 */
typedef struct {
  int kind;                   /* RRTMGPB_1SCL / _2STR / _NSTR */
  int ncol, nlay, ngpt, nband;
  int nmom;                   /* nstr only */
  int top_at_1;               /* set_top_at_1(): is layer 1 the top of the atmosphere? */
  const int* band_lims_gpt;   /* HOST (2,nband), 1-based inclusive; mo_optical_props.F90:183-210 */
  const Float* band_lims_wvn; /* HOST (2,nband), may be NULL (no spectral-consistency checks then) */
  Float* tau;                 /* (ncol,nlay,ngpt) */
  Float* ssa;                 /* 2str, nstr */
  Float* g;                   /* 2str */
  Float* p;                   /* nstr: (nmom,ncol,nlay,ngpt) */
} rrtmgpb_optical_props;

/* ty_source_func_lw: rte/frontend/mo_source_functions.F90:30-38 */
typedef struct {
  int ncol, nlay, ngpt;
  Float* lay_source;     /* (ncol,nlay,ngpt) */
  Float* lev_source;     /* (ncol,nlay+1,ngpt) */
  Float* sfc_source;     /* (ncol,ngpt) */
  Float* sfc_source_Jac; /* (ncol,ngpt) */
} rrtmgpb_source_func_lw;

/* ty_fluxes_broadband: rte/frontend/mo_fluxes.F90:47-54.  NULL = not requested ("not associated"). */
typedef struct {
  Float* flux_up;     /* (ncol,nlay+1) */
  Float* flux_dn;
  Float* flux_net;
  Float* flux_dn_dir; /* SW only */
} rrtmgpb_fluxes_broadband;

/* rte_config_checks(): rte/frontend/mo_rte_config.F90:29-49 */
void rrtmgpb_rte_config_checks(int check_extents, int check_values);

/* ty_optical_props_arry%validate / %delta_scale / %increment: mo_optical_props.F90:478-560,562-613,879-1028.
 * Return 0 on success; otherwise errmsg (RRTMGPB_ERRLEN bytes) holds the reference's message. */
int rrtmgpb_op_validate(const rrtmgpb_optical_props* op, char* errmsg);
int rrtmgpb_op_delta_scale(rrtmgpb_optical_props* op, const Float* forward /* NULL: f = g*g */, char* errmsg);
/* op_io is incremented by op_in ("call op_in%increment(op_io)") */
int rrtmgpb_op_increment(const rrtmgpb_optical_props* op_in, rrtmgpb_optical_props* op_io, char* errmsg);

/* rte_lw(): rte/frontend/mo_rte_lw.F90:79-473.  sfc_emis (nband,ncol); optional: inc_flux (ncol,ngpt),
 * n_gauss_angles (0 = default 1), use_2stream (-1 = default .false.), lw_Ds (ncol,ngpt),
 * flux_up_Jac (ncol,nlay+1).  Fluxes are ty_fluxes_broadband. */
int rrtmgpb_rte_lw(const rrtmgpb_optical_props* optical_props, const rrtmgpb_source_func_lw* sources,
                   const Float* sfc_emis, rrtmgpb_fluxes_broadband* fluxes, const Float* inc_flux,
                   int n_gauss_angles, int use_2stream, const Float* lw_Ds, Float* flux_up_Jac, char* errmsg);
/* As rte_lw with a generic (by-g-point) flux class: returns g-point fluxes (ncol,nlay+1,ngpt).  This is
 * the path lw_solver_2stream needs (it has no broadband outputs). */
int rrtmgpb_rte_lw_bygpoint(const rrtmgpb_optical_props* optical_props, const rrtmgpb_source_func_lw* sources,
                            const Float* sfc_emis, Float* gpt_flux_up, Float* gpt_flux_dn,
                            const Float* inc_flux, int n_gauss_angles, int use_2stream, const Float* lw_Ds,
                            char* errmsg);

/* rte_sw(): rte/frontend/mo_rte_sw.F90:56-394 (mu0 by column).  mu0 (ncol); inc_flux (ncol,ngpt);
 * sfc_alb_dir, sfc_alb_dif (nband,ncol); optional inc_flux_dif (ncol,ngpt). */
int rrtmgpb_rte_sw(const rrtmgpb_optical_props* atmos, const Float* mu0, const Float* inc_flux,
                   const Float* sfc_alb_dir, const Float* sfc_alb_dif, rrtmgpb_fluxes_broadband* fluxes,
                   const Float* inc_flux_dif, char* errmsg);
int rrtmgpb_rte_sw_bygpoint(const rrtmgpb_optical_props* atmos, const Float* mu0, const Float* inc_flux,
                            const Float* sfc_alb_dir, const Float* sfc_alb_dif, Float* gpt_flux_up,
                            Float* gpt_flux_dn, Float* gpt_flux_dir, const Float* inc_flux_dif, char* errmsg);

/* ---------------- ty_gas_optics_rrtmgp ---------------- */
/* Tables as the Fortran object holds them after load() (mo_gas_optics_rrtmgp.F90:46-155,1151-1381).
 * All pointers HOST; they are copied to backend memory once by rrtmgpb_gas_optics_load. */
typedef struct {
  int ngas, nflav, neta, npres, ntemp, nbnd, ngpt;
  int nminorlower, nminorklower, nminorupper, nminorkupper, idx_h2o;
  const int *flavor, *gpoint_flavor, *band_lims_gpt, *gpoint_bands;
  const Float *band_lims_wvn, *press_ref_log, *temp_ref, *vmr_ref;
  Float press_ref_log_delta, temp_ref_min, temp_ref_max, temp_ref_delta, press_ref_min, press_ref_max,
      press_ref_trop_log;
  const Float *kmajor, *kminor_lower, *kminor_upper;
  const int *minor_limits_gpt_lower, *minor_limits_gpt_upper;
  const Bool *minor_scales_with_density_lower, *minor_scales_with_density_upper;
  const Bool *scale_by_complement_lower, *scale_by_complement_upper;
  const int *idx_minor_lower, *idx_minor_upper, *idx_minor_scaling_lower, *idx_minor_scaling_upper;
  const int *kminor_start_lower, *kminor_start_upper;
  /* LW (internal source): NULL for SW */
  const Float *planck_frac, *totplnk;
  int nPlanckTemp;
  Float totplnk_delta;
  /* SW (external source): NULL for LW */
  const Float *krayl, *solar_source;
} rrtmgpb_kdist;

typedef struct rrtmgpb_gas_optics_t rrtmgpb_gas_optics_t;
rrtmgpb_gas_optics_t* rrtmgpb_gas_optics_load(const rrtmgpb_kdist* tables, char* errmsg);
void rrtmgpb_gas_optics_free(rrtmgpb_gas_optics_t* go);
/* Dimensions of a loaded k-distribution (any pointer may be NULL) and the HOST copy of its band_lims_gpt(2,nbnd). */
void rrtmgpb_gas_optics_dims(const rrtmgpb_gas_optics_t* go, int* ngas, int* nbnd, int* ngpt);
const int* rrtmgpb_gas_optics_band_lims_gpt(const rrtmgpb_gas_optics_t* go);
/* Backend address of the loaded kmajor(ntemp,neta,npres+1,ngpt) table.  A host that overwrites the loaded coefficients
 * in place (rrtmgpb_mem_to_backend) passes it to rrtmgpb_tables_changed() afterwards (rrtmgp_b200_ext.h). */
Float* rrtmgpb_gas_optics_kmajor(const rrtmgpb_gas_optics_t* go);
int rrtmgpb_gas_optics_source_is_internal(const rrtmgpb_gas_optics_t* go);

/* gas_optics_int(): mo_gas_optics_rrtmgp.F90:220-330.  play,tlay (ncol,nlay); plev (ncol,nlay+1);
 * tsfc (ncol); vmr (ncol,nlay,ngas) in the k-distribution's gas order (the ty_gas_concs stand-in);
 * optional col_dry (ncol,nlay), tlev (ncol,nlay+1). */
int rrtmgpb_gas_optics_int(const rrtmgpb_gas_optics_t* go, int ncol, int nlay, const Float* play, const Float* plev,
                           const Float* tlay, const Float* tsfc, const Float* vmr,
                           rrtmgpb_optical_props* optical_props, rrtmgpb_source_func_lw* sources,
                           const Float* col_dry, const Float* tlev, char* errmsg);
/* gas_optics_ext(): mo_gas_optics_rrtmgp.F90:337-414.  toa_src (ncol,ngpt). */
int rrtmgpb_gas_optics_ext(const rrtmgpb_gas_optics_t* go, int ncol, int nlay, const Float* play, const Float* plev,
                           const Float* tlay, const Float* vmr, rrtmgpb_optical_props* optical_props,
                           Float* toa_src, const Float* col_dry, char* errmsg);

/* Fused fast path (not in the reference API): gas_optics() followed by clouds%increment(optical_props) in ONE
 * pass that writes only the caller-visible arrays.  `clouds` (by-band, same bands; 1scl or 2str; already
 * delta-scaled if the caller wants that) may be NULL.  Results equal the two separate calls to rounding. */
int rrtmgpb_gas_optics_int_fused(const rrtmgpb_gas_optics_t* go, int ncol, int nlay, const Float* play,
                                 const Float* plev, const Float* tlay, const Float* tsfc, const Float* vmr,
                                 rrtmgpb_optical_props* optical_props, rrtmgpb_source_func_lw* sources,
                                 const Float* col_dry, const Float* tlev, const rrtmgpb_optical_props* clouds,
                                 const rrtmgpb_optical_props* aerosols /* second increment, after clouds; or NULL */,
                                 char* errmsg);
int rrtmgpb_gas_optics_ext_fused(const rrtmgpb_gas_optics_t* go, int ncol, int nlay, const Float* play,
                                 const Float* plev, const Float* tlay, const Float* vmr,
                                 rrtmgpb_optical_props* optical_props, Float* toa_src, const Float* col_dry,
                                 const rrtmgpb_optical_props* clouds, const rrtmgpb_optical_props* aerosols,
                                 char* errmsg);

/* gas_optics() derives top_at_1 from play(1,1) < play(1,nlay) (mo_gas_optics_rrtmgp.F90:258) - two device reads and stream
 * synchronisations per call here.  A driver that knows the orientation may state it for the calling thread (0 / 1;
 * -1: unset again) and keep its launches asynchronous. */
void rrtmgpb_set_top_at_1_hint(int top_at_1);

/* ---------------- express path: state in, broadband fluxes out ---------------- */
/* = gas_optics(play, plev, tlay, tsfc, gas_concs, atmos, sources[, col_dry, tlev]) ; clouds%increment(atmos) ;
 *   rte_lw(atmos, sources, sfc_emis, fluxes[, n_gauss_angles])   with ty_fluxes_broadband
 * (examples/all-sky/rrtmgp_allsky.F90:368-381) without atmos / sources ever existing: no (ncol,nlay,ngpt) array is
 * allocated (include/rrtmgp_b200_ext.h: rrtmgpb_express).  clouds: by-band 1scl / 2str properties or NULL;
 * sfc_emis (nbnd,ncol); the same checks and error strings as the three calls it replaces. */
int rrtmgpb_rte_lw_express(const rrtmgpb_gas_optics_t* go, int ncol, int nlay, const Float* play, const Float* plev,
                           const Float* tlay, const Float* tsfc, const Float* vmr, const Float* col_dry /* or NULL */,
                           const Float* tlev /* or NULL */, const rrtmgpb_optical_props* clouds /* or NULL */,
                           const Float* sfc_emis, int n_gauss_angles, rrtmgpb_fluxes_broadband* fluxes, char* errmsg);
/* = gas_optics(..., atmos, toa_flux) ; clouds%increment(atmos) ; rte_sw(atmos, mu0, toa_flux, sfc_alb_dir, sfc_alb_dif,
 *   fluxes) (rrtmgp_allsky.F90:383-406); clouds are expected delta-scaled already, as there. */
int rrtmgpb_rte_sw_express(const rrtmgpb_gas_optics_t* go, int ncol, int nlay, const Float* play, const Float* plev,
                           const Float* tlay, const Float* vmr, const Float* col_dry /* or NULL */,
                           const rrtmgpb_optical_props* clouds /* or NULL */, const Float* mu0, const Float* sfc_alb_dir,
                           const Float* sfc_alb_dif, rrtmgpb_fluxes_broadband* fluxes, char* errmsg);

/* ---------------- all-sky driver on HOST buffers (CUDA library only) ---------------- */
/* One iteration of the reference's all-sky loop body (examples/all-sky/rrtmgp_allsky.F90:332-409: cloud optics, gas
 * optics, clouds%increment, rte_lw / rte_sw with broadband fluxes) for `ncol` columns whose inputs and outputs live in
 * HOST memory: the columns are processed in chunks of `chunk_cols`, chunk k+1's inputs are uploaded and chunk k-1's
 * fluxes downloaded (separate copy streams, cudaMemcpy2DAsync on the strided column slices) while chunk k computes.
 * Pinned host memory makes the copies asynchronous; pageable memory works, serialised.  This is the entry a host model
 * that keeps its state on the CPU calls, and how a problem larger than device memory (BASELINE config 4) is streamed.
 * Arrays are Fortran-ordered: p_lay, t_lay, lwp, iwp, rel, dei, vmr fields (ncol,nlay); p_lev, t_lev (ncol,nlay+1);
 * t_sfc, mu0 (ncol); emis_sfc (nbnd_lw,ncol), sfc_alb_dir/dif (nbnd_sw,ncol); fluxes (ncol,nlay+1). */
typedef struct {
  int ncol, nlay;
  const Float *p_lay, *p_lev, *t_lay, *t_lev;
  int ngas;                        /* gases in the k-distributions' order (both k-distributions share it) */
  const Float* const* vmr_field;   /* [ngas]: HOST (ncol,nlay) field, or NULL when the gas is well mixed ... */
  const Float* vmr_scalar;         /* [ngas]: ... with this volume mixing ratio */
  const Float *lwp, *iwp, *rel, *dei;   /* all NULL: clear sky */
  const Float *t_sfc, *emis_sfc;        /* LW boundary conditions */
  const Float *mu0, *sfc_alb_dir, *sfc_alb_dif;   /* SW boundary conditions */
} rrtmgpb_allsky_host_inputs;
typedef struct {
  Float *lw_flux_up, *lw_flux_dn, *sw_flux_up, *sw_flux_dn, *sw_flux_dir;   /* HOST (ncol,nlay+1); NULL: not wanted */
} rrtmgpb_allsky_host_fluxes;
/* go_lw / go_sw: either may be NULL (that half is skipped); co_lw / co_sw: cloud optics for the same bands, needed when
 * the cloud inputs are given.  express != 0: rrtmgpb_rte_lw_express / _sw_express per chunk (no (ncol,nlay,ngpt) arrays),
 * else the fused gas optics + rte_lw / rte_sw on chunk-sized planes.  Returns 0, or 1 and the failing call's message. */
struct rrtmgpb_cloud_optics_t;
int rrtmgpb_allsky_stream_host(const rrtmgpb_gas_optics_t* go_lw, const rrtmgpb_gas_optics_t* go_sw,
                               const struct rrtmgpb_cloud_optics_t* co_lw, const struct rrtmgpb_cloud_optics_t* co_sw,
                               const rrtmgpb_allsky_host_inputs* in, const rrtmgpb_allsky_host_fluxes* out, int chunk_cols,
                               int express, char* errmsg);

/* ---------------- ty_cloud_optics_rrtmgp (LUT form) ---------------- */
typedef struct {
  int nbnd, nsize_liq, nsize_ice, nrghice, icergh;
  const int* band_lims_gpt;   /* HOST; NULL = tables are by band */
  const Float* band_lims_wvn; /* HOST (2,nbnd) */
  Float radliq_lwr, radliq_upr, diamice_lwr, diamice_upr;
  const Float *extliq, *ssaliq, *asyliq; /* HOST (nsize_liq, nbnd) */
  const Float *extice, *ssaice, *asyice; /* HOST (nsize_ice, nbnd, nrghice) */
} rrtmgpb_cloud_lut;

typedef struct rrtmgpb_cloud_optics_t rrtmgpb_cloud_optics_t;
rrtmgpb_cloud_optics_t* rrtmgpb_cloud_optics_load(const rrtmgpb_cloud_lut* lut, char* errmsg);
void rrtmgpb_cloud_optics_free(rrtmgpb_cloud_optics_t* co);
/* cloud_optics(): mo_cloud_optics_rrtmgp.F90:256-431.  clwp, ciwp, reliq, dgice (ncol,nlay) */
int rrtmgpb_cloud_optics(const rrtmgpb_cloud_optics_t* co, int ncol, int nlay, const Float* clwp, const Float* ciwp,
                         const Float* reliq, const Float* dgice, rrtmgpb_optical_props* optical_props,
                         char* errmsg);
/* cloud_optics() followed by optical_props%delta_scale() (the SW driver's two calls, rrtmgp_allsky.F90:350-352); with the
 * one-pass kernel the scaling is applied before the by-band arrays are stored (no second pass over them) */
int rrtmgpb_cloud_optics_delta_scaled(const rrtmgpb_cloud_optics_t* co, int ncol, int nlay, const Float* clwp,
                                      const Float* ciwp, const Float* reliq, const Float* dgice,
                                      rrtmgpb_optical_props* optical_props, int delta_scale, char* errmsg);
/* 1 (default): cloud_optics() computes masks, table lookups and the liquid+ice combination in one kernel
 * (rrtmgpb_cloud_optics_from_tables); 0: the reference's kernel-by-kernel sequence with its six intermediates. */
void rrtmgpb_cloud_optics_one_pass(int on);

/* compute_optimal_angles (mo_gas_optics_rrtmgp.F90:1503-1562): secant of the LW transport angle per (column, g-point) from
 * the column transmissivity, the lw_Ds argument of rrtmgpb_rte_lw.  optimal_angle_fit(2,nbnd) comes with the LW
 * k-distribution (load_int :1040-1045; HOST pointer); optimal_angles(ncol_out, ngpt_out) in backend memory. */
int rrtmgpb_gas_optics_set_optimal_angle_fit(rrtmgpb_gas_optics_t* go, const Float* optimal_angle_fit, char* errmsg);
int rrtmgpb_gas_optics_compute_optimal_angles(const rrtmgpb_gas_optics_t* go, const rrtmgpb_optical_props* optical_props,
                                              int ncol_out, int ngpt_out, Float* optimal_angles, char* errmsg);

/* ---------------- McICA cloud sampling (rte/extensions/mo_cloud_sampling.F90) ----------------
 * sampled_mask_max_ran (:125-192), sampled_mask_exp_ran (:205-292), draw_samples (:36-120) with their extent and range
 * checks; every array's extents are passed (C has no size()).  randoms(ngpt,nlay,ncol), cloud_frac(cf_ncol,cf_nlay),
 * overlap_param(op_ncol,op_nlay), cloud_mask(m_ncol,m_nlay,m_ngpt); arrays in backend memory. */
int rrtmgpb_cloud_sampling_mask_max_ran(int ngpt, int nlay, int ncol, const Float* randoms, int cf_ncol, int cf_nlay,
                                        const Float* cloud_frac, int m_ncol, int m_nlay, int m_ngpt, Bool* cloud_mask,
                                        char* errmsg);
int rrtmgpb_cloud_sampling_mask_exp_ran(int ngpt, int nlay, int ncol, const Float* randoms, int cf_ncol, int cf_nlay,
                                        const Float* cloud_frac, int op_ncol, int op_nlay, const Float* overlap_param,
                                        int m_ncol, int m_nlay, int m_ngpt, Bool* cloud_mask, char* errmsg);
int rrtmgpb_cloud_sampling_draw_samples(int m_ncol, int m_nlay, int m_ngpt, const Bool* cloud_mask,
                                        const rrtmgpb_optical_props* clouds, rrtmgpb_optical_props* clouds_sampled,
                                        char* errmsg);

/* ---------------- ty_gas_concs (rte/frontend/gas-optics-template/mo_gas_concentrations.F90) ----------------
 * Concentrations by gas name, stored as a scalar, a profile (nlay) or a field (ncol,nlay) and broadcast on demand;
 * set_vmr copies its argument as the reference does.  Array arguments live in BACKEND memory (device pointers for the
 * CUDA library - like the reference's OpenACC build, which expects device-resident arrays), scalars are passed by value. */
typedef struct rrtmgpb_gas_concs_t rrtmgpb_gas_concs_t;
rrtmgpb_gas_concs_t* rrtmgpb_gc_init(int ngas, const char* const* gas_names, char* errmsg);            /* init :96-124 */
void rrtmgpb_gc_free(rrtmgpb_gas_concs_t* gc);
int rrtmgpb_gc_set_vmr_scalar(rrtmgpb_gas_concs_t* gc, const char* gas, Float w, char* errmsg);          /* :129-191 */
int rrtmgpb_gc_set_vmr_1d(rrtmgpb_gas_concs_t* gc, const char* gas, int nlay, const Float* w, char* errmsg); /* :194-246 */
int rrtmgpb_gc_set_vmr_2d(rrtmgpb_gas_concs_t* gc, const char* gas, int ncol, int nlay, const Float* w,
                          char* errmsg);                                                              /* :249-305 */
/* get_vmr_2d :433-504: array(ncol,nlay) <- the stored concentration, broadcast (rrtmgpb_gas_concs_get_vmr) */
int rrtmgpb_gc_get_vmr(const rrtmgpb_gas_concs_t* gc, const char* gas, int ncol, int nlay, Float* array, char* errmsg);

/* ---------------- ty_aerosol_optics_rrtmgp_merra ---------------- */
/* Tables as load_lut() receives them (mo_aerosol_optics_rrtmgp_merra.F90:99-123): the rh-dependent ones arrive
 * as (nval,nrh,...) and are transposed to (nrh,nval,...) by load (:178-181).  All pointers HOST. */
typedef struct {
  int nbnd, nval, nrh, nbin;
  const Float* band_lims_wvn;       /* (2,nbnd) */
  const Float* merra_aero_bin_lims; /* (2,nbin) */
  const Float* aero_rh;             /* (nrh) */
  const Float* aero_dust_tbl;       /* (nval,nbin,nbnd) */
  const Float* aero_salt_tbl;       /* (nval,nrh,nbin,nbnd) */
  const Float* aero_sulf_tbl;       /* (nval,nrh,nbnd) */
  const Float* aero_bcar_tbl;       /* (nval,nbnd) */
  const Float* aero_bcar_rh_tbl;    /* (nval,nrh,nbnd) */
  const Float* aero_ocar_tbl;       /* (nval,nbnd) */
  const Float* aero_ocar_rh_tbl;    /* (nval,nrh,nbnd) */
} rrtmgpb_aerosol_lut;

typedef struct rrtmgpb_aerosol_optics_t rrtmgpb_aerosol_optics_t;
rrtmgpb_aerosol_optics_t* rrtmgpb_aerosol_optics_load(const rrtmgpb_aerosol_lut* lut, char* errmsg);
void rrtmgpb_aerosol_optics_free(rrtmgpb_aerosol_optics_t* ao);
/* aerosol_optics(): mo_aerosol_optics_rrtmgp_merra.F90:233-424.  aero_type (ncol,nlay) int;
 * aero_size, aero_mass, relhum (ncol,nlay) */
int rrtmgpb_aerosol_optics(const rrtmgpb_aerosol_optics_t* ao, int ncol, int nlay, const int* aero_type,
                           const Float* aero_size, const Float* aero_mass, const Float* relhum,
                           rrtmgpb_optical_props* optical_props, char* errmsg);

#ifdef __cplusplus
}
#endif
#endif /* RRTMGP_B200_FRONTEND_H */
