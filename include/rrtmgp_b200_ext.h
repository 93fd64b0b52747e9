/* rrtmgp_b200_ext.h - extension entry points beside the reference's extern-mode kernel ABI.
 *
 * The reference keeps several O(ncol*nlay*ngpt) loops of the hot path in its Fortran FRONTEND,
 * not behind the bind(C) kernel boundary (SURVEY.md section 8a').  A device-resident path needs
 * them as kernels too.  They are declared here, each citing the frontend lines it replaces.
 * Unlike the Fortran-facing symbols in rte_kernels.h / rrtmgp_kernels.h (all arguments by
 * reference), these are plain C calls: sizes and scalars BY VALUE, arrays as pointers
 * (host or device, same classification rules), Fortran array order, 1-based index VALUES.
 *
 * Also here: the memory/stream plumbing the C++ frontend (rrtmgp_b200_frontend.h) is written
 * against.  The frontend never dereferences array data itself, so the same frontend source links
 * against this CUDA library (device memory) or against the CPU oracle (host memory) - the
 * analogue of the reference's RTE_KERNEL_MODE switch.
 */
#ifndef RRTMGP_B200_EXT_H
#define RRTMGP_B200_EXT_H

#include <stddef.h>
#include "rte_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------- backend plumbing ---------------- */
/* "cuda-sm_100a" for the product library, "cpu-oracle" for the oracle. */
const char* rrtmgpb_backend_name(void);
/* sizeof(Float) of this build: 8 (default) or 4 (-DRTE_USE_SP, the reference's RTE_ENABLE_SP; rte/kernels/mo_rte_kind.F90:28-36).
 * The double-precision product is lib/librte_rrtmgp_b200.so, the single-precision one lib/librte_rrtmgp_b200_sp.so. */
int rrtmgpb_float_bytes(void);
/* Working-memory allocation in the backend's memory space (stream-ordered pool on CUDA). */
void* rrtmgpb_mem_alloc(size_t bytes);
void rrtmgpb_mem_free(void* p);
/* Copies between HOST memory and backend memory (plain memcpy on the oracle). */
void rrtmgpb_mem_to_backend(void* dst_backend, const void* src_host, size_t bytes);
void rrtmgpb_mem_to_host(void* dst_host, const void* src_backend, size_t bytes);
void rrtmgpb_mem_copy(void* dst_backend, const void* src_backend, size_t bytes);
/* CUDA: the launch stream of the CALLING HOST THREAD (every thread has its own; default cudaStreamPerThread, which
 * orders itself against the legacy default stream).  Entry points may be called concurrently from several host
 * threads on disjoint arrays; each thread's work is ordered on its own stream. */
void rrtmgpb_set_stream(void* cuda_stream);
void* rrtmgpb_get_stream(void);
void rrtmgpb_set_device(int device);
/* Blocks until all work queued by this library has finished (no-op on the oracle). */
void rrtmgpb_sync(void);
/* Number of kernel launches issued by this library since the last reset (bench.py gpu_launches). */
long long rrtmgpb_launch_count(int reset);

/* Per-kernel timing with CUDA events on the launch stream.  report: "name count total_ms" lines. */
void rrtmgpb_profile_enable(int on);
int rrtmgpb_profile_report(char* buf, size_t buflen);

/* 1: the kernel-by-kernel entry points (rrtmgp_compute_tau_absorption) may keep g-point-fastest copies of the
 * k-distribution tables, keyed by the kmajor pointer.  Only for callers that never change a table in place while its
 * allocation lives and release it through rrtmgpb_mem_free (the C++ frontend mirror switches it on around its own
 * calls); per host thread, default 0. */
void rrtmgpb_abi_table_cache(int on);
/* Lifetime contract of the g-point-fastest table copies (fused gas optics, and the kernel-by-kernel entry points while
 * rrtmgpb_abi_table_cache(1) is in effect): they are derived from the k-distribution arrays on first use, keyed by the
 * kmajor pointer and reused while every table pointer and dimension is unchanged.  A host that REFILLS a table in
 * place, or frees tables with anything other than rrtmgpb_mem_free() and allocates a new k-distribution of the same
 * shape, must call this before the next gas-optics call: kmajor = that k-distribution's kmajor array, or NULL to drop
 * every cached copy. */
void rrtmgpb_tables_changed(const void* kmajor);

/* ---------------- physical constants ---------------- */
/* replaces mo_gas_optics_constants.F90:42-51 init_constants(); NULL keeps the current value */
void rrtmgpb_init_constants(const Float* gravity, const Float* mol_weight_dry_air,
                            const Float* heat_capacity_dry_air);

/* Gas-optics tau kernels: blocks whose cells do not share table rows (unrelated neighbouring columns) may take a second
 * thread mapping - 4 lanes along the 16 g-points of one cell instead of one cell pair per thread (csrc/kernels/
 * gas_optics_gfast.cuh: tau_band_rows).  1 = on, 0 = off, -1 (default) = the environment's RRTMGPB_TAU_ROWS (0 / 1; unset: AUTOMATIC -
 * the fused entry and rrtmgp_compute_tau_absorption (with cached table copies) sample how often neighbouring cells fall into
 * different temperature / pressure bins and the calling thread's next call picks the kernel instantiation from that); 2..33 = on with that
 * vote threshold (a warp re-maps when fewer than this many lanes share rows between their two cells; 1 means 32).  Same results either way. */
void rrtmgpb_set_gas_optics_rows_path(int on);

/* ---------------- solver options ---------------- */
/* lw_solver_2stream level-source selection.  1 (default of the CUDA library): per-g-point level source as in the
 * reference's accelerator kernels (accel/mo_rte_solver_kernels.F90:958-962); 0 (default of the oracle): the serial
 * DEFAULT kernel's behaviour, every g-point uses g-point 1's level source (mo_rte_solver_kernels.F90:422 passes the
 * rank-3 array to a rank-2 dummy) - kept for bit-parity tests against the serial kernels only. */
void rrtmgpb_set_lw_2stream_lev_source_per_gpt(int on);

/* Solver kernel family: 0 (default) = register-resident warp-systolic kernels when nlay <= 80 (shared-memory
 * tile kernels otherwise); 1 = always the shared-memory tile kernels; 2 = the register kernels, never the
 * warp-specialised SW kernel; 3 = the warp-specialised SW two-stream kernel (csrc/kernels/solver_ws.cuh) wherever it
 * applies (nlay <= 80, even ncol), register kernels elsewhere.  2 and 3 exist for A/B tests: both give bit-identical
 * results. */
void rrtmgpb_set_solver_variant(int variant);
int rrtmgpb_get_solver_variant(void);

/* Test hook for the lean fp64 exp / sqrt / reciprocal / division the solver kernels use (csrc/kernels/fastmath.cuh):
 * e = exp(x), s = sqrt(|x|), r = 1/x, d = (x*x+1)/x, element-wise on n host or device doubles. */
void rrtmgpb_fastmath_probe(int n, const double* x, double* e, double* s, double* r, double* d);

/* ---------------- frontend-resident loops as kernels (SURVEY 8a') ---------------- */
/* replaces the Fortran function get_layer_number / get_col_dry, rte/kernels/mo_gas_optics_utils.F90:127-152
 * (not bind(C) in the reference: api/mo_gas_optics_utils.F90:53-66).  vmr_h2o(ncol,nlay),
 * plev(ncol,nlay+1) -> col_dry(ncol,nlay) */
void rrtmgpb_get_col_dry(int ncol, int nlay, const Float* vmr_h2o, const Float* plev, Float* col_dry);
/* replaces get_layer_mass, rte/kernels/mo_gas_optics_utils.F90:97-125 */
void rrtmgpb_get_layer_mass(int ncol, int nlay, int ngas, const Float* vmr, const Float* plev,
                            const Float* mol_weights, Float m_dry, Float* layer_mass);
/* replaces mo_gas_optics_rrtmgp.F90:594-609: col_gas(:,:,0) = col_dry; col_gas(:,:,i) = vmr(:,:,i)*col_dry.
 * vmr(ncol,nlay,ngas) -> col_gas(ncol,nlay,0:ngas) */
void rrtmgpb_col_gas_from_vmr(int ncol, int nlay, int ngas, const Float* vmr, const Float* col_dry,
                              Float* col_gas);
/* replaces combine_abs_and_rayleigh, mo_gas_optics_rrtmgp.F90:1954-2002.  kind: 1 = 1scl (tau only;
 * ssa, g ignored), 2 = 2str (tau, ssa, g = 0).  Output arrays may alias tau_abs (in-place is safe). */
void rrtmgpb_combine_abs_and_rayleigh(int ncol, int nlay, int ngpt, int kind, const Float* tau_abs,
                                      const Float* tau_rayleigh, Float* tau, Float* ssa, Float* g);
/* replaces the tlev interpolation in source(), mo_gas_optics_rrtmgp.F90:893-911 */
void rrtmgpb_interpolate_tlev(int ncol, int nlay, const Float* play, const Float* plev, const Float* tlay,
                              Float* tlev);
/* replaces mo_gas_optics_rrtmgp.F90:405-411: toa_src(icol,igpt) = solar_source(igpt) */
void rrtmgpb_broadcast_by_gpt(int ncol, int ngpt, const Float* per_gpt, Float* out);
/* replaces expand_and_transpose, rte/frontend/mo_rte_lw.F90:478-501 and mo_rte_sw.F90:399-422:
 * arr_in(nband,ncol) -> arr_out(ncol,ngpt) */
void rrtmgpb_expand_and_transpose(int ncol, int nband, int ngpt, const int* band_lims_gpt,
                                  const Float* arr_in, Float* arr_out);
/* replaces mo_rte_sw.F90:87-93: mu0_bylay(icol,ilay) = mu0(icol) */
void rrtmgpb_broadcast_by_lay(int ncol, int nlay, const Float* per_col, Float* out);
/* replaces ty_gas_concs%get_vmr_2d, rte/frontend/gas-optics-template/mo_gas_concentrations.F90:433-504 (called once per
 * gas by gas_optics, rrtmgp/frontend/mo_gas_optics_rrtmgp.F90:540-545): a concentration stored as (ncol,nlay), (1,nlay)
 * or (1,1) - conc(nc_conc, nl_conc) - is broadcast into array(ncol,nlay) */
void rrtmgpb_gas_concs_get_vmr(int ncol, int nlay, int nc_conc, int nl_conc, const Float* conc, Float* array);
/* replaces the loop of ty_gas_optics_rrtmgp%compute_optimal_angles, rrtmgp/frontend/mo_gas_optics_rrtmgp.F90:1544-1562
 * (SURVEY 8f rank 3): optimal_angles(icol,igpt) = fit(1,band)*exp(-sum_lay tau(icol,:,igpt)) + fit(2,band);
 * tau(ncol,nlay,ngpt), optimal_angle_fit(2,nband) -> optimal_angles(ncol,ngpt), the lw_Ds argument of rte_lw */
void rrtmgpb_compute_optimal_angles(int ncol, int nlay, int ngpt, int nband, const int* band_lims_gpt, const Float* tau,
                                    const Float* optimal_angle_fit, Float* optimal_angles);
/* McICA cloud sampling, rte/extensions/mo_cloud_sampling.F90.  randoms(ngpt,nlay,ncol) in [0,1),
 * cloud_frac(ncol,nlay), overlap_param(ncol,nlay-1) -> cloud_mask(ncol,nlay,ngpt).
 * replaces the column loops of sampled_mask_max_ran (:160-190) and sampled_mask_exp_ran (:250-290) */
void rrtmgpb_sampled_mask_max_ran(int ncol, int nlay, int ngpt, const Float* randoms, const Float* cloud_frac,
                                  Bool* cloud_mask);
void rrtmgpb_sampled_mask_exp_ran(int ncol, int nlay, int ngpt, const Float* randoms, const Float* cloud_frac,
                                  const Float* overlap_param, Bool* cloud_mask);
/* replaces apply_cloud_mask (:298-314), the body of draw_samples: input_field(ncol,nlay,nbnd) by band ->
 * sampled_field(ncol,nlay,ngpt), zero where the mask is false */
void rrtmgpb_apply_cloud_mask(int ncol, int nlay, int nbnd, int ngpt, const int* band_lims_gpt, const Bool* cloud_mask,
                              const Float* input_field, Float* sampled_field);
/* replaces the cloud masks mo_cloud_optics_rrtmgp.F90:334-341 */
void rrtmgpb_cloud_masks(int ncol, int nlay, const Float* clwp, const Float* ciwp, Bool* liqmsk,
                         Bool* icemsk);
/* replaces the liquid+ice combination mo_cloud_optics_rrtmgp.F90:399-424.  kind 1: tau = (ltau-ltaussa)
 * + (itau-itaussa); kind 2: tau, ssa, g with epsilon() guards. */
void rrtmgpb_cloud_combine(int ncol, int nlay, int ngpt, int kind, const Float* ltau, const Float* ltaussa,
                           const Float* ltaussag, const Float* itau, const Float* itaussa,
                           const Float* itaussag, Float* tau, Float* ssa, Float* g);
/* The compute part of ty_cloud_optics_rrtmgp%cloud_optics in one pass: the masks (mo_cloud_optics_rrtmgp.F90:334-341),
 * both compute_cld_from_table calls (:373, :380; ice tables = the icergh slice) and the liquid+ice combination
 * (:399-424), without materialising the six (ncol,nlay,nbnd) intermediates.  kind 1: tau ; kind 2: tau, ssa, g. */
void rrtmgpb_cloud_optics_from_tables(int ncol, int nlay, int nbnd, int kind, const Float* clwp, const Float* ciwp,
                                      const Float* reliq, const Float* dgice, int liq_nsteps, Float liq_step_size,
                                      Float liq_offset, const Float* extliq, const Float* ssaliq, const Float* asyliq,
                                      int ice_nsteps, Float ice_step_size, Float ice_offset, const Float* extice,
                                      const Float* ssaice, const Float* asyice, Float* tau, Float* ssa, Float* g);
/* the same, optionally followed in the same pass by clouds%delta_scale() (rte/frontend/mo_optical_props.F90:565-612 ->
 * delta_scale_2str_k, mo_optical_props_kernels.F90:89-93; 2-stream only), the SW driver's next call
 * (examples/all-sky/rrtmgp_allsky.F90:350-352): saves one read-modify-write pass over the by-band cloud arrays */
void rrtmgpb_cloud_optics_from_tables_ds(int ncol, int nlay, int nbnd, int kind, const Float* clwp, const Float* ciwp,
                                         const Float* reliq, const Float* dgice, int liq_nsteps, Float liq_step_size,
                                         Float liq_offset, const Float* extliq, const Float* ssaliq, const Float* asyliq,
                                         int ice_nsteps, Float ice_step_size, Float ice_offset, const Float* extice,
                                         const Float* ssaice, const Float* asyice, Float* tau, Float* ssa, Float* g,
                                         int delta_scale);
/* replaces compute_all_from_table + the optical-property combination of ty_aerosol_optics_rrtmgp_merra%aerosol_optics,
 * rrtmgp/frontend/mo_aerosol_optics_rrtmgp_merra.F90:436-559 and :385-418 (size-bin search, relative-humidity
 * bracket + linear interpolation, per-type table lookup; kind 1: tau = atau - ataussa; kind 2: tau, ssa, g with
 * epsilon() guards).  type(ncol,nlay) int: 0 none, 1 dust, 2 salt, 3 sulfate, 4 bcar_rh, 5 bcar, 6 ocar_rh, 7 ocar
 * (:50-59); size, mass, rh (ncol,nlay); bin_lims(npair=2,nbin); aero_rh(nrh); tables as the loader holds them:
 * dust(nval,nbin,nbnd) salt(nrh,nval,nbin,nbnd) sulf/bcar_rh/ocar_rh(nrh,nval,nbnd) bcar/ocar(nval,nbnd), nval = 3
 * (ext, ssa, g).  Outputs (ncol,nlay,nbnd). */
void rrtmgpb_aerosol_optics_from_table(int ncol, int nlay, int nval, int nrh, int nbin, int nbnd, int kind,
                                       const int* type, const Float* size, const Float* mass, const Float* rh,
                                       const Float* bin_lims, const Float* aero_rh, const Float* dust_tbl,
                                       const Float* salt_tbl, const Float* sulf_tbl, const Float* bcar_rh_tbl,
                                       const Float* bcar_tbl, const Float* ocar_rh_tbl, const Float* ocar_tbl,
                                       Float* tau, Float* ssa, Float* g);
/* aerosol mask (:343-347) and any_int_vals_outside_2D (:580-600): return 1 if any element violates */
void rrtmgpb_aerosol_mask(int ncol, int nlay, const int* type, Bool* aeromsk);
int rrtmgpb_any_int_vals_outside(size_t n, const int* array, int checkMin, int checkMax);
/* value checks, rte/frontend/mo_rte_util_array_validation.F90:52-...: return 1 if any element violates.
 * mask may be NULL (no mask). */
int rrtmgpb_any_vals_less_than(size_t n, const Float* array, const Bool* mask, Float check_value);
int rrtmgpb_any_vals_outside(size_t n, const Float* array, const Bool* mask, Float checkMin, Float checkMax);

/* ---------------- flux diagnostics beside the broadband reductions ---------------- */
/* The reference compiles these in place (rte/extensions/mo_fluxes_byband.F90:159,184,210: bind(C) names rte_sum_byband,
 * rte_net_byband_full, net_byband_precalc); same by-reference convention as the 45 extern-mode symbols.
 * band_lims(2,nbnd): 1-based g-point limits; spectral_flux(ncol,nlev,ngpt) -> byband_flux(ncol,nlev,nbnd). */
void rte_sum_byband(const int* ncol, const int* nlev, const int* ngpt, const int* nbnd, const int* band_lims,
                    const Float* spectral_flux, Float* byband_flux);
void rte_net_byband_full(const int* ncol, const int* nlev, const int* ngpt, const int* nbnd, const int* band_lims,
                         const Float* spectral_flux_dn, const Float* spectral_flux_up, Float* byband_flux_net);
void net_byband_precalc(const int* ncol, const int* nlev, const int* nbnd, const Float* byband_flux_dn,
                        const Float* byband_flux_up, Float* byband_flux_net);
/* replaces compute_heating_rate_general, rte/extensions/mo_heating_rates.F90:34-64: heating rate [K/s] of every layer
 * from the flux divergence, H = (dF_up - dF_dn) * grav / (cp_dry * dp); fluxes, p_lev (ncol,nlay+1) -> (ncol,nlay) */
void rrtmgpb_heating_rate(int ncol, int nlay, const Float* flux_up, const Float* flux_dn, const Float* p_lev,
                          Float* heating_rate);
/* replaces compute_heating_rate_solar_varmu0 (:66-117): as above, then the layer in which mu0(ncol,nlay) goes from
 * positive to zero is recomputed from the diffuse (total minus direct) net flux */
void rrtmgpb_heating_rate_solar_varmu0(int ncol, int nlay, const Float* flux_up, const Float* flux_dn,
                                       const Float* flux_dir, const Float* p_lev, const Float* mu0, Float* heating_rate);

/* ---------------- simple spectral model gas optics ---------------- */
/* The kernels of the reference's second, data-file-free gas-optics provider (ssm/mo_optics_ssm_kernels.F90:29,83;
 * bind(C) names as there; by-reference convention).  absorption_coeffs(ngas,nnu), play(ncol,nlay),
 * layer_mass(ngas,ncol,nlay) -> tau(ncol,nlay,nnu), pressure broadening play/pref when pref > 0;
 * vmr(ngas,ncol,nlay), plev(ncol,nlay+1), mol_weights(ngas) -> layer_mass(ngas,ncol,nlay).  Its Planck sources use
 * rte_compute_Planck_source_1D/_2D of rte_kernels.h. */
void ssm_compute_tau_absorption(const int* ncol, const int* nlay, const int* nnu, const int* ngas,
                                const Float* absorption_coeffs, const Float* play, const Float* pref,
                                const Float* layer_mass, Float* tau);
void ssm_compute_layer_mass(const int* ncol, const int* nlay, const int* ngas, const Float* vmr, const Float* plev,
                            const Float* mol_weights, const Float* m_dry, Float* layer_mass);

/* ---------------- fused variants used by the device-resident frontend ---------------- */
/* As rrtmgp_compute_tau_absorption but ASSIGNS tau instead of accumulating into a pre-zeroed array:
 * saves the zero_array_3D plane write and the plane read (mo_gas_optics_rrtmgp.F90:637-665,679-706). */
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
    const Float* tlay, const Float* col_gas, const int* jeta, const int* jtemp, const int* jpress, Float* tau);

/* Whole gas_optics() call fused: col_dry, col_gas, interpolation, absorption (major +
 * minor), Rayleigh + combination (when krayl != NULL), the by-band cloud increment (cld_kind 0 none / 1 1scl /
 * 2 2str; cloud arrays (ncol,nlay,nbnd) on the SAME bands) and - when lay_src != NULL - the Planck sources.
 * Only the caller-visible arrays are written: tau[,ssa,g] (op_kind 1 / 2) and the four source arrays; no
 * intermediate of the reference sequence is materialised.  All table pointers in `t` are BACKEND memory. */
typedef struct {
  int ngas, nflav, neta, npres, ntemp, nbnd, ngpt;
  int nminorlower, nminorklower, nminorupper, nminorkupper, idx_h2o;
  const int *flavor, *gpoint_flavor, *band_lims_gpt, *gpoint_bands;
  const Float *press_ref_log, *temp_ref, *vmr_ref;
  Float press_ref_log_delta, temp_ref_min, temp_ref_delta, press_ref_trop_log;
  const Float *kmajor, *kminor_lower, *kminor_upper;
  const int *minor_limits_gpt_lower, *minor_limits_gpt_upper;
  const Bool *minor_scales_with_density_lower, *minor_scales_with_density_upper;
  const Bool *scale_by_complement_lower, *scale_by_complement_upper;
  const int *idx_minor_lower, *idx_minor_upper, *idx_minor_scaling_lower, *idx_minor_scaling_upper;
  const int *kminor_start_lower, *kminor_start_upper;
  const Float* krayl;                  /* NULL: no Rayleigh scattering (LW) */
  const Float *planck_frac, *totplnk;  /* NULL for SW */
  int nPlanckTemp;
  Float totplnk_delta;
} rrtmgpb_gas_tables;

/* 1: stage the k-distribution / Planck-fraction table boxes to shared memory with TMA (cp.async.bulk.tensor)
 * in the fused gas-optics kernels; 0 (default): read them through the read-only global path.
 * RRTMGPB_TMA=1 in the environment turns staging on at start-up. */
void rrtmgpb_set_tma_staging(int on);

/* aer_*: a second by-band increment applied after the cloud one (the driver's aerosols%increment(atmos),
 * examples/all-sky/rrtmgp_allsky.F90:377,398), same kinds. */
void rrtmgpb_gas_optics_fused(const rrtmgpb_gas_tables* t, int ncol, int nlay, const Float* play, const Float* plev,
                              const Float* tlay, const Float* vmr, const Float* col_dry /* or NULL */, int op_kind,
                              Float* tau, Float* ssa, Float* g, int cld_kind, const Float* cld_tau,
                              const Float* cld_ssa, const Float* cld_g, int aer_kind, const Float* aer_tau,
                              const Float* aer_ssa, const Float* aer_g, const Float* tlev, const Float* tsfc,
                              int sfc_lay, Float* sfc_src, Float* lay_src /* NULL: no sources */, Float* lev_src,
                              Float* sfc_source_Jac);

/* ---------------- express path ---------------- */
/* Broadband fluxes straight from the atmospheric state: gas optics (+ by-band cloud increment, + Planck sources) feed
 * the flux solver per band through a scratch that is sized to stay in L2; no (ncol, nlay, ngpt) array is allocated, the
 * footprint is ~6 KB per column instead of ~0.9 MB.  LW when t->krayl == NULL (no-scattering solver, nmus Gauss
 * angles: Ds_host / wts_host are HOST arrays of nmus secants and weights, mo_rte_lw.F90:146-160), else SW (two-stream).
 * Arrays: play, tlay (ncol,nlay); plev, tlev (ncol,nlay+1); tsfc, mu0 (ncol); vmr (ncol,nlay,ngas); col_dry optional;
 * cld_* by band (ncol,nlay,nbnd), kind 0 none / 1 tau / 2 tau,ssa[,g]; sfc_emis_or_alb_dir, sfc_alb_dif (nbnd,ncol);
 * solar_source (ngpt); outputs (ncol,nlay+1).  The oracle's implementation is the reference call sequence on full
 * arrays (gas optics, increment, rte_lw / rte_sw). */
void rrtmgpb_express(const rrtmgpb_gas_tables* t, int ncol, int nlay, int top_at_1, const Float* play, const Float* plev,
                     const Float* tlay, const Float* tlev, const Float* tsfc, const Float* vmr, const Float* col_dry,
                     int cld_kind, const Float* cld_tau, const Float* cld_ssa, const Float* cld_g,
                     const Float* sfc_emis_or_alb_dir, const Float* sfc_alb_dif, const Float* mu0,
                     const Float* solar_source, int nmus, const Float* Ds_host, const Float* wts_host, Float* flux_up,
                     Float* flux_dn, Float* flux_dir);
/* 1: the band-staged pipeline runs (register solvers: nlay within their range); 0: rrtmgpb_express still works, one
 * solver launch per column chunk over all bands (bounded scratch, planes through HBM) */
int rrtmgpb_express_supported(int ncol, int nlay);

#ifdef __cplusplus
}
#endif
#endif /* RRTMGP_B200_EXT_H */
