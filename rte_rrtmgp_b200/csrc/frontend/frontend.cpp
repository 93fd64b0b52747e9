// frontend.cpp - C++ mirror of the reference's Fortran frontend for the all-sky / clear-sky hot path.
//
// Mirrors (same call sequences, argument meaning, error strings):
//   rte/frontend/mo_rte_config.F90:29-49            rte_config_checks
//   rte/frontend/mo_optical_props.F90:562-700,879-1028  delta_scale, validate, increment
//   rte/frontend/mo_rte_lw.F90:79-501               rte_lw (+ expand_and_transpose, Gauss-Jacobi-5 table)
//   rte/frontend/mo_rte_sw.F90:56-422               rte_sw (mu0 by column)
//   rrtmgp/frontend/mo_gas_optics_rrtmgp.F90:220-414,419-745,840-928  gas_optics_int/ext, compute_gas_taus, source
//   rrtmgp/frontend/mo_cloud_optics_rrtmgp.F90:256-431            cloud_optics (LUT)
//   rrtmgp/frontend/mo_aerosol_optics_rrtmgp_merra.F90:99-424     aerosol load_lut, aerosol_optics
//
// Backend-agnostic by construction: this file only sequences extern "C" kernels (rte_kernels.h,
// rrtmgp_kernels.h, rrtmgp_b200_ext.h) and never dereferences an array, so the SAME source is linked
// (a) into librte_rrtmgp_b200.so against the CUDA kernels - the product - and (b) into the oracle's
// CPU shared library, against the C restatement - test infrastructure / CPU baseline.  That is the
// reference's own RTE_KERNEL_MODE idea (one frontend, interchangeable kernel providers).
//
// Differences from the Fortran frontend, all deliberate:
//   * gas concentrations arrive as one (ncol,nlay,ngas) vmr array in k-distribution gas order instead of
//     a ty_gas_concs object;
//   * when the caller wants broadband fluxes the absorption kernel ASSIGNS tau (extension symbol) instead
//     of zero_array + accumulate, saving two plane passes; results are identical.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "rrtmgp_b200_ext.h"
#include "rrtmgp_b200_frontend.h"
#include "rrtmgp_kernels.h"
#include "rte_kernels.h"

namespace {

// process-wide like the module variables of mo_rte_config.F90:25-26; atomics, so host threads may read them concurrently
std::atomic<bool> g_check_extents{true};
std::atomic<bool> g_check_values{true};

int fail(char* errmsg, const std::string& msg) {
  if (errmsg) {
    std::memset(errmsg, 0, RRTMGPB_ERRLEN);
    std::strncpy(errmsg, msg.c_str(), RRTMGPB_ERRLEN - 1);
  }
  return msg.empty() ? 0 : 1;
}
int ok(char* errmsg) { return fail(errmsg, ""); }

// scratch in backend memory, freed at scope exit (stream-ordered on CUDA)
template <typename T>
struct Scratch {
  T* p;
  explicit Scratch(size_t n) : p(static_cast<T*>(rrtmgpb_mem_alloc(n * sizeof(T)))) {}
  ~Scratch() { rrtmgpb_mem_free(p); }
  Scratch(const Scratch&) = delete;
  Scratch& operator=(const Scratch&) = delete;
  operator T*() const { return p; }
};

template <typename T>
T* upload(const T* host, size_t n) {
  if (!host || n == 0) return nullptr;
  T* d = static_cast<T*>(rrtmgpb_mem_alloc(n * sizeof(T)));
  rrtmgpb_mem_to_backend(d, host, n * sizeof(T));
  return d;
}

bool bands_are_equal(const rrtmgpb_optical_props* a, const rrtmgpb_optical_props* b) {
  // mo_optical_props.F90 bands_are_equal: same number of bands and wavenumber limits within 5 spacings
  if (a->nband != b->nband) return false;
  if (!a->band_lims_wvn || !b->band_lims_wvn) return true;
  for (int i = 0; i < 2 * a->nband; ++i) {
    const Float x = a->band_lims_wvn[i], y = b->band_lims_wvn[i];
    const Float d = x > y ? x - y : y - x;
    if (d > (Float)5 * (Float)2.220446049250313e-16 * (x > 0 ? x : -x)) return false;
  }
  return true;
}
bool gpoints_are_equal(const rrtmgpb_optical_props* a, const rrtmgpb_optical_props* b) {
  if (!bands_are_equal(a, b) || a->ngpt != b->ngpt) return false;
  for (int i = 0; i < 2 * a->nband; ++i)
    if (a->band_lims_gpt[i] != b->band_lims_gpt[i]) return false;
  return true;
}

// Gauss-Jacobi-5 quadrature, mo_rte_lw.F90:136-160 (values are mu = cos(theta); D = 1/mu)
const int max_gauss_pts = 4;
const double gauss_mus[4][4] = {{0.6096748751, 0, 0, 0},
                                {0.2509907356, 0.7908473988, 0, 0},
                                {0.1024922169, 0.4417960320, 0.8633751621, 0},
                                {0.0454586727, 0.2322334416, 0.5740198775, 0.9030775973}};
const double gauss_wts[4][4] = {{1.0, 0, 0, 0},
                                {0.2300253764, 0.7699746236, 0, 0},
                                {0.0437820218, 0.3875796738, 0.5686383044, 0},
                                {0.0092068785, 0.1285704278, 0.4323381850, 0.4298845087}};

}  // namespace

extern "C" {

void rrtmgpb_rte_config_checks(int check_extents, int check_values) {
  g_check_extents = check_extents != 0;
  g_check_values = check_values != 0;
}

// ------------------------------------------------------------------------------------------------
// ty_optical_props_arry
// ------------------------------------------------------------------------------------------------
int rrtmgpb_op_validate(const rrtmgpb_optical_props* op, char* errmsg) {
  const size_t n = (size_t)op->ncol * op->nlay * op->ngpt;
  std::string msg;
  if (!op->tau) return fail(errmsg, "validate: tau not allocated/initialized");
  if (op->kind != RRTMGPB_1SCL && (!op->ssa || (op->kind == RRTMGPB_2STR ? !op->g : !op->p)))
    return fail(errmsg, "validate: arrays not allocated/initialized");
  if (g_check_values) {  // mo_optical_props.F90:618-621,652-659,691-698
    if (rrtmgpb_any_vals_less_than(n, op->tau, nullptr, 0)) msg = "validate: tau values out of range";
    if (op->kind != RRTMGPB_1SCL) {
      if (rrtmgpb_any_vals_outside(n, op->ssa, nullptr, 0, 1)) msg = "validate: ssa values out of range";
      if (op->kind == RRTMGPB_2STR && rrtmgpb_any_vals_outside(n, op->g, nullptr, -1, 1))
        msg = "validate: g values out of range";
    }
  }
  return fail(errmsg, msg);
}

int rrtmgpb_op_delta_scale(rrtmgpb_optical_props* op, const Float* forward, char* errmsg) {
  if (op->kind == RRTMGPB_1SCL) return ok(errmsg);  // :573-583 nothing to do
  if (op->kind == RRTMGPB_NSTR) return fail(errmsg, "delta_scale_nstr: Not yet implemented");
  const size_t n = (size_t)op->ncol * op->nlay * op->ngpt;
  if (forward) {  // :596-609
    if (g_check_values && rrtmgpb_any_vals_outside(n, forward, nullptr, 0, 1))
      return fail(errmsg, "delta_scale: values of 'for' out of bounds [0,1]");
    rte_delta_scale_2str_f_k(&op->ncol, &op->nlay, &op->ngpt, op->tau, op->ssa, op->g, forward);
  } else {
    rte_delta_scale_2str_k(&op->ncol, &op->nlay, &op->ngpt, op->tau, op->ssa, op->g);
  }
  return ok(errmsg);
}

int rrtmgpb_op_increment(const rrtmgpb_optical_props* in, rrtmgpb_optical_props* io, char* errmsg) {
  // mo_optical_props.F90:879-1028
  const int ncol = io->ncol, nlay = io->nlay, ngpt = io->ngpt;
  if (!in->tau) return fail(errmsg, "ty_optical_props%increment: Incrementing optical properties aren't initialized");
  if (!io->tau)
    return fail(errmsg, "ty_optical_props%increment: optical properties to be incremented aren't initialized");
  if (!bands_are_equal(in, io))
    return fail(errmsg, "ty_optical_props%increment: optical properties objects have different band structures");
  if (in->ncol != ncol || in->nlay != nlay)
    return fail(errmsg, "ty_optical_props%increment: optical properties objects have different ncol and/or nlay");
  if (gpoints_are_equal(in, io)) {
    switch (io->kind * 10 + in->kind) {
      case 11: rte_increment_1scalar_by_1scalar(&ncol, &nlay, &ngpt, io->tau, in->tau); break;
      case 12: rte_increment_1scalar_by_2stream(&ncol, &nlay, &ngpt, io->tau, in->tau, in->ssa); break;
      case 13: rte_increment_1scalar_by_nstream(&ncol, &nlay, &ngpt, io->tau, in->tau, in->ssa); break;
      case 21: rte_increment_2stream_by_1scalar(&ncol, &nlay, &ngpt, io->tau, io->ssa, in->tau); break;
      case 22: rte_increment_2stream_by_2stream(&ncol, &nlay, &ngpt, io->tau, io->ssa, io->g, in->tau, in->ssa, in->g); break;
      case 23: rte_increment_2stream_by_nstream(&ncol, &nlay, &ngpt, &in->nmom, io->tau, io->ssa, io->g, in->tau, in->ssa, in->p); break;
      case 31: rte_increment_nstream_by_1scalar(&ncol, &nlay, &ngpt, io->tau, io->ssa, in->tau); break;
      case 32: rte_increment_nstream_by_2stream(&ncol, &nlay, &ngpt, &io->nmom, io->tau, io->ssa, io->p, in->tau, in->ssa, in->g); break;
      case 33: rte_increment_nstream_by_nstream(&ncol, &nlay, &ngpt, &io->nmom, &in->nmom, io->tau, io->ssa, io->p, in->tau, in->ssa, in->p); break;
    }
  } else {
    if (in->ngpt != io->nband)
      return fail(errmsg, "ty_optical_props%increment: optical properties objects have incompatible g-point structures");
    const int nb = io->nband;
    int* lims = upload(io->band_lims_gpt, 2 * (size_t)nb);  // kernels read gpt_lims from backend memory
    switch (io->kind * 10 + in->kind) {
      case 11: rte_inc_1scalar_by_1scalar_bybnd(&ncol, &nlay, &ngpt, io->tau, in->tau, &nb, lims); break;
      case 12: rte_inc_1scalar_by_2stream_bybnd(&ncol, &nlay, &ngpt, io->tau, in->tau, in->ssa, &nb, lims); break;
      case 13: rte_inc_1scalar_by_nstream_bybnd(&ncol, &nlay, &ngpt, io->tau, in->tau, in->ssa, &nb, lims); break;
      case 21: rte_inc_2stream_by_1scalar_bybnd(&ncol, &nlay, &ngpt, io->tau, io->ssa, in->tau, &nb, lims); break;
      case 22: rte_inc_2stream_by_2stream_bybnd(&ncol, &nlay, &ngpt, io->tau, io->ssa, io->g, in->tau, in->ssa, in->g, &nb, lims); break;
      case 23: rte_inc_2stream_by_nstream_bybnd(&ncol, &nlay, &ngpt, &in->nmom, io->tau, io->ssa, io->g, in->tau, in->ssa, in->p, &nb, lims); break;
      case 31: rte_inc_nstream_by_1scalar_bybnd(&ncol, &nlay, &ngpt, io->tau, io->ssa, in->tau, &nb, lims); break;
      case 32: rte_inc_nstream_by_2stream_bybnd(&ncol, &nlay, &ngpt, &io->nmom, io->tau, io->ssa, io->p, in->tau, in->ssa, in->g, &nb, lims); break;
      case 33: rte_inc_nstream_by_nstream_bybnd(&ncol, &nlay, &ngpt, &io->nmom, &in->nmom, io->tau, io->ssa, io->p, in->tau, in->ssa, in->p, &nb, lims); break;
    }
    rrtmgpb_mem_free(lims);
  }
  return ok(errmsg);
}

// ------------------------------------------------------------------------------------------------
// rte_lw
// ------------------------------------------------------------------------------------------------
static int rte_lw_impl(const rrtmgpb_optical_props* op, const rrtmgpb_source_func_lw* src, const Float* sfc_emis,
                       rrtmgpb_fluxes_broadband* fluxes, Float* gpt_up_out, Float* gpt_dn_out, const Float* inc_flux,
                       int n_gauss_angles, int use_2stream, const Float* lw_Ds, Float* flux_up_Jac, char* errmsg) {
  const int ncol = op->ncol, nlay = op->nlay, ngpt = op->ngpt, nband = op->nband;
  const size_t ncg = (size_t)ncol * ngpt, nclp = (size_t)ncol * (nlay + 1);
  const bool do_broadband = fluxes != nullptr;  // ty_fluxes_broadband is special-cased, mo_rte_lw.F90:296-313
  const bool do_Jacobians = flux_up_Jac != nullptr;
  std::string msg;
  // ---- error checking :170-262
  if (do_broadband && !(fluxes->flux_up || fluxes->flux_dn || fluxes->flux_net))
    msg = "rte_lw: no space allocated for fluxes";
  if (g_check_extents) {
    if (src->ncol != ncol || src->nlay != nlay || src->ngpt != ngpt)
      msg = "rte_lw: sources and optical properties inconsistently sized";
  }
  if (g_check_values) {
    if (rrtmgpb_any_vals_outside((size_t)nband * ncol, sfc_emis, nullptr, 0, 1))
      msg = "rte_lw: sfc_emis has values < 0 or > 1";
    if (inc_flux && rrtmgpb_any_vals_less_than(ncg, inc_flux, nullptr, 0)) msg = "rte_lw: inc_flux has values < 0";
    if (lw_Ds && rrtmgpb_any_vals_less_than(ncg, lw_Ds, nullptr, 1)) msg = "rte_lw: one or more values of lw_Ds < 1.";
    if (n_gauss_angles > max_gauss_pts)
      msg = "rte_lw: asking for too many quadrature points for no-scattering calculation";
    if (n_gauss_angles < 0)
      msg = "rte_lw: have to ask for at least one quadrature point for no-scattering calculation";
  }
  if (!msg.empty()) return fail(errmsg, msg);
  const int n_quad_angs = n_gauss_angles > 0 ? n_gauss_angles : 1;
  const bool using_2stream = use_2stream > 0;
  if (op->kind == RRTMGPB_1SCL) {
    if (using_2stream) msg = "rte_lw: can't use two-stream methods with only absorption optical depth";
    if (lw_Ds && n_quad_angs != 1) msg = "rte_lw: providing lw_Ds incompatible with specifying n_gauss_angles";
  } else if (op->kind == RRTMGPB_2STR) {
    if (lw_Ds) msg = "rte_lw: lw_Ds not valid when providing scattering optical properties";
    if (using_2stream && n_quad_angs != 1) msg = "rte_lw: using_2stream=true incompatible with specifying n_gauss_angles";
    if (using_2stream && do_Jacobians)
      msg = "rte_lw: can't provide Jacobian of fluxes w.r.t surface temperature with 2-stream";
  } else {
    msg = "rte_lw: lw_solver(...ty_optical_props_nstr...) not yet implemented";
  }
  if (!msg.empty()) return fail(errmsg, msg);

  // ---- boundary conditions :264-282,329
  Scratch<Float> sfc_emis_gpt(ncg);
  Scratch<int> lims(2 * (size_t)nband);
  rrtmgpb_mem_to_backend(lims, op->band_lims_gpt, sizeof(int) * 2 * nband);
  rrtmgpb_expand_and_transpose(ncol, nband, ngpt, lims, sfc_emis, sfc_emis_gpt);
  Float* inc_alloc = nullptr;
  const Float* inc_flux_diffuse = inc_flux;
  if (!inc_flux) {
    inc_alloc = static_cast<Float*>(rrtmgpb_mem_alloc(ncg * sizeof(Float)));
    zero_array_2D(&ncol, &ngpt, inc_alloc);  // :280
    inc_flux_diffuse = inc_alloc;
  }
  // ---- output plumbing :284-322: broadband outputs need both up and down storage
  Float *up_loc = nullptr, *dn_loc = nullptr, *up_tmp = nullptr, *dn_tmp = nullptr, *decoy2 = nullptr;
  if (do_broadband) {
    up_loc = fluxes->flux_up; dn_loc = fluxes->flux_dn;
    if (!up_loc) up_loc = up_tmp = static_cast<Float*>(rrtmgpb_mem_alloc(nclp * sizeof(Float)));
    if (!dn_loc) dn_loc = dn_tmp = static_cast<Float*>(rrtmgpb_mem_alloc(nclp * sizeof(Float)));
  } else {
    decoy2 = static_cast<Float*>(rrtmgpb_mem_alloc(nclp * sizeof(Float)));
    up_loc = dn_loc = decoy2;  // :320-321
  }
  Float* jacobian = do_Jacobians ? flux_up_Jac : (decoy2 ? decoy2 : up_loc);
  // g-point flux arrays are decoys in the broadband case (never touched by the kernels)
  Float* gpt_up = do_broadband ? up_loc : gpt_up_out;
  Float* gpt_dn = do_broadband ? dn_loc : gpt_dn_out;
  const Bool top_at_1 = op->top_at_1 != 0, bb = do_broadband, jac = do_Jacobians;

  if (g_check_values) {
    char verr[RRTMGPB_ERRLEN];
    if (rrtmgpb_op_validate(op, verr)) msg = verr;
  }
  if (msg.empty()) {
    if (op->kind == RRTMGPB_1SCL || !using_2stream) {
      // secants :345-365 and the two lw_solver_noscat call sites :367-378, :411-422
      Scratch<Float> secants(ncg * n_quad_angs);
      std::vector<Float> wts(n_quad_angs);
      if (lw_Ds) {
        rrtmgpb_mem_copy(secants, lw_Ds, ncg * sizeof(Float));
        wts[0] = (Float)gauss_wts[0][0];
      } else {
        for (int imu = 0; imu < n_quad_angs; ++imu) {
          const Float D = (Float)1 / (Float)gauss_mus[n_quad_angs - 1][imu];
          set_to_scalar_2D(&ncol, &ngpt, secants + ncg * imu, &D);
          wts[imu] = (Float)gauss_wts[n_quad_angs - 1][imu];
        }
      }
      Scratch<Float> wts_b(n_quad_angs);
      rrtmgpb_mem_to_backend(wts_b, wts.data(), sizeof(Float) * n_quad_angs);
      const Bool resc = op->kind == RRTMGPB_2STR;
      // the last two arguments are not used when do_rescaling is false but need valid addresses (:378)
      const Float* ssa = resc ? op->ssa : op->tau;
      const Float* g = resc ? op->g : op->tau;
      rte_lw_solver_noscat(&ncol, &nlay, &ngpt, &top_at_1, &n_quad_angs, secants, wts_b, op->tau, src->lay_source,
                           src->lev_source, sfc_emis_gpt, src->sfc_source, inc_flux_diffuse, gpt_up, gpt_dn, &bb,
                           up_loc, dn_loc, &jac, src->sfc_source_Jac, jacobian, &resc, ssa, g);
    } else {
      // two-stream with scattering :388-394; no broadband outputs exist for this kernel:
      // with ty_fluxes_broadband the reference leaves flux_up/flux_dn unfilled; we sum the g-point fluxes.
      Float *gu = gpt_up_out, *gd = gpt_dn_out, *gu_tmp = nullptr, *gd_tmp = nullptr;
      if (do_broadband) {
        gu = gu_tmp = static_cast<Float*>(rrtmgpb_mem_alloc(nclp * ngpt * sizeof(Float)));
        gd = gd_tmp = static_cast<Float*>(rrtmgpb_mem_alloc(nclp * ngpt * sizeof(Float)));
      }
      rte_lw_solver_2stream(&ncol, &nlay, &ngpt, &top_at_1, op->tau, op->ssa, op->g, src->lay_source, src->lev_source,
                            sfc_emis_gpt, src->sfc_source, inc_flux_diffuse, gu, gd);
      if (do_broadband) {
        const int nlev = nlay + 1;
        rte_sum_broadband(&ncol, &nlev, &ngpt, gu, up_loc);
        rte_sum_broadband(&ncol, &nlev, &ngpt, gd, dn_loc);
        rrtmgpb_mem_free(gu_tmp);
        rrtmgpb_mem_free(gd_tmp);
      }
    }
    if (do_broadband && fluxes->flux_net) {  // :439-450
      const int nlev = nlay + 1;
      rte_net_broadband_precalc(&ncol, &nlev, dn_loc, up_loc, fluxes->flux_net);
    }
  }
  rrtmgpb_mem_free(inc_alloc);
  rrtmgpb_mem_free(up_tmp);
  rrtmgpb_mem_free(dn_tmp);
  rrtmgpb_mem_free(decoy2);
  return fail(errmsg, msg);
}

int rrtmgpb_rte_lw(const rrtmgpb_optical_props* op, const rrtmgpb_source_func_lw* src, const Float* sfc_emis,
                   rrtmgpb_fluxes_broadband* fluxes, const Float* inc_flux, int n_gauss_angles, int use_2stream,
                   const Float* lw_Ds, Float* flux_up_Jac, char* errmsg) {
  if (!fluxes) return fail(errmsg, "rte_lw: no space allocated for fluxes");
  return rte_lw_impl(op, src, sfc_emis, fluxes, nullptr, nullptr, inc_flux, n_gauss_angles, use_2stream, lw_Ds,
                     flux_up_Jac, errmsg);
}
int rrtmgpb_rte_lw_bygpoint(const rrtmgpb_optical_props* op, const rrtmgpb_source_func_lw* src, const Float* sfc_emis,
                            Float* gpt_flux_up, Float* gpt_flux_dn, const Float* inc_flux, int n_gauss_angles,
                            int use_2stream, const Float* lw_Ds, char* errmsg) {
  if (!gpt_flux_up || !gpt_flux_dn) return fail(errmsg, "rte_lw: no space allocated for fluxes");
  return rte_lw_impl(op, src, sfc_emis, nullptr, gpt_flux_up, gpt_flux_dn, inc_flux, n_gauss_angles, use_2stream,
                     lw_Ds, nullptr, errmsg);
}

// ------------------------------------------------------------------------------------------------
// rte_sw
// ------------------------------------------------------------------------------------------------
static int rte_sw_impl(const rrtmgpb_optical_props* atmos, const Float* mu0, const Float* inc_flux,
                       const Float* sfc_alb_dir, const Float* sfc_alb_dif, rrtmgpb_fluxes_broadband* fluxes,
                       Float* gpt_up_out, Float* gpt_dn_out, Float* gpt_dir_out, const Float* inc_flux_dif,
                       char* errmsg) {
  const int ncol = atmos->ncol, nlay = atmos->nlay, ngpt = atmos->ngpt, nband = atmos->nband;
  const size_t ncg = (size_t)ncol * ngpt, nclp = (size_t)ncol * (nlay + 1), ncl = (size_t)ncol * nlay;
  const bool do_broadband = fluxes != nullptr;
  const Bool has_dif_bc = inc_flux_dif != nullptr;
  std::string msg;
  if (do_broadband && !(fluxes->flux_up || fluxes->flux_dn || fluxes->flux_net || fluxes->flux_dn_dir))
    msg = "rte_sw: no space allocated for fluxes";
  if (g_check_values) {  // mo_rte_sw.F90:176-191
    if (rrtmgpb_any_vals_outside(ncol, mu0, nullptr, -1, 1)) msg = "rte_sw: one or more mu0 < -1 or > 1";
    if (rrtmgpb_any_vals_less_than(ncg, inc_flux, nullptr, 0)) msg = "rte_sw: one or more inc_flux < 0";
    if (rrtmgpb_any_vals_outside((size_t)nband * ncol, sfc_alb_dir, nullptr, 0, 1))
      msg = "rte_sw: sfc_alb_dir out of bounds [0,1]";
    if (rrtmgpb_any_vals_outside((size_t)nband * ncol, sfc_alb_dif, nullptr, 0, 1))
      msg = "rte_sw: sfc_alb_dif out of bounds [0,1]";
    if (has_dif_bc && rrtmgpb_any_vals_less_than(ncg, inc_flux_dif, nullptr, 0))
      msg = "rte_sw: one or more inc_flux_dif < 0";
  }
  if (!msg.empty()) return fail(errmsg, msg);

  // mu0 is constant with height: mo_rte_sw.F90:87-93
  Scratch<Float> mu0_bylay(ncl);
  rrtmgpb_broadcast_by_lay(ncol, nlay, mu0, mu0_bylay);
  // output plumbing :197-240.  In the broadband case the three g-point flux arguments are decoys; the
  // reference points all three at ONE buffer (:204-207).  We pass one small decoy: the kernels never touch it.
  Float *up_loc, *dn_loc, *dir_loc, *up_tmp = nullptr, *dn_tmp = nullptr, *dir_tmp = nullptr, *decoy = nullptr;
  Float *gpt_up = gpt_up_out, *gpt_dn = gpt_dn_out, *gpt_dir = gpt_dir_out;
  if (do_broadband) {
    up_loc = fluxes->flux_up; dn_loc = fluxes->flux_dn; dir_loc = fluxes->flux_dn_dir;
    if (!up_loc) up_loc = up_tmp = static_cast<Float*>(rrtmgpb_mem_alloc(nclp * sizeof(Float)));
    if (!dn_loc) dn_loc = dn_tmp = static_cast<Float*>(rrtmgpb_mem_alloc(nclp * sizeof(Float)));
    if (!dir_loc) dir_loc = dir_tmp = static_cast<Float*>(rrtmgpb_mem_alloc(nclp * sizeof(Float)));
    decoy = static_cast<Float*>(rrtmgpb_mem_alloc(nclp * sizeof(Float)));
    gpt_up = gpt_dn = gpt_dir = decoy;
  } else {
    decoy = static_cast<Float*>(rrtmgpb_mem_alloc(nclp * sizeof(Float)));
    up_loc = dn_loc = dir_loc = decoy;  // :235-238
  }
  Scratch<Float> alb_dir_gpt(ncg), alb_dif_gpt(ncg);
  Scratch<int> lims(2 * (size_t)nband);
  rrtmgpb_mem_to_backend(lims, atmos->band_lims_gpt, sizeof(int) * 2 * nband);
  rrtmgpb_expand_and_transpose(ncol, nband, ngpt, lims, sfc_alb_dir, alb_dir_gpt);  // :266-267
  rrtmgpb_expand_and_transpose(ncol, nband, ngpt, lims, sfc_alb_dif, alb_dif_gpt);
  Float* dif_alloc = nullptr;
  const Float* inc_flux_diffuse = inc_flux_dif;
  if (!has_dif_bc) {
    dif_alloc = static_cast<Float*>(rrtmgpb_mem_alloc(ncg * sizeof(Float)));
    zero_array_2D(&ncol, &ngpt, dif_alloc);  // :279
    inc_flux_diffuse = dif_alloc;
  }
  if (g_check_values) {
    char verr[RRTMGPB_ERRLEN];
    if (rrtmgpb_op_validate(atmos, verr)) msg = verr;
  }
  const Bool top_at_1 = atmos->top_at_1 != 0, bb = do_broadband;
  if (msg.empty()) {
    if (atmos->kind == RRTMGPB_1SCL) {  // :286-311 direct beam only
      if (do_broadband) {
        msg = "rte_sw: broadband fluxes from 1scl optical properties are not supported by this frontend";
      } else {
        const int nlev = nlay + 1;
        rte_sw_solver_noscat(&ncol, &nlay, &ngpt, &top_at_1, atmos->tau, mu0_bylay, inc_flux, gpt_dir);
        zero_array_3D(&ncol, &nlev, &ngpt, gpt_up);
        rrtmgpb_mem_copy(gpt_dn, gpt_dir, nclp * ngpt * sizeof(Float));
      }
    } else if (atmos->kind == RRTMGPB_2STR) {  // :313-326
      rte_sw_solver_2stream(&ncol, &nlay, &ngpt, &top_at_1, atmos->tau, atmos->ssa, atmos->g, mu0_bylay, alb_dir_gpt,
                            alb_dif_gpt, inc_flux, gpt_up, gpt_dn, gpt_dir, &has_dif_bc, inc_flux_diffuse, &bb,
                            up_loc, dn_loc, dir_loc);
    } else {
      msg = "sw_solver(...ty_optical_props_nstr...) not yet implemented";
    }
    if (msg.empty() && do_broadband && fluxes->flux_net) {  // :346-355
      const int nlev = nlay + 1;
      rte_net_broadband_precalc(&ncol, &nlev, dn_loc, up_loc, fluxes->flux_net);
    }
  }
  rrtmgpb_mem_free(dif_alloc);
  rrtmgpb_mem_free(up_tmp);
  rrtmgpb_mem_free(dn_tmp);
  rrtmgpb_mem_free(dir_tmp);
  rrtmgpb_mem_free(decoy);
  return fail(errmsg, msg);
}

int rrtmgpb_rte_sw(const rrtmgpb_optical_props* atmos, const Float* mu0, const Float* inc_flux,
                   const Float* sfc_alb_dir, const Float* sfc_alb_dif, rrtmgpb_fluxes_broadband* fluxes,
                   const Float* inc_flux_dif, char* errmsg) {
  if (!fluxes) return fail(errmsg, "rte_sw: no space allocated for fluxes");
  return rte_sw_impl(atmos, mu0, inc_flux, sfc_alb_dir, sfc_alb_dif, fluxes, nullptr, nullptr, nullptr, inc_flux_dif,
                     errmsg);
}
int rrtmgpb_rte_sw_bygpoint(const rrtmgpb_optical_props* atmos, const Float* mu0, const Float* inc_flux,
                            const Float* sfc_alb_dir, const Float* sfc_alb_dif, Float* gpt_flux_up,
                            Float* gpt_flux_dn, Float* gpt_flux_dir, const Float* inc_flux_dif, char* errmsg) {
  if (!gpt_flux_up || !gpt_flux_dn || !gpt_flux_dir) return fail(errmsg, "rte_sw: no space allocated for fluxes");
  return rte_sw_impl(atmos, mu0, inc_flux, sfc_alb_dir, sfc_alb_dif, nullptr, gpt_flux_up, gpt_flux_dn, gpt_flux_dir,
                     inc_flux_dif, errmsg);
}

// ------------------------------------------------------------------------------------------------
// ty_gas_optics_rrtmgp
// ------------------------------------------------------------------------------------------------
struct rrtmgpb_gas_optics_t {
  rrtmgpb_kdist h;  // scalars + HOST copies of the small index tables
  std::vector<int> band_lims_gpt_h;
  std::vector<Float> band_lims_wvn_h;
  // backend copies
  int *flavor, *gpoint_flavor, *band_lims_gpt, *gpoint_bands;
  Float *press_ref_log, *temp_ref, *vmr_ref, *kmajor, *kminor_lower, *kminor_upper;
  int *mlg_l, *mlg_u, *im_l, *im_u, *is_l, *is_u, *ks_l, *ks_u;
  Bool *sd_l, *sd_u, *sc_l, *sc_u;
  Float *planck_frac, *totplnk, *krayl, *solar_source;
  Float* optimal_angle_fit = nullptr;  // (2,nbnd), backend; set by rrtmgpb_gas_optics_set_optimal_angle_fit
};

rrtmgpb_gas_optics_t* rrtmgpb_gas_optics_load(const rrtmgpb_kdist* t, char* errmsg) {
  if (!t || !t->kmajor) { fail(errmsg, "ERROR: spectral configuration not loaded"); return nullptr; }
  if ((t->totplnk != nullptr) == (t->solar_source != nullptr)) {
    fail(errmsg, "gas_optics%load: provide either the Planck tables (LW) or the solar source (SW)");
    return nullptr;
  }
  rrtmgpb_gas_optics_t* go = new rrtmgpb_gas_optics_t();
  go->h = *t;
  go->band_lims_gpt_h.assign(t->band_lims_gpt, t->band_lims_gpt + 2 * t->nbnd);
  if (t->band_lims_wvn) go->band_lims_wvn_h.assign(t->band_lims_wvn, t->band_lims_wvn + 2 * t->nbnd);
  const size_t tn = (size_t)t->ntemp * t->neta, nl = (size_t)(t->nminorlower > 0 ? t->nminorlower : 1),
               nu = (size_t)(t->nminorupper > 0 ? t->nminorupper : 1);
  go->flavor = upload(t->flavor, 2 * (size_t)t->nflav);
  go->gpoint_flavor = upload(t->gpoint_flavor, 2 * (size_t)t->ngpt);
  go->band_lims_gpt = upload(t->band_lims_gpt, 2 * (size_t)t->nbnd);
  go->gpoint_bands = upload(t->gpoint_bands, (size_t)t->ngpt);
  go->press_ref_log = upload(t->press_ref_log, (size_t)t->npres);
  go->temp_ref = upload(t->temp_ref, (size_t)t->ntemp);
  go->vmr_ref = upload(t->vmr_ref, 2 * (size_t)(t->ngas + 1) * t->ntemp);
  go->kmajor = upload(t->kmajor, tn * (t->npres + 1) * t->ngpt);
  go->kminor_lower = upload(t->kminor_lower, tn * (t->nminorklower > 0 ? t->nminorklower : 1));
  go->kminor_upper = upload(t->kminor_upper, tn * (t->nminorkupper > 0 ? t->nminorkupper : 1));
  go->mlg_l = upload(t->minor_limits_gpt_lower, 2 * nl);
  go->mlg_u = upload(t->minor_limits_gpt_upper, 2 * nu);
  go->sd_l = upload(t->minor_scales_with_density_lower, nl);
  go->sd_u = upload(t->minor_scales_with_density_upper, nu);
  go->sc_l = upload(t->scale_by_complement_lower, nl);
  go->sc_u = upload(t->scale_by_complement_upper, nu);
  go->im_l = upload(t->idx_minor_lower, nl);
  go->im_u = upload(t->idx_minor_upper, nu);
  go->is_l = upload(t->idx_minor_scaling_lower, nl);
  go->is_u = upload(t->idx_minor_scaling_upper, nu);
  go->ks_l = upload(t->kminor_start_lower, nl);
  go->ks_u = upload(t->kminor_start_upper, nu);
  go->planck_frac = upload(t->planck_frac, t->planck_frac ? tn * (t->npres + 1) * t->ngpt : 0);
  go->totplnk = upload(t->totplnk, t->totplnk ? (size_t)t->nPlanckTemp * t->nbnd : 0);
  go->krayl = upload(t->krayl, t->krayl ? tn * t->ngpt * 2 : 0);
  go->solar_source = upload(t->solar_source, t->solar_source ? (size_t)t->ngpt : 0);
  rrtmgpb_sync();  // host tables may be released by the caller after load() returns
  ok(errmsg);
  return go;
}

void rrtmgpb_gas_optics_free(rrtmgpb_gas_optics_t* go) {
  if (!go) return;
  void* ptrs[] = {go->flavor, go->gpoint_flavor, go->band_lims_gpt, go->gpoint_bands, go->press_ref_log, go->temp_ref,
                  go->vmr_ref, go->kmajor, go->kminor_lower, go->kminor_upper, go->mlg_l, go->mlg_u, go->im_l, go->im_u,
                  go->is_l, go->is_u, go->ks_l, go->ks_u, go->sd_l, go->sd_u, go->sc_l, go->sc_u, go->planck_frac,
                  go->totplnk, go->krayl, go->solar_source};
  for (void* p : ptrs) rrtmgpb_mem_free(p);
  rrtmgpb_mem_free(go->optimal_angle_fit);
  delete go;
}

// optimal_angle_fit(2,nbnd) is part of the LW k-distribution files (load_int, mo_gas_optics_rrtmgp.F90:1040-1045); HOST pointer
int rrtmgpb_gas_optics_set_optimal_angle_fit(rrtmgpb_gas_optics_t* go, const Float* fit, char* errmsg) {
  if (!go || !fit) return fail(errmsg, "gas_optics%load: optimal_angle_fit missing");
  const size_t bytes = 2 * (size_t)go->h.nbnd * sizeof(Float);
  if (!go->optimal_angle_fit) go->optimal_angle_fit = static_cast<Float*>(rrtmgpb_mem_alloc(bytes));
  rrtmgpb_mem_to_backend(go->optimal_angle_fit, fit, bytes);
  return ok(errmsg);
}

// compute_optimal_angles, mo_gas_optics_rrtmgp.F90:1503-1562
int rrtmgpb_gas_optics_compute_optimal_angles(const rrtmgpb_gas_optics_t* go, const rrtmgpb_optical_props* op, int ncol_out,
                                              int ngpt_out, Float* optimal_angles, char* errmsg) {
  const rrtmgpb_kdist& k = go->h;
  if (!go->optimal_angle_fit) return fail(errmsg, "gas_optics%compute_optimal_angles: no optimal_angle_fit in this k-distribution");
  bool same = op->ngpt == k.ngpt && op->nband == k.nbnd;  // gpoints_are_equal, :1532
  for (int i = 0; same && i < 2 * k.nbnd; ++i) same = op->band_lims_gpt[i] == go->band_lims_gpt_h[i];
  if (!same)
    return fail(errmsg, "gas_optics%compute_optimal_angles: optical_props has different spectral discretization than gas_optics");
  if (ncol_out != op->ncol || ngpt_out != op->ngpt)  // :1534-1535
    return fail(errmsg, "gas_optics%compute_optimal_angles: optimal_angles different dimension (ncol)");
  rrtmgpb_compute_optimal_angles(op->ncol, op->nlay, op->ngpt, k.nbnd, go->band_lims_gpt, op->tau, go->optimal_angle_fit,
                                 optimal_angles);
  return ok(errmsg);
}

/* table dimensions and the HOST copy of band_lims_gpt (2,nbnd) of a loaded k-distribution */
void rrtmgpb_gas_optics_dims(const rrtmgpb_gas_optics_t* go, int* ngas, int* nbnd, int* ngpt) {
  if (ngas) *ngas = go->h.ngas;
  if (nbnd) *nbnd = go->h.nbnd;
  if (ngpt) *ngpt = go->h.ngpt;
}
const int* rrtmgpb_gas_optics_band_lims_gpt(const rrtmgpb_gas_optics_t* go) { return go->band_lims_gpt_h.data(); }
/* backend address of the loaded kmajor table: the key of the g-point-fastest copies (rrtmgpb_tables_changed) */
Float* rrtmgpb_gas_optics_kmajor(const rrtmgpb_gas_optics_t* go) { return go ? go->kmajor : nullptr; }
int rrtmgpb_gas_optics_source_is_internal(const rrtmgpb_gas_optics_t* go) { return go->totplnk != nullptr; }

// Interpolation intermediates shared by compute_gas_taus and source (frontend locals in the reference,
// mo_gas_optics_rrtmgp.F90:244-247,455-459)
struct InterpScratch {
  Scratch<int> jtemp, jpress, jeta;
  Scratch<Bool> tropo;
  Scratch<Float> fmajor, fminor, col_mix, col_gas;
  InterpScratch(size_t ncl, int nflav, int ngas)
      : jtemp(ncl), jpress(ncl), jeta(2 * ncl * nflav), tropo(ncl), fmajor(8 * ncl * nflav), fminor(4 * ncl * nflav),
        col_mix(2 * ncl * nflav), col_gas(ncl * (ngas + 1)) {}
};

// compute_gas_taus, mo_gas_optics_rrtmgp.F90:419-745
static int compute_gas_taus(const rrtmgpb_gas_optics_t* go, int ncol, int nlay, const Float* play, const Float* plev,
                            const Float* tlay, const Float* vmr, rrtmgpb_optical_props* op, InterpScratch& s,
                            const Float* col_dry, char* errmsg) {
  const rrtmgpb_kdist& k = go->h;
  const int ngpt = k.ngpt, nband = k.nbnd, ngas = k.ngas, nflav = k.nflav;
  const size_t ncl = (size_t)ncol * nlay;
  std::string msg;
  if (g_check_extents) {  // :491-507
    if (op->ncol != ncol || op->nlay != nlay || op->ngpt != ngpt)
      msg = "gas_optics(): optical properties have the wrong extents";
  }
  if (msg.empty() && g_check_values) {  // :509-521
    if (rrtmgpb_any_vals_outside(ncl, play, nullptr, k.press_ref_min, k.press_ref_max))
      msg = "gas_optics(): array play has values outside range";
    if (rrtmgpb_any_vals_less_than(ncl + ncol, plev, nullptr, 0)) msg = "gas_optics(): array plev has values outside range";
    if (rrtmgpb_any_vals_outside(ncl, tlay, nullptr, k.temp_ref_min, k.temp_ref_max))
      msg = "gas_optics(): array tlay has values outside range";
    if (col_dry && rrtmgpb_any_vals_less_than(ncl, col_dry, nullptr, 0))
      msg = "gas_optics(): array col_dry has values outside range";
  }
  if (!msg.empty()) return fail(errmsg, msg);

  // dry-air column amounts :578-590 and column gas amounts :594-609
  Float* col_dry_alloc = nullptr;
  const Float* col_dry_wk = col_dry;
  if (!col_dry) {
    col_dry_alloc = static_cast<Float*>(rrtmgpb_mem_alloc(ncl * sizeof(Float)));
    rrtmgpb_get_col_dry(ncol, nlay, vmr + ncl * (size_t)(k.idx_h2o - 1), plev, col_dry_alloc);
    col_dry_wk = col_dry_alloc;
  }
  rrtmgpb_col_gas_from_vmr(ncol, nlay, ngas, vmr, col_dry_wk, s.col_gas);
  // :615-633
  rrtmgp_interpolation(&ncol, &nlay, &ngas, &nflav, &k.neta, &k.npres, &k.ntemp, go->flavor, go->press_ref_log,
                       go->temp_ref, &k.press_ref_log_delta, &k.temp_ref_min, &k.temp_ref_delta,
                       &k.press_ref_trop_log, go->vmr_ref, play, tlay, s.col_gas, s.jtemp, s.fmajor, s.fminor,
                       s.col_mix, s.tropo, s.jeta, s.jpress);
  auto tau_abs = [&](Float* tau) {  // :637-665 / :679-706 (zero_array + accumulate == assign)
    // this frontend's tables are immutable and released through rrtmgpb_mem_free: the kernel may keep
    // g-point-fastest copies of them for the duration of this call's lookup (scoped switch, thread-local)
    rrtmgpb_abi_table_cache(1);
    rrtmgpb_compute_tau_absorption_assign(
        ncol, nlay, nband, ngpt, ngas, nflav, k.neta, k.npres, k.ntemp, k.nminorlower, k.nminorklower, k.nminorupper,
        k.nminorkupper, k.idx_h2o, go->gpoint_flavor, go->band_lims_gpt, go->kmajor, go->kminor_lower, go->kminor_upper,
        go->mlg_l, go->mlg_u, go->sd_l, go->sd_u, go->sc_l, go->sc_u, go->im_l, go->im_u, go->is_l, go->is_u, go->ks_l,
        go->ks_u, s.tropo, s.col_mix, s.fmajor, s.fminor, play, tlay, s.col_gas, s.jeta, s.jtemp, s.jpress, tau);
    rrtmgpb_abi_table_cache(0);
  };
  if (go->krayl) {  // :634-677
    Scratch<Float> tau_rayleigh(ncl * ngpt);
    tau_abs(op->tau);  // absorption lands in op->tau, combined in place below
    rrtmgp_compute_tau_rayleigh(&ncol, &nlay, &nband, &ngpt, &ngas, &nflav, &k.neta, &k.npres, &k.ntemp,
                                go->gpoint_flavor, go->band_lims_gpt, go->krayl, &k.idx_h2o, col_dry_wk, s.col_gas,
                                s.fminor, s.jeta, s.tropo, s.jtemp, tau_rayleigh);
    if (op->kind == RRTMGPB_NSTR) {
      rrtmgpb_mem_free(col_dry_alloc);
      return fail(errmsg, "gas_optics(): n-stream optical properties are not supported by this frontend");
    }
    rrtmgpb_combine_abs_and_rayleigh(ncol, nlay, ngpt, op->kind, op->tau, tau_rayleigh, op->tau, op->ssa, op->g);
  } else {
    tau_abs(op->tau);
    if (op->kind == RRTMGPB_2STR) {  // :708-710
      zero_array_3D(&ncol, &nlay, &ngpt, op->ssa);
      zero_array_3D(&ncol, &nlay, &ngpt, op->g);
    } else if (op->kind == RRTMGPB_NSTR) {
      zero_array_3D(&ncol, &nlay, &ngpt, op->ssa);
      zero_array_4D(&op->nmom, &ncol, &nlay, &ngpt, op->p);
    }
  }
  rrtmgpb_mem_free(col_dry_alloc);
  return ok(errmsg);
}

// A driver that already knows the vertical orientation (it holds the pressures on the host) can spare every
// gas-optics call the two device reads + stream synchronisations below: rrtmgpb_set_top_at_1_hint(0 / 1), -1 = unset.
// Per calling thread.
static thread_local int tl_top_at_1_hint = -1;
void rrtmgpb_set_top_at_1_hint(int top_at_1) { tl_top_at_1_hint = top_at_1 < 0 ? -1 : (top_at_1 != 0); }
static int orientation(const Float* play, int ncol, int nlay) {
  if (tl_top_at_1_hint >= 0) return tl_top_at_1_hint;
  // mo_gas_optics_rrtmgp.F90:258: top_at_1 = play(1,1) < play(1,nlay) - needs two values on the host
  Float a = 0, b = 0;
  rrtmgpb_mem_to_host(&a, play, sizeof(Float));
  rrtmgpb_mem_to_host(&b, play + (size_t)ncol * (nlay - 1), sizeof(Float));
  return a < b;
}
static int set_top_at_1(rrtmgpb_optical_props* op, const Float* play, int ncol, int nlay) {
  op->top_at_1 = orientation(play, ncol, nlay);
  return op->top_at_1;
}

int rrtmgpb_gas_optics_int(const rrtmgpb_gas_optics_t* go, int ncol, int nlay, const Float* play, const Float* plev,
                           const Float* tlay, const Float* tsfc, const Float* vmr, rrtmgpb_optical_props* op,
                           rrtmgpb_source_func_lw* sources, const Float* col_dry, const Float* tlev, char* errmsg) {
  const rrtmgpb_kdist& k = go->h;
  if (!go->totplnk) return fail(errmsg, "gas_optics(): no internal (Planck) source tables loaded");
  const int ngpt = k.ngpt, nband = k.nbnd;
  const size_t ncl = (size_t)ncol * nlay;
  set_top_at_1(op, play, ncol, nlay);
  InterpScratch s(ncl, k.nflav, k.ngas);
  if (compute_gas_taus(go, ncol, nlay, play, plev, tlay, vmr, op, s, col_dry, errmsg)) return 1;
  std::string msg;
  if (g_check_extents) {  // :285-291
    if (sources->ncol != ncol || sources->nlay != nlay || sources->ngpt != ngpt)
      msg = "gas_optics%gas_optics: source function arrays inconsistently sized";
  }
  if (msg.empty() && g_check_values) {  // :294-301
    if (rrtmgpb_any_vals_outside(ncol, tsfc, nullptr, k.temp_ref_min, k.temp_ref_max))
      msg = "gas_optics(): array tsfc has values outside range";
    if (tlev && rrtmgpb_any_vals_outside(ncl + ncol, tlev, nullptr, k.temp_ref_min, k.temp_ref_max))
      msg = "gas_optics(): array tlev has values outside range";
  }
  if (!msg.empty()) return fail(errmsg, msg);
  // source(), :840-928
  Float* tlev_alloc = nullptr;
  const Float* tlev_wk = tlev;
  if (!tlev) {
    tlev_alloc = static_cast<Float*>(rrtmgpb_mem_alloc((ncl + ncol) * sizeof(Float)));
    rrtmgpb_interpolate_tlev(ncol, nlay, play, plev, tlay, tlev_alloc);
    tlev_wk = tlev_alloc;
  }
  const int sfc_lay = op->top_at_1 ? nlay : 1;  // :920 merge(nlay, 1, top_at_1)
  rrtmgpb_abi_table_cache(1);   // immutable tables, released through rrtmgpb_mem_free (see tau_abs above)
  rrtmgp_compute_Planck_source(&ncol, &nlay, &nband, &ngpt, &k.nflav, &k.neta, &k.npres, &k.ntemp, &k.nPlanckTemp, tlay,
                               tlev_wk, tsfc, &sfc_lay, s.fmajor, s.jeta, s.tropo, s.jtemp, s.jpress, go->gpoint_bands,
                               go->band_lims_gpt, go->planck_frac, &k.temp_ref_min, &k.totplnk_delta, go->totplnk,
                               go->gpoint_flavor, sources->sfc_source, sources->lay_source, sources->lev_source,
                               sources->sfc_source_Jac);
  rrtmgpb_abi_table_cache(0);
  rrtmgpb_mem_free(tlev_alloc);
  return ok(errmsg);
}

int rrtmgpb_gas_optics_ext(const rrtmgpb_gas_optics_t* go, int ncol, int nlay, const Float* play, const Float* plev,
                           const Float* tlay, const Float* vmr, rrtmgpb_optical_props* op, Float* toa_src,
                           const Float* col_dry, char* errmsg) {
  const rrtmgpb_kdist& k = go->h;
  if (!go->solar_source) return fail(errmsg, "gas_optics(): no external (solar) source loaded");
  const size_t ncl = (size_t)ncol * nlay;
  set_top_at_1(op, play, ncol, nlay);
  InterpScratch s(ncl, k.nflav, k.ngas);
  if (compute_gas_taus(go, ncol, nlay, play, plev, tlay, vmr, op, s, col_dry, errmsg)) return 1;
  rrtmgpb_broadcast_by_gpt(ncol, k.ngpt, go->solar_source, toa_src);  // :405-411
  return ok(errmsg);
}

// ------------------------------------------------------------------------------------------------
// Fused fast path: gas_optics() and the driver's clouds%increment(atmos) in one pass over the planes
// (rrtmgpb_gas_optics_fused).  Same checks, same caller-visible results as the two reference calls.
// ------------------------------------------------------------------------------------------------
static rrtmgpb_gas_tables tables_of(const rrtmgpb_gas_optics_t* go) {
  const rrtmgpb_kdist& k = go->h;
  rrtmgpb_gas_tables t;
  t.ngas = k.ngas; t.nflav = k.nflav; t.neta = k.neta; t.npres = k.npres; t.ntemp = k.ntemp; t.nbnd = k.nbnd;
  t.ngpt = k.ngpt; t.nminorlower = k.nminorlower; t.nminorklower = k.nminorklower; t.nminorupper = k.nminorupper;
  t.nminorkupper = k.nminorkupper; t.idx_h2o = k.idx_h2o;
  t.flavor = go->flavor; t.gpoint_flavor = go->gpoint_flavor; t.band_lims_gpt = go->band_lims_gpt;
  t.gpoint_bands = go->gpoint_bands; t.press_ref_log = go->press_ref_log; t.temp_ref = go->temp_ref;
  t.vmr_ref = go->vmr_ref; t.press_ref_log_delta = k.press_ref_log_delta; t.temp_ref_min = k.temp_ref_min;
  t.temp_ref_delta = k.temp_ref_delta; t.press_ref_trop_log = k.press_ref_trop_log;
  t.kmajor = go->kmajor; t.kminor_lower = go->kminor_lower; t.kminor_upper = go->kminor_upper;
  t.minor_limits_gpt_lower = go->mlg_l; t.minor_limits_gpt_upper = go->mlg_u;
  t.minor_scales_with_density_lower = go->sd_l; t.minor_scales_with_density_upper = go->sd_u;
  t.scale_by_complement_lower = go->sc_l; t.scale_by_complement_upper = go->sc_u;
  t.idx_minor_lower = go->im_l; t.idx_minor_upper = go->im_u; t.idx_minor_scaling_lower = go->is_l;
  t.idx_minor_scaling_upper = go->is_u; t.kminor_start_lower = go->ks_l; t.kminor_start_upper = go->ks_u;
  t.krayl = go->krayl; t.planck_frac = go->planck_frac; t.totplnk = go->totplnk; t.nPlanckTemp = k.nPlanckTemp;
  t.totplnk_delta = k.totplnk_delta;
  return t;
}

static int gas_optics_fused_impl(const rrtmgpb_gas_optics_t* go, int ncol, int nlay, const Float* play,
                                 const Float* plev, const Float* tlay, const Float* tsfc, const Float* vmr,
                                 rrtmgpb_optical_props* op, rrtmgpb_source_func_lw* sources, Float* toa_src,
                                 const Float* col_dry, const Float* tlev, const rrtmgpb_optical_props* clouds,
                                 const rrtmgpb_optical_props* aerosols, char* errmsg) {
  const rrtmgpb_kdist& k = go->h;
  const size_t ncl = (size_t)ncol * nlay;
  std::string msg;
  set_top_at_1(op, play, ncol, nlay);
  // checks of compute_gas_taus (:491-521), gas_optics_int (:285-301) and increment (:893-905,956-961)
  if (g_check_extents) {
    if (op->ncol != ncol || op->nlay != nlay || op->ngpt != k.ngpt)
      msg = "gas_optics(): optical properties have the wrong extents";
    if (sources && (sources->ncol != ncol || sources->nlay != nlay || sources->ngpt != k.ngpt))
      msg = "gas_optics%gas_optics: source function arrays inconsistently sized";
  }
  if (msg.empty() && g_check_values) {
    if (rrtmgpb_any_vals_outside(ncl, play, nullptr, k.press_ref_min, k.press_ref_max))
      msg = "gas_optics(): array play has values outside range";
    if (rrtmgpb_any_vals_less_than(ncl + ncol, plev, nullptr, 0)) msg = "gas_optics(): array plev has values outside range";
    if (rrtmgpb_any_vals_outside(ncl, tlay, nullptr, k.temp_ref_min, k.temp_ref_max))
      msg = "gas_optics(): array tlay has values outside range";
    if (col_dry && rrtmgpb_any_vals_less_than(ncl, col_dry, nullptr, 0))
      msg = "gas_optics(): array col_dry has values outside range";
    if (sources && rrtmgpb_any_vals_outside(ncol, tsfc, nullptr, k.temp_ref_min, k.temp_ref_max))
      msg = "gas_optics(): array tsfc has values outside range";
    if (sources && tlev && rrtmgpb_any_vals_outside(ncl + ncol, tlev, nullptr, k.temp_ref_min, k.temp_ref_max))
      msg = "gas_optics(): array tlev has values outside range";
  }
  if (msg.empty() && op->kind == RRTMGPB_NSTR) msg = "gas_optics(): n-stream optical properties are not supported by this frontend";
  int cld_kind = 0, aer_kind = 0;
  for (int which = 0; which < 2 && msg.empty(); ++which) {  // the checks of increment() (:893-905,956-961) for both
    const rrtmgpb_optical_props* inc = which ? aerosols : clouds;
    if (!inc) continue;
    if (inc->ncol != ncol || inc->nlay != nlay)
      msg = "ty_optical_props%increment: optical properties objects have different ncol and/or nlay";
    else if (inc->nband != op->nband || inc->ngpt != op->nband)
      msg = "ty_optical_props%increment: optical properties objects have incompatible g-point structures";
    else if (inc->kind == RRTMGPB_NSTR)
      msg = "ty_optical_props%increment: n-stream properties are not supported by the fused path";
    (which ? aer_kind : cld_kind) = inc->kind;
  }
  if (!msg.empty()) return fail(errmsg, msg);
  Float* tlev_alloc = nullptr;
  const Float* tlev_wk = tlev;
  if (sources && !tlev) {
    tlev_alloc = static_cast<Float*>(rrtmgpb_mem_alloc((ncl + ncol) * sizeof(Float)));
    rrtmgpb_interpolate_tlev(ncol, nlay, play, plev, tlay, tlev_alloc);
    tlev_wk = tlev_alloc;
  }
  const rrtmgpb_gas_tables t = tables_of(go);
  const int sfc_lay = op->top_at_1 ? nlay : 1;
  rrtmgpb_gas_optics_fused(&t, ncol, nlay, play, plev, tlay, vmr, col_dry, op->kind, op->tau, op->ssa, op->g, cld_kind,
                           clouds ? clouds->tau : nullptr, clouds ? clouds->ssa : nullptr, clouds ? clouds->g : nullptr,
                           aer_kind, aerosols ? aerosols->tau : nullptr, aerosols ? aerosols->ssa : nullptr,
                           aerosols ? aerosols->g : nullptr, tlev_wk, tsfc, sfc_lay, sources ? sources->sfc_source : nullptr,
                           sources ? sources->lay_source : nullptr, sources ? sources->lev_source : nullptr,
                           sources ? sources->sfc_source_Jac : nullptr);
  if (toa_src) rrtmgpb_broadcast_by_gpt(ncol, k.ngpt, go->solar_source, toa_src);
  rrtmgpb_mem_free(tlev_alloc);
  return ok(errmsg);
}

int rrtmgpb_gas_optics_int_fused(const rrtmgpb_gas_optics_t* go, int ncol, int nlay, const Float* play,
                                 const Float* plev, const Float* tlay, const Float* tsfc, const Float* vmr,
                                 rrtmgpb_optical_props* op, rrtmgpb_source_func_lw* sources, const Float* col_dry,
                                 const Float* tlev, const rrtmgpb_optical_props* clouds,
                                 const rrtmgpb_optical_props* aerosols, char* errmsg) {
  if (!go->totplnk) return fail(errmsg, "gas_optics(): no internal (Planck) source tables loaded");
  return gas_optics_fused_impl(go, ncol, nlay, play, plev, tlay, tsfc, vmr, op, sources, nullptr, col_dry, tlev, clouds,
                               aerosols, errmsg);
}

int rrtmgpb_gas_optics_ext_fused(const rrtmgpb_gas_optics_t* go, int ncol, int nlay, const Float* play,
                                 const Float* plev, const Float* tlay, const Float* vmr, rrtmgpb_optical_props* op,
                                 Float* toa_src, const Float* col_dry, const rrtmgpb_optical_props* clouds,
                                 const rrtmgpb_optical_props* aerosols, char* errmsg) {
  if (!go->solar_source) return fail(errmsg, "gas_optics(): no external (solar) source loaded");
  return gas_optics_fused_impl(go, ncol, nlay, play, plev, tlay, nullptr, vmr, op, nullptr, toa_src, col_dry, nullptr,
                               clouds, aerosols, errmsg);
}

// ------------------------------------------------------------------------------------------------
// Express path: gas_optics + clouds%increment(atmos) + rte_lw / rte_sw with ty_fluxes_broadband in ONE
// call that never allocates a (ncol, nlay, ngpt) array.  Same checks and error strings as the calls it replaces
// (mo_gas_optics_rrtmgp.F90:491-521,285-301; mo_optical_props.F90:893-905; mo_rte_lw.F90:170-262; mo_rte_sw.F90:176-191).
// ------------------------------------------------------------------------------------------------
static int express_impl(const rrtmgpb_gas_optics_t* go, int ncol, int nlay, const Float* play, const Float* plev,
                        const Float* tlay, const Float* tsfc, const Float* vmr, const Float* col_dry, const Float* tlev,
                        const rrtmgpb_optical_props* clouds, const Float* sfc_a, const Float* sfc_b, const Float* mu0,
                        int n_gauss_angles, rrtmgpb_fluxes_broadband* fluxes, char* errmsg) {
  const rrtmgpb_kdist& k = go->h;
  const bool sw = go->solar_source != nullptr;
  const char* who = sw ? "rte_sw" : "rte_lw";
  const size_t ncl = (size_t)ncol * nlay, nclp = ncl + ncol;
  std::string msg;
  if (!fluxes || !(fluxes->flux_up || fluxes->flux_dn || fluxes->flux_net || (sw && fluxes->flux_dn_dir)))
    return fail(errmsg, std::string(who) + ": no space allocated for fluxes");
  if (g_check_values) {
    if (rrtmgpb_any_vals_outside(ncl, play, nullptr, k.press_ref_min, k.press_ref_max))
      msg = "gas_optics(): array play has values outside range";
    if (rrtmgpb_any_vals_less_than(nclp, plev, nullptr, 0)) msg = "gas_optics(): array plev has values outside range";
    if (rrtmgpb_any_vals_outside(ncl, tlay, nullptr, k.temp_ref_min, k.temp_ref_max))
      msg = "gas_optics(): array tlay has values outside range";
    if (col_dry && rrtmgpb_any_vals_less_than(ncl, col_dry, nullptr, 0))
      msg = "gas_optics(): array col_dry has values outside range";
    if (!sw && rrtmgpb_any_vals_outside(ncol, tsfc, nullptr, k.temp_ref_min, k.temp_ref_max))
      msg = "gas_optics(): array tsfc has values outside range";
    if (!sw && tlev && rrtmgpb_any_vals_outside(nclp, tlev, nullptr, k.temp_ref_min, k.temp_ref_max))
      msg = "gas_optics(): array tlev has values outside range";
    if (sw) {
      if (rrtmgpb_any_vals_outside(ncol, mu0, nullptr, -1, 1)) msg = "rte_sw: one or more mu0 < -1 or > 1";
      if (rrtmgpb_any_vals_outside((size_t)k.nbnd * ncol, sfc_a, nullptr, 0, 1)) msg = "rte_sw: sfc_alb_dir out of bounds [0,1]";
      if (rrtmgpb_any_vals_outside((size_t)k.nbnd * ncol, sfc_b, nullptr, 0, 1)) msg = "rte_sw: sfc_alb_dif out of bounds [0,1]";
    } else {
      if (rrtmgpb_any_vals_outside((size_t)k.nbnd * ncol, sfc_a, nullptr, 0, 1)) msg = "rte_lw: sfc_emis has values < 0 or > 1";
      if (n_gauss_angles > max_gauss_pts) msg = "rte_lw: asking for too many quadrature points for no-scattering calculation";
      if (n_gauss_angles < 0) msg = "rte_lw: have to ask for at least one quadrature point for no-scattering calculation";
    }
  }
  int cld_kind = 0;
  if (msg.empty() && clouds) {
    if (clouds->ncol != ncol || clouds->nlay != nlay)
      msg = "ty_optical_props%increment: optical properties objects have different ncol and/or nlay";
    else if (clouds->nband != k.nbnd || clouds->ngpt != k.nbnd)
      msg = "ty_optical_props%increment: optical properties objects have incompatible g-point structures";
    else if (clouds->kind == RRTMGPB_NSTR)
      msg = "ty_optical_props%increment: n-stream properties are not supported by the fused path";
    else if (g_check_values) {
      char verr[RRTMGPB_ERRLEN];
      if (rrtmgpb_op_validate(clouds, verr)) msg = verr;
    }
    cld_kind = clouds->kind;
  }
  if (!msg.empty()) return fail(errmsg, msg);
  const int top_at_1 = orientation(play, ncol, nlay);  // play(1,1) < play(1,nlay), mo_gas_optics_rrtmgp.F90:258
  Float* tlev_alloc = nullptr;
  const Float* tlev_wk = tlev;
  if (!sw && !tlev) {
    tlev_alloc = static_cast<Float*>(rrtmgpb_mem_alloc(nclp * sizeof(Float)));
    rrtmgpb_interpolate_tlev(ncol, nlay, play, plev, tlay, tlev_alloc);
    tlev_wk = tlev_alloc;
  }
  const int nmus = n_gauss_angles > 0 ? n_gauss_angles : 1;
  Float Ds[max_gauss_pts], wts[max_gauss_pts];
  for (int imu = 0; imu < nmus; ++imu) {  // mo_rte_lw.F90:146-160,357-365
    Ds[imu] = (Float)1 / (Float)gauss_mus[nmus - 1][imu];
    wts[imu] = (Float)gauss_wts[nmus - 1][imu];
  }
  Float *up = fluxes->flux_up, *dn = fluxes->flux_dn, *dir = fluxes->flux_dn_dir, *up_t = nullptr, *dn_t = nullptr, *dir_t = nullptr;
  if (!up) up = up_t = static_cast<Float*>(rrtmgpb_mem_alloc(nclp * sizeof(Float)));
  if (!dn) dn = dn_t = static_cast<Float*>(rrtmgpb_mem_alloc(nclp * sizeof(Float)));
  if (sw && !dir) dir = dir_t = static_cast<Float*>(rrtmgpb_mem_alloc(nclp * sizeof(Float)));
  const rrtmgpb_gas_tables t = tables_of(go);
  rrtmgpb_express(&t, ncol, nlay, top_at_1, play, plev, tlay, tlev_wk, tsfc, vmr, col_dry, cld_kind,
                  clouds ? clouds->tau : nullptr, clouds ? clouds->ssa : nullptr, clouds ? clouds->g : nullptr, sfc_a, sfc_b,
                  mu0, go->solar_source, nmus, Ds, wts, up, dn, dir);
  if (fluxes->flux_net) {
    const int nlev = nlay + 1;
    rte_net_broadband_precalc(&ncol, &nlev, dn, up, fluxes->flux_net);
  }
  rrtmgpb_mem_free(tlev_alloc); rrtmgpb_mem_free(up_t); rrtmgpb_mem_free(dn_t); rrtmgpb_mem_free(dir_t);
  return ok(errmsg);
}

int rrtmgpb_rte_lw_express(const rrtmgpb_gas_optics_t* go, int ncol, int nlay, const Float* play, const Float* plev,
                           const Float* tlay, const Float* tsfc, const Float* vmr, const Float* col_dry, const Float* tlev,
                           const rrtmgpb_optical_props* clouds, const Float* sfc_emis, int n_gauss_angles,
                           rrtmgpb_fluxes_broadband* fluxes, char* errmsg) {
  if (!go->totplnk) return fail(errmsg, "gas_optics(): no internal (Planck) source tables loaded");
  return express_impl(go, ncol, nlay, play, plev, tlay, tsfc, vmr, col_dry, tlev, clouds, sfc_emis, nullptr, nullptr,
                      n_gauss_angles, fluxes, errmsg);
}

int rrtmgpb_rte_sw_express(const rrtmgpb_gas_optics_t* go, int ncol, int nlay, const Float* play, const Float* plev,
                           const Float* tlay, const Float* vmr, const Float* col_dry, const rrtmgpb_optical_props* clouds,
                           const Float* mu0, const Float* sfc_alb_dir, const Float* sfc_alb_dif,
                           rrtmgpb_fluxes_broadband* fluxes, char* errmsg) {
  if (!go->solar_source) return fail(errmsg, "gas_optics(): no external (solar) source loaded");
  return express_impl(go, ncol, nlay, play, plev, tlay, nullptr, vmr, col_dry, nullptr, clouds, sfc_alb_dir, sfc_alb_dif,
                      mu0, 0, fluxes, errmsg);
}

// ------------------------------------------------------------------------------------------------
// ty_cloud_optics_rrtmgp (LUT)
// ------------------------------------------------------------------------------------------------
struct rrtmgpb_cloud_optics_t {
  rrtmgpb_cloud_lut h;
  std::vector<Float> band_lims_wvn_h;
  int liq_nsteps, ice_nsteps;
  Float liq_step_size, ice_step_size;
  Float *extliq, *ssaliq, *asyliq, *extice, *ssaice, *asyice;  // ice: the icergh slice only
};

rrtmgpb_cloud_optics_t* rrtmgpb_cloud_optics_load(const rrtmgpb_cloud_lut* lut, char* errmsg) {
  if (!lut || !lut->extliq) { fail(errmsg, "cloud optics: no data has been initialized"); return nullptr; }
  if (lut->icergh < 1 || lut->icergh > lut->nrghice) {  // set_ice_roughness
    fail(errmsg, "cloud_optics%set_ice_roughness(): must be > 0");
    return nullptr;
  }
  rrtmgpb_cloud_optics_t* co = new rrtmgpb_cloud_optics_t();
  co->h = *lut;
  if (lut->band_lims_wvn) co->band_lims_wvn_h.assign(lut->band_lims_wvn, lut->band_lims_wvn + 2 * lut->nbnd);
  // mo_cloud_optics_rrtmgp.F90:130-133
  co->liq_nsteps = lut->nsize_liq;
  co->ice_nsteps = lut->nsize_ice;
  co->liq_step_size = (lut->radliq_upr - lut->radliq_lwr) / (Float)(lut->nsize_liq - 1);
  co->ice_step_size = (lut->diamice_upr - lut->diamice_lwr) / (Float)(lut->nsize_ice - 1);
  const size_t nl = (size_t)lut->nsize_liq * lut->nbnd, ni = (size_t)lut->nsize_ice * lut->nbnd;
  co->extliq = upload(lut->extliq, nl);
  co->ssaliq = upload(lut->ssaliq, nl);
  co->asyliq = upload(lut->asyliq, nl);
  const size_t off = ni * (size_t)(lut->icergh - 1);  // this%extice(:,:,this%icergh), :380-385
  co->extice = upload(lut->extice + off, ni);
  co->ssaice = upload(lut->ssaice + off, ni);
  co->asyice = upload(lut->asyice + off, ni);
  rrtmgpb_sync();
  ok(errmsg);
  return co;
}

void rrtmgpb_cloud_optics_free(rrtmgpb_cloud_optics_t* co) {
  if (!co) return;
  void* ptrs[] = {co->extliq, co->ssaliq, co->asyliq, co->extice, co->ssaice, co->asyice};
  for (void* p : ptrs) rrtmgpb_mem_free(p);
  delete co;
}

static std::atomic<int> g_cloud_optics_one_pass{1};  // 0: the reference's kernel-by-kernel sequence (rrtmgpb_cloud_optics_one_pass)

int rrtmgpb_cloud_optics(const rrtmgpb_cloud_optics_t* co, int ncol, int nlay, const Float* clwp, const Float* ciwp,
                         const Float* reliq, const Float* dgice, rrtmgpb_optical_props* op, char* errmsg) {
  return rrtmgpb_cloud_optics_delta_scaled(co, ncol, nlay, clwp, ciwp, reliq, dgice, op, 0, errmsg);
}

int rrtmgpb_cloud_optics_delta_scaled(const rrtmgpb_cloud_optics_t* co, int ncol, int nlay, const Float* clwp,
                                      const Float* ciwp, const Float* reliq, const Float* dgice, rrtmgpb_optical_props* op,
                                      int delta_scale, char* errmsg) {
  const rrtmgpb_cloud_lut& h = co->h;
  const int ngpt = h.nbnd;  // by-band tables: ngpt == nbnd
  const size_t ncl = (size_t)ncol * nlay, n = ncl * ngpt;
  std::string msg;
  if (g_check_extents) {  // :299-313
    if (op->ncol != ncol || op->nlay != nlay) msg = "cloud optics: optical_props have wrong extents";
  }
  if (msg.empty() && g_check_values) {  // :318-322
    if (op->nband != h.nbnd || op->ngpt != ngpt)
      msg = "cloud optics: optical properties don't have the same band structure";
  }
  if (!msg.empty()) return fail(errmsg, msg);
  if (g_check_values) {  // :334-353: the masks are only needed here when the value checks are on
    Scratch<Bool> liqmsk(ncl), icemsk(ncl);
    rrtmgpb_cloud_masks(ncol, nlay, clwp, ciwp, liqmsk, icemsk);
    if (rrtmgpb_any_vals_outside(ncl, reliq, liqmsk, h.radliq_lwr, h.radliq_upr))
      msg = "cloud optics: liquid effective radius is out of bounds";
    if (rrtmgpb_any_vals_outside(ncl, dgice, icemsk, h.diamice_lwr, h.diamice_upr))
      msg = "cloud optics: ice effective diameter is out of bounds";
    if (rrtmgpb_any_vals_less_than(ncl, clwp, liqmsk, 0) || rrtmgpb_any_vals_less_than(ncl, ciwp, icemsk, 0))
      msg = "cloud optics: negative clwp or ciwp where clouds are supposed to be";
  }
  if (!msg.empty()) return fail(errmsg, msg);
  if (op->kind == RRTMGPB_NSTR) return fail(errmsg, "cloud optics: n-stream calculations not yet supported");
  if (g_cloud_optics_one_pass) {  // masks + both table lookups (:373, :380) + combination (:399-424) in one kernel
    rrtmgpb_cloud_optics_from_tables_ds(ncol, nlay, ngpt, op->kind, clwp, ciwp, reliq, dgice, co->liq_nsteps,
                                        co->liq_step_size, h.radliq_lwr, co->extliq, co->ssaliq, co->asyliq, co->ice_nsteps,
                                        co->ice_step_size, h.diamice_lwr, co->extice, co->ssaice, co->asyice, op->tau, op->ssa,
                                        op->g, delta_scale);
    return ok(errmsg);
  }
  Scratch<Bool> liqmsk(ncl), icemsk(ncl);
  rrtmgpb_cloud_masks(ncol, nlay, clwp, ciwp, liqmsk, icemsk);  // :334-341
  Scratch<Float> ltau(n), ltaussa(n), ltaussag(n), itau(n), itaussa(n), itaussag(n);
  rrtmgp_compute_cld_from_table(&ncol, &nlay, &ngpt, liqmsk, clwp, reliq, &co->liq_nsteps, &co->liq_step_size,
                                &h.radliq_lwr, co->extliq, co->ssaliq, co->asyliq, ltau, ltaussa, ltaussag);  // :373
  rrtmgp_compute_cld_from_table(&ncol, &nlay, &ngpt, icemsk, ciwp, dgice, &co->ice_nsteps, &co->ice_step_size,
                                &h.diamice_lwr, co->extice, co->ssaice, co->asyice, itau, itaussa, itaussag);  // :380
  rrtmgpb_cloud_combine(ncol, nlay, ngpt, op->kind, ltau, ltaussa, ltaussag, itau, itaussa, itaussag, op->tau, op->ssa,
                        op->g);  // :399-424
  if (delta_scale) return rrtmgpb_op_delta_scale(op, nullptr, errmsg);
  return ok(errmsg);
}

void rrtmgpb_cloud_optics_one_pass(int on) { g_cloud_optics_one_pass = on ? 1 : 0; }

// ------------------------------------------------------------------------------------------------
// McICA cloud sampling: rte/extensions/mo_cloud_sampling.F90 (extents are passed explicitly: C has no size())
// ------------------------------------------------------------------------------------------------
static int sampled_mask_checks(const char* who, int ngpt, int nlay, int ncol, int cf_ncol, int cf_nlay, const Float* cloud_frac,
                               bool exp_ran, int op_ncol, int op_nlay, const Float* overlap_param, int m_ncol, int m_nlay,
                               int m_ngpt, char* errmsg) {
  const std::string w(who);
  if (ncol != cf_ncol || nlay != cf_nlay)  // :142-145 / :224-227
    return fail(errmsg, w + ": sizes of randoms(ngpt,nlay,ncol) and cloud_frac(ncol,nlay) are inconsistent");
  if (exp_ran && (ncol != op_ncol || nlay - 1 != op_nlay))  // :228-231
    return fail(errmsg, w + ": sizes of randoms(ngpt,nlay,ncol) and overlap_param(ncol,nlay-1) are inconsistent");
  if (ncol != m_ncol || nlay != m_nlay || ngpt != m_ngpt)  // :146-149 / :232-235
    return fail(errmsg, w + ": sizes of randoms(ngpt,nlay,ncol) and cloud_mask(ncol,nlay,ngpt) are inconsistent");
  if (rrtmgpb_any_vals_outside((size_t)ncol * nlay, cloud_frac, nullptr, 0, 1))  // :150-153 / :237-240
    return fail(errmsg, w + ": cloud fraction values out of range [0,1]");
  if (exp_ran && rrtmgpb_any_vals_outside((size_t)ncol * (nlay - 1), overlap_param, nullptr, -1, 1))  // :241-244
    return fail(errmsg, w + ": overlap_param values out of range [-1,1]");
  return ok(errmsg);
}

int rrtmgpb_cloud_sampling_mask_max_ran(int ngpt, int nlay, int ncol, const Float* randoms, int cf_ncol, int cf_nlay,
                                        const Float* cloud_frac, int m_ncol, int m_nlay, int m_ngpt, Bool* cloud_mask,
                                        char* errmsg) {  // sampled_mask_max_ran, :125-192
  if (sampled_mask_checks("sampled_mask_max_ran", ngpt, nlay, ncol, cf_ncol, cf_nlay, cloud_frac, false, 0, 0, nullptr, m_ncol,
                          m_nlay, m_ngpt, errmsg))
    return 1;
  rrtmgpb_sampled_mask_max_ran(ncol, nlay, ngpt, randoms, cloud_frac, cloud_mask);
  return ok(errmsg);
}

int rrtmgpb_cloud_sampling_mask_exp_ran(int ngpt, int nlay, int ncol, const Float* randoms, int cf_ncol, int cf_nlay,
                                        const Float* cloud_frac, int op_ncol, int op_nlay, const Float* overlap_param,
                                        int m_ncol, int m_nlay, int m_ngpt, Bool* cloud_mask, char* errmsg) {  // :205-292
  if (sampled_mask_checks("sampled_mask_exp_ran", ngpt, nlay, ncol, cf_ncol, cf_nlay, cloud_frac, true, op_ncol, op_nlay,
                          overlap_param, m_ncol, m_nlay, m_ngpt, errmsg))
    return 1;
  rrtmgpb_sampled_mask_exp_ran(ncol, nlay, ngpt, randoms, cloud_frac, overlap_param, cloud_mask);
  return ok(errmsg);
}

int rrtmgpb_cloud_sampling_draw_samples(int m_ncol, int m_nlay, int m_ngpt, const Bool* cloud_mask,
                                        const rrtmgpb_optical_props* clouds, rrtmgpb_optical_props* clouds_sampled,
                                        char* errmsg) {  // draw_samples, :36-120
  if (!clouds->tau) return fail(errmsg, "draw_samples: cloud optical properties are not initialized");
  if (!clouds_sampled->tau) return fail(errmsg, "draw_samples: sampled cloud optical properties are not initialized");
  if (clouds->kind == RRTMGPB_NSTR) return fail(errmsg, "draw_samples: sampling isn't implemented yet for ty_optical_props_nstr");
  if (clouds->kind != clouds_sampled->kind)
    return fail(errmsg, "draw_samples: by-band and sampled cloud properties need to be the same variable type");
  if (!bands_are_equal(clouds, clouds_sampled))
    return fail(errmsg, "draw_samples: by-band and sampled cloud properties spectral structure is different");
  const int ncol = clouds->ncol, nlay = clouds->nlay, nbnd = clouds->nband, ngpt = clouds_sampled->ngpt;
  if (m_ncol != ncol || m_nlay != nlay || m_ngpt != ngpt)
    return fail(errmsg, "draw_samples: cloud mask and cloud optical properties have different ncol, nlay and/or ngpt");
  if (clouds_sampled->ncol != ncol || clouds_sampled->nlay != nlay)
    return fail(errmsg, "draw_samples: sampled/unsampled cloud optical properties have different ncol and/or nlay");
  // band limits of the sampled object live on the host: stage them once for the three calls
  Scratch<int> lims(2 * (size_t)nbnd);
  rrtmgpb_mem_to_backend(lims, clouds_sampled->band_lims_gpt, 2 * (size_t)nbnd * sizeof(int));
  rrtmgpb_apply_cloud_mask(ncol, nlay, nbnd, ngpt, lims, cloud_mask, clouds->tau, clouds_sampled->tau);  // :104
  if (clouds->kind == RRTMGPB_2STR) {                                                                     // :108-118
    rrtmgpb_apply_cloud_mask(ncol, nlay, nbnd, ngpt, lims, cloud_mask, clouds->ssa, clouds_sampled->ssa);
    rrtmgpb_apply_cloud_mask(ncol, nlay, nbnd, ngpt, lims, cloud_mask, clouds->g, clouds_sampled->g);
  }
  return ok(errmsg);
}

// ------------------------------------------------------------------------------------------------
// ty_gas_concs: rte/frontend/gas-optics-template/mo_gas_concentrations.F90
// ------------------------------------------------------------------------------------------------
struct rrtmgpb_gas_concs_t {
  struct Conc { int nc = 0, nl = 0; Float* data = nullptr; };  // conc(nc,nl): (1,1), (1,nlay) or (ncol,nlay), backend memory
  std::vector<std::string> names;
  std::vector<Conc> concs;
  int ncol = 0, nlay = 0;
};

static std::string gc_lower_trim(const char* s) {  // lower_case(trim(gas)), :100,:366
  std::string r(s ? s : "");
  while (!r.empty() && r.back() == ' ') r.pop_back();
  size_t b = 0;
  while (b < r.size() && r[b] == ' ') ++b;
  r = r.substr(b);
  for (char& c : r) if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a');
  return r;
}
static std::string gc_trim(const char* s) {  // trim(gas), as the reference's messages print it
  std::string r(s ? s : "");
  while (!r.empty() && r.back() == ' ') r.pop_back();
  return r;
}
static int gc_find(const rrtmgpb_gas_concs_t* gc, const std::string& name) {  // find_gas, :577-590
  for (size_t i = 0; i < gc->names.size(); ++i) if (gc->names[i] == name) return (int)i;
  return -1;
}
// (re)allocate the storage of gas igas as conc(nc,nl) and fill it from w (backend memory) or from the scalar
static void gc_store(rrtmgpb_gas_concs_t* gc, int igas, int nc, int nl, const Float* w, Float scalar) {
  rrtmgpb_gas_concs_t::Conc& c = gc->concs[igas];
  if (c.data && (c.nc != nc || c.nl != nl)) { rrtmgpb_mem_free(c.data); c.data = nullptr; }  // :154-161
  const size_t n = (size_t)nc * nl;
  if (!c.data) c.data = static_cast<Float*>(rrtmgpb_mem_alloc(n * sizeof(Float)));
  c.nc = nc; c.nl = nl;
  if (w) rrtmgpb_mem_copy(c.data, w, n * sizeof(Float));
  else rrtmgpb_mem_to_backend(c.data, &scalar, sizeof(Float));
}

rrtmgpb_gas_concs_t* rrtmgpb_gc_init(int ngas, const char* const* gas_names, char* errmsg) {  // init(), :96-124
  std::vector<std::string> names;
  for (int i = 0; i < ngas; ++i) {
    const std::string n = gc_lower_trim(gas_names[i]);
    if (n.empty()) { fail(errmsg, "ty_gas_concs%init(): must provide non-empty gas names"); return nullptr; }
    for (const std::string& m : names)
      if (m == n) { fail(errmsg, "ty_gas_concs%init(): duplicate gas names aren't allowed"); return nullptr; }
    names.push_back(n);
  }
  rrtmgpb_gas_concs_t* gc = new rrtmgpb_gas_concs_t;
  gc->names = names;
  gc->concs.resize(names.size());
  ok(errmsg);
  return gc;
}

void rrtmgpb_gc_free(rrtmgpb_gas_concs_t* gc) {
  if (!gc) return;
  for (auto& c : gc->concs) rrtmgpb_mem_free(c.data);
  delete gc;
}

int rrtmgpb_gc_set_vmr_scalar(rrtmgpb_gas_concs_t* gc, const char* gas, Float w, char* errmsg) {  // :129-191
  if (w < 0 || w > 1) return fail(errmsg, "ty_gas_concs%set_vmr(): concentrations should be >= 0, <= 1");
  const int igas = gc_find(gc, gc_lower_trim(gas));
  if (igas < 0)
    return fail(errmsg, "ty_gas_concs%set_vmr(): trying to set " + gc_trim(gas) + " but name not provided at initialization");
  gc_store(gc, igas, 1, 1, nullptr, w);
  return ok(errmsg);
}

int rrtmgpb_gc_set_vmr_1d(rrtmgpb_gas_concs_t* gc, const char* gas, int nlay, const Float* w, char* errmsg) {  // :194-246
  if (g_check_values && rrtmgpb_any_vals_outside((size_t)nlay, w, nullptr, 0, 1))
    return fail(errmsg, "ty_gas_concs%set_vmr: concentrations should be >= 0, <= 1");
  if (gc->nlay > 0 && nlay != gc->nlay) return fail(errmsg, "ty_gas_concs%set_vmr: different dimension (nlay)");
  const int igas = gc_find(gc, gc_lower_trim(gas));
  if (igas < 0)
    return fail(errmsg, "ty_gas_concs%set_vmr(): trying to set " + gc_trim(gas) + " but name not provided at initialization");
  gc->nlay = nlay;
  gc_store(gc, igas, 1, nlay, w, 0);
  return ok(errmsg);
}

int rrtmgpb_gc_set_vmr_2d(rrtmgpb_gas_concs_t* gc, const char* gas, int ncol, int nlay, const Float* w, char* errmsg) {  // :249-305
  if (g_check_values && rrtmgpb_any_vals_outside((size_t)ncol * nlay, w, nullptr, 0, 1))
    return fail(errmsg, "ty_gas_concs%set_vmr: concentrations should be >= 0, <= 1");
  if (gc->ncol > 0 && ncol != gc->ncol) return fail(errmsg, "ty_gas_concs%set_vmr: different dimension (ncol)");
  if (gc->nlay > 0 && nlay != gc->nlay) return fail(errmsg, "ty_gas_concs%set_vmr: different dimension (nlay)");
  const int igas = gc_find(gc, gc_lower_trim(gas));
  if (igas < 0)
    return fail(errmsg, "ty_gas_concs%set_vmr(): trying to set " + gc_trim(gas) + " but name not provided at initialization");
  gc->ncol = ncol; gc->nlay = nlay;
  gc_store(gc, igas, ncol, nlay, w, 0);
  return ok(errmsg);
}

int rrtmgpb_gc_get_vmr(const rrtmgpb_gas_concs_t* gc, const char* gas, int ncol, int nlay, Float* array, char* errmsg) {  // :433-504
  const int igas = gc_find(gc, gc_lower_trim(gas));
  const std::string name = gc_trim(gas);
  std::string msg;
  if (igas < 0) msg = "ty_gas_concs%get_vmr; gas " + name + " not found";
  else if (!gc->concs[igas].data) msg = "ty_gas_concs%get_vmr; gas " + name + " concentration hasn't been set";
  if (gc->ncol > 0 && gc->ncol != ncol) msg = "ty_gas_concs%get_vmr; gas " + name + " array is wrong size (ncol)";
  if (gc->nlay > 0 && gc->nlay != nlay) msg = "ty_gas_concs%get_vmr; gas " + name + " array is wrong size (nlay)";
  if (!msg.empty()) return fail(errmsg, msg);
  const rrtmgpb_gas_concs_t::Conc& c = gc->concs[igas];
  rrtmgpb_gas_concs_get_vmr(ncol, nlay, c.nc, c.nl, c.data, array);
  return ok(errmsg);
}

// ------------------------------------------------------------------------------------------------
// ty_aerosol_optics_rrtmgp_merra (LUT): rrtmgp/frontend/mo_aerosol_optics_rrtmgp_merra.F90
// ------------------------------------------------------------------------------------------------
struct rrtmgpb_aerosol_optics_t {
  int nbnd, nval, nrh, nbin;
  std::vector<Float> band_lims_wvn_h;
  Float min_size, max_size;  // merra_aero_bin_lims(1,1), (2,nbin), :291-292
  Float *bin_lims, *aero_rh, *dust, *salt, *sulf, *bcar, *bcar_rh, *ocar, *ocar_rh;  // backend memory
};

// reshape(tbl, shape=(/nrh,nval,.../), order=(/2,1,.../)): swap the first two dimensions (:178-181)
static std::vector<Float> swap_first_two(const Float* src, int nval, int nrh, size_t nrest) {
  std::vector<Float> out((size_t)nval * nrh * nrest);
  for (size_t r = 0; r < nrest; ++r)
    for (int irh = 0; irh < nrh; ++irh)
      for (int iv = 0; iv < nval; ++iv)
        out[irh + (size_t)nrh * (iv + (size_t)nval * r)] = src[iv + (size_t)nval * (irh + (size_t)nrh * r)];
  return out;
}

rrtmgpb_aerosol_optics_t* rrtmgpb_aerosol_optics_load(const rrtmgpb_aerosol_lut* lut, char* errmsg) {
  if (!lut || !lut->aero_dust_tbl || !lut->aero_salt_tbl || !lut->aero_sulf_tbl || !lut->aero_bcar_tbl ||
      !lut->aero_bcar_rh_tbl || !lut->aero_ocar_tbl || !lut->aero_ocar_rh_tbl || !lut->merra_aero_bin_lims ||
      !lut->aero_rh) {
    fail(errmsg, "aerosol optics: no data has been initialized");
    return nullptr;
  }
  if (lut->nval != 3 || lut->nrh < 1 || lut->nbin < 1 || lut->nbnd < 1) {
    fail(errmsg, "aerosol_optics%load_lut(): array aero_salt_tbl isn't consistently sized");
    return nullptr;
  }
  rrtmgpb_aerosol_optics_t* ao = new rrtmgpb_aerosol_optics_t();
  ao->nbnd = lut->nbnd; ao->nval = lut->nval; ao->nrh = lut->nrh; ao->nbin = lut->nbin;
  if (lut->band_lims_wvn) ao->band_lims_wvn_h.assign(lut->band_lims_wvn, lut->band_lims_wvn + 2 * lut->nbnd);
  ao->min_size = lut->merra_aero_bin_lims[0];
  ao->max_size = lut->merra_aero_bin_lims[2 * (lut->nbin - 1) + 1];
  const int nval = lut->nval, nrh = lut->nrh, nbin = lut->nbin, nbnd = lut->nbnd;
  ao->bin_lims = upload(lut->merra_aero_bin_lims, 2 * (size_t)nbin);
  ao->aero_rh = upload(lut->aero_rh, (size_t)nrh);
  ao->dust = upload(lut->aero_dust_tbl, (size_t)nval * nbin * nbnd);
  ao->bcar = upload(lut->aero_bcar_tbl, (size_t)nval * nbnd);
  ao->ocar = upload(lut->aero_ocar_tbl, (size_t)nval * nbnd);
  {
    const std::vector<Float> salt = swap_first_two(lut->aero_salt_tbl, nval, nrh, (size_t)nbin * nbnd);
    const std::vector<Float> sulf = swap_first_two(lut->aero_sulf_tbl, nval, nrh, (size_t)nbnd);
    const std::vector<Float> bcrh = swap_first_two(lut->aero_bcar_rh_tbl, nval, nrh, (size_t)nbnd);
    const std::vector<Float> ocrh = swap_first_two(lut->aero_ocar_rh_tbl, nval, nrh, (size_t)nbnd);
    ao->salt = upload(salt.data(), salt.size());
    ao->sulf = upload(sulf.data(), sulf.size());
    ao->bcar_rh = upload(bcrh.data(), bcrh.size());
    ao->ocar_rh = upload(ocrh.data(), ocrh.size());
    rrtmgpb_sync();  // the host staging vectors die at the end of this scope
  }
  ok(errmsg);
  return ao;
}

void rrtmgpb_aerosol_optics_free(rrtmgpb_aerosol_optics_t* ao) {
  if (!ao) return;
  void* ptrs[] = {ao->bin_lims, ao->aero_rh, ao->dust, ao->salt, ao->sulf, ao->bcar, ao->bcar_rh, ao->ocar, ao->ocar_rh};
  for (void* p : ptrs) rrtmgpb_mem_free(p);
  delete ao;
}

int rrtmgpb_aerosol_optics(const rrtmgpb_aerosol_optics_t* ao, int ncol, int nlay, const int* aero_type,
                           const Float* aero_size, const Float* aero_mass, const Float* relhum,
                           rrtmgpb_optical_props* op, char* errmsg) {
  if (!ao) return fail(errmsg, "aerosol optics: no data has been initialized");  // :278-281
  const size_t ncl = (size_t)ncol * nlay;
  std::string msg;
  if (g_check_extents) {  // :297-310 (the 2-D inputs carry no extents across a C interface)
    if (op->ncol != ncol || op->nlay != nlay) msg = "aerosol optics: optical_props have wrong extents";
    if (!msg.empty()) return fail(errmsg, msg);
  }
  if (g_check_values) {  // :315-323
    if (op->nband != ao->nbnd) msg = "aerosol optics: optical properties don't have the same band structure";
    if (op->band_lims_wvn && !ao->band_lims_wvn_h.empty()) {
      rrtmgpb_optical_props self = *op;
      self.band_lims_wvn = ao->band_lims_wvn_h.data();
      self.nband = ao->nbnd;
      if (!bands_are_equal(&self, op)) msg = "aerosol optics: optical properties don't have the same band structure";
    }
    if (op->nband != op->ngpt) msg = "aerosol optics: optical properties must be requested by band not g-points";
    if (rrtmgpb_any_int_vals_outside(ncl, aero_type, 0, 7)) msg = "aerosol optics: aerosol type is out of bounds";
    if (!msg.empty()) return fail(errmsg, msg);
    Scratch<Bool> aeromsk(ncl);  // :343-347
    rrtmgpb_aerosol_mask(ncol, nlay, aero_type, aeromsk);
    if (rrtmgpb_any_vals_outside(ncl, aero_size, aeromsk, ao->min_size, ao->max_size))  // :352-357
      msg = "aerosol optics: requested aerosol size is out of bounds";
    if (rrtmgpb_any_vals_outside(ncl, relhum, aeromsk, 0, 1))
      msg = "aerosol optics: relative humidity fraction is out of bounds";
    if (!msg.empty()) return fail(errmsg, msg);
  }
  if (op->kind == RRTMGPB_NSTR) return fail(errmsg, "aerosol optics: n-stream calculations not yet supported");  // :419
  rrtmgpb_aerosol_optics_from_table(ncol, nlay, ao->nval, ao->nrh, ao->nbin, ao->nbnd, op->kind, aero_type, aero_size,
                                    aero_mass, relhum, ao->bin_lims, ao->aero_rh, ao->dust, ao->salt, ao->sulf,
                                    ao->bcar_rh, ao->bcar, ao->ocar_rh, ao->ocar, op->tau, op->ssa, op->g);
  return ok(errmsg);
}

}  // extern "C"
