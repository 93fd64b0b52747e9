// util_abi.cu - fills, broadband reductions, Planck helpers and the frontend-glue kernels.
//
// Reference interfaces replaced (extern mode):
//   rte/kernels/api/mo_rte_util_array.F90:23-79          zero_array_*D, set_to_scalar_*D
//   rte/kernels/api/mo_fluxes_broadband_kernels.F90:26-72 rte_sum_broadband, rte_net_broadband_*
//   rte/kernels/api/mo_gas_optics_utils.F90:7-36          rte_compute_Planck_source_1D/_2D
// and the frontend loops listed in include/rrtmgp_b200_ext.h (SURVEY.md section 8a').
// All pure-bandwidth: one coalesced pass, grid-stride, columns innermost.
#include "../kernels/fastmath.cuh"
#include "../kernels/elementwise.cuh"
#include "rte_kernels.h"
#include "rrtmgp_b200_ext.h"

using namespace rrtmgpb;

namespace {

void fill(size_t n, Float* array, Float v) {
  DevArg<Float> a(array, n, Dir::Out);
  Float* p = a;
  if (v == (Float)0) {
    RB_CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(Float), stream()));
    count_launch();
  } else {
    launch_elementwise(n, [=] __device__(size_t i) { p[i] = v; });
  }
}

// physical constants, rte/kernels/mo_gas_optics_constants.F90:10-38 (host copies; passed by value)
struct Constants {
  double boltzmann_k = 1.380649e-23, m_h2o = 0.018016, avogad = 6.02214076e23;
  double planck_h = 6.626075540e-34, lightspeed = 2.99792458e8;
  double m_dry = 0.028964, grav = 9.80665, cp_dry = 1004.64;
} g_const;

__device__ __forceinline__ Float B_nu(Float T, Float nu, Float h, Float c, Float kb) {
  // rte/kernels/mo_gas_optics_utils.F90:36-41
  const Float nu100 = nu * (Float)100;
  return (Float)100 * (Float)2 * h * (nu100 * nu100 * nu100) * (c * c) /
         (exp((h * c * nu * (Float)100) / (kb * T)) - (Float)1);
}

}  // namespace

extern "C" {

// ---------------- fills ----------------
void zero_array_1D(const int* ni, Float* array) { OpName op_name__(__func__); fill((size_t)*ni, array, 0); }
void zero_array_2D(const int* ni, const int* nj, Float* array) { OpName op_name__(__func__); fill((size_t)*ni * *nj, array, 0); }
void zero_array_3D(const int* ni, const int* nj, const int* nk, Float* array) {
  OpName op_name__(__func__);
  fill((size_t)*ni * *nj * *nk, array, 0);
}
void zero_array_4D(const int* ni, const int* nj, const int* nk, const int* nl, Float* array) {
  OpName op_name__(__func__);
  fill((size_t)*ni * *nj * *nk * *nl, array, 0);
}
void set_to_scalar_1D(const int* ni, Float* array, const Float* value) { OpName op_name__(__func__); fill((size_t)*ni, array, *value); }
void set_to_scalar_2D(const int* ni, const int* nj, Float* array, const Float* value) {
  OpName op_name__(__func__);
  fill((size_t)*ni * *nj, array, *value);
}
void set_to_scalar_3D(const int* ni, const int* nj, const int* nk, Float* array, const Float* value) {
  OpName op_name__(__func__);
  fill((size_t)*ni * *nj * *nk, array, *value);
}
void set_to_scalar_4D(const int* ni, const int* nj, const int* nk, const int* nl, Float* array,
                      const Float* value) {
  OpName op_name__(__func__);
  fill((size_t)*ni * *nj * *nk * *nl, array, *value);
}

// ---------------- broadband reductions ----------------
// One thread per (col,lev); sequential sum over g-points 1..ngpt - the reference's order
// (mo_fluxes_broadband_kernels.F90:45-58), so results are reproducible run to run.
void rte_sum_broadband(const int* ncol, const int* nlev, const int* ngpt, const Float* spectral_flux,
                       Float* broadband_flux) {
  OpName op_name__(__func__);
  const size_t n2 = (size_t)*ncol * *nlev;
  const int ng = *ngpt;
  DevArg<Float> in(spectral_flux, n2 * ng, Dir::In), out(broadband_flux, n2, Dir::Out);
  const Float* s = in; Float* b = out;
  launch_elementwise(n2, [=] __device__(size_t c) {
    Float acc = 0;
#pragma unroll 8
    for (int ig = 0; ig < ng; ++ig) acc = acc + s[c + n2 * ig];
    b[c] = acc;
  });
}

// ---------------- simple spectral model gas optics (ssm/mo_optics_ssm_kernels.F90:29-108) ----------------
// compute_tau :29-81: tau(icol,ilay,inu) = sum_igas layer_mass(igas,icol,ilay) * absorption_coeffs(igas,inu)
// [* play/pref when pref > 0].  Thread = cell looping over nu: the cell's ngas masses stay in registers (ngas <= 8;
// more gases re-read them), each tau plane element written once, coalesced.
void ssm_compute_tau_absorption(const int* ncol_, const int* nlay_, const int* nnu_, const int* ngas_,
                                const Float* absorption_coeffs, const Float* play, const Float* pref_,
                                const Float* layer_mass, Float* tau) {
  OpName op_name__(__func__);
  const int nnu = *nnu_, ngas = *ngas_;
  const size_t ncl = (size_t)*ncol_ * *nlay_;
  const Float pref = *pref_;
  DevArg<Float> a_k(absorption_coeffs, (size_t)ngas * nnu, Dir::In), a_p(play, ncl, Dir::In),
      a_m(layer_mass, ncl * ngas, Dir::In), a_t(tau, ncl * nnu, Dir::Out);
  const Float *k = a_k, *pl = a_p, *lm = a_m; Float* t = a_t;
  launch_elementwise(ncl, [=] __device__(size_t c) {
    constexpr int kMaxReg = 8;
    Float m[kMaxReg];
#pragma unroll
    for (int ig = 0; ig < kMaxReg; ++ig) m[ig] = ig < ngas ? lm[ig + (size_t)ngas * c] : (Float)0;
    const Float scale = pref > (Float)0 ? pl[c] : (Float)1;
    for (int inu = 0; inu < nnu; ++inu) {
      const Float* kk = k + (size_t)ngas * inu;
      Float s = 0;  // sum() of the array constructor: left to right from zero
#pragma unroll
      for (int ig = 0; ig < kMaxReg; ++ig)
        if (ig < ngas) s = s + m[ig] * __ldg(kk + ig);
      for (int ig = kMaxReg; ig < ngas; ++ig) s = s + lm[ig + (size_t)ngas * c] * __ldg(kk + ig);
      t[c + ncl * inu] = pref > (Float)0 ? s * scale / pref : s;   // :58-61, :71-72
    }
  });
}

// compute_layer_mass :83-106: layer_mass(igas,icol,ilay) = vmr * (mol_weights(igas)/m_dry) * |dp| / grav
void ssm_compute_layer_mass(const int* ncol_, const int* nlay_, const int* ngas_, const Float* vmr, const Float* plev,
                            const Float* mol_weights, const Float* m_dry_, Float* layer_mass) {
  OpName op_name__(__func__);
  const int ncol = *ncol_, ngas = *ngas_;
  const size_t ncl = (size_t)ncol * *nlay_;
  const Float m_dry = *m_dry_, grav = (Float)g_const.grav;
  DevArg<Float> a_v(vmr, ncl * ngas, Dir::In), a_p(plev, ncl + ncol, Dir::In), a_w(mol_weights, ngas, Dir::In),
      a_o(layer_mass, ncl * ngas, Dir::Out);
  const Float *v = a_v, *pp = a_p, *w = a_w; Float* o = a_o;
  launch_elementwise(ncl * ngas, [=] __device__(size_t i) {
    const int ig = (int)(i % ngas);
    const size_t c = i / ngas;
    o[i] = v[i] * (w[ig] / m_dry) * fabs(pp[c + ncol] - pp[c]) / grav;
  });
}

// ---------------- by-band reductions (rte/extensions/mo_fluxes_byband.F90:159-218) ----------------
// One thread per (col,lev,band); sequential sum over the band's g-points in the reference's order.
void rte_sum_byband(const int* ncol, const int* nlev, const int* ngpt, const int* nbnd, const int* band_lims,
                    const Float* spectral_flux, Float* byband_flux) {
  OpName op_name__(__func__);
  const size_t n2 = (size_t)*ncol * *nlev;
  const int nb = *nbnd;
  DevArg<int> lims(band_lims, 2 * (size_t)nb, Dir::In);
  DevArg<Float> in(spectral_flux, n2 * *ngpt, Dir::In), out(byband_flux, n2 * nb, Dir::Out);
  const int* bl = lims; const Float* s = in; Float* b = out;
  launch_elementwise(n2 * nb, [=] __device__(size_t i) {
    const size_t c = i % n2;
    const int ib = (int)(i / n2);
    const int g0 = bl[2 * ib] - 1, g1 = bl[2 * ib + 1] - 1;
    Float acc = s[c + n2 * g0];                                   // :170
    for (int ig = g0 + 1; ig <= g1; ++ig) acc = acc + s[c + n2 * ig];  // :171-174
    b[i] = acc;
  });
}

void rte_net_byband_full(const int* ncol, const int* nlev, const int* ngpt, const int* nbnd, const int* band_lims,
                         const Float* spectral_flux_dn, const Float* spectral_flux_up, Float* byband_flux_net) {
  OpName op_name__(__func__);
  const size_t n2 = (size_t)*ncol * *nlev;
  const int nb = *nbnd;
  DevArg<int> lims(band_lims, 2 * (size_t)nb, Dir::In);
  DevArg<Float> dn(spectral_flux_dn, n2 * *ngpt, Dir::In), up(spectral_flux_up, n2 * *ngpt, Dir::In),
      out(byband_flux_net, n2 * nb, Dir::Out);
  const int* bl = lims; const Float *d = dn, *u = up; Float* b = out;
  launch_elementwise(n2 * nb, [=] __device__(size_t i) {
    const size_t c = i % n2;
    const int ib = (int)(i / n2);
    const int g0 = bl[2 * ib] - 1, g1 = bl[2 * ib + 1] - 1;
    Float acc = d[c + n2 * g0] - u[c + n2 * g0];                  // :196-198
    for (int ig = g0 + 1; ig <= g1; ++ig) acc = acc + d[c + n2 * ig] - u[c + n2 * ig];  // :199-203, left to right
    b[i] = acc;
  });
}

void net_byband_precalc(const int* ncol, const int* nlev, const int* nbnd, const Float* byband_flux_dn,
                        const Float* byband_flux_up, Float* byband_flux_net) {
  OpName op_name__(__func__);
  const size_t n = (size_t)*ncol * *nlev * *nbnd;
  DevArg<Float> dn(byband_flux_dn, n, Dir::In), up(byband_flux_up, n, Dir::In), out(byband_flux_net, n, Dir::Out);
  const Float *d = dn, *u = up; Float* b = out;
  launch_elementwise(n, [=] __device__(size_t i) { b[i] = d[i] - u[i]; });  // :216
}

// ---------------- heating rates (rte/extensions/mo_heating_rates.F90:34-117) ----------------
void rrtmgpb_heating_rate(int ncol, int nlay, const Float* flux_up, const Float* flux_dn, const Float* p_lev,
                          Float* heating_rate) {
  OpName op_name__(__func__);
  const size_t nc = ncol, ncl = nc * nlay, nclp = nc * (nlay + 1);
  DevArg<Float> fu(flux_up, nclp, Dir::In), fd(flux_dn, nclp, Dir::In), pl(p_lev, nclp, Dir::In), hr(heating_rate, ncl, Dir::Out);
  const Float *u = fu, *d = fd, *pp = pl; Float* h = hr;
  const Float grav = (Float)g_const.grav, cp_dry = (Float)g_const.cp_dry;
  launch_elementwise(ncl, [=] __device__(size_t i) {  // :56-62; level below layer i is i + ncol
    h[i] = (u[i + nc] - u[i] - d[i + nc] + d[i]) * grav / (cp_dry * (pp[i + nc] - pp[i]));
  });
}

void rrtmgpb_heating_rate_solar_varmu0(int ncol, int nlay, const Float* flux_up, const Float* flux_dn,
                                       const Float* flux_dir, const Float* p_lev, const Float* mu0, Float* heating_rate) {
  rrtmgpb_heating_rate(ncol, nlay, flux_up, flux_dn, p_lev, heating_rate);  // :81
  OpName op_name__(__func__);
  const size_t nc = ncol, ncl = nc * nlay, nclp = nc * (nlay + 1);
  DevArg<Float> fu(flux_up, nclp, Dir::In), fd(flux_dn, nclp, Dir::In), fr(flux_dir, nclp, Dir::In), pl(p_lev, nclp, Dir::In),
      m0(mu0, ncl, Dir::In), hr(heating_rate, ncl, Dir::InOut);
  const Float *u = fu, *d = fd, *r = fr, *pp = pl, *mu = m0; Float* h = hr;
  const Float grav = (Float)g_const.grav, cp_dry = (Float)g_const.cp_dry;
  const Float eps = (Float)RB_EPS;
  // one thread per column: the serial searches of :85-104 (any sun below the horizon?  orientation from the last
  // layer of column 1..ncol is a global property in the reference: any(mu0(:,nlay) < eps))
  int* flags = static_cast<int*>(dev_alloc(2 * sizeof(int)));
  RB_CUDA_CHECK(cudaMemsetAsync(flags, 0, 2 * sizeof(int), stream()));
  launch_elementwise(ncl, [=] __device__(size_t i) {
    if (mu[i] < eps) {
      flags[0] = 1;                                   // any_vals_less_than(mu0, epsilon(mu0)), :85-86
      if (i >= ncl - nc) flags[1] = 1;                // any_vals_less_than(mu0(:,nlay), epsilon(mu0)), :100
    }
  });
  launch_elementwise(nc, [=] __device__(size_t icol) {
    if (!flags[0]) return;
    // minloc / maxloc of mu0 over the layers where mu0 > 0 (first occurrence), 1-based; 0 if the mask is empty
    int loc = 0;
    Float best = 0;
    for (int l = 0; l < nlay; ++l) {
      const Float v = mu[icol + nc * l];
      if (!(v > (Float)0)) continue;
      if (loc == 0 || (flags[1] ? v < best : v > best)) { loc = l + 1; best = v; }
    }
    const int ilay = flags[1] ? loc + 1 : loc - 1;      // :102, :104
    if (ilay > 1 && ilay < nlay) {                      // :108
      const size_t i = icol + nc * (size_t)(ilay - 1);
      h[i] = (u[i + nc] - u[i] - d[i + nc] + d[i] + r[i + nc] - r[i]) * grav / (cp_dry * (pp[i + nc] - pp[i]));  // :112-115
    }
  });
  dev_free(flags);
}

void rte_net_broadband_full(const int* ncol, const int* nlev, const int* ngpt, const Float* spectral_flux_dn,
                            const Float* spectral_flux_up, Float* broadband_flux_net) {
  OpName op_name__(__func__);
  const size_t n2 = (size_t)*ncol * *nlev;
  const int ng = *ngpt;
  DevArg<Float> dn(spectral_flux_dn, n2 * ng, Dir::In), up(spectral_flux_up, n2 * ng, Dir::In),
      out(broadband_flux_net, n2, Dir::Out);
  const Float *d = dn, *u = up; Float* b = out;
  launch_elementwise(n2, [=] __device__(size_t c) {
    Float acc = d[c] - u[c];
#pragma unroll 4
    for (int ig = 1; ig < ng; ++ig) acc = acc + (d[c + n2 * ig] - u[c + n2 * ig]);
    b[c] = acc;
  });
}

void rte_net_broadband_precalc(const int* ncol, const int* nlev, const Float* flux_dn, const Float* flux_up,
                               Float* broadband_flux_net) {
  OpName op_name__(__func__);
  const size_t n2 = (size_t)*ncol * *nlev;
  DevArg<Float> dn(flux_dn, n2, Dir::In), up(flux_up, n2, Dir::In), out(broadband_flux_net, n2, Dir::Out);
  const Float *d = dn, *u = up; Float* b = out;
  launch_elementwise(n2, [=] __device__(size_t c) { b[c] = d[c] - u[c]; });
}

// ---------------- Planck helpers (used by the SSM gas optics) ----------------
void rte_compute_Planck_source_2D(const int* ncol, const int* nlay, const int* nnu, const Float* nus,
                                  const Float* dnus, const Float* T, Float* source) {
  OpName op_name__(__func__);
  const size_t n2 = (size_t)*ncol * *nlay;
  const int nn = *nnu;
  DevArg<Float> nu(nus, nn, Dir::In), dnu(dnus, nn, Dir::In), t(T, n2, Dir::In), out(source, n2 * nn, Dir::Out);
  const Float *pn = nu, *pd = dnu, *pt = t; Float* ps = out;
  const Float h = (Float)g_const.planck_h, c = (Float)g_const.lightspeed, kb = (Float)g_const.boltzmann_k;
  launch_elementwise(n2 * nn, [=] __device__(size_t i) {
    const size_t cell = i % n2; const int inu = (int)(i / n2);
    ps[i] = B_nu(pt[cell], pn[inu], h, c, kb) * pd[inu];
  });
}
void rte_compute_Planck_source_1D(const int* ncol, const int* nnu, const Float* nus, const Float* dnus,
                                  const Float* T, Float* source) {
  OpName op_name__(__func__);
  const int one = 1;
  rte_compute_Planck_source_2D(ncol, &one, nnu, nus, dnus, T, source);
}

// ---------------- extension: constants ----------------
void rrtmgpb_init_constants(const Float* gravity, const Float* mol_weight_dry_air,
                            const Float* heat_capacity_dry_air) {
  OpName op_name__(__func__);
  if (gravity) g_const.grav = *gravity;
  if (mol_weight_dry_air) g_const.m_dry = *mol_weight_dry_air;
  if (heat_capacity_dry_air) g_const.cp_dry = *heat_capacity_dry_air;
  fused_set_constants(g_const.grav, g_const.m_dry);
}

// ---------------- extension: frontend glue ----------------
void rrtmgpb_get_col_dry(int ncol, int nlay, const Float* vmr_h2o, const Float* plev, Float* col_dry) {
  OpName op_name__(__func__);
  const size_t ncl = (size_t)ncol * nlay;
  DevArg<Float> v(vmr_h2o, ncl, Dir::In), p(plev, ncl + ncol, Dir::In), o(col_dry, ncl, Dir::Out);
  const Float *pv = v, *pp = p; Float* po = o;
  const Float m_dry = (Float)g_const.m_dry, m_h2o = (Float)g_const.m_h2o, avogad = (Float)g_const.avogad,
              grav = (Float)g_const.grav;
  launch_elementwise(ncl, [=] __device__(size_t k) {  // mo_gas_optics_utils.F90:143-150
    const Float delta_plev = fabs(pp[k] - pp[k + ncol]);
    const Float fact = (Float)1 / ((Float)1 + pv[k]);
    const Float m_air = (m_dry + m_h2o * pv[k]) * fact;
    po[k] = (Float)10 * delta_plev * avogad * fact / ((Float)1000 * m_air * (Float)100 * grav);
  });
}

void rrtmgpb_get_layer_mass(int ncol, int nlay, int ngas, const Float* vmr, const Float* plev,
                            const Float* mol_weights, Float m_dry, Float* layer_mass) {
  OpName op_name__(__func__);
  const size_t ncl = (size_t)ncol * nlay, n = ncl * ngas;
  DevArg<Float> v(vmr, n, Dir::In), p(plev, ncl + ncol, Dir::In), w(mol_weights, ngas, Dir::In),
      o(layer_mass, n, Dir::Out);
  const Float *pv = v, *pp = p, *pw = w; Float* po = o;
  const Float grav = (Float)g_const.grav;
  launch_elementwise(n, [=] __device__(size_t k) {  // mo_gas_optics_utils.F90:114-123
    const int igas = (int)(k % ngas); const size_t cell = k / ngas;
    po[k] = pv[k] * (pw[igas] / m_dry) * fabs(pp[cell + ncol] - pp[cell]) / grav;
  });
}

void rrtmgpb_col_gas_from_vmr(int ncol, int nlay, int ngas, const Float* vmr, const Float* col_dry,
                              Float* col_gas) {
  OpName op_name__(__func__);
  const size_t ncl = (size_t)ncol * nlay;
  DevArg<Float> v(vmr, ncl * ngas, Dir::In), cd(col_dry, ncl, Dir::In), o(col_gas, ncl * (ngas + 1), Dir::Out);
  const Float *pv = v, *pc = cd; Float* po = o;
  launch_elementwise(ncl * (ngas + 1), [=] __device__(size_t k) {
    const size_t c = k % ncl;
    po[k] = (k < ncl) ? pc[c] : pv[k - ncl] * pc[c];
  });
}

void rrtmgpb_combine_abs_and_rayleigh(int ncol, int nlay, int ngpt, int kind, const Float* tau_abs,
                                      const Float* tau_rayleigh, Float* tau, Float* ssa, Float* g) {
  OpName op_name__(__func__);
  const size_t n = (size_t)ncol * nlay * ngpt;
  // tau may alias tau_abs: stage once and treat as in/out in that case
  const bool alias = (tau == tau_abs);
  DevArg<Float> ta(tau_abs, n, Dir::In, !alias), tr(tau_rayleigh, n, Dir::In);
  DevArg<Float> t(tau, n, alias ? Dir::InOut : Dir::Out);
  DevArg<Float> s(ssa, n, Dir::Out, kind == 2), gg(g, n, Dir::Out, kind == 2);
  const Float* pa = alias ? t.get() : ta.get();
  const Float* pr = tr; Float* pt = t; Float* ps = s; Float* pg = gg;
  if (kind == 1) {
    launch_elementwise(n, [=] __device__(size_t i) { pt[i] = pa[i] + pr[i]; });
  } else {
    launch_elementwise(n, [=] __device__(size_t i) {  // mo_gas_optics_rrtmgp.F90:1986-2002
      const Float r = pr[i];
      const Float tt = pa[i] + r;
      ps[i] = (tt > (Float)2 * (Float)RB_TINY) ? r / tt : (Float)0;
      pt[i] = tt;
      pg[i] = (Float)0;
    });
  }
}

void rrtmgpb_interpolate_tlev(int ncol, int nlay, const Float* play, const Float* plev, const Float* tlay,
                              Float* tlev) {
  OpName op_name__(__func__);
  const size_t ncl = (size_t)ncol * nlay, nclp = ncl + ncol;
  DevArg<Float> pl(play, ncl, Dir::In), pv(plev, nclp, Dir::In), tl(tlay, ncl, Dir::In), o(tlev, nclp, Dir::Out);
  const Float *play_ = pl, *plev_ = pv, *tlay_ = tl; Float* tlev_ = o;
  launch_elementwise(nclp, [=] __device__(size_t k) {  // mo_gas_optics_rrtmgp.F90:893-911
    const int l = (int)(k / ncol); const size_t i = k % ncol;
#define A2(a, ll) a[i + (size_t)ncol * (size_t)(ll)]
    Float v;
    if (l == 0) {
      v = A2(tlay_, 0) + (A2(plev_, 0) - A2(play_, 0)) * (A2(tlay_, 1) - A2(tlay_, 0)) / (A2(play_, 1) - A2(play_, 0));
    } else if (l == nlay) {
      v = A2(tlay_, nlay - 1) + (A2(plev_, nlay) - A2(play_, nlay - 1)) * (A2(tlay_, nlay - 1) - A2(tlay_, nlay - 2)) /
                                    (A2(play_, nlay - 1) - A2(play_, nlay - 2));
    } else {
      v = (A2(play_, l - 1) * A2(tlay_, l - 1) * (A2(plev_, l) - A2(play_, l)) +
           A2(play_, l) * A2(tlay_, l) * (A2(play_, l - 1) - A2(plev_, l))) /
          (A2(plev_, l) * (A2(play_, l - 1) - A2(play_, l)));
    }
#undef A2
    tlev_[k] = v;
  });
}

void rrtmgpb_broadcast_by_gpt(int ncol, int ngpt, const Float* per_gpt, Float* out) {
  OpName op_name__(__func__);
  const size_t n = (size_t)ncol * ngpt;
  DevArg<Float> in(per_gpt, ngpt, Dir::In), o(out, n, Dir::Out);
  const Float* pi = in; Float* po = o;
  launch_elementwise(n, [=] __device__(size_t k) { po[k] = pi[k / ncol]; });
}

void rrtmgpb_broadcast_by_lay(int ncol, int nlay, const Float* per_col, Float* out) {
  OpName op_name__(__func__);
  const size_t n = (size_t)ncol * nlay;
  DevArg<Float> in(per_col, ncol, Dir::In), o(out, n, Dir::Out);
  const Float* pi = in; Float* po = o;
  launch_elementwise(n, [=] __device__(size_t k) { po[k] = pi[k % ncol]; });
}

void rrtmgpb_gas_concs_get_vmr(int ncol, int nlay, int nc_conc, int nl_conc, const Float* conc, Float* array) {
  OpName op_name__(__func__);
  const size_t n = (size_t)ncol * nlay;
  DevArg<Float> in(conc, (size_t)nc_conc * nl_conc, Dir::In), o(array, n, Dir::Out);
  const Float* pi = in; Float* po = o;
  // mo_gas_concentrations.F90:464-501: stored as 2D (ncol,nlay), 1D (1,nlay) or scalar (1,1)
  if (nc_conc > 1) launch_elementwise(n, [=] __device__(size_t k) { po[k] = pi[k]; });
  else if (nl_conc > 1) launch_elementwise(n, [=] __device__(size_t k) { po[k] = pi[k / ncol]; });
  else launch_elementwise(n, [=] __device__(size_t k) { po[k] = pi[0]; });
}

void rrtmgpb_compute_optimal_angles(int ncol, int nlay, int ngpt, int nband, const int* band_lims_gpt, const Float* tau,
                                    const Float* optimal_angle_fit, Float* optimal_angles) {
  OpName op_name__(__func__);
  const size_t n = (size_t)ncol * ngpt, ncl = (size_t)ncol * nlay;
  DevArg<int> lims(band_lims_gpt, 2 * (size_t)nband, Dir::In);
  DevArg<Float> t(tau, ncl * ngpt, Dir::In), fit(optimal_angle_fit, 2 * (size_t)nband, Dir::In), o(optimal_angles, n, Dir::Out);
  const int* pl = lims; const Float* pt = t; const Float* pf = fit; Float* po = o;
  // thread = (column, g-point), columns fastest: every layer's read is one coalesced row of the tau plane
  launch_elementwise(n, [=] __device__(size_t k) {
    const int g = (int)(k / ncol) + 1; const size_t i = k % ncol;
    int b = 0;
    while (b < nband - 1 && g > pl[2 * b + 1]) ++b;            // convert_gpt2band
    const Float* tg = pt + i + ncl * (size_t)(g - 1);
    Float tsum = 0;
    for (int l = 0; l < nlay; ++l) tsum = tsum + tg[(size_t)ncol * l];  // mo_gas_optics_rrtmgp.F90:1552-1554, layer order
    const Float trans_total = exp(-tsum);                               // :1555
    po[k] = pf[2 * b] * trans_total + pf[2 * b + 1];                    // :1559-1560
  });
}

}  // extern "C"
namespace {
// arr_in(nband,ncol) -> arr_out(ncol,ngpt): a block stages the bands of 32 columns (one contiguous run of arr_in) in shared
// memory, then every warp writes whole 256-byte rows of arr_out (lane = column).  The straightforward mapping read
// arr_in with a stride of nband words per lane: 0.10 ms per call at 65,536 x 256 on B200, 5x its roofline.
__global__ void __launch_bounds__(256) expand_transpose_kernel(int ncol, int nband, int ngpt, const int* __restrict__ lims,
                                                               const Float* __restrict__ in, Float* __restrict__ out) {
  extern __shared__ Float et_tile[];  // [nband][33]
  const int col0 = blockIdx.x * 32;
  const int ncols = min(32, ncol - col0);
  const Float* src = in + (size_t)nband * col0;
  for (int k = threadIdx.x; k < nband * ncols; k += blockDim.x) et_tile[(k % nband) * 33 + k / nband] = src[k];
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane >= ncols) return;
  for (int g = w + 1; g <= ngpt; g += nw) {  // mo_rte_lw.F90:490-500
    int b = 0;
    while (b < nband - 1 && g > lims[2 * b + 1]) ++b;
    if (g >= lims[2 * b] && g <= lims[2 * b + 1]) out[(size_t)(col0 + lane) + (size_t)ncol * (g - 1)] = et_tile[b * 33 + lane];
  }
}
}  // namespace
extern "C" {

void rrtmgpb_expand_and_transpose(int ncol, int nband, int ngpt, const int* band_lims_gpt,
                                  const Float* arr_in, Float* arr_out) {
  OpName op_name__(__func__);
  const size_t n = (size_t)ncol * ngpt;
  if (n == 0) return;
  DevArg<int> lims(band_lims_gpt, 2 * (size_t)nband, Dir::In);
  DevArg<Float> in(arr_in, (size_t)nband * ncol, Dir::In), o(arr_out, n, Dir::Out);
  const size_t smem = (size_t)nband * 33 * sizeof(Float);
  if (smem > 160 * 1024) {  // (hundreds of bands: the plain mapping)
    const int* pl = lims; const Float* pi = in; Float* po = o;
    launch_elementwise(n, [=] __device__(size_t k) {
      const int g = (int)(k / ncol) + 1; const size_t i = k % ncol;
      int b = 0;
      while (b < nband - 1 && g > pl[2 * b + 1]) ++b;
      if (g >= pl[2 * b] && g <= pl[2 * b + 1]) po[k] = pi[(size_t)b + (size_t)nband * i];
    });
    return;
  }
  KernelTimer timer("rrtmgpb_expand_and_transpose");
  if (smem > 48 * 1024)
    RB_CUDA_CHECK(cudaFuncSetAttribute(expand_transpose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  expand_transpose_kernel<<<ceil_div(ncol, 32), 256, smem, stream()>>>(ncol, nband, ngpt, lims, in, o);
  RB_LAUNCH_CHECK();
}

void rrtmgpb_cloud_masks(int ncol, int nlay, const Float* clwp, const Float* ciwp, Bool* liqmsk, Bool* icemsk) {
  OpName op_name__(__func__);
  const size_t n = (size_t)ncol * nlay;
  DevArg<Float> l(clwp, n, Dir::In), ic(ciwp, n, Dir::In);
  DevArg<Bool> lm(liqmsk, n, Dir::Out), im(icemsk, n, Dir::Out);
  const Float *pl = l, *pi = ic; Bool *plm = lm, *pim = im;
  launch_elementwise(n, [=] __device__(size_t k) { plm[k] = pl[k] > (Float)0; pim[k] = pi[k] > (Float)0; });
}

void rrtmgpb_cloud_combine(int ncol, int nlay, int ngpt, int kind, const Float* ltau, const Float* ltaussa,
                           const Float* ltaussag, const Float* itau, const Float* itaussa,
                           const Float* itaussag, Float* tau, Float* ssa, Float* g) {
  OpName op_name__(__func__);
  const size_t n = (size_t)ncol * nlay * ngpt;
  DevArg<Float> lt(ltau, n, Dir::In), lts(ltaussa, n, Dir::In), ltg(ltaussag, n, Dir::In, kind == 2);
  DevArg<Float> it(itau, n, Dir::In), its(itaussa, n, Dir::In), itg(itaussag, n, Dir::In, kind == 2);
  DevArg<Float> t(tau, n, Dir::Out), s(ssa, n, Dir::Out, kind == 2), gg(g, n, Dir::Out, kind == 2);
  const Float *a = lt, *b = lts, *c = ltg, *d = it, *e = its, *f = itg;
  Float *pt = t, *ps = s, *pg = gg;
  if (kind == 1) {
    launch_elementwise(n, [=] __device__(size_t i) { pt[i] = (a[i] - b[i]) + (d[i] - e[i]); });
  } else {
    launch_elementwise(n, [=] __device__(size_t i) {  // mo_cloud_optics_rrtmgp.F90:412-422
      const Float tt = a[i] + d[i];
      const Float ts = b[i] + e[i];
      pg[i] = (c[i] + f[i]) / fmax((Float)RB_EPS, ts);
      ps[i] = ts / fmax((Float)RB_EPS, tt);
      pt[i] = tt;
    });
  }
}

// Whole cloud_optics() body in one pass (mo_cloud_optics_rrtmgp.F90:334-341 masks, :373/:380 table lookups through
// mo_cloud_optics_rrtmgp_kernels.F90:24-65, :399-424 combination): thread = cell looping over bands; the six
// (ncol,nlay,nbnd) intermediates ltau..itaussag are never materialised.
struct CloudFusedParams {
  int ncol, nlay, nbnd, kind;
  const Float *clwp, *ciwp, *reliq, *dgice;
  int liq_nsteps, ice_nsteps;
  Float liq_step, liq_offset, ice_step, ice_offset;
  const Float *extliq, *ssaliq, *asyliq, *extice, *ssaice, *asyice;
  Float *tau, *ssa, *g;
  int delta_scale;  // 2-stream only: apply delta_scale_2str_k (mo_optical_props_kernels.F90:89-93) before the store
};
__global__ void __launch_bounds__(256) cloud_optics_fused_kernel(const CloudFusedParams p) {
  const size_t ncl = (size_t)p.ncol * p.nlay;
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncl) return;
  const Float lwp = p.clwp[c], iwp = p.ciwp[c];
  const bool lm = lwp > (Float)0, im = iwp > (Float)0;  // :334-341
  int li = 1, ii = 1;
  Float lf = 0, fi = 0;
  if (lm) {
    const Float re = p.reliq[c];
    li = min((int)floor((re - p.liq_offset) / p.liq_step) + 1, p.liq_nsteps - 1);  // kernels :46
    lf = (re - p.liq_offset) / p.liq_step - (Float)(li - 1);                        // kernels :47
  }
  if (im) {
    const Float de = p.dgice[c];
    ii = min((int)floor((de - p.ice_offset) / p.ice_step) + 1, p.ice_nsteps - 1);
    fi = (de - p.ice_offset) / p.ice_step - (Float)(ii - 1);
  }
  for (int b = 0; b < p.nbnd; ++b) {
    Float lt = 0, lts = 0, ltg = 0, it = 0, its = 0, itg = 0;
    if (lm) {  // kernels :52-60
      const Float* tt = p.extliq + (size_t)p.liq_nsteps * b + (li - 1);
      const Float* st = p.ssaliq + (size_t)p.liq_nsteps * b + (li - 1);
      const Float* at = p.asyliq + (size_t)p.liq_nsteps * b + (li - 1);
      lt = lwp * (__ldg(tt) + lf * (__ldg(tt + 1) - __ldg(tt)));
      lts = lt * (__ldg(st) + lf * (__ldg(st + 1) - __ldg(st)));
      ltg = lts * (__ldg(at) + lf * (__ldg(at + 1) - __ldg(at)));
    }
    if (im) {
      const Float* tt = p.extice + (size_t)p.ice_nsteps * b + (ii - 1);
      const Float* st = p.ssaice + (size_t)p.ice_nsteps * b + (ii - 1);
      const Float* at = p.asyice + (size_t)p.ice_nsteps * b + (ii - 1);
      it = iwp * (__ldg(tt) + fi * (__ldg(tt + 1) - __ldg(tt)));
      its = it * (__ldg(st) + fi * (__ldg(st + 1) - __ldg(st)));
      itg = its * (__ldg(at) + fi * (__ldg(at + 1) - __ldg(at)));
    }
    const size_t o = c + ncl * b;
    if (p.kind == 1) {
      p.tau[o] = (lt - lts) + (it - its);  // :402-405
    } else {                               // :412-422
      const Float tt_ = lt + it;
      const Float ts = lts + its;
      // rb_div (kernels/fastmath.cuh): branch-free, <= 1 ulp; the denominators are >= epsilon resp. 3*tiny, finite, normal.
      // The compiler's IEEE division sequence made this HBM-bound kernel compute-bound once the delta scaling was added.
      Float gq = rb_div(ltg + itg, fmax((Float)RB_EPS, ts));
      Float sq = rb_div(ts, fmax((Float)RB_EPS, tt_));
      Float tq = tt_;
      if (p.delta_scale) {  // clouds%delta_scale() (rrtmgp_allsky.F90:352) on the values just computed
        const Float eps3 = (Float)3.0 * (Float)RB_TINY;
        const Float fi = gq * gq;
        const Float wf = sq * fi;
        tq = ((Float)1 - wf) * tq;
        sq = rb_div(sq - wf, fmax(eps3, ((Float)1 - wf)));
        gq = rb_div(gq - fi, fmax(eps3, ((Float)1 - fi)));
      }
      p.g[o] = gq;
      p.ssa[o] = sq;
      p.tau[o] = tq;
    }
  }
}

void rrtmgpb_cloud_optics_from_tables(int ncol, int nlay, int nbnd, int kind, const Float* clwp, const Float* ciwp,
                                      const Float* reliq, const Float* dgice, int liq_nsteps, Float liq_step_size,
                                      Float liq_offset, const Float* extliq, const Float* ssaliq, const Float* asyliq,
                                      int ice_nsteps, Float ice_step_size, Float ice_offset, const Float* extice,
                                      const Float* ssaice, const Float* asyice, Float* tau, Float* ssa, Float* g) {
  rrtmgpb_cloud_optics_from_tables_ds(ncol, nlay, nbnd, kind, clwp, ciwp, reliq, dgice, liq_nsteps, liq_step_size, liq_offset,
                                      extliq, ssaliq, asyliq, ice_nsteps, ice_step_size, ice_offset, extice, ssaice, asyice,
                                      tau, ssa, g, 0);
}

void rrtmgpb_cloud_optics_from_tables_ds(int ncol, int nlay, int nbnd, int kind, const Float* clwp, const Float* ciwp,
                                         const Float* reliq, const Float* dgice, int liq_nsteps, Float liq_step_size,
                                         Float liq_offset, const Float* extliq, const Float* ssaliq, const Float* asyliq,
                                         int ice_nsteps, Float ice_step_size, Float ice_offset, const Float* extice,
                                         const Float* ssaice, const Float* asyice, Float* tau, Float* ssa, Float* g,
                                         int delta_scale) {
  const size_t ncl = (size_t)ncol * nlay, n = ncl * nbnd;
  const size_t nl = (size_t)liq_nsteps * nbnd, ni = (size_t)ice_nsteps * nbnd;
  DevArg<Float> a_lwp(clwp, ncl, Dir::In), a_iwp(ciwp, ncl, Dir::In), a_re(reliq, ncl, Dir::In), a_de(dgice, ncl, Dir::In);
  DevArg<Float> a_el(extliq, nl, Dir::In), a_sl(ssaliq, nl, Dir::In), a_al(asyliq, nl, Dir::In);
  DevArg<Float> a_ei(extice, ni, Dir::In), a_si(ssaice, ni, Dir::In), a_ai(asyice, ni, Dir::In);
  DevArg<Float> o_t(tau, n, Dir::Out), o_s(ssa, n, Dir::Out, kind == 2), o_g(g, n, Dir::Out, kind == 2);
  CloudFusedParams p;
  p.ncol = ncol; p.nlay = nlay; p.nbnd = nbnd; p.kind = kind;
  p.clwp = a_lwp; p.ciwp = a_iwp; p.reliq = a_re; p.dgice = a_de;
  p.liq_nsteps = liq_nsteps; p.ice_nsteps = ice_nsteps; p.liq_step = liq_step_size; p.liq_offset = liq_offset;
  p.ice_step = ice_step_size; p.ice_offset = ice_offset;
  p.extliq = a_el; p.ssaliq = a_sl; p.asyliq = a_al; p.extice = a_ei; p.ssaice = a_si; p.asyice = a_ai;
  p.tau = o_t; p.ssa = o_s; p.g = o_g;
  p.delta_scale = (delta_scale && kind == 2) ? 1 : 0;
  KernelTimer timer("cloud_optics_fused");
  cloud_optics_fused_kernel<<<ceil_div((long long)ncl, 256), 256, 0, stream()>>>(p);
  RB_LAUNCH_CHECK();
}

static int any_flag(size_t n, const Float* array, const Bool* mask, Float lo, Float hi, bool use_hi) {
  DevArg<Float> a(array, n, Dir::In);
  DevArg<Bool> m(mask, n, Dir::In, mask != nullptr);
  int* flag = static_cast<int*>(dev_alloc(sizeof(int)));
  RB_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int), stream()));
  const Float* pa = a; const Bool* pm = mask ? m.get() : nullptr;
  launch_elementwise(n, [=] __device__(size_t i) {
    if (pm && !pm[i]) return;
    const Float v = pa[i];
    if (v < lo || (use_hi && v > hi)) *flag = 1;
  });
  int h = 0;
  RB_CUDA_CHECK(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, stream()));
  RB_CUDA_CHECK(cudaStreamSynchronize(stream()));
  dev_free(flag);
  return h;
}
int rrtmgpb_any_vals_less_than(size_t n, const Float* array, const Bool* mask, Float check_value) {
  OpName op_name__(__func__);
  return any_flag(n, array, mask, check_value, 0, false);
}
int rrtmgpb_any_vals_outside(size_t n, const Float* array, const Bool* mask, Float checkMin, Float checkMax) {
  OpName op_name__(__func__);
  return any_flag(n, array, mask, checkMin, checkMax, true);
}

// ---------------- McICA cloud sampling (rte/extensions/mo_cloud_sampling.F90) ----------------
}  // extern "C"
namespace {
// One thread = (column, group of 4 consecutive g-points) marching down the layers with its random deviates in
// registers: the 4 deviates of a layer are one 32-byte sector of randoms(ngpt,nlay,ncol); the mask is written
// column-fastest (coalesced 1-byte stores).  EXP: exponential-random overlap (:205-292), else maximum-random (:125-192).
// The correlated deviate is evaluated without FMA contraction so that the mask equals the CPU result bit for bit.
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }

template <bool EXP>
void sampled_mask(int ncol, int nlay, int ngpt, const Float* randoms, const Float* cloud_frac, const Float* overlap_param,
                  Bool* cloud_mask) {
  const size_t ncl = (size_t)ncol * nlay;
  DevArg<Float> r(randoms, ncl * ngpt, Dir::In), cf(cloud_frac, ncl, Dir::In);
  DevArg<Float> op(overlap_param, EXP ? (size_t)ncol * (nlay - 1) : 0, Dir::In, EXP);
  DevArg<Bool> m(cloud_mask, ncl * ngpt, Dir::Out);
  const Float* pr = r; const Float* pc = cf; const Float* po = op; Bool* pm = m;
  constexpr int GQ = 4;
  const int ngrp = (ngpt + GQ - 1) / GQ;
  launch_elementwise((size_t)ncol * ngrp, [=] __device__(size_t k) {
    const size_t icol = k % ncol;
    const int g0 = (int)(k / ncol) * GQ, ng = min(GQ, ngpt - g0);
    Float lr[GQ];
    bool prev_cloudy = false, started = false;
    for (int ilay = 0; ilay < nlay; ++ilay) {
      const Float frac = pc[icol + (size_t)ncol * ilay];
      const bool cloudy = frac > (Float)0;  // cloud_mask_layer, :161
      const Float* rl = pr + g0 + (size_t)ngpt * (ilay + (size_t)nlay * icol);
      if (cloudy) {
        Float rho = 0, sq = 0;
        const bool corr = EXP && started && prev_cloudy;
        if (corr) { rho = po[icol + (size_t)ncol * (ilay - 1)]; sq = sqrt(add_rn((Float)1, -mul_rn(rho, rho))); }  // :274-276
#pragma unroll
        for (int q = 0; q < GQ; ++q) {
          if (q < ng) {
            if (corr)  // :275-276, left to right
              lr[q] = add_rn(add_rn(mul_rn(rho, lr[q] - (Float)0.5), mul_rn(sq, rl[q] - (Float)0.5)), (Float)0.5);
            else if (EXP || !(started && prev_cloudy)) lr[q] = rl[q];  // new deviates (:172,:180 / :264,:278)
            pm[icol + (size_t)ncol * ilay + ncl * (size_t)(g0 + q)] = lr[q] > ((Float)1 - frac);  // :173,:181
          }
        }
        started = true;
      } else {
#pragma unroll
        for (int q = 0; q < GQ; ++q)
          if (q < ng) pm[icol + (size_t)ncol * ilay + ncl * (size_t)(g0 + q)] = false;
      }
      prev_cloudy = cloudy;
    }
  });
}
}  // namespace
extern "C" {

void rrtmgpb_sampled_mask_max_ran(int ncol, int nlay, int ngpt, const Float* randoms, const Float* cloud_frac,
                                  Bool* cloud_mask) {
  OpName op_name__(__func__);
  sampled_mask<false>(ncol, nlay, ngpt, randoms, cloud_frac, nullptr, cloud_mask);
}

void rrtmgpb_sampled_mask_exp_ran(int ncol, int nlay, int ngpt, const Float* randoms, const Float* cloud_frac,
                                  const Float* overlap_param, Bool* cloud_mask) {
  OpName op_name__(__func__);
  sampled_mask<true>(ncol, nlay, ngpt, randoms, cloud_frac, overlap_param, cloud_mask);
}

void rrtmgpb_apply_cloud_mask(int ncol, int nlay, int nbnd, int ngpt, const int* band_lims_gpt, const Bool* cloud_mask,
                              const Float* input_field, Float* sampled_field) {
  OpName op_name__(__func__);
  const size_t ncl = (size_t)ncol * nlay, n = ncl * ngpt;
  DevArg<int> lims(band_lims_gpt, 2 * (size_t)nbnd, Dir::In);
  DevArg<Bool> m(cloud_mask, n, Dir::In);
  DevArg<Float> in(input_field, ncl * nbnd, Dir::In), o(sampled_field, n, Dir::Out);
  const int* pl = lims; const Bool* pm = m; const Float* pi = in; Float* po = o;
  launch_elementwise(n, [=] __device__(size_t k) {  // :304-312
    const int g = (int)(k / ncl) + 1; const size_t c = k % ncl;
    int b = 0;
    while (b < nbnd - 1 && g > pl[2 * b + 1]) ++b;
    if (g >= pl[2 * b] && g <= pl[2 * b + 1]) po[k] = pm[k] ? pi[c + ncl * (size_t)b] : (Float)0;
  });
}

// ---------------- aerosol optics (mo_aerosol_optics_rrtmgp_merra.F90) ----------------
void rrtmgpb_aerosol_mask(int ncol, int nlay, const int* type, Bool* aeromsk) {  // :343-347
  OpName op_name__(__func__);
  const size_t n = (size_t)ncol * nlay;
  DevArg<int> t(type, n, Dir::In);
  DevArg<Bool> m(aeromsk, n, Dir::Out);
  const int* pt = t; Bool* pm = m;
  launch_elementwise(n, [=] __device__(size_t k) { pm[k] = pt[k] > 0; });
}

int rrtmgpb_any_int_vals_outside(size_t n, const int* array, int checkMin, int checkMax) {  // :580-600
  OpName op_name__(__func__);
  DevArg<int> a(array, n, Dir::In);
  int* flag = static_cast<int*>(dev_alloc(sizeof(int)));
  RB_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int), stream()));
  const int* pa = a;
  launch_elementwise(n, [=] __device__(size_t i) {
    const int v = pa[i];
    if (v < checkMin || v > checkMax) *flag = 1;
  });
  int h = 0;
  RB_CUDA_CHECK(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, stream()));
  RB_CUDA_CHECK(cudaStreamSynchronize(stream()));
  dev_free(flag);
  return h;
}

// compute_all_from_table (:436-559) fused with the optical-property combination (:385-418): one thread per
// (column, layer) finds its size bin and relative-humidity bracket ONCE (the reference repeats both searches for
// every band), then walks the bands; the three (ncol,nlay,nbnd) temporaries of the reference never exist.
void rrtmgpb_aerosol_optics_from_table(int ncol, int nlay, int nval, int nrh, int nbin, int nbnd, int kind,
                                       const int* type, const Float* size, const Float* mass, const Float* rh,
                                       const Float* bin_lims, const Float* aero_rh, const Float* dust_tbl,
                                       const Float* salt_tbl, const Float* sulf_tbl, const Float* bcar_rh_tbl,
                                       const Float* bcar_tbl, const Float* ocar_rh_tbl, const Float* ocar_tbl,
                                       Float* tau, Float* ssa, Float* g) {
  OpName op_name__(__func__);
  const size_t ncl = (size_t)ncol * nlay, n = ncl * nbnd;
  DevArg<int> a_type(type, ncl, Dir::In);
  DevArg<Float> a_size(size, ncl, Dir::In), a_mass(mass, ncl, Dir::In), a_rh(rh, ncl, Dir::In);
  DevArg<Float> a_lims(bin_lims, 2 * (size_t)nbin, Dir::In), a_arh(aero_rh, (size_t)nrh, Dir::In);
  DevArg<Float> a_dust(dust_tbl, (size_t)nval * nbin * nbnd, Dir::In), a_salt(salt_tbl, (size_t)nrh * nval * nbin * nbnd, Dir::In);
  DevArg<Float> a_sulf(sulf_tbl, (size_t)nrh * nval * nbnd, Dir::In), a_bcrh(bcar_rh_tbl, (size_t)nrh * nval * nbnd, Dir::In);
  DevArg<Float> a_bcar(bcar_tbl, (size_t)nval * nbnd, Dir::In), a_ocrh(ocar_rh_tbl, (size_t)nrh * nval * nbnd, Dir::In);
  DevArg<Float> a_ocar(ocar_tbl, (size_t)nval * nbnd, Dir::In);
  DevArg<Float> o_tau(tau, n, Dir::Out), o_ssa(ssa, n, Dir::Out, kind == 2), o_g(g, n, Dir::Out, kind == 2);
  const int* p_type = a_type;
  const Float *p_size = a_size, *p_mass = a_mass, *p_rh = a_rh, *p_lims = a_lims, *p_arh = a_arh;
  const Float *t_dust = a_dust, *t_salt = a_salt, *t_sulf = a_sulf, *t_bcrh = a_bcrh, *t_bcar = a_bcar,
              *t_ocrh = a_ocrh, *t_ocar = a_ocar;
  Float *p_tau = o_tau, *p_ssa = o_ssa, *p_g = o_g;
  launch_elementwise(ncl, [=] __device__(size_t c) {
    const int itype = p_type[c];
    int ibin = 0;  // 0-based; the LAST bin whose limits bracket the size wins (:457-462)
    int irh1 = 0, irh2 = 0;
    Float rdrh = 0, m = 0;
    if (itype != 0) {
      const Float sz = p_size[c];
      for (int i = 0; i < nbin; ++i)
        if (sz >= __ldg(p_lims + 2 * i) && sz <= __ldg(p_lims + 2 * i + 1)) ibin = i;
      const Float r = p_rh[c];
      int i2 = 1;  // 1-based as in the reference (:466-478)
      while (r > __ldg(p_arh + i2 - 1)) {
        ++i2;
        if (i2 > nrh) break;
      }
      const int i1 = max(1, i2 - 1);
      i2 = min(nrh, i2);
      const Float drh0 = __ldg(p_arh + i2 - 1) - __ldg(p_arh + i1 - 1);
      const Float drh1 = r - __ldg(p_arh + i1 - 1);
      rdrh = (i1 == i2) ? (Float)0 : drh1 / drh0;
      irh1 = i1 - 1; irh2 = i2 - 1;
      m = p_mass[c];
    }
    for (int ibnd = 0; ibnd < nbnd; ++ibnd) {
      Float e = 0, w = 0, asy = 0;  // ext, ssa, g of this (type, bin, rh, band)
      auto lin = [&](const Float* tab) {  // linear_interp_aero_table :571-575 on tab(nrh)
        const Float t1 = __ldg(tab + irh1);
        return t1 + rdrh * (__ldg(tab + irh2) - t1);
      };
      const Float* rh_tab = nullptr;  // (nrh,nval) slice of an rh-dependent table
      const Float* fix_tab = nullptr; // (nval) slice of an rh-independent table
      switch (itype) {
        case 1: fix_tab = t_dust + (size_t)nval * (ibin + (size_t)nbin * ibnd); break;
        case 2: rh_tab = t_salt + (size_t)nrh * nval * (ibin + (size_t)nbin * ibnd); break;
        case 3: rh_tab = t_sulf + (size_t)nrh * nval * ibnd; break;
        case 4: rh_tab = t_bcrh + (size_t)nrh * nval * ibnd; break;
        case 5: fix_tab = t_bcar + (size_t)nval * ibnd; break;
        case 6: rh_tab = t_ocrh + (size_t)nrh * nval * ibnd; break;
        case 7: fix_tab = t_ocar + (size_t)nval * ibnd; break;
        default: break;
      }
      if (rh_tab) { e = lin(rh_tab); w = lin(rh_tab + nrh); asy = lin(rh_tab + 2 * nrh); }
      else if (fix_tab) { e = __ldg(fix_tab); w = __ldg(fix_tab + 1); asy = __ldg(fix_tab + 2); }
      const Float t = m * e, ts = t * w, tsg = ts * asy;  // :502-504
      const size_t o = c + ncl * ibnd;
      if (kind == 1) {
        p_tau[o] = t - ts;                                // :392
      } else {
        p_tau[o] = t;                                     // :405-409
        p_ssa[o] = ts / fmax((Float)RB_EPS, t);
        p_g[o] = tsg / fmax((Float)RB_EPS, ts);
      }
    }
  });
}

}  // extern "C"
