/* ORACLE - TEST INFRASTRUCTURE ONLY.
 * rrtmgpb_allsky_stream_host on the CPU: the same chunked all-sky iteration as the CUDA library's
 * csrc/abi/allsky_stream.cu (reference loop body examples/all-sky/rrtmgp_allsky.F90:332-409), without streams - column
 * slices are gathered into dense arrays and the frontend's calls run on them.  Gives the tests a CPU statement of what
 * the driver computes. */
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "rrtmgp_b200_ext.h"
#include "rrtmgp_b200_frontend.h"
#include "rte_kernels.h"

namespace {
void gather(std::vector<Float>& dst, const Float* src, int ncol, int c0, int n, size_t nrows) {
  dst.resize((size_t)n * nrows);
  for (size_t r = 0; r < nrows; ++r) std::memcpy(&dst[r * n], src + c0 + r * (size_t)ncol, sizeof(Float) * n);
}
void scatter(Float* dst, const std::vector<Float>& src, int ncol, int c0, int n, size_t nrows) {
  for (size_t r = 0; r < nrows; ++r) std::memcpy(dst + c0 + r * (size_t)ncol, &src[r * n], sizeof(Float) * n);
}
}  // namespace

extern "C" int rrtmgpb_allsky_stream_host(const rrtmgpb_gas_optics_t* go_lw, const rrtmgpb_gas_optics_t* go_sw,
                                          const rrtmgpb_cloud_optics_t* co_lw, const rrtmgpb_cloud_optics_t* co_sw,
                                          const rrtmgpb_allsky_host_inputs* in, const rrtmgpb_allsky_host_fluxes* out,
                                          int chunk_cols, int express, char* errmsg) {
  if (errmsg) errmsg[0] = 0;
  auto fail = [&](const std::string& m) { if (errmsg) std::snprintf(errmsg, RRTMGPB_ERRLEN, "%s", m.c_str()); return 1; };
  if (!in || !out || (!go_lw && !go_sw)) return fail("allsky_stream_host: nothing to do");
  const int ncol = in->ncol, nlay = in->nlay, nlev = nlay + 1, ngas = in->ngas;
  const bool clouds = in->lwp != nullptr;
  int nbnd_lw = 0, ngpt_lw = 0, nbnd_sw = 0, ngpt_sw = 0;
  if (go_lw) rrtmgpb_gas_optics_dims(go_lw, nullptr, &nbnd_lw, &ngpt_lw);
  if (go_sw) rrtmgpb_gas_optics_dims(go_sw, nullptr, &nbnd_sw, &ngpt_sw);
  const int nc = std::max(1, std::min(chunk_cols > 0 ? chunk_cols : ncol, ncol));
  std::vector<int> byband_lw(2 * (size_t)std::max(nbnd_lw, 1)), byband_sw(2 * (size_t)std::max(nbnd_sw, 1));
  for (int b = 0; b < nbnd_lw; ++b) byband_lw[2 * b] = byband_lw[2 * b + 1] = b + 1;
  for (int b = 0; b < nbnd_sw; ++b) byband_sw[2 * b] = byband_sw[2 * b + 1] = b + 1;
  char err[RRTMGPB_ERRLEN];
  for (int c0 = 0; c0 < ncol; c0 += nc) {
    const int n = std::min(nc, ncol - c0);
    const size_t nl = (size_t)n * nlay, nlp = (size_t)n * nlev;
    std::vector<Float> p_lay, p_lev, t_lay, t_lev, lwp, iwp, rel, dei, vmr(nl * ngas), tmp;
    gather(p_lay, in->p_lay, ncol, c0, n, nlay); gather(t_lay, in->t_lay, ncol, c0, n, nlay);
    gather(p_lev, in->p_lev, ncol, c0, n, nlev);
    if (in->t_lev) gather(t_lev, in->t_lev, ncol, c0, n, nlev);
    if (clouds) {
      gather(lwp, in->lwp, ncol, c0, n, nlay); gather(iwp, in->iwp, ncol, c0, n, nlay);
      gather(rel, in->rel, ncol, c0, n, nlay); gather(dei, in->dei, ncol, c0, n, nlay);
    }
    for (int g = 0; g < ngas; ++g) {
      if (in->vmr_field && in->vmr_field[g]) { gather(tmp, in->vmr_field[g], ncol, c0, n, nlay); std::copy(tmp.begin(), tmp.end(), vmr.begin() + nl * g); }
      else std::fill(vmr.begin() + nl * g, vmr.begin() + nl * (g + 1), in->vmr_scalar[g]);
    }
    auto props = [&](rrtmgpb_optical_props& o, int kind, int ng, int nb, const int* lims, std::vector<Float>* store) {
      o = rrtmgpb_optical_props{};
      o.kind = kind; o.ncol = n; o.nlay = nlay; o.ngpt = ng; o.nband = nb; o.band_lims_gpt = lims;
      store[0].assign(nl * ng, 0); o.tau = store[0].data();
      if (kind == RRTMGPB_2STR) { store[1].assign(nl * ng, 0); store[2].assign(nl * ng, 0); o.ssa = store[1].data(); o.g = store[2].data(); }
    };
    std::vector<Float> up(nlp), dn(nlp), dir(nlp);
    if (go_lw && (out->lw_flux_up || out->lw_flux_dn)) {
      rrtmgpb_optical_props cld, atm;
      std::vector<Float> cs[3], as[3], lay(nl * ngpt_lw), lev(nlp * ngpt_lw), sfc((size_t)n * ngpt_lw), jac((size_t)n * ngpt_lw);
      props(cld, RRTMGPB_1SCL, nbnd_lw, nbnd_lw, byband_lw.data(), cs);
      if (clouds && rrtmgpb_cloud_optics(co_lw, n, nlay, lwp.data(), iwp.data(), rel.data(), dei.data(), &cld, err)) return fail(err);
      rrtmgpb_fluxes_broadband fl{up.data(), dn.data(), nullptr, nullptr};
      const Float* tsfc = in->t_sfc + c0; const Float* emis = in->emis_sfc + (size_t)nbnd_lw * c0;
      const Float* tl = in->t_lev ? t_lev.data() : nullptr;
      if (express) {
        if (rrtmgpb_rte_lw_express(go_lw, n, nlay, p_lay.data(), p_lev.data(), t_lay.data(), tsfc, vmr.data(), nullptr, tl,
                                   clouds ? &cld : nullptr, emis, 0, &fl, err)) return fail(err);
      } else {
        props(atm, RRTMGPB_1SCL, ngpt_lw, nbnd_lw, rrtmgpb_gas_optics_band_lims_gpt(go_lw), as);
        rrtmgpb_source_func_lw src{n, nlay, ngpt_lw, lay.data(), lev.data(), sfc.data(), jac.data()};
        if (rrtmgpb_gas_optics_int_fused(go_lw, n, nlay, p_lay.data(), p_lev.data(), t_lay.data(), tsfc, vmr.data(), &atm, &src,
                                         nullptr, tl, clouds ? &cld : nullptr, nullptr, err)) return fail(err);
        if (rrtmgpb_rte_lw(&atm, &src, emis, &fl, nullptr, 0, -1, nullptr, nullptr, err)) return fail(err);
      }
      if (out->lw_flux_up) scatter(out->lw_flux_up, up, ncol, c0, n, nlev);
      if (out->lw_flux_dn) scatter(out->lw_flux_dn, dn, ncol, c0, n, nlev);
    }
    if (go_sw && (out->sw_flux_up || out->sw_flux_dn || out->sw_flux_dir)) {
      rrtmgpb_optical_props cld, atm;
      std::vector<Float> cs[3], as[3], toa((size_t)n * ngpt_sw);
      props(cld, RRTMGPB_2STR, nbnd_sw, nbnd_sw, byband_sw.data(), cs);
      if (clouds && rrtmgpb_cloud_optics_delta_scaled(co_sw, n, nlay, lwp.data(), iwp.data(), rel.data(), dei.data(), &cld, 1, err)) return fail(err);
      rrtmgpb_fluxes_broadband fs{up.data(), dn.data(), nullptr, dir.data()};
      const Float* mu0 = in->mu0 + c0; const Float* ad = in->sfc_alb_dir + (size_t)nbnd_sw * c0; const Float* af = in->sfc_alb_dif + (size_t)nbnd_sw * c0;
      if (express) {
        if (rrtmgpb_rte_sw_express(go_sw, n, nlay, p_lay.data(), p_lev.data(), t_lay.data(), vmr.data(), nullptr,
                                   clouds ? &cld : nullptr, mu0, ad, af, &fs, err)) return fail(err);
      } else {
        props(atm, RRTMGPB_2STR, ngpt_sw, nbnd_sw, rrtmgpb_gas_optics_band_lims_gpt(go_sw), as);
        if (rrtmgpb_gas_optics_ext_fused(go_sw, n, nlay, p_lay.data(), p_lev.data(), t_lay.data(), vmr.data(), &atm, toa.data(),
                                         nullptr, clouds ? &cld : nullptr, nullptr, err)) return fail(err);
        if (rrtmgpb_rte_sw(&atm, mu0, toa.data(), ad, af, &fs, nullptr, err)) return fail(err);
      }
      if (out->sw_flux_up) scatter(out->sw_flux_up, up, ncol, c0, n, nlev);
      if (out->sw_flux_dn) scatter(out->sw_flux_dn, dn, ncol, c0, n, nlev);
      if (out->sw_flux_dir) scatter(out->sw_flux_dir, dir, ncol, c0, n, nlev);
    }
  }
  return 0;
}
