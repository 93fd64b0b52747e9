// allsky_stream.cu - the all-sky iteration on HOST buffers, streamed through the device in column chunks
// (include/rrtmgp_b200_frontend.h: rrtmgpb_allsky_stream_host).
//
// Reference workload: the loop body of examples/all-sky/rrtmgp_allsky.F90:332-409.  What a host model that keeps its
// state in CPU memory needs from an accelerator backend is not a kernel but this: inputs cross PCIe once, fluxes come
// back once, and neither transfer is exposed.  Three streams:
//     up      chunk k+1's input slices  (cudaMemcpy2DAsync: a column range of a Fortran (ncol, nlay) array is strided)
//     comp    even chunks: gas_concs broadcast, cloud optics, gas optics, solver  (the C++ frontend's own calls)
//     comp2   odd chunks, concurrently: a chunk's kernels run on grids a few waves long, and the idle tail of one
//             chunk's kernel is filled by the other chunk's (measured: 45.4 -> see DESIGN.md ms of chunked compute)
//     down    finished chunks' five flux slices
// ordered by events; two device input sets, two work-array sets and two device flux sets.  With pinned host memory every copy is
// asynchronous; with pageable memory CUDA stages them (correct, serialised).
#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>
#include "../common.cuh"
#include "rrtmgp_b200_ext.h"
#include "rrtmgp_b200_frontend.h"
#include "rte_kernels.h"

using namespace rrtmgpb;

namespace {

struct Streams {
  cudaStream_t up = nullptr, down = nullptr, comp2 = nullptr;
  cudaEvent_t joined = nullptr;
  cudaEvent_t in_ready[2] = {}, in_free[2] = {}, out_ready[2] = {}, out_free[2] = {};
};
Streams& streams() {
  static thread_local Streams s;
  if (!s.up) {
    RB_CUDA_CHECK(cudaStreamCreateWithFlags(&s.up, cudaStreamNonBlocking));
    RB_CUDA_CHECK(cudaStreamCreateWithFlags(&s.down, cudaStreamNonBlocking));
    RB_CUDA_CHECK(cudaStreamCreateWithFlags(&s.comp2, cudaStreamNonBlocking));
    RB_CUDA_CHECK(cudaEventCreateWithFlags(&s.joined, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
      RB_CUDA_CHECK(cudaEventCreateWithFlags(&s.in_ready[i], cudaEventDisableTiming));
      RB_CUDA_CHECK(cudaEventCreateWithFlags(&s.in_free[i], cudaEventDisableTiming));
      RB_CUDA_CHECK(cudaEventCreateWithFlags(&s.out_ready[i], cudaEventDisableTiming));
      RB_CUDA_CHECK(cudaEventCreateWithFlags(&s.out_free[i], cudaEventDisableTiming));
    }
  }
  return s;
}

// columns [c0, c0+n) of a HOST Fortran (ncol, nrows) array -> dense device (n, nrows), and back
void upload_cols(Float* dst, const Float* src, int ncol, int c0, int n, size_t nrows, cudaStream_t st) {
  RB_CUDA_CHECK(cudaMemcpy2DAsync(dst, (size_t)n * sizeof(Float), src + c0, (size_t)ncol * sizeof(Float),
                                  (size_t)n * sizeof(Float), nrows, cudaMemcpyHostToDevice, st));
}
void download_cols(Float* dst, const Float* src, int ncol, int c0, int n, size_t nrows, cudaStream_t st) {
  RB_CUDA_CHECK(cudaMemcpy2DAsync(dst + c0, (size_t)ncol * sizeof(Float), src, (size_t)n * sizeof(Float),
                                  (size_t)n * sizeof(Float), nrows, cudaMemcpyDeviceToHost, st));
}

struct InSet {   // device copies of one chunk's inputs
  Float *p_lay, *p_lev, *t_lay, *t_lev, *lwp, *iwp, *rel, *dei, *t_sfc, *emis, *mu0, *alb_dir, *alb_dif;
  std::vector<Float*> field;   // [ngas]: uploaded field or nullptr
};
struct OutSet { Float* f[5]; };

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
  if (clouds && ((go_lw && !co_lw) || (go_sw && !co_sw))) return fail("allsky_stream_host: cloud inputs given without cloud optics");
  int nbnd_lw = 0, ngpt_lw = 0, nbnd_sw = 0, ngpt_sw = 0, g1 = ngas, g2 = ngas;
  if (go_lw) rrtmgpb_gas_optics_dims(go_lw, &g1, &nbnd_lw, &ngpt_lw);
  if (go_sw) rrtmgpb_gas_optics_dims(go_sw, &g2, &nbnd_sw, &ngpt_sw);
  if (g1 != ngas || g2 != ngas) return fail("allsky_stream_host: ngas differs from the k-distributions'");
  // Chunk schedule: a SMALL first chunk (a quarter of the nominal width) so that compute starts as soon as possible -
  // its upload is the only one that is not hidden behind compute - then full chunks, then two one-wave chunks (below).  Nominal width: chunk_cols, or by default four waves of the
  // register solvers' 16-column CTAs (2 resident per SM), so that no chunk ends on a nearly empty wave.
  int sms = 148;
  {
    int dev = 0;
    RB_CUDA_CHECK(cudaGetDevice(&dev));
    RB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int wave = 2 * sms * 16;
  const int nc = std::max(1, std::min(chunk_cols > 0 ? chunk_cols : 4 * wave, ncol));
  std::vector<int> starts;   // first column of every chunk, plus ncol
  {
    int c = 0;
    const int first = ncol > nc ? std::max(std::min(wave, nc), nc / 4) : nc;
    starts.push_back(0);
    c = std::min(first, ncol);
    // Optional shapes of the schedule (measured on 1 and 8 GPUs):
    //   RRTMGPB_STREAM_RAMP=1  widths 1, 2, 4 waves before the full chunks: every upload hides behind the previous chunk's
    //                          compute even when the host's PCIe / memory paths are shared by 8 ranks
    //   RRTMGPB_STREAM_TAIL=1  two one-wave chunks at the end: with two compute streams the last chunk of EACH stream ends
    //                          the call, and the exposed download is one wave's fluxes (14 MB) instead of a chunk's (55 MB)
    static const bool ramp = [] { const char* e = std::getenv("RRTMGPB_STREAM_RAMP"); return e && e[0] == '1'; }();
    static const bool want_tail = [] { const char* e = std::getenv("RRTMGPB_STREAM_TAIL"); return e && e[0] == '1'; }();
    const int tail = (want_tail && nc > wave && ncol - c >= 2 * nc) ? wave : 0;
    const int body_end = ncol - 2 * tail;
    int width = (ramp && nc > 2 * wave) ? 2 * wave : nc;
    while (c < body_end) {
      starts.push_back(c);
      c = std::min(c + width, body_end);
      width = std::min(2 * width, nc);
    }
    if (tail) { starts.push_back(body_end); starts.push_back(body_end + tail); }
    starts.push_back(ncol);
  }
  const size_t ncl = (size_t)nc * nlay, nclp = (size_t)nc * nlev;
  Streams& S = streams();
  cudaStream_t comp = stream();
  auto dalloc = [&](size_t n) { return static_cast<Float*>(dev_alloc(std::max<size_t>(n, 1) * sizeof(Float))); };
  // ---- device buffers (stream-ordered pool on the compute stream; the copy streams wait for `start`)
  InSet I[2];
  OutSet O[2];
  std::vector<void*> owned;
  auto take = [&](size_t n) { Float* p = dalloc(n); owned.push_back(p); return p; };
  for (int k = 0; k < 2; ++k) {
    I[k].p_lay = take(ncl); I[k].t_lay = take(ncl); I[k].p_lev = take(nclp); I[k].t_lev = take(nclp);
    I[k].lwp = clouds ? take(ncl) : nullptr; I[k].iwp = clouds ? take(ncl) : nullptr;
    I[k].rel = clouds ? take(ncl) : nullptr; I[k].dei = clouds ? take(ncl) : nullptr;
    I[k].t_sfc = go_lw ? take(nc) : nullptr; I[k].emis = go_lw ? take((size_t)nbnd_lw * nc) : nullptr;
    I[k].mu0 = go_sw ? take(nc) : nullptr; I[k].alb_dir = go_sw ? take((size_t)nbnd_sw * nc) : nullptr;
    I[k].alb_dif = go_sw ? take((size_t)nbnd_sw * nc) : nullptr;
    I[k].field.assign((size_t)ngas, nullptr);
    for (int g = 0; g < ngas; ++g)
      if (in->vmr_field && in->vmr_field[g]) I[k].field[(size_t)g] = take(ncl);
    for (int a = 0; a < 5; ++a) O[k].f[a] = take(nclp);
  }
  // work arrays, one set per concurrent chunk: vmr, by-band cloud properties and (plane path) the chunk-sized planes
  const int* bl_lw = go_lw ? rrtmgpb_gas_optics_band_lims_gpt(go_lw) : nullptr;
  const int* bl_sw = go_sw ? rrtmgpb_gas_optics_band_lims_gpt(go_sw) : nullptr;
  std::vector<int> byband_lw(2 * (size_t)std::max(nbnd_lw, 1)), byband_sw(2 * (size_t)std::max(nbnd_sw, 1));
  for (int b = 0; b < nbnd_lw; ++b) byband_lw[2 * b] = byband_lw[2 * b + 1] = b + 1;
  for (int b = 0; b < nbnd_sw; ++b) byband_sw[2 * b] = byband_sw[2 * b + 1] = b + 1;
  struct Work {
    Float* vmr;
    rrtmgpb_optical_props atm_lw, atm_sw, cld_lw, cld_sw;
    rrtmgpb_source_func_lw src;
    Float* toa;
  } W[2] = {};
  auto props = [&](rrtmgpb_optical_props& o, int kind, int ng, int nb, const int* lims) {
    o.kind = kind; o.ncol = nc; o.nlay = nlay; o.ngpt = ng; o.nband = nb; o.nmom = 0; o.top_at_1 = 0;
    o.band_lims_gpt = lims; o.band_lims_wvn = nullptr;
    o.tau = take(ncl * ng);
    o.ssa = kind == RRTMGPB_2STR ? take(ncl * ng) : nullptr;
    o.g = kind == RRTMGPB_2STR ? take(ncl * ng) : nullptr;
  };
  const int nsets = ncol > nc ? 2 : 1;
  for (int k = 0; k < nsets; ++k) {
    Work& w = W[k];
    w.vmr = take(ncl * ngas);
    if (go_lw && clouds) props(w.cld_lw, RRTMGPB_1SCL, nbnd_lw, nbnd_lw, byband_lw.data());
    if (go_sw && clouds) props(w.cld_sw, RRTMGPB_2STR, nbnd_sw, nbnd_sw, byband_sw.data());
    if (!express) {
      if (go_lw) {
        props(w.atm_lw, RRTMGPB_1SCL, ngpt_lw, nbnd_lw, bl_lw);
        w.src.ncol = nc; w.src.nlay = nlay; w.src.ngpt = ngpt_lw;
        w.src.lay_source = take(ncl * ngpt_lw); w.src.lev_source = take(nclp * ngpt_lw);
        w.src.sfc_source = take((size_t)nc * ngpt_lw); w.src.sfc_source_Jac = take((size_t)nc * ngpt_lw);
      }
      if (go_sw) { props(w.atm_sw, RRTMGPB_2STR, ngpt_sw, nbnd_sw, bl_sw); w.toa = take((size_t)nc * ngpt_sw); }
    }
  }
  cudaStream_t cstream[2] = {comp, nsets > 1 ? S.comp2 : comp};
  // the driver holds the pressures on the host: tell the frontend the orientation (no device reads / syncs per call)
  rrtmgpb_set_top_at_1_hint(in->p_lay[0] < in->p_lay[(size_t)ncol * (nlay - 1)] ? 1 : 0);
  cudaEvent_t start;
  RB_CUDA_CHECK(cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
  RB_CUDA_CHECK(cudaEventRecord(start, comp));
  RB_CUDA_CHECK(cudaStreamWaitEvent(S.up, start, 0));
  RB_CUDA_CHECK(cudaStreamWaitEvent(S.down, start, 0));
  if (nsets > 1) RB_CUDA_CHECK(cudaStreamWaitEvent(S.comp2, start, 0));
  for (int k = 0; k < 2; ++k) {   // both sets start out free
    RB_CUDA_CHECK(cudaEventRecord(S.in_free[k], comp));
    RB_CUDA_CHECK(cudaEventRecord(S.out_free[k], S.down));
  }

  const int nchunk = (int)starts.size() - 1;
  // RRTMGPB_STREAM_TRACE=1: per-chunk timeline (upload / compute / download, ms since the call started) on stderr
  static const bool trace = [] { const char* e = std::getenv("RRTMGPB_STREAM_TRACE"); return e && e[0] == '1'; }();
  struct Tr { cudaEvent_t u0, u1, c0, c1, d0, d1; };
  std::vector<Tr> tr;
  cudaEvent_t t_start = nullptr;
  if (trace) {
    tr.resize((size_t)nchunk);
    for (Tr& t : tr) for (cudaEvent_t* e : {&t.u0, &t.u1, &t.c0, &t.c1, &t.d0, &t.d1}) RB_CUDA_CHECK(cudaEventCreate(e));
    RB_CUDA_CHECK(cudaEventCreate(&t_start));
    RB_CUDA_CHECK(cudaEventRecord(t_start, comp));
  }
  auto upload = [&](int ic) {
    const int k = ic & 1, c0 = starts[(size_t)ic], n = starts[(size_t)ic + 1] - c0;
    RB_CUDA_CHECK(cudaStreamWaitEvent(S.up, S.in_free[k], 0));
    if (trace) RB_CUDA_CHECK(cudaEventRecord(tr[(size_t)ic].u0, S.up));
    upload_cols(I[k].p_lay, in->p_lay, ncol, c0, n, nlay, S.up);
    upload_cols(I[k].t_lay, in->t_lay, ncol, c0, n, nlay, S.up);
    upload_cols(I[k].p_lev, in->p_lev, ncol, c0, n, nlev, S.up);
    if (go_lw && in->t_lev) upload_cols(I[k].t_lev, in->t_lev, ncol, c0, n, nlev, S.up);
    if (clouds) {
      upload_cols(I[k].lwp, in->lwp, ncol, c0, n, nlay, S.up);
      upload_cols(I[k].iwp, in->iwp, ncol, c0, n, nlay, S.up);
      upload_cols(I[k].rel, in->rel, ncol, c0, n, nlay, S.up);
      upload_cols(I[k].dei, in->dei, ncol, c0, n, nlay, S.up);
    }
    for (int g = 0; g < ngas; ++g)
      if (I[k].field[(size_t)g]) upload_cols(I[k].field[(size_t)g], in->vmr_field[g], ncol, c0, n, nlay, S.up);
    auto up1 = [&](Float* d, const Float* h, size_t per_col) {   // (k, ncol) arrays: a column range is contiguous
      RB_CUDA_CHECK(cudaMemcpyAsync(d, h + per_col * c0, per_col * n * sizeof(Float), cudaMemcpyHostToDevice, S.up));
    };
    if (go_lw) { up1(I[k].t_sfc, in->t_sfc, 1); up1(I[k].emis, in->emis_sfc, nbnd_lw); }
    if (go_sw) { up1(I[k].mu0, in->mu0, 1); up1(I[k].alb_dir, in->sfc_alb_dir, nbnd_sw); up1(I[k].alb_dif, in->sfc_alb_dif, nbnd_sw); }
    RB_CUDA_CHECK(cudaEventRecord(S.in_ready[k], S.up));
    if (trace) RB_CUDA_CHECK(cudaEventRecord(tr[(size_t)ic].u1, S.up));
  };

  std::string msg;
  char err[RRTMGPB_ERRLEN];
  upload(0);
  for (int ic = 0; ic < nchunk && msg.empty(); ++ic) {
    const int k = ic & 1, c0 = starts[(size_t)ic], n = starts[(size_t)ic + 1] - c0;
    if (ic + 1 < nchunk) upload(ic + 1);
    cudaStream_t cs = cstream[k];
    rrtmgpb_set_stream(cs);   // the frontend launches on the calling thread's stream: this chunk's
    RB_CUDA_CHECK(cudaStreamWaitEvent(cs, S.in_ready[k], 0));
    RB_CUDA_CHECK(cudaStreamWaitEvent(cs, S.out_free[k], 0));
    const InSet& X = I[k];
    Work& w = W[k % nsets];
    Float* vmr = w.vmr;
    rrtmgpb_optical_props &atm_lw = w.atm_lw, &atm_sw = w.atm_sw, &cld_lw = w.cld_lw, &cld_sw = w.cld_sw;
    rrtmgpb_source_func_lw& src = w.src;
    Float* toa = w.toa;
    if (trace) RB_CUDA_CHECK(cudaEventRecord(tr[(size_t)ic].c0, cs));
    // gas_concs -> vmr(n, nlay, ngas): fields as uploaded, well-mixed gases broadcast (mo_gas_optics_rrtmgp.F90:540-545)
    for (int g = 0; g < ngas; ++g) {
      Float* plane = vmr + (size_t)n * nlay * g;
      if (X.field[(size_t)g]) rrtmgpb_mem_copy(plane, X.field[(size_t)g], (size_t)n * nlay * sizeof(Float));
      else set_to_scalar_2D(&n, &nlay, plane, &in->vmr_scalar[g]);
    }
    rrtmgpb_fluxes_broadband fl{O[k].f[0], O[k].f[1], nullptr, nullptr}, fs{O[k].f[2], O[k].f[3], nullptr, O[k].f[4]};
    cld_lw.ncol = cld_sw.ncol = atm_lw.ncol = atm_sw.ncol = src.ncol = n;
    if (go_lw && (out->lw_flux_up || out->lw_flux_dn)) {
      if (clouds && rrtmgpb_cloud_optics(co_lw, n, nlay, X.lwp, X.iwp, X.rel, X.dei, &cld_lw, err)) { msg = err; break; }
      const Float* tlev = in->t_lev ? X.t_lev : nullptr;
      if (express) {
        if (rrtmgpb_rte_lw_express(go_lw, n, nlay, X.p_lay, X.p_lev, X.t_lay, X.t_sfc, vmr, nullptr, tlev,
                                   clouds ? &cld_lw : nullptr, X.emis, 0, &fl, err)) { msg = err; break; }
      } else {
        if (rrtmgpb_gas_optics_int_fused(go_lw, n, nlay, X.p_lay, X.p_lev, X.t_lay, X.t_sfc, vmr, &atm_lw, &src, nullptr, tlev,
                                         clouds ? &cld_lw : nullptr, nullptr, err)) { msg = err; break; }
        if (rrtmgpb_rte_lw(&atm_lw, &src, X.emis, &fl, nullptr, 0, -1, nullptr, nullptr, err)) { msg = err; break; }
      }
    }
    if (go_sw && (out->sw_flux_up || out->sw_flux_dn || out->sw_flux_dir)) {
      if (clouds && rrtmgpb_cloud_optics_delta_scaled(co_sw, n, nlay, X.lwp, X.iwp, X.rel, X.dei, &cld_sw, 1, err)) { msg = err; break; }
      if (express) {
        if (rrtmgpb_rte_sw_express(go_sw, n, nlay, X.p_lay, X.p_lev, X.t_lay, vmr, nullptr, clouds ? &cld_sw : nullptr, X.mu0,
                                   X.alb_dir, X.alb_dif, &fs, err)) { msg = err; break; }
      } else {
        if (rrtmgpb_gas_optics_ext_fused(go_sw, n, nlay, X.p_lay, X.p_lev, X.t_lay, vmr, &atm_sw, toa, nullptr,
                                         clouds ? &cld_sw : nullptr, nullptr, err)) { msg = err; break; }
        if (rrtmgpb_rte_sw(&atm_sw, X.mu0, toa, X.alb_dir, X.alb_dif, &fs, nullptr, err)) { msg = err; break; }
      }
    }
    RB_CUDA_CHECK(cudaEventRecord(S.in_free[k], cs));
    RB_CUDA_CHECK(cudaEventRecord(S.out_ready[k], cs));
    if (trace) RB_CUDA_CHECK(cudaEventRecord(tr[(size_t)ic].c1, cs));
    // ---- fluxes of this chunk -> host, behind the compute stream
    RB_CUDA_CHECK(cudaStreamWaitEvent(S.down, S.out_ready[k], 0));
    if (trace) RB_CUDA_CHECK(cudaEventRecord(tr[(size_t)ic].d0, S.down));
    Float* dsts[5] = {out->lw_flux_up, out->lw_flux_dn, out->sw_flux_up, out->sw_flux_dn, out->sw_flux_dir};
    for (int a = 0; a < 5; ++a)
      if (dsts[a] && ((a < 2 && go_lw) || (a >= 2 && go_sw))) download_cols(dsts[a], O[k].f[a], ncol, c0, n, nlev, S.down);
    RB_CUDA_CHECK(cudaEventRecord(S.out_free[k], S.down));
    if (trace) RB_CUDA_CHECK(cudaEventRecord(tr[(size_t)ic].d1, S.down));
  }
  rrtmgpb_set_stream(comp);
  rrtmgpb_set_top_at_1_hint(-1);
  if (nsets > 1) {   // join: the caller's stream continues after the auxiliary compute stream
    RB_CUDA_CHECK(cudaEventRecord(S.joined, S.comp2));
    RB_CUDA_CHECK(cudaStreamWaitEvent(comp, S.joined, 0));
  }
  // the call returns when the last fluxes are on the host; the compute stream is ordered behind both copy streams
  RB_CUDA_CHECK(cudaStreamSynchronize(S.down));
  RB_CUDA_CHECK(cudaStreamSynchronize(S.up));
  RB_CUDA_CHECK(cudaStreamSynchronize(comp));
  if (trace && msg.empty()) {
    auto ms = [&](cudaEvent_t e) { float v = 0; RB_CUDA_CHECK(cudaEventElapsedTime(&v, t_start, e)); return v; };
    for (int ic = 0; ic < nchunk; ++ic) {
      const Tr& t = tr[(size_t)ic];
      std::fprintf(stderr, "stream trace: chunk %d cols [%d,%d)  up %.2f-%.2f  compute %.2f-%.2f  down %.2f-%.2f ms\n", ic,
                   starts[(size_t)ic], starts[(size_t)ic + 1], ms(t.u0), ms(t.u1), ms(t.c0), ms(t.c1), ms(t.d0), ms(t.d1));
    }
  }
  if (trace) {
    for (Tr& t : tr) for (cudaEvent_t e : {t.u0, t.u1, t.c0, t.c1, t.d0, t.d1}) cudaEventDestroy(e);
    cudaEventDestroy(t_start);
  }
  for (void* p : owned) dev_free(p);
  RB_CUDA_CHECK(cudaEventDestroy(start));
  return msg.empty() ? 0 : fail(msg);
}
