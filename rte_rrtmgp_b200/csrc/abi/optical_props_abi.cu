// optical_props_abi.cu - delta scaling, the 18 increment variants and the 3 subset extractors.
//
// Replaces rte/kernels/api/mo_optical_props_kernels.F90:36-368 (extern mode); arithmetic follows the
// default kernels rte/kernels/mo_optical_props_kernels.F90:38-706 (eps = 3*tiny guards, operation
// order).  All are read-modify-write streams over (ncol,nlay,ngpt) planes: HBM-bound, one thread per
// element in flat (column-innermost) order, by-band operands broadcast over the band's g-points
// from a g-point -> band map built on the device (no host read of gpt_lims).
#include "../kernels/elementwise.cuh"
#include "rte_kernels.h"

using namespace rrtmgpb;

namespace {

#define OP_EPS ((Float)3.0 * (Float)RB_TINY)

// ---- per-element bodies: i indexes operand 1, j operand 2 (same plane, or the band plane) ----
__device__ __forceinline__ void inc_1s_1s(Float* tau1, const Float* tau2, size_t i, size_t j) {
  tau1[i] = tau1[i] + tau2[j];
}
__device__ __forceinline__ void inc_1s_2s(Float* tau1, const Float* tau2, const Float* ssa2, size_t i, size_t j) {
  tau1[i] = tau1[i] + tau2[j] * ((Float)1 - ssa2[j]);
}
__device__ __forceinline__ void inc_2s_1s(Float* tau1, Float* ssa1, const Float* tau2, size_t i, size_t j) {
  const Float t1 = tau1[i];
  const Float tau12 = t1 + tau2[j];
  ssa1[i] = t1 * ssa1[i] / fmax(OP_EPS, tau12);
  tau1[i] = tau12;
}
// (IEEE divisions on purpose: the branch-free rb_div was measured here and changes nothing - these kernels sit at 0.71
// of the HBM peak on their six-plane read-modify-write traffic, not on the divisions.)
__device__ __forceinline__ void inc_2s_2s(Float* tau1, Float* ssa1, Float* g1, const Float* tau2,
                                          const Float* ssa2, const Float* g2, size_t i, size_t j,
                                          size_t g2stride) {
  const Float t1 = tau1[i], w1 = ssa1[i], t2 = tau2[j], w2 = ssa2[j];
  const Float tau12 = t1 + t2;
  const Float tauscat12 = t1 * w1 + t2 * w2;
  g1[i] = (t1 * w1 * g1[i] + t2 * w2 * g2[j * g2stride]) / fmax(OP_EPS, tauscat12);
  ssa1[i] = tauscat12 / fmax(OP_EPS, tau12);
  tau1[i] = tau12;
}
__device__ __forceinline__ void inc_ns_2s(int nmom1, Float* tau1, Float* ssa1, Float* p1, const Float* tau2,
                                          const Float* ssa2, const Float* g2, size_t i, size_t j) {
  const Float t1 = tau1[i], w1 = ssa1[i], t2 = tau2[j], w2 = ssa2[j], gg = g2[j];
  const Float tau12 = t1 + t2;
  const Float tauscat12 = t1 * w1 + t2 * w2;
  Float mom = gg;
  for (int m = 0; m < nmom1; ++m) {
    if (m > 0) mom = mom * gg;
    const size_t k = (size_t)m + (size_t)nmom1 * i;
    p1[k] = (t1 * w1 * p1[k] + t2 * w2 * mom) / fmax(OP_EPS, tauscat12);
  }
  ssa1[i] = tauscat12 / fmax(OP_EPS, tau12);
  tau1[i] = tau12;
}
__device__ __forceinline__ void inc_ns_ns(int nmom1, int nmom2, Float* tau1, Float* ssa1, Float* p1,
                                          const Float* tau2, const Float* ssa2, const Float* p2, size_t i,
                                          size_t j) {
  const int mom_lim = nmom1 < nmom2 ? nmom1 : nmom2;
  const Float t1 = tau1[i], w1 = ssa1[i], t2 = tau2[j], w2 = ssa2[j];
  const Float tau12 = t1 + t2;
  const Float tauscat12 = t1 * w1 + t2 * w2;
  for (int m = 0; m < mom_lim; ++m) {
    const size_t k = (size_t)m + (size_t)nmom1 * i;
    p1[k] = (t1 * w1 * p1[k] + t2 * w2 * p2[(size_t)m + (size_t)nmom2 * j]) / fmax(OP_EPS, tauscat12);
  }
  ssa1[i] = tauscat12 / fmax(OP_EPS, tau12);
  tau1[i] = tau12;
}

// g-point -> band (0-based) map on the device; -1 for g-points outside every band (untouched, as in
// the reference's loops over gpt_lims).
int* build_gpt2bnd(int ngpt, int nbnd, const int* gpt_lims_dev) {
  int* map = static_cast<int*>(dev_alloc(sizeof(int) * (size_t)ngpt));
  launch_elementwise((size_t)ngpt, [=] __device__(size_t g) {
    int b = -1;
    for (int ib = 0; ib < nbnd; ++ib)
      if ((int)g + 1 >= gpt_lims_dev[2 * ib] && (int)g + 1 <= gpt_lims_dev[2 * ib + 1]) { b = ib; break; }
    map[g] = b;
  });
  return map;
}

// by-band launch: blockIdx.y = g-point (its band is looked up once per block, uniformly), x strides over the cells of
// the plane - no 64-bit division / modulo per element (the flat mapping spent more on i / ncl, i % ncl than on its
// arithmetic: inc_1scalar_by_1scalar_bybnd ran at 0.64 of the HBM peak).  f(i, j): i indexes operand 1, j the band plane.
template <typename F>
__global__ void __launch_bounds__(256) bybnd_kernel(size_t ncl, const int* __restrict__ map, F f) {
  const int bnd = map[blockIdx.y];
  if (bnd < 0) return;  // g-point outside every band: untouched, as in the reference's loops over gpt_lims
  const size_t gi = ncl * (size_t)blockIdx.y, bj = ncl * (size_t)bnd;
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncl; c += (size_t)gridDim.x * blockDim.x)
    f(gi + c, bj + c);
}
template <typename F>
void launch_bybnd(size_t ncl, int ngpt, const int* map, F f) {
  if (ncl == 0 || ngpt <= 0) return;
  KernelTimer timer(tl_op_name ? tl_op_name : "bybnd");
  size_t bx = (ncl + 255) / 256;
  if (bx > 4096) bx = 4096;
  for (int g0 = 0; g0 < ngpt; g0 += 65535) {  // gridDim.y limit
    const int ng = ngpt - g0 < 65535 ? ngpt - g0 : 65535;
    const size_t off = ncl * (size_t)g0;
    bybnd_kernel<<<dim3((unsigned)bx, (unsigned)ng), 256, 0, stream()>>>(
        ncl, map + g0, [=] __device__(size_t i, size_t j) { f(i + off, j); });
    RB_LAUNCH_CHECK();
  }
}

struct Dims {
  size_t ncl, n;
  int ngpt;
  Dims(const int* ncol, const int* nlay, const int* ngpt_) : ncl((size_t)*ncol * *nlay), n(ncl * *ngpt_), ngpt(*ngpt_) {}
};

}  // namespace

extern "C" {

void rte_delta_scale_2str_f_k(const int* ncol, const int* nlay, const int* ngpt, Float* tau, Float* ssa,
                              Float* g, const Float* f) {
  OpName op_name__(__func__);
  Dims d(ncol, nlay, ngpt);
  DevArg<Float> t(tau, d.n, Dir::InOut), s(ssa, d.n, Dir::InOut), gg(g, d.n, Dir::InOut), ff(f, d.n, Dir::In);
  Float *pt = t, *ps = s, *pg = gg; const Float* pf = ff;
  launch_elementwise(d.n, [=] __device__(size_t i) {  // mo_optical_props_kernels.F90:62-66
    const Float w = ps[i], fi = pf[i];
    const Float wf = w * fi;
    pt[i] = ((Float)1 - wf) * pt[i];
    ps[i] = (w - wf) / fmax(OP_EPS, ((Float)1 - wf));
    pg[i] = (pg[i] - fi) / fmax(OP_EPS, ((Float)1 - fi));
  });
}

void rte_delta_scale_2str_k(const int* ncol, const int* nlay, const int* ngpt, Float* tau, Float* ssa, Float* g) {
  OpName op_name__(__func__);
  Dims d(ncol, nlay, ngpt);
  DevArg<Float> t(tau, d.n, Dir::InOut), s(ssa, d.n, Dir::InOut), gg(g, d.n, Dir::InOut);
  Float *pt = t, *ps = s, *pg = gg;
  launch_elementwise(d.n, [=] __device__(size_t i) {  // :89-93
    const Float w = ps[i], gi = pg[i];
    const Float fi = gi * gi;
    const Float wf = w * fi;
    pt[i] = ((Float)1 - wf) * pt[i];
    ps[i] = (w - wf) / fmax(OP_EPS, ((Float)1 - wf));
    pg[i] = (gi - fi) / fmax(OP_EPS, ((Float)1 - fi));
  });
}

// ---------------- same-resolution increments (:116-358) ----------------
void rte_increment_1scalar_by_1scalar(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      const Float* tau2) {
  OpName op_name__(__func__);
  Dims d(ncol, nlay, ngpt);
  DevArg<Float> t1(tau1, d.n, Dir::InOut), t2(tau2, d.n, Dir::In);
  Float* a = t1; const Float* b = t2;
  launch_elementwise(d.n, [=] __device__(size_t i) { inc_1s_1s(a, b, i, i); });
}
static void inc_1scalar_by_scattering(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      const Float* tau2, const Float* ssa2) {
  OpName op_name__(__func__);
  Dims d(ncol, nlay, ngpt);
  DevArg<Float> t1(tau1, d.n, Dir::InOut), t2(tau2, d.n, Dir::In), s2(ssa2, d.n, Dir::In);
  Float* a = t1; const Float *b = t2, *c = s2;
  launch_elementwise(d.n, [=] __device__(size_t i) { inc_1s_2s(a, b, c, i, i); });
}
void rte_increment_1scalar_by_2stream(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      const Float* tau2, const Float* ssa2) {
  OpName op_name__(__func__);
  inc_1scalar_by_scattering(ncol, nlay, ngpt, tau1, tau2, ssa2);
}
void rte_increment_1scalar_by_nstream(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      const Float* tau2, const Float* ssa2) {
  OpName op_name__(__func__);
  inc_1scalar_by_scattering(ncol, nlay, ngpt, tau1, tau2, ssa2);
}
static void inc_scattering_by_1scalar(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      Float* ssa1, const Float* tau2) {
  OpName op_name__(__func__);
  Dims d(ncol, nlay, ngpt);
  DevArg<Float> t1(tau1, d.n, Dir::InOut), s1(ssa1, d.n, Dir::InOut), t2(tau2, d.n, Dir::In);
  Float *a = t1, *b = s1; const Float* c = t2;
  launch_elementwise(d.n, [=] __device__(size_t i) { inc_2s_1s(a, b, c, i, i); });
}
void rte_increment_2stream_by_1scalar(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      Float* ssa1, const Float* tau2) {
  OpName op_name__(__func__);
  inc_scattering_by_1scalar(ncol, nlay, ngpt, tau1, ssa1, tau2);
}
void rte_increment_nstream_by_1scalar(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      Float* ssa1, const Float* tau2) {
  OpName op_name__(__func__);
  inc_scattering_by_1scalar(ncol, nlay, ngpt, tau1, ssa1, tau2);
}
void rte_increment_2stream_by_2stream(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      Float* ssa1, Float* g1, const Float* tau2, const Float* ssa2,
                                      const Float* g2) {
  OpName op_name__(__func__);
  Dims d(ncol, nlay, ngpt);
  DevArg<Float> t1(tau1, d.n, Dir::InOut), s1(ssa1, d.n, Dir::InOut), gg1(g1, d.n, Dir::InOut);
  DevArg<Float> t2(tau2, d.n, Dir::In), s2(ssa2, d.n, Dir::In), gg2(g2, d.n, Dir::In);
  Float *a = t1, *b = s1, *c = gg1; const Float *e = t2, *f = s2, *h = gg2;
  launch_elementwise(d.n, [=] __device__(size_t i) { inc_2s_2s(a, b, c, e, f, h, i, i, 1); });
}
void rte_increment_2stream_by_nstream(const int* ncol, const int* nlay, const int* ngpt, const int* nmom2,
                                      Float* tau1, Float* ssa1, Float* g1, const Float* tau2,
                                      const Float* ssa2, const Float* p2) {
  OpName op_name__(__func__);
  Dims d(ncol, nlay, ngpt);
  const size_t nm2 = (size_t)*nmom2;
  DevArg<Float> t1(tau1, d.n, Dir::InOut), s1(ssa1, d.n, Dir::InOut), gg1(g1, d.n, Dir::InOut);
  DevArg<Float> t2(tau2, d.n, Dir::In), s2(ssa2, d.n, Dir::In), pp2(p2, d.n * nm2, Dir::In);
  Float *a = t1, *b = s1, *c = gg1; const Float *e = t2, *f = s2, *h = pp2;
  launch_elementwise(d.n, [=] __device__(size_t i) { inc_2s_2s(a, b, c, e, f, h, i, i, nm2); });
}
void rte_increment_nstream_by_2stream(const int* ncol, const int* nlay, const int* ngpt, const int* nmom1,
                                      Float* tau1, Float* ssa1, Float* p1, const Float* tau2,
                                      const Float* ssa2, const Float* g2) {
  OpName op_name__(__func__);
  Dims d(ncol, nlay, ngpt);
  const int nm1 = *nmom1;
  DevArg<Float> t1(tau1, d.n, Dir::InOut), s1(ssa1, d.n, Dir::InOut), pp1(p1, d.n * nm1, Dir::InOut);
  DevArg<Float> t2(tau2, d.n, Dir::In), s2(ssa2, d.n, Dir::In), gg2(g2, d.n, Dir::In);
  Float *a = t1, *b = s1, *c = pp1; const Float *e = t2, *f = s2, *h = gg2;
  launch_elementwise(d.n, [=] __device__(size_t i) { inc_ns_2s(nm1, a, b, c, e, f, h, i, i); });
}
void rte_increment_nstream_by_nstream(const int* ncol, const int* nlay, const int* ngpt, const int* nmom1,
                                      const int* nmom2, Float* tau1, Float* ssa1, Float* p1,
                                      const Float* tau2, const Float* ssa2, const Float* p2) {
  OpName op_name__(__func__);
  Dims d(ncol, nlay, ngpt);
  const int nm1 = *nmom1, nm2 = *nmom2;
  DevArg<Float> t1(tau1, d.n, Dir::InOut), s1(ssa1, d.n, Dir::InOut), pp1(p1, d.n * nm1, Dir::InOut);
  DevArg<Float> t2(tau2, d.n, Dir::In), s2(ssa2, d.n, Dir::In), pp2(p2, d.n * nm2, Dir::In);
  Float *a = t1, *b = s1, *c = pp1; const Float *e = t2, *f = s2, *h = pp2;
  launch_elementwise(d.n, [=] __device__(size_t i) { inc_ns_ns(nm1, nm2, a, b, c, e, f, h, i, i); });
}

// ---------------- by-band increments (:366-630) ----------------
#define BYBND_PROLOGUE                                                   \
  Dims d(ncol, nlay, ngpt);                                              \
  const size_t ncl = d.ncl;                                              \
  const size_t nb = ncl * (size_t)*nbnd;                                 \
  DevArg<int> lims(gpt_lims, 2 * (size_t)*nbnd, Dir::In);                \
  int* map = build_gpt2bnd(d.ngpt, *nbnd, lims.get());

void rte_inc_1scalar_by_1scalar_bybnd(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      const Float* tau2, const int* nbnd, const int* gpt_lims) {
  OpName op_name__(__func__);
  BYBND_PROLOGUE
  {
    DevArg<Float> t1(tau1, d.n, Dir::InOut), t2(tau2, nb, Dir::In);
    Float* a = t1; const Float* b = t2;
    launch_bybnd(ncl, d.ngpt, map, [=] __device__(size_t i, size_t j) { inc_1s_1s(a, b, i, j); });
  }
  dev_free(map);
}
static void inc_1scalar_by_scattering_bybnd(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                            const Float* tau2, const Float* ssa2, const int* nbnd,
                                            const int* gpt_lims) {
  OpName op_name__(__func__);
  BYBND_PROLOGUE
  {
    DevArg<Float> t1(tau1, d.n, Dir::InOut), t2(tau2, nb, Dir::In), s2(ssa2, nb, Dir::In);
    Float* a = t1; const Float *b = t2, *c = s2;
    launch_bybnd(ncl, d.ngpt, map, [=] __device__(size_t i, size_t j) { inc_1s_2s(a, b, c, i, j); });
  }
  dev_free(map);
}
void rte_inc_1scalar_by_2stream_bybnd(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      const Float* tau2, const Float* ssa2, const int* nbnd,
                                      const int* gpt_lims) {
  OpName op_name__(__func__);
  inc_1scalar_by_scattering_bybnd(ncol, nlay, ngpt, tau1, tau2, ssa2, nbnd, gpt_lims);
}
void rte_inc_1scalar_by_nstream_bybnd(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      const Float* tau2, const Float* ssa2, const int* nbnd,
                                      const int* gpt_lims) {
  OpName op_name__(__func__);
  inc_1scalar_by_scattering_bybnd(ncol, nlay, ngpt, tau1, tau2, ssa2, nbnd, gpt_lims);
}
static void inc_scattering_by_1scalar_bybnd(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                            Float* ssa1, const Float* tau2, const int* nbnd,
                                            const int* gpt_lims) {
  OpName op_name__(__func__);
  BYBND_PROLOGUE
  {
    DevArg<Float> t1(tau1, d.n, Dir::InOut), s1(ssa1, d.n, Dir::InOut), t2(tau2, nb, Dir::In);
    Float *a = t1, *b = s1; const Float* c = t2;
    launch_bybnd(ncl, d.ngpt, map, [=] __device__(size_t i, size_t j) { inc_2s_1s(a, b, c, i, j); });
  }
  dev_free(map);
}
void rte_inc_2stream_by_1scalar_bybnd(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      Float* ssa1, const Float* tau2, const int* nbnd, const int* gpt_lims) {
  OpName op_name__(__func__);
  inc_scattering_by_1scalar_bybnd(ncol, nlay, ngpt, tau1, ssa1, tau2, nbnd, gpt_lims);
}
void rte_inc_nstream_by_1scalar_bybnd(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      Float* ssa1, const Float* tau2, const int* nbnd, const int* gpt_lims) {
  OpName op_name__(__func__);
  inc_scattering_by_1scalar_bybnd(ncol, nlay, ngpt, tau1, ssa1, tau2, nbnd, gpt_lims);
}
void rte_inc_2stream_by_2stream_bybnd(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      Float* ssa1, Float* g1, const Float* tau2, const Float* ssa2,
                                      const Float* g2, const int* nbnd, const int* gpt_lims) {
  OpName op_name__(__func__);
  BYBND_PROLOGUE
  {
    DevArg<Float> t1(tau1, d.n, Dir::InOut), s1(ssa1, d.n, Dir::InOut), gg1(g1, d.n, Dir::InOut);
    DevArg<Float> t2(tau2, nb, Dir::In), s2(ssa2, nb, Dir::In), gg2(g2, nb, Dir::In);
    Float *a = t1, *b = s1, *c = gg1; const Float *e = t2, *f = s2, *h = gg2;
    launch_bybnd(ncl, d.ngpt, map, [=] __device__(size_t i, size_t j) { inc_2s_2s(a, b, c, e, f, h, i, j, 1); });
  }
  dev_free(map);
}
void rte_inc_2stream_by_nstream_bybnd(const int* ncol, const int* nlay, const int* ngpt, const int* nmom2,
                                      Float* tau1, Float* ssa1, Float* g1, const Float* tau2,
                                      const Float* ssa2, const Float* p2, const int* nbnd,
                                      const int* gpt_lims) {
  OpName op_name__(__func__);
  BYBND_PROLOGUE
  {
    const size_t nm2 = (size_t)*nmom2;
    DevArg<Float> t1(tau1, d.n, Dir::InOut), s1(ssa1, d.n, Dir::InOut), gg1(g1, d.n, Dir::InOut);
    DevArg<Float> t2(tau2, nb, Dir::In), s2(ssa2, nb, Dir::In), pp2(p2, nb * nm2, Dir::In);
    Float *a = t1, *b = s1, *c = gg1; const Float *e = t2, *f = s2, *h = pp2;
    launch_bybnd(ncl, d.ngpt, map, [=] __device__(size_t i, size_t j) { inc_2s_2s(a, b, c, e, f, h, i, j, nm2); });
  }
  dev_free(map);
}
void rte_inc_nstream_by_2stream_bybnd(const int* ncol, const int* nlay, const int* ngpt, const int* nmom1,
                                      Float* tau1, Float* ssa1, Float* p1, const Float* tau2,
                                      const Float* ssa2, const Float* g2, const int* nbnd,
                                      const int* gpt_lims) {
  OpName op_name__(__func__);
  BYBND_PROLOGUE
  {
    const int nm1 = *nmom1;
    DevArg<Float> t1(tau1, d.n, Dir::InOut), s1(ssa1, d.n, Dir::InOut), pp1(p1, d.n * nm1, Dir::InOut);
    DevArg<Float> t2(tau2, nb, Dir::In), s2(ssa2, nb, Dir::In), gg2(g2, nb, Dir::In);
    Float *a = t1, *b = s1, *c = pp1; const Float *e = t2, *f = s2, *h = gg2;
    launch_bybnd(ncl, d.ngpt, map, [=] __device__(size_t i, size_t j) { inc_ns_2s(nm1, a, b, c, e, f, h, i, j); });
  }
  dev_free(map);
}
void rte_inc_nstream_by_nstream_bybnd(const int* ncol, const int* nlay, const int* ngpt, const int* nmom1,
                                      const int* nmom2, Float* tau1, Float* ssa1, Float* p1,
                                      const Float* tau2, const Float* ssa2, const Float* p2,
                                      const int* nbnd, const int* gpt_lims) {
  OpName op_name__(__func__);
  BYBND_PROLOGUE
  {
    const int nm1 = *nmom1, nm2 = *nmom2;
    DevArg<Float> t1(tau1, d.n, Dir::InOut), s1(ssa1, d.n, Dir::InOut), pp1(p1, d.n * nm1, Dir::InOut);
    DevArg<Float> t2(tau2, nb, Dir::In), s2(ssa2, nb, Dir::In), pp2(p2, nb * nm2, Dir::In);
    Float *a = t1, *b = s1, *c = pp1; const Float *e = t2, *f = s2, *h = pp2;
    launch_bybnd(ncl, d.ngpt, map, [=] __device__(size_t i, size_t j) { inc_ns_ns(nm1, nm2, a, b, c, e, f, h, i, j); });
  }
  dev_free(map);
}

// ---------------- subsets (:640-706) ----------------
void rte_extract_subset_dim1_3d(const int* ncol, const int* nlay, const int* ngpt, const Float* array_in,
                                const int* colS, const int* colE, Float* array_out) {
  OpName op_name__(__func__);
  const size_t nc = (size_t)*ncol, nsub = (size_t)(*colE - *colS + 1), nk = (size_t)*nlay * *ngpt;
  const size_t c0 = (size_t)(*colS - 1);
  DevArg<Float> in(array_in, nc * nk, Dir::In), out(array_out, nsub * nk, Dir::Out);
  const Float* a = in; Float* o = out;
  launch_elementwise(nsub * nk, [=] __device__(size_t i) { o[i] = a[(i % nsub) + c0 + nc * (i / nsub)]; });
}
void rte_extract_subset_dim2_4d(const int* nmom, const int* ncol, const int* nlay, const int* ngpt,
                                const Float* array_in, const int* colS, const int* colE, Float* array_out) {
  OpName op_name__(__func__);
  const size_t nm = (size_t)*nmom, nc = (size_t)*ncol, nsub = (size_t)(*colE - *colS + 1),
               nk = (size_t)*nlay * *ngpt;
  const size_t c0 = (size_t)(*colS - 1);
  DevArg<Float> in(array_in, nm * nc * nk, Dir::In), out(array_out, nm * nsub * nk, Dir::Out);
  const Float* a = in; Float* o = out;
  launch_elementwise(nm * nsub * nk, [=] __device__(size_t i) {
    const size_t m = i % nm, r = i / nm;
    o[i] = a[m + nm * ((r % nsub) + c0 + nc * (r / nsub))];
  });
}
void rte_extract_subset_absorption_tau(const int* ncol, const int* nlay, const int* ngpt,
                                       const Float* tau_in, const Float* ssa_in, const int* colS,
                                       const int* colE, Float* tau_out) {
  OpName op_name__(__func__);
  const size_t nc = (size_t)*ncol, nsub = (size_t)(*colE - *colS + 1), nk = (size_t)*nlay * *ngpt;
  const size_t c0 = (size_t)(*colS - 1);
  DevArg<Float> t(tau_in, nc * nk, Dir::In), s(ssa_in, nc * nk, Dir::In), out(tau_out, nsub * nk, Dir::Out);
  const Float *a = t, *b = s; Float* o = out;
  launch_elementwise(nsub * nk, [=] __device__(size_t i) {
    const size_t k = (i % nsub) + c0 + nc * (i / nsub);
    o[i] = a[k] * ((Float)1 - b[k]);
  });
}

}  // extern "C"
