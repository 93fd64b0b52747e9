/* ORACLE - TEST INFRASTRUCTURE ONLY (see rte_solver_ref.c header for the rules).
 *
 * CPU restatement in plain C of /root/reference/rte/kernels/mo_optical_props_kernels.F90:
 * delta scaling (:47-98), the 9 same-resolution increments (:116-358), the 9 by-band increments
 * (:366-630) and the 3 subset extractors (:640-706).  Same C symbols as include/rte_kernels.h.
 *
 * Parity pin: reference tests/rte_optic_prop_unit_tests.F90 (increment by transparent medium,
 * half+half = whole, delta-scale identities), restated in tests/test_rte_optic_prop_unit.py.
 */
#include <float.h>
#include <stddef.h>
#include "rte_kernels.h"

/* :38  eps = 3*tiny(1._wp) */
#define REF_EPS ((Float)3.0 * (sizeof(Float) == 8 ? (Float)DBL_MIN : (Float)FLT_MIN))
#define MAXF(a, b) (((a) > (b)) ? (a) : (b))

/* delta_scale_2str_f_k :47-71 */
void rte_delta_scale_2str_f_k(const int* ncol, const int* nlay, const int* ngpt, Float* tau, Float* ssa,
                              Float* g, const Float* f) {
  const size_t n = (size_t)*ncol * *nlay * *ngpt;
  const Float eps = REF_EPS;
  for (size_t i = 0; i < n; ++i) {
    const Float wf = ssa[i] * f[i];
    tau[i] = ((Float)1 - wf) * tau[i];
    ssa[i] = (ssa[i] - wf) / MAXF(eps, ((Float)1.0 - wf));
    g[i] = (g[i] - f[i]) / MAXF(eps, ((Float)1 - f[i]));
  }
}

/* delta_scale_2str_k :76-98 (f = g*g) */
void rte_delta_scale_2str_k(const int* ncol, const int* nlay, const int* ngpt, Float* tau, Float* ssa,
                            Float* g) {
  const size_t n = (size_t)*ncol * *nlay * *ngpt;
  const Float eps = REF_EPS;
  for (size_t i = 0; i < n; ++i) {
    const Float f = g[i] * g[i];
    const Float wf = ssa[i] * f;
    tau[i] = ((Float)1 - wf) * tau[i];
    ssa[i] = (ssa[i] - wf) / MAXF(eps, ((Float)1.0 - wf));
    g[i] = (g[i] - f) / MAXF(eps, ((Float)1.0 - f));
  }
}

/* ---- same-resolution increments; i2 is the index into the second operand ---- */
/* Per-element bodies shared by the plain and the by-band variants: `i` indexes operand 1
 * (ncol,nlay,ngpt), `j` indexes operand 2 ((ncol,nlay,ngpt) or (ncol,nlay,nbnd)). */
static inline void inc_1s_1s(Float* tau1, const Float* tau2, size_t i, size_t j) {
  tau1[i] = tau1[i] + tau2[j]; /* :128 */
}
static inline void inc_1s_2s(Float* tau1, const Float* tau2, const Float* ssa2, size_t i, size_t j) {
  tau1[i] = tau1[i] + tau2[j] * ((Float)1 - ssa2[j]); /* :147-148 */
}
static inline void inc_2s_1s(Float* tau1, Float* ssa1, const Float* tau2, size_t i, size_t j) {
  const Float tau12 = tau1[i] + tau2[j]; /* :189-191 */
  ssa1[i] = tau1[i] * ssa1[i] / MAXF(REF_EPS, tau12);
  tau1[i] = tau12;
}
static inline void inc_2s_2s(Float* tau1, Float* ssa1, Float* g1, const Float* tau2, const Float* ssa2,
                             const Float* g2, size_t i, size_t j, size_t g2stride) {
  const Float tau12 = tau1[i] + tau2[j]; /* :213-222; g2stride>1 => g2 is p2(1,...) of an n-stream set */
  const Float tauscat12 = tau1[i] * ssa1[i] + tau2[j] * ssa2[j];
  g1[i] = (tau1[i] * ssa1[i] * g1[i] + tau2[j] * ssa2[j] * g2[j * g2stride]) / MAXF(REF_EPS, tauscat12);
  ssa1[i] = tauscat12 / MAXF(REF_EPS, tau12);
  tau1[i] = tau12;
}
static inline void inc_ns_2s(int nmom1, Float* tau1, Float* ssa1, Float* p1, const Float* tau2,
                             const Float* ssa2, const Float* g2, size_t i, size_t j) {
  const Float tau12 = tau1[i] + tau2[j]; /* :302-317 (Henyey-Greenstein moments g^n) */
  const Float tauscat12 = tau1[i] * ssa1[i] + tau2[j] * ssa2[j];
  Float mom = g2[j];
  for (int m = 0; m < nmom1; ++m) {
    if (m > 0) mom = mom * g2[j];
    p1[(size_t)m + (size_t)nmom1 * i] =
        (tau1[i] * ssa1[i] * p1[(size_t)m + (size_t)nmom1 * i] + tau2[j] * ssa2[j] * mom) /
        MAXF(REF_EPS, tauscat12);
  }
  ssa1[i] = tauscat12 / MAXF(REF_EPS, tau12);
  tau1[i] = tau12;
}
static inline void inc_ns_ns(int nmom1, int nmom2, Float* tau1, Float* ssa1, Float* p1, const Float* tau2,
                             const Float* ssa2, const Float* p2, size_t i, size_t j) {
  const int mom_lim = nmom1 < nmom2 ? nmom1 : nmom2; /* :338-354 */
  const Float tau12 = tau1[i] + tau2[j];
  const Float tauscat12 = tau1[i] * ssa1[i] + tau2[j] * ssa2[j];
  for (int m = 0; m < mom_lim; ++m)
    p1[(size_t)m + (size_t)nmom1 * i] = (tau1[i] * ssa1[i] * p1[(size_t)m + (size_t)nmom1 * i] +
                                         tau2[j] * ssa2[j] * p2[(size_t)m + (size_t)nmom2 * j]) /
                                        MAXF(REF_EPS, tauscat12);
  ssa1[i] = tauscat12 / MAXF(REF_EPS, tau12);
  tau1[i] = tau12;
}

#define FULL_LOOP const size_t n = (size_t)*ncol * *nlay * *ngpt; for (size_t i = 0; i < n; ++i)

void rte_increment_1scalar_by_1scalar(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      const Float* tau2) {
  FULL_LOOP inc_1s_1s(tau1, tau2, i, i);
}
void rte_increment_1scalar_by_2stream(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      const Float* tau2, const Float* ssa2) {
  FULL_LOOP inc_1s_2s(tau1, tau2, ssa2, i, i);
}
void rte_increment_1scalar_by_nstream(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      const Float* tau2, const Float* ssa2) {
  FULL_LOOP inc_1s_2s(tau1, tau2, ssa2, i, i); /* :164-171, same arithmetic */
}
void rte_increment_2stream_by_1scalar(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      Float* ssa1, const Float* tau2) {
  FULL_LOOP inc_2s_1s(tau1, ssa1, tau2, i, i);
}
void rte_increment_2stream_by_2stream(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      Float* ssa1, Float* g1, const Float* tau2, const Float* ssa2,
                                      const Float* g2) {
  FULL_LOOP inc_2s_2s(tau1, ssa1, g1, tau2, ssa2, g2, i, i, 1);
}
void rte_increment_2stream_by_nstream(const int* ncol, const int* nlay, const int* ngpt, const int* nmom2,
                                      Float* tau1, Float* ssa1, Float* g1, const Float* tau2,
                                      const Float* ssa2, const Float* p2) {
  FULL_LOOP inc_2s_2s(tau1, ssa1, g1, tau2, ssa2, p2, i, i, (size_t)*nmom2); /* :241-257 uses p2(1,..) */
}
void rte_increment_nstream_by_1scalar(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      Float* ssa1, const Float* tau2) {
  FULL_LOOP inc_2s_1s(tau1, ssa1, tau2, i, i); /* :272-281, p unchanged */
}
void rte_increment_nstream_by_2stream(const int* ncol, const int* nlay, const int* ngpt, const int* nmom1,
                                      Float* tau1, Float* ssa1, Float* p1, const Float* tau2,
                                      const Float* ssa2, const Float* g2) {
  FULL_LOOP inc_ns_2s(*nmom1, tau1, ssa1, p1, tau2, ssa2, g2, i, i);
}
void rte_increment_nstream_by_nstream(const int* ncol, const int* nlay, const int* ngpt, const int* nmom1,
                                      const int* nmom2, Float* tau1, Float* ssa1, Float* p1,
                                      const Float* tau2, const Float* ssa2, const Float* p2) {
  FULL_LOOP inc_ns_ns(*nmom1, *nmom2, tau1, ssa1, p1, tau2, ssa2, p2, i, i);
}

/* ---- by-band increments :366-630; gpt_lims(2,nbnd) holds 1-based inclusive g-point limits ---- */
#define BYBND_LOOP                                                          \
  const size_t ncl = (size_t)*ncol * *nlay;                                 \
  (void)ngpt;                                                               \
  for (int ibnd = 0; ibnd < *nbnd; ++ibnd)                                  \
    for (int igpt = gpt_lims[2 * ibnd] - 1; igpt < gpt_lims[2 * ibnd + 1]; ++igpt) \
      for (size_t c = 0; c < ncl; ++c)

#define I_ (c + ncl * (size_t)igpt)
#define J_ (c + ncl * (size_t)ibnd)

void rte_inc_1scalar_by_1scalar_bybnd(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      const Float* tau2, const int* nbnd, const int* gpt_lims) {
  BYBND_LOOP inc_1s_1s(tau1, tau2, I_, J_);
}
void rte_inc_1scalar_by_2stream_bybnd(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      const Float* tau2, const Float* ssa2, const int* nbnd,
                                      const int* gpt_lims) {
  BYBND_LOOP inc_1s_2s(tau1, tau2, ssa2, I_, J_);
}
void rte_inc_1scalar_by_nstream_bybnd(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      const Float* tau2, const Float* ssa2, const int* nbnd,
                                      const int* gpt_lims) {
  BYBND_LOOP inc_1s_2s(tau1, tau2, ssa2, I_, J_);
}
void rte_inc_2stream_by_1scalar_bybnd(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      Float* ssa1, const Float* tau2, const int* nbnd,
                                      const int* gpt_lims) {
  BYBND_LOOP inc_2s_1s(tau1, ssa1, tau2, I_, J_);
}
void rte_inc_2stream_by_2stream_bybnd(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      Float* ssa1, Float* g1, const Float* tau2, const Float* ssa2,
                                      const Float* g2, const int* nbnd, const int* gpt_lims) {
  BYBND_LOOP inc_2s_2s(tau1, ssa1, g1, tau2, ssa2, g2, I_, J_, 1);
}
void rte_inc_2stream_by_nstream_bybnd(const int* ncol, const int* nlay, const int* ngpt, const int* nmom2,
                                      Float* tau1, Float* ssa1, Float* g1, const Float* tau2,
                                      const Float* ssa2, const Float* p2, const int* nbnd,
                                      const int* gpt_lims) {
  BYBND_LOOP inc_2s_2s(tau1, ssa1, g1, tau2, ssa2, p2, I_, J_, (size_t)*nmom2);
}
void rte_inc_nstream_by_1scalar_bybnd(const int* ncol, const int* nlay, const int* ngpt, Float* tau1,
                                      Float* ssa1, const Float* tau2, const int* nbnd,
                                      const int* gpt_lims) {
  BYBND_LOOP inc_2s_1s(tau1, ssa1, tau2, I_, J_);
}
void rte_inc_nstream_by_2stream_bybnd(const int* ncol, const int* nlay, const int* ngpt, const int* nmom1,
                                      Float* tau1, Float* ssa1, Float* p1, const Float* tau2,
                                      const Float* ssa2, const Float* g2, const int* nbnd,
                                      const int* gpt_lims) {
  BYBND_LOOP inc_ns_2s(*nmom1, tau1, ssa1, p1, tau2, ssa2, g2, I_, J_);
}
void rte_inc_nstream_by_nstream_bybnd(const int* ncol, const int* nlay, const int* ngpt, const int* nmom1,
                                      const int* nmom2, Float* tau1, Float* ssa1, Float* p1,
                                      const Float* tau2, const Float* ssa2, const Float* p2,
                                      const int* nbnd, const int* gpt_lims) {
  BYBND_LOOP inc_ns_ns(*nmom1, *nmom2, tau1, ssa1, p1, tau2, ssa2, p2, I_, J_);
}

/* ---- subsets :640-706; colS/colE are 1-based inclusive ---- */
void rte_extract_subset_dim1_3d(const int* ncol, const int* nlay, const int* ngpt, const Float* array_in,
                                const int* colS, const int* colE, Float* array_out) {
  const size_t nsub = (size_t)(*colE - *colS + 1);
  for (size_t k = 0; k < (size_t)*nlay * *ngpt; ++k)
    for (size_t c = 0; c < nsub; ++c) array_out[c + nsub * k] = array_in[(c + *colS - 1) + (size_t)*ncol * k];
}
void rte_extract_subset_dim2_4d(const int* nmom, const int* ncol, const int* nlay, const int* ngpt,
                                const Float* array_in, const int* colS, const int* colE, Float* array_out) {
  const size_t nsub = (size_t)(*colE - *colS + 1), nm = (size_t)*nmom;
  for (size_t k = 0; k < (size_t)*nlay * *ngpt; ++k)
    for (size_t c = 0; c < nsub; ++c)
      for (size_t m = 0; m < nm; ++m)
        array_out[m + nm * (c + nsub * k)] = array_in[m + nm * ((c + *colS - 1) + (size_t)*ncol * k)];
}
void rte_extract_subset_absorption_tau(const int* ncol, const int* nlay, const int* ngpt,
                                       const Float* tau_in, const Float* ssa_in, const int* colS,
                                       const int* colE, Float* tau_out) {
  const size_t nsub = (size_t)(*colE - *colS + 1);
  for (size_t k = 0; k < (size_t)*nlay * *ngpt; ++k)
    for (size_t c = 0; c < nsub; ++c) {
      const size_t s = (c + *colS - 1) + (size_t)*ncol * k;
      tau_out[c + nsub * k] = tau_in[s] * ((Float)1 - ssa_in[s]);
    }
}
