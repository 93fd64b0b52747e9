// gas_optics_gfast.cuh - gas-optics kernels over G-POINT-FASTEST copies of the k-distribution tables.
//
// The loader's tables are (ntemp, neta, npres+1, ngpt) with the g-point OUTERMOST (stride 60 KB): the 8 entries
// one (cell, g-point) interpolates between sit in 8 different cache lines, and the next g-point needs 8 other
// lines.  Measured on B200 (profiles/r1v5_fused_*): the kernels reading that layout stall ~11 issue slots per
// issued instruction on table loads (L1 hit rate 68%, half of every fetched sector unused).
//
// Here the tables are transposed ONCE per k-distribution (table_cache, gas_optics_fused.cu) to
//       kmajorT / pfracT [row = jt + ntemp*(je + neta*jp)] [g]      kminorT [jt + ntemp*je] [kcol]
//       kraylT [itropo] [jt + ntemp*je] [g]
// so the 16 g-points of a band are ONE 128-byte line per table row: a thread reads its 8 rows with 128-bit
// loads (2 g-points each), every fetched sector is used completely, and neighbouring columns - which share
// table rows whenever their T/p/eta fall into the same table cell - are served by L1 broadcast.
//
//   gas_tau_g_kernel   thread = (cell, band): weights in registers, major + minor + Rayleigh + abs/Rayleigh
//                      combination + by-band cloud increment for the band's g-points, every output written once
//   planck_g_kernel    thread = (column, band, chunk of layers) marching down its layers with the previous
//                      layer's Planck fractions in registers (the reference's pfrac(ncol,nlay,ngpt) temporary,
//                      mo_gas_optics_rrtmgp_kernels.F90:613, never exists)
// Arithmetic is the reference's, expression by expression (lines cited inline).
#pragma once
#include "../common.cuh"
#include "rrtmgp_b200_ext.h"

namespace rrtmgpb {

constexpr int kGThreads = 128;
constexpr int kGG = 16;  // g-points per register super-chunk (one band of the standard k-distributions)

struct CellState {  // struct-of-arrays over cells
  Float *col_dry, *ftemp, *fpress;
  int *jtemp, *jpress;
  Bool* tropo;
};

struct FusedParams {
  rrtmgpb_gas_tables t;
  int ncol, nlay;
  const Float *play, *plev, *tlay, *vmr, *col_dry_in;
  CellState cs;
  const int2 *range_lower, *range_upper;
  // outputs
  int op_kind;  // 1: tau ; 2: tau, ssa, g
  Float *tau, *ssa, *g;
  // optional by-band cloud increment (kind 0 = none; 1 = 1scl tau ; 2 = 2str tau, ssa, g)
  int cld_kind;
  const Float *cld_tau, *cld_ssa, *cld_g;
};

struct PlanckFusedParams {
  FusedParams f;
  const Float *tlev, *tsfc;
  int sfc_lay;
  Float *sfc_src, *lay_src, *lev_src, *sfc_source_Jac;
};

// g-fastest table copies (device memory, owned by the table cache)
struct TablesT {
  const Float *kmajor, *pfrac, *kminor_lower, *kminor_upper, *krayl;
  int gp;        // row pitch of kmajor / pfrac / krayl (ngpt rounded up to a multiple of 2)
  int nkl, nku;  // row pitch of kminor_lower / kminor_upper
  int vec;       // 2: every band / minor interval starts on an even 0-based column and has even length
};

// ---- weights of one flavour for one cell: mo_gas_optics_rrtmgp_kernels.F90:121-168 ----
struct FlavW {
  Float cm[2], fmn[4], fmj[8];
  int je[2];
};

// col_gas(igas) = igas == 0 ? col_dry : vmr(igas)*col_dry   (mo_gas_optics_rrtmgp.F90:594-609)
__device__ __forceinline__ Float col_gas_of(const FusedParams& p, size_t c, size_t ncl, int igas, Float col_dry) {
  return igas == 0 ? col_dry : p.vmr[c + ncl * (size_t)(igas - 1)] * col_dry;
}

__device__ __forceinline__ void flavor_weights(const FusedParams& p, size_t c, size_t ncl, int iflav, int itropo,
                                               int jtemp, Float ftemp, Float fpress, Float col_dry, FlavW& w) {
  const rrtmgpb_gas_tables& t = p.t;
  const int igas_1 = __ldg(t.flavor + 2 * iflav), igas_2 = __ldg(t.flavor + 2 * iflav + 1);
  const Float cg1 = col_gas_of(p, c, ncl, igas_1, col_dry), cg2 = col_gas_of(p, c, ncl, igas_2, col_dry);
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int jt = jtemp + it;
    const Float ratio_eta_half = __ldg(t.vmr_ref + itropo + 2 * (igas_1 + (t.ngas + 1) * (jt - 1))) /
                                 __ldg(t.vmr_ref + itropo + 2 * (igas_2 + (t.ngas + 1) * (jt - 1)));
    const Float colmix = cg1 + ratio_eta_half * cg2;
    const Float eta = (colmix > (Float)2 * (Float)RB_TINY) ? cg1 / colmix : (Float)0.5;
    const Float loceta = eta * (Float)(t.neta - 1);
    w.je[it] = min((int)loceta + 1, t.neta - 1);
    const Float feta = loceta - trunc(loceta);
    const Float ftemp_term = ((Float)(1 - it) + (Float)(2 * it - 1) * ftemp);
    w.cm[it] = colmix;
    w.fmn[2 * it + 0] = ((Float)1 - feta) * ftemp_term;
    w.fmn[2 * it + 1] = feta * ftemp_term;
    w.fmj[4 * it + 0] = ((Float)1 - fpress) * w.fmn[2 * it + 0];
    w.fmj[4 * it + 1] = ((Float)1 - fpress) * w.fmn[2 * it + 1];
    w.fmj[4 * it + 2] = fpress * w.fmn[2 * it + 0];
    w.fmj[4 * it + 3] = fpress * w.fmn[2 * it + 1];
  }
}

struct MinorSet {
  int n;
  const int2* band_range;
  const Float* kminor;  // g-fastest copy
  int pitch;
  const int *limits_gpt, *idx_minor, *idx_scaling, *kminor_start;
  const Bool *scales_with_density, *scale_by_complement;
};

// VEC consecutive table entries starting at p (16-byte aligned when VEC == 2)
template <int VEC>
struct GLoad;
template <>
struct GLoad<1> {
  Float v[1];
  __device__ __forceinline__ GLoad(const Float* p) { v[0] = __ldg(p); }
};
template <>
struct GLoad<2> {
  Float v[2];
  __device__ __forceinline__ GLoad(const Float* p) {
    const Float2 x = __ldg(reinterpret_cast<const Float2*>(p));
    v[0] = x.x; v[1] = x.y;
  }
};

// 3-D interpolation (interpolate3D_byflav :791-801) of n <= kGG consecutive g-points starting at 0-based
// table column g0: out[i] = scale0*(4 terms of row set 0) + scale1*(4 terms of row set 1)
template <int VEC>
__device__ __forceinline__ void interp3d_g(const Float* __restrict__ tab, int gp, int row0, int row1, int s_eta, int s_p,
                                           int g0, int n, const Float (&f)[8], Float scale0, Float scale1,
                                           Float (&out)[kGG]) {
  const Float* a0 = tab + (size_t)row0 * gp + g0;
  const Float* a1 = a0 + (size_t)s_eta * gp;
  const Float* a2 = a0 + (size_t)s_p * gp;
  const Float* a3 = a2 + (size_t)s_eta * gp;
  const Float* b0 = tab + (size_t)row1 * gp + g0;
  const Float* b1 = b0 + (size_t)s_eta * gp;
  const Float* b2 = b0 + (size_t)s_p * gp;
  const Float* b3 = b2 + (size_t)s_eta * gp;
#pragma unroll
  for (int i = 0; i < kGG; i += VEC) {
    if (i < n) {
      const GLoad<VEC> x0(a0 + i), x1(a1 + i), x2(a2 + i), x3(a3 + i), y0(b0 + i), y1(b1 + i), y2(b2 + i), y3(b3 + i);
#pragma unroll
      for (int v = 0; v < VEC; ++v)
        out[i + v] = scale0 * (f[0] * x0.v[v] + f[1] * x1.v[v] + f[2] * x2.v[v] + f[3] * x3.v[v]) +
                     scale1 * (f[4] * y0.v[v] + f[5] * y1.v[v] + f[6] * y2.v[v] + f[7] * y3.v[v]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// tau (+ ssa, g): compute_tau_absorption :176-338, compute_tau_rayleigh :506-565, combine_abs_and_rayleigh
// (mo_gas_optics_rrtmgp.F90:1954-2002) and the by-band increment (mo_optical_props_kernels.F90:366-477)
// ---------------------------------------------------------------------------------------------------
template <bool SW, int VEC>
__global__ void __launch_bounds__(kGThreads, 4) gas_tau_g_kernel(const FusedParams p, const TablesT tt) {
  const size_t ncl = (size_t)p.ncol * p.nlay;
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncl) return;
  const rrtmgpb_gas_tables& t = p.t;
  const int ibnd = blockIdx.y;
  const int bS = __ldg(t.band_lims_gpt + 2 * ibnd), bE = __ldg(t.band_lims_gpt + 2 * ibnd + 1);
  const Float col_dry = p.cs.col_dry[c], ftemp = p.cs.ftemp[c], fpress = p.cs.fpress[c];
  const int jtemp = p.cs.jtemp[c], jpress0 = p.cs.jpress[c];
  const bool tropo = p.cs.tropo[c];
  const int itropo = tropo ? 0 : 1;
  const int jpress = jpress0 + itropo + 1;  // :390
  const int iflav = __ldg(t.gpoint_flavor + itropo + 2 * (bS - 1)) - 1;  // :384 band's first g-point
  FlavW w;
  flavor_weights(p, c, ncl, iflav, itropo, jtemp, ftemp, fpress, col_dry, w);
  const int s_eta = t.ntemp, s_p = t.ntemp * t.neta;
  const int row0 = (jtemp - 1) + s_eta * (w.je[0] - 1) + s_p * (jpress - 2);
  const int row1 = jtemp + s_eta * (w.je[1] - 1) + s_p * (jpress - 2);
  // cloud properties of this (cell, band), if the caller wants them added (mo_optical_props.F90:956-1000)
  Float ct = 0, cw = 0, cg = 0;
  if (p.cld_kind) {
    const size_t cb = c + ncl * (size_t)ibnd;
    ct = p.cld_tau[cb];
    if (p.cld_kind == 2) { cw = p.cld_ssa[cb]; cg = p.cld_g[cb]; }
  }
  const MinorSet ms = tropo ? MinorSet{t.nminorlower, p.range_lower, tt.kminor_lower, tt.nkl, t.minor_limits_gpt_lower,
                                       t.idx_minor_lower, t.idx_minor_scaling_lower, t.kminor_start_lower,
                                       t.minor_scales_with_density_lower, t.scale_by_complement_lower}
                            : MinorSet{t.nminorupper, p.range_upper, tt.kminor_upper, tt.nku, t.minor_limits_gpt_upper,
                                       t.idx_minor_upper, t.idx_minor_scaling_upper, t.kminor_start_upper,
                                       t.minor_scales_with_density_upper, t.scale_by_complement_upper};
  const int2 range = ms.band_range[ibnd];
  const Float play = p.play[c], tlay = p.tlay[c];
  const Float amount_rayl = SW ? col_gas_of(p, c, ncl, t.idx_h2o, col_dry) + col_dry : (Float)0;  // :559

  for (int gS = bS; gS <= bE; gS += kGG) {
    const int n = min(kGG, bE - gS + 1);
    Float acc[kGG];
    // ---- major absorbers: tau = 0 + major (:391 on a zeroed tau) ----
    interp3d_g<VEC>(tt.kmajor, tt.gp, row0, row1, s_eta, s_p, gS - 1, n, w.fmj, w.cm[0], w.cm[1], acc);
    // ---- minor absorbers touching this chunk (:451-498) ----
    for (int imnr = range.x; imnr <= range.y; ++imnr) {
      const int mS = __ldg(ms.limits_gpt + 2 * imnr), mE = __ldg(ms.limits_gpt + 2 * imnr + 1);
      if (mE < gS || mS > gS + n - 1) continue;
      Float scaling = col_gas_of(p, c, ncl, __ldg(ms.idx_minor + imnr), col_dry);
      if (ms.scales_with_density[imnr]) {
        scaling = scaling * ((Float)0.01 * play / tlay);
        const int isc = __ldg(ms.idx_scaling + imnr);
        if (isc > 0) {
          const Float vmr_fact = (Float)1 / col_dry;
          const Float dry_fact = (Float)1 / ((Float)1 + col_gas_of(p, c, ncl, t.idx_h2o, col_dry) * vmr_fact);
          if (ms.scale_by_complement[imnr])
            scaling = scaling * ((Float)1 - col_gas_of(p, c, ncl, isc, col_dry) * vmr_fact * dry_fact);
          else
            scaling = scaling * (col_gas_of(p, c, ncl, isc, col_dry) * vmr_fact * dry_fact);
        }
      }
      const int iflav_m = __ldg(t.gpoint_flavor + itropo + 2 * (mS - 1)) - 1;  // :487
      // the contributor's flavour is the band's flavour for rrtmgp-data (a contributor lives inside one band);
      // otherwise recompute its eta weights
      Float a0 = w.fmn[0], a1 = w.fmn[1], a2 = w.fmn[2], a3 = w.fmn[3];
      int je0 = w.je[0], je1 = w.je[1];
      if (iflav_m != iflav) {
        FlavW wm;
        flavor_weights(p, c, ncl, iflav_m, itropo, jtemp, ftemp, fpress, col_dry, wm);
        a0 = wm.fmn[0]; a1 = wm.fmn[1]; a2 = wm.fmn[2]; a3 = wm.fmn[3];
        je0 = wm.je[0]; je1 = wm.je[1];
      }
      // table column of g-point gS+i: kminor_start + (gS+i - mS) - 1
      const int kcol0 = __ldg(ms.kminor_start + imnr) + (gS - mS) - 1;
      const Float* m0 = ms.kminor + (size_t)((jtemp - 1) + s_eta * (je0 - 1)) * ms.pitch + kcol0;
      const Float* m1 = ms.kminor + (size_t)(jtemp + s_eta * (je1 - 1)) * ms.pitch + kcol0;
      const Float* m0e = m0 + (size_t)s_eta * ms.pitch;
      const Float* m1e = m1 + (size_t)s_eta * ms.pitch;
      const int iS = mS - gS, iE = min(mE - gS, n - 1);  // chunk positions covered by this contributor
#pragma unroll
      for (int i = 0; i < kGG; i += VEC) {
        if (i >= iS && i <= iE) {  // VEC == 2: intervals start even and have even length (TablesT::vec)
          const GLoad<VEC> x0(m0 + i), x1(m0e + i), y0(m1 + i), y1(m1e + i);
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const Float kint = a0 * x0.v[v] + a1 * x1.v[v] + a2 * y0.v[v] + a3 * y1.v[v];  // :757-760
            acc[i + v] = acc[i + v] + scaling * kint;                                      // :493
          }
        }
      }
    }
    // ---- Rayleigh (:554-559), combination, cloud increment, store ----
    const Float* kr = SW ? tt.krayl + (size_t)s_p * tt.gp * itropo + (gS - 1) : nullptr;
    const Float* r0 = SW ? kr + (size_t)((jtemp - 1) + s_eta * (w.je[0] - 1)) * tt.gp : nullptr;
    const Float* r1 = SW ? kr + (size_t)(jtemp + s_eta * (w.je[1] - 1)) * tt.gp : nullptr;
    const Float* r0e = SW ? r0 + (size_t)s_eta * tt.gp : nullptr;
    const Float* r1e = SW ? r1 + (size_t)s_eta * tt.gp : nullptr;
    const size_t o0 = c + ncl * (size_t)(gS - 1);
    const Float eps3 = (Float)3.0 * (Float)RB_TINY;  // mo_optical_props_kernels.F90:38
#pragma unroll
    for (int i0 = 0; i0 < kGG; i0 += VEC) {
      if (i0 >= n) continue;
      Float ray[VEC];
      if (SW) {
        const GLoad<VEC> x0(r0 + i0), x1(r0e + i0), y0(r1 + i0), y1(r1e + i0);
#pragma unroll
        for (int v = 0; v < VEC; ++v)
          ray[v] = (w.fmn[0] * x0.v[v] + w.fmn[1] * x1.v[v] + w.fmn[2] * y0.v[v] + w.fmn[3] * y1.v[v]) * amount_rayl;
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const int i = i0 + v;
        Float tt_ = acc[i], ss = 0, gg = 0;
        if (SW) {
          const Float tray = ray[v];
          tt_ = acc[i] + tray;  // combine :1986-1994
          ss = (tt_ > (Float)2 * (Float)RB_TINY) ? tray / tt_ : (Float)0;
        }
        if (p.op_kind == 1) {
          if (p.cld_kind == 1) tt_ = tt_ + ct;                         // inc_1scalar_by_1scalar_bybnd :379
          else if (p.cld_kind == 2) tt_ = tt_ + ct * ((Float)1 - cw);  // inc_1scalar_by_2stream_bybnd :398
        } else {
          if (p.cld_kind == 1) {                                       // inc_2stream_by_1scalar_bybnd :440-442
            const Float tau12 = tt_ + ct;
            ss = tt_ * ss / fmax(eps3, tau12);
            tt_ = tau12;
          } else if (p.cld_kind == 2) {                                // inc_2stream_by_2stream_bybnd :468-477
            const Float tau12 = tt_ + ct;
            const Float tauscat12 = tt_ * ss + ct * cw;
            gg = (tt_ * ss * gg + ct * cw * cg) / fmax(eps3, tauscat12);
            ss = tauscat12 / fmax(eps3, tau12);
            tt_ = tau12;
          }
        }
        p.tau[o0 + ncl * i] = tt_;
        if (p.op_kind == 2) {
          p.ssa[o0 + ncl * i] = ss;
          p.g[o0 + ncl * i] = gg;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Planck sources: compute_Planck_source :568-710
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ Float planck_band_f(const rrtmgpb_gas_tables& t, Float T, Float delta_r, const Float* tab) {
  const Float val0 = (T - t.temp_ref_min) * delta_r;  // interpolate1D :731-735
  const Float frac = val0 - trunc(val0);
  const int index = min(t.nPlanckTemp - 1, max(1, (int)val0 + 1));
  const Float t0 = __ldg(tab + index - 1), t1 = __ldg(tab + index);
  return t0 + frac * (t1 - t0);
}

// Planck fractions of n g-points (from 0-based column g0) of the band whose first g-point is bS, at cell c (:627-631)
template <int VEC>
__device__ __forceinline__ void pfrac_of_cell(const FusedParams& p, const TablesT& tt, size_t c, size_t ncl, int bS, int g0,
                                              int n, Float (&pf)[kGG]) {
  const rrtmgpb_gas_tables& t = p.t;
  const int itropo = p.cs.tropo[c] ? 0 : 1;
  const int iflav = __ldg(t.gpoint_flavor + itropo + 2 * (bS - 1)) - 1;
  const int jtemp = p.cs.jtemp[c];
  const int jpress = p.cs.jpress[c] + itropo + 1;
  FlavW w;
  flavor_weights(p, c, ncl, iflav, itropo, jtemp, p.cs.ftemp[c], p.cs.fpress[c], p.cs.col_dry[c], w);
  const int s_eta = t.ntemp, s_p = t.ntemp * t.neta;
  const int row0 = (jtemp - 1) + s_eta * (w.je[0] - 1) + s_p * (jpress - 2);
  const int row1 = jtemp + s_eta * (w.je[1] - 1) + s_p * (jpress - 2);
  interp3d_g<VEC>(tt.pfrac, tt.gp, row0, row1, s_eta, s_p, g0, n, w.fmj, (Float)1, (Float)1, pf);
}

template <int VEC>
__global__ void __launch_bounds__(kGThreads, 4) planck_g_kernel(const PlanckFusedParams q, const TablesT tt, int lay_per_chunk) {
  const FusedParams& p = q.f;
  const rrtmgpb_gas_tables& t = p.t;
  const int icol = blockIdx.x * blockDim.x + threadIdx.x;
  if (icol >= p.ncol) return;
  const int ibnd = blockIdx.y;
  const int l0 = blockIdx.z * lay_per_chunk, l1 = min(p.nlay, l0 + lay_per_chunk);
  const size_t ncol = p.ncol, ncl = ncol * p.nlay, nclp = ncol * (p.nlay + 1);
  const int bS = __ldg(t.band_lims_gpt + 2 * ibnd), bE = __ldg(t.band_lims_gpt + 2 * ibnd + 1);
  const Float delta_r = (Float)1.0 / t.totplnk_delta;
  const Float* tab = t.totplnk + (size_t)t.nPlanckTemp * ibnd;
  for (int gS = bS; gS <= bE; gS += kGG) {
    const int n = min(kGG, bE - gS + 1);
    Float pf_prev[kGG];
    if (l0 > 0) pfrac_of_cell<VEC>(p, tt, icol + ncol * (size_t)(l0 - 1), ncl, bS, gS - 1, n, pf_prev);
    for (int ilay = l0; ilay < l1; ++ilay) {
      const size_t c = icol + ncol * ilay;
      Float pf[kGG];
      pfrac_of_cell<VEC>(p, tt, c, ncl, bS, gS - 1, n, pf);
      const Float B_lay = planck_band_f(t, p.tlay[c], delta_r, tab);
      const Float B_lev = planck_band_f(t, q.tlev[c], delta_r, tab);
      const bool is_sfc = (ilay == q.sfc_lay - 1);
      Float B_sfc = 0, B_sfc1 = 0;
      if (is_sfc) {
        const Float ts = q.tsfc[icol];
        B_sfc = planck_band_f(t, ts, delta_r, tab);
        B_sfc1 = planck_band_f(t, ts + (Float)1.0, delta_r, tab);
      }
      Float* lay_c = q.lay_src + c + ncl * (size_t)(gS - 1);
      Float* lev_c = q.lev_src + c + nclp * (size_t)(gS - 1);
#pragma unroll
      for (int i = 0; i < kGG; ++i) {
        if (i < n) {
          lay_c[ncl * i] = pf[i] * B_lay;                                                  // :640
          lev_c[nclp * i] = (ilay == 0) ? pf[i] * B_lev : sqrt(pf_prev[i] * pf[i]) * B_lev;  // :695-701
          if (is_sfc) {
            q.sfc_src[icol + ncol * (size_t)(gS + i - 1)] = pf[i] * B_sfc;                  // :650-653
            q.sfc_source_Jac[icol + ncol * (size_t)(gS + i - 1)] = pf[i] * (B_sfc1 - B_sfc);
          }
          pf_prev[i] = pf[i];
        }
      }
    }
    if (l1 == p.nlay) {  // :703-705
      const Float B_top = planck_band_f(t, q.tlev[icol + ncol * p.nlay], delta_r, tab);
#pragma unroll
      for (int i = 0; i < kGG; ++i)
        if (i < n) q.lev_src[icol + ncol * p.nlay + nclp * (size_t)(gS + i - 1)] = pf_prev[i] * B_top;
    }
  }
}

// out[r*pitch + g] = in[r + nrow*g]   (one-off table transposition)
__global__ void transpose_table_kernel(const Float* __restrict__ in, Float* __restrict__ out, int nrow, int ng, int pitch) {
  __shared__ Float tile[32][33];
  const int r0 = blockIdx.x * 32, g0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + threadIdx.x, g = g0 + j;
    if (r < nrow && g < ng) tile[j][threadIdx.x] = in[(size_t)r + (size_t)nrow * g];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, g = g0 + threadIdx.x;
    if (r < nrow && g < ng) out[(size_t)r * pitch + g] = tile[threadIdx.x][j];
  }
}

}  // namespace rrtmgpb
