// gas_optics_abi.cu - RRTMGP gas-optics and cloud-LUT kernels for sm_100a.
//
// Replaces (extern mode) rrtmgp/kernels/api/mo_gas_optics_rrtmgp_kernels.F90:16,98,170,210 and
// rrtmgp/kernels/api/mo_cloud_optics_rrtmgp_kernels.F90:18; numerics follow the default kernels
// rrtmgp/kernels/mo_gas_optics_rrtmgp_kernels.F90 (cited per function).
//
// Mapping to the hardware (all HBM-/gather-bound, no tensor cores):
//  * thread <-> (column, layer) cell with columns innermost, so every (ncol,nlay,*) load/store of a
//    warp is one or two contiguous 128/256-byte segments;
//  * blockIdx.y <-> spectral band: a thread keeps its band's interpolation weights in registers and
//    walks the band's g-points, accumulating major + minor absorbers in registers, so tau is written
//    ONCE (the reference makes three read-modify-write passes over the tau plane);
//  * k-distribution tables are read through the read-only path; they are <= 16 MB and stay L2-resident.
#include <cstdlib>
#include "../kernels/elementwise.cuh"
#include "rrtmgp_kernels.h"
#include "rrtmgp_b200_ext.h"

using namespace rrtmgpb;

namespace {

constexpr int kCellThreads = 128;
constexpr int kMaxG = 8;   // g-points handled per register chunk (wider bands are walked in chunks; 8 keeps the kernels at 64 registers = 32 warps/SM)

// ------------------------------------------------------------------------------------------
// interpolation: mo_gas_optics_rrtmgp_kernels.F90:37-170
// ------------------------------------------------------------------------------------------
struct InterpParams {
  int ncol, nlay, ngas, nflav, neta, npres, ntemp;
  const int* flavor;
  const Float *press_ref_log, *temp_ref, *vmr_ref, *play, *tlay, *col_gas;
  Float press_ref_log_delta, temp_ref_min, temp_ref_delta, press_ref_trop_log;
  int *jtemp, *jeta, *jpress;
  Float *fmajor, *fminor, *col_mix;
  Bool* tropo;
};

__global__ void __launch_bounds__(kCellThreads) interpolation_kernel(const InterpParams p) {
  const size_t ncl = (size_t)p.ncol * p.nlay;
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncl) return;
  const Float press_ref_trop = exp(p.press_ref_trop_log);                 // :99
  const Float temp_ref_delta_inv = (Float)1.0 / p.temp_ref_delta;         // :100
  const Float press_ref_log_delta_inv = (Float)1.0 / p.press_ref_log_delta;
  const Float tl = p.tlay[c], pl = p.play[c];
  // :106-108 (ftemp uses the UNCLAMPED index; only the memory access is clamped here)
  const int jtemp_ = (int)((tl - (p.temp_ref_min - p.temp_ref_delta)) * temp_ref_delta_inv);
  const int jtemp = min(p.ntemp - 1, max(1, jtemp_));
  const Float ftemp = (tl - __ldg(p.temp_ref + min(p.ntemp, max(1, jtemp_)) - 1)) * temp_ref_delta_inv;
  // :111-114
  const Float locpress = (Float)1 + (log(pl) - __ldg(p.press_ref_log)) * press_ref_log_delta_inv;
  const Float jpress_aint = fmin((Float)(p.npres - 1), fmax((Float)1.0, trunc(locpress)));
  const Float fpress = locpress - jpress_aint;
  const bool tropo = pl > press_ref_trop;                                  // :117
  p.jtemp[c] = jtemp;
  p.jpress[c] = (int)jpress_aint;
  p.tropo[c] = tropo;
  const int itropo = tropo ? 0 : 1;
  for (int iflav = 0; iflav < p.nflav; ++iflav) {                          // :121-168
    const int igas_1 = __ldg(p.flavor + 2 * iflav), igas_2 = __ldg(p.flavor + 2 * iflav + 1);
    const Float cg1 = p.col_gas[c + ncl * igas_1], cg2 = p.col_gas[c + ncl * igas_2];
    const size_t cf = c + ncl * iflav;
    Float fmn[4], fmj[8], cm[2];
    int je[2];
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int jt = jtemp + it;  // 1-based reference temperature index
      const Float ratio_eta_half = __ldg(p.vmr_ref + itropo + 2 * ((size_t)igas_1 + (size_t)(p.ngas + 1) * (jt - 1))) /
                                   __ldg(p.vmr_ref + itropo + 2 * ((size_t)igas_2 + (size_t)(p.ngas + 1) * (jt - 1)));
      const Float colmix = cg1 + ratio_eta_half * cg2;
      const Float eta = (colmix > (Float)2 * (Float)RB_TINY) ? cg1 / colmix : (Float)0.5;  // :147-151
      const Float loceta = eta * (Float)(p.neta - 1);
      je[it] = min((int)loceta + 1, p.neta - 1);
      const Float feta = loceta - trunc(loceta);
      const Float ftemp_term = ((Float)(1 - it) + (Float)(2 * it - 1) * ftemp);  // :157
      cm[it] = colmix;
      fmn[2 * it + 0] = ((Float)1 - feta) * ftemp_term;
      fmn[2 * it + 1] = feta * ftemp_term;
      fmj[4 * it + 0] = ((Float)1 - fpress) * fmn[2 * it + 0];
      fmj[4 * it + 1] = ((Float)1 - fpress) * fmn[2 * it + 1];
      fmj[4 * it + 2] = fpress * fmn[2 * it + 0];
      fmj[4 * it + 3] = fpress * fmn[2 * it + 1];
    }
    // reference layouts: col_mix/jeta (2,col,lay,flav), fminor (2,2,...), fmajor (2,2,2,...): each
    // thread owns 16/8/32/64 contiguous bytes -> 128-bit stores
    reinterpret_cast<Float2*>(p.col_mix)[cf] = make_Float2(cm[0], cm[1]);
    reinterpret_cast<int2*>(p.jeta)[cf] = make_int2(je[0], je[1]);
    Float2* fm = reinterpret_cast<Float2*>(p.fminor) + 2 * cf;
    fm[0] = make_Float2(fmn[0], fmn[1]);
    fm[1] = make_Float2(fmn[2], fmn[3]);
    Float2* fj = reinterpret_cast<Float2*>(p.fmajor) + 4 * cf;
#pragma unroll
    for (int k = 0; k < 4; ++k) fj[k] = make_Float2(fmj[2 * k], fmj[2 * k + 1]);
  }
}

// ------------------------------------------------------------------------------------------
// compute_tau_absorption: major (:345-396, interpolate3D_byflav :765-803) + minor (:402-501,
// interpolate2D_byflav :741-763) fused; one thread per (cell, band).
// Minor-gas layer ranges: the reference derives per-column [first,last] layers from minloc/maxloc
// of play under the tropo mask (:274-285).  For pressure monotonic in height (which the frontend's
// own top_at_1 logic assumes) that is exactly the per-layer predicate tropo / .not.tropo, used here.
// ------------------------------------------------------------------------------------------
struct MinorTables {
  int nminor;
  const int2* band_range;  // per band: [first, last] contributor whose g-point range touches the band (last < first: none)
  const Float* kminor;
  const int *limits_gpt, *idx_minor, *idx_scaling, *kminor_start;
  const Bool *scales_with_density, *scale_by_complement;
};

struct TauAbsParams {
  int ncol, nlay, nbnd, ngpt, ngas, nflav, neta, npres, ntemp, idx_h2o;
  const int *gpoint_flavor, *band_lims_gpt;
  const Float* kmajor;
  MinorTables lower, upper;
  const Bool* tropo;
  const Float *col_mix, *fmajor, *fminor, *play, *tlay, *col_gas;
  const int *jeta, *jtemp, *jpress;
  Float* tau;
  int accumulate;  // 1: tau += (reference contract, caller pre-zeroes); 0: tau = (fused frontend)
  // GF kernels: kmajor / kminor point at g-point-fastest copies (kernels/gas_optics_gfast.cuh) with these row pitches
  int gp, pitch_lower, pitch_upper;
};

// Table strides.  The k-distribution tables of rrtmgp-data all have ntemp = 14, neta = 9, npres+1 = 60;
// with those as compile-time constants every one of the 16 x (gpts per band) table loads of a cell is
// `base pointer + immediate` (no per-load 64-bit address arithmetic - the first version of this kernel spent
// 55% of its issue slots on IMAD/LEA/IADD3, profiles/r1_prof_v1_tile_kernels.txt).  <0,0,0> = runtime dims.
template <int NT, int NE, int NP1>
struct TableDims {
  int nt, ne, np1;
  __device__ __forceinline__ TableDims(int ntemp, int neta, int npres) : nt(NT ? NT : ntemp), ne(NE ? NE : neta), np1(NP1 ? NP1 : npres + 1) {}
  __device__ __forceinline__ int s_eta() const { return NT ? NT : nt; }
  __device__ __forceinline__ int s_p() const { return (NT && NE) ? NT * NE : nt * ne; }
  __device__ __forceinline__ int s_g() const { return (NT && NE && NP1) ? NT * NE * NP1 : nt * ne * np1; }
};

template <int NT, int NE, int NP1, bool GF>
__device__ __forceinline__ void minor_contrib(const MinorTables& m, const int itropo, const TauAbsParams& p,
                                              const TableDims<NT, NE, NP1>& td, size_t c, size_t ncl, int ibnd,
                                              int gS, int gE, Float (&acc)[kMaxG]) {
  // chunk covers 1-based g-points gS..gE
  const Float play = p.play[c], tlay = p.tlay[c];
  const int jtemp = p.jtemp[c];
  const int s_eta = td.s_eta(), s_k = td.s_p();
  // only the contributors that touch this band (scanning all ~30-60 of them per chunk was the longest
  // dependent-load chain of the first version of this kernel)
  const int2 range = m.band_range[ibnd];
  for (int imnr = range.x; imnr <= range.y; ++imnr) {
    const int mS = __ldg(m.limits_gpt + 2 * imnr), mE = __ldg(m.limits_gpt + 2 * imnr + 1);
    if (mE < gS || mS > gE) continue;
    Float scaling = p.col_gas[c + ncl * __ldg(m.idx_minor + imnr)];                       // :461
    if (m.scales_with_density[imnr]) {                                                    // :465-480
      scaling = scaling * ((Float)0.01 * play / tlay);
      const int isc = __ldg(m.idx_scaling + imnr);
      if (isc > 0) {
        const Float vmr_fact = (Float)1 / p.col_gas[c];
        const Float dry_fact = (Float)1 / ((Float)1 + p.col_gas[c + ncl * p.idx_h2o] * vmr_fact);
        if (m.scale_by_complement[imnr])
          scaling = scaling * ((Float)1 - p.col_gas[c + ncl * isc] * vmr_fact * dry_fact);
        else
          scaling = scaling * (p.col_gas[c + ncl * isc] * vmr_fact * dry_fact);
      }
    }
    // flavour of the contributor's FIRST g-point (:487); gpoint_flavor(itropo, gpt)
    const int iflav = __ldg(p.gpoint_flavor + itropo + 2 * (mS - 1)) - 1;
    const size_t cf = c + ncl * iflav;
    const Float2 f01 = reinterpret_cast<const Float2*>(p.fminor)[2 * cf];
    const Float2 f23 = reinterpret_cast<const Float2*>(p.fminor)[2 * cf + 1];
    const int2 je = reinterpret_cast<const int2*>(p.jeta)[cf];
    // table column of chunk slot i is kstart + (gS + i - mS) - 1 = kcol0 + i
    const long long kcol0 = (long long)__ldg(m.kminor_start + imnr) + (gS - mS) - 1;
    if (GF) {  // g-point-fastest copy: row (jt + ntemp*je), consecutive g-points are consecutive addresses
      const int pitch = itropo ? p.pitch_upper : p.pitch_lower;
      const Float* k0 = m.kminor + (size_t)((jtemp - 1) + s_eta * (je.x - 1)) * pitch + kcol0;
      const Float* k1 = m.kminor + (size_t)(jtemp + s_eta * (je.y - 1)) * pitch + kcol0;
      const size_t de = (size_t)s_eta * pitch;
#pragma unroll
      for (int i = 0; i < kMaxG; ++i) {
        const int g = gS + i;
        if (g >= mS && g <= mE && g <= gE) {
          const Float kint = f01.x * __ldg(k0 + i) + f01.y * __ldg(k0 + de + i) +
                             f23.x * __ldg(k1 + i) + f23.y * __ldg(k1 + de + i);              // :757-760
          acc[i] = acc[i] + scaling * kint;                                                   // :493
        }
      }
      continue;
    }
    const Float* k0 = m.kminor + (jtemp - 1) + s_eta * (je.x - 1) + (long long)s_k * kcol0;
    const Float* k1 = m.kminor + jtemp + s_eta * (je.y - 1) + (long long)s_k * kcol0;
#pragma unroll
    for (int i = 0; i < kMaxG; ++i) {
      const int g = gS + i;
      if (g >= mS && g <= mE && g <= gE) {
        const int ko = s_k * i;
        const Float kint = f01.x * __ldg(k0 + ko) + f01.y * __ldg(k0 + ko + s_eta) +
                           f23.x * __ldg(k1 + ko) + f23.y * __ldg(k1 + ko + s_eta);          // :757-760
        acc[i] = acc[i] + scaling * kint;                                                   // :493
      }
    }
  }
}

template <int NT, int NE, int NP1, bool GF = false>
__global__ void __launch_bounds__(kCellThreads, 6) tau_absorption_kernel(const TauAbsParams p) {
  const size_t ncl = (size_t)p.ncol * p.nlay;
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncl) return;
  const TableDims<NT, NE, NP1> td(p.ntemp, p.neta, p.npres);
  const int s_eta = td.s_eta(), s_p = td.s_p(), s_g = td.s_g();
  const int ibnd = blockIdx.y;
  const int bS = __ldg(p.band_lims_gpt + 2 * ibnd), bE = __ldg(p.band_lims_gpt + 2 * ibnd + 1);
  const bool tropo = p.tropo[c];
  const int itropo = tropo ? 0 : 1;
  const int iflav = __ldg(p.gpoint_flavor + itropo + 2 * (bS - 1)) - 1;  // :384 band's first g-point
  const size_t cf = c + ncl * iflav;
  const Float2 cm = reinterpret_cast<const Float2*>(p.col_mix)[cf];
  const Float2* fj = reinterpret_cast<const Float2*>(p.fmajor) + 4 * cf;
  const Float2 f0 = fj[0], f1 = fj[1], f2 = fj[2], f3 = fj[3];
  const int2 je = reinterpret_cast<const int2*>(p.jeta)[cf];
  const int jtemp = p.jtemp[c];
  const int jpress = p.jpress[c] + itropo + 1;  // :390 jpress + itropo (itropo 1/2 in the reference)
  for (int gS = bS; gS <= bE; gS += kMaxG) {
    const int gE = min(bE, gS + kMaxG - 1);
    // k(jtemp, jeta1, jpress-1, gS) and k(jtemp+1, jeta2, jpress-1, gS): everything else is an immediate
    Float* tau_c = p.tau + c + ncl * (size_t)(gS - 1);
    Float acc[kMaxG];
    if (GF) {
      // g-point-fastest copy: row = jt + ntemp*(je + neta*jp); the chunk's g-points are consecutive addresses of a
      // row, so neighbouring g-points share sectors and neighbouring columns share lines
      const size_t gp = (size_t)p.gp, de = (size_t)s_eta * gp, dp = (size_t)s_p * gp;
      const Float* k0 = p.kmajor + (size_t)((jtemp - 1) + s_eta * (je.x - 1) + s_p * (jpress - 2)) * gp + (gS - 1);
      const Float* k1 = p.kmajor + (size_t)(jtemp + s_eta * (je.y - 1) + s_p * (jpress - 2)) * gp + (gS - 1);
#pragma unroll
      for (int i = 0; i < kMaxG; ++i) {
        if (gS + i <= gE) {
          const Float major =  // interpolate3D_byflav :791-801, same association
              cm.x * (f0.x * __ldg(k0 + i) + f0.y * __ldg(k0 + de + i) + f1.x * __ldg(k0 + dp + i) + f1.y * __ldg(k0 + dp + de + i)) +
              cm.y * (f2.x * __ldg(k1 + i) + f2.y * __ldg(k1 + de + i) + f3.x * __ldg(k1 + dp + i) + f3.y * __ldg(k1 + dp + de + i));
          const Float t0 = p.accumulate ? tau_c[ncl * i] : (Float)0;
          acc[i] = t0 + major;                                                               // :391
        } else {
          acc[i] = 0;
        }
      }
    } else {
      const Float* k0 = p.kmajor + (jtemp - 1) + s_eta * (je.x - 1) + s_p * (jpress - 2) + (long long)s_g * (gS - 1);
      const Float* k1 = p.kmajor + jtemp + s_eta * (je.y - 1) + s_p * (jpress - 2) + (long long)s_g * (gS - 1);
#pragma unroll
      for (int i = 0; i < kMaxG; ++i) {
        if (gS + i <= gE) {
          const int go = s_g * i;
          // interpolate3D_byflav :791-801, same association
          const Float major =
              cm.x * (f0.x * __ldg(k0 + go) + f0.y * __ldg(k0 + go + s_eta) +
                      f1.x * __ldg(k0 + go + s_p) + f1.y * __ldg(k0 + go + s_p + s_eta)) +
              cm.y * (f2.x * __ldg(k1 + go) + f2.y * __ldg(k1 + go + s_eta) +
                      f3.x * __ldg(k1 + go + s_p) + f3.y * __ldg(k1 + go + s_p + s_eta));
          const Float t0 = p.accumulate ? tau_c[ncl * i] : (Float)0;
          acc[i] = t0 + major;                                                                 // :391
        } else {
          acc[i] = 0;
        }
      }
    }
    if (tropo) minor_contrib<NT, NE, NP1, GF>(p.lower, 0, p, td, c, ncl, ibnd, gS, gE, acc);
    else       minor_contrib<NT, NE, NP1, GF>(p.upper, 1, p, td, c, ncl, ibnd, gS, gE, acc);
#pragma unroll
    for (int i = 0; i < kMaxG; ++i)
      if (gS + i <= gE) tau_c[ncl * i] = acc[i];
  }
}

// ------------------------------------------------------------------------------------------
// compute_tau_rayleigh :506-565 ; krayl(ntemp,neta,ngpt,2)
// ------------------------------------------------------------------------------------------
struct RaylParams {
  int ncol, nlay, nbnd, ngpt, neta, ntemp, idx_h2o;
  const int *gpoint_flavor, *band_lims_gpt, *jeta, *jtemp;
  const Float *krayl, *col_dry, *col_gas, *fminor;
  const Bool* tropo;
  Float* tau_rayleigh;
};

__global__ void __launch_bounds__(kCellThreads) tau_rayleigh_kernel(const RaylParams p) {
  const size_t ncl = (size_t)p.ncol * p.nlay;
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncl) return;
  const int ibnd = blockIdx.y;
  const int bS = __ldg(p.band_lims_gpt + 2 * ibnd), bE = __ldg(p.band_lims_gpt + 2 * ibnd + 1);
  const int itropo = p.tropo[c] ? 0 : 1;
  const int iflav = __ldg(p.gpoint_flavor + itropo + 2 * (bS - 1)) - 1;
  const size_t cf = c + ncl * iflav;
  const Float2 f01 = reinterpret_cast<const Float2*>(p.fminor)[2 * cf];
  const Float2 f23 = reinterpret_cast<const Float2*>(p.fminor)[2 * cf + 1];
  const int2 je = reinterpret_cast<const int2*>(p.jeta)[cf];
  const int jtemp = p.jtemp[c];
  const size_t s_eta = (size_t)p.ntemp, s_k = (size_t)p.ntemp * p.neta;
  const Float* kr = p.krayl + s_k * (size_t)p.ngpt * itropo;
  const Float* k0 = kr + (size_t)(jtemp - 1) + s_eta * (size_t)(je.x - 1);
  const Float* k1 = kr + (size_t)jtemp + s_eta * (size_t)(je.y - 1);
  const Float amount = p.col_gas[c + ncl * p.idx_h2o] + p.col_dry[c];  // :559
  for (int g = bS; g <= bE; ++g) {
    const size_t ko = s_k * (size_t)(g - 1);
    const Float k = f01.x * __ldg(k0 + ko) + f01.y * __ldg(k0 + ko + s_eta) +
                    f23.x * __ldg(k1 + ko) + f23.y * __ldg(k1 + ko + s_eta);
    p.tau_rayleigh[c + ncl * (size_t)(g - 1)] = k * amount;
  }
}

// ------------------------------------------------------------------------------------------
// compute_Planck_source :568-710.  One thread per (column, band) marching through the layers, so
// the Planck fraction of the previous layer (needed for lev_src = sqrt(pfrac(l-1)*pfrac(l))*B, :699)
// stays in registers and the pfrac(ncol,nlay,ngpt) temporary of the reference (:613) never exists.
// ------------------------------------------------------------------------------------------
struct PlanckParams {
  int ncol, nlay, nbnd, ngpt, nflav, neta, npres, ntemp, nPlanckTemp, sfc_lay;
  const Float *tlay, *tlev, *tsfc, *fmajor, *pfracin, *totplnk;
  const int *jeta, *jtemp, *jpress, *band_lims_gpt, *gpoint_flavor;
  const Bool* tropo;
  Float temp_ref_min, totplnk_delta;
  Float *sfc_src, *lay_src, *lev_src, *sfc_source_Jac;
};

__device__ __forceinline__ Float planck_band(const PlanckParams& p, Float T, Float delta_r, const Float* tab) {
  // interpolate1D :731-735 for ONE band column of totplnk(nPlanckTemp, nbnd)
  const Float val0 = (T - p.temp_ref_min) * delta_r;
  const Float frac = val0 - trunc(val0);
  const int index = min(p.nPlanckTemp - 1, max(1, (int)val0 + 1));
  const Float t0 = __ldg(tab + index - 1), t1 = __ldg(tab + index);
  return t0 + frac * (t1 - t0);
}

template <int NT, int NE, int NP1>
__global__ void __launch_bounds__(kCellThreads, 8) planck_source_kernel(const PlanckParams p) {
  const int icol = blockIdx.x * blockDim.x + threadIdx.x;
  if (icol >= p.ncol) return;
  const TableDims<NT, NE, NP1> td(p.ntemp, p.neta, p.npres);
  const int s_eta = td.s_eta(), s_p = td.s_p(), s_g = td.s_g();
  const int ibnd = blockIdx.y;
  const size_t ncol = p.ncol, ncl = ncol * p.nlay, nclp = ncol * (p.nlay + 1);
  const int bS = __ldg(p.band_lims_gpt + 2 * ibnd), bE = __ldg(p.band_lims_gpt + 2 * ibnd + 1);
  const Float delta_r = (Float)1.0 / p.totplnk_delta;  // :636
  const Float* tab = p.totplnk + (size_t)p.nPlanckTemp * ibnd;
  for (int gS = bS; gS <= bE; gS += kMaxG) {
    const int gE = min(bE, gS + kMaxG - 1);
    Float pf_prev[kMaxG];
#pragma unroll
    for (int i = 0; i < kMaxG; ++i) pf_prev[i] = 0;
    for (int ilay = 0; ilay < p.nlay; ++ilay) {
      const size_t c = icol + ncol * ilay;
      const int itropo = p.tropo[c] ? 0 : 1;
      const int iflav = __ldg(p.gpoint_flavor + itropo + 2 * (bS - 1)) - 1;  // :625
      const size_t cf = c + ncl * iflav;
      const Float2* fj = reinterpret_cast<const Float2*>(p.fmajor) + 4 * cf;
      const Float2 f0 = fj[0], f1 = fj[1], f2 = fj[2], f3 = fj[3];
      const int2 je = reinterpret_cast<const int2*>(p.jeta)[cf];
      const int jtemp = p.jtemp[c];
      const int jpress = p.jpress[c] + itropo + 1;  // :630
      const Float* k0 = p.pfracin + (jtemp - 1) + s_eta * (je.x - 1) + s_p * (jpress - 2) + (long long)s_g * (gS - 1);
      const Float* k1 = p.pfracin + jtemp + s_eta * (je.y - 1) + s_p * (jpress - 2) + (long long)s_g * (gS - 1);
      const Float B_lay = planck_band(p, p.tlay[c], delta_r, tab);          // :661
      const Float B_lev = planck_band(p, p.tlev[c], delta_r, tab);          // :683 (level ilay)
      const bool is_sfc = (ilay == p.sfc_lay - 1);
      Float B_sfc = 0, B_sfc1 = 0;
      if (is_sfc) {                                                          // :642-643
        const Float ts = p.tsfc[icol];
        B_sfc = planck_band(p, ts, delta_r, tab);
        B_sfc1 = planck_band(p, ts + (Float)1.0, delta_r, tab);
      }
      Float* lay_c = p.lay_src + c + ncl * (size_t)(gS - 1);
      Float* lev_c = p.lev_src + c + nclp * (size_t)(gS - 1);
#pragma unroll
      for (int i = 0; i < kMaxG; ++i) {
        if (gS + i <= gE) {
          const int go = s_g * i;
          // interpolate3D_byflav with scaling = (1,1) :627-631
          const Float pf = (Float)1 * (f0.x * __ldg(k0 + go) + f0.y * __ldg(k0 + go + s_eta) +
                                       f1.x * __ldg(k0 + go + s_p) + f1.y * __ldg(k0 + go + s_p + s_eta)) +
                           (Float)1 * (f2.x * __ldg(k1 + go) + f2.y * __ldg(k1 + go + s_eta) +
                                       f3.x * __ldg(k1 + go + s_p) + f3.y * __ldg(k1 + go + s_p + s_eta));
          lay_c[ncl * i] = pf * B_lay;                                                         // :674
          lev_c[nclp * i] = (ilay == 0) ? pf * B_lev : sqrt(pf_prev[i] * pf) * B_lev;          // :695,699
          if (is_sfc) {                                                                        // :651-653
            p.sfc_src[icol + ncol * (size_t)(gS + i - 1)] = pf * B_sfc;
            p.sfc_source_Jac[icol + ncol * (size_t)(gS + i - 1)] = pf * (B_sfc1 - B_sfc);
          }
          pf_prev[i] = pf;
        }
      }
    }
    const Float B_top = planck_band(p, p.tlev[icol + ncol * p.nlay], delta_r, tab);           // :705
#pragma unroll
    for (int i = 0; i < kMaxG; ++i)
      if (gS + i <= gE) p.lev_src[icol + ncol * p.nlay + nclp * (size_t)(gS + i - 1)] = pf_prev[i] * B_top;
  }
}

// ------------------------------------------------------------------------------------------
// compute_cld_from_table: mo_cloud_optics_rrtmgp_kernels.F90:24-65
// ------------------------------------------------------------------------------------------
struct CldParams {
  int ncol, nlay, ngpt, nsteps;
  const Bool* mask;
  const Float *lwp, *re, *tau_table, *ssa_table, *asy_table;
  Float step_size, offset;
  Float *tau, *taussa, *taussag;
};

__global__ void __launch_bounds__(kCellThreads) cld_from_table_kernel(const CldParams p) {
  const size_t ncl = (size_t)p.ncol * p.nlay;
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncl) return;
  const bool m = p.mask[c];
  int index = 1;
  Float fint = 0, lwp = 0;
  if (m) {
    const Float re = p.re[c];
    lwp = p.lwp[c];
    index = min((int)floor((re - p.offset) / p.step_size) + 1, p.nsteps - 1);   // :46
    fint = (re - p.offset) / p.step_size - (Float)(index - 1);                   // :47
  }
  for (int g = 0; g < p.ngpt; ++g) {
    const size_t o = c + ncl * g;
    Float t = 0, ts = 0, tsg = 0;
    if (m) {
      const Float* tt = p.tau_table + (size_t)p.nsteps * g + (index - 1);
      const Float* st = p.ssa_table + (size_t)p.nsteps * g + (index - 1);
      const Float* at = p.asy_table + (size_t)p.nsteps * g + (index - 1);
      t = lwp * (__ldg(tt) + fint * (__ldg(tt + 1) - __ldg(tt)));
      ts = t * (__ldg(st) + fint * (__ldg(st + 1) - __ldg(st)));
      tsg = ts * (__ldg(at) + fint * (__ldg(at + 1) - __ldg(at)));
    }
    p.taussag[o] = tsg;
    p.taussa[o] = ts;
    p.tau[o] = t;
  }
}

// per band: first/last minor contributor whose g-point range intersects the band
__global__ void minor_band_ranges_kernel(int nbnd, const int* band_lims_gpt, int nminor, const int* limits_gpt,
                                         int2* out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbnd) return;
  const int bS = band_lims_gpt[2 * b], bE = band_lims_gpt[2 * b + 1];
  int first = nminor, last = -1;
  for (int i = 0; i < nminor; ++i) {
    const int mS = limits_gpt[2 * i], mE = limits_gpt[2 * i + 1];
    if (mE >= bS && mS <= bE) { first = min(first, i); last = max(last, i); }
  }
  out[b] = make_int2(first, last);
}

void tau_absorption_impl(int ncol, int nlay, int nbnd, int ngpt, int ngas, int nflav, int neta, int npres,
                         int ntemp, int nminorlower, int nminorklower, int nminorupper, int nminorkupper,
                         int idx_h2o, const int* gpoint_flavor, const int* band_lims_gpt, const Float* kmajor,
                         const Float* kminor_lower, const Float* kminor_upper, const int* minor_limits_gpt_lower,
                         const int* minor_limits_gpt_upper, const Bool* minor_scales_with_density_lower,
                         const Bool* minor_scales_with_density_upper, const Bool* scale_by_complement_lower,
                         const Bool* scale_by_complement_upper, const int* idx_minor_lower,
                         const int* idx_minor_upper, const int* idx_minor_scaling_lower,
                         const int* idx_minor_scaling_upper, const int* kminor_start_lower,
                         const int* kminor_start_upper, const Bool* tropo, const Float* col_mix,
                         const Float* fmajor, const Float* fminor, const Float* play, const Float* tlay,
                         const Float* col_gas, const int* jeta, const int* jtemp, const int* jpress, Float* tau,
                         bool accumulate) {
  const size_t ncl = (size_t)ncol * nlay, tn = (size_t)ntemp * neta;
  DevArg<int> a_gf(gpoint_flavor, 2 * (size_t)ngpt, Dir::In), a_bl(band_lims_gpt, 2 * (size_t)nbnd, Dir::In);
  DevArg<Float> a_km(kmajor, tn * (npres + 1) * ngpt, Dir::In), a_kl(kminor_lower, tn * nminorklower, Dir::In),
      a_ku(kminor_upper, tn * nminorkupper, Dir::In);
  DevArg<int> a_ll(minor_limits_gpt_lower, 2 * (size_t)nminorlower, Dir::In),
      a_lu(minor_limits_gpt_upper, 2 * (size_t)nminorupper, Dir::In);
  DevArg<Bool> a_sdl(minor_scales_with_density_lower, nminorlower, Dir::In),
      a_sdu(minor_scales_with_density_upper, nminorupper, Dir::In),
      a_scl(scale_by_complement_lower, nminorlower, Dir::In), a_scu(scale_by_complement_upper, nminorupper, Dir::In);
  DevArg<int> a_iml(idx_minor_lower, nminorlower, Dir::In), a_imu(idx_minor_upper, nminorupper, Dir::In),
      a_isl(idx_minor_scaling_lower, nminorlower, Dir::In), a_isu(idx_minor_scaling_upper, nminorupper, Dir::In),
      a_ksl(kminor_start_lower, nminorlower, Dir::In), a_ksu(kminor_start_upper, nminorupper, Dir::In);
  DevArg<Bool> a_tr(tropo, ncl, Dir::In);
  // (16- / 8-byte alignment: the kernels read these with 128- / 64-bit loads; see DevArg)
  DevArg<Float> a_cm(col_mix, 2 * ncl * nflav, Dir::In, true, 16), a_fj(fmajor, 8 * ncl * nflav, Dir::In, true, 16),
      a_fn(fminor, 4 * ncl * nflav, Dir::In, true, 16), a_pl(play, ncl, Dir::In), a_tl(tlay, ncl, Dir::In),
      a_cg(col_gas, ncl * (ngas + 1), Dir::In);
  DevArg<int> a_je(jeta, 2 * ncl * nflav, Dir::In, true, 8), a_jt(jtemp, ncl, Dir::In), a_jp(jpress, ncl, Dir::In);
  DevArg<Float> a_tau(tau, ncl * ngpt, accumulate ? Dir::InOut : Dir::Out);
  // Device-resident tables whose copies the caller allows to be cached (rrtmgpb_abi_table_cache): the g-point-fastest
  // kernels of the fused path in their ABI instantiation (kernels/gas_optics_gfast.cuh: two cells per thread sharing table
  // loads, 128-bit loads, warp-level TMA staging of shared rows) - same arithmetic, the interpolation state read from
  // the caller's arrays.  B200, 65,536 x 72: LW + SW absorption 15.9 -> see DESIGN.md section 4.
  static const bool gfast_env = [] { const char* e = std::getenv("RRTMGPB_ABI_GFAST"); return !(e && e[0] == '0'); }();
  if (gfast_env && !a_km.staged() && !a_kl.staged() && !a_ku.staged() && !a_gf.staged() && !a_bl.staged() && !a_ll.staged() &&
      !a_lu.staged() &&
      tau_absorption_gfast(ncol, nlay, nbnd, ngpt, ngas, nflav, neta, npres, ntemp, nminorlower, nminorklower, nminorupper,
                           nminorkupper, idx_h2o, a_gf, a_bl, a_km, a_kl, a_ku, a_ll, a_lu, a_sdl, a_sdu, a_scl, a_scu, a_iml,
                           a_imu, a_isl, a_isu, a_ksl, a_ksu, a_tr, a_cm, a_fj, a_fn, a_pl, a_tl, a_cg, a_je, a_jt, a_jp,
                           a_tau, accumulate))
    return;
  TauAbsParams p;
  p.ncol = ncol; p.nlay = nlay; p.nbnd = nbnd; p.ngpt = ngpt; p.ngas = ngas; p.nflav = nflav; p.neta = neta;
  p.npres = npres; p.ntemp = ntemp; p.idx_h2o = idx_h2o;
  p.gpoint_flavor = a_gf; p.band_lims_gpt = a_bl; p.kmajor = a_km;
  int2* ranges = static_cast<int2*>(dev_alloc(sizeof(int2) * 2 * (size_t)nbnd));
  {
    KernelTimer timer("minor_band_ranges");
    minor_band_ranges_kernel<<<ceil_div(nbnd, 32), 32, 0, stream()>>>(nbnd, a_bl, nminorlower, a_ll, ranges);
    RB_LAUNCH_CHECK();
    minor_band_ranges_kernel<<<ceil_div(nbnd, 32), 32, 0, stream()>>>(nbnd, a_bl, nminorupper, a_lu, ranges + nbnd);
    RB_LAUNCH_CHECK();
  }
  p.lower = MinorTables{nminorlower, ranges, a_kl, a_ll, a_iml, a_isl, a_ksl, a_sdl, a_scl};
  p.upper = MinorTables{nminorupper, ranges + nbnd, a_ku, a_lu, a_imu, a_isu, a_ksu, a_sdu, a_scu};
  p.tropo = a_tr; p.col_mix = a_cm; p.fmajor = a_fj; p.fminor = a_fn; p.play = a_pl; p.tlay = a_tl;
  p.col_gas = a_cg; p.jeta = a_je; p.jtemp = a_jt; p.jpress = a_jp; p.tau = a_tau;
  p.accumulate = accumulate ? 1 : 0;
  dim3 grid(ceil_div((long long)ncl, kCellThreads), nbnd);
  // g-point-fastest table copies when the caller allows caching them (device-resident tables only: a staged host
  // table gets a fresh device address on every call)
  const Float *kmT = nullptr, *klT = nullptr, *kuT = nullptr;
  p.gp = p.pitch_lower = p.pitch_upper = 0;
  const bool gf = !a_km.staged() && !a_kl.staged() && !a_ku.staged() &&
                  tables_gfast_abi(a_km, a_kl, a_ku, ntemp, neta, npres, ngpt, nminorklower, nminorkupper, &kmT, &klT, &kuT,
                                   &p.gp, &p.pitch_lower, &p.pitch_upper);
  KernelTimer timer("tau_absorption");
  if (gf) {
    p.kmajor = kmT; p.lower.kminor = klT; p.upper.kminor = kuT;
    tau_absorption_kernel<0, 0, 0, true><<<grid, kCellThreads, 0, stream()>>>(p);
  } else if (ntemp == 14 && neta == 9 && npres == 59) tau_absorption_kernel<14, 9, 60><<<grid, kCellThreads, 0, stream()>>>(p);
  else tau_absorption_kernel<0, 0, 0><<<grid, kCellThreads, 0, stream()>>>(p);
  RB_LAUNCH_CHECK();
  dev_free(ranges);
}

}  // namespace

extern "C" {

void rrtmgp_interpolation(const int* ncol, const int* nlay, const int* ngas, const int* nflav, const int* neta,
                          const int* npres, const int* ntemp, const int* flavor, const Float* press_ref_log,
                          const Float* temp_ref, const Float* press_ref_log_delta, const Float* temp_ref_min,
                          const Float* temp_ref_delta, const Float* press_ref_trop_log, const Float* vmr_ref,
                          const Float* play, const Float* tlay, const Float* col_gas, int* jtemp, Float* fmajor,
                          Float* fminor, Float* col_mix, Bool* tropo, int* jeta, int* jpress) {
  const size_t ncl = (size_t)*ncol * *nlay, nf = (size_t)*nflav;
  DevArg<int> a_fl(flavor, 2 * nf, Dir::In);
  DevArg<Float> a_prl(press_ref_log, *npres, Dir::In), a_tr(temp_ref, *ntemp, Dir::In),
      a_vr(vmr_ref, 2 * (size_t)(*ngas + 1) * *ntemp, Dir::In), a_pl(play, ncl, Dir::In), a_tl(tlay, ncl, Dir::In),
      a_cg(col_gas, ncl * (*ngas + 1), Dir::In);
  DevArg<int> a_jt(jtemp, ncl, Dir::Out), a_je(jeta, 2 * ncl * nf, Dir::Out, true, 8), a_jp(jpress, ncl, Dir::Out);
  DevArg<Float> a_fj(fmajor, 8 * ncl * nf, Dir::Out, true, 16), a_fn(fminor, 4 * ncl * nf, Dir::Out, true, 16),
      a_cm(col_mix, 2 * ncl * nf, Dir::Out, true, 16);
  DevArg<Bool> a_tp(tropo, ncl, Dir::Out);
  InterpParams p;
  p.ncol = *ncol; p.nlay = *nlay; p.ngas = *ngas; p.nflav = *nflav; p.neta = *neta; p.npres = *npres; p.ntemp = *ntemp;
  p.flavor = a_fl; p.press_ref_log = a_prl; p.temp_ref = a_tr; p.vmr_ref = a_vr; p.play = a_pl; p.tlay = a_tl;
  p.col_gas = a_cg; p.press_ref_log_delta = *press_ref_log_delta; p.temp_ref_min = *temp_ref_min;
  p.temp_ref_delta = *temp_ref_delta; p.press_ref_trop_log = *press_ref_trop_log;
  p.jtemp = a_jt; p.jeta = a_je; p.jpress = a_jp; p.fmajor = a_fj; p.fminor = a_fn; p.col_mix = a_cm; p.tropo = a_tp;
  KernelTimer timer("interpolation");
  interpolation_kernel<<<ceil_div((long long)ncl, kCellThreads), kCellThreads, 0, stream()>>>(p);
  RB_LAUNCH_CHECK();
}

void rrtmgp_compute_tau_absorption(
    const int* ncol, const int* nlay, const int* nbnd, const int* ngpt, const int* ngas, const int* nflav,
    const int* neta, const int* npres, const int* ntemp, const int* nminorlower, const int* nminorklower,
    const int* nminorupper, const int* nminorkupper, const int* idx_h2o, const int* gpoint_flavor,
    const int* band_lims_gpt, const Float* kmajor, const Float* kminor_lower, const Float* kminor_upper,
    const int* minor_limits_gpt_lower, const int* minor_limits_gpt_upper,
    const Bool* minor_scales_with_density_lower, const Bool* minor_scales_with_density_upper,
    const Bool* scale_by_complement_lower, const Bool* scale_by_complement_upper, const int* idx_minor_lower,
    const int* idx_minor_upper, const int* idx_minor_scaling_lower, const int* idx_minor_scaling_upper,
    const int* kminor_start_lower, const int* kminor_start_upper, const Bool* tropo, const Float* col_mix,
    const Float* fmajor, const Float* fminor, const Float* play, const Float* tlay, const Float* col_gas,
    const int* jeta, const int* jtemp, const int* jpress, Float* tau) {
  tau_absorption_impl(*ncol, *nlay, *nbnd, *ngpt, *ngas, *nflav, *neta, *npres, *ntemp, *nminorlower,
                      *nminorklower, *nminorupper, *nminorkupper, *idx_h2o, gpoint_flavor, band_lims_gpt, kmajor,
                      kminor_lower, kminor_upper, minor_limits_gpt_lower, minor_limits_gpt_upper,
                      minor_scales_with_density_lower, minor_scales_with_density_upper, scale_by_complement_lower,
                      scale_by_complement_upper, idx_minor_lower, idx_minor_upper, idx_minor_scaling_lower,
                      idx_minor_scaling_upper, kminor_start_lower, kminor_start_upper, tropo, col_mix, fmajor,
                      fminor, play, tlay, col_gas, jeta, jtemp, jpress, tau, /*accumulate=*/true);
}

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
    const Float* tlay, const Float* col_gas, const int* jeta, const int* jtemp, const int* jpress, Float* tau) {
  tau_absorption_impl(ncol, nlay, nbnd, ngpt, ngas, nflav, neta, npres, ntemp, nminorlower, nminorklower,
                      nminorupper, nminorkupper, idx_h2o, gpoint_flavor, band_lims_gpt, kmajor, kminor_lower,
                      kminor_upper, minor_limits_gpt_lower, minor_limits_gpt_upper,
                      minor_scales_with_density_lower, minor_scales_with_density_upper, scale_by_complement_lower,
                      scale_by_complement_upper, idx_minor_lower, idx_minor_upper, idx_minor_scaling_lower,
                      idx_minor_scaling_upper, kminor_start_lower, kminor_start_upper, tropo, col_mix, fmajor,
                      fminor, play, tlay, col_gas, jeta, jtemp, jpress, tau, /*accumulate=*/false);
}

void rrtmgp_compute_tau_rayleigh(const int* ncol, const int* nlay, const int* nbnd, const int* ngpt,
                                 const int* ngas, const int* nflav, const int* neta, const int* npres,
                                 const int* ntemp, const int* gpoint_flavor, const int* band_lims_gpt,
                                 const Float* krayl, const int* idx_h2o, const Float* col_dry,
                                 const Float* col_gas, const Float* fminor, const int* jeta, const Bool* tropo,
                                 const int* jtemp, Float* tau_rayleigh) {
  const size_t ncl = (size_t)*ncol * *nlay, nf = (size_t)*nflav;
  DevArg<int> a_gf(gpoint_flavor, 2 * (size_t)*ngpt, Dir::In), a_bl(band_lims_gpt, 2 * (size_t)*nbnd, Dir::In);
  DevArg<Float> a_kr(krayl, (size_t)*ntemp * *neta * *ngpt * 2, Dir::In), a_cd(col_dry, ncl, Dir::In),
      a_cg(col_gas, ncl * (*ngas + 1), Dir::In), a_fn(fminor, 4 * ncl * nf, Dir::In, true, 16);
  DevArg<int> a_je(jeta, 2 * ncl * nf, Dir::In, true, 8), a_jt(jtemp, ncl, Dir::In);
  DevArg<Bool> a_tp(tropo, ncl, Dir::In);
  DevArg<Float> a_out(tau_rayleigh, ncl * *ngpt, Dir::Out);
  RaylParams p;
  p.ncol = *ncol; p.nlay = *nlay; p.nbnd = *nbnd; p.ngpt = *ngpt; p.neta = *neta; p.ntemp = *ntemp; p.idx_h2o = *idx_h2o;
  p.gpoint_flavor = a_gf; p.band_lims_gpt = a_bl; p.jeta = a_je; p.jtemp = a_jt; p.krayl = a_kr; p.col_dry = a_cd;
  p.col_gas = a_cg; p.fminor = a_fn; p.tropo = a_tp; p.tau_rayleigh = a_out;
  dim3 grid(ceil_div((long long)ncl, kCellThreads), *nbnd);
  KernelTimer timer("tau_rayleigh");
  tau_rayleigh_kernel<<<grid, kCellThreads, 0, stream()>>>(p);
  RB_LAUNCH_CHECK();
}

void rrtmgp_compute_Planck_source(const int* ncol, const int* nlay, const int* nbnd, const int* ngpt,
                                  const int* nflav, const int* neta, const int* npres, const int* ntemp,
                                  const int* nPlanckTemp, const Float* tlay, const Float* tlev,
                                  const Float* tsfc, const int* sfc_lay, const Float* fmajor, const int* jeta,
                                  const Bool* tropo, const int* jtemp, const int* jpress,
                                  const int* gpoint_bands, const int* band_lims_gpt, const Float* pfracin,
                                  const Float* temp_ref_min, const Float* totplnk_delta, const Float* totplnk,
                                  const int* gpoint_flavor, Float* sfc_src, Float* lay_src, Float* lev_src,
                                  Float* sfc_source_Jac) {
  (void)gpoint_bands;
  const size_t nc = (size_t)*ncol, ncl = nc * *nlay, nclp = nc * (*nlay + 1), nf = (size_t)*nflav, ng = (size_t)*ngpt;
  DevArg<Float> a_tl(tlay, ncl, Dir::In), a_tv(tlev, nclp, Dir::In), a_ts(tsfc, nc, Dir::In),
      a_fj(fmajor, 8 * ncl * nf, Dir::In, true, 16), a_pf(pfracin, (size_t)*ntemp * *neta * (*npres + 1) * ng, Dir::In),
      a_tp(totplnk, (size_t)*nPlanckTemp * *nbnd, Dir::In);
  DevArg<int> a_je(jeta, 2 * ncl * nf, Dir::In, true, 8), a_jt(jtemp, ncl, Dir::In), a_jp(jpress, ncl, Dir::In),
      a_bl(band_lims_gpt, 2 * (size_t)*nbnd, Dir::In), a_gf(gpoint_flavor, 2 * ng, Dir::In);
  DevArg<Bool> a_tr(tropo, ncl, Dir::In);
  DevArg<Float> o_sfc(sfc_src, nc * ng, Dir::Out), o_lay(lay_src, ncl * ng, Dir::Out),
      o_lev(lev_src, nclp * ng, Dir::Out), o_jac(sfc_source_Jac, nc * ng, Dir::Out);
  static const bool gfast_env = [] { const char* e = std::getenv("RRTMGPB_ABI_GFAST"); return !(e && e[0] == '0'); }();
  if (gfast_env && !a_pf.staged() && !a_tp.staged() && !a_bl.staged() && !a_gf.staged() &&
      planck_source_gfast(*ncol, *nlay, *nbnd, *ngpt, *nflav, *neta, *npres, *ntemp, *nPlanckTemp, a_tl, a_tv, a_ts, *sfc_lay,
                          a_fj, a_je, a_tr, a_jt, a_jp, a_bl, a_pf, *temp_ref_min, *totplnk_delta, a_tp, a_gf, o_sfc, o_lay,
                          o_lev, o_jac))
    return;
  PlanckParams p;
  p.ncol = *ncol; p.nlay = *nlay; p.nbnd = *nbnd; p.ngpt = *ngpt; p.nflav = *nflav; p.neta = *neta; p.npres = *npres;
  p.ntemp = *ntemp; p.nPlanckTemp = *nPlanckTemp; p.sfc_lay = *sfc_lay;
  p.tlay = a_tl; p.tlev = a_tv; p.tsfc = a_ts; p.fmajor = a_fj; p.pfracin = a_pf; p.totplnk = a_tp;
  p.jeta = a_je; p.jtemp = a_jt; p.jpress = a_jp; p.band_lims_gpt = a_bl; p.gpoint_flavor = a_gf; p.tropo = a_tr;
  p.temp_ref_min = *temp_ref_min; p.totplnk_delta = *totplnk_delta;
  p.sfc_src = o_sfc; p.lay_src = o_lay; p.lev_src = o_lev; p.sfc_source_Jac = o_jac;
  dim3 grid(ceil_div(*ncol, kCellThreads), *nbnd);
  KernelTimer timer("planck_source");
  if (*ntemp == 14 && *neta == 9 && *npres == 59) planck_source_kernel<14, 9, 60><<<grid, kCellThreads, 0, stream()>>>(p);
  else planck_source_kernel<0, 0, 0><<<grid, kCellThreads, 0, stream()>>>(p);
  RB_LAUNCH_CHECK();
}

void rrtmgp_compute_cld_from_table(const int* ncol, const int* nlay, const int* ngpt, const Bool* mask,
                                   const Float* lwp, const Float* re, const int* nsteps,
                                   const Float* step_size, const Float* offset, const Float* tau_table,
                                   const Float* ssa_table, const Float* asy_table, Float* tau, Float* taussa,
                                   Float* taussag) {
  const size_t ncl = (size_t)*ncol * *nlay, nt = (size_t)*nsteps * *ngpt, n = ncl * *ngpt;
  DevArg<Bool> a_m(mask, ncl, Dir::In);
  DevArg<Float> a_l(lwp, ncl, Dir::In), a_r(re, ncl, Dir::In), a_tt(tau_table, nt, Dir::In),
      a_st(ssa_table, nt, Dir::In), a_at(asy_table, nt, Dir::In);
  DevArg<Float> o_t(tau, n, Dir::Out), o_ts(taussa, n, Dir::Out), o_tsg(taussag, n, Dir::Out);
  CldParams p;
  p.ncol = *ncol; p.nlay = *nlay; p.ngpt = *ngpt; p.nsteps = *nsteps; p.mask = a_m; p.lwp = a_l; p.re = a_r;
  p.tau_table = a_tt; p.ssa_table = a_st; p.asy_table = a_at; p.step_size = *step_size; p.offset = *offset;
  p.tau = o_t; p.taussa = o_ts; p.taussag = o_tsg;
  KernelTimer timer("cld_from_table");
  cld_from_table_kernel<<<ceil_div((long long)ncl, kCellThreads), kCellThreads, 0, stream()>>>(p);
  RB_LAUNCH_CHECK();
}

}  // extern "C"
