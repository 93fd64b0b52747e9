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
#include <type_traits>
#include "../common.cuh"
#include "fastmath.cuh"
#include "rrtmgp_b200_ext.h"

namespace rrtmgpb {

constexpr int kGThreads = 128;
constexpr int kGG = 8;  // g-points per register chunk (a 16-g-point band of the standard k-distributions = 2 chunks)

// per-cell state, struct-of-arrays (cell_state_kernel): everything that does not depend on the band
struct CellState {
  Float *col_dry, *ftemp, *fpress;
  Float *pt_scale;   // 0.01*play/tlay                      (:467)
  Float *vmr_fact;   // 1/col_gas(0)                        (:470)
  Float *dry_fact;   // 1/(1 + col_gas(h2o)*vmr_fact)       (:471)
  int *jtemp, *jpress;
  Bool* tropo;
};

// Small per-band / per-contributor records, resolved once per k-distribution on the host (table cache) so that the
// kernels' prologue is not a chain of dependent index loads (band_lims -> gpoint_flavor -> flavor -> vmr_ref ...).
struct BandInfo {
  int bS, bE;                   // 1-based g-point limits
  int iflav[2];                 // 0-based flavour of the band's first g-point, [itropo]   (:384)
  int igas1[2], igas2[2];       // the flavour's two gases, [itropo]                       (:121-122)
  int mfirst[2], mlast[2];      // minor contributors overlapping the band, [itropo] (lower / upper set)
  int mdiff[2];                 // 1: some of them has a flavour other than the band's
  int regular[2];               // 1: the band has kTauRegChunks*kTG = 16 g-points and every contributor covers exactly the band with the
                                //    band's flavour (all rrtmgp-data bands): the kernel's range tests fold away
};
struct MinorInfo {
  int mS, mE;                   // 1-based g-point limits of the contributor
  int igas, isc;                // idx_minor, idx_minor_scaling (0: none)
  int dens, comp;               // minor_scales_with_density, scale_by_complement
  int kstart;                   // kminor_start (1-based table column of g-point mS)
  int iflav, igas1, igas2;      // flavour of g-point mS in this contributor's atmosphere half (:487)
};
struct GasAux {
  const BandInfo* band;
  const MinorInfo *minor_lower, *minor_upper;
  const Float* ratio;           // vmr_ref(itropo,igas1,jt)/vmr_ref(itropo,igas2,jt) as [itropo][iflav][jt]  (:127-128)
};

struct FusedParams {
  rrtmgpb_gas_tables t;
  int ncol, nlay;
  // band sub-range of this launch (express path: the planes of a few bands at a time live in an L2-sized scratch):
  // bands band0 .. band0+nband_sub-1; output planes are indexed from g-point gpt0 (0-based first g-point of band0), so a
  // sub-range writes a dense (ncol, nlay, ng_sub) array.  Whole k-distribution: band0 = 0, nband_sub = nbnd, gpt0 = 0.
  int band0, nband_sub, gpt0;
  const Float *play, *plev, *tlay, *vmr, *col_dry_in;
  CellState cs;
  // outputs
  int op_kind;  // 1: tau ; 2: tau, ssa, g
  int rows_path = 0;    // 1: warps whose cells do not share table rows take the lanes-along-g-points mapping (tau_band_rows)
                        // (this and stg_stride sit in what used to be padding: the kernels' parameter offsets stay put)
  Float *tau, *ssa, *g;
  // optional by-band cloud increment (kind 0 = none; 1 = 1scl tau ; 2 = 2str tau, ssa, g)
  int cld_kind;
  int stg_stride = 0;   // Floats between the per-warp shared-memory slots (table staging / tau_band_rows records)
  const Float *cld_tau, *cld_ssa, *cld_g;
  // optional second by-band increment applied after the cloud one (aerosols), same kinds
  int aer_kind;
  const Float *aer_tau, *aer_ssa, *aer_g;
  // ABI instantiations (rrtmgp_compute_tau_absorption / rrtmgp_compute_Planck_source through the extern symbols): the
  // interpolation state arrives as the arrays rrtmgp_interpolation wrote - col_mix(2,ncol,nlay,nflav),
  // fmajor(2,2,2,ncol,nlay,nflav), fminor(2,2,ncol,nlay,nflav), jeta(2,ncol,nlay,nflav) - and the gas amounts as
  // col_gas(ncol,nlay,0:ngas); cs.jtemp / jpress / tropo point at the caller's arrays, cs.col_dry at col_gas(:,:,0)
  const Float *abi_col_gas = nullptr, *abi_col_mix = nullptr, *abi_fmajor = nullptr, *abi_fminor = nullptr;
  const int* abi_jeta = nullptr;
  int accumulate = 0;   // tau = tau + result: the extern symbol's semantics (the frontend zeroes tau first, :391)
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
  int maxm;      // most minor contributors any band has in one atmosphere half (sizes the tau kernel's scaling slots)
  GasAux aux;
};

// ---- weights of one flavour for one cell: mo_gas_optics_rrtmgp_kernels.F90:121-168 ----
struct FlavW {
  Float cm[2], fmn[4], fmj[8];
  int je[2];
};

// col_gas(igas) = igas == 0 ? col_dry : vmr(igas)*col_dry   (mo_gas_optics_rrtmgp.F90:594-609)
template <bool ABI = false>
__device__ __forceinline__ Float col_gas_of(const FusedParams& p, size_t c, size_t ncl, int igas, Float col_dry) {
  if (ABI) return p.abi_col_gas[c + ncl * (size_t)igas];
  return igas == 0 ? col_dry : p.vmr[c + ncl * (size_t)(igas - 1)] * col_dry;
}

// ratio: the flavour's ratio_eta_half row, indexed by 0-based temperature
__device__ __forceinline__ void flavor_weights_g(const FusedParams& p, size_t c, size_t ncl, int igas_1, int igas_2,
                                                 const Float* __restrict__ ratio, int jtemp, Float ftemp, Float fpress,
                                                 Float col_dry, FlavW& w) {
  const rrtmgpb_gas_tables& t = p.t;
  const Float cg1 = col_gas_of(p, c, ncl, igas_1, col_dry), cg2 = col_gas_of(p, c, ncl, igas_2, col_dry);
  const Float rt0 = __ldg(ratio + jtemp - 1), rt1 = __ldg(ratio + jtemp);
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const Float ratio_eta_half = it ? rt1 : rt0;
    const Float colmix = cg1 + ratio_eta_half * cg2;
    // rb_div (branch-free, 0 ulp from IEEE on the 50,000-argument probe of tests/test_fastmath.py): the compiler's
    // division carries a slow-path branch that splits this prologue into basic blocks.  A last-bit difference could
    // only move eta across a table node, where the piecewise-linear interpolation is continuous.
    const Float eta = (colmix > (Float)2 * (Float)RB_TINY) ? rb_div(cg1, colmix) : (Float)0.5;
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

// ABI instantiations: the same weights read from the arrays rrtmgp_interpolation wrote (flat orders: col_mix it;
// fmajor ie + 2*ip + 4*it = FlavW::fmj's; fminor ie + 2*it = FlavW::fmn's; jeta it), 16- / 8-byte aligned (DevArg)
__device__ __forceinline__ void abi_load_weights(const FusedParams& p, size_t c, size_t ncl, int iflav, FlavW& w) {
  const size_t cf = c + ncl * (size_t)iflav;
  const Float2 cm = reinterpret_cast<const Float2*>(p.abi_col_mix)[cf];
  w.cm[0] = cm.x; w.cm[1] = cm.y;
  const Float2* fj = reinterpret_cast<const Float2*>(p.abi_fmajor) + 4 * cf;
#pragma unroll
  for (int q = 0; q < 4; ++q) { const Float2 v = fj[q]; w.fmj[2 * q] = v.x; w.fmj[2 * q + 1] = v.y; }
  const Float2* fn = reinterpret_cast<const Float2*>(p.abi_fminor) + 2 * cf;
#pragma unroll
  for (int q = 0; q < 2; ++q) { const Float2 v = fn[q]; w.fmn[2 * q] = v.x; w.fmn[2 * q + 1] = v.y; }
  const int2 je = reinterpret_cast<const int2*>(p.abi_jeta)[cf];
  w.je[0] = je.x; w.je[1] = je.y;
}

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

constexpr int kStgRayl = 128, kStgMinor = 192;   // staged rows of a warp: [0,128) kmajor, [128,192) krayl, then 64 per contributor
// the same from a shared-memory copy of the rows (experimental table staging, see gas_tau_g_kernel STAGE)
template <int VEC>
struct SLoad {
  Float v[VEC];
  __device__ __forceinline__ SLoad(const Float* p) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] = p[i];
  }
};

// 3-D interpolation (interpolate3D_byflav :791-801) of n <= kGG consecutive g-points starting at 0-based
// table column g0: out[i] = scale0*(4 terms of row set 0) + scale1*(4 terms of row set 1).  FULL: n == kGG.
template <int VEC, bool FULL>
__device__ __forceinline__ void interp3d_g(const Float* __restrict__ tab, int gp, int row0, int row1, int s_eta, int s_p,
                                           int g0, int n, const Float (&f)[8], Float scale0, Float scale1,
                                           Float (&out)[kGG]) {
  const size_t d_eta = (size_t)s_eta * gp, d_p = (size_t)s_p * gp;
  const Float* a0 = tab + (size_t)row0 * gp + g0;
  const Float* b0 = tab + (size_t)row1 * gp + g0;
#pragma unroll
  for (int i = 0; i < kGG; i += VEC) {
    if (FULL || i < n) {
      const GLoad<VEC> x0(a0 + i), x1(a0 + d_eta + i), x2(a0 + d_p + i), x3(a0 + d_p + d_eta + i);
      const GLoad<VEC> y0(b0 + i), y1(b0 + d_eta + i), y2(b0 + d_p + i), y3(b0 + d_p + d_eta + i);
#pragma unroll
      for (int v = 0; v < VEC; ++v)
        out[i + v] = scale0 * (f[0] * x0.v[v] + f[1] * x1.v[v] + f[2] * x2.v[v] + f[3] * x3.v[v]) +
                     scale1 * (f[4] * y0.v[v] + f[5] * y1.v[v] + f[6] * y2.v[v] + f[7] * y3.v[v]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// tau (+ ssa, g): compute_tau_absorption :176-338, compute_tau_rayleigh :506-565, combine_abs_and_rayleigh
// (mo_gas_optics_rrtmgp.F90:1954-2002) and the by-band increment (mo_optical_props_kernels.F90:366-477).
//
// Grid: 1-D, block = (kTauCells x 128 consecutive cells, band) with the BAND fastest, so the blocks that share a
// cell range (and its per-cell state, vmr and cloud inputs) are resident together and those inputs come from DRAM
// once.  A thread owns kTauCells cells 128 apart (same layer, neighbouring columns).  When they interpolate between
// the same table rows - the rule for neighbouring columns - every table entry is loaded ONCE and applied to all of
// them: the kernel is bound by L1 -> register bandwidth (8..18 table values per output value), and sharing the
// loads divides that traffic, and the latency waited for per output, by kTauCells.  Otherwise each cell is processed
// on its own.
// ---------------------------------------------------------------------------------------------------
#ifndef RB_TAU_CELLS
#define RB_TAU_CELLS 2
#endif
// resident blocks per SM the register budget is set for: the LW instantiations (tau only) run best at 4 (128
// registers; 5.5 -> 5.2 ms at 65,536 x 72 x 256 on B200; 5 blocks: 5.7, 6 blocks: 6.4 - spills), the SW ones (tau, ssa, g +
// Rayleigh + divisions) at 3 (8.4 vs 8.6 ms)
#ifndef RB_TAU_MINB_LW
#define RB_TAU_MINB_LW 4
#endif
#ifndef RB_TAU_MINB_SW
#define RB_TAU_MINB_SW 3
#endif
constexpr int kTauCells = RB_TAU_CELLS;
#ifndef RB_TAU_TG
#define RB_TAU_TG 4
#endif
constexpr int kTG = RB_TAU_TG;          // g-points per register chunk of the tau kernel
constexpr int kTauRegChunks = 16 / kTG;  // chunks of a regular band (16 g-points, all rrtmgp-data bands)

struct TauCell {
  size_t c;       // cell index (clamped into range; `valid` says whether it may be stored)
  bool valid;
  int slot;       // which of the thread's kTauCells cells this is (selects its precomputed minor scalings)
  Float col_dry, ct, cw, cg, amount_rayl;
  Float at, aw, ag;  // second increment (only touched by the AER instantiations)
  FlavW w;
};

// scaling of one minor contributor at cell c (:461-480): col_gas(minor) [* 0.01 p/T [* vmr of the scaling gas or its
// complement]]
template <bool ABI = false>
__device__ __forceinline__ Float minor_scaling(const FusedParams& p, const MinorInfo& mi, size_t c, size_t ncl, Float col_dry) {
  Float sc = col_gas_of<ABI>(p, c, ncl, mi.igas, col_dry);
  if (mi.dens) {
    sc = sc * p.cs.pt_scale[c];
    if (mi.isc > 0) {
      const Float vmr_fact = p.cs.vmr_fact[c], dry_fact = p.cs.dry_fact[c];
      if (mi.comp) sc = sc * ((Float)1 - col_gas_of<ABI>(p, c, ncl, mi.isc, col_dry) * vmr_fact * dry_fact);
      else sc = sc * (col_gas_of<ABI>(p, c, ncl, mi.isc, col_dry) * vmr_fact * dry_fact);
    }
  }
  return sc;
}

// One output value of the tau kernels: absorption tabs, Rayleigh tray (0 without scattering) -> combine_abs_and_rayleigh
// (mo_gas_optics_rrtmgp.F90:1986-1994), the by-band cloud (and, AER, aerosol) increment
// (mo_optical_props_kernels.F90:366-477) and the store at plane offset o: the `finish` lambda of tau_band_cells as a function,
// for the lanes-along-g-points mapping (tau_band_rows) - same expressions, same results (the cells-per-thread mapping keeps
// its lambda: moving it here changed that kernel's instruction schedule and cost 3 %).
template <bool SW, bool AER, int KIND, bool CLD, bool ABI>
__device__ __forceinline__ void tau_finish(const FusedParams& p, Float ct, Float cw, Float cg, Float at, Float aw, Float ag,
                                           bool valid, size_t o, Float tabs, Float tray) {
  const Float eps3 = (Float)3.0 * (Float)RB_TINY;  // mo_optical_props_kernels.F90:38
  Float to = tabs, ss = 0, gg = 0;
  if (SW) {
    to = tabs + tray;
    // rb_div: <= 1 ulp, no special-case code (the denominators below are >= 2*tiny, finite and normal); the IEEE
    // division sequence made up half of this kernel's instructions (profiles/r1_v8_gas_tau_sw.txt)
    ss = (to > (Float)2 * (Float)RB_TINY) ? rb_div(tray, to) : (Float)0;
  }
  // by-band increments of (to[, ss, gg]) by (ct, cw, cg): mo_optical_props_kernels.F90:366-477; the cloud one
  // first, then (AER instantiations) the aerosol one
  const int op_kind = KIND ? (SW ? 2 : 1) : p.op_kind;
  const int cld_kind = KIND ? (SW ? 2 : 1) : p.cld_kind;
  if (op_kind == 1) {
    auto inc1 = [&](int kind, Float ct_, Float cw_) {
      if (kind == 1) to = to + ct_;                          // inc_1scalar_by_1scalar_bybnd :379
      else if (kind == 2) to = to + ct_ * ((Float)1 - cw_);  // inc_1scalar_by_2stream_bybnd :398
    };
    inc1(cld_kind, ct, cw);
    if (AER) inc1(p.aer_kind, at, aw);
    if (valid) p.tau[o] = (ABI && p.accumulate) ? p.tau[o] + to : to;
  } else {
    auto inc2 = [&](int kind, Float ct_, Float cw_, Float cg_) {
      if (KIND == 1 && !CLD && !AER) {                     // ct == 0: tau12 = to, tauscat12 = to*ss, g stays 0
        ss = rb_div(to * ss, fmax(eps3, to));
      } else if (kind == 1) {                                     // inc_2stream_by_1scalar_bybnd :440-442
        const Float tau12 = to + ct_;
        ss = rb_div(to * ss, fmax(eps3, tau12));
        to = tau12;
      } else if (kind == 2) {                              // inc_2stream_by_2stream_bybnd :468-477
        const Float tau12 = to + ct_;
        const Float tauscat12 = to * ss + ct_ * cw_;
        gg = rb_div(to * ss * gg + ct_ * cw_ * cg_, fmax(eps3, tauscat12));
        ss = rb_div(tauscat12, fmax(eps3, tau12));
        to = tau12;
      }
    };
    inc2(cld_kind, ct, cw, cg);
    if (AER) inc2(p.aer_kind, at, aw, ag);
    if (valid) {
      p.tau[o] = to;
      p.ssa[o] = ss;
      p.g[o] = gg;
    }
  }
}

// NC cells that share tropo, jtemp and the table rows (row0, row1 => je[0], je[1]) of band `bi`
// KIND: 0 = optical-property kind and cloud kind read at run time; 1 = the common combination as compile-time constants
// (LW: 1scl tau += 1scl clouds; SW: 2str incremented by 2str clouds), so the epilogue's kind tests fold away.
// CLD (KIND 1 only): false = no cell of the WARP has cloud in this band (ct == 0 in every lane, a warp-uniform fact): the
// by-band increment then reduces, with identical arithmetic, to ssa = (tau*ssa)/max(eps, tau), g = 0 (its numerator
// tau*ssa*0 + 0*cw*cg is exactly 0), tau unchanged - one division per value instead of three.
// STG: the band's table rows were staged to shared memory by the block (regular bands only): stg = [8 major rows: x0..x3,
// y0..y3][16 g-points], then [contributor][4 rows: m0, m0+eta, m1, m1+eta][16 g-points]
template <bool SW, int VEC, int NC, bool AER, int KIND, bool CLD = true, bool STG = false, bool ABI = false>
__device__ __forceinline__ void tau_band_cells(const FusedParams& p, const TablesT& tt, const BandInfo& bi, bool tropo,
                                               int jtemp, int row0, int row1, TauCell (&cell)[NC], const Float* scal,
                                               const Float* stg = nullptr) {
  const rrtmgpb_gas_tables& t = p.t;
  const size_t ncl = (size_t)p.ncol * p.nlay;
  const int itropo = tropo ? 0 : 1;
  const int iflav = tropo ? bi.iflav[0] : bi.iflav[1];
  const int s_eta = t.ntemp, s_p = t.ntemp * t.neta;
  const int je0 = cell[0].w.je[0], je1 = cell[0].w.je[1];
  const MinorInfo* minfo = tropo ? tt.aux.minor_lower : tt.aux.minor_upper;
  const Float* kminor = tropo ? tt.kminor_lower : tt.kminor_upper;
  const int mpitch = tropo ? tt.nkl : tt.nku;
  const int mfirst = tropo ? bi.mfirst[0] : bi.mfirst[1], mlast = tropo ? bi.mlast[0] : bi.mlast[1];
  const Float eps3 = (Float)3.0 * (Float)RB_TINY;  // mo_optical_props_kernels.F90:38

  auto chunk = [&](int gS, int n, auto full_tag, auto reg_tag) {
    constexpr bool FULL = decltype(full_tag)::value;
    constexpr bool REG = decltype(reg_tag)::value;  // regular band: every contributor covers the whole chunk
    Float acc[NC][kTG];
    // ---- major absorbers: tau = 0 + major (:391 on a zeroed tau); interpolate3D_byflav :791-801 ----
    {
      const size_t d_eta = (size_t)s_eta * tt.gp, d_p = (size_t)s_p * tt.gp;
      const Float* a0 = tt.kmajor + (size_t)row0 * tt.gp + (gS - 1);
      const Float* b0 = tt.kmajor + (size_t)row1 * tt.gp + (gS - 1);
#pragma unroll
      for (int i = 0; i < kTG; i += VEC) {
        if (FULL || i < n) {
          using LD = std::conditional_t<STG, SLoad<VEC>, GLoad<VEC>>;
          const int so = (gS - bi.bS) + i;   // STG: position inside the band's 16 staged g-points
          const LD x0(STG ? stg + so : a0 + i), x1(STG ? stg + 16 + so : a0 + d_eta + i),
              x2(STG ? stg + 32 + so : a0 + d_p + i), x3(STG ? stg + 48 + so : a0 + d_p + d_eta + i);
          const LD y0(STG ? stg + 64 + so : b0 + i), y1(STG ? stg + 80 + so : b0 + d_eta + i),
              y2(STG ? stg + 96 + so : b0 + d_p + i), y3(STG ? stg + 112 + so : b0 + d_p + d_eta + i);
#pragma unroll
          for (int k = 0; k < NC; ++k) {
            const Float(&f)[8] = cell[k].w.fmj;
#pragma unroll
            for (int v = 0; v < VEC; ++v)
              acc[k][i + v] = cell[k].w.cm[0] * (f[0] * x0.v[v] + f[1] * x1.v[v] + f[2] * x2.v[v] + f[3] * x3.v[v]) +
                              cell[k].w.cm[1] * (f[4] * y0.v[v] + f[5] * y1.v[v] + f[6] * y2.v[v] + f[7] * y3.v[v]);
          }
        }
      }
    }
    // ---- minor absorbers touching this chunk (:451-498) ----
    for (int imnr = mfirst; imnr <= mlast; ++imnr) {
      const MinorInfo mi = minfo[imnr];
      if (!REG && (mi.mE < gS || mi.mS > gS + n - 1)) continue;
      // scaling of this contributor for every cell: computed once per (cell, band) by the kernel's prologue
      // (minor_scaling) into the thread's own shared-memory slots, not once per chunk
      Float scaling[NC];
#pragma unroll
      for (int k = 0; k < NC; ++k) scaling[k] = scal[((imnr - mfirst) * kTauCells + cell[k].slot) * kGThreads];
      // The contributor's flavour is the band's flavour for rrtmgp-data (a contributor lives inside one band);
      // otherwise its eta weights are recomputed - single-cell path only, the caller does not share rows
      // across cells for such bands (BandInfo::mdiff).
      Float am[NC][4];
      int jm0 = je0, jm1 = je1;
#pragma unroll
      for (int k = 0; k < NC; ++k)
#pragma unroll
        for (int q = 0; q < 4; ++q) am[k][q] = cell[k].w.fmn[q];
      if (!REG && NC == 1 && mi.iflav != iflav) {
        FlavW wm;
        const size_t c = cell[0].c;
        if (ABI) abi_load_weights(p, c, ncl, mi.iflav, wm);
        else flavor_weights_g(p, c, ncl, mi.igas1, mi.igas2, tt.aux.ratio + (size_t)(itropo * t.nflav + mi.iflav) * t.ntemp, jtemp,
                              p.cs.ftemp[c], p.cs.fpress[c], cell[0].col_dry, wm);
#pragma unroll
        for (int q = 0; q < 4; ++q) am[0][q] = wm.fmn[q];
        jm0 = wm.je[0]; jm1 = wm.je[1];
      }
      const int kcol0 = mi.kstart + (gS - mi.mS) - 1;  // table column of g-point gS+i: kminor_start + (gS+i - mS) - 1
      const size_t d_eta = (size_t)s_eta * mpitch;
      const Float* m0 = kminor + (size_t)((jtemp - 1) + s_eta * (jm0 - 1)) * mpitch + kcol0;
      const Float* m1 = kminor + (size_t)(jtemp + s_eta * (jm1 - 1)) * mpitch + kcol0;
      const int iS = mi.mS - gS, iE = min(mi.mE - gS, n - 1);  // chunk positions covered by this contributor
      const bool whole = REG || (FULL && iS <= 0 && iE == kTG - 1);
#pragma unroll
      for (int i = 0; i < kTG; i += VEC) {
        if (whole || (i >= iS && i <= iE)) {  // VEC == 2: intervals start even and have even length (TablesT::vec)
          using LD = std::conditional_t<STG, SLoad<VEC>, GLoad<VEC>>;
          const Float* sm_m = STG ? stg + kStgMinor + (imnr - mfirst) * 64 + (gS - bi.bS) + i : nullptr;
          const LD x0(STG ? sm_m : m0 + i), x1(STG ? sm_m + 16 : m0 + d_eta + i), y0(STG ? sm_m + 32 : m1 + i),
              y1(STG ? sm_m + 48 : m1 + d_eta + i);
#pragma unroll
          for (int k = 0; k < NC; ++k) {
            const Float(&a)[4] = am[k];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
              const Float kint = a[0] * x0.v[v] + a[1] * x1.v[v] + a[2] * y0.v[v] + a[3] * y1.v[v];  // :757-760
              acc[k][i + v] = acc[k][i + v] + scaling[k] * kint;                                     // :493
            }
          }
        }
      }
    }
    // ---- Rayleigh (:554-559), combination (:1986-1994), cloud increment, store ----
    const size_t dr_eta = (size_t)s_eta * tt.gp;
    const Float* kr = SW ? tt.krayl + (size_t)s_p * tt.gp * itropo + (gS - 1) : nullptr;
    const Float* r0 = SW ? kr + (size_t)((jtemp - 1) + s_eta * (je0 - 1)) * tt.gp : nullptr;
    const Float* r1 = SW ? kr + (size_t)(jtemp + s_eta * (je1 - 1)) * tt.gp : nullptr;
    const size_t goff = ncl * (size_t)(gS - 1 - p.gpt0);
    // one output value: absorption tabs, Rayleigh tray (0 without scattering) of cell k at chunk position i
    auto finish = [&](int k, int i, Float tabs, Float tray) {
      Float to = tabs, ss = 0, gg = 0;
      if (SW) {
        to = tabs + tray;
        // rb_div: <= 1 ulp, no special-case code (the denominators below are >= 2*tiny, finite and normal); the IEEE
        // division sequence made up half of this kernel's instructions (profiles/r1_v8_gas_tau_sw.txt)
        ss = (to > (Float)2 * (Float)RB_TINY) ? rb_div(tray, to) : (Float)0;
      }
      const size_t o = cell[k].c + goff + ncl * (size_t)i;
      // by-band increments of (to[, ss, gg]) by (ct, cw, cg): mo_optical_props_kernels.F90:366-477; the cloud one
      // first, then (AER instantiations) the aerosol one
      const int op_kind = KIND ? (SW ? 2 : 1) : p.op_kind;
      const int cld_kind = KIND ? (SW ? 2 : 1) : p.cld_kind;
      if (op_kind == 1) {
        auto inc1 = [&](int kind, Float ct, Float cw) {
          if (kind == 1) to = to + ct;                         // inc_1scalar_by_1scalar_bybnd :379
          else if (kind == 2) to = to + ct * ((Float)1 - cw);  // inc_1scalar_by_2stream_bybnd :398
        };
        inc1(cld_kind, cell[k].ct, cell[k].cw);
        if (AER) inc1(p.aer_kind, cell[k].at, cell[k].aw);
        if (cell[k].valid) p.tau[o] = (ABI && p.accumulate) ? p.tau[o] + to : to;
      } else {
        auto inc2 = [&](int kind, Float ct, Float cw, Float cg) {
          if (KIND == 1 && !CLD && !AER) {                     // ct == 0: tau12 = to, tauscat12 = to*ss, g stays 0
            ss = rb_div(to * ss, fmax(eps3, to));
          } else if (kind == 1) {                                     // inc_2stream_by_1scalar_bybnd :440-442
            const Float tau12 = to + ct;
            ss = rb_div(to * ss, fmax(eps3, tau12));
            to = tau12;
          } else if (kind == 2) {                              // inc_2stream_by_2stream_bybnd :468-477
            const Float tau12 = to + ct;
            const Float tauscat12 = to * ss + ct * cw;
            gg = rb_div(to * ss * gg + ct * cw * cg, fmax(eps3, tauscat12));
            ss = rb_div(tauscat12, fmax(eps3, tau12));
            to = tau12;
          }
        };
        inc2(cld_kind, cell[k].ct, cell[k].cw, cell[k].cg);
        if (AER) inc2(p.aer_kind, cell[k].at, cell[k].aw, cell[k].ag);
        if (cell[k].valid) {
          p.tau[o] = to;
          p.ssa[o] = ss;
          p.g[o] = gg;
        }
      }
    };
#pragma unroll
    for (int i0 = 0; i0 < kTG; i0 += VEC) {
      if (!FULL && i0 >= n) continue;
      if (SW) {
        using LD = std::conditional_t<STG, SLoad<VEC>, GLoad<VEC>>;
        const Float* sm_r = STG ? stg + kStgRayl + (gS - bi.bS) + i0 : nullptr;
        const LD x0(STG ? sm_r : r0 + i0), x1(STG ? sm_r + 16 : r0 + dr_eta + i0), y0(STG ? sm_r + 32 : r1 + i0),
            y1(STG ? sm_r + 48 : r1 + dr_eta + i0);
#pragma unroll
        for (int k = 0; k < NC; ++k) {
          const Float(&a)[4] = cell[k].w.fmn;
#pragma unroll
          for (int v = 0; v < VEC; ++v)
            finish(k, i0 + v, acc[k][i0 + v],
                   (a[0] * x0.v[v] + a[1] * x1.v[v] + a[2] * y0.v[v] + a[3] * y1.v[v]) * cell[k].amount_rayl);
        }
      } else {
#pragma unroll
        for (int k = 0; k < NC; ++k)
#pragma unroll
          for (int v = 0; v < VEC; ++v) finish(k, i0 + v, acc[k][i0 + v], (Float)0);
      }
    }
  };

  if (tropo ? bi.regular[0] : bi.regular[1]) {
    for (int q = 0; q < kTauRegChunks; ++q) chunk(bi.bS + q * kTG, kTG, std::true_type{}, std::true_type{});
  } else {
    for (int gS = bi.bS; gS <= bi.bE; gS += kTG) {
      const int n = min(kTG, bi.bE - gS + 1);
      if (n == kTG) chunk(gS, n, std::true_type{}, std::false_type{});
      else chunk(gS, n, std::false_type{}, std::false_type{});
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Lanes-along-g-points mapping (tau_band_rows) for WARPS whose cells do NOT share table rows - unrelated neighbouring
// columns (RFMIP-like profile sets, BASELINE config 3).  In the cells-per-thread mapping every lane of such a warp reads its
// OWN rows: one LDG.128 is 32 L1 wavefronts (one line per lane, 16 of its 128 bytes used), every line is fetched eight times
// for the band's 16 g-points, and the kernel is bound by the L1 data pipe (84 % busy under ncu).  Here the
// warp first publishes the per-cell state its lanes computed in the prologue (weights, rows, cloud properties: one
// 176-byte record per cell in the warp's own shared-memory slots - the table-staging slots, which such a warp never
// uses; no block barrier), then re-maps: 4 consecutive lanes take the 16 g-points of ONE cell (4 each: RB_ROWS_GPL), so a table
// row is read as two 64-byte requests by 4 lanes - a warp request touches 8 lines instead of 32, each line is fetched twice
// instead of eight times - and the stores of a g-point cover 8 consecutive cells = two full 32-byte sectors.  Regular bands only (16 g-points, every contributor covering the band:
// all rrtmgp-data bands); per-g-point arithmetic is tau_band_cells' expression for expression (same tau_finish), so both
// mappings give the same results (tests/test_gas_optics_rows_path.py: bit for bit).
// ---------------------------------------------------------------------------------------------------
#ifndef RB_ROWS_GPL
#define RB_ROWS_GPL 4
#endif
constexpr int kRowRec = 22;   // Floats per cell record: cm[2], fmj[8], fmn[4], ct, cw, cg, amount_rayl, 6 ints, pad
__host__ __device__ constexpr size_t tau_rows_warp_bytes() { return (size_t)32 * kRowRec * sizeof(double); }

template <bool SW, int KIND, bool ABI>
__device__ __forceinline__ void tau_band_rows(const FusedParams& p, const TablesT& tt, const BandInfo& bi,
                                              const TauCell (&cell)[kTauCells], const bool (&tropo)[kTauCells],
                                              const int (&jtemp)[kTauCells], const int (&row0)[kTauCells],
                                              const int (&row1)[kTauCells], const Float* scal_block, Float* rec_warp) {
  static_assert(sizeof(Float) == 8 || KIND < 0, "double-precision layout (the callers require TablesT::vec == 2)");   // (dependent: checked on instantiation)
  const rrtmgpb_gas_tables& t = p.t;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // GPL g-points per lane (RB_ROWS_GPL): 2 -> eight lanes per cell, a row is one 128-byte request; 4 -> four lanes per cell, two
  // requests of 64 bytes per row but half the per-lane pointer arithmetic per g-point
  constexpr int GPL = RB_ROWS_GPL, LPC = 16 / GPL, CPP = 32 / LPC;   // lanes per cell, cells per pass
  const int j = lane % LPC;    // lane j of a cell's LPC: g-points bS + GPL*j ... bS + GPL*j + GPL - 1
  const int cq = lane / LPC;   // which of the CPP cells of a pass
  const size_t ncl = (size_t)p.ncol * p.nlay;
  const size_t cbase_warp = (size_t)(blockIdx.x / p.nband_sub) * (kTauCells * kGThreads) + (size_t)warp * 32;
  const int s_eta = t.ntemp, s_p = t.ntemp * t.neta;
  const size_t d_eta = (size_t)s_eta * tt.gp, d_p = (size_t)s_p * tt.gp;
  const int gcol = (bi.bS - 1) + GPL * j;  // 0-based table column of this lane's first g-point
#pragma unroll   // (fully unrolled: a run-time k would push the callers' per-cell register arrays to local memory)
  for (int k = 0; k < kTauCells; ++k) {
    __syncwarp();   // every lane is done reading the previous records
    {
      const TauCell& ce = cell[k];
      Float2* r = reinterpret_cast<Float2*>(rec_warp + (size_t)lane * kRowRec);
      r[0] = Float2{ce.w.cm[0], ce.w.cm[1]};
#pragma unroll
      for (int q = 0; q < 4; ++q) r[1 + q] = Float2{ce.w.fmj[2 * q], ce.w.fmj[2 * q + 1]};
      r[5] = Float2{ce.w.fmn[0], ce.w.fmn[1]};
      r[6] = Float2{ce.w.fmn[2], ce.w.fmn[3]};
      r[7] = Float2{ce.ct, ce.cw};
      r[8] = Float2{ce.cg, ce.amount_rayl};
      int4* ri = reinterpret_cast<int4*>(r + 9);
      ri[0] = int4{jtemp[k], row0[k], row1[k], ce.w.je[0]};
      ri[1] = int4{ce.w.je[1], (tropo[k] ? 1 : 0) | (ce.valid ? 2 : 0), 0, 0};
    }
    __syncwarp();
#pragma unroll 1
    for (int it = 0; it < 32 / CPP; ++it) {
      const int cl = it * CPP + cq;
      const Float2* r = reinterpret_cast<const Float2*>(rec_warp + (size_t)cl * kRowRec);
      const int4 i0 = reinterpret_cast<const int4*>(r + 9)[0], i1 = reinterpret_cast<const int4*>(r + 9)[1];
      const int flags = i1.y;
      if (!(flags & 2)) continue;   // beyond the last cell
      const bool tr = flags & 1;
      const int jt = i0.x, r0 = i0.y, r1 = i0.z, je0 = i0.w, je1 = i1.x;
      const Float2 cm = r[0];
      Float f[8], a[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { const Float2 v = r[1 + q]; f[2 * q] = v.x; f[2 * q + 1] = v.y; }
      { const Float2 v = r[5], w = r[6]; a[0] = v.x; a[1] = v.y; a[2] = w.x; a[3] = w.y; }
      Float acc[GPL];
      {  // major absorbers (interpolate3D_byflav :791-801)
        const Float* a0 = tt.kmajor + (size_t)r0 * tt.gp + gcol;
        const Float* b0 = tt.kmajor + (size_t)r1 * tt.gp + gcol;
#pragma unroll
        for (int h = 0; h < GPL; h += 2) {
          const GLoad<2> x0(a0 + h), x1(a0 + d_eta + h), x2(a0 + d_p + h), x3(a0 + d_p + d_eta + h);
          const GLoad<2> y0(b0 + h), y1(b0 + d_eta + h), y2(b0 + d_p + h), y3(b0 + d_p + d_eta + h);
#pragma unroll
          for (int v = 0; v < 2; ++v)
            acc[h + v] = cm.x * (f[0] * x0.v[v] + f[1] * x1.v[v] + f[2] * x2.v[v] + f[3] * x3.v[v]) +
                         cm.y * (f[4] * y0.v[v] + f[5] * y1.v[v] + f[6] * y2.v[v] + f[7] * y3.v[v]);
        }
      }
      {  // minor absorbers (:451-498); regular band: every contributor starts at the band's first g-point
        const MinorInfo* minfo = tr ? tt.aux.minor_lower : tt.aux.minor_upper;
        const Float* kminor = tr ? tt.kminor_lower : tt.kminor_upper;
        const int mpitch = tr ? tt.nkl : tt.nku;
        const int mfirst = tr ? bi.mfirst[0] : bi.mfirst[1], mlast = tr ? bi.mlast[0] : bi.mlast[1];
        const size_t de = (size_t)s_eta * mpitch;
        const Float* m0 = kminor + (size_t)((jt - 1) + s_eta * (je0 - 1)) * mpitch + GPL * j - 1;   // + kstart below
        const Float* m1 = kminor + (size_t)(jt + s_eta * (je1 - 1)) * mpitch + GPL * j - 1;
        const Float* sc = scal_block + (size_t)k * kGThreads + warp * 32 + cl;
        for (int imnr = mfirst; imnr <= mlast; ++imnr) {
          const Float scaling = sc[(size_t)(imnr - mfirst) * kTauCells * kGThreads];
          const int ks = minfo[imnr].kstart;
#pragma unroll
          for (int h = 0; h < GPL; h += 2) {
            const GLoad<2> x0(m0 + ks + h), x1(m0 + de + ks + h), y0(m1 + ks + h), y1(m1 + de + ks + h);
#pragma unroll
            for (int v = 0; v < 2; ++v) {
              const Float kint = a[0] * x0.v[v] + a[1] * x1.v[v] + a[2] * y0.v[v] + a[3] * y1.v[v];  // :757-760
              acc[h + v] = acc[h + v] + scaling * kint;                                             // :493
            }
          }
        }
      }
      Float tray[GPL];
#pragma unroll
      for (int v = 0; v < GPL; ++v) tray[v] = 0;
      const Float2 c01 = r[7], c23 = r[8];   // (ct, cw), (cg, amount_rayl)
      if (SW) {  // Rayleigh (:554-559)
        const Float* kr = tt.krayl + (size_t)s_p * tt.gp * (tr ? 0 : 1) + gcol;
        const Float* q0 = kr + (size_t)((jt - 1) + s_eta * (je0 - 1)) * tt.gp;
        const Float* q1 = kr + (size_t)(jt + s_eta * (je1 - 1)) * tt.gp;
#pragma unroll
        for (int h = 0; h < GPL; h += 2) {
          const GLoad<2> x0(q0 + h), x1(q0 + d_eta + h), y0(q1 + h), y1(q1 + d_eta + h);
#pragma unroll
          for (int v = 0; v < 2; ++v) tray[h + v] = (a[0] * x0.v[v] + a[1] * x1.v[v] + a[2] * y0.v[v] + a[3] * y1.v[v]) * c23.y;
        }
      }
      const size_t o = cbase_warp + (size_t)k * kGThreads + cl + ncl * (size_t)(gcol - p.gpt0);
#pragma unroll
      for (int v = 0; v < GPL; ++v)
        tau_finish<SW, false, KIND, true, ABI>(p, c01.x, c01.y, c23.x, (Float)0, (Float)0, (Float)0, true, o + ncl * (size_t)v, acc[v],
                                               tray[v]);
    }
  }
}

// STAGE: when every cell of a WARP interpolates between the SAME table rows (a regular band; neighbouring columns in the
// same T / p / eta bins - one vote per warp) lane 0 copies those rows - 8 of kmajor, 4 of krayl (SW), 4 per minor
// contributor, 128 bytes each - to the warp's shared-memory slots with cp.async.bulk (the TMA engine, completion on the
// warp's mbarrier) and the warp reads them from there: immediate-offset LDS instead of LDG with 64-bit row addresses, at
// shared-memory latency.  Measured on B200: LW tau 5.18 -> 4.33 ms at 65,536 x 72 x 256 (block-wide
// variant).  Warps whose cells differ take the L1 path below.  RRTMGPB_TABLE_TMA=0 switches it off.
// ROWS: the instantiation that carries the lanes-along-g-points mapping for warps of unrelated columns (tau_band_rows).  A
// separate instantiation because its mere presence changes the register allocation of the cells-per-thread path (measured:
// LW tau 4.38 -> 4.60 ms on the replicated profile with the path compiled in but never taken); the host launches it when
// rrtmgpb_set_gas_optics_rows_path(1) / RRTMGPB_TAU_ROWS=1 says the columns are unrelated.
template <bool SW, int VEC, bool AER, int KIND, bool STAGE = false, bool ABI = false, bool ROWS = false>
__global__ void __launch_bounds__(kGThreads, (SW || AER) ? RB_TAU_MINB_SW : RB_TAU_MINB_LW) gas_tau_g_kernel(const FusedParams p, const TablesT tt) {
  const rrtmgpb_gas_tables& t = p.t;
  const size_t ncl = (size_t)p.ncol * p.nlay;
  const int ibnd = p.band0 + blockIdx.x % p.nband_sub;
  const size_t cbase_raw = (size_t)(blockIdx.x / p.nband_sub) * (kTauCells * kGThreads) + threadIdx.x;
  if (!STAGE && cbase_raw >= ncl) return;
  const bool thread_valid = cbase_raw < ncl;                 // STAGE: every thread stays for the block-wide barriers
  const size_t cbase = thread_valid ? cbase_raw : ncl - 1;
  // minor-contributor scalings of this thread's cells: [contributor of the band][cell slot][thread], lane-private
  // (no barrier: a thread only reads what it wrote); tt.maxm contributors at most
  extern __shared__ __align__(16) unsigned char tau_smem_raw[];
  Float* scal = reinterpret_cast<Float*>(tau_smem_raw) + threadIdx.x;
  const BandInfo bi = tt.aux.band[ibnd];
  const int s_eta = t.ntemp, s_p = t.ntemp * t.neta;
  TauCell cell[kTauCells];
  bool tropo[kTauCells];
  int jtemp[kTauCells], row0[kTauCells], row1[kTauCells];
#pragma unroll
  for (int k = 0; k < kTauCells; ++k) {
    const size_t craw = cbase + (size_t)k * kGThreads;
    TauCell& ce = cell[k];
    ce.valid = thread_valid && craw < ncl;
    ce.slot = k;
    const size_t c = ce.valid ? craw : cbase;  // out-of-range slots shadow the thread's first cell, never store
    ce.c = c;
    ce.col_dry = p.cs.col_dry[c];
    jtemp[k] = p.cs.jtemp[c];
    tropo[k] = p.cs.tropo[c];
    const int itropo = tropo[k] ? 0 : 1;
    const int jpress = p.cs.jpress[c] + itropo + 1;  // :390
    // (explicit selects: indexing the register copy of BandInfo with a run-time itropo would push it to local memory)
    const int iflav = tropo[k] ? bi.iflav[0] : bi.iflav[1];
    if (ABI) abi_load_weights(p, c, ncl, iflav, ce.w);
    else flavor_weights_g(p, c, ncl, tropo[k] ? bi.igas1[0] : bi.igas1[1], tropo[k] ? bi.igas2[0] : bi.igas2[1],
                          tt.aux.ratio + (size_t)(itropo * t.nflav + iflav) * t.ntemp, jtemp[k], p.cs.ftemp[c], p.cs.fpress[c],
                          ce.col_dry, ce.w);
    row0[k] = (jtemp[k] - 1) + s_eta * (ce.w.je[0] - 1) + s_p * (jpress - 2);
    row1[k] = jtemp[k] + s_eta * (ce.w.je[1] - 1) + s_p * (jpress - 2);
    // cloud properties of this (cell, band), if the caller wants them added (mo_optical_props.F90:956-1000)
    ce.ct = 0; ce.cw = 0; ce.cg = 0;
    const int cld_kind = KIND ? (SW ? 2 : 1) : p.cld_kind;
    if (cld_kind) {
      const size_t cb = c + ncl * (size_t)ibnd;
      ce.ct = p.cld_tau[cb];
      if (cld_kind == 2) { ce.cw = p.cld_ssa[cb]; ce.cg = p.cld_g[cb]; }
    }
    ce.at = 0; ce.aw = 0; ce.ag = 0;
    if (AER && p.aer_kind) {
      const size_t cb = c + ncl * (size_t)ibnd;
      ce.at = p.aer_tau[cb];
      if (p.aer_kind == 2) { ce.aw = p.aer_ssa[cb]; ce.ag = p.aer_g[cb]; }
    }
    ce.amount_rayl = SW ? col_gas_of(p, c, ncl, t.idx_h2o, ce.col_dry) + ce.col_dry : (Float)0;  // :559
    {
      const MinorInfo* minfo = tropo[k] ? tt.aux.minor_lower : tt.aux.minor_upper;
      const int mfirst = tropo[k] ? bi.mfirst[0] : bi.mfirst[1], mlast = tropo[k] ? bi.mlast[0] : bi.mlast[1];
      for (int imnr = mfirst; imnr <= mlast; ++imnr)
#ifdef RB_EXPERIMENT_NOSCAL   /* upper bound of what the scaling prologue costs (results are WRONG) */
        scal[((imnr - mfirst) * kTauCells + k) * kGThreads] = ce.col_dry;
#else
        scal[((imnr - mfirst) * kTauCells + k) * kGThreads] = minor_scaling<ABI>(p, minfo[imnr], c, ncl, ce.col_dry);
#endif
    }
  }
  bool shared_rows = true;
#pragma unroll
  for (int k = 1; k < kTauCells; ++k)
    shared_rows = shared_rows && tropo[k] == tropo[0] && row0[k] == row0[0] && row1[k] == row1[0];
  if (tropo[0] ? bi.mdiff[0] : bi.mdiff[1]) shared_rows = false;
  // cloud-free warps (layers above / below the cloud deck, clear regions) take the reduced increment; the test is
  // warp-uniform, so it costs no divergence
  bool cloudy = false;
  if (SW && KIND == 1 && !AER) {
    bool mine = false;
#pragma unroll
    for (int k = 0; k < kTauCells; ++k) mine = mine || cell[k].ct != (Float)0;
    cloudy = __any_sync(__activemask(), mine);
  }
  if (STAGE) {
    // ---- warp-uniform rows?  (lane 0's rows against everybody's: one vote, no block barrier)
    const unsigned full = 0xffffffffu;
    const bool regular = tropo[0] ? bi.regular[0] : bi.regular[1];
    // (the shuffles first, unconditionally: every lane of the warp must execute them)
    const int r0_lane0 = __shfl_sync(full, row0[0], 0), r1_lane0 = __shfl_sync(full, row1[0], 0);
    const int tr_lane0 = __shfl_sync(full, (int)tropo[0], 0);
    const bool same = shared_rows && regular && row0[0] == r0_lane0 && row1[0] == r1_lane0 && (int)tropo[0] == tr_lane0;
    const unsigned votes = ROWS ? __ballot_sync(full, same) : (__all_sync(full, same) ? full : 0u);
    // ---- unrelated columns (fewer than p.rows_path = 32 of the lanes share rows between their own two cells - profiles of one
    // data set coincide in some T / p / eta bins by chance, so the count is rarely near zero; warps in which every lane shares,
    // e.g. stratospheric layers where every profile falls into the same bins, are better off below): the warp takes the
    // lanes-along-g-points mapping; its records overlay the staging slots it will not use
    if constexpr (ROWS) {
      if (bi.regular[0] && bi.regular[1] && __popc(__ballot_sync(full, shared_rows)) < p.rows_path) {   // rows_path = the threshold
        Float* rec = reinterpret_cast<Float*>(tau_smem_raw) + (size_t)tt.maxm * kTauCells * kGThreads + (size_t)(threadIdx.x >> 5) * p.stg_stride;
        tau_band_rows<SW, KIND, ABI>(p, tt, bi, cell, tropo, jtemp, row0, row1, reinterpret_cast<const Float*>(tau_smem_raw), rec);
        return;
      }
    }
    if (votes == full) {
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      __shared__ __align__(8) uint64_t s_bar[kGThreads / 32];
      const int s_eta = t.ntemp, s_p = t.ntemp * t.neta;
      const bool tr = tropo[0];
      const int mfirst = tr ? bi.mfirst[0] : bi.mfirst[1], mlast = tr ? bi.mlast[0] : bi.mlast[1];
      const int nm = mlast >= mfirst ? mlast - mfirst + 1 : 0;
      const int rows_per_warp = kStgMinor / 16 + 4 * tt.maxm;
      Float* stg = reinterpret_cast<Float*>(tau_smem_raw) + (size_t)tt.maxm * kTauCells * kGThreads +
                   (ROWS ? (size_t)warp * p.stg_stride : (size_t)warp * rows_per_warp * 16);
      const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_bar[warp]);
      if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        const unsigned bytes = (unsigned)((8 + (SW ? 4 : 0) + 4 * nm) * 16 * sizeof(Float));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
        auto bulk = [&](Float* dst, const Float* src) {
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
                           "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "n"(16 * (int)sizeof(Float)), "r"(bar)
                       : "memory");
        };
        const Float* a0 = tt.kmajor + (size_t)row0[0] * tt.gp + (bi.bS - 1);
        const Float* b0 = tt.kmajor + (size_t)row1[0] * tt.gp + (bi.bS - 1);
        const size_t d_eta = (size_t)s_eta * tt.gp, d_p = (size_t)s_p * tt.gp;
        bulk(stg, a0); bulk(stg + 16, a0 + d_eta); bulk(stg + 32, a0 + d_p); bulk(stg + 48, a0 + d_p + d_eta);
        bulk(stg + 64, b0); bulk(stg + 80, b0 + d_eta); bulk(stg + 96, b0 + d_p); bulk(stg + 112, b0 + d_p + d_eta);
        const int je0 = cell[0].w.je[0], je1 = cell[0].w.je[1];
        if (SW) {
          const Float* kr = tt.krayl + (size_t)s_p * tt.gp * (tr ? 0 : 1) + (bi.bS - 1);
          const Float* r0 = kr + (size_t)((jtemp[0] - 1) + s_eta * (je0 - 1)) * tt.gp;
          const Float* r1 = kr + (size_t)(jtemp[0] + s_eta * (je1 - 1)) * tt.gp;
          const size_t dr = (size_t)s_eta * tt.gp;
          bulk(stg + kStgRayl, r0); bulk(stg + kStgRayl + 16, r0 + dr); bulk(stg + kStgRayl + 32, r1); bulk(stg + kStgRayl + 48, r1 + dr);
        }
        const MinorInfo* minfo = tr ? tt.aux.minor_lower : tt.aux.minor_upper;
        const Float* kminor = tr ? tt.kminor_lower : tt.kminor_upper;
        const int mpitch = tr ? tt.nkl : tt.nku;
        for (int m = 0; m < nm; ++m) {
          const MinorInfo mi = minfo[mfirst + m];
          const Float* m0 = kminor + (size_t)((jtemp[0] - 1) + s_eta * (je0 - 1)) * mpitch + (mi.kstart - 1);
          const Float* m1 = kminor + (size_t)(jtemp[0] + s_eta * (je1 - 1)) * mpitch + (mi.kstart - 1);
          const size_t de = (size_t)s_eta * mpitch;
          Float* d = stg + kStgMinor + m * 64;
          bulk(d, m0); bulk(d + 16, m0 + de); bulk(d + 32, m1); bulk(d + 48, m1 + de);
        }
      }
      __syncwarp();
      asm volatile(
          "{\n.reg .pred P1;\nWAIT_STG:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra DONE_STG;\nbra WAIT_STG;\nDONE_STG:\n}\n" ::"r"(bar)
          : "memory");
      if (SW && KIND == 1 && !AER && !cloudy)
        tau_band_cells<SW, VEC, kTauCells, AER, KIND, false, true, ABI>(p, tt, bi, tropo[0], jtemp[0], row0[0], row1[0], cell, scal, stg);
      else
        tau_band_cells<SW, VEC, kTauCells, AER, KIND, true, true, ABI>(p, tt, bi, tropo[0], jtemp[0], row0[0], row1[0], cell, scal, stg);
      return;
    }
  }
  if (shared_rows && SW && KIND == 1 && !AER && !cloudy) {
    tau_band_cells<SW, VEC, kTauCells, AER, KIND, false, false, ABI>(p, tt, bi, tropo[0], jtemp[0], row0[0], row1[0], cell, scal);
  } else if (shared_rows) {
    tau_band_cells<SW, VEC, kTauCells, AER, KIND, true, false, ABI>(p, tt, bi, tropo[0], jtemp[0], row0[0], row1[0], cell, scal);
  } else {
#pragma unroll
    for (int k = 0; k < kTauCells; ++k) {
      if (!cell[k].valid) continue;
      TauCell one[1] = {cell[k]};
      tau_band_cells<SW, VEC, 1, AER, KIND, true, false, ABI>(p, tt, bi, tropo[k], jtemp[k], row0[k], row1[k], one, scal);
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

// sqrt(a*b) of two Planck fractions (:699).  RB_PLANCK_FAST_SQRT: the branch-free rb_sqrt (<= 1 ulp; MUFU seed + 6 fp64
// instructions, no slow-path call: 31 call sites with BSSY / CALL / BSYNC otherwise keep the compiler from interleaving
// the g-points of a pass); products below 1e-290 - fractions of a band's Planck function that small carry no energy -
// give 0 instead of a subnormal-range root.
#ifndef RB_PLANCK_FAST_SQRT
#define RB_PLANCK_FAST_SQRT 0
#endif
__device__ __forceinline__ Float planck_geo_mean(Float a, Float b) {
  const Float m = a * b;
#if RB_PLANCK_FAST_SQRT
  return (m > (Float)1.0e-290) ? rb_sqrt(fmax(m, (Float)1.0e-290)) : (Float)0;
#else
  return sqrt(m);
#endif
}

#ifndef RB_PLANCK_PG
#define RB_PLANCK_PG 2
#endif
constexpr int kPG = RB_PLANCK_PG * kGG;  // g-points per pass of the Planck kernel (previous layer's fractions stay in registers)

// interpolation weights and table rows of band `bi` at cell c (:121-168, :390)
template <bool ABI = false>
__device__ __forceinline__ void planck_cell_weights(const FusedParams& p, const TablesT& tt, const BandInfo& bi, size_t c,
                                                    size_t ncl, FlavW& w, int& row0, int& row1) {
  const rrtmgpb_gas_tables& t = p.t;
  const bool tropo = p.cs.tropo[c];
  const int itropo = tropo ? 0 : 1;
  const int jtemp = p.cs.jtemp[c];
  const int jpress = p.cs.jpress[c] + itropo + 1;
  if (ABI) {  // fmajor and jeta as rrtmgp_interpolation wrote them (only these two are used below)
    const size_t cf = c + ncl * (size_t)(tropo ? bi.iflav[0] : bi.iflav[1]);
    const Float2* fj = reinterpret_cast<const Float2*>(p.abi_fmajor) + 4 * cf;
#pragma unroll
    for (int q = 0; q < 4; ++q) { const Float2 v = fj[q]; w.fmj[2 * q] = v.x; w.fmj[2 * q + 1] = v.y; }
    const int2 je = reinterpret_cast<const int2*>(p.abi_jeta)[cf];
    w.je[0] = je.x; w.je[1] = je.y;
  } else
  flavor_weights_g(p, c, ncl, tropo ? bi.igas1[0] : bi.igas1[1], tropo ? bi.igas2[0] : bi.igas2[1],
                   tt.aux.ratio + (size_t)(itropo * t.nflav + (tropo ? bi.iflav[0] : bi.iflav[1])) * t.ntemp, jtemp,
                   p.cs.ftemp[c], p.cs.fpress[c], p.cs.col_dry[c], w);
  const int s_eta = t.ntemp, s_p = t.ntemp * t.neta;
  row0 = (jtemp - 1) + s_eta * (w.je[0] - 1) + s_p * (jpress - 2);
  row1 = jtemp + s_eta * (w.je[1] - 1) + s_p * (jpress - 2);
}

// Planck fractions (:627-631) of the n <= kGG g-points from 0-based table column g0
template <int VEC>
__device__ __forceinline__ void planck_fractions(const FusedParams& p, const TablesT& tt, const FlavW& w, int row0, int row1,
                                                 int g0, int n, Float (&pf)[kGG]) {
  const int s_eta = p.t.ntemp, s_p = p.t.ntemp * p.t.neta;
  if (n == kGG) interp3d_g<VEC, true>(tt.pfrac, tt.gp, row0, row1, s_eta, s_p, g0, n, w.fmj, (Float)1, (Float)1, pf);
  else interp3d_g<VEC, false>(tt.pfrac, tt.gp, row0, row1, s_eta, s_p, g0, n, w.fmj, (Float)1, (Float)1, pf);
}

// Grid: 1-D, block = (128 consecutive columns, chunk of layers, band), band fastest.  A thread marches down its
// layers; per layer the weights are computed once and the band's g-points are produced kGG at a time.
#ifndef RB_PLANCK_MINB
#define RB_PLANCK_MINB 4
#endif
template <int VEC, bool ABI = false>
__global__ void __launch_bounds__(kGThreads, RB_PLANCK_MINB) planck_g_kernel(const PlanckFusedParams q, const TablesT tt, int lay_per_chunk,
                                                                int nchunk) {
  const FusedParams& p = q.f;
  const rrtmgpb_gas_tables& t = p.t;
  const int ibnd = p.band0 + blockIdx.x % p.nband_sub;
  const int rest = blockIdx.x / p.nband_sub;
  const int ichunk = rest % nchunk;
  const int icol = (rest / nchunk) * blockDim.x + threadIdx.x;
  if (icol >= p.ncol) return;
  const int l0 = ichunk * lay_per_chunk, l1 = min(p.nlay, l0 + lay_per_chunk);
  const size_t ncol = p.ncol, ncl = ncol * p.nlay, nclp = ncol * (p.nlay + 1);
  const BandInfo bi = tt.aux.band[ibnd];
  const Float delta_r = (Float)1.0 / t.totplnk_delta;
  const Float* tab = t.totplnk + (size_t)t.nPlanckTemp * ibnd;

  for (int gS = bi.bS; gS <= bi.bE; gS += kPG) {
    const int n = min(kPG, bi.bE - gS + 1);
    Float pf_prev[kPG];
    if (l0 > 0) {  // the chunk's first level also needs the layer above it
      FlavW w;
      int row0, row1;
      planck_cell_weights<ABI>(p, tt, bi, icol + ncol * (size_t)(l0 - 1), ncl, w, row0, row1);
#pragma unroll
      for (int sub = 0; sub < kPG; sub += kGG) {
        if (sub < n) {
          Float pf[kGG];
          planck_fractions<VEC>(p, tt, w, row0, row1, gS - 1 + sub, min(kGG, n - sub), pf);
#pragma unroll
          for (int i = 0; i < kGG; ++i) pf_prev[sub + i] = pf[i];
        }
      }
    }
    for (int ilay = l0; ilay < l1; ++ilay) {
      const size_t c = icol + ncol * ilay;
      FlavW w;
      int row0, row1;
      planck_cell_weights<ABI>(p, tt, bi, c, ncl, w, row0, row1);
      const Float B_lay = planck_band_f(t, p.tlay[c], delta_r, tab);
      const Float B_lev = planck_band_f(t, q.tlev[c], delta_r, tab);
      const bool is_sfc = (ilay == q.sfc_lay - 1);
      Float B_sfc = 0, B_sfc1 = 0;
      if (is_sfc) {
        const Float ts = q.tsfc[icol];
        B_sfc = planck_band_f(t, ts, delta_r, tab);
        B_sfc1 = planck_band_f(t, ts + (Float)1.0, delta_r, tab);
      }
      Float* lay_c = q.lay_src + c + ncl * (size_t)(gS - 1 - p.gpt0);
      Float* lev_c = q.lev_src + c + nclp * (size_t)(gS - 1 - p.gpt0);
#pragma unroll
      for (int sub = 0; sub < kPG; sub += kGG) {
        if (sub < n) {
          const int ns = min(kGG, n - sub);
          Float pf[kGG];
          planck_fractions<VEC>(p, tt, w, row0, row1, gS - 1 + sub, ns, pf);
          auto emit = [&](auto full_tag) {  // full chunk: no per-g-point guards
            constexpr bool FULLC = decltype(full_tag)::value;
#pragma unroll
            for (int i = 0; i < kGG; ++i) {
              if (FULLC || i < ns) {
                // streaming stores: the source planes are next read by the solver, long after they left the caches
                __stcs(lay_c, pf[i] * B_lay);                                                    // :640
                __stcs(lev_c, (ilay == 0) ? pf[i] * B_lev : planck_geo_mean(pf_prev[sub + i], pf[i]) * B_lev);   // :695-701
                lay_c += ncl; lev_c += nclp;
                pf_prev[sub + i] = pf[i];
              }
            }
          };
          if (ns == kGG) emit(std::true_type{});
          else emit(std::false_type{});
          if (is_sfc) {  // one layer of the column only: kept out of the g-point loop above (:650-653)
            Float* sfc_c = q.sfc_src + icol + ncol * (size_t)(gS + sub - 1 - p.gpt0);
            Float* jac_c = q.sfc_source_Jac + icol + ncol * (size_t)(gS + sub - 1 - p.gpt0);
#pragma unroll
            for (int i = 0; i < kGG; ++i) {
              if (i < ns) {
                sfc_c[ncol * (size_t)i] = pf[i] * B_sfc;
                jac_c[ncol * (size_t)i] = pf[i] * (B_sfc1 - B_sfc);
              }
            }
          }
        }
      }
    }
    if (l1 == p.nlay) {  // :703-705
      const Float B_top = planck_band_f(t, q.tlev[icol + ncol * p.nlay], delta_r, tab);
#pragma unroll
      for (int i = 0; i < kPG; ++i)
        if (i < n) q.lev_src[icol + ncol * p.nlay + nclp * (size_t)(gS + i - 1 - p.gpt0)] = pf_prev[i] * B_top;
    }
  }
}

// The ROWS instantiations of gas_tau_g_kernel (double precision, 128-bit table layout, staging slots present), compiled in
// their own translation unit (abi/gas_optics_rows.cu): in the same cubin as the default instantiations they moved those
// 100-230 KB kernels to other code addresses and cost the replicated profile 3 % with instruction-for-instruction
// identical SASS (B200: LW tau 4.38 -> 4.52 ms, SW tau 7.47 -> 7.60 ms).
void launch_tau_rows(const FusedParams& p, const TablesT& tt, unsigned grid, size_t smem, bool sw, int kind, bool abi);

// out[r*pitch + g] = in[r + nrow*g]   (one-off table transposition)
static __global__ void transpose_table_kernel(const Float* __restrict__ in, Float* __restrict__ out, int nrow, int ng, int pitch) {
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
