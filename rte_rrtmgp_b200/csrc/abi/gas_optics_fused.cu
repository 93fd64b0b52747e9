// gas_optics_fused.cu - fused gas-optics fast path used by the device-resident frontend.
//
// The reference sequence for one gas_optics() call (mo_gas_optics_rrtmgp.F90:419-745,840-928) is
//   get_col_dry -> col_gas -> interpolation -> zero(tau) -> tau_absorption (3 RMW passes) -> [tau_rayleigh ->
//   combine_abs_and_rayleigh] -> compute_Planck_source, followed by clouds%increment(atmos) in the driver.
// Its intermediates (col_gas, jtemp, jpress, tropo, jeta, col_mix, fmajor, fminor, tau_rayleigh) are frontend
// locals; only atmos%tau/ssa/g and the Planck sources are visible to the caller.  Materialising them costs
// ~1.2 KB per (column, layer) that every later kernel reads back 16 times (once per band), and the unfused
// kernels are latency-bound on chains of dependent loads (profiles/r1_prof_v4_*).  Here:
//   1. cell_state_kernel      one thread per (col,lay): col_dry, jtemp, ftemp, jpress, fpress, tropo (33 B/cell)
//   2. gas_tau_fused_kernel   one thread per (col,lay) looping over bands: per-flavour weights recomputed in
//                             registers (2 divides + ~30 flops per band), major + minor absorption, Rayleigh,
//                             abs+Rayleigh combination and the by-band cloud increment, each output written once
//   3. planck_fused_kernel    one thread per (col,band) marching through layers (previous layer's Planck
//                             fraction in registers), weights recomputed likewise
// Arithmetic is the reference's, expression by expression (same cited lines as gas_optics_abi.cu), so results
// equal the unfused kernels' to rounding of FMA contraction; tests/test_allsky_parity.py checks fused vs
// unfused vs oracle.
#include <cuda.h>
#include <cstdint>
#include "../kernels/elementwise.cuh"
#include "../kernels/gas_optics_gfast.cuh"
#include <map>
#include <mutex>
#include <vector>
#include "rrtmgp_b200_ext.h"

using namespace rrtmgpb;

namespace {

constexpr int kFThreads = 128;
constexpr int kFG = 8;  // g-points per register chunk

// ---- per-cell state: mo_gas_optics_utils.F90:143-150 (col_dry), mo_gas_optics_rrtmgp_kernels.F90:99-118 ----
__global__ void __launch_bounds__(kFThreads) cell_state_kernel(const FusedParams p, Float m_dry, Float m_h2o,
                                                                Float avogad, Float grav) {
  const size_t ncl = (size_t)p.ncol * p.nlay;
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncl) return;
  const rrtmgpb_gas_tables& t = p.t;
  Float col_dry;
  if (p.col_dry_in) {
    col_dry = p.col_dry_in[c];
  } else {
    const Float vh2o = p.vmr[c + ncl * (size_t)(t.idx_h2o - 1)];
    const Float delta_plev = fabs(p.plev[c] - p.plev[c + p.ncol]);
    const Float fact = (Float)1 / ((Float)1 + vh2o);
    const Float m_air = (m_dry + m_h2o * vh2o) * fact;
    col_dry = (Float)10 * delta_plev * avogad * fact / ((Float)1000 * m_air * (Float)100 * grav);
  }
  const Float press_ref_trop = exp(t.press_ref_trop_log);
  const Float temp_ref_delta_inv = (Float)1.0 / t.temp_ref_delta;
  const Float press_ref_log_delta_inv = (Float)1.0 / t.press_ref_log_delta;
  const Float tl = p.tlay[c], pl = p.play[c];
  const int jtemp_ = (int)((tl - (t.temp_ref_min - t.temp_ref_delta)) * temp_ref_delta_inv);
  const int jtemp = min(t.ntemp - 1, max(1, jtemp_));
  const Float ftemp = (tl - __ldg(t.temp_ref + min(t.ntemp, max(1, jtemp_)) - 1)) * temp_ref_delta_inv;
  const Float locpress = (Float)1 + (log(pl) - __ldg(t.press_ref_log)) * press_ref_log_delta_inv;
  const Float jpress_aint = fmin((Float)(t.npres - 1), fmax((Float)1.0, trunc(locpress)));
  p.cs.col_dry[c] = col_dry;
  p.cs.jtemp[c] = jtemp;
  p.cs.ftemp[c] = ftemp;
  p.cs.jpress[c] = (int)jpress_aint;
  p.cs.fpress[c] = locpress - jpress_aint;
  p.cs.tropo[c] = pl > press_ref_trop;
}

template <int NT, int NE, int NP1>
struct FDims {
  int nt, ne, np1;
  __device__ __forceinline__ FDims(const rrtmgpb_gas_tables& t) : nt(NT ? NT : t.ntemp), ne(NE ? NE : t.neta), np1(NP1 ? NP1 : t.npres + 1) {}
  __device__ __forceinline__ int s_eta() const { return NT ? NT : nt; }
  __device__ __forceinline__ int s_p() const { return (NT && NE) ? NT * NE : nt * ne; }
  __device__ __forceinline__ int s_g() const { return (NT && NE && NP1) ? NT * NE * NP1 : nt * ne * np1; }
};

struct MinorSetO {  // loader-layout tables (legacy kernels)
  int n;
  const int2* band_range;
  const Float* kminor;
  const int *limits_gpt, *idx_minor, *idx_scaling, *kminor_start;
  const Bool *scales_with_density, *scale_by_complement;
};

template <int NT, int NE, int NP1, bool SW>
__global__ void __launch_bounds__(kFThreads, 4) gas_tau_fused_kernel(const FusedParams p) {
  const size_t ncl = (size_t)p.ncol * p.nlay;
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncl) return;
  const rrtmgpb_gas_tables& t = p.t;
  const FDims<NT, NE, NP1> td(t);
  const int s_eta = td.s_eta(), s_p = td.s_p(), s_g = td.s_g();
  const Float col_dry = p.cs.col_dry[c], ftemp = p.cs.ftemp[c], fpress = p.cs.fpress[c];
  const int jtemp = p.cs.jtemp[c], jpress0 = p.cs.jpress[c];
  const bool tropo = p.cs.tropo[c];
  const int itropo = tropo ? 0 : 1;
  const int jpress = jpress0 + itropo + 1;  // :390
  const Float play = p.play[c], tlay = p.tlay[c];
  const MinorSetO ms = tropo ? MinorSetO{t.nminorlower, p.range_lower, t.kminor_lower, t.minor_limits_gpt_lower,
                                       t.idx_minor_lower, t.idx_minor_scaling_lower, t.kminor_start_lower,
                                       t.minor_scales_with_density_lower, t.scale_by_complement_lower}
                            : MinorSetO{t.nminorupper, p.range_upper, t.kminor_upper, t.minor_limits_gpt_upper,
                                       t.idx_minor_upper, t.idx_minor_scaling_upper, t.kminor_start_upper,
                                       t.minor_scales_with_density_upper, t.scale_by_complement_upper};
  const Float amount_rayl = SW ? col_gas_of(p, c, ncl, t.idx_h2o, col_dry) + col_dry : (Float)0;  // :559
  const Float* krayl = SW ? t.krayl + (size_t)s_p * t.ngpt * itropo : nullptr;
  int iflav_cur = -1;
  FlavW w;
  for (int ibnd = 0; ibnd < t.nbnd; ++ibnd) {
    const int bS = __ldg(t.band_lims_gpt + 2 * ibnd), bE = __ldg(t.band_lims_gpt + 2 * ibnd + 1);
    const int iflav = __ldg(t.gpoint_flavor + itropo + 2 * (bS - 1)) - 1;  // :384 band's first g-point
    if (iflav != iflav_cur) {
      flavor_weights(p, c, ncl, iflav, itropo, jtemp, ftemp, fpress, col_dry, w);
      iflav_cur = iflav;
    }
    // cloud properties of this (cell, band), if the caller wants them added (mo_optical_props.F90:956-1000)
    Float ct = 0, cw = 0, cg = 0;
    if (p.cld_kind) {
      const size_t cb = c + ncl * (size_t)ibnd;
      ct = p.cld_tau[cb];
      if (p.cld_kind == 2) { cw = p.cld_ssa[cb]; cg = p.cld_g[cb]; }
    }
    const int2 range = ms.band_range[ibnd];
    for (int gS = bS; gS <= bE; gS += kFG) {
      const int gE = min(bE, gS + kFG - 1);
      const Float* k0 = t.kmajor + (jtemp - 1) + s_eta * (w.je[0] - 1) + s_p * (jpress - 2) + (long long)s_g * (gS - 1);
      const Float* k1 = t.kmajor + jtemp + s_eta * (w.je[1] - 1) + s_p * (jpress - 2) + (long long)s_g * (gS - 1);
      Float acc[kFG];
#pragma unroll
      for (int i = 0; i < kFG; ++i) {
        acc[i] = 0;
        if (gS + i <= gE) {
          const int go = s_g * i;
          const Float major =  // interpolate3D_byflav :791-801
              w.cm[0] * (w.fmj[0] * __ldg(k0 + go) + w.fmj[1] * __ldg(k0 + go + s_eta) +
                         w.fmj[2] * __ldg(k0 + go + s_p) + w.fmj[3] * __ldg(k0 + go + s_p + s_eta)) +
              w.cm[1] * (w.fmj[4] * __ldg(k1 + go) + w.fmj[5] * __ldg(k1 + go + s_eta) +
                         w.fmj[6] * __ldg(k1 + go + s_p) + w.fmj[7] * __ldg(k1 + go + s_p + s_eta));
          acc[i] = (Float)0 + major;  // :391 on a zeroed tau
        }
      }
      // ---- minor absorbers touching this chunk (:451-498) ----
      for (int imnr = range.x; imnr <= range.y; ++imnr) {
        const int mS = __ldg(ms.limits_gpt + 2 * imnr), mE = __ldg(ms.limits_gpt + 2 * imnr + 1);
        if (mE < gS || mS > gE) continue;
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
        // otherwise recompute its eta weights.  (Selecting between two FlavW objects through a pointer would push
        // both into local memory, so the few values needed are copied into scalars instead.)
        Float a0 = w.fmn[0], a1 = w.fmn[1], a2 = w.fmn[2], a3 = w.fmn[3];
        int je0 = w.je[0], je1 = w.je[1];
        if (iflav_m != iflav) {
          FlavW wm;
          flavor_weights(p, c, ncl, iflav_m, itropo, jtemp, ftemp, fpress, col_dry, wm);
          a0 = wm.fmn[0]; a1 = wm.fmn[1]; a2 = wm.fmn[2]; a3 = wm.fmn[3];
          je0 = wm.je[0]; je1 = wm.je[1];
        }
        const long long kcol0 = (long long)__ldg(ms.kminor_start + imnr) + (gS - mS) - 1;
        const Float* m0 = ms.kminor + (jtemp - 1) + s_eta * (je0 - 1) + (long long)s_p * kcol0;
        const Float* m1 = ms.kminor + jtemp + s_eta * (je1 - 1) + (long long)s_p * kcol0;
#pragma unroll
        for (int i = 0; i < kFG; ++i) {
          const int g = gS + i;
          if (g >= mS && g <= mE && g <= gE) {
            const int ko = s_p * i;
            const Float kint = a0 * __ldg(m0 + ko) + a1 * __ldg(m0 + ko + s_eta) +
                               a2 * __ldg(m1 + ko) + a3 * __ldg(m1 + ko + s_eta);  // :757-760
            acc[i] = acc[i] + scaling * kint;                                       // :493
          }
        }
      }
      // ---- Rayleigh, combination, cloud increment, store ----
      const Float* r0 = SW ? krayl + (jtemp - 1) + s_eta * (w.je[0] - 1) + (long long)s_p * (gS - 1) : nullptr;
      const Float* r1 = SW ? krayl + jtemp + s_eta * (w.je[1] - 1) + (long long)s_p * (gS - 1) : nullptr;
      Float* tau_c = p.tau + c + ncl * (size_t)(gS - 1);
#pragma unroll
      for (int i = 0; i < kFG; ++i) {
        if (gS + i <= gE) {
          Float tt = acc[i], ss = 0, gg = 0;
          if (SW) {
            const int ko = s_p * i;
            const Float kr = w.fmn[0] * __ldg(r0 + ko) + w.fmn[1] * __ldg(r0 + ko + s_eta) +
                             w.fmn[2] * __ldg(r1 + ko) + w.fmn[3] * __ldg(r1 + ko + s_eta);
            const Float tray = kr * amount_rayl;                                             // :558-559
            tt = acc[i] + tray;                                                              // combine :1986-1994
            ss = (tt > (Float)2 * (Float)RB_TINY) ? tray / tt : (Float)0;
          }
          const Float eps3 = (Float)3.0 * (Float)RB_TINY;  // mo_optical_props_kernels.F90:38
          if (p.op_kind == 1) {
            if (p.cld_kind == 1) tt = tt + ct;                           // inc_1scalar_by_1scalar_bybnd :379
            else if (p.cld_kind == 2) tt = tt + ct * ((Float)1 - cw);    // inc_1scalar_by_2stream_bybnd :398
          } else {
            if (p.cld_kind == 1) {                                       // inc_2stream_by_1scalar_bybnd :440-442
              const Float tau12 = tt + ct;
              ss = tt * ss / fmax(eps3, tau12);
              tt = tau12;
            } else if (p.cld_kind == 2) {                                // inc_2stream_by_2stream_bybnd :468-477
              const Float tau12 = tt + ct;
              const Float tauscat12 = tt * ss + ct * cw;
              gg = (tt * ss * gg + ct * cw * cg) / fmax(eps3, tauscat12);
              ss = tauscat12 / fmax(eps3, tau12);
              tt = tau12;
            }
          }
          tau_c[ncl * i] = tt;
          if (p.op_kind == 2) {
            p.ssa[c + ncl * (size_t)(gS + i - 1)] = ss;
            p.g[c + ncl * (size_t)(gS + i - 1)] = gg;
          }
        }
      }
    }
  }
}

// ---- Planck sources: compute_Planck_source :568-710 with weights recomputed per (cell, band) ----
template <int NT, int NE, int NP1>
__global__ void __launch_bounds__(kFThreads, 4) planck_fused_kernel(const PlanckFusedParams q) {
  const FusedParams& p = q.f;
  const rrtmgpb_gas_tables& t = p.t;
  const int icol = blockIdx.x * blockDim.x + threadIdx.x;
  if (icol >= p.ncol) return;
  const FDims<NT, NE, NP1> td(t);
  const int s_eta = td.s_eta(), s_p = td.s_p(), s_g = td.s_g();
  const int ibnd = blockIdx.y;
  const size_t ncol = p.ncol, ncl = ncol * p.nlay, nclp = ncol * (p.nlay + 1);
  const int bS = __ldg(t.band_lims_gpt + 2 * ibnd), bE = __ldg(t.band_lims_gpt + 2 * ibnd + 1);
  const Float delta_r = (Float)1.0 / t.totplnk_delta;
  const Float* tab = t.totplnk + (size_t)t.nPlanckTemp * ibnd;
  for (int gS = bS; gS <= bE; gS += kFG) {
    const int gE = min(bE, gS + kFG - 1);
    Float pf_prev[kFG];
#pragma unroll
    for (int i = 0; i < kFG; ++i) pf_prev[i] = 0;
    for (int ilay = 0; ilay < p.nlay; ++ilay) {
      const size_t c = icol + ncol * ilay;
      const int itropo = p.cs.tropo[c] ? 0 : 1;
      const int iflav = __ldg(t.gpoint_flavor + itropo + 2 * (bS - 1)) - 1;
      const int jtemp = p.cs.jtemp[c];
      const int jpress = p.cs.jpress[c] + itropo + 1;
      FlavW w;
      flavor_weights(p, c, ncl, iflav, itropo, jtemp, p.cs.ftemp[c], p.cs.fpress[c], p.cs.col_dry[c], w);
      const Float* k0 = t.planck_frac + (jtemp - 1) + s_eta * (w.je[0] - 1) + s_p * (jpress - 2) + (long long)s_g * (gS - 1);
      const Float* k1 = t.planck_frac + jtemp + s_eta * (w.je[1] - 1) + s_p * (jpress - 2) + (long long)s_g * (gS - 1);
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
      for (int i = 0; i < kFG; ++i) {
        if (gS + i <= gE) {
          const int go = s_g * i;
          const Float pf = (Float)1 * (w.fmj[0] * __ldg(k0 + go) + w.fmj[1] * __ldg(k0 + go + s_eta) +
                                       w.fmj[2] * __ldg(k0 + go + s_p) + w.fmj[3] * __ldg(k0 + go + s_p + s_eta)) +
                           (Float)1 * (w.fmj[4] * __ldg(k1 + go) + w.fmj[5] * __ldg(k1 + go + s_eta) +
                                       w.fmj[6] * __ldg(k1 + go + s_p) + w.fmj[7] * __ldg(k1 + go + s_p + s_eta));
          lay_c[ncl * i] = pf * B_lay;
          lev_c[nclp * i] = (ilay == 0) ? pf * B_lev : sqrt(pf_prev[i] * pf) * B_lev;
          if (is_sfc) {
            q.sfc_src[icol + ncol * (size_t)(gS + i - 1)] = pf * B_sfc;
            q.sfc_source_Jac[icol + ncol * (size_t)(gS + i - 1)] = pf * (B_sfc1 - B_sfc);
          }
          pf_prev[i] = pf;
        }
      }
    }
    const Float B_top = planck_band_f(t, q.tlev[icol + ncol * p.nlay], delta_r, tab);
#pragma unroll
    for (int i = 0; i < kFG; ++i)
      if (gS + i <= gE) q.lev_src[icol + ncol * p.nlay + nclp * (size_t)(gS + i - 1)] = pf_prev[i] * B_top;
  }
}


// =====================================================================================================
// TMA-staged variants.  A block of 128 consecutive cells sits (almost always) in one layer, so the table
// entries it can touch for a band are a small BOX: [tmin, tmin+TB) temperatures x all eta x [pmin, pmin+PB)
// pressure rows x the band's g-points (4 x 9 x 4 x 16 doubles = 18 KB).  One elected thread fetches that box
// with ONE cp.async.bulk.tensor.4d (TMA, completion on an mbarrier), double-buffered across bands, and every
// table read of the block becomes a shared-memory read.  Blocks whose cells do not fit one box (a block that
// straddles distant layers) fall back to the global-load path - same arithmetic either way.
// =====================================================================================================
constexpr int kTB = 4, kPB = 4, kGB = 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::
          "r"(smem_u32(smem_dst)), "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// block-wide min/max of the table rows the block's cells touch
struct BoxRange { int tmin, tmax, pmin, pmax; };
__device__ __forceinline__ void box_reduce(int* red, int jtemp, int prow_lo, bool first_call) {
  // red[0]=tmin red[1]=tmax red[2]=pmin red[3]=pmax ; caller syncs before and after
  atomicMin(&red[0], jtemp);
  atomicMax(&red[1], jtemp + 1);
  atomicMin(&red[2], prow_lo);
  atomicMax(&red[3], prow_lo + 1);
  (void)first_call;
}

template <int NE, bool SW>
__global__ void __launch_bounds__(kFThreads, 3) gas_tau_tma_kernel(const FusedParams p,
                                                                  const __grid_constant__ CUtensorMap tm_kmajor) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Float* box = reinterpret_cast<Float*>(smem_raw);  // [2][kGB][kPB][NE][kTB]
  constexpr int kBoxElems = kGB * kPB * NE * kTB;
  __shared__ uint64_t mbar[2];
  __shared__ int red[4];
  const size_t ncl = (size_t)p.ncol * p.nlay;
  const size_t c_raw = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = c_raw < ncl;
  const size_t c = valid ? c_raw : ncl - 1;
  const rrtmgpb_gas_tables& t = p.t;
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_fence_init();
    red[0] = 1 << 30; red[1] = -1; red[2] = 1 << 30; red[3] = -1;
  }
  const Float col_dry = p.cs.col_dry[c], ftemp = p.cs.ftemp[c], fpress = p.cs.fpress[c];
  const int jtemp = p.cs.jtemp[c], jpress0 = p.cs.jpress[c];
  const bool tropo = p.cs.tropo[c];
  const int itropo = tropo ? 0 : 1;
  const int jpress = jpress0 + itropo + 1;  // :390; table rows jpress-1 and jpress (1-based)
  __syncthreads();
  box_reduce(red, jtemp, jpress - 1, true);
  __syncthreads();
  // the box must start at an EVEN temperature index: TMA needs a 16-byte aligned global start address and the
  // elements are 8 bytes (an odd start raises an illegal-instruction fault)
  const int tmin = ((red[0] - 1) & ~1) + 1, pmin = red[2];
  const bool fits = (red[1] - tmin + 1 <= kTB) && (red[3] - pmin + 1 <= kPB);

  const int s_eta_g = t.ntemp, s_p_g = t.ntemp * t.neta;
  const long long s_g_g = (long long)s_p_g * (t.npres + 1);
  const Float play = p.play[c], tlay = p.tlay[c];
  const MinorSetO ms = tropo ? MinorSetO{t.nminorlower, p.range_lower, t.kminor_lower, t.minor_limits_gpt_lower,
                                       t.idx_minor_lower, t.idx_minor_scaling_lower, t.kminor_start_lower,
                                       t.minor_scales_with_density_lower, t.scale_by_complement_lower}
                            : MinorSetO{t.nminorupper, p.range_upper, t.kminor_upper, t.minor_limits_gpt_upper,
                                       t.idx_minor_upper, t.idx_minor_scaling_upper, t.kminor_start_upper,
                                       t.minor_scales_with_density_upper, t.scale_by_complement_upper};
  const Float amount_rayl = SW ? col_gas_of(p, c, ncl, t.idx_h2o, col_dry) + col_dry : (Float)0;
  const Float* krayl = SW ? t.krayl + (size_t)s_p_g * t.ngpt * itropo : nullptr;

  // flattened list of (band, chunk) work items; item k uses stage k&1
  auto issue = [&](int ibnd, int gS, int k) {
    if (fits && tid == 0) {
      mbar_expect_tx(&mbar[k & 1], (uint32_t)(kBoxElems * sizeof(Float)));
      tma_load_4d(box + (size_t)(k & 1) * kBoxElems, &tm_kmajor, &mbar[k & 1], tmin - 1, 0, pmin - 1, gS - 1);
    }
  };
  int k = 0;
  {
    const int bS0 = __ldg(t.band_lims_gpt);
    issue(0, bS0, 0);
  }
  int iflav_cur = -1;
  FlavW w;
  for (int ibnd = 0; ibnd < t.nbnd; ++ibnd) {
    const int bS = __ldg(t.band_lims_gpt + 2 * ibnd), bE = __ldg(t.band_lims_gpt + 2 * ibnd + 1);
    const int iflav = __ldg(t.gpoint_flavor + itropo + 2 * (bS - 1)) - 1;
    if (iflav != iflav_cur) {
      flavor_weights(p, c, ncl, iflav, itropo, jtemp, ftemp, fpress, col_dry, w);
      iflav_cur = iflav;
    }
    Float ct = 0, cw = 0, cg = 0;
    if (p.cld_kind) {
      const size_t cb = c + ncl * (size_t)ibnd;
      ct = p.cld_tau[cb];
      if (p.cld_kind == 2) { cw = p.cld_ssa[cb]; cg = p.cld_g[cb]; }
    }
    const int2 range = ms.band_range[ibnd];
    for (int gB = bS; gB <= bE; gB += kGB, ++k) {
      // prefetch the next work item into the other stage (its previous contents were released by the
      // __syncthreads() that ended the previous item)
      {
        int nb = ibnd, ng = gB + kGB;
        if (ng > bE) { nb = ibnd + 1; ng = (nb < t.nbnd) ? __ldg(t.band_lims_gpt + 2 * nb) : 0; }
        if (nb < t.nbnd) issue(nb, ng, k + 1);
      }
      if (fits) mbar_wait(&mbar[k & 1], (uint32_t)((k >> 1) & 1));
      const Float* bx = box + (size_t)(k & 1) * kBoxElems;
      const int gBE = min(bE, gB + kGB - 1);
      for (int gS = gB; gS <= gBE; gS += kFG) {
        const int gE = min(gBE, gS + kFG - 1);
        Float acc[kFG];
        if (fits) {
          // shared-memory box: element (t, e, p, g) at t + kTB*(e + NE*(p + kPB*g)), origin (tmin, 1, pmin, gB)
          const Float* k0 = bx + (jtemp - tmin) + kTB * ((w.je[0] - 1) + NE * ((jpress - 1 - pmin) + kPB * (gS - gB)));
          const Float* k1 = bx + (jtemp + 1 - tmin) + kTB * ((w.je[1] - 1) + NE * ((jpress - 1 - pmin) + kPB * (gS - gB)));
          constexpr int s_eta = kTB, s_p = kTB * NE, s_g = kTB * NE * kPB;
#pragma unroll
          for (int i = 0; i < kFG; ++i) {
            acc[i] = 0;
            if (gS + i <= gE) {
              const int go = s_g * i;
              const Float major =
                  w.cm[0] * (w.fmj[0] * k0[go] + w.fmj[1] * k0[go + s_eta] + w.fmj[2] * k0[go + s_p] + w.fmj[3] * k0[go + s_p + s_eta]) +
                  w.cm[1] * (w.fmj[4] * k1[go] + w.fmj[5] * k1[go + s_eta] + w.fmj[6] * k1[go + s_p] + w.fmj[7] * k1[go + s_p + s_eta]);
              acc[i] = (Float)0 + major;
            }
          }
        } else {
          const Float* k0 = t.kmajor + (jtemp - 1) + s_eta_g * (w.je[0] - 1) + s_p_g * (jpress - 2) + s_g_g * (gS - 1);
          const Float* k1 = t.kmajor + jtemp + s_eta_g * (w.je[1] - 1) + s_p_g * (jpress - 2) + s_g_g * (gS - 1);
#pragma unroll
          for (int i = 0; i < kFG; ++i) {
            acc[i] = 0;
            if (gS + i <= gE) {
              const long long go = s_g_g * i;
              const Float major =
                  w.cm[0] * (w.fmj[0] * __ldg(k0 + go) + w.fmj[1] * __ldg(k0 + go + s_eta_g) +
                             w.fmj[2] * __ldg(k0 + go + s_p_g) + w.fmj[3] * __ldg(k0 + go + s_p_g + s_eta_g)) +
                  w.cm[1] * (w.fmj[4] * __ldg(k1 + go) + w.fmj[5] * __ldg(k1 + go + s_eta_g) +
                             w.fmj[6] * __ldg(k1 + go + s_p_g) + w.fmj[7] * __ldg(k1 + go + s_p_g + s_eta_g));
              acc[i] = (Float)0 + major;
            }
          }
        }
        // ---- minor absorbers (:451-498), read-only path ----
        for (int imnr = range.x; imnr <= range.y; ++imnr) {
          const int mS = __ldg(ms.limits_gpt + 2 * imnr), mE = __ldg(ms.limits_gpt + 2 * imnr + 1);
          if (mE < gS || mS > gE) continue;
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
          const int iflav_m = __ldg(t.gpoint_flavor + itropo + 2 * (mS - 1)) - 1;
          Float a0 = w.fmn[0], a1 = w.fmn[1], a2 = w.fmn[2], a3 = w.fmn[3];
          int je0 = w.je[0], je1 = w.je[1];
          if (iflav_m != iflav) {
            FlavW wm;
            flavor_weights(p, c, ncl, iflav_m, itropo, jtemp, ftemp, fpress, col_dry, wm);
            a0 = wm.fmn[0]; a1 = wm.fmn[1]; a2 = wm.fmn[2]; a3 = wm.fmn[3];
            je0 = wm.je[0]; je1 = wm.je[1];
          }
          const long long kcol0 = (long long)__ldg(ms.kminor_start + imnr) + (gS - mS) - 1;
          const Float* m0 = ms.kminor + (jtemp - 1) + s_eta_g * (je0 - 1) + (long long)s_p_g * kcol0;
          const Float* m1 = ms.kminor + jtemp + s_eta_g * (je1 - 1) + (long long)s_p_g * kcol0;
#pragma unroll
          for (int i = 0; i < kFG; ++i) {
            const int g = gS + i;
            if (g >= mS && g <= mE && g <= gE) {
              const int ko = s_p_g * i;
              const Float kint = a0 * __ldg(m0 + ko) + a1 * __ldg(m0 + ko + s_eta_g) +
                                 a2 * __ldg(m1 + ko) + a3 * __ldg(m1 + ko + s_eta_g);
              acc[i] = acc[i] + scaling * kint;
            }
          }
        }
        const Float* r0 = SW ? krayl + (jtemp - 1) + s_eta_g * (w.je[0] - 1) + (long long)s_p_g * (gS - 1) : nullptr;
        const Float* r1 = SW ? krayl + jtemp + s_eta_g * (w.je[1] - 1) + (long long)s_p_g * (gS - 1) : nullptr;
#pragma unroll
        for (int i = 0; i < kFG; ++i) {
          if (gS + i <= gE) {
            Float tt = acc[i], ss = 0, gg = 0;
            if (SW) {
              const int ko = s_p_g * i;
              const Float kr = w.fmn[0] * __ldg(r0 + ko) + w.fmn[1] * __ldg(r0 + ko + s_eta_g) +
                               w.fmn[2] * __ldg(r1 + ko) + w.fmn[3] * __ldg(r1 + ko + s_eta_g);
              const Float tray = kr * amount_rayl;
              tt = acc[i] + tray;
              ss = (tt > (Float)2 * (Float)RB_TINY) ? tray / tt : (Float)0;
            }
            const Float eps3 = (Float)3.0 * (Float)RB_TINY;
            if (p.op_kind == 1) {
              if (p.cld_kind == 1) tt = tt + ct;
              else if (p.cld_kind == 2) tt = tt + ct * ((Float)1 - cw);
            } else {
              if (p.cld_kind == 1) {
                const Float tau12 = tt + ct;
                ss = tt * ss / fmax(eps3, tau12);
                tt = tau12;
              } else if (p.cld_kind == 2) {
                const Float tau12 = tt + ct;
                const Float tauscat12 = tt * ss + ct * cw;
                gg = (tt * ss * gg + ct * cw * cg) / fmax(eps3, tauscat12);
                ss = tauscat12 / fmax(eps3, tau12);
                tt = tau12;
              }
            }
            if (valid) {
              const size_t o = c + ncl * (size_t)(gS + i - 1);
              p.tau[o] = tt;
              if (p.op_kind == 2) { p.ssa[o] = ss; p.g[o] = gg; }
            }
          }
        }
      }
      __syncthreads();  // every thread is done with this stage before it is refilled
    }
  }
}


__global__ void band_ranges_kernel(int nbnd, const int* band_lims_gpt, int nminor, const int* limits_gpt, int2* out) {
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

struct Workspace {
  CellState cs;
  int2* ranges;
  void* block;
};

// physical constants (kept in util_abi.cu; mirrored here through the setter below)
double g_m_dry = 0.028964, g_grav = 9.80665;
const double k_m_h2o = 0.018016, k_avogad = 6.02214076e23;

Workspace prepare(FusedParams& p) {
  const size_t ncl = (size_t)p.ncol * p.nlay;
  const rrtmgpb_gas_tables& t = p.t;
  // one pool allocation: 3 doubles, 2 ints, 1 byte per cell + the band ranges
  const size_t bytes = ncl * (3 * sizeof(Float) + 2 * sizeof(int) + 8) + 2 * (size_t)t.nbnd * sizeof(int2) + 64;
  Workspace w;
  w.block = dev_alloc(bytes);
  Float* f = static_cast<Float*>(w.block);
  w.cs.col_dry = f; w.cs.ftemp = f + ncl; w.cs.fpress = f + 2 * ncl;
  int* ii = reinterpret_cast<int*>(f + 3 * ncl);
  w.cs.jtemp = ii; w.cs.jpress = ii + ncl;
  w.ranges = reinterpret_cast<int2*>(ii + 2 * ncl);
  w.cs.tropo = reinterpret_cast<Bool*>(w.ranges + 2 * t.nbnd);
  p.cs = w.cs;
  p.range_lower = w.ranges;
  p.range_upper = w.ranges + t.nbnd;
  {
    KernelTimer timer("gas_cell_state");
    band_ranges_kernel<<<ceil_div(t.nbnd, 32), 32, 0, stream()>>>(t.nbnd, t.band_lims_gpt, t.nminorlower,
                                                                    t.minor_limits_gpt_lower, w.ranges);
    RB_LAUNCH_CHECK();
    band_ranges_kernel<<<ceil_div(t.nbnd, 32), 32, 0, stream()>>>(t.nbnd, t.band_lims_gpt, t.nminorupper,
                                                                    t.minor_limits_gpt_upper, w.ranges + t.nbnd);
    RB_LAUNCH_CHECK();
    cell_state_kernel<<<ceil_div((long long)ncl, kFThreads), kFThreads, 0, stream()>>>(
        p, (Float)g_m_dry, (Float)k_m_h2o, (Float)k_avogad, (Float)g_grav);
    RB_LAUNCH_CHECK();
  }
  return w;
}

bool std_dims(const rrtmgpb_gas_tables& t) { return t.ntemp == 14 && t.neta == 9 && t.npres == 59; }

// ---- TMA descriptor for a (ntemp, neta, npres+1, ngpt) table with box (kTB, neta, kPB, kGB) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    cudaGetLastError();
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// TMA staging is opt-in (rrtmgpb_set_tma_staging(1) or RRTMGPB_TMA=1): measured on B200 it is slower than the
// L1-cached loads whenever neighbouring columns share table rows, which is the common case.
int g_use_tma = -1;  // -1: decide from the environment
bool tma_enabled() {
  if (g_use_tma < 0) { const char* e = std::getenv("RRTMGPB_TMA"); g_use_tma = (e && e[0] == '1') ? 1 : 0; }
  return g_use_tma == 1;
}

// returns false when the table cannot be described (strides not multiples of 16 B, misaligned base, no driver entry)
bool make_table_tmap(CUtensorMap* tm, const Float* base, const rrtmgpb_gas_tables& t) {
  if (sizeof(Float) != 8 || !tma_enabled()) return false;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc || t.neta != 9 || (t.ntemp * sizeof(Float)) % 16 != 0 || (reinterpret_cast<uintptr_t>(base) % 16) != 0)
    return false;
  const cuuint64_t dims[4] = {(cuuint64_t)t.ntemp, (cuuint64_t)t.neta, (cuuint64_t)(t.npres + 1), (cuuint64_t)t.ngpt};
  const cuuint64_t strides[3] = {(cuuint64_t)t.ntemp * 8, (cuuint64_t)t.ntemp * t.neta * 8,
                                 (cuuint64_t)t.ntemp * t.neta * (t.npres + 1) * 8};
  const cuuint32_t box[4] = {(cuuint32_t)kTB, (cuuint32_t)t.neta, (cuuint32_t)kPB, (cuuint32_t)kGB};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<Float*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}


// ---- g-point-fastest table copies, built once per k-distribution (kernels/gas_optics_gfast.cuh) ----
struct TableCacheEntry {
  TablesT tt;
  std::vector<void*> owned;
  int ntemp, neta, npres, ngpt, nkl, nku;
};
std::mutex g_tc_mutex;
std::map<const void*, TableCacheEntry> g_table_cache;  // key: the loader-layout kmajor pointer

Float* transposed(const Float* in, int nrow, int ng, int pitch, std::vector<void*>& owned) {
  if (!in || nrow <= 0 || ng <= 0) return nullptr;
  Float* out = static_cast<Float*>(dev_alloc((size_t)nrow * pitch * sizeof(Float)));
  RB_CUDA_CHECK(cudaMemsetAsync(out, 0, (size_t)nrow * pitch * sizeof(Float), stream()));
  dim3 grid(ceil_div(nrow, 32), ceil_div(ng, 32)), block(32, 8);
  transpose_table_kernel<<<grid, block, 0, stream()>>>(in, out, nrow, ng, pitch);
  RB_LAUNCH_CHECK();
  owned.push_back(out);
  return out;
}

std::vector<int> host_ints(const int* dev, size_t n) {
  std::vector<int> h(n);
  if (n) RB_CUDA_CHECK(cudaMemcpyAsync(h.data(), dev, n * sizeof(int), cudaMemcpyDeviceToHost, stream()));
  RB_CUDA_CHECK(cudaStreamSynchronize(stream()));
  return h;
}

// 128-bit table loads need every band and every minor-contributor interval to start on an even 0-based
// column and to have even length (true for the rrtmgp-data k-distributions: 16 g-points per band)
bool intervals_even(const std::vector<int>& lims, const std::vector<int>* start) {
  for (size_t i = 0; i + 1 < lims.size(); i += 2) {
    if ((lims[i] & 1) == 0 || ((lims[i + 1] - lims[i] + 1) & 1)) return false;
    if (start && ((*start)[i / 2] & 1) == 0) return false;
  }
  return true;
}

TablesT tables_gfast(const rrtmgpb_gas_tables& t) {
  std::lock_guard<std::mutex> lock(g_tc_mutex);
  auto it = g_table_cache.find(t.kmajor);
  if (it != g_table_cache.end()) {
    const TableCacheEntry& e = it->second;
    if (e.ntemp == t.ntemp && e.neta == t.neta && e.npres == t.npres && e.ngpt == t.ngpt && e.nkl == t.nminorklower &&
        e.nku == t.nminorkupper)
      return e.tt;
    for (void* q : e.owned) dev_free(q);
    g_table_cache.erase(it);
  }
  TableCacheEntry e;
  e.ntemp = t.ntemp; e.neta = t.neta; e.npres = t.npres; e.ngpt = t.ngpt; e.nkl = t.nminorklower; e.nku = t.nminorkupper;
  const int tn = t.ntemp * t.neta, rows = tn * (t.npres + 1);
  TablesT& tt = e.tt;
  tt.gp = (t.ngpt + 1) & ~1;
  tt.nkl = (t.nminorklower + 1) & ~1;
  tt.nku = (t.nminorkupper + 1) & ~1;
  tt.kmajor = transposed(t.kmajor, rows, t.ngpt, tt.gp, e.owned);
  tt.pfrac = transposed(t.planck_frac, rows, t.ngpt, tt.gp, e.owned);
  tt.kminor_lower = transposed(t.kminor_lower, tn, t.nminorklower, tt.nkl, e.owned);
  tt.kminor_upper = transposed(t.kminor_upper, tn, t.nminorkupper, tt.nku, e.owned);
  tt.krayl = nullptr;  // (ntemp, neta, ngpt, 2): the two tropo slices are transposed separately
  if (t.krayl) {
    Float* kr = static_cast<Float*>(dev_alloc((size_t)2 * tn * tt.gp * sizeof(Float)));
    RB_CUDA_CHECK(cudaMemsetAsync(kr, 0, (size_t)2 * tn * tt.gp * sizeof(Float), stream()));
    for (int itropo = 0; itropo < 2; ++itropo) {
      dim3 grid(ceil_div(tn, 32), ceil_div(t.ngpt, 32)), block(32, 8);
      transpose_table_kernel<<<grid, block, 0, stream()>>>(t.krayl + (size_t)tn * t.ngpt * itropo,
                                                           kr + (size_t)tn * tt.gp * itropo, tn, t.ngpt, tt.gp);
      RB_LAUNCH_CHECK();
    }
    e.owned.push_back(kr);
    tt.krayl = kr;
  }
  const std::vector<int> bl = host_ints(t.band_lims_gpt, 2 * (size_t)t.nbnd);
  const std::vector<int> ll = host_ints(t.minor_limits_gpt_lower, 2 * (size_t)t.nminorlower);
  const std::vector<int> lu = host_ints(t.minor_limits_gpt_upper, 2 * (size_t)t.nminorupper);
  const std::vector<int> sl = host_ints(t.kminor_start_lower, (size_t)t.nminorlower);
  const std::vector<int> su = host_ints(t.kminor_start_upper, (size_t)t.nminorupper);
  tt.vec = (sizeof(Float) == 8 && intervals_even(bl, nullptr) && intervals_even(ll, &sl) && intervals_even(lu, &su)) ? 2 : 1;
  g_table_cache[t.kmajor] = e;
  return e.tt;
}

int g_gas_kernels = -1;  // 0: legacy loader-layout kernels, 1: g-point-fastest kernels (default)
bool gfast_enabled() {
  if (g_gas_kernels < 0) { const char* v = std::getenv("RRTMGPB_GAS_KERNELS"); g_gas_kernels = (v && v[0] == '0') ? 0 : 1; }
  return g_gas_kernels == 1;
}

}  // namespace

namespace rrtmgpb {
// called by rrtmgpb_mem_free(): a k-distribution that is being released takes its transposed copies with it
void table_cache_release(const void* key) {
  std::lock_guard<std::mutex> lock(g_tc_mutex);
  auto it = g_table_cache.find(key);
  if (it == g_table_cache.end()) return;
  for (void* q : it->second.owned) dev_free(q);
  g_table_cache.erase(it);
}
void fused_set_constants(double grav, double m_dry) { g_grav = grav; g_m_dry = m_dry; }
}

extern "C" {

void rrtmgpb_set_tma_staging(int on) { g_use_tma = on ? 1 : 0; }

void rrtmgpb_gas_optics_fused(const rrtmgpb_gas_tables* t, int ncol, int nlay, const Float* play, const Float* plev,
                              const Float* tlay, const Float* vmr, const Float* col_dry, int op_kind, Float* tau,
                              Float* ssa, Float* g, int cld_kind, const Float* cld_tau, const Float* cld_ssa,
                              const Float* cld_g, const Float* tlev, const Float* tsfc, int sfc_lay, Float* sfc_src,
                              Float* lay_src, Float* lev_src, Float* sfc_source_Jac) {
  const size_t ncl = (size_t)ncol * nlay;
  FusedParams p;
  p.t = *t;
  p.ncol = ncol; p.nlay = nlay; p.play = play; p.plev = plev; p.tlay = tlay; p.vmr = vmr; p.col_dry_in = col_dry;
  p.op_kind = op_kind; p.tau = tau; p.ssa = ssa; p.g = g;
  p.cld_kind = cld_kind; p.cld_tau = cld_tau; p.cld_ssa = cld_ssa; p.cld_g = cld_g;
  Workspace w = prepare(p);
  const bool sw = t->krayl != nullptr;
  {
    KernelTimer timer(sw ? "gas_tau_fused[sw]" : "gas_tau_fused[lw]");
    const int grid = ceil_div((long long)ncl, kFThreads);
    CUtensorMap tm;
    if (gfast_enabled() && !tma_enabled()) {
      const TablesT tt = tables_gfast(*t);
      dim3 g2(grid, t->nbnd);
      if (sw) {
        if (tt.vec == 2) gas_tau_g_kernel<true, 2><<<g2, kGThreads, 0, stream()>>>(p, tt);
        else gas_tau_g_kernel<true, 1><<<g2, kGThreads, 0, stream()>>>(p, tt);
      } else {
        if (tt.vec == 2) gas_tau_g_kernel<false, 2><<<g2, kGThreads, 0, stream()>>>(p, tt);
        else gas_tau_g_kernel<false, 1><<<g2, kGThreads, 0, stream()>>>(p, tt);
      }
    } else if (make_table_tmap(&tm, t->kmajor, *t)) {
      // TMA-staged major-absorber table (box of kTB x 9 x kPB x kGB doubles per band, double-buffered)
      const size_t smem = (size_t)2 * kGB * kPB * 9 * kTB * sizeof(Float);
      if (sw) {
        RB_CUDA_CHECK(cudaFuncSetAttribute(gas_tau_tma_kernel<9, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gas_tau_tma_kernel<9, true><<<grid, kFThreads, smem, stream()>>>(p, tm);
      } else {
        RB_CUDA_CHECK(cudaFuncSetAttribute(gas_tau_tma_kernel<9, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gas_tau_tma_kernel<9, false><<<grid, kFThreads, smem, stream()>>>(p, tm);
      }
    } else if (std_dims(*t)) {
      if (sw) gas_tau_fused_kernel<14, 9, 60, true><<<grid, kFThreads, 0, stream()>>>(p);
      else gas_tau_fused_kernel<14, 9, 60, false><<<grid, kFThreads, 0, stream()>>>(p);
    } else {
      if (sw) gas_tau_fused_kernel<0, 0, 0, true><<<grid, kFThreads, 0, stream()>>>(p);
      else gas_tau_fused_kernel<0, 0, 0, false><<<grid, kFThreads, 0, stream()>>>(p);
    }
    RB_LAUNCH_CHECK();
  }
  if (lay_src) {
    PlanckFusedParams q;
    q.f = p; q.tlev = tlev; q.tsfc = tsfc; q.sfc_lay = sfc_lay;
    q.sfc_src = sfc_src; q.lay_src = lay_src; q.lev_src = lev_src; q.sfc_source_Jac = sfc_source_Jac;
    KernelTimer timer("planck_fused");
    dim3 grid(ceil_div(ncol, kFThreads), t->nbnd);
    if (gfast_enabled()) {
      const TablesT tt = tables_gfast(*t);
      const int lay_per_chunk = 9, nchunk = ceil_div(nlay, lay_per_chunk);
      dim3 g3(ceil_div(ncol, kGThreads), t->nbnd, nchunk);
      if (tt.vec == 2) planck_g_kernel<2><<<g3, kGThreads, 0, stream()>>>(q, tt, lay_per_chunk);
      else planck_g_kernel<1><<<g3, kGThreads, 0, stream()>>>(q, tt, lay_per_chunk);
    } else if (std_dims(*t)) planck_fused_kernel<14, 9, 60><<<grid, kFThreads, 0, stream()>>>(q);
    else planck_fused_kernel<0, 0, 0><<<grid, kFThreads, 0, stream()>>>(q);
    RB_LAUNCH_CHECK();
  }
  dev_free(w.block);
}

}  // extern "C"
