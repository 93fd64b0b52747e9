// solvers_abi.cu - RTE longwave / shortwave flux solvers for sm_100a.
//
// Replaces (extern mode) rte/kernels/api/mo_rte_solver_kernels.F90:50 (rte_lw_solver_noscat),
// :112 (rte_lw_solver_2stream), :145 (rte_sw_solver_noscat), :176 (rte_sw_solver_2stream).
// Numerics follow the default kernels rte/kernels/mo_rte_solver_kernels.F90 (cited inline).
//
// Design (B200-first, not the reference's accel layout which materialises full (ncol,nlay,ngpt)
// temporaries in global memory and launches one kernel per stage):
//   * A CTA owns a TILE of TC consecutive columns and walks g-points.  Per g-point:
//       phase A (all threads): the per-cell work - the exp/sqrt/divide-heavy part, independent across
//         layers - is done by ALL threads of the CTA, coalesced over columns, and its results
//         (transmittance, sources, R/T) are parked in SHARED memory as [layer][column] tiles;
//       phase B (one warp, lane <-> column): the layer-serial recurrences (transport / adding) run
//         out of shared memory - conflict-free, a few FMAs per layer - one thread per (col, gpt).
//     Every input plane is read from HBM exactly once; nothing but the requested fluxes is written.
//   * Broadband sums accumulate in shared memory in g-point order (the reference's order,
//     mo_rte_solver_kernels.F90:216-218,601-604) - deterministic, no atomics - and are written once.
//   * `intent(out)` arrays whose controlling flag is false are decoys that may alias:
//     they are never read or written here and no pointer is __restrict__.
#include <algorithm>
#include <atomic>
#include <climits>
#include <cstdlib>
#include "../kernels/elementwise.cuh"
#include "../kernels/solver_reg.cuh"
#include "../kernels/solver_ws.cuh"
#include "rte_kernels.h"
#include "rrtmgp_b200_ext.h"

using namespace rrtmgpb;

namespace {

constexpr int kSolverThreads = 256;
// process-wide switches, set once at start-up (atomics: safe to read from concurrently calling host threads)
// lev_source per g-point is the DEFAULT: an accelerator backend should reproduce the reference's accelerator kernels
// (accel/mo_rte_solver_kernels.F90:958-962), not the sequence-association slip of the serial CPU kernel
// (mo_rte_solver_kernels.F90:422), which stays reachable with rrtmgpb_set_lw_2stream_lev_source_per_gpt(0)
static std::atomic<int> g_lw2s_lev_per_gpt{1};
static std::atomic<int> g_solver_variant{0};  // 0: register kernels when nlay <= 144, else tiles; 1: always tiles  // 0: register/warp-systolic kernels when nlay <= 80, else tiles; 1: always tiles

struct Orient {
  int nlay;
  int top_at_1;
  __device__ __forceinline__ int lay(int k) const { return top_at_1 ? k : nlay - 1 - k; }  // k-th layer from the top
  __device__ __forceinline__ int lev(int k) const { return top_at_1 ? k : nlay - k; }      // k-th level from the top
};

__device__ __forceinline__ Float dev_pi() { return (Float)3.14159265358979323846; }  // = acos(-1._wp), :38

// =====================================================================================================
// LW no-scattering (optionally Tang-rescaled), mo_rte_solver_kernels.F90:51-240,248-367,620-844
// =====================================================================================================
struct LwNoscatParams {
  int ncol, nlay, ngpt, top_at_1, nmus;
  const Float *Ds, *weights, *tau, *lay_source, *lev_source, *sfc_emis, *sfc_src, *inc_flux;
  Float *flux_up, *flux_dn;
  int do_broadband;
  Float *bb_up, *bb_dn;
  int do_jac;
  const Float* sfc_srcJac;
  Float* flux_upJac;
  int do_rescaling;
  const Float *ssa, *g;
  int gpt_per_block;
};

template <int TC>
__global__ void __launch_bounds__(kSolverThreads) lw_noscat_kernel(const LwNoscatParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Float* sm = reinterpret_cast<Float*>(smem_raw);
  const int nlay = p.nlay, nlev = nlay + 1;
  const size_t ncol = p.ncol, ncl = ncol * nlay, nclp = ncol * nlev;
  const int lt = nlay * TC, vt = nlev * TC;
  Float* trans = sm;
  Float* sdn = trans + lt;
  Float* sup = sdn + lt;
  Float* nxt = sup + lt;
  Float *An = nullptr, *Cn = nullptr, *rdn = nullptr, *rup = nullptr;
  if (p.do_rescaling) { An = nxt; Cn = An + lt; rdn = Cn + lt; rup = rdn + vt; nxt = rup + vt; }
  Float *acc_up = nullptr, *acc_dn = nullptr, *acc_jac = nullptr;
  if (p.do_broadband) { acc_up = nxt; acc_dn = acc_up + vt; nxt = acc_dn + vt; }
  if (p.do_jac) { acc_jac = nxt; nxt = acc_jac + vt; }

  const int tid = threadIdx.x;
  const int col0 = blockIdx.x * TC;
  const int ncols = min(TC, p.ncol - col0);
  const int gb = blockIdx.y * p.gpt_per_block, ge = min(p.ngpt, gb + p.gpt_per_block);
  const Orient o{nlay, p.top_at_1};
  const Float pi = dev_pi();
  const Float tau_thresh = sqrt(sqrt((Float)RB_EPS));  // :636

  for (int i = tid; i < vt; i += kSolverThreads) {
    if (p.do_broadband) { acc_up[i] = 0; acc_dn[i] = 0; }
    if (p.do_jac) acc_jac[i] = 0;
  }
  __syncthreads();

  for (int g = gb; g < ge; ++g) {
    for (int imu = 0; imu < p.nmus; ++imu) {
      const Float w = p.weights[imu];
      // ---------------- phase A: cell quantities, all threads ----------------
      for (int cell = tid; cell < lt; cell += kSolverThreads) {
        const int c = cell % TC, l = cell / TC;
        if (c < ncols) {
          const size_t col = col0 + c;
          const size_t i3 = col + ncol * l + ncl * g;
          const Float D = p.Ds[col + ncol * ((size_t)g + (size_t)p.ngpt * imu)];
          Float tau_loc;
          if (p.do_rescaling) {  // :154-178
            const Float ssal = p.ssa[i3];
            const Float wb = ssal * ((Float)1 - p.g[i3]) * (Float)0.5;
            const Float scaleTau = ((Float)1 - ssal + wb);
            Cn[cell] = (Float)0.4 * wb / scaleTau;
            tau_loc = p.tau[i3] * D * scaleTau;
          } else {
            tau_loc = p.tau[i3] * D;  // :181
          }
          const Float tr = exp(-tau_loc);
          if (p.do_rescaling) An[cell] = ((Float)1 - tr * tr);
          // lw_source_noscat :652-663
          Float fact;
          if (tau_loc > tau_thresh) fact = ((Float)1 - tr) / tau_loc - tr;
          else fact = tau_loc * ((Float)0.5 + tau_loc * (-(Float)1 / (Float)3 + tau_loc * (Float)1 / (Float)8));
          const Float lay = p.lay_source[i3];
          const size_t iv = col + ncol * l + nclp * g;
          const Float lev_lo = p.lev_source[iv], lev_hi = p.lev_source[iv + ncol];
          const Float s_inc = ((Float)1 - tr) * lev_hi + (Float)2 * fact * (lay - lev_hi);
          const Float s_dec = ((Float)1 - tr) * lev_lo + (Float)2 * fact * (lay - lev_lo);
          trans[cell] = tr;
          sdn[cell] = p.top_at_1 ? s_inc : s_dec;  // :638-644
          sup[cell] = p.top_at_1 ? s_dec : s_inc;
        }
      }
      __syncthreads();
      // ---------------- phase B: layer-serial transport, one thread per (col, gpt) ----------------
      if (tid < ncols) {
        const int c = tid;
        const size_t col = col0 + c;
        const size_t gi = col + ncol * g;
        const Float piw = pi * w;
        Float* fup = p.flux_up + nclp * g;
        Float* fdn = p.flux_dn + nclp * g;
        // record one level value (intensity I) of the up/down stream
        auto rec = [&](Float* acc, Float* gflux, int lev, Float I) {
          if (p.do_broadband) acc[lev * TC + c] += w * I;                       // :216-218 (scaled by pi at the end)
          else if (imu == 0) gflux[col + ncol * lev] = piw * I;                 // :223-224
          else gflux[col + ncol * lev] += piw * I;                              // :356-357
        };
        Float I = p.inc_flux[gi] / (pi * w);                                    // :144
        if (!p.do_rescaling) {
          rec(acc_dn, fdn, o.lev(0), I);
          for (int k = 0; k < nlay; ++k) {                                      // :697-706
            const int i = o.lay(k) * TC + c;
            I = trans[i] * I + sdn[i];
            rec(acc_dn, fdn, o.lev(k + 1), I);
          }
          const Float emis = p.sfc_emis[gi];
          Float Iu = I * ((Float)1 - emis) + emis * p.sfc_src[gi];             // :198-200
          Float Ij = p.do_jac ? emis * p.sfc_srcJac[gi] : (Float)0;             // :202
          rec(acc_up, fup, o.lev(nlay), Iu);
          if (p.do_jac) acc_jac[o.lev(nlay) * TC + c] += w * Ij;
          for (int k = nlay - 1; k >= 0; --k) {                                 // :729-743
            const int i = o.lay(k) * TC + c;
            Iu = trans[i] * Iu + sup[i];
            rec(acc_up, fup, o.lev(k), Iu);
            if (p.do_jac) { Ij = trans[i] * Ij; acc_jac[o.lev(k) * TC + c] += w * Ij; }
          }
        } else {
          // first (plain) downward sweep, keeping every level (:194)
          rdn[o.lev(0) * TC + c] = I;
          for (int k = 0; k < nlay; ++k) {
            const int i = o.lay(k) * TC + c;
            I = trans[i] * I + sdn[i];
            rdn[o.lev(k + 1) * TC + c] = I;
          }
          const Float emis = p.sfc_emis[gi];
          Float Iu = I * ((Float)1 - emis) + emis * p.sfc_src[gi];
          Float Ij = p.do_jac ? emis * p.sfc_srcJac[gi] : (Float)0;
          rup[o.lev(nlay) * TC + c] = Iu;
          if (p.do_jac) acc_jac[o.lev(nlay) * TC + c] += w * Ij;
          // upward sweep with adjustment (:784-793 / :816-826): uses radn_dn at the layer's TOP level
          for (int k = nlay - 1; k >= 0; --k) {
            const int i = o.lay(k) * TC + c;
            const Float adj = Cn[i] * (An[i] * rdn[o.lev(k) * TC + c] - trans[i] * sdn[i] - sup[i]);
            Iu = trans[i] * Iu + sup[i] + adj;
            rup[o.lev(k) * TC + c] = Iu;
            if (p.do_jac) { Ij = trans[i] * Ij; acc_jac[o.lev(k) * TC + c] += w * Ij; }
          }
          // second downward sweep with adjustment (:798-808 / :832-842): radn_up at the layer's TOP
          // level when top_at_1 but at its BOTTOM level otherwise (asymmetric in the reference; kept)
          Float Id = rdn[o.lev(0) * TC + c];
          for (int k = 0; k < nlay; ++k) {
            const int i = o.lay(k) * TC + c;
            const Float Iup_ref = rup[o.lev(p.top_at_1 ? k : k + 1) * TC + c];
            const Float adj = Cn[i] * (An[i] * Iup_ref - trans[i] * sup[i] - sdn[i]);
            Id = trans[i] * Id + sdn[i] + adj;
            rdn[o.lev(k + 1) * TC + c] = Id;
          }
          for (int lev = 0; lev < nlev; ++lev) {
            rec(acc_dn, fdn, lev, rdn[lev * TC + c]);
            rec(acc_up, fup, lev, rup[lev * TC + c]);
          }
        }
      }
      __syncthreads();
    }
  }
  // ---------------- epilogue: spectrally integrated outputs, written once ----------------
  for (int i = tid; i < vt; i += kSolverThreads) {
    const int c = i % TC, lev = i / TC;
    if (c < ncols) {
      const size_t o2 = (size_t)(col0 + c) + ncol * lev;
      if (p.do_broadband) { p.bb_up[o2] = pi * acc_up[i]; p.bb_dn[o2] = pi * acc_dn[i]; }   // :233-236
      if (p.do_jac) p.flux_upJac[o2] = pi * acc_jac[i];                                       // :237-238
    }
  }
}

// =====================================================================================================
// adding (Shonk & Hogan 2008), mo_rte_solver_kernels.F90:1135-1245 - shared by LW and SW two-stream.
// In-place in the shared-memory tiles: Sup is overwritten with denom (beta), AlbB/SrcB receive the
// albedo and source at the level BELOW each layer.  rec(k, fup, fdn) gets the fluxes at the k-th level
// from the top.
// =====================================================================================================
template <int TC, typename Rec>
__device__ __forceinline__ void adding_sweeps(int c, const Orient o, const Float* Rdif, const Float* Tdif, Float* Sup,
                                              const Float* Sdn, Float* AlbB, Float* SrcB, Float albedo_sfc,
                                              Float src_sfc, Float flux_dn_top, Rec rec) {
  const int nlay = o.nlay;
  Float alb = albedo_sfc, src = src_sfc;  // :1166-1168 / :1206-1208
  for (int k = nlay - 1; k >= 0; --k) {    // :1174-1186 / :1214-1226
    const int i = o.lay(k) * TC + c;
    AlbB[i] = alb;
    SrcB[i] = src;
    const Float r = Rdif[i], t = Tdif[i];
    const Float denom = (Float)1 / ((Float)1 - r * alb);
    const Float albn = r + t * t * alb * denom;
    src = Sup[i] + t * denom * (src + alb * Sdn[i]);
    alb = albn;
    Sup[i] = denom;
  }
  Float fdn = flux_dn_top;
  Float fup = fdn * alb + src;             // :1190 / :1230
  rec(0, fup, fdn);
  for (int k = 0; k < nlay; ++k) {         // :1196-1202 / :1236-1243
    const int i = o.lay(k) * TC + c;
    fdn = (Tdif[i] * fdn + Rdif[i] * SrcB[i] + Sdn[i]) * Sup[i];
    fup = fdn * AlbB[i] + SrcB[i];
    rec(k + 1, fup, fdn);
  }
}

// =====================================================================================================
// SW two-stream, mo_rte_solver_kernels.F90:503-609, 985-1127
// =====================================================================================================
struct SwParams {
  int ncol, nlay, ngpt, top_at_1;
  const Float *tau, *ssa, *g, *mu0, *sfc_alb_dir, *sfc_alb_dif, *inc_flux_dir;
  Float *flux_up, *flux_dn, *flux_dir;
  int has_dif_bc;
  const Float* inc_flux_dif;
  int do_broadband;
  Float *bb_up, *bb_dn, *bb_dir;
  int gpt_per_block;
};

template <int TC>
__global__ void __launch_bounds__(kSolverThreads) sw_2stream_kernel(const SwParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Float* sm = reinterpret_cast<Float*>(smem_raw);
  const int nlay = p.nlay, nlev = nlay + 1;
  const size_t ncol = p.ncol, ncl = ncol * nlay, nclp = ncol * nlev;
  const int lt = nlay * TC, vt = nlev * TC;
  Float* Rdif = sm;
  Float* Tdif = Rdif + lt;
  Float* A3 = Tdif + lt;   // Rdir -> source_up -> denom
  Float* A4 = A3 + lt;     // Tdir -> source_dn
  Float* A5 = A4 + lt;     // Tnoscat -> albedo below the layer
  Float* A6 = A5 + lt;     // source below the layer
  Float* nxt = A6 + lt;
  Float *acc_up = nullptr, *acc_dn = nullptr, *acc_dir = nullptr;
  if (p.do_broadband) { acc_up = nxt; acc_dn = acc_up + vt; acc_dir = acc_dn + vt; }

  const int tid = threadIdx.x;
  const int col0 = blockIdx.x * TC;
  const int ncols = min(TC, p.ncol - col0);
  const int gb = blockIdx.y * p.gpt_per_block, ge = min(p.ngpt, gb + p.gpt_per_block);
  const Orient o{nlay, p.top_at_1};
  const Float eps = (Float)RB_EPS;
  const Float min_k = (Float)1.e4 * eps;   // :1005
  const Float min_mu0 = sqrt(eps);         // :1006

  if (p.do_broadband) {
    for (int i = tid; i < vt; i += kSolverThreads) { acc_up[i] = 0; acc_dn[i] = 0; acc_dir[i] = 0; }
  }
  __syncthreads();

  for (int g = gb; g < ge; ++g) {
    // ---------------- phase A: two-stream layer properties, all threads (:1027-1108) ----------------
    for (int cell = tid; cell < lt; cell += kSolverThreads) {
      const int c = cell % TC, l = cell / TC;
      if (c < ncols) {
        const size_t col = col0 + c;
        const size_t i3 = col + ncol * l + ncl * g;
        const Float tau_s = p.tau[i3], w0_s = p.ssa[i3], g_s = p.g[i3];
        const Float mu0 = p.mu0[col + ncol * l];
        const Float gamma1 = ((Float)8 - w0_s * ((Float)5 + (Float)3 * g_s)) * (Float).25;
        const Float gamma2 = (Float)3 * (w0_s * ((Float)1 - g_s)) * (Float).25;
        const Float k = sqrt(fmax((gamma1 - gamma2) * (gamma1 + gamma2), min_k));
        const Float exp_minusktau = exp(-tau_s * k);
        const Float exp_minus2ktau = exp_minusktau * exp_minusktau;
        Float RT_term = (Float)1 / (k * ((Float)1 + exp_minus2ktau) + gamma1 * ((Float)1 - exp_minus2ktau));
        Rdif[cell] = RT_term * gamma2 * ((Float)1 - exp_minus2ktau);
        Tdif[cell] = RT_term * (Float)2 * k * exp_minusktau;
        const Float mu0_s = fmax(min_mu0, mu0);
        const Float k_mu = k * mu0_s;
        const Float om = (Float)1 - k_mu * k_mu;
        RT_term = w0_s * RT_term / (fabs(om) >= eps ? om : eps);
        const Float gamma3 = ((Float)2 - (Float)3 * mu0_s * g_s) * (Float).25;
        const Float gamma4 = (Float)1 - gamma3;
        const Float alpha1 = gamma1 * gamma4 + gamma2 * gamma3;
        const Float alpha2 = gamma1 * gamma3 + gamma2 * gamma4;
        const Float k_gamma3 = k * gamma3;
        const Float k_gamma4 = k * gamma4;
        const Float Tnoscat = exp(-tau_s / mu0_s);
        Float Rdir = RT_term * (((Float)1 - k_mu) * (alpha2 + k_gamma3) -
                                ((Float)1 + k_mu) * (alpha2 - k_gamma3) * exp_minus2ktau -
                                (Float)2.0 * (k_gamma3 - alpha2 * k_mu) * exp_minusktau * Tnoscat);
        Float Tdir = -RT_term * (((Float)1 + k_mu) * (alpha1 + k_gamma4) * Tnoscat -
                                 ((Float)1 - k_mu) * (alpha1 - k_gamma4) * exp_minus2ktau * Tnoscat -
                                 (Float)2.0 * (k_gamma4 + alpha1 * k_mu) * exp_minusktau);
        Rdir = fmax((Float)0, fmin(Rdir, ((Float)1 - Tnoscat)));          // :1107
        Tdir = fmax((Float)0, fmin(Tdir, ((Float)1 - Tnoscat - Rdir)));   // :1108
        // night layers: the reference zeroes source_up/dn where mu0 <= 0 (:1122-1125); zeroing the
        // direct R/T here is equivalent (sources are Rdir*flux, Tdir*flux)
        const bool night = !(mu0 > (Float)0);
        A3[cell] = night ? (Float)0 : Rdir;
        A4[cell] = night ? (Float)0 : Tdir;
        A5[cell] = Tnoscat;
      }
    }
    __syncthreads();
    // ---------------- phase B: direct beam + adding, one thread per (col, gpt) ----------------
    if (tid < ncols) {
      const int c = tid;
      const size_t col = col0 + c;
      const size_t gi = col + ncol * g;
      Float* gup = p.flux_up + nclp * g;
      Float* gdn = p.flux_dn + nclp * g;
      Float* gdir = p.flux_dir + nclp * g;
      Float dir = p.inc_flux_dir[gi] * p.mu0[col + ncol * o.lay(0)];        // :575
      if (p.do_broadband) { acc_dir[o.lev(0) * TC + c] += dir; } else { gdir[col + ncol * o.lev(0)] = dir; }
      const Float dir_top = dir;
      for (int k = 0; k < nlay; ++k) {                                       // :1110-1112
        const int i = o.lay(k) * TC + c;
        const Float s_up = A3[i] * dir;
        const Float s_dn = A4[i] * dir;
        dir = A5[i] * dir;
        A3[i] = s_up;
        A4[i] = s_dn;
        if (p.do_broadband) acc_dir[o.lev(k + 1) * TC + c] += dir; else gdir[col + ncol * o.lev(k + 1)] = dir;
        A5[i] = dir;  // keep the direct flux at the level below layer k for the totals in sweep 3
      }
      const Float src_sfc = (p.mu0[col + ncol * o.lay(nlay - 1)] > (Float)0) ? dir * p.sfc_alb_dir[gi] : (Float)0;  // :1120
      const Float dn_top = p.has_dif_bc ? p.inc_flux_dif[gi] : (Float)0;     // :579-583
      // A5 currently holds direct fluxes; adding() needs a slot for the albedo below each layer.
      // Move the direct flux of each level into the accumulator / global now, then reuse A5.
      // (broadband_dn = diffuse + direct, :603; g-point flux_dn = diffuse + direct, :606)
      auto rec = [&](int k, Float fup, Float fdn) {
        const int lev = o.lev(k);
        if (p.do_broadband) {
          acc_up[lev * TC + c] += fup;                                       // :602
          acc_dn[lev * TC + c] += fdn;                                       // diffuse part of :603
        } else {
          gup[col + ncol * lev] = fup;
          gdn[col + ncol * lev] = fdn + gdir[col + ncol * lev];              // :606 (own earlier write)
        }
      };
      if (p.do_broadband) {
        // direct part of :603, added per g-point before the diffuse part so that each g-point's
        // (diffuse + direct) enters the running sum together
        acc_dn[o.lev(0) * TC + c] += dir_top;
        for (int k = 0; k < nlay; ++k) acc_dn[o.lev(k + 1) * TC + c] += A5[o.lay(k) * TC + c];
      }
      adding_sweeps<TC>(c, o, Rdif, Tdif, A3, A4, A5, A6, p.sfc_alb_dif[gi], src_sfc, dn_top, rec);
    }
    __syncthreads();
  }
  if (p.do_broadband) {
    for (int i = tid; i < vt; i += kSolverThreads) {
      const int c = i % TC, lev = i / TC;
      if (c < ncols) {
        const size_t o2 = (size_t)(col0 + c) + ncol * lev;
        p.bb_up[o2] = acc_up[i];
        p.bb_dn[o2] = acc_dn[i];
        p.bb_dir[o2] = acc_dir[i];
      }
    }
  }
}

// =====================================================================================================
// LW two-stream, mo_rte_solver_kernels.F90:377-440, 854-967
// =====================================================================================================
struct Lw2sParams {
  int ncol, nlay, ngpt, top_at_1, lev_per_gpt;
  const Float *tau, *ssa, *g, *lay_source, *lev_source, *sfc_emis, *sfc_src, *inc_flux;
  Float *flux_up, *flux_dn;
  int gpt_per_block;
};

template <int TC>
__global__ void __launch_bounds__(kSolverThreads) lw_2stream_kernel(const Lw2sParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Float* sm = reinterpret_cast<Float*>(smem_raw);
  const int nlay = p.nlay, nlev = nlay + 1;
  const size_t ncol = p.ncol, ncl = ncol * nlay, nclp = ncol * nlev;
  const int lt = nlay * TC;
  Float* Rdif = sm;
  Float* Tdif = Rdif + lt;
  Float* Sup = Tdif + lt;
  Float* Sdn = Sup + lt;
  Float* AlbB = Sdn + lt;
  Float* SrcB = AlbB + lt;
  const int tid = threadIdx.x;
  const int col0 = blockIdx.x * TC;
  const int ncols = min(TC, p.ncol - col0);
  const int gb = blockIdx.y * p.gpt_per_block, ge = min(p.ngpt, gb + p.gpt_per_block);
  const Orient o{nlay, p.top_at_1};
  const Float pi = dev_pi();
  const Float LW_diff_sec = (Float)1.66f;  // :870 single-precision literal widened to wp

  for (int g = gb; g < ge; ++g) {
    // reference default kernels use g-point 1's level source for every g-point (:422); see
    // rrtmgpb_set_lw_2stream_lev_source_per_gpt()
    const size_t gsrc = p.lev_per_gpt ? g : 0;
    for (int cell = tid; cell < lt; cell += kSolverThreads) {
      const int c = cell % TC, l = cell / TC;
      if (c < ncols) {
        const size_t col = col0 + c;
        const size_t i3 = col + ncol * l + ncl * g;
        const Float tau = p.tau[i3], w0 = p.ssa[i3], gg = p.g[i3];
        // lw_two_stream :879-905
        const Float gamma1 = LW_diff_sec * ((Float)1 - (Float)0.5 * w0 * ((Float)1 + gg));
        const Float gamma2 = LW_diff_sec * (Float)0.5 * w0 * ((Float)1 - gg);
        const Float k = sqrt(fmax((gamma1 - gamma2) * (gamma1 + gamma2), (Float)1.e-12));
        const Float exp_minusktau = exp(-tau * k);
        const Float exp_minus2ktau = exp_minusktau * exp_minusktau;
        const Float RT_term = (Float)1 / (k * ((Float)1 + exp_minus2ktau) + gamma1 * ((Float)1 - exp_minus2ktau));
        const Float rdif = RT_term * gamma2 * ((Float)1 - exp_minus2ktau);
        const Float tdif = RT_term * (Float)2 * k * exp_minusktau;
        // lw_source_2str :938-962
        const size_t iv = col + ncol * l + nclp * gsrc;
        const Float lev_a = p.lev_source[iv], lev_b = p.lev_source[iv + ncol];
        const Float lev_top = p.top_at_1 ? lev_a : lev_b;
        const Float lev_bot = p.top_at_1 ? lev_b : lev_a;
        Float s_up = 0, s_dn = 0;
        if (tau > (Float)1.0e-8) {
          const Float Z = (lev_bot - lev_top) / (tau * (gamma1 + gamma2));
          const Float Zup_top = Z + lev_top;
          const Float Zup_bottom = Z + lev_bot;
          const Float Zdn_top = -Z + lev_top;
          const Float Zdn_bottom = -Z + lev_bot;
          s_up = pi * (Zup_top - rdif * Zdn_top - tdif * Zup_bottom);
          s_dn = pi * (Zdn_bottom - rdif * Zup_bottom - tdif * Zdn_top);
        }
        Rdif[cell] = rdif; Tdif[cell] = tdif; Sup[cell] = s_up; Sdn[cell] = s_dn;
      }
    }
    __syncthreads();
    if (tid < ncols) {
      const int c = tid;
      const size_t col = col0 + c;
      const size_t gi = col + ncol * g;
      Float* gup = p.flux_up + nclp * g;
      Float* gdn = p.flux_dn + nclp * g;
      const Float emis = p.sfc_emis[gi];
      const Float src_sfc = pi * emis * p.sfc_src[gi];  // :965
      auto rec = [&](int k, Float fup, Float fdn) {
        const int lev = o.lev(k);
        gup[col + ncol * lev] = fup;
        gdn[col + ncol * lev] = fdn;
      };
      adding_sweeps<TC>(c, o, Rdif, Tdif, Sup, Sdn, AlbB, SrcB, (Float)1 - emis, src_sfc, p.inc_flux[gi], rec);
    }
    __syncthreads();
  }
}

// =====================================================================================================
// SW direct beam only, mo_rte_solver_kernels.F90:450-494.  One thread per (col, gpt); no reuse, no tile.
// =====================================================================================================
__global__ void __launch_bounds__(kEltThreads) sw_noscat_kernel(int ncol, int nlay, int ngpt, int top_at_1,
                                                                const Float* tau, const Float* mu0,
                                                                const Float* inc_flux_dir, Float* flux_dir) {
  const size_t n = (size_t)ncol * ngpt;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const size_t col = t % ncol, g = t / ncol;
  const size_t nc = ncol, ncl = nc * nlay, nclp = nc * (nlay + 1);
  const Orient o{nlay, top_at_1};
  Float f = inc_flux_dir[t] * mu0[col + nc * o.lay(0)];
  flux_dir[col + nc * o.lev(0) + nclp * g] = f;
  for (int k = 0; k < nlay; ++k) {
    const int l = o.lay(k);
    f = f * exp(-tau[col + nc * l + ncl * g] / mu0[col + nc * l]);
    flux_dir[col + nc * o.lev(k + 1) + nclp * g] = f;
  }
}

// ---- launch helpers -------------------------------------------------------------------------------
int solver_tc_override() {
  static int v = -1;
  if (v < 0) { const char* e = std::getenv("RRTMGPB_SOLVER_TC"); v = e ? std::atoi(e) : 0; }
  return v;
}

// pick the column-tile width: widest of {16, 8, 4} (or the override) whose tiles fit in shared memory
int pick_tc(size_t floats_per_col) {
  const size_t limit = 227 * 1024;
  const int ov = solver_tc_override();
  const int cands[4] = {ov > 0 ? ov : 16, 16, 8, 4};
  for (int tc : cands) {
    if (tc != 32 && tc != 16 && tc != 8 && tc != 4) continue;
    if (floats_per_col * tc * sizeof(Float) <= limit) return tc;
  }
  std::fprintf(stderr, "rte_rrtmgp_b200: nlay too large for the shared-memory solver tiles\n");
  std::abort();
}

int gpt_groups(int ncol, int tc, int ngpt) {
  // enough CTAs for >= 4 waves of 148 SMs when the column count alone does not provide them
  const int tiles = ceil_div(ncol, tc);
  int groups = ceil_div(148 * 4, tiles);
  if (groups < 1) groups = 1;
  if (groups > ngpt) groups = ngpt;
  return groups;
}

template <typename K, typename P>
void launch_tile(K kern, const P& p, dim3 grid, size_t smem, const char* name) {
  KernelTimer timer(name);
  RB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kSolverThreads, smem, stream()>>>(p);
  RB_LAUNCH_CHECK();
}

#define DISPATCH_TC(tc, KERNEL, params, grid, smem)                      \
  switch (tc) {                                                          \
    case 32: launch_tile(KERNEL<32>, params, grid, smem, #KERNEL); break;         \
    case 16: launch_tile(KERNEL<16>, params, grid, smem, #KERNEL); break;         \
    case 8: launch_tile(KERNEL<8>, params, grid, smem, #KERNEL); break;           \
    default: launch_tile(KERNEL<4>, params, grid, smem, #KERNEL); break;          \
  }


// register-resident kernels: chunk length CL = ceil(nlay/8) in {8,9,10}
inline int reg_chunk_len(int nlay, int ncol) {
  if (g_solver_variant.load(std::memory_order_relaxed) == 1) return 0;
  // the register kernels use 32-bit in-plane offsets (kernels/solver_reg.cuh): planes of 2^31 elements or more go to
  // the tile kernels, whose indices are 64-bit
  if ((long long)ncol * (nlay + 1) > (long long)INT_MAX) return 0;
  if (nlay <= 64) return 8;
  if (nlay <= 72) return 9;
  if (nlay <= 80) return 10;
  // 16 lanes per column (8 warps per CTA): 6, 7 or 9 layers per lane
  if (nlay <= 96) return 6;
  if (nlay <= 112) return 7;
  if (nlay <= 144) return 9;
  return 0;
}
inline int reg_lanes(int nlay) { return nlay <= 80 ? 8 : 16; }
inline bool force_pad() {  // RRTMGPB_FORCE_PAD=1: run the FULL = 2 (padded-tile) instantiations even when nlay == lanes*CL (A/B switch)
  static const bool v = [] { const char* e = std::getenv("RRTMGPB_FORCE_PAD"); return e && e[0] == '1'; }();
  return v;
}
inline bool sw_cl8() {  // RRTMGPB_SW_CL9=1: 9 layers per lane for nlay <= 64 in the SW kernel too (A/B switch; default 8, see below)
  static const bool v = [] { const char* e = std::getenv("RRTMGPB_SW_CL9"); return !(e && e[0] == '1'); }();
  return v;
}
inline int reg_minb() {  // experiment switch: resident CTAs per SM the register kernels are compiled for
  static const int v = [] { const char* e = std::getenv("RRTMGPB_REG_MINB"); return (e && e[0] == '3') ? 3 : 2; }();
  return v;
}
inline bool reg_sacc() {  // experiment switch: SW broadband accumulators in shared memory (frees 54 registers)
  static const bool v = [] { const char* e = std::getenv("RRTMGPB_REG_SACC"); return e && e[0] == '1'; }();
  return v;
}
inline bool solver_tma_enabled() {  // RRTMGPB_SOLVER_TMA=0: lane-private cp.async staging instead (A/B switch)
  static const bool v = [] { const char* e = std::getenv("RRTMGPB_SOLVER_TMA"); return !(e && e[0] == '0'); }();
  return v;
}
// warp-specialised SW kernel (kernels/solver_ws.cuh): RRTMGPB_SW_WS=0/1 (A/B switch; measured slower than the register
// kernel, see the header - off by default)
inline bool sw_ws_enabled() {
  static const bool v = [] { const char* e = std::getenv("RRTMGPB_SW_WS"); return e ? e[0] == '1' : false; }();
  const int variant = g_solver_variant.load(std::memory_order_relaxed);
  return variant == 3 || (variant == 0 && v);   // rrtmgpb_set_solver_variant: 2 = register kernels only, 3 = warp-specialised where applicable
}
template <int CL, bool BB, bool MERGED, int NPW>
void launch_sw_ws(const SwRegParams& q, const SwTmaMaps& maps, dim3 grid) {
  auto kern = sw_2stream_ws_kernel<CL, BB, MERGED, NPW>;
  constexpr size_t smem = sw_ws_smem(8 * CL);
  RB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kWsConsumerThreads + 32 * NPW, smem, stream()>>>(q, maps);
}
template <int CL>
void dispatch_sw_ws(const SwRegParams& q, const SwTmaMaps& maps, dim3 grid, bool bb, bool merged) {
#define RB_WS(BBV, MV)                                                   \
  {                                                                      \
    launch_sw_ws<CL, BBV, MV, 8>(q, maps, grid);                         \
  }
  if (bb) { if (merged) RB_WS(true, true) else RB_WS(true, false) }
  else { if (merged) RB_WS(false, true) else RB_WS(false, false) }
#undef RB_WS
}
inline int reg_gpt_groups(int ncol, int ngpt) {
  const int ctas = ceil_div(ncol, (kRegThreads / 32) * kRegCols);
  int groups = ceil_div(148 * 8, ctas);
  if (groups < 1) groups = 1;
  if (groups > ngpt) groups = ngpt;
  return groups;
}
#define DISPATCH_CL(cl, KERNEL, params, grid)                                                    \
  {                                                                                              \
    KernelTimer timer(#KERNEL);                                                                  \
    switch (cl) {                                                                                \
      case 8: KERNEL<8><<<grid, kRegThreads, 0, stream()>>>(params); break;                      \
      case 9: KERNEL<9><<<grid, kRegThreads, 0, stream()>>>(params); break;                      \
      default: KERNEL<10><<<grid, kRegThreads, 0, stream()>>>(params); break;                    \
    }                                                                                            \
    RB_LAUNCH_CHECK();                                                                           \
  }

}  // namespace

// test hook for kernels/fastmath.cuh: out = (exp(x), sqrt(|x|), 1/x, 1/(x*x+1) / x) per element
__global__ void fastmath_probe_kernel(int n, const double* x, double* e, double* s, double* r, double* d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  e[i] = rb_exp<true>(x[i]);
  s[i] = rb_sqrt(fabs(x[i]));
  r[i] = rb_rcp(x[i]);
  d[i] = rb_div(__dadd_rn(__dmul_rn(x[i], x[i]), 1.0), x[i]);  // (no FMA contraction: same numerator as the checker)
}

extern "C" {

void rrtmgpb_fastmath_probe(int n, const double* x, double* e, double* s, double* r, double* d) {
  DevArg<double> ax(x, n, Dir::In);
  DevArg<double> ae(e, n, Dir::Out), as(s, n, Dir::Out), ar(r, n, Dir::Out), ad(d, n, Dir::Out);
  fastmath_probe_kernel<<<ceil_div(n, 256), 256, 0, stream()>>>(n, ax, ae, as, ar, ad);
  RB_LAUNCH_CHECK();
}

void rrtmgpb_set_solver_variant(int v) { g_solver_variant.store(v); }
int rrtmgpb_get_solver_variant(void) { return g_solver_variant.load(); }

void rrtmgpb_set_lw_2stream_lev_source_per_gpt(int on) { g_lw2s_lev_per_gpt.store(on ? 1 : 0); }

void rte_lw_solver_noscat(const int* ncol_, const int* nlay_, const int* ngpt_, const Bool* top_at_1,
                          const int* nmus_, const Float* Ds, const Float* weights, const Float* tau,
                          const Float* lay_source, const Float* lev_source, const Float* sfc_emis,
                          const Float* sfc_src, const Float* inc_flux, Float* flux_up, Float* flux_dn,
                          const Bool* do_broadband, Float* broadband_up, Float* broadband_dn,
                          const Bool* do_Jacobians, const Float* sfc_srcJac, Float* flux_upJac,
                          const Bool* do_rescaling, const Float* ssa, const Float* g) {
  const int ncol = *ncol_, nlay = *nlay_, ngpt = *ngpt_, nmus = *nmus_;
  const bool bb = *do_broadband, jac = *do_Jacobians, resc = *do_rescaling;
  const size_t nc = ncol, ncl = nc * nlay, nclp = nc * (nlay + 1), ncg = nc * ngpt;
  DevArg<Float> a_D(Ds, ncg * nmus, Dir::In), a_w(weights, nmus, Dir::In), a_tau(tau, ncl * ngpt, Dir::In),
      a_lay(lay_source, ncl * ngpt, Dir::In), a_lev(lev_source, nclp * ngpt, Dir::In), a_em(sfc_emis, ncg, Dir::In),
      a_ss(sfc_src, ncg, Dir::In), a_inc(inc_flux, ncg, Dir::In);
  DevArg<Float> a_fu(flux_up, nclp * ngpt, Dir::Out, !bb), a_fd(flux_dn, nclp * ngpt, Dir::Out, !bb);
  DevArg<Float> a_bu(broadband_up, nclp, Dir::Out, bb), a_bd(broadband_dn, nclp, Dir::Out, bb);
  DevArg<Float> a_sj(sfc_srcJac, ncg, Dir::In, jac), a_fj(flux_upJac, nclp, Dir::Out, jac);
  DevArg<Float> a_ssa(ssa, ncl * ngpt, Dir::In, resc), a_g(g, ncl * ngpt, Dir::In, resc);
  LwNoscatParams p;
  p.ncol = ncol; p.nlay = nlay; p.ngpt = ngpt; p.top_at_1 = *top_at_1 ? 1 : 0; p.nmus = nmus;
  p.Ds = a_D; p.weights = a_w; p.tau = a_tau; p.lay_source = a_lay; p.lev_source = a_lev; p.sfc_emis = a_em;
  p.sfc_src = a_ss; p.inc_flux = a_inc; p.flux_up = a_fu; p.flux_dn = a_fd; p.do_broadband = bb; p.bb_up = a_bu;
  p.bb_dn = a_bd; p.do_jac = jac; p.sfc_srcJac = a_sj; p.flux_upJac = a_fj; p.do_rescaling = resc; p.ssa = a_ssa;
  p.g = a_g;
  // ---- Tang rescaling on the register kernels (TMA tiles of the five planes; else the tile kernel below)
  if (const int cl = resc ? reg_chunk_len(nlay, ncol) : 0) {
    LwNoscatRegParams q;
    q.ncol = ncol; q.nlay = nlay; q.ngpt = ngpt; q.top_at_1 = p.top_at_1; q.nmus = nmus; q.Ds = p.Ds;
    q.weights = p.weights; q.tau = p.tau; q.lay_source = p.lay_source; q.lev_source = p.lev_source;
    q.sfc_emis = p.sfc_emis; q.sfc_src = p.sfc_src; q.inc_flux = p.inc_flux; q.flux_up = p.flux_up;
    q.flux_dn = p.flux_dn; q.do_broadband = bb; q.bb_up = p.bb_up; q.bb_dn = p.bb_dn; q.do_jac = jac;
    q.sfc_srcJac = p.sfc_srcJac; q.flux_upJac = p.flux_upJac; q.accumulate = 0; q.group_stride = 0;
    const int groups = (bb || jac) ? 1 : reg_gpt_groups(ncol, ngpt);
    q.gpt_per_block = ceil_div(ngpt, groups);
    dim3 grid(ceil_div(ncol, kRegColsPerCta), ceil_div(ngpt, q.gpt_per_block));
    const int nch = reg_lanes(nlay);
    const int clv = nch == 16 ? cl : (cl <= 9 ? 9 : 10), rows = nch * clv;
    q.tile_rows = rows; q.row0 = p.top_at_1 ? 0 : nlay - rows;
    LwResclTmaMaps maps;
    const bool use_tma = solver_tma_enabled() && make_plane_tmap(&maps.tau, q.tau, ncol, nlay, ngpt, rows) &&
                         make_plane_tmap(&maps.ssa, p.ssa, ncol, nlay, ngpt, rows) && make_plane_tmap(&maps.g, p.g, ncol, nlay, ngpt, rows) &&
                         make_plane_tmap(&maps.lay, q.lay_source, ncol, nlay, ngpt, rows) &&
                         make_plane_tmap(&maps.lev, q.lev_source, ncol, nlay + 1, ngpt, rows + 1);
    if (use_tma) {
      KernelTimer timer("lw_rescl_reg_kernel");
      const size_t smem = lw_rescl_reg_tma_smem(rows, reg_threads(nch));
#define LWRS2(CLV, BBV, JACV, NCHV)                                                                       \
  {                                                                                                       \
    auto kern = lw_rescl_reg_kernel<CLV, BBV, JACV, NCHV>;                                                \
    RB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    kern<<<grid, reg_threads(NCHV), smem, stream()>>>(q, maps);                                           \
  }
#define LWRS(CLV, NCHV)                          \
  if (bb && jac) LWRS2(CLV, true, true, NCHV)    \
  else if (bb) LWRS2(CLV, true, false, NCHV)     \
  else if (jac) LWRS2(CLV, false, true, NCHV)    \
  else LWRS2(CLV, false, false, NCHV)
      if (nch == 16) {
        switch (clv) {
          case 6: LWRS(6, 16); break;
          case 7: LWRS(7, 16); break;
          default: LWRS(9, 16); break;
        }
      } else {
        if (clv == 9) { LWRS(9, 8); } else { LWRS(10, 8); }
      }
#undef LWRS
#undef LWRS2
      RB_LAUNCH_CHECK();
      return;
    }
  }
  if (const int cl = resc ? 0 : reg_chunk_len(nlay, ncol)) {
    LwNoscatRegParams q;
    q.ncol = ncol; q.nlay = nlay; q.ngpt = ngpt; q.top_at_1 = p.top_at_1; q.nmus = nmus; q.Ds = p.Ds;
    q.weights = p.weights; q.tau = p.tau; q.lay_source = p.lay_source; q.lev_source = p.lev_source;
    q.sfc_emis = p.sfc_emis; q.sfc_src = p.sfc_src; q.inc_flux = p.inc_flux; q.flux_up = p.flux_up;
    q.flux_dn = p.flux_dn; q.do_broadband = bb; q.bb_up = p.bb_up; q.bb_dn = p.bb_dn; q.do_jac = jac;
    q.sfc_srcJac = p.sfc_srcJac; q.flux_upJac = p.flux_upJac;
    const int groups = (bb || jac) ? std::min(tl_express.groups, ngpt) : reg_gpt_groups(ncol, ngpt);
    q.gpt_per_block = ceil_div(ngpt, groups);
    q.accumulate = tl_express.accumulate; q.group_stride = tl_express.group_stride;
    dim3 grid(ceil_div(ncol, (kRegThreads / 32) * kRegCols), ceil_div(ngpt, q.gpt_per_block));
    // TMA tile staging of tau / lay_source / lev_source (kernels/tma.cuh) whenever the planes can be described
    LwTmaMaps maps;
    // layers per lane (see the switch below) and the tile height: 8*CL rows when nlay is not a multiple of 8 - the rows
    // beyond the plane arrive zero-filled and act as pass-through layers, so every TMA launch runs the FULL instantiation
    const int nch = reg_lanes(nlay);
    const int clv = nch == 16 ? cl : (cl <= 9 ? 9 : 10), rows = nch * clv;
    q.tile_rows = rows; q.row0 = p.top_at_1 ? 0 : nlay - rows;
    const bool use_tma = solver_tma_enabled() && make_plane_tmap(&maps.tau, q.tau, ncol, nlay, ngpt, rows) &&
                         make_plane_tmap(&maps.lay, q.lay_source, ncol, nlay, ngpt, rows) &&
                         make_plane_tmap(&maps.lev, q.lev_source, ncol, nlay + 1, ngpt, rows + 1);
    {
      KernelTimer timer("lw_noscat_reg_kernel");
#define LWREG2(CLV, BBV, JACV)                                                                              \
  {                                                                                                         \
    if (use_tma && reg_minb() == 2) {                                                                       \
      const size_t smem = lw_noscat_reg_tma_smem(rows);                                                     \
      auto kern = rows == nlay ? (nmus == 1 ? lw_noscat_reg_kernel<CLV, BBV, JACV, 2, true, 1, true>        \
                                            : lw_noscat_reg_kernel<CLV, BBV, JACV, 2, true, 1, false>)      \
                               : (nmus == 1 ? lw_noscat_reg_kernel<CLV, BBV, JACV, 2, true, 2, true>        \
                                            : lw_noscat_reg_kernel<CLV, BBV, JACV, 2, true, 2, false>);     \
      RB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
      kern<<<grid, kRegThreads, smem, stream()>>>(q, maps);                                                 \
    } else {                                                                                                \
      const size_t smem = (size_t)2 * lw_noscat_reg_slots<CLV>() * kRegThreads * sizeof(Float);             \
      auto kern = reg_minb() == 2 ? lw_noscat_reg_kernel<CLV, BBV, JACV, 2, false> : lw_noscat_reg_kernel<CLV, BBV, JACV, 3, false>; \
      RB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
      kern<<<grid, kRegThreads, smem, stream()>>>(q, maps);                                                 \
    }                                                                                                       \
  }
#define LWREG(CLV)                                    \
  if (bb && jac) LWREG2(CLV, true, true)              \
  else if (bb) LWREG2(CLV, true, false)               \
  else if (jac) LWREG2(CLV, false, true)              \
  else LWREG2(CLV, false, false)
      // TMA tiles have 128-byte rows: with CL = 8 the eight chunk lanes of a column read rows 8 apart - the same
      // swizzle phase, an 8-way bank conflict (measured: LW 1.6x slower at 60 layers) - so nlay <= 64 runs CL = 9
      // too (rows 9 apart: conflict free; the extra cells are pass-through padding).  The chunk length fixes the
      // association of the chunk-level scan, so it must not depend on whether TMA is usable (odd ncol):
      // results are bit-identical under column subsetting (tests/test_rte_lw_solver_unit.py).
#define LWREG16_2(CLV, BBV, JACV)                                                                            \
  {                                                                                                         \
    if (use_tma) {                                                                                          \
      const size_t smem = lw_noscat_reg_tma_smem(rows, reg_threads(16));                                    \
      auto kern = lw_noscat_reg_kernel<CLV, BBV, JACV, 1, true, 2, false, 16>;                              \
      RB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
      kern<<<grid, reg_threads(16), smem, stream()>>>(q, maps);                                             \
    } else {                                                                                                \
      const size_t smem = (size_t)2 * lw_noscat_reg_slots<CLV>() * reg_threads(16) * sizeof(Float);         \
      auto kern = lw_noscat_reg_kernel<CLV, BBV, JACV, 1, false, 0, false, 16>;                             \
      RB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
      kern<<<grid, reg_threads(16), smem, stream()>>>(q, maps);                                             \
    }                                                                                                       \
  }
#define LWREG16(CLV)                                    \
  if (bb && jac) LWREG16_2(CLV, true, true)             \
  else if (bb) LWREG16_2(CLV, true, false)              \
  else if (jac) LWREG16_2(CLV, false, true)             \
  else LWREG16_2(CLV, false, false)
      if (nch == 16) {   // 80 < nlay <= 144: 16 lanes per column
        switch (cl) {
          case 6: LWREG16(6); break;
          case 7: LWREG16(7); break;
          default: LWREG16(9); break;
        }
      } else
      switch (cl) {
        case 8:
        case 9: LWREG(9); break;
        default: LWREG(10); break;
      }
#undef LWREG16
#undef LWREG16_2
#undef LWREG
#undef LWREG2
      RB_LAUNCH_CHECK();
    }
    return;
  }
  const int nlev = nlay + 1;
  const size_t per_col = (size_t)3 * nlay + (resc ? 2 * nlay + 2 * nlev : 0) + (bb ? 2 * nlev : 0) + (jac ? nlev : 0);
  const int tc = pick_tc(per_col);
  // spectrally integrated outputs are accumulated inside one CTA in g-point order -> one group
  const int groups = (bb || jac) ? 1 : gpt_groups(ncol, tc, ngpt);
  p.gpt_per_block = ceil_div(ngpt, groups);
  dim3 grid(ceil_div(ncol, tc), ceil_div(ngpt, p.gpt_per_block));
  const size_t smem = per_col * tc * sizeof(Float);
  DISPATCH_TC(tc, lw_noscat_kernel, p, grid, smem);
}

void rte_lw_solver_2stream(const int* ncol_, const int* nlay_, const int* ngpt_, const Bool* top_at_1,
                           const Float* tau, const Float* ssa, const Float* g, const Float* lay_source,
                           const Float* lev_source, const Float* sfc_emis, const Float* sfc_src,
                           const Float* inc_flux, Float* flux_up, Float* flux_dn) {
  const int ncol = *ncol_, nlay = *nlay_, ngpt = *ngpt_;
  const size_t nc = ncol, ncl = nc * nlay, nclp = nc * (nlay + 1), ncg = nc * ngpt;
  DevArg<Float> a_tau(tau, ncl * ngpt, Dir::In), a_ssa(ssa, ncl * ngpt, Dir::In), a_g(g, ncl * ngpt, Dir::In),
      a_lay(lay_source, ncl * ngpt, Dir::In), a_lev(lev_source, nclp * ngpt, Dir::In), a_em(sfc_emis, ncg, Dir::In),
      a_ss(sfc_src, ncg, Dir::In), a_inc(inc_flux, ncg, Dir::In);
  DevArg<Float> a_fu(flux_up, nclp * ngpt, Dir::Out), a_fd(flux_dn, nclp * ngpt, Dir::Out);
  Lw2sParams p;
  p.ncol = ncol; p.nlay = nlay; p.ngpt = ngpt; p.top_at_1 = *top_at_1 ? 1 : 0; p.lev_per_gpt = g_lw2s_lev_per_gpt.load();
  p.tau = a_tau; p.ssa = a_ssa; p.g = a_g; p.lay_source = a_lay; p.lev_source = a_lev; p.sfc_emis = a_em;
  p.sfc_src = a_ss; p.inc_flux = a_inc; p.flux_up = a_fu; p.flux_dn = a_fd;
  if (const int cl = reg_chunk_len(nlay, ncol)) {
    Lw2sRegParams q;
    q.ncol = ncol; q.nlay = nlay; q.ngpt = ngpt; q.top_at_1 = p.top_at_1; q.lev_per_gpt = p.lev_per_gpt;
    q.tau = p.tau; q.ssa = p.ssa; q.g = p.g; q.lay_source = p.lay_source; q.lev_source = p.lev_source;
    q.sfc_emis = p.sfc_emis; q.sfc_src = p.sfc_src; q.inc_flux = p.inc_flux; q.flux_up = p.flux_up;
    q.flux_dn = p.flux_dn;
    q.gpt_per_block = ceil_div(ngpt, reg_gpt_groups(ncol, ngpt));
    dim3 grid(ceil_div(ncol, (kRegThreads / 32) * kRegCols), ceil_div(ngpt, q.gpt_per_block));
    Lw2sTmaMaps maps;
    const int nch = reg_lanes(nlay);
    const int clv = nch == 16 ? cl : (cl <= 9 ? 9 : 10), rows = nch * clv;   // zero-filled padded tiles, see rte_lw_solver_noscat
    q.tile_rows = rows; q.row0 = p.top_at_1 ? 0 : nlay - rows;
    const bool use_tma = solver_tma_enabled() && make_plane_tmap(&maps.tau, q.tau, ncol, nlay, ngpt, rows) &&
                         make_plane_tmap(&maps.ssa, q.ssa, ncol, nlay, ngpt, rows) && make_plane_tmap(&maps.g, q.g, ncol, nlay, ngpt, rows) &&
                         make_plane_tmap(&maps.lay, q.lay_source, ncol, nlay, ngpt, rows) &&
                         make_plane_tmap(&maps.lev, q.lev_source, ncol, nlay + 1, ngpt, rows + 1);
    {
      KernelTimer timer("lw_2stream_reg_kernel");
#define LW2S(CLV)                                                                                          \
  if (use_tma) {                                                                                           \
    const size_t smem = lw_2stream_reg_tma_smem(rows);                                                     \
    auto kern = rows == nlay ? lw_2stream_reg_kernel<CLV, true, 1> : lw_2stream_reg_kernel<CLV, true, 2>;  \
    RB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    kern<<<grid, kRegThreads, smem, stream()>>>(q, maps);                                                  \
  } else {                                                                                                 \
    lw_2stream_reg_kernel<CLV, false, 0><<<grid, kRegThreads, 0, stream()>>>(q, maps);                        \
  }
      // the chunk length fixes the association of the chunk-level scan: it must not depend on use_tma (see
      // rte_lw_solver_noscat); CL = 8 conflicts on the TMA tiles, so nlay <= 64 runs CL = 9 as well
#define LW2S16(CLV)                                                                                        \
  if (use_tma) {                                                                                           \
    const size_t smem = lw_2stream_reg_tma_smem(rows);                                                     \
    auto kern = lw_2stream_reg_kernel<CLV, true, 2, 16>;                                                   \
    RB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    kern<<<grid, reg_threads(16), smem, stream()>>>(q, maps);                                              \
  } else {                                                                                                 \
    lw_2stream_reg_kernel<CLV, false, 0, 16><<<grid, reg_threads(16), 0, stream()>>>(q, maps);             \
  }
      if (nch == 16) {
        switch (cl) {
          case 6: LW2S16(6); break;
          case 7: LW2S16(7); break;
          default: LW2S16(9); break;
        }
      } else
      switch (cl) {
        case 8:
        case 9: LW2S(9); break;
        default: LW2S(10); break;
      }
#undef LW2S16
#undef LW2S
      RB_LAUNCH_CHECK();
    }
    return;
  }
  const size_t per_col = (size_t)6 * nlay;
  const int tc = pick_tc(per_col);
  const int groups = gpt_groups(ncol, tc, ngpt);
  p.gpt_per_block = ceil_div(ngpt, groups);
  dim3 grid(ceil_div(ncol, tc), ceil_div(ngpt, p.gpt_per_block));
  DISPATCH_TC(tc, lw_2stream_kernel, p, grid, per_col * tc * sizeof(Float));
}

void rte_sw_solver_noscat(const int* ncol_, const int* nlay_, const int* ngpt_, const Bool* top_at_1,
                          const Float* tau, const Float* mu0, const Float* inc_flux_dir, Float* flux_dir) {
  const int ncol = *ncol_, nlay = *nlay_, ngpt = *ngpt_;
  const size_t nc = ncol, ncl = nc * nlay, nclp = nc * (nlay + 1), ncg = nc * ngpt;
  DevArg<Float> a_tau(tau, ncl * ngpt, Dir::In), a_mu(mu0, ncl, Dir::In), a_inc(inc_flux_dir, ncg, Dir::In);
  DevArg<Float> a_out(flux_dir, nclp * ngpt, Dir::Out);
  KernelTimer timer("sw_noscat_kernel");
  sw_noscat_kernel<<<ceil_div((long long)ncg, kEltThreads), kEltThreads, 0, stream()>>>(
      ncol, nlay, ngpt, *top_at_1 ? 1 : 0, a_tau, a_mu, a_inc, a_out);
  RB_LAUNCH_CHECK();
}

void rte_sw_solver_2stream(const int* ncol_, const int* nlay_, const int* ngpt_, const Bool* top_at_1,
                           const Float* tau, const Float* ssa, const Float* g, const Float* mu0,
                           const Float* sfc_alb_dir, const Float* sfc_alb_dif, const Float* inc_flux_dir,
                           Float* flux_up, Float* flux_dn, Float* flux_dir, const Bool* has_dif_bc,
                           const Float* inc_flux_dif, const Bool* do_broadband, Float* broadband_up,
                           Float* broadband_dn, Float* broadband_dir) {
  const int ncol = *ncol_, nlay = *nlay_, ngpt = *ngpt_;
  const bool bb = *do_broadband, bc = *has_dif_bc;
  const size_t nc = ncol, ncl = nc * nlay, nclp = nc * (nlay + 1), ncg = nc * ngpt;
  DevArg<Float> a_tau(tau, ncl * ngpt, Dir::In), a_ssa(ssa, ncl * ngpt, Dir::In), a_g(g, ncl * ngpt, Dir::In),
      a_mu(mu0, ncl, Dir::In), a_ad(sfc_alb_dir, ncg, Dir::In), a_af(sfc_alb_dif, ncg, Dir::In),
      a_inc(inc_flux_dir, ncg, Dir::In), a_dif(inc_flux_dif, ncg, Dir::In, bc);
  // g-point outputs are decoys (and may alias one another) when do_broadband is true: mo_rte_sw.F90:204-207
  DevArg<Float> a_fu(flux_up, nclp * ngpt, Dir::Out, !bb), a_fd(flux_dn, nclp * ngpt, Dir::Out, !bb),
      a_fr(flux_dir, nclp * ngpt, Dir::Out, !bb);
  DevArg<Float> a_bu(broadband_up, nclp, Dir::Out, bb), a_bd(broadband_dn, nclp, Dir::Out, bb),
      a_br(broadband_dir, nclp, Dir::Out, bb);
  SwParams p;
  p.ncol = ncol; p.nlay = nlay; p.ngpt = ngpt; p.top_at_1 = *top_at_1 ? 1 : 0;
  p.tau = a_tau; p.ssa = a_ssa; p.g = a_g; p.mu0 = a_mu; p.sfc_alb_dir = a_ad; p.sfc_alb_dif = a_af;
  p.inc_flux_dir = a_inc; p.flux_up = a_fu; p.flux_dn = a_fd; p.flux_dir = a_fr; p.has_dif_bc = bc;
  p.inc_flux_dif = a_dif; p.do_broadband = bb; p.bb_up = a_bu; p.bb_dn = a_bd; p.bb_dir = a_br;
  if (const int cl = reg_chunk_len(nlay, ncol)) {
    SwRegParams q;
    q.ncol = ncol; q.nlay = nlay; q.ngpt = ngpt; q.top_at_1 = p.top_at_1; q.tau = p.tau; q.ssa = p.ssa; q.g = p.g;
    q.mu0 = p.mu0; q.sfc_alb_dir = p.sfc_alb_dir; q.sfc_alb_dif = p.sfc_alb_dif; q.inc_flux_dir = p.inc_flux_dir;
    q.flux_up = p.flux_up; q.flux_dn = p.flux_dn; q.flux_dir = p.flux_dir; q.has_dif_bc = bc;
    q.inc_flux_dif = p.inc_flux_dif; q.do_broadband = bb; q.bb_up = p.bb_up; q.bb_dn = p.bb_dn; q.bb_dir = p.bb_dir;
    const int groups = bb ? std::min(tl_express.groups, ngpt) : reg_gpt_groups(ncol, ngpt);
    q.gpt_per_block = ceil_div(ngpt, groups);
    q.accumulate = tl_express.accumulate; q.group_stride = tl_express.group_stride;
    dim3 grid(ceil_div(ncol, (kRegThreads / 32) * kRegCols), ceil_div(ngpt, q.gpt_per_block));
    // TMA tile staging of tau / ssa / g (kernels/tma.cuh) whenever the planes can be described
    SwTmaMaps maps;
    const int nch = reg_lanes(nlay);
    // nlay <= 64 keeps 8 layers per lane here (the LW kernels run 9: with CL = 8 the eight lanes of a column read tile rows
    // 8 apart - one swizzle phase, 8-way bank conflicts): this kernel is fp64-bound, and 64 slots with conflicts beat 72
    // without (B200, 65,536 x 60, select-free padded tiles: 12.8 vs 13.9 ms)
    const int cl_sw = (nch == 8 && cl == 8 && !sw_cl8()) ? 9 : cl;
    const int rows = nch * cl_sw;   // zero-filled padded tiles, see rte_lw_solver_noscat
    q.tile_rows = rows; q.row0 = p.top_at_1 ? 0 : nlay - rows;
    const bool use_tma = solver_tma_enabled() && make_plane_tmap(&maps.tau, q.tau, ncol, nlay, ngpt, rows) &&
                         make_plane_tmap(&maps.ssa, q.ssa, ncol, nlay, ngpt, rows) && make_plane_tmap(&maps.g, q.g, ncol, nlay, ngpt, rows);
    if (use_tma && nch == 8 && sw_ws_enabled()) {
      // warp-specialised kernel: same shapes as the 8-lane TMA instantiations; MERGED follows the shape as below
      KernelTimer timer("sw_2stream_ws_kernel");
      const bool merged = rows == nlay && !force_pad();
      switch (cl_sw) {
        case 8: dispatch_sw_ws<8>(q, maps, grid, bb, merged); break;
        case 9: dispatch_sw_ws<9>(q, maps, grid, bb, merged); break;
        default: dispatch_sw_ws<10>(q, maps, grid, bb, merged); break;
      }
      RB_LAUNCH_CHECK();
      return;
    }
    {
      KernelTimer timer("sw_2stream_reg_kernel");
#define SWREG2(CLV, BBV)                                                                                    \
  {                                                                                                         \
    const bool lean = BBV && reg_minb() == 3;                                                               \
    const bool sacc = BBV && reg_sacc();                                                                    \
    if (use_tma && !lean) {                                                                                 \
      const size_t smem = sacc ? sw_reg_tma_smem<CLV, true>(rows) : sw_reg_tma_smem<CLV, false>(rows);      \
      auto kern = sacc ? sw_2stream_reg_kernel<CLV, BBV, 2, BBV, true, 2>                                   \
                       : ((rows == nlay && !force_pad()) ? sw_2stream_reg_kernel<CLV, BBV, 2, false, true, 1> \
                                       : sw_2stream_reg_kernel<CLV, BBV, 2, false, true, 2>);               \
      RB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
      kern<<<grid, kRegThreads, smem, stream()>>>(q, maps);                                                 \
    } else {                                                                                                \
      const size_t smem = (size_t)(lean ? sw_reg_smem_slots<CLV, true>() : sw_reg_smem_slots<CLV, false>()) * \
                          kRegThreads * sizeof(Float);                                                      \
      auto kern = lean ? sw_2stream_reg_kernel<CLV, BBV, 3, BBV, false> : sw_2stream_reg_kernel<CLV, BBV, 2, false, false>; \
      RB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
      kern<<<grid, kRegThreads, smem, stream()>>>(q, maps);                                                 \
    }                                                                                                       \
  }
#define SWREG(CLV) \
  if (bb) SWREG2(CLV, true) else SWREG2(CLV, false)
      // (CL = 8 conflicts on the TMA tiles as in rte_lw_solver_noscat, but this kernel is fp64-bound: the padding of
      // CL = 9 costs more than the conflicts - measured 39.2 vs 37.7 ms at 131,072 x 60)
#define SWREG16_2(CLV, BBV)                                                                                 \
  {                                                                                                         \
    if (use_tma) {                                                                                          \
      const size_t smem = sw_reg_tma_smem<CLV, false>(rows, reg_threads(16));                               \
      auto kern = sw_2stream_reg_kernel<CLV, BBV, 1, false, true, 2, 16>;                                   \
      RB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
      kern<<<grid, reg_threads(16), smem, stream()>>>(q, maps);                                             \
    } else {                                                                                                \
      const size_t smem = (size_t)sw_reg_smem_slots<CLV, false>() * reg_threads(16) * sizeof(Float);        \
      auto kern = sw_2stream_reg_kernel<CLV, BBV, 1, false, false, 0, 16>;                                  \
      RB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
      kern<<<grid, reg_threads(16), smem, stream()>>>(q, maps);                                             \
    }                                                                                                       \
  }
#define SWREG16(CLV) \
  if (bb) SWREG16_2(CLV, true) else SWREG16_2(CLV, false)
      if (nch == 16) {
        switch (cl) {
          case 6: SWREG16(6); break;
          case 7: SWREG16(7); break;
          default: SWREG16(9); break;
        }
      } else
      switch (cl_sw) {
        case 8: SWREG(8); break;
        case 9: SWREG(9); break;
        default: SWREG(10); break;
      }
#undef SWREG16
#undef SWREG16_2
#undef SWREG
#undef SWREG2
      RB_LAUNCH_CHECK();
    }
    return;
  }
  const int nlev = nlay + 1;
  const size_t per_col = (size_t)6 * nlay + (bb ? 3 * nlev : 0);
  const int tc = pick_tc(per_col);
  const int groups = bb ? 1 : gpt_groups(ncol, tc, ngpt);
  p.gpt_per_block = ceil_div(ngpt, groups);
  dim3 grid(ceil_div(ncol, tc), ceil_div(ngpt, p.gpt_per_block));
  DISPATCH_TC(tc, sw_2stream_kernel, p, grid, per_col * tc * sizeof(Float));
}

}  // extern "C"
