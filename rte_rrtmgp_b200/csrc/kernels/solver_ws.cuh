// solver_ws.cuh - warp-specialised SW two-stream solver (mo_rte_solver_kernels.F90:503-609, 985-1245).
//
// The register kernel (solver_reg.cuh) runs phase A (the exp / sqrt / divide-heavy two-stream cell algebra, ~2/3 of the
// fp64 work, embarrassingly parallel) and phase B (direct beam + adding as chunk-level scans: dependent chains, warp
// shuffles) in the SAME warps, at 254 registers and two warps per scheduler: whenever both warps of a scheduler sit in
// the latency-bound phase B the fp64 pipe idles (69 % active under ncu).  Here the two phases live in DIFFERENT warps of
// one CTA (1 CTA per SM, 16 columns, loop over g-points):
//   producers  NPW warps, few registers (setmaxnreg.dec): thread = up to CPT fixed cells (tile row, column) of the
//              16-column x ROWS tile; per g-point they read tau / ssa / g of their cells from the TMA-filled stage,
//              run sw_two_stream_cell() and write Rdif, Tdif, Rdir, Tdir, Tnoscat back IN PLACE (planes 0-2 of the
//              stage are the inputs, planes 3-4 are extra).  Everything that depends on (column, layer) only - mu0_s,
//              3*mu0_s, the refined reciprocal of mu0_s, the mu0 > 0 test - is hoisted out of the g-point loop into
//              registers (bit-identical to the register kernel: same operations on the same operands).
//   consumers  4 warps, many registers (setmaxnreg.inc): lane = (column, chunk of CL layers) exactly as in the register
//              kernel; per g-point they copy the 5*CL coefficients of their chunk into registers, hand the stage back,
//              and run the register kernel's phase B unchanged (same code: adding_reg, the prefix-scan hand-overs).
//   stages     NST-deep ring of (5 planes x ROWS x 16) tiles; per stage three mbarriers: full (TMA bytes landed),
//              coef (every producer warp has written its cells), free (every consumer warp has copied its cells).
//              One producer lane issues the TMA for g-point n+2 at the top of iteration n.
// Results are bit-identical to sw_2stream_reg_kernel on the same shape (tests/test_kernels_parity.py).
#pragma once
#include "solver_reg.cuh"

namespace rrtmgpb {

constexpr int kWsStages = 4;
constexpr int kWsPrefetch = 2;         // TMA issue distance in g-points
constexpr int kWsConsumerThreads = 128;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// orders this thread's (and, cumulatively, observed) generic-proxy accesses to shared memory before later async-proxy
// (TMA) accesses: the stage a TMA refills was last WRITTEN by generic stores (the producers' coefficients)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__host__ __device__ constexpr size_t ws_tile_bytes(int rows) { return ((size_t)rows * kTmaCols * sizeof(Float) + 1023) & ~(size_t)1023; }
__host__ __device__ constexpr size_t sw_ws_smem(int rows) {
  return (size_t)kWsStages * 5 * ws_tile_bytes(rows) + (size_t)2 * 4 * kWsConsumerThreads * sizeof(Float) + 3 * kWsStages * sizeof(uint64_t);
}
// Register budgets after setmaxnreg (multiples of 8; consumers*128 + producers*32*NPW must not exceed what the LAUNCH
// allocated, which is ptxas' choice under __launch_bounds__, not 64K / threads: a budget beyond it makes setmaxnreg.inc
// wait forever).  Measured on B200 (65,536 x 72 x 224, profiles/r2_ws_ab.jsonl; the register kernel takes 13.45 ms):
//   12 producer warps (512 threads x 128; consumers 224, producers 96)   19.7 ms
//    8 producer warps (384 threads x 168; consumers 232, producers 136)  15.1 ms   <- the only budget kept
//    4 producer warps (256 threads x 255; nothing to move)               20.9 ms
// i.e. NOT faster than the register kernel: with the broadband accumulators the consumers need ~200 registers, so an
// SM holds four consumer warps and at most eight useful producer warps - three warps per scheduler, one of them the
// latency-bound consumer whose ~140 dependent fp64 operations + shuffles per g-point are the critical path.  The kernel
// stays as an opt-in variant (rrtmgpb_set_solver_variant(3) / RRTMGPB_SW_WS=1) with its parity tests.
template <int NPW> struct WsRegs;
template <> struct WsRegs<8>  { static constexpr int consumer = 232, producer = 136; };

template <int CL, bool BB, bool MERGED, int NPW>
__global__ void __launch_bounds__(kWsConsumerThreads + 32 * NPW, 1) sw_2stream_ws_kernel(const SwRegParams p,
                                                                                        const __grid_constant__ SwTmaMaps tm) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int NST = kWsStages, PF = kWsPrefetch;
  constexpr int kRegChunks = 8, kRegCols = 4, kRegThreads = kWsConsumerThreads;
  constexpr int ROWS = kRegChunks * CL;            // == p.tile_rows: the tile always fills the lanes (rows outside the plane arrive zero-filled)
  constexpr int NPT = 32 * NPW;                    // producer threads
  constexpr int NCELL = ROWS * kTmaCols;
  constexpr int CPT = (NCELL + NPT - 1) / NPT;     // cells per producer thread
  constexpr size_t tileb = ws_tile_bytes(ROWS);
  constexpr int tile_elems = (int)(tileb / sizeof(Float));
  constexpr size_t stageb = 5 * tileb;
  Float* sm = reinterpret_cast<Float*>(smem_raw + NST * stageb);            // consumers' cp.async slots [2][4][128]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sm + (size_t)2 * 4 * kRegThreads);
  uint64_t* coef_bar = full_bar + NST;
  uint64_t* free_bar = full_bar + 2 * NST;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nlay = p.nlay, nlev = nlay + 1;
  const size_t ncol = p.ncol, nclp = ncol * nlev;
  const int row0 = p.row0;
  const int gb = blockIdx.y * p.gpt_per_block, ge = min(p.ngpt, gb + p.gpt_per_block);
  const int niter = ge - gb;
  const int cta_col0 = blockIdx.x * kTmaCols;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&coef_bar[s], NPW);
      mbar_init(&free_bar[s], kRegThreads / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp >= kRegThreads / 32) {
    // =============================== producers: phase A ===============================
    if constexpr (WsRegs<NPW>::producer > 0) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(WsRegs<NPW>::producer));
    const int pt = threadIdx.x - kRegThreads;
    const Float min_mu0 = sqrt((Float)RB_EPS);  // :1006
    int eoff[CPT];
    Float mu0_s[CPT], mu0_3[CPT], r_mu0[CPT];
    bool lit[CPT];
#pragma unroll
    for (int m = 0; m < CPT; ++m) {
      const int idx = min(pt + NPT * m, NCELL - 1);
      const int r = idx / kTmaCols, cc = idx % kTmaCols;
      eoff[m] = tile_off(r, cc);
      const int l = min(max(row0 + r, 0), nlay - 1);                 // padding rows: any valid mu0 (their cells are zero)
      const int col = min(cta_col0 + cc, p.ncol - 1);
      const Float mu0 = p.mu0[(size_t)col + ncol * l];
      mu0_s[m] = fmax(min_mu0, mu0);
      mu0_3[m] = (Float)3 * mu0_s[m];
      r_mu0[m] = rb_rcp1(mu0_s[m]);
      lit[m] = mu0 > (Float)0;                                       // :1122-1125: no source for diffuse light where mu0 <= 0
    }
    auto issue = [&](int n) {
      const int s = n % NST;
      mbar_expect_tx(&full_bar[s], (uint32_t)(3 * ROWS * kTmaCols * sizeof(Float)));
      unsigned char* dst = smem_raw + (size_t)s * stageb;
      tma_load_tile(dst, &tm.tau, &full_bar[s], cta_col0, row0, gb + n);
      tma_load_tile(dst + tileb, &tm.ssa, &full_bar[s], cta_col0, row0, gb + n);
      tma_load_tile(dst + 2 * tileb, &tm.g, &full_bar[s], cta_col0, row0, gb + n);
    };
    const bool issuer = pt == 0;
    if (issuer) {
      for (int n = 0; n < PF && n < niter; ++n) issue(n);
    }
    for (int n = 0; n < niter; ++n) {
      const int s = n % NST;
      if (issuer && n + PF < niter) {
        const int n2 = n + PF;
        if (n2 >= NST) {  // the stage's previous tenant (g-point n2 - NST) must have been copied out by every consumer warp
          mbar_wait(&free_bar[n2 % NST], (uint32_t)((n2 / NST - 1) & 1));
          fence_proxy_async_smem();
        }
        issue(n2);
      }
      __syncwarp();
      mbar_wait(&full_bar[s], (uint32_t)((n / NST) & 1));
      Float* st = reinterpret_cast<Float*>(smem_raw + (size_t)s * stageb);
#pragma unroll
      for (int m = 0; m < CPT; ++m) {
        if (CPT * NPT == NCELL || pt + NPT * m < NCELL) {   // (warp-uniform: NCELL and NPT are multiples of 32)
          Float* e = st + eoff[m];
          const Float tau_s = e[0], w0_s = e[tile_elems], g_s = e[2 * tile_elems];
          Float Rdif, Tdif, Rdir, Tdir, Tnoscat;
          sw_two_stream_cell(tau_s, w0_s, g_s, mu0_s[m], mu0_3[m], r_mu0[m], MERGED, Rdif, Tdif, Rdir, Tdir, Tnoscat);
          e[0] = Rdif;
          e[tile_elems] = Tdif;
          e[2 * tile_elems] = lit[m] ? Rdir : (Float)0;
          e[3 * tile_elems] = lit[m] ? Tdir : (Float)0;
          e[4 * tile_elems] = Tnoscat;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&coef_bar[s]);
    }
    return;
  }

  // =============================== consumers: phase B ===============================
  if constexpr (WsRegs<NPW>::consumer > 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(WsRegs<NPW>::consumer));
  const int c = lane / kRegChunks, j = lane % kRegChunks;
  const int cw = warp * kRegCols + c;
  const int col_raw = cta_col0 + cw;
  const bool col_ok = col_raw < p.ncol;
  const int col = col_ok ? col_raw : p.ncol - 1;
  const RegOrient o{nlay, p.top_at_1};
  const int k0 = j * CL;
  constexpr int NS = 4, BC0 = 0;
  auto prefetch = [&](int g, int s2) {
    const size_t gi = (size_t)col + ncol * g;
    cp_async_f(RB_SLOT(sm, NS, s2, BC0 + 0), p.sfc_alb_dir + gi);
    cp_async_f(RB_SLOT(sm, NS, s2, BC0 + 1), p.sfc_alb_dif + gi);
    cp_async_f(RB_SLOT(sm, NS, s2, BC0 + 2), p.inc_flux_dir + gi);
    if (p.has_dif_bc) cp_async_f(RB_SLOT(sm, NS, s2, BC0 + 3), p.inc_flux_dif + gi);
  };
  constexpr int NACC = BB ? CL : 1;
  Float acc_up[NACC], acc_dn[NACC], acc_dir[NACC];
  Float acc_up_top = 0, acc_dn_top = 0, acc_dir_top = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) { acc_up[i] = 0; acc_dn[i] = 0; acc_dir[i] = 0; }
  auto acc_add = [&](int which, int i, Float v) {
    if (which == 0) acc_up[NACC > 1 ? i : 0] += v;
    else if (which == 1) acc_dn[NACC > 1 ? i : 0] += v;
    else acc_dir[NACC > 1 ? i : 0] += v;
  };
  const Float mu0_top = p.mu0[(size_t)col + ncol * o.lay(0)];
  const Float mu0_sfc = p.mu0[(size_t)col + ncol * o.lay(nlay - 1)];
  int roff[CL];   // element offsets of this lane's cells inside a plane of the stage
#pragma unroll
  for (int i = 0; i < CL; ++i)
    roff[i] = tile_off(o.lay(k0 + i) - row0, cw);

  if (niter > 0) prefetch(gb, 0);
  cp_async_commit();
  for (int n = 0; n < niter; ++n) {
    const int g = gb + n, s = n % NST, sb = n & 1;
    if (n + 1 < niter) prefetch(g + 1, sb ^ 1);
    cp_async_commit();
    cp_async_wait<1>();   // boundary values of g have landed (slots are lane-private)
    Float* gup = p.flux_up + nclp * g;
    Float* gdn = p.flux_dn + nclp * g;
    Float* gdir = p.flux_dir + nclp * g;
    Float R[CL], T[CL], A3[CL], A4[CL], A5[CL];  // Rdif, Tdif, Rdir->src_up, Tdir->src_dn, Tnoscat->direct flux
    mbar_wait(&coef_bar[s], (uint32_t)((n / NST) & 1));
    {
      const Float* st = reinterpret_cast<const Float*>(smem_raw + (size_t)s * stageb);
#pragma unroll
      for (int i = 0; i < CL; ++i) {
        const Float* e = st + roff[i];
        R[i] = e[0]; T[i] = e[tile_elems]; A3[i] = e[2 * tile_elems]; A4[i] = e[3 * tile_elems]; A5[i] = e[4 * tile_elems];
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&free_bar[s]);   // this warp's cells are in registers: the stage may be refilled
    const Float alb_dir = *RB_SLOT(sm, NS, sb, BC0 + 0), alb_dif = *RB_SLOT(sm, NS, sb, BC0 + 1);
    const Float dir_top_g = *RB_SLOT(sm, NS, sb, BC0 + 2) * mu0_top;                  // :575
    const Float dn_top = p.has_dif_bc ? *RB_SLOT(sm, NS, sb, BC0 + 3) : (Float)0;     // :579-583
    // ---------------- phase B1: direct beam and its sources, :1110-1112 (as in sw_2stream_reg_kernel) ----------------
    Float dir = dir_top_g;
    if (j == 0) {
      if (BB) { acc_dir_top += dir; acc_dn_top += dir; }
      else if (col_ok) gdir[(size_t)col + ncol * o.lev(0)] = dir;
    }
    {
      Float P = 1;
#pragma unroll
      for (int i = 0; i < CL; ++i) P = A5[i] * P;
      Float out;
      dir = product_handoff_down<kRegChunks>(j, P, dir_top_g, out);
#pragma unroll
      for (int i = 0; i < CL; ++i) {
        const Float s_up = A3[i] * dir, s_dn = A4[i] * dir;
        dir = A5[i] * dir;
        A3[i] = s_up;
        A4[i] = s_dn;
        if (BB) { acc_add(2, i, dir); acc_add(1, i, dir); }  // :604, direct part of :603
        else if (k0 + i < nlay && col_ok) gdir[(size_t)col + ncol * o.lev(k0 + i + 1)] = dir;
        A5[i] = dir;  // direct flux below layer k0+i, for the g-point totals (:606)
      }
    }
    const Float src_sfc = (mu0_sfc > (Float)0) ? dir * alb_dir : (Float)0;  // :1120
    auto top = [&](Float fup, Float fdn) {
      if (BB) { acc_up_top += fup; acc_dn_top += fdn; }                        // :602, diffuse part of :603
      else if (col_ok) {
        const size_t q = (size_t)col + ncol * o.lev(0);
        gup[q] = fup;
        gdn[q] = fdn + dir_top_g;                                             // :606
      }
    };
    auto lev = [&](int i, Float fup, Float fdn) {
      if (BB) { acc_add(0, i, fup); acc_add(1, i, fdn); }
      else if (k0 + i < nlay && col_ok) {
        const size_t q = (size_t)col + ncol * o.lev(k0 + i + 1);
        gup[q] = fup;
        gdn[q] = fdn + A5[i];                                                 // :606
      }
    };
    adding_reg<CL, kRegChunks>(j, R, T, A3, A4, alb_dif, src_sfc, dn_top, top, lev);
  }
  if (BB && col_ok) {
    const size_t goff = p.group_stride * blockIdx.y;  // 0 unless the express path splits a launch's g-points
    Float *bu = p.bb_up + goff, *bd = p.bb_dn + goff, *br = p.bb_dir + goff;
    auto put = [&](Float* dst, size_t o2, Float v) { dst[o2] = p.accumulate ? dst[o2] + v : v; };
#pragma unroll
    for (int i = 0; i < CL; ++i) {
      const int klev = k0 + i + 1;
      if (klev <= nlay) {
        const size_t o2 = (size_t)col + ncol * o.lev(klev);
        put(bu, o2, acc_up[NACC > 1 ? i : 0]); put(bd, o2, acc_dn[NACC > 1 ? i : 0]); put(br, o2, acc_dir[NACC > 1 ? i : 0]);
      }
    }
    if (j == 0) {
      const size_t o2 = (size_t)col + ncol * o.lev(0);
      put(bu, o2, acc_up_top); put(bd, o2, acc_dn_top); put(br, o2, acc_dir_top);
    }
  }
}

}  // namespace rrtmgpb
