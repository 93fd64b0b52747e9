// solver_reg.cuh - register-resident, warp-systolic RTE solvers (the hot-path variants).
//
// Layout of the work inside a warp (32 lanes):
//     lane = c*8 + j      c in 0..3  : one of the warp's 4 consecutive columns
//                         j in 0..7  : a CHUNK of CL consecutive layers, counted from the top
// so a warp owns 4 (column, g-point) recurrences at a time and every lane owns CL cells of one of them.
//   inputs   TMA (kernels/tma.cuh): one thread issues one cp.async.bulk.tensor per input plane - the (16 columns x
//            nlay) tile of a g-point - into a two-stage, 128B-swizzled shared-memory ring, completion on an
//            mbarrier; the stage is handed back for g+2 after a __syncthreads() once every warp has read it.
//            Fallback (odd ncol, single precision): every lane copies the inputs of its CL cells for g+1 into its
//            own shared-memory slots with cp.async (LDGSTS), lane-private, completion by cp.async.wait_group;
//   phase A  every lane computes its CL cells (the exp / sqrt / divide-heavy two-stream or Planck-source
//            algebra) into REGISTERS as STRAIGHT-LINE code (branch-free math from fastmath.cuh, padding cells by
//            selects): CL independent cells per lane give the instruction-level parallelism that hides the fp64
//            latencies.  FULL instantiations (nlay == 8*CL) drop the padding selects, clamps and guards;
//   phase B  the layer-serial recurrences (transport / direct beam / adding) as a chunk-level scan: every lane
//            composes the CL layers of its chunk into one map, the 8 chunks of a column are chained with 8
//            hand-overs (one warp shuffle each), and every lane replays its own layers from the incoming state
//            with the reference's per-layer expressions.
//   Broadband sums live in registers of the lane that owns the level, accumulated in g-point order (the
//   reference's order, mo_rte_solver_kernels.F90:216-218,601-604): deterministic, no atomics.
// Layers beyond nlay (padding of the last chunk) are exact pass-through cells (T = 1, R = 0, no source).
// Every input plane is read from HBM once (TMA: 128-byte rows of 16 consecutive columns).
//
// Numerics follow rte/kernels/mo_rte_solver_kernels.F90 (line numbers cited inline).  Two documented
// reassociations w.r.t. the reference (results differ by ~1 ulp of the affected term):
//   * adding: flux_dn(l+1) = a*flux_dn(l) + b with a = Tdif*denom, b = (Rdif*src + src_dn)*denom
//     instead of (Tdif*flux_dn + Rdif*src + src_dn)*denom (:1197-1199)
//   * SW broadband_dn adds the direct beam before the diffuse flux of the same g-point (:603)
#pragma once
#include "../common.cuh"
#include "fastmath.cuh"
#include "tma.cuh"

namespace rrtmgpb {

// A CTA always owns 16 consecutive columns (= one TMA tile row of 128 bytes).  NCH lanes share a column (template
// parameter of every kernel below): 8 lanes x CL <= 10 layers for nlay <= 80 (4 warps per CTA), 16 lanes x CL <= 9
// layers for 80 < nlay <= 144 (8 warps per CTA) - the chunk-level scans simply take one more step.
constexpr int kRegColsPerCta = 16;
__host__ __device__ constexpr int reg_threads(int nch) { return kRegColsPerCta * nch; }
constexpr int kRegThreads = reg_threads(8);  // the 8-lane family (host-side defaults)
constexpr int kRegChunks = 8;
constexpr int kRegCols = 4;

struct RegOrient {
  int nlay, top_at_1;
  __device__ __forceinline__ int lay(int k) const { return top_at_1 ? k : nlay - 1 - k; }
  __device__ __forceinline__ int lev(int k) const { return top_at_1 ? k : nlay - k; }
};

__device__ __forceinline__ Float reg_pi() { return (Float)3.14159265358979323846; }

// 8-byte asynchronous global->shared copy (LDGSTS); lane-private destination
__device__ __forceinline__ void cp_async_f(Float* smem_dst, const Float* gmem_src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(s), "l"(gmem_src), "n"((int)sizeof(Float)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// slot v of this thread in stage s: [s][v][thread] - consecutive threads hit consecutive 8-byte words
#define RB_SLOT(base, nslots, s, v) ((base) + ((size_t)((s) * (nslots) + (v)) * kRegThreads + threadIdx.x))


// ---------------------------------------------------------------------------------------------------
// Chunk-level scan.  Every layer recurrence here is a linear (affine or projective) map of the chain state,
// so a lane can COMPOSE the CL layers of its chunk into one map with all 32 lanes busy, the 8 chunks of a
// column are then chained with 8 cheap hand-overs (one shuffle + a couple of FMAs each), and finally every
// lane REPLAYS its own layers from the now-known incoming state - again with all lanes busy - using the
// reference's per-layer formulas.  Per-level results are therefore computed with exactly the reference's
// expressions; only the state entering each chunk carries the (rounding-level) reassociation of the
// composed map.  This replaces sweeps that ran with 4 of 32 lanes active (55% of the solver's issue slots,
// profiles/r1_prof_v4_*).
// ---------------------------------------------------------------------------------------------------
// x_out = A*x_in + B, chunks chained top -> bottom (j = 0 first).  Returns the state entering this lane's chunk;
// `out` receives the state leaving it.  x_top must hold the boundary value on every lane of the column.
//
// The chain is a prefix scan of affine maps over the 8 chunk lanes of a column (Hillis-Steele, 3 steps of
// 2 shuffles + 2 multiply-adds, then one shuffle to pass the result on) instead of 8 dependent
// shuffle + multiply-add steps: the hand-overs were the longest serial sections of a g-point (three of them per
// g-point in the SW kernel), during which a warp issues next to nothing.  Composition order differs from the serial
// chain only in the association of the products and sums (rounding level, like the chunk composition itself).
#ifndef RB_HANDOFF_SCAN
#define RB_HANDOFF_SCAN 1
#endif
// 1: the scan steps are branch-free - lanes without a partner d chunks away combine with the IDENTITY map (selected on
// the integer pipe) instead of skipping the step under a divergent branch (BRA/BSYNC pairs and partially active fp64
// instructions in round 1's SASS: the fp64 pipe is busy for the whole warp either way)
#ifndef RB_SCAN_SELECT
#define RB_SCAN_SELECT 1
#endif
// 1: adding's upward replay carries (albedo, source) in homogeneous form through the lane's layers, so the per-layer
// reciprocal leaves the 9-step dependent chain: the chain is 2 multiply-adds deep per layer, the reciprocals of all
// levels are then independent of each other (see adding_reg)
#ifndef RB_ADD_HOMOG
#define RB_ADD_HOMOG 0   // measured on B200: 13.56 (1) vs 13.51 ms (0) for the SW solver - no gain, so the reference's per-layer form stays
#endif
// 1: Rdir/Tdir's RT_term*w0/(1 - k^2 mu0^2) shares ONE reciprocal with RT_term itself (see sw cell)
#ifndef RB_SW_MERGED_DIV
#define RB_SW_MERGED_DIV 1
#endif
// 1: the SW kernel's FULL = 2 (zero-filled padded tiles) instantiations carry NO padding select at all: with the two
// divisions kept separate, a zero row gives Tdif = rcp(4)*2*2*exp(-0) = 1 EXACTLY (k = sqrt(4) = 2, rt_den = 4 - every
// factor a power of two), so the padding is pass-through by arithmetic alone
// Measured on B200 (65,536 x 60 / x 72 forced onto the padded instantiation): ONE select per cell (on Tdif) costs the
// SW kernel 17 % (16.28 vs 13.87 ms: the kernel lives on interleaving nine cells at the register limit), none 3 %.
// To keep results independent of WHICH instantiation runs (even / odd ncol decides whether TMA is usable), the choice
// between the shared and the separate reciprocal follows the SHAPE, not the instantiation: shared when nlay fills the
// lanes exactly (FULL = 1, and FULL = 0 on such shapes), separate otherwise (FULL = 2, and FULL = 0 on such shapes).
// 1: max(sqrt(eps), mu0) kept in the lane's mu0 slot and the mu0 > 0 tests as one bit per cell in a register (both depend
// on (column, layer) only): two fp64 compares and a select pair fewer per cell and g-point, bit-identical results - and
// 16.17 instead of 13.45 ms on B200 (255 registers: like every other change that adds a live value to this kernel, it
// costs the compiler the interleaving of the nine cells).  Off.
#ifndef RB_SW_MU0_HOIST
#define RB_SW_MU0_HOIST 0
#endif
#ifndef RB_PAD_NOSELECT
#define RB_PAD_NOSELECT 1
#endif
template <int NCH>
__device__ __forceinline__ Float affine_handoff_down(int j, Float A, Float B, Float x_top, Float& out) {
#if RB_HANDOFF_SCAN
  Float PA = A, PB = B;  // composed map of chunks (j-d+1 .. j): x -> PA*x + PB
#pragma unroll
  for (int d = 1; d < NCH; d <<= 1) {
    Float qa = __shfl_up_sync(0xffffffffu, PA, d, NCH);
    Float qb = __shfl_up_sync(0xffffffffu, PB, d, NCH);
#if RB_SCAN_SELECT
    qa = (j >= d) ? qa : (Float)1; qb = (j >= d) ? qb : (Float)0;
    PB = PA * qb + PB; PA = PA * qa;
#else
    if (j >= d) { PB = PA * qb + PB; PA = PA * qa; }  // this map after the d chunks above it
#endif
  }
  out = PA * x_top + PB;
  const Float got = __shfl_up_sync(0xffffffffu, out, 1, NCH);
  return j == 0 ? x_top : got;
#else
  Float xin = x_top;
  out = 0;
  for (int jj = 0; jj < NCH; ++jj) {
    const Float got = __shfl_up_sync(0xffffffffu, out, 1);
    if (j == jj) {
      if (jj > 0) xin = got;
      out = A * xin + B;
    }
  }
  return xin;
#endif
}
// direct beam: B == 0, the scan carries the products only
template <int NCH>
__device__ __forceinline__ Float product_handoff_down(int j, Float A, Float x_top, Float& out) {
#if RB_HANDOFF_SCAN
  Float PA = A;
#pragma unroll
  for (int d = 1; d < NCH; d <<= 1) {
    Float qa = __shfl_up_sync(0xffffffffu, PA, d, NCH);
#if RB_SCAN_SELECT
    qa = (j >= d) ? qa : (Float)1;
    PA = PA * qa;
#else
    if (j >= d) PA = PA * qa;
#endif
  }
  out = PA * x_top;
  const Float got = __shfl_up_sync(0xffffffffu, out, 1, NCH);
  return j == 0 ? x_top : got;
#else
  return affine_handoff_down<NCH>(j, A, (Float)0, x_top, out);
#endif
}
// chunks chained bottom -> top (j = 7 first); x_bottom needs to be valid on the lane of the last chunk only
template <int NCH>
__device__ __forceinline__ Float affine_handoff_up(int j, Float A, Float B, Float x_bottom, Float& out) {
#if RB_HANDOFF_SCAN
  const Float xb = __shfl_sync(0xffffffffu, x_bottom, NCH - 1, NCH);
  Float PA = A, PB = B;  // composed map of chunks (j+d-1 .. j), applied bottom first
#pragma unroll
  for (int d = 1; d < NCH; d <<= 1) {
    Float qa = __shfl_down_sync(0xffffffffu, PA, d, NCH);
    Float qb = __shfl_down_sync(0xffffffffu, PB, d, NCH);
#if RB_SCAN_SELECT
    qa = (j + d < NCH) ? qa : (Float)1; qb = (j + d < NCH) ? qb : (Float)0;
    PB = PA * qb + PB; PA = PA * qa;
#else
    if (j + d < NCH) { PB = PA * qb + PB; PA = PA * qa; }
#endif
  }
  out = PA * xb + PB;
  const Float got = __shfl_down_sync(0xffffffffu, out, 1, NCH);
  return j == NCH - 1 ? xb : got;
#else
  Float xin = x_bottom;
  out = 0;
  for (int jj = NCH - 1; jj >= 0; --jj) {
    const Float got = __shfl_down_sync(0xffffffffu, out, 1);
    if (j == jj) {
      if (jj < NCH - 1) xin = got;
      out = A * xin + B;
    }
  }
  return xin;
#endif
}

// ---------------------------------------------------------------------------------------------------
// LW no-scattering (mo_rte_solver_kernels.F90:51-240, 620-745), without Tang rescaling.
// ---------------------------------------------------------------------------------------------------
struct LwNoscatRegParams {
  int ncol, nlay, ngpt, top_at_1, nmus;
  const Float *Ds, *weights, *tau, *lay_source, *lev_source, *sfc_emis, *sfc_src, *inc_flux;
  Float *flux_up, *flux_dn;
  int do_broadband;
  Float *bb_up, *bb_dn;
  int do_jac;
  const Float* sfc_srcJac;
  Float* flux_upJac;
  int gpt_per_block;
  // express path (spectrally integrated outputs only): the g-points of a launch may be split over blockIdx.y, every
  // group adding into its OWN copy of the outputs (bb + blockIdx.y * group_stride; summed afterwards in a fixed order:
  // deterministic, no atomics), and a launch may add to what earlier launches (other bands) left there
  int accumulate;
  size_t group_stride;
  // FULL = 2 instantiations (TMA, nlay not a multiple of 8): the TMA box is 8*CL (+1) rows tall and starts at row0
  // (negative for bottom-up columns); rows outside the plane arrive zero-filled, so the tile is addressed like a full
  // one - no clamped row indices - and padding cells are turned into exact pass-through cells by selects on their
  // RESULTS only (FSEL pairs on the integer pipe).  tile_rows = rows of a layer tile (nlay when FULL = 1).
  int tile_rows, row0;
};

template <int CL>
__host__ __device__ constexpr int lw_noscat_reg_slots() { return 3 * CL + 1 + 5; }  // tau, lay, lev(+1), emis, sfc_src, inc_flux, jac, D(angle 1)

// TMA variant (see the SW kernel and kernels/tma.cuh): tau, lay_source (nlay rows) and lev_source (nlay+1 rows) tiles
// by cp.async.bulk.tensor, two stages; the five per-(column, g-point) values keep their lane-private cp.async slots.
struct LwTmaMaps { CUtensorMap tau, lay, lev; };
// RB_LW_STAGES: depth of the TMA ring of lw_noscat_reg_kernel (g-points in flight per CTA).  Neither the fp64 pipe (53 %) nor
// issue (51 %) nor DRAM (4.8 of 6.9 TB/s) is saturated under ncu, which suggested too few bytes in flight (two stages keep
// ~57 KB per SM outstanding); a third stage doubles that and changes nothing: 5.84 vs 5.85 ms on B200 (and 60 layers in padded
// 72-row tiles take the same 5.8 ms as 72): the kernel's time is per g-point iteration - dependent-issue latency of the
// scans with two warps per scheduler - not per byte.  2 stays.
#ifndef RB_LW_STAGES
#define RB_LW_STAGES 2
#endif
__host__ __device__ inline size_t lw_noscat_reg_tma_smem(int nlay, int nthreads = kRegThreads) {
  return RB_LW_STAGES * (2 * tile_bytes(nlay) + tile_bytes(nlay + 1)) + (size_t)(RB_LW_STAGES * 5) * nthreads * sizeof(Float) +
         RB_LW_STAGES * sizeof(uint64_t);
}

// FULL = 1: nlay == 8*CL, every lane's cells are real layers - the padding selects and tests fold away at compile time;
// FULL = 2: zero-filled padded tiles (LwNoscatRegParams::tile_rows); 0: clamped addressing (cp.async fallback).
// ONEMU: a single quadrature angle (the default of rte_lw): the loop over angles folds away.
template <int CL, bool BB, bool JAC, int MINB = 3, bool TMA = false, int FULL = 0, bool ONEMU = false, int NCH = 8>
__global__ void __launch_bounds__(reg_threads(NCH), NCH == 8 ? MINB : 1) lw_noscat_reg_kernel(const LwNoscatRegParams p,
                                                                           const __grid_constant__ LwTmaMaps tm) {
  // no static shared memory: the swizzled TMA tiles need the dynamic window to start 1024-byte aligned
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int kRegChunks = NCH, kRegCols = 32 / NCH, kRegThreads = reg_threads(NCH);  // lanes per column, columns per warp, threads per CTA
  constexpr bool FULLG = FULL == 1;   // no level beyond nlay exists: padding selects and store guards fold away
  const int row0 = (FULL == 2) ? p.row0 : 0, tile_rows = (FULL == 2) ? p.tile_rows : p.nlay;
  const size_t tb_lay = TMA ? tile_bytes(tile_rows) : 0, tb_lev = TMA ? tile_bytes(tile_rows + 1) : 0;
  const size_t stageb = 2 * tb_lay + tb_lev;
  const int te_lay = (int)(tb_lay / sizeof(Float));
  constexpr int NSTG = TMA ? RB_LW_STAGES : 2;                     // ring depth (tiles and boundary-value slots)
  Float* sm = reinterpret_cast<Float*>(smem_raw + NSTG * stageb);  // cp.async slots
  constexpr int NS = TMA ? 5 : lw_noscat_reg_slots<CL>();
  constexpr int BC0 = TMA ? -1 : 3 * CL;                          // boundary-value slots are BC0+1 .. BC0+5
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sm + (size_t)NSTG * NS * kRegThreads);  // TMA: [NSTG] mbarriers
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = lane / kRegChunks, j = lane % kRegChunks;
  const int col_raw = (blockIdx.x * (kRegThreads / 32) + warp) * kRegCols + c;
  const bool col_ok = col_raw < p.ncol;
  const int col = col_ok ? col_raw : p.ncol - 1;  // out-of-range lanes shadow the last column, never store
  const int nlay = p.nlay, nlev = nlay + 1;
  const size_t ncol = p.ncol, ncl = ncol * nlay, nclp = ncol * nlev;
  const RegOrient o{nlay, p.top_at_1};
  const Float pi = reg_pi();
  const Float tau_thresh = sqrt(sqrt((Float)RB_EPS));  // :636
  const int k0 = j * CL;                                // first layer (from the top) of this lane's chunk
  const int gb = blockIdx.y * p.gpt_per_block, ge = min(p.ngpt, gb + p.gpt_per_block);
  // in-plane offsets are 32-bit (ncol*(nlay+1) < 2^31); plane bases are 64-bit
  const int lay_step = p.top_at_1 ? p.ncol : -p.ncol;
  const int off_lay0 = col + p.ncol * o.lay(min(k0, nlay - 1));  // layer k0
  const int off_lev0 = col + p.ncol * o.lev(min(k0, nlay));      // level k0

  const int cta_col0 = blockIdx.x * (kRegThreads / 32) * kRegCols;  // first column of this CTA (= tile origin)
  const int cw = warp * kRegCols + c;                                // this lane's column inside the tile
  auto prefetch = [&](int g, int s) {
    if (TMA) {
      if (threadIdx.x == 0) {
        mbar_expect_tx(&full_bar[s], (uint32_t)((3 * tile_rows + 1) * kTmaCols * sizeof(Float)));
        unsigned char* dst = smem_raw + (size_t)s * stageb;
        tma_load_tile(dst, &tm.tau, &full_bar[s], cta_col0, row0, g);
        tma_load_tile(dst + tb_lay, &tm.lay, &full_bar[s], cta_col0, row0, g);
        tma_load_tile(dst + 2 * tb_lay, &tm.lev, &full_bar[s], cta_col0, row0, g);
      }
    } else {
      const Float* tau_g = p.tau + ncl * g;
      const Float* lay_g = p.lay_source + ncl * g;
      const Float* lev_g = p.lev_source + nclp * g;
#pragma unroll
      for (int i = 0; i < CL; ++i) {
        const int off = (k0 + i < nlay) ? off_lay0 + lay_step * i : off_lay0;
        cp_async_f(RB_SLOT(sm, NS, s, i), tau_g + off);
        cp_async_f(RB_SLOT(sm, NS, s, CL + i), lay_g + off);
      }
#pragma unroll
      for (int i = 0; i <= CL; ++i) {
        const int off = (k0 + i <= nlay) ? off_lev0 + lay_step * i : off_lev0;
        cp_async_f(RB_SLOT(sm, NS, s, 2 * CL + i), lev_g + off);
      }
    }
    const size_t gi = (size_t)col + ncol * g;
    cp_async_f(RB_SLOT(sm, NS, s, BC0 + 1), p.sfc_emis + gi);
    cp_async_f(RB_SLOT(sm, NS, s, BC0 + 2), p.sfc_src + gi);
    cp_async_f(RB_SLOT(sm, NS, s, BC0 + 3), p.inc_flux + gi);
    if (JAC) cp_async_f(RB_SLOT(sm, NS, s, BC0 + 4), p.sfc_srcJac + gi);
    cp_async_f(RB_SLOT(sm, NS, s, BC0 + 5), p.Ds + gi);
  };

  // broadband accumulators: slot i <-> level k0+i+1 (below layer k0+i); *_top <-> level 0 (lane j == 0)
  Float acc_up[BB ? CL : 1], acc_dn[BB ? CL : 1], acc_jac[JAC ? CL : 1];
  Float acc_up_top = 0, acc_dn_top = 0, acc_jac_top = 0;
#pragma unroll
  for (int i = 0; i < CL; ++i) {
    if (BB) { acc_up[BB ? i : 0] = 0; acc_dn[BB ? i : 0] = 0; }
    if (JAC) acc_jac[JAC ? i : 0] = 0;
  }

  if (TMA) {
    if (threadIdx.x == 0) {
#pragma unroll
      for (int i = 0; i < NSTG; ++i) mbar_init(&full_bar[i], 1);
      mbar_fence_init();
    }
    __syncthreads();
  }
  if (gb < ge) prefetch(gb, 0);
  cp_async_commit();
  if (TMA) {  // NSTG g-points in flight; the stage of g is refilled for g+NSTG at the end of iteration g
#pragma unroll
    for (int i = 1; i < NSTG; ++i) {
      if (gb + i < ge) prefetch(gb + i, i);
      cp_async_commit();
    }
  }
  for (int g = gb; g < ge; ++g) {
    const int s = TMA ? (g - gb) % NSTG : (g - gb) & 1;
    if (TMA) {
      cp_async_wait<NSTG - 1>();                                       // boundary values of g (groups g .. g+NSTG-1 outstanding)
      mbar_wait(&full_bar[s], (uint32_t)(((g - gb) / NSTG) & 1));      // tiles of g
    } else {
      if (g + 1 < ge) prefetch(g + 1, s ^ 1);
      cp_async_commit();
      cp_async_wait<1>();  // everything but the newest group (stage s^1) has landed; slots are lane-private
    }
    const Float* tile_tau = reinterpret_cast<const Float*>(smem_raw + (size_t)s * stageb);
    const Float* tile_lev = tile_tau + 2 * te_lay;
    const Float emis = *RB_SLOT(sm, NS, s, BC0 + 1), ssrc = *RB_SLOT(sm, NS, s, BC0 + 2),
                inc = *RB_SLOT(sm, NS, s, BC0 + 3);
    const Float sjac = JAC ? *RB_SLOT(sm, NS, s, BC0 + 4) : (Float)0;
    Float* fup = p.flux_up + nclp * g;
    Float* fdn = p.flux_dn + nclp * g;
    const int nmus = ONEMU ? 1 : p.nmus;
    for (int imu = 0; imu < nmus; ++imu) {
      const Float w = p.weights[imu];
      const Float piw = pi * w;
      const Float D = (imu == 0) ? *RB_SLOT(sm, NS, s, BC0 + 5)
                               : p.Ds[(size_t)col + ncol * ((size_t)g + (size_t)p.ngpt * imu)];
      // ---------------- phase A: CL cells per lane, in registers ----------------
      Float tr[CL], sd[CL], su[CL];
      Float Btop = TMA ? *tile_at(tile_lev, o.lev(FULL ? k0 : min(k0, nlay)) - row0, cw) : *RB_SLOT(sm, NS, s, 2 * CL);
#pragma unroll
      for (int i = 0; i < CL; ++i) {
        const Float Bbot = TMA ? *tile_at(tile_lev, o.lev(FULL ? k0 + i + 1 : min(k0 + i + 1, nlay)) - row0, cw) : *RB_SLOT(sm, NS, s, 2 * CL + i + 1);
        {  // straight-line (see the SW kernel): padding cells become pass-through cells by selects
          const bool live = FULLG || k0 + i < nlay;
          const Float* e_lay = TMA ? tile_at(tile_tau, o.lay(FULL ? k0 + i : min(k0 + i, nlay - 1)) - row0, cw) : nullptr;
          const Float tau_loc = (TMA ? e_lay[0] : *RB_SLOT(sm, NS, s, i)) * D;     // :181
          const Float t = rb_exp(-tau_loc);                                           // :182
          // :652-656, both branches evaluated (the divisor is clamped where the series is selected anyway)
          const Float fact_big = rb_div((Float)1 - t, fmax(tau_loc, tau_thresh)) - t;
          const Float fact_small = tau_loc * ((Float)0.5 + tau_loc * (-(Float)1 / (Float)3 + tau_loc * (Float)1 / (Float)8));
          const Float fact = (tau_loc > tau_thresh) ? fact_big : fact_small;
          const Float lay = TMA ? e_lay[te_lay] : *RB_SLOT(sm, NS, s, CL + i);
          // :660-663; source_dn uses the Planck source at the layer's BOTTOM level, source_up at its TOP
          // level in either orientation (:638-644)
          const Float sdn = ((Float)1 - t) * Bbot + (Float)2 * fact * (lay - Bbot);
          const Float sup = ((Float)1 - t) * Btop + (Float)2 * fact * (lay - Btop);
          // FULL = 2: a zero-filled padding row (tau = 0, zero sources) gives t = 1 and sdn = sup = 0 EXACTLY
          // (exp(-0) = 1, the series branch of fact is 0): no select needed
          constexpr bool SEL = FULL == 0;
          sd[i] = (!SEL || live) ? sdn : (Float)0;
          su[i] = (!SEL || live) ? sup : (Float)0;
          tr[i] = (!SEL || live) ? t : (Float)1;
        }
        Btop = Bbot;
      }
      // g-point flux of one level (only when spectrally resolved output is requested)
      auto store = [&](Float* gflux, int klev, Float I) {
        if ((!FULLG && klev > nlay) || !col_ok) return;
        Float* q = gflux + (size_t)col + ncol * o.lev(klev);
        *q = (imu == 0) ? piw * I : *q + piw * I;                                  // :223-224, :356-357
      };
      // ---------------- phase B1: downward transport, :697-706 (chunk-level scan) ----------------
      const Float I_top = inc / (pi * w);                                           // :144
      if (j == 0) {
        if (BB) acc_dn_top += w * I_top; else store(fdn, 0, I_top);
      }
      Float I;
      {
        Float A = 1, B = 0;
#pragma unroll
        for (int i = 0; i < CL; ++i) { A = tr[i] * A; B = tr[i] * B + sd[i]; }      // composed map of the chunk
        Float out;
        I = affine_handoff_down<kRegChunks>(j, A, B, I_top, out);
#pragma unroll
        for (int i = 0; i < CL; ++i) {                                              // replay: the reference's recurrence
          I = tr[i] * I + sd[i];
          if (BB) acc_dn[BB ? i : 0] += w * I;                                       // :218 (scaled by pi at the end)
          else store(fdn, k0 + i + 1, I);
        }
      }
      // surface: the lane of the last chunk holds the intensity at the surface (:198-202)
      Float Iu = I * ((Float)1 - emis) + emis * ssrc;
      Float Ij = emis * sjac;
      // ---------------- phase B2: upward transport, :729-743 (chunk-level scan) ----------------
      {
        Float A = 1, B = 0;
#pragma unroll
        for (int i = CL - 1; i >= 0; --i) { A = tr[i] * A; B = tr[i] * B + su[i]; }
        Float out, outj;
        Iu = affine_handoff_up<kRegChunks>(j, A, B, Iu, out);
        if (JAC) Ij = affine_handoff_up<kRegChunks>(j, A, (Float)0, Ij, outj);
#pragma unroll
        for (int i = CL - 1; i >= 0; --i) {
          // the incoming value sits at the level below layer k0+i: record it, then cross the layer
          // (accumulator slots of padding cells are never written out, so the sums need no guard: branch-free)
          if (BB) acc_up[BB ? i : 0] += w * Iu;
          else if (FULLG || k0 + i < nlay) store(fup, k0 + i + 1, Iu);
          if (JAC) acc_jac[JAC ? i : 0] += w * Ij;
          Iu = tr[i] * Iu + su[i];
          if (JAC) Ij = tr[i] * Ij;
        }
      }
      if (j == 0) {
        if (BB) acc_up_top += w * Iu; else store(fup, 0, Iu);
        if (JAC) acc_jac_top += w * Ij;
      }
    }
    if (TMA) {  // every warp of the CTA is done with stage s (all angles): hand it back to the TMA for g+2
      __syncthreads();
      if (g + NSTG < ge) prefetch(g + NSTG, s);
      cp_async_commit();
    }
  }
  // ---------------- epilogue: spectrally integrated outputs (:233-238) ----------------
  if (col_ok && (BB || JAC)) {
    const size_t goff = p.group_stride * blockIdx.y;  // 0 unless the express path splits a launch's g-points
    Float *bu = p.bb_up + goff, *bd = p.bb_dn + goff, *bj = p.flux_upJac + goff;
    auto put = [&](Float* dst, size_t o2, Float v) { dst[o2] = p.accumulate ? dst[o2] + v : v; };
#pragma unroll
    for (int i = 0; i < CL; ++i) {
      const int klev = k0 + i + 1;
      if (FULLG || klev <= nlay) {
        const size_t o2 = (size_t)col + ncol * o.lev(klev);
        if (BB) { put(bu, o2, pi * acc_up[BB ? i : 0]); put(bd, o2, pi * acc_dn[BB ? i : 0]); }
        if (JAC) put(bj, o2, pi * acc_jac[JAC ? i : 0]);
      }
    }
    if (j == 0) {
      const size_t o2 = (size_t)col + ncol * o.lev(0);
      if (BB) { put(bu, o2, pi * acc_up_top); put(bd, o2, pi * acc_dn_top); }
      if (JAC) put(bj, o2, pi * acc_jac_top);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// LW no-scattering with Tang rescaling (mo_rte_solver_kernels.F90:148-178, 206-210, 753-844): the default LW path for
// scattering (2-stream) optical properties when use_2stream is false (mo_rte_lw.F90:395-423).  TMA tiles only (five
// planes: tau, ssa, g, lay_source, lev_source; zero-filled padded tiles as in FULL = 2 above).
// Three sweeps, each an affine chain solved as a chunk-level scan:
//   1. plain downward transport                     I_dn1(l+1) = t*I_dn1(l) + S_dn                              (:697-706)
//   2. upward transport with the adjustment term    I_up(top)  = t*I_up(bot) + S_up + Cn*(An*I_dn1(top) - t*S_dn - S_up)
//   3. second downward transport                    I_dn(bot)  = t*I_dn(top) + S_dn + Cn*(An*I_up(X)   - t*S_up - S_dn)
//      with X = the layer's TOP level when top_at_1 but its BOTTOM level otherwise (:801-804 vs :835-838: the reference
//      is orientation-asymmetric as written; replicated per branch, not symmetrised).
// The offsets of sweeps 2 and 3 are known once the previous sweep's radiances are known at every level, so the chain
// stays affine.  Only the upward Jacobian is propagated (:791-792, 824-825).
// ---------------------------------------------------------------------------------------------------
struct LwResclTmaMaps { CUtensorMap tau, ssa, g, lay, lev; };
__host__ __device__ inline size_t lw_rescl_reg_tma_smem(int rows, int nthreads) {
  return 2 * (4 * tile_bytes(rows) + tile_bytes(rows + 1)) + (size_t)(2 * 5) * nthreads * sizeof(Float) + 2 * sizeof(uint64_t);
}

template <int CL, bool BB, bool JAC, int NCH = 8>
__global__ void __launch_bounds__(reg_threads(NCH), NCH == 8 ? 2 : 1) lw_rescl_reg_kernel(const LwNoscatRegParams p,
                                                                                         const __grid_constant__ LwResclTmaMaps tm) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int kRegChunks = NCH, kRegCols = 32 / NCH, kRegThreads = reg_threads(NCH);
  const int row0 = p.row0, tile_rows = p.tile_rows;
  const size_t tb_lay = tile_bytes(tile_rows), tb_lev = tile_bytes(tile_rows + 1);
  const size_t stageb = 4 * tb_lay + tb_lev;
  const int te_lay = (int)(tb_lay / sizeof(Float));
  Float* sm = reinterpret_cast<Float*>(smem_raw + 2 * stageb);  // lane-private cp.async slots: emis, sfc_src, inc_flux, jac, D
  constexpr int NS = 5;
  constexpr int BC0 = -1;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sm + (size_t)2 * NS * kRegThreads);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = lane / kRegChunks, j = lane % kRegChunks;
  const int col_raw = (blockIdx.x * (kRegThreads / 32) + warp) * kRegCols + c;
  const bool col_ok = col_raw < p.ncol;
  const int col = col_ok ? col_raw : p.ncol - 1;
  const int nlay = p.nlay, nlev = nlay + 1;
  const size_t ncol = p.ncol, nclp = ncol * nlev;
  const RegOrient o{nlay, p.top_at_1};
  const Float pi = reg_pi();
  const Float tau_thresh = sqrt(sqrt((Float)RB_EPS));  // :636
  const int k0 = j * CL;
  const int gb = blockIdx.y * p.gpt_per_block, ge = min(p.ngpt, gb + p.gpt_per_block);
  const int cta_col0 = blockIdx.x * (kRegThreads / 32) * kRegCols;
  const int cw = warp * kRegCols + c;
  auto prefetch = [&](int g, int s) {
    if (threadIdx.x == 0) {
      mbar_expect_tx(&full_bar[s], (uint32_t)((5 * tile_rows + 1) * kTmaCols * sizeof(Float)));
      unsigned char* dst = smem_raw + (size_t)s * stageb;
      tma_load_tile(dst, &tm.tau, &full_bar[s], cta_col0, row0, g);
      tma_load_tile(dst + tb_lay, &tm.ssa, &full_bar[s], cta_col0, row0, g);
      tma_load_tile(dst + 2 * tb_lay, &tm.g, &full_bar[s], cta_col0, row0, g);
      tma_load_tile(dst + 3 * tb_lay, &tm.lay, &full_bar[s], cta_col0, row0, g);
      tma_load_tile(dst + 4 * tb_lay, &tm.lev, &full_bar[s], cta_col0, row0, g);
    }
    const size_t gi = (size_t)col + ncol * g;
    cp_async_f(RB_SLOT(sm, NS, s, BC0 + 1), p.sfc_emis + gi);
    cp_async_f(RB_SLOT(sm, NS, s, BC0 + 2), p.sfc_src + gi);
    cp_async_f(RB_SLOT(sm, NS, s, BC0 + 3), p.inc_flux + gi);
    if (JAC) cp_async_f(RB_SLOT(sm, NS, s, BC0 + 4), p.sfc_srcJac + gi);
    cp_async_f(RB_SLOT(sm, NS, s, BC0 + 5), p.Ds + gi);
  };
  Float acc_up[BB ? CL : 1], acc_dn[BB ? CL : 1], acc_jac[JAC ? CL : 1];
  Float acc_up_top = 0, acc_dn_top = 0, acc_jac_top = 0;
#pragma unroll
  for (int i = 0; i < CL; ++i) {
    if (BB) { acc_up[BB ? i : 0] = 0; acc_dn[BB ? i : 0] = 0; }
    if (JAC) acc_jac[JAC ? i : 0] = 0;
  }
  if (threadIdx.x == 0) {
    mbar_init(&full_bar[0], 1);
    mbar_init(&full_bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (gb < ge) prefetch(gb, 0);
  cp_async_commit();
  if (gb + 1 < ge) prefetch(gb + 1, 1);
  cp_async_commit();
  for (int g = gb; g < ge; ++g) {
    const int s = (g - gb) & 1;
    cp_async_wait<1>();
    mbar_wait(&full_bar[s], (uint32_t)(((g - gb) >> 1) & 1));
    const Float* tile_tau = reinterpret_cast<const Float*>(smem_raw + (size_t)s * stageb);
    const Float* tile_lev = tile_tau + 4 * te_lay;
    const Float emis = *RB_SLOT(sm, NS, s, BC0 + 1), ssrc = *RB_SLOT(sm, NS, s, BC0 + 2), inc = *RB_SLOT(sm, NS, s, BC0 + 3);
    const Float sjac = JAC ? *RB_SLOT(sm, NS, s, BC0 + 4) : (Float)0;
    Float* fup = p.flux_up + nclp * g;
    Float* fdn = p.flux_dn + nclp * g;
    for (int imu = 0; imu < p.nmus; ++imu) {
      const Float w = p.weights[imu];
      const Float piw = pi * w;
      const Float D = (imu == 0) ? *RB_SLOT(sm, NS, s, BC0 + 5) : p.Ds[(size_t)col + ncol * ((size_t)g + (size_t)p.ngpt * imu)];
      // ---------------- phase A ----------------
      Float tr[CL], sd[CL], su[CL], an[CL], cn[CL];
      Float Btop = *tile_at(tile_lev, o.lev(k0) - row0, cw);
#pragma unroll
      for (int i = 0; i < CL; ++i) {
        const Float Bbot = *tile_at(tile_lev, o.lev(k0 + i + 1) - row0, cw);
        const bool live = k0 + i < nlay;
        const Float* e_lay = tile_at(tile_tau, o.lay(k0 + i) - row0, cw);
        const Float ssal = e_lay[te_lay], gg = e_lay[2 * te_lay];
        const Float wb = ssal * ((Float)1 - gg) * (Float)0.5;                      // :161
        const Float scaleTau = ((Float)1 - ssal + wb);                             // :165
        const Float cnv = rb_div((Float)0.4 * wb, live ? scaleTau : (Float)1);     // :169
        const Float tau_loc = e_lay[0] * D * scaleTau;                             // :173
        const Float t = rb_exp(-tau_loc);                                          // :175
        const Float fact_big = rb_div((Float)1 - t, fmax(tau_loc, tau_thresh)) - t;
        const Float fact_small = tau_loc * ((Float)0.5 + tau_loc * (-(Float)1 / (Float)3 + tau_loc * (Float)1 / (Float)8));
        const Float fact = (tau_loc > tau_thresh) ? fact_big : fact_small;         // :652-656
        const Float lay = e_lay[3 * te_lay];
        const Float sdn = ((Float)1 - t) * Bbot + (Float)2 * fact * (lay - Bbot);  // :660-663
        const Float sup = ((Float)1 - t) * Btop + (Float)2 * fact * (lay - Btop);
        sd[i] = live ? sdn : (Float)0;
        su[i] = live ? sup : (Float)0;
        tr[i] = live ? t : (Float)1;
        an[i] = live ? ((Float)1 - t * t) : (Float)0;                              // :176
        cn[i] = live ? cnv : (Float)0;
        Btop = Bbot;
      }
      auto store = [&](Float* gflux, int klev, Float I) {
        if (klev > nlay || !col_ok) return;
        Float* q = gflux + (size_t)col + ncol * o.lev(klev);
        *q = (imu == 0) ? piw * I : *q + piw * I;
      };
      // ---------------- sweep 1: plain downward transport; radiance kept at every level of the chunk ----------------
      const Float I_top = inc / (pi * w);                                           // :144
      Float dn1[CL + 1];   // dn1[i] = first-pass radiance ABOVE layer k0+i, dn1[CL] below the chunk's last layer
      {
        Float A = 1, B = 0;
#pragma unroll
        for (int i = 0; i < CL; ++i) { A = tr[i] * A; B = tr[i] * B + sd[i]; }
        Float out;
        Float I = affine_handoff_down<kRegChunks>(j, A, B, I_top, out);
        dn1[0] = I;
#pragma unroll
        for (int i = 0; i < CL; ++i) { I = tr[i] * I + sd[i]; dn1[i + 1] = I; }
      }
      // surface (:198-202), on the lane of the last chunk
      Float Iu = dn1[CL] * ((Float)1 - emis) + emis * ssrc;
      Float Ij = emis * sjac;
      // ---------------- sweep 2: upward with adjustment (:786-792 / :819-825) ----------------
      Float up[CL + 1];    // up[i] = radiance ABOVE layer k0+i (after crossing it), up[CL] = entering the chunk from below
      {
        Float su2[CL];
#pragma unroll
        for (int i = 0; i < CL; ++i) su2[i] = su[i] + cn[i] * (an[i] * dn1[i] - tr[i] * sd[i] - su[i]);
        Float A = 1, B = 0;
#pragma unroll
        for (int i = CL - 1; i >= 0; --i) { A = tr[i] * A; B = tr[i] * B + su2[i]; }
        Float out, outj;
        Iu = affine_handoff_up<kRegChunks>(j, A, B, Iu, out);
        if (JAC) Ij = affine_handoff_up<kRegChunks>(j, A, (Float)0, Ij, outj);
        up[CL] = Iu;
#pragma unroll
        for (int i = CL - 1; i >= 0; --i) {
          if (BB) acc_up[BB ? i : 0] += w * Iu;
          else if (k0 + i < nlay) store(fup, k0 + i + 1, Iu);
          if (JAC) acc_jac[JAC ? i : 0] += w * Ij;
          Iu = tr[i] * Iu + su2[i];
          if (JAC) Ij = tr[i] * Ij;
          up[i] = Iu;
        }
      }
      if (j == 0) {
        if (BB) acc_up_top += w * Iu; else store(fup, 0, Iu);
        if (JAC) acc_jac_top += w * Ij;
      }
      // ---------------- sweep 3: second downward transport with adjustment (:801-804 / :835-838) ----------------
      if (j == 0) {
        if (BB) acc_dn_top += w * I_top; else store(fdn, 0, I_top);
      }
      {
        Float sd2[CL];
#pragma unroll
        for (int i = 0; i < CL; ++i) {
          const Float upx = p.top_at_1 ? up[i] : up[i + 1];   // the layer's top level, or (bottom-up columns) its bottom level
          sd2[i] = sd[i] + cn[i] * (an[i] * upx - tr[i] * su[i] - sd[i]);
        }
        Float A = 1, B = 0;
#pragma unroll
        for (int i = 0; i < CL; ++i) { A = tr[i] * A; B = tr[i] * B + sd2[i]; }
        Float out;
        Float I = affine_handoff_down<kRegChunks>(j, A, B, I_top, out);
#pragma unroll
        for (int i = 0; i < CL; ++i) {
          I = tr[i] * I + sd2[i];
          if (BB) acc_dn[BB ? i : 0] += w * I;
          else store(fdn, k0 + i + 1, I);
        }
      }
    }
    __syncthreads();
    if (g + 2 < ge) prefetch(g + 2, s);
    cp_async_commit();
  }
  if (col_ok && (BB || JAC)) {
#pragma unroll
    for (int i = 0; i < CL; ++i) {
      const int klev = k0 + i + 1;
      if (klev <= nlay) {
        const size_t o2 = (size_t)col + ncol * o.lev(klev);
        if (BB) { p.bb_up[o2] = pi * acc_up[BB ? i : 0]; p.bb_dn[o2] = pi * acc_dn[BB ? i : 0]; }
        if (JAC) p.flux_upJac[o2] = pi * acc_jac[JAC ? i : 0];
      }
    }
    if (j == 0) {
      const size_t o2 = (size_t)col + ncol * o.lev(0);
      if (BB) { p.bb_up[o2] = pi * acc_up_top; p.bb_dn[o2] = pi * acc_dn_top; }
      if (JAC) p.flux_upJac[o2] = pi * acc_jac_top;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// adding (Shonk & Hogan 2008), mo_rte_solver_kernels.F90:1135-1245, on register chunks.
// In:  R[i] = Rdif, T[i] = Tdif, SU[i] = src_up, SD[i] = src_dn of layer k0+i (from the top).
// The upward sweep overwrites them with what the downward sweep needs:
//       T[i] <- a = Tdif*denom       SD[i] <- b = (Rdif*src_below + src_dn)*denom
//       R[i] <- albedo below layer   SU[i] <- source below layer
// top(fup, fdn): fluxes at the top level (lane j == 0); lev(i, fup, fdn): at the level below layer k0+i,
// called with a compile-time-constant i so that callers can index register arrays with it.
// ---------------------------------------------------------------------------------------------------
template <int CL, int NCH, typename Top, typename Lev>
__device__ __forceinline__ void adding_reg(int j, Float (&R)[CL], Float (&T)[CL], Float (&SU)[CL], Float (&SD)[CL],
                                           Float albedo_sfc, Float src_sfc, Float flux_dn_top, Top top, Lev lev) {
  // ---- upward pass (:1166-1186) as a chunk-level scan.  One layer maps the (albedo, source) below it to the
  // pair above it projectively: with v = (alb, src, 1)^T,
  //        [ t^2 - r^2        0   r   ]
  //   v' ~ [ t*sdn - sup*r    t   sup ] v ,   divided by its third component (1 - r*alb),
  //        [ -r               0   1   ]
  // and products of such matrices keep the zero pattern, so a chunk composes into 7 numbers.
  Float ma = 1, mb = 0, mc = 0, md = 1, me = 0, mf = 0, mg = 1;
#pragma unroll
  for (int i = CL - 1; i >= 0; --i) {
    const Float r = R[i], t = T[i], sup = SU[i], sdn = SD[i];
    const Float xa = t * t - r * r, xc = t * sdn - sup * r;
    const Float na = xa * ma + r * mf, nb = xa * mb + r * mg;
    const Float nc = xc * ma + t * mc + sup * mf, ne = xc * mb + t * me + sup * mg;
    const Float nf = mf - r * ma, ng = mg - r * mb;
    ma = na; mb = nb; mc = nc; md = t * md; me = ne; mf = nf; mg = ng;
  }
  // State entering this lane's chunk from below (:1166-1168 for the last chunk).  The hand-over carries the pair in
  // HOMOGENEOUS form (a, s, d) with albedo = a/d, source = s/d: applying a chunk's matrix is then 7 multiply-adds with
  // no division inside the 8-step serial chain (the reciprocal used to be its longest link); every lane normalises
  // the state it received once, in parallel with the others.
  Float alb = albedo_sfc, src = src_sfc;
#if RB_HANDOFF_SCAN
  {
    // prefix scan of the chunk matrices from the bottom (3 steps; see affine_handoff_down): after it lane j holds the
    // product of the matrices of chunks j .. 7, bottom one applied first
#pragma unroll
    for (int d = 1; d < NCH; d <<= 1) {
      Float la = __shfl_down_sync(0xffffffffu, ma, d, NCH), lb = __shfl_down_sync(0xffffffffu, mb, d, NCH);
      Float lc = __shfl_down_sync(0xffffffffu, mc, d, NCH), ld = __shfl_down_sync(0xffffffffu, md, d, NCH);
      Float le = __shfl_down_sync(0xffffffffu, me, d, NCH), lf = __shfl_down_sync(0xffffffffu, mf, d, NCH);
      Float lg = __shfl_down_sync(0xffffffffu, mg, d, NCH);
#if RB_SCAN_SELECT
      const bool on = j + d < NCH;  // lanes without a partner combine with the identity matrix
      la = on ? la : (Float)1; lb = on ? lb : (Float)0; lc = on ? lc : (Float)0; ld = on ? ld : (Float)1;
      le = on ? le : (Float)0; lf = on ? lf : (Float)0; lg = on ? lg : (Float)1;
      {
#else
      if (j + d < NCH) {  // (this lane's chunks) after (the d chunk groups below them)
#endif
        const Float na = ma * la + mb * lf, nb = ma * lb + mb * lg;
        const Float nc = mc * la + md * lc + me * lf, ne = mc * lb + md * le + me * lg;
        const Float nf = mf * la + mg * lf, ng = mf * lb + mg * lg;
        ma = na; mb = nb; mc = nc; md = md * ld; me = ne; mf = nf; mg = ng;
      }
    }
    // the surface state lives on the lane of the last chunk (src_sfc comes from its direct beam)
    const Float a_s = __shfl_sync(0xffffffffu, albedo_sfc, NCH - 1, NCH);
    const Float s_s = __shfl_sync(0xffffffffu, src_sfc, NCH - 1, NCH);
    // state leaving this lane's chunk upwards, homogeneous; the lane above takes it over
    const Float ha = ma * a_s + mb, hs = mc * a_s + md * s_s + me, hd = mf * a_s + mg;
    const Float ia = __shfl_down_sync(0xffffffffu, ha, 1, NCH);
    const Float is = __shfl_down_sync(0xffffffffu, hs, 1, NCH);
    const Float id = __shfl_down_sync(0xffffffffu, hd, 1, NCH);
    if (j < NCH - 1) {
      const Float inv = rb_rcp(id);
      alb = ia * inv;
      src = is * inv;
    } else {
      alb = a_s;  // exactly the surface values (:1166-1168)
      src = s_s;
    }
  }
#else
  {
    Float ha = 0, hs = 0, hd = 0;
    Float ia = albedo_sfc, is = src_sfc, id = 1;  // incoming state of this lane, homogeneous
    for (int jj = NCH - 1; jj >= 0; --jj) {
      const Float a_b = __shfl_down_sync(0xffffffffu, ha, 1);
      const Float s_b = __shfl_down_sync(0xffffffffu, hs, 1);
      const Float d_b = __shfl_down_sync(0xffffffffu, hd, 1);
      if (j == jj) {
        if (jj < NCH - 1) { ia = a_b; is = s_b; id = d_b; }
        ha = ma * ia + mb * id;
        hs = mc * ia + md * is + me * id;
        hd = mf * ia + mg * id;
      }
    }
    const Float inv = rb_rcp(id);  // id = 1 for the last chunk: alb, src are then exactly the surface values
    alb = ia * inv;
    src = is * inv;
  }
#endif
#if RB_ADD_HOMOG
  // Upward replay in HOMOGENEOUS form.  With alb = A/D, src = S/D the layer step (:1174-1186) reads
  //     D' = D - r*A            (= D*(1 - r*alb))
  //     A' = r*D' + t^2*A       (albedo above  = r + t^2*alb/(1 - r*alb))
  //     S' = sup*D' + t*(S + A*sdn)
  // - two dependent multiply-adds per layer and no reciprocal inside the chain (the per-layer form's chain carries a
  // reciprocal, ~5 dependent operations, per layer: the longest serial stretch of the kernel, during which a warp
  // issues next to nothing).  The reciprocals 1/D of the CL+1 levels are then independent of one another, and
  // denom = 1/(1 - r*alb) = D/D' needs no reciprocal of its own.  Per-level albedo and source carry the rounding of
  // CL <= 10 homogeneous steps (a few ulp) instead of the reference's level-by-level quotient.
  {
    Float hA[CL + 1], hS[CL + 1], hD[CL + 1];  // state BELOW layer i (slot i), slot CL... see below: slot k = below layer k-1
    // slot CL = below the lane's last layer (incoming), slot i = above layer i = below layer i-1
    hA[CL] = alb; hS[CL] = src; hD[CL] = (Float)1;
#pragma unroll
    for (int i = CL - 1; i >= 0; --i) {
      const Float r = R[i], t = T[i];
      hD[i] = hD[i + 1] - r * hA[i + 1];
      hA[i] = r * hD[i] + (t * t) * hA[i + 1];
      hS[i] = SU[i] * hD[i] + t * (hS[i + 1] + hA[i + 1] * SD[i]);
    }
    Float inv[CL + 1];
    inv[CL] = (Float)1;
#pragma unroll
    for (int i = 0; i < CL; ++i) inv[i] = rb_rcp(hD[i]);
#pragma unroll
    for (int i = CL - 1; i >= 0; --i) {
      const Float r = R[i], t = T[i], sdn = SD[i];
      const Float alb_i = hA[i + 1] * inv[i + 1], src_i = hS[i + 1] * inv[i + 1];  // below layer i
      const Float denom = hD[i + 1] * inv[i];                                        // 1/(1 - r*alb_i)
      R[i] = alb_i;
      SU[i] = src_i;
      T[i] = t * denom;
      SD[i] = (r * src_i + sdn) * denom;
    }
    alb = hA[0] * inv[0];
    src = hS[0] * inv[0];
  }
#else
  // replay with the reference's per-layer expressions, every lane on its own chunk
#pragma unroll
  for (int i = CL - 1; i >= 0; --i) {  // :1174-1186
    const Float r = R[i], t = T[i], sup = SU[i], sdn = SD[i];
    const Float denom = rb_rcp((Float)1 - r * alb);
    const Float a = t * denom;
    R[i] = alb;
    SU[i] = src;
    T[i] = a;
    SD[i] = (r * src + sdn) * denom;
    const Float albn = r + t * t * alb * denom;
    src = sup + a * (src + alb * sdn);
    alb = albn;
  }
#endif
  // lane j == 0 now holds albedo and source at the top of the domain
  if (j == 0) top(flux_dn_top * alb + src, flux_dn_top);  // :1190
  // ---- downward pass (:1196-1202): fdn' = a*fdn + b per layer, an affine chain
  Float A = 1, B = 0;
#pragma unroll
  for (int i = 0; i < CL; ++i) { A = T[i] * A; B = T[i] * B + SD[i]; }
  Float out;
  Float fdn = affine_handoff_down<NCH>(j, A, B, flux_dn_top, out);
#pragma unroll
  for (int i = 0; i < CL; ++i) {
    fdn = T[i] * fdn + SD[i];
    lev(i, fdn * R[i] + SU[i], fdn);
  }
}

// ---------------------------------------------------------------------------------------------------
// One SW two-stream cell (mo_rte_solver_kernels.F90:1027-1108): layer reflectance / transmittance for diffuse and direct
// light.  Shared by the register kernel below and the warp-specialised kernel (kernels/solver_ws.cuh), so that both run
// the same arithmetic.  mu0_s = max(sqrt(eps), mu0), mu0_3 = 3*mu0_s and r_mu0 = rb_rcp1(mu0_s) depend on (column, layer)
// only; callers that keep a cell across g-points hoist them.
// merged: RT_term = 1/rt_den (:1052) and RT_term*w0/om_s (:1071) from ONE reciprocal, 1/(rt_den*om_s): a multiplication
// each instead of a second division sequence (9 fp64 instructions); rt_den in [k, 2*max(k, gamma1)] and |om_s| >= eps
// keep the product finite and normal; both quotients stay within 2 ulp of the reference's.  Not merged: a zero cell
// (tau = ssa = g = 0) gives Tdif = rcp(4)*2*2*exp(-0) = 1 EXACTLY, which the zero-filled padded tiles rely on.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sw_two_stream_cell(Float tau_s, Float w0_s, Float g_s, Float mu0_s, Float mu0_3, Float r_mu0,
                                                   bool merged, Float& Rdif, Float& Tdif, Float& Rdir, Float& Tdir,
                                                   Float& Tnoscat) {
  const Float eps = (Float)RB_EPS;
  const Float min_k = (Float)1.e4 * eps;  // :1005
  const Float gamma1 = ((Float)8 - w0_s * ((Float)5 + (Float)3 * g_s)) * (Float).25;
  const Float gamma2 = (Float)3 * (w0_s * ((Float)1 - g_s)) * (Float).25;
  const Float kk = rb_sqrt(fmax((gamma1 - gamma2) * (gamma1 + gamma2), min_k));
  const Float exp_minusktau = rb_exp(-tau_s * kk);
  const Float exp_minus2ktau = exp_minusktau * exp_minusktau;
  const Float k_mu = kk * mu0_s;
  const Float om = (Float)1 - k_mu * k_mu;
  const Float om_s = fabs(om) >= eps ? om : eps;                                  // :1071-1073
  const Float rt_den = kk * ((Float)1 + exp_minus2ktau) + gamma1 * ((Float)1 - exp_minus2ktau);
  const Float rt_inv = rb_rcp(merged ? rt_den * om_s : rt_den);
  Float RT_term = merged ? om_s * rt_inv : rt_inv;
  Rdif = RT_term * gamma2 * ((Float)1 - exp_minus2ktau);
  Tdif = RT_term * (Float)2 * kk * exp_minusktau;
  RT_term = merged ? w0_s * rt_inv : rb_div(w0_s * rt_inv, om_s);
  const Float gamma3 = ((Float)2 - mu0_3 * g_s) * (Float).25;
  const Float gamma4 = (Float)1 - gamma3;
  const Float alpha1 = gamma1 * gamma4 + gamma2 * gamma3;
  const Float alpha2 = gamma1 * gamma3 + gamma2 * gamma4;
  const Float k_gamma3 = kk * gamma3;
  const Float k_gamma4 = kk * gamma4;
  Tnoscat = rb_exp<true>(-rb_div_r(tau_s, mu0_s, r_mu0));  // flushes to 0 (night columns, mu0_s = sqrt(eps))
  Rdir = RT_term * (((Float)1 - k_mu) * (alpha2 + k_gamma3) -
                    ((Float)1 + k_mu) * (alpha2 - k_gamma3) * exp_minus2ktau -
                    (Float)2.0 * (k_gamma3 - alpha2 * k_mu) * exp_minusktau * Tnoscat);
  Tdir = -RT_term * (((Float)1 + k_mu) * (alpha1 + k_gamma4) * Tnoscat -
                     ((Float)1 - k_mu) * (alpha1 - k_gamma4) * exp_minus2ktau * Tnoscat -
                     (Float)2.0 * (k_gamma4 + alpha1 * k_mu) * exp_minusktau);
  Rdir = fmax((Float)0, fmin(Rdir, ((Float)1 - Tnoscat)));         // :1107
  Tdir = fmax((Float)0, fmin(Tdir, ((Float)1 - Tnoscat - Rdir)));  // :1108
}

// ---------------------------------------------------------------------------------------------------
// SW two-stream (mo_rte_solver_kernels.F90:503-609, 985-1127)
// ---------------------------------------------------------------------------------------------------
struct SwRegParams {
  int ncol, nlay, ngpt, top_at_1;
  const Float *tau, *ssa, *g, *mu0, *sfc_alb_dir, *sfc_alb_dif, *inc_flux_dir;
  Float *flux_up, *flux_dn, *flux_dir;
  int has_dif_bc;
  const Float* inc_flux_dif;
  int do_broadband;
  Float *bb_up, *bb_dn, *bb_dir;
  int gpt_per_block;
  int accumulate;        // express path, see LwNoscatRegParams
  size_t group_stride;
  int tile_rows, row0;   // zero-filled padded tiles, see LwNoscatRegParams
};

template <int CL>
__host__ __device__ constexpr int sw_reg_slots() { return 3 * CL + 4; }  // tau, ssa, g, alb_dir, alb_dif, inc_dir, inc_dif

// LEAN (broadband only): the per-level broadband accumulators live in lane-private shared-memory slots instead
// of registers and the input prefetch is single-stage (issued when phase A has consumed the slots), so that the
// kernel fits 168 registers and 68 KB of shared memory: 3 resident CTAs (12 warps) per SM instead of 2.
template <int CL, bool LEAN>
__host__ __device__ constexpr int sw_reg_smem_slots() { return (LEAN ? 1 : 2) * sw_reg_slots<CL>() + CL + (LEAN ? 3 * CL : 0); }

// TMA variant: the three (16 columns x nlay) input tiles of a g-point arrive by cp.async.bulk.tensor (kernels/tma.cuh),
// two stages; the four per-(column, g-point) boundary values keep their lane-private cp.async slots.
struct SwTmaMaps { CUtensorMap tau, ssa, g; };
template <int CL, bool LEAN>
__host__ __device__ inline size_t sw_reg_tma_smem(int nlay, int nthreads = kRegThreads) {
  return 2 * 3 * tile_bytes(nlay) + (size_t)(2 * 4 + CL + (LEAN ? 3 * CL : 0)) * nthreads * sizeof(Float) + 2 * sizeof(uint64_t);
}

// FULL: nlay == 8*CL, every lane's cells are real layers - the padding selects and tests fold away at compile time.
template <int CL, bool BB, int MINB = 3, bool LEAN = false, bool TMA = false, int FULL = 0, int NCH = 8>
__global__ void __launch_bounds__(reg_threads(NCH), NCH == 8 ? MINB : 1) sw_2stream_reg_kernel(const SwRegParams p,
                                                                            const __grid_constant__ SwTmaMaps tm) {
  static_assert(!LEAN || BB, "LEAN is a broadband-only variant");
  // no static shared memory in this kernel: the swizzled TMA tiles need the dynamic window to start 1024-byte aligned
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int kRegChunks = NCH, kRegCols = 32 / NCH, kRegThreads = reg_threads(NCH);  // lanes per column, columns per warp, threads per CTA
  constexpr bool FULLG = FULL == 1;                             // see LwNoscatRegParams::tile_rows
  const int row0 = (FULL == 2) ? p.row0 : 0, tile_rows = (FULL == 2) ? p.tile_rows : p.nlay;
  const size_t tileb = TMA ? tile_bytes(tile_rows) : 0;        // bytes of one tile
  const int tile_elems = (int)(tileb / sizeof(Float));
  const Float* tiles = reinterpret_cast<const Float*>(smem_raw);   // TMA: [stage][plane][row][16]
  Float* sm = reinterpret_cast<Float*>(smem_raw + 2 * 3 * tileb);  // cp.async slots
  constexpr int NS = TMA ? 4 : sw_reg_slots<CL>();                // slots per stage
  constexpr int BC0 = TMA ? 0 : 3 * CL;                           // first boundary-value slot
  constexpr int NSTAGE = (LEAN && !TMA) ? 1 : 2;
  Float* sm_mu0 = sm + (size_t)NSTAGE * NS * kRegThreads;  // [CL][thread], loaded once
  Float* sm_acc = sm_mu0 + (size_t)CL * kRegThreads + threadIdx.x;  // LEAN: [3][CL][thread]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sm_mu0 + (size_t)(CL + (LEAN ? 3 * CL : 0)) * kRegThreads);  // TMA: [2] mbarriers
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = lane / kRegChunks, j = lane % kRegChunks;
  const int col_raw = (blockIdx.x * (kRegThreads / 32) + warp) * kRegCols + c;
  const bool col_ok = col_raw < p.ncol;
  const int col = col_ok ? col_raw : p.ncol - 1;
  const int nlay = p.nlay, nlev = nlay + 1;
  const size_t ncol = p.ncol, ncl = ncol * nlay, nclp = ncol * nlev;
  const RegOrient o{nlay, p.top_at_1};
  const Float eps = (Float)RB_EPS;
  const Float min_k = (Float)1.e4 * eps;  // :1005
  const Float min_mu0 = sqrt(eps);        // :1006
  const int k0 = j * CL;
  const int gb = blockIdx.y * p.gpt_per_block, ge = min(p.ngpt, gb + p.gpt_per_block);
  const int lay_step = p.top_at_1 ? p.ncol : -p.ncol;
  const int off_lay0 = col + p.ncol * o.lay(min(k0, nlay - 1));

  const int cta_col0 = blockIdx.x * (kRegThreads / 32) * kRegCols;  // first column of this CTA (= tile origin)
  const int cw = warp * kRegCols + c;                                // this lane's column inside the tile
  auto prefetch = [&](int g, int s) {
    if (TMA) {
      if (threadIdx.x == 0) {
        mbar_expect_tx(&full_bar[s], (uint32_t)(3 * tile_rows * kTmaCols * sizeof(Float)));
        unsigned char* dst = smem_raw + (size_t)s * 3 * tileb;
        tma_load_tile(dst, &tm.tau, &full_bar[s], cta_col0, row0, g);
        tma_load_tile(dst + tileb, &tm.ssa, &full_bar[s], cta_col0, row0, g);
        tma_load_tile(dst + 2 * tileb, &tm.g, &full_bar[s], cta_col0, row0, g);
      }
    } else {
      const Float* tau_g = p.tau + ncl * g;
      const Float* ssa_g = p.ssa + ncl * g;
      const Float* g_g = p.g + ncl * g;
#pragma unroll
      for (int i = 0; i < CL; ++i) {
        const int off = (k0 + i < nlay) ? off_lay0 + lay_step * i : off_lay0;
        cp_async_f(RB_SLOT(sm, NS, s, i), tau_g + off);
        cp_async_f(RB_SLOT(sm, NS, s, CL + i), ssa_g + off);
        cp_async_f(RB_SLOT(sm, NS, s, 2 * CL + i), g_g + off);
      }
    }
    const size_t gi = (size_t)col + ncol * g;
    cp_async_f(RB_SLOT(sm, NS, s, BC0 + 0), p.sfc_alb_dir + gi);
    cp_async_f(RB_SLOT(sm, NS, s, BC0 + 1), p.sfc_alb_dif + gi);
    cp_async_f(RB_SLOT(sm, NS, s, BC0 + 2), p.inc_flux_dir + gi);
    if (p.has_dif_bc) cp_async_f(RB_SLOT(sm, NS, s, BC0 + 3), p.inc_flux_dif + gi);
  };

  constexpr int NACC = (BB && !LEAN) ? CL : 1;
  Float acc_up[NACC], acc_dn[NACC], acc_dir[NACC];
  Float acc_up_top = 0, acc_dn_top = 0, acc_dir_top = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) { acc_up[i] = 0; acc_dn[i] = 0; acc_dir[i] = 0; }
  if (LEAN) {
#pragma unroll
    for (int i = 0; i < 3 * CL; ++i) sm_acc[i * kRegThreads] = 0;
  }
  // broadband accumulation of level slot i (compile-time constant): which = 0 up, 1 dn, 2 dir
  auto acc_add = [&](int which, int i, Float v) {
    if (LEAN) sm_acc[(which * CL + i) * kRegThreads] += v;
    else if (which == 0) acc_up[NACC > 1 ? i : 0] += v;
    else if (which == 1) acc_dn[NACC > 1 ? i : 0] += v;
    else acc_dir[NACC > 1 ? i : 0] += v;
  };

  if (TMA) {
    if (threadIdx.x == 0) {
      mbar_init(&full_bar[0], 1);
      mbar_init(&full_bar[1], 1);
      mbar_fence_init();
    }
    __syncthreads();
  }
  if (gb < ge) prefetch(gb, 0);
  cp_async_commit();
  if (TMA) {  // TMA keeps two g-points in flight: the stage of g is refilled for g+2 once every warp has read it
    if (gb + 1 < ge) prefetch(gb + 1, 1);
    cp_async_commit();
  }
  // (Tried: mu0_s, 3*mu0_s and 1/mu0_s precomputed per (column, layer) in three shared-memory planes - five fp64
  // instructions fewer per cell, but 255 registers and two more LDS per cell: 14.3 -> 16.9 ms on B200.)
  // RB_SW_MU0_HOIST: the slot holds max(sqrt(eps), mu0) (:1065) and the mu0 > 0 tests (:1122-1125) are one bit per cell
  // in a register - both depend on (column, layer) only; per cell and g-point that is one fp64 compare + select pair and
  // one fp64 compare less (same values, bit-identical results)
  unsigned lit_mask = 0;
#pragma unroll
  for (int i = 0; i < CL; ++i) {
    const Float m = p.mu0[(k0 + i < nlay) ? off_lay0 + lay_step * i : off_lay0];
    sm_mu0[i * kRegThreads + threadIdx.x] = RB_SW_MU0_HOIST ? fmax(min_mu0, m) : m;
    if (m > (Float)0) lit_mask |= 1u << i;
  }
  const Float mu0_top = p.mu0[(size_t)col + ncol * o.lay(0)];
  const Float mu0_sfc = p.mu0[(size_t)col + ncol * o.lay(nlay - 1)];

  for (int g = gb; g < ge; ++g) {
    const int s = (LEAN && !TMA) ? 0 : (g - gb) & 1;
    if (TMA) {
      cp_async_wait<1>();                                         // boundary values of g (groups g, g+1 outstanding)
      mbar_wait(&full_bar[s], (uint32_t)(((g - gb) >> 1) & 1));   // tiles of g
    } else if (!LEAN) {
      if (g + 1 < ge) prefetch(g + 1, s ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    const Float* tile_s = tiles + (size_t)s * 3 * tile_elems;
    Float* gup = p.flux_up + nclp * g;
    Float* gdn = p.flux_dn + nclp * g;
    Float* gdir = p.flux_dir + nclp * g;
    // ---------------- phase A: two-stream layer properties (:1027-1108) ----------------
    Float R[CL], T[CL], A3[CL], A4[CL], A5[CL];  // Rdif, Tdif, Rdir->src_up, Tdir->src_dn, Tnoscat->direct flux
    // Straight-line per cell (no branches, no calls): padding cells beyond nlay compute on the clamped inputs the
    // prefetch gave them and are turned into pass-through cells by selects at the end, so that the compiler can
    // interleave the CL independent cells of a lane.
#pragma unroll
    for (int i = 0; i < CL; ++i) {
      constexpr bool NOSEL = RB_PAD_NOSELECT && FULL == 2;
      const bool live = FULLG || NOSEL || k0 + i < nlay;
      Float tau_s, w0_s, g_s;
      if (TMA) {
        const Float* e = tile_at(tile_s, o.lay(FULL ? k0 + i : min(k0 + i, nlay - 1)) - row0, cw);
        tau_s = e[0]; w0_s = e[tile_elems]; g_s = e[2 * tile_elems];
      } else {
        tau_s = *RB_SLOT(sm, NS, s, i); w0_s = *RB_SLOT(sm, NS, s, CL + i); g_s = *RB_SLOT(sm, NS, s, 2 * CL + i);
      }
      const Float mu0 = sm_mu0[i * kRegThreads + threadIdx.x];
      // FULL = 1: shared reciprocal; FULL = 2: separate; FULL = 0 (clamped cp.async inputs, any shape): whichever the
      // TMA instantiation of the same shape uses (uniform at run time), so that results do not depend on the path
      const bool MERGED = RB_SW_MERGED_DIV && ((FULL == 1) ? true : ((FULL == 2) ? !NOSEL : (RB_PAD_NOSELECT ? (NCH == 8 && nlay == kRegChunks * CL) : true)));   // (16-lane TMA launches are all FULL = 2)
      const Float mu0_s = RB_SW_MU0_HOIST ? mu0 : fmax(min_mu0, mu0);
      Float Rdif, Tdif, Rdir, Tdir, Tnoscat;
      sw_two_stream_cell(tau_s, w0_s, g_s, mu0_s, (Float)3 * mu0_s, rb_rcp1(mu0_s), MERGED, Rdif, Tdif, Rdir, Tdir, Tnoscat);
      // FULL = 2: a zero-filled padding row (tau = ssa = g = 0) gives Rdif = 0 (factor 1 - exp(-0)), Rdir = Tdir = 0
      // (factor ssa) and Tnoscat = exp(-0) = 1 EXACTLY; only Tdif = RT_term*2k is 1 to rounding - one select, on T
      constexpr bool SEL = FULL == 0;
      const bool lit = (!SEL || live) && (RB_SW_MU0_HOIST ? ((lit_mask >> i) & 1u) != 0 : (mu0 > (Float)0));  // :1122-1125: no source for diffuse light where mu0 <= 0
      R[i] = (!SEL || live) ? Rdif : (Float)0;
      T[i] = live ? Tdif : (Float)1;
      A3[i] = lit ? Rdir : (Float)0;
      A4[i] = lit ? Tdir : (Float)0;
      A5[i] = (!SEL || live) ? Tnoscat : (Float)1;
    }
    const Float alb_dir = *RB_SLOT(sm, NS, s, BC0 + 0), alb_dif = *RB_SLOT(sm, NS, s, BC0 + 1);
    const Float dir_top_g = *RB_SLOT(sm, NS, s, BC0 + 2) * mu0_top;                  // :575
    const Float dn_top = p.has_dif_bc ? *RB_SLOT(sm, NS, s, BC0 + 3) : (Float)0;     // :579-583
    if (LEAN && !TMA) {  // every slot of the (single) stage has been consumed: refill it while phase B runs
      if (g + 1 < ge) prefetch(g + 1, 0);
      cp_async_commit();
    }
    if (TMA) {  // every warp of the CTA has read stage s into registers: hand it back to the TMA for g+2
      __syncthreads();
      if (g + 2 < ge) prefetch(g + 2, s);
      cp_async_commit();
    }
    // ---------------- phase B1: direct beam and its sources, :1110-1112 ----------------
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
        // (accumulator slots of padding cells are never written out, so the sums need no guard: branch-free)
        if (BB) { acc_add(2, i, dir); acc_add(1, i, dir); }  // :604, direct part of :603
        else if ((FULLG || k0 + i < nlay) && col_ok) gdir[(size_t)col + ncol * o.lev(k0 + i + 1)] = dir;
        A5[i] = dir;  // direct flux below layer k0+i, for the g-point totals (:606)
      }
    }
    // the lane of the last chunk holds the direct flux at the surface (:1120)
    const Float src_sfc = (mu0_sfc > (Float)0) ? dir * alb_dir : (Float)0;
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
      else if ((FULLG || k0 + i < nlay) && col_ok) {
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
      if (FULLG || klev <= nlay) {
        const size_t o2 = (size_t)col + ncol * o.lev(klev);
        if (LEAN) {
          put(bu, o2, sm_acc[i * kRegThreads]); put(bd, o2, sm_acc[(CL + i) * kRegThreads]);
          put(br, o2, sm_acc[(2 * CL + i) * kRegThreads]);
        } else {
          put(bu, o2, acc_up[NACC > 1 ? i : 0]); put(bd, o2, acc_dn[NACC > 1 ? i : 0]); put(br, o2, acc_dir[NACC > 1 ? i : 0]);
        }
      }
    }
    if (j == 0) {
      const size_t o2 = (size_t)col + ncol * o.lev(0);
      put(bu, o2, acc_up_top); put(bd, o2, acc_dn_top); put(br, o2, acc_dir_top);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// LW two-stream (mo_rte_solver_kernels.F90:377-440, 854-967): g-point fluxes only
// ---------------------------------------------------------------------------------------------------
struct Lw2sRegParams {
  int ncol, nlay, ngpt, top_at_1, lev_per_gpt;
  const Float *tau, *ssa, *g, *lay_source, *lev_source, *sfc_emis, *sfc_src, *inc_flux;
  Float *flux_up, *flux_dn;
  int gpt_per_block;
  int tile_rows, row0;   // zero-filled padded tiles, see LwNoscatRegParams
};

// TMA variant: tau, ssa, g, lay_source (nlay rows) and lev_source (nlay+1 rows) tiles per g-point, two stages.
struct Lw2sTmaMaps { CUtensorMap tau, ssa, g, lay, lev; };
__host__ __device__ inline size_t lw_2stream_reg_tma_smem(int nlay) {
  return 2 * (4 * tile_bytes(nlay) + tile_bytes(nlay + 1)) + 2 * sizeof(uint64_t);
}

template <int CL, bool TMA = false, int FULL = 0, int NCH = 8>
__global__ void __launch_bounds__(reg_threads(NCH), NCH == 8 ? (TMA ? 2 : 3) : 1) lw_2stream_reg_kernel(const Lw2sRegParams p,
                                                                                  const __grid_constant__ Lw2sTmaMaps tm) {
  // no static shared memory: the swizzled TMA tiles need the dynamic window to start 1024-byte aligned
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int kRegChunks = NCH, kRegCols = 32 / NCH, kRegThreads = reg_threads(NCH);  // lanes per column, columns per warp, threads per CTA
  constexpr bool FULLG = FULL == 1;    // see LwNoscatRegParams::tile_rows
  const int row0 = (FULL == 2) ? p.row0 : 0, tile_rows = (FULL == 2) ? p.tile_rows : p.nlay;
  const size_t tb_lay = TMA ? tile_bytes(tile_rows) : 0, tb_lev = TMA ? tile_bytes(tile_rows + 1) : 0;
  const size_t stageb = 4 * tb_lay + tb_lev;
  const int te_lay = (int)(tb_lay / sizeof(Float));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + 2 * stageb);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = lane / kRegChunks, j = lane % kRegChunks;
  const int col_raw = (blockIdx.x * (kRegThreads / 32) + warp) * kRegCols + c;
  const bool col_ok = col_raw < p.ncol;
  const size_t col = col_ok ? col_raw : p.ncol - 1;
  const int nlay = p.nlay, nlev = nlay + 1;
  const size_t ncol = p.ncol, ncl = ncol * nlay, nclp = ncol * nlev;
  const RegOrient o{nlay, p.top_at_1};
  const Float pi = reg_pi();
  const Float LW_diff_sec = (Float)1.66f;  // :870 single-precision literal widened to wp
  const int k0 = j * CL;
  const int gb = blockIdx.y * p.gpt_per_block, ge = min(p.ngpt, gb + p.gpt_per_block);
  const int cta_col0 = blockIdx.x * (kRegThreads / 32) * kRegCols;
  const int cw = warp * kRegCols + c;

  auto issue = [&](int g, int s) {  // TMA only
    if (threadIdx.x == 0) {
      mbar_expect_tx(&full_bar[s], (uint32_t)((5 * tile_rows + 1) * kTmaCols * sizeof(Float)));
      unsigned char* dst = smem_raw + (size_t)s * stageb;
      tma_load_tile(dst, &tm.tau, &full_bar[s], cta_col0, row0, g);
      tma_load_tile(dst + tb_lay, &tm.ssa, &full_bar[s], cta_col0, row0, g);
      tma_load_tile(dst + 2 * tb_lay, &tm.g, &full_bar[s], cta_col0, row0, g);
      tma_load_tile(dst + 3 * tb_lay, &tm.lay, &full_bar[s], cta_col0, row0, g);
      tma_load_tile(dst + 4 * tb_lay, &tm.lev, &full_bar[s], cta_col0, row0, p.lev_per_gpt ? g : 0);  // quirk :422
    }
  };
  if (TMA) {
    if (threadIdx.x == 0) {
      mbar_init(&full_bar[0], 1);
      mbar_init(&full_bar[1], 1);
      mbar_fence_init();
    }
    __syncthreads();
    if (gb < ge) issue(gb, 0);
    if (gb + 1 < ge) issue(gb + 1, 1);
  }

  for (int g = gb; g < ge; ++g) {
    const int s = (g - gb) & 1;
    const size_t gi = col + ncol * g;
    const size_t gsrc = p.lev_per_gpt ? g : 0;  // reference default-kernel quirk (:422), see the ABI header
    Float* gup = p.flux_up + nclp * g;
    Float* gdn = p.flux_dn + nclp * g;
    const Float* tile_tau = reinterpret_cast<const Float*>(smem_raw + (size_t)s * stageb);
    const Float* tile_lev = tile_tau + 4 * te_lay;
    if (TMA) mbar_wait(&full_bar[s], (uint32_t)(((g - gb) >> 1) & 1));
    Float R[CL], T[CL], SU[CL], SD[CL], Blev[CL + 1];
#pragma unroll
    for (int i = 0; i <= CL; ++i) {
      const int kk = FULL ? k0 + i : min(k0 + i, nlay);
      Blev[i] = TMA ? *tile_at(tile_lev, o.lev(kk) - row0, cw) : p.lev_source[col + ncol * o.lev(kk) + nclp * gsrc];
    }
    // straight-line per cell (see sw_2stream_reg_kernel): padding cells become pass-through cells by selects
#pragma unroll
    for (int i = 0; i < CL; ++i) {
      const bool live = FULLG || k0 + i < nlay;
      const int lay = o.lay(FULL ? k0 + i : min(k0 + i, nlay - 1));
      Float tau, w0, gg;
      if (TMA) {
        const Float* e = tile_at(tile_tau, lay - row0, cw);
        tau = e[0]; w0 = e[te_lay]; gg = e[2 * te_lay];
      } else {
        const size_t i3 = col + ncol * lay + ncl * g;
        tau = p.tau[i3]; w0 = p.ssa[i3]; gg = p.g[i3];
      }
      const Float gamma1 = LW_diff_sec * ((Float)1 - (Float)0.5 * w0 * ((Float)1 + gg));   // :879
      const Float gamma2 = LW_diff_sec * (Float)0.5 * w0 * ((Float)1 - gg);                // :880
      const Float kk = rb_sqrt(fmax((gamma1 - gamma2) * (gamma1 + gamma2), (Float)1.e-12));   // :885
      const Float exp_minusktau = rb_exp(-tau * kk);
      const Float exp_minus2ktau = exp_minusktau * exp_minusktau;
      const Float RT_term = rb_rcp(kk * ((Float)1 + exp_minus2ktau) + gamma1 * ((Float)1 - exp_minus2ktau));
      const Float rdif = RT_term * gamma2 * ((Float)1 - exp_minus2ktau);
      const Float tdif = RT_term * (Float)2 * kk * exp_minusktau;
      const Float lev_top = Blev[i], lev_bot = Blev[i + 1];
      // :947-957; the divisor is clamped where the source is switched off anyway (tau <= 1e-8)
      constexpr bool SEL = FULL == 0;   // FULL = 2: zero-filled padding rows have tau = 0 -> no source, rdif = 0 exactly; tdif by select
      const bool has_src = (!SEL || live) && tau > (Float)1.0e-8;
      const Float Z = rb_div(lev_bot - lev_top, fmax(tau, (Float)1.0e-8) * (gamma1 + gamma2));
      const Float Zup_top = Z + lev_top;
      const Float Zup_bottom = Z + lev_bot;
      const Float Zdn_top = -Z + lev_top;
      const Float Zdn_bottom = -Z + lev_bot;
      const Float s_up = pi * (Zup_top - rdif * Zdn_top - tdif * Zup_bottom);
      const Float s_dn = pi * (Zdn_bottom - rdif * Zup_bottom - tdif * Zdn_top);
      R[i] = (!SEL || live) ? rdif : (Float)0;
      T[i] = live ? tdif : (Float)1;
      SU[i] = has_src ? s_up : (Float)0;
      SD[i] = has_src ? s_dn : (Float)0;
    }
    if (TMA) {  // every warp of the CTA has read stage s into registers: hand it back to the TMA for g+2
      __syncthreads();
      if (g + 2 < ge) issue(g + 2, s);
    }
    const Float emis = p.sfc_emis[gi];
    auto top = [&](Float fup, Float fdn) {
      if (!col_ok) return;
      const size_t q = col + ncol * o.lev(0);
      gup[q] = fup;
      gdn[q] = fdn;
    };
    auto lev = [&](int i, Float fup, Float fdn) {
      if ((!FULLG && k0 + i >= nlay) || !col_ok) return;
      const size_t q = col + ncol * o.lev(k0 + i + 1);
      gup[q] = fup;
      gdn[q] = fdn;
    };
    adding_reg<CL, kRegChunks>(j, R, T, SU, SD, (Float)1 - emis, pi * emis * p.sfc_src[gi], p.inc_flux[gi], top, lev);
  }
}

}  // namespace rrtmgpb
