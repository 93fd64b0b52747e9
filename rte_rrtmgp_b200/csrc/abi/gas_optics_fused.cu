// gas_optics_fused.cu - fused gas-optics fast path used by the device-resident frontend.
//
// The reference sequence for one gas_optics() call (mo_gas_optics_rrtmgp.F90:419-745,840-928) is
//   get_col_dry -> col_gas -> interpolation -> zero(tau) -> tau_absorption (3 RMW passes) -> [tau_rayleigh ->
//   combine_abs_and_rayleigh] -> compute_Planck_source, followed by clouds%increment(atmos) in the driver.
// Its intermediates (col_gas, jtemp, jpress, tropo, jeta, col_mix, fmajor, fminor, tau_rayleigh) are frontend
// locals; only atmos%tau/ssa/g and the Planck sources are visible to the caller.  Materialising them costs
// ~1.2 KB per (column, layer) that every later kernel reads back once per band.  Here:
//   1. cell_state_kernel   one thread per (col,lay): col_dry, jtemp, ftemp, jpress, fpress, tropo and the
//                          band-independent factors of the minor-gas scaling (57 B/cell)
//   2. gas_tau_g_kernel    thread = (cell, band)                  } kernels/gas_optics_gfast.cuh, reading
//   3. planck_g_kernel     thread = (column, band, layer chunk)   } g-point-fastest copies of the tables
// Arithmetic is the reference's, expression by expression (same cited lines as gas_optics_abi.cu), so results
// equal the unfused kernels' to rounding of FMA contraction; tests/test_allsky_parity.py checks fused vs
// unfused vs oracle.
//
// Table cache: the g-point-fastest copies and the small per-band / per-contributor records are built on first
// use of a k-distribution (keyed by its kmajor pointer) and dropped when that table is released through
// rrtmgpb_mem_free().
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <atomic>
#include <map>
#include <mutex>
#include <vector>
#include "../kernels/elementwise.cuh"
#include "../kernels/gas_optics_gfast.cuh"
#include "rrtmgp_b200_ext.h"
#include "rte_kernels.h"

using namespace rrtmgpb;

namespace {

constexpr int kFThreads = 128;

// ---- per-cell state: mo_gas_optics_utils.F90:143-150 (col_dry), mo_gas_optics_rrtmgp_kernels.F90:99-118,
// and the per-cell factors of the minor scaling :467-471 ----
// stat (optional): a sampled count of neighbouring cells (consecutive columns of a layer) that fall into different
// temperature / pressure bins or atmosphere halves - {differing pairs, pairs looked at} - from every 64th block; the host
// reads the PREVIOUS call's counts to choose the tau kernels' thread mapping (launch_tau: automatic rows path)
constexpr int kStatEvery = 64;
__device__ __forceinline__ void neighbour_stat(int* stat, int jtemp, int jpress, bool tropo) {
  if (!stat || blockIdx.x % kStatEvery != 0) return;
  const unsigned act = __activemask();
  const int lane = threadIdx.x & 31;
  const int key = jtemp | (jpress << 8) | (tropo ? (1 << 16) : 0);
  const int nxt = __shfl_down_sync(act, key, 1);
  const bool has_next = lane < 31 && ((act >> (lane + 1)) & 1u);
  const unsigned differ = __ballot_sync(act, has_next && nxt != key), pairs = __ballot_sync(act, has_next);
  if (lane == __ffs(act) - 1) {
    atomicAdd(stat, __popc(differ));
    atomicAdd(stat + 1, __popc(pairs));
  }
}
__global__ void __launch_bounds__(kFThreads) cell_state_kernel(const FusedParams p, Float m_dry, Float m_h2o,
                                                                Float avogad, Float grav, int* stat) {
  const size_t ncl = (size_t)p.ncol * p.nlay;
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncl) return;
  const rrtmgpb_gas_tables& t = p.t;
  const Float vh2o = p.vmr[c + ncl * (size_t)(t.idx_h2o - 1)];
  Float col_dry;
  if (p.col_dry_in) {
    col_dry = p.col_dry_in[c];
  } else {
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
  const Float vmr_fact = (Float)1 / col_dry;                               // :470
  p.cs.pt_scale[c] = (Float)0.01 * pl / tl;                                // :467
  p.cs.vmr_fact[c] = vmr_fact;
  p.cs.dry_fact[c] = (Float)1 / ((Float)1 + (vh2o * col_dry) * vmr_fact);  // :471, col_gas(h2o) = vmr*col_dry
  neighbour_stat(stat, jtemp, (int)jpress_aint, pl > press_ref_trop);
}

// per host thread: device counters of cell_state_kernel and their pinned host mirror (copied back stream-ordered after every
// call; read - never waited for - by the next call)
struct RowsStat {
  int* d = nullptr;
  int* h = nullptr;
  int device = -1;
};
RowsStat& rows_stat() {
  thread_local RowsStat st;
  int dev = 0;
  RB_CUDA_CHECK(cudaGetDevice(&dev));
  if (st.device != dev) {   // first use on this thread, or the thread moved to another device (the old pair is abandoned)
    RB_CUDA_CHECK(cudaMalloc(&st.d, 2 * sizeof(int)));
    RB_CUDA_CHECK(cudaHostAlloc(&st.h, 2 * sizeof(int), cudaHostAllocDefault));
    st.h[0] = st.h[1] = 0;
    st.device = dev;
  }
  return st;
}

struct Workspace {
  CellState cs;
  void* block;
};

// physical constants (kept in util_abi.cu; mirrored here through the setter below)
double g_m_dry = 0.028964, g_grav = 9.80665;
const double k_m_h2o = 0.018016, k_avogad = 6.02214076e23;

Workspace prepare(FusedParams& p) {
  const size_t ncl = (size_t)p.ncol * p.nlay;
  // one pool allocation: 6 doubles, 2 ints, 1 byte per cell
  const size_t bytes = ncl * (6 * sizeof(Float) + 2 * sizeof(int) + 1) + 64;
  Workspace w;
  w.block = dev_alloc(bytes);
  Float* f = static_cast<Float*>(w.block);
  w.cs.col_dry = f; w.cs.ftemp = f + ncl; w.cs.fpress = f + 2 * ncl;
  w.cs.pt_scale = f + 3 * ncl; w.cs.vmr_fact = f + 4 * ncl; w.cs.dry_fact = f + 5 * ncl;
  int* ii = reinterpret_cast<int*>(f + 6 * ncl);
  w.cs.jtemp = ii; w.cs.jpress = ii + ncl;
  w.cs.tropo = reinterpret_cast<Bool*>(ii + 2 * ncl);
  p.cs = w.cs;
  {
    KernelTimer timer("gas_cell_state");
    RowsStat& st = rows_stat();
    RB_CUDA_CHECK(cudaMemsetAsync(st.d, 0, 2 * sizeof(int), stream()));
    cell_state_kernel<<<ceil_div((long long)ncl, kFThreads), kFThreads, 0, stream()>>>(
        p, (Float)g_m_dry, (Float)k_m_h2o, (Float)k_avogad, (Float)g_grav, st.d);
    RB_LAUNCH_CHECK();
    RB_CUDA_CHECK(cudaMemcpyAsync(st.h, st.d, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream()));
  }
  return w;
}

// ---- g-point-fastest table copies, built once per k-distribution (kernels/gas_optics_gfast.cuh) ----
struct TableCacheEntry {
  TablesT tt;
  std::vector<void*> owned;
  int ntemp, neta, npres, ngpt, nkl, nku, nbnd, nflav;
  // every array the copies and the host-resolved records were derived from: an entry is only reused when ALL of them
  // are the same allocations (a new k-distribution that happens to reuse the kmajor address does not match unless
  // it reuses all of them - and then rrtmgpb_tables_changed() is the caller's explicit invalidation)
  const void* src[16];
};
void table_sources(const rrtmgpb_gas_tables& t, const void* (&src)[16]) {
  const void* v[16] = {t.kmajor, t.planck_frac, t.kminor_lower, t.kminor_upper, t.krayl, t.band_lims_gpt, t.gpoint_flavor,
                       t.flavor, t.vmr_ref, t.minor_limits_gpt_lower, t.minor_limits_gpt_upper, t.kminor_start_lower,
                       t.kminor_start_upper, t.idx_minor_lower, t.idx_minor_upper, t.totplnk};
  for (int i = 0; i < 16; ++i) src[i] = v[i];
}
std::mutex g_tc_mutex;
// key: the loader-layout kmajor pointer (the Planck-only entry of the extern symbols: its pfracin pointer).  Slot 0: the
// fused path (complete table sets); slot 1: sets assembled from the arguments of the extern symbols (no flavor / vmr_ref:
// the interpolation weights arrive as arrays), kept apart so that a host alternating between both does not evict
std::map<const void*, TableCacheEntry> g_table_cache_slots[2];

void transpose_into(const Float* in, Float* out, int nrow, int ng, int pitch) {
  dim3 grid(ceil_div(nrow, 32), ceil_div(ng, 32)), block(32, 8);
  transpose_table_kernel<<<grid, block, 0, stream()>>>(in, out, nrow, ng, pitch);
  RB_LAUNCH_CHECK();
}

Float* transposed(const Float* in, int nslice, int nrow, int ng, int pitch, std::vector<void*>& owned) {
  if (!in || nrow <= 0 || ng <= 0) return nullptr;
  const size_t n = (size_t)nslice * nrow * pitch;
  Float* out = static_cast<Float*>(dev_alloc(n * sizeof(Float)));
  RB_CUDA_CHECK(cudaMemsetAsync(out, 0, n * sizeof(Float), stream()));
  for (int s = 0; s < nslice; ++s) transpose_into(in + (size_t)s * nrow * ng, out + (size_t)s * nrow * pitch, nrow, ng, pitch);
  owned.push_back(out);
  return out;
}

template <typename T>
std::vector<T> to_host(const T* dev, size_t n) {
  std::vector<T> h(n);
  if (n) RB_CUDA_CHECK(cudaMemcpyAsync(h.data(), dev, n * sizeof(T), cudaMemcpyDeviceToHost, stream()));
  RB_CUDA_CHECK(cudaStreamSynchronize(stream()));
  return h;
}
template <typename T>
const T* to_device(const std::vector<T>& h, std::vector<void*>& owned) {
  T* d = static_cast<T*>(dev_alloc((h.size() ? h.size() : 1) * sizeof(T)));
  if (!h.empty()) RB_CUDA_CHECK(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, stream()));
  RB_CUDA_CHECK(cudaStreamSynchronize(stream()));  // h dies with the caller
  owned.push_back(d);
  return d;
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

struct MinorHost {
  std::vector<int> lim, idx, isc, ks;
  std::vector<unsigned char> dens, comp;  // Bool is 1 byte (rte_types.h)
};
MinorHost minor_host(int n, const int* lim, const int* idx, const int* isc, const int* ks, const Bool* dens, const Bool* comp) {
  MinorHost m;
  m.lim = to_host(lim, 2 * (size_t)n); m.idx = to_host(idx, (size_t)n); m.isc = to_host(isc, (size_t)n);
  m.ks = to_host(ks, (size_t)n); m.dens = to_host(reinterpret_cast<const unsigned char*>(dens), (size_t)n);
  m.comp = to_host(reinterpret_cast<const unsigned char*>(comp), (size_t)n);
  return m;
}

TablesT tables_gfast(const rrtmgpb_gas_tables& t, int slot = 0) {
  std::lock_guard<std::mutex> lock(g_tc_mutex);
  std::map<const void*, TableCacheEntry>& g_table_cache = g_table_cache_slots[slot];
  const void* key = t.kmajor ? static_cast<const void*>(t.kmajor) : static_cast<const void*>(t.planck_frac);
  auto it = g_table_cache.find(key);
  if (it != g_table_cache.end()) {
    const TableCacheEntry& e = it->second;
    const void* src[16];
    table_sources(t, src);
    if (e.ntemp == t.ntemp && e.neta == t.neta && e.npres == t.npres && e.ngpt == t.ngpt && e.nkl == t.nminorklower &&
        e.nku == t.nminorkupper && e.nbnd == t.nbnd && e.nflav == t.nflav && std::equal(src, src + 16, e.src))
      return e.tt;
    for (void* q : e.owned) dev_free(q);
    g_table_cache.erase(it);
  }
  TableCacheEntry e;
  e.ntemp = t.ntemp; e.neta = t.neta; e.npres = t.npres; e.ngpt = t.ngpt; e.nkl = t.nminorklower; e.nku = t.nminorkupper;
  e.nbnd = t.nbnd; e.nflav = t.nflav;
  table_sources(t, e.src);
  const int tn = t.ntemp * t.neta, rows = tn * (t.npres + 1);
  TablesT& tt = e.tt;
  tt.maxm = 1;
  tt.gp = (t.ngpt + 1) & ~1;
  tt.nkl = (t.nminorklower + 1) & ~1;
  tt.nku = (t.nminorkupper + 1) & ~1;
  tt.kmajor = transposed(t.kmajor, 1, rows, t.ngpt, tt.gp, e.owned);
  tt.pfrac = transposed(t.planck_frac, 1, rows, t.ngpt, tt.gp, e.owned);
  tt.kminor_lower = transposed(t.kminor_lower, 1, tn, t.nminorklower, tt.nkl, e.owned);
  tt.kminor_upper = transposed(t.kminor_upper, 1, tn, t.nminorkupper, tt.nku, e.owned);
  tt.krayl = transposed(t.krayl, 2, tn, t.ngpt, tt.gp, e.owned);  // (ntemp, neta, ngpt, 2): one slice per tropo half

  // ---- small records, resolved on the host ----
  const std::vector<int> bl = to_host(t.band_lims_gpt, 2 * (size_t)t.nbnd);
  const std::vector<int> gf = to_host(t.gpoint_flavor, 2 * (size_t)t.ngpt);
  // (extern-symbol table sets carry no flavor / vmr_ref: the records derived from them are only read by the fused path)
  const std::vector<int> fl = t.flavor ? to_host(t.flavor, 2 * (size_t)t.nflav) : std::vector<int>(2 * (size_t)t.nflav, 0);
  const std::vector<Float> vr = t.vmr_ref ? to_host(t.vmr_ref, 2 * (size_t)(t.ngas + 1) * t.ntemp)
                                          : std::vector<Float>(2 * (size_t)(t.ngas + 1) * t.ntemp, (Float)1);
  const MinorHost mh[2] = {
      minor_host(t.nminorlower, t.minor_limits_gpt_lower, t.idx_minor_lower, t.idx_minor_scaling_lower, t.kminor_start_lower,
                 t.minor_scales_with_density_lower, t.scale_by_complement_lower),
      minor_host(t.nminorupper, t.minor_limits_gpt_upper, t.idx_minor_upper, t.idx_minor_scaling_upper, t.kminor_start_upper,
                 t.minor_scales_with_density_upper, t.scale_by_complement_upper)};
  const int nminor[2] = {t.nminorlower, t.nminorupper};
  std::vector<BandInfo> bands(t.nbnd);
  for (int b = 0; b < t.nbnd; ++b) {
    BandInfo& bi = bands[b];
    bi.bS = bl[2 * b]; bi.bE = bl[2 * b + 1];
    for (int a = 0; a < 2; ++a) {
      bi.iflav[a] = gf[a + 2 * (bi.bS - 1)] - 1;
      bi.igas1[a] = fl[2 * bi.iflav[a]]; bi.igas2[a] = fl[2 * bi.iflav[a] + 1];
      int first = nminor[a], last = -1;
      for (int i = 0; i < nminor[a]; ++i)
        if (mh[a].lim[2 * i + 1] >= bi.bS && mh[a].lim[2 * i] <= bi.bE) { first = std::min(first, i); last = std::max(last, i); }
      bi.mfirst[a] = first; bi.mlast[a] = last; bi.mdiff[a] = 0;
      tt.maxm = std::max(tt.maxm, last - first + 1);
    }
  }
  std::vector<MinorInfo> minfo[2];
  for (int a = 0; a < 2; ++a) {
    minfo[a].resize(nminor[a]);
    for (int i = 0; i < nminor[a]; ++i) {
      MinorInfo& mi = minfo[a][i];
      mi.mS = mh[a].lim[2 * i]; mi.mE = mh[a].lim[2 * i + 1];
      mi.igas = mh[a].idx[i]; mi.isc = mh[a].isc[i];
      mi.dens = mh[a].dens[i] ? 1 : 0; mi.comp = mh[a].comp[i] ? 1 : 0;
      mi.kstart = mh[a].ks[i];
      mi.iflav = (mi.mS >= 1 && mi.mS <= t.ngpt) ? gf[a + 2 * (mi.mS - 1)] - 1 : 0;
      mi.igas1 = fl[2 * mi.iflav]; mi.igas2 = fl[2 * mi.iflav + 1];
      for (int b = 0; b < t.nbnd; ++b)
        if (i >= bands[b].mfirst[a] && i <= bands[b].mlast[a] && mi.iflav != bands[b].iflav[a]) bands[b].mdiff[a] = 1;
    }
    for (int b = 0; b < t.nbnd; ++b) {  // regular band: 16 g-points, every contributor covers exactly the band
      BandInfo& bi = bands[b];
      bool reg = (bi.bE - bi.bS + 1 == kTauRegChunks * kTG) && !bi.mdiff[a];
      for (int i = bi.mfirst[a]; reg && i <= bi.mlast[a]; ++i)
        reg = minfo[a][i].mS == bi.bS && minfo[a][i].mE == bi.bE;
      bi.regular[a] = reg ? 1 : 0;
    }
  }
  // ratio_eta_half = vmr_ref(itropo,igas_1,jt) / vmr_ref(itropo,igas_2,jt), mo_gas_optics_rrtmgp_kernels.F90:127-128
  std::vector<Float> ratio((size_t)2 * t.nflav * t.ntemp);
  for (int a = 0; a < 2; ++a)
    for (int f = 0; f < t.nflav; ++f)
      for (int jt = 0; jt < t.ntemp; ++jt) {
        const Float num = vr[a + 2 * (fl[2 * f] + (size_t)(t.ngas + 1) * jt)];
        const Float den = vr[a + 2 * (fl[2 * f + 1] + (size_t)(t.ngas + 1) * jt)];
        ratio[((size_t)a * t.nflav + f) * t.ntemp + jt] = num / den;
      }
  tt.aux.band = to_device(bands, e.owned);
  tt.aux.minor_lower = to_device(minfo[0], e.owned);
  tt.aux.minor_upper = to_device(minfo[1], e.owned);
  tt.aux.ratio = to_device(ratio, e.owned);
  tt.vec = (sizeof(Float) == 8 && intervals_even(bl, nullptr) && intervals_even(mh[0].lim, &mh[0].ks) &&
            intervals_even(mh[1].lim, &mh[1].ks)) ? 2 : 1;
  RB_CUDA_CHECK(cudaStreamSynchronize(stream()));  // other host threads (other streams) may use the copies next
  g_table_cache[key] = e;
  return e.tt;
}

// ---- the same g-point-fastest copies for the kernel-by-kernel ABI entry points (gas_optics_abi.cu) ----
// Only kmajor / kminor_* are needed there (the interpolation weights arrive as arguments).  Keyed by the kmajor
// pointer like the fused path's cache, so it is only valid for callers that keep a table immutable while its
// allocation lives: off unless rrtmgpb_abi_table_cache(1) was called (the C++ frontend mirror does, its tables are
// released through rrtmgpb_mem_free which drops the copies).
struct AbiCacheEntry {
  TablesT tt;
  std::vector<void*> owned;
  int ntemp, neta, npres, ngpt, nkl, nku;
};
std::map<const void*, AbiCacheEntry> g_abi_cache;
thread_local int g_abi_cache_on = 0;
}  // namespace

namespace rrtmgpb {
bool tables_gfast_abi(const Float* kmajor, const Float* kminor_lower, const Float* kminor_upper, int ntemp, int neta,
                      int npres, int ngpt, int nkl, int nku, const Float** kmajorT, const Float** kminorT_lower,
                      const Float** kminorT_upper, int* gp, int* pitch_lower, int* pitch_upper) {
  if (!g_abi_cache_on) return false;
  std::lock_guard<std::mutex> lock(g_tc_mutex);
  auto it = g_abi_cache.find(kmajor);
  if (it != g_abi_cache.end()) {
    const AbiCacheEntry& e = it->second;
    if (!(e.ntemp == ntemp && e.neta == neta && e.npres == npres && e.ngpt == ngpt && e.nkl == nkl && e.nku == nku)) {
      for (void* q : e.owned) dev_free(q);
      g_abi_cache.erase(it);
      it = g_abi_cache.end();
    }
  }
  if (it == g_abi_cache.end()) {
    AbiCacheEntry e;
    e.ntemp = ntemp; e.neta = neta; e.npres = npres; e.ngpt = ngpt; e.nkl = nkl; e.nku = nku;
    const int tn = ntemp * neta;
    e.tt.gp = (ngpt + 1) & ~1; e.tt.nkl = (nkl + 1) & ~1; e.tt.nku = (nku + 1) & ~1;
    e.tt.kmajor = transposed(kmajor, 1, tn * (npres + 1), ngpt, e.tt.gp, e.owned);
    e.tt.kminor_lower = transposed(kminor_lower, 1, tn, nkl, e.tt.nkl, e.owned);
    e.tt.kminor_upper = transposed(kminor_upper, 1, tn, nku, e.tt.nku, e.owned);
    RB_CUDA_CHECK(cudaStreamSynchronize(stream()));  // other host threads (other streams) may use the copies next
    it = g_abi_cache.emplace(kmajor, e).first;
  }
  const TablesT& tt = it->second.tt;
  *kmajorT = tt.kmajor; *kminorT_lower = tt.kminor_lower; *kminorT_upper = tt.kminor_upper;
  *gp = tt.gp; *pitch_lower = tt.nkl; *pitch_upper = tt.nku;
  return tt.kmajor != nullptr;
}
}  // namespace rrtmgpb

namespace {
}  // namespace

namespace rrtmgpb {
// called by rrtmgpb_mem_free(): a k-distribution that is being released takes its transposed copies with it
void table_cache_release(const void* key) {
  std::lock_guard<std::mutex> lock(g_tc_mutex);
  for (auto& g_table_cache : g_table_cache_slots) {
    auto it = g_table_cache.find(key);
    if (it == g_table_cache.end()) continue;
    for (void* q : it->second.owned) dev_free(q);
    g_table_cache.erase(it);
  }
}
void table_cache_release_abi(const void* key) {
  std::lock_guard<std::mutex> lock(g_tc_mutex);
  auto it = g_abi_cache.find(key);
  if (it == g_abi_cache.end()) return;
  for (void* q : it->second.owned) dev_free(q);
  g_abi_cache.erase(it);
}
void fused_set_constants(double grav, double m_dry) { g_grav = grav; g_m_dry = m_dry; }
}

extern "C" void rrtmgpb_tables_changed(const void* kmajor) {
  if (kmajor) {
    table_cache_release(kmajor);
    table_cache_release_abi(kmajor);
    return;
  }
  std::lock_guard<std::mutex> lock(g_tc_mutex);
  for (auto& g_table_cache : g_table_cache_slots) {
    for (auto& kv : g_table_cache) for (void* q : kv.second.owned) dev_free(q);
    g_table_cache.clear();
  }
  for (auto& kv : g_abi_cache) for (void* q : kv.second.owned) dev_free(q);
  g_abi_cache.clear();
}

namespace {
// gas_tau_g_kernel for the bands p.band0 .. p.band0 + p.nband_sub - 1
std::atomic<int> g_tau_rows{-1};
void launch_tau(const FusedParams& p, const TablesT& tt, bool abi = false) {
  const rrtmgpb_gas_tables* t = &p.t;
  const size_t ncl = (size_t)p.ncol * p.nlay;
  const bool sw = t->krayl != nullptr;
  const int op_kind = p.op_kind, cld_kind = p.cld_kind, aer_kind = p.aer_kind;
  KernelTimer timer(abi ? "tau_absorption" : (sw ? "gas_tau_fused[sw]" : "gas_tau_fused[lw]"));
  const unsigned grid = (unsigned)((long long)ceil_div((long long)ncl, kTauCells * kGThreads) * p.nband_sub);
  // lane-private slots for the per-(cell, band) minor scalings: [contributor][cell slot][thread]
  size_t smem = (size_t)tt.maxm * kTauCells * kGThreads * sizeof(Float);
  // table rows staged per warp by cp.async.bulk when the warp's cells share them (gas_tau_g_kernel STAGE); needs the
  // 128-bit layout (tt.vec == 2) and 16-g-point chunks; RRTMGPB_TABLE_TMA=0 switches it off
  static const bool stage_env = [] { const char* e = std::getenv("RRTMGPB_TABLE_TMA"); return !(e && e[0] == '0'); }();
  const bool stage = stage_env && tt.vec == 2 && kTG * kTauRegChunks == 16;   // (also with a second, aerosol increment: config 5)
  // lanes-along-g-points mapping for warps of unrelated columns (tau_band_rows): its records overlay the warp's staging
  // slots.  RRTMGPB_TAU_ROWS=0/1 (A/B switch)
  // rows_env: RRTMGPB_TAU_ROWS = 0 (never) / 1 (always the ROWS instantiations) / unset (-1: automatic - the ROWS instantiations
  // when, in the previous call's sample, more than one in eight neighbouring cells fell into different T / p bins)
  static const int rows_env_raw = [] { const char* e = std::getenv("RRTMGPB_TAU_ROWS"); return (e && (e[0] == '0' || e[0] == '1')) ? e[0] - '0' : -1; }();
  int rows_auto = 0;
  {
    const RowsStat& st = rows_stat();
    const int differ = st.h[0], pairs = st.h[1];   // (a torn or stale read only delays the switch by a call)
    rows_auto = (pairs > 0 && differ * 8 > pairs) ? 1 : 0;
  }
  const bool rows_env = rows_env_raw < 0 ? rows_auto != 0 : rows_env_raw != 0;
  const int rows_set = g_tau_rows.load(std::memory_order_relaxed);   // rrtmgpb_set_gas_optics_rows_path: -1 = environment
  FusedParams pr = p;
  // 0: off; otherwise the vote threshold of the ROWS instantiations: a warp re-maps when fewer than this many of its lanes share
  // table rows between their own two cells (1 = the default, 32: any lane that does not; 2..33 = that threshold, for
  // experiments).  B200, 65,536 x 60 distinct columns, LW / SW tau in ms: off 9.52 / 13.67; two g-points per lane: 12: 8.19 / 13.09,
  // 20: 7.37 / 12.16, 24: 7.03 / 11.66, 28: 6.92 / 11.36, 33 (every warp, even uniform ones): 8.34 / 13.81; four g-points per lane
  // (RB_ROWS_GPL = 4, shipped): 24: 6.85 / 10.35, 28: 6.69 / 9.78, 30: 6.63 / 9.54, 32: 6.52 / 9.24; eight per lane, 28: 7.33 / 9.79
  const int rows_req = stage ? (rows_set < 0 ? (rows_env ? 1 : 0) : rows_set) : 0;
  pr.rows_path = rows_req <= 0 ? 0 : (rows_req == 1 ? 32 : std::min(rows_req, 33));
  {
    const size_t per_warp = std::max((size_t)(kStgMinor / 16 + 4 * tt.maxm) * 16 * sizeof(Float), pr.rows_path ? tau_rows_warp_bytes() : (size_t)0);
    pr.stg_stride = (int)(per_warp / sizeof(Float));
    if (stage) smem += (size_t)(kGThreads / 32) * per_warp;
  }
// KIND 1: the common kinds as compile-time constants (LW 1scl += 1scl clouds, SW 2str += 2str clouds, no aerosols)
#define GAS_TAU_LAUNCH1(SWV, VECV, AERV, KINDV)                                                                   \
  do {                                                                                                            \
    if (stage && VECV == 2 && !AERV && pr.rows_path) {   /* the ROWS instantiations live in gas_optics_rows.cu */ \
      launch_tau_rows(pr, tt, grid, smem, SWV, KINDV, false);                                                     \
      break;                                                                                                      \
    }                                                                                                             \
    auto kern = (stage && VECV == 2) ? gas_tau_g_kernel<SWV, VECV, AERV, KINDV, (VECV == 2)>                      \
                                     : gas_tau_g_kernel<SWV, VECV, AERV, KINDV, false>;                           \
    if (smem > 48 * 1024) RB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<grid, kGThreads, smem, stream()>>>(pr, tt);                                                             \
  } while (0)
#define GAS_TAU_LAUNCH(SWV, VECV)                                                                     \
  if (aer_kind) GAS_TAU_LAUNCH1(SWV, VECV, true, 0);                                                  \
  else if (op_kind == (SWV ? 2 : 1) && cld_kind == (SWV ? 2 : 1)) GAS_TAU_LAUNCH1(SWV, VECV, false, 1); \
  else GAS_TAU_LAUNCH1(SWV, VECV, false, 0)
  if (abi) {  // absorption only, the interpolation state from the caller's arrays (extern symbol rrtmgp_compute_tau_absorption)
#define GAS_TAU_LAUNCH_ABI(VECV, STGV)                                                                            \
  do {                                                                                                            \
    if (STGV && pr.rows_path) {                                                                                   \
      launch_tau_rows(pr, tt, grid, smem, false, 0, true);                                                        \
      break;                                                                                                      \
    }                                                                                                             \
    auto kern = gas_tau_g_kernel<false, VECV, false, 0, STGV, true>;                                              \
    if (smem > 48 * 1024) RB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<grid, kGThreads, smem, stream()>>>(pr, tt);                                                             \
  } while (0)
    if (tt.vec == 2 && stage) GAS_TAU_LAUNCH_ABI(2, true);
    else if (tt.vec == 2) GAS_TAU_LAUNCH_ABI(2, false);
    else GAS_TAU_LAUNCH_ABI(1, false);
#undef GAS_TAU_LAUNCH_ABI
  } else if (sw) {
    if (tt.vec == 2) { GAS_TAU_LAUNCH(true, 2); } else { GAS_TAU_LAUNCH(true, 1); }
  } else {
    if (tt.vec == 2) { GAS_TAU_LAUNCH(false, 2); } else { GAS_TAU_LAUNCH(false, 1); }
  }
#undef GAS_TAU_LAUNCH
#undef GAS_TAU_LAUNCH1
  RB_LAUNCH_CHECK();
}

void launch_planck(const PlanckFusedParams& q, const TablesT& tt, bool abi = false) {
  KernelTimer timer(abi ? "planck_source" : "planck_fused");
  // layers a thread marches through (its first level needs the Planck fractions of the layer above: 1/lay_per_chunk
  // redundant work; B200, 65,536 x 72: 9 -> 5.40 ms, 12 -> 5.31, 18 -> 5.24, 36 -> 5.22); RRTMGPB_PLANCK_CHUNK overrides
  static const int chunk_env = [] { const char* e = std::getenv("RRTMGPB_PLANCK_CHUNK"); return e ? std::atoi(e) : 0; }();
  const int lay_per_chunk = chunk_env > 0 ? chunk_env : 18, nchunk = ceil_div(q.f.nlay, lay_per_chunk);
  const unsigned grid = (unsigned)((long long)ceil_div(q.f.ncol, kGThreads) * nchunk * q.f.nband_sub);
  if (abi) {
    if (tt.vec == 2) planck_g_kernel<2, true><<<grid, kGThreads, 0, stream()>>>(q, tt, lay_per_chunk, nchunk);
    else planck_g_kernel<1, true><<<grid, kGThreads, 0, stream()>>>(q, tt, lay_per_chunk, nchunk);
  } else if (tt.vec == 2) planck_g_kernel<2><<<grid, kGThreads, 0, stream()>>>(q, tt, lay_per_chunk, nchunk);
  else planck_g_kernel<1><<<grid, kGThreads, 0, stream()>>>(q, tt, lay_per_chunk, nchunk);
  RB_LAUNCH_CHECK();
}
}  // namespace

// ---- the extern symbols' tau_absorption / Planck_source on the g-point-fastest kernels (ABI instantiations) ----
namespace {
// the three per-cell factors of the minor scaling (:467-471) from the caller's play, tlay, col_gas
__global__ void __launch_bounds__(kFThreads) abi_cell_prep_kernel(size_t ncl, int idx_h2o, const Float* __restrict__ play,
                                                                   const Float* __restrict__ tlay, const Float* __restrict__ col_gas,
                                                                   const int* __restrict__ jtemp, const int* __restrict__ jpress,
                                                                   const Bool* __restrict__ tropo, Float* pt_scale, Float* vmr_fact,
                                                                   Float* dry_fact, int* stat) {
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncl) return;
  const Float vf = (Float)1 / col_gas[c];                                        // :470, col_gas(:,:,0) = col_dry
  pt_scale[c] = (Float)0.01 * play[c] / tlay[c];                                 // :467
  vmr_fact[c] = vf;
  dry_fact[c] = (Float)1 / ((Float)1 + col_gas[c + ncl * (size_t)idx_h2o] * vf); // :471
  neighbour_stat(stat, jtemp[c], jpress[c], tropo[c]);   // (see cell_state_kernel: the next call's choice of thread mapping)
}
}  // namespace

namespace rrtmgpb {
// All array arguments are DEVICE pointers (the callers have classified / staged them).  Returns false when the caller has
// not allowed cached table copies (rrtmgpb_abi_table_cache) - the loader-layout kernels of gas_optics_abi.cu run then.
bool tau_absorption_gfast(int ncol, int nlay, int nbnd, int ngpt, int ngas, int nflav, int neta, int npres, int ntemp,
                          int nminorlower, int nminorklower, int nminorupper, int nminorkupper, int idx_h2o,
                          const int* gpoint_flavor, const int* band_lims_gpt, const Float* kmajor, const Float* kminor_lower,
                          const Float* kminor_upper, const int* minor_limits_gpt_lower, const int* minor_limits_gpt_upper,
                          const Bool* minor_scales_with_density_lower, const Bool* minor_scales_with_density_upper,
                          const Bool* scale_by_complement_lower, const Bool* scale_by_complement_upper,
                          const int* idx_minor_lower, const int* idx_minor_upper, const int* idx_minor_scaling_lower,
                          const int* idx_minor_scaling_upper, const int* kminor_start_lower, const int* kminor_start_upper,
                          const Bool* tropo, const Float* col_mix, const Float* fmajor, const Float* fminor, const Float* play,
                          const Float* tlay, const Float* col_gas, const int* jeta, const int* jtemp, const int* jpress,
                          Float* tau, bool accumulate) {
  if (!g_abi_cache_on) return false;
  rrtmgpb_gas_tables t{};
  t.ngas = ngas; t.nflav = nflav; t.neta = neta; t.npres = npres; t.ntemp = ntemp; t.nbnd = nbnd; t.ngpt = ngpt;
  t.nminorlower = nminorlower; t.nminorupper = nminorupper; t.nminorklower = nminorklower; t.nminorkupper = nminorkupper;
  t.idx_h2o = idx_h2o; t.gpoint_flavor = gpoint_flavor; t.band_lims_gpt = band_lims_gpt; t.kmajor = kmajor;
  t.kminor_lower = kminor_lower; t.kminor_upper = kminor_upper; t.minor_limits_gpt_lower = minor_limits_gpt_lower;
  t.minor_limits_gpt_upper = minor_limits_gpt_upper; t.minor_scales_with_density_lower = minor_scales_with_density_lower;
  t.minor_scales_with_density_upper = minor_scales_with_density_upper; t.scale_by_complement_lower = scale_by_complement_lower;
  t.scale_by_complement_upper = scale_by_complement_upper; t.idx_minor_lower = idx_minor_lower; t.idx_minor_upper = idx_minor_upper;
  t.idx_minor_scaling_lower = idx_minor_scaling_lower; t.idx_minor_scaling_upper = idx_minor_scaling_upper;
  t.kminor_start_lower = kminor_start_lower; t.kminor_start_upper = kminor_start_upper;
  const TablesT tt = tables_gfast(t, 1);
  const size_t ncl = (size_t)ncol * nlay;
  Float* prep = static_cast<Float*>(dev_alloc(3 * ncl * sizeof(Float)));
  {
    KernelTimer timer("tau_absorption_cell_prep");
    RowsStat& st = rows_stat();
    RB_CUDA_CHECK(cudaMemsetAsync(st.d, 0, 2 * sizeof(int), stream()));
    abi_cell_prep_kernel<<<ceil_div((long long)ncl, kFThreads), kFThreads, 0, stream()>>>(ncl, idx_h2o, play, tlay, col_gas, jtemp, jpress,
                                                                                             tropo, prep, prep + ncl, prep + 2 * ncl, st.d);
    RB_LAUNCH_CHECK();
    RB_CUDA_CHECK(cudaMemcpyAsync(st.h, st.d, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream()));
  }
  FusedParams p{};
  p.t = t;
  p.ncol = ncol; p.nlay = nlay; p.band0 = 0; p.nband_sub = nbnd; p.gpt0 = 0;
  p.play = play; p.tlay = tlay;
  p.cs.col_dry = const_cast<Float*>(col_gas);   // slice 0 of col_gas(ncol,nlay,0:ngas); read-only in the ABI instantiations
  p.cs.pt_scale = prep; p.cs.vmr_fact = prep + ncl; p.cs.dry_fact = prep + 2 * ncl;
  p.cs.jtemp = const_cast<int*>(jtemp); p.cs.jpress = const_cast<int*>(jpress); p.cs.tropo = const_cast<Bool*>(tropo);
  p.op_kind = 1; p.tau = tau; p.cld_kind = 0; p.aer_kind = 0;
  p.abi_col_gas = col_gas; p.abi_col_mix = col_mix; p.abi_fmajor = fmajor; p.abi_fminor = fminor; p.abi_jeta = jeta;
  p.accumulate = accumulate ? 1 : 0;
  launch_tau(p, tt, /*abi=*/true);
  dev_free(prep);
  return true;
}

bool planck_source_gfast(int ncol, int nlay, int nbnd, int ngpt, int nflav, int neta, int npres, int ntemp, int nPlanckTemp,
                         const Float* tlay, const Float* tlev, const Float* tsfc, int sfc_lay, const Float* fmajor,
                         const int* jeta, const Bool* tropo, const int* jtemp, const int* jpress, const int* band_lims_gpt,
                         const Float* pfracin, Float temp_ref_min, Float totplnk_delta, const Float* totplnk,
                         const int* gpoint_flavor, Float* sfc_src, Float* lay_src, Float* lev_src, Float* sfc_source_Jac) {
  if (!g_abi_cache_on) return false;
  rrtmgpb_gas_tables t{};
  t.nflav = nflav; t.neta = neta; t.npres = npres; t.ntemp = ntemp; t.nbnd = nbnd; t.ngpt = ngpt;
  t.gpoint_flavor = gpoint_flavor; t.band_lims_gpt = band_lims_gpt; t.planck_frac = pfracin; t.totplnk = totplnk;
  t.nPlanckTemp = nPlanckTemp; t.totplnk_delta = totplnk_delta; t.temp_ref_min = temp_ref_min;
  const TablesT tt = tables_gfast(t, 1);
  PlanckFusedParams q{};
  FusedParams& p = q.f;
  p.t = t;
  p.ncol = ncol; p.nlay = nlay; p.band0 = 0; p.nband_sub = nbnd; p.gpt0 = 0;
  p.tlay = tlay;
  p.cs.jtemp = const_cast<int*>(jtemp); p.cs.jpress = const_cast<int*>(jpress); p.cs.tropo = const_cast<Bool*>(tropo);
  p.abi_fmajor = fmajor; p.abi_jeta = jeta;
  q.tlev = tlev; q.tsfc = tsfc; q.sfc_lay = sfc_lay;
  q.sfc_src = sfc_src; q.lay_src = lay_src; q.lev_src = lev_src; q.sfc_source_Jac = sfc_source_Jac;
  launch_planck(q, tt, /*abi=*/true);
  return true;
}
}  // namespace rrtmgpb

extern "C" {

void rrtmgpb_abi_table_cache(int on) { g_abi_cache_on = on ? 1 : 0; }
void rrtmgpb_set_gas_optics_rows_path(int on) { g_tau_rows.store(on); }

/* kept so that programs linked against earlier builds still resolve it: table staging is no longer selectable */
void rrtmgpb_set_tma_staging(int on) { (void)on; }

void rrtmgpb_gas_optics_fused(const rrtmgpb_gas_tables* t, int ncol, int nlay, const Float* play, const Float* plev,
                              const Float* tlay, const Float* vmr, const Float* col_dry, int op_kind, Float* tau,
                              Float* ssa, Float* g, int cld_kind, const Float* cld_tau, const Float* cld_ssa,
                              const Float* cld_g, int aer_kind, const Float* aer_tau, const Float* aer_ssa,
                              const Float* aer_g, const Float* tlev, const Float* tsfc, int sfc_lay, Float* sfc_src,
                              Float* lay_src, Float* lev_src, Float* sfc_source_Jac) {
  FusedParams p;
  p.t = *t;
  p.ncol = ncol; p.nlay = nlay; p.play = play; p.plev = plev; p.tlay = tlay; p.vmr = vmr; p.col_dry_in = col_dry;
  p.band0 = 0; p.nband_sub = t->nbnd; p.gpt0 = 0;
  p.op_kind = op_kind; p.tau = tau; p.ssa = ssa; p.g = g;
  p.cld_kind = cld_kind; p.cld_tau = cld_tau; p.cld_ssa = cld_ssa; p.cld_g = cld_g;
  p.aer_kind = aer_kind; p.aer_tau = aer_tau; p.aer_ssa = aer_ssa; p.aer_g = aer_g;
  const TablesT tt = tables_gfast(*t);
  Workspace w = prepare(p);
  launch_tau(p, tt);
  if (lay_src) {
    PlanckFusedParams q;
    q.f = p; q.tlev = tlev; q.tsfc = tsfc; q.sfc_lay = sfc_lay;
    q.sfc_src = sfc_src; q.lay_src = lay_src; q.lev_src = lev_src; q.sfc_source_Jac = sfc_source_Jac;
    launch_planck(q, tt);
  }
  dev_free(w.block);
}

}  // extern "C"

// ====================================================================================================
// Express path: broadband fluxes straight from the atmospheric state.  No (ncol, nlay, ngpt) array
// exists in HBM-sized form: columns are processed in chunks and, inside a chunk, a few bands at a time - gas optics
// (+ cloud increment, + Planck sources) write the planes of those bands into a scratch that is sized to stay in the
// 126 MB L2, and the register solver reads them straight back (TMA) and ADDS its spectrally integrated fluxes to the
// chunk's flux arrays (band after band, in band order: deterministic).  The chunk's g-points can be split over grid
// rows to fill the SMs; every row owns a copy of the flux arrays, summed in a fixed order at the end of the chunk.
// Kernels are the headline path's (gas_tau_g_kernel / planck_g_kernel on a band sub-range, *_reg_kernel with
// accumulate): per-g-point arithmetic is identical, only the association of the broadband sums differs
// (sum over bands of per-band sums).  What it buys: the footprint of a step drops from ~0.9 MB to ~6 KB per column
// (millions of columns stay resident) and the HBM traffic by the same factor.
// ====================================================================================================
namespace {

int env_int(const char* name, int dflt) {
  const char* e = std::getenv(name);
  return (e && *e) ? std::atoi(e) : dflt;
}

// dense copy of columns [c0, c0+n) of a Fortran (ncol, nrows) array, and back
void gather_cols(Float* dst, const Float* src, int ncol, int c0, int n, size_t nrows) {
  RB_CUDA_CHECK(cudaMemcpy2DAsync(dst, (size_t)n * sizeof(Float), src + c0, (size_t)ncol * sizeof(Float),
                                  (size_t)n * sizeof(Float), nrows, cudaMemcpyDeviceToDevice, stream()));
}

// out(c0 + i, l) = sum over the `groups` partial copies, in group order
__global__ void reduce_groups_kernel(int n, int nlev, int ncol, int c0, int groups, size_t stride, int narr,
                                     const Float* part, Float* o0, Float* o1, Float* o2) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t per = (size_t)n * nlev;
  if (k >= per) return;
  const int i = (int)(k % n), l = (int)(k / n);
  for (int a = 0; a < narr; ++a) {
    Float* out = a == 0 ? o0 : (a == 1 ? o1 : o2);
    if (!out) continue;
    const Float* q = part + (size_t)a * per + k;   // arrays of one group are contiguous: [group][array][n*nlev]
    Float acc = q[0];
    for (int g = 1; g < groups; ++g) acc = acc + q[(size_t)g * stride];
    out[(size_t)c0 + i + (size_t)ncol * l] = acc;
  }
}

struct ExpressPlan {
  int nc, bands_per_group, rows, nstreams;
};
// columns per chunk / bands per solver launch / grid rows / concurrent chunks.  Measured on B200 (65,536 x 72, LW+SW,
// profiles/r2_express_sweep.jsonl; the plane path takes 39.4 ms): launches whose scratch fits L2 (2,368 columns x 1
// band = 64 MB) leave most SMs idle - the Planck kernel has 76 blocks there - and the step takes 125 ms; 9,472 x 2
// bands 54 ms; 16,384 x 2 bands 52.8 ms on one stream, 47.7 ms with two chunks in flight on two streams (one chunk's
// tail overlaps the other's head); 32,768 x 4 bands on two streams 45.1 ms.  None of these kernels is HBM-bound (the
// solvers are fp64-bound, the gas optics L1-bound), so L2 residency buys nothing and small launches cost; what the
// express path buys is the footprint (2.9 GB instead of 54 GB at this size).  RRTMGPB_EXPRESS_NC / _BANDS / _ROWS /
// _STREAMS override.
ExpressPlan express_plan(int ncol, int nlay) {
  ExpressPlan pl;
  pl.nc = env_int("RRTMGPB_EXPRESS_NC", 32768);
  pl.bands_per_group = env_int("RRTMGPB_EXPRESS_BANDS", 4);
  pl.rows = env_int("RRTMGPB_EXPRESS_ROWS", 1);
  pl.nstreams = std::max(1, std::min(2, env_int("RRTMGPB_EXPRESS_STREAMS", 2)));
  if (pl.nc > ncol) pl.nc = ncol;
  if (pl.nc < 1) pl.nc = 1;
  if (pl.nc % 2) pl.nc += (pl.nc < ncol) ? 1 : 0;  // even chunk widths keep the scratch planes TMA-describable (16-byte row stride)
  (void)nlay;
  return pl;
}

}  // namespace

extern "C" {

int rrtmgpb_express_supported(int ncol, int nlay) {
  (void)ncol;
  return (rrtmgpb_get_solver_variant() == 0 && nlay <= 144) ? 1 : 0;   // register solvers (accumulate mode)
}

// see include/rrtmgp_b200_ext.h
void rrtmgpb_express(const rrtmgpb_gas_tables* t, int ncol, int nlay, int top_at_1, const Float* play, const Float* plev,
                     const Float* tlay, const Float* tlev, const Float* tsfc, const Float* vmr, const Float* col_dry,
                     int cld_kind, const Float* cld_tau, const Float* cld_ssa, const Float* cld_g,
                     const Float* sfc_emis_or_alb_dir, const Float* sfc_alb_dif, const Float* mu0,
                     const Float* solar_source, int nmus, const Float* Ds_host, const Float* wts_host, Float* flux_up,
                     Float* flux_dn, Float* flux_dir) {
  OpName op_name__(__func__);
  const bool sw = t->krayl != nullptr;
  const int ngpt = t->ngpt, nbnd = t->nbnd, ngas = t->ngas, nlev = nlay + 1;
  // the caller's arrays: used in place when they live on the device, staged once per call otherwise
  const size_t Ncl = (size_t)ncol * nlay, Nclp = (size_t)ncol * nlev;
  DevArg<Float> a_play(play, Ncl, Dir::In), a_plev(plev, Nclp, Dir::In), a_tlay(tlay, Ncl, Dir::In),
      a_tlev(tlev, Nclp, Dir::In, !sw), a_tsfc(tsfc, ncol, Dir::In, !sw), a_vmr(vmr, Ncl * ngas, Dir::In),
      a_cd(col_dry, Ncl, Dir::In, col_dry != nullptr), a_ct(cld_tau, Ncl * nbnd, Dir::In, cld_kind != 0),
      a_cw(cld_ssa, Ncl * nbnd, Dir::In, cld_kind == 2), a_cg(cld_g, Ncl * nbnd, Dir::In, cld_kind == 2 && sw),
      a_sa(sfc_emis_or_alb_dir, (size_t)nbnd * ncol, Dir::In), a_sb(sfc_alb_dif, (size_t)nbnd * ncol, Dir::In, sw),
      a_mu0(mu0, ncol, Dir::In, sw), a_sol(solar_source, ngpt, Dir::In, sw);
  DevArg<Float> a_fu(flux_up, Nclp, Dir::Out), a_fd(flux_dn, Nclp, Dir::Out), a_fr(flux_dir, Nclp, Dir::Out, sw);
  play = a_play; plev = a_plev; tlay = a_tlay; tlev = a_tlev; tsfc = a_tsfc; vmr = a_vmr; col_dry = a_cd;
  cld_tau = a_ct; cld_ssa = a_cw; cld_g = a_cg; sfc_emis_or_alb_dir = a_sa; sfc_alb_dif = a_sb; mu0 = a_mu0;
  solar_source = a_sol; flux_up = a_fu; flux_dn = a_fd; flux_dir = a_fr;
  struct Trust { Trust() { tl_trust_device_ptrs = true; } ~Trust() { tl_trust_device_ptrs = false; } } trust__;  // see DevArg
  const TablesT tt = tables_gfast(*t);
  const std::vector<int> bl = to_host(t->band_lims_gpt, 2 * (size_t)nbnd);
  const bool staged = rrtmgpb_express_supported(ncol, nlay) != 0;   // else: one launch per chunk over all bands
  ExpressPlan pl = express_plan(ncol, nlay);
  if (!staged) { pl.bands_per_group = nbnd; pl.rows = 1; }
  const int nc = pl.nc;
  int ng_max = 0;
  for (int b0 = 0; b0 < nbnd; b0 += pl.bands_per_group) {
    const int b1 = std::min(nbnd, b0 + pl.bands_per_group);
    ng_max = std::max(ng_max, bl[2 * (b1 - 1) + 1] - bl[2 * b0] + 1);
  }
  const int narr = sw ? 3 : 2;
  const size_t ncl = (size_t)nc * nlay, nclp = (size_t)nc * nlev;
  // ---- scratch, allocated once per concurrent chunk (stream-ordered pool, on the caller's stream)
  const size_t gstride = (size_t)narr * nclp;
  const int nslots = (ncol > nc) ? pl.nstreams : 1;
  struct Slot { Float *planes, *in2d, *bnd, *part, *decoy, *mu0_lay; } slots[2] = {};
  for (int k = 0; k < nslots; ++k) {
    Slot& sl = slots[k];
    sl.planes = static_cast<Float*>(dev_alloc((ncl * 2 + nclp) * (size_t)ng_max * sizeof(Float)));  // SW tau, ssa, g; LW tau, lay, lev
    sl.in2d = static_cast<Float*>(dev_alloc((ncl * (3 + (size_t)ngas + 3 * (size_t)nbnd) + 2 * nclp) * sizeof(Float)));
    sl.bnd = static_cast<Float*>(dev_alloc((size_t)nc * ngpt * 4 * sizeof(Float) + (size_t)nc * ng_max * (nmus + 2) * sizeof(Float)));
    sl.part = static_cast<Float*>(dev_alloc(gstride * pl.rows * sizeof(Float)));
    sl.decoy = static_cast<Float*>(dev_alloc(nclp * sizeof(Float)));
    sl.mu0_lay = static_cast<Float*>(dev_alloc(ncl * sizeof(Float)));
  }
  Float* wts_dev = static_cast<Float*>(dev_alloc((size_t)std::max(nmus, 1) * sizeof(Float)));
  if (!sw) RB_CUDA_CHECK(cudaMemcpyAsync(wts_dev, wts_host, (size_t)nmus * sizeof(Float), cudaMemcpyHostToDevice, stream()));
  // ---- chunks alternate between the caller's stream and an auxiliary one (fork / join by events): the small grids at
  // the end of one chunk's kernels overlap the next chunk's
  cudaStream_t s_main = stream(), s_aux = s_main;
  static thread_local cudaStream_t tl_aux = nullptr;
  static thread_local cudaEvent_t tl_fork = nullptr, tl_join = nullptr;
  if (nslots > 1) {
    if (!tl_aux) {
      RB_CUDA_CHECK(cudaStreamCreateWithFlags(&tl_aux, cudaStreamNonBlocking));
      RB_CUDA_CHECK(cudaEventCreateWithFlags(&tl_fork, cudaEventDisableTiming));
      RB_CUDA_CHECK(cudaEventCreateWithFlags(&tl_join, cudaEventDisableTiming));
    }
    s_aux = tl_aux;
    RB_CUDA_CHECK(cudaEventRecord(tl_fork, s_main));
    RB_CUDA_CHECK(cudaStreamWaitEvent(s_aux, tl_fork, 0));
  }

  int ichunk = 0;
  for (int c0 = 0; c0 < ncol; c0 += nc, ++ichunk) {
    const Slot& sl = slots[ichunk % nslots];
    rrtmgpb_set_stream((ichunk % nslots) ? s_aux : s_main);   // every launch below goes to this chunk's stream
    Float *planes = sl.planes, *in2d = sl.in2d, *bnd = sl.bnd, *part = sl.part, *decoy = sl.decoy, *mu0_lay = sl.mu0_lay;
    // chunk-local inputs
    Float* c_play = in2d; Float* c_tlay = c_play + ncl; Float* c_cd = c_tlay + ncl; Float* c_vmr = c_cd + ncl;
    Float* c_ct = c_vmr + ncl * ngas; Float* c_cw = c_ct + ncl * nbnd; Float* c_cg = c_cw + ncl * nbnd;
    Float* c_plev = c_cg + ncl * nbnd; Float* c_tlev = c_plev + nclp;
    // boundary arrays of the chunk on all g-points: a g-point sub-range of an (nc, ngpt) array is a contiguous slab
    Float* b_a = bnd; Float* b_b = b_a + (size_t)nc * ngpt; Float* b_toa = b_b + (size_t)nc * ngpt; Float* b_zero = b_toa + (size_t)nc * ngpt;
    Float* b_Ds = b_zero + (size_t)nc * ngpt; Float* b_sfc = b_Ds + (size_t)nc * ng_max * nmus; Float* b_jac = b_sfc + (size_t)nc * ng_max;
    const int n = std::min(nc, ncol - c0);
    const size_t nl = (size_t)n * nlay, nlp = (size_t)n * nlev;
    // ---- gather the chunk's columns (dense (n, nlay[, k]) copies of the strided slices)
    gather_cols(c_play, play, ncol, c0, n, nlay);
    gather_cols(c_tlay, tlay, ncol, c0, n, nlay);
    gather_cols(c_plev, plev, ncol, c0, n, nlev);
    gather_cols(c_vmr, vmr, ncol, c0, n, (size_t)nlay * ngas);
    if (col_dry) gather_cols(c_cd, col_dry, ncol, c0, n, nlay);
    if (cld_kind) {
      gather_cols(c_ct, cld_tau, ncol, c0, n, (size_t)nlay * nbnd);
      if (cld_kind == 2) {
        gather_cols(c_cw, cld_ssa, ncol, c0, n, (size_t)nlay * nbnd);
        if (sw) gather_cols(c_cg, cld_g, ncol, c0, n, (size_t)nlay * nbnd);
      }
    }
    if (!sw) gather_cols(c_tlev, tlev, ncol, c0, n, nlev);
    // ---- per-cell state of the chunk
    FusedParams p;
    p.t = *t;
    p.ncol = n; p.nlay = nlay; p.play = c_play; p.plev = c_plev; p.tlay = c_tlay; p.vmr = c_vmr;
    p.col_dry_in = col_dry ? c_cd : nullptr;
    p.op_kind = sw ? 2 : 1;
    p.cld_kind = cld_kind; p.cld_tau = c_ct; p.cld_ssa = c_cw; p.cld_g = c_cg;
    p.aer_kind = 0; p.aer_tau = p.aer_ssa = p.aer_g = nullptr;
    Workspace w = prepare(p);
    // ---- boundary conditions of the chunk (mo_rte_lw.F90:264-282, mo_rte_sw.F90:266-280)
    rrtmgpb_expand_and_transpose(n, nbnd, ngpt, t->band_lims_gpt, sfc_emis_or_alb_dir + (size_t)nbnd * c0, b_a);
    RB_CUDA_CHECK(cudaMemsetAsync(b_zero, 0, (size_t)n * ngpt * sizeof(Float), stream()));
    if (sw) {
      rrtmgpb_expand_and_transpose(n, nbnd, ngpt, t->band_lims_gpt, sfc_alb_dif + (size_t)nbnd * c0, b_b);
      rrtmgpb_broadcast_by_gpt(n, ngpt, solar_source, b_toa);
      rrtmgpb_broadcast_by_lay(n, nlay, mu0 + c0, mu0_lay);
    }
    RB_CUDA_CHECK(cudaMemsetAsync(part, 0, gstride * pl.rows * sizeof(Float), stream()));
    int ng_last = -1;
    for (int b0 = 0; b0 < nbnd; b0 += pl.bands_per_group) {
      const int b1 = std::min(nbnd, b0 + pl.bands_per_group);
      const int g0 = bl[2 * b0] - 1, ng = bl[2 * (b1 - 1) + 1] - g0;   // 0-based first g-point, count
      p.band0 = b0; p.nband_sub = b1 - b0; p.gpt0 = g0;
      Float* s_tau = planes; Float* s_b = s_tau + nl * ng; Float* s_c = s_b + nl * ng;
      p.tau = s_tau; p.ssa = sw ? s_b : nullptr; p.g = sw ? s_c : nullptr;
      launch_tau(p, tt);
      const Bool top = top_at_1 != 0, yes = 1, no = 0;
      tl_express.accumulate = 1; tl_express.groups = staged ? pl.rows : 1; tl_express.group_stride = gstride;
      if (!staged) tl_express.accumulate = 0;
      if (sw) {
        rte_sw_solver_2stream(&n, &nlay, &ng, &top, s_tau, s_b, s_c, mu0_lay, b_a + (size_t)n * g0, b_b + (size_t)n * g0,
                              b_toa + (size_t)n * g0, decoy, decoy, decoy, &no, b_zero, &yes, part, part + nlp, part + 2 * nlp);
      } else {
        PlanckFusedParams q;
        q.f = p; q.tlev = c_tlev; q.tsfc = tsfc + c0; q.sfc_lay = top_at_1 ? nlay : 1;
        q.sfc_src = b_sfc; q.lay_src = s_b; q.lev_src = s_c; q.sfc_source_Jac = b_jac;
        launch_planck(q, tt);
        if (ng != ng_last) {   // secants: the same value for every column and g-point of an angle (mo_rte_lw.F90:357-365)
          for (int imu = 0; imu < nmus; ++imu) {
            const int n2 = ng;
            set_to_scalar_2D(&n, &n2, b_Ds + (size_t)n * ng * imu, &Ds_host[imu]);
          }
          ng_last = ng;
        }
        rte_lw_solver_noscat(&n, &nlay, &ng, &top, &nmus, b_Ds, wts_dev, s_tau, s_b, s_c, b_a + (size_t)n * g0, b_sfc,
                             b_zero, decoy, decoy, &yes, part, part + nlp, &no, b_jac, decoy, &no, s_tau, s_tau);
      }
      tl_express = ExpressSolverMode();
    }
    // ---- sum the grid rows' copies (fixed order) into the caller's arrays
    {
      KernelTimer timer("express_reduce");
      // partial layout per row: [array a][n*nlev]; rows are gstride apart (sized for nc columns)
      reduce_groups_kernel<<<ceil_div((long long)nlp, 256), 256, 0, stream()>>>(
          n, nlev, ncol, c0, staged ? pl.rows : 1, gstride, narr, part, flux_up, flux_dn, sw ? flux_dir : nullptr);
      RB_LAUNCH_CHECK();
    }
    dev_free(w.block);
  }
  rrtmgpb_set_stream(s_main);
  if (nslots > 1) {   // join: the caller's stream continues after the auxiliary stream's chunks
    RB_CUDA_CHECK(cudaEventRecord(tl_join, s_aux));
    RB_CUDA_CHECK(cudaStreamWaitEvent(s_main, tl_join, 0));
  }
  dev_free(wts_dev);
  for (int k = 0; k < nslots; ++k) {
    dev_free(slots[k].mu0_lay); dev_free(slots[k].decoy); dev_free(slots[k].part); dev_free(slots[k].bnd);
    dev_free(slots[k].in2d); dev_free(slots[k].planes);
  }
}

}  // extern "C"
