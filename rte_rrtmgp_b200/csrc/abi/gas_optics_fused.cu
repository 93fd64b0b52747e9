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
#include <map>
#include <mutex>
#include <vector>
#include "../kernels/elementwise.cuh"
#include "../kernels/gas_optics_gfast.cuh"
#include "rrtmgp_b200_ext.h"

using namespace rrtmgpb;

namespace {

constexpr int kFThreads = 128;

// ---- per-cell state: mo_gas_optics_utils.F90:143-150 (col_dry), mo_gas_optics_rrtmgp_kernels.F90:99-118,
// and the per-cell factors of the minor scaling :467-471 ----
__global__ void __launch_bounds__(kFThreads) cell_state_kernel(const FusedParams p, Float m_dry, Float m_h2o,
                                                                Float avogad, Float grav) {
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
    cell_state_kernel<<<ceil_div((long long)ncl, kFThreads), kFThreads, 0, stream()>>>(
        p, (Float)g_m_dry, (Float)k_m_h2o, (Float)k_avogad, (Float)g_grav);
    RB_LAUNCH_CHECK();
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
std::map<const void*, TableCacheEntry> g_table_cache;  // key: the loader-layout kmajor pointer

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

TablesT tables_gfast(const rrtmgpb_gas_tables& t) {
  std::lock_guard<std::mutex> lock(g_tc_mutex);
  auto it = g_table_cache.find(t.kmajor);
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
  const std::vector<int> fl = to_host(t.flavor, 2 * (size_t)t.nflav);
  const std::vector<Float> vr = to_host(t.vmr_ref, 2 * (size_t)(t.ngas + 1) * t.ntemp);
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
  g_table_cache[t.kmajor] = e;
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
  auto it = g_table_cache.find(key);
  if (it == g_table_cache.end()) return;
  for (void* q : it->second.owned) dev_free(q);
  g_table_cache.erase(it);
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
  for (auto& kv : g_table_cache) for (void* q : kv.second.owned) dev_free(q);
  for (auto& kv : g_abi_cache) for (void* q : kv.second.owned) dev_free(q);
  g_table_cache.clear();
  g_abi_cache.clear();
}

extern "C" {

void rrtmgpb_abi_table_cache(int on) { g_abi_cache_on = on ? 1 : 0; }

/* kept so that programs linked against earlier builds still resolve it: table staging is no longer selectable */
void rrtmgpb_set_tma_staging(int on) { (void)on; }

void rrtmgpb_gas_optics_fused(const rrtmgpb_gas_tables* t, int ncol, int nlay, const Float* play, const Float* plev,
                              const Float* tlay, const Float* vmr, const Float* col_dry, int op_kind, Float* tau,
                              Float* ssa, Float* g, int cld_kind, const Float* cld_tau, const Float* cld_ssa,
                              const Float* cld_g, int aer_kind, const Float* aer_tau, const Float* aer_ssa,
                              const Float* aer_g, const Float* tlev, const Float* tsfc, int sfc_lay, Float* sfc_src,
                              Float* lay_src, Float* lev_src, Float* sfc_source_Jac) {
  const size_t ncl = (size_t)ncol * nlay;
  FusedParams p;
  p.t = *t;
  p.ncol = ncol; p.nlay = nlay; p.play = play; p.plev = plev; p.tlay = tlay; p.vmr = vmr; p.col_dry_in = col_dry;
  p.op_kind = op_kind; p.tau = tau; p.ssa = ssa; p.g = g;
  p.cld_kind = cld_kind; p.cld_tau = cld_tau; p.cld_ssa = cld_ssa; p.cld_g = cld_g;
  p.aer_kind = aer_kind; p.aer_tau = aer_tau; p.aer_ssa = aer_ssa; p.aer_g = aer_g;
  const TablesT tt = tables_gfast(*t);
  Workspace w = prepare(p);
  const bool sw = t->krayl != nullptr;
  {
    KernelTimer timer(sw ? "gas_tau_fused[sw]" : "gas_tau_fused[lw]");
    const unsigned grid = (unsigned)((long long)ceil_div((long long)ncl, kTauCells * kGThreads) * t->nbnd);
    // lane-private slots for the per-(cell, band) minor scalings: [contributor][cell slot][thread]
    const size_t smem = (size_t)tt.maxm * kTauCells * kGThreads * sizeof(Float);
// KIND 1: the common kinds as compile-time constants (LW 1scl += 1scl clouds, SW 2str += 2str clouds, no aerosols)
#define GAS_TAU_LAUNCH1(SWV, VECV, AERV, KINDV)                                                                   \
  do {                                                                                                            \
    auto kern = gas_tau_g_kernel<SWV, VECV, AERV, KINDV>;                                                         \
    if (smem > 48 * 1024) RB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<grid, kGThreads, smem, stream()>>>(p, tt);                                                             \
  } while (0)
#define GAS_TAU_LAUNCH(SWV, VECV)                                                                     \
  if (aer_kind) GAS_TAU_LAUNCH1(SWV, VECV, true, 0);                                                  \
  else if (op_kind == (SWV ? 2 : 1) && cld_kind == (SWV ? 2 : 1)) GAS_TAU_LAUNCH1(SWV, VECV, false, 1); \
  else GAS_TAU_LAUNCH1(SWV, VECV, false, 0)
    if (sw) {
      if (tt.vec == 2) { GAS_TAU_LAUNCH(true, 2); } else { GAS_TAU_LAUNCH(true, 1); }
    } else {
      if (tt.vec == 2) { GAS_TAU_LAUNCH(false, 2); } else { GAS_TAU_LAUNCH(false, 1); }
    }
#undef GAS_TAU_LAUNCH
#undef GAS_TAU_LAUNCH1
    RB_LAUNCH_CHECK();
  }
  if (lay_src) {
    PlanckFusedParams q;
    q.f = p; q.tlev = tlev; q.tsfc = tsfc; q.sfc_lay = sfc_lay;
    q.sfc_src = sfc_src; q.lay_src = lay_src; q.lev_src = lev_src; q.sfc_source_Jac = sfc_source_Jac;
    KernelTimer timer("planck_fused");
    // layers a thread marches through (its first level needs the Planck fractions of the layer above: 1/lay_per_chunk
    // redundant work; B200, 65,536 x 72: 9 -> 5.40 ms, 12 -> 5.31, 18 -> 5.24, 36 -> 5.22); RRTMGPB_PLANCK_CHUNK overrides
    static const int chunk_env = [] { const char* e = std::getenv("RRTMGPB_PLANCK_CHUNK"); return e ? std::atoi(e) : 0; }();
    const int lay_per_chunk = chunk_env > 0 ? chunk_env : 18, nchunk = ceil_div(nlay, lay_per_chunk);
    const unsigned grid = (unsigned)((long long)ceil_div(ncol, kGThreads) * nchunk * t->nbnd);
    if (tt.vec == 2) planck_g_kernel<2><<<grid, kGThreads, 0, stream()>>>(q, tt, lay_per_chunk, nchunk);
    else planck_g_kernel<1><<<grid, kGThreads, 0, stream()>>>(q, tt, lay_per_chunk, nchunk);
    RB_LAUNCH_CHECK();
  }
  dev_free(w.block);
}

}  // extern "C"
