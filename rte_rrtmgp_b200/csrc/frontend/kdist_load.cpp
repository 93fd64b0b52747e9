// kdist_load.cpp - ty_gas_optics_rrtmgp%load without Fortran or netCDF (include/rrtmgp_b200_kdist.h).
//
// Host-only restatement of rrtmgp/frontend/mo_gas_optics_rrtmgp.F90: load_int :938-1030, load_ext :1038-1145,
// init_abs_coeffs :1151-1381, check_key_species_present_init :1383-1397, rewrite_key_species_pair :1568-1576,
// create_flavor :1598-1632, create_idx_minor :1637-1657, create_idx_minor_scaling :1661-1675,
// create_key_species_reduce :1752-1786, reduce_minor_arrays :1790-1907, create_gpoint_flavor :1930-1946,
// set_solar_variability :760-798, set_tsi :800-835; ty_optical_props%init (band -> g-point map),
// rte/frontend/mo_optical_props.F90:240-302; string_loc_in_array, mo_gas_optics_util_string.F90:71-87.
// No array data leaves the host here: the result is the rrtmgpb_kdist that rrtmgpb_gas_optics_load uploads.
// Compiled into the product library and (the same file) into the oracle, like frontend.cpp.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "rrtmgp_b200_kdist.h"

namespace {

int fail(char* errmsg, const std::string& msg) {
  if (errmsg) { std::snprintf(errmsg, 128, "%s", msg.c_str()); }
  return 1;
}

// lower_case(trim(s)); Fortran character variables are blank padded, so trailing blanks never matter
std::string lc_trim(const char* s) {
  std::string r = s ? s : "";
  while (!r.empty() && (r.back() == ' ' || r.back() == '\0')) r.pop_back();
  for (char& c : r) c = (char)std::tolower((unsigned char)c);
  return r;
}
// string_loc_in_array :71-87: 1-based position, -1 when absent
int loc_in(const std::string& s, const std::vector<std::string>& arr) {
  const std::string key = lc_trim(s.c_str());
  for (size_t i = 0; i < arr.size(); ++i)
    if (key == lc_trim(arr[i].c_str())) return (int)i + 1;
  return -1;
}
std::vector<std::string> strings(const char* const* p, int n) {
  std::vector<std::string> v((size_t)std::max(n, 0));
  for (int i = 0; i < n; ++i) v[(size_t)i] = p && p[i] ? p[i] : "";
  return v;
}

struct MinorRed {  // one atmosphere half after reduce_minor_arrays
  std::vector<Float> kminor;  // (ntemp, neta, tot_g)
  std::vector<std::string> gases, scaling_gas;
  std::vector<int> limits, kstart, idx_minor, idx_scaling;
  std::vector<unsigned char> dens, comp;  // Bool is one byte (rte_types.h); std::vector<bool> has no data()
  int tot_g = 0;
};

// reduce_minor_arrays :1790-1907.  kminor_atm(ncontrib, neta, ntemp) as on disk -> kminor_red(ntemp, neta, tot_g)
MinorRed reduce_minor(const std::vector<std::string>& available, const std::vector<std::string>& gas_minor,
                      const std::vector<std::string>& identifier_minor, const Float* kminor_atm, int ncontrib, int neta,
                      int ntemp, const std::vector<std::string>& minor_gases_atm, const int* limits, const Bool* dens,
                      const std::vector<std::string>& scaling_gas_atm, const Bool* comp, const int* kstart) {
  MinorRed r;
  const int nm = (int)minor_gases_atm.size();
  std::vector<char> present((size_t)nm, 0);
  for (int i = 0; i < nm; ++i) {
    const int idx_mnr = loc_in(minor_gases_atm[(size_t)i], identifier_minor);           // :1850
    present[(size_t)i] = idx_mnr > 0 && loc_in(gas_minor[(size_t)idx_mnr - 1], available) > 0;   // :1853
    if (present[(size_t)i]) r.tot_g += limits[2 * i + 1] - limits[2 * i] + 1;           // :1855
  }
  std::vector<Float> red_t;  // kminor_atm_red_t(tot_g, neta, ntemp)
  const size_t plane = (size_t)neta * ntemp;
  red_t.assign((size_t)r.tot_g * plane, (Float)0);
  int n_elim = 0;
  for (int i = 0; i < nm; ++i) {                                                         // :1884-1900
    const int ng = limits[2 * i + 1] - limits[2 * i] + 1;
    if (!present[(size_t)i]) { n_elim += ng; continue; }
    r.gases.push_back(minor_gases_atm[(size_t)i]);
    r.scaling_gas.push_back(scaling_gas_atm[(size_t)i]);
    r.dens.push_back(dens[i] ? 1 : 0);
    r.comp.push_back(comp[i] ? 1 : 0);
    r.limits.push_back(limits[2 * i]);
    r.limits.push_back(limits[2 * i + 1]);
    const int ks = kstart[i] - n_elim;                                                   // :1890
    r.kstart.push_back(ks);
    for (int j = 0; j < ng; ++j)                                                         // :1892-1895
      for (size_t q = 0; q < plane; ++q)
        red_t[(size_t)(ks + j - 1) + (size_t)r.tot_g * q] = kminor_atm[(size_t)(kstart[i] + j - 1) + (size_t)ncontrib * q];
  }
  // RESHAPE(..., ORDER=(/3,2,1/)) :1904: red(it, ie, k) = red_t(k, ie, it)
  r.kminor.assign((size_t)std::max(r.tot_g, 1) * plane, (Float)0);
  for (int k = 0; k < r.tot_g; ++k)
    for (int ie = 0; ie < neta; ++ie)
      for (int it = 0; it < ntemp; ++it)
        r.kminor[(size_t)it + (size_t)ntemp * (ie + (size_t)neta * k)] = red_t[(size_t)k + (size_t)r.tot_g * (ie + (size_t)neta * it)];
  return r;
}

}  // namespace

struct rrtmgpb_kdist_loaded {
  rrtmgpb_kdist t;
  std::vector<std::string> gas_names;
  std::vector<char> is_key;
  std::vector<int> flavor, gpoint_flavor, band_lims_gpt, gpoint_bands;
  std::vector<Float> band_lims_wvn, press_ref, press_ref_log, temp_ref, vmr_ref, kmajor, planck_frac, totplnk, krayl,
      optimal_angle_fit, solar_quiet, solar_facular, solar_sunspot, solar_source;
  MinorRed lower, upper;
};

extern "C" {

rrtmgpb_kdist_loaded* rrtmgpb_kdist_reduce(const rrtmgpb_kdist_raw* raw, int navailable,
                                           const char* const* available_gases, char* errmsg) {
  if (errmsg) errmsg[0] = 0;
  if (!raw || !raw->kmajor) { fail(errmsg, "ERROR: spectral configuration not loaded"); return nullptr; }
  const bool lw = raw->totplnk != nullptr;
  if (lw == (raw->solar_source_quiet != nullptr)) {
    fail(errmsg, "gas_optics%load: provide either the Planck tables (LW) or the solar source (SW)");
    return nullptr;
  }
  const int ntemp = raw->ntemp, npres = raw->npres, neta = raw->nmixingfracs, nbnd = raw->nbnd, ngpt = raw->ngpt;
  auto kd = new rrtmgpb_kdist_loaded();
  auto bail = [&](const std::string& m) { fail(errmsg, m); delete kd; return (rrtmgpb_kdist_loaded*)nullptr; };

  // ---- ty_optical_props%init(band_lims_wavenum, band2gpt): mo_optical_props.F90:254-300
  for (int i = 0; i < 2 * nbnd; ++i)
    if (raw->bnd_limits_wavenumber[i] < 0) return bail("optical_props%init(): band_lims_wvn has values <  0., respectively");
  for (int i = 0; i < 2 * nbnd; ++i)
    if (raw->bnd_limits_gpt[i] < 1) return bail("optical_props%init(): band_lims_gpt has values < 1");
  kd->band_lims_wvn.assign(raw->bnd_limits_wavenumber, raw->bnd_limits_wavenumber + 2 * nbnd);
  kd->band_lims_gpt.assign(raw->bnd_limits_gpt, raw->bnd_limits_gpt + 2 * nbnd);
  kd->gpoint_bands.assign((size_t)ngpt, 0);   // gpt2band :292-299
  for (int b = 0; b < nbnd; ++b)
    for (int g = kd->band_lims_gpt[2 * b]; g <= kd->band_lims_gpt[2 * b + 1] && g <= ngpt; ++g) kd->gpoint_bands[(size_t)g - 1] = b + 1;

  // ---- which gases of the k-distribution does the host provide?  :1219-1233
  const std::vector<std::string> gas_names = strings(raw->gas_names, raw->nabsorbers);
  const std::vector<std::string> available = strings(available_gases, navailable);
  for (const std::string& g : gas_names)
    if (loc_in(g, available) > 0) kd->gas_names.push_back(g);
  const int ngas = (int)kd->gas_names.size();

  // ---- vmr_ref(2, 0:ngas, ntemp)  :1236-1245 ; on disk (atmos_layer, absorber_ext, temperature)
  const int next = raw->nextabsorbers;
  kd->vmr_ref.assign((size_t)2 * (ngas + 1) * ntemp, (Float)0);
  for (int it = 0; it < ntemp; ++it)
    for (int a = 0; a < 2; ++a) {
      kd->vmr_ref[(size_t)a + 2 * ((size_t)0 + (size_t)(ngas + 1) * it)] = raw->vmr_ref[(size_t)a + 2 * ((size_t)0 + (size_t)next * it)];
      for (int i = 1; i <= ngas; ++i) {
        const int idx = loc_in(kd->gas_names[(size_t)i - 1], gas_names);   // 1-based in gas_names -> absorber_ext idx+1
        kd->vmr_ref[(size_t)a + 2 * ((size_t)i + (size_t)(ngas + 1) * it)] = raw->vmr_ref[(size_t)a + 2 * ((size_t)idx + (size_t)next * it)];
      }
    }

  // ---- minor contributors  :1250-1282
  const std::vector<std::string> gas_minor = strings(raw->gas_minor, raw->nminorabsorbers);
  const std::vector<std::string> identifier_minor = strings(raw->identifier_minor, raw->nminorabsorbers);
  kd->lower = reduce_minor(available, gas_minor, identifier_minor, raw->kminor_lower, raw->ncontributors_lower, neta, ntemp,
                           strings(raw->minor_gases_lower, raw->nminor_absorber_intervals_lower), raw->minor_limits_gpt_lower,
                           raw->minor_scales_with_density_lower, strings(raw->scaling_gas_lower, raw->nminor_absorber_intervals_lower),
                           raw->scale_by_complement_lower, raw->kminor_start_lower);
  kd->upper = reduce_minor(available, gas_minor, identifier_minor, raw->kminor_upper, raw->ncontributors_upper, neta, ntemp,
                           strings(raw->minor_gases_upper, raw->nminor_absorber_intervals_upper), raw->minor_limits_gpt_upper,
                           raw->minor_scales_with_density_upper, strings(raw->scaling_gas_upper, raw->nminor_absorber_intervals_upper),
                           raw->scale_by_complement_upper, raw->kminor_start_upper);

  // ---- arrays not reduced: kmajor(gpt, eta, p+1, T) -> (T, eta, p+1, gpt)  :1296-1299
  auto to_kernel_layout = [&](const Float* src) {
    std::vector<Float> out((size_t)ntemp * neta * (npres + 1) * ngpt);
    for (int g = 0; g < ngpt; ++g)
      for (int ip = 0; ip <= npres; ++ip)
        for (int ie = 0; ie < neta; ++ie)
          for (int it = 0; it < ntemp; ++it)
            out[(size_t)it + (size_t)ntemp * (ie + (size_t)neta * (ip + (size_t)(npres + 1) * g))] =
                src[(size_t)g + (size_t)ngpt * (ie + (size_t)neta * (ip + (size_t)(npres + 1) * it))];
    return out;
  };
  kd->press_ref.assign(raw->press_ref, raw->press_ref + npres);
  kd->temp_ref.assign(raw->temp_ref, raw->temp_ref + ntemp);
  kd->kmajor = to_kernel_layout(raw->kmajor);
  if ((raw->rayl_lower != nullptr) != (raw->rayl_upper != nullptr))
    return bail("rayl_lower and rayl_upper must have the same allocation status");   // :1303-1306
  if (raw->rayl_lower) {  // krayl(T, eta, gpt, 2) from rayl_*(gpt, eta, T)  :1307-1316
    kd->krayl.assign((size_t)ntemp * neta * ngpt * 2, (Float)0);
    for (int a = 0; a < 2; ++a) {
      const Float* src = a ? raw->rayl_upper : raw->rayl_lower;
      for (int g = 0; g < ngpt; ++g)
        for (int ie = 0; ie < neta; ++ie)
          for (int it = 0; it < ntemp; ++it)
            kd->krayl[(size_t)it + (size_t)ntemp * (ie + (size_t)neta * (g + (size_t)ngpt * a))] =
                src[(size_t)g + (size_t)ngpt * (ie + (size_t)neta * it)];
    }
  }
  // ---- post processing  :1320-1326
  kd->press_ref_log.resize((size_t)npres);
  for (int i = 0; i < npres; ++i) kd->press_ref_log[(size_t)i] = std::log(kd->press_ref[(size_t)i]);

  // ---- index of the gas behind every minor contributor / its scaling gas  :1329-1333
  for (MinorRed* m : {&kd->lower, &kd->upper}) {
    for (size_t i = 0; i < m->gases.size(); ++i) {
      const int idx_mnr = loc_in(m->gases[i], identifier_minor);
      m->idx_minor.push_back(loc_in(gas_minor[(size_t)idx_mnr - 1], kd->gas_names));   // :1654
      m->idx_scaling.push_back(loc_in(m->scaling_gas[i], kd->gas_names));               // :1672, -1: no interacting gas
    }
  }

  // ---- key species -> flavours  :1340-1349
  const int nks = 2 * 2 * nbnd;
  std::vector<int> ks_red((size_t)nks);
  std::vector<char> key_present(gas_names.size(), 1);
  for (int i = 0; i < nks; ++i) {   // create_key_species_reduce :1770-1783
    const int k = raw->key_species[i];
    if (k != 0) {
      ks_red[(size_t)i] = loc_in(gas_names[(size_t)k - 1], kd->gas_names);
      if (ks_red[(size_t)i] == -1) key_present[(size_t)k - 1] = 0;
    } else {
      ks_red[(size_t)i] = 0;
    }
  }
  {  // check_key_species_present_init :1389-1395 (the message lists the missing gases in reverse order)
    std::string missing;
    for (size_t i = 0; i < key_present.size(); ++i)
      if (!key_present[i]) {  // ' ' // trim(gas_names(i)) // trim(err_message)
        std::string g = gas_names[i];
        while (!g.empty() && g.back() == ' ') g.pop_back();
        missing = " " + g + missing;
      }
    if (!missing.empty()) return bail("gas_optics: required gases" + missing + " are not provided");
  }
  auto rewrite = [](int a, int b, int* out) {   // :1568-1576: (0,0) -> (2,2)
    if (a == 0 && b == 0) { out[0] = 2; out[1] = 2; } else { out[0] = a; out[1] = b; }
  };
  std::vector<int>& flavor = kd->flavor;   // create_flavor :1598-1632: unique pairs in (band, atmosphere) order
  for (int ibnd = 0; ibnd < nbnd; ++ibnd)
    for (int iatm = 0; iatm < 2; ++iatm) {
      int pr[2];
      rewrite(ks_red[(size_t)0 + 2 * (iatm + 2 * (size_t)ibnd)], ks_red[(size_t)1 + 2 * (iatm + 2 * (size_t)ibnd)], pr);
      bool seen = false;
      for (size_t f = 0; f + 1 < flavor.size(); f += 2) seen = seen || (flavor[f] == pr[0] && flavor[f + 1] == pr[1]);
      if (!seen) { flavor.push_back(pr[0]); flavor.push_back(pr[1]); }
    }
  const int nflav = (int)flavor.size() / 2;
  kd->gpoint_flavor.assign((size_t)2 * ngpt, -1);   // create_gpoint_flavor :1930-1946
  for (int g = 0; g < ngpt; ++g)
    for (int iatm = 0; iatm < 2; ++iatm) {
      const int ibnd = kd->gpoint_bands[(size_t)g] - 1;
      int pr[2];
      rewrite(ks_red[(size_t)0 + 2 * (iatm + 2 * (size_t)ibnd)], ks_red[(size_t)1 + 2 * (iatm + 2 * (size_t)ibnd)], pr);
      for (int f = 0; f < nflav; ++f)
        if (flavor[2 * (size_t)f] == pr[0] && flavor[2 * (size_t)f + 1] == pr[1]) { kd->gpoint_flavor[(size_t)iatm + 2 * (size_t)g] = f + 1; break; }
    }
  kd->is_key.assign((size_t)ngas, 0);   // :1364-1372
  for (int v : flavor)
    if (v != 0 && v <= ngas) kd->is_key[(size_t)v - 1] = 1;

  // ---- the kernel-facing record
  rrtmgpb_kdist& t = kd->t;
  std::memset(&t, 0, sizeof t);
  t.ngas = ngas; t.nflav = nflav; t.neta = neta; t.npres = npres; t.ntemp = ntemp; t.nbnd = nbnd; t.ngpt = ngpt;
  t.nminorlower = (int)kd->lower.gases.size(); t.nminorklower = kd->lower.tot_g;
  t.nminorupper = (int)kd->upper.gases.size(); t.nminorkupper = kd->upper.tot_g;
  t.idx_h2o = loc_in("h2o", kd->gas_names);   // :580
  t.flavor = flavor.data(); t.gpoint_flavor = kd->gpoint_flavor.data(); t.band_lims_gpt = kd->band_lims_gpt.data();
  t.gpoint_bands = kd->gpoint_bands.data(); t.band_lims_wvn = kd->band_lims_wvn.data();
  t.press_ref_log = kd->press_ref_log.data(); t.temp_ref = kd->temp_ref.data(); t.vmr_ref = kd->vmr_ref.data();
  t.temp_ref_min = kd->temp_ref[0]; t.temp_ref_max = kd->temp_ref[(size_t)ntemp - 1];                 // :1353-1356
  t.press_ref_min = kd->press_ref[(size_t)npres - 1]; t.press_ref_max = kd->press_ref[0];
  t.press_ref_log_delta = (std::log(t.press_ref_min) - std::log(t.press_ref_max)) / (Float)(npres - 1);   // :1359
  t.temp_ref_delta = (t.temp_ref_max - t.temp_ref_min) / (Float)(ntemp - 1);                            // :1360
  t.press_ref_trop_log = std::log(raw->press_ref_trop);                                                   // :1326
  t.kmajor = kd->kmajor.data(); t.kminor_lower = kd->lower.kminor.data(); t.kminor_upper = kd->upper.kminor.data();
  // (contributor vectors keep one addressable element when empty, so every pointer stays valid)
  for (MinorRed* m : {&kd->lower, &kd->upper})
    if (m->gases.empty()) {
      m->limits = {1, 0}; m->kstart = {1}; m->idx_minor = {1}; m->idx_scaling = {0}; m->dens = {0}; m->comp = {0};
    }
  t.minor_limits_gpt_lower = kd->lower.limits.data(); t.minor_limits_gpt_upper = kd->upper.limits.data();
  static_assert(sizeof(Bool) == 1, "Bool is a 1-byte logical");
  t.minor_scales_with_density_lower = reinterpret_cast<const Bool*>(kd->lower.dens.data());
  t.minor_scales_with_density_upper = reinterpret_cast<const Bool*>(kd->upper.dens.data());
  t.scale_by_complement_lower = reinterpret_cast<const Bool*>(kd->lower.comp.data());
  t.scale_by_complement_upper = reinterpret_cast<const Bool*>(kd->upper.comp.data());
  t.idx_minor_lower = kd->lower.idx_minor.data(); t.idx_minor_upper = kd->upper.idx_minor.data();
  t.idx_minor_scaling_lower = kd->lower.idx_scaling.data(); t.idx_minor_scaling_upper = kd->upper.idx_scaling.data();
  t.kminor_start_lower = kd->lower.kstart.data(); t.kminor_start_upper = kd->upper.kstart.data();
  t.krayl = kd->krayl.empty() ? nullptr : kd->krayl.data();
  if (lw) {   // load_int :1018-1029
    kd->totplnk.assign(raw->totplnk, raw->totplnk + (size_t)raw->ntemp_planck * nbnd);
    kd->planck_frac = to_kernel_layout(raw->plank_fraction);
    if (raw->optimal_angle_fit)
      kd->optimal_angle_fit.assign(raw->optimal_angle_fit, raw->optimal_angle_fit + (size_t)raw->nfit_coeffs * nbnd);
    t.totplnk = kd->totplnk.data(); t.planck_frac = kd->planck_frac.data(); t.nPlanckTemp = raw->ntemp_planck;
    t.totplnk_delta = (t.temp_ref_max - t.temp_ref_min) / (Float)(raw->ntemp_planck - 1);   // :1029
  } else {    // load_ext :1118-1143
    kd->solar_quiet.assign(raw->solar_source_quiet, raw->solar_source_quiet + ngpt);
    kd->solar_facular.assign(raw->solar_source_facular, raw->solar_source_facular + ngpt);
    kd->solar_sunspot.assign(raw->solar_source_sunspot, raw->solar_source_sunspot + ngpt);
    kd->solar_source.assign((size_t)ngpt, (Float)0);
    t.solar_source = kd->solar_source.data();
    char msg[128];
    if (rrtmgpb_kdist_set_solar_variability(kd, raw->mg_default, raw->sb_default, (Float)-1, msg)) return bail(msg);   // :1143
  }
  return kd;
}

void rrtmgpb_kdist_loaded_free(rrtmgpb_kdist_loaded* kd) { delete kd; }
const rrtmgpb_kdist* rrtmgpb_kdist_loaded_tables(const rrtmgpb_kdist_loaded* kd) { return kd ? &kd->t : nullptr; }
const char* rrtmgpb_kdist_loaded_gas_name(const rrtmgpb_kdist_loaded* kd, int i) {
  return (kd && i >= 0 && i < (int)kd->gas_names.size()) ? kd->gas_names[(size_t)i].c_str() : nullptr;
}
int rrtmgpb_kdist_loaded_is_key(const rrtmgpb_kdist_loaded* kd, int i) {
  return (kd && i >= 0 && i < (int)kd->is_key.size()) ? kd->is_key[(size_t)i] : 0;
}
const Float* rrtmgpb_kdist_loaded_optimal_angle_fit(const rrtmgpb_kdist_loaded* kd) {
  return (kd && !kd->optimal_angle_fit.empty()) ? kd->optimal_angle_fit.data() : nullptr;
}

int rrtmgpb_kdist_set_tsi(rrtmgpb_kdist_loaded* kd, Float tsi, char* errmsg) {   // :800-835
  if (errmsg) errmsg[0] = 0;
  if (tsi < 0) return fail(errmsg, "tsi out of range");
  Float norm = 0;
  for (Float v : kd->solar_source) norm += v;
  norm = (Float)1 / norm;
  for (Float& v : kd->solar_source) v = v * tsi * norm;
  return 0;
}

int rrtmgpb_kdist_set_solar_variability(rrtmgpb_kdist_loaded* kd, Float mg_index, Float sb_index, Float tsi, char* errmsg) {
  if (errmsg) errmsg[0] = 0;
  const Float a_offset = (Float)0.1495954, b_offset = (Float)0.00066696;   // :776-777
  std::string msg;
  if (mg_index < 0) msg = "mg_index out of range";
  if (sb_index < 0) msg = "sb_index out of range";
  if (!msg.empty()) return fail(errmsg, msg);
  for (size_t g = 0; g < kd->solar_source.size(); ++g)   // :788-792
    kd->solar_source[g] = kd->solar_quiet[g] + (mg_index - a_offset) * kd->solar_facular[g] + (sb_index - b_offset) * kd->solar_sunspot[g];
  if (tsi >= 0) return rrtmgpb_kdist_set_tsi(kd, tsi, errmsg);   // present(tsi)
  return 0;
}

}  // extern "C"
