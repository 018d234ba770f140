// model.h -- host-side spinmodel_helper / weight_helper for S=1/2 XXZ bonds and transverse-field
// sites (reference: looper/model.h:38-127, looper/weight_impl.h:62-88,90-188,349-423).  The loop equations
//   -offset + v1 + v3 = -Jz/4,  -offset + v0 + v2 = +Jz/4,  v0 + v1 = |Jxy|/2
// are solved as in weight_impl.h (standard solution, or the "ergodic" one for FORCE_SCATTER = a).
#pragma once
#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "lattice.h"
#include "parameters.h"

namespace looper {

struct bond_parameter_xxz {
  double c, jxy, jz;
  bond_parameter_xxz(double c_ = 0, double jxy_ = 0, double jz_ = 0) : c(c_), jxy(jxy_), jz(jz_) {}
};

struct xxz_bond_weight_helper {
  static const int num_graphs = 4;
  int sign = 1;
  double offset = 0;
  double v[4] = {0, 0, 0, 0};
  xxz_bond_weight_helper() {}
  xxz_bond_weight_helper(const bond_parameter_xxz& p, double force_scatter = 0) { init(p, force_scatter); }
  void init(const bond_parameter_xxz& p, double a) {
    auto pos = [](double x) { return x > 0 ? x : 0.0; };
    a = std::min(1.0, std::max(0.0, a));
    sign = (p.jxy <= 0) ? 1 : -1;
    const double jxy = std::abs(p.jxy), jz = p.jz;
    v[0] = v[1] = v[2] = v[3] = 0;
    if (jxy + std::abs(jz) > 1e-10) {
      if (jxy - jz > 2 * a * jxy) {
        v[0] = pos(std::min(jxy / 2, (jxy + jz) / 4));
        v[1] = pos(std::min(jxy / 2, (jxy - jz) / 4));
        v[2] = pos(-(jxy - jz) / 2);
        v[3] = pos(-(jxy + jz) / 2);
      } else {
        v[0] = (1 - a) * jxy / 2;
        v[1] = a * jxy / 2;
        v[2] = -((1 - 2 * a) * jxy - jz) / 2;
      }
    }
    offset = weight() / 2;
  }
  double weight() const { return v[0] + v[1] + v[2] + v[3]; }
  bool has_weight() const { return weight() > 1e-10; }
};

// weight_impl.h:62-88: site graph weight v0 = |Hx|/2, offset = v0
struct site_weight_helper {
  int sign = 1;
  double offset = 0;
  double v[1] = {0};
  site_weight_helper() {}
  explicit site_weight_helper(double hx) { sign = hx >= 0 ? 1 : -1; v[0] = std::abs(hx) / 2; offset = v[0]; }
  double weight() const { return v[0]; }
  bool has_weight() const { return weight() > 1e-10; }
};

class spinmodel_helper {
public:
  spinmodel_helper() {}
  spinmodel_helper(const Parameters& p, const lattice_helper& lat) { init(p, lat); }
  // Couplings of the ALPS "spin" model: J (default 1), Jxy, Jz per bond, Gamma per site; a key with the type
  // of the bond / site appended (Jz0, Jxy1, Gamma0: model_parameter.h) overrides the plain one on that type.
  void init(const Parameters& p, const lattice_helper& lat) {
    const virtual_graph& g = lat.vg();
    const int nb = num_bonds(g), ns = num_sites(g);
    int nbt = 1, nst = 1;
    for (int t : g.bond_type) nbt = std::max(nbt, t + 1);
    for (int t : g.site_type) nst = std::max(nst, t + 1);
    check_supported(p, nbt, nst);
    const double J = p.value_or_default<double>("J", 1.0);
    const double a = p.value_or_default<double>("FORCE_SCATTER", 0.0);
    auto typed = [&](const char* key, int t, double def) {
      return p.value_or_default<double>(key + std::to_string(t), p.value_or_default<double>(key, def));
    };
    std::vector<xxz_bond_weight_helper> bw(nbt);
    for (int t = 0; t < nbt; ++t) bw[t].init(bond_parameter_xxz(0, typed("Jxy", t, J), typed("Jz", t, J)), a);
    std::vector<site_weight_helper> sw(nst);
    for (int t = 0; t < nst; ++t) sw[t] = site_weight_helper(typed("Gamma", t, 0.0));
    // Signs.  The offdiagonal weights must be made positive by ONE gauge e_s = +-1 (a rotation by pi about z
    // on the sites with e_s = -1): e_s e_t = sign of every bond with an offdiagonal term, e_s = sign of the
    // field on every site that has one (up to a global flip per connected component).  The antiferromagnet on
    // a bipartite lattice passes (sublattice rotation), with a STAGGERED transverse field too
    // (extras/transmag, check/transmag-3); with a uniform field it does not: a genuine sign problem,
    // which the reference carries as a sign (dispatch<..., SIGN, ...>) and this path refuses.
    {
      std::vector<std::vector<std::pair<int, int> > > adj(ns);
      for (int b = 0; b < nb; ++b) {
        const xxz_bond_weight_helper& w = bw[g.bond_type[b]];
        if (w.v[0] + w.v[1] <= 1e-10) continue;   // no offdiagonal term: the bond carries no sign
        adj[source(b, g)].push_back(std::make_pair(target(b, g), w.sign));
        adj[target(b, g)].push_back(std::make_pair(source(b, g), w.sign));
      }
      std::vector<int> e(ns, 0), stack;
      for (int s0 = 0; s0 < ns; ++s0) {
        if (e[s0]) continue;
        e[s0] = 1;
        stack.assign(1, s0);
        int field = 0;   // e_s * sign(Gamma_s) of the component, once a site with a field was seen
        while (!stack.empty()) {
          const int s = stack.back();
          stack.pop_back();
          const site_weight_helper& f = sw[g.site_type[s]];
          if (f.has_weight()) {
            if (!field) field = e[s] * f.sign;
            else if (field != e[s] * f.sign)
              throw std::invalid_argument("negative sign (transverse field against the sign structure of the exchange, e.g. a uniform field on an antiferromagnet) is not supported");
          }
          for (const std::pair<int, int>& n : adj[s]) {
            if (!e[n.first]) { e[n.first] = e[s] * n.second; stack.push_back(n.first); }
            else if (e[n.first] != e[s] * n.second) throw std::invalid_argument("negative sign (frustration) is not supported");
          }
        }
      }
    }
    weights_.assign(4 * size_t(nb), 0.0);
    gw_ = 0;
    offset_ = 0;
    for (int b = 0; b < nb; ++b) {
      const xxz_bond_weight_helper& w = bw[g.bond_type[b]];
      for (int k = 0; k < 4; ++k) weights_[4 * size_t(b) + k] = w.v[k];
      gw_ += w.weight();
      offset_ += w.offset;
    }
    site_weights_.assign(size_t(ns), 0.0);
    site_weight_ = 0;
    uniform_sites_ = true;
    for (int s = 0; s < ns; ++s) {
      const site_weight_helper& f = sw[g.site_type[s]];
      if (!f.has_weight()) { if (s > 0 && site_weights_[0] != 0) uniform_sites_ = false; continue; }
      site_weights_[s] = f.weight();
      if (s > 0 && site_weights_[s] != site_weights_[0]) uniform_sites_ = false;
      gw_ += f.weight();
      offset_ += f.offset;
    }
    if (uniform_sites_ && ns > 0) site_weight_ = site_weights_[0];
  }
  // A drop-in must refuse what it does not implement instead of quietly simulating something else: the
  // parameters of the ALPS "spin" model (model_parameter.h) that change the Hamiltonian but have no
  // counterpart on the accelerated path are errors, not ignored keys.
  static void check_supported(const Parameters& p, int bond_types = 1, int site_types = 1) {
    const std::string model = p.value_or_default("MODEL", "spin");
    if (model != "spin") throw std::invalid_argument("MODEL '" + model + "' is not supported (only the ALPS \"spin\" model: XXZ bonds + transverse field)");
    for (const char* k : {"local_S", "S"})
      if (p.value_or_default<double>(k, 0.5) != 0.5) throw std::invalid_argument(std::string(k) + " != 1/2 is outside the accelerated path");
    if (p.value_or_default<double>("h", 0.0) != 0.0)
      throw std::invalid_argument("longitudinal fields are outside the accelerated path");
    if (p.value_or_default<double>("D", 0.0) != 0.0) throw std::invalid_argument("single-ion anisotropy D needs S > 1/2");
    for (const auto& kv : p.items()) {
      const std::string& k = kv.first;
      if (k == "Jx" || k == "Jy") throw std::invalid_argument("parameter " + k + ": XYZ couplings are not supported");
      // type-dependent couplings: Jxy<t>, Jz<t> on a bond type and Gamma<t> on a site type the lattice has are
      // read by init(); every other suffixed coupling (a type the lattice does not have, J<t>, h<t>, D<t>, J', ...)
      // would change the Hamiltonian unnoticed
      for (const char* stem : {"Jxy", "Jz", "J", "Gamma", "h", "D"}) {
        const size_t n = std::char_traits<char>::length(stem);
        if (!(k.size() > n && k.compare(0, n, stem) == 0 && k.find_first_not_of("0123456789'", n) == std::string::npos)) continue;
        const bool digits = k.find_first_not_of("0123456789", n) == std::string::npos;
        const int t = digits ? std::atoi(k.c_str() + n) : -1;
        const std::string st(stem);
        if (digits && (st == "Jxy" || st == "Jz") && t < bond_types) continue;
        if (digits && st == "Gamma" && t < site_types) continue;
        if (digits && (st == "h" || st == "D") && p.value_or_default<double>(k, 0.0) == 0.0) continue;
        throw std::invalid_argument("parameter " + k + ": this site- or bond-type dependent coupling is not supported");
      }
    }
  }
  double graph_weight() const { return gw_; }       // model.h:114, graph_impl.h:694
  double energy_offset() const { return offset_; }  // model.h:84, weight_impl.h:396-404
  bool is_signed() const { return false; }
  bool has_field() const { return false; }
  const std::vector<double>& bond_weights() const { return weights_; }
  double site_weight() const { return site_weight_; }   // |Hx|/2 if it is the same on every site (0: none, or not uniform)
  bool has_site_weights() const { for (double w : site_weights_) if (w > 0) return true; return false; }
  bool uniform_site_weights() const { return uniform_sites_; }
  const std::vector<double>& site_weights() const { return site_weights_; }   // |Hx_s|/2 per site

private:
  std::vector<double> weights_, site_weights_;
  bool uniform_sites_ = true;
  double gw_ = 0, offset_ = 0, site_weight_ = 0;
};

}  // namespace looper
