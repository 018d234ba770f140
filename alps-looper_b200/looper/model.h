// model.h -- host-side spinmodel_helper / weight_helper for S=1/2 XXZ bonds and transverse-field
// sites (reference: looper/model.h:38-127, looper/weight_impl.h:62-88,90-188,349-423).  The loop equations
//   -offset + v1 + v3 = -Jz/4,  -offset + v0 + v2 = +Jz/4,  v0 + v1 = |Jxy|/2
// are solved as in weight_impl.h (standard solution, or the "ergodic" one for FORCE_SCATTER = a).
#pragma once
#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <string>
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
  void init(const Parameters& p, const lattice_helper& lat) {
    check_supported(p);
    const double J = p.value_or_default<double>("J", 1.0);
    const double jxy = p.value_or_default<double>("Jxy", J), jz = p.value_or_default<double>("Jz", J);
    // transverse field Gamma (ALPS "spin" model: H -= Gamma Sx): site graphs
    site_weight_helper sw(p.value_or_default<double>("Gamma", 0.0));
    if (sw.sign < 0) throw std::invalid_argument("negative sign (Gamma < 0) is not supported");
    const double a = p.value_or_default<double>("FORCE_SCATTER", 0.0);
    const int nb = num_bonds(lat.vg());
    xxz_bond_weight_helper w(bond_parameter_xxz(0, jxy, jz), a);
    if (nb > 0 && w.sign < 0 && !lat.is_bipartite()) throw std::invalid_argument("negative sign (frustration) is not supported");
    if (nb > 0 && w.sign < 0 && sw.has_weight())
      throw std::invalid_argument("negative sign (antiferromagnetic Jxy with a transverse field) is not supported");
    weights_.assign(4 * size_t(nb), 0.0);
    gw_ = 0;
    offset_ = 0;
    for (int b = 0; b < nb; ++b) {
      for (int g = 0; g < 4; ++g) weights_[4 * size_t(b) + g] = w.v[g];
      gw_ += w.weight();
      offset_ += w.offset;
    }
    site_weight_ = sw.weight();
    if (sw.has_weight()) {
      const int ns = num_sites(lat.vg());
      gw_ += ns * sw.weight();
      offset_ += ns * sw.offset;
    }
  }
  // A drop-in must refuse what it does not implement instead of quietly simulating something else: the
  // parameters of the ALPS "spin" model (model_parameter.h) that change the Hamiltonian but have no
  // counterpart on the accelerated path are errors, not ignored keys.
  static void check_supported(const Parameters& p) {
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
      // type-dependent couplings (Jz0, Jxy1, Gamma0, h1, J', ...): the C ABI takes per-bond and per-site weights
      // (lq_model.bond_weights / site_weights), this host mirror only fills them uniformly
      for (const char* stem : {"Jxy", "Jz", "J", "Gamma", "h", "D"}) {
        const size_t n = std::char_traits<char>::length(stem);
        if (k.size() > n && k.compare(0, n, stem) == 0 &&
            k.find_first_not_of("0123456789'", n) == std::string::npos)
          throw std::invalid_argument("parameter " + k + ": site- or bond-type dependent couplings are not supported by this host mirror");
      }
    }
  }
  double graph_weight() const { return gw_; }       // model.h:114, graph_impl.h:694
  double energy_offset() const { return offset_; }  // model.h:84, weight_impl.h:396-404
  bool is_signed() const { return false; }
  bool has_field() const { return false; }
  const std::vector<double>& bond_weights() const { return weights_; }
  double site_weight() const { return site_weight_; }   // uniform |Hx|/2 (0: no site graphs)

private:
  std::vector<double> weights_;
  double gw_ = 0, offset_ = 0, site_weight_ = 0;
};

}  // namespace looper
