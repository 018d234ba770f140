// lattice.h -- host-side lattice_helper for the lattices the hot path is benchmarked on
// (reference: looper/lattice.h:290-371 lattice_helper, :49-62 source/target, :85 gauge_t,
// :576-675 virtual graph -- identical to the real graph for S=1/2).  ALPS XML lattice libraries are
// out of scope; LATTICE selects a built-in generator instead.
#pragma once
#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>
#include "parameters.h"

namespace looper {

struct virtual_graph {
  int nsites = 0;
  std::vector<int> src, dst;
  std::vector<double> gauge;
  int dims[3] = {0, 0, 0};
  std::vector<double> bond_vector_relative;  // 3 per bond (lattice.h bond_vector_relative_t, stiffness.h:63)
  int dimension = 0;
  std::vector<int> site_type, bond_type;     // ALPS vertex / edge types (all 0 on the plain lattices): the model's
                                             // couplings may depend on them (Jz0, Jxy1, Gamma0, ...: model.h)
};

inline int num_sites(const virtual_graph& g) { return g.nsites; }
inline int num_bonds(const virtual_graph& g) { return int(g.src.size()); }
inline int source(int b, const virtual_graph& g) { return g.src[b]; }
inline int target(int b, const virtual_graph& g) { return g.dst[b]; }
inline double gauge(int s, const virtual_graph& g) { return g.gauge[s]; }

class lattice_helper {
public:
  lattice_helper() {}
  explicit lattice_helper(const Parameters& p) { init(p); }
  void init(const Parameters& p) {
    const std::string name = p.value_or_default("LATTICE", "chain lattice");
    const int L = p.value_or_default<int>("L", 8);
    const int W = p.value_or_default<int>("W", L);
    const int H = p.value_or_default<int>("H", W);
    std::vector<int> ext;
    bool open_rung = false;
    if (name == "chain lattice") ext = {L};
    else if (name == "square lattice") ext = {L, W};
    else if (name == "simple cubic lattice") ext = {L, W, H};
    else if (name == "ladder") { ext = {L, 2}; open_rung = true; }
    else if (name == "alternating chain lattice") {
      // extras/transmag/alternating_chain.xml.in: a chain of L sites with a two-site unit cell -- vertex types
      // 0, 1, 0, 1, ... and edge types 0 (inside a cell), 1 (between cells)
      if (L % 2 || L < 4) throw std::invalid_argument("alternating chain lattice: L must be even and at least 4");
      build_hypercubic({L});
      for (int s = 0; s < L; ++s) vg_.site_type[s] = s % 2;
      for (int b = 0; b < num_bonds(vg_); ++b) vg_.bond_type[b] = vg_.src[b] % 2;
      return;
    }
    else if (name == "site") {   // a single site without bonds (check/site-*, extras/transmag: a spin in a field)
      vg_ = virtual_graph();
      vg_.nsites = 1;
      vg_.gauge.assign(1, 1.0);
      vg_.site_type.assign(1, 0);
      bipartite_ = true;
      return;
    }
    else throw std::invalid_argument("unknown LATTICE '" + name + "' (built-in: chain lattice, square lattice, simple cubic lattice, ladder, alternating chain lattice, site; ALPS lattice libraries are not read)");
    // A periodic direction of extent 2 is a DOUBLE bond in ALPS (test/lattice.op: "anisotropic square lattice",
    // L = 2, W = 4 has 16 bonds); the generator below makes a single one, which is what the ladder's open rungs
    // are.  Refused rather than simulated with half the coupling.
    for (int e : ext)
      if (e == 2 && !open_rung) throw std::invalid_argument("a periodic lattice direction of extent 2 (a double bond in ALPS) is not supported");
    if (open_rung && L == 2) throw std::invalid_argument("a periodic lattice direction of extent 2 (a double bond in ALPS) is not supported");
    build_hypercubic(ext);
  }
  // periodic hypercubic lattice, site = x + L0 (y + L1 z); direction-major bond order
  void build_hypercubic(const std::vector<int>& ext) {
    vg_ = virtual_graph();
    int n = 1;
    for (int e : ext) { if (e < 2) throw std::invalid_argument("lattice extent < 2"); n *= e; }
    vg_.nsites = n;
    vg_.dimension = int(std::min<size_t>(ext.size(), 3));
    bipartite_ = true;
    for (size_t k = 0; k < ext.size() && k < 3; ++k) { vg_.dims[k] = ext[k]; if (ext[k] % 2) bipartite_ = false; }
    vg_.gauge.assign(n, 0.0);
    int stride = 1;
    for (size_t k = 0; k < ext.size(); ++k) {
      for (int s = 0; s < n; ++s) {
        const int c = (s / stride) % ext[k];
        if (ext[k] == 2 && c == 1) continue;  // a ring of two sites has one bond
        vg_.src.push_back(s);
        vg_.dst.push_back(s + stride * (((c + 1) % ext[k]) - c));
        for (size_t x = 0; x < 3; ++x) vg_.bond_vector_relative.push_back(x == k ? 1.0 / ext[k] : 0.0);   // over the extent
      }
      stride *= ext[k];
    }
    vg_.site_type.assign(n, 0);
    vg_.bond_type.assign(vg_.src.size(), 0);
    if (bipartite_)
      for (int s = 0; s < n; ++s) {
        int par = 0, r = s;
        for (int e : ext) { par += r % e; r /= e; }
        vg_.gauge[s] = (par % 2 == 0) ? 1.0 : -1.0;
      }
  }
  const virtual_graph& vg() const { return vg_; }
  const virtual_graph& rg() const { return vg_; }  // S = 1/2: virtual graph == real graph
  double volume() const { return vg_.nsites; }
  bool is_bipartite() const { return bipartite_; }

private:
  virtual_graph vg_;
  bool bipartite_ = false;
};

inline bool is_bipartite(const lattice_helper& l) { return l.is_bipartite(); }
inline int max_virtual_sites(const lattice_helper&) { return 1; }

}  // namespace looper
