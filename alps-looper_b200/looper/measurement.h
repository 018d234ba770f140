// measurement.h -- host side of the measurement framework for the accelerated path:
// observable_set (stand-in for alps::ObservableSet with a binning error analysis) and the
// commit() arithmetic of the estimators the GPU fills in
//   energy          looper/energy.h:74-83
//   susceptibility  looper/susceptibility.h:199-254 (improved collector, path-integral and SSE forms)
// and the evaluated observables of energy.h:89-102 (specific heat) and susceptibility.h:340-376 (Binder ratios)
// The per-cluster sums themselves (susceptibility.h:117-155,182-198) are computed on the device
// (csrc/lq_kernels.cuh k_estimate / k_collect); lq_collector carries the 14 collector sums.
#pragma once
#include <cmath>
#include <cstdint>
#include <istream>
#include <map>
#include <ostream>
#include <string>
#include <utility>
#include <vector>
#include "../../include/lq.h"

namespace looper {

inline double power2(double x) { return x * x; }
inline double power4(double x) { return power2(power2(x)); }
inline double dip(double x, double y) { return y > 0 ? x / y : 0.0; }  // divide_if_positive.h:42

// scalar observable with logarithmic binning (mean, error from the largest level with >= 32 bins)
class observable {
public:
  void operator<<(double x) {
    ++n_; sum_ += x;
    // jackknife bins for the evaluated observables (at most 128 bins of equal size; when they are full,
    // neighbours are merged and the bin size doubles)
    jcur_ += x;
    if (++jfill_ == jsize_) {
      jb_.push_back(jcur_); jcur_ = 0; jfill_ = 0;
      if (jb_.size() == 128) {
        for (size_t i = 0; i < 64; ++i) jb_[i] = jb_[2 * i] + jb_[2 * i + 1];
        jb_.resize(64);
        jsize_ *= 2;
      }
    }
    size_t l = 0;
    double v = x;
    for (;;) {
      if (lv_.size() <= l) lv_.push_back(level());
      level& L = lv_[l];
      L.s += v; L.s2 += v * v; ++L.n;
      if (!L.have) { L.pend = v; L.have = true; break; }
      v = 0.5 * (L.pend + v); L.have = false; ++l;
    }
  }
  unsigned long count() const { return n_; }
  double mean() const { return n_ ? sum_ / n_ : 0.0; }
  const std::vector<double>& bin_sums() const { return jb_; }   // complete jackknife bins (sums)
  unsigned long bin_size() const { return jsize_; }
  double naive_error() const { return lv_.empty() ? 0.0 : err(lv_[0]); }
  double error() const {
    double e = 0;
    for (const level& L : lv_) if (L.n >= 32) e = std::max(e, err(L));
    return e > 0 ? e : naive_error();
  }
  // checkpoint payload: the whole binning state (the ALPS scheduler dumps its ObservableSet with the worker)
  void save(std::ostream& os) const {
    const uint64_t nl = lv_.size();
    os.write(reinterpret_cast<const char*>(&n_), sizeof n_);
    os.write(reinterpret_cast<const char*>(&sum_), sizeof sum_);
    os.write(reinterpret_cast<const char*>(&nl), sizeof nl);
    for (const level& L : lv_) {
      const double v[3] = {L.s, L.s2, L.pend};
      const uint64_t m[2] = {L.n, L.have ? 1u : 0u};
      os.write(reinterpret_cast<const char*>(v), sizeof v);
      os.write(reinterpret_cast<const char*>(m), sizeof m);
    }
    const uint64_t jh[3] = {jb_.size(), jsize_, jfill_};
    os.write(reinterpret_cast<const char*>(jh), sizeof jh);
    os.write(reinterpret_cast<const char*>(&jcur_), sizeof jcur_);
    os.write(reinterpret_cast<const char*>(jb_.data()), std::streamsize(jb_.size() * sizeof(double)));
  }
  void load(std::istream& is) {
    uint64_t nl = 0;
    is.read(reinterpret_cast<char*>(&n_), sizeof n_);
    is.read(reinterpret_cast<char*>(&sum_), sizeof sum_);
    is.read(reinterpret_cast<char*>(&nl), sizeof nl);
    if (!is || nl > 64) { is.setstate(std::ios::failbit); return; }
    lv_.assign(nl, level());
    for (level& L : lv_) {
      double v[3];
      uint64_t m[2];
      is.read(reinterpret_cast<char*>(v), sizeof v);
      is.read(reinterpret_cast<char*>(m), sizeof m);
      L.s = v[0]; L.s2 = v[1]; L.pend = v[2]; L.n = m[0]; L.have = m[1] != 0;
    }
    uint64_t jh[3] = {0, 1, 0};
    is.read(reinterpret_cast<char*>(jh), sizeof jh);
    is.read(reinterpret_cast<char*>(&jcur_), sizeof jcur_);
    if (!is || jh[0] > 128 || jh[1] == 0) { is.setstate(std::ios::failbit); return; }
    jb_.assign(size_t(jh[0]), 0.0);
    jsize_ = (unsigned long)jh[1]; jfill_ = (unsigned long)jh[2];
    is.read(reinterpret_cast<char*>(jb_.data()), std::streamsize(jb_.size() * sizeof(double)));
  }
  double tau() const {  // integrated autocorrelation time estimate
    const double e0 = naive_error(), e = error();
    return e0 > 0 ? 0.5 * (power2(e / e0) - 1) : 0.0;
  }
private:
  struct level { double s = 0, s2 = 0, pend = 0; unsigned long n = 0; bool have = false; };
  static double err(const level& L) {
    if (L.n < 2) return 0.0;
    const double m = L.s / L.n, var = L.s2 / L.n - m * m;
    return var > 0 ? std::sqrt(var / (L.n - 1)) : 0.0;
  }
  unsigned long n_ = 0;
  double sum_ = 0;
  std::vector<level> lv_;
  std::vector<double> jb_;
  unsigned long jsize_ = 1, jfill_ = 0;
  double jcur_ = 0;
};

class observable_set {
public:
  observable& operator[](const std::string& name) {
    auto it = obs_.find(name);
    if (it == obs_.end()) { order_.push_back(name); return obs_[name]; }
    return it->second;
  }
  bool has(const std::string& name) const { return obs_.count(name) != 0; }
  const observable& at(const std::string& name) const { return obs_.at(name); }
  const std::vector<std::string>& names() const { return order_; }
  void save(std::ostream& os) const {
    const uint64_t n = order_.size();
    os.write(reinterpret_cast<const char*>(&n), sizeof n);
    for (const std::string& name : order_) {
      const uint64_t len = name.size();
      os.write(reinterpret_cast<const char*>(&len), sizeof len);
      os.write(name.data(), std::streamsize(len));
      obs_.at(name).save(os);
    }
  }
  void load(std::istream& is) {
    uint64_t n = 0;
    is.read(reinterpret_cast<char*>(&n), sizeof n);
    for (uint64_t k = 0; is && k < n && n < 4096; ++k) {
      uint64_t len = 0;
      is.read(reinterpret_cast<char*>(&len), sizeof len);
      if (!is || len > 256) { is.setstate(std::ios::failbit); return; }
      std::string name(size_t(len), ' ');
      is.read(&name[0], std::streamsize(len));
      (*this)[name].load(is);
    }
  }
  void print(std::ostream& os) const {
    for (const std::string& n : order_) {
      const observable& o = obs_.at(n);
      os << n << ": " << o.mean() << " +/- " << o.error() << "; tau = " << o.tau() << "\n";
    }
    for (const std::string& n : eorder_) os << n << ": " << eval_.at(n).first << " +/- " << eval_.at(n).second << "\n";
  }
  // Evaluated observables (stand-in for alps::RealObsevaluator arithmetic, energy.h:89-102,
  // susceptibility.h:340-376): `f` of the means of the named observables, error by the jackknife over their
  // common bins (the operands are committed together, once per sweep, so their bins line up).  Returns false
  // -- like the reference's try/catch around every evaluation -- if an operand is missing or has < 2 bins.
  template <class F>
  bool evaluate(const std::string& name, const std::vector<std::string>& operands, F f) {
    std::vector<const observable*> o;
    for (const std::string& n : operands) {
      if (!has(n)) return false;
      o.push_back(&obs_.at(n));
    }
    const size_t nb = o[0]->bin_sums().size();
    if (nb < 2) return false;
    for (const observable* x : o)
      if (x->bin_sums().size() != nb || x->bin_size() != o[0]->bin_size()) return false;
    const double per_bin = double(o[0]->bin_size());
    std::vector<double> tot(o.size(), 0.0), arg(o.size());
    for (size_t k = 0; k < o.size(); ++k)
      for (double b : o[k]->bin_sums()) tot[k] += b;
    for (size_t k = 0; k < o.size(); ++k) arg[k] = tot[k] / (per_bin * nb);
    const double value = f(arg);
    double s = 0, s2 = 0;
    for (size_t i = 0; i < nb; ++i) {
      for (size_t k = 0; k < o.size(); ++k) arg[k] = (tot[k] - o[k]->bin_sums()[i]) / (per_bin * (nb - 1));
      const double fi = f(arg);
      s += fi; s2 += fi * fi;
    }
    const double var = s2 / nb - (s / nb) * (s / nb);
    if (!eval_.count(name)) eorder_.push_back(name);
    eval_[name] = std::make_pair(value, var > 0 ? std::sqrt((nb - 1) * var) : 0.0);
    return true;
  }
  bool has_evaluated(const std::string& name) const { return eval_.count(name) != 0; }
  std::pair<double, double> evaluated(const std::string& name) const { return eval_.at(name); }   // value, error
private:
  std::map<std::string, observable> obs_;
  std::vector<std::string> order_;
  std::map<std::string, std::pair<double, double> > eval_;
  std::vector<std::string> eorder_;
};

struct energy {
  static void init_observables(observable_set& m) { m["Energy"]; m["Energy Density"]; m["Energy^2"]; }
  static void commit(observable_set& m, const lq_collector& c, double beta, double vol, double sign = 1) {
    m["Energy"] << sign * c.ene;
    m["Energy Density"] << sign * c.ene / vol;
    m["Energy^2"] << sign * (power2(c.ene) - c.nop / power2(beta));
  }
  // energy.h:89-102: "Specific Heat" = beta^2 (<E^2> - <E>^2) / volume
  static void evaluate(observable_set& m) {
    if (!m.has("Inverse Temperature") || !m.has("Volume")) return;
    const double beta = m["Inverse Temperature"].mean(), vol = m["Volume"].mean();
    m.evaluate("Specific Heat", {"Energy", "Energy^2"},
               [=](const std::vector<double>& x) { return beta * beta * (x[1] - x[0] * x[0]) / vol; });
  }
};

struct susceptibility {
  static void commit(observable_set& m, const lq_collector& c, double beta, double vol, bool bipartite,
                     bool sse = false, double sign = 1) {
    const double nop = c.nop;
    m["Magnetization"] << 0.0;
    m["Magnetization Density"] << 0.0;
    m["|Magnetization|"] << sign * std::abs(c.umag0);
    m["|Magnetization Density|"] << sign * std::abs(c.umag0) / vol;
    m["Magnetization^2"] << sign * c.umag2;
    m["Magnetization Density^2"] << sign * c.umag2 / power2(vol);
    m["Magnetization^4"] << sign * (3 * power2(c.umag2) - 2 * c.umag4);
    m["Magnetization Density^4"] << sign * (3 * power2(c.umag2) - 2 * c.umag4) / power4(vol);
    m["Susceptibility"] << (sse ? sign * beta * (dip(c.umag, nop) + c.umag2) / (nop + 1) / vol
                                : sign * beta * c.umag / vol);
    m["Generalized Magnetization^2"] << sign * c.usize2;
    m["Generalized Magnetization Density^2"] << sign * c.usize2 / power2(vol);
    m["Generalized Magnetization^4"] << sign * (3 * power2(c.usize2) - 2 * c.usize4);
    m["Generalized Magnetization Density^4"] << sign * (3 * power2(c.usize2) - 2 * c.usize4) / power4(vol);
    m["Generalized Susceptibility"] << (sse ? sign * beta * (dip(c.usize, nop) + c.usize2) / (nop + 1) / vol
                                            : sign * beta * c.usize / vol);
    if (!bipartite) return;
    m["Staggered Magnetization"] << 0.0;
    m["Staggered Magnetization Density"] << 0.0;
    m["|Staggered Magnetization|"] << sign * std::abs(c.smag0);
    m["|Staggered Magnetization Density|"] << sign * std::abs(c.smag0) / vol;
    m["Staggered Magnetization^2"] << sign * c.smag2;
    m["Staggered Magnetization Density^2"] << sign * c.smag2 / power2(vol);
    m["Staggered Magnetization^4"] << sign * (3 * power2(c.smag2) - 2 * c.smag4);
    m["Staggered Magnetization Density^4"] << sign * (3 * power2(c.smag2) - 2 * c.smag4) / power4(vol);
    m["Staggered Susceptibility"] << (sse ? sign * beta * (dip(c.smag, nop) + c.smag2) / (nop + 1) / vol
                                          : sign * beta * c.smag / vol);
    m["Generalized Staggered Magnetization^2"] << sign * c.ssize2;
    m["Generalized Staggered Magnetization Density^2"] << sign * c.ssize2 / power2(vol);
    m["Generalized Staggered Magnetization^4"] << sign * (3 * power2(c.ssize2) - 2 * c.ssize4);
    m["Generalized Staggered Magnetization Density^4"] << sign * (3 * power2(c.ssize2) - 2 * c.ssize4) / power4(vol);
    m["Generalized Staggered Susceptibility"] << (sse ? sign * beta * (dip(c.ssize, nop) + c.ssize2) / (nop + 1) / vol
                                                      : sign * beta * c.ssize / vol);
  }
};

// susceptibility.h:340-376: the four Binder ratios <m^2>^2 / <m^4>
inline void evaluate_susceptibility(observable_set& m) {
  for (const char* what : {"Magnetization", "Staggered Magnetization", "Generalized Magnetization", "Generalized Staggered Magnetization"}) {
    const std::string w(what);
    m.evaluate("Binder Ratio of " + w, {w + "^2", w + "^4"},
               [](const std::vector<double>& x) { return x[1] != 0 ? x[0] * x[0] / x[1] : 0.0; });
  }
}

// stiffness.h:118-133
struct stiffness {
  static void commit(observable_set& m, const lq_collector& c, double beta, int dim, double sign = 1) {
    if (dim > 0) m["Stiffness"] << sign * c.w2 / (beta * dim);
  }
};

// transmag.h:95-110: length of the clusters cut by a site operator
struct transverse_magnetization {
  static void commit(observable_set& m, const lq_collector& c, double vol, double sign = 1) {
    m["Transverse Magnetization"] << 0.5 * sign * c.tlen;
    m["Transverse Magnetization Density"] << 0.5 * sign * c.tlen / vol;
  }
};

}  // namespace looper
