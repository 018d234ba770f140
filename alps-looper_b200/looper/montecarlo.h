// montecarlo.h -- mc_steps and temperature bookkeeping of the worker
// (reference: looper/montecarlo.h:36-73 mc_steps; looper/temperature.h:35-119: T / BETA and the
// piecewise-linear annealing schedule T_START_i / T_DURATION_i during thermalisation).
#pragma once
#include <string>
#include <utility>
#include <vector>
#include "parameters.h"

namespace looper {

class mc_steps {
public:
  mc_steps() {}
  explicit mc_steps(const Parameters& p)
      : sweeps_(p.value_or_default<unsigned>("SWEEPS", 65536u)),
        therm_(p.defined("THERMALIZATION") ? p.value_or_default<unsigned>("THERMALIZATION", 0u) : (sweeps_ >> 3)) {}
  mc_steps& operator++() { ++mcs_; return *this; }
  unsigned operator()() const { return mcs_; }
  bool can_work() const { return mcs_ < therm_ || mcs_ - therm_ < sweeps_; }
  bool is_thermalized() const { return mcs_ >= therm_; }
  double progress() const { return double(mcs_) / double(therm_ + sweeps_); }
  unsigned thermalization() const { return therm_; }
  unsigned sweeps() const { return sweeps_; }
  void set(unsigned mcs) { mcs_ = mcs; }
private:
  unsigned mcs_ = 0, sweeps_ = 0, therm_ = 0;
};

class temperature {
public:
  temperature() {}
  explicit temperature(const Parameters& p) {
    if (p.defined("T")) final_ = p.value_or_default<double>("T", 1.0);
    else if (p.defined("BETA")) final_ = 1.0 / p.value_or_default<double>("BETA", 1.0);
    else final_ = 1.0;
    unsigned at = 0;
    for (int n = 0;; ++n) {  // annealing stages (temperature.h:55-80)
      const std::string ns = std::to_string(n);
      if (!(p.defined("T_START_" + ns) && p.defined("T_DURATION_" + ns))) break;
      const unsigned d = p.value_or_default<unsigned>("T_DURATION_" + ns, 0u);
      seq_.push_back({at, p.value_or_default<double>("T_START_" + ns, final_)});
      at += d;
    }
    end_ = at;
    current_ = final_;
  }
  void set_beta(double beta) { final_ = current_ = 1.0 / beta; seq_.clear(); end_ = 0; }
  unsigned annealing_steps() const { return end_; }
  // temperature at Monte Carlo step mcs: linear between stage starts, final after the last stage
  double operator()(unsigned mcs) const {
    if (seq_.empty() || mcs >= end_) return final_;
    size_t k = 0;
    while (k + 1 < seq_.size() && seq_[k + 1].first <= mcs) ++k;
    const unsigned t0 = seq_[k].first, t1 = (k + 1 < seq_.size()) ? seq_[k + 1].first : end_;
    const double T0 = seq_[k].second, T1 = (k + 1 < seq_.size()) ? seq_[k + 1].second : final_;
    return T0 + (T1 - T0) * double(mcs - t0) / double(t1 - t0);
  }
  double final() const { return final_; }
private:
  double final_ = 1, current_ = 1;
  unsigned end_ = 0;
  std::vector<std::pair<unsigned, double>> seq_;
};

}  // namespace looper
