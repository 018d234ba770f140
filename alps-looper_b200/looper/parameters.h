// parameters.h -- stand-in for alps::Parameters as consumed by the loop worker
// (keys: doc/index.md:96-147; loop.ip).  "KEY = value;" or "KEY = value" per line, '#' and '//'
// comments, values may be quoted.  Host-side plumbing only.
#pragma once
#include <cstdlib>
#include <istream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace looper {

class Parameters {
public:
  Parameters() {}
  bool defined(const std::string& k) const { return kv_.count(k) != 0; }
  std::string& operator[](const std::string& k) { return kv_[k]; }
  const std::string& get(const std::string& k) const {
    auto it = kv_.find(k);
    if (it == kv_.end()) throw std::invalid_argument("parameter " + k + " not defined");
    return it->second;
  }
  template <class T>
  T value_or_default(const std::string& k, T def) const {
    auto it = kv_.find(k);
    if (it == kv_.end()) return def;
    std::istringstream is(it->second);
    T v;
    if (!(is >> v)) throw std::invalid_argument("parameter " + k + " = '" + it->second + "' is not readable");
    return v;
  }
  std::string value_or_default(const std::string& k, const char* def) const {
    auto it = kv_.find(k);
    return it == kv_.end() ? std::string(def) : it->second;
  }
  template <class T>
  void set(const std::string& k, T v) {
    std::ostringstream os;
    os.precision(17);
    os << v;
    kv_[k] = os.str();
  }
  // reads one parameter block (up to a '{' / end of stream)
  void parse(std::istream& in) {
    std::string line;
    while (std::getline(in, line)) {
      auto c = line.find('#');
      if (c != std::string::npos) line.erase(c);
      c = line.find("//");
      if (c != std::string::npos) line.erase(c);
      // statements end at ';' -- outside quotes: ALGORITHM = "loop; sse" is one value (loop.op:334)
      std::vector<std::string> stmts(1);
      bool quoted = false;
      for (char ch : line) {
        if (ch == '"') quoted = !quoted;
        if (ch == ';' && !quoted) stmts.emplace_back();
        else stmts.back().push_back(ch);
      }
      for (const std::string& stmt : stmts) {
        auto eq = stmt.find('=');
        if (eq == std::string::npos) continue;
        std::string k = trim(stmt.substr(0, eq)), v = trim(stmt.substr(eq + 1));
        if (v.size() >= 2 && v.front() == '"' && v.back() == '"') v = v.substr(1, v.size() - 2);
        if (!k.empty()) kv_[k] = v;
      }
    }
  }
  const std::map<std::string, std::string>& items() const { return kv_; }

private:
  static std::string trim(const std::string& s) {
    const char* ws = " \t\r\n{}";
    auto b = s.find_first_not_of(ws);
    if (b == std::string::npos) return "";
    auto e = s.find_last_not_of(ws);
    return s.substr(b, e - b + 1);
  }
  std::map<std::string, std::string> kv_;
};

}  // namespace looper
