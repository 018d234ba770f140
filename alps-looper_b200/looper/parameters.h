// parameters.h -- stand-in for alps::Parameters as consumed by the loop worker
// (keys: doc/index.md:96-147; loop.ip, check/*, extras/*/*.ip).  The ALPS parameter-file conventions the
// reference's own input files rely on are kept: statements "KEY = value" separated by newline, ';' or ',';
// '#' and '//' comments; quoted values; numeric values may be EXPRESSIONS over numbers, other parameters
// and pi ("T = 1/L", "local_S = 1/2"); a '{ ... }' block is one TASK that inherits what was defined outside
// the blocks before it ("{ T = 0.1 } { T = 0.2 }": two tasks) -- parse_tasks().  Host-side plumbing only.
#pragma once
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <istream>
#include <iterator>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
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
    return convert(it->second, k, static_cast<T*>(nullptr));
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
  // reads everything into this one set, blocks included, later statements winning
  void parse(std::istream& in) {
    scan(in, [&](int ev, const std::string& k, const std::string& v) { if (ev == STATEMENT) kv_[k] = v; });
  }
  // ALPS parameter file -> one Parameters per task: every '{ ... }' block on top of `base` and of the statements
  // outside blocks that precede it; a file without blocks is one task.  task_keys() lists what the block itself set.
  static std::vector<Parameters> parse_tasks(std::istream& in, const Parameters& base = Parameters()) {
    std::vector<Parameters> tasks;
    Parameters global = base, cur;
    bool in_block = false;
    scan(in, [&](int ev, const std::string& k, const std::string& v) {
      if (ev == BLOCK_BEGIN) {
        if (in_block) throw std::invalid_argument("nested '{' in the parameter file");
        cur = global; cur.task_keys_.clear(); in_block = true;
      } else if (ev == BLOCK_END) {
        if (!in_block) throw std::invalid_argument("'}' without '{' in the parameter file");
        tasks.push_back(cur); in_block = false;
      } else if (in_block) { cur.kv_[k] = v; cur.task_keys_.push_back(k); }
      else global.kv_[k] = v;
    });
    if (in_block) throw std::invalid_argument("missing '}' at the end of the parameter file");
    if (tasks.empty()) tasks.push_back(global);
    return tasks;
  }
  const std::vector<std::string>& task_keys() const { return task_keys_; }
  const std::map<std::string, std::string>& items() const { return kv_; }

private:
  enum { STATEMENT, BLOCK_BEGIN, BLOCK_END };
  template <class F>
  static void scan(std::istream& in, F emit) {
    const std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    std::string stmt;
    bool quoted = false;
    auto flush = [&]() {
      auto eq = stmt.find('=');
      if (eq != std::string::npos) {
        std::string k = trim(stmt.substr(0, eq)), v = trim(stmt.substr(eq + 1));
        if (v.size() >= 2 && v.front() == '"' && v.back() == '"') v = v.substr(1, v.size() - 2);
        if (!k.empty()) emit(STATEMENT, k, v);
      }
      stmt.clear();
    };
    for (size_t i = 0; i < text.size(); ++i) {
      const char ch = text[i];
      if (quoted) {                      // ALGORITHM = "loop; sse" is one value (loop.op:334)
        if (ch == '"') quoted = false;
        if (ch == '\n') { quoted = false; flush(); } else stmt.push_back(ch);
        continue;
      }
      if (ch == '#' || (ch == '/' && i + 1 < text.size() && text[i + 1] == '/')) {   // comment: to the end of the line
        while (i < text.size() && text[i] != '\n') ++i;
        flush();
      } else if (ch == '"') { quoted = true; stmt.push_back(ch); }
      else if (ch == ';' || ch == ',' || ch == '\n') flush();
      else if (ch == '{') { flush(); emit(BLOCK_BEGIN, "", ""); }
      else if (ch == '}') { flush(); emit(BLOCK_END, "", ""); }
      else stmt.push_back(ch);
    }
    flush();
  }
  static std::string trim(const std::string& s) {
    const char* ws = " \t\r\n";
    auto b = s.find_first_not_of(ws);
    if (b == std::string::npos) return "";
    auto e = s.find_last_not_of(ws);
    return s.substr(b, e - b + 1);
  }
  // a plain value, else (numbers only) an expression
  template <class T>
  T convert(const std::string& text, const std::string& k, T*) const {
    {
      std::istringstream is(text);
      T v;
      if ((is >> v) && (is >> std::ws).eof()) return v;
    }
    if (!std::is_arithmetic<T>::value) throw std::invalid_argument("parameter " + k + " = '" + text + "' is not readable");
    return from_double<T>(evaluate(text, k, 0));
  }
  std::string convert(const std::string& text, const std::string&, std::string*) const { return text; }
  template <class T>
  static T from_double(double x) {
    if (std::is_integral<T>::value) {
      if (std::abs(x - std::round(x)) > 1e-9 * std::max(1.0, std::abs(x))) throw std::invalid_argument("an integer parameter evaluates to a fraction");
      return static_cast<T>(std::llround(x));
    }
    return static_cast<T>(x);
  }
  // expr := term (('+' | '-') term)* ; term := factor (('*' | '/') factor)* ;
  // factor := number | '(' expr ')' | '-' factor | '+' factor | name      (name: another parameter, or pi)
  double evaluate(const std::string& text, const std::string& key, int depth) const {
    if (depth > 16) throw std::invalid_argument("parameter " + key + " refers to itself");
    size_t pos = 0;
    auto bad = [&]() { return std::invalid_argument("parameter " + key + " = '" + text + "' is not readable"); };
    auto skip = [&]() { while (pos < text.size() && std::isspace((unsigned char)text[pos])) ++pos; };
    std::function<double()> expr, term, factor;
    factor = [&]() -> double {
      skip();
      if (pos >= text.size()) throw bad();
      const char c = text[pos];
      if (c == '(') { ++pos; const double v = expr(); skip(); if (pos >= text.size() || text[pos] != ')') throw bad(); ++pos; return v; }
      if (c == '-') { ++pos; return -factor(); }
      if (c == '+') { ++pos; return factor(); }
      if (std::isdigit((unsigned char)c) || c == '.') {
        const char* b = text.c_str() + pos;
        char* e = nullptr;
        const double v = std::strtod(b, &e);
        if (e == b) throw bad();
        pos += size_t(e - b);
        return v;
      }
      if (std::isalpha((unsigned char)c) || c == '_') {
        size_t e = pos;
        while (e < text.size() && (std::isalnum((unsigned char)text[e]) || text[e] == '_')) ++e;
        const std::string name = text.substr(pos, e - pos);
        pos = e;
        if (name == "pi" || name == "Pi" || name == "PI") return 3.14159265358979323846;
        auto it = kv_.find(name);
        if (it == kv_.end() || name == key) throw bad();
        return evaluate(it->second, name, depth + 1);
      }
      throw bad();
    };
    term = [&]() -> double {
      double v = factor();
      for (;;) {
        skip();
        if (pos < text.size() && text[pos] == '*') { ++pos; v *= factor(); }
        else if (pos < text.size() && text[pos] == '/') { ++pos; v /= factor(); }
        else return v;
      }
    };
    expr = [&]() -> double {
      double v = term();
      for (;;) {
        skip();
        if (pos < text.size() && text[pos] == '+') { ++pos; v += term(); }
        else if (pos < text.size() && text[pos] == '-') { ++pos; v -= term(); }
        else return v;
      }
    };
    const double v = expr();
    skip();
    if (pos != text.size()) throw bad();
    return v;
  }
  std::map<std::string, std::string> kv_;
  std::vector<std::string> task_keys_;
};

}  // namespace looper
