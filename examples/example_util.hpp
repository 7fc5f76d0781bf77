// example_util.hpp -- tiny helpers shared by the example drivers (option parsing, wall-clock timer, JSON line).
#ifndef LPMX_EXAMPLE_UTIL_HPP
#define LPMX_EXAMPLE_UTIL_HPP
#include <chrono>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>

struct Options {
  std::map<std::string, std::string> kv;
  Options(int argc, char** argv) {
    for (int i = 1; i < argc; ++i) {
      std::string k = argv[i];
      if (k == "-h" || k == "--help") kv["help"] = "1";
      else if (i + 1 < argc) kv[k] = argv[++i];
    }
  }
  bool has(const std::string& k) const { return kv.count(k) > 0; }
  int get_int(const std::string& k, int dflt) const { return has(k) ? std::atoi(kv.at(k).c_str()) : dflt; }
  double get_real(const std::string& k, double dflt) const { return has(k) ? std::atof(kv.at(k).c_str()) : dflt; }
  std::string get_str(const std::string& k, const std::string& dflt) const { return has(k) ? kv.at(k) : dflt; }
};

struct Timer {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  double seconds() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};
#endif
