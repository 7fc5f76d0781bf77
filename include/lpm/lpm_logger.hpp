// lpm/lpm_logger.hpp -- a console logger with the call surface the mesh and remesh classes use
// (Logger<>::debug/info/warn/error with "{}" placeholders; src/lpm_logger.hpp:60-140 wraps spdlog, which is not needed here).
#ifndef LPM_SHIM_LOGGER_HPP
#define LPM_SHIM_LOGGER_HPP

#include <cstdio>
#include <sstream>
#include <string>

namespace Lpm {

struct Log {
  enum Level { debug = 0, info = 1, warn = 2, error = 3, none = 4 };
};

class Logger {
 public:
  explicit Logger(const std::string& name = "lpm", const Log::Level level = Log::info) : name_(name), level_(level) {}
  void set_level(const Log::Level l) { level_ = l; }
  template <typename... Args>
  void debug(const char* fmt, const Args&... args) const { emit(Log::debug, "debug", fmt, args...); }
  template <typename... Args>
  void info(const char* fmt, const Args&... args) const { emit(Log::info, "info", fmt, args...); }
  template <typename... Args>
  void warn(const char* fmt, const Args&... args) const { emit(Log::warn, "warning", fmt, args...); }
  template <typename... Args>
  void error(const char* fmt, const Args&... args) const { emit(Log::error, "error", fmt, args...); }
  /// number of messages emitted at `level` or above (tests look at warnings)
  int count(const Log::Level level) const {
    int n = 0;
    for (int l = level; l < Log::none; ++l) n += counts_[l];
    return n;
  }

 private:
  static void format(std::ostringstream& ss, const char* fmt) { ss << fmt; }
  template <typename T, typename... Rest>
  static void format(std::ostringstream& ss, const char* fmt, const T& v, const Rest&... rest) {
    for (; *fmt; ++fmt) {
      if (fmt[0] == '{' && fmt[1] == '}') {
        ss << v;
        format(ss, fmt + 2, rest...);
        return;
      }
      ss << *fmt;
    }
  }
  template <typename... Args>
  void emit(const Log::Level l, const char* tag, const char* fmt, const Args&... args) const {
    ++counts_[l];
    if (l < level_) return;
    std::ostringstream ss;
    format(ss, fmt, args...);
    std::printf("[%s] [%s] %s\n", name_.c_str(), tag, ss.str().c_str());
  }
  std::string name_;
  Log::Level level_;
  mutable int counts_[4] = {0, 0, 0, 0};
};

}  // namespace Lpm
#endif
