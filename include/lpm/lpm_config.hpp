// lpm/lpm_config.hpp -- basic types, constants, error handling and the process-wide engine handle of the
// C++ API shim over the C ABI (include/lpmx.h).
//
// The shim keeps the reference's public names so that the reference's example drivers port line by line:
//   Real / Index / Int                     LpmConfig.h.in:31-32
//   constants::PI                          src/lpm_constants.hpp:11
//   LPM_REQUIRE / LPM_REQUIRE_MSG          src/lpm_assert.hpp:22-30,54-56 (throws std::runtime_error on host)
// Views live in host memory (as in the reference's OpenMP build); every O(N^2) call goes through the C ABI into
// the sm_100a kernels.  There is no CPU fallback: Engine::get() throws when no B200 is present.
#ifndef LPM_SHIM_CONFIG_HPP
#define LPM_SHIM_CONFIG_HPP

#define LPM_NULL_IDX (-1)      /* LpmConfig.h.in:20 */
#define LPM_MAX_AMR_LIMIT 6     /* LpmConfig.h.in:19 */

#include <cmath>
#include <cstdlib>
#include <sstream>
#include <stdexcept>
#include <string>

#include "../lpmx.h"

namespace Lpm {

typedef double Real;
typedef int Index;
typedef int Int;

namespace constants {
static constexpr Real PI = 3.1415926535897932384626433832795027975;
static constexpr Real ZERO_TOL = 2.220446049250313e-16;  // FloatingPoint<Real>::zero_tol, lpm_floating_point.hpp:22
static constexpr int NULL_IND = -1;                       // lpm_constants.hpp:31
}  // namespace constants

template <typename T>
inline T square(const T& x) { return x * x; }
template <typename T>
inline T cube(const T& x) { return x * x * x; }

#define LPM_REQUIRE_MSG(cond, msg)                                                                         \
  do {                                                                                                     \
    if (!(cond)) {                                                                                         \
      std::ostringstream _ss;                                                                              \
      _ss << "LPM_REQUIRE failed: " << #cond << " (" << __FILE__ << ":" << __LINE__ << ") " << (msg);      \
      throw std::runtime_error(_ss.str());                                                                 \
    }                                                                                                      \
  } while (0)
#define LPM_REQUIRE(cond) LPM_REQUIRE_MSG(cond, "")

/// Process-wide engine handle (one GPU per process; device from $LPMX_DEVICE or $LOCAL_RANK, default 0).
class Engine {
 public:
  static lpmx_handle_t get() { return instance().h_; }
  static void check(int rc, const char* where) {
    if (rc != LPMX_OK) {
      std::ostringstream ss;
      ss << where << ": " << lpmx_error_name(rc) << " -- " << lpmx_last_error_string(instance().h_);
      throw std::runtime_error(ss.str());
    }
  }
  static long launch_count() {
    long n = 0;
    lpmx_launch_count(get(), &n);
    return n;
  }
  static void sync() { check(lpmx_sync(get()), "lpmx_sync"); }

 private:
  lpmx_handle_t h_ = nullptr;
  Engine() {
    int dev = 0;
    if (const char* e = std::getenv("LPMX_DEVICE")) dev = std::atoi(e);
    else if (const char* l = std::getenv("LOCAL_RANK")) dev = std::atoi(l);
    const int rc = lpmx_create(&h_, dev);
    if (rc != LPMX_OK)
      throw std::runtime_error(std::string("lpmx_create failed: ") + lpmx_error_name(rc) +
                               " (the engine needs an sm_100 GPU; there is no CPU fallback)");
  }
  ~Engine() { lpmx_destroy(h_); }
  static Engine& instance() {
    static Engine e;
    return e;
  }
};

}  // namespace Lpm
#endif
