// lpm/lpm_coriolis.hpp -- CoriolisSphere (src/lpm_coriolis.hpp:154-195)
#ifndef LPM_SHIM_CORIOLIS_HPP
#define LPM_SHIM_CORIOLIS_HPP

#include "lpm_config.hpp"

namespace Lpm {

struct CoriolisSphere {
  Real Omega;
  explicit CoriolisSphere(const Real Omg = 2 * constants::PI) : Omega(Omg) {}
  template <typename PtType>
  Real f(const PtType& xyz) const { return 2 * Omega * xyz[2]; }
  template <typename UType>
  Real dfdt(const UType& u) const { return 2 * Omega * u[2]; }
  template <typename XType, typename UType>
  Real grad_f_cross_u(const XType& x, const UType& u) const { return -2 * Omega * (-u[0] * x[1] + u[1] * x[0]); }
};

}  // namespace Lpm
#endif
