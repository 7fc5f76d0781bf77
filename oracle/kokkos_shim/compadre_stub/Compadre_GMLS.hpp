// Compadre_GMLS.hpp / Compadre_Evaluator.hpp / Compadre_Operators.hpp / Compadre_PointCloudSearch.hpp -- DECLARATIONS-ONLY stand-in
// for the Compadre 1.6.2 interface the reference's headers mention (lpm_compadre.hpp, mesh/lpm_compadre_remesh{,_impl}.hpp).
// TEST INFRASTRUCTURE (oracle/_ref only); our own code written from the reference's call sites, not from Compadre's sources.
// Purpose: let /root/reference/src/lpm_incompressible2d{,_impl}.hpp and lpm_incompressible2d_rk2_impl.hpp -- which include the
// Compadre-based remesh headers -- be compiled IN PLACE so that Incompressible2DRK2::advance_timestep_impl (no Compadre call on
// its path) can be run and compared with the oracle.  Nothing here computes anything: every member that would need Compadre
// aborts.  The GMLS values themselves stay "parity unpinned" (DESIGN.md section 3).
#ifndef ORACLE_SHIM_COMPADRE_STUB_HPP
#define ORACLE_SHIM_COMPADRE_STUB_HPP
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "Kokkos_Core.hpp"

namespace Compadre {
[[noreturn]] inline void oracle_stub_abort(const char* what) {
  std::fprintf(stderr, "Compadre stand-in: %s called, but Compadre is not available in the oracle build\n", what);
  std::abort();
}
enum ReconstructionSpace { ScalarTaylorPolynomial, VectorTaylorPolynomial, VectorOfScalarClonesTaylorPolynomial };
enum ProblemType { STANDARD, MANIFOLD };
enum DenseSolverType { QR, LU };
enum ConstraintType { NO_CONSTRAINT, NEUMANN_GRAD_SCALAR };
enum WeightingFunctionType { Power, Gaussian, CubicSpline };
enum TargetOperation {
  ScalarPointEvaluation,
  VectorPointEvaluation,
  LaplacianOfScalarPointEvaluation,
  GradientOfScalarPointEvaluation,
  DivergenceOfVectorPointEvaluation,
  CurlOfVectorPointEvaluation,
  GaussianCurvaturePointEvaluation
};
struct SamplingFunctional {
  int id;
  constexpr bool operator==(const SamplingFunctional& o) const { return id == o.id; }
};
constexpr SamplingFunctional PointSample{0}, VectorPointSample{1}, ManifoldVectorPointSample{2};

class GMLS {
 public:
  GMLS(ReconstructionSpace, SamplingFunctional, int /*poly order*/, int /*dimension*/ = 3, const char* /*solver*/ = "QR",
       const char* /*problem*/ = "STANDARD", const char* /*constraint*/ = "NO_CONSTRAINT", int /*manifold order*/ = 2) {}
  GMLS(ReconstructionSpace, SamplingFunctional, SamplingFunctional, int, int = 3, const char* = "QR", const char* = "STANDARD",
       const char* = "NO_CONSTRAINT", int = 2) {}
  template <class... A>
  GMLS(ReconstructionSpace, SamplingFunctional, int, int, DenseSolverType, ProblemType, ConstraintType, A...) {}
  template <class... A>
  GMLS(ReconstructionSpace, SamplingFunctional, SamplingFunctional, int, int, DenseSolverType, ProblemType, ConstraintType, A...) {}
  static int getNP(const int m, const int dimension = 3, ReconstructionSpace = ScalarTaylorPolynomial) {
    // number of monomials of total degree <= m in `dimension` variables (the published definition)
    if (dimension == 3) return (m + 1) * (m + 2) * (m + 3) / 6;
    if (dimension == 2) return (m + 1) * (m + 2) / 2;
    return m + 1;
  }
  template <class... A>
  void setProblemData(A&&...) { oracle_stub_abort("GMLS::setProblemData"); }
  template <class... A>
  void addTargets(A&&...) { oracle_stub_abort("GMLS::addTargets"); }
  template <class... A>
  void setWeightingType(A&&...) { oracle_stub_abort("GMLS::setWeightingType"); }
  template <class... A>
  void setWeightingParameter(A&&...) { oracle_stub_abort("GMLS::setWeightingParameter"); }
  template <class... A>
  void setCurvatureWeightingType(A&&...) { oracle_stub_abort("GMLS::setCurvatureWeightingType"); }
  template <class... A>
  void setCurvatureWeightingParameter(A&&...) { oracle_stub_abort("GMLS::setCurvatureWeightingParameter"); }
  template <class... A>
  void setReferenceOutwardNormalDirection(A&&...) { oracle_stub_abort("GMLS::setReferenceOutwardNormalDirection"); }
  template <class... A>
  void generateAlphas(A&&...) { oracle_stub_abort("GMLS::generateAlphas"); }
};

class Evaluator {
 public:
  explicit Evaluator(GMLS*) {}
  template <class DataT, class Mem, class V>
  Kokkos::View<DataT> applyAlphasToDataAllComponentsAllTargetSites(const V&, TargetOperation, SamplingFunctional = PointSample,
                                                                   bool = true, int = 0) const {
    oracle_stub_abort("Evaluator::applyAlphasToDataAllComponentsAllTargetSites");
  }
};

template <class V>
class PointCloudSearch {
 public:
  explicit PointCloudSearch(const V&, int = 3) {}
  template <class... A>
  std::size_t generate2DNeighborListsFromKNNSearch(A&&...) { oracle_stub_abort("PointCloudSearch::generate2DNeighborListsFromKNNSearch"); }
  template <class... A>
  std::size_t generate2DNeighborListsFromRadiusSearch(A&&...) { oracle_stub_abort("PointCloudSearch::generate2DNeighborListsFromRadiusSearch"); }
};
template <class V>
PointCloudSearch<V> CreatePointCloudSearch(const V& v, int dim = 3) {
  return PointCloudSearch<V>(v, dim);
}
}  // namespace Compadre
#endif
