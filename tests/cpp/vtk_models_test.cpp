// vtk_models_test.cpp -- host-only: vtk_mesh_interface(model) for a planar SWE model and a spherical Incompressible2D model
// (constructors and the init_* functors never touch the engine).  Usage: vtk_models_test <swe.vtp> <ic2d.vtp>
#include <cstdio>

#include "lpm/lpm.hpp"

using namespace Lpm;

struct Lat0 {
  template <typename CV>
  Real operator()(const CV& x) const { return SphereGeometry::latitude(x); }
  std::string name() const { return "lat0"; }
};

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  try {
    {
      PolyMeshParameters<QuadRectSeed> params(2, 6.0);
      SWE<QuadRectSeed> plane(params, CoriolisBetaPlane(0.5, 0.1));
      plane.init_surface(PlanarGaussianMountain(), PlanarGaussianSurfacePerturbation());
      vtk_mesh_interface(plane).write(argv[1]);
    }
    {
      PolyMeshParameters<CubedSphereSeed> params(2);
      Incompressible2D<CubedSphereSeed> sphere(params, CoriolisSphere(), 0.0);
      sphere.init_vorticity(GaussianVortexSphere());
      Lat0 lat0;
      sphere.allocate_tracer(lat0);
      sphere.init_tracer(lat0);
      vtk_mesh_interface(sphere).write(vtk_frame_name(std::string(argv[2]) + "_", 7));
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 3;
  }
  return 0;
}
