// io_formats_test.cpp -- test driver (built by tests/test_io_formats.py): writes a .vtp and a .m file for a small mesh
// through the shim's VtkPolymeshInterface / write_*_matlab so the Python side can parse them.  Host only (no GPU).
//   usage: io_formats_test <cubed|icos> <depth> <out.vtp> <out.m>
#include <fstream>
#include <string>

#include "lpm/lpm.hpp"
#include "lpm/lpm_matlab_io.hpp"
#include "lpm/lpm_vtk_io.hpp"

using namespace Lpm;

template <typename Seed>
int run(int depth, const std::string& vtp, const std::string& mfile) {
  PolyMeshParameters<Seed> params(depth, 1.0, 0, 0);
  PolyMesh2d<Seed> mesh(params);
  const Index nv = mesh.n_vertices_host(), nf = mesh.n_faces_host();
  scalar_view_type vs("vertex_scalar", nv), fs("face_scalar", nf);
  vec3_view_type vv("vertex_vector", nv), fv("face_vector", nf);
  for (Index i = 0; i < nv; ++i) {
    vs(i) = 0.1 * i + 1.0 / 3.0;
    for (int k = 0; k < 3; ++k) vv(i, k) = mesh.vertices.phys_crds.view(i, k) * (k + 1);
  }
  for (Index i = 0; i < nf; ++i) {
    fs(i) = -0.5 * i;
    for (int k = 0; k < 3; ++k) fv(i, k) = mesh.faces.phys_crds.view(i, k) - k;
  }
  VtkPolymeshInterface<Seed> vtk(mesh);
  vtk.add_scalar_point_data(vs);
  vtk.add_vector_point_data(vv, "renamed_vector");
  vtk.add_scalar_cell_data(fs);
  vtk.add_vector_cell_data(fv);
  vtk.write(vtp);
  std::ofstream m(mfile);
  std::vector<Real> t = {1.0, 2.5, 1.0 / 3.0, 1e-7, 123456789.0};
  write_vector_matlab(m, "t", t);
  write_vector_matlab(m, "area", mesh.faces.area);
  write_array_matlab(m, "xyz", mesh.vertices.phys_crds.view);
  return 0;
}

int main(int argc, char** argv) {
  if (argc != 5) return 2;
  const int depth = std::stoi(argv[2]);
  return std::string(argv[1]) == "icos" ? run<IcosTriSphereSeed>(depth, argv[3], argv[4])
                                         : run<CubedSphereSeed>(depth, argv[3], argv[4]);
}
