// lpm/lpm_vtk_interfaces.hpp -- vtk_mesh_interface(model): the VtkPolymeshInterface of a model with every field attached,
// in the reference's order and under the reference's array names (the view labels).
//   vtk_mesh_interface(const Incompressible2D<Seed>&)   src/lpm_incompressible2d_impl.hpp:298-317
//   vtk_mesh_interface(const SWE<Seed>&)                 src/lpm_swe_impl.hpp:489-530
// The reference's SWE version lists the four velocity-gradient entries for every geometry; on the sphere those fields
// do not exist in this shim (the spherical kernels return double_dot only), so they are written for planar models only.
#ifndef LPM_SHIM_VTK_INTERFACES_HPP
#define LPM_SHIM_VTK_INTERFACES_HPP

#include "lpm_plane.hpp"
#include "lpm_vtk_io.hpp"

namespace Lpm {

template <typename SeedType>
VtkPolymeshInterface<SeedType> vtk_mesh_interface(const Incompressible2D<SeedType>& ic2d) {
  VtkPolymeshInterface<SeedType> vtk(ic2d.mesh);
  vtk.add_scalar_point_data(ic2d.rel_vort_passive.view);
  vtk.add_scalar_point_data(ic2d.stream_fn_passive.view);
  vtk.add_vector_point_data(ic2d.velocity_passive.view);
  vtk.add_vector_point_data(ic2d.ref_crds_passive.view);
  vtk.add_scalar_cell_data(ic2d.rel_vort_active.view);
  vtk.add_scalar_cell_data(ic2d.stream_fn_active.view);
  vtk.add_vector_cell_data(ic2d.velocity_active.view);
  vtk.add_vector_cell_data(ic2d.ref_crds_active.view);
  vtk.add_scalar_cell_data(ic2d.ftle.view);
  for (const auto& tracer : ic2d.tracer_passive) {
    vtk.add_scalar_point_data(tracer.second.view, tracer.first);
    vtk.add_scalar_cell_data(ic2d.tracer_active.at(tracer.first).view, tracer.first);
  }
  return vtk;
}

template <typename SeedType>
VtkPolymeshInterface<SeedType> vtk_mesh_interface(const SWE<SeedType>& swe) {
  constexpr bool planar = std::is_same<typename SeedType::geo, PlaneGeometry>::value;
  VtkPolymeshInterface<SeedType> vtk(swe.mesh);
  vtk.add_scalar_point_data(swe.rel_vort_passive.view);
  vtk.add_scalar_point_data(swe.pot_vort_passive.view);
  vtk.add_scalar_point_data(swe.div_passive.view);
  vtk.add_scalar_point_data(swe.surf_passive.view);
  vtk.add_scalar_point_data(swe.surf_lap_passive.view);
  vtk.add_scalar_point_data(swe.depth_passive.view);
  vtk.add_scalar_point_data(swe.double_dot_passive.view);
  if constexpr (planar) {
    vtk.add_scalar_point_data(swe.du1dx1_passive.view);
    vtk.add_scalar_point_data(swe.du1dx2_passive.view);
    vtk.add_scalar_point_data(swe.du2dx1_passive.view);
    vtk.add_scalar_point_data(swe.du2dx2_passive.view);
  }
  vtk.add_scalar_point_data(swe.bottom_passive.view);
  vtk.add_scalar_point_data(swe.stream_fn_passive.view);
  vtk.add_scalar_point_data(swe.potential_passive.view);
  vtk.add_vector_point_data(swe.velocity_passive.view);
  vtk.add_scalar_cell_data(swe.rel_vort_active.view);
  vtk.add_scalar_cell_data(swe.pot_vort_active.view);
  vtk.add_scalar_cell_data(swe.div_active.view);
  vtk.add_scalar_cell_data(swe.surf_active.view);
  vtk.add_scalar_cell_data(swe.depth_active.view);
  vtk.add_scalar_cell_data(swe.surf_lap_active.view);
  vtk.add_scalar_cell_data(swe.double_dot_active.view);
  if constexpr (planar) {
    vtk.add_scalar_cell_data(swe.du1dx1_active.view);
    vtk.add_scalar_cell_data(swe.du1dx2_active.view);
    vtk.add_scalar_cell_data(swe.du2dx1_active.view);
    vtk.add_scalar_cell_data(swe.du2dx2_active.view);
  }
  vtk.add_scalar_cell_data(swe.bottom_active.view);
  vtk.add_vector_cell_data(swe.velocity_active.view);
  vtk.add_scalar_cell_data(swe.mass_active.view);
  vtk.add_scalar_cell_data(swe.stream_fn_active.view);
  vtk.add_scalar_cell_data(swe.potential_active.view);
  for (const auto& tracer : swe.tracer_passive) {
    vtk.add_scalar_point_data(tracer.second.view, tracer.first);
    vtk.add_scalar_cell_data(swe.tracer_active.at(tracer.first).view, tracer.first);
  }
  return vtk;
}

/// frame file name of the reference drivers: <root><zero-filled counter>.vtp (zero_fill_str + vtp_suffix, src/util/lpm_string_util.hpp)
inline std::string vtk_frame_name(const std::string& root, const int counter) {
  char buf[16];
  std::snprintf(buf, sizeof(buf), "%04d", counter);
  return root + buf + ".vtp";
}

}  // namespace Lpm
#endif
