// lpm/lpm.hpp -- umbrella header of the C++ API shim (reference class names over the lpmx C ABI).
#ifndef LPM_SHIM_HPP
#define LPM_SHIM_HPP
#include "lpm_bve_sphere.hpp"
#include "lpm_compadre_remesh.hpp"
#include "lpm_config.hpp"
#include "lpm_coords.hpp"
#include "lpm_coriolis.hpp"
#include "lpm_error.hpp"
#include "lpm_ftle.hpp"
#include "lpm_gallery.hpp"
#include "lpm_geometry.hpp"
#include "lpm_incompressible2d.hpp"
#include "lpm_logger.hpp"
#include "lpm_plane.hpp"
#include "lpm_polymesh2d.hpp"
#include "lpm_refinement.hpp"
#include "lpm_swe.hpp"
#include "lpm_views.hpp"
#include "lpm_vtk_interfaces.hpp"
#endif
