// LpmConfig.h -- what CMake would generate from /root/reference/LpmConfig.h.in (oracle/_ref only):
// Index = int, Real = double (LpmConfig.h.in:31-32); no optional packages.
#ifndef LPM_CONFIG_H
#define LPM_CONFIG_H
// MeshSeed<Seed>::full_filename() does std::string(LPM_MESH_SEED_DIR) (src/mesh/lpm_mesh_seed.cpp:210): the directory comes from
// $LPM_ORACLE_SEED_DIR when set (oracle/ref_mesh.py points it at seed files rewritten from tests/golden/seed_tables.npz on
// machines without /root/reference, e.g. the GPU box), else the reference's own directory.
#include <cstdlib>
inline const char* oracle_shim_seed_dir() {
  const char* e = std::getenv("LPM_ORACLE_SEED_DIR");
  return (e && *e) ? e : "/root/reference/mesh_seeds";
}
#define LPM_MESH_SEED_DIR oracle_shim_seed_dir()
#define LPM_TEST_DATA_DIR "/root/reference/tests/data"
#define LPM_MAX_AMR_LIMIT 6
#define LPM_NULL_IDX -1
#include <map>
#include <string>
#include "Kokkos_Core.hpp"
namespace Lpm {
typedef int Index;
typedef double Real;
typedef int Int;
typedef unsigned Uint;
typedef short Short;
typedef std::map<std::string, std::string> metadata_type;
}  // namespace Lpm
#endif
