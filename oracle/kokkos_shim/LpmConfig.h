// LpmConfig.h -- what CMake would generate from /root/reference/LpmConfig.h.in (oracle/_ref only):
// Index = int, Real = double (LpmConfig.h.in:31-32); no optional packages.
#ifndef LPM_CONFIG_H
#define LPM_CONFIG_H
#define LPM_MESH_SEED_DIR "/root/reference/mesh_seeds"
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
