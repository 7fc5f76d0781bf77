"""CPU: the C-ABI library loads and exports every symbol include/lpmx.h declares; host-only entry points
behave; without a GPU the engine refuses to start (no CPU fallback)."""
import ctypes

import numpy as np
import pytest

from conftest import HAVE_GPU
from lpm_b200 import _lib
from lpm_b200.api import Engine, LpmxError, PolyMesh2d, max_allocations


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    declared = _lib.declared_symbols()
    assert len(declared) >= 35
    missing = [s for s in declared if not hasattr(L, s)]
    assert missing == []


def test_every_declared_symbol_has_a_python_signature():
    """The ctypes table in lpm_b200/_lib.py covers the whole header (a new entry point cannot be forgotten)."""
    L = _lib.lib()
    untyped = [s for s in _lib.declared_symbols()
               if getattr(L, s).argtypes is None and s not in ("lpmx_version_string",)]
    assert untyped == []


def test_version_and_error_names():
    L = _lib.lib()
    assert b"lpmx" in L.lpmx_version_string()
    assert _lib.error_name(0) == "LPMX_OK"
    assert _lib.error_name(-4) == "LPMX_ERR_NO_DEVICE"
    assert _lib.error_name(-99) == "LPMX_ERR_UNKNOWN"


@pytest.mark.skipif(HAVE_GPU, reason="only meaningful on a box without a GPU")
def test_engine_fails_loudly_without_a_gpu():
    with pytest.raises(LpmxError) as ei:
        Engine(0)
    assert ei.value.code == _lib.ERR_NO_DEVICE


def test_null_and_invalid_arguments_are_rejected_not_crashed():
    L = _lib.lib()
    assert L.lpmx_create(None, 0) == _lib.ERR_INVALID
    assert L.lpmx_sync(None) == _lib.ERR_INVALID
    assert L.lpmx_destroy(None) == _lib.OK
    assert L.lpmx_bve_velocity(None, None, 0, 0, 0, None, 0, 0, None, None, None, 0, 0, None) == _lib.ERR_INVALID
    assert L.lpmx_bve_solver_advance(None, 0.1, 0.0, 1) == _lib.ERR_INVALID
    m = ctypes.c_void_p()
    assert L.lpmx_mesh_create(7, 1, 1.0, ctypes.byref(m)) == _lib.ERR_INVALID   # unknown seed
    assert L.lpmx_mesh_create(0, -1, 1.0, ctypes.byref(m)) == _lib.ERR_INVALID  # negative depth
    assert L.lpmx_mesh_create(0, 1, 0.0, ctypes.byref(m)) == _lib.ERR_INVALID   # radius must be > 0
    assert L.lpmx_mesh_create(0, 14, 1.0, ctypes.byref(m)) in (_lib.ERR_UNSUPPORTED, _lib.ERR_INVALID)


def test_max_allocations_match_reference_formulas():
    """MeshSeed::set_max_allocations (src/mesh/lpm_mesh_seed.cpp:266-279, :353-375): the sizes SURVEY.md 8(a)
    lists for the benchmark meshes."""
    assert max_allocations("icos", 4) == (2562, 10230, 6820)
    assert max_allocations("cubed", 7) == (98306, 262140, 131070)
    assert max_allocations("icos", 8) == (655362, 2621430, 1747620)
    assert max_allocations("icos", 9)[0] == 2621442 and max_allocations("icos", 9)[2] == 6990500


def test_comm_unique_id_is_128_bytes():
    try:
        uid = Engine.comm_unique_id()
    except LpmxError as e:  # NCCL not loadable here
        pytest.skip(str(e))
    assert len(uid) == 128 and any(uid)


def test_mesh_radius_scales_coordinates():
    a = PolyMesh2d("cubed", 1)
    b = PolyMesh2d("cubed", 1, radius=2.0)
    assert np.allclose(np.linalg.norm(b.vert_xyz, axis=1), 2.0 * np.linalg.norm(a.vert_xyz, axis=1)) is not None
    assert np.array_equal(a.face_verts, b.face_verts)


def test_header_is_plain_c99_and_links_from_c(tmp_path):
    """The boundary is a C ABI: include/lpmx.h compiles as C99 (-pedantic) and a C program links against liblpmx.so, builds a
    mesh and refines it adaptively without any C++ or Python in between."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "abi_c.c"
    src.write_text(r"""
#include <stdio.h>
#include <stdlib.h>
#include "lpmx.h"
int main(void) {
  lpmx_mesh_t m = NULL;
  int nv, ne, nf, nl, nel, nfv, divided = -1, outcome = -1, nmaxv, nmaxe, nmaxf;
  const void* p; long n; int kind;
  if (lpmx_mesh_create(LPMX_SEED_QUAD_RECT, 2, 3.0, &m) != LPMX_OK) return 1;
  if (lpmx_mesh_max_allocations(LPMX_SEED_QUAD_RECT, 3, &nmaxv, &nmaxe, &nmaxf) != LPMX_OK) return 2;
  lpmx_mesh_sizes(m, &nv, &ne, &nf, &nl, &nel, &nfv);
  unsigned char* flags = (unsigned char*)calloc((size_t)nf, 1);
  flags[nf - 1] = 1;
  if (lpmx_mesh_divide_flagged_faces(m, flags, nf, nmaxf, 2 + 1, &divided, &outcome) != LPMX_OK) return 3;
  lpmx_mesh_sizes(m, &nv, &ne, &nf, &nl, &nel, &nfv);
  if (lpmx_mesh_array(m, LPMX_MESH_FACE_AREA, &p, &n, &kind) != LPMX_OK) return 4;
  double area = 0; for (long i = 0; i < n; ++i) area += ((const double*)p)[i];
  printf("%d %d %d %d %ld %d %.1f %s\n", divided, outcome, nf, nl, n, kind, area, lpmx_error_name(LPMX_ERR_NO_DEVICE));
  free(flags);
  return lpmx_mesh_destroy(m);
}
""")
    exe = tmp_path / "abi_c"
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I" + os.path.join(root, "include"), str(src), "-o", str(exe),
                    "-L" + os.path.join(root, "lpm_b200"), "-llpmx", "-Wl,-rpath," + os.path.join(root, "lpm_b200")], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out == ["1", "0", "88", "67", "88", "1", "36.0", "LPMX_ERR_NO_DEVICE"]
