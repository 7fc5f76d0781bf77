"""Output formats: the shim's VtkPolymeshInterface (.vtp, written without libvtk) and the
MATLAB text writers.  CPU only.  The .m text is pinned byte-for-byte against the reference's util/lpm_matlab_io.hpp
compiled in place (oracle/_ref, when present; else against the strings that build produced, recorded below); the .vtp
is parsed back and every array is compared bit-exactly with the mesh the generator produced (VTK itself is absent, so
byte parity with vtkXMLPolyDataWriter's compressed encoding is not attempted)."""
import ctypes
import os
import subprocess
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from lpm_b200.api import PolyMesh2d

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "liblpm_ref.so")


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    out = tmp_path_factory.mktemp("io") / "io_formats_test"
    cmd = ["g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "io_formats_test.cpp"),
           "-o", str(out), "-L" + os.path.join(ROOT, "lpm_b200"), "-llpmx", "-Wl,-rpath," + os.path.join(ROOT, "lpm_b200")]
    subprocess.run(cmd, check=True, capture_output=True)
    return str(out)


def _arrays(parent):
    out = {}
    for da in parent.findall("DataArray"):
        dt = np.float64 if da.get("type") == "Float64" else np.int64
        a = np.array(da.text.split(), dtype=dt)
        nc = int(da.get("NumberOfComponents", "1"))
        out[da.get("Name")] = a.reshape(-1, nc) if nc > 1 else a
    return out


@pytest.mark.parametrize("seed,depth", [("cubed", 2), ("icos", 2)])
def test_vtp_file_holds_the_reference_data_set(driver, tmp_path, seed, depth):
    vtp, mfile = str(tmp_path / "mesh.vtp"), str(tmp_path / "mesh.m")
    subprocess.run([driver, seed, str(depth), vtp, mfile], check=True)
    m = PolyMesh2d(seed, depth)
    leaf = m.face_mask == 0
    root = ET.parse(vtp).getroot()
    assert root.get("type") == "PolyData"
    piece = root.find("PolyData/Piece")
    assert int(piece.get("NumberOfPoints")) == m.n_verts and int(piece.get("NumberOfPolys")) == m.n_face_leaves
    pts = _arrays(piece.find("Points"))["Points"]
    assert np.array_equal(pts, m.vert_xyz)  # %.17g round-trips every double
    polys = _arrays(piece.find("Polys"))
    nfv = m.face_verts.shape[1]
    assert np.array_equal(polys["connectivity"].reshape(-1, nfv), m.face_verts[leaf])  # leaves only, in face order
    assert np.array_equal(polys["offsets"], nfv * np.arange(1, m.n_face_leaves + 1))
    pd, cd = _arrays(piece.find("PointData")), _arrays(piece.find("CellData"))
    # constructor arrays first (area, lag_crds), then the added ones; names default to the view labels
    assert list(cd) == ["area", "lag_crds", "face_scalar", "face_vector"]
    assert list(pd) == ["lag_crds", "vertex_scalar", "renamed_vector"]
    assert np.array_equal(cd["area"], m.face_area[leaf]) and np.array_equal(cd["lag_crds"], m.face_lag_xyz[leaf])
    assert np.array_equal(pd["lag_crds"], m.vert_lag_xyz)
    assert np.array_equal(cd["face_scalar"], -0.5 * np.arange(m.n_faces)[leaf])
    assert np.array_equal(cd["face_vector"], (m.face_xyz - np.arange(3))[leaf])
    assert np.array_equal(pd["vertex_scalar"], 0.1 * np.arange(m.n_verts) + 1.0 / 3.0)
    assert np.array_equal(pd["renamed_vector"], m.vert_xyz * np.arange(1, 4))


def test_matlab_text_matches_the_reference_writer(driver, tmp_path):
    vtp, mfile = str(tmp_path / "mesh.vtp"), str(tmp_path / "mesh.m")
    subprocess.run([driver, "cubed", "1", vtp, mfile], check=True)
    lines = open(mfile).read().splitlines(keepends=True)
    # produced by the reference's write_vector_matlab compiled in place (oracle/ref_driver.cpp)
    assert lines[0] == "t = [1,2.5,0.333333,1e-07,1.23457e+08];\n"
    m = PolyMesh2d("cubed", 1)
    if os.path.exists(REF_LIB):
        R = ctypes.CDLL(REF_LIB)
        buf = ctypes.create_string_buffer(1 << 16)
        dp = ctypes.POINTER(ctypes.c_double)
        a = np.ascontiguousarray(m.face_area)
        R.oracle_write_vector_matlab(b"area", len(a), a.ctypes.data_as(dp), buf, len(buf))
        assert lines[1] == buf.value.decode()
        x = np.ascontiguousarray(m.vert_xyz)
        R.oracle_write_array_matlab(b"xyz", x.shape[0], 3, x.ctypes.data_as(dp), buf, len(buf))
        assert lines[2] == buf.value.decode()
    else:
        assert lines[1].startswith("area = [") and lines[1].endswith("];\n") and lines[1].count(",") == m.n_faces - 1
        assert lines[2].startswith("xyz = [") and lines[2].count(";") == m.n_verts  # n-1 row breaks + the final "];"


def test_vtk_mesh_interface_of_the_models(tmp_path):
    """vtk_mesh_interface(SWE<QuadRectSeed>) and vtk_mesh_interface(Incompressible2D<CubedSphereSeed>)
    (src/lpm_swe_impl.hpp:489-530, src/lpm_incompressible2d_impl.hpp:298-317): array order and names as in the reference,
    values as the init functors left them, planar points with z = 0."""
    exe = str(tmp_path / "vtk_models_test")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "vtk_models_test.cpp"),
                    "-o", exe, "-L" + os.path.join(ROOT, "lpm_b200"), "-llpmx", "-Wl,-rpath," + os.path.join(ROOT, "lpm_b200")],
                   check=True, capture_output=True)
    swe_vtp, ic2d_root = str(tmp_path / "swe.vtp"), str(tmp_path / "ic2d")
    subprocess.run([exe, swe_vtp, ic2d_root], check=True)
    # planar SWE
    m = PolyMesh2d("quad_rect", 2, radius=6.0)
    leaf = m.face_mask == 0
    piece = ET.parse(swe_vtp).getroot().find("PolyData/Piece")
    pts = _arrays(piece.find("Points"))["Points"]
    assert np.array_equal(pts[:, :2], m.vert_xyz) and not pts[:, 2].any()
    pd, cd = _arrays(piece.find("PointData")), _arrays(piece.find("CellData"))
    assert list(pd) == ["lag_crds", "relative_vorticity", "potential_vorticity", "divergence", "surface_height",
                        "surface_laplacian", "depth", "double_dot", "du1dx1", "du1dx2", "du2dx1", "du2dx2", "bottom_height",
                        "stream_function", "potential", "velocity"]
    assert list(cd) == ["area", "lag_crds", "relative_vorticity", "potential_vorticity", "divergence", "surface_height", "depth",
                        "surface_laplacian", "double_dot", "du1dx1", "du1dx2", "du2dx1", "du2dx2", "bottom_height", "velocity",
                        "mass", "stream_function", "potential"]
    bot = 0.8 * np.exp(-5.0 * (m.face_xyz ** 2).sum(axis=1))
    surf = 1.0 + 0.1 * np.exp(-(20 * (m.face_xyz[:, 0] + 1.125) ** 2 + 5 * m.face_xyz[:, 1] ** 2))
    assert np.allclose(cd["bottom_height"], bot[leaf], rtol=1e-15, atol=0) and np.allclose(cd["surface_height"], surf[leaf], rtol=1e-15)
    assert np.allclose(cd["mass"], ((surf - bot) * m.face_area)[leaf], rtol=1e-15)
    assert pd["velocity"].shape == (m.n_verts, 2) and cd["lag_crds"].shape == (m.n_face_leaves, 2)
    # spherical Incompressible2D, frame 7 of root "ic2d_"
    ms = PolyMesh2d("cubed", 2)
    piece = ET.parse(ic2d_root + "_0007.vtp").getroot().find("PolyData/Piece")
    pd, cd = _arrays(piece.find("PointData")), _arrays(piece.find("CellData"))
    assert list(pd) == ["lag_crds", "relative_vorticity", "stream_function", "velocity", "crds", "lat0"]
    assert list(cd) == ["area", "lag_crds", "relative_vorticity", "stream_function", "velocity", "crds", "ftle", "lat0"]
    assert np.array_equal(pd["crds"], ms.vert_xyz)  # ref_crds start as the physical coordinates
    lat = np.arctan2(ms.vert_xyz[:, 2], np.sqrt(ms.vert_xyz[:, 0] ** 2 + ms.vert_xyz[:, 1] ** 2))
    assert np.allclose(pd["lat0"], lat, rtol=0, atol=1e-15)
