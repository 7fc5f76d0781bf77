"""The C++ API shim (include/lpm/*.hpp: the reference's class names over the C ABI) and the four config drivers in
examples/ (bve_rotation, sphere_rh54, sphere_gaussian_vortex, sphere_swe_tc2).  CPU: they build with the host compiler
alone and refuse to run without a GPU.  GPU: they run the reference's smoke configurations
(examples/CMakeLists.txt:143-167: `bve_rotation -d 3 -dt 0.01 -tf 0.03`) and pass their own acceptance checks."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "examples", "_build")
EXAMPLES = ["bve_rotation", "sphere_rh54", "sphere_gaussian_vortex", "sphere_swe_tc2", "plane_gravity_wave",
            "plane_colliding_dipoles"]


@pytest.fixture(scope="module")
def built():
    from lpm_b200 import build
    build.build()
    subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True, capture_output=True)
    return BUILD


def test_examples_build_with_the_host_compiler(built):
    for name in EXAMPLES:
        assert os.access(os.path.join(built, name), os.X_OK), name


def test_examples_refuse_to_run_without_a_gpu(built):
    from conftest import HAVE_GPU
    if HAVE_GPU:
        pytest.skip("a GPU is present")
    p = subprocess.run([os.path.join(built, "bve_rotation"), "-d", "2"], capture_output=True, text=True)
    assert p.returncode == 2
    assert "LPMX_ERR_NO_DEVICE" in p.stderr and "no CPU fallback" in p.stderr


def _run(built, name, *args):
    p = subprocess.run([os.path.join(built, name), *args], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1]
    return json.loads(line), p.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("seed", ["cubed", "icos"])
def test_bve_rotation_ctest_case(built, seed):
    out, log = _run(built, "bve_rotation", "-s", seed, "-d", "3", "-dt", "0.01", "-tf", "0.03")
    assert out["steps"] == 3 and out["gpu_launches"] > 0
    assert out["vel_l2"] < 0.1 and out["pos_l2"] < 1e-2  # first-order quadrature at depth 3
    assert "tfinal (velocity)" in log


@pytest.mark.gpu
def test_bve_rotation_converges_with_depth(built):
    e3, _ = _run(built, "bve_rotation", "-d", "3", "-dt", "0.01", "-tf", "0.01")
    e5, _ = _run(built, "bve_rotation", "-d", "5", "-dt", "0.005", "-tf", "0.01")  # same Courant number
    assert e5["vel_l2"] < 0.5 * e3["vel_l2"]


@pytest.mark.gpu
def test_bve_rotation_refuses_courant_number_above_one(built):
    """examples/bve_rotation.cpp:111-114: LPM_REQUIRE(cr < 1)."""
    p = subprocess.run([os.path.join(built, "bve_rotation"), "-d", "5", "-dt", "0.01", "-tf", "0.02"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 3 and "exceeds 1" in p.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sphere_rh54", "sphere_gaussian_vortex"])
def test_ic2d_examples(built, name):
    out, log = _run(built, name, "-d", "4", "-tf", "0.05", "-n", "5")
    assert out["steps"] == 5 and abs(out["t"] - 0.05) < 1e-12 and out["gpu_launches"] > 0
    assert out["ke_drift"] < 1e-2


@pytest.mark.gpu
@pytest.mark.parametrize("seed", ["cubed", "icos"])
def test_swe_tc2_example(built, seed):
    out, log = _run(built, "sphere_swe_tc2", "-s", seed, "-d", "3", "-tf", "0.02", "-n", "4")
    assert out["steps"] == 4 and out["gpu_launches"] > 0
    assert out["depth_l2"] < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("strategy", ["direct", "indirect"])
def test_ic2d_example_with_remeshing(built, strategy):
    """examples/sphere_rh54.cpp:255-300: rebuild the particle set every remesh_interval steps (CompadreRemesh uniform_*);
    the RH54 wave is steady in shape, so energy and enstrophy must survive two remeshes."""
    ref, _ = _run(built, "sphere_rh54", "-d", "4", "-tf", "0.06", "-n", "6")
    out, log = _run(built, "sphere_rh54", "-d", "4", "-tf", "0.06", "-n", "6", "-rm", "3", "-rs", strategy)
    assert "remeshes: 2" in log
    assert out["ke_drift"] < 5e-3 and out["enstrophy_drift"] < 5e-3
    assert abs(out["ke_drift"] - ref["ke_drift"]) < 5e-3


@pytest.mark.gpu
@pytest.mark.parametrize("lap", ["gmls", "exact"])
def test_swe_tc2_example_surface_laplacians(built, lap):
    """sphere_swe_tc2 with the reference's configuration (GMLS of order 4, examples/sphere_swe_tc2.cpp:136-139), evaluated
    on the device, and with the closed-form Laplacian: both keep the steady state."""
    out, log = _run(built, "sphere_swe_tc2", "-d", "4", "-tf", "0.02", "-n", "4", "-lap", lap)
    assert out["steps"] == 4 and out["depth_l2"] < 1e-3 and out["zeta_l2"] < 1e-2


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sphere_gaussian_vortex", "sphere_rh54"])
def test_ic2d_examples_with_adaptive_refinement(built, name):
    """examples/sphere_gaussian_vortex.cpp:89-118 / sphere_rh54.cpp:118-147: -amr 2 refines where |zeta| A exceeds a fraction
    of its maximum on the uniform mesh (flags on the device, division on the host), then steps on the mixed-level mesh."""
    uni, _ = _run(built, name, "-d", "3", "-tf", "0.03", "-n", "3")
    out, log = _run(built, name, "-d", "3", "-tf", "0.03", "-n", "3", "-amr", "2", "-c", "0.25")
    assert "amr is enabled with limit 2" in log and "faces divided" in log
    assert out["n_leaves"] > uni["n_leaves"] and out["max_level"] == 3 + 2 + 1
    assert out["steps"] == 3 and out["ke_drift"] < 1e-2


@pytest.mark.gpu
@pytest.mark.parametrize("strategy", ["direct", "indirect"])
def test_gaussian_vortex_adaptive_remesh(built, strategy):
    """examples/sphere_gaussian_vortex.cpp:205-245: with AMR on, every remesh builds a fresh uniform mesh, interpolates, then
    refines it adaptively (CompadreRemesh::adaptive_direct_remesh / adaptive_indirect_remesh)."""
    out, log = _run(built, "sphere_gaussian_vortex", "-d", "3", "-tf", "0.04", "-n", "4", "-amr", "1", "-c", "0.25", "-rm", "2",
                    "-rs", strategy)
    assert "remeshes: 2" in log
    assert out["max_level"] == 3 + 1 + 1 and out["ke_drift"] < 2e-2


@pytest.mark.gpu
def test_rh54_remesh_triggered_by_ftle(built):
    """examples/sphere_rh54.cpp:247-258: -rt ftle remeshes when the maximum FTLE exceeds -ftle."""
    out, log = _run(built, "sphere_rh54", "-d", "3", "-tf", "0.2", "-n", "8", "-rt", "ftle", "-ftle", "0.05")
    assert "triggered by ftle" in log and "remeshes: 0" not in log


@pytest.mark.gpu
@pytest.mark.parametrize("seed", ["quad", "tri"])
def test_plane_gravity_wave_example(built, seed):
    """examples/plane_gravity_wave.cpp with its defaults scaled down (depth 4 -> 3): SWE<QuadRectSeed> + SWERK4 over a Gaussian
    mountain.  The wave spreads: the crest drops, the depth stays positive, mass (a Lagrangian invariant) is conserved."""
    out, log = _run(built, "plane_gravity_wave", "-s", seed, "-d", "3", "-tf", "0.25", "-n", "5")
    assert out["steps"] == 5 and out["gpu_launches"] > 0
    assert out["mass_drift"] < 1e-14
    assert out["surf_max"] < out["surf_max0"] and out["min_depth"] > 0.1 and 0 < out["max_speed"] < 1.0
    assert "SWERK4: dt = 0.05" in log and "PlanarGaussianMountain" in log


@pytest.mark.gpu
def test_plane_colliding_dipoles_example(built):
    """examples/plane_colliding_dipoles.cpp: uniform mesh, then with two adaptive levels driven by the circulation and
    vorticity-variation flags; the total vorticity of the two opposite dipoles is zero."""
    uni, _ = _run(built, "plane_colliding_dipoles", "-d", "4", "-tf", "0.1", "-n", "4")
    assert uni["steps"] == 4 and abs(uni["total_vorticity"]) < 1e-12 and uni["ke_drift"] < 1e-2
    out, log = _run(built, "plane_colliding_dipoles", "-d", "4", "-tf", "0.1", "-n", "4", "-amr", "2", "-c", "0.2", "-zv", "0.3")
    assert "amr is enabled with limit 2" in log and "vorticity variation refinement count" in log
    assert out["n_leaves"] > uni["n_leaves"] and out["max_level"] == 4 + 2 + 1
    assert out["ke_drift"] < 1e-2


@pytest.mark.gpu
def test_drivers_write_vtp_frames(built, tmp_path):
    """-o <root> -of <n>: a .vtp frame of the whole model at t = 0 and after every n-th step, as the reference's LPM_USE_VTK
    blocks do (vtk_mesh_interface + VtkPolymeshInterface::write)."""
    import xml.etree.ElementTree as ET
    root = str(tmp_path / "gw")
    out, _ = _run(built, "plane_gravity_wave", "-d", "3", "-tf", "0.2", "-n", "4", "-o", root, "-of", "2")
    frames = sorted(f for f in os.listdir(tmp_path) if f.startswith("gw_quad_rect3_"))
    assert frames == ["gw_quad_rect3_0000.vtp", "gw_quad_rect3_0001.vtp", "gw_quad_rect3_0002.vtp"]
    piece = ET.parse(os.path.join(tmp_path, frames[-1])).getroot().find("PolyData/Piece")
    assert int(piece.get("NumberOfPoints")) == out["n_verts"] and int(piece.get("NumberOfPolys")) == out["n_leaves"]
    names = [da.get("Name") for da in piece.find("CellData").findall("DataArray")]
    assert "surface_height" in names and "du1dx1" in names and "mass" in names
    surf = [da for da in piece.find("CellData").findall("DataArray") if da.get("Name") == "surface_height"][0]
    vals = [float(t) for t in surf.text.split()]
    assert abs(max(vals) - out["surf_max"]) < 1e-8 and abs(min(vals) - out["surf_min"]) < 1e-8
    root2 = str(tmp_path / "rh")
    _run(built, "sphere_rh54", "-d", "3", "-tf", "0.02", "-n", "2", "-o", root2)
    assert sorted(f for f in os.listdir(tmp_path) if f.startswith("rh_")) == [f"rh_cubed_sphere3_000{k}.vtp" for k in range(3)]
