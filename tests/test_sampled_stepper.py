"""CPU: the sampled-vertex restatement of one BVERK4 step (tests/sampled_stepper.py, used by the full-size GPU parity test)
against the oracle's whole-step C restatement, which tests/test_oracle_golden.py pins against the compiled reference."""
import numpy as np

from conftest import field_rel_err
from lpm_b200 import gallery
from sampled_stepper import bve_rk4_step_sampled


def test_sampled_vertex_step_equals_the_whole_step(oracle, meshes):
    m = meshes("cubed", 4)
    f = gallery.RossbyHaurwitz54()
    f.set_stationary_wave_speed()
    vz, fz = f(m.vert_xyz), f(m.face_xyz)
    dt, Omega = 0.02, 2 * np.pi
    vu = oracle.bve_velocity(m.vert_xyz, m.face_xyz, fz, m.face_area, m.face_mask)
    fu = oracle.bve_velocity(None, m.face_xyz, fz, m.face_area, m.face_mask, collocated=True)
    st = [m.vert_xyz.copy(), vz.copy(), vu, m.face_xyz.copy(), fz.copy(), fu]
    oracle.bve_rk4_step(dt, Omega, *st, m.face_area, m.face_mask, n_steps=1)
    idx = np.sort(np.random.default_rng(3).choice(m.n_verts, 200, replace=False))
    r = bve_rk4_step_sampled(oracle, m, vz, fz, dt, Omega, idx)
    # same arithmetic up to the association of the update expression: a few ulp
    assert field_rel_err(r["vert_xyz"], st[0][idx]) <= 1e-15
    assert field_rel_err(r["vert_zeta"], st[1][idx]) <= 1e-15
    assert field_rel_err(r["vert_vel"], st[2][idx]) <= 1e-14
    assert field_rel_err(r["face_xyz"], st[3]) <= 1e-15
    assert field_rel_err(r["face_zeta"], st[4]) <= 1e-15
    assert field_rel_err(r["face_vel"], st[5]) <= 1e-14
    # the quirk is visible: with Omega != 0 the textbook update differs
    assert np.abs(r["face_zeta"] - fz).max() > 0
