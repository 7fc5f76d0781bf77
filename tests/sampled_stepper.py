"""Test helper: one BVERK4::advance_timestep (/root/reference/src/lpm_bve_rk4_impl.hpp:63-167) restated in numpy around the
oracle's velocity sums, with the VERTEX targets restricted to a sample.  Vertices are never sources, so a vertex's RK stages
depend only on its own state and on the faces'; all faces are stepped because each stage's face state is the next stage's
source set.  Lets the full-size configs (cubed-7: 229 376 targets) meet the oracle at the cost of 4 face evaluations + 5 sampled
vertex evaluations.  tests/test_sampled_stepper.py checks it against oracle_bve_rk4_step (the C restatement of the whole step)."""
import numpy as np


def bve_rk4_step_sampled(oracle, m, vert_zeta, face_zeta, dt, Omega, idx, L=None):
    fa, fm = m.face_area, m.face_mask
    vx, vz = np.ascontiguousarray(m.vert_xyz[idx]), np.ascontiguousarray(vert_zeta[idx])
    fx, fz = m.face_xyz.copy(), np.ascontiguousarray(face_zeta).copy()

    def vel(tx, sx, sz):
        return (oracle.bve_velocity(tx, sx, sz, fa, fm, L=L), oracle.bve_velocity(None, sx, sz, fa, fm, collocated=True, L=L))

    def tend(u):  # BVEVorticityTendency: dzeta = -2 Omega w dt
        return -2.0 * Omega * u[:, 2] * dt

    vu, fu = vel(vx, fx, fz)  # BVESphere::init_velocity
    kx_v, kz_v, kx_f, kz_f = [], [], [], []
    wvx, wvz, wfx, wfz = vx, vz, fx, fz
    for stage, c in enumerate((0.5, 0.5, 1.0, None)):
        if stage > 0:
            vu, fu = vel(wvx, wfx, wfz)
        kx_v.append(dt * vu), kz_v.append(tend(vu)), kx_f.append(dt * fu), kz_f.append(tend(fu))
        if c is not None:  # KokkosBlas::update(1, x, c, k, 0, work)
            wvx, wvz = 1.0 * vx + c * kx_v[-1], 1.0 * vz + c * kz_v[-1]
            wfx, wfz = 1.0 * fx + c * kx_f[-1], 1.0 * fz + c * kz_f[-1]
    sixth, third = 1.0 / 6.0, 1.0 / 3.0
    vx = vx + (sixth * (kx_v[0] + kx_v[3]) + third * (kx_v[1] + kx_v[2]))
    vz = vz + (sixth * (kz_v[0] + kz_v[3]) + third * (kz_v[1] + kz_v[2]))
    fx = fx + (sixth * (kx_f[0] + kx_f[3]) + third * (kx_f[1] + kx_f[2]))
    fz = fz + (sixth * (kz_f[0] + kz_f[3]) + third * (kz_f[1] + kz_f[3]))  # quirk A-i: facevort4 in the facevort3 slot (:155-157)
    vu, fu = vel(vx, fx, fz)
    return {"vert_xyz": vx, "vert_zeta": vz, "vert_vel": vu, "face_xyz": fx, "face_zeta": fz, "face_vel": fu}
