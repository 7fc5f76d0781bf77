import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from lpm_b200.api import PolyMesh2d, Engine
from lpm_b200 import gallery
e = Engine(0)
m = PolyMesh2d('cubed', 4)
f = gallery.GaussianVortexSphere()
vz, fz = f(m.vert_xyz), f(m.face_xyz)
dt = 0.5 / 15; Om = 2 * np.pi
for eps in (0.0, 0.05):
    pu, pp = O.ic2d_sums(m.vert_xyz, m.face_xyz, fz, m.face_area, m.face_mask, eps=eps)
    au, ap = O.ic2d_sums(None, m.face_xyz, fz, m.face_area, m.face_mask, eps=eps, targets_are_sources=True)
    st = [m.vert_xyz.copy(), vz.copy(), pu, pp, m.face_xyz.copy(), fz.copy(), au, ap]
    for n in (1, 2, 3):
        ref = [a.copy() for a in st]; got = [a.copy() for a in st]
        O.ic2d_rk2_step(dt, Om, eps, *ref, m.face_area, m.face_mask, n_steps=n)
        e.ic2d_rk2_step(dt, Om, eps, *got, m.face_area, m.face_mask, n_steps=n)
        print('eps', eps, 'steps', n, ' '.join('%s %.2e' % (nm, np.abs(a - b).max()) for nm, a, b in zip('px pz pu pp ax az au ap'.split(), got, ref)))
