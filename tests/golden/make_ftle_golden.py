"""Golden FTLE values of the REFERENCE's ComputeFTLE<CubedSphereSeed> / ComputeFTLE<QuadRectSeed> and get_max_ftle
(src/mesh/lpm_ftle.hpp, compiled in place: oracle/_ref/liblpm_ref.so) on the seeded cases of tests/ftle_cases.py
-> tests/golden/ref_ftle.npz.      python tests/golden/make_ftle_golden.py   (2>/dev/null silences the reference's
own debug warnings about its planar branch)"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ftle_cases  # noqa: E402
from lpm_b200.api import PolyMesh2d  # noqa: E402
from oracle import oracle as O  # noqa: E402

if __name__ == "__main__":
    R = ctypes.CDLL(O.REF_LIB)
    out = {}
    for name, case in (("cubed3", ftle_cases.sphere_case(PolyMesh2d("cubed", 3))), ("plane12", ftle_cases.plane_case())):
        f, fp, mx = O.ftle(**case, L=R)
        out[name + "_ftle"], out[name + "_face_phys"], out[name + "_max"] = f, fp, np.array(mx)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_ftle.npz"), **out)
    print({k: v.shape for k, v in out.items()})
