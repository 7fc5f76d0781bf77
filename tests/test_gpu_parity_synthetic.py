"""GPU parity on BASELINE configs[4] (synthetic direct-sum Biot-Savart sweep: N i.i.d. collocated particles, tools/synthetic_sweep.py).

On these inputs the kernel's formula is ill-conditioned in the reference as well: the nearest pairs of N random points have
d = 1 - x.y ~ 1/N, and rounding x.y to double perturbs d by 2^-53 whatever the evaluation order, i.e. that pair's term by
2^-53 / d relative.  So the reference's own FP64 arithmetic (BVEFaceVelocity compiled in place, oracle/_ref) sits 1e-10..1e-8
away from the extended-precision value of the same formula on the same doubles, and two correct double-precision evaluations
(the reference's per-pair cross(x, y) / d, this engine's x cross sum(G y / d) with FMA chains) differ from each other by the same
order.  north_star's 1e-12 cannot be met by ANY pair of FP64 implementations here -- the reference's OpenMP and CUDA builds
included -- and holds on the quasi-uniform meshes (every other test).  What is asserted on 4 096 sampled targets:
  * the engine is no farther from the long-double value than the reference arithmetic is (x 4: both distances are set by the
    few closest pairs of the sample and scatter by that much between N = 1e5 and 1e6), and
  * the engine and the reference arithmetic agree to within the sum of their distances from it;
all three numbers are logged (LPMX_PARITY_LOG) and quoted in DESIGN.md section 5."""
import ctypes
import os
import sys

import numpy as np
import pytest

from conftest import check_err
from lpm_b200.api import BVESolver

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

pytestmark = pytest.mark.gpu


def _rel(a, b, scale):
    return float(np.linalg.norm(a - b, axis=1).max() / scale)


@pytest.mark.parametrize("n", [100_000, 1_000_000])
def test_synthetic_collocated_velocity_sampled_against_reference_arithmetic(engine, oracle, n):
    from synthetic_sweep import particles
    x, zeta, area = particles(n)
    mask = np.zeros(n, dtype=np.uint8)
    s = BVESolver(engine, 0, n)
    s.set_state(None, None, None, x, zeta, None, area, mask)
    s.init_velocity()
    vel = np.zeros((n, 3))
    s.get_state(None, None, None, None, None, vel)
    s.close()
    idx = np.sort(np.random.default_rng(20261018).choice(n, 4096, replace=False)).astype(np.int32)
    L = ctypes.CDLL(oracle.REF_LIB) if os.path.exists(oracle.REF_LIB) else None  # the reference functor itself, else the port
    ref = oracle.bve_velocity_subset(idx, x, zeta, area, mask, L=L)
    ld = oracle.bve_velocity_subset(idx, x, zeta, area, mask, long_double=True)
    scale = float(np.linalg.norm(ld, axis=1).max())
    e_gpu_ld, e_ref_ld, e_gpu_ref = _rel(vel[idx], ld, scale), _rel(ref, ld, scale), _rel(vel[idx], ref, scale)
    check_err(f"N={n} reference FP64 vs long double (conditioning of the inputs)", e_ref_ld, 1e-6)  # logged; a property of the inputs
    check_err(f"N={n} engine vs long double", e_gpu_ld, max(1e-12, 4 * e_ref_ld))
    check_err(f"N={n} engine vs reference FP64", e_gpu_ref, max(1e-12, e_gpu_ld + e_ref_ld))
    assert np.isfinite(vel).all()
