import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


HAVE_GPU = _have_gpu()


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/lpm_oracle.c), built on demand.  Test infrastructure only."""
    from oracle import oracle as o
    o.lib()
    return o


@pytest.fixture(scope="session")
def engine():
    """The product engine on cuda:0.  GPU tests must go through the C ABI; no fallback exists."""
    if not HAVE_GPU:
        pytest.skip("no CUDA device")
    from lpm_b200.api import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="session")
def meshes():
    from lpm_b200.api import PolyMesh2d
    cache = {}

    def get(seed, depth):
        key = (seed, depth)
        if key not in cache:
            cache[key] = PolyMesh2d(seed, depth)
        return cache[key]
    return get


def field_rel_err(a, b, sel=None):
    """Field-relative max-norm used for every FP64 tolerance in this suite:
    max_i |a_i - b_i|_2 / max_i |b_i|_2 over the selected rows."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if sel is not None:
        a, b = a[sel], b[sel]
    if a.ndim == 1:
        a, b = a[:, None], b[:, None]
    num = np.sqrt(((a - b) ** 2).sum(axis=1)).max()
    den = np.sqrt((b ** 2).sum(axis=1)).max()
    if den == 0.0:
        return 0.0 if num == 0.0 else np.inf
    return num / den


def check_err(name, err, tol):
    """assert err <= tol, and append the observed error to $LPMX_PARITY_LOG (one JSON object per line) when that variable is
    set: the GPU visits collect the table of measured parity errors this way (profiles/r2*_parity_errors.jsonl)."""
    path = os.environ.get("LPMX_PARITY_LOG")
    if path:
        import json
        test = os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0]
        with open(path, "a") as f:
            f.write(json.dumps({"test": test, "quantity": name, "err": float(err), "tol": float(tol)}) + "\n")
    assert err <= tol, (name, err, tol)
