import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    import numpy as np
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    geo = dict(vertices=z["vertices"], triangles=z["triangles"], mu0=z["mu0"], mu1=z["mu1"],
               edges=z["triangles"][:, [0, 1, 1, 2, 2, 0]].reshape(-1, 2))
    kw = {str(k): float(v) for k, v in zip(z["kw_keys"], z["kw_vals"])}
    if "nit" in kw:
        kw["nit"] = int(kw["nit"])
    for flag in ("check_kkt_step_by_step", "is_palm", "is_constant_scaling"):
        if flag in kw:
            kw[flag] = bool(kw[flag])
    return z, geo, int(z["n_time"]), kw


@pytest.fixture
def golden():
    return load_golden
