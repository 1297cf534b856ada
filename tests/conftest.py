import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_devices():
    try:
        from jointforces_b200 import _lib
        return _lib.device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip: only auto-skip when the
    # user did not ask for gpu tests explicitly.
    if "gpu" in (config.getoption("-m") or ""):
        return
    if _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# ---- shared helpers ------------------------------------------------------------------------------
COLLAGEN12 = (1645.0, 0.0008, 0.0075, 0.033)
LINEAR = (2364.0, 1e30, 1e30, 1e30)
FIBRIN = (2091.0, 0.002, 1e30, 1e30)
BATCH_A = (1449.0, 0.00215, 0.032, 0.055)


def make_case(lf_iso=0.667, pressure=100.0, method="delaunay"):
    """Synthetic spherical inclusion + the reference's boundary conditions for one pressure."""
    from jointforces_b200.mesh import generate_spherical_inclusion
    from jointforces_b200.simulation import spherical_boundary_conditions
    coords, tets = generate_spherical_inclusion(lf_iso=lf_iso, method=method)
    r_in, r_out = 100e-6, 10000e-6
    bc = spherical_boundary_conditions(coords, r_in, r_out, pressure)
    movable = np.any(np.isnan(bc["displacement"]), axis=1)
    return dict(coords=coords, tets=tets.astype(np.int32), bc=bc, movable=movable,
                f_target=np.nan_to_num(bc["forces"]), r_in=r_in, r_out=r_out, pressure=pressure)


@pytest.fixture(scope="session")
def beams():
    from jointforces_b200.solver import build_beams
    return build_beams(300)
