import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "user-eph_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

REFERENCE = "/root/reference"   # exists in the development container only (never on the GPU box)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """CPU-side libraries (oracle, host side, harness) are built on demand; the CUDA library is cross-compiled too."""
    from eph_b200 import _paths
    need = [p for p in (_paths.lib_path("engine"), _paths.lib_path("fix")) if not os.path.exists(p)]
    if need:
        _paths.build_all()
    import eph_harness
    if not os.path.exists(eph_harness.LIB):
        eph_harness.build()
    from oracle import oracle
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        oracle.build()


@pytest.fixture(scope="session")
def has_reference_tree():
    return os.path.isdir(REFERENCE)


@pytest.fixture(scope="session")
def ref():
    from oracle import reference
    if not reference.available():
        pytest.skip("oracle/_ref/libeph_ref.so not built (needs /root/reference at build time)")
    return reference


@pytest.fixture(scope="session")
def synth_beta_1(tmp_path_factory):
    from eph_harness import harness as H
    p = tmp_path_factory.mktemp("beta") / "synth1.beta"
    return str(H.write_beta_file(p, H.synthetic_knots(1, n_beta=5001, drho=0.01)))


@pytest.fixture(scope="session")
def synth_beta_4(tmp_path_factory):
    from eph_harness import harness as H
    p = tmp_path_factory.mktemp("beta") / "synth4.beta"
    return str(H.write_beta_file(p, H.synthetic_knots(4, n_beta=5001, drho=0.01)))


@pytest.fixture(scope="session")
def ni_trunc_beta():
    return os.path.join(GOLDEN, "Ni_trunc.beta")


@pytest.fixture(scope="session")
def sys500():
    from eph_harness import harness as H
    return H.make_system(5)


def gpu_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
