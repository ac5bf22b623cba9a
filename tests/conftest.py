import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ge():
    import __graft_entry__ as g

    return g


@pytest.fixture(scope="session")
def pkg(ge):
    return ge.load_package()


@pytest.fixture(scope="session")
def oracle(ge):
    o = ge.load_oracle()
    o.build()
    o.set_threads(os.cpu_count() or 1)
    return o


@pytest.fixture(scope="session")
def synth(pkg):
    from importlib import import_module

    return import_module("sift_gpu_b200.synth")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    d = os.path.join(ROOT, "tests", "golden")

    def load(name):
        return np.load(os.path.join(d, name + ".npz"))

    return load


@pytest.fixture(scope="session")
def sift(pkg):
    """One shared device workspace for the GPU tests (2448^2 covers every fixture)."""
    s = pkg.Sift(2448, 2448, max_batch=2, max_kp_per_frame=1 << 15, device=0)
    yield s
    s.close()
