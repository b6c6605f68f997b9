import os
import sys

import pytest

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before anything starts CUDA (see csrc/chain_core.cu)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def pkg():
    return entry.load_package()


@pytest.fixture(scope="session")
def synth(pkg):
    from mm2gb_b200 import synth as s
    return s


@pytest.fixture(scope="session")
def po():
    return entry.load_oracle()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def ctx(pkg):
    """A small shared chaining context on cuda:0 (map-ont parameters)."""
    c = pkg.ChainContext(pkg.map_ont_misc(), device=0, max_anchors=1 << 21, max_reads=1 << 14, n_slots=2)
    yield c
    c.close()
