"""pytest configuration: `gpu` marker, repo root on sys.path, lazily-built artefacts.

-m "not gpu": oracle vs golden vectors, host logic, C-ABI symbol export (runs in the CPU container)
-m gpu      : parity of the CUDA path (through the C ABI) against the oracle (runs on a B200)
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def port():
    """the plain-C restatement oracle (built on demand; gcc only)"""
    from oracle import cpu_parth
    cpu_parth.build("port")
    return cpu_parth


@pytest.fixture(scope="session")
def ref():
    """the unmodified reference compiled into oracle/_ref (None when not available)"""
    from oracle import cpu_parth
    if not cpu_parth.available("reference"):
        if os.path.isdir("/root/reference/src/Parth"):
            cpu_parth.build("reference")
        else:
            return None
    return cpu_parth


@pytest.fixture(scope="session")
def gpu_api():
    from parth_b200 import api
    api.load_library()
    return api
