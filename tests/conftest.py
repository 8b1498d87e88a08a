"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"`  : oracle vs golden vectors, host logic, C-ABI symbol checks (no GPU needed).
`-m gpu`        : parity tests proper -- the CUDA path through the C-ABI vs the oracle.
"""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference behind ctypes; only present where oracle/_ref was built."""
    from oracle.pyoracle import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    return Reference()
