"""Test configuration: `gpu` marker, import paths.

`-m "not gpu"` tests run without a GPU (oracle vs golden vectors, host logic, ABI).
`-m gpu` tests are the parity tests proper: CUDA path vs oracle, through the C ABI."""
import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# puts montecarlo-risk-engine_b200/ on sys.path so that `from controller.controller import ...`
# resolves exactly like the reference's tests/pytests/context.py does for src/
importlib.import_module("montecarlo-risk-engine_b200")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def hazards():
    """CDS-bootstrapped hazard curve used by the reference's CVA tests (tests/pytests/test_cva.py:17-32)."""
    return {0.5: 0.006402303360855854, 1.0: 0.01553038972325307, 2.0: 0.009729741230773657,
            3.0: 0.015552544648116201, 4.0: 0.021196186202801115, 5.0: 0.02284319986706472,
            7.0: 0.010111423894480876, 10.0: 0.00613267811172937, 15.0: 0.0036969930706003337,
            20.0: 0.003791311459217732}
