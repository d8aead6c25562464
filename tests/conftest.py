import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "needs_ref: needs oracle/_ref built from /root/reference")


@pytest.fixture(scope="session", autouse=True)
def _build_test_infra():
    """Build the oracle restatement / fake libjpeg front end if missing (gcc only)."""
    import oracle_lib
    oracle_lib.ensure_built()
    yield
