import os
import sys

import pytest

# Guard mode for the whole test session: every device array the library allocates carries canary bytes on both sides
# (checked when it is freed; tests/test_zz_guard.py reads the verdict last).  Must be set before the library is used.
os.environ.setdefault("SB_GUARD", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import sb_oracle
    sb_oracle.build()
    return sb_oracle
