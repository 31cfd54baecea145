import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def g1_index():
    from ibwa_b200 import bwt_restore_bwt
    return (bwt_restore_bwt(os.path.join(GOLDEN, "g1.bwt")), bwt_restore_bwt(os.path.join(GOLDEN, "g1.rbwt")))
