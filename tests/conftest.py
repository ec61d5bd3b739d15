import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def hn():
    from hnumo_loader import hnumo_b200
    return hnumo_b200


@pytest.fixture(scope="session")
def hn_lib(hn):
    hn.build_library()
    return hn.load_library()
