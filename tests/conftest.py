import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the reference checkout under /root/reference (container only)")


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "tests", "files"))


needs_reference = pytest.mark.skipif(not have_reference(), reason="/root/reference is not mounted here")
