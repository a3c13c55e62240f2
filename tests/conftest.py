import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (authoring container only)")


def has_reference():
    return os.path.isdir(os.path.join(REFERENCE, "src", "modules"))


@pytest.fixture(scope="session")
def synth_w():
    from canonswap_b200 import synth
    return synth.synth_weights()
