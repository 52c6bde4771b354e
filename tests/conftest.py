import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the read-only reference checkout at /root/reference")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir(REFERENCE)
    for it in items:
        if "reference" in it.keywords and not have_ref:
            it.add_marker(pytest.mark.skip(reason="/root/reference is not present on this machine"))


@pytest.fixture(scope="session")
def aid():
    import aid_b200
    return aid_b200


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
