import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "reference: needs the live reference tree at /root/reference")


def pytest_collection_modifyitems(config, items):
    from oracle import ref_shim
    have_ref = ref_shim.reference_root() is not None
    skip_ref = pytest.mark.skip(reason="/root/reference not present on this box")
    for it in items:
        if "reference" in it.keywords and not have_ref:
            it.add_marker(skip_ref)


GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
