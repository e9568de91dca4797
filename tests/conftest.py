import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (authoring container only)")


def pytest_collection_modifyitems(config, items):
    from oracle.ref_import import reference_available
    if not reference_available():
        skip = pytest.mark.skip(reason="/root/reference not present")
        for it in items:
            if "reference" in it.keywords:
                it.add_marker(skip)


GOLDEN = os.path.join(ROOT, "tests", "golden")
