import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def manifest(golden_dir):
    import json
    with open(os.path.join(golden_dir, "manifest.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def built_lib():
    """Build (if stale) and return the path of libapnetg.so; nvcc cross-compiles without a GPU."""
    from animateportrait_b200.build import build
    return build()
