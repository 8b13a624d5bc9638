import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(HERE, "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def assets_dir():
    return os.path.join(HERE, "golden", "assets")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def dz():
    """The product package with its CUDA library loaded; GPU tests must fail (not skip) if it is missing."""
    sys.path.insert(0, ROOT)
    import dzahui_b200
    from dzahui_b200 import _lib, build
    if not os.environ.get("DZB_LIB"):
        build.build()   # no-op unless the library is missing or its sources changed since it was built (content hash)
    _lib.load()
    return dzahui_b200
