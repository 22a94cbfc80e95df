import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """the CPU checker (oracle/liboracle.so); TEST INFRASTRUCTURE, never used by the product"""
    from oracle.oracle import Oracle, build
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        build()
    return Oracle()


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def lib():
    """the product's C-ABI library; built in-tree if missing (nvcc cross-compiles without a GPU)"""
    from atrip_b200 import capi
    if not os.path.exists(capi.lib_path()):
        capi.build_library()
    return capi.load_library()


def fh(s):
    return float.fromhex(s)
