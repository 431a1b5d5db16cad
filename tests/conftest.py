import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def _ensure_native_library():
    """The tests call the product through libqcdsim.so.  On a fresh checkout (the library is git-ignored)
    build it first -- nvcc cross-compiles sm_100a without a GPU.  build.py is loaded by path because
    importing the package itself requires the library."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_qcd_build", os.path.join(ROOT, "qcdenoise_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if not os.path.exists(mod.LIB):       # missing only: file times do not survive every copy of the tree
        mod.build()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    _ensure_native_library()


@pytest.fixture(scope="session")
def goldens():
    with open(os.path.join(GOLDEN_DIR, "witness_goldens.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ref_graphs():
    with open(os.path.join(GOLDEN_DIR, "graphs.json")) as f:
        return json.load(f)
