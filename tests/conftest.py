import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a); run with `-m gpu` on the GPU box")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree shared library (built here if missing; nvcc cross-compiles without a GPU)."""
    from workoutdetector_b200 import _lib
    _lib.build()
    return _lib.load()


@pytest.fixture(scope="session")
def count_oracle_c():
    """oracle/_build/libcount_oracle.so (plain-C restatement of pred_to_count) via ctypes."""
    import ctypes as C
    import subprocess
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libcount_oracle.so"))
    lib.oracle_count_reps.restype = C.c_int
    lib.oracle_count_reps.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_void_p]
    return lib
