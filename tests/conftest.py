"""pytest configuration: `-m gpu` = parity tests proper (need a B200), `-m "not gpu"` = oracle / host logic / ABI."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py
    oracle_py.build(ref=True)
    return oracle_py.Oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference TUs compiled against oracle/shim (only where /root/reference or a prebuilt .so exists)."""
    from oracle import oracle_py
    oracle_py.build(ref=True)
    if not os.path.exists(oracle_py.REF_SO):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    return oracle_py.Reference()


@pytest.fixture(scope="session")
def lib():
    import bayesembler_b200 as bb
    from bayesembler_b200 import _build
    _build.build()
    bb.load_library()
    return bb


@pytest.fixture(scope="session")
def facade(lib):
    """tests/cpp/facade_driver.cpp: C-callable test driver around the C++ host facade, built with g++ against the library."""
    import ctypes as C
    import subprocess
    from oracle import oracle_py
    SRC = os.path.join(ROOT, "tests", "cpp", "facade_driver.cpp")
    HPP = os.path.join(ROOT, "bayesembler_b200", "host", "bayesembler_b200.hpp")
    OUT = os.path.join(ROOT, "tests", "cpp", "libfacade_test.so")
    libdir = os.path.join(ROOT, "bayesembler_b200")
    newest = max(os.path.getmtime(p) for p in (SRC, HPP, os.path.join(ROOT, "include", "bayesembler_b200.h")))
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < newest:
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=gnu++17", "-fPIC", "-shared", "-Wall", "-Werror", "-o", OUT, SRC,
                               "-L" + libdir, "-lbayesembler_b200", "-Wl,-rpath," + libdir])
    lib.load_library()
    L = C.CDLL(OUT)
    L.facade_gtf.restype = C.c_int
    L.facade_gtf.argtypes = oracle_py.GTF_ARGTYPES
    L.facade_gtf_from_accumulators.restype = C.c_int
    return L
