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
