import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc_lib():
    import orc

    return orc.oracle()


@pytest.fixture(scope="session")
def ref_lib():
    import orc

    if not orc.ref_available():
        pytest.skip("oracle/_ref/libcspm_ref.so not built (needs /root/reference)")
    return orc.ref()
