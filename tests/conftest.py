import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The reference's own code compiled in place (oracle/_ref); skipped if it was not built."""
    from oracle.pyoracle import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref not built (no /root/reference on this box and no prebuilt copy)")
    return Reference()
