import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import bindings
    lib = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(lib):
        bindings.build(ref=False)
    return bindings.Oracle()


@pytest.fixture(scope="session")
def ref():
    """The compiled reference (oracle/_ref).  Built here when /root/reference is mounted, else used prebuilt."""
    from oracle import bindings
    if not bindings.Ref.available() and os.path.isdir("/root/reference/include/hashdag"):
        bindings.build(ref=True)
    if not bindings.Ref.available():
        pytest.skip("oracle/_ref not built (reference tree not mounted)")
    return bindings.Ref()


@pytest.fixture(scope="session")
def hd():
    """The product library; GPU tests fail loudly (not skip) if it cannot be loaded."""
    import vkhashdag_b200 as v
    v.lib()
    return v
