import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import support
    return support.load_oracle()


@pytest.fixture(scope="session")
def ref():
    import support
    lib = support.load_ref()
    if lib is None:
        pytest.skip("oracle/_ref/libhalmd_ref.so not built (reference sources absent at build time)")
    return lib


@pytest.fixture(scope="session")
def hx():
    """The product's C ABI (ctypes); building is __graft_entry__.build()'s job, not a silent fallback."""
    from halmd_b200 import _lib
    return _lib


@pytest.fixture(scope="session")
def cuda(hx, oracle):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test started without a CUDA device")
    torch.cuda.set_device(0)
    import support
    support.install_device_rcp(hx, oracle)
    return torch
