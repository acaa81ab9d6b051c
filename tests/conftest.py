import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """Build (or reuse) libmindrec_b200.so; nvcc cross-compiles without a GPU."""
    from mindrec_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def cuda(built_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _seed_torch_per_test(request):
    """Every test starts from a torch RNG state derived from its own id (CPU and CUDA generators): a draw does not
    depend on which tests ran before it, so a test that passes once passes always, whatever is added around it."""
    import zlib
    import torch
    torch.manual_seed(zlib.crc32(request.node.nodeid.encode()) & 0x7fffffff)
    yield
