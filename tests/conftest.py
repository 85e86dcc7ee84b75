import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_be():
    import oracle
    return oracle.load()


@pytest.fixture(scope="session")
def hostemu_be():
    from tests import _hostemu
    return _hostemu.load()


@pytest.fixture
def hostemu_coop_be(hostemu_be):
    """the g++ build of the WARP-COOPERATIVE kernel arithmetic (32 virtual lanes run phase by phase)"""
    hostemu_be.dll.hostemu_set_coop(1)
    yield hostemu_be
    hostemu_be.dll.hostemu_set_coop(0)


@pytest.fixture
def hostemu_dynamic_be(hostemu_be):
    """the opt-in persistent variant (lane-level refill from an instance queue): heavy workspace-slot reuse"""
    hostemu_be.dll.hostemu_set_dynamic(1)
    yield hostemu_be
    hostemu_be.dll.hostemu_set_dynamic(0)


@pytest.fixture(scope="session")
def gpu_be():
    import ratilqr_b200 as R
    be = R.new_backend(0)  # raises loudly if the .so is missing or no device can be opened
    yield be
    be.close()


@pytest.fixture(params=["oracle", "hostemu", "hostemu_coop", pytest.param("gpu", marks=pytest.mark.gpu)])
def backend(request):
    """Every provider of the C ABI: the CPU oracle, the g++ build of the kernel arithmetic, the CUDA library."""
    return request.getfixturevalue(request.param + "_be")


@pytest.fixture(params=["oracle", pytest.param("gpu", marks=pytest.mark.gpu)])
def full_backend(request):
    """Providers that export the complete ABI (PETS refit / solve included)."""
    return request.getfixturevalue(request.param + "_be")
