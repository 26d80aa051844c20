import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import __graft_entry__ as graft  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def pkg():
    return graft.load_package()


@pytest.fixture(scope="session")
def oracle_lib():
    """Path of the CPU oracle (built on demand; test infrastructure only)."""
    if not os.path.exists(graft.ORACLE_LIB):
        graft.build_oracle()
    return graft.ORACLE_LIB


@pytest.fixture(scope="session")
def engine_lib():
    """Path of the CUDA engine.  Never falls back to anything else."""
    if not os.path.exists(graft.LIB):
        graft.build_engine()
    return graft.LIB


@pytest.fixture(scope="session")
def dev_lib():
    """Development build of the engine: the product sources + measurement / self-test kernels."""
    if not os.path.exists(graft.DEV_LIB):
        graft.build_engine(dev=True)
    return graft.DEV_LIB


# Every behavioural test of the reference is run against BOTH libraries through the same
# marshalling code: the oracle on CPU (pins the oracle), the engine on the GPU (parity).
@pytest.fixture(params=["oracle", pytest.param("engine", marks=pytest.mark.gpu)])
def backend(request):
    return request.getfixturevalue("oracle_lib" if request.param == "oracle" else "engine_lib")
