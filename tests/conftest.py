import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_NAME = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session", autouse=True)
def _library_built():
    """The tests bind the in-tree libisac_b200.so: build it (nvcc cross-compiles without a GPU) when a fresh checkout has none."""
    build = importlib.import_module(PKG_NAME + ".build")
    if not os.path.exists(build.LIB_PATH):
        build.build_library(verbose=False)
    assert os.path.exists(build.LIB_PATH)


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module(PKG_NAME)


@pytest.fixture(scope="session")
def workloads():
    return importlib.import_module(PKG_NAME + ".workloads")


@pytest.fixture(scope="session")
def gpu():
    """Session fixture for -m gpu tests: the CUDA library must load and a device must exist."""
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test selected but no CUDA device is visible")
    lib = importlib.import_module(PKG_NAME + "._lib")
    return lib.get_context(0)
