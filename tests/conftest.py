import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# no TabPFNv2 checkpoint exists offline: the suite runs on the seeded random init, which the package only hands out
# on explicit request (npe_pfn_b200/weights.py::PFNWeights.default)
os.environ.setdefault("NPE_PFN_B200_ALLOW_RANDOM_INIT", "1")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")


@pytest.fixture(scope="session")
def weights():
    from npe_pfn_b200.weights import PFNWeights
    return PFNWeights.random_init()


@pytest.fixture(scope="session")
def engine(weights):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from npe_pfn_b200.engine import Engine
    return Engine(weights=weights, max_slots=16)
