import os, sys
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    from tests import util
    return np.load(util.GOLDEN)


@pytest.fixture(scope="session")
def cams():
    from tests import util
    return util.dtu_cameras()
