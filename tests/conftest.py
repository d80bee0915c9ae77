import os
import sys

import pytest

# the multi-tile tests run several handles (7 streams each) on one device, some of them with kernels that wait for a
# neighbouring tile: every stream needs its own hardware queue (the library asks for the same when it is loaded first)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")  # a lazy kernel load waits for an idle device (see bcd_b200_tile_setup)

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="session")
def crop_a():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "testraw_crop_a.npz")))


@pytest.fixture(scope="session")
def crop_b():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "testraw_crop_b.npz")))


@pytest.fixture(scope="session")
def crop_c():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "testraw_crop_c.npz")))


def inputs_of(d):
    return {k: d[k] for k in ("color", "nsamples", "hist", "cov")}
