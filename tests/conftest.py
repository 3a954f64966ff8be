import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import orc
    return orc.Oracle()


@pytest.fixture(scope="session")
def reference():
    """The reference's own code (T1).  Prebuilt .so travels to the GPU box."""
    from oracle import orc
    orc.build()
    if not orc.have_ref():
        pytest.skip("oracle/_ref/libmaveric_ref.so not available (needs /root/reference to build)")
    return orc.Reference()


@pytest.fixture(scope="session")
def image0():
    return dict(np.load(os.path.join(GOLDEN, "ref_image0.npz")))


@pytest.fixture(scope="session")
def kat():
    return dict(np.load(os.path.join(GOLDEN, "ref_kat.npz")))


@pytest.fixture(scope="session")
def synth():
    import maveric_slam_b200  # noqa: F401
    from maveric_slam_b200 import synth as s
    return s


@pytest.fixture(scope="session")
def tracker():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import maveric_slam_b200  # noqa: F401
    from maveric_slam_b200 import tracking
    return tracking.Tracker(0)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.int32)
